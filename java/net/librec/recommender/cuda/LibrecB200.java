// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK / jni.h).  Reference-side binding a LibRec
// maintainer adds: a 1:1 JNI forwarder to include/librec_b200.h.  See INTEGRATION.md.
package net.librec.recommender.cuda;

import java.nio.ByteBuffer;

/** Thin JNI surface over liblibrec_b200.so (C ABI, include/librec_b200.h). */
final class LibrecB200 {
    static { System.loadLibrary("librec_b200_jni"); }   // ~150-line C file, INTEGRATION.md section 3

    static final int MODEL_BIASEDMF = 0, MODEL_PMF = 1, MODEL_BPR = 2, MODEL_RANKSGD = 3;
    static final int UPDATE_ATOMIC = 0, UPDATE_HOGWILD = 1, UPDATE_REFERENCE_ORDER = 2;

    // every native returns the lrk_status; 0 == LRK_OK
    static native long create(int device, int model, int numFactors, int updateMode, long seed, int topnPath);
    static native int destroy(long h);
    static native String lastError(long h);
    static native ByteBuffer hostAlloc(long bytes);          // lrk_host_alloc -> direct, pinned buffer
    static native int hostFree(ByteBuffer b);
    static native int setTrainCsr(long h, int numUsers, int numItems, ByteBuffer rowptr, ByteBuffer col, ByteBuffer val);
    static native int setFactors(long h, ByteBuffer P, ByteBuffer Q, ByteBuffer bu, ByteBuffer bi, double globalMean);
    static native int getFactors(long h, ByteBuffer P, ByteBuffer Q, ByteBuffer bu, ByteBuffer bi);
    static native int sgdEpoch(long h, float lr, float regU, float regI, double regB, int epochIdx, double[] lossOut);
    static native int predictPairs(long h, int[] users, int[] items, long n, double[] out);
    static native int evalRating(long h, int numUsers, long[] rowptr, int[] col, double[] val, double minRate,
                                 double maxRate, double[] predOut, double[] rmseMaeOut);
    static native int topn(long h, int[] usersOrNull, int nq, int topN, int excludeTrain, int[] outItems,
                           double[] outScores, int[] outCounts);
    static native int commUniqueId(byte[] out128);
    static native int commInit(long h, int rank, int world, byte[] id128);
}
