// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK).  Reference-side binding a LibRec maintainer adds: the native declarations that
// java/librec_b200_jni.c forwards 1:1 to include/librec_b200.h.  See INTEGRATION.md.
package net.librec.recommender.cuda;

import java.nio.ByteBuffer;

/** Thin JNI surface over liblibrec_b200.so (C ABI, include/librec_b200.h).  Every int result is the lrk_status; 0 == LRK_OK. */
final class LibrecB200 {
    static { System.loadLibrary("librec_b200_jni"); }

    static final int MODEL_BIASEDMF = 0, MODEL_PMF = 1, MODEL_BPR = 2, MODEL_RANKSGD = 3, MODEL_GBPR = 4, MODEL_SVDPP = 5, MODEL_AOBPR = 6,
            MODEL_WRMF = 7, MODEL_EALS = 8;
    static final int UPDATE_ATOMIC = 0, UPDATE_HOGWILD = 1, UPDATE_REFERENCE_ORDER = 2;
    static final int ERR_DIVERGED = -5;

    // lifecycle
    static native String version();
    static native int abiVersion();
    static native int deviceCount();
    static native long create(int device, int model, int numFactors, int updateMode, long seed, int topnPath);   // 0 on failure
    static native long createMulti(int[] devices, int model, int numFactors, int updateMode, long seed, int topnPath);   // rec.cuda.devices
    static native int destroy(long h);
    static native String lastError(long h);
    static native int setStream(long h, long cudaStream);
    static native int synchronize(long h);
    static native ByteBuffer hostAlloc(long bytes);          // lrk_host_alloc -> direct, pinned buffer (null on failure)
    static native int hostFree(ByteBuffer b);
    // staging (direct buffers in native byte order)
    static native int setTrainCsr(long h, int numUsers, int numItems, ByteBuffer rowptr, ByteBuffer col, ByteBuffer val);
    static native int setFactors(long h, ByteBuffer P, ByteBuffer Q, ByteBuffer bu, ByteBuffer bi, double globalMean);
    static native int getFactors(long h, ByteBuffer P, ByteBuffer Q, ByteBuffer bu, ByteBuffer bi);
    static native int stageStats(long h, long[] out4);
    static native int setParam(long h, byte[] nameUtf8, double value);      // "gbpr.rho", "gbpr.gsize"
    static native int setMatrix(long h, byte[] nameUtf8, ByteBuffer values);   // "svdpp.y"
    static native int getMatrix(long h, byte[] nameUtf8, ByteBuffer values);
    // training
    static native int sgdEpoch(long h, float lr, float regU, float regI, double regB, int epochIdx, double[] lossOut);
    static native int sgdEpochs(long h, int nEpochs, float lr, float decay, float maxLr, float regU, float regI, double regB,
                                int firstEpochIdx, double[] lossesOut);
    static native int lastEpochMs(long h, float[] out1);
    static native int sgdSafeguardState(long h, long[] concDivAndRollbacks);
    static native int launchCount(long h, long[] out1);
    static native int bprPeekSamples(long h, int epochIdx, long first, long n, int[] out3n);
    static native int debugStream(long h, ByteBuffer su, ByteBuffer si, ByteBuffer sr, ByteBuffer units, long maxUnits, long[] nUnitsOut1);
    // prediction
    static native int predictPairs(long h, int[] users, int[] items, long n, double[] out);
    static native int evalRating(long h, int numUsers, ByteBuffer rowptr, ByteBuffer col, ByteBuffer val, double minRate,
                                 double maxRate, ByteBuffer predOutOrNull, double[] rmseMaeOut);
    // top-N
    static native int topn(long h, int[] usersOrNull, int nq, int topN, int excludeTrain, ByteBuffer outItems,
                           ByteBuffer outScores, ByteBuffer outCounts);
    static native int evalRanking(long h, int topN, ByteBuffer tRowptr, ByteBuffer tCol, ByteBuffer tVal, ByteBuffer outItemsOrNull,
                                  ByteBuffer outScoresOrNull, ByteBuffer outCountsOrNull, double[] outMeasures8);
    static native int topnStats(long h, long[] fastAndFallbackUsers, float[] msOut1);
    static native int topnPhaseMs(long h, float[] out6);
    // measurement aid
    static native int probeL2(long h, long workingSetBytes, int rowFloats, double[] outGbps3);
    // multi-GPU (one JVM per GPU)
    static native int commUniqueId(byte[] out128);
    static native int commInit(long h, int rank, int world, byte[] id128);
}
