/* librec_b200_jni.c -- the JNI forwarder between net.librec.recommender.cuda.LibrecB200 (java/net/librec/recommender/cuda/
 * LibrecB200.java) and the C ABI of include/librec_b200.h.  One function per export, no logic of its own.
 *
 *   gcc -shared -fPIC -O2 -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude java/librec_b200_jni.c \
 *       -Llibrec_b200/_lib -llibrec_b200 -o liblibrec_b200_jni.so
 *
 * The build image has no JDK, so here the file is only compile-checked against tests/stubs/jni.h
 * (tests/test_abi_symbols.py::test_jni_forwarder_compiles, which also checks that every lrk_* export is forwarded).
 *
 * Conventions: large inputs / outputs travel in direct ByteBuffers (pinned through lrk_host_alloc, native byte order) -- no copy;
 * small ones in Java arrays pinned with GetPrimitiveArrayCritical for the duration of the call (the native side does not call
 * back into the JVM).  Every function returns the lrk_status; 0 == LRK_OK. */
#include <jni.h>
#include <stdint.h>
#include <string.h>
#include "librec_b200.h"

#define LRK_JNI(name) JNIEXPORT JNICALL Java_net_librec_recommender_cuda_LibrecB200_##name
#define H(x) ((lrk_handle_t)(intptr_t)(x))
#define BUF(b) ((b) ? (*env)->GetDirectBufferAddress(env, (b)) : NULL)
#define PIN(a) ((a) ? (*env)->GetPrimitiveArrayCritical(env, (a), NULL) : NULL)
#define UNPIN(a, p, mode) do { if (a) (*env)->ReleasePrimitiveArrayCritical(env, (a), (p), (mode)); } while (0)

/* ---- lifecycle ---------------------------------------------------------------------- */
jstring LRK_JNI(version)(JNIEnv* env, jclass c) { (void)c; return (*env)->NewStringUTF(env, lrk_version()); }
jint LRK_JNI(abiVersion)(JNIEnv* env, jclass c) { (void)env; (void)c; return lrk_abi_version(); }
jint LRK_JNI(deviceCount)(JNIEnv* env, jclass c) { (void)env; (void)c; return lrk_device_count(); }

jlong LRK_JNI(create)(JNIEnv* env, jclass c, jint device, jint model, jint numFactors, jint updateMode, jlong seed, jint topnPath) {
    lrk_config_t cfg;
    lrk_handle_t h = NULL;
    (void)env; (void)c;
    memset(&cfg, 0, sizeof cfg);
    cfg.device = device; cfg.model = model; cfg.num_factors = numFactors; cfg.update_mode = updateMode;
    cfg.seed = (uint64_t)seed; cfg.topn_path = topnPath;
    return lrk_create(&cfg, &h) == LRK_OK ? (jlong)(intptr_t)h : 0;     /* 0: lastError(0) has the text */
}
jlong LRK_JNI(createMulti)(JNIEnv* env, jclass c, jintArray devices, jint model, jint numFactors, jint updateMode, jlong seed, jint topnPath) {
    lrk_config_t cfg;
    lrk_handle_t h = NULL;
    int32_t* d = (int32_t*)PIN(devices);
    const jsize n = (*env)->GetArrayLength(env, devices);
    int rc;
    (void)c;
    memset(&cfg, 0, sizeof cfg);
    cfg.model = model; cfg.num_factors = numFactors; cfg.update_mode = updateMode; cfg.seed = (uint64_t)seed; cfg.topn_path = topnPath;
    rc = lrk_create_multi(&cfg, d, (int32_t)n, &h);
    UNPIN(devices, d, JNI_ABORT);
    return rc == LRK_OK ? (jlong)(intptr_t)h : 0;
}
jint LRK_JNI(destroy)(JNIEnv* env, jclass c, jlong h) { (void)env; (void)c; return lrk_destroy(H(h)); }
jstring LRK_JNI(lastError)(JNIEnv* env, jclass c, jlong h) { (void)c; return (*env)->NewStringUTF(env, lrk_last_error(H(h))); }
jint LRK_JNI(setStream)(JNIEnv* env, jclass c, jlong h, jlong cudaStream) { (void)env; (void)c; return lrk_set_stream(H(h), (void*)(intptr_t)cudaStream); }
jint LRK_JNI(synchronize)(JNIEnv* env, jclass c, jlong h) { (void)env; (void)c; return lrk_synchronize(H(h)); }

jobject LRK_JNI(hostAlloc)(JNIEnv* env, jclass c, jlong bytes) {
    void* p = NULL;
    (void)c;
    if (bytes < 0 || lrk_host_alloc(&p, (uint64_t)bytes) != LRK_OK) return NULL;
    return (*env)->NewDirectByteBuffer(env, p, bytes);
}
jint LRK_JNI(hostFree)(JNIEnv* env, jclass c, jobject buf) { (void)c; return lrk_host_free(BUF(buf)); }

/* ---- staging ------------------------------------------------------------------------ */
jint LRK_JNI(setTrainCsr)(JNIEnv* env, jclass c, jlong h, jint numUsers, jint numItems, jobject rowptr, jobject col, jobject val) {
    (void)c;
    return lrk_set_train_csr(H(h), numUsers, numItems, (const int64_t*)BUF(rowptr), (const int32_t*)BUF(col), (const double*)BUF(val));
}
jint LRK_JNI(setFactors)(JNIEnv* env, jclass c, jlong h, jobject P, jobject Q, jobject bu, jobject bi, jdouble globalMean) {
    (void)c;
    return lrk_set_factors(H(h), (const double*)BUF(P), (const double*)BUF(Q), (const double*)BUF(bu), (const double*)BUF(bi), globalMean);
}
jint LRK_JNI(getFactors)(JNIEnv* env, jclass c, jlong h, jobject P, jobject Q, jobject bu, jobject bi) {
    (void)c;
    return lrk_get_factors(H(h), (double*)BUF(P), (double*)BUF(Q), (double*)BUF(bu), (double*)BUF(bi));
}
jint LRK_JNI(stageStats)(JNIEnv* env, jclass c, jlong h, jlongArray out4) {
    int64_t v[4] = {0, 0, 0, 0};
    const int rc = lrk_stage_stats(H(h), v);
    (void)c;
    if (rc == LRK_OK) (*env)->SetLongArrayRegion(env, out4, 0, 4, (const jlong*)v);
    return rc;
}

jint LRK_JNI(setParam)(JNIEnv* env, jclass c, jlong h, jbyteArray nameUtf8, jdouble value) {
    char name[64];
    jsize n = (*env)->GetArrayLength(env, nameUtf8);
    (void)c;
    if (n > (jsize)sizeof name - 1) n = (jsize)sizeof name - 1;
    (*env)->GetByteArrayRegion(env, nameUtf8, 0, n, (jbyte*)name);
    name[n] = 0;
    return lrk_set_param(H(h), name, value);
}

static void jni_copy_name(JNIEnv* env, jbyteArray nameUtf8, char* name, jsize cap) {
    jsize n = (*env)->GetArrayLength(env, nameUtf8);
    if (n > cap - 1) n = cap - 1;
    (*env)->GetByteArrayRegion(env, nameUtf8, 0, n, (jbyte*)name);
    name[n] = 0;
}
jint LRK_JNI(setMatrix)(JNIEnv* env, jclass c, jlong h, jbyteArray nameUtf8, jobject values) {
    char name[64];
    (void)c;
    jni_copy_name(env, nameUtf8, name, (jsize)sizeof name);
    return lrk_set_matrix(H(h), name, (const double*)BUF(values));
}
jint LRK_JNI(getMatrix)(JNIEnv* env, jclass c, jlong h, jbyteArray nameUtf8, jobject values) {
    char name[64];
    (void)c;
    jni_copy_name(env, nameUtf8, name, (jsize)sizeof name);
    return lrk_get_matrix(H(h), name, (double*)BUF(values));
}

/* ---- training ----------------------------------------------------------------------- */
jint LRK_JNI(sgdEpoch)(JNIEnv* env, jclass c, jlong h, jfloat lr, jfloat regU, jfloat regI, jdouble regB, jint epochIdx, jdoubleArray lossOut) {
    double loss = 0.0;
    const int rc = lrk_sgd_epoch(H(h), lr, regU, regI, regB, epochIdx, &loss);
    (void)c;
    (*env)->SetDoubleArrayRegion(env, lossOut, 0, 1, &loss);     /* also on LRK_ERR_DIVERGED: isConverged() throws the reference's exception */
    return rc;
}
jint LRK_JNI(sgdEpochs)(JNIEnv* env, jclass c, jlong h, jint nEpochs, jfloat lr, jfloat decay, jfloat maxLr, jfloat regU, jfloat regI,
                        jdouble regB, jint firstEpochIdx, jdoubleArray lossesOut) {
    double* l = (double*)PIN(lossesOut);
    const int rc = lrk_sgd_epochs(H(h), nEpochs, lr, decay, maxLr, regU, regI, regB, firstEpochIdx, l);
    (void)c;
    UNPIN(lossesOut, l, 0);
    return rc;
}
jint LRK_JNI(lastEpochMs)(JNIEnv* env, jclass c, jlong h, jfloatArray out1) {
    float ms = 0.f;
    const int rc = lrk_last_epoch_ms(H(h), &ms);
    (void)c;
    if (rc == LRK_OK) (*env)->SetFloatArrayRegion(env, out1, 0, 1, &ms);
    return rc;
}
jint LRK_JNI(sgdSafeguardState)(JNIEnv* env, jclass c, jlong h, jlongArray out2) {
    int32_t div = 1; int64_t rb = 0; jlong v[2];
    const int rc = lrk_sgd_safeguard_state(H(h), &div, &rb);
    (void)c;
    v[0] = div; v[1] = rb;
    if (rc == LRK_OK) (*env)->SetLongArrayRegion(env, out2, 0, 2, v);
    return rc;
}
jint LRK_JNI(launchCount)(JNIEnv* env, jclass c, jlong h, jlongArray out1) {
    uint64_t n = 0; jlong v;
    const int rc = lrk_launch_count(H(h), &n);
    (void)c;
    v = (jlong)n;
    if (rc == LRK_OK) (*env)->SetLongArrayRegion(env, out1, 0, 1, &v);
    return rc;
}
jint LRK_JNI(bprPeekSamples)(JNIEnv* env, jclass c, jlong h, jint epochIdx, jlong first, jlong n, jintArray out3n) {
    int32_t* o = (int32_t*)PIN(out3n);
    const int rc = lrk_bpr_peek_samples(H(h), epochIdx, first, n, o);
    (void)c;
    UNPIN(out3n, o, 0);
    return rc;
}

jint LRK_JNI(debugStream)(JNIEnv* env, jclass c, jlong h, jobject su, jobject si, jobject sr, jobject units, jlong maxUnits, jlongArray nUnitsOut1) {
    int64_t n = 0; jlong v;
    const int rc = lrk_debug_stream(H(h), (int32_t*)BUF(su), (int32_t*)BUF(si), (float*)BUF(sr), (int32_t*)BUF(units), maxUnits, &n);
    (void)c;
    v = (jlong)n;
    (*env)->SetLongArrayRegion(env, nUnitsOut1, 0, 1, &v);
    return rc;
}

/* ---- prediction --------------------------------------------------------------------- */
jint LRK_JNI(predictPairs)(JNIEnv* env, jclass c, jlong h, jintArray users, jintArray items, jlong n, jdoubleArray out) {
    int32_t* u = (int32_t*)PIN(users); int32_t* i = (int32_t*)PIN(items); double* o = (double*)PIN(out);
    const int rc = lrk_predict_pairs(H(h), u, i, n, o);
    (void)c;
    UNPIN(out, o, 0); UNPIN(items, i, JNI_ABORT); UNPIN(users, u, JNI_ABORT);
    return rc;
}
jint LRK_JNI(evalRating)(JNIEnv* env, jclass c, jlong h, jint numUsers, jobject rowptr, jobject col, jobject val, jdouble minRate,
                         jdouble maxRate, jobject predOutOrNull, jdoubleArray rmseMaeOut) {
    double r[2] = {0.0, 0.0};
    const int rc = lrk_eval_rating(H(h), numUsers, (const int64_t*)BUF(rowptr), (const int32_t*)BUF(col), (const double*)BUF(val),
                                   minRate, maxRate, (double*)BUF(predOutOrNull), &r[0], &r[1]);
    (void)c;
    if (rc == LRK_OK) (*env)->SetDoubleArrayRegion(env, rmseMaeOut, 0, 2, r);
    return rc;
}

/* ---- top-N ranking ------------------------------------------------------------------ */
jint LRK_JNI(topn)(JNIEnv* env, jclass c, jlong h, jintArray usersOrNull, jint nq, jint topN, jint excludeTrain, jobject outItems,
                   jobject outScores, jobject outCounts) {
    int32_t* u = (int32_t*)PIN(usersOrNull);
    const int rc = lrk_topn(H(h), u, nq, topN, excludeTrain, (int32_t*)BUF(outItems), (double*)BUF(outScores), (int32_t*)BUF(outCounts));
    (void)c;
    UNPIN(usersOrNull, u, JNI_ABORT);
    return rc;
}
jint LRK_JNI(evalRanking)(JNIEnv* env, jclass c, jlong h, jint topN, jobject tRowptr, jobject tCol, jobject tVal, jobject outItemsOrNull,
                          jobject outScoresOrNull, jobject outCountsOrNull, jdoubleArray outMeasures8) {
    double m[8];
    const int rc = lrk_eval_ranking(H(h), topN, (const int64_t*)BUF(tRowptr), (const int32_t*)BUF(tCol), (const double*)BUF(tVal),
                                    (int32_t*)BUF(outItemsOrNull), (double*)BUF(outScoresOrNull), (int32_t*)BUF(outCountsOrNull), m);
    (void)c;
    if (rc == LRK_OK) (*env)->SetDoubleArrayRegion(env, outMeasures8, 0, 8, m);
    return rc;
}
jint LRK_JNI(topnStats)(JNIEnv* env, jclass c, jlong h, jlongArray fastFallbackOut2, jfloatArray msOut1) {
    int64_t a = 0, b = 0; float ms = 0.f; jlong v[2];
    const int rc = lrk_topn_stats(H(h), &a, &b, &ms);
    (void)c;
    v[0] = a; v[1] = b;
    if (rc == LRK_OK) { (*env)->SetLongArrayRegion(env, fastFallbackOut2, 0, 2, v); (*env)->SetFloatArrayRegion(env, msOut1, 0, 1, &ms); }
    return rc;
}
jint LRK_JNI(topnPhaseMs)(JNIEnv* env, jclass c, jlong h, jfloatArray out6) {
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int rc = lrk_topn_phase_ms(H(h), v);
    (void)c;
    if (rc == LRK_OK) (*env)->SetFloatArrayRegion(env, out6, 0, 6, v);
    return rc;
}

/* ---- measurement aid ------------------------------------------------------------------ */
jint LRK_JNI(probeL2)(JNIEnv* env, jclass c, jlong h, jlong workingSetBytes, jint rowFloats, jdoubleArray outGbps3) {
    double v[3] = {0.0, 0.0, 0.0};
    const int rc = lrk_probe_l2(H(h), (uint64_t)workingSetBytes, rowFloats, v);
    (void)c;
    if (rc == LRK_OK) (*env)->SetDoubleArrayRegion(env, outGbps3, 0, 3, v);
    return rc;
}

/* ---- multi-GPU ---------------------------------------------------------------------- */
jint LRK_JNI(commUniqueId)(JNIEnv* env, jclass c, jbyteArray out128) {
    uint8_t id[128];
    const int rc = lrk_comm_unique_id(id);
    (void)c;
    if (rc == LRK_OK) (*env)->SetByteArrayRegion(env, out128, 0, 128, (const jbyte*)id);
    return rc;
}
jint LRK_JNI(commInit)(JNIEnv* env, jclass c, jlong h, jint rank, jint world, jbyteArray id128) {
    uint8_t id[128];
    (void)c;
    (*env)->GetByteArrayRegion(env, id128, 0, 128, (jbyte*)id);
    return lrk_comm_init(H(h), rank, world, id);
}
