/* sfm_jni.c -- JNI glue for libsparkfm_b200.so: one native method per C-ABI export of
 * include/sparkfm_b200.h, for the JVM the reference actually pins (Scala 2.10 / Spark 1.2.0 =>
 * Java 7/8, /root/reference/build.sbt:7-11), where java.lang.foreign does not exist.
 * Scala side: scala/io/edstud/spark/fm/gpu/SfmJni.scala (`@native def` per function below).
 *
 * Conventions
 *   handle           jlong (the sfm_handle*), 0 = none
 *   bulk arrays      direct java.nio buffers in native byte order (ByteBuffer.allocateDirect /
 *                    sfm_host_alloc'd pinned memory wrapped by hostAlloc below); null = NULL.
 *                    Nothing is copied: the address goes straight to the library.
 *   scalar outputs   1-element (or n-element) primitive arrays, written with Set*ArrayRegion
 *   return value     the library's int32 status; sfm_last_error(h) has the message
 *
 * Build (on a machine with a JDK):
 *   gcc -O2 -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       jni/sfm_jni.c -Lsparkfm_b200 -lsparkfm_b200 -Wl,-rpath,'$ORIGIN' -o sparkfm_b200/libsfm_jni.so
 * The build image has no JDK: tests/test_host.py only syntax-checks this file against
 * jni/compile_check/jni.h and checks that every export of the header is bound here. */
#include <jni.h>
#include <stdint.h>
#include <string.h>

#include "sparkfm_b200.h"

#define H(x) ((sfm_handle*)(intptr_t)(x))
#define FN(name) JNIEXPORT JNICALL Java_io_edstud_spark_fm_gpu_SfmJni_##name

static void* buf(JNIEnv* e, jobject b) { return b ? (*e)->GetDirectBufferAddress(e, b) : NULL; }
static void out_d(JNIEnv* e, jdoubleArray a, const double* v, int n) { if (a) (*e)->SetDoubleArrayRegion(e, a, 0, n, v); }
static void out_l(JNIEnv* e, jlongArray a, const int64_t* v, int n) { if (a) (*e)->SetLongArrayRegion(e, a, 0, n, (const jlong*)v); }
static void out_i(JNIEnv* e, jintArray a, const int32_t* v, int n) { if (a) (*e)->SetIntArrayRegion(e, a, 0, n, (const jint*)v); }
static void out_f(JNIEnv* e, jfloatArray a, const float* v, int n) { if (a) (*e)->SetFloatArrayRegion(e, a, 0, n, v); }

/* ---- library / host memory ---- */
jint FN(abiVersion)(JNIEnv* e, jclass c) { return sfm_abi_version(); }
jstring FN(statusString)(JNIEnv* e, jclass c, jint status) { return (*e)->NewStringUTF(e, sfm_status_string(status)); }
jint FN(deviceCount)(JNIEnv* e, jclass c) { return sfm_device_count(); }
/* pinned host memory: address in addrOut[0]; wrapHost(addr, bytes) views it as a direct ByteBuffer */
jint FN(hostAlloc)(JNIEnv* e, jclass c, jlong bytes, jlongArray addrOut) {
    void* p = NULL;
    const jint rc = sfm_host_alloc(&p, (uint64_t)bytes);
    const int64_t a = (int64_t)(intptr_t)p;
    out_l(e, addrOut, &a, 1);
    return rc;
}
jobject FN(wrapHost)(JNIEnv* e, jclass c, jlong addr, jlong bytes) {
    return addr ? (*e)->NewDirectByteBuffer(e, (void*)(intptr_t)addr, bytes) : NULL;
}
jint FN(hostFree)(JNIEnv* e, jclass c, jlong addr) { return sfm_host_free((void*)(intptr_t)addr); }

/* ---- handle ---- */
jint FN(create)(JNIEnv* e, jclass c, jint task, jint k, jint k0, jint k1, jint device, jlong nSlots,
                jfloat reg0, jfloat regw, jfloat regv, jfloat stepSize, jfloat miniBatchFraction,
                jint samplerMode, jlong samplerSeed, jlongArray handleOut) {
    sfm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = SFM_ABI_VERSION;
    cfg.task = task; cfg.k = k; cfg.k0 = k0; cfg.k1 = k1; cfg.device = device; cfg.n_slots = nSlots;
    cfg.reg0 = reg0; cfg.regw = regw; cfg.regv = regv; cfg.step_size = stepSize;
    cfg.mini_batch_fraction = miniBatchFraction; cfg.sampler_mode = samplerMode;
    cfg.sampler_seed = (uint64_t)samplerSeed;
    sfm_handle* h = NULL;
    const jint rc = sfm_create(&cfg, &h);
    const int64_t a = (int64_t)(intptr_t)h;
    out_l(e, handleOut, &a, 1);
    return rc;
}
jint FN(destroy)(JNIEnv* e, jclass c, jlong h) { return sfm_destroy(H(h)); }
jstring FN(lastError)(JNIEnv* e, jclass c, jlong h) { return (*e)->NewStringUTF(e, sfm_last_error(H(h))); }
/* ints: task,k,k0,k1,device,sampler_mode; longs: n_slots,sampler_seed; floats: reg0,regw,regv,step,fraction */
jint FN(getConfig)(JNIEnv* e, jclass c, jlong h, jintArray ints6, jlongArray longs2, jfloatArray floats5) {
    sfm_config cfg;
    const jint rc = sfm_get_config(H(h), &cfg);
    if (rc == SFM_OK) {
        const int32_t iv[6] = {cfg.task, cfg.k, cfg.k0, cfg.k1, cfg.device, cfg.sampler_mode};
        const int64_t lv[2] = {cfg.n_slots, (int64_t)cfg.sampler_seed};
        const float fv[5] = {cfg.reg0, cfg.regw, cfg.regv, cfg.step_size, cfg.mini_batch_fraction};
        out_i(e, ints6, iv, 6); out_l(e, longs2, lv, 2); out_f(e, floats5, fv, 5);
    }
    return rc;
}
jint FN(setHyper)(JNIEnv* e, jclass c, jlong h, jfloat reg0, jfloat regw, jfloat regv, jfloat step, jfloat frac) {
    return sfm_set_hyper(H(h), reg0, regw, regv, step, frac);
}

/* ---- model ---- */
jint FN(initModel)(JNIEnv* e, jclass c, jlong h, jdouble mean, jdouble stdev, jlong seed) {
    return sfm_init_model(H(h), mean, stdev, (uint64_t)seed);
}
jint FN(setModel)(JNIEnv* e, jclass c, jlong h, jfloat w0, jobject w, jobject v) {
    return sfm_set_model(H(h), w0, (const float*)buf(e, w), (const float*)buf(e, v));
}
jint FN(getModel)(JNIEnv* e, jclass c, jlong h, jfloatArray w0Out, jobject w, jobject v) {
    float w0 = 0.f;
    const jint rc = sfm_get_model(H(h), &w0, (float*)buf(e, w), (float*)buf(e, v));
    out_f(e, w0Out, &w0, 1);
    return rc;
}
jint FN(setModelF64)(JNIEnv* e, jclass c, jlong h, jdouble w0, jobject w, jobject v) {
    return sfm_set_model_f64(H(h), w0, (const double*)buf(e, w), (const double*)buf(e, v));
}
jint FN(getModelF64)(JNIEnv* e, jclass c, jlong h, jdoubleArray w0Out, jobject w, jobject v) {
    double w0 = 0.0;
    const jint rc = sfm_get_model_f64(H(h), &w0, (double*)buf(e, w), (double*)buf(e, v));
    out_d(e, w0Out, &w0, 1);
    return rc;
}
jint FN(save)(JNIEnv* e, jclass c, jlong h, jstring path) {
    const char* p = (*e)->GetStringUTFChars(e, path, NULL);
    const jint rc = sfm_save(H(h), p);
    (*e)->ReleaseStringUTFChars(e, path, p);
    return rc;
}
jint FN(load)(JNIEnv* e, jclass c, jstring path, jint device, jlongArray handleOut) {
    const char* p = (*e)->GetStringUTFChars(e, path, NULL);
    sfm_handle* h = NULL;
    const jint rc = sfm_load(p, device, &h);
    (*e)->ReleaseStringUTFChars(e, path, p);
    const int64_t a = (int64_t)(intptr_t)h;
    out_l(e, handleOut, &a, 1);
    return rc;
}

/* ---- scoring / data set ---- */
jint FN(predict)(JNIEnv* e, jclass c, jlong h, jobject rowPtr, jobject idx, jobject val, jlong nRows, jobject out) {
    return sfm_predict(H(h), (const int64_t*)buf(e, rowPtr), (const int32_t*)buf(e, idx),
                       (const float*)buf(e, val), nRows, (float*)buf(e, out));
}
jint FN(loadDataset)(JNIEnv* e, jclass c, jlong h, jobject rowPtr, jobject idx, jobject val, jobject label,
                     jlong nRows, jlong globalRowOffset) {
    return sfm_load_dataset(H(h), (const int64_t*)buf(e, rowPtr), (const int32_t*)buf(e, idx),
                            (const float*)buf(e, val), (const float*)buf(e, label), nRows, globalRowOffset);
}
jint FN(unloadDataset)(JNIEnv* e, jclass c, jlong h) { return sfm_unload_dataset(H(h)); }
jint FN(synthCtrDataset)(JNIEnv* e, jclass c, jlong h, jlong nRows, jlong globalRowOffset, jint nFields,
                         jobject fieldLog2Card, jobject zipfCdf, jobject zipfCdfOff, jlong seed) {
    return sfm_synth_ctr_dataset(H(h), nRows, globalRowOffset, nFields, (const int32_t*)buf(e, fieldLog2Card),
                                 (const uint32_t*)buf(e, zipfCdf), (const int64_t*)buf(e, zipfCdfOff),
                                 (uint64_t)seed);
}
jint FN(getDatasetRows)(JNIEnv* e, jclass c, jlong h, jlong rowLo, jlong rowHi, jobject rowPtr, jobject idx,
                        jobject val, jobject label) {
    return sfm_get_dataset_rows(H(h), rowLo, rowHi, (int64_t*)buf(e, rowPtr), (int32_t*)buf(e, idx),
                                (float*)buf(e, val), (float*)buf(e, label));
}
jint FN(datasetInfo)(JNIEnv* e, jclass c, jlong h, jlongArray rowsNnzOut, jintArray maxIndexOut) {
    int64_t v[2] = {0, 0};
    int32_t mx = -1;
    const jint rc = sfm_dataset_info(H(h), &v[0], &v[1], &mx);
    out_l(e, rowsNnzOut, v, 2); out_i(e, maxIndexOut, &mx, 1);
    return rc;
}
jint FN(predictResident)(JNIEnv* e, jclass c, jlong h, jlong rowLo, jlong rowHi, jobject out) {
    return sfm_predict_resident(H(h), rowLo, rowHi, (float*)buf(e, out));
}
jint FN(evaluate)(JNIEnv* e, jclass c, jlong h, jdoubleArray metrics5) {
    double m[5] = {0, 0, 0, 0, 0};
    const jint rc = sfm_evaluate(H(h), m);
    out_d(e, metrics5, m, 5);
    return rc;
}
jint FN(evaluateAuc)(JNIEnv* e, jclass c, jlong h, jdoubleArray out3) {
    double m[3] = {0, 0, 0};
    const jint rc = sfm_evaluate_auc(H(h), m);
    out_d(e, out3, m, 3);
    return rc;
}

/* ---- learner (FMLearn.learn, fm/FMLearn.scala:10-16) ---- */
jint FN(trainStep)(JNIEnv* e, jclass c, jlong h, jobject rowIds, jlong nIds, jlong iter,
                   jdoubleArray lossOut, jlongArray batchOut) {
    double loss = 0.0;
    int64_t batch = 0;
    const jint rc = sfm_train_step(H(h), (const int64_t*)buf(e, rowIds), nIds, iter, &loss, &batch);
    out_d(e, lossOut, &loss, 1); out_l(e, batchOut, &batch, 1);
    return rc;
}
jint FN(trainStepCsr)(JNIEnv* e, jclass c, jlong h, jobject rowPtr, jobject idx, jobject val, jobject label,
                      jlong nRows, jlong iter, jdoubleArray lossOut, jlongArray batchOut) {
    double loss = 0.0;
    int64_t batch = 0;
    const jint rc = sfm_train_step_csr(H(h), (const int64_t*)buf(e, rowPtr), (const int32_t*)buf(e, idx),
                                       (const float*)buf(e, val), (const float*)buf(e, label), nRows, iter,
                                       &loss, &batch);
    out_d(e, lossOut, &loss, 1); out_l(e, batchOut, &batch, 1);
    return rc;
}
jint FN(stageCsr)(JNIEnv* e, jclass c, jlong h, jint slot, jobject rowPtr, jobject idx, jobject val,
                  jobject label, jlong nRows) {
    return sfm_stage_csr(H(h), slot, (const int64_t*)buf(e, rowPtr), (const int32_t*)buf(e, idx),
                         (const float*)buf(e, val), (const float*)buf(e, label), nRows);
}
jint FN(trainStepStaged)(JNIEnv* e, jclass c, jlong h, jint slot, jlong iter, jdoubleArray lossOut,
                         jlongArray batchOut) {
    double loss = 0.0;
    int64_t batch = 0;
    const jint rc = sfm_train_step_staged(H(h), slot, iter, &loss, &batch);
    out_d(e, lossOut, &loss, 1); out_l(e, batchOut, &batch, 1);
    return rc;
}
jint FN(stageOnehot)(JNIEnv* e, jclass c, jlong h, jint slot, jobject packedIdx, jobject labelBits,
                     jobject labelF32, jlong nRows, jint m, jint idBits) {
    return sfm_stage_onehot(H(h), slot, (const uint32_t*)buf(e, packedIdx), (const uint32_t*)buf(e, labelBits),
                            (const float*)buf(e, labelF32), nRows, m, idBits);
}
jint FN(packOnehot)(JNIEnv* e, jclass c, jobject idx, jobject label, jlong nRows, jint m, jint idBits,
                    jobject packedIdx, jobject labelBits) {
    return sfm_pack_onehot((const int32_t*)buf(e, idx), (const float*)buf(e, label), nRows, m, idBits,
                           (uint32_t*)buf(e, packedIdx), (uint32_t*)buf(e, labelBits));
}
jint FN(train)(JNIEnv* e, jclass c, jlong h, jlong firstIter, jlong nIters, jobject lossHistory) {
    return sfm_train(H(h), firstIter, nIters, (double*)buf(e, lossHistory));
}
jint FN(sampleRows)(JNIEnv* e, jclass c, jlong seed, jlong iter, jdouble fraction, jlong rowLo, jlong rowHi,
                    jobject out, jlongArray nOut) {
    int64_t n = 0;
    const jint rc = sfm_sample_rows((uint64_t)seed, iter, fraction, rowLo, rowHi, (int64_t*)buf(e, out), &n);
    out_l(e, nOut, &n, 1);
    return rc;
}
jint FN(partitionRows)(JNIEnv* e, jclass c, jlong seed, jlong nParts, jlong part, jlong rowLo, jlong rowHi,
                       jobject out, jlongArray nOut) {
    int64_t n = 0;
    const jint rc = sfm_partition_rows((uint64_t)seed, nParts, part, rowLo, rowHi, (int64_t*)buf(e, out), &n);
    out_l(e, nOut, &n, 1);
    return rc;
}
jint FN(gradient)(JNIEnv* e, jclass c, jlong h, jobject rowIds, jlong nIds, jobject gradV, jobject gradW,
                  jfloatArray gradW0Out, jdoubleArray lossSumOut, jlongArray batchOut) {
    float g0 = 0.f;
    double ls = 0.0;
    int64_t batch = 0;
    const jint rc = sfm_gradient(H(h), (const int64_t*)buf(e, rowIds), nIds, (float*)buf(e, gradV),
                                 (float*)buf(e, gradW), &g0, &ls, &batch);
    out_f(e, gradW0Out, &g0, 1); out_d(e, lossSumOut, &ls, 1); out_l(e, batchOut, &batch, 1);
    return rc;
}

/* ---- the reference's own learner: ALS.learn (fm/lib/ALS.scala:15-75) ---- */
jint FN(alsSweep)(JNIEnv* e, jclass c, jlong h, jint flags, jdoubleArray rmseOut) {
    double r = 0.0;
    const jint rc = sfm_als_sweep(H(h), flags, &r);
    out_d(e, rmseOut, &r, 1);
    return rc;
}
jint FN(alsResiduals)(JNIEnv* e, jclass c, jlong h, jobject out, jlong n) {
    return sfm_als_residuals(H(h), (double*)buf(e, out), n);
}

/* ---- multi-GPU (replaces the driver-side .sum(), Model.scala:14) ---- */
jint FN(commUniqueId)(JNIEnv* e, jclass c, jbyteArray id128) {
    uint8_t id[SFM_UNIQUE_ID_BYTES];
    const jint rc = sfm_comm_unique_id(id);
    if (rc == SFM_OK) (*e)->SetByteArrayRegion(e, id128, 0, SFM_UNIQUE_ID_BYTES, (const jbyte*)id);
    return rc;
}
jint FN(commInit)(JNIEnv* e, jclass c, jlong h, jbyteArray id128, jint rank, jint worldSize) {
    uint8_t id[SFM_UNIQUE_ID_BYTES];
    (*e)->GetByteArrayRegion(e, id128, 0, SFM_UNIQUE_ID_BYTES, (jbyte*)id);
    return sfm_comm_init(H(h), id, rank, worldSize);
}
jint FN(commInfo)(JNIEnv* e, jclass c, jlong h, jintArray rankWorldOut) {
    int32_t v[2] = {0, 1};
    const jint rc = sfm_comm_info(H(h), &v[0], &v[1]);
    out_i(e, rankWorldOut, v, 2);
    return rc;
}
jint FN(commMode)(JNIEnv* e, jclass c, jlong h, jintArray modeOut) {
    int32_t m = 0;
    const jint rc = sfm_comm_mode(H(h), &m);
    out_i(e, modeOut, &m, 1);
    return rc;
}
jint FN(commBroadcastModel)(JNIEnv* e, jclass c, jlong h) { return sfm_comm_broadcast_model(H(h)); }

/* ---- LibFM text (FMUtils.loadLibFMFile / saveAsLibFMFile, fm/FMUtils.scala:23-74) ---- */
/* counts5Out: n_rows, nnz, dimension, err_line, (unused) */
jint FN(parseLibfm)(JNIEnv* e, jclass c, jobject text, jlong len, jint numFeatures, jlongArray counts4Out,
                    jobject label, jobject rowPtr, jobject idx, jobject val) {
    int64_t nr = 0, nnz = 0, errl = 0;
    int32_t dim = 0;
    const jint rc = sfm_parse_libfm((const char*)buf(e, text), (uint64_t)len, numFeatures, &nr, &nnz, &dim,
                                    (double*)buf(e, label), (int64_t*)buf(e, rowPtr), (int32_t*)buf(e, idx),
                                    (double*)buf(e, val), &errl);
    const int64_t v[4] = {nr, nnz, dim, errl};
    out_l(e, counts4Out, v, 4);
    return rc;
}
jint FN(formatLibfm)(JNIEnv* e, jclass c, jobject label, jobject rowPtr, jobject idx, jobject val, jlong nRows,
                     jobject out, jlong cap, jlongArray neededOut) {
    uint64_t need = 0;
    const jint rc = sfm_format_libfm((const double*)buf(e, label), (const int64_t*)buf(e, rowPtr),
                                     (const int32_t*)buf(e, idx), (const double*)buf(e, val), nRows,
                                     (char*)buf(e, out), (uint64_t)cap, &need);
    const int64_t n = (int64_t)need;
    out_l(e, neededOut, &n, 1);
    return rc;
}

/* ---- counters / timing ---- */
jint FN(statsGet)(JNIEnv* e, jclass c, jlong h, jlongArray counters8, jdoubleArray ms7) {
    sfm_stats s;
    const jint rc = sfm_stats_get(H(h), &s);
    if (rc == SFM_OK) {
        const int64_t cv[8] = {s.train_steps, s.train_rows, s.train_nnz, s.predict_rows, s.predict_nnz,
                               s.kernel_launches, s.h2d_bytes, s.d2h_bytes};
        const double mv[7] = {s.ms_forward, s.ms_sort, s.ms_reduce, s.ms_allreduce, s.ms_update,
                              s.ms_predict, s.ms_total_train};
        out_l(e, counters8, cv, 8); out_d(e, ms7, mv, 7);
    }
    return rc;
}
jint FN(statsReset)(JNIEnv* e, jclass c, jlong h) { return sfm_stats_reset(H(h)); }
jint FN(setPhaseTiming)(JNIEnv* e, jclass c, jlong h, jint enabled) { return sfm_set_phase_timing(H(h), enabled); }
jint FN(synchronize)(JNIEnv* e, jclass c, jlong h) { return sfm_synchronize(H(h)); }
jint FN(timerStart)(JNIEnv* e, jclass c, jlong h) { return sfm_timer_start(H(h)); }
jint FN(timerStop)(JNIEnv* e, jclass c, jlong h, jfloatArray msOut) {
    float ms = 0.f;
    const jint rc = sfm_timer_stop(H(h), &ms);
    out_f(e, msOut, &ms, 1);
    return rc;
}
