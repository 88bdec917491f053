/* JNI shim over include/genomix_gb.h for edu.uci.ics.genomix.hyracks.gpu.GenomixGb (see java/README.md).
 * Exceptions mirror what the reference's pure-Java path throws. */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>
#include "genomix_gb.h"

#define CTX(h) ((gx_ctx*)(intptr_t)(h))

static void throw_gx(JNIEnv* env, gx_ctx* c, int st) {
    const char* cls = st == GX_ERR_FORMAT ? "java/lang/IllegalStateException"
                    : st == GX_ERR_NUMBER ? "java/lang/NumberFormatException"
                    : (st == GX_ERR_READ_TOO_SHORT || st == GX_ERR_READID_RANGE || st == GX_ERR_INVALID)
                          ? "java/lang/IllegalArgumentException"
                          : "edu/uci/ics/hyracks/api/exceptions/HyracksDataException";
    jclass ex = (*env)->FindClass(env, cls);
    if (!ex) {   /* class not on the path (FindClass left a NoClassDefFoundError pending): fall back to a JDK class */
        (*env)->ExceptionClear(env);
        ex = (*env)->FindClass(env, "java/lang/RuntimeException");
        if (!ex) return;
    }
    (*env)->ThrowNew(env, ex, gx_last_error(c));
}

JNIEXPORT jlong JNICALL Java_edu_uci_ics_genomix_hyracks_gpu_GenomixGb_create(JNIEnv* env, jclass k, jint kmer, jint device,
                                                                               jint rank, jint nRanks, jlong expectedKmers) {
    gx_config cfg = {0};
    cfg.abi_version = GX_ABI_VERSION;
    cfg.kmer_length = kmer;
    cfg.device = device;
    cfg.rank = rank;
    cfg.n_ranks = nRanks;
    cfg.expected_kmers = (uint64_t)expectedKmers;
    gx_ctx* c = NULL;
    int st = gx_create(&cfg, &c);
    if (st) { throw_gx(env, NULL, st); return 0; }
    return (jlong)(intptr_t)c;
}

JNIEXPORT void JNICALL Java_edu_uci_ics_genomix_hyracks_gpu_GenomixGb_pushLines(JNIEnv* env, jclass k, jlong h, jobject buf, jint n) {
    int st = gx_push_lines(CTX(h), (const uint8_t*)(*env)->GetDirectBufferAddress(env, buf), (size_t)n);
    if (st) throw_gx(env, CTX(h), st);
}

JNIEXPORT void JNICALL Java_edu_uci_ics_genomix_hyracks_gpu_GenomixGb_pushFastq(JNIEnv* env, jclass k, jlong h, jobject r1, jint n1,
                                                                                jobject r2, jint n2, jlong firstRecord) {
    const uint8_t* p2 = r2 ? (const uint8_t*)(*env)->GetDirectBufferAddress(env, r2) : NULL;
    int st = gx_push_fastq(CTX(h), (const uint8_t*)(*env)->GetDirectBufferAddress(env, r1), (size_t)n1, p2, (size_t)n2,
                           (uint64_t)firstRecord);
    if (st) throw_gx(env, CTX(h), st);
}

JNIEXPORT void JNICALL Java_edu_uci_ics_genomix_hyracks_gpu_GenomixGb_finish(JNIEnv* env, jclass k, jlong h) {
    int st = gx_finish(CTX(h));
    if (st) throw_gx(env, CTX(h), st);
}

/* Hyracks frames must be array-backed (AggregateKmerAggregateFactory.java:99 reads buffer.array()) */
JNIEXPORT jint JNICALL Java_edu_uci_ics_genomix_hyracks_gpu_GenomixGb_nextFrame(JNIEnv* env, jclass k, jlong h, jlongArray cursor,
                                                                                jbyteArray frame, jint frameSize) {
    jlong cur;
    (*env)->GetLongArrayRegion(env, cursor, 0, 1, &cur);
    /* gx_next_frame blocks on device-to-host copies: it fills a native scratch frame, which is then copied into the Java
     * array -- no JNI critical region is held across the blocking call (that would stall the garbage collector) */
    uint8_t* f = (uint8_t*)malloc((size_t)frameSize);
    if (!f) { throw_gx(env, NULL, GX_ERR_NOMEM); return -1; }
    int32_t n = 0;
    uint64_t c64 = (uint64_t)cur;
    int st = gx_next_frame(CTX(h), &c64, f, frameSize, &n);
    if (!st) (*env)->SetByteArrayRegion(env, frame, 0, frameSize, (const jbyte*)f);
    free(f);
    if (st) { throw_gx(env, CTX(h), st); return -1; }
    cur = (jlong)c64;
    (*env)->SetLongArrayRegion(env, cursor, 0, 1, &cur);
    return n;
}

JNIEXPORT jint JNICALL Java_edu_uci_ics_genomix_hyracks_gpu_GenomixGb_nextRecords(JNIEnv* env, jclass k, jlong h, jlongArray cursor,
                                                                                  jobject buf, jint cap) {
    jlong cur;
    (*env)->GetLongArrayRegion(env, cursor, 0, 1, &cur);
    uint64_t c64 = (uint64_t)cur;
    size_t used = 0;
    int st = gx_next_records(CTX(h), &c64, (uint8_t*)(*env)->GetDirectBufferAddress(env, buf), (size_t)cap, &used);
    if (st) { throw_gx(env, CTX(h), st); return -1; }
    cur = (jlong)c64;
    (*env)->SetLongArrayRegion(env, cursor, 0, 1, &cur);
    return (jint)used;
}

JNIEXPORT void JNICALL Java_edu_uci_ics_genomix_hyracks_gpu_GenomixGb_writeSequenceFile(JNIEnv* env, jclass k, jlong h, jstring path,
                                                                                        jint nParts, jint part) {
    const char* p = (*env)->GetStringUTFChars(env, path, NULL);
    uint64_t written = 0;
    int st = gx_write_sequence_file(CTX(h), p, NULL, nParts, part, &written);
    (*env)->ReleaseStringUTFChars(env, path, p);
    if (st) throw_gx(env, CTX(h), st);
}

JNIEXPORT jbyteArray JNICALL Java_edu_uci_ics_genomix_hyracks_gpu_GenomixGb_mgUniqueId(JNIEnv* env, jclass k) {
    uint8_t id[128];
    if (gx_mg_unique_id(id)) { throw_gx(env, NULL, GX_ERR_CUDA); return NULL; }
    jbyteArray out = (*env)->NewByteArray(env, 128);
    (*env)->SetByteArrayRegion(env, out, 0, 128, (const jbyte*)id);
    return out;
}

JNIEXPORT void JNICALL Java_edu_uci_ics_genomix_hyracks_gpu_GenomixGb_mgInit(JNIEnv* env, jclass k, jlong h, jbyteArray id) {
    jbyte buf[128];
    (*env)->GetByteArrayRegion(env, id, 0, 128, buf);
    int st = gx_mg_init(CTX(h), (const uint8_t*)buf);
    if (st) throw_gx(env, CTX(h), st);
}

JNIEXPORT void JNICALL Java_edu_uci_ics_genomix_hyracks_gpu_GenomixGb_mgExchange(JNIEnv* env, jclass k, jlong h) {
    int st = gx_mg_exchange(CTX(h));
    if (st) throw_gx(env, CTX(h), st);
}

JNIEXPORT void JNICALL Java_edu_uci_ics_genomix_hyracks_gpu_GenomixGb_destroy(JNIEnv* env, jclass k, jlong h) { gx_destroy(CTX(h)); }
