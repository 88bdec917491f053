package edu.uci.ics.genomix.hyracks.gpu;

import java.nio.ByteBuffer;

/** Native binding of libgenomix_gb.so (include/genomix_gb.h) through jni/genomix_gb_jni.c. One ctx per task thread. */
public final class GenomixGb {
    static {
        System.loadLibrary("genomix_gb_jni");
    }

    private GenomixGb() {
    }

    public static native long create(int kmerLength, int device, int rank, int nRanks, long expectedKmers);

    /** lines exactly as ReadsKeyValueParserFactory.parse sees them, '\n' terminated, in a direct buffer */
    public static native void pushLines(long ctx, ByteBuffer direct, int nBytes);

    /** whole-record fastq chunks (direct buffers); r2 may be null; ids are 4*(firstRecord+i)+2 */
    public static native void pushFastq(long ctx, ByteBuffer r1, int n1, ByteBuffer r2, int n2, long firstRecord);

    public static native void finish(long ctx);

    /** fills one Hyracks frame with (Kmer, Node) tuples; returns the tuple count, 0 at the end */
    public static native int nextFrame(long ctx, long[] cursor, byte[] frame, int frameSize);

    /** copies whole SequenceFile records (recordLength|keyLength|VKmer|Node) into a direct buffer; returns bytes */
    public static native int nextRecords(long ctx, long[] cursor, ByteBuffer direct, int capacity);

    public static native void writeSequenceFile(long ctx, String path, int nParts, int part);

    public static native byte[] mgUniqueId();

    public static native void mgInit(long ctx, byte[] ncclUniqueId128);

    public static native void mgExchange(long ctx);

    public static native void destroy(long ctx);
}
