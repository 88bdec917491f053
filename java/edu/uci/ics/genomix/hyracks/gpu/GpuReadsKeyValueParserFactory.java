package edu.uci.ics.genomix.hyracks.gpu;

import java.nio.ByteBuffer;

import org.apache.hadoop.io.LongWritable;
import org.apache.hadoop.io.Text;
import org.apache.hadoop.mapred.JobConf;

import edu.uci.ics.genomix.data.config.GenomixJobConf;
import edu.uci.ics.hyracks.api.comm.IFrameWriter;
import edu.uci.ics.hyracks.api.context.IHyracksTaskContext;
import edu.uci.ics.hyracks.api.exceptions.HyracksDataException;
import edu.uci.ics.hyracks.dataflow.common.comm.util.FrameUtils;
import edu.uci.ics.hyracks.hdfs.api.IKeyValueParser;
import edu.uci.ics.hyracks.hdfs.api.IKeyValueParserFactory;
import edu.uci.ics.hyracks.hdfs.dataflow.ConfFactory;

/**
 * Drop-in for ReadsKeyValueParserFactory
 * (genomix-hyracks/src/main/java/edu/uci/ics/genomix/hyracks/graph/dataflow/ReadsKeyValueParserFactory.java:48-267):
 * same interface (IKeyValueParserFactory<LongWritable, Text>), same record descriptor (field 0 = Kmer bytes, field 1 =
 * Node bytes), but the tuples it emits are the FINAL aggregated (Kmer, Node) pairs of this partition: k-mer extraction,
 * group-by and the hash shuffle all happen inside libgenomix_gb on the GPU. With it in place the stock sort / aggregate /
 * M:N-merge operators downstream only ever see unique keys (init, never aggregate) and the fast plan
 * (JobGenBuildBrujinGraphGpu) drops them.
 */
public class GpuReadsKeyValueParserFactory implements IKeyValueParserFactory<LongWritable, Text> {
    private static final long serialVersionUID = 1L;
    private static final int LINE_BUFFER_BYTES = 64 << 20;

    private final ConfFactory confFactory;
    private final int nNodes;
    private final int nPartitions;

    public GpuReadsKeyValueParserFactory(JobConf conf, int nNodes, int nPartitions) throws HyracksDataException {
        this.confFactory = new ConfFactory(conf);
        this.nNodes = nNodes;
        this.nPartitions = nPartitions;
    }

    @Override
    public IKeyValueParser<LongWritable, Text> createKeyValueParser(final IHyracksTaskContext ctx)
            throws HyracksDataException {
        final int k = Integer.parseInt(confFactory.getConf().get(GenomixJobConf.KMER_LENGTH));
        final int partition = ctx.getTaskAttemptId().getTaskId().getPartition();
        // node-major partition layout (JobGen.java:67-69): this node's (partition / nNodes)-th partition -> its own GPU
        final long gx = GenomixGb.create(k, partition / nNodes, partition, nPartitions, 0L);
        if (nPartitions > 1) {
            try {
                // partition 0 creates the NCCL id (it has the GPU and the library), everybody else picks it up
                byte[] ncclUniqueId = partition == 0 ? NcclIdExchange.publish(confFactory.getConf(), GenomixGb.mgUniqueId())
                        : NcclIdExchange.await(confFactory.getConf());
                GenomixGb.mgInit(gx, ncclUniqueId); // collective: returns once every partition has joined
                if (partition == 0) {
                    NcclIdExchange.done(confFactory.getConf());
                }
            } catch (java.io.IOException e) {
                GenomixGb.destroy(gx);
                throw new HyracksDataException(e);
            }
        }
        final ByteBuffer lines = ByteBuffer.allocateDirect(LINE_BUFFER_BYTES);
        final ByteBuffer frame = ctx.allocateFrame(); // heap, array-backed, like the reference (:68-70)
        final int frameSize = ctx.getFrameSize();

        return new IKeyValueParser<LongWritable, Text>() {
            @Override
            public void open(IFrameWriter writer) throws HyracksDataException {
            }

            @Override
            public void parse(LongWritable key, Text value, IFrameWriter writer, String fileString)
                    throws HyracksDataException {
                if (lines.remaining() < value.getLength() + 1) {
                    flush();
                }
                lines.put(value.getBytes(), 0, value.getLength());
                lines.put((byte) '\n');
            }

            private void flush() {
                if (lines.position() > 0) {
                    GenomixGb.pushLines(gx, lines, lines.position()); // throws what the reference's parse() throws
                    lines.clear();
                }
            }

            @Override
            public void close(IFrameWriter writer) throws HyracksDataException {
                try {
                    flush();
                    if (nPartitions > 1) {
                        GenomixGb.mgExchange(gx); // collective over all partition tasks of the job
                    }
                    GenomixGb.finish(gx);
                    long[] cursor = { 0 };
                    while (GenomixGb.nextFrame(gx, cursor, frame.array(), frameSize) > 0) {
                        FrameUtils.flushFrame(frame, writer); // same hand-off as ReadsKeyValueParserFactory.java:241,263
                    }
                } finally {
                    GenomixGb.destroy(gx);
                }
            }
        };
    }
}
