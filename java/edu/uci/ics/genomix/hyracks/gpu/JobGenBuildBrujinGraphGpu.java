package edu.uci.ics.genomix.hyracks.gpu;

import java.io.IOException;
import java.util.Map;

import edu.uci.ics.genomix.data.config.GenomixJobConf;
import edu.uci.ics.genomix.hyracks.graph.dataflow.KmerNodePairSequenceWriterFactory;
import edu.uci.ics.genomix.hyracks.graph.dataflow.ReadsKeyValueParserFactory;
import edu.uci.ics.genomix.hyracks.graph.job.JobGenBuildBrujinGraph;
import edu.uci.ics.hyracks.api.client.NodeControllerInfo;
import edu.uci.ics.hyracks.api.constraints.PartitionConstraintHelper;
import edu.uci.ics.hyracks.api.exceptions.HyracksDataException;
import edu.uci.ics.hyracks.api.exceptions.HyracksException;
import edu.uci.ics.hyracks.api.job.JobSpecification;
import edu.uci.ics.hyracks.dataflow.std.connectors.OneToOneConnectorDescriptor;
import edu.uci.ics.hyracks.hdfs.dataflow.HDFSReadOperatorDescriptor;
import edu.uci.ics.hyracks.hdfs.dataflow.HDFSWriteOperatorDescriptor;
import edu.uci.ics.hyracks.hdfs.scheduler.Scheduler;

/**
 * The fast plan for genomix.conf.hyracksGroupby=GPU: HDFS read (GPU parser) -> HDFS write (the unchanged
 * KmerNodePairSequenceWriterFactory). It replaces the six-operator plan of JobGenBuildBrujinGraph.assignJob
 * (genomix-hyracks/src/main/java/edu/uci/ics/genomix/hyracks/graph/job/JobGenBuildBrujinGraph.java:79-143): the external
 * sort, both pre-clustered group-bys and the M:N hash-partitioning merging connector run inside libgenomix_gb.
 */
public class JobGenBuildBrujinGraphGpu extends JobGenBuildBrujinGraph {
    private static final long serialVersionUID = 1L;
    private final int gpusPerNode;
    private final int numPartitionPerMachine;

    public JobGenBuildBrujinGraphGpu(GenomixJobConf job, Scheduler scheduler, final Map<String, NodeControllerInfo> ncMap,
            int numPartitionPerMachine, int gpusPerNode) throws HyracksDataException {
        super(job, scheduler, ncMap, numPartitionPerMachine);
        this.gpusPerNode = gpusPerNode;
        this.numPartitionPerMachine = numPartitionPerMachine;
    }

    @Override
    public JobSpecification assignJob(JobSpecification jobSpec) throws HyracksException { // public, like JobGenBuildBrujinGraph.java:79
        try {
            int nPartitions = readSchedule.length;
            // JobGen.java:67-69 lays the partitions out node-major (partition p runs on node p % nNodes and is that node's
            // (p / nNodes)-th partition): one partition per GPU, never two NCCL ranks on one device
            int nNodes = ncNodeNames.length / numPartitionPerMachine;
            if (numPartitionPerMachine > gpusPerNode) {
                throw new IllegalArgumentException("threadsPerMachine (" + numPartitionPerMachine + ") exceeds the GPUs per machine ("
                        + gpusPerNode + "): the GPU graph build runs one partition per GPU");
            }
            // the client that generates the job needs neither a GPU nor NCCL: partition 0's task creates the NCCL id and
            // publishes it next to the job's output (NcclIdExchange), the other partitions wait for it
            HDFSReadOperatorDescriptor read = new HDFSReadOperatorDescriptor(jobSpec,
                    ReadsKeyValueParserFactory.readKmerOutputRec, hadoopJobConfFactory.getConf(), getInputSplit(), readSchedule,
                    new GpuReadsKeyValueParserFactory(hadoopJobConfFactory.getConf(), nNodes, nPartitions));
            PartitionConstraintHelper.addAbsoluteLocationConstraint(jobSpec, read, ncNodeNames);

            HDFSWriteOperatorDescriptor write = new HDFSWriteOperatorDescriptor(jobSpec, hadoopJobConfFactory.getConf(),
                    new KmerNodePairSequenceWriterFactory(hadoopJobConfFactory.getConf()));
            PartitionConstraintHelper.addAbsoluteLocationConstraint(jobSpec, write, ncNodeNames);

            jobSpec.connect(new OneToOneConnectorDescriptor(jobSpec), read, 0, write, 0);
            jobSpec.addRoot(write);
            return jobSpec;
        } catch (IOException e) {
            throw new HyracksException(e);
        }
    }
}
