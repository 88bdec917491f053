package edu.uci.ics.genomix.hyracks.gpu;

import java.io.IOException;

import org.apache.hadoop.fs.FSDataInputStream;
import org.apache.hadoop.fs.FSDataOutputStream;
import org.apache.hadoop.fs.FileSystem;
import org.apache.hadoop.fs.Path;
import org.apache.hadoop.mapred.FileOutputFormat;
import org.apache.hadoop.mapred.JobConf;

/**
 * Hands the 128-byte NCCL unique id from partition 0's task to the other partition tasks of one graph-build job through
 * the job's own file system (the tasks share nothing else: the client that generated the job has neither a GPU nor NCCL).
 * The id sits next to the job's output as <output>/../_genomix_gb_nccl_id.<output name>; it is written to a temporary name
 * and renamed, so that a reader never sees a partial file, and removed again by partition 0 once every rank is connected
 * (GpuReadsKeyValueParserFactory calls done() after gx_mg_init returned, which is collective).
 */
final class NcclIdExchange {
    private static final long POLL_MS = 50, TIMEOUT_MS = 10 * 60 * 1000;

    private NcclIdExchange() {
    }

    private static Path idPath(JobConf conf) {
        Path out = FileOutputFormat.getOutputPath(conf);
        return new Path(out.getParent(), "_genomix_gb_nccl_id." + out.getName());
    }

    static byte[] publish(JobConf conf, byte[] id) throws IOException {
        Path p = idPath(conf), tmp = new Path(p.getParent(), p.getName() + ".tmp");
        FileSystem fs = p.getFileSystem(conf);
        fs.delete(p, false); // a leftover of an earlier, failed job
        FSDataOutputStream o = fs.create(tmp, true);
        o.write(id);
        o.close();
        if (!fs.rename(tmp, p)) {
            throw new IOException("cannot publish the NCCL id at " + p);
        }
        return id;
    }

    static byte[] await(JobConf conf) throws IOException {
        Path p = idPath(conf);
        FileSystem fs = p.getFileSystem(conf);
        long deadline = System.currentTimeMillis() + TIMEOUT_MS;
        while (!fs.exists(p)) {
            if (System.currentTimeMillis() > deadline) {
                throw new IOException("partition 0 did not publish the NCCL id at " + p);
            }
            try {
                Thread.sleep(POLL_MS);
            } catch (InterruptedException e) {
                throw new IOException(e);
            }
        }
        byte[] id = new byte[128];
        FSDataInputStream in = fs.open(p);
        in.readFully(id);
        in.close();
        return id;
    }

    static void done(JobConf conf) throws IOException {
        Path p = idPath(conf);
        p.getFileSystem(conf).delete(p, false);
    }
}
