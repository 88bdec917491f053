"""torchrun worker for tests/test_gpu_multirank.py: every rank builds from its shard, rank 0 checks the union of
all ranks' records against the oracle on the whole input."""
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import genomix_b200 as gx
    from genomix_b200 import multigpu

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = pickle.load(open(sys.argv[1], "rb"))
    results = []
    for k, text in cases:
        shard = multigpu.shard_lines(text, rank, world)
        stream = multigpu.build_graph_distributed(k, shard, dist, local)
        recs = gx.types.canonical_records(stream)
        gathered = [None] * world
        dist.all_gather_object(gathered, recs)
        if rank == 0:
            union = {}
            for part in gathered:
                for key, val in part.items():
                    assert key not in union, "a key was emitted by two ranks"
                    union[key] = val
            results.append(union)
    if rank == 0:
        pickle.dump(results, open(sys.argv[2], "wb"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
