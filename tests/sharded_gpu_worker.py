"""torchrun worker for the 2-GPU sharded test (tests/test_gpu_parity.py::test_sharded_two_gpus):
each rank loads one golden shard on its own GPU, rank 0's batch is broadcast over NCCL, the merged result
must equal the reference's answer on the unsharded collection (tests/golden/col3_k4.res)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLDEN, read_patterns  # noqa: E402
from e2fm_b200 import sharded  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    assert world == 2
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    with open(os.path.join(GOLDEN, "key.bin"), "rb") as f:
        key = f.read()
    with open(os.path.join(GOLDEN, f"col3_k4_shard{rank}of2.efm"), "rb") as f:
        raw = f.read()
    first_seq, _, text_base = sharded.plan_shards([30011, 20003, 9001], world, 4)[rank]
    col = sharded.ShardedCollection(raw, key, local, first_seq, text_base)
    res = oracle.read_result_file(os.path.join(GOLDEN, "col3_k4.res"))
    all_pats = read_patterns(os.path.join(GOLDEN, "col3_k4.pat"))
    for L in (8, 12, 20, 31):
        keep = np.array([i for i, p in enumerate(all_pats) if len(p) == L])
        shape = (len(keep), L)
        pats0 = None
        if rank == 0:
            pats0 = torch.from_numpy(np.frombuffer(b"".join(all_pats[i] for i in keep), dtype=np.uint8).reshape(shape).copy())
        for locate, pieces in ((True, None), (False, None), (True, 3)):
            m = col.search_batch(pats0, shape, locate, pieces)
            assert np.array_equal(m.counts.cpu().numpy(), res["counts"][keep].astype(np.int64)), (rank, L)
            if locate:
                want_pos, want_seq, want_off = [], [], []
                for i in keep:
                    a, b = int(res["pos_offsets"][i]), int(res["pos_offsets"][i + 1])
                    want_pos += list(res["positions"][a:b]); want_seq += list(res["seq"][a:b]); want_off += list(res["off"][a:b])
                assert np.array_equal(m.positions.cpu().numpy(), np.array(want_pos, dtype=np.int64)), (rank, L)
                assert np.array_equal(m.seq.cpu().numpy(), np.array(want_seq, dtype=np.int64)), (rank, L)
                assert np.array_equal(m.off.cpu().numpy(), np.array(want_off, dtype=np.int64)), (rank, L)
    col.close()
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
