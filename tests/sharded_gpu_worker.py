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
from e2fm_b200 import capi, sharded  # noqa: E402
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
    # ---- the same through the C ABI: e2fm_shardset_* owns the NCCL communicator, every rank supplies ITS slice of the batch
    ids = [capi.shardset_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ss = capi.ShardSet(col.ix, ids[0], rank, world, first_seq, text_base)
    for L in (8, 12, 20, 31):
        keep = [i for i, p in enumerate(all_pats) if len(p) == L]
        acgt = [i for i in keep if all(ch in b"ACGT" for ch in all_pats[i])]
        for packed, sel in ((False, keep), (True, acgt)):
            if not sel:
                continue
            arr = np.frombuffer(b"".join(all_pats[i] for i in sel), dtype=np.uint8).reshape(len(sel), L).copy()
            lo, cnt = capi.shardset_slice(len(sel), world, rank)
            mine = np.ascontiguousarray(arr[lo:lo + cnt])
            if packed:
                mine, bad = capi.pack_patterns(mine)
                assert bad == 0
            sel = np.array(sel)
            for locate in (True, False):
                try:
                    r = ss.search(mine.reshape(-1), len(sel), L, packed, locate)
                except capi.E2FMError:
                    assert packed and L == 8, (L, packed)     # too short for a whole seed in every shift: ASCII entry points only
                    continue
                out = r.fetch()
                mysel = sel[lo:lo + cnt]                      # the result covers this rank's slice of the batch
                assert np.array_equal(out["counts"], res["counts"][mysel]), (rank, L, packed, "shardset counts")
                if locate:
                    segs = [(int(res["pos_offsets"][i]), int(res["pos_offsets"][i + 1])) for i in mysel]
                    cat = lambda a: np.concatenate([a[x:y] for x, y in segs])  # noqa: E731
                    assert np.array_equal(np.diff(out["pos_offsets"].astype(np.int64)), np.array([b - a for a, b in segs])), (rank, L)
                    assert np.array_equal(out["positions"], cat(res["positions"])), (rank, L, packed, "shardset positions")
                    assert np.array_equal(out["seq"], cat(res["seq"])) and np.array_equal(out["off"], cat(res["off"])), (rank, L)
                r.free()
    assert ss.last_ms()["total"] > 0
    ss.close()
    col.close()
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
