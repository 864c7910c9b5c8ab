"""The N>1 path on CPU: world_size-2 gloo groups drive the same merge code the GPUs run over NCCL
(e2fm_b200/sharded.py). Per-shard results come from the oracle here; the merged result must equal the
reference's answer on the UNSHARDED collection (tests/golden/col3_k4.res), position for position."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT, read_patterns
from e2fm_b200 import sharded


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_plan_shards_is_contiguous_and_balanced():
    plan = sharded.plan_shards([4_641_652] * 100, 8, 4)
    assert [p[1] for p in plan] in ([13, 12, 13, 12, 13, 12, 13, 12], [12, 13, 12, 13, 12, 13, 12, 13]) or sum(p[1] for p in plan) == 100
    assert plan[0][0] == 0 and all(plan[i][0] + plan[i][1] == plan[i + 1][0] for i in range(7))
    assert all(11 <= p[1] <= 14 for p in plan)
    # text bases are the padded lengths of everything before the shard
    pl = sharded.padded_length(4_641_652, 4)
    assert all(p[2] == p[0] * pl for p in plan)
    # more ranks than sequences: trailing shards are empty, nothing is lost
    plan = sharded.plan_shards([10, 20], 4, 4)
    assert sum(p[1] for p in plan) == 2
    assert sharded.plan_shards([30011, 20003, 9001], 2, 4) == [(0, 1, 0), (1, 2, 30016)]


def test_plan_shards_never_starves_a_rank():
    """skewed lengths: one long sequence must not close several groups at once (an empty shard has no index to build and its
    rank would hang the collective); with at least `world` sequences every rank gets one or more"""
    import random
    assert [p[1] for p in sharded.plan_shards([100, 1, 1], 3, 4)] == [1, 1, 1]
    assert [p[1] for p in sharded.plan_shards([1000, 10, 10, 10], 4, 4)] == [1, 1, 1, 1]
    assert [p[1] for p in sharded.plan_shards([1, 1, 1000], 3, 4)] == [1, 1, 1]
    rnd = random.Random(5)
    for _ in range(300):
        world = rnd.randint(1, 8)
        lens = [rnd.choice([1, 3, 10, 1000, 10 ** 6]) for _ in range(rnd.randint(1, 30))]
        plan = sharded.plan_shards(lens, world, 4)
        assert len(plan) == world and plan[0][0] == 0 and sum(p[1] for p in plan) == len(lens)
        assert all(plan[i][0] + plan[i][1] == plan[i + 1][0] for i in range(world - 1))
        if len(lens) >= world:
            assert all(p[1] >= 1 for p in plan), (lens, world, plan)
        base = 0
        for first, cnt, tb in plan:
            assert tb == base
            base += sum(sharded.padded_length(L, 4) for L in lens[first:first + cnt])


def _worker(rank, world, port, locate, q):
    import sys
    sys.path.insert(0, ROOT)
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with open(os.path.join(GOLDEN, "key.bin"), "rb") as f:
            key = f.read()
        with open(os.path.join(GOLDEN, f"col3_k4_shard{rank}of2.efm"), "rb") as f:
            ix = oracle.OracleIndex(f.read(), key)
        all_pats = read_patterns(os.path.join(GOLDEN, "col3_k4.pat"))
        keep = [i for i, p in enumerate(all_pats) if len(p) == 8]       # a fixed-length batch, as the GPU path broadcasts
        shape = (len(keep), 8)
        pats0 = torch.from_numpy(np.frombuffer(b"".join(all_pats[i] for i in keep), dtype=np.uint8).reshape(shape).copy()) if rank == 0 else None
        pats = sharded.broadcast_patterns(pats0, shape, torch.device("cpu"))
        counts, offs, pos = ix.search_batch(pats.numpy(), True)
        seq = np.zeros(len(pos), dtype=np.int64)
        off = np.zeros(len(pos), dtype=np.int64)
        for j, p in enumerate(pos):
            s_, o_ = ix.map_position(int(p))
            seq[j] = s_ if s_ != 0xFFFFFFFF else -1
            off[j] = o_
        local = sharded.LocalResult(torch.from_numpy(counts.astype(np.int64)), torch.from_numpy(offs.astype(np.int64)),
                                    torch.from_numpy(pos.astype(np.int64)), torch.from_numpy(seq), torch.from_numpy(off))
        first_seq, _, text_base = sharded.plan_shards([30011, 20003, 9001], world, 4)[rank]
        m = sharded.merge_results(local, first_seq, text_base, locate=locate)
        q.put((rank, keep, m.counts.numpy(), m.pos_offsets.numpy(), m.positions.numpy(), m.seq.numpy(), m.off.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("locate", [True, False])
def test_sharded_merge_equals_unsharded_reference(oracle_mod, locate):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, locate, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = oracle_mod.read_result_file(os.path.join(GOLDEN, "col3_k4.res"))
    for rank, keep, counts, offs, pos, seq, off in got:
        keep = np.array(keep)
        assert np.array_equal(counts, res["counts"][keep].astype(np.int64)), rank
        if not locate:
            assert pos.size == 0
            continue
        want_pos, want_seq, want_off, want_offs = [], [], [], [0]
        for i in keep:
            a, b = int(res["pos_offsets"][i]), int(res["pos_offsets"][i + 1])
            want_pos += list(res["positions"][a:b]); want_seq += list(res["seq"][a:b]); want_off += list(res["off"][a:b])
            want_offs.append(len(want_pos))
        assert np.array_equal(offs, np.array(want_offs, dtype=np.int64))
        assert np.array_equal(pos, np.array(want_pos, dtype=np.int64))       # unsharded global positions, already sorted
        assert np.array_equal(seq, np.array(want_seq, dtype=np.int64))
        assert np.array_equal(off, np.array(want_off, dtype=np.int64))
    assert sum(int(res["counts"][i]) for i in got[0][1]) >= 20
