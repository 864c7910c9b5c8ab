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


def test_shardset_slices_cover_the_batch_once():
    """e2fm_shardset_slice: contiguous slices of ceil(n / world) patterns, every pattern in exactly one (host-only call)"""
    from e2fm_b200 import capi
    for n in (0, 1, 7, 8, 9, 1000, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            seen = 0
            for r in range(world):
                lo, cnt = capi.shardset_slice(n, world, r)
                assert lo == min(n, -(-n // world) * r) and lo == seen
                seen += cnt
            assert seen == n


@pytest.mark.parametrize("world", [1, 2, 3])
def test_shardset_merge_model(world, oracle_mod):
    """The arithmetic of shard_merge_kernel (csrc/shardset.cu) on the CPU: rank q receives, from every shard r, the shard's CSR
    offsets of slice q and the occurrence triples between its first and last offset; the merged offset of a pattern is the SUM over
    the shards of (offset - offset of the slice's first pattern), shard r's occurrences follow those of shards 0..r-1. Fed with the
    oracle's answers on the golden shards (incl. an EMPTY shard when world = 3 and a sequence index of -1), the result must equal
    the reference on the unsharded collection."""
    from conftest import read_patterns
    key = open(os.path.join(GOLDEN, "key.bin"), "rb").read()
    res = oracle_mod.read_result_file(os.path.join(GOLDEN, "col3_k4.res"))
    pats = read_patterns(os.path.join(GOLDEN, "col3_k4.pat"))
    keep = [i for i, p in enumerate(pats) if len(p) in (12, 20)]
    if world == 1:
        shards = [(os.path.join(GOLDEN, "col3_k4.efm"), 0, 0)]
    else:
        plan = sharded.plan_shards([30011, 20003, 9001], 2, 4)
        shards = [(os.path.join(GOLDEN, f"col3_k4_shard{r}of2.efm"), plan[r][0], plan[r][2]) for r in range(2)]
    local = []
    for path, first_seq, text_base in shards:
        orc = oracle_mod.OracleIndex(open(path, "rb").read(), key)
        off, pos, seq, so = [0], [], [], []
        for i in keep:
            _, p = orc.search(pats[i], True)
            for x in p:
                s, o = orc.map_position(int(x))
                pos.append(int(x) + text_base); seq.append(s + first_seq if s >= 0 else -1); so.append(o)
            off.append(len(pos))
        orc.close()
        local.append((np.array(off), np.array(pos, dtype=np.int64), np.array(seq, dtype=np.int64), np.array(so, dtype=np.int64)))
    if world == 3:                                   # a shard without a single occurrence
        local.append((np.zeros(len(keep) + 1, dtype=np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64)))
        local[0][2][:1] = -1                          # and a sequence index the reference left at -1 travels unchanged
    n = len(keep)
    per = -(-n // world)
    for q in range(world):
        lo, hi = min(n, per * q), min(n, per * (q + 1))
        recv = [(off[lo:hi + 1], pos[off[lo]:off[hi]], seq[off[lo]:off[hi]], so[off[lo]:off[hi]]) for off, pos, seq, so in local]
        m_csr = sum(o - o[0] for o, _, _, _ in recv)
        m_pos = np.zeros(int(m_csr[-1]), dtype=np.int64); m_seq = m_pos.copy(); m_so = m_pos.copy()
        for i in range(hi - lo):
            dst = int(m_csr[i])
            for o, p_, s_, f_ in recv:
                b, c = int(o[i] - o[0]), int(o[i + 1] - o[i])
                m_pos[dst:dst + c], m_seq[dst:dst + c], m_so[dst:dst + c] = p_[b:b + c], s_[b:b + c], f_[b:b + c]
                dst += c
        want_pos = np.concatenate([res["positions"][int(res["pos_offsets"][i]):int(res["pos_offsets"][i + 1])] for i in keep[lo:hi]] or [np.zeros(0)])
        want_seq = np.concatenate([res["seq"][int(res["pos_offsets"][i]):int(res["pos_offsets"][i + 1])] for i in keep[lo:hi]] or [np.zeros(0)])
        assert np.array_equal(np.diff(m_csr), [int(res["pos_offsets"][i + 1] - res["pos_offsets"][i]) for i in keep[lo:hi]])
        assert np.array_equal(m_pos, want_pos.astype(np.int64))
        if not (world == 3 and q == 0):
            assert np.array_equal(m_seq, want_seq.astype(np.int64))
        else:
            assert m_seq[0] == -1 and np.array_equal(m_seq[1:], want_seq.astype(np.int64)[1:])
