"""bench.py's driver-facing contract, as far as it can be checked without a GPU: the reference arm runs the unmodified
reference on the host cores and prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "e2fm_ref")), reason="oracle/_ref not built")
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "ecoli", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "patterns/sec" and d["unit"] == "patterns/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "patterns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "configs[1]" in d["config"]["workload"]
    assert d["config"]["sample_patterns"] == d["config"]["patterns"] == 1_000_000   # the whole batch: the CPU finishes it in seconds


def test_reference_arm_on_other_ranks_is_silent():
    """under torchrun only rank 0 runs the CPU reference; the other ranks exit 0 without output"""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cpu_affinity_helper_never_fails_and_restores():
    """bench.bind_to_gpu_numa is an optimisation: without a GPU (or without a NUMA node for it) it reports why and changes nothing"""
    sys.path.insert(0, ROOT)
    import bench
    before = os.sched_getaffinity(0)
    restore, note = bench.bind_to_gpu_numa(0)
    try:
        assert isinstance(note, str) and note
        if restore is None:
            assert os.sched_getaffinity(0) == before
        else:
            assert os.sched_getaffinity(0) <= before and restore == before
    finally:
        os.sched_setaffinity(0, before)


def test_shared_host_buffers_defer_page_locking(tmp_path):
    """SharedHost(defer=True) hands out the shared arrays unlocked, so that every rank can first-touch its own part; world 1 uses the
    library's pinned allocator, which needs the CUDA runtime — here only the bookkeeping of the deferred mode is checked"""
    sys.path.insert(0, ROOT)
    import numpy as np

    import bench

    class NoCapi:
        pass
    sh = bench.SharedHost(f"pytest_{os.getpid()}", 2, 0, NoCapi(), defer=True)
    try:
        a = sh.array("x", (4, 8), np.uint32)
        assert a.shape == (4, 8) and a.dtype == np.uint32 and len(sh._pending) == 1
        a[1] = 7                                           # writable before it is page-locked
        other = bench.SharedHost(f"pytest_{os.getpid()}", 2, 1, NoCapi(), defer=True)
        b = other.array("x", (4, 8), np.uint32)
        assert b[1, 3] == 7                                # the same pages
    finally:
        for s_ in (sh,):
            s_._pending = []
            for name in list(s_.arrays):
                try:
                    os.remove(f"/dev/shm/e2fm_bench_{s_.tag}_{name}")
                except OSError:
                    pass
