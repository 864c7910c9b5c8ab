#!/usr/bin/env python
"""bench.py — batched pattern search throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload human|collection100|chr1|ecoli|config1] [--impl reference]
                    [--sharded]

Default workload = BASELINE.json configs[3], the north-star configuration: 3.1 Gbp synthetic collection (24 sequences),
10 M patterns of length 50, count + locate. A step = one pass of the hot path over the whole 10 M-pattern batch.

  N = 1    one replica of the index searches the whole batch.
  N > 1    STRONG scaling: every GPU holds a replica, the one batch is partitioned by pattern (contiguous slices), results land
           back in ONE host buffer (POSIX shared memory, page-locked by every rank): replicas + pattern partition, the
           throughput mode of SURVEY.md 8(e). `--sharded` times the other mode (collection sharded by sequence, patterns
           broadcast, NCCL merge) and is also reported as the `sharded` object of the default line when N > 1.

  value         whole-job patterns/s, patterns already resident in HBM (CUDA events on the library's stream, max over ranks)
  e2e           same batch through the C ABI from HOST buffers: H2D of the patterns and D2H of what EFMCollection::search
                returns (per-pattern occurrence offsets, sequence index, offset in the sequence) inside the timed region
  roofline      dominant kernel: 32 B x the sectors it has to gather / its CUDA-event time, against the measured HBM copy peak
                AND against the random-gather rate measured on this box for a working set of the index's size
  parity        the first patterns of the TIMED batch compared with the unmodified reference's own CPU search (exit 1 on mismatch)
  cpu_baseline  the unmodified reference (oracle/_ref/e2fm_ref) on the host cores over a bounded sample of the same batch
--impl reference times that CPU reference as the main line (rank 0 only).
The index is always built by the reference's own CPU builder (north star: index construction stays on the CPU) and cached
under data/bench_cache together with the pattern batch.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from e2fm_b200 import synth  # noqa: E402

CACHE = os.path.join(ROOT, "data", "bench_cache")

HG = [248_956_422, 242_193_529, 198_295_559, 190_214_555, 181_538_259, 170_805_979, 159_345_973, 145_138_636,
      138_394_717, 133_797_422, 135_086_622, 133_275_309, 114_364_328, 107_043_718, 101_991_189, 90_338_345,
      83_257_441, 80_373_285, 58_617_616, 64_444_167, 46_709_983, 50_818_468, 156_040_895, 57_227_415]
WORKLOADS = {
    # BASELINE.json configs[3]: human-genome-sized 3.1 Gbp collection (24 sequences, hg38-like proportions), 10M patterns len 50
    "human": dict(seq_lens=HG, seed=44, n_patterns=10_000_000, pat_len=50, locate=True,
                  name="configs[3]: 3.1 Gbp human-genome-sized synthetic collection (24 sequences), 10M patterns len 50, count+locate"),
    # BASELINE.json configs[1]: E. coli-sized 4.6 Mbp synthetic genome, 1M patterns len 32, count only, 1 GPU
    "ecoli": dict(seq_lens=[4_641_652], seed=43, n_patterns=1_000_000, pat_len=32, locate=False,
                  name="configs[1]: 4.6 Mbp synthetic genome, 1M patterns len 32, count"),
    # BASELINE.json configs[0]: 10 Mbp random ACGT, 10k patterns len 20, count+locate
    "config1": dict(seq_lens=[10_000_000], seed=42, n_patterns=10_000, pat_len=20, locate=True,
                    name="configs[0]: 10 Mbp random ACGT, 10k patterns len 20, count+locate"),
    # BASELINE.json configs[2] (whole collection in one index per GPU)
    "collection100": dict(seq_lens=[4_641_652] * 100, seed=1000, n_patterns=10_000_000, pat_len=24, locate=True,
                          name="configs[2]: 100 x 4.6 Mbp synthetic genomes, 10M patterns len 24, count+locate"),
    # BASELINE.json configs[4]: chr1-sized 248 Mbp single sequence, locate-heavy short patterns (len 12)
    "chr1": dict(seq_lens=[248_956_422], seed=45, n_patterns=1_000_000, pat_len=12, locate=True,
                 name="configs[4]: 248 Mbp chr1-sized synthetic sequence, 1M patterns len 12, count+locate"),
    # a sixth of configs[3] (same pattern shape) for quicker HBM-bound kernel iterations
    "mid50": dict(seq_lens=[30_000_000] * 16, seed=46, n_patterns=4_000_000, pat_len=50, locate=True,
                  name="16 x 30 Mbp synthetic sequences, 4M patterns len 50, count+locate"),
    # a collection100 shard-sized index for quicker HBM-bound runs
    "collection25": dict(seq_lens=[4_641_652] * 25, seed=1000, n_patterns=4_000_000, pat_len=24, locate=True,
                         name="25 x 4.6 Mbp synthetic genomes, 4M patterns len 24, count+locate"),
}
K, BUCKET, MRP = 4, 4096, 2   # the authors' regime (Main.cpp:469,747)
PARITY_PATTERNS = 2000
L2_BYTES = 126 << 20


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_JSON_OUT = None


def claim_stdout():
    """Libraries (NCCL's version banner, for one) write to stdout; the driver wants ONE JSON line there. Keep the real stdout
    for that line and send everything else to stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _JSON_OUT or sys.stdout
    print(json.dumps(line), file=out, flush=True)


def ref_bin():
    p = os.path.join(ROOT, "oracle", "_ref", "e2fm_ref")
    if not os.path.exists(p):
        raise SystemExit("oracle/_ref/e2fm_ref is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "in the authoring container (the reference index builder is required)")
    return p


def wait_for(path, timeout=3600):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise SystemExit("timed out waiting for " + path)
        time.sleep(0.5)


def workload_lens(w):
    return synth.safe_lengths(w["seq_lens"], K, 100 // MRP)


def make_sequences(w, lens=None):
    lens = lens or workload_lens(w)
    seqs = []
    for i, L in enumerate(lens):
        seqs += synth.random_sequences([L], w["seed"] + i)
    return seqs


def build_index(fa_seqs, efm, keyfile, threads):
    t0 = time.time()
    fa = efm[:-4] + ".fa"
    synth.write_fasta(fa, fa_seqs)
    r = subprocess.run([ref_bin(), "build", fa, efm + ".tmp", keyfile, str(K), str(BUCKET), str(MRP), str(threads)],
                       capture_output=True, text=True)
    if r.returncode != 0 or not os.path.exists(efm + ".tmp"):
        raise SystemExit("reference index build failed: " + r.stderr[-400:])
    os.replace(efm + ".tmp", efm)
    os.remove(fa)
    log(f"[bench] reference built {efm} ({os.path.getsize(efm) / 1e6:.1f} MB) in {time.time() - t0:.1f}s")


def prepare(workload: str, rank: int = 0):
    """Synthetic collection -> reference-built index + the pattern batch, both cached on disk (rank 0 makes them, the
    other ranks wait). The batch is SURVEY.md 8(d)'s: substrings sampled inside sequences plus 5 % uniformly random strings,
    shuffled, cut to the configured size."""
    w = WORKLOADS[workload]
    os.makedirs(CACHE, exist_ok=True)
    base = os.path.join(CACHE, f"{workload}_k{K}_b{BUCKET}_m{MRP}")
    keyfile = os.path.join(CACHE, "key.bin")
    efm = base + ".efm"
    patf = base + f"_p{w['n_patterns']}x{w['pat_len']}.npy"
    if rank == 0:
        if not os.path.exists(keyfile):
            with open(keyfile + ".tmp", "wb") as f:
                f.write(synth.test_key())
            os.replace(keyfile + ".tmp", keyfile)
        if not (os.path.exists(efm) and os.path.exists(patf)):
            seqs = make_sequences(w)
            if not os.path.exists(patf):
                t0 = time.time()
                pats = synth.sample_patterns(seqs, w["n_patterns"], w["pat_len"], w["seed"] + 1)
                perm = np.random.Generator(np.random.PCG64(w["seed"] + 2)).permutation(len(pats))[:w["n_patterns"]]
                np.save(patf + ".tmp.npy", pats[perm])
                os.replace(patf + ".tmp.npy", patf)
                del pats
                log(f"[bench] pattern batch {patf} in {time.time() - t0:.1f}s")
            if not os.path.exists(efm):
                build_index(seqs, efm, keyfile, min(32, os.cpu_count() or 2))
            del seqs
    else:
        wait_for(keyfile)
        wait_for(patf)
        wait_for(efm)
    return dict(w=w, efm=efm, keyfile=keyfile, patf=patf, workload=workload)


def load_patterns(prep):
    return np.load(prep["patf"], mmap_mode="r")


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every ~10 ms DURING the timed region (the same counters
    `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints, B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.sm, self.reasons, self.power = [], set(), []
        self.max_sm = None
        self._stop = False
        self._thr = None

    def _loop(self):
        import pynvml as nv
        try:
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except (ValueError, IndexError):
                    idx = self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            while True:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k_, bit in names.items():
                    if r & bit:
                        self.reasons.add(k_)
                if self._stop:
                    break
                time.sleep(0.01)
        except Exception as e:   # noqa: BLE001
            self.reasons.add("nvml unavailable: " + type(e).__name__)

    def start(self):
        import threading
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop = True
        if self._thr:
            self._thr.join(timeout=5)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm, "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None, "reasons": sorted(self.reasons)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------ CPU reference

def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def reference_procs(prep) -> int:
    """processes the host can run side by side: one per core, but each holds the whole index plus what it decodes lazily
    (about 12x the index file: bucketText u32 + charactersPositions u64 per super-character)"""
    cores = os.cpu_count() or 1
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    per_proc = 12 * os.path.getsize(prep["efm"]) + (256 << 20)
    return int(max(1, min(cores, 32, (avail // 2) // per_proc)))


def reference_sample_size(prep, procs: int) -> int:
    """Bounded sample: the whole batch when the CPU finishes it in seconds, else a fixed leading part.
    The cold pass costs one bucket decode per bucket touched, so big indexes get fewer patterns per process."""
    w = prep["w"]
    size = os.path.getsize(prep["efm"])
    if size <= (64 << 20):
        per_proc = w["n_patterns"] if w["pat_len"] >= 20 else 2_000   # the full batch takes the CPU a few seconds
    else:
        per_proc = 500 if w["pat_len"] >= 20 else 100
    return int(min(w["n_patterns"], procs * per_proc))


def run_reference(prep, pats: np.ndarray, locate: bool, procs: int, tmpdir: str, warm_passes: int = 1):
    """P independent processes of the unmodified reference, each over a contiguous 1/P slice (the reference is serial and not
    re-entrant per object); 1 matcher thread each; pass 0 is cold (lazy bucket decode), the others warm.
    Returns (per-pass wall ms = max over processes, [result dict per slice], [slice index arrays], wall seconds)."""
    from oracle import oracle
    slices = np.array_split(np.arange(len(pats)), procs)
    ps = []
    t0 = time.time()
    for i, sl in enumerate(slices):
        pf = os.path.join(tmpdir, f"p{i}.txt")
        synth.write_patterns(pf, np.ascontiguousarray(pats[sl[0]:sl[-1] + 1]) if len(sl) else pats[:0])
        ps.append(subprocess.Popen([ref_bin(), "search", prep["efm"], prep["keyfile"], pf, os.path.join(tmpdir, f"r{i}.res"),
                                    "1" if locate else "0", "1", str(1 + warm_passes)], stdout=subprocess.PIPE,
                                   stderr=subprocess.PIPE, text=True))
    pass_ms = np.zeros(1 + warm_passes)
    for p in ps:
        out, err = p.communicate()
        line = [l for l in out.splitlines() if l.startswith("{")]
        if p.returncode != 0 or not line:
            for q in ps:
                if q.poll() is None:
                    q.kill()
            raise RuntimeError(f"reference search failed (exit code {p.returncode}): {(err or out)[-300:].strip()}")
        pass_ms = np.maximum(pass_ms, np.array(json.loads(line[-1])["pass_ms"], dtype=np.float64))
    res = [oracle.read_result_file(os.path.join(tmpdir, f"r{i}.res")) for i in range(len(slices))]
    return pass_ms, res, slices, time.time() - t0


def reference_arm(args, prep):
    w = prep["w"]
    cores = os.cpu_count() or 1
    procs = reference_procs(prep)
    sample = reference_sample_size(prep, procs)
    pats = np.ascontiguousarray(load_patterns(prep)[:sample])
    with tempfile.TemporaryDirectory() as td:
        t_all = time.time()
        pass_ms, _res, _sl, _wall = run_reference(prep, pats, w["locate"], procs, td, warm_passes=args.warmup + args.steps)
        wall_all = time.time() - t_all
    timed = pass_ms[1 + args.warmup:]
    ms = float(np.mean(timed))
    value = sample / (ms / 1e3)
    cold = sample / (pass_ms[0] / 1e3)
    desc = (f"first {sample} of the {w['n_patterns']} patterns (len {w['pat_len']}) split over {procs} processes of the unmodified "
            f"reference, 1 matcher thread each; a step = one warm pass over that sample (buckets already decoded by the cold pass, "
            f"which ran at {cold:.0f} patterns/s)")
    line = {
        "impl": "reference", "metric": "patterns/sec", "value": value, "unit": "patterns/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": w["name"], "mode": "count+locate" if w["locate"] else "count", "k": K, "bucket_size": BUCKET,
                   "mrp": MRP, "patterns": w["n_patterns"], "pattern_len": w["pat_len"], "sample_patterns": sample},
        "cpu_baseline": {"value": value, "unit": "patterns/s", "cores": procs, "kind": "reference", "sample": desc,
                         "cold_value": cold, "host_cores": cores, "cpu_model": cpu_model(), "per_process": value / procs},
        "e2e": {"value": value, "unit": "patterns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall_all,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ shared host buffer (N > 1: one batch, one result, one host)

class SharedHost:
    """The batch and its result live in ONE host buffer per array: POSIX shared memory mapped by every rank and page-locked
    with cudaHostRegister, so that each GPU copies its slice of the patterns straight out of it and its slice of the result
    straight into it. With one rank it is plain pinned memory from the library."""

    def __init__(self, tag: str, world: int, rank: int, capi, defer: bool = False):
        self.tag, self.world, self.rank, self.capi = tag, world, rank, capi
        self.arrays = {}
        self._keep = []
        # defer: the shared files are page-locked by pin(), AFTER every rank has written its own part — a tmpfs page lives on the
        # NUMA node of the CPU that touches it first, and page-locking touches every page
        self.defer = defer
        self._pending = []

    def array(self, name: str, shape, dtype):
        dtype = np.dtype(dtype)
        shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        if self.world == 1:
            pa = self.capi.PinnedArray(shape, dtype)
            self._keep.append(pa)
            self.arrays[name] = pa.array
            return pa.array
        import torch
        nbytes = max(int(np.prod(shape, dtype=np.int64)) * dtype.itemsize, 4096)
        path = f"/dev/shm/e2fm_bench_{self.tag}_{name}"
        if self.rank == 0:
            with open(path + ".tmp", "wb") as f:
                f.truncate(nbytes)
            os.replace(path + ".tmp", path)
        else:
            wait_for(path, 600)
        mm = np.memmap(path, dtype=np.uint8, mode="r+", shape=(nbytes,))
        if self.defer:
            self._pending.append((name, mm, nbytes))
        else:
            rc = torch.cuda.cudart().cudaHostRegister(mm.ctypes.data, nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister({name}) failed: {rc}")
        self._keep.append(mm)
        a = mm[:int(np.prod(shape, dtype=np.int64)) * dtype.itemsize].view(dtype).reshape(shape)
        self.arrays[name] = a
        return a

    def pin(self):
        import torch
        for name, mm, nbytes in self._pending:
            rc = torch.cuda.cudart().cudaHostRegister(mm.ctypes.data, nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister({name}) failed: {rc}")
        self._pending = []

    def close(self):
        if self.world > 1:
            import torch
            for mm in self._keep:
                try:
                    torch.cuda.cudart().cudaHostUnregister(mm.ctypes.data)
                except Exception:
                    pass
            if self.rank == 0:
                for name in self.arrays:
                    try:
                        os.remove(f"/dev/shm/e2fm_bench_{self.tag}_{name}")
                    except OSError:
                        pass
        self.arrays.clear()
        self._keep.clear()


# ------------------------------------------------------------------ GPU arm

def compare_with_reference(ref_res, slices, got, n_checked):
    """bit-exact comparison of counts, occurrence counts, (sequence, offset) and global positions of the first patterns"""
    mism = 0
    first_bad = None
    for res, sl in zip(ref_res, slices):
        if len(sl) == 0:
            continue
        a, b = int(sl[0]), int(sl[-1]) + 1
        ok = np.array_equal(res["counts"], got["counts"][a:b])
        if got.get("pos_offsets") is not None and ok:
            o0, o1 = int(got["pos_offsets"][a]), int(got["pos_offsets"][b])
            ok = np.array_equal(np.diff(res["pos_offsets"].astype(np.int64)), np.diff(got["pos_offsets"][a:b + 1].astype(np.int64)))
            ok = ok and np.array_equal(res["seq"], got["seq"][o0:o1]) and np.array_equal(res["off"], got["off"][o0:o1])
            if ok and got.get("positions") is not None:
                ok = np.array_equal(res["positions"], got["positions"][o0:o1])
        if not ok:
            # count the patterns that differ
            for j in range(a, b):
                c_ok = int(res["counts"][j - a]) == int(got["counts"][j])
                if got.get("pos_offsets") is not None:
                    r0, r1 = int(res["pos_offsets"][j - a]), int(res["pos_offsets"][j - a + 1])
                    g0, g1 = int(got["pos_offsets"][j]), int(got["pos_offsets"][j + 1])
                    c_ok = c_ok and np.array_equal(res["seq"][r0:r1], got["seq"][g0:g1]) and np.array_equal(res["off"][r0:r1], got["off"][g0:g1])
                if not c_ok:
                    mism += 1
                    if first_bad is None:
                        first_bad = j
    return {"checked": int(n_checked), "mismatches": int(mism), "first_mismatch": first_bad}


def bind_to_gpu_numa(local_rank):
    """Run this process on the CPUs of the NUMA node its GPU hangs off, so that the host buffers it first touches (its slice of the
    shared batch, its result region) are local to that GPU's PCIe root: with every rank on node 0 the copy rate of the far GPUs
    halves (round-1 SCALE). Returns (affinity to restore, description) or (None, why not)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None, "no NUMA node reported for the GPU"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        before = os.sched_getaffinity(0)
        allowed = before & cpus
        if not allowed:
            return None, f"NUMA node {node} has no CPU this process may use"
        os.sched_setaffinity(0, allowed)
        return before, f"NUMA node {node} ({len(allowed)} CPUs) of GPU {bus}"
    except Exception as e:   # noqa: BLE001  (affinity is an optimisation, never a reason to fail)
        return None, f"not bound: {type(e).__name__}: {e}"[:120]


def gpu_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from e2fm_b200 import capi

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev_t = torch.device("cuda", local_rank)

    prep = prepare(args.workload, rank)
    if world > 1:
        dist.barrier()
    w = prep["w"]
    locate = w["locate"]
    with open(prep["efm"], "rb") as f:
        raw = f.read()
    t0 = time.time()
    ix = capi.DeviceIndex(raw, synth.test_key(), local_rank)
    load_s = time.time() - t0
    del raw
    if args.walk:
        ix.set_option(capi.OPT_DENSE, 0)
    if args.chunk:
        ix.set_option(capi.OPT_CHUNK, args.chunk)
    if args.pipe:
        ix.set_option(capi.OPT_PIPE, args.pipe)
    if args.taper:
        ix.set_option(capi.OPT_TAPER, 1)
    info = ix.info
    log(f"[bench r{rank}] index loaded in {load_s:.2f}s: n={info.text_length} buckets={info.n_buckets} "
        f"device_bytes={info.device_bytes / 1e6:.1f} MB load_ms={[round(x, 2) for x in info.load_ms]}")

    # host buffers are first touched from here on: on the CPUs next to this rank's GPU
    restore_affinity, affinity_note = (None, "off") if args.no_numa else bind_to_gpu_numa(local_rank)
    log(f"[bench r{rank}] cpu affinity: {affinity_note}")
    # strong scaling: the ONE batch is partitioned by pattern; rank r owns the contiguous slice [lo, hi)
    npat_all, L = w["n_patterns"], w["pat_len"]
    lo, npat = capi.shardset_slice(npat_all, world, rank)     # contiguous slices of ceil(n / world) patterns
    hi = lo + npat
    all_pats = load_patterns(prep)
    # 2-bit packed patterns (e2fm_search_batch_packed) when the seed search takes this batch shape, else ASCII
    stride = capi.packed_stride(L)
    packed_api = not args.ascii
    if packed_api:
        try:
            probe, _ = capi.pack_patterns(np.ascontiguousarray(all_pats[:1]))
            ix.search_packed(probe, 1, L, locate).free()
        except capi.E2FMError:
            packed_api = False
    unit = stride if packed_api else L
    shm = SharedHost(f"{os.environ.get('MASTER_PORT', '0')}_{args.workload}", world, rank, capi, defer=True)
    h_pats = shm.array("patterns", (npat_all, unit), np.uint8)
    if world > 1:
        dist.barrier()
    mine = np.ascontiguousarray(all_pats[lo:hi])
    if packed_api:
        _, bad = capi.pack_patterns(mine, h_pats[lo:hi].reshape(-1))
        assert bad == 0, "the synthetic batch holds only A C G T"
    else:
        h_pats[lo:hi] = mine
    # what travels over PCIe in the end-to-end call: the same 2-bit codes without the padding (ceil(L / 4) bytes per pattern,
    # E2FM_PACKED_TIGHT; the device widens them stage by stage)
    tight_api = packed_api and args.e2e_api == "hits" and not args.no_tight
    h_tight = None
    if tight_api:
        tstride = capi.packed_tight_stride(L)
        h_tight = shm.array("patterns_tight", (npat_all, tstride), np.uint8)
        _, bad = capi.pack_patterns_tight(mine, h_tight[lo:hi].reshape(-1))
        assert bad == 0
    # result buffers of the whole batch, ONE host buffer per array (every rank writes its own region of it): what
    # EFMCollection::search returns, in 32 bits. Region r holds rank r's occurrences; its offsets index that region.
    per_rank = -(-npat_all // world)
    occ_cap = int((per_rank + 1) * max(2.0, 1.3 * args.occ_per_pattern)) + 1024
    h_counts = shm.array("counts", npat_all, np.uint32)
    h_noff = shm.array("occ_offsets", (world, per_rank + 2), np.uint32) if locate else None
    h_seq = shm.array("seq", (world, occ_cap), np.uint32) if locate else None
    h_off = shm.array("off", (world, occ_cap), np.uint32) if locate else None
    if world > 1:                              # every rank touches its own regions, then all page-lock the shared buffers
        h_counts[lo:hi] = 0
        if locate:
            h_noff[rank] = 0
            h_seq[rank] = 0
            h_off[rank] = 0
        dist.barrier()
        shm.pin()
        dist.barrier()
    d_pats = torch.from_numpy(np.array(h_pats[lo:hi])).to(dev_t)
    del mine
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev_t)   # > 126 MB L2
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    def l2_flush():
        if os.environ.get("E2FM_BENCH_NOFLUSH"):
            return
        flush.add_(1)
        torch.cuda.synchronize()

    def step_device():
        if packed_api:
            r = ix.search_packed_device(d_pats.data_ptr(), npat, L, locate)
        else:
            r = ix.search_device_fixed(d_pats.data_ptr(), npat, L, locate)
        ms = ix.last_batch_ms()
        ctr = ix.last_batch_counters()
        tot = r.total
        r.free()
        return ms, ctr, tot

    e2e_dev_ms = []
    e2e_state = {}

    def step_e2e():
        """host patterns -> host result through ONE call of the C ABI (e2fm_search_hits: H2D, search and D2H pipelined by stage);
        --e2e-api fetch: the two-call form (e2fm_search_batch_packed / _fixed, then e2fm_result_fetch_compact)"""
        t0 = time.perf_counter()
        if args.e2e_api == "hits":
            # locate: a pattern's count is its number of occurrences (every match is located), so only the offsets travel back
            total = ix.search_hits((h_tight if tight_api else h_pats)[lo:hi].reshape(-1), npat, L, 2 if tight_api else int(packed_api), locate,
                                   None if locate else h_counts[lo:hi],
                                   h_noff[rank][:npat + 1] if locate else None, h_seq[rank] if locate else None,
                                   h_off[rank] if locate else None)
        else:
            if packed_api:
                r = ix.search_packed(h_pats[lo:hi].reshape(-1), npat, L, locate)
            else:
                r = ix.search_fixed(h_pats[lo:hi], locate)
            total = r.total
            if total > occ_cap:
                raise SystemExit("bench: occurrence buffer too small; pass a larger --occ-per-pattern")
            if locate:
                r.fetch_compact(None, h_noff[rank][:npat + 1], h_seq[rank][:total], h_off[rank][:total])
            else:
                r.fetch_compact(h_counts[lo:hi], None, None, None)
            r.free()
        dt = time.perf_counter() - t0
        e2e_dev_ms.append(ix.last_batch_ms()["total"])
        ctr = ix.last_batch_counters()
        e2e_state.update(total=total, stages=ctr.get("hits_stages", 0), launches=ctr.get("launches", 0))
        return dt, total

    for _ in range(args.warmup):
        l2_flush(); step_device()
        l2_flush(); step_e2e()

    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if not os.environ.get("E2FM_BENCH_NOSAMPLER"):
        sampler.start()
    wall0 = time.perf_counter()
    phases = ["total", "backward_search", "walk", "expand", "results", "encode", "firstcheck"]
    ph_ms = {k_: 0.0 for k_ in phases}
    ctr_sum = {}
    occ_total = 0
    for _ in range(args.steps):
        l2_flush()
        ms, ctr, tot = step_device()
        for k_ in phases:
            ph_ms[k_] += ms[k_]
        for k_, v in ctr.items():
            ctr_sum[k_] = ctr_sum.get(k_, 0) + v
        occ_total += tot
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall_dev = time.perf_counter() - wall0
    e2e_s = 0.0
    e2e_launches = 0
    for _ in range(args.steps):
        l2_flush()
        if world > 1:
            dist.barrier()     # a step = the whole batch: all ranks start it together
        dt, _tot = step_e2e()
        e2e_s += dt
        e2e_launches += e2e_state.get("launches", 0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()

    # what the host link gives this rank: the H2D copy of its own slice alone (rank 0) and with every rank copying at once —
    # GPUs behind one PCIe switch share its uplink, which bounds the end-to-end figure at N > 1 whatever the kernels do
    def h2d_rate():
        src = torch.from_numpy((h_tight if tight_api else h_pats)[lo:hi].reshape(-1))
        dst = torch.empty(src.numel(), dtype=torch.uint8, device=dev_t)
        best = None
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return src.numel() / best / 1e9
    pcie = {}
    if world > 1:
        dist.barrier()
        if rank == 0:
            pcie["h2d_gbs_rank0_alone"] = round(h2d_rate(), 1)
        dist.barrier()
        mine_gbs = torch.tensor([h2d_rate()], dtype=torch.float64, device=dev_t)
        dist.all_reduce(mine_gbs, op=dist.ReduceOp.MIN)
        pcie["h2d_gbs_slowest_rank_all_copying"] = round(float(mine_gbs[0]), 1)
    else:
        pcie["h2d_gbs_rank0_alone"] = round(h2d_rate(), 1)
    if restore_affinity is not None:           # the reference processes of the parity / CPU-baseline leg and the shard builds use every core
        os.sched_setaffinity(0, restore_affinity)

    # max over ranks (times), sum over ranks (work)
    t = torch.tensor([ph_ms["total"], e2e_s * 1e3, ph_ms["backward_search"], ph_ms["walk"]], dtype=torch.float64, device=dev_t)
    cnt = torch.tensor([occ_total, ctr_sum.get("launches", 0), e2e_launches], dtype=torch.int64, device=dev_t)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max = float(t[0]), float(t[1])
    occ_all, launches_all, e2e_launches_all = int(cnt[0]), int(cnt[1]), int(cnt[2])
    steps = args.steps

    # ---- parity of the TIMED batch + CPU baseline (rank 0, outside the timed region): the unmodified reference searches the first
    # patterns of the batch on the host cores; its answers must equal what the last timed e2e step left in the host buffers
    parity = None
    cpu_base = None
    if rank == 0:
        cores = os.cpu_count() or 1
        procs = reference_procs(prep)
        sample = min(npat, max(PARITY_PATTERNS, 0 if args.no_cpu_baseline else reference_sample_size(prep, procs)))
        # what the last timed e2e step left in the host buffers, plus the global positions of the same patterns from an untimed
        # 64-bit fetch (the timed arrays must agree with it)
        chk = ix.search_packed(h_pats[lo:lo + sample].reshape(-1), sample, L, locate) if packed_api else ix.search_fixed(h_pats[lo:lo + sample], locate)
        full = chk.fetch()
        chk.free()
        # (locate: the timed call brings back occurrences, not counts; the counts compared with the reference are the 64-bit fetch's)
        got = {"counts": full["counts"] if locate else h_counts[:sample].astype(np.uint64)}
        assert np.array_equal(full["counts"], got["counts"]), "timed counts differ from the 64-bit fetch"
        if locate:
            got.update(pos_offsets=h_noff[0][:sample + 1], seq=h_seq[0], off=h_off[0], positions=full["positions"])
            tot = int(h_noff[0][sample])
            assert np.array_equal(full["pos_offsets"], h_noff[0][:sample + 1]) and np.array_equal(full["seq"], h_seq[0][:tot]) and \
                np.array_equal(full["off"], h_off[0][:tot]), "timed occurrence arrays differ from the 64-bit fetch"
        try:
            with tempfile.TemporaryDirectory() as td:
                pass_ms, ref_res, slices, wall = run_reference(prep, np.ascontiguousarray(all_pats[:sample]), locate, procs, td)
            parity = compare_with_reference(ref_res, slices, got, sample)
            parity["against"] = "oracle/_ref/e2fm_ref search (the unmodified reference) on the first patterns of the timed batch"
            v, c = sample / (pass_ms[1] / 1e3), sample / (pass_ms[0] / 1e3)
            cpu_base = {"value": v, "unit": "patterns/s", "cores": procs, "kind": "reference", "host_cores": cores, "cold_value": c,
                        "cpu_model": cpu_model(), "per_process": v / procs,
                        "sample": f"first {sample} patterns of the same batch over {procs} processes of the unmodified reference "
                                  f"(1 matcher thread each), warm pass; cold pass {c:.0f} patterns/s; {wall:.1f}s wall"}
        except RuntimeError as e:
            parity = {"checked": 0, "mismatches": None, "error": str(e)}

    # ---- N > 1: the other multi-GPU mode, the collection sharded by sequence, on the same batch
    sharded_obj = None
    if world > 1 and not args.no_sharded and len(w["seq_lens"]) >= world:
        replica = {"lo": lo, "npat": npat}
        if locate:
            tot = int(h_noff[rank][npat])
            replica.update(noff=h_noff[rank][:npat + 1].copy(), seq=h_seq[rank][:tot].copy(), off=h_off[rank][:tot].copy())
        else:
            replica.update(counts=h_counts[lo:hi].copy())
        ix.close()
        ix = None
        try:
            sharded_obj = sharded_measure(args, rank, world, local_rank, prep, replica)
        except Exception as e:   # noqa: BLE001  (the replica line stands on its own)
            sharded_obj = {"error": f"{type(e).__name__}: {e}"[:300]}
            log(f"[bench r{rank}] sharded mode failed: {sharded_obj['error']}")

    if rank == 0:
        peak, peak_kind = measured_peak()
        value = npat_all * steps / (dev_ms_max / 1e3)
        e2e_value = npat_all * steps / (e2e_ms_max / 1e3)
        roofline = make_roofline(args, capi, info, local_rank, ctr_sum, ph_ms, steps, peak, peak_kind, npat)
        if sharded_obj is not None:
            roofline["note"] += "; the gather microbenchmark ran after the sharded leg freed the replica"
        h2d = int(npat * (tstride if tight_api else unit))
        d2h = int(((npat + 1) * 4 + (occ_total // max(steps, 1)) * 8) if locate else npat * 4)
        line = {
            "metric": "patterns/sec", "value": value, "unit": "patterns/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": w["name"], "mode": "count+locate" if locate else "count", "k": K, "bucket_size": BUCKET, "mrp": MRP,
                       "patterns": npat_all, "pattern_len": L,
                       "parallelism": (f"{world} index replicas, the one batch partitioned by pattern ({npat} per GPU), results gathered in one "
                                       f"shared host buffer" if world > 1 else "one index replica, whole batch"),
                       "l2": "256 MB flush between timed iterations", "cpu_affinity_rank0": affinity_note,
                       "locate_tables": "per-query walks from marked rows" if args.walk else "dense sa[]/text[] built at load",
                       "patterns_in": "2-bit packed (16 B per 64 bases)" if packed_api else "ASCII", "seed_chars": int(info.seed_chars),
                       "seed_table_MB": int(info.seed_table_mb), "verification_records_MB": int(info.vrec_mb),
                       "index_device_MB": round(info.device_bytes / 1e6, 1), "index_load_s": round(load_s, 3)},
            "occurrences_per_s": (occ_all / (dev_ms_max / 1e3)) if locate else None,
            "e2e": {"value": e2e_value, "unit": "patterns/s", "h2d_bytes_per_step": h2d * world if world > 1 else h2d,
                    "d2h_bytes_per_step": d2h * world if world > 1 else d2h, "ms_per_step": e2e_ms_max / steps,
                    "device_timeline_ms_per_step": float(np.mean(e2e_dev_ms[-steps:])) if e2e_dev_ms else None,
                    "host_link": pcie,
                    "what": ((f"2-bit packed patterns without padding (E2FM_PACKED_TIGHT, {tstride} B per pattern)" if tight_api else
                              "2-bit packed patterns (e2fm_search_batch_packed, 16 B per 64 bases)" if packed_api else "ASCII patterns") +
                             " from page-locked host memory; " +
                             ("per-pattern occurrence offsets, sequence index and offset in the sequence of every occurrence (what "
                              "EFMCollection::search returns, FMCollection.h:23-44), 32 bits each" if locate else "32-bit counts") +
                             " back into page-locked host memory; " +
                             (f"one call, e2fm_search_hits, {e2e_state.get('stages', 0)} pipeline stages" if args.e2e_api == "hits"
                              else "two calls: search, then e2fm_result_fetch_compact")),
                    "api": "e2fm_search_hits" if args.e2e_api == "hits" else "e2fm_search_batch_* + e2fm_result_fetch_compact"},
            "gpu_launches": launches_all + e2e_launches_all,
            "gpu_launches_by_leg": {"device_resident": launches_all, "end_to_end": e2e_launches_all},
            "roofline": roofline,
            "parity": parity,
            "ops_per_pattern": {k_: ctr_sum.get(k_, 0) / (npat * steps) for k_ in ("rank", "lf", "mark", "rank_sectors", "walk_gathers",
                                                                                    "complex_patterns")},
            "phase_ms_per_step": {"encode_or_pack": ph_ms["encode"] / steps, "backward_or_seed_search": ph_ms["backward_search"] / steps,
                                  "scan_expand": ph_ms["expand"] / steps, "firstcheck": ph_ms["firstcheck"] / steps,
                                  "walk_resolve_or_seed_verify": ph_ms["walk"] / steps, "results": ph_ms["results"] / steps},
            "clocks": clocks, "wall_s_timed_device_loop": wall_dev,
        }
        if cpu_base is not None and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_base
        if world == 1 and not args.no_cpp:
            line["e2e_cpp"] = cpp_host_arm(prep, local_rank)
        if sharded_obj is not None:
            line["sharded"] = sharded_obj
        emit(line)
    if ix is not None:
        ix.close()
    shm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and parity is not None and parity.get("mismatches"):
        log(f"[bench] PARITY FAILURE: {parity}")
        sys.exit(1)


def cpp_host_arm(prep, device):
    """the same batch through the C++ host classes (e2fm_b200/host): std::vector<std::string> in, the 32-bit CSR or the reference's
    std::vector<SearchResult> out; host wall clock of the whole call, measured by e2fm_bench_host"""
    exe = os.path.join(ROOT, "e2fm_b200", "e2fm_bench_host")
    if not os.path.exists(exe):
        return {"error": "e2fm_b200/e2fm_bench_host not built"}
    w = prep["w"]
    with open(prep["patf"], "rb") as f:
        np.lib.format.read_magic(f)
        np.lib.format.read_array_header_1_0(f)
        skip = f.tell()
    r = subprocess.run([exe, prep["efm"], prep["keyfile"], prep["patf"], str(w["n_patterns"]), str(w["pat_len"]),
                        "locate" if w["locate"] else "count", "3", str(device), str(skip)], capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not line:
        return {"error": (r.stderr or r.stdout)[-300:]}
    return json.loads(line[-1])


def make_roofline(args, capi, info, local_rank, ctr_sum, ph_ms, steps, peak, peak_kind, npat=0):
    """Roofline of the dominant kernel. This layout's contract (it replaces SURVEY.md 8(d)'s table, which prices the
    REFERENCE's five-to-seven-sector operations): 32 B per 32-byte sector the kernel has to gather (seed-table sectors, verification
    records, rank sectors, sa / text / record gathers), 2 B x 2 per row streamed by the wildcard scan.

    The ceiling of such a kernel is not the copy bandwidth but the RANDOM-GATHER rate of the part, measured on this box for a working
    set of the index's size by e2fm_debug_gather_bench (47 G gathers/s from 1 GB up on a B200: every gather activates a DRAM row and,
    without a hint, moves a whole 128-byte line, i.e. 6.0 TB/s of DRAM traffic = 0.92 of the copy peak; profiles/exp/). So:
        achieved = 32 B x sectors / kernel time          peak = 32 B x measured gathers/s          frac = achieved / peak
    and next to it the same against the HBM copy peak of MEASURED_PEAKS.json (hbm_copy_peak), with the ncu DRAM traffic of the
    kernel (`traffic`, bytes per launch of `launch_patterns` patterns) when profiles/ holds a capture of this workload."""
    resident = info.device_bytes < L2_BYTES
    # random-gather rate of THIS box for a working set of the index's size (power of two >= the image, at least 16 MB)
    ws_buf = 16 << 20
    while ws_buf < info.device_bytes and ws_buf < (64 << 30):
        ws_buf *= 2
    try:
        gg = capi.gather_bench(ws_buf, 1 << 29, local_rank)
    except Exception as e:   # noqa: BLE001
        gg = None
        log(f"[bench] gather microbenchmark failed: {e}")
    seeded = ctr_sum.get("seeded", 0) > 0
    k3 = "seed_lookup_kernel" if seeded else "backward_search_kernel"
    k4 = ("seed_verify_rec_kernel" if info.vrec_mb else "seed_verify_kernel") if seeded else ("walk_kernel" if args.walk else "resolve_kernel")
    lay = {k3: 32.0 * ctr_sum.get("rank_sectors", 0), k4: 32.0 * ctr_sum.get("walk_gathers", 0),
           "firstcheck_kernel": 4.0 * ctr_sum.get("scan_rows", 0)}
    tms = {k3: ph_ms["backward_search"], k4: ph_ms["walk"], "firstcheck_kernel": ph_ms["firstcheck"]}
    gpeak = gg * 32.0 if gg else None
    kern = {}
    for name in lay:
        t = tms[name]
        ach = lay[name] / (t / 1e3) / 1e9 if t else 0.0
        kern[name] = {"ms_per_step": t / steps, "algorithmic_bytes_per_step": lay[name] / steps, "achieved_gbs": ach,
                      "frac_of_random_gather_rate": (ach / gpeak) if (gpeak and name != "firstcheck_kernel") else None,
                      "frac_of_hbm_copy_peak": ach / peak}
    dom = max(tms, key=lambda n: tms[n])
    # ncu DRAM traffic of the dominant kernel: one launch = one chunk of the device-resident batch
    chunk = min(npat, 4 << 20) if npat else 0
    traffic = None
    short = {"ecoli": "ecoli", "collection25": "c25", "collection100": "c100", "config1": "config1"}.get(args.workload, args.workload)
    try:
        import glob
        for tf in sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_*.json"))):
            with open(tf) as f:
                tj = json.load(f)
            for name, d in tj.items():
                if name.startswith(f"prof_{short}_") and dom in d:
                    traffic = d[dom]
    except Exception:
        traffic = None
    achieved = kern[dom]["achieved_gbs"]
    per_launch = (chunk / npat) if npat else 1.0
    t_launch_s = tms[dom] / steps / 1e3 * per_launch
    ref_bytes = 160.0 * ctr_sum.get("rank", 0) + 224.0 * ctr_sum.get("lf", 0) + 64.0 * ctr_sum.get("mark", 0)
    use_gather = gpeak is not None and dom != "firstcheck_kernel"
    roof = {"bound": "l2" if resident else "hbm", "kernel": dom, "achieved": achieved, "peak": gpeak if use_gather else peak, "unit": "GB/s",
            "frac": achieved / (gpeak if use_gather else peak), "traffic": traffic,
            "peak_source": ((f"measured in this run: {gg:.1f} G random 32-byte-sector gathers/s over a {ws_buf >> 20} MB buffer (e2fm_debug_gather_bench, "
                             f"2^29 independent 8-byte loads at hashed addresses, best of 3) x 32 B; the HBM copy peak is in hbm_copy_peak")
                            if use_gather else f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})"),
            "algorithmic_bytes_per_launch": kern[dom]["algorithmic_bytes_per_step"] * per_launch, "launch_patterns": chunk,
            "hbm_copy_peak": {"peak": peak, "source": "MEASURED_PEAKS.json hbm_gbs (" + peak_kind + ")", "frac_sectors": achieved / peak,
                              "dram_traffic_gbs": (traffic / t_launch_s / 1e9) if (traffic and t_launch_s) else None,
                              "frac_dram_traffic": (traffic / t_launch_s / 1e9 / peak) if (traffic and t_launch_s) else None,
                              "random_gather_rate_as_frac": (gpeak / peak) if gpeak else None},
            "random_gather": None if not gg else {"giga_gathers_per_s": gg, "buffer_MB": ws_buf >> 20},
            "note": ("the device image fits the 126 MB L2: the gather rate is that of an L2-resident working set of the same size" if resident else
                     "achieved = 32 B x sectors this layout must gather, per step, over the kernel's CUDA-event time; traffic = ncu dram read+write "
                     "of one launch of the same kernel (profiles/traffic_*.json, launch_patterns patterns)"),
            "reference_model": {"bytes_per_step": ref_bytes / steps,
                                "note": "SURVEY.md 8(d): 160 B/rank + 224 B/lf + 64 B/mark x the REFERENCE operations this batch stands for; "
                                        "not a measure of this layout (seed table, verification records, one sector per rank, no walks)"},
            "kernels": kern}
    return roof


def shard_plan(w, world):
    """contiguous sequence groups balanced by length; every shard's own super-text length must dodge the reference's load-time quirks
    (SURVEY.md 8 quirks 2, 3): lengthen the last sequence of an unlucky shard by k bases until all shards are clean (deterministic,
    so every rank agrees). Returns (lens, plan, adjusted?)."""
    from e2fm_b200 import sharded
    lens = workload_lens(w)
    lens0 = list(lens)
    for _ in range(1000):
        plan = sharded.plan_shards(lens, world, K)
        bad = None
        for (f0, c0, _b) in plan:
            sl = lens[f0:f0 + c0]
            if c0 and synth.safe_lengths(sl, K, 100 // MRP) != sl:
                bad = f0 + c0 - 1
                break
        if bad is None:
            break
        lens[bad] += K
    return lens, plan, lens != lens0


def sharded_measure(args, rank, world, local_rank, prep, replica=None):
    """Collection sharded by sequence over the ranks, through the C ABI (e2fm_shardset_*): rank r indexes sequence
    group r with the reference builder, supplies 1/N of the batch from the one host buffer, every rank searches the whole batch
    against its shard, and the answers for slice r travel to rank r over NCCL (reduce-scatter of the counts, one send/recv round for
    the occurrence lists), from where they go back to the host buffer. A step = host patterns -> host results, wall clock, max over
    ranks. `replica`: the host arrays the replica arm left for the same batch; the sharded answers must equal them."""
    import torch
    import torch.distributed as dist

    from e2fm_b200 import capi

    dev_t = torch.device("cuda", local_rank)
    w = prep["w"]
    locate = w["locate"]
    lens, plan, adjusted = shard_plan(w, world)
    first, cnt, text_base = plan[rank]
    keyfile = prep["keyfile"]
    efm = os.path.join(CACHE, f"{prep['workload']}_shard{rank}of{world}_k{K}_b{BUCKET}_m{MRP}.efm")
    t_build = time.time()
    if not os.path.exists(efm):
        seqs = []
        for i in range(first, first + cnt):
            seqs += synth.random_sequences([lens[i]], w["seed"] + i)
        build_index(seqs, efm, keyfile, max(2, min(32, (os.cpu_count() or 2) // world)))
        del seqs
    t_build = time.time() - t_build
    with open(efm, "rb") as f:
        raw = f.read()
    ix = capi.DeviceIndex(raw, synth.test_key(), local_rank)
    del raw
    ids = [capi.shardset_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(ids, src=0)
    ss = capi.ShardSet(ix, ids[0], rank, world, first, text_base)
    npat_all, L = w["n_patterns"], w["pat_len"]
    lo, npat = capi.shardset_slice(npat_all, world, rank)
    per = -(-npat_all // world)
    all_pats = load_patterns(prep)
    packed_api = not args.ascii
    if packed_api:
        try:
            probe, _ = capi.pack_patterns(np.ascontiguousarray(all_pats[:1]))
            ix.search_packed(probe, 1, L, locate).free()
        except capi.E2FMError:
            packed_api = False
    unit = capi.packed_stride(L) if packed_api else L
    shm = SharedHost(f"{os.environ.get('MASTER_PORT', '0')}_{prep['workload']}_sh", world, rank, capi)
    h_pats = shm.array("patterns", (npat_all, unit), np.uint8)
    occ_cap = int((per + 1) * max(2.0, 1.3 * args.occ_per_pattern)) + 1024
    h_counts = shm.array("counts", (world, per + 1), np.uint32)
    h_noff = shm.array("occ_offsets", (world, per + 2), np.uint32) if locate else None
    h_seq = shm.array("seq", (world, occ_cap), np.uint32) if locate else None
    h_off = shm.array("off", (world, occ_cap), np.uint32) if locate else None
    if world > 1:
        dist.barrier()
    mine = np.ascontiguousarray(all_pats[lo:lo + npat])
    if packed_api:
        _, bad = capi.pack_patterns(mine, h_pats[lo:lo + npat].reshape(-1))
        assert bad == 0
    else:
        h_pats[lo:lo + npat] = mine
    del mine
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev_t)
    state = {}

    def step():
        flush.add_(1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        r = ss.search(h_pats[lo:lo + npat].reshape(-1), npat_all, L, packed_api, locate)
        total = r.total
        if total > occ_cap:
            raise SystemExit("bench: occurrence buffer too small; pass a larger --occ-per-pattern")
        if locate:
            r.fetch_compact(None, h_noff[rank][:npat + 1], h_seq[rank][:total], h_off[rank][:total])
        else:
            r.fetch_compact(h_counts[rank][:npat], None, None, None)
        r.free()
        dt = time.perf_counter() - t0
        state.update(total=total, ms=ss.last_ms())
        return dt

    for _ in range(max(args.warmup, 3)):
        step()
    tot_s = 0.0
    phases = []
    for _ in range(args.steps):
        tot_s += step()
        phases.append(state["ms"])
    # the same answers as the replicas gave for this rank's slice of the same batch?
    mism = 0
    if replica is not None and replica.get("lo") == lo and replica.get("npat") == npat:
        if locate:
            a, b = replica["noff"][:npat + 1], h_noff[rank][:npat + 1]
            t = int(b[npat])
            same = np.array_equal(a, b) and np.array_equal(replica["seq"][:t], h_seq[rank][:t]) and np.array_equal(replica["off"][:t], h_off[rank][:t])
        else:
            same = np.array_equal(replica["counts"], h_counts[rank][:npat])
        mism = 0 if same else 1
    t = torch.tensor([tot_s, float(mism), float(state["total"]), t_build], dtype=torch.float64, device=dev_t)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms = float(tmax[0]) * 1e3 / args.steps
    out = {
        "value": npat_all / (ms / 1e3), "unit": "patterns/s", "ms_per_step": ms, "scaling": "strong",
        "parallelism": f"collection sharded by sequence over {world} GPUs (e2fm_shardset_*): each rank supplies 1/{world} of the batch from the host, "
                       "ncclAllGather of the batch, every rank searches all patterns on its shard, counts by ncclReduceScatter, occurrence lists "
                       "by one grouped ncclSend/ncclRecv round, merged answers of slice r back to the host from rank r",
        "timing": "wall clock, host patterns -> host results, barrier in front, max over ranks",
        "shards": [list(x) for x in plan], "lengths_adjusted_for_shard_quirks": adjusted,
        "phase_ms_rank0": {k_: float(np.mean([p[k_] for p in phases])) for k_ in (phases[0] if phases else {})},
        "occurrences": int(t[2]), "shard_index_build_s_max": float(tmax[3]),
        "index_device_MB_rank0": round(ix.info.device_bytes / 1e6, 1),
        "same_answers_as_replicas": (None if replica is None else {"ranks_that_differ": int(t[1])}),
        "patterns_in": "2-bit packed" if packed_api else "ASCII",
    }
    ss.close()
    ix.close()
    shm.close()
    return out


def sharded_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    prep = prepare(args.workload, rank)
    if world > 1:
        dist.barrier()
    sh = sharded_measure(args, rank, world, local_rank, prep)
    if rank == 0:
        w = prep["w"]
        line = {"metric": "patterns/sec", "value": sh["value"], "unit": "patterns/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": sh["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
                "data": "synthetic",
                "config": {"workload": w["name"], "mode": "count+locate" if w["locate"] else "count", "k": K, "bucket_size": BUCKET, "mrp": MRP,
                           "patterns": w["n_patterns"], "pattern_len": w["pat_len"], "parallelism": sh["parallelism"],
                           "l2": "256 MB flush between timed iterations"},
                "e2e": {"value": sh["value"], "unit": "patterns/s", "ms_per_step": sh["ms_per_step"]},
                "sharded": sh, "gpu_launches": None}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="e2fm_b200", choices=["e2fm_b200", "reference"])
    ap.add_argument("--workload", default="human", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpp", action="store_true", help="skip the e2e_cpp figure (C++ host classes)")
    ap.add_argument("--occ-per-pattern", type=float, default=1.0, help="expected occurrences per pattern (sizes the host result buffers)")
    ap.add_argument("--pieces", type=int, default=0, help="sharded mode: pipeline pieces per batch (0 = automatic)")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the sharded-by-sequence leg of the default line")
    ap.add_argument("--sharded", action="store_true", help="shard the collection by sequence over the GPUs and merge over NCCL")
    ap.add_argument("--chunk", type=int, default=0, help="patterns per internal chunk (0 = library default)")
    ap.add_argument("--taper", action="store_true", help="halving pipeline stages for host batches")
    ap.add_argument("--pipe", type=int, default=0, help="patterns per pipeline stage of host batches (0 = library default)")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the CPUs of its GPU's NUMA node")
    ap.add_argument("--no-tight", action="store_true", help="end-to-end call on the 16-byte-word packing instead of the tight one")
    ap.add_argument("--ascii", action="store_true", help="ASCII patterns through e2fm_search_batch_fixed even when the packed entry point would take the batch")
    ap.add_argument("--e2e-api", default="hits", choices=["hits", "fetch"], help="end-to-end leg: the pipelined one-call form or search + fetch")
    ap.add_argument("--walk", action="store_true", help="locate / verify by per-query walks from the marked rows instead of the dense tables")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)
    if args.workload == "chr1" and args.occ_per_pattern < 20:
        args.occ_per_pattern = 20.0
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        reference_arm(args, prepare(args.workload, 0))
        return
    claim_stdout()
    if args.sharded:
        sharded_arm(args, rank, world, local_rank)
    else:
        gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
