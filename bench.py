#!/usr/bin/env python
"""bench.py — batched pattern search throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ecoli|collection100|config1] [--impl reference]

A step = one pass of the hot path over one batch of synthetic patterns (count, or count+locate).
  value      whole-job patterns/s with the pattern batch already resident in HBM (device time, CUDA events on
             the library's stream, max over ranks)
  e2e        the same through the C ABI with HOST (pinned) buffers: H2D of the patterns and D2H of the counts
             (and positions) inside the timed region
  roofline   dominant kernel: algorithmic bytes (160 B/rank + 224 B/lf + 64 B/mark, operation
             counts measured by the kernels and pinned to the oracle's by tests) / its CUDA-event time
  cpu_baseline  the UNMODIFIED reference (oracle/_ref/e2fm_ref) on the host cores over a bounded sample
--impl reference times that CPU reference as the main line.
The index is always built by the reference's own CPU builder (north star: index construction stays on the CPU).
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

WORKLOADS = {
    # BASELINE.json configs[1]: E. coli-sized 4.6 Mbp synthetic genome, 1M patterns len 32, count only, 1 GPU
    "ecoli": dict(seq_lens=[4_641_652], seed=43, n_patterns=1_000_000, pat_len=32, locate=False,
                  name="configs[1]: 4.6 Mbp synthetic genome, 1M patterns len 32, count"),
    # BASELINE.json configs[0]: 10 Mbp random ACGT, 10k patterns len 20, count+locate
    "config1": dict(seq_lens=[10_000_000], seed=42, n_patterns=10_000, pat_len=20, locate=True,
                    name="configs[0]: 10 Mbp random ACGT, 10k patterns len 20, count+locate"),
    # BASELINE.json configs[2] on ONE GPU (whole collection in one index): larger than L2, HBM-bound
    "collection100": dict(seq_lens=[4_641_652] * 100, seed=1000, n_patterns=10_000_000, pat_len=24, locate=True,
                          name="configs[2]: 100 x 4.6 Mbp synthetic genomes, 10M patterns len 24, count+locate"),
    # BASELINE.json configs[4]: chr1-sized 248 Mbp single sequence, locate-heavy short patterns (len 12)
    "chr1": dict(seq_lens=[248_956_422], seed=45, n_patterns=1_000_000, pat_len=12, locate=True,
                 name="configs[4]: 248 Mbp chr1-sized synthetic sequence, 1M patterns len 12, count+locate"),
    # BASELINE.json configs[3]: human-genome-sized 3.1 Gbp collection (24 sequences, hg38-like proportions), 10M patterns len 50
    "human": dict(seq_lens=[248_956_422, 242_193_529, 198_295_559, 190_214_555, 181_538_259, 170_805_979, 159_345_973, 145_138_636,
                            138_394_717, 133_797_422, 135_086_622, 133_275_309, 114_364_328, 107_043_718, 101_991_189, 90_338_345,
                            83_257_441, 80_373_285, 58_617_616, 64_444_167, 46_709_983, 50_818_468, 156_040_895, 57_227_415],
                  seed=44, n_patterns=10_000_000, pat_len=50, locate=True,
                  name="configs[3]: 3.1 Gbp human-genome-sized synthetic collection (24 sequences), 10M patterns len 50, count+locate"),
    # a collection100 shard-sized index for quicker HBM-bound runs
    "collection25": dict(seq_lens=[4_641_652] * 25, seed=1000, n_patterns=4_000_000, pat_len=24, locate=True,
                         name="25 x 4.6 Mbp synthetic genomes, 4M patterns len 24, count+locate"),
}
K, BUCKET, MRP = 4, 4096, 2   # the authors' regime (Main.cpp:469,747)


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


def prepare(workload: str, rank: int = 0):
    """Synthetic collection + reference-built index (cached on disk) -> dict"""
    w = WORKLOADS[workload]
    os.makedirs(CACHE, exist_ok=True)
    lens = synth.safe_lengths(w["seq_lens"], K, 100 // MRP)
    seqs = []
    for i, L in enumerate(lens):
        seqs += synth.random_sequences([L], w["seed"] + i)
    base = os.path.join(CACHE, f"{workload}_k{K}_b{BUCKET}_m{MRP}")
    keyfile = os.path.join(CACHE, "key.bin")
    efm = base + ".efm"
    if rank == 0:
        if not os.path.exists(keyfile):
            with open(keyfile, "wb") as f:
                f.write(synth.test_key())
        if not os.path.exists(efm):
            t0 = time.time()
            fa = base + ".fa"
            synth.write_fasta(fa, seqs)
            threads = str(min(32, os.cpu_count() or 2))
            r = subprocess.run([ref_bin(), "build", fa, efm + ".tmp", keyfile, str(K), str(BUCKET), str(MRP), threads],
                               capture_output=True, text=True)
            if r.returncode != 0 or not os.path.exists(efm + ".tmp"):
                raise SystemExit("reference index build failed: " + r.stderr[-400:])
            os.replace(efm + ".tmp", efm)
            os.remove(fa)
            log(f"[bench] reference built {efm} ({os.path.getsize(efm) / 1e6:.1f} MB) in {time.time() - t0:.1f}s")
    return dict(w=w, seqs=seqs, efm=efm, keyfile=keyfile)


def wait_for(path, timeout=3600):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise SystemExit("timed out waiting for " + path)
        time.sleep(0.5)


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


def run_reference_sample(prep, pats: np.ndarray, locate: bool, procs: int, tmpdir: str):
    """P independent processes of the unmodified reference, each over a 1/P slice (the reference is serial and not
    re-entrant per object); 1 matcher thread each; 2 passes: cold (lazy bucket decode) and warm.
    Returns (patterns/s warm aggregate, patterns/s cold aggregate, wall seconds)."""
    slices = np.array_split(np.arange(len(pats)), procs)
    ps = []
    t0 = time.time()
    for i, sl in enumerate(slices):
        pf = os.path.join(tmpdir, f"p{i}.txt")
        synth.write_patterns(pf, pats[sl])
        ps.append(subprocess.Popen([ref_bin(), "search", prep["efm"], prep["keyfile"], pf, os.path.join(tmpdir, f"r{i}.res"),
                                    "1" if locate else "0", "1", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    cold = warm = 0.0
    for p, sl in zip(ps, slices):
        out, err = p.communicate()
        line = [l for l in out.splitlines() if l.startswith("{")]
        if p.returncode != 0 or not line:
            for q in ps:
                if q.poll() is None:
                    q.kill()
            raise RuntimeError(f"reference search failed (exit code {p.returncode}): {(err or out)[-300:].strip()}")
        j = json.loads(line[-1])
        cold = max(cold, j["pass_ms"][0])
        warm = max(warm, j["pass_ms"][1])
    return len(pats) / (warm / 1e3), len(pats) / (cold / 1e3), time.time() - t0


def reference_arm(args, prep):
    w = prep["w"]
    cores = os.cpu_count() or 1
    procs = reference_procs(prep)
    big = os.path.getsize(prep["efm"]) > (200 << 20)
    per_proc = (500 if big else 4000) if w["pat_len"] >= 20 else 400
    sample = min(w["n_patterns"], procs * per_proc)
    pats = synth.sample_patterns(prep["seqs"], sample, w["pat_len"], w["seed"] + 1)[:sample]
    vals, colds = [], []
    with tempfile.TemporaryDirectory() as td:
        t_all = time.time()
        for step in range(args.warmup + args.steps):
            v, c, wall = run_reference_sample(prep, pats, w["locate"], procs, td)
            if step >= args.warmup:
                vals.append(v); colds.append(c)
        wall_all = time.time() - t_all
    value = float(np.mean(vals))
    desc = (f"{sample} of {w['n_patterns']} patterns (len {w['pat_len']}) split over {procs} processes of the unmodified "
            f"reference, 1 matcher thread each; value = warm pass (2nd pass, buckets already decoded), cold pass = "
            f"{float(np.mean(colds)):.0f} patterns/s")
    line = {
        "impl": "reference", "metric": "patterns/sec", "value": value, "unit": "patterns/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": w["name"], "mode": "count+locate" if w["locate"] else "count", "k": K, "bucket_size": BUCKET,
                   "mrp": MRP, "sample_patterns": sample},
        "cpu_baseline": {"value": value, "unit": "patterns/s", "cores": procs, "kind": "reference", "sample": desc,
                         "cold_value": float(np.mean(colds)), "host_cores": cores, "cpu_model": cpu_model(),
                         "per_process": value / procs},
        "e2e": {"value": value, "unit": "patterns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall_all,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm

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
    wait_for(prep["efm"])
    w = prep["w"]
    locate = w["locate"]
    with open(prep["efm"], "rb") as f:
        raw = f.read()
    t0 = time.time()
    ix = capi.DeviceIndex(raw, synth.test_key(), local_rank)
    load_s = time.time() - t0
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

    # weak scaling: every rank searches its own batch of the workload's size against its replica of the index
    # (single-sequence workloads do not shard: replicas only)
    npat, L = w["n_patterns"], w["pat_len"]
    pats = synth.sample_patterns(prep["seqs"], npat, L, w["seed"] + 1 + 7919 * rank)[:npat]
    pin_p = capi.PinnedArray((npat, L), np.uint8)
    pin_p.array[:] = pats
    pin_counts = capi.PinnedArray(npat, np.uint64)
    d_pats = torch.from_numpy(pats).to(dev_t)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev_t)   # > 126 MB L2
    torch.cuda.synchronize()

    def l2_flush():
        if os.environ.get("E2FM_BENCH_NOFLUSH"):
            return
        flush.add_(1)
        torch.cuda.synchronize()

    def step_device():
        r = ix.search_device_fixed(d_pats.data_ptr(), npat, L, locate)
        ms = ix.last_batch_ms()
        ctr = ix.last_batch_counters()
        tot = r.total
        r.free()
        return ms, ctr, tot

    pin_pos = {}
    e2e_dev_ms = []

    def step_e2e():
        t0 = time.perf_counter()
        r = ix.search_stream(pin_p.array, locate, pin_counts.array)
        if locate:
            need = max(r.total, 1)
            if pin_pos.get("n", 0) < need:
                for k_ in ("pos", "off"):
                    if k_ in pin_pos:
                        pin_pos[k_].free()
                pin_pos["pos"] = capi.PinnedArray(int(need * 1.2) + 16, np.uint64)
                pin_pos["off"] = capi.PinnedArray(npat + 1, np.uint64)
                pin_pos["n"] = int(need * 1.2) + 16
            r.fetch_positions(pos_offsets=pin_pos["off"].array, positions=pin_pos["pos"].array[:r.total])
        else:
            pass   # count mode: the counts were streamed into pin_counts chunk by chunk (e2fm_search_batch_stream)
        dt = time.perf_counter() - t0
        tot = r.total
        r.free()
        e2e_dev_ms.append(ix.last_batch_ms()["total"])
        return dt, tot

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
    dev_ms, bs_ms, walk_ms, exp_ms, res_ms, enc_ms = 0.0, 0.0, 0.0, 0.0, 0.0, 0.0
    ctr_sum = {"rank": 0, "lf": 0, "mark": 0, "launches": 0, "tasks": 0, "rank_sectors": 0, "walk_gathers": 0, "scan_rows": 0,
               "chunks_redone": 0}
    fc_ms = 0.0
    occ_total = 0
    for _ in range(args.steps):
        l2_flush()
        ms, ctr, tot = step_device()
        dev_ms += ms["total"]; bs_ms += ms["backward_search"]; walk_ms += ms["walk"]; exp_ms += ms["expand"]; res_ms += ms["results"]; enc_ms += ms["encode"]; fc_ms += ms["firstcheck"]
        for k_ in ctr_sum:
            ctr_sum[k_] += ctr[k_]
        occ_total += tot
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall_dev = time.perf_counter() - wall0
    e2e_s = 0.0
    for _ in range(args.steps):
        l2_flush()
        dt, _tot = step_e2e()
        e2e_s += dt
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()

    # max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1e3, bs_ms, walk_ms], dtype=torch.float64, device=dev_t)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, bs_ms_max, walk_ms_max = [float(x) for x in t.tolist()]

    if rank == 0:
        peak, peak_kind = measured_peak()
        steps = args.steps
        # the random-gather roofline of THIS box for a working set of this index's size: independent 8-byte gathers over a
        # buffer as large as the device image (power of two, at least 256 MB), one 32-byte sector each
        gbuf = 256 << 20
        while gbuf < info.device_bytes and gbuf < (64 << 30):
            gbuf *= 2
        try:
            gg = capi.gather_bench(gbuf, 1 << 29, local_rank)
        except Exception as e:   # noqa: BLE001
            gg = None
            log(f"[bench] gather microbenchmark failed: {e}")
        # ... and, when the image is smaller than that floor (L2-resident indexes), the same measurement over a buffer of the
        # image's own size: what a dependent-gather kernel could reach out of the L2
        gg_ws, ws_buf = None, 16 << 20
        try:
            while ws_buf < info.device_bytes:
                ws_buf *= 2
            if ws_buf < gbuf:
                gg_ws = capi.gather_bench(ws_buf, 1 << 29, local_rank)
        except Exception as e:   # noqa: BLE001
            gg_ws = None
            log(f"[bench] working-set gather microbenchmark failed: {e}")
        value = world * npat * steps / (dev_ms_max / 1e3)
        e2e_value = world * npat * steps / (e2e_ms_max / 1e3)
        # algorithmic bytes, SURVEY.md 8(d): rank 160 B, lf 224 B, mark 64 B per reference operation
        # SURVEY.md 8(d) reference-model bytes: only meaningful where the reference's operations are executed one for one
        bytes_bs = 160.0 * ctr_sum["rank"]
        bytes_walk = 224.0 * ctr_sum["lf"] + 64.0 * ctr_sum["mark"]
        # what THIS layout has to touch at minimum: 32 B per distinct sector gathered by the backward search,
        # 32 B per gather of the walk / resolve kernel, 2 B per row streamed (twice) by the wildcard scan
        lay = {"backward_search_kernel": 32.0 * ctr_sum["rank_sectors"],
               "walk_kernel": 32.0 * ctr_sum["walk_gathers"],
               "firstcheck_kernel": 4.0 * ctr_sum["scan_rows"]}
        tms = {"backward_search_kernel": bs_ms, "walk_kernel": walk_ms, "firstcheck_kernel": fc_ms}
        ref = {"backward_search_kernel": bytes_bs, "walk_kernel": bytes_walk, "firstcheck_kernel": 0.0}
        kern = {}
        for name in lay:
            t = tms[name]
            kern[name] = {"ms_per_step": t / steps, "alg_bytes_per_step": ref[name] / steps,
                          "achieved_gbs": ref[name] / (t / 1e3) / 1e9 if t else 0.0,
                          "layout_sector_bytes_per_step": lay[name] / steps,
                          "layout_sector_gbs": lay[name] / (t / 1e3) / 1e9 if t else 0.0}
        dom = max(tms, key=lambda n: tms[n])
        traffic = None
        short = {"ecoli": "ecoli", "collection25": "c25", "collection100": "c100", "config1": "config1"}.get(args.workload, args.workload)
        try:
            import glob
            tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_*.json")))
            if tfiles:
                with open(tfiles[-1]) as f:
                    tj = json.load(f)
                want = ("walk_kernel" if args.walk else "resolve_kernel") if dom == "walk_kernel" else dom
                for name, d in tj.items():
                    if name.startswith(f"prof_{short}_") and want in d:
                        traffic = d[want]
        except Exception:
            traffic = None
        # roofline.achieved = 32 B x the sectors THIS layout has to gather / the kernel's CUDA-event time;
        # reference_model = SURVEY.md 8(d)'s per-operation bytes of the REFERENCE's structure x the operations counted
        roofline = {"bound": "hbm", "kernel": dom + (" (resolve_kernel: dense tables)" if dom == "walk_kernel" and not args.walk else ""),
                    "achieved": kern[dom]["layout_sector_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kern[dom]["layout_sector_gbs"] / peak, "traffic": traffic, "peak_source": peak_kind,
                    "algorithmic_bytes_per_launch": kern[dom]["layout_sector_bytes_per_step"],
                    "reference_model": {"bytes_per_launch": kern[dom]["alg_bytes_per_step"], "gbs": kern[dom]["achieved_gbs"],
                                        "frac": kern[dom]["achieved_gbs"] / peak,
                                        "note": "160 B/rank + 224 B/lf + 64 B/mark (SURVEY.md 8(d)) x operations the kernels executed; "
                                                "above the peak by construction: one sector per rank instead of five, and with the dense "
                                                "tables the walks are not executed at all"},
                    "random_gather": None if not gg else {
                        "giga_gathers_per_s": gg, "sector_gbs": gg * 32.0, "frac_of_hbm_peak": gg * 32.0 / peak,
                        "achieved_frac_of_random_gather": kern[dom]["layout_sector_gbs"] / (gg * 32.0),
                        "buffer_MB": gbuf >> 20,
                        "working_set": None if not gg_ws else {
                            "buffer_MB": ws_buf >> 20, "giga_gathers_per_s": gg_ws,
                            "achieved_frac": kern[dom]["layout_sector_gbs"] / (gg_ws * 32.0),
                            "note": "the device image fits the L2: same microbenchmark over a buffer of the image's size"},
                        "how": "independent 8-byte loads at hashed addresses of a buffer the size of the device image (power of two, "
                               ">= 256 MB), 2^29 gathers, best of 3 (e2fm_debug_gather_bench); one 32-byte sector per gather"},
                    "note": "achieved = 32 B x sectors this layout must gather (distinct rank sectors / row-task gathers), per launch, "
                            "over the kernel's average CUDA-event duration; traffic = ncu dram read+write of one launch (profiles/)",
                    "kernels": kern}
        line = {
            "metric": "patterns/sec", "value": value, "unit": "patterns/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": w["name"], "mode": "count+locate" if locate else "count", "k": K, "bucket_size": BUCKET, "mrp": MRP,
                       "patterns_per_gpu": npat, "pattern_len": L, "parallelism": f"replicas x{world} (pattern batch per GPU)",
                       "l2": "256 MB flush between timed iterations",
                       "locate_tables": "per-query walks from marked rows" if args.walk else "dense sa[]/text[] built at load",
                       "index_device_MB": round(info.device_bytes / 1e6, 1), "index_load_s": round(load_s, 3)},
            "occurrences_per_s": (world * occ_total / (dev_ms_max / 1e3)) if locate else None,
            "e2e": {"value": e2e_value, "unit": "patterns/s", "h2d_bytes_per_step": int(npat * L),
                    "d2h_bytes_per_step": int(npat * 8 + ((npat + 1) * 8 + (occ_total // max(steps, 1)) * 8 if locate else 0)),
                    "ms_per_step": e2e_ms_max / steps,
                    "device_timeline_ms_per_step": float(np.mean(e2e_dev_ms[-steps:])) if e2e_dev_ms else None},
            "gpu_launches": int(ctr_sum["launches"]),
            "roofline": roofline,
            "ops_per_pattern": {"rank": ctr_sum["rank"] / (npat * steps), "lf": ctr_sum["lf"] / (npat * steps),
                                "mark": ctr_sum["mark"] / (npat * steps)},
            "phase_ms_per_step": {"encode": enc_ms / steps, "backward_search": bs_ms / steps, "scan_expand": exp_ms / steps,
                                  "firstcheck": fc_ms / steps, "walk_or_resolve": walk_ms / steps, "results": res_ms / steps},
            "clocks": clocks, "wall_s_timed_device_loop": wall_dev,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            procs = reference_procs(prep)
            big = os.path.getsize(prep["efm"]) > (200 << 20)
            per_proc = (500 if big else 4000) if L >= 20 else 400
            sample = min(npat, procs * per_proc)
            try:
                with tempfile.TemporaryDirectory() as td:
                    v, c, wall = run_reference_sample(prep, pats[:sample], locate, procs, td)
            except RuntimeError as e:
                v = c = None
                wall = 0.0
                line["cpu_baseline_error"] = str(e)
            line["cpu_baseline"] = {"value": v, "unit": "patterns/s", "cores": procs, "kind": "reference", "host_cores": cores,
                                    "cold_value": c, "cpu_model": cpu_model(), "per_process": (v / procs) if v else None,
                                    "sample": f"first {sample} patterns of the same batch over {procs} processes of the unmodified "
                                              f"reference (1 matcher thread each), warm pass; cold pass {c} patterns/s; {wall:.1f}s wall"}
        emit(line)
    ix.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def sharded_arm(args, rank, world, local_rank):
    """Collection sharded by sequence over the ranks: rank r indexes sequence group r with the reference
    builder, every rank searches the whole (broadcast) batch against its shard, counts and positions are merged over NCCL.
    Total work is fixed as N grows -> strong scaling."""
    import torch
    import torch.distributed as dist

    from e2fm_b200 import capi, sharded

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local_rank))
    dev_t = torch.device("cuda", local_rank)
    w = WORKLOADS[args.workload]
    locate = w["locate"]
    # every shard's own super-text length must dodge the reference's load-time quirks: lengthen the
    # last sequence of an unlucky shard by k bases until all shards are clean (deterministic, so every rank agrees)
    lens = synth.safe_lengths(w["seq_lens"], K, 100 // MRP)
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
    first, cnt, text_base = plan[rank]
    os.makedirs(CACHE, exist_ok=True)
    keyfile = os.path.join(CACHE, f"key_r{rank}.bin")
    with open(keyfile, "wb") as f:
        f.write(synth.test_key())
    seqs = []
    for i in range(first, first + cnt):
        seqs += synth.random_sequences([lens[i]], w["seed"] + i)
    efm = os.path.join(CACHE, f"{args.workload}_shard{rank}of{world}_k{K}_b{BUCKET}_m{MRP}.efm")
    if not os.path.exists(efm):
        t0 = time.time()
        fa = efm[:-4] + ".fa"
        # every shard must dodge the reference's load-time quirks on its own
        synth.write_fasta(fa, seqs)
        threads = str(max(4, min(32, (os.cpu_count() or 2) // world)))
        r = subprocess.run([ref_bin(), "build", fa, efm + ".tmp", keyfile, str(K), str(BUCKET), str(MRP), threads], capture_output=True, text=True)
        if r.returncode != 0:
            raise SystemExit("reference shard build failed: " + r.stderr[-400:])
        os.replace(efm + ".tmp", efm)
        os.remove(fa)
        log(f"[bench r{rank}] reference built shard {efm} in {time.time() - t0:.1f}s")
    with open(efm, "rb") as f:
        raw = f.read()
    col = sharded.ShardedCollection(raw, synth.test_key(), local_rank, first, text_base)
    if args.walk:
        col.ix.set_option(capi.OPT_DENSE, 0)
    npat, L = w["n_patterns"], w["pat_len"]
    pats0 = None
    if rank == 0:
        # patterns are sampled from the WHOLE collection: rank 0 regenerates the sequences it does not index
        all_seqs = []
        for i in range(len(lens)):
            all_seqs += synth.random_sequences([lens[i]], w["seed"] + i)
        pats0 = torch.from_numpy(synth.sample_patterns(all_seqs, npat, L, w["seed"] + 1)[:npat]).pin_memory()
        del all_seqs
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev_t)
    dist.barrier()

    def step():
        flush.add_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        m = col.search_batch(pats0, (npat, L), locate, args.pieces or None)
        e1.record()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        phases.append(dict(col.last_phase_ms))
        return dt, int(m.positions.numel()), int(m.counts.sum().item())

    phases = []
    for _ in range(args.warmup):
        step()
    phases.clear()
    sampler = ClockSampler(local_rank)
    dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    tot_s, occ, cnt_sum = 0.0, 0, 0
    for _ in range(args.steps):
        dt, o, c = step()
        tot_s += dt
        occ, cnt_sum = o, c
    dist.barrier()
    clocks = sampler.stop()
    t = torch.tensor([tot_s], dtype=torch.float64, device=dev_t)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item()) * 1e3 / args.steps
        line = {
            "metric": "patterns/sec", "value": npat / (ms / 1e3), "unit": "patterns/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": w["name"], "mode": "count+locate" if locate else "count", "k": K, "bucket_size": BUCKET, "mrp": MRP,
                       "patterns": npat, "pattern_len": L, "parallelism": f"collection sharded by sequence over {world} GPUs; "
                       "patterns broadcast, counts all-reduced, positions all-gathered (NCCL)", "shards": plan,
                       "l2": "256 MB flush between timed iterations", "timing": "wall clock around broadcast + search + merge, device synchronised, max over ranks"},
            "occurrences_per_s": occ / (ms / 1e3) if locate else None, "total_count": cnt_sum,
            "phase_ms_rank0": {k_: float(np.mean([p[k_] for p in phases])) for k_ in (phases[0] if phases else {})},
            "index_device_MB_rank0": round(col.ix.info.device_bytes / 1e6, 1),
            "e2e": {"value": npat / (ms / 1e3), "unit": "patterns/s", "h2d_bytes_per_step": npat * L, "d2h_bytes_per_step": 0,
                    "note": "patterns start in pinned host memory on rank 0; the merged result stays on the GPUs"},
            "gpu_launches": None, "clocks": clocks,
        }
        emit(line)
    col.close()
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="e2fm_b200", choices=["e2fm_b200", "reference"])
    ap.add_argument("--workload", default="ecoli", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pieces", type=int, default=0, help="sharded mode: pipeline pieces per batch (0 = automatic)")
    ap.add_argument("--sharded", action="store_true", help="shard the collection by sequence over the GPUs and merge over NCCL")
    ap.add_argument("--chunk", type=int, default=0, help="patterns per internal chunk (0 = library default)")
    ap.add_argument("--taper", action="store_true", help="halving pipeline stages for host batches")
    ap.add_argument("--pipe", type=int, default=0, help="patterns per pipeline stage of host batches (0 = library default)")
    ap.add_argument("--walk", action="store_true", help="locate / verify by per-query walks from the marked rows instead of the dense tables")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)
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
