"""Turn the ncu captures a gpurun call brought back (gpurun_out/*.ncu-rep, launches_*.csv) into the small text
summaries committed under profiles/. Usage: python profiles/summarize.py <round-tag>   (run in the authoring container)."""
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % (integer: add/xor/shf/lop3)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % (IMAD)"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors (L1)"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors from SMs"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (warps/issue)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
]


def raw_rows(rep):
    if rep.endswith(".csv"):
        with open(rep) as f:
            out = f.read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    traffic = {}
    reps = {}
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"*_{tag}.ncu-rep"))):
        reps[os.path.basename(rep)[:-8]] = rep
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"*_{tag}.raw.csv"))):     # exported on the GPU box (capture_r2.sh)
        reps[os.path.basename(rep)[:-8]] = rep
    for name, rep in sorted(reps.items()):
        hdr, units, rows = raw_rows(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        with open(os.path.join(OUT, name + ".txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on  ({os.path.basename(rep)}; per-launch, cold caches, serialised)\n")
            for r in rows:
                kn = r[idx["Kernel Name"]]
                f.write(f"\n== {kn}\n")
                for key, label in KEYS:
                    if key in idx:
                        f.write(f"  {label:48s} {r[idx[key]]:>18s} {units[idx[key]]}\n")
                try:
                    rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
                    wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
                    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                    tb = rd * mult[units[idx["dram__bytes_read.sum"]]] + wr * mult[units[idx["dram__bytes_write.sum"]]]
                    short = kn.split("(")[0].split("::")[-1].replace("void ", "").split("<")[0]
                    traffic.setdefault(name, {})[short] = tb
                    f.write(f"  {'traffic = dram read + write (bytes)':48s} {tb:18.0f}\n")
                except Exception:
                    pass
        print("wrote", name + ".txt")
    for lc in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"launches_*_{tag}.csv"))):
        rows = [r for r in csv.reader(open(lc)) if len(r) > 10]
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        name = os.path.basename(lc)[:-4] + ".txt"
        tot = {}
        order = []
        for r in rows[1:]:
            k = r[ki].split("(")[0]
            v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
            if k not in tot:
                tot[k] = [0, 0.0]
                order.append(k)
            tot[k][0] += 1
            tot[k][1] += v
        allus = sum(v[1] for v in tot.values())
        with open(os.path.join(OUT, name), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none : every launch of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline`\n")
            f.write("# (cold-cache, serialised launches: compare SHARES, not absolutes)\n")
            f.write(f"{'kernel':80s} {'launches':>8s} {'total us':>12s} {'share':>7s}\n")
            for k in order:
                f.write(f"{k[:80]:80s} {tot[k][0]:8d} {tot[k][1]:12.1f} {100 * tot[k][1] / allus:6.1f}%\n")
        print("wrote", name)
    if not traffic:
        return
    with open(os.path.join(OUT, f"traffic_{tag}.json"), "w") as f:
        json.dump(traffic, f, indent=1)


if __name__ == "__main__":
    main()
