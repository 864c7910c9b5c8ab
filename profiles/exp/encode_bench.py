"""Bucket text encoder (e2fm_encode_buckets) on a bench workload's index: kernel time against the oracle's C restatement of
Bucket::encodeText on one host core, with every bucket text compared with the bytes the reference builder wrote.
Usage: python profiles/exp/encode_bench.py [workload]   (default mid50: 480 Mbp, k = 4, bucket 4096 -> 29 k buckets)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from e2fm_b200 import capi, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "mid50"
prep = bench.prepare(wl, 0)
key = synth.test_key()
raw = open(prep["efm"], "rb").read()
t0 = time.time()
ix = O.OracleIndex(raw, key)
B, bs = int(ix.info.n_buckets), int(ix.info.bucket_size)
per = int(ix.info.superbucket_size) // bs
infos = [ix.bucket_info(b) for b in range(B)]
codes = np.concatenate([ix.bucket_codes(b) for b in range(B)])
alpha = np.array([i["alpha"] for i in infos], dtype=np.uint32)
print(f"[encode] {wl}: {len(codes)} rows, {B} buckets of {bs}; oracle decode {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
best = None
for rep in range(4):
    r = capi.encode_buckets(key, codes, bs, 0, per, B, alpha)
    best = r["device_ms"] if best is None else min(best, r["device_ms"])
# parity: every bucket text against the file
mism = 0
for b, i in enumerate(infos):
    nbytes = (i["clen"] * i["tbits"] + 7) // 8
    off = int(r["offsets"][b])
    want = raw[i["text_bit"] // 8: i["text_bit"] // 8 + nbytes] if i["alpha"] > 1 else b""
    if int(r["clen"][b]) != i["clen"] or r["payload"][off:off + len(want)].tobytes() != want:
        mism += 1
# CPU: the oracle's restatement, one core, a bounded sample of buckets
sample = list(range(0, B, max(1, B // 2000)))
t0 = time.perf_counter()
rows = 0
for b in sample:
    i = infos[b]
    if i["alpha"] > 1:
        O.pack_sequence(O.encode_bucket(key, b, codes[b * bs: b * bs + i["length"]], i["alpha"], bool(i["is_odd"])), i["tbits"])
    rows += i["length"]
cpu_s = time.perf_counter() - t0
print(json.dumps({"workload": wl, "rows": int(len(codes)), "buckets": B, "bucket_size": bs,
                  "gpu_kernels_ms": best, "gpu_rows_per_s": len(codes) / (best / 1e3),
                  "payload_MB": r["payload_bytes"] / 1e6,
                  "parity": {"buckets_checked": B, "mismatches": mism, "against": "the bytes the unmodified reference builder wrote"},
                  "cpu_port": {"rows_per_s": rows / cpu_s, "cores": 1, "sample": f"{len(sample)} buckets, oracle/orc_encode_bucket + orc_pack_sequence"}}))
