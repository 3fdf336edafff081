#!/usr/bin/env python
"""Config 5 of BASELINE.json: one scene of 1 M spheres + 100 K boxes + ground; grid broad phase on the GPU, the oracle's
sort-and-sweep on the host.  Prints one JSON line (bodies/s, pairs, roofline of the pipeline with B_alg = 128 B per
body + 8 B per pair)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from physicssimulator_b200 import scenes, _native
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
b = scenes.c5_big()
n = len(b["mass"])
_native.broadphase_pairs(b, aabb_mode=mode, capacity=1 << 24)  # warm-up (context, allocations)
ms = []
for _ in range(5):
    pairs, kms = _native.broadphase_pairs(b, aabb_mode=mode, capacity=1 << 24)
    ms.append(kms)
kms = float(np.median(ms))
from oracle import oracle
t = time.perf_counter(); ref = oracle.overlap_pairs(b, aabb_mode=mode, method=1, cap=1 << 24); cpu_s = time.perf_counter() - t
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
alg = 128.0 * n + 8.0 * len(pairs)
print(json.dumps({"workload": "c5: 1M spheres + 100K boxes + ground, aabb_mode=%d" % mode, "bodies": n, "pairs": int(len(pairs)),
                  "identical_to_oracle": bool(np.array_equal(pairs, ref)), "kernel_ms": kms, "bodies_per_s": n / (kms * 1e-3),
                  "roofline": {"bound": "hbm", "algorithmic_bytes": alg, "achieved_GBs": alg / (kms * 1e-3) / 1e9, "peak_GBs": peak,
                               "frac": alg / (kms * 1e-3) / 1e9 / peak},
                  "cpu_oracle_sweep_s": cpu_s, "cpu_bodies_per_s": n / cpu_s}))
