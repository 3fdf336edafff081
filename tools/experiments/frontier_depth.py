#!/usr/bin/env python
"""CPU model of two schedules of the batch-wide rounds on settled C3 worlds (oracle states): the static levels of
ps_wide.cuh (a pair waits for the previous near pair of each of its bodies) against a frontier in which a near pair that
does NOT collide releases its successors at once (it changes nothing), so only colliding pairs cost a full round."""
import math, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import P, scenes, to_native_config
from oracle.oracle import OracleWorld

def world(seed, settle):
    protos, cfg = scenes.c3_mixed128(seed)
    b = P.build_world(protos, cfg.StepConfig.GravitationalAccelerationMagnitude, cfg.StepConfig.GravityDirection)
    nc = to_native_config(cfg, record_contacts=1)
    o = OracleWorld(b, nc)
    o.step(settle)
    st0 = o.get_state()
    o.step(1)
    c = o.contacts()
    return b, st0, c, cfg

def near_pairs(b, st, margin):
    n = len(b["mass"]); pos = st["pos"]; stat = np.isinf(b["mass"])
    r = np.where(b["shape"] == 0, b["dims"][:, 0], 0.5 * np.linalg.norm(b["dims"], axis=1))
    rb = np.where(b["shape"] == 0, math.sqrt(3) * b["dims"][:, 0], r)
    out = []
    for i in range(n):
        for j in range(i + 1, n):
            if stat[i] and stat[j]: continue
            ri = (rb[i] if b["shape"][j] != 0 else r[i]) + (0 if stat[i] else margin)
            rj = (rb[j] if b["shape"][i] != 0 else r[j]) + (0 if stat[j] else margin)
            if np.linalg.norm(pos[i] - pos[j]) <= ri + rj: out.append((i, j))
    return out, stat

def schedules(pairs, stat, hit):
    n_body = len(stat)
    # static levels
    last = [0] * n_body; lv = []
    for (i, j) in pairs:
        l = max(0 if stat[i] else last[i], 0 if stat[j] else last[j]) + 1
        lv.append(l)
        if not stat[i]: last[i] = l
        if not stat[j]: last[j] = l
    # frontier
    lastp = [-1] * n_body; pend = [0] * len(pairs); succ = [[] for _ in pairs]
    for k, (i, j) in enumerate(pairs):
        for bdy in (i, j):
            if stat[bdy]: continue
            if lastp[bdy] >= 0: pend[k] += 1; succ[lastp[bdy]].append(k)
            lastp[bdy] = k
    frontier = [k for k in range(len(pairs)) if pend[k] == 0]
    macro = micro = 0; done = 0; micro_per_macro = []; hits_per_macro = []
    while done < len(pairs):
        macro += 1; hits = []; m = 0
        while frontier:
            m += 1; nxt = []
            for k in frontier:
                if pairs[k] in hit: hits.append(k)
                else:
                    done += 1
                    for s in succ[k]:
                        pend[s] -= 1
                        if pend[s] == 0: nxt.append(s)
            frontier = nxt
        micro += m; micro_per_macro.append(m); hits_per_macro.append(len(hits))
        for k in hits:
            done += 1
            for s in succ[k]:
                pend[s] -= 1
                if pend[s] == 0: frontier.append(s)
    return max(lv), macro, micro, micro_per_macro, hits_per_macro

if __name__ == "__main__":
    seeds = [int(x) for x in sys.argv[1:]] or [0, 1, 2, 3]
    for s in seeds:
        b, st, c, cfg = world(s, 200)
        ids = c["pairs"]
        hit = set((int(i), int(j)) for i, j in ids)
        for margin in (0.03, 0.06):
            pairs, stat = near_pairs(b, st, margin)
            lv, macro, micro, mpm, hpm = schedules(pairs, stat, hit)
            print(f"seed {s} margin {margin}: near {len(pairs)} colliding {len(hit)} | static levels {lv} | frontier: macro rounds {macro}, micro rounds {micro}  micro/macro {mpm}  hits/macro {hpm}")

def batch(seeds, margin=0.05):
    import multiprocessing as mp
    with mp.Pool(16) as pool:
        res = pool.map(_one, [(s, margin) for s in seeds])
    L = max(r[0] for r in res); M = max(r[1] for r in res)
    tot_micro = 0; per = []
    for m in range(M):
        mm = max((r[3][m] if m < len(r[3]) else 0) for r in res); per.append(mm); tot_micro += mm
    print(f"{len(seeds)} worlds: static levels max {L} (mean {np.mean([r[0] for r in res]):.1f}) | frontier macro max {M} (mean {np.mean([r[1] for r in res]):.1f}), batch-wide micro rounds {tot_micro}: {per}")
def _one(a):
    s, margin = a
    b, st, c, cfg = world(s, 200)
    hit = set((int(i), int(j)) for i, j in c["pairs"])
    pairs, stat = near_pairs(b, st, margin)
    return schedules(pairs, stat, hit)
