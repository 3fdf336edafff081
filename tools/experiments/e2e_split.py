#!/usr/bin/env python
"""End-to-end rate (pinned host state in -> one step -> pinned host state out) when the worlds of a batch are spread over K
handles, each driven by its own host thread: the copies of one handle overlap the rounds of the others (a step of the
batch-wide path is a latency-bound chain that leaves most of the GPU idle)."""
import argparse, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from physicssimulator_b200.simulator import BatchSimulator
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3"); ap.add_argument("--worlds", type=int, default=16384)
ap.add_argument("--settle", type=int, default=200); ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--handles", default="1,2,4,8"); ap.add_argument("--device", action="store_true", help="state stays on the device: ps_batch_step only")
a = ap.parse_args()
bench.bind_near_gpu(0)
for K in [int(x) for x in a.handles.split(",")]:
    per = a.worlds // K
    sims, hosts = [], []
    for k in range(K):
        built, cfg, joints = bench.build_worlds(a.workload, k * per, per, min(16, os.cpu_count() or 1))
        sim = BatchSimulator(None, cfg, joints=joints, device=0, _prebuilt=built)
        sim.batch.step(a.settle)
        st = sim.batch.get_state()
        pinned = {n: torch.from_numpy(v).pin_memory() for n, v in st.items()}
        hosts.append({n: v.numpy() for n, v in pinned.items()}); sims.append((sim, pinned))
    n_dyn = bench.WORKLOADS[a.workload][2] * a.worlds
    for (sim, _), h in zip(sims, hosts): sim.batch.step_host(h, 1, out=h)
    torch.cuda.synchronize()
    bar = threading.Barrier(K + 1)
    def run(sim, h):
        bar.wait()
        for _ in range(a.steps):
            if a.device: sim.batch.step(1)
            else: sim.batch.step_host(h, 1, out=h)
        bar.wait()
    ths = [threading.Thread(target=run, args=(s[0], h)) for s, h in zip(sims, hosts)]
    for t in ths: t.start()
    bar.wait(); t0 = time.perf_counter(); bar.wait(); dt = time.perf_counter() - t0
    for t in ths: t.join()
    print(f"handles {K}: {1e3 * dt / a.steps:8.3f} ms per step of all {a.worlds} worlds, {n_dyn * a.steps / dt:.4g} body-steps/s {'device-resident' if a.device else 'end to end'}", flush=True)
    for s, _ in sims: s.close()
    del sims, hosts
