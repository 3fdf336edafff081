#!/usr/bin/env python
"""bench.py — body-steps/s of the batched multi-world stepper on N B200s (BASELINE.json metric).

A "step" is one simulation step (Simulator.updateSimulation: resolveAll then update) of EVERY world of the
batch.  Workload: BASELINE.json configs[2] ("c3": 16384 independent worlds x (64 boxes + 63 spheres + static
ground), friction on — the configuration the 1/2/4/8-GPU metric is quoted on; it fits one GPU), stepped from
creation for --settle steps (timed as part of the rollout figure), then W warm-up and K timed steps on the pile
at rest.  `--workload c2|c4` select configs[1] / configs[3]; at N=1 their short runs are appended to the same
line under "other_workloads".  Weak scaling: every rank owns the same number of worlds (world seeds continue
across ranks), no data-path collective; the statistics block is all-reduced once over NCCL after the timed
region.  After the timed region 16 worlds of the batch are re-stepped from creation on the CPU oracle and
compared bit for bit with the GPU batch ("parity_sample", outside every timed region).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA)
  python bench.py --impl reference ...                      the reference's CPU path: the oracle port
                                                            (the F# original cannot be built here: no dotnet)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

B_ALG = 416.0  # algorithmic bytes per body-step: 144 state read + 144 written + 128 parameters

WORKLOADS = {
    # name: (worlds per GPU, builder, dynamic bodies per world, description)
    "c2": (4096, "c2_spheres64", 64, "configs[1]: 4096 worlds x (64 spheres + static ground), EnableFriction=false"),
    "c3": (16384, "c3_mixed128", 127, "configs[2]: 16384 worlds x (64 boxes + 63 spheres + static ground), friction"),
    "c4": (8192, "c4_chain32", 32, "configs[3]: 8192 worlds x 32-link jointed box chains, restore_joints=1"),
}


def _build_chunk(job):
    name, w0, w1 = job
    from physicssimulator_b200 import scenes, prototypes as P
    built, joints, cfg = [], [], None
    fn = getattr(scenes, WORKLOADS[name][1])
    for w in range(w0, w1):
        r = fn(w)
        protos, cfg = r[0], r[1]
        sc = cfg.StepConfig
        built.append(P.build_world(protos, sc.GravitationalAccelerationMagnitude, sc.GravityDirection))
        if len(r) > 2:
            joints.append(r[2])
    return built, cfg, joints


def build_worlds(name, first_world, n_worlds, procs=1):
    """prototypes -> built bodies of worlds [first_world, first_world + n_worlds) (host side of the reference:
    RigidBodyPrototype.build); `procs` worker processes (the scene generators are plain Python)"""
    if procs <= 1 or n_worlds < 64:
        built, cfg, joints = _build_chunk((name, first_world, first_world + n_worlds))
        return built, cfg, (joints or None)
    import multiprocessing as mp
    per = (n_worlds + 4 * procs - 1) // (4 * procs)
    jobs = [(name, first_world + k, min(first_world + n_worlds, first_world + k + per)) for k in range(0, n_worlds, per)]
    with mp.get_context("fork").Pool(procs) as pool:
        parts = pool.map(_build_chunk, jobs)
    built = [b for p in parts for b in p[0]]
    joints = [j for p in parts for j in p[2]]
    return built, parts[0][1], (joints or None)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_sample(name, settle, steps, threads, first_world=0, budget_s=20.0):
    """the oracle (CPU restatement of the reference) on a bounded sample of the same workload"""
    from oracle.oracle import OracleWorld, batch_step
    from physicssimulator_b200.configuration import to_native_config
    n_dyn = WORKLOADS[name][2]
    # calibrate on one world so that settle + timed work stays near the budget
    built, cfg, joints = build_worlds(name, first_world, 1)
    jn = _native_joints(built, joints)
    w0 = OracleWorld(built[0], to_native_config(cfg), joints=jn[0] if jn else None)
    t = time.perf_counter(); w0.step(3); per_step = (time.perf_counter() - t) / 3
    n_sample = int(max(threads, min(64 * threads, budget_s * threads / max(per_step * (settle + steps), 1e-9))))
    n_sample = max(threads, (n_sample // threads) * threads)
    built, cfg, joints = build_worlds(name, first_world, n_sample)
    jn = _native_joints(built, joints)
    worlds = [OracleWorld(b, to_native_config(cfg), joints=jn[k] if jn else None) for k, b in enumerate(built)]
    batch_step(worlds, settle, threads)
    t = time.perf_counter()
    batch_step(worlds, steps, threads)
    dt = time.perf_counter() - t
    return {"value": n_dyn * n_sample * steps / dt, "unit": "body-steps/s", "cores": threads, "kind": "port",
            "sample": f"{n_sample} worlds of {name} (seeds {first_world}..), {settle} settle steps untimed then {steps} timed steps, "
                      f"oracle/ps_oracle.cpp on {threads} host threads; 'port' = CPU restatement of the F# reference (no dotnet in the image)",
            "seconds": dt}, worlds


def _native_joints(built, joints):
    if not joints:
        return None
    from physicssimulator_b200 import prototypes as P
    out = []
    for b, js in zip(built, joints):
        out.append([(a, bb) + P.ball_socket_local_anchors(b, a, bb, anchor) for (a, bb, anchor) in js])
    return out


def bind_near_gpu(local):
    """pin this rank to the CPU cores of its GPU's NUMA node BEFORE any pinned host buffer is allocated (first touch puts
    the pages on that node): the end-to-end leg moves the whole state over PCIe both ways every step"""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{dev}"
        node = int(open(base + "/numa_node").read().strip())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"pci": dev, "numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001  (containers without sysfs PCI entries: stay unbound, say so)
        return {"unbound": str(e)[:80]}


def workload_config(name, per_gpu, settle, world, sched=None):
    """the `config` object of the JSON line: identical keys in both arms (ours / --impl reference)"""
    return {"workload": f"{name}: {WORKLOADS[name][3]}", "worlds_per_gpu": per_gpu, "dynamic_bodies_per_world": WORKLOADS[name][2],
            "settle_steps": settle, "gpus": world,
            "l2": "256 MiB buffer written between timed iterations (outside the CUDA-event pairs)",
            "parallelism": f"worlds sharded over {world} GPU(s), no data-path collective"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    name = args.workload
    per_gpu = args.worlds or WORKLOADS[name][0]
    # each "step" = one simulation step of the bounded sample; warm-up = settle + W steps
    res, _ = cpu_sample(name, args.settle + max(args.warmup, 3), args.steps, threads)
    line = {"impl": "reference", "metric": "body-steps/sec", "value": res["value"], "unit": "body-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * res["seconds"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, per_gpu, args.settle, args.gpus),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "a bounded SAMPLE of the config's worlds (worlds are independent: the rate per world does not depend on the batch size) on this box's host threads; "
                    "at N > 1 the same host cores serve all N GPUs, so this value does not grow with N",
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def parity_sample(name, sim, first_world, per_gpu, n_steps, k=16):
    """K worlds of the batch re-stepped from creation on the CPU oracle for the number of steps the GPU batch has
    done, compared bit for bit (dynamic bodies: position, velocity, orientation, angular momentum).  Checker only:
    runs after the timed regions."""
    from oracle.oracle import OracleWorld, batch_step
    from physicssimulator_b200.configuration import to_native_config
    k = min(k, per_gpu)
    idx = sorted({(q * per_gpu) // k + ((q * 7919) % max(1, per_gpu // k)) for q in range(k)})
    worlds, masks = [], []
    for w in idx:
        built, cfg, joints = build_worlds(name, first_world + w, 1)
        jn = _native_joints(built, joints)
        worlds.append(OracleWorld(built[0], to_native_config(cfg), joints=jn[0] if jn else None))
        masks.append(~np.isinf(built[0]["mass"]))
    t = time.perf_counter()
    batch_step(worlds, n_steps, min(len(worlds), os.cpu_count() or 1))
    secs = time.perf_counter() - t
    bad = []
    for w, o, m in zip(idx, worlds, masks):
        got, ref = sim.batch.get_state(w, w + 1), o.get_state()
        for f in ("pos", "vel", "rot", "ang_mom"):
            a, b = np.ascontiguousarray(got[f][m]), np.ascontiguousarray(ref[f][m])
            same = (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
            if not same.all():
                bad.append((int(w), f))
    return {"worlds": len(idx), "world_ids": [int(first_world + w) for w in idx], "steps_from_creation": n_steps, "ok": not bad,
            "mismatches": bad[:8], "compared": "pos, vel, rot, ang_mom of every dynamic body, bit for bit (NaN = NaN)",
            "oracle_seconds": secs}


def run_workload(name, args, rank, local, world, main_line):
    """one workload on this rank's GPU: rollout from creation, warm-up, K timed steps, parity sample, end-to-end legs"""
    import torch
    import torch.distributed as dist
    from physicssimulator_b200.simulator import BatchSimulator
    from physicssimulator_b200.sharding import allreduce_stats

    per_gpu = args.worlds or WORKLOADS[name][0]
    n_dyn = WORKLOADS[name][2]
    steps = args.steps if main_line else max(3, min(args.steps, 10))
    warm = max(args.warmup, 3)
    settle = args.settle if args.settle >= 0 else WORKLOAD_SETTLE[name]
    rollout_to = max(args.rollout, settle + warm + steps) if main_line else settle + warm + steps

    procs = max(1, min(16, (os.cpu_count() or 1) // max(1, world)))
    built, cfg, joints = build_worlds(name, rank * per_gpu, per_gpu, procs)
    sim = BatchSimulator(None, cfg, joints=joints, device=local, _prebuilt=built)
    del built
    n_bodies, n_dynamic = sim.batch.num_bodies()
    assert n_dynamic == n_dyn * per_gpu
    # a non-default stream of our own: the kernels are launched on it through ps_batch_step_async and the CUDA
    # events below are recorded on it (torch.cuda.Event sees only torch's current stream)
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    E = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def timed_block(n):  # n steps, one call each, no flush: part of the rollout figure
        a, b = E(), E()
        a.record()
        for _ in range(n):
            sim.batch.step_async(1, stream)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    done = 0
    roll_ms = timed_block(settle); done += settle          # steps [0, settle) from creation
    roll_ms += timed_block(warm); done += warm             # W warm-up steps (untimed for the headline)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    ev = [(E(), E()) for _ in range(steps)]
    sampler = ClockSampler(local)
    if rank == 0 and main_line:
        sampler.start()
        time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = sim.batch.schedule()["kernel_launches"]
    t0 = time.perf_counter()
    for k in range(steps):
        flush.fill_(k & 0xff)           # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record()
        sim.batch.step_async(1, stream)  # one step of every world
        ev[k][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    done += steps
    sched = sim.batch.schedule()
    gpu_launches = sched["kernel_launches"] - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    roll_ms += dev_ms
    clocks = sampler.stop() if (rank == 0 and main_line) else None
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tmax.item())
    value = n_dyn * per_gpu * world * steps / (dev_ms_max * 1e-3)
    kernel_ms = dev_ms / steps  # this rank's average step duration

    if args.profile_step and main_line:  # one settled step for `ncu --profile-from-start off`; outside every timed region
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        sim.batch.step_async(1, stream)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        done += 1

    # the rest of the rollout from creation (what a caller of Simulator.Start sees: bodies falling, landing, settling)
    if rollout_to > done:
        roll_ms += timed_block(rollout_to - done)
        done = rollout_to
    troll = torch.tensor([roll_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(troll, op=dist.ReduceOp.MAX)
    rollout = {"steps_from_creation": done, "ms_per_step": float(troll.item()) / done,
               "value": n_dyn * per_gpu * world * done / (float(troll.item()) * 1e-3), "unit": "body-steps/s",
               "what": "average over EVERY step from creation (free fall, landing, settling, rest), device-timed, max over ranks; "
                       "the headline `value` is the K timed steps on the pile at rest"}
    if args.profile_step and main_line:
        rollout["what"] += " (one profiled step of this run is not in the time but in the step count)"

    stats_local = sim.Stats()
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_sample(name, sim, rank * per_gpu, per_gpu, done, 16 if main_line else 4)

    # end to end through the public API with HOST buffers: H2D of the step's input state from pinned memory,
    # the step, D2H of the resulting state — all inside the timed region
    st = sim.batch.get_state()
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in st.items()}
    host = {k: v.numpy() for k, v in pinned.items()}
    e2e_steps = max(3, min(steps, 10))
    e2e_seq_s = None
    if main_line:
        sim.batch.set_state(host); sim.batch.step(1); sim.batch.get_state(out=host)  # warm
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(e2e_steps):  # informational: the three calls one after the other
            sim.batch.set_state(host)
            sim.batch.step(1)
            sim.batch.get_state(out=host)
        e2e_seq_s = time.perf_counter() - t1
    sim.batch.step_host(host, 1, out=host)  # warm (creates the chunk streams)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        # the reference-facing call: state in host memory -> one step -> state in host memory (ps_batch_step_host)
        sim.batch.step_host(host, 1, out=host)
    e2e_s = time.perf_counter() - t1
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_dyn * per_gpu * world * e2e_steps / float(te.item())
    bytes_state = n_bodies * 18 * 8

    stats = allreduce_stats(stats_local)  # the only collective of the path: NCCL all-reduce of the stats block
    res = {"name": name, "per_gpu": per_gpu, "n_dyn": n_dyn, "steps": steps, "warm": warm, "settle": settle, "value": value,
           "ms_per_step": dev_ms_max / steps, "kernel_ms": kernel_ms, "sched": sched, "gpu_launches": gpu_launches, "clocks": clocks,
           "rollout": rollout, "parity": parity, "stats": stats, "wall": wall,
           "e2e": {"value": e2e_value, "unit": "body-steps/s", "h2d_bytes_per_step": bytes_state, "d2h_bytes_per_step": bytes_state,
                   "steps": e2e_steps, "ms_per_step": 1e3 * float(te.item()) / e2e_steps,
                   "what": "ps_batch_step_host(pinned host state in, 1 step, pinned host state out) per step: H2D + step + D2H inside the timed region, pipelined over world chunks"}}
    copy_ms = max(res["e2e"]["ms_per_step"] - kernel_ms, 1e-9)
    res["e2e"]["copy_GBps_both_ways"] = 2 * bytes_state / (copy_ms * 1e-3) / 1e9
    res["e2e"]["bound"] = ("host<->device copies: (e2e - device step) time moves h2d + d2h bytes at the rate above.  On the batch-wide path the call "
                           "splits the worlds into 4 chunks, each with its own chain of rounds on its own stream: the copies of one chunk overlap the "
                           "rounds of the others, but four staggered chains share the GPU less well than one (DESIGN.md 4.3)")
    if e2e_seq_s is not None:
        res["e2e"]["sequential_calls_value_this_rank"] = n_dyn * per_gpu * e2e_steps / e2e_seq_s
        res["e2e"]["sequential_calls_what"] = "ps_batch_set_state + ps_batch_step(1) + ps_batch_get_state, one after the other (informational)"
    sim.close()
    del sim, flush, pinned, host
    torch.cuda.empty_cache()
    return res


WORKLOAD_SETTLE = {"c2": 200, "c3": 200, "c4": 200}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--worlds", type=int, default=0, help="worlds per GPU (default: the config's)")
    ap.add_argument("--settle", type=int, default=200, help="steps from creation before the warm-up (timed only for the rollout figure)")
    ap.add_argument("--rollout", type=int, default=400, help="the rollout figure averages steps [0, this) from creation")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of 16 sampled worlds")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other workloads (N=1 only)")
    ap.add_argument("--profile-step", action="store_true",
                    help="bracket ONE extra step after the timed region with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local)
    affinity = bind_near_gpu(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = args.workload
    r = run_workload(name, args, rank, local, world, True)
    others = {}
    if world == 1 and not args.no_extra and not args.profile_step:
        for other in sorted(WORKLOADS):
            if other == name:
                continue
            o = run_workload(other, args, rank, local, world, False)
            others[other] = {"workload": WORKLOADS[other][3], "value": o["value"], "unit": "body-steps/s", "ms_per_step": o["ms_per_step"],
                             "steps": o["steps"], "settle_steps": o["settle"], "path": o["sched"]["path"], "e2e_value": o["e2e"]["value"],
                             "roofline_frac": o["n_dyn"] * o["per_gpu"] * B_ALG / (o["kernel_ms"] * 1e-3) / 1e9 / hbm_peak()[0],
                             "parity_sample": o["parity"], "nan_worlds": o["stats"]["nan_worlds"],
                             "valid": o["stats"]["nan_worlds"] == 0}

    peak, peak_src = hbm_peak()
    per_gpu, n_dyn, sched, kernel_ms = r["per_gpu"], r["n_dyn"], r["sched"], r["kernel_ms"]
    achieved = n_dyn * per_gpu * B_ALG / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_note = None, "no ncu capture matches this run (profiles/traffic.json)"
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
        if isinstance(t, dict) and t.get("worlds") == per_gpu and t.get("settle") == r["settle"] and t.get("path") == sched["path"]:
            traffic = t["dram_bytes_per_step"]
            traffic_note = f"ncu dram__bytes_read.sum + dram__bytes_write.sum of one step of this workload at the same settle count ({t.get('source', '')})"
    except Exception:
        pass

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            cpu, _ = cpu_sample(name, r["settle"] + r["warm"], r["steps"], os.cpu_count() or 1)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cfgd = workload_config(name, per_gpu, r["settle"], world)  # the same object in both arms
        schedule = {"path": sched["path"], "launch": ("fused: 1 step_kernel launch per step, %d lanes per world" % sched["lanes_per_world"]) if sched["path"] == "fused"
                    else "wide: per-stage kernels over the whole batch + the fused kernel for flagged worlds"}
        line = {
            "metric": "body-steps/sec", "value": r["value"], "unit": "body-steps/s", "n_gpus": world, "steps": r["steps"],
            "warmup": r["warm"], "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfgd,
            "schedule": schedule,
            "e2e": r["e2e"],
            "gpu_launches": r["gpu_launches"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "psk::step_kernel" if sched["path"] == "fused" else "psw::wide_* stages (one step of the whole batch)",
                         "algorithmic_bytes_per_body_step": B_ALG, "kernel_ms": kernel_ms, "peak_source": peak_src,
                         "traffic_note": traffic_note,
                         "note": "the per-world ordered contact solve is a dependent FP64 chain: the HBM roofline is the metric's denominator, not the kernel's limiter"},
            "rollout": r["rollout"],
            "parity_sample": r["parity"],
            "other_workloads": others or None,
            "cpu_baseline": cpu,
            "clocks": r["clocks"],
            "stats": r["stats"],
            "valid": r["stats"]["nan_worlds"] == 0,
            "wall_s_timed_region": r["wall"],
            "affinity_rank0": affinity,
        }
        if r["stats"]["nan_worlds"]:
            line["invalid_reason"] = f"{r['stats']['nan_worlds']} worlds hold NaN/inf state: the rate is not a meaningful workload rate"
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s"


if __name__ == "__main__":
    main()
