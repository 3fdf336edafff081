import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, bench
from physicssimulator_b200.simulator import BatchSimulator
bench.bind_near_gpu(0)
built, cfg, joints = bench.build_worlds("c3", 0, 16384, 16)
sim = BatchSimulator(None, cfg, joints=joints, device=0, _prebuilt=built)
sim.batch.step(200)
st = sim.batch.get_state()
pinned = {n: torch.from_numpy(v).pin_memory() for n, v in st.items()}
host = {n: v.numpy() for n, v in pinned.items()}
for _ in range(3): sim.batch.step_host(host, 1, out=host)
t0 = time.perf_counter()
for _ in range(8): sim.batch.step_host(host, 1, out=host)
print("ms per call", 1e3 * (time.perf_counter() - t0) / 8)
