"""Scratch probe (GPU): where does the pipelined e2e step time go?"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from stp3_b200.utils import synthetic as syn
from stp3_b200.models.stp3 import GraphedPerception, PipelinedPerception

dev = torch.device("cuda:0")
cfg = syn.CONFIGS["perceive"]
b = 4
prob = bench.make_problem(cfg, b, 0)
inp = prob["inp"]
host = {k: inp[k].pin_memory() for k in ("feat", "depth_logits", "intrinsics", "extrinsics", "future_egomotion")}
model = bench.build_model(dev)
args_h = (host["feat"], host["depth_logits"], host["intrinsics"], host["extrinsics"], host["future_egomotion"])
with torch.no_grad():
    pipe = PipelinedPerception(model, b, cfg.n_cameras, depth=2, device=dev)

def run(n, label, collect_lag=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cpu_submit = 0.0
    for k in range(n):
        s0 = time.perf_counter()
        pipe.submit(*args_h)
        cpu_submit += time.perf_counter() - s0
        if k >= collect_lag:
            pipe.collect()
    while pipe.inflight:
        pipe.collect()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    print(f"{label}: {dt:.2f} ms/step, cpu in submit {cpu_submit / n * 1e3:.2f} ms")

for _ in range(3):
    pipe.submit(*args_h); pipe.collect()
run(20, "pipelined depth2")
run(20, "depth2 but collect immediately (serial)", collect_lag=0)
t0 = time.perf_counter()
for _ in range(20):
    model.prepare_inputs(host["intrinsics"], host["extrinsics"], host["future_egomotion"])
print("prepare_inputs cpu ms", (time.perf_counter() - t0) / 20 * 1e3, "threads", torch.get_num_threads())
torch.set_num_threads(4)
t0 = time.perf_counter()
for _ in range(20):
    model.prepare_inputs(host["intrinsics"], host["extrinsics"], host["future_egomotion"])
print("prepare_inputs cpu ms (4 threads)", (time.perf_counter() - t0) / 20 * 1e3)
run(20, "pipelined depth2, 4 cpu threads")
# raw copy timing
s = torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s):
    for _ in range(10):
        pipe.slots[0].static["feat"].copy_(host["feat"], non_blocking=True)
        pipe.slots[0].static["depth_logits"].copy_(host["depth_logits"], non_blocking=True)
torch.cuda.synchronize(); print("H2D 54MB ms", (time.perf_counter() - t0) / 10 * 1e3)
