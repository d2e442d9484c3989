#!/usr/bin/env python
"""bench.py — BEV frames/s of the ST-P3 camera->BEV hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lift_splat|perceive] [--batch b_per_gpu]
  python bench.py --impl reference ...     # the reference's CPU path (oracle port) on the host cores

One "step" = one pass of the hot path over one batch of synthetic samples (6 cameras x 3 frames, 200x200 BEV);
one "frame" of the metric = one sample's BEV output.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from stp3_b200.utils import geometry as G  # noqa: E402
from stp3_b200.utils import synthetic as syn  # noqa: E402

METRIC = "bev_frames_per_sec"
UNIT = "frames/s"


def algorithmic_bytes_lift_splat(cfg, batch):
    """SURVEY.md §8d / BASELINE.md §3: read features + depth logits once, write every output frame once."""
    S, N, C, D = cfg.receptive_field, cfg.n_cameras, cfg.out_channels, cfg.n_depth
    Hf, Wf = cfg.feat_hw
    X, Y = cfg.bev_xy
    per_sample = 4 * S * N * Hf * Wf * (C + D) + 4 * S * C * X * Y + 4 * S * N * 25
    return per_sample * batch


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi sampled every 100 ms DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_problem(cfg, batch, seed):
    inp = syn.lift_inputs(cfg, batch, seed=seed)
    cam_M, cam_t, ego_R, ego_t = G.lift_matrices(inp["intrinsics"], inp["extrinsics"], inp["future_egomotion"])
    xs, ys, ds = G.frustum_axes(cfg.final_dim, cfg.downsample, cfg.d_bound)
    res, start, dim = G.calculate_birds_eye_view_parameters(cfg.x_bound, cfg.y_bound, cfg.z_bound)
    return dict(inp=inp, mats=(cam_M, cam_t, ego_R, ego_t), axes=(xs, ys, ds), res=res, start=start, dim=dim,
                off=G.bev_offset(start, res))


# ------------------------------------------------------------------------------------------------ reference arm
def reference_step_lift_splat(cfg, prob):
    """One sample through the op-for-op CPU port of the reference's lift-splat (oracle/torch_port.py)."""
    from oracle import torch_port as TP
    inp = prob["inp"]
    xs, ys, ds = prob["axes"]
    with torch.no_grad():
        return TP.lift_splat(inp["feat"][:1], inp["depth_logits"][:1], inp["intrinsics"][:1], inp["extrinsics"][:1],
                             inp["future_egomotion"][:1], xs, ys, ds, prob["res"], prob["start"], prob["dim"],
                             cfg.discount)


def time_reference(cfg, prob, steps, warmup):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    for _ in range(warmup):
        reference_step_lift_splat(cfg, prob)
    t0 = time.perf_counter()
    for _ in range(steps):
        reference_step_lift_splat(cfg, prob)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return 1.0 / dt, dt, cores


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob = make_problem(cfg, 1, seed=0)
    steps = min(args.steps, 3)
    warmup = min(args.warmup, 1)
    fps, dt, cores = time_reference(cfg, prob, steps, warmup)
    sample = f"1 sample (6 cam x {cfg.receptive_field} t) per step, {steps} timed step(s) after {warmup} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload_name, "batch_per_step": 1, "path": "oracle/torch_port.py (op-for-op CPU port of the reference; /root/reference cannot travel to the GPU box)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lift_splat", choices=["lift_splat"])
    ap.add_argument("--batch", type=int, default=4, help="samples per GPU per step (perceive config 4: 32 / 8 GPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = syn.CONFIGS["perceive"]
    args.workload_name = "lift_splat: 6 cam x 3 t x (28x60x48 frustum) -> 200x200x64 BEV, ego-warp + discount (BASELINE configs[2] lift-splat stage)"

    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return

    from stp3_b200 import ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl b200) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    W, K, b = max(args.warmup, 3), args.steps, args.batch

    prob = make_problem(cfg, b, seed=rank)
    inp = prob["inp"]
    host = {k: inp[k].pin_memory() for k in ("feat", "depth_logits")}
    host_mats = [m.pin_memory() for m in prob["mats"]]
    xs, ys, ds = (a.to(dev) for a in prob["axes"])
    d_feat, d_depth = host["feat"].to(dev), host["depth_logits"].to(dev)
    d_mats = [m.to(dev) for m in host_mats]
    X, Y = cfg.bev_xy
    out = torch.empty((b, cfg.receptive_field, cfg.out_channels, X, Y), dtype=torch.float32, device=dev)
    host_out = torch.empty(out.shape, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_resident():
        ops.lift_splat(d_feat, d_depth, *d_mats, xs, ys, ds, prob["off"], prob["res"], prob["dim"], cfg.discount,
                       out=out)

    def step_e2e():
        f = host["feat"].to(dev, non_blocking=True)
        d = host["depth_logits"].to(dev, non_blocking=True)
        mats = [m.to(dev, non_blocking=True) for m in host_mats]
        ops.lift_splat(f, d, *mats, xs, ys, ds, prob["off"], prob["res"], prob["dim"], cfg.discount, out=out)
        host_out.copy_(out, non_blocking=True)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for s, e in evs:
            flush.zero_()                      # evict L2 between timed iterations (not timed)
            s.record(); fn(); e.record()
        barrier()
        total_ms = sum(s.elapsed_time(e) for s, e in evs)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms

    for _ in range(W):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed(step_resident, K)
    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, K)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        ms_per_step = total_ms / K
        frames = b * world
        value = frames / (ms_per_step * 1e-3)
        pk = peaks()
        alg = algorithmic_bytes_lift_splat(cfg, b)
        achieved = alg / (ms_per_step * 1e-3) / 1e9
        h2d = sum(t.numel() * t.element_size() for t in list(host.values()) + host_mats)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload_name, "samples_per_gpu_per_step": b, "global_batch": frames,
                       "cameras": cfg.n_cameras, "frames": cfg.receptive_field, "bev": [X, Y],
                       "channels": cfg.out_channels, "depth_bins": cfg.n_depth, "parallelism": f"dp{world} (batch sharded, no collective)",
                       "l2": "flushed between timed iterations (256 MiB write)"},
            "clocks": clocks,
            "e2e": {"value": frames / (e2e_ms / K * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": host_out.numel() * 4},
            "gpu_launches": 2 * K * 2,   # scatter + finalize kernels per step (no memset), resident and e2e timed regions
            "roofline": {"bound": "hbm", "kernel": "lift-splat (scatter + finalize, one C-ABI call)",
                         "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / pk["hbm_gbs"], "peak_source": pk["source"], "traffic": None,
                         "algorithmic_bytes_per_step": alg},
        }
        if world == 1 and not args.no_cpu_baseline:
            fps, dt, cores = time_reference(cfg, make_problem(cfg, 1, seed=0), 1, 0)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "1 sample through oracle/torch_port.py (op-for-op CPU port of the reference lift-splat), 1 run"}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
