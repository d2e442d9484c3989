#!/usr/bin/env python
"""bench.py — BEV frames/s of the ST-P3 camera->BEV perception hot path on B200 (Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload perceive|lift_splat] [--batch b_per_gpu]
  python bench.py --impl reference ...     # the reference's CPU path (oracle port) on the host cores

One "step" = one pass of the hot path over one batch of synthetic samples (6 cameras x 3 frames, 200x200 BEV);
one "frame" of the metric = one sample's BEV perception output.  Rank 0 prints ONE JSON line.

  perceive   (default; BASELINE configs[3], the configuration the metric is quoted on): encoder outputs ->
             lift-splat -> ego-motion + 3-D temporal block + DeepLab head -> BEV decoder heads (segmentation,
             pedestrian, hdmap logits).  The EfficientNet trunk is third-party, absent from the image and excluded on
             both arms: inputs enter as the trunk-head outputs (context features + depth logits).
  lift_splat (BASELINE configs[2] lift stage): lift-splat only, output = (B,3,64,200,200) BEV features.
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
# SURVEY.md §8d / BASELINE.md §3: 2*MAC, unpadded channels, counted on the reference modules
GFLOP_TEMPORAL, GFLOP_DECODER = 134.7, 59.3
WORKLOADS = {
    "perceive": "perceive: 6 cam x 3 t x (28x60x48 frustum) -> lift-splat -> ego-warp + 3-D temporal block + DeepLab "
                "head -> BEV decoder heads (seg/ped/hdmap), 200x200x64 BEV (BASELINE configs[3]; EfficientNet trunk "
                "excluded on both arms)",
    "lift_splat": "lift_splat: 6 cam x 3 t x (28x60x48 frustum) -> 200x200x64 BEV, ego-warp + discount "
                  "(BASELINE configs[2] lift-splat stage)",
    "stress": "stress: 6 cam x 5 t x (28x60x96 frustum), C=128 -> lift-splat -> 4 temporal blocks + DeepLab head -> BEV "
              "decoder heads, 400x400 BEV (BASELINE configs[4]; a robustness / maximum-size run, not the headline metric)",
}
# the whole perception path runs for these workloads (the others stop after the lift-splat)
PERCEPTION = ("perceive", "stress")


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


def build_model(device=None, lcfg=None):
    """Random-init (seeded, machine-independent) perception model; the trunk is not needed (inputs enter after it)."""
    from stp3_b200.config import get_cfg
    from stp3_b200.models.stp3 import STP3
    over = None
    if lcfg is not None and lcfg is not syn.CONFIGS["perceive"]:
        over = {"LIFT": {"X_BOUND": list(lcfg.x_bound), "Y_BOUND": list(lcfg.y_bound), "Z_BOUND": list(lcfg.z_bound),
                         "D_BOUND": list(lcfg.d_bound)},
                "TIME_RECEPTIVE_FIELD": lcfg.receptive_field, "MODEL": {"ENCODER": {"OUT_CHANNELS": lcfg.out_channels}}}
    cfg = get_cfg(over)
    with torch.no_grad():
        model = STP3(cfg, backbone=torch.nn.Identity())
        geo = {k: getattr(model, k).detach().clone() for k in ("frustum", "bev_resolution", "bev_start_position", "bev_dimension")}
        syn.init_exact(model, seed=0)
        for k, v in geo.items():
            getattr(model, k).copy_(v)
    model.eval()
    return model.to(device) if device is not None else model


# ------------------------------------------------------------------------------------------------ reference arm
def reference_step(workload, cfg, prob, model):
    """One sample through the CPU port of the reference's path (oracle/torch_port.py + oracle/torch_dense.py: the same
    ATen operator sequence as the reference, which cannot travel to the GPU box)."""
    from oracle import torch_port as TP
    inp = prob["inp"]
    xs, ys, ds = prob["axes"]
    with torch.no_grad():
        bev = TP.lift_splat(inp["feat"][:1], inp["depth_logits"][:1], inp["intrinsics"][:1], inp["extrinsics"][:1],
                            inp["future_egomotion"][:1], xs, ys, ds, prob["res"], prob["start"], prob["dim"],
                            cfg.discount)
        if workload == "lift_splat":
            return bev
        from oracle import torch_dense as TD
        ego = inp["future_egomotion"][:1]
        ego = torch.cat([torch.zeros_like(ego[:, :1]), ego[:, :-1]], 1)
        x = torch.cat([bev, ego.view(1, -1, 6, 1, 1).expand(1, ego.shape[1], 6, *bev.shape[-2:])], dim=2)
        return TD.decoder(TD.temporal_model(x, model.temporal_model), model.decoder)


def pick_threads(cfg, prob):
    """The eager CPU path is dominated by small ops and slows down when oversubscribed: give it its best thread count."""
    from oracle import torch_port as TP
    inp = prob["inp"]
    xs, ys, ds = prob["axes"]
    cores = os.cpu_count() or 1
    best, best_t = cores, float("inf")
    for n in sorted({min(cores, c) for c in (8, 16, 32, 64, cores)}):
        torch.set_num_threads(n)
        t0 = time.perf_counter()
        with torch.no_grad():   # one frame of one sample
            TP.lift_splat(inp["feat"][:1, :1], inp["depth_logits"][:1, :1], inp["intrinsics"][:1, :1],
                          inp["extrinsics"][:1, :1], inp["future_egomotion"][:1, :1], xs, ys, ds, prob["res"],
                          prob["start"], prob["dim"], cfg.discount)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def time_reference(workload, cfg, steps, warmup):
    prob = make_problem(cfg, 1, seed=0)
    model = build_model(lcfg=cfg) if workload in PERCEPTION else None
    threads = pick_threads(cfg, prob)
    for _ in range(warmup):
        reference_step(workload, cfg, prob, model)
    t0 = time.perf_counter()
    for _ in range(steps):
        reference_step(workload, cfg, prob, model)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return 1.0 / dt, dt, threads


def run_reference_arm(args, cfg):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    steps, warmup = min(args.steps, 3), min(args.warmup, 1)
    fps, dt, threads = time_reference(args.workload, cfg, steps, warmup)
    sample = (f"1 sample (6 cam x {cfg.receptive_field} t) per step, {steps} timed step(s) after {warmup} warm-up, "
              f"{threads} of {os.cpu_count()} host threads (best of a short sweep)")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "batch_per_step": 1,
                   "path": "oracle/torch_port.py + oracle/torch_dense.py: op-for-op CPU port of the reference "
                           "(/root/reference is pure Python and cannot travel to the GPU box)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="perceive", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=4, help="samples per GPU per step (perceive config 4: 32 / 8 GPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of one CUDA graph")
    ap.add_argument("--no-pipeline", action="store_true", help="e2e: serialise H2D, compute and D2H of every step")
    ap.add_argument("--profiler-range", action="store_true",
                    help="bracket the resident timed steps with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    args = ap.parse_args()
    cfg = syn.CONFIGS["stress" if args.workload == "stress" else "perceive"]

    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return

    from stp3_b200 import ops
    # the host side of a step is a few hundred floats of calibration math (torch.inverse etc.); with the default
    # 64-128 intra-op threads every such call pays a multi-millisecond OpenMP wake-up, so keep the pool small
    torch.set_num_threads(min(4, os.cpu_count() or 1))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl b200) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    W, K, b = max(args.warmup, 3), args.steps, args.batch
    perceive = args.workload in PERCEPTION

    prob = make_problem(cfg, b, seed=rank)          # every rank works on its own shard of the global batch
    inp = prob["inp"]
    host = {k: inp[k].pin_memory() for k in ("feat", "depth_logits", "intrinsics", "extrinsics", "future_egomotion")}
    host_mats = [m.pin_memory() for m in prob["mats"]]
    xs, ys, ds = (a.to(dev) for a in prob["axes"])
    d_feat, d_depth = host["feat"].to(dev), host["depth_logits"].to(dev)
    d_mats = [m.to(dev) for m in host_mats]
    X, Y = cfg.bev_xy
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    model = build_model(dev, cfg) if perceive else None
    graphed = None
    if perceive and not args.no_graph:
        from stp3_b200.models.stp3 import GraphedPerception
        graphed = GraphedPerception(model, b, cfg.n_cameras, dev)
        graphed(d_feat, d_depth, inp["intrinsics"], inp["extrinsics"], inp["future_egomotion"])   # fill the static inputs
    if perceive:
        host_out = {"segmentation": torch.empty((b, cfg.receptive_field, 2, X, Y)).pin_memory(),
                    "pedestrian": torch.empty((b, cfg.receptive_field, 2, X, Y)).pin_memory(),
                    "hdmap": torch.empty((b, 4, X, Y)).pin_memory()}
    else:
        out = torch.empty((b, cfg.receptive_field, cfg.out_channels, X, Y), dtype=torch.float32, device=dev)
        host_out = {"bev": torch.empty(out.shape, dtype=torch.float32).pin_memory()}

    def step_resident():
        with torch.no_grad():
            if graphed is not None:
                graphed.graph.replay()            # inputs already resident in the graph's static buffers
                return graphed.out
            if perceive:
                return model.forward_features(d_feat, d_depth, inp["intrinsics"], inp["extrinsics"], inp["future_egomotion"])
            ops.lift_splat(d_feat, d_depth, *d_mats, xs, ys, ds, prob["off"], prob["res"], prob["dim"], cfg.discount, out=out)
            return {"bev": out}

    def step_e2e():
        """The call a user makes: host (pinned) inputs in, host results out."""
        with torch.no_grad():
            if graphed is not None:
                res = graphed(host["feat"], host["depth_logits"], host["intrinsics"], host["extrinsics"],
                              host["future_egomotion"])
            elif perceive:
                f = host["feat"].to(dev, non_blocking=True)
                d = host["depth_logits"].to(dev, non_blocking=True)
                res = model.forward_features(f, d, host["intrinsics"], host["extrinsics"], host["future_egomotion"])
            else:
                f = host["feat"].to(dev, non_blocking=True)
                d = host["depth_logits"].to(dev, non_blocking=True)
                mats = [m.to(dev, non_blocking=True) for m in host_mats]
                ops.lift_splat(f, d, *mats, xs, ys, ds, prob["off"], prob["res"], prob["dim"], cfg.discount, out=out)
                res = {"bev": out}
            for k, t in host_out.items():
                t.copy_(res[k], non_blocking=True)

    from stp3_b200 import parallel

    def barrier():
        parallel.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for s, e in evs:
            flush.zero_()                      # evict L2 between timed iterations (not timed)
            s.record(); fn(); e.record()
        barrier()
        return parallel.max_over_ranks(sum(s.elapsed_time(e) for s, e in evs), dev)

    def timed_pipelined(pipe, steps):
        """End-to-end throughput with the copies of neighbouring steps overlapped with compute (PipelinedPerception):
        every step still moves its inputs host->device and its results device->host inside the timed region."""
        args_h = (host["feat"], host["depth_logits"], host["intrinsics"], host["extrinsics"], host["future_egomotion"])
        pipe.between_steps = flush.zero_            # L2 eviction between steps, on the compute stream
        for _ in range(3):
            pipe.submit(*args_h); pipe.collect()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for k in range(steps):
            pipe.submit(*args_h)
            if k >= 1:
                pipe.collect()
        pipe.collect()
        t1.record()
        barrier()
        return parallel.max_over_ranks(t0.elapsed_time(t1), dev)

    for _ in range(W):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if args.profiler_range:
        torch.cuda.cudart().cudaProfilerStart()
    total_ms = timed(step_resident, K)
    if args.profiler_range:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    e2e_mode = "serial: H2D -> compute -> D2H per step"
    if graphed is not None and not args.no_pipeline:
        from stp3_b200.models.stp3 import PipelinedPerception
        with torch.no_grad():
            pipe = PipelinedPerception(model, b, cfg.n_cameras, depth=2, device=dev)
        e2e_ms = timed_pipelined(pipe, K)
        e2e_mode = "pipelined (depth 2): copies of steps i-1 / i+1 overlap the CUDA graph of step i"
    else:
        for _ in range(2):
            step_e2e()
        e2e_ms = timed(step_e2e, K)
    clocks = sampler.stop() if rank == 0 else None

    # per-stage device time: each stage captured as its own CUDA graph (no launch gaps), timed with CUDA events
    stage_ms = {}
    if perceive:
        stage_ms = time_stages(model, graphed.static if graphed is not None else None, d_feat, d_depth, inp, flush, dev)

    if rank == 0:
        ms_per_step = total_ms / K
        frames = b * world
        value = frames / (ms_per_step * 1e-3)
        pk = peaks()
        alg = algorithmic_bytes_lift_splat(cfg, b)
        h2d = sum(host[k].numel() * host[k].element_size() for k in ("feat", "depth_logits"))
        h2d += sum(t.numel() * t.element_size() for t in host_mats)
        d2h = sum(t.numel() * 4 for t in host_out.values())
        ls_ms = stage_ms.get("lift_splat", ms_per_step)
        roof_ls = {"bound": "hbm", "kernel": "lift-splat (scatter + finalize, one C-ABI call)",
                   "achieved": alg / (ls_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                   "frac": alg / (ls_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "peak_source": pk["source"],
                   # dram__bytes_read+write of scatter (105.2 MB) + finalize (154.3 MB) for this workload at B=4,
                   # profiles/r01_ncu_liftsplat_v10_summary.txt (one ncu --set full capture; scales with B)
                   "traffic": int(259.5e6 * b / 4) if args.workload == "perceive" else None,
                   "algorithmic_bytes_per_step": alg, "ms": ls_ms}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (dense layers: bf16x3 split products, fp32 accumulate)" if perceive else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "samples_per_gpu_per_step": b, "global_batch": frames,
                       "cameras": cfg.n_cameras, "frames": cfg.receptive_field, "bev": [X, Y],
                       "channels": cfg.out_channels, "depth_bins": cfg.n_depth,
                       "parallelism": f"dp{world} (batch sharded, no collective on the forward path)",
                       "l2": "flushed between timed iterations (256 MiB write)", "weights": "random init, seeded"},
            "clocks": clocks,
            "e2e": {"value": frames / (e2e_ms / K * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "mode": e2e_mode},
        }
        if perceive:
            dense_ms = stage_ms.get("temporal_model", 0.0) + stage_ms.get("decoder", 0.0)
            # the flop count was taken on the reference modules for the perceive configuration only
            flops = (GFLOP_TEMPORAL + GFLOP_DECODER) * 1e9 * b if args.workload == "perceive" else None
            ach = flops / (dense_ms * 1e-3) / 1e12 if flops and dense_ms > 0 else None
            line["stage_ms"] = stage_ms
            # every extra temporal block (stress: 4 instead of 2) adds 5 convs + 3 small kernels
            line["gpu_launches"] = 2 * K * (LAUNCHES_PER_PERCEIVE_STEP + 8 * max(0, cfg.receptive_field - 3))
            line["roofline"] = {"bound": "tensor", "kernel": "conv_igemm_kernel<BN, PAIR, STACK> family (temporal model + decoder, 41 launches/step)",
                                "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                "frac": ach / pk["bf16_tflops_sustained"] if ach else None, "peak_source": pk["source"],
                                # dram bytes of the 22 conv launches captured in profiles/r01_ncu_conv_v10_summary.txt
                                # (the temporal model's 16 + the first 6 decoder convs, B=4): 4093 MB read + 2330 MB written
                                "traffic": int(6423e6 * b / 4) if args.workload == "perceive" else None, "traffic_note": "22 of 41 launches (ncu --set full, cold cache)",
                                "algorithmic_flops_per_step": flops, "ms": dense_ms,
                                "note": "algorithmic 2*MAC flops of the fp32 layers; the kernel issues 3 bf16 MMAs per product (hi*hi+hi*lo+lo*hi) to hold 1e-3 parity"}
            line["roofline_lift_splat"] = roof_ls
        else:
            line["gpu_launches"] = 2 * K * 2
            line["roofline"] = roof_ls
        if perceive and os.environ.get("STP3_TUNE_REPORT"):
            from stp3_b200 import dense
            for desc, times in dense.TUNE_LOG:
                print("TUNE", desc.ljust(44), "  ".join(f"{k}:{v * 1e3:7.1f}us" for k, v in times.items()), file=sys.stderr)
        if world == 1 and not args.no_cpu_baseline:
            fps, dt, threads = time_reference(args.workload, cfg, 1, 0)       # picks its own (best) thread count
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"1 sample through the op-for-op CPU port of the reference (oracle/torch_port.py"
                                              f"{' + oracle/torch_dense.py' if perceive else ''}), 1 run, {threads} threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def time_stages(model, static, d_feat, d_depth, inp, flush, dev, reps=5):
    """GPU time of the three stages of STP3.forward_device, each replayed as its own CUDA graph."""
    from stp3_b200 import dense, ops
    from stp3_b200.models.stp3 import STP3  # noqa: F401
    if static is None:
        h = {k: v.to(dev) for k, v in model.prepare_inputs(inp["intrinsics"], inp["extrinsics"], inp["future_egomotion"]).items()}
        static = dict(feat=d_feat, depth_logits=d_depth, **h)
    B, S = static["feat"].shape[:2]
    X, Y = model.bev_size
    C = model.encoder_out_channels
    off, res, dim = model._bev_host()
    planes = torch.empty((2, B, S, X, Y, C), dtype=torch.bfloat16, device=dev)
    hold = {}

    def lift():
        r = ops.lift_splat(static["feat"], static["depth_logits"], static["cam_M"], static["cam_t"], static["ego_R"],
                           static["ego_t"], *model._axes(), off, res, dim, float(model.discount), workspace=model._ws,
                           out_hilo=planes, pool_sum=True)
        hold["sums"] = r[1].view(B * S, C)

    def temporal():
        hold["states"] = model.temporal_model.forward_hl(dense.HL(planes[0], planes[1], C), const=static["const"],
                                                         sums=hold["sums"])

    def decoder():
        hold["out"] = model.decoder.forward_hl(hold["states"])

    out = {}
    with torch.no_grad():
        for name, fn in (("lift_splat", lift), ("temporal_model", temporal), ("decoder", decoder)):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            tot = 0.0
            for _ in range(reps):
                flush.zero_()
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); g.replay(); c.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(c)
            out[name] = tot / reps
    return out


# kernels of this repository launched by one perceive step (profiles/r01_launches_perceive_v9.csv): 41 conv_igemm,
# lift-splat scatter + finalize + pool reduce, 3 upsample, 3 pool_bias, 2 small_linear, 2 col_sum_reduce
LAUNCHES_PER_PERCEIVE_STEP = 54


if __name__ == "__main__":
    main()
