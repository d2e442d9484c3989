"""Seeded synthetic inputs for the lift-splat / perception hot path.

There are no datasets or checkpoints in the build image, so benchmarks and parity tests use a ring
of N pinhole cameras around the ego vehicle, nuScenes-like ego-motion, post-ReLU context features
and Gaussian depth logits.  Everything is generated on the CPU with a fixed seed so that the build
container (where the reference runs) and the GPU box see bit-identical tensors.
"""
import math
import zlib
from dataclasses import dataclass, field
from typing import Tuple

import torch


@dataclass
class LiftSplatConfig:
    """The subset of the reference cfg the lift-splat reads (stp3/config.py:58-86)."""
    x_bound: Tuple[float, float, float] = (-50.0, 50.0, 0.5)
    y_bound: Tuple[float, float, float] = (-50.0, 50.0, 0.5)
    z_bound: Tuple[float, float, float] = (-10.0, 10.0, 20.0)
    d_bound: Tuple[float, float, float] = (2.0, 50.0, 1.0)
    final_dim: Tuple[int, int] = (224, 480)
    downsample: int = 8
    out_channels: int = 64
    discount: float = 0.5
    n_cameras: int = 6
    receptive_field: int = 3

    @property
    def feat_hw(self):
        return self.final_dim[0] // self.downsample, self.final_dim[1] // self.downsample

    @property
    def n_depth(self):
        return len(torch.arange(*self.d_bound, dtype=torch.float))

    @property
    def bev_xy(self):
        return (int((self.x_bound[1] - self.x_bound[0]) / self.x_bound[2]),
                int((self.y_bound[1] - self.y_bound[0]) / self.y_bound[2]))


# BASELINE.json configs
CONFIGS = {
    # 1 camera, 1 timestep, 50x50 BEV, C=64, D=32 (plumbing)
    "plumbing": LiftSplatConfig(x_bound=(-12.5, 12.5, 0.5), y_bound=(-12.5, 12.5, 0.5),
                                d_bound=(2.0, 34.0, 1.0), n_cameras=1, receptive_field=1),
    # 6 cameras, 1 timestep, 200x200, lift-splat only
    "lift_splat": LiftSplatConfig(receptive_field=1),
    # 6 cameras x 3 frames, 200x200 (+ ego warp + temporal fusion): perceive config
    "perceive": LiftSplatConfig(),
    # stress: 6cam x 5t, 400x400, D=96, C=128
    "stress": LiftSplatConfig(x_bound=(-100.0, 100.0, 0.5), y_bound=(-100.0, 100.0, 0.5),
                              d_bound=(2.0, 98.0, 1.0), out_channels=128, receptive_field=5),
    # CARLA-like resolution 0.2 m (true division matters)
    "carla_res": LiftSplatConfig(x_bound=(-20.0, 20.0, 0.2), y_bound=(-20.0, 20.0, 0.2),
                                 d_bound=(2.0, 26.0, 1.0), n_cameras=4, receptive_field=3,
                                 final_dim=(128, 128), out_channels=16),
    # tiny case for pure-python checks
    "tiny": LiftSplatConfig(x_bound=(-8.0, 8.0, 0.5), y_bound=(-8.0, 8.0, 0.5), d_bound=(2.0, 10.0, 1.0),
                            final_dim=(32, 48), out_channels=5, n_cameras=2, receptive_field=3),
}


def _rz(yaw):
    c, s = math.cos(yaw), math.sin(yaw)
    return torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64)


def camera_rig(cfg: LiftSplatConfig, batch: int, gen: torch.Generator, focal_jitter=True, random_pose=False,
               tilt_deg: float = 0.0):
    """intrinsics (B,S,N,3,3), extrinsics (B,S,N,4,4) fp32.  Ring of level pinhole cameras:
    yaw = 2*pi*n/N, position (1.5cos, 1.5sin, 1.5) m, camera z -> ego forward.
    tilt_deg > 0: every camera of every sample gets a fixed (same for all frames, like a real calibration) random
    roll / pitch / yaw error, each U[-tilt_deg, tilt_deg] degrees -- nuScenes calibrations are about 1 degree off
    level, which is what decides how many frustum points of an image column share a BEV pillar."""
    S, N = cfg.receptive_field, cfg.n_cameras
    H, W = cfg.final_dim
    cam2ego_axes = torch.tensor([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]], dtype=torch.float64)
    intr = torch.zeros(batch, S, N, 3, 3, dtype=torch.float64)
    extr = torch.zeros(batch, S, N, 4, 4, dtype=torch.float64)
    for n in range(N):
        yaw = 2.0 * math.pi * n / N
        R = _rz(yaw) @ cam2ego_axes
        t = torch.tensor([1.5 * math.cos(yaw), 1.5 * math.sin(yaw), 1.5], dtype=torch.float64)
        extr[:, :, n, :3, :3] = R
        extr[:, :, n, :3, 3] = t
        extr[:, :, n, 3, 3] = 1.0
    jit = torch.rand(batch, S, N, generator=gen, dtype=torch.float64) if focal_jitter else torch.zeros(batch, S, N, dtype=torch.float64)
    f = 0.55 * W * (1.0 + 0.01 * jit)
    intr[..., 0, 0] = f
    intr[..., 1, 1] = f
    intr[..., 0, 2] = W / 2.0
    intr[..., 1, 2] = H / 2.0
    intr[..., 2, 2] = 1.0
    if tilt_deg > 0.0:
        ang = (torch.rand(batch, N, 3, generator=gen, dtype=torch.float64) - 0.5) * 2.0 * math.radians(tilt_deg)
        for b in range(batch):
            for n in range(N):
                ax, ay, az = ang[b, n].tolist()
                Rx = torch.tensor([[1, 0, 0], [0, math.cos(ax), -math.sin(ax)], [0, math.sin(ax), math.cos(ax)]], dtype=torch.float64)
                Ry = torch.tensor([[math.cos(ay), 0, math.sin(ay)], [0, 1, 0], [-math.sin(ay), 0, math.cos(ay)]], dtype=torch.float64)
                extr[b, :, n, :3, :3] = _rz(az) @ Ry @ Rx @ extr[b, :, n, :3, :3]
    if random_pose:  # perturb every camera by a random small rotation + offset (property tests)
        ang = (torch.rand(batch, S, N, 3, generator=gen, dtype=torch.float64) - 0.5) * 0.6
        for idx in torch.cartesian_prod(torch.arange(batch), torch.arange(S), torch.arange(N)):
            b, s, n = idx.tolist()
            ax, ay, az = ang[b, s, n].tolist()
            Rx = torch.tensor([[1, 0, 0], [0, math.cos(ax), -math.sin(ax)], [0, math.sin(ax), math.cos(ax)]], dtype=torch.float64)
            Ry = torch.tensor([[math.cos(ay), 0, math.sin(ay)], [0, 1, 0], [-math.sin(ay), 0, math.cos(ay)]], dtype=torch.float64)
            extr[b, s, n, :3, :3] = _rz(az) @ Ry @ Rx @ extr[b, s, n, :3, :3]
        extr[..., :3, 3] += (torch.rand(batch, S, N, 3, generator=gen, dtype=torch.float64) - 0.5)
    return intr.float(), extr.float()


def egomotion(cfg: LiftSplatConfig, batch: int, gen: torch.Generator):
    """future_egomotion (B,S,6) = U[-.5,.5] * (2, 2, 0.1, 0.01, 0.01, 0.3): ~0.5 s of nuScenes motion."""
    scale = torch.tensor([2.0, 2.0, 0.1, 0.01, 0.01, 0.3])
    u = torch.rand(batch, cfg.receptive_field, 6, generator=gen) - 0.5
    return (u * scale).float()


def exact_gauss(shape, gen: torch.Generator) -> torch.Tensor:
    """Approximately N(0, 1.15^2) samples that are bit-identical on every machine: the sum of four
    14-bit integers from the (integer, ISA-independent) mt19937 stream, centred and scaled by a power
    of two.  torch.randn is not used because its vectorised Box-Muller differs between AVX2/AVX-512."""
    s = torch.randint(0, 1 << 14, (4, *shape), generator=gen, dtype=torch.int32).sum(0)
    return (s - (1 << 15)).float() / float(1 << 13)


def lift_inputs(cfg: LiftSplatConfig, batch: int, seed: int = 0, focal_jitter=True, random_pose=False,
                tilt_deg: float = 0.0):
    """Everything the lift-splat consumes, entering at the encoder's outputs:
      feat (B,S,N,C,Hf,Wf) = relu(randn)   (UpsamplingConcat ends in ReLU, convolutions.py:195)
      depth_logits (B,S,N,D,Hf,Wf) = 2*randn
      intrinsics, extrinsics, future_egomotion."""
    gen = torch.Generator().manual_seed(seed)
    S, N, C = cfg.receptive_field, cfg.n_cameras, cfg.out_channels
    Hf, Wf = cfg.feat_hw
    D = cfg.n_depth
    feat = exact_gauss((batch, S, N, C, Hf, Wf), gen).relu_()
    depth = exact_gauss((batch, S, N, D, Hf, Wf), gen) * 2.0
    intr, extr = camera_rig(cfg, batch, gen, focal_jitter=focal_jitter, random_pose=random_pose, tilt_deg=tilt_deg)
    ego = egomotion(cfg, batch, gen)
    return {"feat": feat, "depth_logits": depth, "intrinsics": intr, "extrinsics": extr,
            "future_egomotion": ego}


def stack_samples(cfg: LiftSplatConfig, seeds, **kw):
    """A batch assembled from single-sample draws: sample i == lift_inputs(cfg, 1, seed=seeds[i]).  Parity fixtures
    are recorded per sample seed, so any batch built this way can be checked sample by sample."""
    parts = [lift_inputs(cfg, 1, seed=int(s), **kw) for s in seeds]
    return {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}


def init_exact(module, seed=0):
    """Deterministic, machine-independent parameters/buffers keyed by state-dict name (integer RNG, power-of-two
    scaling): there are no checkpoints in the image, so benchmarks and parity tests run on seeded random weights,
    and the build container (reference) and the GPU box (drop-in) must hold bit-identical values.  BatchNorm running
    statistics are randomised so that the BN folding is exercised."""
    sd = module.state_dict()
    for name in sorted(sd):
        t = sd[name]
        gen = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) & 0x7FFFFFFF)
        if name.endswith("num_batches_tracked"):
            continue
        if name.endswith("running_var"):
            v = 0.5 + torch.randint(0, 1 << 14, t.shape, generator=gen).float() / float(1 << 14)
        elif name.endswith("running_mean"):
            v = exact_gauss(t.shape, gen) * 0.125
        elif t.dim() == 1 and name.endswith("weight"):          # BatchNorm gamma
            v = 1.0 + exact_gauss(t.shape, gen) * 0.125
        elif t.dim() == 1:                                       # biases / BatchNorm beta
            v = exact_gauss(t.shape, gen) * 0.125
        else:                                                    # conv weights ~ N(0, 2/fan_in), power-of-two scale
            fan_in = t[0].numel()
            scale = 2.0 ** round(math.log2((2.0 / fan_in) ** 0.5 / 1.155))
            v = exact_gauss(t.shape, gen) * scale
        t.copy_(v.to(t.dtype))
    return module
