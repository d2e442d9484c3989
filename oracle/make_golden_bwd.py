"""Generates tests/golden/lift_splat_bwd_*.npz: gradients of the UNMODIFIED reference's lift-splat (softmax over depth,
outer product, VoxelsSumming.backward, discount recurrence; stp3.py:214-301, geometry.py:299-330) obtained with its own
autograd on CPU in the build container, for loss = sum(bev * W) with a seeded W.  Test infrastructure; run by hand:
    python -m oracle.make_golden_bwd
Small cases store every tensor; `carla_res` (true division, 4 cameras) stores a 20k-entry sample of each gradient."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402
from oracle.make_golden import run_reference, OUT  # noqa: E402
from stp3_b200.utils import synthetic as syn  # noqa: E402

CASES = [("tiny_randpose", "tiny", 2, 11, True, True), ("tiny_level", "tiny", 1, 12, False, True),
         ("carla_res", "carla_res", 1, 3, True, False)]


def loss_weights(shape, seed=99):
    return syn.exact_gauss(shape, torch.Generator().manual_seed(seed))


def main():
    ref = load_reference()
    for name, cfg_name, batch, seed, rp, full in CASES:
        cfg = syn.CONFIGS[cfg_name]
        inp = syn.lift_inputs(cfg, batch, seed=seed, random_pose=rp)
        feat = inp["feat"].clone().requires_grad_(True)
        depth = inp["depth_logits"].clone().requires_grad_(True)
        bev = run_reference(ref, cfg, dict(inp, feat=feat, depth_logits=depth))[3]
        W = loss_weights(bev.shape)
        (bev * W).sum().backward()
        rec = dict(config=cfg_name, batch=batch, seed=seed, random_pose=rp, w_seed=99)
        gf, gd = feat.grad, depth.grad
        if full:
            rec.update(grad_feat=gf.numpy(), grad_depth=gd.numpy())
        else:
            g = torch.Generator().manual_seed(77)
            for key, t in (("grad_feat", gf), ("grad_depth", gd)):
                idx = torch.randint(0, t.numel(), (20_000,), generator=g)
                rec[key + "_index"] = idx.numpy()
                rec[key + "_value"] = t.reshape(-1)[idx].numpy()
                rec[key + "_max"] = float(t.abs().max())
        path = os.path.join(OUT, f"lift_splat_bwd_{name}.npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: |g_feat| max {float(gf.abs().max()):.4g}, |g_depth| max {float(gd.abs().max()):.4g} -> {path} "
              f"({os.path.getsize(path) / 1e3:.0f} kB)")


if __name__ == "__main__":
    main()
