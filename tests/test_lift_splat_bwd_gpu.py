"""GPU parity of the lift-splat backward (stp3_lift_splat_bwd) against (a) the gradients the
UNMODIFIED reference's autograd produced (tests/golden/lift_splat_bwd_*.npz) and (b) the fp64 backward oracle on other
shapes.  Bar: 1e-3 of max (north_star's floating-point tolerance); measured ~1e-6."""
import numpy as np
import pytest
import torch

from oracle import lift_splat_oracle as O
from oracle.make_golden_bwd import loss_weights
from stp3_b200 import ops
from stp3_b200.utils import geometry as G
from stp3_b200.utils import synthetic as syn
from tests.test_bwd_oracle_cpu import bwd_case, check

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", ["tiny_randpose", "tiny_level", "carla_res"])
def test_backward_matches_reference_autograd(name):
    cfg, inp, g, fwd = bwd_case(name)
    mats = [torch.from_numpy(fwd[k]) for k in ("cam_M", "cam_t", "ego_R", "ego_t")]
    axes = [torch.from_numpy(fwd[k]) for k in ("xs", "ys", "ds")]
    off, res, dim = torch.from_numpy(fwd["bev_offset"]), torch.from_numpy(fwd["bev_resolution"]), torch.from_numpy(fwd["bev_dimension"])
    X, Y = int(dim[0]), int(dim[1])
    B, S, _, C = inp["feat"].shape[:4]
    W = loss_weights((B, S, C, X, Y), int(g["w_seed"])).to(DEV)
    gf, gd = ops.lift_splat_backward(W, inp["feat"].to(DEV), inp["depth_logits"].to(DEV), *mats, *axes, off, res, dim, cfg.discount)
    torch.cuda.synchronize()
    check(g, "grad_feat", gf.double().cpu().numpy(), 1e-4)
    check(g, "grad_depth", gd.double().cpu().numpy(), 1e-4)


@pytest.mark.parametrize("cfg_name,batch,rp,use_dd", [("plumbing", 1, False, True), ("plumbing", 2, True, False),
                                                      ("lift_splat", 1, True, True)])
def test_backward_matches_fp64_oracle(cfg_name, batch, rp, use_dd):
    """C = 64 (one full channel pass), D = 32 / 48, 28x60 feature maps; with and without the depth distribution."""
    cfg = syn.CONFIGS[cfg_name]
    inp = syn.lift_inputs(cfg, batch, seed=21, random_pose=rp)
    cam_M, cam_t, ego_R, ego_t = G.lift_matrices(inp["intrinsics"], inp["extrinsics"], inp["future_egomotion"])
    xs, ys, ds = G.frustum_axes(cfg.final_dim, cfg.downsample, cfg.d_bound)
    res, start, dim = G.calculate_birds_eye_view_parameters(cfg.x_bound, cfg.y_bound, cfg.z_bound)
    off = G.bev_offset(start, res)
    feat, depth = inp["feat"].to(DEV), inp["depth_logits"].to(DEV)
    out, ranks = ops.lift_splat(feat, depth, cam_M, cam_t, ego_R, ego_t, xs, ys, ds, off, res, dim, cfg.discount,
                                use_depth_distribution=use_dd, return_ranks=True)
    W = loss_weights(tuple(out.shape), 5)
    gf, gd = ops.lift_splat_backward(W.to(DEV), feat, depth, cam_M, cam_t, ego_R, ego_t, xs, ys, ds, off, res, dim,
                                     cfg.discount, use_depth_distribution=use_dd)
    torch.cuda.synchronize()
    ogf, ogd = O.lift_splat_backward(W.numpy(), inp["feat"].numpy(), inp["depth_logits"].numpy(), ranks.cpu().numpy(),
                                     cfg.discount, use_depth_distribution=use_dd)
    assert np.abs(gf.double().cpu().numpy() - ogf).max() <= 1e-5 * np.abs(ogf).max()
    if use_dd:
        assert np.abs(gd.double().cpu().numpy() - ogd).max() <= 1e-5 * np.abs(ogd).max()
    else:
        assert gd is None


def test_autograd_through_the_drop_in_module():
    """STP3.projection_to_birds_eye_view is differentiable w.r.t. features and depth logits (the trainer's use,
    trainer.py:115 -> stp3.py:303-318): loss.backward() through the CUDA forward/backward pair equals the oracle."""
    from stp3_b200.config import get_cfg
    from stp3_b200.models.stp3 import STP3
    cfg = get_cfg({"LIFT": {"X_BOUND": [-8.0, 8.0, 0.5], "Y_BOUND": [-8.0, 8.0, 0.5], "D_BOUND": [2.0, 10.0, 1.0]},
                   "IMAGE": {"FINAL_DIM": (32, 48)}})
    lcfg = syn.LiftSplatConfig(x_bound=(-8.0, 8.0, 0.5), y_bound=(-8.0, 8.0, 0.5), d_bound=(2.0, 10.0, 1.0),
                               final_dim=(32, 48), out_channels=64, n_cameras=2, receptive_field=3)
    inp = syn.lift_inputs(lcfg, 1, seed=9, random_pose=True)
    model = STP3(cfg, backbone=torch.nn.Identity()).to(DEV)
    feat = inp["feat"].to(DEV).requires_grad_(True)
    depth = inp["depth_logits"].to(DEV).requires_grad_(True)
    bev = model.projection_to_birds_eye_view(feat, depth, inp["intrinsics"], inp["extrinsics"], inp["future_egomotion"])
    W = loss_weights(tuple(bev.shape), 3)
    (bev * W.to(DEV)).sum().backward()
    cam_M, cam_t, ego_R, ego_t = G.lift_matrices(inp["intrinsics"], inp["extrinsics"], inp["future_egomotion"])
    xs, ys, ds = G.frustum_axes(lcfg.final_dim, lcfg.downsample, lcfg.d_bound)
    res, start, dim = G.calculate_birds_eye_view_parameters(lcfg.x_bound, lcfg.y_bound, lcfg.z_bound)
    ora = O.lift_splat(inp["feat"].numpy(), inp["depth_logits"].numpy(), cam_M.numpy(), cam_t.numpy(), ego_R.numpy(),
                       ego_t.numpy(), xs.numpy(), ys.numpy(), ds.numpy(), G.bev_offset(start, res).numpy(), res.numpy(),
                       dim.numpy(), lcfg.discount)
    ogf, ogd = O.lift_splat_backward(W.numpy(), inp["feat"].numpy(), inp["depth_logits"].numpy(), ora["rank"], lcfg.discount)
    assert np.abs(feat.grad.double().cpu().numpy() - ogf).max() <= 1e-5 * np.abs(ogf).max()
    assert np.abs(depth.grad.double().cpu().numpy() - ogd).max() <= 1e-5 * np.abs(ogd).max()
