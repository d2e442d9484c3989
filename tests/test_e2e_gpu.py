"""GPU parity at the HEADLINE size: the drop-in STP3 on one perceive-config sample (6 cameras x 3 frames, 200x200x64
BEV; every dilated ASPP tap live) against tests/golden/e2e_perceive_*.npz -- the unmodified reference run end to end
(lift-splat -> TemporalModel -> Decoder) and the fp64 oracle on the same inputs (oracle/make_golden_e2e.py).

Bars (north_star): voxel ranks bit-exact (SHA-256 of the reference's rank array); BEV features, temporal-model output
and head logits within 1e-3 of max of the reference; and, tighter, within 3e-4 of the fp64 oracle (the reference's
own fp32 path sits 2e-4 from it)."""
import numpy as np
import pytest
import torch

from stp3_b200 import ops
from stp3_b200.utils import geometry as G
from stp3_b200.utils import synthetic as syn
from tests.helpers import E2E_CASES, E2E_KEYS, e2e_errors, load_e2e_case, sha

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BAR_REF, BAR_ORACLE = 1e-3, 3e-4


@pytest.fixture(scope="module")
def model():
    import bench
    return bench.build_model(torch.device(DEV), syn.CONFIGS["perceive"])


@pytest.mark.parametrize("name", E2E_CASES)
def test_lift_splat_full_size_ranks_and_bev(name):
    cfg, inp, g = load_e2e_case(name)
    cam_M, cam_t, ego_R, ego_t = G.lift_matrices(inp["intrinsics"], inp["extrinsics"], inp["future_egomotion"])
    for a, k in ((cam_M, "cam_M"), (cam_t, "cam_t"), (ego_R, "ego_R"), (ego_t, "ego_t")):
        # the host matrices decide the ranks bit for bit: use the fixture's (they are equal on same-ISA hosts)
        assert np.allclose(a.numpy(), g[k], rtol=1e-6, atol=1e-7)
    mats = [torch.from_numpy(g[k]) for k in ("cam_M", "cam_t", "ego_R", "ego_t")]
    xs, ys, ds = G.frustum_axes(cfg.final_dim, cfg.downsample, cfg.d_bound)
    res, start, dim = G.calculate_birds_eye_view_parameters(cfg.x_bound, cfg.y_bound, cfg.z_bound)
    out, ranks = ops.lift_splat(inp["feat"].to(DEV), inp["depth_logits"].to(DEV), *mats, xs, ys, ds,
                                G.bev_offset(start, res), res, dim, cfg.discount, return_ranks=True)
    torch.cuda.synchronize()
    assert sha(ranks.cpu().numpy()) == str(g["rank_sha"]), "voxel ranks differ from the reference's"
    e_o, e_r = e2e_errors(g, "bev", out.cpu().numpy())
    assert e_o <= 1e-5 and e_r <= BAR_REF, (e_o, e_r)


@pytest.mark.parametrize("name", E2E_CASES)
def test_forward_features_full_size_vs_reference(model, name):
    cfg, inp, g = load_e2e_case(name)
    with torch.no_grad():
        out = model.forward_features(inp["feat"].to(DEV), inp["depth_logits"].to(DEV), inp["intrinsics"],
                                     inp["extrinsics"], inp["future_egomotion"])
        torch.cuda.synchronize()
    for k in E2E_KEYS:
        e_o, e_r = e2e_errors(g, k, out[k].cpu().numpy())
        print(f"{name} {k}: {e_o:.2e} of max vs fp64 oracle, {e_r:.2e} vs reference fp32")
        assert e_o <= BAR_ORACLE and e_r <= BAR_REF, (k, e_o, e_r)


def test_temporal_model_full_size_states(model):
    """The temporal model's own output (the decoder's input) at 200x200 against the reference's."""
    from stp3_b200 import dense
    cfg, inp, g = load_e2e_case("level")
    with torch.no_grad():
        h = {k: v.to(DEV) for k, v in model.prepare_inputs(inp["intrinsics"], inp["extrinsics"], inp["future_egomotion"]).items()}
        X, Y = model.bev_size
        planes = torch.empty((2, 1, 3, X, Y, 64), dtype=torch.bfloat16, device=DEV)
        off, res, dim = model._bev_host()
        r = ops.lift_splat(inp["feat"].to(DEV), inp["depth_logits"].to(DEV), h["cam_M"], h["cam_t"], h["ego_R"], h["ego_t"],
                           *model._axes(), off, res, dim, float(model.discount), workspace=model._ws, out_hilo=planes,
                           pool_sum=True)
        states = model.temporal_model.forward_hl(dense.HL(planes[0], planes[1], 64), const=h["const"], sums=r[1].view(3, 64))
        y = dense.to_f32(states, 0, 64)
        torch.cuda.synchronize()
    e_o, e_r = e2e_errors(g, "states", y.cpu().numpy())
    assert e_o <= BAR_ORACLE and e_r <= BAR_REF, (e_o, e_r)


def test_graphed_batch4_full_size(model):
    """GraphedPerception at the benched batch (4 samples/GPU): sample 0 is the fixture's sample, checked against the
    reference; the other samples must equal their own batch-1 forward (batch independence at full size)."""
    from stp3_b200.models.stp3 import GraphedPerception
    cfg, inp0, g = load_e2e_case("level")
    batch = syn.stack_samples(cfg, [int(g["seed"]), 1, 2, 3])
    assert torch.equal(batch["feat"][:1], inp0["feat"])
    with torch.no_grad():
        gp = GraphedPerception(model, 4, cfg.n_cameras, DEV)
        out = gp(batch["feat"].to(DEV), batch["depth_logits"].to(DEV), batch["intrinsics"], batch["extrinsics"],
                 batch["future_egomotion"])
        torch.cuda.synchronize()
        out = {k: out[k].clone() for k in E2E_KEYS}
        for k in E2E_KEYS:
            e_o, e_r = e2e_errors(g, k, out[k][:1].cpu().numpy())
            assert e_o <= BAR_ORACLE and e_r <= BAR_REF, (k, e_o, e_r)
        i = 2
        one = model.forward_features(batch["feat"][i:i + 1].to(DEV), batch["depth_logits"][i:i + 1].to(DEV),
                                     batch["intrinsics"][i:i + 1], batch["extrinsics"][i:i + 1],
                                     batch["future_egomotion"][i:i + 1])
        for k in E2E_KEYS:
            assert (out[k][i:i + 1] - one[k]).abs().max() <= 1e-4 * one[k].abs().max(), k
