"""CPU: pins oracle/lift_splat_oracle.py to the outputs of the reference itself (tests/golden/*.npz, made by
oracle/make_golden.py from the unmodified reference).  Voxel ranks / geometry: bitwise.  BEV: against the
reference's fp32 cumsum-trick output within that trick's own error (SURVEY.md headline fact 6)."""
import numpy as np
import pytest

from oracle import lift_splat_oracle as O
from tests.helpers import LIFT_CASES, load_lift_case, sha


def run_oracle(cfg, inp, g):
    return O.lift_splat(inp["feat"].numpy(), inp["depth_logits"].numpy(), g["cam_M"], g["cam_t"], g["ego_R"],
                        g["ego_t"], g["xs"], g["ys"], g["ds"], g["bev_offset"], g["bev_resolution"],
                        g["bev_dimension"], cfg.discount)


@pytest.mark.parametrize("name", LIFT_CASES)
def test_oracle_matches_reference(name):
    cfg, inp, g = load_lift_case(name)
    out = run_oracle(cfg, inp, g)
    # voxel ranks (and therefore the mask) are bit-exact with the reference
    assert sha(out["rank"]) == str(g["rank_sha"])
    assert sha(out["geom"]) == str(g["geom_sha"])
    assert int((out["rank"] >= 0).sum()) == int(g["n_kept"])
    bev = out["bev"]
    scale = np.abs(bev).max()
    if "bev" in g:
        ref = g["bev"].astype(np.float64)
        assert np.array_equal(out["rank"], g["rank"])
        err = np.abs(bev - ref).max()
    else:
        ref = g["bev_sample_value"].astype(np.float64)
        err = np.abs(bev.reshape(-1)[g["bev_sample_index"]] - ref).max()
    # reference fp32 prefix-sum error: ~eps * (running prefix) ; 3e-4 of the max is ample, a wrong pillar is O(1)
    assert err <= 3e-4 * scale, (err, scale)
    tc = bev.sum(axis=(-1, -2))
    assert np.allclose(tc, g["bev_tc_sum"], rtol=2e-4, atol=1e-3 * scale)


def test_host_parameters_restated():
    """bev params / offset restatements agree with the values the reference produced."""
    for name in ("plumbing", "carla_res", "perceive"):
        cfg, _, g = load_lift_case(name)
        res, start, dim = O.bev_params(cfg.x_bound, cfg.y_bound, cfg.z_bound)
        assert np.array_equal(res, g["bev_resolution"]) and np.array_equal(start, g["bev_start_position"])
        assert np.array_equal(dim, g["bev_dimension"])
        assert np.array_equal(O.bev_offset(res, start), g["bev_offset"])


def test_host_matrices_same_machine():
    """cam_M / pose matrices: product host code and oracle make the same torch calls -> identical bits on the
    machine they run on; against the build container's fixture they agree to 1 ulp-level tolerance
    (LAPACK / vectorised sin-cos are ISA dependent)."""
    from stp3_b200.utils import geometry as G
    import torch
    cfg, inp, g = load_lift_case("tiny_randpose")
    a = G.lift_matrices(inp["intrinsics"], inp["extrinsics"], inp["future_egomotion"])
    b = O.host_matrices(g["intrinsics"], g["extrinsics"], g["future_egomotion"])
    for x, y, key in zip(a, b, ("cam_M", "cam_t", "ego_R", "ego_t")):
        assert np.array_equal(x.numpy(), y)
        assert np.allclose(y, g[key], rtol=1e-6, atol=1e-7)
    xs, ys, ds = G.frustum_axes(cfg.final_dim, cfg.downsample, cfg.d_bound)
    assert np.allclose(xs.numpy(), g["xs"], rtol=1e-6) and np.array_equal(ds.numpy(), g["ds"])
