"""CPU: pins the backward restatement (oracle/lift_splat_oracle.py::lift_splat_backward) to the gradients the
UNMODIFIED reference's own autograd produced (tests/golden/lift_splat_bwd_*.npz, oracle/make_golden_bwd.py)."""
import os

import numpy as np
import pytest

from oracle import lift_splat_oracle as O
from oracle.make_golden_bwd import loss_weights
from stp3_b200.utils import geometry as G
from stp3_b200.utils import synthetic as syn
from tests.helpers import GOLDEN


def bwd_case(name):
    g = dict(np.load(os.path.join(GOLDEN, f"lift_splat_bwd_{name}.npz"), allow_pickle=False))
    fwd = dict(np.load(os.path.join(GOLDEN, f"lift_splat_{name}.npz"), allow_pickle=False))      # host matrices of the case
    cfg = syn.CONFIGS[str(g["config"])]
    inp = syn.lift_inputs(cfg, int(g["batch"]), seed=int(g["seed"]), random_pose=bool(g["random_pose"]))
    return cfg, inp, g, fwd


def check(g, key, got, rel):
    if key in g:
        ref = g[key].astype(np.float64)
        err = np.abs(got - ref).max() / np.abs(ref).max()
    else:
        ref = g[key + "_value"].astype(np.float64)
        err = np.abs(got.reshape(-1)[g[key + "_index"]] - ref).max() / float(g[key + "_max"])
    assert err <= rel, (key, err)
    return err


@pytest.mark.parametrize("name", ["tiny_randpose", "tiny_level", "carla_res"])
def test_backward_oracle_matches_reference_autograd(name):
    cfg, inp, g, fwd = bwd_case(name)
    xs, ys, ds = fwd["xs"], fwd["ys"], fwd["ds"]
    ora = O.lift_splat(inp["feat"].numpy(), inp["depth_logits"].numpy(), fwd["cam_M"], fwd["cam_t"], fwd["ego_R"], fwd["ego_t"],
                       xs, ys, ds, fwd["bev_offset"], fwd["bev_resolution"], fwd["bev_dimension"], cfg.discount)
    W = loss_weights(ora["bev"].shape, int(g["w_seed"])).numpy()
    gf, gd = O.lift_splat_backward(W, inp["feat"].numpy(), inp["depth_logits"].numpy(), ora["rank"], cfg.discount)
    check(g, "grad_feat", gf, 2e-5)          # the reference's gradients are fp32
    check(g, "grad_depth", gd, 2e-5)
