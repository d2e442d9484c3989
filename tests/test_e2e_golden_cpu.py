"""CPU: pins the composed oracle (numpy fp64 lift-splat -> torch fp64 TemporalModel -> Decoder) at the HEADLINE size
against tests/golden/e2e_perceive_level.npz, i.e. against what the unmodified reference produced end to end on the
same single perceive-config sample (200x200 BEV, dilations 12/24/36 live)."""
import copy

import numpy as np
import torch

from oracle import lift_splat_oracle as O
from oracle import torch_dense as TD
from stp3_b200.utils import geometry as G
from tests.helpers import E2E_KEYS, e2e_errors, load_e2e_case, sha


def test_oracle_end_to_end_matches_reference_at_full_size():
    import bench
    cfg, inp, g = load_e2e_case("level")
    mats = [g[k] for k in ("cam_M", "cam_t", "ego_R", "ego_t")]
    xs, ys, ds = G.frustum_axes(cfg.final_dim, cfg.downsample, cfg.d_bound)
    res, start, dim = G.calculate_birds_eye_view_parameters(cfg.x_bound, cfg.y_bound, cfg.z_bound)
    ora = O.lift_splat(inp["feat"].numpy(), inp["depth_logits"].numpy(), *mats, xs.numpy(), ys.numpy(), ds.numpy(),
                       G.bev_offset(start, res).numpy(), res.numpy(), dim.numpy(), cfg.discount)
    assert sha(ora["rank"].astype(np.int32).reshape(inp["depth_logits"].shape)) == str(g["rank_sha"])
    e_o, e_r = e2e_errors(g, "bev", ora["bev"])
    assert e_o <= 1e-12 and e_r <= 2e-4, (e_o, e_r)            # the reference's fp32 cumsum trick is ~4e-5 off
    with torch.no_grad():
        model = bench.build_model(lcfg=cfg).double()
        X, Y = cfg.bev_xy
        ego = inp["future_egomotion"].double()
        ego = torch.cat([torch.zeros_like(ego[:, :1]), ego[:, :-1]], 1)
        x = torch.cat([torch.from_numpy(ora["bev"]), ego.view(1, -1, 6, 1, 1).expand(1, ego.shape[1], 6, X, Y)], dim=2)
        states = TD.temporal_model(x, model.temporal_model)
        out = TD.decoder(states, model.decoder)
    e_o, e_r = e2e_errors(g, "states", states.numpy())
    assert e_o <= 1e-9 and e_r <= 5e-4, (e_o, e_r)
    for k in E2E_KEYS:
        e_o, e_r = e2e_errors(g, k, out[k].numpy())
        assert e_o <= 1e-9 and e_r <= 5e-4, (k, e_o, e_r)      # reference fp32 vs fp64: ~2e-4 of max on the logits
