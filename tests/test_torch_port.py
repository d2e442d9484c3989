"""CPU: the op-for-op torch port (bench.py's reference arm / cpu_baseline) reproduces the reference's own fp32
output stored in the golden fixtures."""
import numpy as np
import pytest
import torch

from oracle import torch_port as TP
from tests.helpers import load_lift_case


@pytest.mark.parametrize("name", ["tiny_randpose", "tiny_level", "plumbing"])
def test_port_matches_reference_output(name):
    cfg, inp, g = load_lift_case(name)
    t = torch.as_tensor
    out = TP.lift_splat(inp["feat"], inp["depth_logits"], t(g["intrinsics"]), t(g["extrinsics"]),
                        t(g["future_egomotion"]), t(g["xs"]), t(g["ys"]), t(g["ds"]), t(g["bev_resolution"]),
                        t(g["bev_start_position"]), g["bev_dimension"], cfg.discount).numpy()
    ref = g["bev"]
    # same ATen sequence; argsort is unstable so the cumsum order (and the last bits) may differ
    assert np.abs(out - ref).max() <= 2e-4 * np.abs(ref).max()
    occupied = np.abs(ref).sum(axis=2) > 0
    assert np.array_equal(np.abs(out).sum(axis=2) > 0, occupied)
