"""Shared helpers for the parity tests (test infrastructure)."""
import hashlib
import os

import numpy as np

from stp3_b200.utils import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LIFT_CASES = ["tiny_randpose", "tiny_level", "plumbing", "carla_res", "lift_splat", "perceive", "stress"]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_lift_case(name):
    """Fixture + regenerated (machine-independent) inputs; asserts the inputs hash to the recorded value."""
    g = dict(np.load(os.path.join(GOLDEN, f"lift_splat_{name}.npz"), allow_pickle=False))
    cfg = syn.CONFIGS[str(g["config"])]
    inp = syn.lift_inputs(cfg, int(g["batch"]), seed=int(g["seed"]), random_pose=bool(g["random_pose"]))
    assert sha(inp["feat"].numpy()) == str(g["feat_sha"]), "synthetic generator is not reproducing the fixture inputs"
    assert sha(inp["depth_logits"].numpy()) == str(g["depth_sha"])
    return cfg, inp, g
