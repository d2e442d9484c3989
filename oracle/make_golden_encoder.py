"""Generates tests/golden/dense_encoder_*.npz from the UNMODIFIED reference (oracle/ref_loader.py), CPU, build container:

  dense_encoder_head.npz   DeepLabHead(160 -> 160, hidden 64) + UpsamplingConcat(216 -> 48): the depth head
                           (encoder.py:31-35) on synthetic r4 / r3 endpoints
  dense_encoder_full.npz   the reference's Encoder.get_features_depth itself (encoder.py:57-97: endpoint bookkeeping,
                           feature head -> 64 channels, depth head -> 48 bins) run on a stub trunk
                           (oracle/stub_trunk.py; the real EfficientNet is third-party and absent) for two 224x480 images

Test infrastructure; run by hand:  python -m oracle.make_golden_encoder"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402
from oracle.make_golden_dense import dense_input  # noqa: E402
from oracle.stub_trunk import StubEfficientNetB4  # noqa: E402
from oracle import torch_dense as TD  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ref = load_reference()
    C = ref.convolutions
    with torch.no_grad():
        h = TD.init_exact(C.DeepLabHead(160, 160, hidden_channel=64), seed=11).eval()
        u = TD.init_exact(C.UpsamplingConcat(216, 48), seed=12).eval()
        r_hi, r_lo = dense_input((2, 160, 14, 30), 21), dense_input((2, 56, 28, 60), 22)
        y = u(h(r_hi), r_lo)
        np.savez_compressed(os.path.join(OUT, "dense_encoder_head.npz"), out=y.numpy())
        print("depth head", tuple(y.shape), float(y.abs().max()))

        # the reference Encoder around a stub trunk: __init__ wants the EfficientNet download, so the instance is
        # assembled by hand with exactly the attributes its methods read (encoder.py:12-36)
        enc = ref.encoder.Encoder.__new__(ref.encoder.Encoder)
        torch.nn.Module.__init__(enc)
        enc.D, enc.C, enc.use_depth_distribution, enc.downsample, enc.version = 48, 64, True, 8, 'b4'
        enc.backbone = StubEfficientNetB4()
        enc.delete_unused_layers()
        enc.depth_layer_1 = C.DeepLabHead(160, 160, hidden_channel=64)
        enc.depth_layer_2 = C.UpsamplingConcat(216, 48)
        enc.feature_layer_1 = C.DeepLabHead(160, 160, hidden_channel=64)
        enc.feature_layer_2 = C.UpsamplingConcat(216, 64)
        TD.init_exact(enc, seed=31).eval()
        img = dense_input((2, 3, 224, 480), 32)
        feat, depth = enc(img)
        np.savez_compressed(os.path.join(OUT, "dense_encoder_full.npz"), feature=feat.numpy(), depth=depth.numpy(),
                            seed=31, in_seed=32)
        print("encoder", tuple(feat.shape), tuple(depth.shape), float(feat.abs().max()), float(depth.abs().max()))


if __name__ == "__main__":
    main()
