"""GPU parity of the drop-in dense modules (TemporalBlock, DeepLabHead, TemporalModel, Decoder, UpsamplingAdd and the
memory-bound helpers) against the plain-PyTorch oracle evaluated in float64 on the CPU with the same weights, and
against the golden outputs of the unmodified reference.  Bar: head logits / features within 1e-3 relative (north_star);
measured ~1e-5."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import torch_dense as TD
from oracle.make_golden_dense import GATES_ALL, GATES_PERCEIVE, dense_input
from stp3_b200 import dense
from stp3_b200.layers.convolutions import DeepLabHead, UpsamplingAdd
from stp3_b200.layers.temporal import TemporalBlock
from stp3_b200.models.decoder import Decoder
from stp3_b200.models.temporal_model import TemporalModel
from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REL = 1e-3


def close(y, ref, rel=REL):
    ref = ref.double()
    err = (y.double().cpu() - ref).abs().max().item()
    scale = max(ref.abs().max().item(), 1e-6)
    assert err <= rel * scale, (err, scale, err / scale)
    return err / scale


def f64(m):
    import copy
    return copy.deepcopy(m).double()


def test_layout_roundtrip_and_spatial_sum():
    x = dense_input((2, 3, 70, 13, 21), 1).to(DEV)
    h = dense.from_f32(x)
    assert h.hi.shape == (2, 3, 13, 21, 128) and float(h.hi[..., 70:].float().abs().max()) == 0.0
    y = dense.to_f32(h, 0, 70)
    assert (y - x).abs().max() <= 2e-5 * x.abs().max()            # hi+lo keeps 16 mantissa bits
    s = dense.spatial_sum(h)
    ref = x.double().sum(dim=(-1, -2)).view(6, 70)
    assert torch.allclose(s[:, :70].double(), ref, rtol=1e-4, atol=1e-3)
    hcl = dense.from_f32(x.permute(0, 1, 3, 4, 2).contiguous(), channels_last=True)
    assert torch.equal(hcl.hi, h.hi) and torch.equal(hcl.lo, h.lo)


def test_upsampling_add():
    with torch.no_grad():
        up = TD.init_exact(UpsamplingAdd(128, 64), seed=3).eval()
        x, skip = dense_input((3, 128, 9, 11), 2), dense_input((3, 64, 18, 22), 3)
        ref = TD.upsampling_add(x.double(), skip.double(), f64(up))
        y = up.to(DEV)(x.to(DEV), skip.to(DEV))
    close(y, ref)


@pytest.mark.parametrize("cin,cout", [
    (70, 64), (64, 64),       # the two blocks of the perceive configuration
    (134, 64),                # stress configuration (C = 128 + 6 ego-motion channels): 67-channel paths, 216-channel concat
    (20, 24),                 # narrow: both mid paths share one 64-channel block
    (128, 128),               # identity skip with 64-channel paths
    (200, 96),                # 100-channel paths
])
def test_temporal_block(cin, cout):
    H, W = 20, 28
    with torch.no_grad():
        blk = TD.init_exact(TemporalBlock(cin, cout, use_pyramid_pooling=True, pool_sizes=[(2, H, W)]), seed=4).eval()
        x = dense_input((2, cin, 3, H, W), 7)
        ref = TD.temporal_block(x.double(), f64(blk))
        y = blk.to(DEV)(x.to(DEV))
    close(y, ref)


def test_deeplab_head():
    with torch.no_grad():
        head = TD.init_exact(DeepLabHead(64, 64, hidden_channel=128), seed=5).eval()
        x = dense_input((3, 64, 30, 44), 8)
        ref = TD.deeplab_head(x.double(), f64(head))
        y = head.to(DEV)(x.to(DEV))
    close(y, ref)


def _temporal_golden():
    g = dict(np.load(os.path.join(GOLDEN, "dense_temporal_model.npz")))
    x = dense_input((1, 3, 70, int(g["H"]), int(g["W"])), int(g["in_seed"]))
    x[:, :, 64:] = x[:, :, 64:, :1, :1]
    return g, x


def test_temporal_model_vs_reference_golden_and_fp64():
    g, x = _temporal_golden()
    with torch.no_grad():
        tm = TD.init_exact(TemporalModel(70, 3, (int(g["H"]), int(g["W"])), start_out_channels=64), seed=int(g["seed"])).eval()
        ref64 = TD.temporal_model(x.double(), f64(tm))
        y = tm.to(DEV)(x.to(DEV))
    close(y, ref64)
    close(y, torch.from_numpy(g["out"]))                  # the reference's own fp32 output


def test_temporal_model_stress_shape_vs_reference_golden():
    """BASELINE configs[4] shape (134 input channels, receptive field 5) against the reference's own fp32 output."""
    g = dict(np.load(os.path.join(GOLDEN, "dense_temporal_model_stress.npz")))
    H, W = int(g["H"]), int(g["W"])
    x = dense_input((1, 5, 134, H, W), int(g["in_seed"]))
    x[:, :, 128:] = x[:, :, 128:, :1, :1]
    with torch.no_grad():
        tm = TD.init_exact(TemporalModel(134, 5, (H, W), start_out_channels=64), seed=int(g["seed"])).eval()
        y = tm.to(DEV)(x.to(DEV))
    close(y, torch.from_numpy(g["out"]))


def test_temporal_model_constant_channels_as_bias():
    """The fused path never materialises the 6 broadcast ego-motion channels (stp3.py:145-152)."""
    g, x = _temporal_golden()
    with torch.no_grad():
        tm = TD.init_exact(TemporalModel(70, 3, (int(g["H"]), int(g["W"])), start_out_channels=64), seed=int(g["seed"])).eval().to(DEV)
        tm.model[0].n_const = 6
        const = x[:, :, 64:, 0, 0].reshape(3, 6).contiguous().to(DEV)
        y = dense.to_f32(tm.forward_hl(dense.from_f32(x[:, :, :64].to(DEV)), const=const), 0, 64)
    close(y, torch.from_numpy(g["out"]))


@pytest.mark.parametrize("name,gates", [("perceive", GATES_PERCEIVE), ("all", GATES_ALL)])
def test_decoder_vs_reference_golden_and_fp64(name, gates):
    g = dict(np.load(os.path.join(GOLDEN, f"dense_decoder_{name}.npz")))
    x = dense_input((1, 3, 64, int(g["H"]), int(g["W"])), int(g["in_seed"]))
    with torch.no_grad():
        dec = TD.init_exact(Decoder(64, 2, 3, 2, gates), seed=int(g["seed"])).eval()
        ref64 = TD.decoder(x.double(), f64(dec))
        out = dec.to(DEV)(x.to(DEV))
    for k, v in ref64.items():
        if v is None:
            assert out[k] is None
            continue
        assert out[k].shape == v.shape, (k, out[k].shape, v.shape)
        close(out[k], v)
        close(out[k], torch.from_numpy(g[k]))


def test_decoder_batch2_full_resolution_shapes():
    with torch.no_grad():
        dec = TD.init_exact(Decoder(64, 2, 3, 2, GATES_PERCEIVE), seed=9).eval().to(DEV)
        out = dec(dense_input((2, 3, 64, 200, 200), 10).to(DEV))
    assert out["segmentation"].shape == (2, 3, 2, 200, 200) and out["hdmap"].shape == (2, 4, 200, 200)
    assert all(torch.isfinite(v).all() for v in out.values() if v is not None)


def test_training_mode_is_refused():
    blk = TemporalBlock(64, 64).to(DEV)
    with pytest.raises(NotImplementedError):
        blk(torch.zeros(1, 64, 2, 8, 8, device=DEV))


def test_encoder_heads_vs_reference_golden():
    """DeepLabHead(160->160, hidden 64) + UpsamplingConcat(216->48): the depth head of the image encoder
    (encoder.py:31-35, 88-95) against the output of the reference's own modules with the same weights."""
    from stp3_b200.layers.convolutions import UpsamplingConcat
    g = dict(np.load(os.path.join(GOLDEN, "dense_encoder_head.npz")))
    with torch.no_grad():
        h = TD.init_exact(DeepLabHead(160, 160, hidden_channel=64), seed=11).eval().to(DEV)
        u = TD.init_exact(UpsamplingConcat(216, 48), seed=12).eval().to(DEV)
        r_hi, r_lo = dense_input((2, 160, 14, 30), 21).to(DEV), dense_input((2, 56, 28, 60), 22).to(DEV)
        y = u(h(r_hi), r_lo)
    close(y, torch.from_numpy(g["out"]))


def test_temporal_model_receptive_field_5():
    """Stress configuration: S=5 -> four TemporalBlocks (temporal_model.py:13)."""
    H, W = 16, 24
    with torch.no_grad():
        tm = TD.init_exact(TemporalModel(70, 5, (H, W), start_out_channels=64), seed=13).eval()
        x = dense_input((1, 5, 70, H, W), 14)
        ref = TD.temporal_model(x.double(), f64(tm))
        y = tm.to(DEV)(x.to(DEV))
    close(y, ref)


def test_encoder_forward_vs_reference_golden():
    """The drop-in Encoder (trunk endpoint bookkeeping + feature head -> 64 channels + depth head -> 48 bins) against
    the reference's own Encoder.get_features_depth on the same stub trunk and weights (oracle/make_golden_encoder.py):
    the reference layout, and the channels-last hand-off to the lift-splat."""
    from oracle.stub_trunk import StubEfficientNetB4
    from stp3_b200.config import get_cfg
    from stp3_b200.models.encoder import Encoder
    g = dict(np.load(os.path.join(GOLDEN, "dense_encoder_full.npz")))
    with torch.no_grad():
        enc = Encoder(get_cfg().MODEL.ENCODER, D=48, backbone=StubEfficientNetB4())
        enc = TD.init_exact(enc, seed=int(g["seed"])).eval()
        img = dense_input((2, 3, 224, 480), int(g["in_seed"]))
        # the (third-party) trunk runs where the reference ran it -- on the CPU, in fp32 (cuDNN would use TF32) -- so both
        # sides see bit-identical endpoints; the drop-in's endpoint bookkeeping (encoder.py:57-86) is what is exercised
        r_lo, r_hi = enc.trunk(img)
        assert r_lo.shape == (2, 56, 28, 60) and r_hi.shape == (2, 160, 14, 30)
        enc = enc.to(DEV)
        r_lo, r_hi = r_lo.to(DEV), r_hi.to(DEV)
        feat, depth = enc.heads_f32(r_lo, r_hi)
        close(feat, torch.from_numpy(g["feature"]))
        close(depth, torch.from_numpy(g["depth"]))
        feat_cl, depth2 = enc.heads_f32(r_lo, r_hi, channels_last=True)
        assert feat_cl.shape == (2, 28, 60, 64)
        close(feat_cl.permute(0, 3, 1, 2), torch.from_numpy(g["feature"]))
        # two runs differ by the summation order of the global-pool branch's atomics (stp3_spatial_sum), ~1e-5
        assert (feat_cl.permute(0, 3, 1, 2) - feat).abs().max() <= 1e-4 * feat.abs().max()
        assert (depth2 - depth).abs().max() <= 1e-4 * depth.abs().max()
        # the path through hi/lo planes + layout conversion gives the same values
        f_hl, d_hl = enc.heads_hl(r_lo, r_hi)
        assert (dense.to_f32(f_hl, 0, 64).squeeze(1) - feat).abs().max() <= 1e-4 * feat.abs().max()
        # Encoder.forward end to end on the device (trunk included) stays within TF32 noise of the same values
        f2, d2 = enc(img.to(DEV))
        assert (f2 - feat).abs().max() <= 2e-2 * feat.abs().max()


@pytest.mark.parametrize("B,H,W", [(1, 40, 40), (2, 30, 44), (1, 200, 200), (3, 100, 100)])
def test_aspp_fused_kernel_vs_oracle(B, H, W, monkeypatch):
    """DeepLabHead(64 -> 64, hidden 128) with the ASPP branches + projection as ONE back-to-back kernel
    (stp3_aspp_fused_fwd) against the fp64 oracle and against the unfused launch sequence: dilation 12 / 24 / 36 taps live
    at 200x200, skipped padding taps and ragged edge tiles at the small sizes."""
    monkeypatch.setenv("STP3_ASPP_FUSED", "1")
    with torch.no_grad():
        head = TD.init_exact(DeepLabHead(64, 64, hidden_channel=128), seed=17).eval()
        x = dense_input((B, 64, H, W), 18)
        ref = TD.deeplab_head(x.double(), f64(head))
        y = head.to(DEV)(x.to(DEV))
        assert "fused" in head.packed()
        err = close(y, ref)
        monkeypatch.setenv("STP3_ASPP_FUSED", "0")
        head2 = TD.init_exact(DeepLabHead(64, 64, hidden_channel=128), seed=17).eval().to(DEV)
        y2 = head2(x.to(DEV))
        assert "fused" not in head2.packed()
        assert (y - y2).abs().max() <= 2e-5 * y2.abs().max(), "fused and unfused ASPP disagree"
    print(f"aspp fused {B}x{H}x{W}: {err:.2e} of max vs fp64 oracle")


@pytest.mark.parametrize("B,H,W", [(1, 24, 40), (2, 50, 36), (1, 200, 200)])
def test_temporal_block_fused_tail_vs_oracle(B, H, W, monkeypatch):
    """TemporalModel(70 -> 64, S = 3) with each block's tail (spatio-temporal convolutions of both paths, path 2, aggregation,
    pyramid-pooling bias, projection / identity residual, output sums) in the back-to-back kernel stp3_block_fused_fwd:
    block 1 = three chains of 35 channels + projection + ego-motion channels as per-image biases, block 2 = block-diagonal
    chain + identity residual.  Against the fp64 oracle on the materialised 70-channel input and against the unfused
    launch sequence, incl. ragged tiles and the frame-0 causal padding."""
    x = dense_input((B, 3, 70, H, W), 41)
    x[:, :, 64:] = x[:, :, 64:, :1, :1]                        # the six ego-motion channels are spatially constant
    outs = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("STP3_BLOCK_FUSED", fused)
        with torch.no_grad():
            tm = TD.init_exact(TemporalModel(70, 3, (H, W), start_out_channels=64), seed=42).eval()
            if fused == "1":
                ref = TD.temporal_model(x.double(), f64(tm))
            tm = tm.to(DEV)
            tm.model[0].n_const = 6
            const = x[:, :, 64:, 0, 0].reshape(B * 3, 6).contiguous().to(DEV)
            y = dense.to_f32(tm.forward_hl(dense.from_f32(x[:, :, :64].to(DEV)), const=const), 0, 64)
            assert ("tail" in tm.model[0].packed()) == (fused == "1") and ("tail" in tm.model[1].packed()) == (fused == "1")
        outs[fused] = y
        err = close(y, ref)
        print(f"temporal model {B}x{H}x{W} fused={fused}: {err:.2e} of max vs fp64 oracle")
    assert (outs["1"] - outs["0"]).abs().max() <= 5e-5 * outs["0"].abs().max(), "fused and unfused block tails disagree"
