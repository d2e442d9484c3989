"""Host-side logic of the dense path checked on the CPU: weight packing / tap tables of stp3_b200.dense.pack_conv and the
channel-window bookkeeping of the temporal block.  The packed tensors are evaluated with a plain-torch emulation of the
kernel's implicit GEMM (sum over taps of shifted-window x weight-slice products) and compared with torch's own
convolution; no CUDA involved."""
import pytest
import torch
import torch.nn.functional as F

from stp3_b200 import dense
from stp3_b200.layers.temporal import TemporalBlock


def emulate(x, pc, T_causal=True):
    """x (B,T,C,H,W) fp64 -> (B,T,bn,Ho,Wo) fp64 with the packed weights: out[p] = sum_tap W_tap^T x[p*stride + d_tap]."""
    B, T, C, H, W = x.shape
    nt, kbs, two, bn, kb = pc.w.shape
    assert two == 2 and kb == 64 and nt == len(pc.taps) and kbs * 64 == pc.cin_p
    w = (pc.w[:, :, 0].double() + pc.w[:, :, 1].double())            # hi + lo, (nt, kbs, bn, 64)
    w = w.permute(0, 2, 1, 3).reshape(nt, bn, pc.cin_p)               # (nt, bn, cin_p)
    xp = torch.zeros(B, T, pc.cin_p, H, W, dtype=torch.float64)
    xp[:, :, :C] = x
    s = pc.stride
    Ho, Wo = (H + s - 1) // s, (W + s - 1) // s
    out = torch.zeros(B, T, bn, Ho, Wo, dtype=torch.float64)
    for i, (dt, dy, dx) in enumerate(pc.taps):
        for t in range(T):
            if not 0 <= t + dt < T:
                continue                                              # zero padding in time
            for oy in range(Ho):
                iy = oy * s + dy
                if not 0 <= iy < H:
                    continue
                for ox in range(Wo):
                    ix = ox * s + dx
                    if 0 <= ix < W:
                        out[:, t, :, oy, ox] += xp[:, t + dt, :, iy, ix] @ w[i].T
    return out + pc.bias.double().view(1, 1, bn, 1, 1)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


@pytest.mark.parametrize("cin,cout,k,stride,dil", [(5, 7, 3, 1, 1), (8, 4, 7, 2, 1), (6, 70, 3, 1, 2), (3, 3, 3, 2, 1),
                                                   (70, 35, 1, 1, 1)])
def test_pack_conv_2d_reproduces_conv2d(cin, cout, k, stride, dil):
    x = rnd(1, 2, cin, 9, 8, seed=1).double()
    w, b = rnd(cout, cin, k, k, seed=2, scale=0.3), rnd(cout, seed=3)
    pc = dense.pack_conv(w, b, stride=stride, dilation=dil)
    assert pc.bn in (64, 128, 256) and pc.cin_p % 64 == 0 and pc.cout == cout
    y = emulate(x, pc)
    ref = F.conv2d(x.view(2, cin, 9, 8), w.double(), b.double(), stride=stride, padding=(k - 1) * dil // 2, dilation=dil)
    assert torch.allclose(y[0, :, :cout], ref, atol=3e-5, rtol=0)               # weights carry 16 mantissa bits (hi + lo)
    assert float(y[0, :, cout:].abs().max()) == 0.0                             # padded output columns: zero weights and bias


def test_stride2_taps_are_grouped_by_row_parity():
    """Taps of one kernel column whose dy advance by the stride are consecutive (the kernel shares one activation load
    between them): 7x7 stride 2 -> dy = -3,-1,1,3 then -2,0,2 for every dx."""
    pc = dense.pack_conv(rnd(4, 4, 7, 7, seed=4), torch.zeros(4), stride=2)
    assert len(pc.taps) == 49
    for c in range(7):
        col = pc.taps[7 * c:7 * c + 7]
        assert [t[1] for t in col] == [-3, -1, 1, 3, -2, 0, 2] and len({t[2] for t in col}) == 1
    pc1 = dense.pack_conv(rnd(4, 4, 3, 3, seed=5), torch.zeros(4))
    assert [t[1] for t in pc1.taps[:3]] == [-1, 0, 1]                           # stride 1: consecutive dy


def test_causal_conv3d_taps_and_values():
    x = rnd(1, 3, 6, 6, 7, seed=6).double()
    w, b = rnd(5, 6, 2, 3, 3, seed=7, scale=0.3), rnd(5, seed=8)
    pc = dense.pack_conv(w, b)
    assert sorted({t[0] for t in pc.taps}) == [-1, 0]                           # time padded on the left only
    y = emulate(x, pc)
    xp = F.pad(x.permute(0, 2, 1, 3, 4), (1, 1, 1, 1, 1, 0))
    ref = F.conv3d(xp, w.double(), b.double()).permute(0, 2, 1, 3, 4)
    assert torch.allclose(y[:, :, :5], ref, atol=3e-5, rtol=0)


def test_in_layout_maps_logical_channels_to_physical_windows():
    """Aggregation conv of a temporal block: three 35-channel groups living at offsets 0 / 40 / 80 of a 128-channel tensor."""
    w, b = rnd(64, 105, 1, 1, seed=9, scale=0.3), rnd(64, seed=10)
    pc = dense.pack_conv(w, b, in_layout=[(0, 35, 0), (35, 35, 40), (70, 35, 80)], cin_p=128)
    x = torch.zeros(1, 1, 128, 3, 3, dtype=torch.float64)
    xl = rnd(1, 1, 105, 3, 3, seed=11).double()
    x[:, :, 0:35], x[:, :, 40:75], x[:, :, 80:115] = xl[:, :, :35], xl[:, :, 35:70], xl[:, :, 70:]
    y = emulate(x, pc)
    ref = F.conv2d(xl.view(1, 105, 3, 3), w.double(), b.double())
    assert torch.allclose(y[0, :, :64], ref, atol=3e-5, rtol=0)


@pytest.mark.parametrize("cin,cout,nc", [(70, 64, 6), (64, 64, 0), (134, 64, 6), (20, 24, 0), (200, 96, 0)])
def test_temporal_block_packing_layout(cin, cout, nc):
    """Channel windows of the packed temporal block for the reference's blocks, the stress block and odd widths."""
    with torch.no_grad():
        blk = TemporalBlock(cin, cout, use_pyramid_pooling=True, pool_sizes=[(2, 8, 8)]).eval()
        blk.n_const = nc
        P = blk._pack()
    half, o = cin // 2, P["o"]
    assert o % 8 == 0 and o >= half and P["ap"] % 64 == 0 and P["ap"] >= 3 * o
    assert P["agg"].cin_p == P["ap"] and P["b"].cin_p == P["hp"] == P["c"].cin_p
    assert P["a1"].bn == P["nmid"] and P["nmid"] in (64, 128, 256)
    assert (P["m1"] == half) == (2 * half <= 64)                                 # both mid paths share one K block iff they fit
    if nc:
        assert P["a1_c"].shape == (P["nmid"], nc) and P["a2_c"].shape == (P["a2"].bn, nc)
    merged = ("a12" in P) or ("a2p" in P)
    assert merged == (half <= 64 and ((blk.projection is None and P["nmid"] == 64) or
                                      (blk.projection is not None and cout <= 64)))
    for k in ("a12", "a2p"):
        if k in P:
            assert P[k].bn == 128 and P[k].cout == 128
