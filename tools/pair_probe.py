"""Scratch probe (GPU): single-CTA tilings vs the CTA-pair tiling (tcgen05.mma.cta_group::2) on the hot-path layer shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stp3_b200 import dense

dev = "cuda:0"
B, T, H, W = 4, 3, 200, 200


def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n * 1e3


def run(name, cin, cout, k, dil=1, stride=1, hw=(H, W)):
    x = dense.HL.zeros(B, T, hw[0], hw[1], cin, dev)
    x.hi.normal_(); x.lo.normal_(std=0.01)
    w = torch.randn(cout, cin, k, k, device=dev) * 0.05
    pc = dense.pack_conv(w, torch.zeros(cout, device=dev), dilation=dil, stride=stride)
    groupable = k == 3 and dil == 1 and stride == 1
    cfgs = [(2, 1), (3, 1)] + ([(2, 3), (3, 3)] if groupable else [])
    if cout <= 64:
        cfgs += [(ns, g + 8) for ns, g in cfgs]
    out = "  ".join(f"{c}: {t(lambda: dense.conv(x, pc, relu=True, tune=c)):7.1f}us" for c in cfgs)
    print(f"{name:28s} {out}", flush=True)


run("1x1 64->64", 64, 64, 1)
run("1x1 64->128", 64, 128, 1)
run("1x1 512->128", 512, 128, 1)
run("3x3 64->64", 64, 64, 3)
run("3x3 d12 64->128", 64, 128, 3, dil=12)
run("3x3 128->128", 128, 128, 3)
run("7x7 s2 64->64", 64, 64, 7, stride=2)
run("3x3 128->128 @50", 128, 128, 3, hw=(50, 50))
run("3x3 256->256 @25", 256, 256, 3, hw=(25, 25))
