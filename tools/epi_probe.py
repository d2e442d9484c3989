"""Scratch probe (GPU): epilogue-bound 1x1 layers -- plain / per-image bias / concat window / residual."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stp3_b200 import dense

dev = "cuda:0"
B, T, H, W = 4, 3, 200, 200
quick = len(sys.argv) > 1 and sys.argv[1] == "ncu"


def t(fn, n=10):
    if quick:
        fn(); torch.cuda.synchronize(); return 0.0
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n * 1e3


x = dense.HL.zeros(B, T, H, W, 64, dev); x.hi.normal_(); x.lo.normal_(std=0.01)
for cout in (64, 128):
    w = torch.randn(cout, 64, 1, 1, device=dev) * 0.1
    pc = dense.pack_conv(w, torch.zeros(cout, device=dev))
    dn = dense.HL.empty(B, T, H, W, cout, dev, cp=cout)
    cat = dense.HL.empty(B, T, H, W, 512, dev, cp=512)
    res = dense.HL.zeros(B, T, H, W, cout, dev); res.hi.normal_()
    ib = torch.randn(B * T, cout, device=dev)
    for tune in ((1, 1), (2, 1), (3, 1)):
        print(f"bn{cout} {tune} plain   : {t(lambda: dense.conv(x, pc, out=dn, relu=True, tune=tune)):7.1f} us", flush=True)
        print(f"bn{cout} {tune} img_bias: {t(lambda: dense.conv(x, pc, out=dn, relu=True, img_bias=ib, tune=tune)):7.1f} us", flush=True)
        print(f"bn{cout} {tune} into CAT: {t(lambda: dense.conv(x, pc, out=cat, out_coff=128, relu=True, tune=tune)):7.1f} us", flush=True)
        print(f"bn{cout} {tune} residual: {t(lambda: dense.conv(x, pc, out=dn, relu=True, residual=res, res_after_act=True, tune=tune)):7.1f} us", flush=True)
        if quick: break
