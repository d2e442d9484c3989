"""Scratch probe (GPU): does writing a 128-channel window of a 512-channel concat tensor cost more than a dense tensor?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stp3_b200 import dense

dev = "cuda:0"
B, T, H, W = 4, 3, 200, 200
x = dense.HL.zeros(B, T, H, W, 64, dev)
x.hi.normal_(); x.lo.normal_(std=0.01)
w = torch.randn(128, 64, 1, 1, device=dev) * 0.1
pc1 = dense.pack_conv(w, torch.zeros(128, device=dev))
w3 = torch.randn(128, 64, 3, 3, device=dev) * 0.05
pcd = dense.pack_conv(w3, torch.zeros(128, device=dev), dilation=12)
cat = dense.HL.empty(B, T, H, W, 512, dev, cp=512)
dn = dense.HL.empty(B, T, H, W, 128, dev, cp=128)

def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n * 1e3

print("1x1 64->128 into dense   :", t(lambda: dense.conv(x, pc1, out=dn, relu=True)), "us")
print("1x1 64->128 into CAT[0]  :", t(lambda: dense.conv(x, pc1, out=cat, out_coff=0, relu=True)), "us")
print("1x1 64->128 into CAT[384]:", t(lambda: dense.conv(x, pc1, out=cat, out_coff=384, relu=True)), "us")
print("3x3 d12 64->128 dense    :", t(lambda: dense.conv(x, pcd, out=dn, relu=True)), "us")
print("3x3 d12 64->128 CAT[128] :", t(lambda: dense.conv(x, pcd, out=cat, out_coff=128, relu=True)), "us")
wp = torch.randn(128, 512, 1, 1, device=dev) * 0.05
pcp = dense.pack_conv(wp, torch.zeros(128, device=dev))
print("1x1 512->128 from CAT    :", t(lambda: dense.conv(cat, pcp, out=dn, relu=True)), "us")
