#!/usr/bin/env python
"""The unfused resample/gather path on device-resident buffers (quantise -> tile scan -> pull -> gather) at 2^24
particles: the case behind bench.py's `roofline_resample_gather`; run under ncu for per-kernel times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import stochy_jl_b200 as sb
N = 1 << 24
g = torch.Generator(device="cuda"); g.manual_seed(3)
lw = torch.randn(N, device="cuda", dtype=torch.float32, generator=g) * 2.0
src = torch.randint(0, 1 << 30, (N,), device="cuda", dtype=torch.int32, generator=g)
anc = torch.empty(N + 2048, device="cuda", dtype=torch.int32)
dst = torch.empty(N, device="cuda", dtype=torch.int32)
torch.cuda.synchronize()
with sb.Engine(precision="fp32") as eng:
    for i in range(6):
        eng.dev_resample_systematic(lw.data_ptr(), sb._native.W_LOG_F32, N, 0.1 + 0.01 * i, N, anc.data_ptr())
        eng.dev_gather(src.data_ptr(), 4, N, anc.data_ptr(), N, dst.data_ptr())
    eng.sync()
print("ok", bool(torch.equal(dst, src[anc[:N].long()])))
