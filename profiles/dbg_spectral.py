"""Timing of the SpectralCF step on an ml-100k-sized problem (N = 2 625 nodes, d = 100, 2 layers, batch 256): the dense
operator is random (the kernel's cost does not depend on its values), CUDA events around 50 steps."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurec_b200 import ops  # noqa: E402

nu, ni, d, K, bs = 943, 1682, 100, 2, 256
N = nu + ni
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn((N, N), device="cuda", generator=g) / N
At = A.t().contiguous()
E = torch.randn((N, d), device="cuda", generator=g) * 0.1
W = torch.randn((K, d, d), device="cuda", generator=g) * 0.1
all_emb = torch.zeros((N, d * (K + 1)), device="cuda"); G = torch.zeros_like(all_emb)
touched = torch.zeros(N, dtype=torch.int32, device="cuda")
gE, gW = torch.zeros_like(E), torch.zeros_like(W)
sE, sW = (torch.zeros_like(E), torch.zeros_like(E)), (torch.zeros_like(W), torch.zeros_like(W))
work = ops.spectralcf_work(N, d, K)
users = torch.randint(0, nu, (bs,), device="cuda", dtype=torch.int32)
pos = torch.randint(0, ni, (bs,), device="cuda", dtype=torch.int32)
neg = torch.randint(0, ni, (bs,), device="cuda", dtype=torch.int32)
loss = torch.zeros(1, device="cuda")


def step(s):
    ops.spectralcf_grad(nu, A, At, E, W, "sigmoid", users, pos, neg, "bpr", 1e-3, all_emb, G, touched, gE, gW, work, loss)
    ops.opt_apply_multi("adam", [(E, gE, sE[0], sE[1], None, True), (W, gW, sW[0], sW[1], None, True)], s + 1,
                        [1e-3, 0.9, 0.999, 1e-8])


for s in range(5):
    step(s)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for s in range(50):
    step(s + 5)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 50
flops = 4 * 2.0 * N * N * d + 6 * 2.0 * N * d * d
print("SpectralCF step (N=%d, d=%d, K=%d, batch %d): %.3f ms, %.1f TFLOP/s fp32, %.0f triplets/s, loss finite %s" % (
    N, d, K, bs, ms, flops / ms / 1e9, bs / ms * 1e3, bool(torch.isfinite(loss).item())))
