"""Timing of the persistent MF epoch kernel on ml-100k (BASELINE config 1 / pointwise variant):
us per step with sampling + shuffling inside the timed region.  python profiles/dbg_epoch.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neurec_b200 import ops

z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ml100k_split.npz"))
nu, ni = int(z["num_users"]), int(z["num_items"])
tp, ti = z["train_indptr"].astype(np.int64), z["train_indices"].astype(np.int32)
pu = np.repeat(np.arange(nu, dtype=np.int32), np.diff(tp))
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
dtp, dti, dpu = dev(tp), dev(ti), dev(pu)


def run(pairwise, dim, bs, neg_num, loss, reps=20):
    g = torch.Generator(device="cuda").manual_seed(1)
    U = torch.randn(nu, dim, device="cuda", generator=g) * 0.01
    V = torch.randn(ni, dim, device="cuda", generator=g) * 0.01
    zl = torch.zeros_like
    gU, gV, mU, vU, mV, vV = zl(U), zl(V), zl(U), zl(U), zl(V), zl(V)
    tU = torch.zeros(nu, dtype=torch.int32, device="cuda"); tV = torch.zeros(ni, dtype=torch.int32, device="cuda")
    n = len(pu) * (1 if pairwise else neg_num + 1)
    steps = (n + bs - 1) // bs
    ws = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(3)]
    sl = torch.zeros(steps, device="cuda")
    pows = torch.tensor([0.9, 0.999], device="cuda")
    def epoch(e):
        ops.mf_epoch_fused(U, V, dtp, dti, dpu, dti, neg_num, pairwise, True, False, 2018, e, bs, 0, steps, loss, 0.0,
                           "adam", [1e-3, 0.9, 0.999, 1e-8], pows, gU, gV, tU, tV, mU, vU, mV, vV, 1 + e * steps,
                           ws[0], ws[1], ws[2], sl)
    for e in range(3):
        epoch(e)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for e in range(reps):
        epoch(3 + e)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%s dim %d bs %d: %d steps/epoch, %.3f ms/epoch, %.2f us/step, %.1f M samples/s, last loss %.4f" %
          ("pairwise" if pairwise else "pointwise", dim, bs, steps, ms, 1e3 * ms / steps, n / ms / 1e3,
           float(sl.sum()) / steps), flush=True)


run(True, 64, 512, 1, "bpr")
if os.environ.get("NRC_EPOCH_ONLY") is None:
    run(True, 128, 512, 1, "bpr")
    run(False, 32, 256, 4, "cross_entropy")
    run(True, 64, 4096, 1, "bpr")

if os.environ.get("NRC_EPOCH_ONLY") is None:
    import subprocess

    def sub(label, **env):
        e = dict(os.environ, NRC_EPOCH_ONLY="1", **{k: str(v) for k, v in env.items()})
        out = subprocess.run([sys.executable, __file__], env=e, capture_output=True, text=True).stdout.strip().splitlines()
        print("  %-70s %s" % (label, out[0] if out else "?"), flush=True)
    for mode in (0, 1, 2):
        sub("one-barrier kernel, NRC_BAR_MODE=%d" % mode, NRC_BAR_MODE=mode)
    for mode in (0, 1, 2):
        sub("two-barrier kernel, NRC_BAR_MODE=%d" % mode, NRC_BAR_MODE=mode, NRC_EPOCH_TWO_BARRIER=1)
    sub("two-barrier kernel, barriers only (both phases skipped)", NRC_EPOCH_DBG=3)
    sub("two-barrier kernel, optimizer phase + barriers", NRC_EPOCH_DBG=1)
    sub("two-barrier kernel, gradient phase + barriers", NRC_EPOCH_DBG=2)
