"""Round-2 kernel timings outside bench.py: NeuMF persistent epoch, gowalla SpMM (fast / exact order),
LightGCN step.  python profiles/dbg_round2.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from neurec_b200 import ops


def ev(fn, reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


w = bench.NeumfMl100k(0); w.setup()
ms = ev(lambda: w.run_steps(w.spe), 5)
print("NeuMF ml-100k persistent epoch: %.3f ms / %d steps = %.2f us/step, %.1f M samples/s" %
      (ms, w.spe, 1e3 * ms / w.spe, w.spe * w.batch / ms / 1e3), flush=True)
del w
g = bench.LightgcnGowalla(0); g.setup()
fn, nbytes = g.spmm_kernel()
for exact in (False, True):
    ops.spmm_set_exact(exact)
    t = bench.graph_time(fn)
    print("gowalla SpMM (%s order): %.1f us, %.0f GB/s algorithmic" % ("exact" if exact else "fast", t * 1e6, nbytes / t / 1e9), flush=True)
ops.spmm_set_exact(False)
ms = ev(lambda: g.run_steps(50), 3)
print("LightGCN gowalla: %.3f ms/step (50 steps incl. epoch_build)" % (ms / 50), flush=True)
