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
EVAL_ONLY = bool(os.environ.get("NRC_EVAL_ONLY"))
ms = ev(lambda: w.run_steps(w.spe), 1 if EVAL_ONLY else 5)
print("NeuMF ml-100k persistent epoch%s: %.3f ms / %d steps = %.2f us/step, %.1f M samples/s" %
      (" [NRC_EPOCH_DBG=%s]" % os.environ["NRC_EPOCH_DBG"] if os.environ.get("NRC_EPOCH_DBG") else "", ms, w.spe,
       1e3 * ms / w.spe, w.spe * w.batch / ms / 1e3), flush=True)
if os.environ.get("NRC_NCF_ONLY"):
    sys.exit(0)
import subprocess
for bits, what in () if EVAL_ONLY else ((15, "barriers only"), (14, "samples + barriers"), (13, "weight gradients + barriers"), (11, "tables + barriers"),
                   (7, "weight staging + barriers")):
    e = dict(os.environ, NRC_EPOCH_DBG=str(max(bits, 0)), NRC_NCF_ONLY="1")
    out = subprocess.run([sys.executable, __file__], env=e, capture_output=True, text=True).stdout.strip().splitlines()
    print("   %-32s %s" % (what, out[0] if out else "?"), flush=True)
# NeuMF evaluation (predict over all items + mask + top-K + metrics): the fast scoring kernel vs the generic one
import time
users = torch.arange(w.d["num_users"], dtype=torch.int32, device="cuda")
tp, ti, sp, si = (bench.dev(w.d[k]) for k in ("train_indptr", "train_indices", "test_indptr", "test_indices"))


def neumf_eval():
    sc = ops.ncf_scores(w.shape, w.P, users)
    ops.mask_rows(sc, users, tp, ti)
    return ops.eval_score_matrix(sc, sp, si, bench.METRICS, 20)


ms = ev(neumf_eval, 5)
print("NeuMF ml-100k evaluation (943 users x 1682 items): %.3f ms = %.2f M users/s" % (ms, 943 / ms / 1e3), flush=True)
sc = ops.ncf_scores(w.shape, w.P, users)
for name, fn in (("ncf_scores", lambda: ops.ncf_scores(w.shape, w.P, users)), ("mask_rows", lambda: ops.mask_rows(sc, users, tp, ti)),
                 ("eval_score_matrix", lambda: ops.eval_score_matrix(sc, sp, si, bench.METRICS, 20))):
    print("   %-20s %.1f us (stream time incl. host launch gaps)" % (name, 1e3 * ev(fn, 20)), flush=True)
try:
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        neumf_eval()
    print("   whole evaluation as a CUDA graph: %.1f us" % (1e3 * ev(g_.replay, 20)), flush=True)
except Exception as ex:          # measurement aid only
    print("   graph capture of the evaluation failed: %r" % (ex,), flush=True)
if EVAL_ONLY:
    sys.exit(0)
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
