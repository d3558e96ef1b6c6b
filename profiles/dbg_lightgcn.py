"""Why does the LightGCN timed region come out at 0.49 or at 0.82 ms/step?  Runs the bench workload's region
(100 steps) several times in one process, with and without the L2 flush in front, device-timed (CUDA events) and
wall-clock, and times the SpMM alone in between."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

w = bench.LightgcnGowalla(0)
w.setup()
bench.warm_up(lambda: w.run_steps(5), 1)
K = int(os.environ.get("K", "100"))
for rep in range(8):
    flush = rep % 2 == 0
    torch.cuda.synchronize()
    if flush:
        bench.flush_l2()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    w.run_steps(K)
    t1 = time.perf_counter()
    b.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("region %d (%s): device %.3f ms/step, host enqueue %.3f ms/step, wall %.3f ms/step" % (
        rep, "L2 flushed" if flush else "no flush", a.elapsed_time(b) / K, (t1 - t0) * 1e3 / K, (t2 - t0) * 1e3 / K), flush=True)
fn, _ = w.spmm_kernel()
print("SpMM alone (graph replay): %.1f us" % (bench.graph_time(fn) * 1e6))
