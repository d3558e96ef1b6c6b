import sys, ctypes, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from neurec_b200 import _lib, ops
w = bench.LightgcnGowalla(0); w.setup_device()
U, V = w.eval_tables()
users = torch.arange(w.d["num_users"], dtype=torch.int32, device="cuda")
res = ops.eval_mf(U, V, users, w.tp, w.ti, w.sp, w.si, bench.METRICS, 20)
c = ctypes.c_int32(0); _lib.load().nrc_eval_last_undecided(ctypes.byref(c)); print("undecided users:", c.value, "of", users.numel())
m1 = ops.mean_rows(res).cpu().numpy(); print("mean_rows == np.mean:", np.array_equal(m1, np.mean(res.cpu().numpy(), axis=0)))
import time
for name, fn in [("eval_mf", lambda: ops.eval_mf(U, V, users, w.tp, w.ti, w.sp, w.si, bench.METRICS, 20)), ("mean_rows", lambda: ops.mean_rows(res))]:
    fn(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); print(name, (time.perf_counter()-t)/5*1e3, "ms")
w2 = bench.BprmfMl100k(0); w2.setup_device()
u2 = torch.arange(943, dtype=torch.int32, device="cuda")
ops.eval_mf(w2.dU, w2.dV, u2, w2.tp, w2.ti, w2.sp, w2.si, bench.METRICS, 20)
_lib.load().nrc_eval_last_undecided(ctypes.byref(c)); print("ml-100k undecided:", c.value)
