"""Where does the tensor-core evaluator beat the SIMT one?  users x items x dim sweep."""
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from neurec_b200 import ops
def csr(rs, nu, ni, deg):
    idx = np.sort(rs.randint(0, ni - deg, (nu, deg)), 1) + np.arange(deg)
    return np.arange(nu + 1, dtype=np.int64) * deg, idx.reshape(-1).astype(np.int32)
def run(nu, ni, dim, K=20):
    g = torch.Generator(device="cuda").manual_seed(1)
    U = torch.randn(nu, dim, device="cuda", generator=g) * 0.1
    V = torch.randn(ni, dim, device="cuda", generator=g) * 0.1
    rs = np.random.RandomState(2)
    tp, ti = csr(rs, nu, ni, 27); sp, si = csr(rs, nu, ni, 7)
    d = lambda a: torch.from_numpy(a).cuda()
    args = (U, V, torch.arange(nu, dtype=torch.int32, device="cuda"), d(tp), d(ti), d(sp), d(si), [1, 2, 3, 4, 5], K)
    out = {}
    for name, fn in (("simt", ops.eval_mf), ("tc", ops.eval_mf_tc)):
        r = fn(*args); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            r = fn(*args)
        e1.record(); torch.cuda.synchronize()
        out[name] = (e0.elapsed_time(e1) / 3, r)
    same = torch.equal(out["simt"][1], out["tc"][1])
    print("users %6d items %7d dim %3d: simt %8.3f ms  tc %8.3f ms  (x%.2f)  identical=%s replays=%d" % (
        nu, ni, dim, out["simt"][0], out["tc"][0], out["simt"][0] / out["tc"][0], same, ops.eval_last_undecided()))
for cfg in ((943, 1682, 64), (6040, 3706, 64), (29858, 40981, 64), (29858, 40981, 128), (8192, 16384, 64), (50000, 100000, 64), (20000, 262144, 128)):
    run(*cfg)
