import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from neurec_b200 import ops
def random_csr(rs, num_rows, num_cols, degrees):
    rows = [np.unique(rs.randint(0, num_cols, int(k))) for k in degrees]
    indptr = np.zeros(num_rows + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    return indptr, np.concatenate(rows).astype(np.int32)
ALL = [1, 2, 3, 4, 5]
def run(nu, ni, dim=128, K=20, simt=True):
    g = torch.Generator(device="cuda").manual_seed(1)
    U = torch.randn(nu, dim, device="cuda", generator=g) * 0.1
    V = torch.randn(ni, dim, device="cuda", generator=g) * 0.1
    rs = np.random.RandomState(2)
    tp, ti = random_csr(rs, nu, ni, np.full(nu, 50)); sp, si = random_csr(rs, nu, ni, np.full(nu, 10))
    d = lambda a: torch.from_numpy(a).cuda()
    args = (U, V, torch.arange(nu, dtype=torch.int32, device="cuda"), d(tp), d(ti), d(sp), d(si), ALL, K)
    for name, fn in (("tc", ops.eval_mf_tc), ("simt", ops.eval_mf)):
        if name == "simt" and not simt: continue
        fn(*args); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*args); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("%s users %d items %d: %.2f ms  %.3g users/s  %.1f TFLOP/s" % (name, nu, ni, ms, nu / ms * 1e3, 2.0 * nu * ni * dim / ms / 1e9))
run(4096, 1_000_000)
run(18944, 2_000_000, simt=False)
