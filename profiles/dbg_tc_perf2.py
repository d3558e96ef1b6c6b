import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from neurec_b200 import ops
def random_csr(rs, num_rows, num_cols, degrees):
    rows = [np.unique(rs.randint(0, num_cols, int(k))) for k in degrees]
    indptr = np.zeros(num_rows + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    return indptr, np.concatenate(rows).astype(np.int32)
def run(nu, ni, dim=128, K=20):
    g = torch.Generator(device="cuda").manual_seed(1)
    U = torch.randn(nu, dim, device="cuda", generator=g) * 0.1
    V = torch.randn(ni, dim, device="cuda", generator=g) * 0.1
    rs = np.random.RandomState(2)
    tp, ti = random_csr(rs, nu, ni, np.full(nu, 50)); sp, si = random_csr(rs, nu, ni, np.full(nu, 10))
    d = lambda a: torch.from_numpy(a).cuda()
    args = (U, V, torch.arange(nu, dtype=torch.int32, device="cuda"), d(tp), d(ti), d(sp), d(si), [1,2,3,4,5], K)
    ops.eval_mf_tc(*args); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.eval_mf_tc(*args); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("tc users %d items %d: %.2f ms  %.1f TFLOP/s  %.2f us/tile" % (nu, ni, ms, 2.0 * nu * ni * dim / ms / 1e9, ms * 1e3 / ((ni + 255) // 256)))
if len(sys.argv) > 1:
    run(int(sys.argv[1]), int(sys.argv[2]))
else:
    for nu, ni in ((4096, 200_000), (9472, 200_000), (18944, 200_000), (18944, 1_000_000), (4096, 2_000_000)):
        run(nu, ni)
