import sys, ctypes, numpy as np, torch
sys.path.insert(0, "/root/repo")
from neurec_b200 import ops, _lib
def random_csr(rs, num_rows, num_cols, degrees):
    rows = [np.unique(rs.randint(0, num_cols, int(k))) for k in degrees]
    indptr = np.zeros(num_rows + 1, np.int64); indptr[1:] = np.cumsum([len(r) for r in rows])
    return indptr, np.concatenate(rows).astype(np.int32)
nu, ni, dim, K = 4096, 1_000_000, 128, 20
g = torch.Generator(device="cuda").manual_seed(1)
U = torch.randn(nu, dim, device="cuda", generator=g) * 0.1
V = torch.randn(ni, dim, device="cuda", generator=g) * 0.1
rs = np.random.RandomState(2)
tp, ti = random_csr(rs, nu, ni, np.full(nu, 50)); sp, si = random_csr(rs, nu, ni, np.full(nu, 10))
d = lambda a: torch.from_numpy(a).cuda()
args = (U, V, torch.arange(nu, dtype=torch.int32, device="cuda"), d(tp), d(ti), d(sp), d(si), [1,2,3,4,5], K)
ops.eval_mf_tc(*args); torch.cuda.synchronize()
c = ctypes.c_int32(0); _lib.load().nrc_eval_last_undecided(ctypes.byref(c)); print("undecided/overflowed users:", c.value, "of", nu)
ops.eval_mf_tc(*args); torch.cuda.synchronize()
