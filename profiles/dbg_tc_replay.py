import sys, ctypes, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import oracle
from neurec_b200 import ops, _lib
from test_gpu_tc_eval import _problem, dev, ALL
from conftest import random_csr
lib = _lib.load()
lib.nrc_eval_force_exact(1)
cases = ((200, 6000, 64, 20, False), (130, 3000, 128, 7, False), (140, 5000, 64, 20, True), (64, 45, 64, 31, False))
which = [int(a) for a in sys.argv[1:]] or range(len(cases))
for c in which:
    nu, ni, dim, K, ints = cases[c]
    U, V, tp, ti, sp, si = _problem(nu, ni, dim, 17 + ni, int_tables=ints)
    if ni < 100:
        rs = np.random.RandomState(4)
        tp, ti = random_csr(rs, nu, ni, rs.randint(0, 20, nu)); sp, si = random_csr(rs, nu, ni, rs.randint(1, 6, nu))
    users = np.arange(nu, dtype=np.int32)
    want, wranks = oracle.eval_mf(U, V, users, tp, ti, sp, si, ALL, K, return_ranks=True)
    got, ranks = ops.eval_mf_tc(dev(U), dev(V), dev(users), dev(tp), dev(ti), dev(sp), dev(si), ALL, K, return_ranks=True)
    torch.cuda.synchronize()
    n = ctypes.c_int32(0); lib.nrc_eval_last_undecided(ctypes.byref(n))
    print(c, cases[c], "ranks", np.array_equal(ranks.cpu().numpy(), wranks), "res", np.array_equal(got.cpu().numpy(), want), "replayed", n.value, flush=True)
