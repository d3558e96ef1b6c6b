import sys, ctypes, numpy as np, torch
sys.path.insert(0, "/root/repo")
from neurec_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
for K, sw in ((16, 0), (64, 0), (128, 0), (64, 1), (128, 1)):
    A = torch.randn(128, K, device="cuda").bfloat16(); B = torch.randn(256, K, device="cuda").bfloat16()
    out = torch.zeros(128, 256, device="cuda")
    rc = lib.nrc_tc_gemm_debug(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), K, sw, ctypes.c_void_p(out.data_ptr()), None)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    err = (out - ref).abs().max().item()
    print("K", K, "swizzle", sw, "rc", rc, "max abs err", err, "ref max", ref.abs().max().item(), "nonzero", int((out != 0).sum()))
