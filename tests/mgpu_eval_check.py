"""Run under torchrun with N >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_eval_check.py

Checks that the user-sharded evaluator (contiguous user blocks per rank, NCCL all-gather of the
per-user rows) prints the SAME string as the single-GPU evaluation, bit for bit."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from neurec_b200 import ops
    from neurec_b200.data import Dataset
    from neurec_b200.evaluator import ProxyEvaluator
    z = np.load(os.path.join(ROOT, "tests", "golden", "ml100k_split.npz"))
    shape = (int(z["num_users"]), int(z["num_items"]))
    mk = lambda p, i: sp.csr_matrix((np.ones(len(z[i]), np.float32), z[i].astype(np.int32), z[p].astype(np.int64)), shape=shape)
    ds = Dataset.from_csr("ml-100k", mk("train_indptr", "train_indices"), mk("test_indptr", "test_indices"))
    rng = np.random.RandomState(1)
    U = torch.from_numpy((rng.randn(shape[0], 64) * .01).astype(np.float32)).cuda()
    V = torch.from_numpy((rng.randn(shape[1], 64) * .01).astype(np.float32)).cuda()

    class M:
        def get_eval_tables(self):
            return U, V
    ev = ProxyEvaluator(ds.get_user_train_dict(), ds.get_user_test_dict(), None,
                        metric=["Precision", "Recall", "NDCG", "MAP", "MRR"], top_k=[10, 20], batch_size=128)
    sharded = ev.evaluate(M())
    uni = ev.evaluator
    users = list(uni.user_pos_test.keys())
    rows = uni._evaluate_fused(M(), users)                 # every rank: all users, no sharding
    final = ops.mean_rows(rows).cpu().numpy().reshape(5, 20)[:, [9, 19]].reshape(-1)
    single = "\t".join([("%.8f" % x).ljust(12) for x in final])
    ok = sharded == single
    print("rank %d/%d sharded == single-GPU string: %s" % (dist.get_rank(), dist.get_world_size(), ok), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
