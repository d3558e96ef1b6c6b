"""UniEvaluator on the sm_100a evaluator kernels.

Mirror of the reference's evaluator/backend/cpp/uni_evaluator.py:16-157 (constructor, metric
names and ids, ``metrics_info()`` / ``evaluate()`` string formats, ``test_users`` override).
What changes is where the work happens:
  * models that expose ``get_eval_tables() -> (user_table, item_table)`` (MF, LightGCN: predict is
    U[users] . V^T) are evaluated by ONE fused launch of ``nrc_eval_mf`` over all test users --
    scores never leave the SM -- or, for catalogues of >= 16 384 items, by ``nrc_eval_mf_tc``
    (tcgen05 candidate pass + exact re-score; same bits);
  * any other model keeps the reference flow per batch: ``model.predict(batch_users, None)`` ->
    [B, num_items] scores -> train items masked to -inf (``nrc_mask_rows``) ->
    ``nrc_eval_score_matrix``;
  * ``rec.evaluate.neg > 0`` (``user_neg_test``): candidate lists padded with -inf, truth = first
    len(pos) positions (uni_evaluator.py:123-131), then the same score-matrix kernel.
The per-user rows are averaged by ``nrc_mean_rows`` in numpy's fp32 order (uni_evaluator.py:150).
"""
import numpy as np
import torch

from .. import ops
from ..util.tool import pad_sequences, typeassert
from . import sharded
from .abstract_evaluator import AbstractEvaluator

metric_dict = {"Precision": 1, "Recall": 2, "MAP": 3, "NDCG": 4, "MRR": 5}
re_metric_dict = {v: k for k, v in metric_dict.items()}


def _dict_to_device_csr(d, num_rows):
    # the reference turns every list into a set (cpp_evaluator.pyx:33-36): duplicates collapse
    rows = {u: np.unique(np.asarray(list(items), dtype=np.int32)) for u, items in d.items()}
    ptr = np.zeros(num_rows + 1, dtype=np.int64)
    for u, r in rows.items():
        ptr[u + 1] = len(r)
    ptr = np.cumsum(ptr)
    idx = np.empty(int(ptr[-1]), dtype=np.int32)
    for u, r in rows.items():
        idx[ptr[u]:ptr[u + 1]] = r
    if idx.size == 0:
        idx = np.zeros(1, np.int32)
    return torch.from_numpy(ptr).cuda(), torch.from_numpy(idx).cuda()


def _rows_to_device_csr(rows):
    rows = [np.unique(np.asarray(list(r), dtype=np.int32)) for r in rows]
    ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(r) for r in rows])
    idx = np.concatenate(rows).astype(np.int32) if rows and ptr[-1] > 0 else np.zeros(1, np.int32)
    return torch.from_numpy(ptr).cuda(), torch.from_numpy(np.ascontiguousarray(idx)).cuda()


class UniEvaluator(AbstractEvaluator):
    @typeassert(user_train_dict=dict, user_test_dict=(dict, None.__class__))
    def __init__(self, user_train_dict, user_test_dict, user_neg_test=None, metric=None, top_k=50,
                 batch_size=1024, num_thread=8):
        super(UniEvaluator, self).__init__()
        if metric is None:
            metric = ["Precision", "Recall", "MAP", "NDCG", "MRR"]
        elif isinstance(metric, str):
            metric = [metric]
        elif not isinstance(metric, (set, tuple, list)):
            raise TypeError("The type of 'metric' (%s) is invalid!" % metric.__class__.__name__)
        for m in metric:
            if m not in metric_dict:
                raise ValueError("There is not the metric named '%s'!" % metric)
        self.user_pos_train = user_train_dict
        self.user_pos_test = user_test_dict
        self.user_neg_test = user_neg_test
        self.metrics_num = len(metric)
        self.metrics = [metric_dict[m] for m in metric]
        self.num_thread = num_thread          # kept for signature compatibility; unused on the GPU
        self.batch_size = batch_size
        self.max_top = top_k if isinstance(top_k, int) else max(top_k)
        self.top_show = np.arange(top_k) + 1 if isinstance(top_k, int) else np.sort(top_k)
        self._csr = None

    def metrics_info(self):
        shown = ["\t".join([("%s@" % re_metric_dict[m] + str(k)).ljust(12) for k in self.top_show])
                 for m in self.metrics]
        return "metrics:\t%s" % "\t".join(shown)

    # -------------------------------------------------------------------------------
    def _device_csr(self):
        if self._csr is None:
            keys = list(self.user_pos_train.keys()) + list(self.user_pos_test.keys())
            n = int(max(keys)) + 1
            self._csr = (_dict_to_device_csr(self.user_pos_train, n), _dict_to_device_csr(self.user_pos_test, n))
        return self._csr

    def evaluate(self, model, test_users=None):
        test_users = test_users if test_users is not None else list(self.user_pos_test.keys())
        if not isinstance(test_users, (list, tuple, set, np.ndarray)):
            raise TypeError("'test_user' must be a list, tuple, set or numpy array!")
        test_users = list(test_users)
        # multi-GPU: each rank scores a contiguous block of the users (tables are replicated),
        # rows are all-gathered back in user order => same string for any world size
        rank, ws = sharded.world()
        n_total = len(test_users)
        if ws > 1:
            a, b = sharded.local_slice(n_total, rank, ws)
            test_users = test_users[a:b]
        if not test_users:
            rows = torch.zeros((0, self.metrics_num * self.max_top), dtype=torch.float32, device="cuda")
        elif self.user_neg_test is not None:
            rows = self._evaluate_candidates(model, test_users)
        elif hasattr(model, "get_eval_tables"):
            rows = self._evaluate_fused(model, test_users)
        else:
            rows = self._evaluate_generic(model, test_users)
        if ws > 1:
            rows = sharded.gather_rows(rows, n_total)
        final = ops.mean_rows(rows).cpu().numpy()                      # uni_evaluator.py:150
        final = np.reshape(final, [self.metrics_num, self.max_top])[:, self.top_show - 1].reshape(-1)
        return "\t".join([("%.8f" % x).ljust(12) for x in final])      # uni_evaluator.py:156

    def _evaluate_fused(self, model, test_users):
        (trp, tri), (tep, tei) = self._device_csr()
        U, V = model.get_eval_tables()
        users = torch.as_tensor(np.asarray(test_users, dtype=np.int32)).cuda()
        # Large catalogues go through the tensor-core path (bit-identical results).
        return ops.eval_mf_auto(U, V, users, trp, tri, tep, tei, self.metrics, self.max_top)

    def _evaluate_generic(self, model, test_users):
        (trp, tri), _ = self._device_csr()
        out = []
        for off in range(0, len(test_users), self.batch_size):
            bu = test_users[off:off + self.batch_size]
            score = model.predict(bu, None)                            # (B, N)
            if not (isinstance(score, torch.Tensor) and score.is_cuda):
                score = torch.from_numpy(np.ascontiguousarray(np.asarray(score, dtype=np.float32))).cuda()
            score = score.to(torch.float32).contiguous()
            users = torch.as_tensor(np.asarray(bu, dtype=np.int32)).cuda()
            ops.mask_rows(score, users, trp, tri)                      # uni_evaluator.py:140-143
            tp, ti = _rows_to_device_csr([self.user_pos_test[u] for u in bu])
            out.append(ops.eval_score_matrix(score, tp, ti, self.metrics, self.max_top))
        return torch.cat(out, dim=0)

    def _evaluate_candidates(self, model, test_users):
        out = []
        for off in range(0, len(test_users), self.batch_size):
            bu = test_users[off:off + self.batch_size]
            cands = [list(self.user_pos_test[u]) + self.user_neg_test[u] for u in bu]
            truth = [range(len(self.user_pos_test[u])) for u in bu]     # uni_evaluator.py:125
            score = model.predict(bu, cands)
            score = pad_sequences([np.asarray(s.cpu() if isinstance(s, torch.Tensor) else s) for s in score],
                                  value=-np.inf, dtype=np.float32)      # uni_evaluator.py:128
            tp, ti = _rows_to_device_csr(truth)
            out.append(ops.eval_score_matrix(torch.from_numpy(score).cuda(), tp, ti, self.metrics,
                                             self.max_top))
        return torch.cat(out, dim=0)
