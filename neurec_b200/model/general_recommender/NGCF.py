"""NGCF: plug-in mirror of the reference's model/general_recommender/NGCF.py:14-332 (alg_type 'ngcf')
on the CSR SpMM + the fused dense layer kernels (csrc/ngcf.cu): same constructor, configuration keys,
log lines and predict contract; message dropout is always on, also at evaluation (NGCF.py:193)."""
from time import time

import numpy as np
import torch

from ... import ops
from ...data import PairwiseSampler
from ...util import timer
from ..AbstractRecommender import AbstractRecommender
from .._engine import OptimizerState, get_initializer
from .LightGCN import bipartite_adjacency


class NGCF(AbstractRecommender):
    def __init__(self, sess, dataset, conf):
        super(NGCF, self).__init__(dataset, conf)
        self.learning_rate = conf["learning_rate"]
        self.learner = conf["learner"]
        self.batch_size = conf["batch_size"]
        self.emb_dim = conf["embedding_size"]
        self.weight_size = list(conf["layer_size"])
        self.n_layers = len(self.weight_size)
        self.num_epochs = conf["epochs"]
        self.reg = conf["reg"]
        self.adj_type = conf["adj_type"]
        self.alg_type = conf["alg_type"]
        self.node_dropout_flag = conf["node_dropout_flag"]
        self.node_dropout_ratio = conf["node_dropout_ratio"]
        self.mess_dropout_ratio = conf["mess_dropout_ratio"]
        self.n_fold = 100                     # NGCF.py:33; the slabs are one CSR pass here
        self.embed_init_method = conf["embed_init_method"]
        self.weight_init_method = conf["weight_init_method"]
        self.stddev = conf["stddev"]
        self.verbose = conf["verbose"]
        self.dataset = dataset
        self.num_users, self.num_items = dataset.num_users, dataset.num_items
        if self.alg_type not in ("ngcf",):
            raise NotImplementedError("alg_type '%s': only 'ngcf' (NGCF.py:160-202) is built on the device" % self.alg_type)
        if self.node_dropout_flag is True:
            raise NotImplementedError("node dropout (NGCF.py:334-356) is not built; conf default is False")
        self.norm_adj = self.get_adj_mat()
        self.n_nonzero_elems = self.norm_adj.count_nonzero()
        self.sess = sess

    @timer
    def get_adj_mat(self):
        """NGCF.py:299-332: the same four matrices as LightGCN.create_adj_mat ('norm' = D^-1 (A + I))."""
        users, items = self.dataset.get_train_interactions()
        return bipartite_adjacency(users, items, self.num_users, self.num_items, self.adj_type,
                                   verbose=False).tocsr()

    def build_graph(self):
        gen = torch.Generator().manual_seed(2017)
        e_init = get_initializer(self.embed_init_method, self.stddev, gen)
        w_init = get_initializer(self.weight_init_method, self.stddev, gen)
        self.shape = ops.NgcfShape.make(self.num_users, self.num_items, self.emb_dim, self.weight_size)
        self.ego_embeddings = torch.cat([e_init([self.num_users, self.emb_dim]),
                                         e_init([self.num_items, self.emb_dim])], dim=0).cuda()
        self.logger.info("using xavier initialization")             # NGCF.py:271
        dims = [self.emb_dim] + self.weight_size
        parts = []
        for k in range(self.n_layers):                              # NGCF.py:286-297: W_gc, b_gc, W_bi, b_bi
            for shp in ([dims[k], dims[k + 1]], [1, dims[k + 1]], [dims[k], dims[k + 1]], [1, dims[k + 1]]):
                parts.append(w_init(shp).reshape(-1))
        self.weights = torch.cat(parts).cuda()
        assert self.weights.numel() == self.shape.weights_size()
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

        def csr_dev(A):
            A = A.tocsr().astype(np.float32)
            A.sort_indices()
            return (t(A.indptr.astype(np.int64)), t(A.indices.astype(np.int32)), t(A.data)), \
                t(np.argsort(-np.diff(A.indptr), kind="stable").astype(np.int32))
        self._csr, self._order = csr_dev(self.norm_adj)
        self._tcsr, self._torder = csr_dev(self.norm_adj.T)          # 'norm' is not symmetric
        N, dt = self.shape.n_nodes, self.shape.d_total
        z = torch.zeros
        self._all = z((N, dt), device="cuda")
        self._G = z((N, dt), device="cuda")
        self._gE, self._gW = torch.zeros_like(self.ego_embeddings), torch.zeros_like(self.weights)
        self._work = torch.empty(self.shape.work_floats(), dtype=torch.float32, device="cuda")
        self._masks = torch.empty(self.shape.mask_floats(), dtype=torch.float32, device="cuda")
        self.opt = OptimizerState(self.learner, self.learning_rate)
        self._sE, self._sW = self.opt.slots_like(self.ego_embeddings), self.opt.slots_like(self.weights)
        self._draws = 0

    def _draw_masks(self):
        keep = 1.0 - self.mess_dropout_ratio
        if keep >= 1.0:
            return None, 1.0
        ops.dropout_mask(self._masks.numel(), keep, 2017, self._draws, out=self._masks)
        self._draws += 1
        return self._masks, keep

    def _step(self, users, pos, neg, loss2):
        masks, keep = self._draw_masks()
        ops.ngcf_grad(self.shape, self._csr, self._order, self._tcsr, self._torder, self.ego_embeddings, self.weights,
                      masks, keep, users, pos, neg, self.reg, self._all, self._G, self._gE, self._gW, self._work, loss2)
        hyper = list(self.opt.hyper)
        if self.opt.kind == "adam":
            hyper[0] = float(self.opt.lr_t(1)[0])
        # every NGCF variable has a dense gradient (E_0 through concat + SpMM): dense-gradient formulas
        ops.opt_apply_multi(self.opt.kind, [(self.ego_embeddings, self._gE, self._sE[0], self._sE[1], None, True),
                                            (self.weights, self._gW, self._sW[0], self._sW[1], None, True)],
                            self.opt.take_stamps(1), hyper)

    def train_model(self):
        self.logger.info(self.evaluator.metrics_info())
        data_iter = PairwiseSampler(self.dataset, neg_num=1, batch_size=self.batch_size, shuffle=True)
        for epoch in range(1, self.num_epochs + 1):
            start = time()
            users, pos, neg = data_iter.device_epoch()
            steps = len(data_iter)
            loss2 = torch.zeros(2, device="cuda")
            for s in range(steps):
                sl = slice(s * self.batch_size, (s + 1) * self.batch_size)
                self._step(users[sl], pos[sl], neg[sl], loss2)
            total_loss = float(loss2.sum().item())
            self.logger.info("[iter %d : loss : %f, time: %f]" % (epoch, total_loss / steps, time() - start))
            if epoch % self.verbose == 0:
                self.logger.info("epoch %d:\t%s" % (epoch, self.evaluate()))

    @timer
    def evaluate(self):
        # NGCF.py:144-146: one more forward (dropout still on) gives the tables that are ranked
        masks, keep = self._draw_masks()
        ops.ngcf_forward(self.shape, self._csr, self._order, self.ego_embeddings, self.weights, masks, keep, self._all,
                         self._work)
        self._cur_user_embeddings = self._all[:self.num_users].contiguous()
        self._cur_item_embeddings = self._all[self.num_users:].contiguous()
        return self.evaluator.evaluate(self)

    def get_eval_tables(self):
        return self._cur_user_embeddings, self._cur_item_embeddings

    def predict(self, user_ids, candidate_items_userids=None):
        u = torch.as_tensor(np.asarray(user_ids, dtype=np.int32)).cuda()
        ratings = ops.mf_scores(self._cur_user_embeddings, self._cur_item_embeddings, u).cpu().numpy()
        if candidate_items_userids is not None:                        # NGCF.py:153-158
            ratings = [ratings[idx][np.asarray(items)] for idx, items in enumerate(candidate_items_userids)]
        return ratings
