"""LightGCN: plug-in mirror of the reference's model/general_recommender/LightGCN.py:16-192 on
the CSR SpMM kernel (forward and backward propagation every step, dense Adam over E_0)."""
import numpy as np
import scipy.sparse as sp
import torch

from ... import ops
from ...data import PairwiseSampler
from ...util import timer
from ..AbstractRecommender import AbstractRecommender
from .._engine import OptimizerState, get_initializer


def bipartite_adjacency(users, items, n_users, n_items, adj_type, verbose=True):
    """create_adj_mat of the reference (LightGCN.py:35-78; NGCF.py:299-322 builds the same four
    matrices): [[0, R], [R^T, 0]] over users+items and its normalisations, fp64 scipy calls in the
    reference's order.  'plain' A; 'norm' D^-1 (A + I); 'gcmc' D^-1 A; 'pre' D^-1/2 A D^-1/2;
    anything else D^-1 A + I."""
    say = print if verbose else (lambda *a: None)
    u = np.asarray(users, dtype=np.int32)
    i = np.asarray(items, dtype=np.int32)
    n = n_users + n_items
    half = sp.csr_matrix((np.ones_like(u, dtype=np.float32), (u, i + n_users)), shape=(n, n))
    adj = half + half.T

    def row_normalised(a):
        deg = np.array(a.sum(1))
        with np.errstate(divide="ignore"):
            inv = np.power(deg, -1).flatten()
        inv[np.isinf(inv)] = 0.
        say("generate single-normalized adjacency matrix.")
        return sp.diags(inv).dot(a).tocoo()

    if adj_type == "plain":
        out = adj
        say("use the plain adjacency matrix")
    elif adj_type == "norm":
        out = row_normalised(adj + sp.eye(n))
        say("use the normalized adjacency matrix")
    elif adj_type == "gcmc":
        out = row_normalised(adj)
        say("use the gcmc adjacency matrix")
    elif adj_type == "pre":
        deg = np.array(adj.sum(1))
        with np.errstate(divide="ignore"):
            inv = np.power(deg, -0.5).flatten()
        inv[np.isinf(inv)] = 0.
        d = sp.diags(inv)
        out = d.dot(adj).dot(d)
        say("use the pre adjcency matrix")
    else:
        out = row_normalised(adj) + sp.eye(n)
        say("use the mean adjacency matrix")
    return out


class LightGCN(AbstractRecommender):
    def __init__(self, sess, dataset, config):
        super(LightGCN, self).__init__(dataset, config)
        self.lr = config["lr"]
        self.reg = config["reg"]
        self.emb_dim = config["embed_size"]
        self.batch_size = config["batch_size"]
        self.epochs = config["epochs"]
        self.n_layers = config["n_layers"]
        self.dataset = dataset
        self.n_users, self.n_items = dataset.num_users, dataset.num_items
        self.user_pos_train = dataset.get_user_train_dict(by_time=False)
        self.all_users = list(self.user_pos_train.keys())
        self.norm_adj = self.create_adj_mat(config["adj_type"])
        self.sess = sess

    @timer
    def create_adj_mat(self, adj_type):
        """LightGCN.py:35-78: bipartite adjacency and its normalisations, fp64 scipy like the
        reference; cast to fp32 when uploaded (LightGCN.py:151-154)."""
        users, items = self.dataset.get_train_interactions()
        return bipartite_adjacency(users, items, self.n_users, self.n_items, adj_type)

    def build_graph(self):
        gen = torch.Generator().manual_seed(2017)
        init = get_initializer("xavier_uniform", 0.0, gen)          # LightGCN.py:87-89
        e0 = torch.cat([init([self.n_users, self.emb_dim]), init([self.n_items, self.emb_dim])], dim=0)
        self.ego_embeddings = e0.cuda()                              # concat(user, item), :135
        A = self.norm_adj.tocoo().astype(np.float32).tocsr()         # :151-154
        A.sort_indices()
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        self._csr = (t(A.indptr.astype(np.int64)), t(A.indices.astype(np.int32)), t(A.data))
        sym = abs(A - A.T).max() == 0
        if sym:
            self._tcsr = None
        else:
            AT = A.T.tocsr()
            AT.sort_indices()
            self._tcsr = (t(AT.indptr.astype(np.int64)), t(AT.indices.astype(np.int32)), t(AT.data))
        self._row_order = t(np.argsort(-np.diff(A.indptr), kind="stable").astype(np.int32))
        z = lambda: torch.zeros_like(self.ego_embeddings)
        self._m, self._v = z(), z()
        self._e_final, self._g_final, self._g_e0 = z(), z(), z()
        self._work = (z(), z())
        self.opt = OptimizerState("adam", self.lr)                   # LightGCN.py:130
        self._final_ready = False

    @property
    def user_embeddings_final(self):
        return self._e_final[:self.n_users]

    @property
    def item_embeddings_final(self):
        return self._e_final[self.n_users:]

    def train_model(self):
        data_iter = PairwiseSampler(self.dataset, neg_num=1, batch_size=self.batch_size, shuffle=True)
        self.logger.info(self.evaluator.metrics_info())
        for epoch in range(self.epochs):
            users, pos, neg = data_iter.device_epoch()
            steps = len(data_iter)
            loss2 = torch.empty(max(steps, 1), 2, dtype=torch.float32, device="cuda")
            ops.lightgcn_train_epoch(self._csr, self._tcsr, self._row_order, self.n_users, self.n_items,
                                     self.n_layers, self.ego_embeddings, self._m, self._v, users, pos,
                                     neg, self.batch_size, self.reg, self.opt.lr_t(steps), self.opt.hyper,
                                     self._e_final, self._g_final, self._g_e0, self._work, loss2)
            result = self.evaluate_model()
            self.logger.info("epoch %d:\t%s" % (epoch, result))

    def evaluate_model(self):
        # LightGCN.py:183-185: snapshot the propagated tables (assign_opt), then evaluate
        ops.lightgcn_propagate(self._csr[0], self._csr[1], self._csr[2], self._row_order,
                               self.ego_embeddings, self.n_layers, self._e_final, self._work)
        return self.evaluator.evaluate(self)

    def get_eval_tables(self):
        return self.user_embeddings_final.contiguous(), self.item_embeddings_final.contiguous()

    def predict(self, users, candidate_items=None):
        u = torch.as_tensor(np.asarray(users, dtype=np.int32)).cuda()
        U, V = self.get_eval_tables()
        ratings = ops.mf_scores(U, V, u).cpu().numpy()
        if candidate_items is not None:
            ratings = [ratings[idx][u_item] for idx, u_item in enumerate(candidate_items)]
        return ratings
