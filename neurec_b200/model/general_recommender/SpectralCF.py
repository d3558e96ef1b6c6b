"""SpectralCF: spectral collaborative filtering (Zheng et al., RecSys 2018).

Plug-in mirror of the reference's model/general_recommender/SpectralCF.py:16-160.  The constructor builds the
same constant operator with the same numpy calls -- A = I + bipartite adjacency, L = I - D^-1 A, eig(L),
A_hat = U U^T + U diag(lamda) U^T cast to fp32 (:37-43,67-69,108-128) -- once, on the host, as the reference does;
everything per step runs on the device: ``nrc_spectralcf_grad`` (the K spectral convolutions forward, the pairwise
loss on the concatenated rows, the whole backward) + one dense TF-1.12 optimizer launch over the embedding table
and the filters.  A_hat is dense (users + items)^2 fp32: 27 MB on ml-100k, L2-resident.
"""
import warnings
from time import time

import numpy as np
import torch

from ... import ops
from ...data import PairwiseSampler
from ...util import timer
from ..AbstractRecommender import AbstractRecommender
from .._engine import OptimizerState, get_initializer


def spectral_operator(train_matrix):
    """A_hat of SpectralCF.__init__ / _create_inference (SpectralCF.py:37-43,67-69)."""
    graph = np.asarray(train_matrix.toarray(), dtype=np.float32)
    nu, ni = graph.shape
    n = nu + ni
    A = np.identity(n, dtype=np.float32)                     # adjacient_matrix(self_connection=True), :108-114
    A[:nu, nu:] += graph
    A[nu:, :nu] += graph.T
    D = A.sum(axis=1)                                        # degree_matrix, :116-119
    L = np.identity(n, dtype=np.float32) - np.dot(np.diag(np.power(D, -1)), A)     # laplacian_matrix(True), :121-128
    lamda, U = np.linalg.eig(L)                              # :41
    A_hat = np.dot(U, U.T) + np.dot(np.dot(U, np.diag(lamda)), U.T)                # :67
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                      # a complex eig result loses its imaginary part, as in :69
        return np.ascontiguousarray(A_hat.astype(np.float32))


class SpectralCF(AbstractRecommender):
    def __init__(self, sess, dataset, conf):
        super(SpectralCF, self).__init__(dataset, conf)
        self.learning_rate = conf["learning_rate"]
        self.learner = conf["learner"]
        self.batch_size = conf["batch_size"]
        self.num_layers = conf["num_layers"]
        self.activation = conf["activation"]
        self.embedding_size = conf["embedding_size"]
        self.num_epochs = conf["epochs"]
        self.reg = conf["reg"]
        self.loss_function = conf["loss_function"]
        self.dropout = conf["dropout"]
        self.embed_init_method = conf["embed_init_method"]
        self.weight_init_method = conf["weight_init_method"]
        self.stddev = conf["stddev"]
        self.verbose = conf["verbose"]
        self.dataset = dataset
        self.num_users = dataset.num_users
        self.num_items = dataset.num_items
        self.A_hat = spectral_operator(dataset.train_matrix)
        self.sess = sess

    def build_graph(self):
        if self.loss_function.lower() not in ("bpr", "hinge", "square"):
            raise Exception("please choose a suitable loss function")          # learner.py:27-28
        if self.activation not in ops.ACT_IDS:
            raise NotImplementedError("ERROR")                                   # tool.py:32-33
        gen = torch.Generator().manual_seed(2017)
        e_init = get_initializer(self.embed_init_method, self.stddev, gen)
        w_init = get_initializer(self.weight_init_method, self.stddev, gen)
        d, K, N = self.embedding_size, self.num_layers, self.num_users + self.num_items
        self.embeddings = torch.cat([e_init([self.num_users, d]), e_init([self.num_items, d])], dim=0).cuda()   # :50-56
        self.filters = torch.stack([w_init([d, d]) for _ in range(K)]).cuda() if K else \
            torch.zeros((0, d, d), device="cuda")                                                                 # :58-61
        self._A = torch.from_numpy(self.A_hat).cuda()
        self._At = self._A.t().contiguous()
        z = torch.zeros
        self._all = z((N, d * (K + 1)), device="cuda")
        self._G = z((N, d * (K + 1)), device="cuda")
        self._touched = z(N, dtype=torch.int32, device="cuda")
        self._gE, self._gW = torch.zeros_like(self.embeddings), torch.zeros_like(self.filters)
        self._work = ops.spectralcf_work(N, d, K)
        self.opt = OptimizerState(self.learner, self.learning_rate)
        self._sE, self._sW = self.opt.slots_like(self.embeddings), self.opt.slots_like(self.filters)

    def _step(self, users, pos, neg, loss_out):
        ops.spectralcf_grad(self.num_users, self._A, self._At, self.embeddings, self.filters, self.activation, users, pos,
                            neg, self.loss_function, self.reg, self._all, self._G, self._touched, self._gE, self._gW,
                            self._work, loss_out)
        hyper = list(self.opt.hyper)
        if self.opt.kind == "adam":
            hyper[0] = float(self.opt.lr_t(1)[0])
        segs = [(self.embeddings, self._gE, self._sE[0], self._sE[1], None, True)]
        if self.num_layers:
            segs.append((self.filters, self._gW, self._sW[0], self._sW[1], None, True))
        ops.opt_apply_multi(self.opt.kind, segs, self.opt.take_stamps(1), hyper)

    def train_model(self):
        self.logger.info(self.evaluator.metrics_info())
        data_iter = PairwiseSampler(self.dataset, neg_num=1, batch_size=self.batch_size, shuffle=True)
        for epoch in range(1, self.num_epochs + 1):
            start = time()
            users, pos, neg = data_iter.device_epoch()
            steps = len(data_iter)
            loss = torch.zeros(1, device="cuda")
            for s in range(steps):
                sl = slice(s * self.batch_size, (s + 1) * self.batch_size)
                self._step(users[sl], pos[sl], neg[sl], loss)
            self.logger.info("[iter %d : loss : %f, time: %f]" % (epoch, float(loss.item()) / steps, time() - start))
            if epoch % self.verbose == 0:
                self.logger.info("epoch %d:\t%s" % (epoch, self.evaluate()))

    @timer
    def evaluate(self):
        ops.spectralcf_forward(self._A, self.embeddings, self.filters, self.activation, self._all, self._work)   # :141-144
        self._cur_user_embeddings = self._all[:self.num_users].contiguous()
        self._cur_item_embeddings = self._all[self.num_users:].contiguous()
        return self.evaluator.evaluate(self)

    def get_eval_tables(self):
        return self._cur_user_embeddings, self._cur_item_embeddings

    def predict(self, user_ids, candidate_items_userids=None):
        u = torch.as_tensor(np.asarray(user_ids, dtype=np.int32)).cuda()
        ratings = ops.mf_scores(self._cur_user_embeddings, self._cur_item_embeddings, u).cpu().numpy()
        if candidate_items_userids is not None:
            ratings = [ratings[idx][np.asarray(items)] for idx, items in enumerate(candidate_items_userids)]
        return ratings
