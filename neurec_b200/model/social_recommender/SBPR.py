"""SBPR: social BPR (Zhao et al., CIKM 2014).

Plug-in mirror of the reference's model/social_recommender/SBPR.py:17-166 on the sm_100a kernels:
  * ``_get_SocialItemsSet`` (:39-49) runs once at construction as two sparse products on the host;
  * ``_get_pairwise_all_data`` + ``DataIterator(shuffle=True)`` (:103-149) -- per epoch and per positive a
    social item, a negative outside train + social items and the weight s_uk -- are ONE device kernel
    (``nrc_sbpr_epoch_build``: keyed-bijection order + counter-based draws, nothing on the host);
  * the batch loop (:111-121) is ``nrc_sbpr_train_epoch``: per batch the fused gather -> three scores (+ item
    bias) -> two pairwise losses -> gradients kernel and one TF-1.12 optimizer launch over user table, item
    table and bias.
"""
from time import time

import numpy as np
import scipy.sparse as sp
import torch

from ... import ops
from ...data.sampler import _EPOCH_COUNTER_NEXT
from ...util import timer
from ..AbstractRecommender import SocialAbstractRecommender
from .._engine import OptimizerState, get_initializer


def social_items_csr(train_matrix, social_matrix):
    """SBPR._get_SocialItemsSet (SBPR.py:39-49): items of a user's trusted users that the user has not interacted
    with -> CSR with ascending rows.  Users without train items get an empty row (the reference iterates train_dict)."""
    train = sp.csr_matrix(train_matrix, dtype=np.float32)
    train.data[:] = 1.0
    trust = sp.csr_matrix(social_matrix, dtype=np.float32)
    trust.data[:] = 1.0
    reach = (trust @ train).tocsr()
    reach.data[:] = 1.0
    social = (reach - reach.multiply(train)).tocsr()
    social.eliminate_zeros()
    has_train = np.diff(train.indptr) > 0
    social = sp.diags(has_train.astype(np.float32)) @ social
    social = social.tocsr()
    social.eliminate_zeros()
    social.sort_indices()
    return social.indptr.astype(np.int64), social.indices.astype(np.int32)


class SBPR(SocialAbstractRecommender):
    def __init__(self, sess, dataset, conf):
        super(SBPR, self).__init__(dataset, conf)
        self.learning_rate = conf["learning_rate"]
        self.embedding_size = conf["embedding_size"]
        self.learner = conf["learner"]
        self.loss_function = conf["loss_function"]
        self.num_epochs = conf["num_epochs"]
        self.reg_mf = conf["reg_mf"]
        self.batch_size = conf["batch_size"]
        self.init_method = conf["init_method"]
        self.stddev = conf["stddev"]
        self.verbose = conf["verbose"]
        self.dataset = dataset
        self.num_users = dataset.num_users
        self.num_items = dataset.num_items
        self.userids = self.dataset.userids
        self.seed = 2018
        self.sess = sess
        train = sp.csr_matrix(dataset.train_matrix)
        train.sort_indices()
        trust = sp.csr_matrix(self.social_matrix)
        trust.sort_indices()
        tptr, tidx = train.indptr.astype(np.int64), train.indices.astype(np.int32)
        sptr, sidx = social_items_csr(train, trust)
        # the positives of the users that have social items, in dict (ascending-user) order (SBPR.py:125-131)
        eligible = np.diff(sptr) > 0
        deg = np.diff(tptr)
        pos_users = np.repeat(np.arange(self.num_users, dtype=np.int32), np.where(eligible, deg, 0))
        keep = np.repeat(eligible, deg)
        self._n = int(keep.sum())
        self._max_excluded = int((deg + np.diff(sptr))[eligible].max()) if eligible.any() else 0
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        self._d = {"tptr": dev(tptr), "tidx": dev(tidx), "sptr": dev(sptr), "sidx": dev(sidx),
                   "fptr": dev(trust.indptr.astype(np.int64)), "fidx": dev(trust.indices.astype(np.int32)),
                   "users": dev(pos_users), "pos": dev(tidx[keep])}

    def build_graph(self):
        loss = self.loss_function.lower()
        if loss not in ("bpr", "hinge", "square"):
            raise Exception("please choose a suitable loss function")      # learner.py:27-28
        self._loss = loss
        gen = torch.Generator().manual_seed(2017)
        init = get_initializer(self.init_method, self.stddev, gen)
        self.user_embeddings = init([self.num_users, self.embedding_size]).cuda()     # SBPR.py:59-64
        self.item_embeddings = init([self.num_items, self.embedding_size]).cuda()
        self.bias = init([self.num_items]).cuda()
        self.opt = OptimizerState(self.learner, self.learning_rate)
        z = torch.zeros_like
        self._g = [z(self.user_embeddings), z(self.item_embeddings), z(self.bias)]
        self._slots = [self.opt.slots_like(t) for t in (self.user_embeddings, self.item_embeddings, self.bias)]
        self._tU = torch.zeros(self.num_users, dtype=torch.int32, device="cuda")
        self._tV = torch.zeros(self.num_items, dtype=torch.int32, device="cuda")
        self._step_loss = None

    def _get_pairwise_all_data(self, shuffle=True):
        """(user_input, item_input_pos, item_input_social, item_input_neg, suk_input) of one epoch, already in
        DataIterator's shuffled order, as CUDA tensors (SBPR.py:103-106,123-149)."""
        d = self._d
        return ops.sbpr_epoch_build(d["tptr"], d["tidx"], d["sptr"], d["sidx"], d["fptr"], d["fidx"], d["users"],
                                    d["pos"], self.num_items, self._max_excluded, shuffle, self.seed,
                                    _EPOCH_COUNTER_NEXT())

    def _train_epoch(self):
        users, pos, soc, neg, suk = self._get_pairwise_all_data()
        steps = (self._n + self.batch_size - 1) // self.batch_size
        if self._step_loss is None or self._step_loss.numel() < steps:
            self._step_loss = torch.empty(max(steps, 1), dtype=torch.float32, device="cuda")
        (gU, gV, gB), ((s0U, s1U), (s0V, s1V), (s0B, s1B)) = self._g, self._slots
        ops.sbpr_train_epoch(self.user_embeddings, self.item_embeddings, self.bias, users, pos, soc, neg, suk,
                             self.batch_size, self._loss, self.reg_mf, self.opt.kind, self.opt.lr_t(steps),
                             self.opt.hyper, gU, gV, gB, self._tU, self._tV, s0U, s1U, s0V, s1V, s0B, s1B,
                             self.opt.take_stamps(steps), self._step_loss)
        return float(self._step_loss[:steps].sum().item())

    def train_model(self):
        self.logger.info(self.evaluator.metrics_info())
        for epoch in range(self.num_epochs):
            start = time()
            total_loss = self._train_epoch()
            self.logger.info("[iter %d : loss : %f, time: %f]" % (epoch, total_loss / max(self._n, 1), time() - start))
            if epoch % self.verbose == 0:
                self.logger.info("epoch %d:\t%s" % (epoch, self.evaluate()))

    @timer
    def evaluate(self):
        return self.evaluator.evaluate(self)

    def get_eval_tables(self):
        # predict ignores the item bias, exactly like the reference (SBPR.py:151-166)
        return self.user_embeddings, self.item_embeddings

    def predict(self, user_ids, candidate_items_userids=None):
        users = torch.as_tensor(np.asarray(user_ids, dtype=np.int32)).cuda()
        ratings = ops.mf_scores(self.user_embeddings, self.item_embeddings, users).cpu().numpy()
        if candidate_items_userids is not None:
            ratings = [r[items] for r, items in zip(ratings, candidate_items_userids)]
        return ratings
