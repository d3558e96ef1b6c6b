"""MF: BPRMF (is_pairwise=True, loss_function=bpr) and pointwise "GMF"-style MF.

Plug-in mirror of the reference's model/general_recommender/MF.py:16-134 (same constructor,
configuration keys, log lines, predict contract) with the TensorFlow graph replaced by the fused
sm_100a step: ``build_graph`` allocates the tables / optimizer slots / gradient accumulators in
HBM, ``train_model`` runs ONE persistent launch per epoch (``nrc_mf_epoch_fused``: shuffle +
negative sampling + every batch of the epoch), reading back only the per-step losses.
"""
from time import time

import numpy as np
import torch

from ... import ops
from ...data import PairwiseSampler, PointwiseSampler
from ...util import timer
from ..AbstractRecommender import AbstractRecommender
from .._engine import OptimizerState, get_initializer


class MF(AbstractRecommender):
    def __init__(self, sess, dataset, conf):
        super(MF, self).__init__(dataset, conf)
        self.learning_rate = conf["learning_rate"]
        self.embedding_size = conf["embedding_size"]
        self.learner = conf["learner"]
        self.loss_function = conf["loss_function"]
        self.is_pairwise = conf["is_pairwise"]
        self.num_epochs = conf["epochs"]
        self.reg_mf = conf["reg_mf"]
        self.batch_size = conf["batch_size"]
        self.verbose = conf["verbose"]
        self.num_negatives = conf["num_negatives"]
        self.init_method = conf["init_method"]
        self.stddev = conf["stddev"]
        self.dataset = dataset
        self.num_users = dataset.num_users
        self.num_items = dataset.num_items
        self.sess = sess                      # unused: there is no TF session

    def build_graph(self):
        # MF.py:78-82: variables (49-52), loss (62-72) and optimizer (74-76) become device state
        gen = torch.Generator().manual_seed(2017)
        init = get_initializer(self.init_method, self.stddev, gen)
        self.user_embeddings = init([self.num_users, self.embedding_size]).cuda()
        self.item_embeddings = init([self.num_items, self.embedding_size]).cuda()
        loss = self.loss_function.lower()
        allowed = ("bpr", "hinge", "square") if self.is_pairwise is True else ("cross_entropy", "square")
        if loss not in allowed:
            raise Exception("please choose a suitable loss function")      # learner.py:27-28,39-40
        self._loss = loss
        self.opt = OptimizerState(self.learner, self.learning_rate)
        U, V = self.user_embeddings, self.item_embeddings
        self._gU, self._gV = torch.zeros_like(U), torch.zeros_like(V)
        self._s0U, self._s1U = self.opt.slots_like(U)
        self._s0V, self._s1V = self.opt.slots_like(V)
        self._tU = torch.zeros(self.num_users, dtype=torch.int32, device="cuda")
        self._tV = torch.zeros(self.num_items, dtype=torch.int32, device="cuda")
        self._ws = self._step_loss = None

    def _train_epoch(self, data_iter):
        if self.is_pairwise is True and data_iter.neg_num != 1:
            raise ValueError("MF trains on one negative per positive (MF.py:88)")
        steps = len(data_iter)
        d, a = data_iter.epoch_args()
        if self._ws is None or self._ws[0].numel() < data_iter._n_samples():
            n = data_iter._n_samples()
            mk = lambda: torch.empty(n, dtype=torch.int32, device="cuda")
            self._ws = (mk(), mk(), mk())
        if self._step_loss is None or self._step_loss.numel() < steps:
            self._step_loss = torch.empty(max(steps, 1), dtype=torch.float32, device="cuda")
        ops.mf_epoch_fused(self.user_embeddings, self.item_embeddings, d["ptr"], d["idx"], d["users"], d["pos"],
                           a["neg_num"], self.is_pairwise is True, a["shuffle"], a["drop_last"], a["seed"],
                           a["epoch"], self.batch_size, 0, steps, self._loss, self.reg_mf, self.opt.kind,
                           self.opt.hyper, self.opt.device_pows(), self._gU, self._gV, self._tU, self._tV,
                           self._s0U, self._s1U, self._s0V, self._s1V, self.opt.take_stamps(steps),
                           self._ws[0], self._ws[1], self._ws[2], self._step_loss)
        self.opt.lr_t(steps)              # keep the host mirror of the beta powers in step
        return float(self._step_loss[:steps].sum().item())

    def train_model(self):
        self.logger.info(self.evaluator.metrics_info())
        if self.is_pairwise is True:
            data_iter = PairwiseSampler(self.dataset, neg_num=1, batch_size=self.batch_size, shuffle=True)
        else:
            data_iter = PointwiseSampler(self.dataset, neg_num=self.num_negatives,
                                         batch_size=self.batch_size, shuffle=True)
        for epoch in range(1, self.num_epochs + 1):
            start = time()
            total_loss = self._train_epoch(data_iter)
            self.logger.info("[iter %d : loss : %f, time: %f]" % (epoch, total_loss / len(data_iter),
                                                                  time() - start))
            if epoch % self.verbose == 0:
                self.logger.info("epoch %d:\t%s" % (epoch, self.evaluate()))

    @timer
    def evaluate(self):
        return self.evaluator.evaluate(self)

    def get_eval_tables(self):
        """Fast path of UniEvaluator: predict is user_embeddings[users] . item_embeddings^T."""
        return self.user_embeddings, self.item_embeddings

    def predict(self, user_ids, candidate_items=None):
        users = torch.as_tensor(np.asarray(user_ids, dtype=np.int32)).cuda()
        ratings = ops.mf_scores(self.user_embeddings, self.item_embeddings, users).cpu().numpy()
        if candidate_items is not None:
            ratings = [r[items] for r, items in zip(ratings, candidate_items)]   # MF.py:123-124
        return ratings
