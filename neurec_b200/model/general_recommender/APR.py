"""APR: adversarial personalized ranking (He et al., SIGIR 2018) as the reference runs it.

Plug-in mirror of the reference's model/general_recommender/APR.py:15-168.  What the reference's graph
TRAINS is ``learner.optimizer(self.learner, self.loss, ...)`` (APR.py:120-122): the plain BPR loss
``sum softplus(-(x_ui - x_uj))`` (:80-81) -- not ``opt_loss``; ``reg``, ``reg_adv`` and the adversarial
branch (:86-92) never reach the optimizer, and ``train_model`` (:132-151) never runs ``update_P`` /
``update_Q`` (:103-104,117-118).  So an APR epoch is a BPRMF epoch with reg = 0 and takes the same
persistent launch (``nrc_mf_epoch_fused``: shuffle + negative sampling + every step).  The adversarial
ops exist here as they do in the reference's graph: ``delta_P`` / ``delta_Q`` tables and
``update_adversarial(batch)`` (the ``_create_adversarial`` ops, random or gradient-based), built from
``nrc_mf_pairwise_grad`` + ``nrc_l2_normalize_rows``.
"""
from time import time

import numpy as np
import torch

from ... import ops
from ...data import PairwiseSampler
from ...util import timer
from ..AbstractRecommender import AbstractRecommender
from .._engine import OptimizerState, get_initializer


class APR(AbstractRecommender):
    def __init__(self, sess, dataset, conf):
        super(APR, self).__init__(dataset, conf)
        self.learning_rate = conf["learning_rate"]
        self.embedding_size = conf["embedding_size"]
        self.learner = conf["learner"]
        self.num_epochs = conf["epochs"]
        self.eps = conf["eps"]
        self.adv = conf["adv"]
        self.adver = conf["adver"]
        self.adv_epoch = conf["adv_epoch"]
        self.reg = conf["reg"]
        self.reg_adv = conf["reg_adv"]
        self.batch_size = conf["batch_size"]
        self.init_method = conf["init_method"]
        self.stddev = conf["stddev"]
        self.verbose = conf["verbose"]
        self.dataset = dataset
        self.num_users = dataset.num_users
        self.num_items = dataset.num_items
        self.sess = sess

    def build_graph(self):
        gen = torch.Generator().manual_seed(2017)
        self._gen = gen
        init = get_initializer(self.init_method, self.stddev, gen)
        self.embedding_P = init([self.num_users, self.embedding_size]).cuda()      # APR.py:46-49
        self.embedding_Q = init([self.num_items, self.embedding_size]).cuda()
        self.delta_P = torch.zeros_like(self.embedding_P)                          # APR.py:51-54, trainable=False
        self.delta_Q = torch.zeros_like(self.embedding_Q)
        self.opt = OptimizerState(self.learner, self.learning_rate)
        P, Q = self.embedding_P, self.embedding_Q
        self._gP, self._gQ = torch.zeros_like(P), torch.zeros_like(Q)
        self._s0P, self._s1P = self.opt.slots_like(P)
        self._s0Q, self._s1Q = self.opt.slots_like(Q)
        self._tP = torch.zeros(self.num_users, dtype=torch.int32, device="cuda")
        self._tQ = torch.zeros(self.num_items, dtype=torch.int32, device="cuda")
        self._ws = self._step_loss = None

    def update_adversarial(self, bat_users=None, bat_items_pos=None, bat_items_neg=None):
        """The ops of _create_adversarial (APR.py:92-118): delta = l2_normalize(rows) * eps with rows either
        truncated-normal noise (adv == "random") or the gradient of the BPR loss of one batch w.r.t. the
        embedding tables (adv == "grad", rows outside the batch normalise to 0)."""
        if self.adv == "random":
            noise = get_initializer("tnormal", 0.01, self._gen)
            for delta, n in ((self.delta_P, self.num_users), (self.delta_Q, self.num_items)):
                ops.l2_normalize_rows(noise([n, self.embedding_size]).cuda(), self.eps, delta)
        elif self.adv == "grad":
            t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32)).cuda()
            scratch = torch.zeros(1, device="cuda")
            ops.mf_pairwise_grad(self.embedding_P, self.embedding_Q, t(bat_users), t(bat_items_pos), t(bat_items_neg),
                                 "bpr", 0.0, self._gP, self._gQ, self._tP, self._tQ, 0, scratch)
            ops.l2_normalize_rows(self._gP, self.eps, self.delta_P)
            ops.l2_normalize_rows(self._gQ, self.eps, self.delta_Q)
            self._gP.zero_(); self._gQ.zero_()          # the accumulators must be clean for the next step

    def _train_epoch(self, data_iter):
        steps = len(data_iter)
        d, a = data_iter.epoch_args()
        n = data_iter._n_samples()
        if self._ws is None or self._ws[0].numel() < n:
            mk = lambda: torch.empty(n, dtype=torch.int32, device="cuda")
            self._ws = (mk(), mk(), mk())
        if self._step_loss is None or self._step_loss.numel() < steps:
            self._step_loss = torch.empty(max(steps, 1), dtype=torch.float32, device="cuda")
        # self.optimizer minimises self.loss (APR.py:120-122): BPR softplus sum, no regulariser
        ops.mf_epoch_fused(self.embedding_P, self.embedding_Q, d["ptr"], d["idx"], d["users"], d["pos"],
                           a["neg_num"], True, a["shuffle"], a["drop_last"], a["seed"], a["epoch"], self.batch_size, 0,
                           steps, "bpr", 0.0, self.opt.kind, self.opt.hyper, self.opt.device_pows(), self._gP,
                           self._gQ, self._tP, self._tQ, self._s0P, self._s1P, self._s0Q, self._s1Q,
                           self.opt.take_stamps(steps), self._ws[0], self._ws[1], self._ws[2], self._step_loss)
        self.opt.lr_t(steps)
        return float(self._step_loss[:steps].sum().item())

    def train_model(self):
        self.logger.info(self.evaluator.metrics_info())
        data_iter = PairwiseSampler(self.dataset, neg_num=1, batch_size=self.batch_size, shuffle=True)
        for epoch in range(1, self.num_epochs + 1):
            start = time()
            total_loss = self._train_epoch(data_iter)
            self.logger.info("[iter %d : loss : %f, time: %f]" % (epoch, total_loss / len(data_iter), time() - start))
            if epoch % self.verbose == 0:
                self.logger.info("epoch %d:\t%s" % (epoch, self.evaluate()))

    @timer
    def evaluate(self):
        return self.evaluator.evaluate(self)

    def get_eval_tables(self):
        return self.embedding_P, self.embedding_Q

    def predict(self, user_ids, candidate_items_userids=None):
        users = torch.as_tensor(np.asarray(user_ids, dtype=np.int32)).cuda()
        ratings = ops.mf_scores(self.embedding_P, self.embedding_Q, users).cpu().numpy()
        if candidate_items_userids is not None:
            ratings = [r[items] for r, items in zip(ratings, candidate_items_userids)]   # APR.py:160-166
        return ratings
