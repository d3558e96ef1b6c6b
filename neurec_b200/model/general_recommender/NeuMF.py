"""NeuMF (GMF + MLP) and, through the same code, MLP: plug-in mirrors of the reference's
model/general_recommender/NeuMF.py:16-169 and MLP.py:15-142 on the fused NCF kernels."""
import pickle
from time import time

import numpy as np
import torch

from ... import ops
from ...data import PairwiseSampler, PointwiseSampler
from ...util import timer
from ..AbstractRecommender import AbstractRecommender
from .._engine import OptimizerState, get_initializer

_KEYS = ("mf_user", "mf_item", "mlp_user", "mlp_item", "dense")


class _NCFBase(AbstractRecommender):
    has_mf = True

    def _common_init(self, sess, dataset, conf):
        self.layers = conf["layers"]
        self.learning_rate = conf["learning_rate"]
        self.learner = conf["learner"]
        self.loss_function = conf["loss_function"]
        self.num_epochs = conf["epochs"]
        self.num_negatives = conf["num_neg"]
        self.batch_size = conf["batch_size"]
        self.verbose = conf["verbose"]
        self.is_pairwise = conf["is_pairwise"]
        self.init_method = conf["init_method"]
        self.stddev = conf["stddev"]
        self.num_users = dataset.num_users
        self.num_items = dataset.num_items
        self.dataset = dataset
        self.sess = sess

    def _n_towers(self):
        return 1

    def _build(self, mf_dim, reg_mf, reg_mlp, pretrained=None):
        gen = torch.Generator().manual_seed(2017)
        init = get_initializer(self.init_method, self.stddev, gen)
        mlp_dim = int(self.layers[0] / 2)                        # NeuMF.py:58 / MLP.py:48
        self.shape = ops.NcfShape.make(self.num_users, self.num_items, mf_dim, self.layers, self._n_towers())
        P = {"mf_user": None, "mf_item": None}
        if mf_dim:
            P["mf_user"] = init([self.num_users, mf_dim]).cuda()
            P["mf_item"] = init([self.num_items, mf_dim]).cuda()
        P["mlp_user"] = init([self.num_users, mlp_dim]).cuda()
        P["mlp_item"] = init([self.num_items, mlp_dim]).cuda()
        if pretrained is not None:                               # NeuMF.py:62-67
            for k, v in zip(("mf_user", "mf_item"), pretrained[0]):
                P[k] = torch.as_tensor(np.asarray(v, np.float32)).cuda()
            for k, v in zip(("mlp_user", "mlp_item"), pretrained[1]):
                P[k] = torch.as_tensor(np.asarray(v, np.float32)).cuda()
        # tf.layers.dense defaults: glorot_uniform kernel, zeros bias; packed per tower
        dense = torch.zeros(self.shape.dense_size(), dtype=torch.float32)
        tower = dense.numel() // self.shape.n_towers
        off, inn = 0, 2 * mlp_dim
        for out in self.layers:
            lim = (6.0 / (inn + out)) ** 0.5
            for t in range(self.shape.n_towers):
                w = (torch.rand(inn * out, generator=gen) * 2 - 1) * lim
                dense[t * tower + off:t * tower + off + inn * out] = w
            off += inn * out + out
            inn = out
        P["dense"] = dense.cuda()
        self.params = P
        loss = self.loss_function.lower()
        allowed = ("bpr", "hinge", "square") if self.is_pairwise is True else ("cross_entropy", "square")
        if loss not in allowed:
            raise Exception("please choose a suitable loss function")
        self._loss, self._reg_mf, self._reg_mlp = loss, reg_mf, reg_mlp
        self.opt = OptimizerState(self.learner, self.learning_rate)
        self._G = {k: (torch.zeros_like(v) if v is not None else None) for k, v in P.items()}
        self._S0, self._S1 = {}, {}
        for k, v in P.items():
            self._S0[k], self._S1[k] = (None, None) if v is None else self.opt.slots_like(v)
        self._tU = torch.zeros(self.num_users, dtype=torch.int32, device="cuda")
        self._tI = torch.zeros(self.num_items, dtype=torch.int32, device="cuda")

    def _make_sampler(self):
        if self.is_pairwise is True:
            return PairwiseSampler(self.dataset, neg_num=1, batch_size=self.batch_size, shuffle=True)
        return PointwiseSampler(self.dataset, neg_num=self.num_negatives, batch_size=self.batch_size,
                                shuffle=True)

    def _train_epoch(self, data_iter):
        """One epoch = ONE persistent launch (nrc_ncf_epoch_fused): shuffle + negative sampling + every
        batch of NeuMF.py:131-147 / MLP.py:104-120; only the per-step losses come back."""
        steps = len(data_iter)
        d, a = data_iter.epoch_args()
        n = data_iter._n_samples()
        if getattr(self, "_ws", None) is None or self._ws[0].numel() < n:
            self._ws = tuple(torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(3))
        if getattr(self, "_step_loss", None) is None or self._step_loss.numel() < steps:
            self._step_loss = torch.empty(max(steps, 1), dtype=torch.float32, device="cuda")
        ops.ncf_epoch_fused(self.shape, self.params, d["ptr"], d["idx"], d["users"], d["pos"], a["neg_num"],
                            self.is_pairwise is True, a["shuffle"], a["drop_last"], a["seed"], a["epoch"],
                            self.batch_size, 0, steps, self._loss, self._reg_mf, self._reg_mlp, self.opt.kind,
                            self.opt.hyper, self.opt.device_pows(), self._G, self._S0, self._S1, self._tU, self._tI,
                            self.opt.take_stamps(steps), self._ws[0], self._ws[1], self._ws[2], self._step_loss)
        self.opt.lr_t(steps)              # keep the host mirror of the beta powers in step
        return float(self._step_loss[:steps].sum().item())

    def predict(self, user_ids, candidate_items_user_ids=None):
        # NeuMF.py:158-168: one forward over all items per user (tower 0 = self.output)
        users = torch.as_tensor(np.asarray(user_ids, dtype=np.int32)).cuda()
        scores = ops.ncf_scores(self.shape, self.params, users)
        if candidate_items_user_ids is not None:
            s = scores.cpu().numpy()
            return [s[b][np.asarray(items)] for b, items in enumerate(candidate_items_user_ids)]
        return scores          # CUDA tensor [B, num_items]; UniEvaluator consumes it in place


class NeuMF(_NCFBase):
    def __init__(self, sess, dataset, conf):
        super(NeuMF, self).__init__(dataset, conf)
        self.embedding_size = conf["embedding_size"]
        self.reg_mf = conf["reg_mf"]
        self.reg_mlp = conf["reg_mlp"]
        self.mf_pretrain = conf["mf_pretrain"]
        self.mlp_pretrain = conf["mlp_pretrain"]
        self._common_init(sess, dataset, conf)

    def _n_towers(self):
        # pairwise NeuMF re-instantiates tf.layers.dense for the negative tower (NeuMF.py:81-82,90)
        return 2 if self.is_pairwise is True else 1

    def build_graph(self):
        try:                                                     # NeuMF.py:108-117
            pre = []
            with open(self.mf_pretrain, "rb") as fin:
                pre.append(pickle.load(fin, encoding="utf-8"))
            with open(self.mlp_pretrain, "rb") as fin:
                pre.append(pickle.load(fin, encoding="utf-8"))
            self.logger.info("load pretrained params successful!")
        except Exception:
            pre = None
            self.logger.info("load pretrained params unsuccessful!")
        self._build(self.embedding_size, self.reg_mf, self.reg_mlp, pre)

    def train_model(self):
        self.logger.info(self.evaluator.metrics_info())
        data_iter = self._make_sampler()
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
