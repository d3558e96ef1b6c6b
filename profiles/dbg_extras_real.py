"""The SURVEY 8(f) rank-3 plug-ins on their real datasets (from the committed golden splits): epoch time and ranking
quality over a few epochs.  SBPR on Ciao (6 596 users, 107 320 items, 221 734 train interactions, 113 530 trust pairs),
APR and SpectralCF on ml-100k.  Orientation numbers for DESIGN.md 3c, not BASELINE metrics."""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir("/tmp")
from neurec_b200.data import Dataset, PairwiseSampler  # noqa: E402
import importlib  # noqa: E402
AR = importlib.import_module("neurec_b200.model.AbstractRecommender")

GOLDEN = os.path.join(ROOT, "tests", "golden")


class Conf(dict):
    def params_str(self):
        return "dbg"


BASE = {"metric": ["Precision", "Recall", "NDCG", "MAP", "MRR"], "group_view": None, "topk": [10, 20], "test_batch_size": 128,
        "num_thread": 8, "data.convert.separator": "\t"}


def dataset(npz, name):
    z = np.load(os.path.join(GOLDEN, npz))
    shape = (int(z["num_users"]), int(z["num_items"]))
    mk = lambda p, i: sp.csr_matrix((np.ones(len(z[i]), np.float32), z[i].astype(np.int32), z[p].astype(np.int64)), shape=shape)
    return Dataset.from_csr(name, mk("train_indptr", "train_indices"), mk("test_indptr", "test_indices")), z


def ndcg10(model):
    s = model.evaluate()                # the model's own evaluate(): SpectralCF / NGCF propagate first
    return float(s.split()[4])          # Precision@10 @20 Recall@10 @20 NDCG@10 ...


def run(name, model, epoch_fn, epochs):
    model.build_graph()
    print("%s: NDCG@10 before training %.5f" % (name, ndcg10(model)), flush=True)
    for e in range(epochs):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        loss = epoch_fn()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if e in (0, epochs // 2, epochs - 1):
            print("%s: epoch %d  %.1f ms  loss %.4f  NDCG@10 %.5f" % (name, e + 1, dt * 1e3, loss, ndcg10(model)), flush=True)


ONLY = sys.argv[1] if len(sys.argv) > 1 else ""

# ---- SBPR on Ciao
from neurec_b200.model.social_recommender.SBPR import SBPR  # noqa: E402
ds, z = dataset("ciao_split.npz", "Ciao_u5_s2")
trust = sp.csr_matrix((np.ones(len(z["trust_indices"]), np.int64), z["trust_indices"].astype(np.int32), z["trust_indptr"].astype(np.int64)),
                      shape=(ds.num_users, ds.num_users))
orig = AR.SocialAbstractRecommender.__init__


def social_init(self, dataset, conf):            # the trust CSR comes from the golden file instead of conf["social_file"]
    AR.AbstractRecommender.__init__(self, dataset, conf)
    self.social_matrix = trust


AR.SocialAbstractRecommender.__init__ = social_init
conf = Conf(BASE, recommender="SBPR", learning_rate=0.001, embedding_size=16, learner="adam", loss_function="bpr", num_epochs=1,
            reg_mf=0.01, batch_size=512, init_method="normal", stddev=0.01, verbose=1)
m = SBPR(None, ds, conf)
if ONLY in ("", "sbpr"):
    run("SBPR / Ciao (%d samples per epoch, %d steps of 512)" % (m._n, (m._n + 511) // 512), m, lambda: m._train_epoch() / m._n, 12)
AR.SocialAbstractRecommender.__init__ = orig

# ---- APR on ml-100k
from neurec_b200.model.general_recommender.APR import APR  # noqa: E402
ds, _ = dataset("ml100k_split.npz", "ml-100k")
conf = Conf(BASE, recommender="APR", learning_rate=0.001, embedding_size=64, learner="adam", epochs=1, eps=0.5, adv="grad", adver=1,
            adv_epoch=0, reg=0.0, reg_adv=1.0, batch_size=512, init_method="tnormal", stddev=0.01, verbose=1)
m = APR(None, ds, conf)
it = PairwiseSampler(ds, neg_num=1, batch_size=512, shuffle=True)
if ONLY in ("", "apr"):
    run("APR / ml-100k (157 steps of 512)", m, lambda: m._train_epoch(it) / len(it), 30)

# ---- SpectralCF on ml-100k (the operator is a 2 625 x 2 625 eigendecomposition on the host, once)
from neurec_b200.model.general_recommender.SpectralCF import SpectralCF  # noqa: E402
conf = Conf(BASE, recommender="SpectralCF", learning_rate=0.001, learner="adam", batch_size=256, num_layers=2, activation="sigmoid",
            embedding_size=100, epochs=1, reg=0.001, loss_function="BPR", dropout=0.0, embed_init_method="xavier_normal",
            weight_init_method="xavier_normal", stddev=0.01, verbose=1)
if ONLY not in ("", "spectral"):
    sys.exit(0)
t0 = time.perf_counter()
m = SpectralCF(None, ds, conf)
print("SpectralCF: operator built on the host in %.1f s" % (time.perf_counter() - t0), flush=True)
it = PairwiseSampler(ds, neg_num=1, batch_size=256, shuffle=True)


def spectral_epoch():
    users, pos, neg = it.device_epoch()
    loss = torch.zeros(1, device="cuda")
    for s in range(len(it)):
        sl = slice(s * 256, (s + 1) * 256)
        m._step(users[sl], pos[sl], neg[sl], loss)
    return float(loss.item()) / len(it)


run("SpectralCF / ml-100k (314 steps of 256)", m, spectral_epoch, 6)
