"""Dataset: load -> filter -> split -> id remap -> CSR, with the reference's on-disk cache.

Mirror of the reference's data/dataset.py:16-289 and data/utils.py:11-105 (same configuration
keys, file formats, `_tmp_<dataset>/` cache with md5 guard, same public attributes and getters).
One-time host preprocessing (pandas / scipy), not part of the accelerated path: it exists so
that the engine and the oracle see byte-identical splits.
"""
import hashlib
import math
import os

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix

from ..util.logger import Logger
from ..util.tool import csr_to_user_dict, csr_to_user_dict_bytime, randint_choice

_FORMATS = {"UIRT": ["user", "item", "rating", "time"], "UIR": ["user", "item", "rating"],
            "UI": ["user", "item"], "UIT": ["user", "item", "time"]}


def check_md5(file_name):
    if not os.path.isfile(file_name):
        raise FileNotFoundError("There is not file named '%s'!" % file_name)
    with open(file_name, "rb") as fin:
        return hashlib.md5(fin.read()).hexdigest()


def filter_data(data, user_min=None, item_min=None):
    """data/utils.py:27-39."""
    data.dropna(how="any", inplace=True)
    if item_min is not None and item_min > 0:
        cnt = data["item"].value_counts(sort=False)
        data = data[data["item"].map(lambda x: cnt[x] >= item_min)]
    if user_min is not None and user_min > 0:
        cnt = data["user"].value_counts(sort=False)
        data = data[data["user"].map(lambda x: cnt[x] >= user_min)]
    return data


def _split(data, by_time, cut):
    data.sort_values(by=["user", "time" if by_time else "item"], inplace=True)
    first, second = [], []
    for _, u_data in data.groupby(by=["user"]):
        a, b = cut(u_data, by_time)
        first.append(a)
        if b is not None:
            second.append(b)
    return pd.concat(first, ignore_index=True), pd.concat(second, ignore_index=True)


def split_by_ratio(data, ratio=0.8, by_time=True):
    """data/utils.py:63-82: per user, shuffle (unless by_time) and cut at ceil(ratio * n)."""
    def cut(u, by_time):
        if not by_time:
            u = u.sample(frac=1)
        idx = math.ceil(ratio * len(u))
        return u.iloc[:idx], u.iloc[idx:]
    return _split(data, by_time, cut)


def split_by_loo(data, by_time=True):
    """data/utils.py:85-105: last (or a random) interaction of users with > 3 interactions."""
    def cut(u, by_time):
        if len(u) <= 3:
            return u, None
        if not by_time:
            u = u.sample(frac=1)
        return u.iloc[:-1], u.iloc[-1:]
    return _split(data, by_time, cut)


class Dataset(object):
    def __init__(self, conf):
        self.train_matrix = self.test_matrix = self.time_matrix = self.negative_matrix = None
        self.userids = self.itemids = None
        self.num_users = self.num_items = None
        self.dataset_name = conf["data.input.dataset"]
        self._load_data(conf)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_csr(cls, name, train_matrix, test_matrix, time_matrix=None, negative_matrix=None):
        """Build a Dataset from ready-made CSR matrices (tests, benchmarks)."""
        self = cls.__new__(cls)
        self.dataset_name = name
        self.train_matrix, self.test_matrix = train_matrix.tocsr(), test_matrix.tocsr()
        self.time_matrix, self.negative_matrix = time_matrix, negative_matrix
        self.num_users, self.num_items = self.train_matrix.shape
        self.num_ratings = int(self.train_matrix.nnz + self.test_matrix.nnz)
        self.userids = {u: u for u in range(self.num_users)}
        self.itemids = {i: i for i in range(self.num_items)}
        return self

    def _get_data_path(self, config):
        path = config["data.input.path"]
        ori_prefix = os.path.join(path, self.dataset_name)
        saved_prefix = "%s_%s_u%d_i%d" % (self.dataset_name, config["splitter"], config["user_min"],
                                          config["item_min"])
        if "by_time" in config and config["by_time"] is True:
            saved_prefix += "_by_time"
        return ori_prefix, os.path.join(path, "_tmp_" + self.dataset_name, saved_prefix)

    @staticmethod
    def _check_saved_data(splitter, ori_prefix, saved_prefix):
        if splitter in ("loo", "ratio"):
            ori_md5 = [check_md5(ori_prefix + ".rating")]
        elif splitter == "given":
            ori_md5 = [check_md5(ori_prefix + ".train"), check_md5(ori_prefix + ".test")]
        else:
            raise ValueError("'%s' is an invalid splitter!" % splitter)
        ok = False
        if os.path.isfile(saved_prefix + ".md5"):
            with open(saved_prefix + ".md5") as fin:
                ok = [line.strip() for line in fin.readlines()] == ori_md5
        return ok and all(os.path.isfile(saved_prefix + ext)
                          for ext in (".train", ".test", ".user2id", ".item2id"))

    def _load_data(self, config):
        file_format = config["data.column.format"]
        if file_format not in _FORMATS:
            raise ValueError("'%s' is an invalid data column format!" % file_format)
        ori_prefix, saved_prefix = self._get_data_path(config)
        splitter, sep, columns = config["splitter"], config["data.convert.separator"], _FORMATS[file_format]
        if self._check_saved_data(splitter, ori_prefix, saved_prefix):
            print("load saved data...")
            read = lambda ext, names: pd.read_csv(saved_prefix + ext, sep=sep, header=None, names=names)
            train_data, test_data = read(".train", columns), read(".test", columns)
            um, im = read(".user2id", ["user", "id"]), read(".item2id", ["item", "id"])
            self.userids = dict(zip(um["user"], um["id"]))
            self.itemids = dict(zip(im["item"], im["id"]))
        else:
            print("split and save data...")
            by_time = config["by_time"] if file_format in {"UIRT", "UIT"} else False
            train_data, test_data = self._split_data(ori_prefix, saved_prefix, columns, by_time, config)
        all_data = pd.concat([train_data, test_data])
        self.num_users = int(max(all_data["user"])) + 1
        self.num_items = int(max(all_data["item"])) + 1
        self.num_ratings = len(all_data)
        if file_format in {"UI", "UIT"}:
            tr_r, te_r = [1.0] * len(train_data), [1.0] * len(test_data)
        else:
            tr_r, te_r = train_data["rating"], test_data["rating"]
        shape = (self.num_users, self.num_items)
        self.train_matrix = csr_matrix((tr_r, (train_data["user"], train_data["item"])), shape=shape)
        self.test_matrix = csr_matrix((te_r, (test_data["user"], test_data["item"])), shape=shape)
        if file_format in {"UIRT", "UIT"}:
            self.time_matrix = csr_matrix((train_data["time"], (train_data["user"], train_data["item"])),
                                          shape=shape)
        self.negative_matrix = self._load_test_neg_items(all_data, config, saved_prefix)

    def _split_data(self, ori_prefix, saved_prefix, columns, by_time, config):
        splitter, sep = config["splitter"], config["data.convert.separator"]
        os.makedirs(os.path.dirname(saved_prefix), exist_ok=True)
        if splitter in ("loo", "ratio"):
            rating_file = ori_prefix + ".rating"
            data = filter_data(pd.read_csv(rating_file, sep=sep, header=None, names=columns),
                               user_min=config["user_min"], item_min=config["item_min"])
            if splitter == "ratio":
                train_data, test_data = split_by_ratio(data, ratio=config["ratio"], by_time=by_time)
            else:
                train_data, test_data = split_by_loo(data, by_time=by_time)
            md5 = [check_md5(rating_file)]
        elif splitter == "given":
            train_data = pd.read_csv(ori_prefix + ".train", sep=sep, header=None, names=columns)
            test_data = pd.read_csv(ori_prefix + ".test", sep=sep, header=None, names=columns)
            md5 = [check_md5(ori_prefix + ".train"), check_md5(ori_prefix + ".test")]
        else:
            raise ValueError("'%s' is an invalid splitter!" % splitter)
        with open(saved_prefix + ".md5", "w") as out:
            out.writelines("\n".join(md5))
        all_data = pd.concat([train_data, test_data])
        uu, ui = all_data["user"].unique(), all_data["item"].unique()
        self.userids = pd.Series(data=range(len(uu)), index=uu).to_dict()
        self.itemids = pd.Series(data=range(len(ui)), index=ui).to_dict()
        for frame in (train_data, test_data):
            frame["user"] = frame["user"].map(self.userids)
            frame["item"] = frame["item"].map(self.itemids)
        np.savetxt(saved_prefix + ".train", train_data, fmt="%d", delimiter=sep)
        np.savetxt(saved_prefix + ".test", test_data, fmt="%d", delimiter=sep)
        np.savetxt(saved_prefix + ".user2id", [[k, v] for k, v in self.userids.items()], fmt="%s", delimiter=sep)
        np.savetxt(saved_prefix + ".item2id", [[k, v] for k, v in self.itemids.items()], fmt="%s", delimiter=sep)
        neg_file = ori_prefix + ".neg"
        if os.path.isfile(neg_file):
            rows = []
            with open(neg_file) as fin:
                for line in fin.readlines():
                    parts = line.strip().split(sep)
                    rows.append([self.userids[parts[0]]] + [self.itemids[i] for i in parts[1:]])
            np.savetxt("%s.neg%d" % (saved_prefix, len(rows[0]) - 1), rows, fmt="%d", delimiter=sep)
        remapped = pd.concat([train_data, test_data])
        self.num_users = int(max(remapped["user"])) + 1
        self.num_items = int(max(remapped["item"])) + 1
        self.num_ratings = len(remapped)
        logger = Logger(saved_prefix + ".info")
        logger.info(os.path.basename(saved_prefix))
        logger.info(self.__str__())
        return train_data, test_data

    def _load_test_neg_items(self, all_data, config, saved_prefix):
        number_neg, sep = config["rec.evaluate.neg"], config["data.convert.separator"]
        if not number_neg or number_neg <= 0:
            return None
        neg_file = "%s.neg%d" % (saved_prefix, number_neg)
        if not os.path.isfile(neg_file):
            rows = []
            for user, u_data in all_data.groupby("user"):
                rows.append([user] + list(randint_choice(self.num_items, size=number_neg, replace=False,
                                                         exclusion=u_data["item"].tolist())))
            neg_items = pd.DataFrame(rows)
            np.savetxt(neg_file, neg_items, fmt="%d", delimiter=sep)
        else:
            neg_items = pd.read_csv(neg_file, sep=sep, header=None)
        users, items = [], []
        for line in neg_items.values:
            users.extend([line[0]] * (len(line) - 1))
            items.extend(line[1:])
        return csr_matrix(([1] * len(users), (users, items)), shape=(self.num_users, self.num_items))

    # ------------------------------------------------------------------ reference getters
    def __str__(self):
        nu, ni, nr = self.num_users, self.num_items, self.num_ratings
        sparsity = 1 - 1.0 * nr / (nu * ni)
        return "\n".join(["Dataset name: %s" % self.dataset_name,
                          "The number of users: %d" % nu,
                          "The number of items: %d" % ni,
                          "The number of ratings: %d" % nr,
                          "Average actions of users: %.2f" % (1.0 * nr / nu),
                          "Average actions of items: %.2f" % (1.0 * nr / ni),
                          "The sparsity of the dataset: %.6f%%" % (sparsity * 100)])

    __repr__ = __str__

    def get_user_train_dict(self, by_time=False):
        if by_time:
            return csr_to_user_dict_bytime(self.time_matrix, self.train_matrix)
        return csr_to_user_dict(self.train_matrix)

    def get_user_test_dict(self):
        return csr_to_user_dict(self.test_matrix)

    def get_user_test_neg_dict(self):
        return None if self.negative_matrix is None else csr_to_user_dict(self.negative_matrix)

    def get_train_interactions(self):
        coo = self.train_matrix.tocoo()
        return coo.row.tolist(), coo.col.tolist()

    def to_csr_matrix(self):
        return self.train_matrix.copy()
