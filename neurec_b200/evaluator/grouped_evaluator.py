"""GroupedEvaluator: ranking metrics per user group, groups = ranges of the number of TRAINING
interactions (reference evaluator/grouped_evaluator.py:23-112)."""
from collections import OrderedDict

import numpy as np

from ..util.tool import typeassert
from .abstract_evaluator import AbstractEvaluator
from .uni_evaluator import UniEvaluator


class GroupedEvaluator(AbstractEvaluator):
    @typeassert(user_train_dict=dict, user_test_dict=dict, group_view=list)
    def __init__(self, user_train_dict, user_test_dict, user_neg_test=None, metric=None,
                 group_view=None, top_k=50, batch_size=1024, num_thread=8):
        super(GroupedEvaluator, self).__init__()
        if not isinstance(group_view, list):
            raise TypeError("The type of 'group_view' must be `list`!")
        self.evaluator = UniEvaluator(user_train_dict, user_test_dict, user_neg_test, metric=metric,
                                      top_k=top_k, batch_size=batch_size, num_thread=num_thread)
        self.user_pos_train, self.user_pos_test = user_train_dict, user_test_dict
        edges = [0] + group_view
        info = [("(%d,%d]:" % (lo, hi)).ljust(12) for lo, hi in zip(edges[:-1], edges[1:])]
        users = list(user_test_dict.keys())
        n_train = [len(user_train_dict[u]) for u in users]
        # grouped_evaluator.py:67: bucket (lo, hi]; users beyond the last edge are dropped (:75-77)
        group_idx = np.searchsorted(edges[1:], n_train)
        self.grouped_user = OrderedDict()
        for g in sorted(set(group_idx.tolist())):
            if g < len(info):
                self.grouped_user[info[g]] = [u for u, gi in zip(users, group_idx) if gi == g]
        if not self.grouped_user:
            raise ValueError("The splitting of user groups is not suitable!")

    def metrics_info(self):
        return self.evaluator.metrics_info()

    def evaluate(self, model):
        out = ""
        for group, users in self.grouped_user.items():
            out = "%s\n%s\t%s" % (out, group, self.evaluator.evaluate(model, users))
        return out
