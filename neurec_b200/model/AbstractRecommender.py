"""AbstractRecommender: the model plug-in base class (reference model/AbstractRecommender.py:9-45).
Sequential / social bases are outside the hot path."""
import os
import time

from ..evaluator import ProxyEvaluator
from ..util.logger import Logger

# ProxyEvaluator keyword <- configuration key (AbstractRecommender.py:25-32)
_EVALUATOR_OPTIONS = (("metric", "metric"), ("group_view", "group_view"), ("top_k", "topk"),
                      ("batch_size", "test_batch_size"), ("num_thread", "num_thread"))


def _create_logger(config, data_name):
    # AbstractRecommender.py:9-20: log/<dataset>/<model>/<dataset>_<params[:150]>_<ts>.log
    run_id = "%s_%.8f" % (("%s_%s" % (data_name, config.params_str()))[:150], time.time())
    return Logger(os.path.join("log", data_name, config["recommender"], run_id + ".log"))


class AbstractRecommender(object):
    """Every model gets its evaluator (train / test / negative-test dicts of the dataset + the
    evaluation options of NeuRec.properties) and its logger, and logs dataset and configuration."""

    def __init__(self, dataset, conf):
        splits = (dataset.get_user_train_dict(), dataset.get_user_test_dict(), dataset.get_user_test_neg_dict())
        options = {kw: conf[key] for kw, key in _EVALUATOR_OPTIONS}
        self.evaluator = ProxyEvaluator(*splits, **options)
        self.logger = _create_logger(conf, dataset.dataset_name)
        for what in (dataset, conf):
            self.logger.info(what)

    def build_graph(self):
        raise NotImplementedError

    def train_model(self):
        raise NotImplementedError

    def predict(self, user_ids, items):
        raise NotImplementedError
