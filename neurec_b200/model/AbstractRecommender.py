"""AbstractRecommender: the model plug-in base class (reference model/AbstractRecommender.py:9-45).
Sequential / social bases are outside the hot path."""
import os
import time

from ..evaluator import ProxyEvaluator
from ..util.logger import Logger


def _create_logger(config, data_name):
    # AbstractRecommender.py:9-20: log/<dataset>/<model>/<dataset>_<params[:150]>_<ts>.log
    param_str = "%s_%s" % (data_name, config.params_str())
    run_id = "%s_%.8f" % (param_str[:150], time.time())
    log_dir = os.path.join("log", data_name, config["recommender"])
    return Logger(os.path.join(log_dir, run_id + ".log"))


class AbstractRecommender(object):
    def __init__(self, dataset, conf):
        self.evaluator = ProxyEvaluator(dataset.get_user_train_dict(),
                                        dataset.get_user_test_dict(),
                                        dataset.get_user_test_neg_dict(),
                                        metric=conf["metric"],
                                        group_view=conf["group_view"],
                                        top_k=conf["topk"],
                                        batch_size=conf["test_batch_size"],
                                        num_thread=conf["num_thread"])
        self.logger = _create_logger(conf, dataset.dataset_name)
        self.logger.info(dataset)
        self.logger.info(conf)

    def build_graph(self):
        raise NotImplementedError

    def train_model(self):
        raise NotImplementedError

    def predict(self, user_ids, items):
        raise NotImplementedError
