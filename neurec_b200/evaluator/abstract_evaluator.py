"""Base class of all evaluators (reference evaluator/abstract_evaluator.py:4-35)."""


class AbstractEvaluator(object):
    def __init__(self):
        pass

    def metrics_info(self):
        raise NotImplementedError

    def evaluate(self, model):
        raise NotImplementedError
