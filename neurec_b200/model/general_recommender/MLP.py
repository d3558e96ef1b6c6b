"""MLP: plug-in mirror of the reference's model/general_recommender/MLP.py:15-142 (shared Dense
objects for the positive and negative towers; a new sampler every epoch, MLP.py:97-101)."""
from time import time

from .NeuMF import _NCFBase


class MLP(_NCFBase):
    def __init__(self, sess, dataset, conf):
        super(MLP, self).__init__(dataset, conf)
        self.reg_mlp = conf["reg_mlp"]
        self._common_init(sess, dataset, conf)

    def build_graph(self):
        self._build(0, 0.0, self.reg_mlp, None)

    def train_model(self):
        self.logger.info(self.evaluator.metrics_info())
        for epoch in range(1, self.num_epochs + 1):
            data_iter = self._make_sampler()                     # MLP.py:97-101
            start = time()
            total_loss = self._train_epoch(data_iter)
            self.logger.info("[iter %d : loss : %f, time: %f]" % (epoch, total_loss / len(data_iter),
                                                                  time() - start))
            if epoch % self.verbose == 0:
                self.logger.info("epoch %d:\t%s" % (epoch, self.evaluate()))

    def evaluate(self):
        return self.evaluator.evaluate(self)
