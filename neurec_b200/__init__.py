"""neurec_b200 -- B200-native (sm_100a) hot path of NeuRec behind the reference's plug-in surface.

``neurec_b200.ops`` are the tensor-facing wrappers of the C ABI (include/neurec_b200.h);
``neurec_b200.util / data / evaluator / model`` mirror the reference's Python interface
(Configurator, DataIterator, samplers, ProxyEvaluator, AbstractRecommender, MF/MLP/NeuMF/
LightGCN) on top of those kernels.  There is no CPU implementation in this package.
"""
__version__ = "0.1.0"
