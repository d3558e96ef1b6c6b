"""Host-side mirror of the reference's `evaluator` package; the backend is the CUDA library
(the reference picks cpp or python at import, evaluator/backend/__init__.py:1-6)."""
from .proxy_evaluator import ProxyEvaluator
from .grouped_evaluator import GroupedEvaluator
from .uni_evaluator import UniEvaluator
