"""B200-native batched racing-line evaluator (periodic cubic fit -> resample with heading / turn radius / bank
-> exact-schedule QSS speed profile -> lap time) behind the reference's Python API.  See DESIGN.md."""
from ._lib import StoError, QSS_MEMO, QSS_PLAIN  # noqa: F401

__all__ = ["StoError", "QSS_MEMO", "QSS_PLAIN", "BatchedLineEvaluator"]


def __getattr__(name):   # torch is imported lazily: the host mirrors work without it
    if name == "BatchedLineEvaluator":
        from .evaluator import BatchedLineEvaluator
        return BatchedLineEvaluator
    raise AttributeError(name)
