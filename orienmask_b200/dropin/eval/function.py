"""Drop-in for ``eval.function`` (the builder resolves ``batched_nms`` by name in this module)."""
from orienmask_b200.function import batched_nms, nms  # noqa: F401
