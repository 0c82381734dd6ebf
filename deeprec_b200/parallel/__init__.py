from . import strategy  # noqa: F401
from .collective import CollectiveStrategy, DistStrategy, enable_distributed_strategy  # noqa: F401
