from .metrics import StreamingAccuracy, StreamingAUC  # noqa: F401
from .tracing import Timeline, nvtx_range  # noqa: F401
from .trainer import Trainer  # noqa: F401
