from .metrics import StreamingAccuracy, StreamingAUC  # noqa: F401
from .tracing import Timeline, nvtx_range  # noqa: F401
from .trainer import Trainer  # noqa: F401

from .streams import mark_target_node, stream, target_nodes  # noqa: E402,F401
from .health import FaultInjector, HeartbeatMonitor, InjectedFault, StepWatchdog  # noqa: E402,F401
from .summary import SummaryHook  # noqa: E402,F401
