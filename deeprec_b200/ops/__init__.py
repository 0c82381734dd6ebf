from . import embedding_ops  # noqa: F401
from . import sparse_ops  # noqa: F401,E402
