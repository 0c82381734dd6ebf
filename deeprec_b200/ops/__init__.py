from . import embedding_ops  # noqa: F401
