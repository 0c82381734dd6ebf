from . import strategy  # noqa: F401
