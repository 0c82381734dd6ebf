from .saver import BundleReader, BundleWriter, IncrementalSaver, Saver, latest_checkpoint  # noqa: F401
