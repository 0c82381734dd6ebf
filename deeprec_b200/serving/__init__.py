from .export import export_delta, export_saved_model  # noqa: F401
from .processor import Processor, ProcessorGroup, decode_response, encode_request  # noqa: F401
from .session_group import SessionGroup  # noqa: F401
