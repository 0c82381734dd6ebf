from .export import (export_delta, export_delta_module, export_delta_program, export_saved_model, export_saved_model_module,  # noqa: F401
                     export_saved_model_program, export_zoo_model,
                     load_zoo_model)
from .processor import Processor, ProcessorGroup, decode_response, encode_request  # noqa: F401
from .session_group import SessionGroup  # noqa: F401
