"""Top-level alias of cold_compress_b200.generation_utils, so that the reference's own scripts (`from generation_utils import ...`)
resolve against this package when this directory comes first on sys.path."""
from cold_compress_b200.generation_utils import *  # noqa: F401,F403
from cold_compress_b200 import generation_utils as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
