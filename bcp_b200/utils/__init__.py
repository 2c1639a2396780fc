from . import BCP_utils, losses  # noqa: F401
