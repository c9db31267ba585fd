from . import losses, penalties  # noqa: F401
from .function import Function  # noqa: F401
