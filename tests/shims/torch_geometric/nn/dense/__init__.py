from . import linear  # noqa: F401
