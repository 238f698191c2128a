from . import resource  # noqa: F401
