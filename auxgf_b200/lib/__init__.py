from . import agf2  # noqa: F401
