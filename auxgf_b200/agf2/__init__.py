from .optragf2 import OptRAGF2  # noqa: F401
from .optuagf2 import OptUAGF2  # noqa: F401
