from .aux import Aux  # noqa: F401
from . import energy  # noqa: F401
from .energy import energy_2body_aux, energy_mp2_aux  # noqa: F401
