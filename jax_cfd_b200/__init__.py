"""Import shim: the package sources live in `jax-cfd_b200/` (a hyphen cannot be imported)."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 'jax-cfd_b200'))

from . import _lib  # noqa: E402,F401
from . import grids  # noqa: E402,F401
from . import boundaries  # noqa: E402,F401
from . import array_utils  # noqa: E402,F401
from . import fast_diagonalization  # noqa: E402,F401
from . import filter_utils  # noqa: E402,F401
from . import advection  # noqa: E402,F401
from . import diffusion  # noqa: E402,F401
from . import forcings  # noqa: E402,F401
from . import pressure  # noqa: E402,F401
from . import time_stepping  # noqa: E402,F401
from . import equations  # noqa: E402,F401
from . import subgrid_models  # noqa: E402,F401
from . import funcutils  # noqa: E402,F401
from . import initial_conditions  # noqa: E402,F401
from . import resize  # noqa: E402,F401
from . import distributed  # noqa: E402,F401
from ._engine import diagnostics, get_plan, clear_plans  # noqa: E402,F401
from ._lib import DeviceArray, CfdError  # noqa: E402,F401
