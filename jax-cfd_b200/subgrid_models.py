"""jax_cfd/base/subgrid_models.py: the explicit Smagorinsky builder (config #5)."""
from __future__ import annotations

from . import equations
from . import forcings
from ._engine import ForcingFn, SmagorinskyTerm


def explicit_smagorinsky_navier_stokes(dt, cs, forcing, **kwargs):
  """subgrid_models.py:188-213: the eddy-viscosity acceleration enters as the LAST forcing term."""
  smag = ForcingFn([SmagorinskyTerm(cs)])
  forcing = smag if forcing is None else forcings.sum_forcings(forcing, smag)
  return equations.semi_implicit_navier_stokes(dt=dt, forcing=forcing, **kwargs)
