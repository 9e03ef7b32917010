"""jax_cfd/base/time_stepping.py: forward RK time steppers for incompressible Navier-Stokes."""
from __future__ import annotations

import dataclasses
from typing import Callable, Sequence

import numpy as np

from . import _engine
from . import grids


class ExplicitNavierStokesODE:
  """time_stepping.py:26-45."""

  def __init__(self, explicit_terms, pressure_projection):
    self.explicit_terms = explicit_terms
    self.pressure_projection = pressure_projection


@dataclasses.dataclass
class ButcherTableau:
  """time_stepping.py:48-56."""
  a: Sequence[Sequence[float]]
  b: Sequence[float]

  def __post_init__(self):
    if len(self.a) + 1 != len(self.b):
      raise ValueError('inconsistent Butcher tableau')


def _is_native(equation) -> bool:
  return (isinstance(equation.explicit_terms, _engine.NativeExplicitTerms) and
          getattr(equation, 'native_projection', False))


def _tree_axpy(u0, ks, coefs):
  """u0 + sum coef*k for host pytrees (tuples of GridVariable / arrays / scalars)."""
  def comb(x, *ys):
    if isinstance(x, grids.GridVariable):
      data = x.data + sum(c * y.data for c, y in zip(coefs, ys))
      return grids.GridVariable(grids.GridArray(data, x.offset, x.grid), x.bc)
    if isinstance(x, grids.GridArray):
      return grids.GridArray(x.data + sum(c * y.data for c, y in zip(coefs, ys)), x.offset, x.grid)
    return x + sum(c * y for c, y in zip(coefs, ys))
  if isinstance(u0, (tuple, list)):
    return type(u0)(comb(x, *[k[i] for k in ks]) for i, x in enumerate(u0))
  return comb(u0, *ks)


def navier_stokes_rk(tableau: ButcherTableau, equation: ExplicitNavierStokesODE,
                     time_step: float) -> Callable:
  """time_stepping.py:59-106: the reference (non-fast-projection) RK scheme, projection after
  every stage.  With the native ODE of `equations.semi_implicit_navier_stokes` every F, P and
  stage combination is a CUDA kernel; a single-stage tableau collapses to the fused step."""
  dt = time_step
  F, P = equation.explicit_terms, equation.pressure_projection
  a, b = tableau.a, tableau.b
  num_steps = len(b)
  native = _is_native(equation)
  if native and num_steps == 1 and b[0] == 1:
    # one fused call per step; a stepper dt that differs from the builder's dt (which stays
    # bound inside the convection term) gets its own parameter block
    return equation.fused_step.with_time_step(dt)

  def combine(u0, ks, coefs):
    pairs = [(c, k) for c, k in zip(coefs, ks) if c]
    if native:
      return _engine.axpy(u0, [k for _, k in pairs], [dt * c for c, _ in pairs])
    return _tree_axpy(u0, [k for _, k in pairs], [dt * c for c, _ in pairs])

  def step_fn(u0):
    u = [None] * num_steps
    k = [None] * num_steps
    u[0] = u0
    k[0] = F(u0)
    for i in range(1, num_steps):
      u_star = combine(u0, k[:i], a[i - 1][:i])
      u[i] = P(u_star)
      k[i] = F(u[i])
    u_star = combine(u0, k, b)
    return P(u_star)

  return step_fn


def forward_euler(equation, time_step):
  """time_stepping.py:109-118."""
  return navier_stokes_rk(ButcherTableau(a=[], b=[1]), equation, time_step)


def midpoint_rk2(equation, time_step):
  """time_stepping.py:121-131."""
  return navier_stokes_rk(ButcherTableau(a=[[1 / 2]], b=[0, 1]), equation=equation,
                          time_step=time_step)


def heun_rk2(equation, time_step):
  """time_stepping.py:134-144."""
  return navier_stokes_rk(ButcherTableau(a=[[1]], b=[1 / 2, 1 / 2]), equation=equation,
                          time_step=time_step)


def classic_rk4(equation, time_step):
  """time_stepping.py:147-158."""
  return navier_stokes_rk(
      ButcherTableau(a=[[1 / 2], [0, 1 / 2], [0, 0, 1]], b=[1 / 6, 1 / 3, 1 / 3, 1 / 6]),
      equation=equation, time_step=time_step)
