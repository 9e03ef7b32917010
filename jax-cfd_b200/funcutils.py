"""jax_cfd/base/funcutils.py: repeated / trajectory, composing with the native step."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

from . import _engine
from . import _lib


def repeated(f: Callable, steps: int) -> Callable:
  """funcutils.py:82-88.  A native step is advanced `steps` times inside one C call."""
  if isinstance(f, _engine.NativeStep):
    return lambda v: f.advance(v, steps)

  def f_repeated(x):
    for _ in range(steps):
      x = f(x)
    return x
  return f_repeated


def _stack(frames):
  first = frames[0]
  if isinstance(first, (tuple, list)):
    return type(first)(_stack([fr[i] for fr in frames]) for i in range(len(first)))
  from . import grids
  if isinstance(first, grids.GridVariable):
    return grids.GridVariable(_stack([fr.array for fr in frames]), first.bc)
  if isinstance(first, grids.GridArray):
    return grids.GridArray(_stack([fr.data for fr in frames]), first.offset, first.grid)
  if isinstance(first, _lib.DeviceArray):
    # frames stay on the device: one (steps, *shape) array filled by device-to-device copies that
    # are ordered on the stream the steps run on -- no host round trip, no synchronisation per frame
    # (the reference's scan stacks its outputs on the device the same way, funcutils.py:65-79)
    out = _lib.DeviceArray((len(frames),) + first.shape, first.dtype, device=first.device)
    for k, fr in enumerate(frames):
      _lib.check(_lib.lib().cfd_memcpy_d2d(out.ptr + k * first.nbytes, fr.ptr, first.nbytes, None))
    return out
  if _lib.is_device_array(first):
    if hasattr(first, 'new_empty'):  # torch CUDA tensors: stacked by their own library, on the device
      import torch
      return torch.stack(list(frames))
    return np.stack([np.asarray(fr.cpu().numpy()) for fr in frames])
  return np.stack([np.asarray(fr) for fr in frames])


def trajectory(step_fn: Callable, steps: int, post_process: Callable = lambda x: x, *,
               start_with_input: bool = False) -> Callable:
  """funcutils.py:95-126: returns (final_state, stacked post-processed frames).  Device-resident
  frames are stacked on the device (no host copy inside the loop); host frames with NumPy."""
  def multistep(values):
    frames = []
    x = values
    for _ in range(steps):
      nxt = step_fn(x)
      frames.append(post_process(x) if start_with_input else post_process(nxt))
      x = nxt
    return x, _stack(frames)
  return multistep
