"""Slab-decomposed multi-GPU stepping (one process per GPU), SURVEY.md section 8(e).

The reference has no distributed solver; this module is the host side of the new path.  Each rank
owns `grid.shape[0] / world` rows of every field.  All data movement between GPUs happens inside
the CUDA kernels over NVLink (csrc/multi_gpu.cu); the host only exchanges one 64-byte CUDA-IPC
handle per rank at start-up -- `exchange` is any all-gather of bytes (default:
torch.distributed.all_gather_object, plumbing only).
"""
from __future__ import annotations

import ctypes
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _engine
from . import _lib
from . import grids
from ._lib import check, lib


def slab_rows(nx: int, rank: int, world: int) -> Tuple[int, int]:
  """[start, stop) of the rows owned by `rank` (axis 0 split evenly)."""
  if nx % world:
    raise ValueError(f'axis 0 ({nx}) is not divisible by the number of ranks ({world})')
  n = nx // world
  return rank * n, (rank + 1) * n


def line_range(ny: int, rank: int, world: int) -> Tuple[int, int]:
  """[start, stop) of the packed ky lines (Ny/2 of them) whose x-direction transform `rank` owns."""
  my = ny // 2
  if my % world:
    raise ValueError(f'Ny/2 ({my}) is not divisible by the number of ranks ({world})')
  n = my // world
  return rank * n, (rank + 1) * n


def check_decomposition(shape: Sequence[int], world: int) -> None:
  """The constraints cfd_dist_plan_create[_nd] enforces, checked on the host (same messages)."""
  if world not in (1, 2, 4, 8):
    raise ValueError('world size must be 1, 2, 4 or 8')
  if len(shape) == 3:
    n0, n1, n2 = shape
    nloc = n0 // world
    if n0 % world or nloc < 16 or nloc & (nloc - 1):
      raise ValueError('local slab must be a power of two >= 16 planes')
    if n1 % 8 or n2 % 64:
      raise ValueError('slab-decomposed 3-D grids need N1 % 8 == 0 and N2 % 64 == 0')
    if n1 % (16 * world):
      raise ValueError('axis 1 must be divisible by 16 x the number of ranks')
    if n0 > (1 << 14):
      raise NotImplementedError('global axis 0 longer than 16384 is not supported in 3-D')
    return
  nx, ny = shape
  nloc = nx // world
  if nx % world or nloc < 16 or nloc & (nloc - 1):
    raise ValueError('local slab must be a power of two >= 16 rows')
  if (ny // 2) % world or (ny // 2 // world) % 16:
    raise ValueError('Ny/2 lines must split evenly over the ranks')
  if nx > (1 << 15):
    raise NotImplementedError('global axis 0 longer than 32768 is not supported')


def torch_exchange(blob: bytes) -> List[bytes]:
  import torch.distributed as dist  # plumbing: bootstrap only
  out = [None] * dist.get_world_size()
  dist.all_gather_object(out, blob)
  return out


def tcp_exchange(rank: int, world: int, addr: str = '127.0.0.1', port: int = 29655,
                 timeout: float = 120.0) -> Callable[[bytes], List[bytes]]:
  """A torch-free all-gather of small byte strings over TCP (standard library only): rank 0
  listens on (addr, port), collects one blob per rank and sends every rank the full list.  All the
  host-side communication the slab-decomposed step ever needs is one such exchange of 64-byte
  CUDA-IPC handles at start-up."""
  import socket
  import struct
  import time

  def recv_exact(sock, n):
    buf = b''
    while len(buf) < n:
      part = sock.recv(n - len(buf))
      if not part:
        raise ConnectionError('peer closed the connection during the handle exchange')
      buf += part
    return buf

  def exchange(blob: bytes) -> List[bytes]:
    if world == 1:
      return [blob]
    if rank == 0:
      srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
      srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
      srv.bind((addr, port))
      srv.listen(world)
      srv.settimeout(timeout)
      blobs, conns = {0: blob}, []
      try:
        while len(blobs) < world:
          c, _ = srv.accept()
          c.settimeout(timeout)
          r, n = struct.unpack('!ii', recv_exact(c, 8))
          blobs[r] = recv_exact(c, n)
          conns.append(c)
        out = [blobs[r] for r in range(world)]
        payload = b''.join(struct.pack('!i', len(b)) + b for b in out)
        for c in conns:
          c.sendall(payload)
      finally:
        for c in conns:
          c.close()
        srv.close()
      return out
    deadline = time.time() + timeout
    while True:
      try:
        c = socket.create_connection((addr, port), timeout=timeout)
        break
      except OSError:
        if time.time() > deadline:
          raise
        time.sleep(0.05)
    try:
      c.sendall(struct.pack('!ii', rank, len(blob)) + blob)
      out = []
      for _ in range(world):
        (n,) = struct.unpack('!i', recv_exact(c, 4))
        out.append(recv_exact(c, n))
    finally:
      c.close()
    return out

  return exchange


def default_exchange(rank: int, world: int) -> Callable[[bytes], List[bytes]]:
  """torch.distributed when the caller has already initialised it (plumbing only), else the
  torch-free TCP rendezvous on MASTER_ADDR / MASTER_PORT + 1."""
  import os
  import sys
  td = sys.modules.get('torch.distributed')
  if td is not None and td.is_available() and td.is_initialized():
    return torch_exchange
  return tcp_exchange(rank, world, os.environ.get('MASTER_ADDR', '127.0.0.1'),
                      int(os.environ.get('MASTER_PORT', '29654')) + 1)


class SlabStepper:
  """Distributed counterpart of `funcutils.repeated(equations.semi_implicit_navier_stokes(...))`."""

  def __init__(self, grid: grids.Grid, dt: float, density: float, viscosity: Optional[float],
               forcing=None, *, rank: int, world: int, device: int,
               exchange: Optional[Callable[[bytes], List[bytes]]] = None):
    if grid.ndim not in (2, 3):
      raise NotImplementedError('slab decomposition is implemented for 2-D and 3-D grids')
    check_decomposition(grid.shape, world)
    _lib.require_device()
    if exchange is None:
      exchange = default_exchange(rank, world)
    self._exchange = exchange
    self.grid, self.rank, self.world, self.device = grid, rank, world, device
    self.rows = slab_rows(grid.shape[0], rank, world)
    self.local_shape = (self.rows[1] - self.rows[0],) + tuple(grid.shape[1:])
    check(lib().cfd_set_device(device))
    nd = grid.ndim
    shape = (ctypes.c_int64 * nd)(*grid.shape)
    step = (ctypes.c_double * nd)(*grid.step)
    self.handle = ctypes.c_void_p()
    check(lib().cfd_dist_plan_create_nd(ctypes.byref(self.handle), nd, shape, step, rank, world, device))
    nb = lib().cfd_dist_handle_bytes()
    blob = ctypes.create_string_buffer(nb)
    check(lib().cfd_dist_export(self.handle, blob))
    blobs = exchange(bytes(blob.raw))
    assert len(blobs) == world and all(len(b) == nb for b in blobs)
    allb = ctypes.create_string_buffer(b''.join(blobs), nb * world)
    check(lib().cfd_dist_connect(self.handle, allb))
    self.params, self._keep = _engine.make_params(grid, dt, density, viscosity,
                                                  _engine.as_forcing(forcing))
    self.stream = _lib.Stream()

  def _staging(self):
    if getattr(self, '_stage', None) is None:
      self._stage = [_lib.DeviceArray(self.local_shape) for _ in range(self.grid.ndim)]
    return self._stage

  def load(self, v_local):
    """v_local: the rank's rows (planes in 3-D) of every velocity component: numpy (ideally pinned)
    or device arrays of local_shape."""
    arrs = []
    for a, stage in zip(v_local, self._staging()):
      assert tuple(a.shape) == self.local_shape, (a.shape, self.local_shape)
      if _lib.is_device_array(a):
        arrs.append(a)
      else:
        a = np.ascontiguousarray(a, np.float32)
        check(lib().cfd_memcpy_h2d(stage.ptr, a.ctypes.data, a.nbytes, self.stream.handle))
        arrs.append(stage)
    check(lib().cfd_dist_load(self.handle, self.stream.handle, _lib.ptr_array(arrs)))
    self.stream.sync()

  def advance(self, nsteps: int):
    check(lib().cfd_dist_advance(self.handle, self.stream.handle, nsteps, ctypes.byref(self.params)))

  def store(self, want_q: bool = False, host_out=None):
    """Materialises the projected local rows.  host_out: optional pair of (pinned) numpy arrays to
    receive them; otherwise device arrays are returned."""
    if host_out is not None:
      outs = self._staging()
    else:
      outs = [_lib.DeviceArray(self.local_shape) for _ in range(self.grid.ndim)]
    q = _lib.DeviceArray(self.local_shape) if want_q else None
    check(lib().cfd_dist_store(self.handle, self.stream.handle, _lib.ptr_array(outs),
                               None if q is None else q.ptr))
    if host_out is not None:
      for h, d in zip(host_out, outs):
        check(lib().cfd_memcpy_d2h(h.ctypes.data, d.ptr, d.nbytes, self.stream.handle))
    self.stream.sync()
    check(lib().cfd_dist_check(self.handle))
    if host_out is not None:
      return (host_out, q) if want_q else host_out
    return (outs, q) if want_q else outs

  def profile(self, nsteps: int = 2):
    """Mean CUDA-event time per launch of every kernel of `nsteps` steps (all ranks must call)."""
    names = (ctypes.c_char_p * 24)()
    ms = (ctypes.c_float * 24)()
    nk = ctypes.c_int(0)
    check(lib().cfd_dist_profile(self.handle, self.stream.handle, nsteps, ctypes.byref(self.params), 24,
                                 ms, names, ctypes.byref(nk)))
    return {names[i].decode(): float(ms[i]) for i in range(nk.value)}

  def sync(self):
    """Waits for the enqueued steps and raises if a device-side wait on a peer ever timed out."""
    self.stream.sync()
    check(lib().cfd_dist_check(self.handle))

  def close(self):
    """Collective: every rank must call it.  The peers' kernels may still be reading this rank's
    buffers, so the ranks meet on the host (one more exchange) before the memory is released."""
    if getattr(self, 'handle', None):
      try:
        self.stream.sync()
        if self.world > 1:
          self._exchange(b'close')
      finally:
        lib().cfd_plan_destroy(self.handle)
        self.handle = None

  def __del__(self):
    # not collective (the peers may be gone already): release without the meeting
    h = getattr(self, 'handle', None)
    if h and _lib._lib is not None:
      try:
        _lib._lib.cfd_plan_destroy(h)
      except Exception:  # interpreter shutdown
        pass
      self.handle = None
