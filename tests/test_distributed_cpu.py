"""Host-side logic of the slab decomposition, including a world_size-2 gloo run (CPU only)."""
import os
import sys

import numpy as np
import pytest

import jax_cfd_b200 as cfd
from jax_cfd_b200 import distributed as D


def test_partition_arithmetic():
  assert D.slab_rows(8192, 0, 8) == (0, 1024) and D.slab_rows(8192, 7, 8) == (7168, 8192)
  assert D.line_range(8192, 3, 4) == (3072, 4096)
  rows = [D.slab_rows(256, r, 4) for r in range(4)]
  assert rows[0][0] == 0 and rows[-1][1] == 256 and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
  with pytest.raises(ValueError):
    D.slab_rows(100, 0, 8)
  D.check_decomposition((8192, 8192), 8)
  D.check_decomposition((16384, 8192), 2)
  D.check_decomposition((32768, 32768), 8)
  with pytest.raises(ValueError):
    D.check_decomposition((64, 64), 8)       # 8 rows per rank
  with pytest.raises(ValueError):
    D.check_decomposition((8192, 64), 4)     # 8 lines per rank
  with pytest.raises(NotImplementedError):
    D.check_decomposition((65536, 32768), 8)
  # 3-D slabs (config #5): the constraints of cfd_dist_plan_create_nd, same messages
  D.check_decomposition((512, 512, 512), 8)
  D.check_decomposition((64, 32, 64), 2)
  with pytest.raises(ValueError, match='power of two >= 16 planes'):
    D.check_decomposition((64, 128, 64), 8)        # 8 planes per rank
  with pytest.raises(ValueError, match='N2 % 64'):
    D.check_decomposition((128, 128, 32), 2)       # the marching stencil's tile is 8 x 64
  with pytest.raises(ValueError, match='16 x the number of ranks'):
    D.check_decomposition((256, 64, 64), 8)        # (kz, ky) lines must split evenly over the ranks
  with pytest.raises(ValueError):
    D.check_decomposition((512, 512, 512), 3)


def _worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  import torch.distributed as dist
  dist.init_process_group('gloo', rank=rank, world_size=world)
  blob = bytes([rank]) * 64
  blobs = D.torch_exchange(blob)
  ok = blobs == [bytes([r]) * 64 for r in range(world)]
  # every rank derives a consistent, disjoint cover of rows and lines
  rows = D.slab_rows(256, rank, world)
  lines = D.line_range(512, rank, world)
  allr = [None] * world
  dist.all_gather_object(allr, (rows, lines))
  ok = ok and sorted(r for r, _ in allr) == [D.slab_rows(256, r, world) for r in range(world)]
  ok = ok and sum(l[1] - l[0] for _, l in allr) == 256
  q.put((rank, ok))
  dist.destroy_process_group()


def test_handle_exchange_world2_gloo():
  import torch.multiprocessing as mp
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = 29500 + (os.getpid() % 400)
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=120) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  assert sorted(res) == [(0, True), (1, True)]


def _tcp_worker(rank, world, port, q):
  ex = D.tcp_exchange(rank, world, '127.0.0.1', port, timeout=60)
  first = ex(bytes([rank + 1]) * 64)
  second = ex(b'close')  # the collective close() of SlabStepper reuses the same rendezvous
  q.put((rank, first == [bytes([r + 1]) * 64 for r in range(world)] and second == [b'close'] * world))


@pytest.mark.parametrize('world', [2, 3])
def test_torch_free_tcp_exchange(world):
  """The host side of the slab decomposition without torch: one all-gather of 64-byte handles over
  a standard-library TCP rendezvous ('ship a torch-free exchange')."""
  import multiprocessing as mp
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = 30100 + (os.getpid() % 300) + world
  procs = [ctx.Process(target=_tcp_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = [q.get(timeout=120) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  assert sorted(res) == [(r, True) for r in range(world)]
