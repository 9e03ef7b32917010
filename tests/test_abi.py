"""The C-ABI library loads and exports every symbol include/cfd_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
  text = open(os.path.join(ROOT, 'include', 'cfd_b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(cfd_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_header_symbol():
  from jax_cfd_b200 import _lib
  lib = _lib.lib()
  syms = header_symbols()
  assert len(syms) >= 25
  for s in syms:
    assert hasattr(lib, s), f'{s} declared in include/cfd_b200.h but not exported'
  assert set(syms) == set(_lib.EXPORTED_SYMBOLS), set(syms) ^ set(_lib.EXPORTED_SYMBOLS)
  assert b'sm_100a' in lib.cfd_version()


def test_params_struct_layout_matches_header():
  """sizeof(cfd_params) as laid out by ctypes == the C compiler's (checked via a tiny C probe)."""
  import subprocess
  import tempfile
  from jax_cfd_b200 import _lib
  src = '#include <stdio.h>\n#include "cfd_b200.h"\nint main(){printf("%zu %zu", sizeof(cfd_params), sizeof(cfd_diag));return 0;}\n'
  with tempfile.TemporaryDirectory() as d:
    c = os.path.join(d, 'p.c')
    open(c, 'w').write(src)
    exe = os.path.join(d, 'p')
    subprocess.run(['/usr/bin/gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
  assert int(out[0]) == ctypes.sizeof(_lib.Params)
  assert int(out[1]) == ctypes.sizeof(_lib.Diag)


def test_compute_fails_loudly_without_gpu():
  import numpy as np
  import jax_cfd_b200 as cfd
  if cfd._lib.lib().cfd_device_count() > 0:
    pytest.skip('a GPU is present')
  g = cfd.grids.Grid((64, 64), domain=((0, 1), (0, 1)))
  step = cfd.equations.semi_implicit_navier_stokes(1.0, 1e-3, 0.01, g)
  bc = cfd.boundaries.periodic_boundary_conditions(2)
  v = tuple(cfd.grids.GridVariable(cfd.grids.GridArray(np.zeros((64, 64), np.float32), o, g), bc)
            for o in g.cell_faces)
  with pytest.raises(cfd.CfdError):
    step(v)
