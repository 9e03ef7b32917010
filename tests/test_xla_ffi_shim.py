"""csrc/xla_ffi_shim.cc (the XLA typed-FFI handlers a JAX host registers) compiled against a
stand-in for xla/ffi/api/ffi.h and driven the way XLA would call it.  XLA / JAX are absent from this
image, so this is how the handler LOGIC is exercised: argument validation and error mapping on the
CPU box; on a GPU the handlers must reproduce cfd_repeated / cfd_project bit for bit for nsteps =
1, 2, 3 (even chains included), batched operands, a matmul plan and 3-D, without writing operands."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, 'jax-cfd_b200', 'lib')


def _build(tmp_path):
  exe = str(tmp_path / 'xla_shim_harness')
  cmd = ['g++', '-O1', '-std=c++17', '-I' + os.path.join(ROOT, 'tests', 'host', 'xla_stub'),
         '-I' + os.path.join(ROOT, 'include'), '-I/usr/local/cuda/include',
         os.path.join(ROOT, 'jax-cfd_b200', 'csrc', 'xla_ffi_shim.cc'),
         os.path.join(ROOT, 'tests', 'host', 'xla_shim_harness.cc'),
         '-L' + LIBDIR, '-lcfd_b200', '-Wl,-rpath,' + LIBDIR, '-L/usr/local/cuda/lib64', '-lcudart', '-o', exe]
  subprocess.run(cmd, check=True)
  return exe


needs_toolchain = pytest.mark.skipif(
    shutil.which('g++') is None or not os.path.exists(os.path.join(LIBDIR, 'libcfd_b200.so')),
    reason='needs g++ and the built libcfd_b200.so')


@needs_toolchain
def test_handlers_validate_arguments_and_map_errors(tmp_path):
  out = subprocess.run([_build(tmp_path), 'cpu'], capture_output=True, text=True)
  assert out.returncode == 0 and 'XLA SHIM CPU PASS' in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
@needs_toolchain
def test_handlers_reproduce_the_c_abi_bitwise(tmp_path):
  out = subprocess.run([_build(tmp_path), 'gpu'], capture_output=True, text=True, timeout=600)
  assert out.returncode == 0 and 'XLA SHIM GPU PASS' in out.stdout, out.stdout + out.stderr
