"""Host emulation of the shared-memory line FFT machinery (csrc/fft_smem.cuh): every radix
schedule the kernels instantiate (incl. the balanced radix-32 schedule of the 8192-point x lines
and the reduced-load twiddle scheme) against a double-precision DFT.  No GPU needed: the device
functions are compiled for the CPU and the threads of a line are run pass by pass."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which('g++') is None or not os.path.exists('/usr/local/cuda/include/cuda_runtime.h'),
                    reason='needs g++ and the CUDA headers')
def test_line_fft_schedules_match_dft(tmp_path):
  exe = str(tmp_path / 'fft_emulate')
  subprocess.run(['g++', '-O1', '-std=c++17', '-w', '-I/usr/local/cuda/include',
                  os.path.join(ROOT, 'tests', 'host', 'fft_emulate.cpp'), '-o', exe], check=True)
  out = subprocess.run([exe], capture_output=True, text=True)
  assert out.returncode == 0, out.stdout
  assert 'FftPlan<13,5,5,true>' in out.stdout and 'worst' in out.stdout
