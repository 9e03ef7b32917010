python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "paired_spectrum or (matches_oracle and 32768)" 2>&1 | tail -3
run() { # name, workload, env...
  name=$1; wl=$2; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/r2m_$name.json') if l.startswith('{')][-1])
  k = d.get('kernel_ms_rank0') or d.get('roofline_step',{}).get('kernel_ms')
  print('$name', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), {a:round(b,3) for a,b in k.items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/r2m_$name.err').read()[-1500:])
PY
}
run paired_cluster K32768 CFD_SLAB_SHAPE=32768x16384 CFD_X15=cluster
run k8192_tx84 K8192 A=1
run k8192_tx64 K8192 CFD_EXPLICIT_TX=64
run k8192_tx90 K8192 CFD_EXPLICIT_TX=90
run k8192_tx72 K8192 CFD_EXPLICIT_TX=72
