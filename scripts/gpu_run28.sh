run() { # name, workload, steps, warmup, env...
  name=$1; wl=$2; st=$3; wu=$4; shift; shift; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --gpus 1 --steps $st --warmup $wu --no-cpu-baseline > gpurun_out/r2y_$name.json 2> gpurun_out/r2y_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/r2y_$name.json') if l.startswith('{')][-1])
  k = d.get('kernel_ms_rank0') or d.get('roofline_step',{}).get('kernel_ms')
  print('$name', 'ms/step', round(d['ms_per_step'],5), 'value', round(d['value'],2), {a:round(b,4) for a,b in k.items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/r2y_$name.err').read()[-1500:])
PY
}
run k8192 K8192 20 5 A=1
run k8192_p1s1 K8192 20 5 CFD_T_PAIRED=1 CFD_FFT_ROWS_SHIFT=1
run k8192_p1s0 K8192 20 5 CFD_T_PAIRED=1
run k8192_p0s1 K8192 20 5 CFD_FFT_ROWS_SHIFT=1
run k8192_p1s2 K8192 20 5 CFD_T_PAIRED=1 CFD_FFT_ROWS_SHIFT=2
run d2048 D2048 200 10 A=1
