python scripts/state_hash.py 2>&1 | tail -2 | head -1
CFD_B200_LIB=$PWD/jax-cfd_b200/lib/libcfd_b200_l8.so python scripts/state_hash.py 2>&1 | tail -2 | head -1
run() { # name, workload, steps, warmup, env...
  name=$1; wl=$2; st=$3; wu=$4; shift; shift; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --gpus 1 --steps $st --warmup $wu --no-cpu-baseline > gpurun_out/r3g_$name.json 2> gpurun_out/r3g_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/r3g_$name.json') if l.startswith('{')][-1])
  k = d.get('kernel_ms_rank0') or d.get('roofline_step',{}).get('kernel_ms')
  print('$name', 'ms/step', round(d['ms_per_step'],5), 'value', round(d['value'],2), {a:round(b,4) for a,b in k.items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/r3g_$name.err').read()[-1500:])
PY
}
run e1024 E1024 20 5 A=1
run e1024_l8 E1024 20 5 CFD_B200_LIB=$PWD/jax-cfd_b200/lib/libcfd_b200_l8.so
run k256_l8 K256 200 10 CFD_B200_LIB=$PWD/jax-cfd_b200/lib/libcfd_b200_l8.so
run d2048_l8 D2048 200 10 CFD_B200_LIB=$PWD/jax-cfd_b200/lib/libcfd_b200_l8.so
