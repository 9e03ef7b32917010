python scripts/state_hash.py 2>&1 | tail -1
run() { # name, workload, steps, warmup, env...
  name=$1; wl=$2; st=$3; wu=$4; shift; shift; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --gpus 1 --steps $st --warmup $wu --no-cpu-baseline > gpurun_out/r3f_$name.json 2> gpurun_out/r3f_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/r3f_$name.json') if l.startswith('{')][-1])
  k = d.get('kernel_ms_rank0') or d.get('roofline_step',{}).get('kernel_ms')
  print('$name', 'ms/step', round(d['ms_per_step'],5), 'value', round(d['value'],2), {a:round(b,4) for a,b in k.items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/r3f_$name.err').read()[-1500:])
PY
}
run tgv512 TGV512 10 3 A=1
run tgv512_x3 TGV512 10 3 CFD_B200_LIB=$PWD/jax-cfd_b200/lib/libcfd_b200_x3.so
run tgv512_x4 TGV512 10 3 CFD_B200_LIB=$PWD/jax-cfd_b200/lib/libcfd_b200_x4.so
