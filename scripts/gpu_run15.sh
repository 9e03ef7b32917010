# N GPUs: push-mode variants at the weak-scaling shape
N=$1; TAG=${2:-r2n}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 900 $TR --master-port 29520 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_n${N}_$name.json 2> gpurun_out/${TAG}_n${N}_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/${TAG}_n${N}_$name.json') if l.startswith('{')][-1])
  print('$name', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'parity', (d.get('parity') or {}).get('bitwise'), {k:round(v,3) for k,v in d['kernel_ms_rank0'].items()})
  if d.get('k32768'): print('   k32768', round(d['k32768']['ms_per_step'],3), round(d['k32768']['value'],1), {k:round(v,3) for k,v in d['k32768']['kernel_ms_rank0'].items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/${TAG}_n${N}_$name.err').read()[-1500:])
PY
}
run default A=1
run parts2 CFD_DIST_STENCIL_PARTS=2
run parts1 CFD_DIST_STENCIL_PARTS=1
run pull CFD_DIST_MODE=pull
