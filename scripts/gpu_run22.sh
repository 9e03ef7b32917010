# final N-GPU check: bitwise workers, the weak-scaling bench line (+ k32768 at 8), TGV512 strong scaling
N=$1; TAG=${2:-r2u}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/mgpu_worker.py 2048 1024 4 2>&1 | grep -E "MGPU|rror" | tail -2
timeout 300 $TR --master-port 29512 tests/mgpu_worker.py 128 128 64 2 2>&1 | grep -E "MGPU|rror" | tail -2
run() { # name, env... [--workload W]
  name=$1; shift
  EXTRA=""
  if [ "$2" = "--workload" ]; then EXTRA="--workload $3"; set -- "$1"; fi
  env "$@" timeout 900 $TR --master-port 29520 bench.py --gpus $N --steps 20 --warmup 5 $EXTRA > gpurun_out/${TAG}_n${N}_$name.json 2> gpurun_out/${TAG}_n${N}_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/${TAG}_n${N}_$name.json') if l.startswith('{')][-1])
  print('$name', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],2), 'parity', (d.get('parity') or {}).get('bitwise'), {k:round(v,3) for k,v in d['kernel_ms_rank0'].items()})
  if d.get('k32768'): print('   k32768', round(d['k32768']['ms_per_step'],3), round(d['k32768']['value'],1), {k:round(v,3) for k,v in d['k32768']['kernel_ms_rank0'].items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/${TAG}_n${N}_$name.err').read()[-1500:])
PY
}
run default A=1
run tgv512 A=1 --workload TGV512
