TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/mgpu_worker.py 256 512 6 2>&1 | grep -E "bitwise|MGPU|rror" 
timeout 300 $TR --master-port 29512 tests/mgpu_worker.py 2048 1024 4 2>&1 | grep -E "bitwise|MGPU|rror"
timeout 300 $TR --master-port 29513 tests/mgpu_worker.py 1024 16384 3 2>&1 | grep -E "bitwise|MGPU|rror"
timeout 300 $TR --master-port 29514 tests/mgpu_worker.py 32768 64 3 2>&1 | grep -E "bitwise|MGPU|rror"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 $TR --master-port 29520 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_n2_$name.json 2> gpurun_out/r2e_n2_$name.err
  python - <<PY
import json
try:
  d=json.loads([l for l in open('gpurun_out/r2e_n2_$name.json') if l.startswith('{')][-1])
  print('$name', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'parity', (d.get('parity') or {}).get('bitwise'), {k:round(v,3) for k,v in d['kernel_ms_rank0'].items()})
except Exception as e:
  print('$name FAILED', e); print(open('gpurun_out/r2e_n2_$name.err').read()[-1500:])
PY
}
run push A=1
run push_c8 CFD_DIST_COPY_CTAS=8
run push_c48 CFD_DIST_COPY_CTAS=48
run push_b1 CFD_DIST_BLOCKS=1
run push_b2 CFD_DIST_BLOCKS=2
run push_b8 CFD_DIST_BLOCKS=8
