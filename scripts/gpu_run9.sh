# 3-D slab decomposition: world-1 code path, world-2 bitwise check, TGV512 at 1 and 2 GPUs
N=${1:-2}
python -m pytest tests -m gpu -q -x -k "3d or slab_stepper" 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/mgpu_worker.py 64 32 64 3 2>&1 | grep -E "bitwise|MGPU|rror|Trace" | head -20
timeout 300 $TR --master-port 29512 tests/mgpu_worker.py 128 128 64 2 2>&1 | grep -E "bitwise|MGPU|rror|Trace" | head -20
timeout 300 $TR --master-port 29513 tests/mgpu_worker.py 2048 1024 4 2>&1 | grep -E "bitwise|MGPU|rror|Trace" | head -20
timeout 600 python bench.py --workload TGV512 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_tgv512_n1.json 2> gpurun_out/r2h_tgv512_n1.err
timeout 600 $TR --master-port 29520 bench.py --workload TGV512 --gpus $N --steps 10 --warmup 3 > gpurun_out/r2h_tgv512_n$N.json 2> gpurun_out/r2h_tgv512_n$N.err
python - <<PY
import json
for f in ['gpurun_out/r2h_tgv512_n1', 'gpurun_out/r2h_tgv512_n$N']:
  try:
    d=json.loads([l for l in open(f+'.json') if l.startswith('{')][-1])
    k = d.get('kernel_ms_rank0') or d.get('roofline_step',{}).get('kernel_ms')
    print(f, 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],2), 'parity', (d.get('parity') or {}).get('bitwise'), {a:round(b,3) for a,b in (k or {}).items()})
  except Exception as e:
    print(f, 'FAILED', e); print(open(f+'.err').read()[-2500:])
PY
