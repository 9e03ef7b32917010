TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
CFD_T_PAIRED=1 timeout 300 $TR --master-port 29511 tests/mgpu_worker.py 32768 256 2 2>&1 | grep -E "bitwise|MGPU|rror" | tail -5
timeout 300 $TR --master-port 29512 tests/mgpu_worker.py 32768 256 2 2>&1 | grep -E "bitwise|MGPU|rror" | tail -5
