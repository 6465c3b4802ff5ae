#!/bin/bash
# 2 (or 4) GPUs at the final state: quick parity, the driver's bench command (tuned), R2C, reference CLI testcase 1 at full size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/mgpu_parity.py --quick > gpurun_out/r02d_mgpu${N}_parity.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/r02d_mgpu${N}_parity.log
grep -c "^ok" gpurun_out/r02d_mgpu${N}_parity.log; grep "FAIL\|failed\|Error\|error" gpurun_out/r02d_mgpu${N}_parity.log | head -10
timeout 600 $TR --master-port 29512 tests/cli.py slab -nx 1024 -ny 512 -nz 512 -t 1 -d -snd Streams > gpurun_out/r02d_cli_t1_n${N}.log 2>&1; echo "cli t1 rc=$?"; grep Result gpurun_out/r02d_cli_t1_n${N}.log
B="--no-e2e --no-cpu --steps 10 --warmup 3"
timeout 900 $TR --master-port 29513 tools/multi_bench.py "tuned:$B" "r2c_tuned:$B --transform r2c" "sync:$B --send Sync" > gpurun_out/r02d_mb${N}.log 2>&1; echo "multi_bench rc=$?"
grep -v "^\[\|^\*\|^Setting\|NCCL version\|^$\|UserWarning\|e_in = " gpurun_out/r02d_mb${N}.log | cut -c1-3000
timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02d_bench_n${N}.json 2> gpurun_out/r02d_bench_n${N}.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02d_bench_n${N}.json').read().strip().splitlines()[-1])
print('bench', d['ms_per_step'], d['value'], 'roofline', d['roofline']['kernel'][:60], round(d['roofline']['frac'],3), 'e2e', d['e2e']['ms_per_step'], d['config']['parity']['ok'])
PY
timeout 300 $TR --master-port 29515 bench.py --impl reference --gpus $N --steps 3 --warmup 1 | cut -c1-400
