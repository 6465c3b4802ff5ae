#!/bin/bash
# The reference's validation sweep (testcase 4, double precision, every comm/send method: jobs/*/{slab,pencil}/validation.json)
# through tests/launch_jobs.py --in-process on N GPUs.  Usage: bash tools/session_sweep.sh <ngpus> [max-size]
N=${1:-2}; MAX=${2:-1024}
mkdir -p gpurun_out
JOBS="tests/jobs/validation_slab.json"
[ -f tests/jobs/validation_pencil_${N}gpu.json ] && JOBS="$JOBS tests/jobs/validation_pencil_${N}gpu.json"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  tests/launch_jobs.py $JOBS --in-process --max-size $MAX --report gpurun_out/validation_sweep_n$N.json > gpurun_out/validation_sweep_n$N.log 2>&1
echo "sweep rc=$?"
grep -c "^SWEEP ok" gpurun_out/validation_sweep_n$N.log; grep "^SWEEP FAIL\|^SWEEP summary\|rror" gpurun_out/validation_sweep_n$N.log | head -20
tail -3 gpurun_out/validation_sweep_n$N.log
