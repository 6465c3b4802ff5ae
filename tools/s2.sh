#!/bin/bash
# GPU session 2 (two GPUs): multi-rank parity (incl. full-size vs cuFFT), exchange variants of the slab at N=2,
# NVLink baselines and the NVLink counters of the exchanging pass under ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tests/mgpu_parity.py ${PARITY_ARGS:---quick} > gpurun_out/r02_mgpu${N}_parity.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/r02_mgpu${N}_parity.log
grep -c "^ok" gpurun_out/r02_mgpu${N}_parity.log; grep "FAIL\|failed\|Error\|error" gpurun_out/r02_mgpu${N}_parity.log | head -20
timeout 600 $TR --master-port 29512 tests/mgpu_parity.py --cufft > gpurun_out/r02_mgpu${N}_cufft.log 2>&1; echo "cufft rc=$?" | tee -a gpurun_out/r02_mgpu${N}_cufft.log
grep "full-size\|failed\|Error" gpurun_out/r02_mgpu${N}_cufft.log | head -20
timeout 300 python tools/peer_bw.py 256 > gpurun_out/r02_peer_bw_${N}.log 2>&1; cat gpurun_out/r02_peer_bw_${N}.log | tail -2
B="--no-e2e --no-cpu --steps 10 --warmup 3"
timeout 1500 $TR --master-port 29513 tools/multi_bench.py \
  "sync:$B --send Sync" \
  "a2a:$B --send Sync --comm All2All" \
  "streams96:$B" \
  "streams96_notma:$B DFFT_TMA=0" \
  "streams96_invplain:$B DFFT_BLOCKED_INV=0" \
  "streams96_bulk:$B DFFT_BULK_STORE=1" \
  "streams48_bulk:$B DFFT_BULK_STORE=1 DFFT_XCHG_CTAS=48" \
  "streams148:$B DFFT_XCHG_CTAS=148" \
  "streams48:$B DFFT_XCHG_CTAS=48" \
  "streams96_c8:$B DFFT_OVL_CHUNKS=8" \
  "narrow4_296:$B DFFT_BLOCKED=4 DFFT_XCHG_WIDE=0 DFFT_XCHG_CTAS=296" \
  "narrow4_148:$B DFFT_BLOCKED=4 DFFT_XCHG_WIDE=0 DFFT_XCHG_CTAS=148" \
  "narrow4_148_bulk:$B DFFT_BLOCKED=4 DFFT_XCHG_WIDE=0 DFFT_XCHG_CTAS=148 DFFT_BULK_STORE=1" \
  "narrow4_74_bulk:$B DFFT_BLOCKED=4 DFFT_XCHG_WIDE=0 DFFT_XCHG_CTAS=74 DFFT_BULK_STORE=1" \
  "sync_narrow4:$B --send Sync DFFT_BLOCKED=4 DFFT_XCHG_WIDE=0" \
  "sync_narrow4_bulk:$B --send Sync DFFT_BLOCKED=4 DFFT_XCHG_WIDE=0 DFFT_BULK_STORE=1" \
  "sync_bulk:$B --send Sync DFFT_BULK_STORE=1" \
  "r2c_streams:$B --transform r2c" \
  "r2c_sync:$B --transform r2c --send Sync" \
  "r2c_streams_bulk:$B --transform r2c DFFT_BULK_STORE=1" \
  > gpurun_out/r02_mb${N}.log 2>&1; echo "multi_bench rc=$?"
cat gpurun_out/r02_mb${N}.log | grep -v "^\[" | cut -c1-400
# NVLink counters of the exchanging kernel (single process, kernel replay is safe)
timeout 120 python tools/nvlink_probe.py 1024 8 32 f64 > gpurun_out/r02_nvlink_probe.log 2>&1; tail -2 gpurun_out/r02_nvlink_probe.log
ncu --query-metrics 2>/dev/null | grep -i "nvl" | head -40 > gpurun_out/r02_ncu_nvl_metrics.txt
timeout 300 ncu --section Nvlink --section Nvlink_Tables --section Nvlink_Topology --metrics regex:nvl.x__bytes,gpu__time_duration.sum --clock-control none -k regex:fft_c2c -s 2 -c 1 -o gpurun_out/r02_nvlink_probe python tools/nvlink_probe.py 1024 8 32 f64 > gpurun_out/r02_nvlink_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_nvlink_ncu.log
