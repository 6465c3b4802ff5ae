#!/bin/bash
# GPU session 1 (one GPU): full parity suite, per-pass throughput of the default build and of the experiment
# variants, default bench line.  Everything is logged in full under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_s1_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r02_s1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_s1_pytest.log
tail -5 gpurun_out/r02_s1_pytest.log
for v in "" e8 e8r32 r32; do
  lib=""; [ -n "$v" ] && lib="$PWD/distributedfft_b200/libdfft_$v.so"
  [ -n "$v" ] && [ ! -f "$lib" ] && continue
  DFFT_LIB=$lib timeout 300 python tools/axis_bench.py --prec f64 --sizes 512,1024,2048 --tag "_r02_$v" > gpurun_out/r02_s1_axis_f64_$v.log 2>&1
  cat gpurun_out/r02_s1_axis_f64_$v.log
done
for v in "" r32; do lib=""; [ -n "$v" ] && lib="$PWD/distributedfft_b200/libdfft_$v.so"; DFFT_LIB=$lib timeout 300 python tools/axis_bench.py --prec f32 --sizes 1024,2048,4096 --elems 29 --tag "_r02_$v" > gpurun_out/r02_s1_axis_f32_$v.log 2>&1; cat gpurun_out/r02_s1_axis_f32_$v.log; done
timeout 600 python bench.py > gpurun_out/r02_s1_bench.json 2> gpurun_out/r02_s1_bench.err
echo "bench rc=$?"; cat gpurun_out/r02_s1_bench.json | cut -c1-1500
