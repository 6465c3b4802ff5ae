#!/bin/bash
# Final-state validation on one GPU: the driver's own round-end sequence (GPU test suite, smoke, one bench line).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/validate_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/validate_pytest.log; tail -5 gpurun_out/validate_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --no-cpu > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/validate_bench.json').read().strip().splitlines()[-1])
print('bench ms', round(d['ms_per_step'],4), 'inv', round(d['config']['ms_inverse'],4), [(q['step'],round(q['ms'],3),round(q['gbs'])) for q in d['roofline']['all_passes']], 'e2e', d['e2e']['ms_per_step'], 'cufft', d['cufft_1gpu_ms'], 'parity', d['config']['parity']['ok'])
PY
