#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02_final_pytest.log
python bench.py > gpurun_out/r02_final_bench.log 2>&1
python scripts/bench_configs.py > gpurun_out/r02_final_configs.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/r02_final_pytest.log | tail -5; tail -1 gpurun_out/r02_final_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['kernels_ms_per_step'])"; cat gpurun_out/r02_final_configs.log
python scripts/bench_c4.py > gpurun_out/r02_final_c4.log 2>&1; cat gpurun_out/r02_final_c4.log
