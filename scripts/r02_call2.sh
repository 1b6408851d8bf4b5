#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_call4_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02_call4_pytest.log
python scripts/time_step.py > gpurun_out/r02_call4_time_step.log 2>&1
python bench.py > gpurun_out/r02_call4_bench.log 2>&1
tail -5 gpurun_out/r02_call4_pytest.log; cat gpurun_out/r02_call4_time_step.log; tail -1 gpurun_out/r02_call4_bench.log | cut -c1-400
