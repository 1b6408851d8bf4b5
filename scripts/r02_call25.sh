#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_call25_pytest.log 2>&1; tail -3 gpurun_out/r02_call25_pytest.log
python scripts/soak.py 2>&1 | grep soak
python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; tail -c 2200 gpurun_out/r02c_bench_n1.json
python scripts/bench_configs.py > gpurun_out/r02c_configs.log 2>&1; python scripts/bench_c4.py >> gpurun_out/r02c_configs.log 2>&1; grep "^C" gpurun_out/r02c_configs.log
