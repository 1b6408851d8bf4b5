#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02_call11_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02_call11_pytest.log
{ python scripts/tracked_fraction.py; RR_LIB=real_robots_b200/variants/lib_physnofma.so python scripts/tracked_fraction.py; } > gpurun_out/r02_call11_tracked.log 2>&1
{ python scripts/time_render.py; RR_LIB=real_robots_b200/variants/lib_tma1.so python scripts/time_render.py; RR_LIB=real_robots_b200/variants/lib_flat7.so python scripts/time_render.py; python scripts/time_render.py; } > gpurun_out/r02_call11_render.log 2>&1
grep -E "passed|failed|^FAILED|contact-rich|touch readings" gpurun_out/r02_call11_pytest.log | tail -6; cat gpurun_out/r02_call11_tracked.log gpurun_out/r02_call11_render.log
