#!/bin/bash
# round-2 GPU call 1: is pybullet obtainable on the GPU box through the sanctioned sources, GPU tests,
# physics timing A/B, bench line.  Only pip's configured sources and the offline wheelhouse are tried: an install from the
# public index needs a human's authorisation in this environment and was not attempted.
mkdir -p gpurun_out
L=gpurun_out/r02_pybullet_attempt.log
{
  echo "# is pybullet + gym obtainable on the GPU box?  ($(date -u +%FT%TZ))"; echo '$ python -m pip --version'; python -m pip --version
  echo '$ pip config list'; python -m pip config list
  echo '$ timeout 60 python -m pip install pybullet gym   (configured sources only)'; timeout 60 python -m pip install pybullet gym; echo "exit code $?"
  echo '$ timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse pybullet gym'; timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse pybullet gym; echo "exit code $?"
  echo '$ ls /opt/wheelhouse | grep -i -E "bullet|gym"'; ls /opt/wheelhouse 2>/dev/null | grep -i -E "bullet|gym"; echo "exit code $?"
  echo '$ find / -xdev -iname "*bullet*"'; find / -xdev -iname "*bullet*" -not -path "/proc/*" 2>/dev/null | head -20
  echo '$ python -c "import pybullet, gym"'; python -c "import pybullet, gym"; echo "exit code $?"
} > $L 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r02_call1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02_call1_pytest.log
python scripts/time_step.py > gpurun_out/r02_call1_time_step.log 2>&1
RR_LIB=real_robots_b200/variants/lib_solve7.so python scripts/time_step.py >> gpurun_out/r02_call1_time_step.log 2>&1
python bench.py > gpurun_out/r02_call1_bench.log 2>&1
tail -3 gpurun_out/r02_call1_pytest.log; cat gpurun_out/r02_call1_time_step.log; tail -1 gpurun_out/r02_call1_bench.log | cut -c1-600
