#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/r02_call12_multi.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02_call12_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02_call12_bench2.log 2> gpurun_out/r02_call12_bench2.err
tail -3 gpurun_out/r02_call12_multi.log; tail -1 gpurun_out/r02_call12_bench2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['d2h_ceiling_gbs'], d['e2e']['frac_of_ceiling'], d['collective_check'])"
