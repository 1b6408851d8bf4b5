#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 scripts/bench_c5.py --envs 8192 > gpurun_out/r02_c5_n$N.json 2> gpurun_out/r02_c5_n$N.err
tail -1 gpurun_out/r02_bench_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['d2h_ceiling_gbs'], d['e2e']['frac_of_ceiling'], d['collective_check'])"
tail -1 gpurun_out/r02_c5_n$N.json
