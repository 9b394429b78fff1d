#!/bin/bash
set -u
N=${1:-4}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { local to=$1; shift; timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
echo "== bench --gpus $N"; run 600 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; grep "^\[rank" gpurun_out/bench_n$N.err; python -c "
import json
for l in open('gpurun_out/bench_n$N.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus','parity','gpu_launches')}); print(d['roofline']['kernel_ms'], d.get('sharded_host'))
"
echo "== again, 100 steps"; run 600 bench.py --gpus $N --steps 100 --warmup 3 2> gpurun_out/bench_n${N}b.err | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d['roofline']['kernel_ms'], d.get('sharded_host'))
"; grep "^\[rank" gpurun_out/bench_n${N}b.err
echo "== without the NVML sampler"; SJB200_BENCH_NVML=0 run 600 bench.py --gpus $N --steps 30 --warmup 3 2> gpurun_out/bench_n${N}c.err | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d['roofline']['kernel_ms'], d.get('sharded_host'))
"; grep "^\[rank" gpurun_out/bench_n${N}c.err
