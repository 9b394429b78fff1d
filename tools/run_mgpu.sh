#!/bin/bash
# multi-GPU round: bench.py under torchrun (the driver's own launch line), the parity gate on adversarial cuts, configs[2]
# usage: gpurun --gpus N -- 'bash tools/run_mgpu.sh N'
set -u
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
run() { local to=$1; shift; timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
echo "== check (adversarial cuts, $N ranks)"; run 600 bench.py --gpus $N --check 2>&1 | tail -3
echo "== bench --gpus $N"; run 900 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; python -c "
import json,sys
for l in open('gpurun_out/bench_n$N.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','n_gpus','parity','gpu_launches')}); print(d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e'])
"; tail -3 gpurun_out/bench_n$N.err
echo "== bench --gpus 1 (same box)"; timeout 600 python bench.py --steps 30 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print({k:d[k] for k in ('value','ms_per_step','parity')}, d['roofline']['kernel_ms'])"
echo "== ndjson_1g --gpus $N"; run 900 bench.py --gpus $N --config ndjson_1g --steps 10 2>&1 | tail -2 | cut -c1-1200
if [ "$N" = "8" ] || [ "$N" = "4" ] || [ "$N" = "2" ]; then echo "== concat_8g --gpus $N"; run 1500 bench.py --gpus $N --config concat_8g --steps 3 2>&1 | tail -2 | cut -c1-1200; fi
echo "== gloo/NCCL protocol test"; timeout 600 python -m pytest tests/test_sharding_gloo.py -q 2>&1 | tail -2
ls gpurun_out | head -40
