#!/bin/bash
# bench line only, under torchrun on N GPUs (refresh of the multi-GPU numbers after a kernel change). usage: bash tools/gpu_sessions/gpu_round2_k.sh N
cd "$(dirname "$0")/../.."
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/k_bench_$N.json 2> gpurun_out/k_bench_$N.err
echo "bench rc $?"
tail -c 600 gpurun_out/k_bench_$N.json
