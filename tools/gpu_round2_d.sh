#!/bin/bash
# multi-GPU session: whole GPU suite (incl. the 2-GPU gather test and the multi-device C++ call), bench under torchrun,
# kernel-vs-gather measurement. usage: bash tools/gpu_round2_d.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/d_topo_$N.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/d_full.log 2>&1
  echo "rc $?" >> gpurun_out/d_full.log
  tail -n 5 gpurun_out/d_full.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/d_bench_$N.json 2> gpurun_out/d_bench_$N.err
echo "bench rc $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/nccl_gather_check.py > gpurun_out/d_gather_$N.jsonl 2> gpurun_out/d_gather_$N.err
echo "gather rc $?"
cat gpurun_out/d_gather_$N.jsonl
grep -n 'device(s)' -A1 gpurun_out/d_full.log | head -4
