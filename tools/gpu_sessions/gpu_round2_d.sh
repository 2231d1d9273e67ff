#!/bin/bash
# multi-GPU session: (N = 2: whole GPU suite incl. the 2-GPU gather test and the multi-device C++ call,) bench under
# torchrun, kernel-vs-gather measurement (N = 8: also on 4 of the GPUs). usage: bash tools/gpu_round2_d.sh N
cd "$(dirname "$0")/../.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/d_topo_$N.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/d_full.log 2>&1
  echo "rc $?" >> gpurun_out/d_full.log
  tail -n 5 gpurun_out/d_full.log
else
  timeout 900 python -m pytest tests -m gpu -q -k "nccl_gather or cpp_drop_in" > gpurun_out/d_multi_tests_$N.log 2>&1
  tail -n 3 gpurun_out/d_multi_tests_$N.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/d_bench_$N.json 2> gpurun_out/d_bench_$N.err
echo "bench rc $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/nccl_gather_check.py > gpurun_out/d_gather_$N.jsonl 2> gpurun_out/d_gather_$N.err
echo "gather rc $?"
grep "^{" gpurun_out/d_gather_$N.jsonl
if [ "$N" = "8" ]; then
  for M in 4 2; do
    CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((M-1))) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $M --master-addr 127.0.0.1 --master-port 29543 tools/nccl_gather_check.py > gpurun_out/d_gather_$M.jsonl 2> gpurun_out/d_gather_$M.err
    grep "^{" gpurun_out/d_gather_$M.jsonl
    CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((M-1))) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $M --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $M --steps 10 --warmup 3 > gpurun_out/d_bench_$M.json 2> gpurun_out/d_bench_$M.err
  done
  # one C++ call over all 8 GPUs (host tensors)
  TCHEM_DEVICES=all timeout 300 python -m pytest tests/test_cpp_host.py -q -m gpu -k drop_in -s 2>&1 | grep -E "device\(s\)|all devices|passed|failed" | head -5
fi
