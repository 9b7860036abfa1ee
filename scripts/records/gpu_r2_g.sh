#!/bin/bash
# Round 2, GPU visit G (2 GPUs): the partitioned mode as the driver runs it (bench.py --gpus 2 under torchrun), the
# 2-GPU NCCL parity test, and the fixed single-GPU tests
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity.log gpurun_out/summary.txt
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
for f in test_gpu_partition test_gpu_parity_80k test_gpu_e2e_mad; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" >> gpurun_out/summary.txt
  tail -n 3 gpurun_out/$f.log
done
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k "operator_level or directed" > gpurun_out/ops_ctx.log 2>&1
echo "ops_ctx exit=$?" >> gpurun_out/summary.txt; tail -n 2 gpurun_out/ops_ctx.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench 2gpu exit=$?" >> gpurun_out/summary.txt; cat gpurun_out/bench_2gpu.json; tail -n 5 gpurun_out/bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 5 --warmup 3 --no-overlap --no-mode-a --no-e2e > gpurun_out/bench_2gpu_1stream.json 2> gpurun_out/bench_2gpu_1stream.err
echo "bench 2gpu 1stream exit=$?" >> gpurun_out/summary.txt; cat gpurun_out/bench_2gpu_1stream.json; tail -n 3 gpurun_out/bench_2gpu_1stream.err
cat gpurun_out/summary.txt
