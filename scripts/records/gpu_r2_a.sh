#!/bin/bash
# Round 2, GPU visit A: new parity tests (reference drivers over the drop-in, 81,920-face unmodified-oracle step),
# end-to-end MAD product arms (both weight sets, both product paths), host facts for the CPU arm.
set -u
mkdir -p gpurun_out/e2e
rm -f gpurun_out/parity.log gpurun_out/summary.txt
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
(nproc; free -g; lscpu | head -20) > gpurun_out/host.txt 2>&1
for f in test_reference_driver test_gpu_parity_80k; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" >> gpurun_out/summary.txt
  tail -n 5 gpurun_out/$f.log
done
for cfg in default cad; do
  for path in dualstep dropin; do
    timeout 600 python scripts/e2e_mad.py --arm product --config $cfg --path $path \
      --out gpurun_out/e2e/product_${cfg}_${path}.json 2> gpurun_out/e2e/product_${cfg}_${path}.log
    echo "e2e $cfg $path exit=$?" >> gpurun_out/summary.txt
    tail -n 2 gpurun_out/e2e/product_${cfg}_${path}.log
  done
done
cat gpurun_out/summary.txt
