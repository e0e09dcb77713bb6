#!/bin/bash
# Runs the GPU kernel parity tests group by group, each in its own process (a device trap in one
# group must not poison the CUDA context of the others). Output -> gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/kernel_parity.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for grp in "linear_plain" "linear and not plain" "conv2d" "attention" "groupnorm or layernorm" "upsample or softmax or silu or fold or cfg or quantize"; do
  tag=$(echo "$grp" | tr ' ' '_')
  echo "=== $grp"
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "$grp" -p no:cacheprovider --tb=line 2>&1 | tail -n 40 | tee "gpurun_out/kernels_${tag}.log"
done
