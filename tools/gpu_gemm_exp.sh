#!/bin/bash
# GEMM experiment visit: kernel parity for the GEMM paths, the sweep, the trace
mkdir -p gpurun_out
python stable-diffusion-webui-distributed_b200/b200sd/build.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider --tb=short -k "linear or conv or geglu" 2>&1 | tail -5
timeout 300 python tools/gemm_sweep.py --one 2>&1 | tee gpurun_out/gemm_sweep.log
bash tools/gpu_trace.sh 2>&1 | grep -A22 "geglu" | head -30
