#!/bin/bash
# trace visit: GEMM timeline with a trace-enabled build, then restore the normal build
mkdir -p gpurun_out
B200SD_NVCC_EXTRA="-DB200SD_GEMM_TRACE_ENABLE=1" python stable-diffusion-webui-distributed_b200/b200sd/build.py --force > gpurun_out/build_trace.log 2>&1
timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace.log 2>&1
python stable-diffusion-webui-distributed_b200/b200sd/build.py --force > gpurun_out/build.log 2>&1
cat gpurun_out/gemm_trace.log
