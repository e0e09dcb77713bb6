#!/bin/bash
# multi-GPU visit (gpurun --gpus N): in-process World sharding test at tiny + SD1.5 size, log kept for profiles/
mkdir -p gpurun_out
python stable-diffusion-webui-distributed_b200/b200sd/build.py > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/mgpu_sd15.log
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -s -p no:cacheprovider --tb=short 2>&1 | tail -30 >> gpurun_out/mgpu_sd15.log
cat gpurun_out/mgpu_sd15.log
