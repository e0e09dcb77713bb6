#!/bin/bash
# Final check of a tree on one B200: the whole GPU suite, then smoke().
mkdir -p gpurun_out
python stable-diffusion-webui-distributed_b200/b200sd/build.py > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short --durations=8 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
tail -25 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log
exit 0
