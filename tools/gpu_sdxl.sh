#!/bin/bash
# SDXL visit: parity tests + bench --model sdxl (N=1)
mkdir -p gpurun_out
python stable-diffusion-webui-distributed_b200/b200sd/build.py > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests/test_sdxl_gpu.py -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | tail -25 > gpurun_out/pytest_sdxl.log
timeout 900 python bench.py --model sdxl --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sdxl.json 2> gpurun_out/bench_sdxl.err
tail -8 gpurun_out/pytest_sdxl.log; head -c 500 gpurun_out/bench_sdxl.json; echo; grep -v "^DISTRIBUTED" gpurun_out/bench_sdxl.err | tail -5
grep xl gpurun_out/engine_parity.jsonl
