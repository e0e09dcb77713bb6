#!/bin/bash
# 1-GPU visit: whole GPU suite, smoke, bench (+ batch sweep artefact), img2img workload line, SDXL line
mkdir -p gpurun_out
python stable-diffusion-webui-distributed_b200/b200sd/build.py > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -12 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --gpus 1 --steps ${BENCH_STEPS:-3} --warmup 3 --sweep-out gpurun_out/sweep_n1.json --sweep-batches 1,2,4,8,16,64 > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "$EXTRA" = "1" ]; then
  timeout 600 python bench.py --workload img2img --steps 3 --warmup 3 --no-cpu-baseline --no-stock > gpurun_out/bench_img2img.json 2> gpurun_out/bench_img2img.err
  timeout 900 python bench.py --model sdxl --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sdxl.json 2> gpurun_out/bench_sdxl.err
fi
tail -3 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; head -c 300 gpurun_out/bench.json; echo
[ "$EXTRA" = "1" ] && head -c 250 gpurun_out/bench_img2img.json && echo && head -c 250 gpurun_out/bench_sdxl.json && echo
exit 0
