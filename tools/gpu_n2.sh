#!/bin/bash
# 2-GPU visit: the quick re-check tests + bench at N=2 under torchrun (value, e2e, world_e2e, strong_scaling)
mkdir -p gpurun_out
python stable-diffusion-webui-distributed_b200/b200sd/build.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider --tb=short \
  -k "PLMS or inpainting or euler_a_graph or multi or sharded" 2>&1 | tail -15 > gpurun_out/pytest_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps ${BENCH_STEPS:-3} --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -5 gpurun_out/pytest_n2.log; head -c 600 gpurun_out/bench_n2.json; echo; tail -5 gpurun_out/bench_n2.err
