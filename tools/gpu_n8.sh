#!/bin/bash
# N-GPU visit (gpurun --gpus N): multi-GPU World test + bench under torchrun with world keys and the batch sweep
N=${NGPU:-8}
mkdir -p gpurun_out
python stable-diffusion-webui-distributed_b200/b200sd/build.py > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/mgpu_n${N}.log
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -s -p no:cacheprovider --tb=short 2>&1 | tail -12 >> gpurun_out/mgpu_n${N}.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $N --steps ${BENCH_STEPS:-3} --warmup 3 --sweep-out gpurun_out/sweep_n${N}.json --sweep-batches ${SWEEP:-1,2,4,8,16} \
  > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
tail -4 gpurun_out/mgpu_n${N}.log; head -c 400 gpurun_out/bench_n${N}.json; echo; grep -v "^DISTRIBUTED\|SD_CKPT\|SD_TOKENIZER" gpurun_out/bench_n${N}.err | tail -8
