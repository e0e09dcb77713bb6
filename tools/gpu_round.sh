#!/bin/bash
# One GPU-box visit: gpu test suite, smoke, bench (device + e2e + cpu baseline), ncu launch list + full captures.
# Pair-mode (tcgen05 cta_group::2) GEMM is exercised separately from the 1-CTA path so a failure in one does not
# take the other's numbers down:  B200SD_PAIR=0 never, =1 whenever legal, unset = heuristic.
mkdir -p gpurun_out
rm -f gpurun_out/kernel_parity.jsonl gpurun_out/engine_parity.jsonl
python stable-diffusion-webui-distributed_b200/b200sd/build.py > gpurun_out/build.log 2>&1
if [ "$SKIP_TESTS" != "1" ]; then
  B200SD_PAIR=0 timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
  B200SD_PAIR=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider --tb=line -k "linear or conv" 2>&1 | tail -30 > gpurun_out/pytest_pair.log
  timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -p no:cacheprovider --tb=line 2>&1 | tail -12 > gpurun_out/pytest_engine_heuristic.log
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
fi
B200SD_PAIR=0 timeout 900 python bench.py --gpus 1 --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_nopair.json 2> gpurun_out/bench_nopair.err
timeout 900 python bench.py --gpus 1 --steps ${BENCH_STEPS:-3} --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "$SKIP_NCU" != "1" ]; then
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_unet.csv python tools/ncu_cases.py unet > gpurun_out/ncu_unet.log 2>&1
  for c in ${NCU_CASES:-conv attn geglu proj gn}; do
    timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -o gpurun_out/prof_$c -f python tools/ncu_cases.py $c > gpurun_out/ncu_$c.log 2>&1
  done
fi
tail -3 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_pair.log; tail -2 gpurun_out/pytest_engine_heuristic.log
tail -1 gpurun_out/smoke.log; head -c 300 gpurun_out/bench_nopair.json; echo; head -c 300 gpurun_out/bench.json; echo; tail -2 gpurun_out/bench.err
