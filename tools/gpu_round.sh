#!/bin/bash
# One GPU-box visit: full gpu test suite, smoke, bench (device + e2e + cpu baseline), ncu launch list + full captures.
mkdir -p gpurun_out
rm -f gpurun_out/kernel_parity.jsonl gpurun_out/engine_parity.jsonl
if [ "$SKIP_TESTS" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
fi
timeout 900 python bench.py --gpus 1 --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "$SKIP_NCU" != "1" ]; then
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_unet.csv python tools/ncu_cases.py unet > gpurun_out/ncu_unet.log 2>&1
  for c in ${NCU_CASES:-conv attn geglu proj gn}; do
    timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -o gpurun_out/prof_$c -f python tools/ncu_cases.py $c > gpurun_out/ncu_$c.log 2>&1
  done
fi
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; head -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
