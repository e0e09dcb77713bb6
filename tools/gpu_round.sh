#!/bin/bash
# One GPU-box visit: gpu test suite, smoke, bench (device + e2e + cpu baseline), ncu launch list + full captures.
# Pair-mode (tcgen05 cta_group::2) GEMM is opt-in (B200SD_PAIR=1) and exercised separately so a failure there does not
# take the default path's numbers down.
mkdir -p gpurun_out
rm -f gpurun_out/kernel_parity.jsonl gpurun_out/engine_parity.jsonl
python stable-diffusion-webui-distributed_b200/b200sd/build.py > gpurun_out/build.log 2>&1
if [ "$SKIP_TESTS" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
  B200SD_PAIR=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider --tb=line -k "linear or conv" 2>&1 | tail -30 > gpurun_out/pytest_pair.log
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
fi
timeout 900 python bench.py --gpus 1 --steps ${BENCH_STEPS:-3} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
if [ "$SWEEPS" = "1" ]; then
  timeout 300 python tools/attn_sweep.py --one > gpurun_out/attn_sweep.log 2>&1
  timeout 300 python tools/gemm_sweep.py --one > gpurun_out/gemm_sweep.log 2>&1
  B200SD_PAIR=1 timeout 300 python tools/gemm_sweep.py --one >> gpurun_out/gemm_sweep.log 2>&1
fi
if [ "$SKIP_NCU" != "1" ]; then
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_unet.csv python tools/ncu_cases.py unet > gpurun_out/ncu_unet.log 2>&1
  for c in ${NCU_CASES:-conv attn geglu proj gn}; do
    timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -o gpurun_out/prof_$c -f python tools/ncu_cases.py $c > gpurun_out/ncu_$c.log 2>&1
  done
fi
if [ "$TRAFFIC" = "1" ]; then
  # DRAM bytes of every GEMM/conv launch of one UNet evaluation at bench.py's batch -> roofline.traffic
  NCU_NB=64 timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none -k regex:gemm_conv --csv --log-file gpurun_out/traffic_gemm.csv \
      python tools/ncu_cases.py unet > gpurun_out/ncu_traffic.log 2>&1
fi
tail -3 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_pair.log
tail -1 gpurun_out/smoke.log; head -c 300 gpurun_out/bench.json; echo; tail -2 gpurun_out/bench.err
if [ "$SWEEPS" = "1" ]; then cat gpurun_out/attn_sweep.log gpurun_out/gemm_sweep.log; fi
exit 0
