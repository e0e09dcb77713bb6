#!/bin/bash
# GroupNorm experiments of round 2: parity tests of both paths, the per-shape timing table (two kernels with the statistics
# grid walking forward / backward, one pass), then the whole step with the statistics order forward vs backward.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gn_gpu.txt 2>&1
timeout 400 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "groupnorm" -p no:cacheprovider --tb=short 2>&1 | tail -n 40 | tee gpurun_out/gn_tests.log
timeout 200 python tools/norm_sweep.py --gn-modes 48 2>&1 | tee gpurun_out/gn_modes.txt
if [ -z "$SKIP_BENCH" ]; then
  for rev in 0 1; do
    B200SD_GN_REVERSE=$rev timeout 300 python bench.py --steps 3 --warmup 3 --no-world --no-stock --no-cpu-baseline --no-e2e \
      > gpurun_out/gn_bench_rev$rev.json 2> gpurun_out/gn_bench_rev$rev.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/gn_bench_rev$rev.json").read().strip().splitlines()[-1])
print("reverse=$rev value", round(d["value"], 3), "ms/step", round(d["ms_per_step"], 1), "clk", d["clocks"].get("sm_mhz"),
      "gn", d["roofline"]["hbm_kernels"].get("groupnorm", {}).get("ms"), "breakdown", d.get("unet_eval_breakdown_ms"))
PY
  done
fi
