#!/bin/bash
# The most tolerance-sensitive SD1.5-size parity tests + the third-party VAE test + smoke(), for a short GPU visit.
mkdir -p gpurun_out
timeout 130 python -m pytest tests/test_engine_gpu.py -m gpu -q -p no:cacheprovider --tb=short --durations=5 \
  -k "third_party or vae_decode_parity or unet_eval_parity or tight_unet or txt2img_parity or img2img_parity" 2>&1 | tail -25 > gpurun_out/pytest_subset.log
timeout 60 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
tail -25 gpurun_out/pytest_subset.log; tail -1 gpurun_out/smoke.log
exit 0
