#!/bin/bash
# attention experiment visit: rebuild the library with each B200SD_ATTN_H2POLY_MASK and time the dominant shape
mkdir -p gpurun_out; : > gpurun_out/attn_exp.log
for m in ${MASKS:-0 0x1111 0x5252 0x5555 0x7777}; do
  B200SD_NVCC_EXTRA="-DB200SD_ATTN_H2POLY_MASK=$m" python stable-diffusion-webui-distributed_b200/b200sd/build.py --force > gpurun_out/build.log 2>&1
  echo "mask $m: $(timeout 120 python tools/attn_sweep.py --one 2>&1 | tail -1)" | tee -a gpurun_out/attn_exp.log
  ATTN_NB=64 timeout 120 python tools/attn_sweep.py --one 2>&1 | tail -1 | sed "s/^/   nb=64 /" | tee -a gpurun_out/attn_exp.log
done
