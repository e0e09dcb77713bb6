#!/bin/bash
# PDL visit: full GPU suite with programmatic dependent launch on, then the batch sweep with it on and off
mkdir -p gpurun_out
python stable-diffusion-webui-distributed_b200/b200sd/build.py > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short -x --deselect tests/test_sdxl_gpu.py::test_sdxl_base_parity_1024 2>&1 | tail -15 > gpurun_out/pytest_pdl.log
tail -4 gpurun_out/pytest_pdl.log
for pdl in 1 0; do
  B200SD_PDL=$pdl timeout 600 python bench.py --sweep ${SWEEP:-1,4,32} --steps 3 --warmup 3 > gpurun_out/sweep_pdl$pdl.json 2> gpurun_out/sweep_pdl$pdl.err
  echo "PDL=$pdl"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sweep_pdl$pdl.json').read().strip().splitlines()[-1])
    for r in d['sweep']: print(r['per_gpu_batch'], round(r['value'],2), 'img/s', round(r['ms_per_step'],1), 'ms', r['clocks']['sm_mhz'])
except Exception as e: print('ERR', e); print(open('gpurun_out/sweep_pdl$pdl.err').read()[-1500:])
PY
done
timeout 1200 python -m pytest tests/test_sdxl_gpu.py -m gpu -q -p no:cacheprovider --tb=short -k sdxl_base 2>&1 | tail -8 > gpurun_out/pytest_sdxl.log
tail -4 gpurun_out/pytest_sdxl.log; grep -i xl gpurun_out/engine_parity.jsonl
