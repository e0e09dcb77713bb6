"""BASELINE config 1 with the UNMODIFIED reference dispatcher (build container only: /root/reference does not travel).

    python tools/ref_http_dispatch_container.py [--model sd15|tiny] [--steps 20] [--hw 64] > profiles/r02_reference_http_dispatch_container.json

The reference's own `DistributedScript.before_process -> World.optimize_jobs -> Worker.request (HTTP) ->
postprocess_batch_list -> postprocess` (scripts/distributed.py:185-357, scripts/spartan/worker.py:288-504), imported from
/root/reference under the host stub (tests/hoststub: `modules.*`, `gradio`; pydantic -> pydantic.v1), drives
  * the master's share in-process (hoststub process_images_inner -> the fp32 oracle), and
  * one worker over real HTTP: `bench.py --serve-cpu-oracle` = server/sdapi.py whose executor is the fp32 oracle,
for txt2img 512x512 batch 2, 20 DDIM timesteps, both on this container's host cores.  bench.py --impl reference runs the
same thing on the GPU box with this repo's mirror of the dispatcher; this file is the evidence that the mirror and the
original time the same path.  Separate process from everything else: the reference's module names (`scripts.*`) collide
with this repo's.
"""
import argparse
import json
import logging.handlers
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REFERENCE_DIR", "/root/reference")
tmp = tempfile.mkdtemp(prefix="refdispatch_")
os.environ["HOSTSTUB_CONFIG_DIR"] = tmp
# the reference's `scripts` package must win over this repo's: REF first, the repo's extension dir only for b200sd/server
sys.path[:0] = [os.path.join(ROOT, "tests", "hoststub"), REF, ROOT]

import pydantic.v1  # noqa: E402

sys.modules["pydantic"] = pydantic.v1
_Orig = logging.handlers.RotatingFileHandler


class _Redirected(_Orig):
    def __init__(self, filename, *a, **k):
        super().__init__(os.path.join(tmp, os.path.basename(str(filename))), *a, **k)


logging.handlers.RotatingFileHandler = _Redirected
import signal  # noqa: E402

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="sd15")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--hw", type=int, default=64)
    args = ap.parse_args()
    import modules.processing as processing
    import modules.scripts as mscripts
    from scripts import distributed as ref_distributed          # /root/reference/scripts/distributed.py
    from scripts.spartan import pmodels, shared as sh
    assert ref_distributed.__file__.startswith(REF), ref_distributed.__file__
    logging.getLogger("distributed").setLevel(logging.ERROR)
    sys.path.append(os.path.join(ROOT, "stable-diffusion-webui-distributed_b200"))   # b200sd only (after REF)
    import bench
    import requests
    cores = bench.usable_cpus()
    threads = max(1, cores // 2)
    port = bench._free_port()
    child = subprocess.Popen([sys.executable, os.path.join(ROOT, "bench.py"), "--serve-cpu-oracle", str(port), "--threads",
                              str(threads), "--model", args.model], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    try:
        for _ in range(600):
            try:
                if requests.get(f"http://127.0.0.1:{port}/sdapi/v1/memory", timeout=1).status_code == 200:
                    break
            except requests.RequestException:
                time.sleep(0.2)
        eng = bench.OracleCPUEngine(threads, args.model)
        from b200sd.factory import synthetic_tokens

        def master_generator(p, n):
            v = eng.clip_cfg.vocab
            u8 = eng.txt2img(synthetic_tokens([p.prompt] * p.batch_size, v), synthetic_tokens([p.negative_prompt] * p.batch_size, v),
                             p.seeds[0], steps=p.steps, cfg_scale=p.cfg_scale, height=p.height, width=p.width)
            return [u8[i].permute(2, 0, 1).float() / 255.0 for i in range(u8.shape[0])]

        processing.MASTER_GENERATOR = master_generator
        sh.benchmark_payload = pmodels.Benchmark_Payload()
        Script = ref_distributed.DistributedScript
        w = Script.world
        w.master().avg_ipm, w.master().benchmarked = 1.0, True
        wk = w.add_worker(address="127.0.0.1", port=port, label="cpu-worker", avg_ipm=1.0, master=False)
        wk.benchmarked = True
        w.benchmark = lambda *a, **k: None
        w.job_timeout = 3600
        script = Script()
        script.args_from = script.args_to = 0
        p = processing.StableDiffusionProcessingTxt2Img(
            prompt="a synthetic benchmark prompt", negative_prompt="", seed=1000, subseed=1, subseed_strength=0, batch_size=2,
            n_iter=1, steps=args.steps, width=args.hw * 8, height=args.hw * 8, sampler_name="DDIM", cfg_scale=7.0,
            scripts=mscripts.ScriptRunner([script]), script_args=[])
        t0 = time.perf_counter()
        out = processing.process_images(p)
        dt = time.perf_counter() - t0
        res = {"what": "UNMODIFIED reference dispatcher (" + ref_distributed.__file__ + ") under tests/hoststub: master + 1 HTTP "
                       "worker, both the fp32 oracle on host cores", "model": args.model, "batch": 2, "timesteps": args.steps,
               "resolution": f"{args.hw * 8}x{args.hw * 8}", "images_returned": len(out.images), "seeds": list(p.seeds),
               "sec_per_request": dt, "images_per_s": len(out.images) / dt, "cores": cores, "threads_per_process": threads,
               "torch": torch.__version__, "where": "build container (no GPU)"}
        print(json.dumps(res))
    finally:
        child.terminate()


if __name__ == "__main__":
    main()
