#!/usr/bin/env python
"""bench.py — images/sec of the batch-sharded SD1.5 txt2img path (BASELINE.json metric) on N GPUs of one node.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the numeric path on the box's host cores (oracle port)

One "step" = one whole txt2img request of the per-GPU batch: CLIP encode, 20 DDIM timesteps (19 UNet evaluations
on [cond | uncond]) and the VAE decode to uint8, plus — for N > 1 — the single NCCL all-gather of the images.
Ranks own disjoint image indices (seed + k), weights are replicated, there is no per-step collective: scaling "weak"
(per-GPU batch fixed at --per-gpu-batch, default 32 = BASELINE.json configs[1]'s batch on one GPU).

  value : images/s with prompts/noise already resident in HBM (device-timed with CUDA events, max over ranks)
  e2e   : the same request issued through the reference-facing plugin surface — DistributedScript hooks driving a
          LocalGPUWorker.request() — from HOST buffers: prompt strings/tokens and per-image CPU-RNG noise are
          uploaded, decoded uint8 images are copied back into `worker.response` inside the timed region
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
EXT = os.path.join(ROOT, "stable-diffusion-webui-distributed_b200")
for p in (ROOT, EXT, os.path.join(ROOT, "tests", "hoststub")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

STEPS_DDIM = 20
CFG_SCALE = 7.0
HW = 64  # 512 x 512 images

# algorithmic FLOPs (SURVEY.md App. D / BASELINE.md §3), TFLOP
UNET_TFLOP_PER_SAMPLE_EVAL = 0.8033
VAE_TFLOP_PER_IMAGE = 2.5145
CLIP_TFLOP_PER_SEQ = 0.0133
ATTN_TFLOP_PER_SAMPLE_EVAL = 0.1225


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"], "hbm_gbs": d["hbm_gbs"],
                "source": "measured"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def usable_cpus() -> int:
    """host cores this process may really use: affinity mask, capped by the cgroup CPU quota (containers often
    advertise every core of the host in os.cpu_count() while being limited to a few)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, q // per))
        except (OSError, ValueError, IndexError):
            continue
    return max(1, n)


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_arm(n_evals: int, threads: int, unet_reps: int = 1):
    """The reference's numeric path on host cores: the fp32 oracle (oracle/sd_oracle.py, kind "port").
    Bounded sample: `unet_reps` UNet evaluations on [cond|uncond] of ONE image at 64x64 latents + one VAE decode + two
    CLIP encodes; images/s = 1 / (n_evals * t_unet + t_vae + t_clip)."""
    from b200sd import config as C, synth
    from oracle import sd_oracle as O
    torch.set_num_threads(threads)
    cfgs = (C.SD15_UNET, C.SD15_VAE, C.SD15_CLIP)
    sd = synth.make_state_dict(*cfgs, seed=0)
    tok = O.random_prompt_tokens(1)
    neg = O.empty_prompt_tokens(1)
    with torch.no_grad():
        t0 = time.perf_counter()
        cond = O.clip_text_encode(sd, cfgs[2], tok)
        unc = O.clip_text_encode(sd, cfgs[2], neg)
        t_clip = time.perf_counter() - t0
        x = O.per_image_noise(1000, 1, (4, HW, HW))
        ts = []
        for _ in range(unet_reps):
            t0 = time.perf_counter()
            O.cfg_eps(lambda a, t, c: O.unet_forward(sd, cfgs[0], a, t, c), x, 651, cond, unc, CFG_SCALE)
            ts.append(time.perf_counter() - t0)
        t_unet = min(ts)
        t0 = time.perf_counter()
        O.to_uint8(O.vae_decode(sd, cfgs[1], x))
        t_vae = time.perf_counter() - t0
    sec_per_image = n_evals * t_unet + t_vae + t_clip
    return {"value": 1.0 / sec_per_image, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"1 image: {unet_reps} CFG UNet eval(s) ({t_unet:.2f}s each) x {n_evals} + 1 VAE decode ({t_vae:.2f}s) + "
                      f"2 CLIP encodes ({t_clip:.2f}s), fp32 torch on {threads} host threads, composed to 1 image",
            "sec_per_image": sec_per_image}


# --------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--per-gpu-batch", type=int, default=32)
    ap.add_argument("--model", default="sd15", choices=["sd15", "tiny"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank, world, local = dist_env()
    n_evals = STEPS_DDIM - 1
    workload = (f"SD1.5 txt2img 512x512 fp16, {STEPS_DDIM} DDIM timesteps = {n_evals} CFG UNet evaluations + VAE decode, "
                f"per-GPU batch {args.per_gpu_batch}, batch-sharded by image index, synthetic seeded weights, "
                f"random-token prompts")
    config = {"workload": workload, "per_gpu_batch": args.per_gpu_batch, "global_batch": args.per_gpu_batch * world,
              "resolution": "512x512", "sampler": "DDIM", "timesteps": STEPS_DDIM, "unet_evals": n_evals,
              "cfg_scale": CFG_SCALE, "parallelism": f"dp{world} (batch index sharding, one all-gather at the end)",
              "l2": "every step streams far more than the 126 MB L2 (activations of one UNet eval at batch 64 exceed 10 GB)"}

    if args.impl == "reference":
        if rank != 0:
            return
        threads = usable_cpus()
        t0 = time.perf_counter()
        best = None
        for _ in range(max(1, min(args.steps, 2))):
            r = cpu_reference_arm(n_evals, threads, unet_reps=1)
            best = r if best is None or r["value"] > best["value"] else best
        line = {"impl": "reference", "metric": "images/sec SD1.5 512x512 txt2img", "value": best["value"],
                "unit": "images/s", "n_gpus": args.gpus, "gpus_used": 0, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": best["sec_per_image"] * 1000.0, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": config,
                "cpu_baseline": {k: best[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": best["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "wall_s": time.perf_counter() - t0}
        print(json.dumps(line))
        return

    import torch.distributed as dist
    from b200sd import engine as E, factory, ops
    from b200sd.sharding import all_gather_images
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    b = args.per_gpu_batch
    eng = factory.default_engine_factory(dev, args.model)
    vocab = eng.clip_cfg.vocab
    g = torch.Generator().manual_seed(1234 + rank)
    tokens = torch.cat([torch.full((b, 1), vocab - 2), torch.randint(0, vocab - 3, (b, 75), generator=g),
                        torch.full((b, 1), vocab - 1)], dim=1)
    neg = torch.full((b, 77), vocab - 1)
    neg[:, 0] = vocab - 2
    seed0 = 1000 + rank * b   # global image index -> seed (reference: seed + images owned by earlier jobs)
    x_T = E.per_image_noise(seed0, b, (4, HW, HW))[0]
    tokens_d, neg_d, x_T_d = tokens.to(dev), neg.to(dev), x_T.to(dev)

    def step_device():
        cond = eng.encode_prompts(tokens_d)
        unc = eng.encode_prompts(neg_d)
        lat = eng.sample(cond, unc, x_T_d, STEPS_DDIM, CFG_SCALE, "DDIM")
        u8 = eng.decode(lat, HW, HW)
        if world > 1:
            all_gather_images(u8, [b] * world)
        return u8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_device()
    barrier()
    launches0 = ops.LAUNCHES
    replays0 = eng.graph_replayed_launches
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    gpu_launches = (ops.LAUNCHES - launches0) + (eng.graph_replayed_launches - replays0)
    t = torch.tensor([elapsed_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * b * args.steps / (elapsed_ms / 1000.0)

    # ---------------- e2e through the plugin surface (host buffers, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        import modules.processing as processing
        import modules.scripts as mscripts
        from modules.shared import cmd_opts
        import logging
        from scripts.distributed import DistributedScript
        from scripts.spartan import pmodels, shared as sh
        logging.getLogger("distributed").setLevel(logging.ERROR)
        w = DistributedScript.world
        sh.benchmark_payload = pmodels.Benchmark_Payload()
        wk = w.add_local_gpus(lambda d: eng, devices=[local], avg_ipm=600.0)[0]
        wk.benchmarked = True
        w.thin_client_mode = True
        w.benchmark = lambda *a, **k: None      # speeds are set above; do not re-benchmark inside the timed region
        script = DistributedScript()
        script.args_from = script.args_to = 0

        def step_plugin():
            p = processing.StableDiffusionProcessingTxt2Img(
                prompt="synthetic", negative_prompt="", seed=seed0, subseed=1, subseed_strength=0, batch_size=b, n_iter=1,
                steps=STEPS_DDIM, width=HW * 8, height=HW * 8, sampler_name="DDIM", cfg_scale=CFG_SCALE,
                scripts=mscripts.ScriptRunner([script]), script_args=[])
            p.prompt_tokens = tokens.tolist()       # host token ids ride along in the payload (p.__dict__)
            return processing.process_images(p)

        for _ in range(2):
            out = step_plugin()
        assert len(out.images) == b, len(out.images)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_plugin()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * b * args.steps / dt, "unit": "images/s",
               "h2d_bytes_per_step": int(tokens.numel() * 8 + neg.numel() * 8 + x_T.numel() * 4),
               "d2h_bytes_per_step": int(b * HW * 8 * HW * 8 * 3),
               "path": "hoststub process_images -> DistributedScript.before_process -> LocalGPUWorker.request -> "
                       "postprocess_batch_list -> postprocess (thin-client world, 1 local GPU per rank)"}

    # ---------------- roofline of the dominant kernel: per-launch CUDA-event timing of one eager UNet evaluation
    roof, roof_attn, breakdown = kernel_rooflines(eng, b, peaks())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    tflop_per_image = 2 * n_evals * UNET_TFLOP_PER_SAMPLE_EVAL + VAE_TFLOP_PER_IMAGE + 2 * CLIP_TFLOP_PER_SEQ
    line = {
        "metric": "images/sec SD1.5 512x512 txt2img", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "config": config, "clocks": clk, "gpu_launches": int(gpu_launches), "e2e": e2e,
        "roofline": roof, "roofline_attention": roof_attn, "unet_eval_breakdown_ms": breakdown,
        "step_roofline": {"bound": "tensor", "achieved": value * tflop_per_image / world, "peak": pk["tflops_sustained"],
                          "unit": "TFLOP/s", "frac": value * tflop_per_image / world / pk["tflops_sustained"],
                          "tflop_per_image": tflop_per_image, "peak_source": pk["source"] + " (sustained)"},
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = {k: v for k, v in cpu_reference_arm(n_evals, usable_cpus()).items()
                                if k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def kernel_rooflines(eng, b, pk):
    """Per-launch CUDA-event durations of every kernel class in ONE eager UNet evaluation (same shapes as the timed
    region).  roofline = the dominant kernel (gemm_conv_tc_kernel: all convs and linears), algorithmic FLOPs summed
    over its launches / summed duration, against the measured sustained bf16 tensor peak."""
    from b200sd import ops
    plan = eng.plan(b, HW, HW)
    ops.select_step(plan.table, plan.step * 0, plan.unet.cur_bias)
    torch.cuda.synchronize()
    recs = []
    exps = [0.0]   # exponentials of all attention launches (one per S element)
    for fn, a, k in plan.unet.ops:
        name = getattr(fn, "__name__", "op")
        name = "groupnorm" if name == "<lambda>" else name
        flop = 0.0
        if name == "linear":
            m = a[0].numel() // a[0].shape[-1]
            flop = 2.0 * m * a[1].shape[0] * a[1].shape[1]
        elif name == "conv2d":
            m = a[2].numel() // a[2].shape[-1]
            flop = 2.0 * m * a[1].shape[0] * a[1].shape[1]
        elif name == "attention":
            bq, sq, skv, heads, d = a[0].shape[0], a[0].shape[1], a[1].shape[1], a[4], a[5]
            flop = 4.0 * bq * heads * sq * skv * d
            exps[0] += float(bq) * heads * sq * skv
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(*a, **k)
        e1.record()
        recs.append((name, flop, e0, e1))
    torch.cuda.synchronize()
    agg = {}
    for name, flop, e0, e1 in recs:
        d = agg.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1; d[1] += e0.elapsed_time(e1); d[2] += flop
    tc = [agg.get("linear", [0, 0, 0]), agg.get("conv2d", [0, 0, 0])]
    n_l, ms, fl = (sum(x[i] for x in tc) for i in range(3))
    peak = pk["tflops_sustained"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("gemm_conv_tc_kernel_dram_bytes_per_launch")
    roof = {"kernel": "gemm_conv_tc_kernel (all conv2d + linear launches of one UNet evaluation)", "bound": "tensor",
            "achieved": fl / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": fl / (ms * 1e-3) / 1e12 / peak,
            "traffic": traffic, "launches": n_l, "avg_launch_us": ms * 1e3 / max(1, n_l),
            "algorithmic_flop_per_launch": fl / max(1, n_l), "peak_source": pk["source"] + " sustained bf16 cuBLAS"}
    at = agg.get("attention", [0, 1e-9, 0])
    roof_attn = {"kernel": "attention_tc_kernel", "bound": "tensor", "achieved": at[2] / (at[1] * 1e-3) / 1e12, "peak": peak,
                 "unit": "TFLOP/s", "frac": at[2] / (at[1] * 1e-3) / 1e12 / peak, "launches": at[0],
                 "note": "QK^T + PV FLOPs; d=40 heads make this kernel exp-throughput (MUFU) bound, see DESIGN.md",
                 # the pipe that actually bounds it: one MUFU.EX2 per S element, 16 per clock and SM (tools/xu_probe.cu
                 # measured 4.47 T/s at 1.9 GHz on 148 SMs)
                 "exp_rate": {"achieved": exps[0] / (at[1] * 1e-3) / 1e12, "peak": 4.47, "unit": "T exp/s",
                              "frac": exps[0] / (at[1] * 1e-3) / 1e12 / 4.47, "peak_source": "measured (xu_probe, boost clock)"}}
    breakdown = {k: round(v[1], 3) for k, v in agg.items()}
    # HBM-bound kernel classes: algorithmic bytes (GroupNorm 2 reads + 1 write of 45.1 M elements per
    # sample-evaluation, LayerNorm 1 read + 1 write of 34.7 M) over the summed CUDA-event durations, against the measured
    # copy bandwidth
    hbm = {}
    for name, elems, bpe in (("groupnorm", 45.1e6, 6), ("layernorm", 34.7e6, 4)):
        if name in agg and agg[name][1] > 0:
            gbs = elems * 2 * b * bpe / (agg[name][1] * 1e-3) / 1e9
            hbm[name] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                         "launches": agg[name][0], "algorithmic_bytes": elems * 2 * b * bpe}
    roof["hbm_kernels"] = hbm
    return roof, roof_attn, breakdown


if __name__ == "__main__":
    main()
