"""SDEngine — the local executor a Worker binds instead of POST /sdapi/v1/txt2img (reference worker.py:432-435).

One engine per GPU: packed weights, per-shape execution plans (UNet step graph, VAE decode graph), samplers.
Images of one request are independent given (prompt, seed + k) — the property the reference's seed offsetting
relies on (scripts/distributed.py:297-305) — so a request's batch is sharded across engines with no per-step
exchange; results meet once, at the end (bench: one NCCL all-gather; plugin path: the collector thread join).
"""
import contextlib
import math
import threading
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .clip_text import ClipText
from .config import CLIPConfig, UNetConfig, VAEConfig
from .unet_exec import TimeEmbedding, UNetProgram, UNetWeights
from .vae_exec import VAEDecoderProgram, VAEDecoderWeights, VAEEncoderProgram, VAEEncoderWeights

MAX_STEPS = 256
_CAPTURE_LOCK = threading.Lock()  # CUDA graph captures are serialised across the per-device worker threads


# ------------------------------------------------------------------------------------------------ schedules
def alphas_cumprod() -> torch.Tensor:
    """ldm linear schedule (linear_start 0.00085, linear_end 0.012, 1000 steps), fp64 math, fp32 storage."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    return torch.cumprod(1.0 - betas, dim=0).to(torch.float32)


def ddim_plan(steps: int) -> Tuple[List[float], List[List[float]]]:
    """sdwui sd_samplers_timesteps(_impl).ddim, eta = 0.  Returns (timesteps, coef rows) in execution order;
    `steps` timesteps give steps-1 UNet evaluations (index len-1 .. 1)."""
    ac = alphas_cumprod().double()
    ts = torch.clamp(torch.arange(0, 1000, 1000 // steps) + 1, 0, 999)
    a = ac[ts]
    a_prev = ac[torch.cat([ts.new_zeros(1), ts[:-1]])]
    t_out, rows = [], []
    for i in range(len(ts) - 1, 0, -1):
        at, ap = float(a[i]), float(a_prev[i])
        t_out.append(float(ts[i]))
        rows.append([math.sqrt(at), math.sqrt(1 - at), math.sqrt(ap), math.sqrt(1 - ap)])
    return t_out, rows


def ddim_img2img_plan(steps: int, denoising_strength: float):
    """sdwui sd_samplers_timesteps.sample_img2img: t_enc = int(min(d, 0.999) * steps); DDIM runs on timesteps[:t_enc]
    (t_enc - 1 UNet evaluations) from x = init * sqrt(a[ts[t_enc]]) + noise * sqrt(1 - a[ts[t_enc]]).
    Returns (sqrt_a_start, sqrt_1m_a_start, timesteps, coef rows)."""
    ac = alphas_cumprod().double()
    ts = torch.clamp(torch.arange(0, 1000, 1000 // steps) + 1, 0, 999)
    t_enc = max(1, min(int(min(denoising_strength, 0.999) * steps), len(ts) - 1))
    a_start = float(ac[ts[t_enc]])
    sub = ts[:t_enc]
    a = ac[sub]
    a_prev = ac[torch.cat([sub.new_zeros(1), sub[:-1]])]
    t_out, rows = [], []
    for i in range(len(sub) - 1, 0, -1):
        at, ap = float(a[i]), float(a_prev[i])
        t_out.append(float(sub[i]))
        rows.append([math.sqrt(at), math.sqrt(1 - at), math.sqrt(ap), math.sqrt(1 - ap)])
    return math.sqrt(a_start), math.sqrt(1 - a_start), t_out, rows


def euler_a_plan(steps: int, scheduler: str = "uniform", sigmas=None):
    """k-diffusion sample_euler_ancestral over CompVisDenoiser sigmas.  Returns (timesteps, coef rows, sigma0):
    row = [sigma, sigma_down, sigma_up, c_in of the NEXT step].  `sigmas` = (table ending in 0, log-sigma table of the
    model) runs an explicit schedule instead (img2img: the tail of the full one)."""
    sig, log_sig = sigmas if sigmas is not None else kdiffusion_sigmas(steps, scheduler)
    steps = len(sig) - 1

    t_out, rows = [], []
    for i in range(steps):
        s, sn = float(sig[i]), float(sig[i + 1])
        up = min(sn, (sn ** 2 * (s ** 2 - sn ** 2) / s ** 2) ** 0.5)
        down = (sn ** 2 - up ** 2) ** 0.5
        t_out.append(sigma_to_t(s, log_sig))
        rows.append([s, down, up, 1.0 / math.sqrt(sn * sn + 1.0)])
    return t_out, rows, float(sig[0])


def euler_plan(steps: int, scheduler: str = "uniform", sigmas=None):
    """k-diffusion sample_euler (s_churn = 0) on the same sigmas: the ancestral step with sigma_up = 0, i.e.
    sigma_down = sigma_next and no noise — same coefficient rows, same kernel."""
    t_out, rows, sigma0 = euler_a_plan(steps, scheduler, sigmas)
    for i, r in enumerate(rows):
        sn = (r[1] ** 2 + r[2] ** 2) ** 0.5   # sigma_next = sqrt(down^2 + up^2)
        rows[i] = [r[0], sn, 0.0, r[3]]
    return t_out, rows, sigma0


def kdiffusion_sigmas(steps: int, scheduler: str = "uniform") -> Tuple[torch.Tensor, torch.Tensor]:
    """sigmas[steps + 1] (last 0) as sdwui's KDiffusionSampler.get_sigmas builds them, and the model's log-sigma table.
    "uniform": k-diffusion DiscreteSchedule.get_sigmas (timesteps linspace(999, 0, steps), log-sigma interpolation);
    "karras": get_sigmas_karras(steps, sigma_min = sigmas[0], sigma_max = sigmas[-1], rho = 7);
    "exponential" (= "polyexponential", rho 1): get_sigmas_exponential — log-sigmas equally spaced between the same bounds;
    "sgm_uniform": sdwui sd_schedulers.sgm_uniform — timesteps linspace(t(sigma_max), t(sigma_min), steps + 1)[:-1]."""
    ac = alphas_cumprod().double()
    sig_all = ((1 - ac) / ac) ** 0.5
    log_sig = sig_all.log()
    if scheduler == "karras":
        ramp = torch.linspace(0, 1, steps, dtype=torch.float64)
        lo, hi = float(sig_all[0]) ** (1 / 7.0), float(sig_all[-1]) ** (1 / 7.0)
        sig = (hi + ramp * (lo - hi)) ** 7.0
    elif scheduler in ("exponential", "polyexponential"):
        sig = torch.linspace(math.log(float(sig_all[-1])), math.log(float(sig_all[0])), steps, dtype=torch.float64).exp()
    elif scheduler in ("uniform", "sgm_uniform"):
        # t(sigma_max) = 999 and t(sigma_min) = 0 on the model's own table
        tt = torch.linspace(999, 0, steps + 1, dtype=torch.float64)[:-1] if scheduler == "sgm_uniform" else \
            torch.linspace(999, 0, steps, dtype=torch.float64)
        lo, hi = tt.floor().long(), tt.ceil().long()
        wgt = tt - lo
        sig = ((1 - wgt) * log_sig[lo] + wgt * log_sig[hi]).exp()
    else:
        raise ValueError(f"scheduler {scheduler!r} is not implemented")
    return torch.cat([sig, sig.new_zeros(1)]), log_sig


def sigma_to_t(s: float, log_sig: torch.Tensor) -> float:
    """k-diffusion DiscreteSchedule.sigma_to_t (quantize = False)"""
    ls = math.log(s)
    dists = ls - log_sig
    low = int((dists >= 0).cumsum(0).argmax().clamp(max=len(log_sig) - 2))
    l0, h0 = float(log_sig[low]), float(log_sig[low + 1])
    w = min(max((l0 - ls) / (l0 - h0), 0.0), 1.0)
    return (1 - w) * low + w * (low + 1)


def kdiffusion_img2img_sigmas(steps: int, denoising_strength: float, scheduler: str):
    """sdwui KDiffusionSampler.sample_img2img: t_enc = int(min(d, 0.999) * steps); the sampler runs on
    sigmas[steps - t_enc - 1:] (t_enc + 1 UNet evaluations) from x = init + noise * sigma_sched[0]."""
    sig, log_sig = kdiffusion_sigmas(steps, scheduler)
    t_enc = int(min(denoising_strength, 0.999) * steps)
    return sig[steps - t_enc - 1:], log_sig


def dpmpp_2m_plan(steps: int, scheduler: str = "karras", sigmas=None):
    """k-diffusion sample_dpmpp_2m.  Returns (timesteps, coef rows of 8, sigma0);
    row = [sigma, sigma_next / sigma, c1, c2, c_in of the NEXT step, 0, 0, 0] with denoised_d = c1 * denoised - c2 * old."""
    sig, log_sig = sigmas if sigmas is not None else kdiffusion_sigmas(steps, scheduler)
    steps = len(sig) - 1
    t_out, rows = [], []
    for i in range(steps):
        s, sn = float(sig[i]), float(sig[i + 1])
        c1, c2 = 1.0, 0.0
        if i > 0 and sn > 0:
            h = math.log(s / sn)
            h_last = math.log(float(sig[i - 1]) / s)
            r = h_last / h
            c1, c2 = 1.0 + 1.0 / (2.0 * r), 1.0 / (2.0 * r)
        t_out.append(sigma_to_t(s, log_sig))
        rows.append([s, sn / s, c1, c2, 1.0 / math.sqrt(sn * sn + 1.0), 0.0, 0.0, 0.0])
    return t_out, rows, float(sig[0])


# sampler names of the sdwui API -> (method, scheduler).  sdwui >= 1.9 sends the scheduler separately ("scheduler" key,
# "Automatic" = the sampler's default, which is Karras for DPM++ 2M); older versions fold it into the name.
SAMPLERS = {"DDIM": ("ddim", None), "Euler a": ("euler_a", "uniform"), "Euler": ("euler", "uniform"),
            "DPM++ 2M": ("dpmpp_2m", "karras"), "DPM++ 2M Karras": ("dpmpp_2m", "karras")}


# API scheduler labels (sdwui >= 1.9 sd_schedulers.schedulers: label or name) -> kdiffusion_sigmas scheduler
SCHEDULERS = {"Uniform": "uniform", "uniform": "uniform", "Karras": "karras", "karras": "karras",
              "Exponential": "exponential", "exponential": "exponential", "Polyexponential": "polyexponential",
              "polyexponential": "polyexponential", "SGM Uniform": "sgm_uniform", "sgm_uniform": "sgm_uniform"}


def resolve_sampler(name: str, scheduler: Optional[str] = None):
    """(method, scheduler) for an API sampler name + optional API scheduler label; ValueError if not implemented."""
    if name not in SAMPLERS:
        raise ValueError(f"sampler {name!r} is not implemented on the local executor")
    method, default = SAMPLERS[name]
    if default is not None and scheduler not in (None, "", "Automatic"):   # DDIM takes no sigma schedule
        label = SCHEDULERS.get(scheduler, SCHEDULERS.get(str(scheduler).lower()))
        if label is None:
            raise ValueError(f"scheduler {scheduler!r} is not implemented on the local executor")
        return method, label
    return method, default


def slerp(val: float, low: torch.Tensor, high: torch.Tensor) -> torch.Tensor:
    """sdwui modules/rng.py slerp, applied to one image's [C, H, W] noise (norms and angles along dim 1, as upstream)"""
    low_norm = low / torch.norm(low, dim=1, keepdim=True)
    high_norm = high / torch.norm(high, dim=1, keepdim=True)
    dot = (low_norm * high_norm).sum(1)
    if float(dot.mean()) > 0.9995:
        return low * val + high * (1 - val)
    omega = torch.acos(dot)
    so = torch.sin(omega)
    return (torch.sin((1.0 - val) * omega) / so).unsqueeze(1) * low + (torch.sin(val * omega) / so).unsqueeze(1) * high


def per_image_noise(seed: int, n: int, shape, draws: int = 1, subseed: Optional[int] = None,
                    subseed_strength: float = 0.0) -> torch.Tensor:
    """sdwui ImageRNG with randn_source = 'CPU': image k owns torch.Generator('cpu').manual_seed(seed + k);
    `draws` successive tensors per image (x_T, then ancestral noises).  Variation seeds (ImageRNG.first): with a
    non-zero `subseed_strength` the FIRST draw is slerp(strength, noise(seed + k), noise(subseed + k)).
    Returns [draws, n, *shape] fp32 (host)."""
    out = torch.empty((draws, n, *shape), dtype=torch.float32)
    for k in range(n):
        g = torch.Generator(device="cpu").manual_seed(int(seed) + k)
        for d in range(draws):
            out[d, k] = torch.randn(shape, generator=g, dtype=torch.float32)
        if subseed is not None and subseed_strength != 0:
            sg = torch.Generator(device="cpu").manual_seed(int(subseed) + k)
            out[0, k] = slerp(float(subseed_strength), out[0, k], torch.randn(shape, generator=sg, dtype=torch.float32))
    return out


# ------------------------------------------------------------------------------------------------ plans
class Plan:
    """Everything shape-dependent for (images per call b, latent h x w) on one device."""

    def __init__(self, eng: "SDEngine", b: int, h: int, w: int, vae_chunk: int):
        dev = eng.device
        self.b, self.h, self.w = b, h, w
        self.unet = UNetProgram(eng.unet_w, 2 * b, h, w)
        self.vae_chunk = min(vae_chunk, b)
        self.vae = VAEDecoderProgram(eng.vae_w, self.vae_chunk, h, w)
        self.x = torch.zeros((b, h * w, 4), device=dev, dtype=torch.float32)
        self.step = torch.zeros((1,), device=dev, dtype=torch.int32)
        self.coef = torch.zeros((MAX_STEPS, 4), device=dev, dtype=torch.float32)
        self.table = torch.zeros((MAX_STEPS, eng.unet_w.emb_total), device=dev, dtype=torch.float32)
        self.coef8 = torch.zeros((MAX_STEPS, 8), device=dev, dtype=torch.float32)   # DPM++ 2M rows
        self.old = torch.zeros((b, h * w, 4), device=dev, dtype=torch.float32)        # its previous x0 prediction
        self.init = torch.zeros((b, h * w, 4), device=dev, dtype=torch.float32)     # inpainting: clean init latents
        self.latmask = torch.ones((h * w,), device=dev, dtype=torch.float32)        # ... and the latent mask (1 = repaint)
        self.noise = None
        self.graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        self.graph_launches: Dict[str, int] = {}

    # one sampler step = select this step's biases, UNet on [cond | uncond], CFG + update + repack
    def step_ddim(self, cfg_scale: float):
        ops.select_step(self.table, self.step, self.unet.cur_bias)
        self.unet.run()
        ops.cfg_ddim_step(self.unet.eps, self.x, self.unet.xin, cfg_scale, self.coef, self.step)

    def step_ddim_masked(self, cfg_scale: float):
        """inpainting (sdwui CFGDenoiserTimesteps, mask_before_denoising): the kept region of x is replaced by the clean
        init latent before every model call"""
        ops.blend_latent(self.x, self.init, self.latmask)
        ops.pack_unet_input(self.x, self.unet.xin, 1.0)
        self.step_ddim(cfg_scale)

    def step_euler_a(self, cfg_scale: float):
        ops.select_step(self.table, self.step, self.unet.cur_bias)
        self.unet.run()
        ops.cfg_euler_a_step(self.unet.eps, self.x, self.noise, self.unet.xin, cfg_scale, self.coef, self.step)

    def step_euler(self, cfg_scale: float):  # sigma_up == 0 in every coefficient row: the kernel needs no noise
        ops.select_step(self.table, self.step, self.unet.cur_bias)
        self.unet.run()
        ops.cfg_euler_a_step(self.unet.eps, self.x, None, self.unet.xin, cfg_scale, self.coef, self.step)

    def step_dpmpp_2m(self, cfg_scale: float):
        ops.select_step(self.table, self.step, self.unet.cur_bias)
        self.unet.run()
        ops.cfg_dpmpp_2m_step(self.unet.eps, self.x, self.old, self.unet.xin, cfg_scale, self.coef8, self.step)


class SDEngine:
    _require_cuda = True  # tests/test_programs_cpu.py flips this together with an emulated ops module

    def _ctx(self):
        return torch.cuda.device(self.device) if self.device.type == "cuda" else contextlib.nullcontext()

    def __init__(self, sd: Dict[str, torch.Tensor], unet_cfg: UNetConfig, vae_cfg: VAEConfig, clip_cfg: CLIPConfig,
                 device="cuda:0", dtype=torch.float16, use_graphs: bool = True, vae_chunk: int = 8):
        self.device = torch.device(device)
        if self.device.type != "cuda" and self._require_cuda:
            raise RuntimeError("SDEngine needs a CUDA device: the hot path is sm_100a kernels only (no CPU fallback)")
        self.dtype = dtype
        self.unet_cfg, self.vae_cfg, self.clip_cfg = unet_cfg, vae_cfg, clip_cfg
        self.use_graphs = use_graphs
        self.vae_chunk = vae_chunk
        with self._ctx():
            self.unet_w = UNetWeights(sd, unet_cfg, self.device, dtype)
            self.vae_w = VAEDecoderWeights(sd, vae_cfg, self.device, dtype)
            self.vae_enc_w = VAEEncoderWeights(sd, vae_cfg, self.device, dtype)
            self.clip = ClipText(sd, clip_cfg, self.device, dtype)
            self.temb = TimeEmbedding(self.unet_w)
        self.plans: Dict[Tuple[int, int, int], Plan] = {}
        self.encoders: Dict[Tuple[int, int, int], VAEEncoderProgram] = {}
        self.interrupted = False
        self.variation = (None, 0.0)   # (subseed, subseed_strength) of the request being served: sdwui variation seeds
        self._cap_stream = None
        self.last_unet_evals = 0
        self.graph_replayed_launches = 0   # b200sd kernels launched through graph replays (bench.py gpu_launches)

    def _capture_stream(self):
        """torch.cuda.graph's default capture stream is ONE process-wide stream, created on whichever device captured
        first — a second engine on another device would capture (and then run) its kernels on that other device.
        Every engine captures on a stream of its own device."""
        if self._cap_stream is None:
            self._cap_stream = torch.cuda.Stream(device=self.device)
        return self._cap_stream

    def plan(self, b: int, h: int, w: int) -> Plan:
        key = (b, h, w)
        down = 2 ** (len(self.unet_cfg.channel_mult) - 1)
        if b < 1 or h < down or w < down or h % down or w % down:
            # the UNet halves the latent len(channel_mult) - 1 times and concatenates skip tensors on the way up: upstream
            # ldm fails with a size mismatch for other sizes, here it is refused before any buffer is built
            raise ValueError(f"latent size {h}x{w} (batch {b}) is not a positive multiple of {down}: image sides must be "
                             f"multiples of {8 * down} pixels")
        if key not in self.plans:
            with self._ctx():
                self.plans[key] = Plan(self, b, h, w, self.vae_chunk)
        return self.plans[key]

    @torch.no_grad()
    def encode_prompts(self, tokens: torch.Tensor) -> torch.Tensor:
        with self._ctx():
            return self.clip(tokens)

    def _graph(self, plan: Plan, name: str, fn):
        """Run fn eagerly once (per-device kernel attribute setup must not happen under capture), then capture."""
        if not self.use_graphs or self.device.type != "cuda":
            return None
        if name not in plan.graphs:
            # eager warm-up on scratch state: save what the step mutates
            saved = (plan.x.clone(), plan.step.clone(), plan.unet.xin.clone())
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                fn()
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            # One LocalGPUWorker thread per device may be building its graphs at the same time: capture one at a time,
            # and in thread-local error mode so that another thread's allocations / launches on ITS device do not
            # invalidate this capture (the default "global" mode does).
            with _CAPTURE_LOCK:
                l0 = ops.LAUNCHES
                with torch.cuda.graph(g, stream=self._capture_stream(), capture_error_mode="thread_local"):
                    fn()
                plan.graph_launches[name] = ops.LAUNCHES - l0   # b200sd kernels inside one replay
            plan.graphs[name] = g
            plan.x.copy_(saved[0]); plan.step.copy_(saved[1]); plan.unet.xin.copy_(saved[2])
        return plan.graphs[name]

    @torch.no_grad()
    def sample(self, cond: torch.Tensor, uncond: torch.Tensor, x_T: torch.Tensor, steps: int, cfg_scale: float,
               sampler: str = "DDIM", noises: Optional[torch.Tensor] = None, schedule=None,
               scheduler: Optional[str] = None, sigmas=None, inpaint=None) -> torch.Tensor:
        """cond/uncond [b, 77, ctx] fp16 on device, x_T [b, 4, h, w] fp32 (host or device): the start latents.
        `schedule` = (timesteps, coef rows) overrides the full DDIM schedule (img2img starts part-way); `sigmas` =
        (sigma table ending in 0, model log-sigmas) does the same for the k-diffusion samplers, and x_T is then the
        ALREADY NOISED start (init + noise * sigmas[0]), not unit noise.
        Returns the final latents fp32 [b, h*w, 4] (NHWC, a view of plan state)."""
        b, _, h, w = x_T.shape
        with self._ctx():
            plan = self.plan(b, h, w)
            plan.unet.set_context(torch.cat([cond, uncond]).to(self.dtype).contiguous())
            graph_name = f"{sampler}:{cfg_scale}"
            if sampler == "DDIM":
                ts, rows = schedule if schedule is not None else ddim_plan(steps)
                scale0, in0 = 1.0, 1.0
                step_fn = lambda: plan.step_ddim(cfg_scale)  # noqa: E731
                if inpaint is not None:   # (clean init latents [b, 4, h, w], latent mask [h * w])
                    plan.init.copy_(inpaint[0].to(self.device, torch.float32).permute(0, 2, 3, 1).reshape(b, h * w, 4))
                    plan.latmask.copy_(inpaint[1].to(self.device, torch.float32).reshape(-1))
                    step_fn = lambda: plan.step_ddim_masked(cfg_scale)  # noqa: E731
                    graph_name += ":mask"
            elif sampler == "Euler a":
                ts, rows, sigma0 = euler_a_plan(steps, resolve_sampler(sampler, scheduler)[1], sigmas)
                scale0, in0 = (sigma0 if sigmas is None else 1.0), 1.0 / math.sqrt(sigma0 * sigma0 + 1.0)
                step_fn = lambda: plan.step_euler_a(cfg_scale)  # noqa: E731
                if noises is None:
                    raise ValueError("Euler a needs the per-image ancestral noises")
                plan.noise = noises.to(self.device, torch.float32).permute(0, 1, 3, 4, 2).reshape(len(rows), b, h * w, 4).contiguous()
            elif sampler == "Euler":
                ts, rows, sigma0 = euler_plan(steps, resolve_sampler(sampler, scheduler)[1], sigmas)
                scale0, in0 = (sigma0 if sigmas is None else 1.0), 1.0 / math.sqrt(sigma0 * sigma0 + 1.0)
                step_fn = lambda: plan.step_euler(cfg_scale)  # noqa: E731
            elif sampler in SAMPLERS and SAMPLERS[sampler][0] == "dpmpp_2m":
                ts, rows, sigma0 = dpmpp_2m_plan(steps, resolve_sampler(sampler, scheduler)[1], sigmas)
                scale0, in0 = (sigma0 if sigmas is None else 1.0), 1.0 / math.sqrt(sigma0 * sigma0 + 1.0)
                step_fn = lambda: plan.step_dpmpp_2m(cfg_scale)  # noqa: E731
            else:
                raise ValueError(f"sampler {sampler!r} is not implemented on the local executor")
            n_evals = len(ts)
            if n_evals > MAX_STEPS:
                raise ValueError("too many steps")
            plan.table[:n_evals].copy_(self.temb.table(torch.tensor(ts, dtype=torch.float32)))
            (plan.coef8 if len(rows[0]) == 8 else plan.coef)[:n_evals].copy_(torch.tensor(rows, dtype=torch.float32))
            plan.x.copy_((x_T.to(self.device, torch.float32) * scale0).permute(0, 2, 3, 1).reshape(b, h * w, 4))
            plan.step.zero_()
            ops.pack_unet_input(plan.x, plan.unet.xin, in0)
            if inpaint is not None and sampler != "DDIM":
                raise ValueError("inpainting masks are implemented for the DDIM sampler only")
            g = self._graph(plan, graph_name, step_fn)
            self.last_unet_evals = 0
            for _ in range(n_evals):
                if self.interrupted:
                    break
                if g is not None:
                    g.replay()
                    self.graph_replayed_launches += plan.graph_launches[graph_name]
                else:
                    step_fn()
                self.last_unet_evals += 1
            if inpaint is not None:   # processing.py sample(): samples * nmask + init_latent * mask
                ops.blend_latent(plan.x, plan.init, plan.latmask)
            return plan.x

    @torch.no_grad()
    def decode(self, latents: torch.Tensor, h: int, w: int) -> torch.Tensor:
        """latents fp32 [b, h*w, 4] (scaled) -> uint8 [b, 8h, 8w, 3] on device."""
        b = latents.shape[0]
        with self._ctx():
            plan = self.plan(b, h, w)
            vae = plan.vae
            out = torch.empty((b, vae.out_h * vae.out_w, 3), device=self.device, dtype=torch.uint8)
            c = plan.vae_chunk
            for i in range(0, b, c):
                chunk = latents[i:i + c]
                if chunk.shape[0] < c:  # ragged tail: pad with the last latent, drop the surplus images
                    chunk = torch.cat([chunk, chunk[-1:].expand(c - chunk.shape[0], -1, -1)]).contiguous()
                vae.set_latents(chunk.contiguous(), self.vae_cfg.scale_factor)
                if self.use_graphs and self.device.type == "cuda":
                    if "vae" not in plan.graphs:
                        vae.run()
                        torch.cuda.current_stream().synchronize()
                        g = torch.cuda.CUDAGraph()
                        with _CAPTURE_LOCK:
                            l0 = ops.LAUNCHES
                            with torch.cuda.graph(g, stream=self._capture_stream(), capture_error_mode="thread_local"):
                                vae.run()
                            plan.graph_launches["vae"] = ops.LAUNCHES - l0
                        plan.graphs["vae"] = g
                    plan.graphs["vae"].replay()
                    self.graph_replayed_launches += plan.graph_launches["vae"]
                else:
                    vae.run()
                n = min(c, b - i)
                out[i:i + n].copy_(vae.u8[:n])
            return out.reshape(b, vae.out_h, vae.out_w, 3)

    @torch.no_grad()
    def encode(self, images_u8: torch.Tensor) -> torch.Tensor:
        """images uint8 [b, H, W, 3] (host or device) -> scaled latents fp32 [b, 4, H/f, W/f] (posterior mean), in
        chunks of `vae_chunk` images."""
        b, hh, ww, _ = images_u8.shape
        with self._ctx():
            c = min(self.vae_chunk, b)
            key = (c, hh, ww)
            if key not in self.encoders:
                self.encoders[key] = VAEEncoderProgram(self.vae_enc_w, c, hh, ww)
            enc = self.encoders[key]
            imgs = images_u8.to(self.device).reshape(b, hh * ww, 3)
            out = torch.empty((b, enc.lat_h * enc.lat_w, 4), device=self.device, dtype=torch.float32)
            for i in range(0, b, c):
                chunk = imgs[i:i + c]
                n = chunk.shape[0]
                enc.img_u8[:n].copy_(chunk)
                if n < c:
                    enc.img_u8[n:].copy_(chunk[-1:].expand(c - n, -1, -1))
                enc.run()
                out[i:i + n].copy_(enc.latents[:n])
            return out.reshape(b, enc.lat_h, enc.lat_w, 4).permute(0, 3, 1, 2).contiguous()

    @torch.no_grad()
    def img2img(self, tokens: torch.Tensor, neg_tokens: torch.Tensor, seed: int, init_u8: torch.Tensor,
                denoising_strength: float = 0.75, steps: int = 20, cfg_scale: float = 7.0, sampler: str = "DDIM",
                scheduler: Optional[str] = None, latmask: Optional[torch.Tensor] = None,
                inpainting_fill: int = 1) -> torch.Tensor:
        """img2img: VAE-encode the init images (posterior mean), noise them to t_enc, run the remaining part of the
        sampler's schedule, decode.  init_u8 uint8 [b, H, W, 3].  Returns uint8 [b, H, W, 3] on device.
        `latmask` fp32 [h * w] (b200sd.inpaint.prepare_mask): inpainting — the region with latmask 0 keeps the init
        latents at every step (DDIM only); the caller composites the original pixels back (inpaint.apply_overlays).
        `inpainting_fill` 2 ("latent noise") / 3 ("latent nothing") replace the repainted region of the init latents by
        the request's start noise / by zeros first (sdwui Img2Img.init); 0 ("fill") is image-space work the caller does
        before the call (inpaint.fill_masked), 1 keeps the original content."""
        b = tokens.shape[0]
        cond = self.encode_prompts(tokens)
        uncond = self.encode_prompts(neg_tokens)
        init = self.encode(init_u8)
        _, _, h, w = init.shape
        if latmask is not None and inpainting_fill in (2, 3):
            nm = latmask.to(self.device, torch.float32).reshape(1, 1, h, w)
            init = init * (1.0 - nm)
            if inpainting_fill == 2:   # create_random_tensors(shape, seeds): the same first draw the sampler starts from
                init = init + per_image_noise(seed, b, (4, h, w), 1, *self.variation)[0].to(self.device) * nm
        lat = self._sample_from(init, cond, uncond, seed, denoising_strength, steps, cfg_scale, sampler, scheduler,
                                inpaint=None if latmask is None else (init, latmask))
        return self.decode(lat, h, w)

    def _sample_from(self, init: torch.Tensor, cond, uncond, seed: int, denoising_strength: float, steps: int,
                     cfg_scale: float, sampler: str, scheduler: Optional[str], inpaint=None) -> torch.Tensor:
        """the img2img half of a sampler (also the second pass of the hires fix): `init` [b, 4, h, w] latents on the device,
        fresh per-image noise from `seed`, start at the noise level of t_enc.
        DDIM: sdwui sd_samplers_timesteps.sample_img2img; k-diffusion samplers: KDiffusionSampler.sample_img2img."""
        b, _, h, w = init.shape
        method, sched = resolve_sampler(sampler, scheduler)
        if method == "ddim":
            noise = per_image_noise(seed, b, (4, h, w), 1, *self.variation)[0].to(self.device)
            sa, s1a, ts, rows = ddim_img2img_plan(steps, denoising_strength)
            return self.sample(cond, uncond, init * sa + noise * s1a, steps, cfg_scale, "DDIM", schedule=(ts, rows),
                               inpaint=inpaint)
        if inpaint is not None:
            raise ValueError("inpainting masks are implemented for the DDIM sampler only")
        sig, log_sig = kdiffusion_img2img_sigmas(steps, denoising_strength, sched)
        n_evals = len(sig) - 1
        draws = 1 + (n_evals if method == "euler_a" else 0)
        nz = per_image_noise(seed, b, (4, h, w), draws, *self.variation)
        x0 = init + nz[0].to(self.device) * float(sig[0])
        return self.sample(cond, uncond, x0, n_evals, cfg_scale, sampler, noises=nz[1:] if draws > 1 else None,
                           scheduler=scheduler, sigmas=(sig, log_sig))

    @torch.no_grad()
    def txt2img_hires(self, tokens: torch.Tensor, neg_tokens: torch.Tensor, seed: int, steps: int = 20,
                      cfg_scale: float = 7.0, height: int = 512, width: int = 512, hr_scale: float = 2.0,
                      hr_steps: int = 0, denoising_strength: float = 0.7, sampler: str = "DDIM",
                      scheduler: Optional[str] = None) -> torch.Tensor:
        """txt2img with sdwui's hires fix and the "Latent" upscaler (StableDiffusionProcessingTxt2Img.sample /
        sample_hr_pass): first pass at (height, width), bilinear resize of the latents to hr_scale x, a fresh per-image
        noise of the large shape from the same seeds, then the same sampler's img2img half from t_enc with `hr_steps`
        (0 = `steps`) steps, decode at the large size.  Returns uint8 [b, H*hr, W*hr, 3] on device."""
        b = tokens.shape[0]
        h, w = height // 8, width // 8
        h2, w2 = int(height * hr_scale) // 8, int(width * hr_scale) // 8
        cond = self.encode_prompts(tokens)
        uncond = self.encode_prompts(neg_tokens)
        draws = 1 + (steps if sampler == "Euler a" else 0)
        nz = per_image_noise(seed, b, (4, h, w), draws, *self.variation)
        lat = self.sample(cond, uncond, nz[0], steps, cfg_scale, sampler, noises=nz[1:] if draws > 1 else None,
                          scheduler=scheduler)
        with self._ctx():
            up = torch.empty((b, h2 * w2, 4), device=self.device, dtype=torch.float32)
            ops.resize_latent_bilinear(lat.contiguous(), up, h, w, h2, w2)
        init = up.reshape(b, h2, w2, 4).permute(0, 3, 1, 2)
        lat2 = self._sample_from(init, cond, uncond, seed, denoising_strength, hr_steps or steps, cfg_scale, sampler, scheduler)
        return self.decode(lat2, h2, w2)

    @torch.no_grad()
    def txt2img(self, tokens: torch.Tensor, neg_tokens: torch.Tensor, seed: int, steps: int = 20, cfg_scale: float = 7.0,
                height: int = 512, width: int = 512, sampler: str = "DDIM", scheduler: Optional[str] = None) -> torch.Tensor:
        """Whole request for this engine's share: returns uint8 [b, H, W, 3] on device."""
        b = tokens.shape[0]
        h, w = height // 8, width // 8
        cond = self.encode_prompts(tokens)
        uncond = self.encode_prompts(neg_tokens)
        draws = 1 + (steps if sampler == "Euler a" else 0)
        nz = per_image_noise(seed, b, (4, h, w), draws, *self.variation)
        lat = self.sample(cond, uncond, nz[0], steps, cfg_scale, sampler, noises=nz[1:] if draws > 1 else None,
                          scheduler=scheduler)
        return self.decode(lat, h, w)
