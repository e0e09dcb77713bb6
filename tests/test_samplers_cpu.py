"""Every sampler of the reference's ETA table (scripts/spartan/worker.py:75-94) on the executor's generic stage machine
(b200sd/samplers.py: coefficient algebra, b200sd_cfg_eps + b200sd_latent_lincomb) against the oracle's step-by-step
restatement of k-diffusion / sdwui (oracle/sd_oracle.py run_sampler), with b200sd.ops emulated in fp32 on the CPU:
agreement to rounding — any mismatch is an algebra or plumbing bug."""
import math

import pytest
import torch

import ops_emulator

ALL = ["Euler", "Euler a", "LMS", "Heun", "DPM2", "DPM2 a", "DPM++ 2S a", "DPM++ 2M", "DPM++ SDE", "DPM fast", "DPM adaptive",
       "LMS Karras", "DPM2 Karras", "DPM2 a Karras", "DPM++ 2S a Karras", "DPM++ 2M Karras", "DPM++ SDE Karras", "PLMS"]


@pytest.fixture()
def env(monkeypatch):
    from b200sd import config as C, engine as E, ops, synth
    from oracle import sd_oracle as O
    ops_emulator.install(monkeypatch, ops)
    monkeypatch.setattr(E.SDEngine, "_require_cuda", False)
    cfgs = (C.TINY_UNET, C.TINY_VAE, C.TINY_CLIP)
    sd = synth.make_state_dict(*cfgs, seed=0)
    eng = E.SDEngine(sd, *cfgs, device="cpu", dtype=torch.float32, use_graphs=False, vae_chunk=2)
    b = 2
    tok, neg = O.random_prompt_tokens(b, vocab_hi=997), O.empty_prompt_tokens(b, vocab_hi=997)
    cond, unc = O.clip_text_encode(sd, cfgs[2], tok), O.clip_text_encode(sd, cfgs[2], neg)
    unet = lambda x, t, c: O.unet_forward(sd, cfgs[0], x, t, c)  # noqa: E731
    return E, O, eng, cond, unc, unet, b


def test_reference_table_is_covered():
    """no sampler name the reference's ETA model knows falls back any more"""
    import os, re
    from b200sd import engine as E
    src = open(os.path.join(os.path.dirname(__file__), "..", "stable-diffusion-webui-distributed_b200", "scripts", "spartan",
                            "worker.py")).read()
    table = re.search(r"other_to_euler_a\s*=\s*\{(.*?)\}", src, re.S).group(1)
    names = re.findall(r'"([^"]+)"\s*:', table)
    assert len(names) == 18
    for n in names + ["Euler a"]:
        assert n in E.SAMPLERS, n
    assert set(ALL) | {"DDIM"} >= set(names)


@pytest.mark.parametrize("name", ALL)
def test_txt2img_sampler_matches_oracle(env, name):
    E, O, eng, cond, unc, unet, b = env
    hw, steps = 8, 7
    pr = eng.program(name, None, steps)
    nz = E.per_image_noise(4100, b, (4, hw, hw), 1 + pr.draws)
    with torch.no_grad():
        ref = O.run_sampler(name, unet, cond, unc, 7.0, steps, nz[0], list(nz[1:]))
    lat = eng.run_program(cond, unc, pr.start(nz[0]), pr, 7.0, noises=nz[1:] if pr.draws else None)
    z = lat.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    assert float((z - ref).abs().max()) <= 2e-3 * float(ref.abs().max()), name
    if pr.n_evals is not None:
        assert eng.last_unet_evals == pr.n_evals
    two = {"Heun", "DPM2", "DPM2 a", "DPM++ 2S a", "DPM++ SDE"}
    base = name.replace(" Karras", "")
    if base in two:
        assert eng.last_unet_evals == 2 * steps - 1      # the step to sigma 0 is a single Euler evaluation
    elif base in ("DPM fast", "LMS", "Euler", "Euler a", "DPM++ 2M"):
        assert eng.last_unet_evals == steps
    elif base == "PLMS":
        # len(timesteps) - 1 iterations, the first with two evaluations (7 steps give 8 timesteps: 1000 // 7 = 142)
        assert eng.last_unet_evals == len(O.ddim_timesteps(steps))
    else:
        assert base == "DPM adaptive" and eng.last_unet_evals % 3 == 0 and eng.last_unet_evals >= 3


@pytest.mark.parametrize("name", ["Heun", "DPM2 a Karras", "DPM++ SDE", "LMS", "DPM fast", "PLMS", "DPM++ 2S a", "DPM adaptive"])
def test_img2img_half_matches_oracle(env, name):
    """the tail of every sampler from a noised init (hires second pass and img2img)"""
    E, O, eng, cond, unc, unet, b = env
    hw, steps, d = 8, 9, 0.6
    g = torch.Generator().manual_seed(5)
    init = torch.randn((b, 4, hw, hw), generator=g) * 0.7
    pr = eng.program(name, None, steps, denoise=d)
    nz = E.per_image_noise(4200, b, (4, hw, hw), 1 + pr.draws)
    with torch.no_grad():
        ref = O.run_sampler(name, unet, cond, unc, 7.0, steps, nz[0], list(nz[1:]), init=init, denoising_strength=d)
    lat = eng.run_program(cond, unc, pr.start(nz[0], init), pr, 7.0, noises=nz[1:] if pr.draws else None)
    z = lat.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    assert float((z - ref).abs().max()) <= 2e-3 * float(ref.abs().max()), name


@pytest.mark.parametrize("name", ["Euler a", "DPM++ 2M Karras", "Euler", "Heun", "DPM++ 2S a Karras", "PLMS", "LMS", "DPM++ SDE"])
def test_masked_sampling_matches_oracle(env, name):
    """inpainting on the k-diffusion samplers (sdwui CFGDenoiser: the DENOISED prediction is blended with the init latents)
    and on PLMS (the model INPUT is blended), plus the final blend of processing.py"""
    E, O, eng, cond, unc, unet, b = env
    hw, steps, d = 8, 8, 0.75
    g = torch.Generator().manual_seed(6)
    init = torch.randn((b, 4, hw, hw), generator=g) * 0.7
    nmask = (torch.rand((hw, hw), generator=g) > 0.5).float()
    pr = eng.program(name, None, steps, denoise=d, masked=True)
    nz = E.per_image_noise(4300, b, (4, hw, hw), 1 + pr.draws)
    with torch.no_grad():
        ref = O.run_sampler(name, unet, cond, unc, 7.0, steps, nz[0], list(nz[1:]), init=init, denoising_strength=d,
                            mask=(init, nmask[None, None]))
        ref = ref * nmask[None, None] + init * (1 - nmask[None, None])
    lat = eng.run_program(cond, unc, pr.start(nz[0], init), pr, 7.0, noises=nz[1:] if pr.draws else None,
                          inpaint=(init, nmask.reshape(-1)))
    z = lat.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    assert float((z - ref).abs().max()) <= 2e-3 * float(ref.abs().max()), name
    keep = (nmask == 0)[None, None].expand_as(z)
    assert torch.equal(z[keep], init[keep])


def test_generic_twins_equal_the_fused_kernels(env):
    """Euler / Euler a / DPM++ 2M exist twice (fused per-step kernel; stage list used when a mask rides along): same result"""
    E, O, eng, cond, unc, unet, b = env
    hw, steps = 8, 6
    for name in ("Euler", "Euler a", "DPM++ 2M"):
        fused, twin = eng.program(name, None, steps), eng.program(name, None, steps, masked=True)
        assert fused.fused is not None and twin.fused is None and twin.sp is not None
        nz = E.per_image_noise(4400, b, (4, hw, hw), 1 + fused.draws)
        a = eng.run_program(cond, unc, fused.start(nz[0]), fused, 7.0, noises=nz[1:] if fused.draws else None).clone()
        c = eng.run_program(cond, unc, twin.start(nz[0]), twin, 7.0, noises=nz[1:] if twin.draws else None)
        assert float((a - c).abs().max()) <= 1e-4 * float(a.abs().max()), name


def test_sde_noise_pair_has_the_brownian_covariance():
    """n1, n2 of a DPM++ SDE step: unit variances and cov = sqrt(len1 / len_total), as increments of one Brownian motion"""
    from oracle import sd_oracle as O
    g = torch.Generator().manual_seed(1)
    z1, z2 = torch.randn(200000, generator=g), torch.randn(200000, generator=g)
    s, ss, sn = 3.0, 1.7, 0.9
    n1, n2 = O.brownian_pair(z1, z2, s, ss, sn)
    assert abs(float(n1.var()) - 1) < 0.02 and abs(float(n2.var()) - 1) < 0.02
    assert abs(float((n1 * n2).mean()) - math.sqrt((s - ss) / (s - sn))) < 0.02


def test_graph_warmup_must_not_shift_history():
    """engine._graph runs a stage eagerly once before capturing it: the state it saves and restores must include the
    multistep history (found on the GPU: PLMS's first multistep stage was shifted twice)"""
    import inspect
    from b200sd import engine as E
    src = inspect.getsource(E.SDEngine._graph)
    assert "plan.lat.items()" in src and "plan.old" in src


# ------------------------------------------------------------------------------------------------ the product vs diffusion theory
# tests/test_oracle_pins_cpu.py pins the ORACLE's samplers to the exact probability-flow solution / data distribution of
# Gaussian data.  Here the PRODUCT's sampler machinery (coefficient algebra of b200sd/samplers.py, the fused per-step kernels'
# schedules, Program.start's scaling, the stage loop of the engine) meets the same theory directly, without the oracle in
# between: the UNet program of the plan is replaced by the closed-form optimal eps-predictor, everything else runs as shipped.
_S_DATA = 0.8


def _theory_run(env, monkeypatch, name, steps, hw=16, seed=77):
    E, O, eng, cond, unc, unet, b = env
    _, log_sig = O.model_sigmas()
    seen = {}
    table = eng.temb.table

    def recording_table(ts, y=None):
        seen["ts"] = [float(v) for v in ts]
        return table(ts, y)
    monkeypatch.setattr(eng.temb, "table", recording_table)
    plan = eng.plan(b, hw, hw)

    def optimal_eps():   # eps(x_t, t) = x_t sqrt(1 - a) / (a s^2 + 1 - a), a = alphas_cumprod at (fractional) t
        t = torch.tensor(seen["ts"][int(plan.step.item())], dtype=torch.float64).clamp(0, 999)
        lo = t.floor().long().clamp(max=998)
        w = t - lo
        sg = ((1 - w) * log_sig[lo] + w * log_sig[lo + 1]).exp()
        a = 1 / (1 + sg * sg)
        k = float(torch.sqrt(1 - a) / (a * _S_DATA ** 2 + 1 - a))
        plan.unet.eps[:, :, :4] = plan.unet.xin[:, :, :4].float() * k
    monkeypatch.setattr(plan.unet, "run", optimal_eps)
    pr = eng.program(name, None, steps)
    nz = E.per_image_noise(seed, b, (4, hw, hw), 1 + pr.draws)
    lat = eng.run_program(cond, unc, pr.start(nz[0]), pr, 1.0, noises=nz[1:] if pr.draws else None)
    return pr, nz[0].permute(0, 2, 3, 1).reshape(b, hw * hw, 4), lat.clone()


@pytest.mark.parametrize("name,err40,order", [
    ("DDIM", 0.05, 0.75), ("Euler", 0.06, 1), ("Heun", 0.003, 2), ("DPM2", 0.012, 2), ("DPM2 Karras", 0.003, 2),
    ("DPM++ 2M", 0.006, 2), ("DPM++ 2M Karras", 0.006, 2), ("LMS", 5e-4, 3), ("LMS Karras", 2e-3, 3), ("PLMS", 0.008, 1.2)])
def test_product_deterministic_samplers_solve_the_probability_flow_ode(env, monkeypatch, name, err40, order):
    E, O = env[0], env[1]
    errs = []
    for steps in (20, 40):
        pr, noise0, lat = _theory_run(env, monkeypatch, name, steps)
        s = _S_DATA
        if name in ("DDIM", "PLMS"):    # timestep samplers stop at alphas_cumprod[timesteps[0]] (DDIM) / [0] (PLMS)
            ac, ts = O.alphas_cumprod().double(), O.ddim_timesteps(steps)
            a_T, a_0 = float(ac[ts[-1]]), float(ac[ts[0]] if name == "DDIM" else ac[0])
            exact = noise0 * math.sqrt((a_0 * s * s + 1 - a_0) / (a_T * s * s + 1 - a_T))
        else:
            s0 = pr.noise_scale          # sigma_0 of the schedule the program was built on
            exact = noise0 * s0 * s / math.sqrt(s * s + s0 * s0)
        errs.append(float((lat - exact).norm() / exact.norm()))
    assert errs[1] <= err40 and errs[0] / errs[1] >= 0.8 * 2 ** order, (name, errs)


@pytest.mark.parametrize("name,tol", [("DPM fast", 0.01), ("DPM adaptive", 0.05)])
def test_product_dpm_solver_samplers_solve_the_probability_flow_ode(env, monkeypatch, name, tol):
    pr, noise0, lat = _theory_run(env, monkeypatch, name, 20)
    s, s0 = _S_DATA, pr.noise_scale
    exact = noise0 * s0 * s / math.sqrt(s * s + s0 * s0)
    assert float((lat - exact).norm() / exact.norm()) <= tol


@pytest.mark.parametrize("name,lo,hi", [("Euler a", 0.78, 1.04), ("DPM2 a", 0.90, 1.12), ("DPM++ 2S a", 0.90, 1.08),
                                        ("DPM++ SDE", 0.88, 1.08), ("DPM++ SDE Karras", 0.90, 1.09)])
def test_product_stochastic_samplers_end_on_the_data_distribution(env, monkeypatch, name, lo, hi):
    pr, noise0, lat = _theory_run(env, monkeypatch, name, 40, hw=32)
    ratio = float(lat.var()) / _S_DATA ** 2
    assert lo <= ratio <= hi and abs(float(lat.mean())) <= 0.04, (name, ratio, float(lat.mean()))
