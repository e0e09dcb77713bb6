"""Host-side logic of the executor on CPU: the UNet / VAE programs, weight packing, per-step bias table and sampler
bookkeeping are run with b200sd.ops replaced by a torch emulation (tests/ops_emulator.py) and compared with the
fp32 oracle.  fp32 buffers, so agreement is to rounding (1e-4): any mismatch is a plumbing bug, not precision."""
import pytest
import torch

import ops_emulator


@pytest.fixture()
def env(monkeypatch):
    from b200sd import config as C, engine as E, ops, synth
    from oracle import sd_oracle as O
    ops_emulator.install(monkeypatch, ops)
    monkeypatch.setattr(E.SDEngine, "_require_cuda", False)
    cfgs = (C.TINY_UNET, C.TINY_VAE, C.TINY_CLIP)
    sd = synth.make_state_dict(*cfgs, seed=0)
    eng = E.SDEngine(sd, *cfgs, device="cpu", dtype=torch.float32, use_graphs=False, vae_chunk=2)
    return C, E, O, cfgs, sd, eng


def test_layouts_match_oracle(env):
    C, E, O, cfgs, sd, eng = env
    assert C.unet_layout(C.SD15_UNET) == O.unet_layout(O.SD15_UNET)
    assert C.unet_layout(C.TINY_UNET) == O.unet_layout(O.TINY_UNET)
    assert sum(v.numel() for k, v in sd.items() if k.startswith(C.UNET_PREFIX)) > 0


def test_schedules_match_oracle(env):
    C, E, O, cfgs, sd, eng = env
    for steps in (5, 20, 50):
        ts, rows = E.ddim_plan(steps)
        ref = O.ddim_coefficients(steps)
        assert len(ts) == steps - 1 == len(ref)
        for t, r, o in zip(ts, rows, ref):
            assert t == o[0] and all(abs(a - b) < 1e-12 for a, b in zip(r, o[1:]))
        ts, rows, s0 = E.euler_a_plan(steps)
        ref = O.euler_a_coefficients(steps)
        for t, r, o in zip(ts, rows, ref):
            assert abs(t - o[0]) < 1e-9 and abs(r[0] - o[1]) < 1e-12 and abs(r[1] - o[2]) < 1e-12
            assert abs(r[2] - o[3]) < 1e-12 and abs(r[3] - o[5]) < 1e-12
    assert torch.equal(E.per_image_noise(5, 3, (4, 8, 8))[0], O.per_image_noise(5, 3, (4, 8, 8)))


@pytest.mark.parametrize("b,hw", [(1, 8), (2, 16)])
def test_unet_program_matches_oracle(env, b, hw):
    C, E, O, cfgs, sd, eng = env
    from b200sd import ops
    tok = O.random_prompt_tokens(b, vocab_hi=997)
    neg = O.empty_prompt_tokens(b, vocab_hi=997)
    cond, unc = O.clip_text_encode(sd, cfgs[2], tok), O.clip_text_encode(sd, cfgs[2], neg)
    x = O.per_image_noise(1000, b, (4, hw, hw))
    with torch.no_grad():
        ref = O.unet_forward(sd, cfgs[0], torch.cat([x, x]), torch.full((2 * b,), 651.0), torch.cat([cond, unc]))
    plan = eng.plan(b, hw, hw)
    plan.unet.set_context(torch.cat([cond, unc]))
    plan.table[:1].copy_(eng.temb.table(torch.tensor([651.0])))
    plan.x.copy_(x.permute(0, 2, 3, 1).reshape(b, hw * hw, 4))
    ops.pack_unet_input(plan.x, plan.unet.xin, 1.0)
    ops.select_step(plan.table, plan.step, plan.unet.cur_bias)
    plan.unet.run()
    got = plan.unet.eps[..., :4].reshape(2 * b, hw, hw, 4).permute(0, 3, 1, 2)
    assert float(plan.unet.eps[..., 4:].abs().max()) == 0.0
    assert float((got - ref).abs().max()) <= 2e-4 * float(ref.abs().max())


def test_clip_text_matches_oracle(env):
    C, E, O, cfgs, sd, eng = env
    tok = O.random_prompt_tokens(3, vocab_hi=997)
    assert torch.allclose(eng.encode_prompts(tok), O.clip_text_encode(sd, cfgs[2], tok), atol=1e-4, rtol=1e-4)


def test_txt2img_matches_oracle_ddim(env):
    C, E, O, cfgs, sd, eng = env
    b, hw, steps = 3, 8, 5  # b=3 with vae_chunk=2 exercises the ragged decode tail
    tok = O.random_prompt_tokens(b, vocab_hi=997)
    neg = O.empty_prompt_tokens(b, vocab_hi=997)
    with torch.no_grad():
        ref_u8, ref_z, ref_dec = O.txt2img(sd, *cfgs, tok, neg, seed=1000, steps=steps, height=hw * 8, width=hw * 8)
    got = eng.txt2img(tok, neg, seed=1000, steps=steps, cfg_scale=7.0, height=hw * 8, width=hw * 8, sampler="DDIM")
    assert eng.last_unet_evals == steps - 1
    z = eng.plan(b, hw, hw).x.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    assert float((z - ref_z).abs().max()) <= 1e-3 * float(ref_z.abs().max())
    assert got.shape == ref_u8.shape
    d = (got.int() - ref_u8.int()).abs()
    assert float((d <= 1).float().mean()) == 1.0 and float((d == 0).float().mean()) > 0.99


def test_sample_matches_oracle_euler_a(env):
    C, E, O, cfgs, sd, eng = env
    b, hw, steps = 2, 8, 4
    tok = O.random_prompt_tokens(b, vocab_hi=997)
    neg = O.empty_prompt_tokens(b, vocab_hi=997)
    cond, unc = O.clip_text_encode(sd, cfgs[2], tok), O.clip_text_encode(sd, cfgs[2], neg)
    nz = E.per_image_noise(2000, b, (4, hw, hw), 1 + steps)
    with torch.no_grad():
        ref = O.sample_euler_a(lambda x, t, c: O.unet_forward(sd, cfgs[0], x, t, c), nz[0], cond, unc, steps, 7.0,
                               list(nz[1:]))
    lat = eng.sample(cond, unc, nz[0], steps, 7.0, "Euler a", noises=nz[1:])
    z = lat.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    assert float((z - ref).abs().max()) <= 1e-3 * float(ref.abs().max())


def test_sample_matches_oracle_euler(env):
    """sdwui "Euler" = k-diffusion sample_euler with s_churn 0: the ancestral kernel with sigma_up = 0 rows, no noise"""
    C, E, O, cfgs, sd, eng = env
    b, hw, steps = 2, 8, 5
    ts, rows, s0 = E.euler_plan(steps)
    ts_a, rows_a, s0_a = E.euler_a_plan(steps)
    assert ts == ts_a and s0 == s0_a and all(r[2] == 0.0 for r in rows)
    assert all(abs(r[1] ** 2 - (ra[1] ** 2 + ra[2] ** 2)) < 1e-9 for r, ra in zip(rows, rows_a))
    tok = O.random_prompt_tokens(b, vocab_hi=997)
    neg = O.empty_prompt_tokens(b, vocab_hi=997)
    cond, unc = O.clip_text_encode(sd, cfgs[2], tok), O.clip_text_encode(sd, cfgs[2], neg)
    nz = E.per_image_noise(3000, b, (4, hw, hw), 1)
    with torch.no_grad():
        ref = O.sample_euler(lambda x, t, c: O.unet_forward(sd, cfgs[0], x, t, c), nz[0], cond, unc, steps, 7.0)
    lat = eng.sample(cond, unc, nz[0], steps, 7.0, "Euler")
    z = lat.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    assert float((z - ref).abs().max()) <= 1e-3 * float(ref.abs().max())


@pytest.mark.parametrize("sched,karras", [(None, True), ("Karras", True), ("Uniform", False)])
def test_sample_matches_oracle_dpmpp_2m(env, sched, karras):
    """sdwui "DPM++ 2M": k-diffusion sample_dpmpp_2m, Karras schedule by default (sdwui >= 1.9) or as the API asks"""
    C, E, O, cfgs, sd, eng = env
    b, hw, steps = 2, 8, 6
    ts, rows, s0 = E.dpmpp_2m_plan(steps, "karras" if karras else "uniform")
    sig, log_sig = O.sigmas_karras(steps) if karras else O.karras_sigmas_compvis(steps)
    assert abs(s0 - float(sig[0])) < 1e-12 and len(rows) == steps and rows[0][3] == 0.0 and rows[-1][3] == 0.0
    assert all(abs(r[0] - float(sig[i])) < 1e-12 and abs(r[1] * r[0] - float(sig[i + 1])) < 1e-12 for i, r in enumerate(rows))
    assert all(abs(r[2] - r[3] - 1.0) < 1e-12 for r in rows)      # c1 - c2 == 1: a constant x0 prediction is a fixed point
    assert all(abs(t - O.sigma_to_t(float(sig[i]), log_sig)) < 1e-9 for i, t in enumerate(ts))
    tok = O.random_prompt_tokens(b, vocab_hi=997)
    neg = O.empty_prompt_tokens(b, vocab_hi=997)
    cond, unc = O.clip_text_encode(sd, cfgs[2], tok), O.clip_text_encode(sd, cfgs[2], neg)
    nz = E.per_image_noise(3100, b, (4, hw, hw), 1)
    with torch.no_grad():
        ref = O.sample_dpmpp_2m(lambda x, t, c: O.unet_forward(sd, cfgs[0], x, t, c), nz[0], cond, unc, steps, 7.0,
                                karras=karras)
    lat = eng.sample(cond, unc, nz[0], steps, 7.0, "DPM++ 2M", scheduler=sched)
    assert eng.last_unet_evals == steps
    z = lat.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    assert float((z - ref).abs().max()) <= 1e-3 * float(ref.abs().max())


def test_sizes_the_unet_cannot_take_are_refused(env):
    """upstream ldm dies with a tensor size mismatch when the latent is not a multiple of 2^(levels-1); here the request
    is refused before any buffer exists (the worker turns it into InvalidWorkerResponse, like a remote 500)"""
    C, E, O, cfgs, sd, eng = env
    down = 2 ** (len(cfgs[0].channel_mult) - 1)
    with pytest.raises(ValueError):
        eng.plan(1, down + 1, down)
    with pytest.raises(ValueError):
        eng.plan(0, down, down)
    assert eng.plan(1, down, 2 * down).b == 1


@pytest.mark.parametrize("label,fn", [("Exponential", "sigmas_exponential"), ("Polyexponential", "sigmas_exponential"),
                                      ("SGM Uniform", "sigmas_sgm_uniform"), ("Karras", "sigmas_karras")])
def test_euler_on_the_api_schedulers(env, label, fn):
    """sdwui >= 1.9 sends the noise schedule separately: the k-diffusion samplers run on any of the implemented ones"""
    C, E, O, cfgs, sd, eng = env
    b, hw, steps = 2, 8, 5
    sig, log_sig = getattr(O, fn)(steps)
    mine, _ = E.kdiffusion_sigmas(steps, E.SCHEDULERS[label])
    assert torch.allclose(mine, sig, rtol=1e-12, atol=1e-12)
    assert all(float(sig[i]) > float(sig[i + 1]) for i in range(steps))
    tok = O.random_prompt_tokens(b, vocab_hi=997)
    neg = O.empty_prompt_tokens(b, vocab_hi=997)
    cond, unc = O.clip_text_encode(sd, cfgs[2], tok), O.clip_text_encode(sd, cfgs[2], neg)
    nz = E.per_image_noise(3200, b, (4, hw, hw), 1)
    with torch.no_grad():
        ref = O.sample_euler_sigmas(lambda x, t, c: O.unet_forward(sd, cfgs[0], x, t, c), nz[0], cond, unc, sig, log_sig, 7.0)
    lat = eng.sample(cond, unc, nz[0], steps, 7.0, "Euler", scheduler=label)
    z = lat.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    assert float((z - ref).abs().max()) <= 1e-3 * float(ref.abs().max())


def test_variation_seeds(env):
    """subseed / subseed_strength (sdwui ImageRNG.first): strength 0 leaves the noise alone, 1 gives the subseed's noise,
    in between the spherical interpolation of the oracle; later draws (ancestral noise) are untouched"""
    C, E, O, cfgs, sd, eng = env
    shape = (4, 8, 8)
    base = E.per_image_noise(10, 3, shape, 2)
    assert torch.equal(E.per_image_noise(10, 3, shape, 2, 500, 0.0), base)
    full = E.per_image_noise(10, 3, shape, 2, 500, 1.0)
    # sdwui processing.py: with variation seeds EVERY image uses the base seed (all_seeds[k] = seed) and only the subseed
    # advances (all_subseeds[k] = subseed + k): the later draws of all images are image 0's
    assert torch.allclose(full[0], E.per_image_noise(500, 3, shape)[0], atol=1e-5)
    assert all(torch.equal(full[1, k], base[1, 0]) for k in range(3))
    mid = E.per_image_noise(10, 3, shape, 2, 500, 0.3)
    assert torch.allclose(mid[0], O.per_image_noise(10, 3, shape, subseed=500, subseed_strength=0.3), atol=1e-6)
    assert not torch.allclose(mid[0], base[0]) and all(torch.equal(mid[1, k], base[1, 0]) for k in range(3))
    s0 = E.slerp(0.3, base[0, 0], E.per_image_noise(501, 1, shape)[0, 0])   # image 1: noise(seed) with subnoise(subseed + 1)
    assert torch.allclose(mid[0, 1], s0, atol=1e-6)
    # through the request path: the engine's variation attribute changes the start noise only when set
    tok, neg = O.random_prompt_tokens(2, vocab_hi=997), O.empty_prompt_tokens(2, vocab_hi=997)
    a = eng.txt2img(tok, neg, seed=5, steps=3, cfg_scale=7.0, height=64, width=64, sampler="DDIM").clone()
    eng.variation = (900, 0.5)
    b = eng.txt2img(tok, neg, seed=5, steps=3, cfg_scale=7.0, height=64, width=64, sampler="DDIM").clone()
    eng.variation = (None, 0.0)
    c = eng.txt2img(tok, neg, seed=5, steps=3, cfg_scale=7.0, height=64, width=64, sampler="DDIM")
    assert torch.equal(a, c) and not torch.equal(a, b)


def test_prompts_use_a_local_clip_tokenizer_when_given(env, tmp_path, monkeypatch):
    """SD_TOKENIZER=<dir with vocab.json + merges.txt> switches the worker from hashed word ids to real BPE ids"""
    pytest.importorskip("transformers")
    import json
    from b200sd import factory
    vocab = {"<|startoftext|>": 0, "<|endoftext|>": 1}
    for i, ch in enumerate("abcdefghijklmnopqrstuvwxyz"):
        vocab[ch], vocab[ch + "</w>"] = 2 + 2 * i, 3 + 2 * i
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    (tmp_path / "merges.txt").write_text("#version: 0.2\n")
    hashed = factory.synthetic_tokens(["ab c", ""], 997)
    assert hashed.shape == (2, 77) and int(hashed[0, 0]) == 995 and int(hashed[1, 1]) == 996
    monkeypatch.setenv("SD_TOKENIZER", str(tmp_path))
    ids = factory.synthetic_tokens(["ab c", ""], 997)
    assert ids.shape == (2, 77) and ids[0, :6].tolist() == [0, 2, 5, 7, 1, 1] and ids[1, :3].tolist() == [0, 1, 1]
    with pytest.raises(ValueError):
        factory.synthetic_tokens(["z"], 40)


def test_sampler_names_resolve(env):
    C, E, O, cfgs, sd, eng = env
    assert E.resolve_sampler("DPM++ 2M") == ("dpmpp_2m", "karras") == E.resolve_sampler("DPM++ 2M Karras")
    assert E.resolve_sampler("DPM++ 2M", "Automatic") == ("dpmpp_2m", "karras")
    assert E.resolve_sampler("DPM++ 2M", "Uniform") == ("dpmpp_2m", "uniform")
    assert E.resolve_sampler("Euler a") == ("euler_a", "uniform") and E.resolve_sampler("DDIM") == ("ddim", None)
    with pytest.raises(ValueError):
        E.resolve_sampler("UniPC")
    assert E.resolve_sampler("DPM++ 2M", "Exponential") == ("dpmpp_2m", "exponential")
    assert E.resolve_sampler("Euler a", "SGM Uniform") == ("euler_a", "sgm_uniform")
    assert E.resolve_sampler("DDIM", "Karras") == ("ddim", None)   # the timestep samplers take no sigma schedule
    with pytest.raises(ValueError):
        E.resolve_sampler("DPM++ 2M", "Align Your Steps")


def test_hires_fix_matches_oracle(env):
    """first pass -> bilinear latent resize -> DDIM img2img from t_enc at the large size -> decode"""
    C, E, O, cfgs, sd, eng = env
    b = 2
    tok = O.random_prompt_tokens(b, vocab_hi=997)
    neg = O.empty_prompt_tokens(b, vocab_hi=997)
    with torch.no_grad():
        ref_u8, ref_x = O.txt2img_hires(sd, *cfgs, tok, neg, seed=77, steps=5, cfg_scale=7.0, height=64, width=64,
                                        hr_scale=2.0, hr_steps=6, denoising_strength=0.6)
    got = eng.txt2img_hires(tok, neg, seed=77, steps=5, cfg_scale=7.0, height=64, width=64, hr_scale=2.0, hr_steps=6,
                            denoising_strength=0.6)
    assert got.shape == ref_u8.shape and eng.last_unet_evals == int(0.6 * 6) - 1
    d = (got.int() - ref_u8.int()).abs()
    assert float((d <= 1).float().mean()) == 1.0 and float((d == 0).float().mean()) > 0.99


@pytest.mark.parametrize("sampler,sched,method,sig_fn", [("Euler", None, "euler", "karras_sigmas_compvis"),
                                                         ("Euler a", None, "euler_a", "karras_sigmas_compvis"),
                                                         ("DPM++ 2M", None, "dpmpp_2m", "sigmas_karras"),
                                                         ("DPM++ 2M", "Exponential", "dpmpp_2m", "sigmas_exponential")])
def test_img2img_on_kdiffusion_samplers(env, sampler, sched, method, sig_fn):
    """img2img (and the hires second pass) with the k-diffusion samplers: the tail sigmas[steps - t_enc - 1:] of the
    sampler's schedule from init + noise * sigma (sdwui KDiffusionSampler.sample_img2img)"""
    C, E, O, cfgs, sd, eng = env
    b, size, steps, d = 2, 32, 8, 0.6
    g = torch.Generator().manual_seed(99)
    init = torch.randint(0, 256, (b, size, size, 3), generator=g, dtype=torch.uint8)
    tok = O.random_prompt_tokens(b, vocab_hi=997)
    neg = O.empty_prompt_tokens(b, vocab_hi=997)
    cond, unc = O.clip_text_encode(sd, cfgs[2], tok), O.clip_text_encode(sd, cfgs[2], neg)
    sig, log_sig = getattr(O, sig_fn)(steps)
    sched_sig = O.kdiff_img2img_sigmas(sig, steps, d)
    n_evals = len(sched_sig) - 1
    assert n_evals == int(d * steps) + 1
    with torch.no_grad():
        lat0 = O.vae_encode_mean(sd, cfgs[1], O.image_to_model_input(init)) * cfgs[1].scale_factor
        nz = E.per_image_noise(555, b, tuple(lat0.shape[1:]), 1 + (n_evals if method == "euler_a" else 0))
        ref = O.sample_kdiff_img2img(lambda x, t, c: O.unet_forward(sd, cfgs[0], x, t, c), lat0, list(nz), cond, unc,
                                     sched_sig, log_sig, 7.0, method)
        ref_u8 = O.to_uint8(O.vae_decode(sd, cfgs[1], ref / cfgs[1].scale_factor))
    got = eng.img2img(tok, neg, 555, init, d, steps=steps, cfg_scale=7.0, sampler=sampler, scheduler=sched)
    assert eng.last_unet_evals == n_evals and got.shape == ref_u8.shape
    h, w = lat0.shape[2:]
    z = eng.plan(b, h, w).x.reshape(b, h, w, 4).permute(0, 3, 1, 2)
    assert float((z - ref).abs().max()) <= 1e-3 * float(ref.abs().max())
    dd = (got.int() - ref_u8.int()).abs()
    assert float((dd <= 1).float().mean()) == 1.0 and float((dd == 0).float().mean()) > 0.99


def test_hires_fix_on_a_kdiffusion_sampler(env):
    """hires fix with "Euler": first pass sample_euler, bilinear latent resize, then the tail of the same sampler's
    schedule from the noised upscaled latents (sdwui sample_hr_pass -> sampler.sample_img2img)"""
    C, E, O, cfgs, sd, eng = env
    b, steps, hr_steps, d = 2, 5, 6, 0.5
    tok = O.random_prompt_tokens(b, vocab_hi=997)
    neg = O.empty_prompt_tokens(b, vocab_hi=997)
    cond, unc = O.clip_text_encode(sd, cfgs[2], tok), O.clip_text_encode(sd, cfgs[2], neg)
    unet = lambda x, t, c: O.unet_forward(sd, cfgs[0], x, t, c)  # noqa: E731
    with torch.no_grad():
        x = O.sample_euler(unet, O.per_image_noise(88, b, (4, 8, 8)), cond, unc, steps, 7.0)
        up = torch.nn.functional.interpolate(x, size=(16, 16), mode="bilinear", antialias=False)
        sig, log_sig = O.karras_sigmas_compvis(hr_steps)
        ref = O.sample_kdiff_img2img(unet, up, [O.per_image_noise(88, b, (4, 16, 16))], cond, unc,
                                     O.kdiff_img2img_sigmas(sig, hr_steps, d), log_sig, 7.0, "euler")
        ref_u8 = O.to_uint8(O.vae_decode(sd, cfgs[1], ref / cfgs[1].scale_factor))
    got = eng.txt2img_hires(tok, neg, seed=88, steps=steps, cfg_scale=7.0, height=64, width=64, hr_scale=2.0,
                            hr_steps=hr_steps, denoising_strength=d, sampler="Euler")
    assert got.shape == ref_u8.shape and eng.last_unet_evals == int(d * hr_steps) + 1
    dd = (got.int() - ref_u8.int()).abs()
    assert float((dd <= 1).float().mean()) == 1.0 and float((dd == 0).float().mean()) > 0.99


@pytest.mark.parametrize("blur,invert", [(4, False), (0, True)])
def test_inpainting_matches_oracle(env, blur, invert):
    """img2img with a mask: mask pipeline on the host, the kept region forced back to the init latents
    before every DDIM evaluation and after the last one, original pixels composited back through the blurred mask"""
    from PIL import Image, ImageDraw
    from b200sd import inpaint as inp
    C, E, O, cfgs, sd, eng = env
    b, size, steps = 2, 32, 8
    g = torch.Generator().manual_seed(77)
    init = torch.randint(0, 256, (b, size, size, 3), generator=g, dtype=torch.uint8)
    mask_img = Image.new("L", (size, size), 0)
    ImageDraw.Draw(mask_img).ellipse((6, 8, 24, 26), fill=255)
    tok = O.random_prompt_tokens(b, vocab_hi=997)
    neg = O.empty_prompt_tokens(b, vocab_hi=997)
    with torch.no_grad():
        ref_u8, ref_x = O.img2img_inpaint(sd, *cfgs, tok, neg, 4242, init, mask_img, 0.75, steps=steps, mask_blur=blur,
                                          invert=invert)
    down = 2 ** (len(cfgs[1].ch_mult) - 1)
    m = inp.prepare_mask(mask_img, size, size, size // down, size // down, mask_blur=blur, invert=invert)
    assert m.latmask.shape == ((size // down) ** 2,) and 0 < float(m.latmask.mean()) < 1
    assert set(m.latmask.unique().tolist()) <= {0.0, 1.0}
    got = eng.img2img(tok, neg, 4242, init, 0.75, steps=steps, cfg_scale=7.0, latmask=m.latmask)
    h = size // down
    z = eng.plan(b, h, h).x.reshape(b, h, h, 4).permute(0, 3, 1, 2)
    assert float((z - ref_x).abs().max()) <= 1e-3 * float(ref_x.abs().max())
    # the kept region carries the init latents exactly
    lat0 = eng.encode(init)
    keep = (m.latmask.reshape(h, h) == 0)[None, None].expand_as(z)
    assert torch.equal(z[keep], lat0[keep])
    final = inp.apply_overlays(got.cpu(), inp.overlays_for(init, m))
    d = (final.int() - ref_u8.int()).abs()
    assert float((d <= 1).float().mean()) == 1.0 and float((d == 0).float().mean()) > 0.99
    # the k-diffusion samplers take masks too (tests/test_samplers_cpu.py compares them with the oracle): kept region intact
    got_e = eng.img2img(tok, neg, 1, init, 0.75, steps=steps, cfg_scale=7.0, sampler="Euler", latmask=m.latmask)
    z_e = eng.plan(b, h, h).x.reshape(b, h, h, 4).permute(0, 3, 1, 2)
    assert got_e.shape == got.shape and torch.equal(z_e[keep], lat0[keep])


@pytest.mark.parametrize("fill", [0, 2, 3])
def test_inpainting_fill_modes(env, fill):
    """inpainting_fill 0 ("fill": image-space blur cascade before encoding), 2 ("latent noise"), 3 ("latent nothing")"""
    from PIL import Image, ImageDraw
    from b200sd import inpaint as inp
    C, E, O, cfgs, sd, eng = env
    b, size, steps = 2, 32, 6
    g = torch.Generator().manual_seed(78)
    init = torch.randint(0, 256, (b, size, size, 3), generator=g, dtype=torch.uint8)
    mask_img = Image.new("L", (size, size), 0)
    ImageDraw.Draw(mask_img).rectangle((8, 8, 23, 25), fill=255)
    tok = O.random_prompt_tokens(b, vocab_hi=997)
    neg = O.empty_prompt_tokens(b, vocab_hi=997)
    with torch.no_grad():
        ref_u8, ref_x = O.img2img_inpaint(sd, *cfgs, tok, neg, 99, init, mask_img, 0.75, steps=steps, mask_blur=2,
                                          inpainting_fill=fill)
    down = 2 ** (len(cfgs[1].ch_mult) - 1)
    m = inp.prepare_mask(mask_img, size, size, size // down, size // down, mask_blur=2)
    enc_in = inp.fill_masked(init, m) if fill == 0 else init
    if fill == 0:
        assert not torch.equal(enc_in, init)
    got = eng.img2img(tok, neg, 99, enc_in, 0.75, steps=steps, cfg_scale=7.0, latmask=m.latmask, inpainting_fill=fill)
    h = size // down
    z = eng.plan(b, h, h).x.reshape(b, h, h, 4).permute(0, 3, 1, 2)
    assert float((z - ref_x).abs().max()) <= 1e-3 * float(ref_x.abs().max())
    final = inp.apply_overlays(got.cpu(), inp.overlays_for(init, m))
    d = (final.int() - ref_u8.int()).abs()
    assert float((d <= 1).float().mean()) == 1.0 and float((d == 0).float().mean()) > 0.99


def test_img2img_matches_oracle(env):
    """VAE encoder program (asymmetric stride-2 padding) + DDIM started at t_enc + decode."""
    C, E, O, cfgs, sd, eng = env
    b, size, steps = 3, 32, 8   # the reduced VAE is f2: 32 px -> 16 x 16 latents; b=3 exercises the ragged chunk
    g = torch.Generator().manual_seed(4321)
    init = torch.randint(0, 256, (b, size, size, 3), generator=g, dtype=torch.uint8)
    tok = O.random_prompt_tokens(b, vocab_hi=997)
    neg = O.empty_prompt_tokens(b, vocab_hi=997)
    with torch.no_grad():
        ref_u8, ref_x, ref_init = O.img2img(sd, *cfgs, tok, neg, 1000, init, 0.75, steps=steps)
    lat0 = eng.encode(init)
    assert float((lat0 - ref_init).abs().max()) <= 1e-3 * float(ref_init.abs().max())
    sa, s1a, ts, rows = E.ddim_img2img_plan(steps, 0.75)
    rsa, rs1a, rrows = O.ddim_img2img_coefficients(steps, 0.75)
    assert abs(sa - rsa) < 1e-12 and abs(s1a - rs1a) < 1e-12 and len(rows) == len(rrows) == int(0.75 * steps) - 1
    got = eng.img2img(tok, neg, 1000, init, 0.75, steps=steps, cfg_scale=7.0)
    assert eng.last_unet_evals == len(rrows)
    assert got.shape == ref_u8.shape
    d = (got.int() - ref_u8.int()).abs()
    assert float((d <= 1).float().mean()) == 1.0 and float((d == 0).float().mean()) > 0.99


@pytest.mark.parametrize("size,box,proc", [((96, 64), (20, 10, 50, 30), (64, 64)), ((64, 96), (2, 60, 30, 94), (64, 32)),
                                            ((80, 80), (0, 0, 80, 12), (32, 64))])
def test_only_masked_inpainting_geometry(size, box, proc):
    """inpaint_full_res ("Only masked", reference worker.py:406-410 forwards the field): crop region, processing-size init,
    latent mask, overlay and the final paste-back equal the oracle's restatement of sdwui's Img2Img.init / apply_overlay"""
    import numpy as np
    from PIL import Image, ImageDraw
    from b200sd import inpaint as inp
    from oracle import sd_oracle as O
    g = torch.Generator().manual_seed(3)
    w, h = size
    init = Image.fromarray(torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8).numpy(), "RGB")
    mask = Image.new("L", size, 0)
    ImageDraw.Draw(mask).ellipse(box, fill=255)
    pw, ph = proc
    crop, paste_to, proc_init, latmask, overlay = O.only_masked_setup(init, mask, pw, ph, ph // 8, pw // 8, mask_blur=2, padding=6)
    m = inp.prepare_mask_only_masked(mask, pw, ph, ph // 8, pw // 8, mask_blur=2, padding=6)
    assert m.crop == tuple(crop) and m.paste_to == tuple(paste_to)
    assert torch.equal(m.latmask.reshape(ph // 8, pw // 8), latmask)
    mine = inp.crop_init_images([init], m)
    assert torch.equal(mine[0], torch.from_numpy(np.array(proc_init)))
    full = torch.from_numpy(np.asarray(init))[None]
    ov = inp.overlays_for(full, m)
    assert np.array_equal(np.array(ov[0]), np.array(overlay))
    gen = torch.randint(0, 256, (1, ph, pw, 3), generator=g, dtype=torch.uint8)
    ref = O.apply_overlay(Image.fromarray(gen[0].numpy(), "RGB"), paste_to, overlay)
    out = inp.apply_overlays(gen, ov, m.paste_to)
    assert out.shape == (1, h, w, 3) and torch.equal(out[0], torch.from_numpy(np.array(ref)))
    assert inp.prepare_mask_only_masked(Image.new("L", size, 0), pw, ph, ph // 8, pw // 8, mask_blur=0) is None


def test_only_masked_crop_geometry_random():
    """b200sd.inpaint.masked_region / grow_to_aspect (vectorised) against the oracle's loop-by-loop restatement of sdwui's
    masking.get_crop_region / expand_crop_region on a few hundred random masks, paddings and aspect ratios — including
    regions that touch the borders, single-pixel masks and boxes that must be shifted back inside the picture"""
    import numpy as np
    from hypothesis import given, settings, strategies as st
    from b200sd import inpaint as inp
    from oracle import sd_oracle as O

    @settings(max_examples=400, deadline=None)
    @given(w=st.integers(8, 96), h=st.integers(8, 96), pad=st.integers(0, 40), pw=st.sampled_from([32, 64, 96, 128]),
           ph=st.sampled_from([32, 64, 96, 128]), data=st.data())
    def run(w, h, pad, pw, ph, data):
        x1 = data.draw(st.integers(0, w - 1))
        x2 = data.draw(st.integers(x1 + 1, w))
        y1 = data.draw(st.integers(0, h - 1))
        y2 = data.draw(st.integers(y1 + 1, h))
        mask = np.zeros((h, w), dtype=np.uint8)
        mask[y1:y2, x1:x2] = data.draw(st.sampled_from([1, 128, 255]))
        if data.draw(st.booleans()):   # a second blob somewhere else
            mask[data.draw(st.integers(0, h - 1)), data.draw(st.integers(0, w - 1))] = 255
        region = inp.masked_region(mask, pad)
        assert region == tuple(O.get_crop_region(mask, pad))
        assert inp.grow_to_aspect(region, pw, ph, w, h) == tuple(O.expand_crop_region(region, pw, ph, w, h))

    run()
    assert inp.masked_region(np.zeros((5, 7), dtype=np.uint8), 3) is None
