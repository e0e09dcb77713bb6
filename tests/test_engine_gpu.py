"""GPU parity of the whole numeric path (UNet eval, sampler loop, VAE decode, uint8 images) against the fp32
PyTorch oracle (oracle/sd_oracle.py) on the same seeded synthetic weights and inputs.

Stated tolerances (fp16 kernels with fp32 accumulation vs fp32 oracle; "parity unpinned" — see oracle header):
  * one UNet evaluation:        max |d eps| <= 3e-2 * max |eps|   and  rms(d) <= 5e-3 * rms(eps)
  * VAE decode (float image):   max |d| <= 3e-2, mean |d| <= 3e-3   (image range [-1, 1])
  * uint8 image after a full sampler run: mean |d| <= 1.5 LSB, >= 97 % of pixels within 2 LSB (the truncating
    uint8 cast turns fp16 noise into +-1 LSB flips; sampler steps compound rounding differences)
"""
import json
import os

import pytest
import torch

from kutil import OUT_DIR

pytestmark = pytest.mark.gpu


def _record(name, **kw):
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, "engine_parity.jsonl"), "a") as f:
        f.write(json.dumps(dict(name=name, **kw)) + "\n")


@pytest.fixture(scope="module")
def mods():
    from b200sd import config, engine, synth
    from oracle import sd_oracle
    return config, engine, synth, sd_oracle


def _setup(mods, size):
    C, E, S, O = mods
    if size == "tiny":
        cfgs = (C.TINY_UNET, C.TINY_VAE, C.TINY_CLIP)
        vocab_hi = 997
    else:
        cfgs = (C.SD15_UNET, C.SD15_VAE, C.SD15_CLIP)
        vocab_hi = 49405
    sd = S.make_state_dict(*cfgs, seed=0)
    eng = E.SDEngine(sd, *cfgs, device="cuda:0", use_graphs=False)
    dsd = {k: v.cuda() for k, v in sd.items()}
    return cfgs, sd, dsd, eng, vocab_hi


_CACHE = {}


def _get(mods, size):
    if size not in _CACHE:
        _CACHE.clear()
        torch.cuda.empty_cache()
        _CACHE[size] = _setup(mods, size)
    return _CACHE[size]


@pytest.mark.parametrize("size,b,hw", [("tiny", 2, 16), ("tiny", 1, 32), ("sd15", 1, 64), ("sd15", 3, 64)])
def test_unet_eval_parity(mods, size, b, hw):
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, size)
    tok = O.random_prompt_tokens(b, vocab_hi=vocab_hi)
    neg = O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    cond32 = O.clip_text_encode(dsd, cfgs[2], tok.cuda())
    unc32 = O.clip_text_encode(dsd, cfgs[2], neg.cuda())
    x = O.per_image_noise(1000, b, (4, hw, hw)).cuda()
    t_val = 651.0
    with torch.no_grad():
        ref = O.unet_forward(dsd, cfgs[0], torch.cat([x, x]), torch.full((2 * b,), t_val, device="cuda"),
                             torch.cat([cond32, unc32]))
    plan = eng.plan(b, hw, hw)
    plan.unet.set_context(torch.cat([cond32, unc32]).half().contiguous())
    table = eng.temb.table(torch.tensor([t_val]))
    plan.table[:1].copy_(table)
    plan.step.zero_()
    plan.x.copy_(x.permute(0, 2, 3, 1).reshape(b, hw * hw, 4))
    from b200sd import ops
    ops.pack_unet_input(plan.x, plan.unet.xin, 1.0)
    ops.select_step(plan.table, plan.step, plan.unet.cur_bias)
    plan.unet.run()
    torch.cuda.synchronize()
    got = plan.unet.eps[..., :4].float().reshape(2 * b, hw, hw, 4).permute(0, 3, 1, 2)
    d = (got - ref).abs()
    rel_max = float(d.max() / ref.abs().max())
    rel_rms = float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    _record(f"unet_eval {size} b{b} hw{hw}", rel_max=rel_max, rel_rms=rel_rms, ref_absmax=float(ref.abs().max()),
            nan=bool(torch.isnan(got).any()))
    assert not torch.isnan(got).any()
    assert rel_max <= 3e-2 and rel_rms <= 5e-3, (rel_max, rel_rms)


@pytest.mark.parametrize("size,b,hw", [("tiny", 2, 16), ("sd15", 2, 64)])
def test_vae_decode_parity(mods, size, b, hw):
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, size)
    z = O.per_image_noise(77, b, (4, hw, hw)).cuda() * 0.18215 * 4.0
    with torch.no_grad():
        ref = O.vae_decode(dsd, cfgs[1], z / cfgs[1].scale_factor)
    lat = z.permute(0, 2, 3, 1).reshape(b, hw * hw, 4).contiguous()
    u8 = eng.decode(lat, hw, hw)
    torch.cuda.synchronize()
    vae = eng.plan(b, hw, hw).vae
    got = vae.img[:b, :, :3].float().reshape(b, vae.out_h, vae.out_w, 3).permute(0, 3, 1, 2)
    d = (got - ref).abs()
    ref_u8 = O.to_uint8(ref)
    du8 = (u8.int() - ref_u8.int()).abs().float()
    _record(f"vae_decode {size} b{b} hw{hw}", max_abs=float(d.max()), mean_abs=float(d.mean()),
            u8_mean=float(du8.mean()), u8_max=float(du8.max()), u8_within1=float((du8 <= 1).float().mean()),
            ref_std=float(ref.std()), sat_frac=float(((ref_u8 == 0) | (ref_u8 == 255)).float().mean()))
    assert float(d.max()) <= 3e-2 and float(d.mean()) <= 3e-3
    assert float((du8 <= 1).float().mean()) >= 0.999


@pytest.mark.parametrize("size,b,hw,steps,graphs", [("tiny", 2, 16, 6, False), ("tiny", 2, 16, 6, True),
                                                     ("sd15", 2, 64, 20, True)])
def test_txt2img_parity(mods, size, b, hw, steps, graphs):
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, size)
    eng.use_graphs = graphs
    tok = O.random_prompt_tokens(b, vocab_hi=vocab_hi)
    neg = O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    with torch.no_grad():
        ref_u8, ref_z, ref_dec = O.txt2img(dsd, *cfgs, tok, neg, seed=1000, steps=steps, cfg_scale=7.0, height=hw * 8,
                                           width=hw * 8, device="cuda")
    got = eng.txt2img(tok, neg, seed=1000, steps=steps, cfg_scale=7.0, height=hw * 8, width=hw * 8, sampler="DDIM")
    torch.cuda.synchronize()
    eng.use_graphs = False
    assert eng.last_unet_evals == len(O.ddim_timesteps(steps)) - 1  # 19 for the 20-step configs
    plan = eng.plan(b, hw, hw)
    z = plan.x.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    dz = (z - ref_z).abs()
    du8 = (got.int() - ref_u8.int()).abs().float()
    _record(f"txt2img {size} b{b} hw{hw} steps{steps} graphs{graphs}", z_rel_max=float(dz.max() / ref_z.abs().max()),
            z_rel_rms=float(dz.pow(2).mean().sqrt() / ref_z.pow(2).mean().sqrt()), u8_mean=float(du8.mean()),
            u8_max=float(du8.max()), u8_within2=float((du8 <= 2).float().mean()),
            u8_exact=float((du8 == 0).float().mean()), ref_mean=float(ref_u8.float().mean()),
            sat_frac=float(((ref_u8 == 0) | (ref_u8 == 255)).float().mean()))
    assert got.shape == ref_u8.shape
    assert float(du8.mean()) <= 1.5 and float((du8 <= 2).float().mean()) >= 0.97


def test_sd15_sampler_is_bit_reproducible(mods):
    """No kernel on the path uses floating-point atomics or an unordered reduction: the same request twice gives the
    same bytes (also a sharp race detector: 148 persistent CTAs x thousands of tiles per run)."""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "sd15")
    tok = O.random_prompt_tokens(3, vocab_hi=vocab_hi)
    neg = O.empty_prompt_tokens(3, vocab_hi=vocab_hi)
    eng.use_graphs = True
    runs = [eng.txt2img(tok, neg, seed=77, steps=8, cfg_scale=7.0, height=512, width=512, sampler="DDIM").clone()
            for _ in range(3)]
    eng.use_graphs = False
    torch.cuda.synchronize()
    assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])


@pytest.mark.parametrize("size,px", [("tiny", 128), ("sd15", 512)])
def test_images_do_not_depend_on_their_batch(mods, size, px):
    """image k is a function of (prompt, seed + k) only — the premise of sharding a request's batch over
    workers.  Here it holds bit for bit: no kernel's reduction order for one image depends on the other images of the
    launch (GroupNorm's CTA split is per image shape, GEMM K loops and attention rows are per element), so the last two
    images of a batch of 5 equal a batch of 2 started at seed + 3."""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, size)
    tok = O.random_prompt_tokens(1, vocab_hi=vocab_hi)
    neg = O.empty_prompt_tokens(1, vocab_hi=vocab_hi)
    eng.use_graphs = True
    five = eng.txt2img(tok.expand(5, -1), neg.expand(5, -1), seed=300, steps=6, cfg_scale=7.0, height=px, width=px).clone()
    two = eng.txt2img(tok.expand(2, -1), neg.expand(2, -1), seed=303, steps=6, cfg_scale=7.0, height=px, width=px).clone()
    eng.use_graphs = False
    torch.cuda.synchronize()
    assert torch.equal(five[3:], two)


@pytest.mark.parametrize("size,b,px,steps", [("tiny", 3, 64, 8), ("sd15", 2, 512, 20)])
def test_img2img_parity(mods, size, b, px, steps):
    """config C3: VAE encode (posterior mean) + noise to t_enc + DDIM remainder + decode, vs the fp32 oracle."""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, size)
    g = torch.Generator().manual_seed(4321)
    init = torch.randint(0, 256, (b, px, px, 3), generator=g, dtype=torch.uint8)
    tok = O.random_prompt_tokens(b, vocab_hi=vocab_hi)
    neg = O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    with torch.no_grad():
        ref_u8, ref_x, ref_init = O.img2img(dsd, *cfgs, tok, neg, 1000, init, 0.75, steps=steps, device="cuda")
    lat0 = eng.encode(init)
    enc_rel = float((lat0 - ref_init).abs().max() / ref_init.abs().max())
    eng.use_graphs = True
    got = eng.img2img(tok, neg, 1000, init, 0.75, steps=steps, cfg_scale=7.0)
    torch.cuda.synchronize()
    eng.use_graphs = False
    assert eng.last_unet_evals == int(0.75 * steps) - 1
    du8 = (got.int() - ref_u8.int()).abs().float()
    _record(f"img2img {size} b{b} {px}px steps{steps}", enc_rel_max=enc_rel, u8_mean=float(du8.mean()),
            u8_max=float(du8.max()), u8_within2=float((du8 <= 2).float().mean()), u8_exact=float((du8 == 0).float().mean()))
    assert enc_rel <= 2e-2
    assert float(du8.mean()) <= 1.5 and float((du8 <= 2).float().mean()) >= 0.97


def test_hires_fix_parity_tiny(mods):
    """hires fix, "Latent" upscaler (bilinear latent resize kernel + DDIM img2img second pass)"""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "tiny")
    b = 2
    tok = O.random_prompt_tokens(b, vocab_hi=vocab_hi)
    neg = O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    with torch.no_grad():
        ref_u8, ref_x = O.txt2img_hires(dsd, *cfgs, tok, neg, seed=900, steps=6, cfg_scale=7.0, height=64, width=64,
                                        hr_scale=2.0, hr_steps=8, denoising_strength=0.7, device="cuda")
    got = eng.txt2img_hires(tok, neg, seed=900, steps=6, cfg_scale=7.0, height=64, width=64, hr_scale=2.0, hr_steps=8,
                            denoising_strength=0.7)
    torch.cuda.synchronize()
    du8 = (got.int() - ref_u8.int()).abs().float()
    _record("hires tiny", u8_mean=float(du8.mean()), u8_max=float(du8.max()), u8_exact=float((du8 == 0).float().mean()))
    assert got.shape == ref_u8.shape
    assert float(du8.mean()) <= 1.5 and float((du8 <= 2).float().mean()) >= 0.97


def test_euler_parity_tiny(mods):
    """sdwui "Euler" (k-diffusion sample_euler, s_churn 0) through the worker-facing txt2img call"""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "tiny")
    b, hw, steps = 2, 16, 6
    tok = O.random_prompt_tokens(b, vocab_hi=vocab_hi)
    neg = O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    x_T = O.per_image_noise(2100, b, (4, hw, hw))
    cond32 = O.clip_text_encode(dsd, cfgs[2], tok.cuda())
    unc32 = O.clip_text_encode(dsd, cfgs[2], neg.cuda())
    with torch.no_grad():
        ref = O.sample_euler(lambda x, t, c: O.unet_forward(dsd, cfgs[0], x, t, c), x_T.cuda(), cond32, unc32, steps, 7.0)
    got_u8 = eng.txt2img(tok, neg, seed=2100, steps=steps, cfg_scale=7.0, height=hw * 8, width=hw * 8, sampler="Euler")
    torch.cuda.synchronize()
    z = eng.plan(b, hw, hw).x.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    rel = float((z - ref).abs().max() / ref.abs().max())
    _record("euler tiny", z_rel_max=rel)
    assert rel <= 3e-2 and got_u8.shape[0] == b and got_u8.dtype == torch.uint8


@pytest.mark.parametrize("name,sched,karras", [("DPM++ 2M", None, True), ("DPM++ 2M Karras", None, True),
                                               ("DPM++ 2M", "Uniform", False)])
def test_dpmpp_2m_parity_tiny(mods, name, sched, karras):
    """sdwui "DPM++ 2M" (k-diffusion sample_dpmpp_2m) with the Karras / uniform noise schedules, through txt2img"""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "tiny")
    b, hw, steps = 2, 16, 7
    tok = O.random_prompt_tokens(b, vocab_hi=vocab_hi)
    neg = O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    x_T = O.per_image_noise(2200, b, (4, hw, hw))
    cond32 = O.clip_text_encode(dsd, cfgs[2], tok.cuda())
    unc32 = O.clip_text_encode(dsd, cfgs[2], neg.cuda())
    with torch.no_grad():
        ref = O.sample_dpmpp_2m(lambda x, t, c: O.unet_forward(dsd, cfgs[0], x, t, c), x_T.cuda(), cond32, unc32, steps, 7.0,
                                karras=karras)
    got_u8 = eng.txt2img(tok, neg, seed=2200, steps=steps, cfg_scale=7.0, height=hw * 8, width=hw * 8, sampler=name,
                         scheduler=sched)
    torch.cuda.synchronize()
    assert eng.last_unet_evals == steps
    z = eng.plan(b, hw, hw).x.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    rel = float((z - ref).abs().max() / ref.abs().max())
    _record(f"dpmpp_2m tiny {name} {sched}", z_rel_max=rel)
    assert rel <= 3e-2 and got_u8.shape[0] == b and got_u8.dtype == torch.uint8


def test_img2img_dpmpp_2m_parity_tiny(mods):
    """img2img on a k-diffusion sampler: the tail of the Karras schedule from init + noise * sigma (t_enc + 1 evaluations)"""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "tiny")
    b, size, steps, d = 2, 64, 8, 0.6
    g = torch.Generator().manual_seed(31)
    init = torch.randint(0, 256, (b, size, size, 3), generator=g, dtype=torch.uint8)
    tok = O.random_prompt_tokens(b, vocab_hi=vocab_hi)
    neg = O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    cond32 = O.clip_text_encode(dsd, cfgs[2], tok.cuda())
    unc32 = O.clip_text_encode(dsd, cfgs[2], neg.cuda())
    sig, log_sig = O.sigmas_karras(steps)
    sched = O.kdiff_img2img_sigmas(sig, steps, d)
    with torch.no_grad():
        lat0 = O.vae_encode_mean(dsd, cfgs[1], O.image_to_model_input(init.cuda())) * cfgs[1].scale_factor
        nz = O.per_image_noise(777, b, tuple(lat0.shape[1:])).cuda()
        ref = O.sample_kdiff_img2img(lambda x, t, c: O.unet_forward(dsd, cfgs[0], x, t, c), lat0, [nz], cond32, unc32, sched,
                                     log_sig, 7.0, "dpmpp_2m")
    got_u8 = eng.img2img(tok, neg, 777, init, d, steps=steps, cfg_scale=7.0, sampler="DPM++ 2M")
    torch.cuda.synchronize()
    assert eng.last_unet_evals == len(sched) - 1 == int(d * steps) + 1
    h, w = lat0.shape[2:]
    z = eng.plan(b, h, w).x.reshape(b, h, w, 4).permute(0, 3, 1, 2)
    rel = float((z - ref).abs().max() / ref.abs().max())
    _record("img2img dpmpp_2m tiny", z_rel_max=rel)
    assert rel <= 3e-2 and tuple(got_u8.shape) == (b, size, size, 3)


def test_euler_a_parity_tiny(mods):
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "tiny")
    b, hw, steps = 2, 16, 5
    tok = O.random_prompt_tokens(b, vocab_hi=vocab_hi)
    neg = O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    nz = E.per_image_noise(2000, b, (4, hw, hw), 1 + steps)
    cond32 = O.clip_text_encode(dsd, cfgs[2], tok.cuda())
    unc32 = O.clip_text_encode(dsd, cfgs[2], neg.cuda())
    with torch.no_grad():
        ref = O.sample_euler_a(lambda x, t, c: O.unet_forward(dsd, cfgs[0], x, t, c), nz[0].cuda(), cond32, unc32, steps,
                               7.0, [n.cuda() for n in nz[1:]])
    lat = eng.sample(cond32.half(), unc32.half(), nz[0], steps, 7.0, "Euler a", noises=nz[1:])
    torch.cuda.synchronize()
    z = lat.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    rel = float((z - ref).abs().max() / ref.abs().max())
    _record("euler_a tiny", z_rel_max=rel)
    assert rel <= 3e-2
