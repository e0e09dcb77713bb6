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


@pytest.mark.parametrize("size,b,hw", [("tiny", 2, 16), ("sd15", 2, 64)])
def test_vae_against_third_party_autoencoder(mods, size, b, hw):
    """the CUDA VAE decoder and encoder against an implementation that is neither ours nor the oracle: the `Decoder` /
    `Encoder` classes of the FLUX autoencoder shipped in torchtitan (the ldm autoencoder with ldm's module names; the same
    state dict loads with strict=True), fp32 on the same GPU"""
    A = pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")
    import torch.nn.functional as F
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, size)
    vcfg = cfgs[1]
    f = 2 ** (len(vcfg.ch_mult) - 1)
    pre = "first_stage_model."
    dec = A.Decoder(ch=vcfg.ch, out_ch=3, ch_mult=list(vcfg.ch_mult), num_res_blocks=vcfg.num_res_blocks, in_channels=3,
                    resolution=hw * f, z_channels=vcfg.z_channels).eval().float().cuda()
    dec.load_state_dict({k[len(pre + "decoder."):]: v.float() for k, v in dsd.items() if k.startswith(pre + "decoder.")},
                        strict=True)
    enc = A.Encoder(resolution=hw * f, in_channels=3, ch=vcfg.ch, ch_mult=list(vcfg.ch_mult),
                    num_res_blocks=vcfg.num_res_blocks, z_channels=vcfg.z_channels).eval().float().cuda()
    enc.load_state_dict({k[len(pre + "encoder."):]: v.float() for k, v in dsd.items() if k.startswith(pre + "encoder.")},
                        strict=True)
    z = O.per_image_noise(78, b, (4, hw, hw)).cuda() * vcfg.scale_factor * 4.0
    with torch.no_grad():
        ref = dec(F.conv2d(z / vcfg.scale_factor, dsd[pre + "post_quant_conv.weight"].float(),
                           dsd[pre + "post_quant_conv.bias"].float()))
    ref_u8 = O.to_uint8(ref)
    u8 = eng.decode(z.permute(0, 2, 3, 1).reshape(b, hw * hw, 4).contiguous(), hw, hw)
    torch.cuda.synchronize()
    du8 = (u8.int() - ref_u8.int()).abs().float()
    # the encoder on that picture: posterior mean, scaled
    with torch.no_grad():
        x = ref_u8.float().permute(0, 3, 1, 2) / 127.5 - 1.0     # O.image_to_model_input
        mom = F.conv2d(enc(x), dsd[pre + "quant_conv.weight"].float(), dsd[pre + "quant_conv.bias"].float())
        ref_lat = mom.chunk(2, dim=1)[0] * vcfg.scale_factor
    lat = eng.encode(ref_u8)
    torch.cuda.synchronize()
    enc_rel = float((lat - ref_lat).abs().max() / ref_lat.abs().max())
    _record(f"vae vs third-party autoencoder {size} b{b} hw{hw}", u8_mean=float(du8.mean()), u8_max=float(du8.max()),
            u8_within1=float((du8 <= 1).float().mean()), enc_rel_max=enc_rel)
    assert float((du8 <= 1).float().mean()) >= 0.999 and float(du8.max()) <= 2
    assert enc_rel <= 1e-2, enc_rel


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


# ----------------------------------------------------------------------------------------------------------------
# round 2: what is measured is what is asserted.  Round-1 measurements of the uint8 images of full sampler runs: max 1 LSB,
# mean 0.09-0.15 LSB, 85-91 % of pixels identical (gpurun_out/engine_parity.jsonl); one SD1.5 UNet evaluation rel-rms 1.4e-3.
U8_MAX, U8_MEAN, U8_EXACT = 2, 0.3, 0.80
UNET_REL_RMS = 3e-3


def _u8_check(name, got, ref_u8, u8_max=U8_MAX, **extra):
    du8 = (got.int().cpu() - ref_u8.int().cpu()).abs().float()
    rec = dict(u8_mean=float(du8.mean()), u8_max=float(du8.max()), u8_exact=float((du8 == 0).float().mean()), **extra)
    _record(name, **rec)
    assert got.shape == ref_u8.shape
    assert rec["u8_max"] <= u8_max and rec["u8_mean"] <= U8_MEAN and rec["u8_exact"] >= U8_EXACT, rec
    return rec


def test_tight_unet_and_image_tolerances_sd15(mods):
    """the round-1 tests above keep their loose historical bounds; this one asserts what is actually measured"""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "sd15")
    b, hw = 2, 64
    tok, neg = O.random_prompt_tokens(b, vocab_hi=vocab_hi), O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    cond32, unc32 = O.clip_text_encode(dsd, cfgs[2], tok.cuda()), O.clip_text_encode(dsd, cfgs[2], neg.cuda())
    x = O.per_image_noise(1000, b, (4, hw, hw)).cuda()
    with torch.no_grad():
        ref = O.unet_forward(dsd, cfgs[0], torch.cat([x, x]), torch.full((2 * b,), 651.0, device="cuda"), torch.cat([cond32, unc32]))
    from b200sd import ops
    plan = eng.plan(b, hw, hw)
    plan.unet.set_context(torch.cat([cond32, unc32]).half().contiguous())
    plan.table[:1].copy_(eng.temb.table(torch.tensor([651.0])))
    plan.step.zero_()
    plan.x.copy_(x.permute(0, 2, 3, 1).reshape(b, hw * hw, 4))
    ops.pack_unet_input(plan.x, plan.unet.xin, 1.0)
    ops.select_step(plan.table, plan.step, plan.unet.cur_bias)
    plan.unet.run()
    got = plan.unet.eps[..., :4].float().reshape(2 * b, hw, hw, 4).permute(0, 3, 1, 2)
    rel_rms = float((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    _record("unet_eval tight sd15", rel_rms=rel_rms)
    assert rel_rms <= UNET_REL_RMS
    with torch.no_grad():
        ref_u8, _, _ = O.txt2img(dsd, *cfgs, tok, neg, seed=1000, steps=20, cfg_scale=7.0, height=512, width=512, device="cuda")
    eng.use_graphs = True
    got_u8 = eng.txt2img(tok, neg, seed=1000, steps=20, cfg_scale=7.0, height=512, width=512, sampler="DDIM")
    eng.use_graphs = False
    _u8_check("txt2img tight sd15 ddim", got_u8, ref_u8)


def _oracle_run(O, dsd, cfgs, name, tok, neg, seed, steps, hw, draws_n, init=None, denoise=None, mask=None):
    from b200sd import engine as E
    b = tok.shape[0]
    cond32, unc32 = O.clip_text_encode(dsd, cfgs[2], tok.cuda()), O.clip_text_encode(dsd, cfgs[2], neg.cuda())
    nz = E.per_image_noise(seed, b, (4, hw, hw), 1 + draws_n).cuda()
    unet = lambda x, t, c: O.unet_forward(dsd, cfgs[0], x, t, c)  # noqa: E731
    with torch.no_grad():
        z = O.run_sampler(name, unet, cond32, unc32, 7.0, steps, nz[0], list(nz[1:]), init=init, denoising_strength=denoise,
                          mask=mask)
    return z


@pytest.mark.parametrize("name", ["Euler a", "DPM++ 2M Karras", "Heun", "DPM++ SDE Karras", "PLMS", "LMS"])
def test_sd15_sampler_parity(mods, name):
    """SD1.5-size, 512x512, 20 steps, CUDA graphs on: the samplers beyond DDIM against the oracle's k-diffusion / sdwui
    restatement — latents and decoded uint8 images"""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "sd15")
    b, hw, steps = 2, 64, 20
    tok, neg = O.random_prompt_tokens(b, vocab_hi=vocab_hi), O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    pr = eng.program(name, None, steps)
    ref_z = _oracle_run(O, dsd, cfgs, name, tok, neg, 5100, steps, hw, pr.draws)
    with torch.no_grad():
        ref_u8 = O.to_uint8(O.vae_decode(dsd, cfgs[1], ref_z / cfgs[1].scale_factor))
    eng.use_graphs = True
    got = eng.txt2img(tok, neg, seed=5100, steps=steps, cfg_scale=7.0, height=512, width=512, sampler=name)
    eng.use_graphs = False
    z = eng.plan(b, hw, hw).x.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    rel_rms = float((z - ref_z).pow(2).mean().sqrt() / ref_z.pow(2).mean().sqrt())
    _u8_check(f"sd15 sampler {name}", got, ref_u8, z_rel_rms=rel_rms, evals=eng.last_unet_evals)
    assert rel_rms <= 2e-2


@pytest.mark.parametrize("name", ["Euler", "LMS", "Heun", "DPM2", "DPM2 a", "DPM++ 2S a", "DPM++ SDE", "DPM fast", "DPM adaptive",
                                  "LMS Karras", "DPM2 Karras", "DPM2 a Karras", "DPM++ 2S a Karras", "DPM++ SDE Karras", "PLMS"])
def test_every_reference_sampler_on_gpu_tiny(mods, name):
    """reference scripts/spartan/worker.py:75-94: no name of the table falls back; each runs as stage graphs of
    b200sd_cfg_eps + b200sd_latent_lincomb and matches the oracle"""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "tiny")
    b, hw, steps = 2, 16, 7
    tok, neg = O.random_prompt_tokens(b, vocab_hi=vocab_hi), O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    pr = eng.program(name, None, steps)
    ref_z = _oracle_run(O, dsd, cfgs, name, tok, neg, 5200, steps, hw, pr.draws)
    eng.use_graphs = True
    eng.txt2img(tok, neg, seed=5200, steps=steps, cfg_scale=7.0, height=hw * 8, width=hw * 8, sampler=name)
    eng.use_graphs = False
    z = eng.plan(b, hw, hw).x.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    rel = float((z - ref_z).abs().max() / ref_z.abs().max())
    _record(f"tiny sampler {name}", z_rel_max=rel, evals=eng.last_unet_evals)
    assert rel <= 3e-2, rel


def test_euler_a_graph_reads_this_requests_noise(mods):
    """the Euler a step graph used to bake the FIRST request's noise tensor address.  Two requests with
    different seeds and step counts on one engine, graphs on, must equal the eager runs."""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "tiny")
    b, hw = 2, 16
    tok, neg = O.random_prompt_tokens(b, vocab_hi=vocab_hi), O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    reqs = [(11, 5), (9999, 9), (12345, 40), (7, 5)]     # 40 steps outgrows the 32-row noise stack: graphs are rebuilt
    eng.use_graphs = True
    with_graphs = [eng.txt2img(tok, neg, seed=s, steps=n, cfg_scale=7.0, height=hw * 8, width=hw * 8, sampler="Euler a").clone()
                   for s, n in reqs]
    eng.use_graphs = False
    eager = [eng.txt2img(tok, neg, seed=s, steps=n, cfg_scale=7.0, height=hw * 8, width=hw * 8, sampler="Euler a").clone()
             for s, n in reqs]
    for a, c in zip(with_graphs, eager):
        assert torch.equal(a, c)
    assert not torch.equal(with_graphs[0], with_graphs[3])


def test_step_graphs_per_plan_are_bounded(mods, monkeypatch):
    """every (sampler stage structure, cfg scale) captures a step graph: a client walking through cfg scales must not grow
    a plan without bound — the least recently used graphs are dropped and rebuilt on demand, with unchanged results"""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "tiny")
    monkeypatch.setattr(E, "MAX_GRAPHS", 4)
    b, hw = 2, 16
    tok, neg = O.random_prompt_tokens(b, vocab_hi=vocab_hi), O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    run = lambda cfg, name: eng.txt2img(tok, neg, seed=31, steps=5, cfg_scale=cfg, height=hw * 8, width=hw * 8,  # noqa: E731
                                        sampler=name).clone()
    eng.use_graphs = True
    try:
        first = {(cfg, name): run(cfg, name) for name in ("Euler", "Heun") for cfg in (3.0, 4.5, 6.0, 7.5)}
        plan = eng.plan(b, hw, hw)
        assert len(plan.graphs) <= 4 and set(plan.graph_launches) <= set(plan.graphs)
        again = {k: run(*k) for k in first}     # most of these graphs were evicted in the meantime
    finally:
        eng.use_graphs = False
    for k in first:
        assert torch.equal(first[k], again[k]), k
    assert torch.equal(first[(7.5, "Heun")], run(7.5, "Heun"))   # and equal to the eager run


def test_sd15_hires_fix_parity(mods):
    """hires fix at SD1.5 size: 256x256 first pass, Latent upscale x2, Euler a second pass from t_enc at 512x512"""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "sd15")
    b, steps, hr_steps, d = 2, 12, 10, 0.6
    tok, neg = O.random_prompt_tokens(b, vocab_hi=vocab_hi), O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    pr1, pr2 = eng.program("Euler a", None, steps), eng.program("Euler a", None, hr_steps, denoise=d)
    z1 = _oracle_run(O, dsd, cfgs, "Euler a", tok, neg, 5300, steps, 32, pr1.draws)
    up = torch.nn.functional.interpolate(z1, size=(64, 64), mode="bilinear", antialias=False)
    z2 = _oracle_run(O, dsd, cfgs, "Euler a", tok, neg, 5300, hr_steps, 64, pr2.draws, init=up, denoise=d)
    with torch.no_grad():
        ref_u8 = O.to_uint8(O.vae_decode(dsd, cfgs[1], z2 / cfgs[1].scale_factor))
    eng.use_graphs = True
    got = eng.txt2img_hires(tok, neg, seed=5300, steps=steps, cfg_scale=7.0, height=256, width=256, hr_scale=2.0,
                            hr_steps=hr_steps, denoising_strength=d, sampler="Euler a")
    eng.use_graphs = False
    _u8_check("sd15 hires euler_a", got, ref_u8)


def test_sd15_img2img_kdiffusion_parity(mods):
    """img2img at 512x512 on a k-diffusion sampler (DPM++ 2M Karras): encode, noise to sigma_sched[0], the schedule's tail"""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "sd15")
    b, steps, d = 2, 20, 0.75
    g = torch.Generator().manual_seed(4322)
    init = torch.randint(0, 256, (b, 512, 512, 3), generator=g, dtype=torch.uint8)
    tok, neg = O.random_prompt_tokens(b, vocab_hi=vocab_hi), O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    with torch.no_grad():
        lat0 = O.vae_encode_mean(dsd, cfgs[1], O.image_to_model_input(init.cuda())) * cfgs[1].scale_factor
    z = _oracle_run(O, dsd, cfgs, "DPM++ 2M Karras", tok, neg, 5400, steps, 64, 0, init=lat0, denoise=d)
    with torch.no_grad():
        ref_u8 = O.to_uint8(O.vae_decode(dsd, cfgs[1], z / cfgs[1].scale_factor))
    eng.use_graphs = True
    got = eng.img2img(tok, neg, 5400, init, d, steps=steps, cfg_scale=7.0, sampler="DPM++ 2M Karras")
    eng.use_graphs = False
    assert eng.last_unet_evals == int(d * steps) + 1
    _u8_check("sd15 img2img dpmpp_2m_karras", got, ref_u8)


@pytest.mark.parametrize("sampler", ["DDIM", "Euler a"])
def test_sd15_inpainting_parity(mods, sampler):
    """inpainting at SD1.5 size against the ORACLE (round 1 compared the engine with itself): mask pipeline, per-step blend
    (DDIM: model input; Euler a: denoised prediction), final blend, overlay composite"""
    from PIL import Image, ImageDraw
    from b200sd import inpaint as inp
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "sd15")
    b, steps, d = 2, 20, 0.75
    g = torch.Generator().manual_seed(4323)
    init = torch.randint(0, 256, (b, 512, 512, 3), generator=g, dtype=torch.uint8)
    mask_img = Image.new("L", (512, 512), 0)
    ImageDraw.Draw(mask_img).ellipse((120, 150, 400, 380), fill=255)
    tok, neg = O.random_prompt_tokens(b, vocab_hi=vocab_hi), O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    m = inp.prepare_mask(mask_img, 512, 512, 64, 64, mask_blur=4)
    if sampler == "DDIM":
        with torch.no_grad():
            ref_u8, _ = O.img2img_inpaint(sd, *cfgs, tok, neg, 5500, init, mask_img, d, steps=steps, mask_blur=4, device="cuda")
    else:
        with torch.no_grad():
            lat0 = O.vae_encode_mean(dsd, cfgs[1], O.image_to_model_input(init.cuda())) * cfgs[1].scale_factor
        latmask, overlay_mask = O.inpaint_masks(mask_img, 512, 512, 64, 64, 4, False)
        nmask = latmask[None, None].cuda()
        pr = eng.program(sampler, None, steps, denoise=d, masked=True)
        z = _oracle_run(O, dsd, cfgs, sampler, tok, neg, 5500, steps, 64, pr.draws, init=lat0, denoise=d, mask=(lat0, nmask))
        z = z * nmask + lat0 * (1 - nmask)
        with torch.no_grad():
            gen = O.to_uint8(O.vae_decode(dsd, cfgs[1], z / cfgs[1].scale_factor)).cpu()
        ref_u8 = torch.stack([torch.from_numpy(__import__("numpy").array(O.apply_overlay(
            Image.fromarray(gen[k].numpy(), "RGB"), None, inp.overlays_for(init[k:k + 1], m)[0]))) for k in range(b)])
    eng.use_graphs = True
    got = eng.img2img(tok, neg, 5500, init, d, steps=steps, cfg_scale=7.0, sampler=sampler, latmask=m.latmask).cpu()
    eng.use_graphs = False
    final = inp.apply_overlays(got, inp.overlays_for(init, m))
    # the latent mask is a hard 0/1 edge: a few pixels next to it decode 3 LSB apart (measured: max 3 for both samplers,
    # mean 0.05 LSB, 95 % identical) — the mean / exact bounds stay the tight ones.  The bound on the handful of edge pixels
    # is 5, not 4: the measurement predates the oracle's switch from cuDNN's default TF32 convolutions to IEEE fp32 (the
    # round's GPU minutes ended before this test could be re-measured; the SD1.5 tests that were re-measured moved by
    # 0.004 LSB in the mean and not at all in the maximum).
    _u8_check(f"sd15 inpainting {sampler}", final, ref_u8, u8_max=5)
    assert torch.equal(final[:, :60], init[:, :60])     # far outside the blurred mask: the original pixels


def test_bench_batch_spot_check_against_oracle(mods):
    """the benchmark's configuration itself — batch 32, 20 DDIM timesteps, graphs on — with images 0, 13 and 31 checked
    against the oracle run on each of them alone (image k depends on seed + k only)"""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "sd15")
    b = 32
    tok, neg = O.random_prompt_tokens(b, vocab_hi=vocab_hi), O.empty_prompt_tokens(b, vocab_hi=vocab_hi)
    eng.use_graphs = True
    got = eng.txt2img(tok, neg, seed=1000, steps=20, cfg_scale=7.0, height=512, width=512, sampler="DDIM").cpu()
    eng.use_graphs = False
    for k in (0, 13, 31):
        with torch.no_grad():
            ref_u8, _, _ = O.txt2img(dsd, *cfgs, tok[k:k + 1], neg[k:k + 1], seed=1000 + k, steps=20, cfg_scale=7.0, height=512,
                                     width=512, device="cuda")
        _u8_check(f"bench batch32 image {k}", got[k:k + 1], ref_u8)
    eng.plans.pop((32, 64, 64), None)
    torch.cuda.empty_cache()


def test_a_shard_of_a_large_batch_is_bit_identical_sd15(mods):
    """found at 8 GPUs in round 2: 17 images whole vs the shards World.optimize_jobs makes of them (3, 2, 2, ...) differed by
    1 LSB in 8 % of the pixels — the CLIP text tower is library GEMMs whose kernel choice (split-K) follows the batch, so the
    same prompt came out different in the last bit.  The conditioner now encodes unique prompts in fixed-size calls."""
    C, E, S, O = mods
    cfgs, sd, dsd, eng, vocab_hi = _get(mods, "sd15")
    tok = O.random_prompt_tokens(1, vocab_hi=vocab_hi)
    neg = O.empty_prompt_tokens(1, vocab_hi=vocab_hi)
    eng.use_graphs = True
    whole = eng.txt2img(tok.expand(17, -1), neg.expand(17, -1), seed=7000, steps=8, cfg_scale=7.0, height=512, width=512).clone()
    a = eng.txt2img(tok.expand(3, -1), neg.expand(3, -1), seed=7000, steps=8, cfg_scale=7.0, height=512, width=512).clone()
    c = eng.txt2img(tok.expand(2, -1), neg.expand(2, -1), seed=7015, steps=8, cfg_scale=7.0, height=512, width=512).clone()
    eng.use_graphs = False
    assert torch.equal(whole[:3], a) and torch.equal(whole[15:], c)
    # distinct prompts per image (the benchmark's case): still a function of (prompt, seed + k) only
    toks = O.random_prompt_tokens(17, vocab_hi=vocab_hi)
    e17 = eng.encode_prompts(toks)
    e2 = eng.encode_prompts(toks[9:11])
    assert torch.equal(e17[9:11], e2)
    eng.plans.pop((17, 64, 64), None)
    torch.cuda.empty_cache()
