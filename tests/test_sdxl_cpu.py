"""SDXL-base topology (BASELINE config 4) on the CPU with b200sd.ops emulated: per-level transformer depth, d_head 64,
Linear proj_in/out, label_emb vector conditioning (per-sample time-embedding rows), two-tower conditioner — against the
oracle's restatement of sgm's UNetModel / GeneralConditioner at reduced width (same topology)."""
import pytest
import torch

import ops_emulator


@pytest.fixture()
def env(monkeypatch):
    from b200sd import config as C, engine as E, ops, synth
    from oracle import sd_oracle as O
    ops_emulator.install(monkeypatch, ops)
    monkeypatch.setattr(E.SDEngine, "_require_cuda", False)
    cfgs = (C.TINYXL_UNET, C.TINYXL_VAE, C.TINYXL_CLIP)
    ocfgs = (O.TINYXL_UNET, O.TINYXL_VAE, O.TINYXL_CLIP)
    sd = synth.make_state_dict(*cfgs, seed=0)
    eng = E.SDEngine(sd, *cfgs, device="cpu", dtype=torch.float32, use_graphs=False, vae_chunk=2)
    return C, E, O, cfgs, ocfgs, sd, eng


def test_sdxl_configs_and_parameter_count():
    """SDXL-base UNet 2 567.5 M parameters; topology equal in product and oracle"""
    from b200sd import config as C
    from oracle import sd_oracle as O
    assert C.unet_layout(C.SDXL_UNET) == O.unet_layout(O.SDXL_UNET)
    inputs, middle, outputs = C.unet_layout(C.SDXL_UNET)
    cfg = C.SDXL_UNET
    ted = cfg.time_embed_dim
    n = (320 * ted + ted) + (ted * ted + ted) + (cfg.adm_in_channels * ted + ted) + (ted * ted + ted)

    def res(cin, cout):
        return 2 * cin + cin * cout * 9 + cout + ted * cout + cout + 2 * cout + cout * cout * 9 + cout + \
            ((cin * cout + cout) if cin != cout else 0)

    def attn(c, depth):
        blk = 6 * c + (3 * c * c + c * c + c) + (c * c + 2 * cfg.context_dim * c + c * c + c) + (c * 8 * c + 8 * c) + (4 * c * c + c)
        return 2 * c + 2 * (c * c + c) + depth * blk

    for blk in inputs + [middle] + outputs:
        for layer in blk:
            if layer[0] == "conv_in":
                n += layer[1] * layer[2] * 9 + layer[2]
            elif layer[0] == "res":
                n += res(layer[1], layer[2])
            elif layer[0] == "attn":
                n += attn(layer[1], layer[2])
            else:
                n += layer[1] * layer[1] * 9 + layer[1]
    n += 2 * 320 + 320 * 4 * 9 + 4
    assert abs(n / 1e6 - 2567.5) < 1.0, n / 1e6


def _conds(O, sd, ocfgs, tok, neg, px):
    ctx_c, y_c = O.sdxl_conditioner(sd, ocfgs[2], tok, px, px)
    ctx_u, y_u = O.sdxl_conditioner(sd, ocfgs[2], neg, px, px, zero_txt=True)
    return ctx_c, y_c, ctx_u, y_u


def test_sdxl_conditioner_matches_oracle(env):
    C, E, O, cfgs, ocfgs, sd, eng = env
    tok, neg = O.random_prompt_tokens(2, vocab_hi=997), O.empty_prompt_tokens(2, vocab_hi=997)
    ctx_c, y_c, ctx_u, y_u = _conds(O, sd, ocfgs, tok, neg, 96)
    c, u = eng._conds(tok, neg, 96, 96)
    assert c.ctx.shape == (2, 77, 128) and c.y.shape == (2, cfgs[0].adm_in_channels)
    assert torch.allclose(c.ctx, ctx_c, atol=1e-4, rtol=1e-4) and torch.allclose(c.y, y_c, atol=1e-4, rtol=1e-4)
    assert float(u.ctx.abs().max()) == 0.0 and torch.allclose(u.y, y_u, atol=1e-5) and float(u.y[:, :64].abs().max()) == 0.0
    # a non-empty negative prompt is encoded, not zeroed
    c2, u2 = eng._conds(tok, tok, 96, 96)
    assert torch.equal(c2.ctx, u2.ctx)


@pytest.mark.parametrize("b,hw", [(1, 8), (2, 16)])
def test_sdxl_unet_program_matches_oracle(env, b, hw):
    C, E, O, cfgs, ocfgs, sd, eng = env
    from b200sd import ops
    tok, neg = O.random_prompt_tokens(b, vocab_hi=997), O.empty_prompt_tokens(b, vocab_hi=997)
    ctx_c, y_c, ctx_u, y_u = _conds(O, sd, ocfgs, tok, neg, hw * 8)
    x = O.per_image_noise(1000, b, (4, hw, hw))
    with torch.no_grad():
        ref = O.unet_forward(sd, ocfgs[0], torch.cat([x, x]), torch.full((2 * b,), 651.0), torch.cat([ctx_c, ctx_u]),
                             y=torch.cat([y_c, y_u]))
    plan = eng.plan(b, hw, hw)
    assert plan.unet.per_sample and plan.unet.cur_bias.numel() == 2 * b * eng.unet_w.emb_total
    plan.unet.set_context(torch.cat([ctx_c, ctx_u]))
    plan.table[:1].copy_(eng.temb.table(torch.tensor([651.0]), torch.cat([y_c, y_u])))
    plan.x.copy_(x.permute(0, 2, 3, 1).reshape(b, hw * hw, 4))
    ops.pack_unet_input(plan.x, plan.unet.xin, 1.0)
    ops.select_step(plan.table, plan.step, plan.unet.cur_bias)
    plan.unet.run()
    got = plan.unet.eps[..., :4].reshape(2 * b, hw, hw, 4).permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= 2e-4 * float(ref.abs().max())


def test_sdxl_txt2img_euler_a_matches_oracle(env):
    """config C4's sampler (Euler a) end to end: conditioner, per-sample embedding table over all steps, sampler, VAE
    decode with the SDXL scale factor"""
    C, E, O, cfgs, ocfgs, sd, eng = env
    b, hw, steps = 2, 8, 5
    tok, neg = O.random_prompt_tokens(b, vocab_hi=997), O.empty_prompt_tokens(b, vocab_hi=997)
    ctx_c, y_c, ctx_u, y_u = _conds(O, sd, ocfgs, tok, neg, hw * 8)
    y = torch.cat([y_c, y_u])
    unet = lambda x, t, c: O.unet_forward(sd, ocfgs[0], x, t, c, y=y)  # noqa: E731
    nz = E.per_image_noise(77, b, (4, hw, hw), 1 + steps)
    with torch.no_grad():
        z = O.run_sampler("Euler a", unet, ctx_c, ctx_u, 7.0, steps, nz[0], list(nz[1:]))
        ref_u8 = O.to_uint8(O.vae_decode(sd, ocfgs[1], z / ocfgs[1].scale_factor))
    got = eng.txt2img(tok, neg, seed=77, steps=steps, cfg_scale=7.0, height=hw * 8, width=hw * 8, sampler="Euler a")
    lat = eng.plan(b, hw, hw).x.reshape(b, hw, hw, 4).permute(0, 3, 1, 2)
    assert float((lat - z).abs().max()) <= 1e-3 * float(z.abs().max())
    d = (got.int() - ref_u8.int()).abs()
    assert float((d <= 1).float().mean()) == 1.0 and float((d == 0).float().mean()) > 0.99
