"""Independent pins of the numeric oracle (the reference holds no arithmetic, so the oracle is pinned wherever an
implementation that is NOT ours exists in this image).

* CLIP text tower: `oracle.sd_oracle.clip_text_encode` vs `transformers.CLIPTextModel` (the very class ldm's
  FrozenCLIPEmbedder wraps as `cond_stage_model.transformer`) on the same random weights and tokens.
* Schedules: the alphas / sigmas tables against their closed forms, DDIM / DPM++ 2M coefficient identities.
* Round 2: the VAE decoder / encoder vs the `Decoder` / `Encoder` of the FLUX autoencoder shipped in torchtitan (the ldm
  autoencoder written by a third party, ldm's key names, strict load); the UNet's attention vs torch.nn.MultiheadAttention, its ResBlock's main path vs FLUX's ResnetBlock;
  the timestep embedding vs FLUX's.
The UNet as a whole and the sampler update rules have no such counterpart offline (diffusers / ldm / k-diffusion are not
installed): they stay "parity unpinned" (DESIGN.md §2).
"""
import math

import pytest
import torch

from oracle import sd_oracle as O


@pytest.mark.parametrize("cfg", [O.TINY_CLIP, O.CLIPConfig(vocab=2000, width=128, layers=3, heads=4)])
def test_clip_text_matches_transformers(cfg):
    transformers = pytest.importorskip("transformers")
    hf_cfg = transformers.CLIPTextConfig(vocab_size=cfg.vocab + 3, hidden_size=cfg.width, intermediate_size=4 * cfg.width,
                                         num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                                         max_position_embeddings=cfg.ctx, hidden_act="quick_gelu", layer_norm_eps=1e-5,
                                         eos_token_id=cfg.vocab + 2, bos_token_id=cfg.vocab + 1, pad_token_id=cfg.vocab + 2)
    torch.manual_seed(7)
    model = transformers.CLIPTextModel(hf_cfg).eval().float()
    with torch.no_grad():   # make every tensor non-trivial (HF initialises biases / LayerNorm to 0 / 1)
        for p in model.parameters():
            p.add_(0.05 * torch.randn_like(p))
    sd = {"cond_stage_model.transformer." + k: v.detach().clone() for k, v in model.state_dict().items()}
    tokens = O.random_prompt_tokens(3, seed=11, vocab_hi=cfg.vocab)
    with torch.no_grad():
        ref = model(input_ids=tokens).last_hidden_state
        got = O.clip_text_encode(sd, cfg, tokens)
    assert got.shape == ref.shape == (3, 77, cfg.width)
    assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


def test_schedule_tables_closed_forms():
    ac = O.alphas_cumprod().double()
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2   # ldm "linear" schedule
    assert torch.allclose(ac, torch.cumprod(1 - betas, 0), rtol=1e-6, atol=0)
    sig, log_sig = O.karras_sigmas_compvis(20)
    assert len(sig) == 21 and float(sig[-1]) == 0.0 and abs(float(sig[0]) - float(((1 - ac[-1]) / ac[-1]) ** 0.5)) < 1e-9
    ks, _ = O.sigmas_karras(20)
    assert abs(float(ks[0]) - float(sig[0])) < 1e-9 and abs(float(ks[-2]) - float(((1 - ac[0]) / ac[0]) ** 0.5)) < 1e-9
    assert all(float(ks[i]) > float(ks[i + 1]) for i in range(20))
    # sigma_to_t inverts the table at the integer timesteps
    for t in (0, 17, 500, 998):
        assert abs(O.sigma_to_t(math.exp(float(log_sig[t])), log_sig) - t) < 1e-6


def test_samplers_on_analytic_denoisers():
    """closed-form checks of the sampler restatements with synthetic eps models (the conditioning is unused):
    eps == 0 -> the x0 prediction is x itself, every sampler must leave x at x_T * sigma_0;
    eps == x_in / (c_in * sigma) -> the x0 prediction is 0, DPM++ 2M / Euler contract x by sigma_{i+1} / sigma_i per step and
    end exactly at 0."""
    c = torch.zeros(1, 1, 1)
    xT = torch.randn(2, 4, 8, 8)
    zero = lambda x, t, ctx: torch.zeros_like(x)  # noqa: E731
    for karras in (True, False):
        sig, _ = O.sigmas_karras(6) if karras else O.karras_sigmas_compvis(6)
        got = O.sample_dpmpp_2m(zero, xT, c.expand(2, 1, 1), c.expand(2, 1, 1), 6, 3.0, karras=karras)
        assert torch.allclose(got, xT * float(sig[0]), rtol=1e-5, atol=1e-6)
    sig, log_sig = O.karras_sigmas_compvis(6)
    assert torch.allclose(O.sample_euler(zero, xT, c.expand(2, 1, 1), c.expand(2, 1, 1), 6, 3.0), xT * float(sig[0]), rtol=1e-5)

    def to_zero(x, t, ctx):   # x arrives scaled by c_in = 1/sqrt(sigma^2 + 1); t -> sigma through the model's table
        tt = float(t[0])
        lo = int(math.floor(tt))
        w = tt - lo
        s = math.exp((1 - w) * float(log_sig[lo]) + w * float(log_sig[min(lo + 1, 999)]))
        return x * math.sqrt(s * s + 1.0) / s

    assert float(O.sample_dpmpp_2m(to_zero, xT, c.expand(2, 1, 1), c.expand(2, 1, 1), 6, 3.0, karras=True).abs().max()) < 1e-4
    assert float(O.sample_euler(to_zero, xT, c.expand(2, 1, 1), c.expand(2, 1, 1), 6, 3.0).abs().max()) < 1e-4


# ------------------------------------------------------------------------------------------------ round 2: more pins
def _flux_ae():
    """torchtitan ships Black Forest Labs' FLUX autoencoder (torchtitan/experiments/flux/model/autoencoder.py): the ldm /
    taming `Encoder` / `Decoder` (ResnetBlock, single-head AttnBlock, asymmetric-pad Downsample, nearest Upsample) written
    by someone else, with ldm's own module names — i.e. ldm's state-dict keys."""
    return pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")


@pytest.mark.parametrize("cfg,hw", [(O.TINY_VAE, 8), (O.VAEConfig(ch=64, ch_mult=(1, 2, 4, 4), num_res_blocks=2), 4)])
def test_vae_decoder_matches_independent_implementation(cfg, hw):
    """oracle.vae_decode vs the FLUX / ldm `Decoder` class: strict state-dict load (every key name and shape of
    `first_stage_model.decoder.*` is the third party's), same input, same output"""
    A = _flux_ae()
    from b200sd import synth
    sd = synth.make_vae_state_dict(cfg, seed=3) if hasattr(synth, "make_vae_state_dict") else \
        {k: v for k, v in synth.make_state_dict(O.TINY_UNET, cfg, O.TINY_CLIP, seed=3).items() if k.startswith("first_stage_model.")}
    dec = A.Decoder(ch=cfg.ch, out_ch=3, ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks, in_channels=3,
                    resolution=hw * 2 ** (len(cfg.ch_mult) - 1), z_channels=cfg.z_channels).eval().float()
    weights = {k[len("first_stage_model.decoder."):]: v.float() for k, v in sd.items()
               if k.startswith("first_stage_model.decoder.")}
    dec.load_state_dict(weights, strict=True)
    g = torch.Generator().manual_seed(5)
    z = torch.randn((2, cfg.z_channels, hw, hw), generator=g)
    with torch.no_grad():
        got = O.vae_decode({k: v.float() for k, v in sd.items()}, cfg, z)
        pq = torch.nn.functional.conv2d(z, sd["first_stage_model.post_quant_conv.weight"].float(),
                                        sd["first_stage_model.post_quant_conv.bias"].float())
        ref = dec(pq)
    assert got.shape == ref.shape == (2, 3, hw * 2 ** (len(cfg.ch_mult) - 1), hw * 2 ** (len(cfg.ch_mult) - 1))
    assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("cfg,px", [(O.TINY_VAE, 64), (O.VAEConfig(ch=64, ch_mult=(1, 2, 4, 4), num_res_blocks=2), 32)])
def test_vae_encoder_matches_independent_implementation(cfg, px):
    """oracle.vae_encode_mean vs the FLUX / ldm `Encoder` class + quant_conv + the mean half of the moments"""
    A = _flux_ae()
    from b200sd import synth
    sd = {k: v.float() for k, v in synth.make_state_dict(O.TINY_UNET, cfg, O.TINY_CLIP, seed=4).items()
          if k.startswith("first_stage_model.")}
    enc = A.Encoder(resolution=px, in_channels=3, ch=cfg.ch, ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks,
                    z_channels=cfg.z_channels).eval().float()
    enc.load_state_dict({k[len("first_stage_model.encoder."):]: v for k, v in sd.items()
                         if k.startswith("first_stage_model.encoder.")}, strict=True)
    g = torch.Generator().manual_seed(6)
    x = torch.rand((2, 3, px, px), generator=g) * 2 - 1
    with torch.no_grad():
        got = O.vae_encode_mean(sd, cfg, x)
        moments = torch.nn.functional.conv2d(enc(x), sd["first_stage_model.quant_conv.weight"],
                                             sd["first_stage_model.quant_conv.bias"])
        ref = moments.chunk(2, dim=1)[0]
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


def test_timestep_embedding_matches_independent_implementation():
    """ldm's timestep_embedding(t, dim) == FLUX's timestep_embedding(t, dim, time_factor=1) (cos | sin halves, max_period
    10000; FLUX feeds t in [0, 1] and scales by 1000 inside)"""
    L = pytest.importorskip("torchtitan.experiments.flux.model.layers")
    t = torch.tensor([1.0, 51.0, 500.5, 999.0])
    for dim in (320, 64):
        ref = L.timestep_embedding(t, dim, max_period=10000, time_factor=1.0)
        assert torch.allclose(O.timestep_embedding(t, dim), ref.float(), atol=2e-5, rtol=0)


@pytest.mark.parametrize("ctx_dim,n_ctx", [(None, None), (96, 77)])
def test_unet_attention_matches_torch_multihead_attention(ctx_dim, n_ctx):
    """ldm CrossAttention (to_q / to_k / to_v without bias, heads as contiguous channel chunks, scale d^-0.5, to_out with
    bias) is torch.nn.MultiheadAttention with separate projection weights — an implementation that is not ours"""
    c, heads, b, n = 64, 4, 2, 50
    g = torch.Generator().manual_seed(8)
    kd = c if ctx_dim is None else ctx_dim
    sd = {"a.to_q.weight": torch.randn((c, c), generator=g) * 0.2, "a.to_k.weight": torch.randn((c, kd), generator=g) * 0.2,
          "a.to_v.weight": torch.randn((c, kd), generator=g) * 0.2, "a.to_out.0.weight": torch.randn((c, c), generator=g) * 0.2,
          "a.to_out.0.bias": torch.randn((c,), generator=g) * 0.2}
    x = torch.randn((b, n, c), generator=g)
    ctx = None if ctx_dim is None else torch.randn((b, n_ctx, ctx_dim), generator=g)
    mha = torch.nn.MultiheadAttention(c, heads, bias=True, batch_first=True, kdim=kd, vdim=kd).eval()
    with torch.no_grad():
        if ctx_dim is None:   # same embed dims: one packed in-projection
            mha.in_proj_weight.copy_(torch.cat([sd["a.to_q.weight"], sd["a.to_k.weight"], sd["a.to_v.weight"]]))
        else:
            mha.q_proj_weight.copy_(sd["a.to_q.weight"])
            mha.k_proj_weight.copy_(sd["a.to_k.weight"])
            mha.v_proj_weight.copy_(sd["a.to_v.weight"])
        mha.in_proj_bias.zero_()
        mha.out_proj.weight.copy_(sd["a.to_out.0.weight"])
        mha.out_proj.bias.copy_(sd["a.to_out.0.bias"])
        kv = x if ctx is None else ctx
        ref, _ = mha(x, kv, kv, need_weights=False)
        for sdpa in (False, True):
            old, O.USE_SDPA = O.USE_SDPA, sdpa
            try:
                got = O.cross_attention(sd, "a", x, ctx, heads)
            finally:
                O.USE_SDPA = old
            assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("cin,cout", [(64, 64), (64, 128)])
def test_unet_resblock_main_path_matches_independent_resnet_block(cin, cout):
    """openaimodel.ResBlock with a silent embedding path (emb_layers weight and bias zero) is the autoencoder's ResnetBlock
    at eps 1e-5: GN32 - SiLU - conv3x3, GN32 - SiLU - conv3x3, 1x1 skip when the width changes.  The embedding path itself
    (h + Linear(SiLU(emb)) broadcast over pixels, before the second GroupNorm) is checked as a shift of the first conv's
    bias."""
    A = _flux_ae()
    g = torch.Generator().manual_seed(9)
    blk = A.ResnetBlock(in_channels=cin, out_channels=cout).eval().float()
    blk.norm1.eps = blk.norm2.eps = 1e-5
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    w = blk.state_dict()
    names = {"norm1": "in_layers.0", "conv1": "in_layers.2", "norm2": "out_layers.0", "conv2": "out_layers.3",
             "nin_shortcut": "skip_connection"}
    sd = {"r." + names[k.split(".")[0]] + "." + k.split(".")[1]: v.clone() for k, v in w.items()}
    emb_dim = 32
    sd["r.emb_layers.1.weight"] = torch.zeros((cout, emb_dim))
    sd["r.emb_layers.1.bias"] = torch.zeros((cout,))
    x = torch.randn((2, cin, 8, 8), generator=g)
    emb = torch.randn((2, emb_dim), generator=g)
    with torch.no_grad():
        assert float((O.res_block(sd, "r", x, emb) - blk(x)).abs().max()) <= 2e-5
        # a per-channel embedding contribution equals the same shift of conv1's bias (same for every sample here)
        shift = torch.randn((cout,), generator=g) * 0.3
        sd["r.emb_layers.1.bias"] = shift.clone()
        blk.conv1.bias.add_(shift)
        assert float((O.res_block(sd, "r", x, emb) - blk(x)).abs().max()) <= 2e-5


def _hf_clip(transformers, vocab, width, layers, heads, act, proj=None):
    cfg = transformers.CLIPTextConfig(vocab_size=vocab + 3, hidden_size=width, intermediate_size=4 * width,
                                      num_hidden_layers=layers, num_attention_heads=heads, max_position_embeddings=77,
                                      hidden_act=act, layer_norm_eps=1e-5, eos_token_id=vocab + 2, bos_token_id=vocab + 1,
                                      pad_token_id=vocab + 2, **({} if proj is None else {"projection_dim": proj}))
    torch.manual_seed(13)
    model = (transformers.CLIPTextModel if proj is None else transformers.CLIPTextModelWithProjection)(cfg).eval().float()
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn_like(p))
    return model


def test_sdxl_clip_l_hidden_layer_matches_transformers():
    """sgm FrozenCLIPEmbedder(layer="hidden", layer_idx=11): hidden_states[11] of transformers.CLIPTextModel"""
    transformers = pytest.importorskip("transformers")
    cfg = O.CLIPConfig(vocab=1500, width=96, layers=4, heads=4)
    model = _hf_clip(transformers, cfg.vocab, cfg.width, cfg.layers, cfg.heads, "quick_gelu")
    sd = {"t." + k: v.detach().clone() for k, v in model.state_dict().items()}
    tokens = O.random_prompt_tokens(3, seed=12, vocab_hi=cfg.vocab)
    with torch.no_grad():
        hs = model(input_ids=tokens, output_hidden_states=True).hidden_states
        for idx in (cfg.layers - 1, 1):
            got = O.clip_text_hidden(sd, cfg, tokens, idx, "t.text_model.")
            assert float((got - hs[idx]).abs().max()) <= 2e-5 * max(1.0, float(hs[idx].abs().max()))


def test_sdxl_open_clip_tower_matches_transformers_with_projection():
    """sgm FrozenOpenCLIPEmbedder2 (open_clip text transformer: packed in_proj, GELU, penultimate layer, pooled =
    ln_final(last)[EOS] @ text_projection) is transformers.CLIPTextModelWithProjection(hidden_act="gelu") under the
    open_clip -> HF key mapping every SDXL checkpoint converter uses"""
    transformers = pytest.importorskip("transformers")
    cfg = O.CLIPConfig(vocab=1500, width=64, layers=2, heads=2, xl_width=96, xl_layers=4, xl_heads=4, xl_proj=80)
    model = _hf_clip(transformers, cfg.vocab, cfg.xl_width, cfg.xl_layers, cfg.xl_heads, "gelu", proj=cfg.xl_proj)
    hf = {k: v.detach().clone() for k, v in model.state_dict().items()}
    sd = {"m.token_embedding.weight": hf["text_model.embeddings.token_embedding.weight"],
          "m.positional_embedding": hf["text_model.embeddings.position_embedding.weight"],
          "m.ln_final.weight": hf["text_model.final_layer_norm.weight"], "m.ln_final.bias": hf["text_model.final_layer_norm.bias"],
          "m.text_projection": hf["text_projection.weight"].t().contiguous()}
    for i in range(cfg.xl_layers):
        h, o = f"text_model.encoder.layers.{i}.", f"m.transformer.resblocks.{i}."
        sd[o + "attn.in_proj_weight"] = torch.cat([hf[h + f"self_attn.{n}_proj.weight"] for n in "qkv"])
        sd[o + "attn.in_proj_bias"] = torch.cat([hf[h + f"self_attn.{n}_proj.bias"] for n in "qkv"])
        for a, b_ in (("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"),
                      ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
            sd[o + a + ".weight"], sd[o + a + ".bias"] = hf[h + b_ + ".weight"], hf[h + b_ + ".bias"]
    tokens = O.random_prompt_tokens(3, seed=14, vocab_hi=cfg.vocab)
    with torch.no_grad():
        out = model(input_ids=tokens, output_hidden_states=True)
        pen, pooled = O.open_clip_text(sd, cfg, tokens, "m.")
    assert float((pen - out.hidden_states[-2]).abs().max()) <= 2e-5 * max(1.0, float(out.hidden_states[-2].abs().max()))
    assert float((pooled - out.text_embeds).abs().max()) <= 2e-5 * max(1.0, float(out.text_embeds.abs().max()))


def test_unet_transformer_block_wiring_matches_torch_decoder_layer():
    """BasicTransformerBlock = a pre-LN decoder layer: x + self_attn(LN1 x), x + cross_attn(LN2 x, context), x + FF(LN3 x).
    torch.nn.TransformerDecoderLayer(norm_first=True) supplies the wiring, the LayerNorms and both attentions (cross
    attention re-made with kdim = context width); the feed-forward is its linear1 -> activation -> linear2 with GEGLU
    (value half times GELU of the gate half) as the activation."""
    c, heads, ctx_dim, b, n, n_ctx = 64, 4, 96, 2, 40, 77
    torch.manual_seed(21)
    layer = torch.nn.TransformerDecoderLayer(c, heads, dim_feedforward=8 * c, dropout=0.0, batch_first=True, norm_first=True,
                                             activation=lambda h: h.chunk(2, dim=-1)[0] * torch.nn.functional.gelu(h.chunk(2, dim=-1)[1]))
    layer.multihead_attn = torch.nn.MultiheadAttention(c, heads, batch_first=True, kdim=ctx_dim, vdim=ctx_dim)
    layer.linear2 = torch.nn.Linear(4 * c, c)
    layer = layer.eval().float()
    with torch.no_grad():
        for p in layer.parameters():
            p.copy_(torch.randn_like(p) * 0.15)
        layer.self_attn.in_proj_bias.zero_()
        layer.multihead_attn.in_proj_bias.zero_()
    q, k, v = layer.self_attn.in_proj_weight.detach().chunk(3)
    sd = {"b.attn1.to_q.weight": q, "b.attn1.to_k.weight": k, "b.attn1.to_v.weight": v,
          "b.attn1.to_out.0.weight": layer.self_attn.out_proj.weight.detach(), "b.attn1.to_out.0.bias": layer.self_attn.out_proj.bias.detach(),
          "b.attn2.to_q.weight": layer.multihead_attn.q_proj_weight.detach(), "b.attn2.to_k.weight": layer.multihead_attn.k_proj_weight.detach(),
          "b.attn2.to_v.weight": layer.multihead_attn.v_proj_weight.detach(),
          "b.attn2.to_out.0.weight": layer.multihead_attn.out_proj.weight.detach(), "b.attn2.to_out.0.bias": layer.multihead_attn.out_proj.bias.detach(),
          "b.ff.net.0.proj.weight": layer.linear1.weight.detach(), "b.ff.net.0.proj.bias": layer.linear1.bias.detach(),
          "b.ff.net.2.weight": layer.linear2.weight.detach(), "b.ff.net.2.bias": layer.linear2.bias.detach()}
    for i in (1, 2, 3):
        sd[f"b.norm{i}.weight"], sd[f"b.norm{i}.bias"] = getattr(layer, f"norm{i}").weight.detach(), getattr(layer, f"norm{i}").bias.detach()
    x, ctx = torch.randn(b, n, c), torch.randn(b, n_ctx, ctx_dim)
    with torch.no_grad():
        ref = layer(x, ctx)
        got = O.transformer_block(sd, "b", x, ctx, heads)
    assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


# ------------------------------------------------------------------------------------------------ samplers vs diffusion theory
# For Gaussian data x ~ N(0, s^2) the optimal denoiser is known in closed form: eps(x_t, t) = x_t sqrt(1 - a) / (a s^2 + 1 - a)
# (a = alphas_cumprod at t), i.e. denoised(x, sigma) = x s^2 / (s^2 + sigma^2) in k-diffusion's variables.  The probability-flow
# ODE then has the exact solution x(sigma) = x(sigma_0) sqrt((s^2 + sigma^2) / (s^2 + sigma_0^2)), and every stochastic sampler
# must end on samples of the data distribution (variance s^2).  No third-party code needed: a wrong coefficient anywhere in a
# sampler restatement shows up as a wrong limit or a wrong convergence order.
_S_DATA = 0.8


def _gaussian_eps_model(x, t, ctx):
    _, log_sig = O.model_sigmas()
    tt = t.double().clamp(0, 999)
    lo = tt.floor().long().clamp(max=998)
    w = tt - lo
    sg = ((1 - w) * log_sig[lo] + w * log_sig[lo + 1]).exp()
    a = 1 / (1 + sg * sg)
    k = (torch.sqrt(1 - a) / (a * _S_DATA ** 2 + 1 - a)).float()
    return x * k[:, None, None, None]


def _flow_error(name, steps, noise0):
    c = torch.zeros(noise0.shape[0], 1, 1)
    x = O.run_sampler(name, _gaussian_eps_model, c, c, 1.0, steps, noise0, None)
    s = _S_DATA
    if name == "PLMS":   # a timestep sampler: it stops at alphas_cumprod[0], not at sigma = 0
        ac, ts = O.alphas_cumprod().double(), O.ddim_timesteps(steps)
        a_T, a_0 = float(ac[ts[-1]]), float(ac[0])
        exact = noise0 * math.sqrt((a_0 * s * s + 1 - a_0) / (a_T * s * s + 1 - a_T))
    else:
        karras = name.endswith("Karras") or name == "DPM++ 2M"
        s0 = float((O.sigmas_karras(steps) if karras else O.karras_sigmas_compvis(steps))[0][0])
        exact = noise0 * s0 * s / math.sqrt(s * s + s0 * s0)
    return float((x - exact).norm() / exact.norm())


@pytest.mark.parametrize("name,err40,order", [
    ("Euler", 0.06, 1), ("Heun", 0.003, 2), ("DPM2", 0.012, 2), ("DPM2 Karras", 0.003, 2), ("DPM++ 2M", 0.006, 2),
    ("DPM++ 2M Karras", 0.006, 2), ("LMS", 5e-4, 3), ("LMS Karras", 2e-3, 3), ("PLMS", 0.008, 1.2)])
def test_deterministic_samplers_solve_the_probability_flow_ode(name, err40, order):
    """error against the exact solution: small at 40 steps and falling with the sampler's order from 20 to 40 steps"""
    noise0 = torch.randn((2, 4, 16, 16), generator=torch.Generator().manual_seed(0))
    e20, e40 = _flow_error(name, 20, noise0), _flow_error(name, 40, noise0)
    assert e40 <= err40, (name, e20, e40)
    assert e20 / e40 >= 0.8 * 2 ** order, (name, e20, e40)    # halving the step divides the error by ~2^order


@pytest.mark.parametrize("name,tol", [("DPM fast", 0.01), ("DPM adaptive", 0.05)])
def test_dpm_solver_samplers_solve_the_probability_flow_ode(name, tol):
    """DPM fast (fixed n evaluations) and DPM adaptive (rtol 0.05 step control: the error is the tolerance's, whatever
    `steps` says)"""
    noise0 = torch.randn((2, 4, 16, 16), generator=torch.Generator().manual_seed(0))
    assert _flow_error(name, 20, noise0) <= tol


@pytest.mark.parametrize("name,lo,hi", [("Euler a", 0.80, 1.02), ("DPM2 a", 0.93, 1.10), ("DPM++ 2S a", 0.92, 1.06),
                                        ("DPM++ SDE", 0.90, 1.06), ("DPM2 a Karras", 0.93, 1.07),
                                        ("DPM++ 2S a Karras", 0.93, 1.06), ("DPM++ SDE Karras", 0.93, 1.07)])
def test_stochastic_samplers_end_on_the_data_distribution(name, lo, hi):
    """40 steps of an ancestral / SDE sampler from pure noise: zero mean and the data's variance (Euler a, first order,
    approaches it from below: 0.73 at 20 steps, 0.86 at 40)"""
    g = torch.Generator().manual_seed(1)
    shape = (4, 4, 32, 32)
    noise0 = torch.randn(shape, generator=g)
    draws = [torch.randn(shape, generator=g) for _ in range(160)]
    c = torch.zeros(shape[0], 1, 1)
    x = O.run_sampler(name, _gaussian_eps_model, c, c, 1.0, 40, noise0, draws)
    ratio = float(x.var()) / _S_DATA ** 2
    assert lo <= ratio <= hi and abs(float(x.mean())) <= 0.03, (name, ratio, float(x.mean()))


def test_ddim_solves_the_probability_flow_ode():
    """sdwui's DDIM (eta 0) on the same Gaussian-data denoiser: first order, ends at alphas_cumprod[timesteps[0]]"""
    noise0 = torch.randn((2, 4, 16, 16), generator=torch.Generator().manual_seed(0))
    c = torch.zeros(2, 1, 1)
    errs = []
    for steps in (20, 40):
        x = O.sample_ddim(_gaussian_eps_model, noise0, c, c, steps, 1.0)
        ac, ts = O.alphas_cumprod().double(), O.ddim_timesteps(steps)
        a_T, a_0, s = float(ac[ts[-1]]), float(ac[ts[0]]), _S_DATA
        exact = noise0 * math.sqrt((a_0 * s * s + 1 - a_0) / (a_T * s * s + 1 - a_T))
        errs.append(float((x - exact).norm() / exact.norm()))
    assert errs[1] <= 0.05 and errs[0] / errs[1] >= 1.6, errs


# ------------------------------------------------------------------------------------------------ known-answer: parameter counts
def _shape_only_state_dict(cfgs):
    """b200sd.synth.make_state_dict with every random draw replaced by a meta tensor: shapes and key names, no storage"""
    from unittest import mock
    from b200sd import synth
    real = torch.randn

    def meta_randn(*shape, **kw):
        shape = shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape
        return torch.empty(tuple(shape), device="meta")
    with mock.patch.object(torch, "randn", meta_randn):
        sd = synth.make_state_dict(*cfgs, seed=0)
    assert torch.randn is real
    return sd


def _count(sd, prefix):
    return sum(v.numel() for k, v in sd.items() if k.startswith(prefix))


def test_parameter_counts_equal_the_published_models():
    """Known answers the restatements cannot have been tuned to: the parameter totals of the released models.  The state
    dict both the product and the oracle consume (same keys, same shapes; see test_*_program_matches_oracle) has, for the
    full-size configurations, exactly
      SD1.5  UNet 859 520 964 - kl-f8 VAE 83 653 863 - CLIP ViT-L/14 text model 123 060 480
      SDXL   UNet 2 567 463 684 - OpenCLIP ViT-bigG/14 text tower 694 659 840 (+ the logit_scale scalar upstream keeps)
    parameters.  Every block, its input width after the skip concatenations, every attention depth and head projection
    enters these sums: a missing or extra layer, or a wrong channel count anywhere, changes them."""
    from b200sd import config as C
    sd = _shape_only_state_dict((C.SD15_UNET, C.SD15_VAE, C.SD15_CLIP))
    assert _count(sd, C.UNET_PREFIX) == 859_520_964
    assert _count(sd, C.VAE_PREFIX) == 83_653_863
    assert _count(sd, C.CLIP_PREFIX) == 123_060_480
    assert len(sd) == sum(1 for k in sd if k.startswith((C.UNET_PREFIX, C.VAE_PREFIX, C.CLIP_PREFIX)))
    xl = _shape_only_state_dict((C.SDXL_UNET, C.SDXL_VAE, C.SDXL_CLIP))
    assert _count(xl, C.UNET_PREFIX) == 2_567_463_684
    assert _count(xl, C.VAE_PREFIX) == 83_653_863
    assert _count(xl, C.XL_PREFIX0) == 123_060_480
    assert _count(xl, C.XL_PREFIX1) == 694_659_840


def test_schedule_known_constants():
    """numbers every SD1.x user has seen: k-diffusion reports sigma_min 0.0292 / sigma_max 14.6146 for this schedule, the
    last alphas_cumprod is 0.00466, and sdwui's 20-step DDIM visits 1, 51, ..., 951"""
    sig, _ = O.model_sigmas()
    assert abs(float(sig[0]) - 0.0291675) < 2e-6 and abs(float(sig[-1]) - 14.614642) < 2e-5
    ac = O.alphas_cumprod()
    assert abs(float(ac[-1]) - 0.0046600) < 2e-6 and abs(float(ac[0]) - 0.99915) < 1e-6
    assert O.ddim_timesteps(20).tolist() == list(range(1, 1000, 50))
    assert O.SD15_VAE.scale_factor == 0.18215 and O.SDXL_VAE.scale_factor == 0.13025
    assert O.SDXL_UNET.adm_in_channels == 1280 + 6 * 256
