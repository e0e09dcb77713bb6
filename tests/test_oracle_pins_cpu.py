"""Independent pins of the numeric oracle (the reference holds no arithmetic, so the oracle is pinned wherever an
implementation that is NOT ours exists in this image).

* CLIP text tower: `oracle.sd_oracle.clip_text_encode` vs `transformers.CLIPTextModel` (the very class ldm's
  FrozenCLIPEmbedder wraps as `cond_stage_model.transformer`) on the same random weights and tokens.
* Schedules: the alphas / sigmas tables against their closed forms, DDIM / DPM++ 2M coefficient identities.
The UNet and VAE restatements have no such counterpart offline (diffusers / ldm / k-diffusion are not installed): they stay
"parity unpinned" (DESIGN.md §2).
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
