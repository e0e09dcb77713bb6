"""oracle/sd_oracle.py — plain PyTorch fp32 restatement of the numeric path the reference dispatches to.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg — never by the product path (stable-diffusion-webui-distributed_b200/).

PARITY PARTLY PINNED (the UNet's top-level wiring and sdwui's step conventions: UNPINNED): the reference
(papuSpartan/stable-diffusion-webui-distributed @ 8fd65ebd) contains none of this arithmetic and ships no tests or golden
vectors.  Its call sites into the numeric path are
  scripts/spartan/world.py:196   process_images(p)                     (master's share / sample_master)
  scripts/spartan/worker.py:432  session.post(.../sdapi/v1/txt2img|img2img)  (remote sdwui: UNet x steps, VAE)
The arithmetic lives in un-vendored third parties (AUTOMATIC1111 sdwui -> CompVis `ldm` openaimodel / model.py /
attention.py, k-diffusion sampling.py); none is installable offline and the extension pins no version.  This file
restates their published algorithms with ldm state_dict key names so a real checkpoint loads.
Pins against implementations that are NOT ours and exist in this image (tests/test_oracle_pins_cpu.py, 2e-5 on shared
random weights): the CLIP text tower equals `transformers.CLIPTextModel` — the class ldm's FrozenCLIPEmbedder wraps; the
VAE decoder and encoder equal the `Decoder` / `Encoder` classes of Black Forest Labs' FLUX autoencoder as shipped in
torchtitan (torchtitan/experiments/flux/model/autoencoder.py: the ldm / taming autoencoder with ldm's own module names —
loaded with strict=True, so every key name and shape of `first_stage_model.{encoder,decoder}.*` is the third party's);
the UNet's attention equals `torch.nn.MultiheadAttention` with separate projection weights; the ResBlock's main path equals FLUX's
`ResnetBlock` at eps 1e-5; the timestep embedding equals FLUX's; BasicTransformerBlock's wiring equals torch.nn.TransformerDecoderLayer(norm_first=True); both SDXL text towers
equal transformers (CLIP-L hidden_states[11], OpenCLIP bigG as CLIPTextModelWithProjection under the open_clip -> HF key
mapping); the schedule tables equal their closed forms.  Samplers vs diffusion theory: on the closed-form optimal denoiser
of Gaussian data every deterministic sampler (all names of the reference's table + DDIM) converges to the exact
probability-flow solution with its order, and every ancestral / SDE sampler ends on the data distribution.  Known
answers: the full-size state dicts have exactly the released models' parameter totals (SD1.5 UNet 859 520 964, VAE
83 653 863, CLIP-L 123 060 480, SDXL UNet 2 567 463 684, bigG text 694 659 840).  What has no
independent counterpart offline (diffusers, ldm, sgm, k-diffusion are not installed) and stays an unpinned restatement:
the UNet's top-level wiring (block lists, skip concatenations, SpatialTransformer reshapes, SDXL label_emb) and the
conventions theory does not fix (sdwui's step counts, img2img t_enc, which noise draw feeds which step).

Everything here is NCHW fp32 (or whatever dtype/device the caller's tensors have), functional over a dict of
parameters.  Function docstrings name the upstream symbol they follow.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import math

import torch
import torch.nn.functional as F

# "fp32" has to mean fp32 when the oracle runs on a GPU (the -m gpu tests and smoke() put it next to the CUDA path): PyTorch
# lets cuDNN run fp32 convolutions on TF32 tensor cores by default (10-bit mantissas — the precision class of the fp16 path
# under test).  Importing the oracle turns that off for the process; matmuls are IEEE fp32 by default already.
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

SD = Dict[str, torch.Tensor]
# bench.py's stock-PyTorch comparator flips this: attention through F.scaled_dot_product_attention (the fused library
# kernel a stock fp16 pipeline would use) instead of the explicit softmax(q k^T) v the parity checks run
USE_SDPA = False


# ------------------------------------------------------------------------------------------------ configs
@dataclass
class UNetConfig:
    """ldm/modules/diffusionmodules/openaimodel.py::UNetModel ctor args (v1-inference.yaml)."""
    in_channels: int = 4
    out_channels: int = 4
    model_channels: int = 320
    channel_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    attention_levels: Tuple[int, ...] = (0, 1, 2)   # levels (index into channel_mult) that carry transformers
    num_heads: int = 8
    context_dim: int = 768
    transformer_depth: int = 1
    # sgm/modules/diffusionmodules/openaimodel.py::UNetModel (sd_xl_base.yaml): transformer_depth per level,
    # num_head_channels, use_linear_in_transformer, adm_in_channels (label_emb)
    transformer_depths: Optional[Tuple[int, ...]] = None
    middle_depth: Optional[int] = None
    num_head_channels: int = 0
    linear_proj: bool = False
    adm_in_channels: int = 0

    @property
    def time_embed_dim(self):
        return 4 * self.model_channels

    def depth(self, level):
        if self.transformer_depths is not None:
            return self.transformer_depths[level]
        return self.transformer_depth if level in self.attention_levels else 0

    def heads(self, channels):
        return channels // self.num_head_channels if self.num_head_channels else self.num_heads


@dataclass
class VAEConfig:
    """ldm/modules/diffusionmodules/model.py Encoder/Decoder args (kl-f8)."""
    ch: int = 128
    ch_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    z_channels: int = 4
    out_ch: int = 3
    scale_factor: float = 0.18215


@dataclass
class CLIPConfig:
    vocab: int = 49408
    width: int = 768
    layers: int = 12
    heads: int = 12
    ctx: int = 77
    xl_width: int = 0      # SDXL: second text tower (OpenCLIP ViT-bigG), see sdxl_conditioner
    xl_layers: int = 0
    xl_heads: int = 0
    xl_proj: int = 0
    size_embed_dim: int = 256


SD15_UNET = UNetConfig()
SD15_VAE = VAEConfig()
SD15_CLIP = CLIPConfig()
# reduced-width models with the same topology, for fast CPU tests
TINY_UNET = UNetConfig(model_channels=64, num_heads=2, context_dim=64)
TINY_VAE = VAEConfig(ch=64, ch_mult=(1, 2), num_res_blocks=1)
TINY_CLIP = CLIPConfig(vocab=1000, width=64, layers=2, heads=2)
SDXL_UNET = UNetConfig(channel_mult=(1, 2, 4), transformer_depths=(0, 2, 10), middle_depth=10, num_head_channels=64,
                       context_dim=2048, linear_proj=True, adm_in_channels=2816)
SDXL_VAE = VAEConfig(scale_factor=0.13025)
SDXL_CLIP = CLIPConfig(xl_width=1280, xl_layers=32, xl_heads=20, xl_proj=1280)
TINYXL_UNET = UNetConfig(model_channels=64, channel_mult=(1, 2, 4), transformer_depths=(0, 1, 2), middle_depth=2,
                         num_head_channels=64, context_dim=128, linear_proj=True, adm_in_channels=64 + 6 * 32)
TINYXL_VAE = VAEConfig(ch=64, ch_mult=(1, 2), num_res_blocks=1, scale_factor=0.13025)
TINYXL_CLIP = CLIPConfig(vocab=1000, width=64, layers=3, heads=2, xl_width=64, xl_layers=3, xl_heads=2, xl_proj=64,
                         size_embed_dim=32)


# ------------------------------------------------------------------------------------------------ UNet topology
def unet_layout(cfg: UNetConfig):
    """Block list of UNetModel.__init__: returns (input_blocks, middle, output_blocks); each block is a list of
    ('conv_in', cin, cout) | ('res', cin, cout) | ('attn', c, depth) | ('down', c) | ('up', c)."""
    mc = cfg.model_channels
    inputs = [[("conv_in", cfg.in_channels, mc)]]
    chans = [mc]
    ch = mc
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            blk = [("res", ch, mult * mc)]
            ch = mult * mc
            if cfg.depth(level):
                blk.append(("attn", ch, cfg.depth(level)))
            inputs.append(blk)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            inputs.append([("down", ch)])
            chans.append(ch)
    middle = [("res", ch, ch), ("attn", ch, cfg.middle_depth if cfg.middle_depth is not None else cfg.transformer_depth),
              ("res", ch, ch)]
    outputs = []
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            blk = [("res", ch + ich, mult * mc)]
            ch = mult * mc
            if cfg.depth(level):
                blk.append(("attn", ch, cfg.depth(level)))
            if level and i == cfg.num_res_blocks:
                blk.append(("up", ch))
            outputs.append(blk)
    return inputs, middle, outputs


# ------------------------------------------------------------------------------------------------ UNet forward
def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """ldm/modules/diffusionmodules/util.py::timestep_embedding (cos first, then sin)."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(x, sd, key, eps):
    return F.group_norm(x.float(), 32, sd[key + ".weight"].float(), sd[key + ".bias"].float(), eps).to(x.dtype)


def res_block(sd: SD, p: str, x, emb):
    """openaimodel.ResBlock._forward: GN32-SiLU-conv3x3, + Linear(SiLU(emb)), GN32-SiLU-conv3x3, + skip."""
    h = F.silu(_gn(x, sd, p + ".in_layers.0", 1e-5))
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = F.silu(_gn(h, sd, p + ".out_layers.0", 1e-5))
    h = F.conv2d(h, sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if p + ".skip_connection.weight" in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def cross_attention(sd: SD, p: str, x, context, heads: int):
    """ldm/modules/attention.py::CrossAttention.forward: softmax(q k^T * d^-0.5) v, to_out with bias."""
    ctx = x if context is None else context
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(ctx, sd[p + ".to_k.weight"])
    v = F.linear(ctx, sd[p + ".to_v.weight"])
    b, n, c = q.shape
    d = c // heads
    q, k, v = (t.reshape(b, -1, heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
    if USE_SDPA:
        out = F.scaled_dot_product_attention(q, k, v)
    else:
        sim = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
        out = torch.matmul(torch.softmax(sim.float(), dim=-1).to(v.dtype), v)
    out = out.permute(0, 2, 1, 3).reshape(b, n, c)
    return F.linear(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def _ln(x, sd, key):
    return F.layer_norm(x.float(), (x.shape[-1],), sd[key + ".weight"].float(), sd[key + ".bias"].float(), 1e-5).to(x.dtype)


def transformer_block(sd: SD, p: str, x, context, heads: int):
    """attention.py::BasicTransformerBlock._forward (attn1 self, attn2 cross, GEGLU feed-forward)."""
    x = cross_attention(sd, p + ".attn1", _ln(x, sd, p + ".norm1"), None, heads) + x
    x = cross_attention(sd, p + ".attn2", _ln(x, sd, p + ".norm2"), context, heads) + x
    h = F.linear(_ln(x, sd, p + ".norm3"), sd[p + ".ff.net.0.proj.weight"], sd[p + ".ff.net.0.proj.bias"])
    a, g = h.chunk(2, dim=-1)
    h = a * F.gelu(g)
    return F.linear(h, sd[p + ".ff.net.2.weight"], sd[p + ".ff.net.2.bias"]) + x


def spatial_transformer(sd: SD, p: str, x, context, heads: int, depth: int = 1, use_linear: bool = False):
    """attention.py::SpatialTransformer.forward: conv proj_in/out (SD1.x) or, with use_linear (SD2 / SDXL), Linear
    layers applied after / before the (b c h w <-> b (h w) c) rearrange."""
    b, c, h, w = x.shape
    x_in = x
    x = _gn(x, sd, p + ".norm", 1e-6)
    if not use_linear:
        x = F.conv2d(x, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    if use_linear:
        x = F.linear(x, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    for i in range(depth):
        x = transformer_block(sd, f"{p}.transformer_blocks.{i}", x, context, heads)
    if use_linear:
        x = F.linear(x, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    if not use_linear:
        x = F.conv2d(x, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x + x_in


def _run_block(sd, cfg, prefix, blk, h, emb, context):
    for j, layer in enumerate(blk):
        p = f"{prefix}.{j}"
        kind = layer[0]
        if kind == "conv_in":
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif kind == "res":
            h = res_block(sd, p, h, emb)
        elif kind == "attn":
            h = spatial_transformer(sd, p, h, context, cfg.heads(layer[1]), layer[2], cfg.linear_proj)
        elif kind == "down":
            h = F.conv2d(h, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
        elif kind == "up":
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
    return h


def unet_forward(sd: SD, cfg: UNetConfig, x, t, context, prefix: str = "model.diffusion_model.", y=None):
    """openaimodel.UNetModel.forward(x, timesteps, context, y) -> eps.  y [N, adm_in_channels]: SDXL's vector conditioning,
    emb = time_embed(t_emb) + label_emb(y)."""
    sdp = _Prefixed(sd, prefix)
    inputs, middle, outputs = unet_layout(cfg)
    emb = timestep_embedding(t, cfg.model_channels).to(x.dtype)
    emb = F.linear(emb, sdp["time_embed.0.weight"], sdp["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sdp["time_embed.2.weight"], sdp["time_embed.2.bias"])
    if cfg.adm_in_channels:
        le = F.linear(y.to(x.dtype), sdp["label_emb.0.0.weight"], sdp["label_emb.0.0.bias"])
        emb = emb + F.linear(F.silu(le), sdp["label_emb.0.2.weight"], sdp["label_emb.0.2.bias"])
    hs = []
    h = x
    for i, blk in enumerate(inputs):
        h = _run_block(sdp, cfg, f"input_blocks.{i}", blk, h, emb, context)
        hs.append(h)
    h = _run_block(sdp, cfg, "middle_block", middle, h, emb, context)
    for i, blk in enumerate(outputs):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(sdp, cfg, f"output_blocks.{i}", blk, h, emb, context)
    h = F.silu(_gn(h, sdp, "out.0", 1e-5))
    return F.conv2d(h, sdp["out.2.weight"], sdp["out.2.bias"], padding=1)


class _Prefixed:
    """dict view adding a key prefix (ldm checkpoints prefix the UNet with 'model.diffusion_model.')."""

    def __init__(self, sd, prefix):
        self.sd, self.prefix = sd, prefix

    def __getitem__(self, k):
        return self.sd[self.prefix + k]

    def __contains__(self, k):
        return (self.prefix + k) in self.sd


# ------------------------------------------------------------------------------------------------ VAE
def _vae_res(sd, p, x):
    """model.py::ResnetBlock.forward (swish, GN eps 1e-6, temb unused)."""
    h = F.conv2d(F.silu(_gn(x, sd, p + ".norm1", 1e-6)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(F.silu(_gn(h, sd, p + ".norm2", 1e-6)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if p + ".nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    return x + h


def _vae_attn(sd, p, x):
    """model.py::AttnBlock.forward: single head, d = C, softmax(q^T k * C^-0.5)."""
    h = _gn(x, sd, p + ".norm", 1e-6)
    q = F.conv2d(h, sd[p + ".q.weight"], sd[p + ".q.bias"])
    k = F.conv2d(h, sd[p + ".k.weight"], sd[p + ".k.bias"])
    v = F.conv2d(h, sd[p + ".v.weight"], sd[p + ".v.bias"])
    b, c, hh, ww = q.shape
    if USE_SDPA:
        qs, ks, vs = (t.reshape(b, 1, c, hh * ww).transpose(-1, -2) for t in (q, k, v))
        h = F.scaled_dot_product_attention(qs, ks, vs).transpose(-1, -2).reshape(b, c, hh, ww)
        return x + F.conv2d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.softmax((torch.bmm(q, k) * (c ** -0.5)).float(), dim=2).to(q.dtype)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + F.conv2d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])


def vae_decode(sd: SD, cfg: VAEConfig, z, prefix: str = "first_stage_model."):
    """AutoencoderKL.decode: post_quant_conv then model.py::Decoder.forward. z is the UNscaled latent."""
    s = _Prefixed(sd, prefix)
    z = F.conv2d(z, s["post_quant_conv.weight"], s["post_quant_conv.bias"])
    h = F.conv2d(z, s["decoder.conv_in.weight"], s["decoder.conv_in.bias"], padding=1)
    h = _vae_res(s, "decoder.mid.block_1", h)
    h = _vae_attn(s, "decoder.mid.attn_1", h)
    h = _vae_res(s, "decoder.mid.block_2", h)
    nlev = len(cfg.ch_mult)
    for lvl in reversed(range(nlev)):
        for i in range(cfg.num_res_blocks + 1):
            h = _vae_res(s, f"decoder.up.{lvl}.block.{i}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, s[f"decoder.up.{lvl}.upsample.conv.weight"], s[f"decoder.up.{lvl}.upsample.conv.bias"], padding=1)
    h = F.silu(_gn(h, s, "decoder.norm_out", 1e-6))
    return F.conv2d(h, s["decoder.conv_out.weight"], s["decoder.conv_out.bias"], padding=1)


def vae_encode_mean(sd: SD, cfg: VAEConfig, x, prefix: str = "first_stage_model."):
    """AutoencoderKL.encode -> DiagonalGaussianDistribution.mean (the oracle uses the mean, not a sample:
    upstream samples with the global RNG)."""
    s = _Prefixed(sd, prefix)
    h = F.conv2d(x, s["encoder.conv_in.weight"], s["encoder.conv_in.bias"], padding=1)
    nlev = len(cfg.ch_mult)
    for lvl in range(nlev):
        for i in range(cfg.num_res_blocks):
            h = _vae_res(s, f"encoder.down.{lvl}.block.{i}", h)
        if lvl != nlev - 1:
            h = F.pad(h, (0, 1, 0, 1))
            h = F.conv2d(h, s[f"encoder.down.{lvl}.downsample.conv.weight"], s[f"encoder.down.{lvl}.downsample.conv.bias"], stride=2)
    h = _vae_res(s, "encoder.mid.block_1", h)
    h = _vae_attn(s, "encoder.mid.attn_1", h)
    h = _vae_res(s, "encoder.mid.block_2", h)
    h = F.silu(_gn(h, s, "encoder.norm_out", 1e-6))
    h = F.conv2d(h, s["encoder.conv_out.weight"], s["encoder.conv_out.bias"], padding=1)
    moments = F.conv2d(h, s["quant_conv.weight"], s["quant_conv.bias"])
    return moments.chunk(2, dim=1)[0]


# ------------------------------------------------------------------------------------------------ CLIP text
def clip_text_encode(sd: SD, cfg: CLIPConfig, tokens, prefix: str = "cond_stage_model.transformer.text_model."):
    """CLIP ViT-L/14 text tower (transformers CLIPTextModel): causal mask, quick-gelu MLP, final LN; last hidden state."""
    s = _Prefixed(sd, prefix)
    x = s["embeddings.token_embedding.weight"][tokens] + s["embeddings.position_embedding.weight"][None, :tokens.shape[1]]
    n = tokens.shape[1]
    mask = torch.full((n, n), float("-inf"), device=x.device, dtype=x.dtype).triu(1)
    d = cfg.width // cfg.heads
    for i in range(cfg.layers):
        p = f"encoder.layers.{i}"
        h = F.layer_norm(x, (cfg.width,), s[p + ".layer_norm1.weight"], s[p + ".layer_norm1.bias"], 1e-5)
        q = F.linear(h, s[p + ".self_attn.q_proj.weight"], s[p + ".self_attn.q_proj.bias"])
        k = F.linear(h, s[p + ".self_attn.k_proj.weight"], s[p + ".self_attn.k_proj.bias"])
        v = F.linear(h, s[p + ".self_attn.v_proj.weight"], s[p + ".self_attn.v_proj.bias"])
        b = x.shape[0]
        q, k, v = (t.reshape(b, n, cfg.heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
        att = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5 + mask, dim=-1) @ v
        att = att.permute(0, 2, 1, 3).reshape(b, n, cfg.width)
        x = x + F.linear(att, s[p + ".self_attn.out_proj.weight"], s[p + ".self_attn.out_proj.bias"])
        h = F.layer_norm(x, (cfg.width,), s[p + ".layer_norm2.weight"], s[p + ".layer_norm2.bias"], 1e-5)
        h = F.linear(h, s[p + ".mlp.fc1.weight"], s[p + ".mlp.fc1.bias"])
        h = h * torch.sigmoid(1.702 * h)
        x = x + F.linear(h, s[p + ".mlp.fc2.weight"], s[p + ".mlp.fc2.bias"])
    return F.layer_norm(x, (cfg.width,), s["final_layer_norm.weight"], s["final_layer_norm.bias"], 1e-5)


def clip_text_hidden(sd: SD, cfg: CLIPConfig, tokens, layer_idx: int, prefix: str):
    """transformers CLIPTextModel(output_hidden_states=True).hidden_states[layer_idx]: the residual stream after
    `layer_idx` encoder layers, no final LayerNorm (sgm FrozenCLIPEmbedder layer="hidden", layer_idx=11 for SDXL)."""
    s = _Prefixed(sd, prefix)
    x = s["embeddings.token_embedding.weight"][tokens] + s["embeddings.position_embedding.weight"][None, :tokens.shape[1]]
    n = tokens.shape[1]
    mask = torch.full((n, n), float("-inf"), device=x.device, dtype=x.dtype).triu(1)
    d = cfg.width // cfg.heads
    for i in range(layer_idx):
        p = f"encoder.layers.{i}"
        h = F.layer_norm(x, (cfg.width,), s[p + ".layer_norm1.weight"], s[p + ".layer_norm1.bias"], 1e-5)
        q = F.linear(h, s[p + ".self_attn.q_proj.weight"], s[p + ".self_attn.q_proj.bias"])
        k = F.linear(h, s[p + ".self_attn.k_proj.weight"], s[p + ".self_attn.k_proj.bias"])
        v = F.linear(h, s[p + ".self_attn.v_proj.weight"], s[p + ".self_attn.v_proj.bias"])
        b = x.shape[0]
        q, k, v = (t.reshape(b, n, cfg.heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
        att = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5 + mask, dim=-1) @ v
        x = x + F.linear(att.permute(0, 2, 1, 3).reshape(b, n, cfg.width), s[p + ".self_attn.out_proj.weight"],
                         s[p + ".self_attn.out_proj.bias"])
        h = F.layer_norm(x, (cfg.width,), s[p + ".layer_norm2.weight"], s[p + ".layer_norm2.bias"], 1e-5)
        h = F.linear(h, s[p + ".mlp.fc1.weight"], s[p + ".mlp.fc1.bias"])
        x = x + F.linear(h * torch.sigmoid(1.702 * h), s[p + ".mlp.fc2.weight"], s[p + ".mlp.fc2.bias"])
    return x


def open_clip_text(sd: SD, cfg: CLIPConfig, tokens, prefix: str):
    """sgm FrozenOpenCLIPEmbedder2 (arch ViT-bigG-14, layer="penultimate", legacy=False, always_return_pooled):
    open_clip text transformer (pre-LN resblocks, packed in_proj, GELU MLP, causal mask).  Returns
    (penultimate residual stream [B, 77, W] — no ln_final —, pooled [B, proj] = ln_final(last)[EOS] @ text_projection,
    EOS = argmax of the token ids)."""
    s = _Prefixed(sd, prefix)
    w, heads, n = cfg.xl_width, cfg.xl_heads, tokens.shape[1]
    x = s["token_embedding.weight"][tokens] + s["positional_embedding"][None, :n]
    mask = torch.full((n, n), float("-inf"), device=x.device, dtype=x.dtype).triu(1)
    d = w // heads
    b = x.shape[0]
    penultimate = None
    for i in range(cfg.xl_layers):
        p = f"transformer.resblocks.{i}"
        if i == cfg.xl_layers - 1:
            penultimate = x
        h = F.layer_norm(x, (w,), s[p + ".ln_1.weight"], s[p + ".ln_1.bias"], 1e-5)
        q, k, v = F.linear(h, s[p + ".attn.in_proj_weight"], s[p + ".attn.in_proj_bias"]).chunk(3, dim=-1)
        q, k, v = (t.reshape(b, n, heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
        att = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5 + mask, dim=-1) @ v
        x = x + F.linear(att.permute(0, 2, 1, 3).reshape(b, n, w), s[p + ".attn.out_proj.weight"], s[p + ".attn.out_proj.bias"])
        h = F.layer_norm(x, (w,), s[p + ".ln_2.weight"], s[p + ".ln_2.bias"], 1e-5)
        x = x + F.linear(F.gelu(F.linear(h, s[p + ".mlp.c_fc.weight"], s[p + ".mlp.c_fc.bias"])), s[p + ".mlp.c_proj.weight"],
                         s[p + ".mlp.c_proj.bias"])
    last = F.layer_norm(x, (w,), s["ln_final.weight"], s["ln_final.bias"], 1e-5)
    pooled = last[torch.arange(b, device=x.device), tokens.argmax(dim=-1)] @ s["text_projection"]
    return penultimate, pooled


def sdxl_conditioner(sd: SD, cfg: CLIPConfig, tokens, width: int, height: int, zero_txt: bool = False,
                     crop=(0, 0)):
    """sgm GeneralConditioner as sdwui feeds it (sd_models_xl.get_learned_conditioning): crossattn = cat(CLIP-L hidden
    layer 11, bigG penultimate) [B, 77, 2048]; vector = cat(bigG pooled, Fourier(original_size h, w), Fourier(crop top,
    left), Fourier(target_size h, w)) [B, 2816], each scalar through timestep_embedding(., 256).  zero_txt: sdwui's
    force_zero_embeddings=['txt'] for an all-empty negative prompt — both text outputs are zeros, the size part stays."""
    b = tokens.shape[0]
    h0 = clip_text_hidden(sd, cfg, tokens, cfg.layers - 1, "conditioner.embedders.0.transformer.text_model.")
    h1, pooled = open_clip_text(sd, cfg, tokens, "conditioner.embedders.1.model.")
    ctx = torch.cat([h0, h1], dim=-1)
    if zero_txt:
        ctx, pooled = torch.zeros_like(ctx), torch.zeros_like(pooled)
    scal = torch.tensor([height, width, crop[0], crop[1], height, width], dtype=torch.float32, device=ctx.device)
    emb = timestep_embedding(scal, cfg.size_embed_dim).reshape(1, -1).expand(b, -1).to(ctx.dtype)
    return ctx, torch.cat([pooled, emb], dim=-1)


# ------------------------------------------------------------------------------------------------ schedule / samplers
def alphas_cumprod() -> torch.Tensor:
    """ldm DDPM.register_schedule('linear', 1000, 0.00085, 0.012): fp64 math, stored fp32."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    return torch.cumprod(1.0 - betas, dim=0).to(torch.float32)


def ddim_timesteps(steps: int) -> torch.Tensor:
    """sdwui sd_samplers_timesteps: uniform discretisation, clip(arange(0,1000,1000//steps)+1, 0, 999)."""
    return torch.clamp(torch.arange(0, 1000, 1000 // steps) + 1, 0, 999)


def ddim_coefficients(steps: int) -> List[Tuple[int, float, float, float, float]]:
    """Per executed step, in execution order: (t, sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)) with eta=0.
    sdwui sd_samplers_timesteps_impl.ddim: alphas_prev = alphas_cumprod[pad(ts[:-1], (1, 0))] (index 0 pads to
    alphas_cumprod[0]); `for i in trange(len(ts) - 1): index = len(ts) - 1 - i` — i.e. index runs from len(ts)-1
    down to 1, so `steps - 1` UNet evaluations are executed for `steps` timesteps (19 for the 20-step configs)."""
    ac = alphas_cumprod().double()
    ts = ddim_timesteps(steps)
    alphas = ac[ts]
    alphas_prev = ac[torch.cat([ts[:1] * 0, ts[:-1]])]
    out = []
    for i in range(len(ts) - 1, 0, -1):
        a_t, a_p = float(alphas[i]), float(alphas_prev[i])
        out.append((int(ts[i]), math.sqrt(a_t), math.sqrt(1.0 - a_t), math.sqrt(a_p), math.sqrt(1.0 - a_p)))
    return out


def cfg_eps(unet, x, t: int, cond, uncond, cfg_scale: float):
    """sdwui CFGDenoiser: one UNet call on cat([x, x]) with cat([cond, uncond]); uncond + s * (cond - uncond)."""
    b = x.shape[0]
    tt = torch.full((2 * b,), float(t), device=x.device)
    e = unet(torch.cat([x, x]), tt, torch.cat([cond, uncond]))
    ec, eu = e[:b], e[b:]
    return eu + cfg_scale * (ec - eu)


def sample_ddim(unet, x_T, cond, uncond, steps: int, cfg_scale: float, start_index: int = 0):
    x = x_T
    for (t, sa, s1a, sap, s1ap) in ddim_coefficients(steps)[start_index:]:
        e = cfg_eps(unet, x, t, cond, uncond, cfg_scale)
        x0 = (x - s1a * e) / sa
        x = sap * x0 + s1ap * e
    return x


def karras_sigmas_compvis(steps: int):
    """k-diffusion DiscreteSchedule.get_sigmas(n): t = linspace(999, 0, n), log-sigma interpolation, append 0."""
    ac = alphas_cumprod().double()
    sig = ((1 - ac) / ac) ** 0.5
    log_sig = sig.log()
    t = torch.linspace(len(sig) - 1, 0, steps, dtype=torch.float64)
    lo = t.floor().long()
    hi = t.ceil().long()
    w = t - lo
    s = ((1 - w) * log_sig[lo] + w * log_sig[hi]).exp()
    return torch.cat([s, s.new_zeros(1)]), log_sig


def sigma_to_t(sigma: float, log_sig: torch.Tensor) -> float:
    """k-diffusion DiscreteSchedule.sigma_to_t (quantize=False): fractional timestep by log-sigma interpolation."""
    ls = math.log(sigma)
    dists = ls - log_sig
    low = int((dists >= 0).cumsum(0).argmax().clamp(max=len(log_sig) - 2))
    high = low + 1
    lo, hi = float(log_sig[low]), float(log_sig[high])
    w = min(max((lo - ls) / (lo - hi), 0.0), 1.0)
    return (1 - w) * low + w * high


def euler_a_coefficients(steps: int):
    """Per step: (t, sigma, sigma_down, sigma_up, c_in, c_in_next) — k-diffusion sample_euler_ancestral + CompVisDenoiser."""
    sig, log_sig = karras_sigmas_compvis(steps)
    out = []
    for i in range(steps):
        s, sn = float(sig[i]), float(sig[i + 1])
        up = min(sn, (sn ** 2 * (s ** 2 - sn ** 2) / s ** 2) ** 0.5)
        down = (sn ** 2 - up ** 2) ** 0.5
        out.append((sigma_to_t(s, log_sig), s, down, up, 1.0 / math.sqrt(s * s + 1.0), 1.0 / math.sqrt(sn * sn + 1.0)))
    return out


def sample_euler_a(unet, x_T, cond, uncond, steps: int, cfg_scale: float, noises):
    """noises[i] is the N(0,1) draw used after step i (per-image generators upstream; injected here)."""
    coefs = euler_a_coefficients(steps)
    x = x_T * coefs[0][1]
    for i, (t, s, down, up, c_in, _) in enumerate(coefs):
        e = cfg_eps(unet, x * c_in, t, cond, uncond, cfg_scale)
        x = x + e * (down - s)       # d = (x - denoised) / sigma = eps ; x += d * (sigma_down - sigma)
        if up > 0:
            x = x + noises[i] * up
    return x


def sample_euler(unet, x_T, cond, uncond, steps: int, cfg_scale: float):
    """k-diffusion sample_euler with s_churn = 0 (sdwui "Euler"): d = (x - denoised) / sigma = eps,
    x += d * (sigma_next - sigma) on the same CompVisDenoiser sigmas as Euler a."""
    coefs = euler_a_coefficients(steps)
    x = x_T * coefs[0][1]
    for i, (t, s, down, up, c_in, _) in enumerate(coefs):
        sn = math.sqrt(down * down + up * up)
        e = cfg_eps(unet, x * c_in, t, cond, uncond, cfg_scale)
        x = x + e * (sn - s)
    return x


def sigmas_karras(steps: int):
    """k-diffusion get_sigmas_karras(n, sigma_min, sigma_max, rho = 7) with sdwui's bounds (the model's own smallest
    and largest sigma: sd_samplers_kdiffusion.KDiffusionSampler.get_sigmas, use_old_karras_scheduler_sigmas off)."""
    ac = alphas_cumprod().double()
    sig = ((1 - ac) / ac) ** 0.5
    ramp = torch.linspace(0, 1, steps, dtype=torch.float64)
    min_inv_rho, max_inv_rho = float(sig[0]) ** (1 / 7.0), float(sig[-1]) ** (1 / 7.0)
    s = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** 7.0
    return torch.cat([s, s.new_zeros(1)]), sig.log()


def sigmas_exponential(steps: int):
    """k-diffusion get_sigmas_exponential(n, sigma_min, sigma_max) with the model's own bounds (sdwui "Exponential";
    "Polyexponential" is the same schedule at its default rho = 1)."""
    ac = alphas_cumprod().double()
    sig = ((1 - ac) / ac) ** 0.5
    s = torch.linspace(math.log(float(sig[-1])), math.log(float(sig[0])), steps, dtype=torch.float64).exp()
    return torch.cat([s, s.new_zeros(1)]), sig.log()


def sigmas_sgm_uniform(steps: int):
    """sdwui sd_schedulers.sgm_uniform: sigmas at timesteps linspace(t(sigma_max), t(sigma_min), n + 1)[:-1], then 0."""
    ac = alphas_cumprod().double()
    sig = ((1 - ac) / ac) ** 0.5
    log_sig = sig.log()
    start, end = sigma_to_t(float(sig[-1]), log_sig), sigma_to_t(float(sig[0]), log_sig)
    out = []
    for t in torch.linspace(start, end, steps + 1, dtype=torch.float64)[:-1].tolist():   # DiscreteSchedule.t_to_sigma
        lo = int(math.floor(t))
        hi = int(math.ceil(t))
        w = t - lo
        out.append(math.exp((1 - w) * float(log_sig[lo]) + w * float(log_sig[hi])))
    return torch.tensor(out + [0.0], dtype=torch.float64), log_sig


def sample_euler_sigmas(unet, x_T, cond, uncond, sig, log_sig, cfg_scale: float):
    """k-diffusion sample_euler (s_churn 0) over an explicit sigma schedule"""
    x = x_T * float(sig[0])
    for i in range(len(sig) - 1):
        s, sn = float(sig[i]), float(sig[i + 1])
        e = cfg_eps(unet, x * (1.0 / math.sqrt(s * s + 1.0)), sigma_to_t(s, log_sig), cond, uncond, cfg_scale)
        x = x + e * (sn - s)
    return x


def sample_dpmpp_2m(unet, x_T, cond, uncond, steps: int, cfg_scale: float, karras: bool = True):
    """k-diffusion sample_dpmpp_2m (sdwui "DPM++ 2M" / "DPM++ 2M Karras") around the eps-prediction CompVisDenoiser:
    denoised = x - sigma * eps(x * c_in, t(sigma)); written with t = -log(sigma) exactly as upstream."""
    sig, log_sig = sigmas_karras(steps) if karras else karras_sigmas_compvis(steps)
    x = x_T * float(sig[0])
    old = None
    for i in range(steps):
        s, sn = float(sig[i]), float(sig[i + 1])
        e = cfg_eps(unet, x * (1.0 / math.sqrt(s * s + 1.0)), sigma_to_t(s, log_sig), cond, uncond, cfg_scale)
        denoised = x - s * e
        t = -math.log(s)
        t_next = -math.log(sn) if sn > 0 else math.inf
        h = t_next - t
        ratio = sn / s                      # sigma_fn(t_next) / sigma_fn(t)
        if old is None or sn == 0:
            x = ratio * x - math.expm1(-h) * denoised
        else:
            h_last = t - (-math.log(float(sig[i - 1])))
            r = h_last / h
            dd = (1 + 1 / (2 * r)) * denoised - (1 / (2 * r)) * old
            x = ratio * x - math.expm1(-h) * dd
        old = denoised
    return x


def kdiff_img2img_sigmas(sig: torch.Tensor, steps: int, denoising_strength: float) -> torch.Tensor:
    """sdwui KDiffusionSampler.sample_img2img: t_enc = int(min(d, 0.999) * steps) (setup_img2img_steps without
    img2img_fix_steps); sigma_sched = sigmas[steps - t_enc - 1:]; the start is x = init + noise * sigma_sched[0]."""
    t_enc = int(min(denoising_strength, 0.999) * steps)
    return sig[steps - t_enc - 1:]


def sample_kdiff_img2img(unet, init, noises, cond, uncond, sig_sched, log_sig, cfg_scale: float, method: str):
    """the tail of a k-diffusion sampler from a noised init latent.  noises[0] noises the start, noises[1:] are the
    ancestral draws (method "euler_a"); methods: "euler", "euler_a", "dpmpp_2m"."""
    x = init + noises[0] * float(sig_sched[0])
    old = None
    for i in range(len(sig_sched) - 1):
        s, sn = float(sig_sched[i]), float(sig_sched[i + 1])
        e = cfg_eps(unet, x * (1.0 / math.sqrt(s * s + 1.0)), sigma_to_t(s, log_sig), cond, uncond, cfg_scale)
        if method == "euler":
            x = x + e * (sn - s)
        elif method == "euler_a":
            up = min(sn, (sn ** 2 * (s ** 2 - sn ** 2) / s ** 2) ** 0.5)
            down = (sn ** 2 - up ** 2) ** 0.5
            x = x + e * (down - s)
            if up > 0:
                x = x + noises[1 + i] * up
        elif method == "dpmpp_2m":
            denoised = x - s * e
            if old is None or sn == 0:
                x = (sn / s) * x + (1 - sn / s) * denoised
            else:
                h, h_last = math.log(s / sn), math.log(float(sig_sched[i - 1]) / s)
                r = h_last / h
                x = (sn / s) * x + (1 - sn / s) * ((1 + 1 / (2 * r)) * denoised - (1 / (2 * r)) * old)
            old = denoised
        else:
            raise ValueError(method)
    return x


# ------------------------------------------------------------------------------------------------ k-diffusion, the rest
# Restated in k-diffusion's own shape: a `model(x, sigma) -> denoised` callable (CompVisDenoiser around the CFG'd eps
# network) and samplers over a sigma table.  (k-diffusion sampling.py / external.py; sdwui sd_samplers_cfg_denoiser.py)
def kdiff_model(unet, cond, uncond, cfg_scale: float, log_sig: torch.Tensor, mask=None):
    """CompVisDenoiser.forward inside sdwui's CFGDenoiser: denoised = x + eps(x * c_in, t(sigma)) * c_out with
    c_in = 1 / sqrt(sigma^2 + 1), c_out = -sigma.  `mask` = (init_latent, nmask): CFGDenoiser's last lines for the
    k-diffusion samplers (mask_before_denoising False): denoised = init_latent * mask + nmask * denoised."""
    def model(x, sigma: float):
        e = cfg_eps(unet, x * (1.0 / math.sqrt(sigma * sigma + 1.0)), sigma_to_t(sigma, log_sig), cond, uncond, cfg_scale)
        den = x - sigma * e
        if mask is not None:
            init, nmask = mask
            den = init * (1.0 - nmask) + nmask * den
        return den
    return model


def to_d(x, sigma: float, denoised):
    """k-diffusion to_d"""
    return (x - denoised) / sigma


def get_ancestral_step(sigma_from: float, sigma_to: float, eta: float = 1.0):
    """k-diffusion get_ancestral_step"""
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


def k_sample_euler(model, x, sigmas):
    for i in range(len(sigmas) - 1):
        s, sn = float(sigmas[i]), float(sigmas[i + 1])
        d = to_d(x, s, model(x, s))
        x = x + d * (sn - s)
    return x


def k_sample_euler_ancestral(model, x, sigmas, noises):
    for i in range(len(sigmas) - 1):
        s, sn = float(sigmas[i]), float(sigmas[i + 1])
        denoised = model(x, s)
        down, up = get_ancestral_step(s, sn)
        x = x + to_d(x, s, denoised) * (down - s)
        if sn > 0:
            x = x + noises[i] * up
    return x


def k_sample_heun(model, x, sigmas):
    """sample_heun with s_churn = 0"""
    for i in range(len(sigmas) - 1):
        s, sn = float(sigmas[i]), float(sigmas[i + 1])
        d = to_d(x, s, model(x, s))
        dt = sn - s
        if sn == 0:
            x = x + d * dt
        else:
            x_2 = x + d * dt
            d_2 = to_d(x_2, sn, model(x_2, sn))
            x = x + (d + d_2) / 2 * dt
    return x


def k_sample_dpm_2(model, x, sigmas):
    """sample_dpm_2 with s_churn = 0"""
    for i in range(len(sigmas) - 1):
        s, sn = float(sigmas[i]), float(sigmas[i + 1])
        d = to_d(x, s, model(x, s))
        if sn == 0:
            x = x + d * (sn - s)
        else:
            sigma_mid = math.exp(0.5 * math.log(s) + 0.5 * math.log(sn))     # sigma.log().lerp(sigma_next.log(), 0.5).exp()
            x_2 = x + d * (sigma_mid - s)
            d_2 = to_d(x_2, sigma_mid, model(x_2, sigma_mid))
            x = x + d_2 * (sn - s)
    return x


def k_sample_dpm_2_ancestral(model, x, sigmas, noises):
    """sample_dpm_2_ancestral, eta = s_noise = 1; noises[k] = k-th noise_sampler call"""
    k = 0
    for i in range(len(sigmas) - 1):
        s, sn = float(sigmas[i]), float(sigmas[i + 1])
        denoised = model(x, s)
        down, up = get_ancestral_step(s, sn)
        d = to_d(x, s, denoised)
        if down == 0:
            x = x + d * (down - s)
        else:
            sigma_mid = math.exp(0.5 * math.log(s) + 0.5 * math.log(down))
            x_2 = x + d * (sigma_mid - s)
            d_2 = to_d(x_2, sigma_mid, model(x_2, sigma_mid))
            x = x + d_2 * (down - s)
            x = x + noises[k] * up
            k += 1
    return x


def k_sample_dpmpp_2s_ancestral(model, x, sigmas, noises):
    """sample_dpmpp_2s_ancestral, eta = s_noise = 1"""
    sigma_fn = lambda t: math.exp(-t)  # noqa: E731
    t_fn = lambda sigma: -math.log(sigma)  # noqa: E731
    k = 0
    for i in range(len(sigmas) - 1):
        s, sn = float(sigmas[i]), float(sigmas[i + 1])
        denoised = model(x, s)
        down, up = get_ancestral_step(s, sn)
        if down == 0:
            x = x + to_d(x, s, denoised) * (down - s)
        else:
            t, t_next = t_fn(s), t_fn(down)
            r = 1 / 2
            h = t_next - t
            sm = t + r * h
            x_2 = (sigma_fn(sm) / sigma_fn(t)) * x - math.expm1(-h * r) * denoised
            denoised_2 = model(x_2, sigma_fn(sm))
            x = (sigma_fn(t_next) / sigma_fn(t)) * x - math.expm1(-h) * denoised_2
        if sn > 0:
            x = x + noises[k] * up
            k += 1
    return x


def k_sample_dpmpp_2m(model, x, sigmas):
    t_fn = lambda sigma: -math.log(sigma)  # noqa: E731
    old = None
    for i in range(len(sigmas) - 1):
        s, sn = float(sigmas[i]), float(sigmas[i + 1])
        denoised = model(x, s)
        if old is None or sn == 0:
            x = (sn / s) * x + (1 - sn / s) * denoised if sn > 0 else denoised
        else:
            t, t_next = t_fn(s), t_fn(sn)
            h = t_next - t
            h_last = t - t_fn(float(sigmas[i - 1]))
            r = h_last / h
            denoised_d = (1 + 1 / (2 * r)) * denoised - (1 / (2 * r)) * old
            x = (sn / s) * x - math.expm1(-h) * denoised_d
        old = denoised
    return x


def linear_multistep_coeff(order: int, t, i: int, j: int) -> float:
    """k-diffusion linear_multistep_coeff"""
    from scipy import integrate
    if order - 1 > i:
        raise ValueError(f"Order {order} too high for step {i}")

    def fn(tau):
        prod = 1.0
        for k in range(order):
            if j == k:
                continue
            prod *= (tau - t[i - k]) / (t[i - j] - t[i - k])
        return prod

    return integrate.quad(fn, t[i], t[i + 1], epsrel=1e-4)[0]


def k_sample_lms(model, x, sigmas, order: int = 4):
    """sample_lms"""
    sig = [float(v) for v in sigmas]
    ds = []
    for i in range(len(sig) - 1):
        d = to_d(x, sig[i], model(x, sig[i]))
        ds.append(d)
        if len(ds) > order:
            ds.pop(0)
        cur_order = min(i + 1, order)
        coeffs = [linear_multistep_coeff(cur_order, sig, i, j) for j in range(cur_order)]
        x = x + sum(coeff * dd for coeff, dd in zip(coeffs, reversed(ds)))
    return x


def brownian_pair(z1, z2, s: float, ss: float, sn: float):
    """The two noise_sampler calls of one DPM++ SDE step, (sigma -> sigma_s) and (sigma -> sigma_next).  Upstream:
    BrownianTreeNoiseSampler, (W(t1) - W(t0)) / sqrt|t1 - t0| of ONE Brownian motion on the sigma axis (torchsde, absent
    offline), so the second value contains the first interval's increment.  Restated from two independent N(0,1) draws:
    z1 drives [sigma_s, sigma], z2 drives [sigma_next, sigma_s]."""
    n1 = z1
    n2 = (math.sqrt(s - ss) * z1 + math.sqrt(ss - sn) * z2) / math.sqrt(s - sn)
    return n1, n2


def k_sample_dpmpp_sde(model, x, sigmas, draws, eta: float = 1.0, s_noise: float = 1.0, r: float = 1 / 2):
    """sample_dpmpp_sde; draws[2k], draws[2k + 1] feed the k-th step that is not the final Euler step (brownian_pair)"""
    sigma_fn = lambda t: math.exp(-t)  # noqa: E731
    t_fn = lambda sigma: -math.log(sigma)  # noqa: E731
    k = 0
    for i in range(len(sigmas) - 1):
        s, sn = float(sigmas[i]), float(sigmas[i + 1])
        denoised = model(x, s)
        if sn == 0:
            x = x + to_d(x, s, denoised) * (sn - s)
        else:
            t, t_next = t_fn(s), t_fn(sn)
            h = t_next - t
            sm = t + h * r
            fac = 1 / (2 * r)
            n1, n2 = brownian_pair(draws[2 * k], draws[2 * k + 1], s, sigma_fn(sm), sn)
            k += 1
            sd, su = get_ancestral_step(sigma_fn(t), sigma_fn(sm), eta)
            s_ = t_fn(sd)
            x_2 = (sigma_fn(s_) / sigma_fn(t)) * x - math.expm1(t - s_) * denoised
            x_2 = x_2 + n1 * s_noise * su
            denoised_2 = model(x_2, sigma_fn(sm))
            sd, su = get_ancestral_step(sigma_fn(t), sigma_fn(t_next), eta)
            t_next_ = t_fn(sd)
            denoised_d = (1 - fac) * denoised + fac * denoised_2
            x = (sigma_fn(t_next_) / sigma_fn(t)) * x - math.expm1(t - t_next_) * denoised_d
            x = x + n2 * s_noise * su
    return x


class DPMSolver:
    """k-diffusion DPMSolver (eta = 0 paths): t = -log sigma, eps = (x - model(x, sigma)) / sigma"""

    def __init__(self, model):
        self.model = model
        self.nfe = 0

    @staticmethod
    def sigma(t):
        return math.exp(-t)

    def eps(self, cache, key, x, t):
        if key in cache:
            return cache[key], cache
        self.nfe += 1
        e = (x - self.model(x, self.sigma(t))) / self.sigma(t)
        return e, {key: e, **cache}

    def dpm_solver_1_step(self, x, t, t_next, cache=None):
        cache = {} if cache is None else cache
        h = t_next - t
        eps, cache = self.eps(cache, "eps", x, t)
        return x - self.sigma(t_next) * math.expm1(h) * eps, cache

    def dpm_solver_2_step(self, x, t, t_next, r1=1 / 2, cache=None):
        cache = {} if cache is None else cache
        h = t_next - t
        eps, cache = self.eps(cache, "eps", x, t)
        s1 = t + r1 * h
        u1 = x - self.sigma(s1) * math.expm1(r1 * h) * eps
        eps_r1, cache = self.eps(cache, "eps_r1", u1, s1)
        x_2 = x - self.sigma(t_next) * math.expm1(h) * eps - self.sigma(t_next) / (2 * r1) * math.expm1(h) * (eps_r1 - eps)
        return x_2, cache

    def dpm_solver_3_step(self, x, t, t_next, r1=1 / 3, r2=2 / 3, cache=None):
        cache = {} if cache is None else cache
        h = t_next - t
        eps, cache = self.eps(cache, "eps", x, t)
        s1 = t + r1 * h
        s2 = t + r2 * h
        u1 = x - self.sigma(s1) * math.expm1(r1 * h) * eps
        eps_r1, cache = self.eps(cache, "eps_r1", u1, s1)
        u2 = x - self.sigma(s2) * math.expm1(r2 * h) * eps - \
            self.sigma(s2) * (r2 / r1) * (math.expm1(r2 * h) / (r2 * h) - 1) * (eps_r1 - eps)
        eps_r2, cache = self.eps(cache, "eps_r2", u2, s2)
        x_3 = x - self.sigma(t_next) * math.expm1(h) * eps - self.sigma(t_next) / r2 * (math.expm1(h) / h - 1) * (eps_r2 - eps)
        return x_3, cache

    def dpm_solver_fast(self, x, t_start, t_end, nfe):
        m = math.floor(nfe / 3) + 1
        ts = torch.linspace(t_start, t_end, m + 1, dtype=torch.float64).tolist()
        orders = [3] * (m - 2) + [2, 1] if nfe % 3 == 0 else [3] * (m - 1) + [nfe % 3]
        for i in range(len(orders)):
            t, t_next = ts[i], ts[i + 1]
            if orders[i] == 1:
                x, _ = self.dpm_solver_1_step(x, t, t_next)
            elif orders[i] == 2:
                x, _ = self.dpm_solver_2_step(x, t, t_next)
            else:
                x, _ = self.dpm_solver_3_step(x, t, t_next)
        return x

    def dpm_solver_adaptive(self, x, t_start, t_end, order=3, rtol=0.05, atol=0.0078, h_init=0.05, pcoeff=0.0, icoeff=1.0,
                            dcoeff=0.0, accept_safety=0.81):
        """forward in t, eta = 0; returns (x, info)"""
        b1, b2, b3 = (pcoeff + icoeff + dcoeff) / order, -(pcoeff + 2 * dcoeff) / order, dcoeff / order
        errs, h = [], h_init
        s = t_start
        x_prev = x
        info = dict(steps=0, nfe=0, n_accept=0, n_reject=0)
        while s < t_end - 1e-5:
            cache = {}
            t = min(t_end, s + h)
            eps, cache = self.eps(cache, "eps", x, s)
            x_low, cache = self.dpm_solver_2_step(x, s, t, r1=1 / 3, cache=cache)
            x_high, cache = self.dpm_solver_3_step(x, s, t, cache=cache)
            delta = torch.maximum(torch.full_like(x_low, atol), rtol * torch.maximum(x_low.abs(), x_prev.abs()))
            error = float(torch.linalg.norm((x_low - x_high) / delta) / x.numel() ** 0.5)
            inv_error = 1 / (error + 1e-8)                     # PIDStepSizeController.propose_step
            if not errs:
                errs = [inv_error, inv_error, inv_error]
            errs[0] = inv_error
            factor = errs[0] ** b1 * errs[1] ** b2 * errs[2] ** b3
            factor = 1 + math.atan(factor - 1)
            accept = factor >= accept_safety
            if accept:
                errs[2] = errs[1]
                errs[1] = errs[0]
            h *= factor
            if accept:
                x_prev = x_low
                x = x_high
                s = t
                info["n_accept"] += 1
            else:
                info["n_reject"] += 1
            info["nfe"] += order
            info["steps"] += 1
        return x, info


def k_sample_dpm_fast(model, x, sigma_min: float, sigma_max: float, n: int):
    """sample_dpm_fast"""
    return DPMSolver(model).dpm_solver_fast(x, -math.log(sigma_max), -math.log(sigma_min), n)


def k_sample_dpm_adaptive(model, x, sigma_min: float, sigma_max: float):
    """sample_dpm_adaptive (defaults)"""
    return DPMSolver(model).dpm_solver_adaptive(x, -math.log(sigma_max), -math.log(sigma_min))


def sample_plms(unet, x, cond, uncond, timesteps, cfg_scale: float, mask=None):
    """sdwui sd_samplers_timesteps_impl.plms over ascending `timesteps` (the model there is the CFG'd eps network;
    `mask` = (init_latent, nmask): CFGDenoiser with mask_before_denoising — the input is blended before every call)."""
    ac = alphas_cumprod().double()
    ts = [int(t) for t in timesteps]
    alphas = [float(ac[t]) for t in ts]
    alphas_prev = [float(ac[0])] + [float(ac[t]) for t in ts[:-1]]

    def model(xx, t):
        if mask is not None:
            init, nmask = mask
            xx = init * (1.0 - nmask) + nmask * xx
        return cfg_eps(unet, xx, t, cond, uncond, cfg_scale)

    old_eps = []
    for i in range(len(ts) - 1):
        index = len(ts) - 1 - i
        t_here, t_next = ts[index], ts[max(index - 1, 0)]
        a_t, a_prev = alphas[index], alphas_prev[index]

        def get_x_prev(e_t):
            pred_x0 = (x - math.sqrt(1 - a_t) * e_t) / math.sqrt(a_t)
            return math.sqrt(a_prev) * pred_x0 + math.sqrt(1.0 - a_prev) * e_t

        e_t = model(x, t_here)
        if len(old_eps) == 0:
            e_t_next = model(get_x_prev(e_t), t_next)
            e_t_prime = (e_t + e_t_next) / 2
        elif len(old_eps) == 1:
            e_t_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_t_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_t_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        x_prev = get_x_prev(e_t_prime)
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
        x = x_prev
    return x


def model_sigmas():
    """CompVisDenoiser.sigmas / log_sigmas of the SD schedule"""
    ac = alphas_cumprod().double()
    sig = ((1 - ac) / ac) ** 0.5
    return sig, sig.log()


def run_sampler(name: str, unet, cond, uncond, cfg_scale: float, steps: int, noise0, draws=None, init=None,
                denoising_strength: Optional[float] = None, mask=None):
    """One sampling run as sdwui dispatches it by sampler NAME (sd_samplers_kdiffusion.samplers_k_diffusion /
    sd_samplers_timesteps): txt2img from noise0 (x = noise0 * sigmas[0]), or — with `init` and `denoising_strength` — the
    img2img half (KDiffusionSampler.sample_img2img: sigma_sched = sigmas[steps - t_enc - 1:], x = init + noise0 * sigma_sched[0];
    timestep samplers: timesteps[:t_enc], x = init sqrt(a) + noise0 sqrt(1 - a)).  draws: per-image N(0,1) tensors consumed
    by the stochastic samplers in call order.  mask = (init_latent, nmask) for inpainting.  Returns the final latents
    (before the caller's final mask blend)."""
    karras = name.endswith(" Karras") or name == "DPM++ 2M"
    base = name[:-len(" Karras")] if name.endswith(" Karras") else name
    if base in ("DDIM", "PLMS"):
        ts = ddim_timesteps(steps)
        if init is not None:
            ac = alphas_cumprod().double()
            t_enc = max(1, min(int(min(denoising_strength, 0.999) * steps), len(ts) - 1))
            a = float(ac[ts[t_enc]])
            x = init * math.sqrt(a) + noise0 * math.sqrt(1 - a)
            ts = ts[:t_enc]
        else:
            x = noise0
        if base == "PLMS":
            return sample_plms(unet, x, cond, uncond, ts, cfg_scale, mask)
        raise ValueError("DDIM: use sample_ddim / img2img")
    sig, log_sig = sigmas_karras(steps) if karras else karras_sigmas_compvis(steps)
    if init is not None:
        sig = kdiff_img2img_sigmas(sig, steps, denoising_strength)
        x = init + noise0 * float(sig[0])
    else:
        x = noise0 * float(sig[0])
    model = kdiff_model(unet, cond, uncond, cfg_scale, log_sig, mask)
    if base == "Euler":
        return k_sample_euler(model, x, sig)
    if base == "Euler a":
        return k_sample_euler_ancestral(model, x, sig, draws)
    if base == "Heun":
        return k_sample_heun(model, x, sig)
    if base == "DPM2":
        return k_sample_dpm_2(model, x, sig)
    if base == "DPM2 a":
        return k_sample_dpm_2_ancestral(model, x, sig, draws)
    if base == "DPM++ 2S a":
        return k_sample_dpmpp_2s_ancestral(model, x, sig, draws)
    if base == "DPM++ 2M":
        return k_sample_dpmpp_2m(model, x, sig)
    if base == "DPM++ SDE":
        return k_sample_dpmpp_sde(model, x, sig, draws)
    if base == "LMS":
        return k_sample_lms(model, x, sig)
    if base in ("DPM fast", "DPM adaptive"):
        all_sig, _ = model_sigmas()
        lo, hi = (float(all_sig[0]), float(all_sig[-1])) if init is None else (float(sig[-2]), float(sig[0]))
        if base == "DPM fast":
            return k_sample_dpm_fast(model, x, lo, hi, len(sig) - 1)
        return k_sample_dpm_adaptive(model, x, lo, hi)[0]
    raise ValueError(name)


# ------------------------------------------------------------------------------------------------ images / rng
def per_image_noise(seed: int, n: int, shape, subseed_offset: int = 0, subseed: Optional[int] = None,
                    subseed_strength: float = 0.0) -> torch.Tensor:
    """sdwui rng.ImageRNG with randn_source='CPU': image k is drawn from its own generator seeded all_seeds[k]; with
    variation seeds (subseed_strength != 0) ImageRNG.first() returns slerp(strength, noise, subnoise(subseed + k)).
    processing.py: all_seeds[k] = seed + (k if subseed_strength == 0 else 0), all_subseeds[k] = subseed + k — with
    variation seeds every image shares the base seed and only the subseed advances."""
    out = []
    variation = subseed is not None and subseed_strength != 0
    for k in range(n):
        g = torch.Generator(device="cpu").manual_seed(int(seed) + (0 if variation else k))
        noise = torch.randn(shape, generator=g, dtype=torch.float32)
        if subseed is not None and subseed_strength != 0:
            sub = torch.randn(shape, generator=torch.Generator(device="cpu").manual_seed(int(subseed) + k), dtype=torch.float32)
            # modules/rng.py slerp: unit vectors / angles along dim 1 of the [C, H, W] tensor
            ln, hn = noise / torch.norm(noise, dim=1, keepdim=True), sub / torch.norm(sub, dim=1, keepdim=True)
            dot = (ln * hn).sum(1)
            if float(dot.mean()) > 0.9995:
                noise = noise * subseed_strength + sub * (1 - subseed_strength)
            else:
                om = torch.acos(dot)
                noise = (torch.sin((1.0 - subseed_strength) * om) / torch.sin(om)).unsqueeze(1) * noise + \
                        (torch.sin(subseed_strength * om) / torch.sin(om)).unsqueeze(1) * sub
        out.append(noise)
    return torch.stack(out)


def random_prompt_tokens(n: int, seed: int = 1234, vocab_hi: int = 49405) -> torch.Tensor:
    """SURVEY.md §8d synthetic prompts: [BOS] + 75 random ids + [EOS]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    body = torch.randint(0, vocab_hi, (n, 75), generator=g)
    bos = torch.full((n, 1), vocab_hi + 1)
    eos = torch.full((n, 1), vocab_hi + 2)
    return torch.cat([bos, body, eos], dim=1)


def empty_prompt_tokens(n: int, vocab_hi: int = 49405) -> torch.Tensor:
    t = torch.full((n, 77), vocab_hi + 2)
    t[:, 0] = vocab_hi + 1
    return t


def to_uint8(decoded: torch.Tensor) -> torch.Tensor:
    """sdwui process_images_inner: clamp((x+1)/2, 0, 1) -> (255 * x).astype(uint8) (truncation), HWC."""
    x = torch.clamp((decoded.float() + 1.0) / 2.0, 0.0, 1.0)
    return (255.0 * x).permute(0, 2, 3, 1).to(torch.uint8)


def txt2img(sd: SD, unet_cfg, vae_cfg, clip_cfg, tokens, neg_tokens, seed: int, steps: int = 20, cfg_scale: float = 7.0,
            height: int = 512, width: int = 512, sampler: str = "DDIM", device="cpu"):
    """End-to-end oracle for one batch: returns (uint8 images [B,H,W,3], final latents, decoded float)."""
    b = tokens.shape[0]
    dsd = {k: v.to(device) for k, v in sd.items()}
    cond = clip_text_encode(dsd, clip_cfg, tokens.to(device))
    uncond = clip_text_encode(dsd, clip_cfg, neg_tokens.to(device))
    x_T = per_image_noise(seed, b, (unet_cfg.in_channels, height // 8, width // 8)).to(device)

    def unet(x, t, c):
        return unet_forward(dsd, unet_cfg, x, t, c)

    if sampler == "DDIM":
        z = sample_ddim(unet, x_T, cond, uncond, steps, cfg_scale)
    else:
        raise ValueError(sampler)
    dec = vae_decode(dsd, vae_cfg, z / vae_cfg.scale_factor)
    return to_uint8(dec), z, dec


# ------------------------------------------------------------------------------------------------ img2img
def image_to_model_input(img_u8: torch.Tensor) -> torch.Tensor:
    """sdwui StableDiffusionProcessingImg2Img.init: uint8 HWC -> float NCHW in [-1, 1] (2 * x/255 - 1)."""
    return img_u8.permute(0, 3, 1, 2).float().div(255.0).mul(2.0).sub(1.0)


def ddim_img2img_coefficients(steps: int, denoising_strength: float):
    """sdwui sd_samplers_timesteps.sample_img2img: t_enc = int(min(d, 0.999) * steps); the sampler runs on
    timesteps[:t_enc] (so t_enc - 1 UNet evaluations), starting from
    x = init * sqrt(a[ts[t_enc]]) + noise * sqrt(1 - a[ts[t_enc]]).
    Returns (sqrt_a_start, sqrt_1m_a_start, rows) with rows as in ddim_coefficients."""
    ac = alphas_cumprod().double()
    ts = ddim_timesteps(steps)
    t_enc = int(min(denoising_strength, 0.999) * steps)
    t_enc = max(1, min(t_enc, len(ts) - 1))
    a_start = float(ac[ts[t_enc]])
    sub = ts[:t_enc]
    alphas = ac[sub]
    alphas_prev = ac[torch.cat([sub[:1] * 0, sub[:-1]])]
    rows = []
    for i in range(len(sub) - 1, 0, -1):
        a_t, a_p = float(alphas[i]), float(alphas_prev[i])
        rows.append((int(sub[i]), math.sqrt(a_t), math.sqrt(1.0 - a_t), math.sqrt(a_p), math.sqrt(1.0 - a_p)))
    return math.sqrt(a_start), math.sqrt(1.0 - a_start), rows


def txt2img_hires(sd: SD, unet_cfg, vae_cfg, clip_cfg, tokens, neg_tokens, seed: int, steps: int = 20,
                  cfg_scale: float = 7.0, height: int = 512, width: int = 512, hr_scale: float = 2.0, hr_steps: int = 0,
                  denoising_strength: float = 0.7, device="cpu"):
    """sdwui hires fix with the "Latent" upscaler (processing.py StableDiffusionProcessingTxt2Img.sample_hr_pass):
    DDIM first pass -> F.interpolate(bilinear, antialias=False) of the latents -> fresh per-image noise of the large shape
    from the same seeds -> DDIM img2img from t_enc -> decode.  Returns (uint8 images, final latents)."""
    b = tokens.shape[0]
    h, w = height // 8, width // 8
    h2, w2 = int(height * hr_scale) // 8, int(width * hr_scale) // 8
    dsd = {k: v.to(device) for k, v in sd.items()}
    cond = clip_text_encode(dsd, clip_cfg, tokens.to(device))
    uncond = clip_text_encode(dsd, clip_cfg, neg_tokens.to(device))
    unet = lambda a, tt, c: unet_forward(dsd, unet_cfg, a, tt, c)  # noqa: E731
    x = sample_ddim(unet, per_image_noise(seed, b, (4, h, w)).to(device), cond, uncond, steps, cfg_scale)
    up = torch.nn.functional.interpolate(x, size=(h2, w2), mode="bilinear", antialias=False)
    noise = per_image_noise(seed, b, (4, h2, w2)).to(device)
    sa, s1a, rows = ddim_img2img_coefficients(hr_steps or steps, denoising_strength)
    x = up * sa + noise * s1a
    for (t, c_sa, c_s1a, c_sap, c_s1ap) in rows:
        e = cfg_eps(unet, x, t, cond, uncond, cfg_scale)
        x0 = (x - c_s1a * e) / c_sa
        x = c_sap * x0 + c_s1ap * e
    dec = vae_decode(dsd, vae_cfg, x / vae_cfg.scale_factor)
    return to_uint8(dec), x


def img2img(sd: SD, unet_cfg, vae_cfg, clip_cfg, tokens, neg_tokens, seed: int, init_u8: torch.Tensor,
            denoising_strength: float = 0.75, steps: int = 20, cfg_scale: float = 7.0, device="cpu"):
    """End-to-end img2img oracle: encode (posterior MEAN, see vae_encode_mean) -> noise to t_enc -> DDIM -> decode."""
    b = tokens.shape[0]
    dsd = {k: v.to(device) for k, v in sd.items()}
    cond = clip_text_encode(dsd, clip_cfg, tokens.to(device))
    uncond = clip_text_encode(dsd, clip_cfg, neg_tokens.to(device))
    x_in = image_to_model_input(init_u8.to(device))
    init = vae_encode_mean(dsd, vae_cfg, x_in) * vae_cfg.scale_factor
    noise = per_image_noise(seed, b, tuple(init.shape[1:])).to(device)
    sa, s1a, rows = ddim_img2img_coefficients(steps, denoising_strength)
    x = init * sa + noise * s1a
    for (t, c_sa, c_s1a, c_sap, c_s1ap) in rows:
        e = cfg_eps(lambda a, tt, c: unet_forward(dsd, unet_cfg, a, tt, c), x, t, cond, uncond, cfg_scale)
        x0 = (x - c_s1a * e) / c_sa
        x = c_sap * x0 + c_s1ap * e
    dec = vae_decode(dsd, vae_cfg, x / vae_cfg.scale_factor)
    return to_uint8(dec), x, init


# ------------------------------------------------------------------------------------------------ inpainting
def inpaint_masks(mask_img, width: int, height: int, lat_h: int, lat_w: int, mask_blur: int = 4, invert: bool = False):
    """sdwui StableDiffusionProcessingImg2Img.init, "whole picture" branch (inpaint_full_res False), mask_round True:
    binary mask -> optional invert -> cv2.GaussianBlur along x then y (kernel 2*int(2.5*blur+0.5)+1, sigma = blur) ->
    resize to the image -> overlay mask = clip(2 * mask, 0, 255); latent mask = round(bicubic resize of the mask / 255).
    Returns (latmask [lat_h, lat_w] float tensor, overlay mask 'L' PIL image)."""
    import cv2
    import numpy as np
    from PIL import Image, ImageOps
    if mask_img.mode == "RGBA" and mask_img.getextrema()[-1] != (255, 255):
        m = mask_img.split()[-1].convert("L").point(lambda v: 255 if v > 128 else 0)
    else:
        m = mask_img.convert("L")
    if invert:
        m = ImageOps.invert(m)
    if mask_blur > 0:
        ksize = 2 * int(2.5 * mask_blur + 0.5) + 1
        a = cv2.GaussianBlur(np.array(m), (ksize, 1), mask_blur)
        m = Image.fromarray(cv2.GaussianBlur(a, (1, ksize), mask_blur))
    if m.size != (width, height):
        m = m.resize((width, height), resample=Image.LANCZOS)
    overlay = Image.fromarray(np.clip(np.array(m).astype(np.float32) * 2, 0, 255).astype(np.uint8))
    lat = np.array(m.convert("RGB").resize((lat_w, lat_h)), dtype=np.float32)[..., 0] / 255.0
    return torch.from_numpy(np.around(lat)), overlay


def img2img_inpaint(sd: SD, unet_cfg, vae_cfg, clip_cfg, tokens, neg_tokens, seed: int, init_u8: torch.Tensor, mask_img,
                    denoising_strength: float = 0.75, steps: int = 20, cfg_scale: float = 7.0, mask_blur: int = 4,
                    invert: bool = False, inpainting_fill: int = 1, device="cpu"):
    """DDIM inpainting as sdwui runs it (CFGDenoiserTimesteps.mask_before_denoising): before EVERY model call the region
    to keep is replaced by the clean init latent, x = x * nmask + init * mask; once more after sampling; decode; then
    apply_overlay pastes the original pixels back through the blurred mask.  Returns (uint8 images, final latents)."""
    import numpy as np
    from PIL import Image, ImageOps
    b, hh, ww = tokens.shape[0], init_u8.shape[1], init_u8.shape[2]
    sd = {k: v.to(device) for k, v in sd.items()}
    cond = clip_text_encode(sd, clip_cfg, tokens.to(device))
    uncond = clip_text_encode(sd, clip_cfg, neg_tokens.to(device))
    enc_in = init_u8
    if inpainting_fill == 0:   # modules/masking.py fill(): cascade of blurs of the surroundings, with the BLURRED mask
        from PIL import ImageFilter
        import cv2
        m = mask_img.convert("L") if not (mask_img.mode == "RGBA" and mask_img.getextrema()[-1] != (255, 255)) else \
            mask_img.split()[-1].convert("L").point(lambda v: 255 if v > 128 else 0)
        if invert:
            m = ImageOps.invert(m)
        if mask_blur > 0:
            ksize = 2 * int(2.5 * mask_blur + 0.5) + 1
            m = Image.fromarray(cv2.GaussianBlur(cv2.GaussianBlur(np.array(m), (ksize, 1), mask_blur), (1, ksize), mask_blur))
        filled = []
        for k in range(b):
            image = Image.fromarray(init_u8[k].numpy(), "RGB")
            mod = Image.new("RGBA", image.size)
            masked = Image.new("RGBa", image.size)
            masked.paste(image.convert("RGBA").convert("RGBa"), mask=ImageOps.invert(m))
            for radius, repeats in [(256, 1), (64, 1), (16, 2), (4, 4), (2, 2), (0, 1)]:
                blurred = masked.filter(ImageFilter.GaussianBlur(radius)).convert("RGBA")
                for _ in range(repeats):
                    mod.alpha_composite(blurred)
            filled.append(torch.from_numpy(np.array(mod.convert("RGB"))))
        enc_in = torch.stack(filled)
    init = vae_encode_mean(sd, vae_cfg, image_to_model_input(enc_in.to(device))) * vae_cfg.scale_factor
    latmask, overlay_mask = inpaint_masks(mask_img, ww, hh, init.shape[2], init.shape[3], mask_blur, invert)
    nmask = latmask[None, None].to(init.dtype).to(device)
    mask = 1.0 - nmask
    noise = per_image_noise(seed, b, tuple(init.shape[1:])).to(device)
    if inpainting_fill == 2:     # "latent noise"
        init = init * mask + noise * nmask
    elif inpainting_fill == 3:   # "latent nothing"
        init = init * mask
    sa, s1a, rows = ddim_img2img_coefficients(steps, denoising_strength)
    x = init * sa + noise * s1a
    for (t, c_sa, c_s1a, c_sap, c_s1ap) in rows:
        x = x * nmask + init * mask
        e = cfg_eps(lambda a, tt, c: unet_forward(sd, unet_cfg, a, tt, c), x, t, cond, uncond, cfg_scale)
        x0 = (x - c_s1a * e) / c_sa
        x = c_sap * x0 + c_s1ap * e
    x = x * nmask + init * mask
    gen = to_uint8(vae_decode(sd, vae_cfg, x / vae_cfg.scale_factor)).cpu()
    out = torch.empty_like(gen)
    inv = ImageOps.invert(overlay_mask.convert("L"))
    for k in range(b):
        image = Image.fromarray(init_u8[k].numpy(), "RGB")
        masked = Image.new("RGBa", (image.width, image.height))
        masked.paste(image.convert("RGBA").convert("RGBa"), mask=inv)
        img = Image.fromarray(gen[k].numpy(), "RGB").convert("RGBA")
        img.alpha_composite(masked.convert("RGBA"))
        out[k] = torch.from_numpy(np.array(img.convert("RGB")))
    return out, x



# ------------------------------------------------------------------------------------------------ "only masked" inpainting
def get_crop_region(mask, pad: int = 0):
    """sdwui modules/masking.py get_crop_region (numpy mask [h, w]): columns / rows that are entirely zero are trimmed"""
    h, w = mask.shape
    crop_left = 0
    for i in range(w):
        if not (mask[:, i] == 0).all():
            break
        crop_left += 1
    crop_right = 0
    for i in reversed(range(w)):
        if not (mask[:, i] == 0).all():
            break
        crop_right += 1
    crop_top = 0
    for i in range(h):
        if not (mask[i] == 0).all():
            break
        crop_top += 1
    crop_bottom = 0
    for i in reversed(range(h)):
        if not (mask[i] == 0).all():
            break
        crop_bottom += 1
    return (int(max(crop_left - pad, 0)), int(max(crop_top - pad, 0)), int(min(w - crop_right + pad, w)),
            int(min(h - crop_bottom + pad, h)))


def expand_crop_region(crop_region, processing_width, processing_height, image_width, image_height):
    """sdwui modules/masking.py expand_crop_region"""
    x1, y1, x2, y2 = crop_region
    ratio_crop_region = (x2 - x1) / (y2 - y1)
    ratio_processing = processing_width / processing_height
    if ratio_crop_region > ratio_processing:
        desired_height = (x2 - x1) / ratio_processing
        desired_height_diff = int(desired_height - (y2 - y1))
        y1 -= desired_height_diff // 2
        y2 += desired_height_diff - desired_height_diff // 2
        if y2 >= image_height:
            diff = y2 - image_height
            y2 -= diff
            y1 -= diff
        if y1 < 0:
            y2 -= y1
            y1 -= y1
        if y2 >= image_height:
            y2 = image_height
    else:
        desired_width = (y2 - y1) * ratio_processing
        desired_width_diff = int(desired_width - (x2 - x1))
        x1 -= desired_width_diff // 2
        x2 += desired_width_diff - desired_width_diff // 2
        if x2 >= image_width:
            diff = x2 - image_width
            x2 -= diff
            x1 -= diff
        if x1 < 0:
            x2 -= x1
            x1 -= x1
        if x2 >= image_width:
            x2 = image_width
    return x1, y1, x2, y2


def resize_image(resize_mode: int, im, width: int, height: int):
    """sdwui modules/images.py resize_image without an upscaler: 0 just resize, 1 crop and resize, 2 resize and fill"""
    from PIL import Image
    if resize_mode == 0:
        return im.resize((width, height), resample=Image.LANCZOS)
    ratio = width / height
    src_ratio = im.width / im.height
    if resize_mode == 1:
        src_w = width if ratio > src_ratio else im.width * height // im.height
        src_h = height if ratio <= src_ratio else im.height * width // im.width
        resized = im.resize((src_w, src_h), resample=Image.LANCZOS)
        res = Image.new("RGB", (width, height))
        res.paste(resized, box=(width // 2 - src_w // 2, height // 2 - src_h // 2))
        return res
    src_w = width if ratio < src_ratio else im.width * height // im.height
    src_h = height if ratio >= src_ratio else im.height * width // im.width
    resized = im.resize((src_w, src_h), resample=Image.LANCZOS)
    res = Image.new("RGB", (width, height))
    res.paste(resized, box=(width // 2 - src_w // 2, height // 2 - src_h // 2))
    if ratio < src_ratio:
        fill_height = height // 2 - src_h // 2
        if fill_height > 0:
            res.paste(resized.resize((width, fill_height), box=(0, 0, width, 0)), box=(0, 0))
            res.paste(resized.resize((width, fill_height), box=(0, resized.height, width, resized.height)),
                      box=(0, fill_height + src_h))
    elif ratio > src_ratio:
        fill_width = width // 2 - src_w // 2
        if fill_width > 0:
            res.paste(resized.resize((fill_width, height), box=(0, 0, 0, height)), box=(0, 0))
            res.paste(resized.resize((fill_width, height), box=(resized.width, 0, resized.width, height)),
                      box=(fill_width + src_w, 0))
    return res


def only_masked_setup(init_image, mask_img, width: int, height: int, lat_h: int, lat_w: int, mask_blur: int = 4,
                      invert: bool = False, padding: int = 32):
    """sdwui StableDiffusionProcessingImg2Img.init with inpaint_full_res ("Only masked"): returns
    (crop_region, paste_to, processing-size init image (PIL RGB), latent mask tensor [lat_h, lat_w], overlay image RGBA)."""
    import cv2
    import numpy as np
    from PIL import Image, ImageOps
    if mask_img.mode == "RGBA" and mask_img.getextrema()[-1] != (255, 255):
        image_mask = mask_img.split()[-1].convert("L").point(lambda v: 255 if v > 128 else 0)
    else:
        image_mask = mask_img.convert("L")
    if invert:
        image_mask = ImageOps.invert(image_mask)
    if mask_blur > 0:
        ks = 2 * int(2.5 * mask_blur + 0.5) + 1
        image_mask = Image.fromarray(cv2.GaussianBlur(np.array(image_mask), (ks, 1), mask_blur))
        image_mask = Image.fromarray(cv2.GaussianBlur(np.array(image_mask), (1, ks), mask_blur))
    mask_for_overlay = image_mask
    mask = image_mask.convert("L")
    crop_region = get_crop_region(np.array(mask), padding)
    crop_region = expand_crop_region(crop_region, width, height, mask.width, mask.height)
    x1, y1, x2, y2 = crop_region
    image_mask = resize_image(2, mask.crop(crop_region), width, height)
    paste_to = (x1, y1, x2 - x1, y2 - y1)
    image = init_image.convert("RGB")
    image_masked = Image.new("RGBa", (image.width, image.height))
    image_masked.paste(image.convert("RGBA").convert("RGBa"), mask=ImageOps.invert(mask_for_overlay.convert("L")))
    overlay = image_masked.convert("RGBA")
    image = resize_image(2, image.crop(crop_region), width, height)
    latmask = image_mask.convert("RGB").resize((lat_w, lat_h))
    latmask = np.moveaxis(np.array(latmask, dtype=np.float32), 2, 0) / 255
    latmask = np.around(latmask[0])
    return crop_region, paste_to, image, torch.from_numpy(latmask), overlay


def apply_overlay(image, paste_loc, overlay):
    """sdwui processing.py apply_overlay"""
    from PIL import Image
    if paste_loc is not None:
        x, y, w, h = paste_loc
        base_image = Image.new("RGBA", (overlay.width, overlay.height))
        image = resize_image(1, image, w, h)
        base_image.paste(image, (x, y))
        image = base_image
    image = image.convert("RGBA")
    image.alpha_composite(overlay)
    return image.convert("RGB")
