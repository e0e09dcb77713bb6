"""Model hyper-parameters and block topology (ldm UNetModel / kl-f8 AutoencoderKL / CLIP text tower).

Mirrors upstream ctor arguments (v1-inference.yaml); key names produced by `unet_layout` follow the ldm
state_dict so real checkpoints map 1:1.
"""
from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    model_channels: int = 320
    channel_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    attention_levels: Tuple[int, ...] = (0, 1, 2)
    num_heads: int = 8
    context_dim: int = 768
    transformer_depth: int = 1

    @property
    def time_embed_dim(self) -> int:
        return 4 * self.model_channels


@dataclass(frozen=True)
class VAEConfig:
    ch: int = 128
    ch_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    z_channels: int = 4
    out_ch: int = 3
    scale_factor: float = 0.18215


@dataclass(frozen=True)
class CLIPConfig:
    vocab: int = 49408
    width: int = 768
    layers: int = 12
    heads: int = 12
    ctx: int = 77


SD15_UNET = UNetConfig()
SD15_VAE = VAEConfig()
SD15_CLIP = CLIPConfig()
# same topology, reduced width: fast tests (every channel count stays a multiple of 64)
TINY_UNET = UNetConfig(model_channels=64, num_heads=2, context_dim=64)
TINY_VAE = VAEConfig(ch=64, ch_mult=(1, 2), num_res_blocks=1)
TINY_CLIP = CLIPConfig(vocab=1000, width=64, layers=2, heads=2)

UNET_PREFIX = "model.diffusion_model."
VAE_PREFIX = "first_stage_model."
CLIP_PREFIX = "cond_stage_model.transformer.text_model."


def unet_layout(cfg: UNetConfig):
    """(input_blocks, middle_block, output_blocks); a block is a list of layer tuples:
    ('conv_in', cin, cout) | ('res', cin, cout) | ('attn', c) | ('down', c) | ('up', c)."""
    mc = cfg.model_channels
    inputs = [[("conv_in", cfg.in_channels, mc)]]
    skip_ch = [mc]
    ch = mc
    last = len(cfg.channel_mult) - 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("res", ch, mult * mc)]
            ch = mult * mc
            if level in cfg.attention_levels:
                layers.append(("attn", ch))
            inputs.append(layers)
            skip_ch.append(ch)
        if level != last:
            inputs.append([("down", ch)])
            skip_ch.append(ch)
    middle = [("res", ch, ch), ("attn", ch), ("res", ch, ch)]
    outputs = []
    for level in range(last, -1, -1):
        mult = cfg.channel_mult[level]
        for i in range(cfg.num_res_blocks + 1):
            layers = [("res", ch + skip_ch.pop(), mult * mc)]
            ch = mult * mc
            if level in cfg.attention_levels:
                layers.append(("attn", ch))
            if level > 0 and i == cfg.num_res_blocks:
                layers.append(("up", ch))
            outputs.append(layers)
    return inputs, middle, outputs
