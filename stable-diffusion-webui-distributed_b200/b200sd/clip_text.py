"""CLIP ViT-L/14 text tower in plain torch (fp16 on the GPU).

Runs once per request (cond + uncond prompts), 0.04 % of an image's FLOPs — SURVEY.md §8 a-ext x13 keeps it out of
kernel scope ("run in torch").  Same math as transformers' CLIPTextModel: causal mask, quick-gelu MLP, final LN,
last hidden state (sdwui CLIP_stop_at_last_layers = 1).
"""
from typing import Dict

import torch
import torch.nn.functional as F

from .config import CLIP_PREFIX, CLIPConfig


class ClipText:
    def __init__(self, sd: Dict[str, torch.Tensor], cfg: CLIPConfig, device, dtype=torch.float16):
        self.cfg, self.device, self.dtype = cfg, device, dtype
        n = len(CLIP_PREFIX)
        self.w = {k[n:]: v.to(device=device, dtype=dtype) for k, v in sd.items() if k.startswith(CLIP_PREFIX)}

    @torch.no_grad()
    def __call__(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens int64 [B, 77] -> [B, 77, width]"""
        w, cfg = self.w, self.cfg
        tokens = tokens.to(self.device)
        b, n = tokens.shape
        x = w["embeddings.token_embedding.weight"][tokens] + w["embeddings.position_embedding.weight"][None, :n]
        mask = torch.full((n, n), float("-inf"), device=self.device, dtype=torch.float32).triu(1)
        d = cfg.width // cfg.heads
        for i in range(cfg.layers):
            p = f"encoder.layers.{i}"
            h = F.layer_norm(x, (cfg.width,), w[p + ".layer_norm1.weight"], w[p + ".layer_norm1.bias"], 1e-5)
            q, k, v = (F.linear(h, w[f"{p}.self_attn.{m}_proj.weight"], w[f"{p}.self_attn.{m}_proj.bias"])
                       .reshape(b, n, cfg.heads, d).permute(0, 2, 1, 3) for m in "qkv")
            att = torch.softmax((q @ k.transpose(-1, -2)).float() * d ** -0.5 + mask, dim=-1).to(x.dtype) @ v
            att = att.permute(0, 2, 1, 3).reshape(b, n, cfg.width)
            x = x + F.linear(att, w[p + ".self_attn.out_proj.weight"], w[p + ".self_attn.out_proj.bias"])
            h = F.layer_norm(x, (cfg.width,), w[p + ".layer_norm2.weight"], w[p + ".layer_norm2.bias"], 1e-5)
            h = F.linear(h, w[p + ".mlp.fc1.weight"], w[p + ".mlp.fc1.bias"])
            h = h * torch.sigmoid(1.702 * h)
            x = x + F.linear(h, w[p + ".mlp.fc2.weight"], w[p + ".mlp.fc2.bias"])
        return F.layer_norm(x, (cfg.width,), w["final_layer_norm.weight"], w["final_layer_norm.bias"], 1e-5)
