"""kl-f8 VAE decoder as a static program of libb200sd kernels (NHWC fp16).

Stands in for upstream `AutoencoderKL.decode` (ldm/modules/diffusionmodules/model.py::Decoder) + sdwui's
clamp/255/uint8 conversion — the "final VAE decode" of the north star (SURVEY.md §8 a-ext x11).

The single-head d=C mid-block attention is expressed with the GEMM kernel: per image S = q k^T, row softmax,
V^T = Wv h^T (so P.V needs no transpose), O = P V^T^T; the v bias is folded into proj_out's bias (softmax rows sum
to one, so P (V0 + 1 bv^T) = P V0 + bv exactly).
"""
from typing import Dict, List

import torch

from . import ops
from .config import VAE_PREFIX, VAEConfig
from .unet_exec import Pool
from .weights import pack_conv


class VAEDecoderWeights:
    def __init__(self, sd: Dict[str, torch.Tensor], cfg: VAEConfig, device, dtype=torch.float16):
        self.cfg, self.device, self.dtype = cfg, device, dtype
        self.t: Dict[str, torch.Tensor] = {}
        p = VAE_PREFIX

        def dev(t, dt=None):
            return t.to(device=device, dtype=dt or dtype).contiguous()

        def conv(name, key, cin_pad=0, cout_pad=0):
            self.t[name + ".w"] = dev(pack_conv(sd[p + key + ".weight"], cin_pad, cout_pad))
            b = sd[p + key + ".bias"]
            if cout_pad > b.numel():
                b = torch.cat([b, b.new_zeros(cout_pad - b.numel())])
            self.t[name + ".b"] = dev(b, torch.float32)

        def norm(name, key):
            self.t[name + ".g"] = dev(sd[p + key + ".weight"], torch.float32)
            self.t[name + ".beta"] = dev(sd[p + key + ".bias"], torch.float32)

        def res(name, key, cin, cout):
            norm(name + ".gn1", key + ".norm1")
            conv(name + ".conv1", key + ".conv1")
            norm(name + ".gn2", key + ".norm2")
            conv(name + ".conv2", key + ".conv2")
            if cin != cout:
                conv(name + ".skip", key + ".nin_shortcut")

        # post_quant_conv: 1x1 z->z, both sides padded to 64 channels
        conv("post_quant", "post_quant_conv", cin_pad=64, cout_pad=64)
        nlev = len(cfg.ch_mult)
        cin = cfg.ch * cfg.ch_mult[-1]
        conv("conv_in", "decoder.conv_in", cin_pad=64)
        res("mid1", "decoder.mid.block_1", cin, cin)
        norm("attn.norm", "decoder.mid.attn_1.norm")
        for n in ("q", "k", "v", "proj_out"):
            conv("attn." + n, "decoder.mid.attn_1." + n)
        wp = sd[p + "decoder.mid.attn_1.proj_out.weight"].reshape(cin, cin).double()
        bv = sd[p + "decoder.mid.attn_1.v.bias"].double()
        self.t["attn.proj_out.b"] = dev((sd[p + "decoder.mid.attn_1.proj_out.bias"].double() + wp @ bv).float(),
                                        torch.float32)
        res("mid2", "decoder.mid.block_2", cin, cin)
        self.levels = []
        for lvl in reversed(range(nlev)):
            cout = cfg.ch * cfg.ch_mult[lvl]
            blocks = []
            for b in range(cfg.num_res_blocks + 1):
                res(f"up{lvl}.{b}", f"decoder.up.{lvl}.block.{b}", cin, cout)
                blocks.append((cin, cout))
                cin = cout
            if lvl != 0:
                conv(f"up{lvl}.upsample", f"decoder.up.{lvl}.upsample.conv")
            self.levels.append((lvl, blocks))
        norm("norm_out", "decoder.norm_out")
        conv("conv_out", "decoder.conv_out", cout_pad=32)
        self.mid_ch = cfg.ch * cfg.ch_mult[-1]
        self.out_ch = cin


class VAEDecoderProgram:
    """Decode `b` latents of size h x w -> uint8 [b, 8h*.., 3].  run() is allocation- and sync-free."""

    def __init__(self, w: VAEDecoderWeights, b: int, h: int, wd: int):
        self.w, self.b, self.h, self.wd = w, b, h, wd
        self.dev, self.dt = w.device, w.dtype
        self.pool = Pool(self.dev, self.dt)
        self.zin = torch.zeros((2 * b, h * wd, 64), device=self.dev, dtype=self.dt)  # pack_unet_input writes both halves
        self.ops: List = []
        self.gn_stats: List = []
        self._build()
        self.stats_all = torch.zeros((len(self.gn_stats), b, 32, 2), device=self.dev, dtype=torch.float32)
        for i, holder in enumerate(self.gn_stats):
            holder[0] = self.stats_all[i]

    def _emit(self, fn, *a, **k):
        self.ops.append((fn, a, k))

    def _gn(self, x, out, name, silu):
        holder = [None]
        self.gn_stats.append(holder)
        g, beta = self.w.t[name + ".g"], self.w.t[name + ".beta"]
        self._emit(lambda: ops.groupnorm(x, out, holder[0], g, beta, 32, 1e-6, silu))

    def _res(self, name, x, cin, cout, h, wd):
        b, hw, t = self.b, h * wd, self.w.t
        a = self.pool.get(b, hw, cin)
        self._gn(x, a, name + ".gn1", True)
        c1 = self.pool.get(b, hw, cout)
        self._emit(ops.conv2d, a.unflatten(1, (h, wd)), t[name + ".conv1.w"], c1, ksize=3, bias=t[name + ".conv1.b"])
        self.pool.put(a)
        a2 = self.pool.get(b, hw, cout)
        self._gn(c1, a2, name + ".gn2", True)
        self.pool.put(c1)
        if cin != cout:
            s = self.pool.get(b, hw, cout)
            self._emit(ops.linear, x, t[name + ".skip.w"], s, bias=t[name + ".skip.b"])
        else:
            s = x
        out = self.pool.get(b, hw, cout)
        self._emit(ops.conv2d, a2.unflatten(1, (h, wd)), t[name + ".conv2.w"], out, ksize=3, bias=t[name + ".conv2.b"],
                   residual=s)
        self.pool.put(a2)
        if s is not x:
            self.pool.put(s)
        return out

    def _attn(self, x, c, h, wd):
        b, s, t = self.b, h * wd, self.w.t
        hn = self.pool.get(b, s, c)
        self._gn(x, hn, "attn.norm", False)
        q = self.pool.get(b, s, c)
        k = self.pool.get(b, s, c)
        self._emit(ops.linear, hn, t["attn.q.w"], q, bias=t["attn.q.b"])
        self._emit(ops.linear, hn, t["attn.k.w"], k, bias=t["attn.k.b"])
        o = self.pool.get(b, s, c)
        scores = self.pool.get(s, s)
        vt = self.pool.get(c, s)
        for i in range(b):
            self._emit(ops.linear, q[i], k[i], scores)                  # S = q k^T        [s, s]
            self._emit(ops.softmax_rows_, scores, float(c) ** -0.5)     # softmax(S * c^-1/2)
            self._emit(ops.linear, t["attn.v.w"], hn[i], vt)            # V^T = Wv h^T      [c, s]
            self._emit(ops.linear, scores, vt, o[i])                    # O = P V           [s, c]
        out = self.pool.get(b, s, c)
        self._emit(ops.linear, o, t["attn.proj_out.w"], out, bias=t["attn.proj_out.b"], residual=x)
        for tmp in (hn, q, k, o, scores, vt):
            self.pool.put(tmp)
        return out

    def _build(self):
        b, h, wd, t = self.b, self.h, self.wd, self.w.t
        c = self.w.mid_ch
        z = self.pool.get(b, h * wd, 64)
        self._emit(ops.linear, self.zin[:b], t["post_quant.w"], z, bias=t["post_quant.b"])
        x = self.pool.get(b, h * wd, c)
        self._emit(ops.conv2d, z.unflatten(1, (h, wd)), t["conv_in.w"], x, ksize=3, bias=t["conv_in.b"])
        self.pool.put(z)
        for stage in ("mid1", "attn", "mid2"):
            y = self._attn(x, c, h, wd) if stage == "attn" else self._res(stage, x, c, c, h, wd)
            self.pool.put(x)
            x = y
        for lvl, blocks in self.w.levels:
            for i, (cin, cout) in enumerate(blocks):
                y = self._res(f"up{lvl}.{i}", x, cin, cout, h, wd)
                self.pool.put(x)
                x = y
                c = cout
            if lvl != 0:
                up = self.pool.get(b, 4 * h * wd, c)
                self._emit(ops.upsample2x, x.unflatten(1, (h, wd)), up.unflatten(1, (2 * h, 2 * wd)))
                self.pool.put(x)
                h, wd = 2 * h, 2 * wd
                x = self.pool.get(b, h * wd, c)
                self._emit(ops.conv2d, up.unflatten(1, (h, wd)), t[f"up{lvl}.upsample.w"], x, ksize=3,
                           bias=t[f"up{lvl}.upsample.b"])
                self.pool.put(up)
        a = self.pool.get(b, h * wd, c)
        self._gn(x, a, "norm_out", True)
        self.pool.put(x)
        self.img = torch.empty((b, h * wd, 32), device=self.dev, dtype=self.dt)   # RGB in channels 0..2, in [-1, 1]
        self._emit(ops.conv2d, a.unflatten(1, (h, wd)), t["conv_out.w"], self.img, ksize=3, bias=t["conv_out.b"])
        self.pool.put(a)
        self.out_h, self.out_w = h, wd
        self.u8 = torch.empty((b, h * wd, 3), device=self.dev, dtype=torch.uint8)
        self._emit(ops.quantize_u8, self.img, self.u8)

    def set_latents(self, x: torch.Tensor, scale_factor: float):
        """x fp32 [b, h*w, 4] (scaled latents): writes x / scale_factor into the padded fp16 input."""
        ops.pack_unet_input(x, self.zin, 1.0 / scale_factor)

    def run(self):
        self.stats_all.zero_()
        for fn, a, k in self.ops:
            fn(*a, **k)
        return self.u8
