"""Time the self-attention kernel at its dominant UNet shape (batch 16, 4096 tokens, 8 heads, d = 40).

    python tools/attn_sweep.py            # the default shape plus kv lengths 1024 / 2048 / 8192, one child process each
    python tools/attn_sweep.py --one      # one timing with the current environment (ATTN_NB, ATTN_SKV)
    ATTN_SKV=8192 python tools/attn_sweep.py --one    # longer kv (per-CTA start-up cost shows as time/kv-tile)
Build-time experiment knobs go through B200SD_NVCC_EXTRA (e.g. -DB200SD_ATTN_POLY_STRIDE=4, -DB200SD_WAIT_HINT_NS=...).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-webui-distributed_b200"))



def one():
    import torch
    from b200sd import ops
    nb = int(os.environ.get("ATTN_NB", "16"))
    sq = 4096
    skv = int(os.environ.get("ATTN_SKV", "4096"))
    tok = max(sq, skv)
    qkv = torch.zeros((nb, tok, 3, 8, 64), device="cuda", dtype=torch.half)
    qkv[..., :40] = (torch.randn((nb, tok, 3, 8, 40), device="cuda")).half()
    qkv[:, :, 2, :, 40] = 1.0  # ones column of V -> row sums from the P.V MMA (as the UNet uses it)
    flat = qkv.reshape(nb, tok, 1536)
    q, k, v = flat[:, :sq, :512], flat[:, :skv, 512:1024], flat[:, :skv, 1024:]
    if skv != tok or sq != tok:  # the op wants batch stride == rows * pitch
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    o = torch.empty((nb, sq, 320), device="cuda", dtype=torch.half)
    for _ in range(3):
        ops.attention(q, k, v, o, 8, 40, 64, 40 ** -0.5, v_ones_col=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        ops.attention(q, k, v, o, 8, 40, 64, 40 ** -0.5, v_ones_col=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    # reference check against torch on one (batch, head)
    qq, kk, vv = (t[0].reshape(-1, 8, 64)[:, 3, :40].float() for t in (q, k, v))
    ref = torch.softmax(qq @ kk.T * 40 ** -0.5, dim=-1) @ vv
    err = float((o[0].reshape(sq, 8, 40)[:, 3].float() - ref).abs().max())
    exps = nb * 8 * sq * skv
    print(f"nb={nb} skv={skv}: {ms:.4f} ms  ({exps / ms / 1e9:.2f} Texp/s)  max_err={err:.2e}", flush=True)


def main():
    if "--one" in sys.argv:
        return one()
    base = {k: v for k, v in os.environ.items() if not k.startswith("B200SD_ATTN")}
    me = [sys.executable, os.path.abspath(__file__), "--one"]
    for skv in ("1024", "2048", "8192"):
        subprocess.run(me, env=dict(base, ATTN_SKV=skv), check=False, timeout=300)
    subprocess.run(me, env=base, check=False, timeout=300)


if __name__ == "__main__":
    main()
