"""Time the self-attention kernel at its dominant UNet shape (batch 16, 4096 tokens, 8 heads, d = 40) — one process per
tuning configuration, because the library reads B200SD_ATTN_* once.

    python tools/attn_sweep.py            # run every configuration in a child process, print one line each
    python tools/attn_sweep.py --one      # time the configuration given by the current environment
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-webui-distributed_b200"))

CONFIGS = [  # (S buffers, P buffers, "k,v" ring depths)
    ("3", "3", "3,2"), ("3", "3", "2,2"), ("3", "2", "3,2"), ("3", "2", "2,2"), ("2", "2", "2,2"), ("2", "3", "2,2"),
    ("3", "2", "4,3"),
]


def one():
    import torch
    from b200sd import ops
    nb = int(os.environ.get("ATTN_NB", "16"))
    qkv = torch.zeros((nb, 4096, 3, 8, 64), device="cuda", dtype=torch.half)
    qkv[..., :40] = (torch.randn((nb, 4096, 3, 8, 40), device="cuda")).half()
    qkv[:, :, 2, :, 40] = 1.0  # ones column of V -> row sums from the P.V MMA (as the UNet uses it)
    flat = qkv.reshape(nb, 4096, 1536)
    q, k, v = flat[..., :512], flat[..., 512:1024], flat[..., 1024:]
    o = torch.empty((nb, 4096, 320), device="cuda", dtype=torch.half)
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
    qq, kk, vv = (t[0].reshape(4096, 8, 64)[:, 3, :40].float() for t in (q, k, v))
    ref = torch.softmax(qq @ kk.T * 40 ** -0.5, dim=-1) @ vv
    err = float((o[0].reshape(4096, 8, 40)[:, 3].float() - ref).abs().max())
    exps = nb * 8 * 4096 * 4096
    print(f"sbufs={os.environ.get('B200SD_ATTN_SBUFS', '-')} pbufs={os.environ.get('B200SD_ATTN_PBUFS', '-')} ring={os.environ.get('B200SD_ATTN_RING', '-')} "
          f"nb={nb}: {ms:.4f} ms  ({exps / ms / 1e9:.2f} Texp/s)  max_err={err:.2e}", flush=True)


def main():
    if "--one" in sys.argv:
        return one()
    for sb, pb, ring in CONFIGS:
        env = dict(os.environ, B200SD_ATTN_SBUFS=sb, B200SD_ATTN_PBUFS=pb, B200SD_ATTN_RING=ring)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, check=False, timeout=300)
    subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env={k: v for k, v in os.environ.items()
                                                                              if not k.startswith("B200SD_ATTN")},
                   check=False, timeout=300)


if __name__ == "__main__":
    main()
