"""Achieved HBM bandwidth of the GroupNorm (statistics + apply) and LayerNorm kernels at the UNet's shapes (bench batch:
64 = 32 images x CFG), and of GroupNorm issued per image chunk — a chunk whose tensor fits the 126 MB L2 is read from
HBM once (statistics) and from L2 the second time (apply).

    python tools/norm_sweep.py [chunk_images ...]      default chunks: 64 (one launch pair), 32, 16, 8
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-webui-distributed_b200"))
import torch  # noqa: E402

from b200sd import ops  # noqa: E402

GN_SHAPES = [(64, 4096, 320), (64, 4096, 640), (64, 4096, 960), (64, 1024, 640), (64, 1024, 1280), (64, 1024, 1920),
             (64, 256, 1280), (64, 256, 2560), (64, 64, 1280), (64, 64, 2560)]
LN_SHAPES = [(64 * 4096, 320), (64 * 1024, 640), (64 * 256, 1280)]


def timed(fn, reps=10):
    for _ in range(2):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    chunks = [int(a) for a in sys.argv[1:]] or [64, 32, 16, 8]
    dev = "cuda"
    print("GroupNorm (+SiLU): shape, then per chunking: ms and GB/s of algorithmic traffic (6 B/element)")
    tot = {c: 0.0 for c in chunks}
    for nb, hw, c in GN_SHAPES:
        bufs = [torch.randn((nb, hw, c), device=dev).half() for _ in range(3)]
        outs = [torch.empty_like(b) for b in bufs]
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        stats = torch.zeros(ops.groupnorm_stats_floats(nb, hw, c, 32), device=dev)
        line = f"  [{nb},{hw},{c}] {nb * hw * c * 2 / 1e6:7.1f} MB:"
        for ch in chunks:
            def run(i, ch=ch):
                x, y = bufs[i % 3], outs[i % 3]
                for b0 in range(0, nb, ch):
                    ops.groupnorm(x[b0:b0 + ch], y[b0:b0 + ch], stats, gamma, beta, 32, 1e-5, True)
            ms = timed(run)
            tot[ch] += ms
            line += f"  chunk {ch:2d}: {ms:6.3f} ms {nb * hw * c * 6 / ms / 1e6:6.0f} GB/s"
        print(line)
        del bufs, outs
    print("  sum over these shapes: " + "  ".join(f"chunk {c}: {t:.3f} ms" for c, t in tot.items()))
    print("LayerNorm: rows x C, ms, GB/s (4 B/element)")
    for rows, c in LN_SHAPES:
        bufs = [torch.randn((rows, c), device=dev).half() for _ in range(3)]
        outs = [torch.empty_like(b) for b in bufs]
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        ms = timed(lambda i: ops.layernorm(bufs[i % 3], outs[i % 3], gamma, beta))
        print(f"  [{rows},{c}] {rows * c * 2 / 1e6:7.1f} MB: {ms:6.3f} ms {rows * c * 4 / ms / 1e6:6.0f} GB/s")
        del bufs, outs


if __name__ == "__main__":
    main()
