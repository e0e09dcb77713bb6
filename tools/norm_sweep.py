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
# every GroupNorm site of the SD1.5 UNet: (HW, C, sites) — 61 per evaluation
GN_SITES = [(4096, 320, 13), (4096, 640, 2), (4096, 960, 1), (1024, 320, 1), (1024, 640, 11), (1024, 960, 1), (1024, 1280, 1),
            (1024, 1920, 1), (256, 640, 1), (256, 1280, 11), (256, 1920, 1), (256, 2560, 2), (64, 1280, 12), (64, 2560, 3)]
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


def gn_modes(argv):
    """--gn-modes: statistics + apply (mode 1) against the one-pass kernel (mode 2) at several slab sizes, one process"""
    from b200sd import _lib
    dev = "cuda"
    kbs = [int(a) for a in argv] or [48]
    # (label, mode, slab KB of the one-pass kernel, statistics kernel walks the tensor back to front)
    cols = [("2 kernels fwd", 1, None, 0), ("2 kernels rev", 1, None, 1)] + [(f"1-pass {kb} KB", 2, kb, 1) for kb in kbs]
    tot = {c[0]: 0.0 for c in cols}
    print("GroupNorm (+SiLU), UNet batch 64: ms per call (GB/s at 4 B/element = 1 read + 1 write); the sum weights every "
          "shape by its number of sites in one SD1.5 UNet evaluation")
    for hw, c, sites in GN_SITES:
        nb = 64
        bufs = [torch.randn((nb, hw, c), device=dev).half() for _ in range(3)]
        outs = [torch.empty_like(b) for b in bufs]
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        line = f"  [{nb},{hw},{c}] {nb * hw * c * 2 / 1e6:7.1f} MB:"
        for label, mode, kb, rev in cols:
            _lib.check(_lib.lib().b200sd_debug_gn_config(kb or 0, rev), "gn config")
            if kb is not None:
                ops._GN_FLOATS.clear()
                ops._GN_FUSED.clear()
            if mode == 2 and not ops.groupnorm_is_fused(nb, hw, c, 32, torch.float16):
                line += f"  {label}: n/a"
                mode = 1   # what the executor would run
            stats = torch.zeros(ops.groupnorm_stats_floats(nb, hw, c, 32), device=dev)

            def run(i, mode=mode, stats=stats):
                ops.groupnorm(bufs[i % 3], outs[i % 3], stats, gamma, beta, 32, 1e-5, True, mode=mode)
            ms = timed(run)
            tot[label] += ms * sites
            line += f"  {label}: {ms:6.3f} ms {nb * hw * c * 4 / ms / 1e6:5.0f}"
        print(line + f"   x{sites}", flush=True)
        del bufs, outs
    print("  per UNet evaluation (61 sites): " + "  ".join(f"{k}: {t:.3f} ms" for k, t in tot.items()))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--gn-modes":
        return gn_modes(sys.argv[2:])
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
