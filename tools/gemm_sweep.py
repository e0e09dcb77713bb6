"""Sensitivity sweep of the persistent tcgen05 GEMM / implicit-GEMM conv at the UNet's dominant shapes: operand-ring
depth (B200SD_GEMM_STAGES), CTA-pair mode (B200SD_PAIR) and tile width.  One child process per environment, because the
library reads its knobs once.

    python tools/gemm_sweep.py            # all configurations
    python tools/gemm_sweep.py --one      # the configuration of the current environment
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-webui-distributed_b200"))


def timeit(fn, n=10):
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


def one():
    import torch
    from b200sd import ops
    nb = int(os.environ.get("SWEEP_NB", "32"))
    tag = f"stages={os.environ.get('B200SD_GEMM_STAGES', '-')} pair={os.environ.get('B200SD_PAIR', '-')}"

    def rnd(*shape, scale=1.0):
        return (torch.randn(shape, device="cuda") * scale).half()

    m = nb * 4096
    out = []
    # 3x3 conv 320 -> 320 at 64x64 (the resnet-block conv of the top level)
    x, w, o = rnd(nb, 64, 64, 320), rnd(320, 2880, scale=0.02), torch.empty((m, 320), device="cuda", dtype=torch.half)
    b = torch.randn(320, device="cuda")
    for bn in (160, 64):
        us = timeit(lambda: ops.conv2d(x, w, o, ksize=3, bias=b, block_n=bn))
        out.append(f"conv3x3_320 bn{bn}: {us:8.1f} us {2 * m * 320 * 2880 / us / 1e6:7.1f} TF/s")
    # 3x3 conv 640 -> 640 at 32x32, 1280 -> 1280 at 16x16
    for c, hw in ((640, 32), (1280, 16)):
        xx, ww = rnd(nb, hw, hw, c), rnd(c, 9 * c, scale=0.02)
        oo = torch.empty((nb * hw * hw, c), device="cuda", dtype=torch.half)
        for bn in ((160, 128) if c == 640 else (256, 160, 128)):
            us = timeit(lambda: ops.conv2d(xx, ww, oo, ksize=3, block_n=bn))
            out.append(f"conv3x3_{c} bn{bn}: {us:8.1f} us {2 * nb * hw * hw * c * 9 * c / us / 1e6:7.1f} TF/s")
    # linears at 64x64 tokens: K = 320
    a = rnd(m, 320)
    for n, bns in ((320, (160, 64)), (1536, (256, 192, 128))):
        ww, oo = rnd(n, 320, scale=0.05), torch.empty((m, n), device="cuda", dtype=torch.half)
        for bn in bns:
            us = timeit(lambda: ops.linear(a, ww, oo, block_n=bn))
            out.append(f"linear K320 N{n} bn{bn}: {us:8.1f} us {2 * m * n * 320 / us / 1e6:7.1f} TF/s")
    # proj with bias + residual, GEGLU, FF2 (K = 1280)
    ww, oo, r = rnd(320, 320, scale=0.05), torch.empty((m, 320), device="cuda", dtype=torch.half), rnd(m, 320)
    us = timeit(lambda: ops.linear(a, ww, oo, bias=b, residual=r))
    out.append(f"proj K320 N320 +res: {us:8.1f} us {2 * m * 320 * 320 / us / 1e6:7.1f} TF/s  {3 * m * 640 / us / 1e6:6.2f} TB/s")
    wg, og, bg = rnd(2560, 320, scale=0.05), torch.empty((m, 1280), device="cuda", dtype=torch.half), torch.randn(2560, device="cuda")
    for bn in (256, 128):
        us = timeit(lambda: ops.linear(a, wg, og, bias=bg, flags=ops.EPI_GEGLU, block_n=bn))
        out.append(f"geglu K320 N2560 bn{bn}: {us:8.1f} us {2 * m * 2560 * 320 / us / 1e6:7.1f} TF/s")
    w2 = rnd(320, 1280, scale=0.03)
    us = timeit(lambda: ops.linear(og, w2, oo, bias=b, residual=r))
    out.append(f"ff2 K1280 N320 +res: {us:8.1f} us {2 * m * 320 * 1280 / us / 1e6:7.1f} TF/s")
    print(f"[{tag} nb={nb}]\n  " + "\n  ".join(out), flush=True)


def main():
    if "--one" in sys.argv:
        return one()
    base = {k: v for k, v in os.environ.items() if not k.startswith("B200SD_")}
    for env in ({}, {"B200SD_GEMM_STAGES": "3"}, {"B200SD_GEMM_STAGES": "2"}, {"B200SD_PAIR": "1"},
                {"B200SD_PAIR": "1", "B200SD_GEMM_STAGES": "3"}):
        subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=dict(base, **env), check=False, timeout=300)


if __name__ == "__main__":
    main()
