"""Tile-width probe for the q|k|v / q projection shapes of the narrow head pitch (round 2): time ops.linear per block_n."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-webui-distributed_b200"))
import torch  # noqa: E402

from b200sd import ops  # noqa: E402


def t(m, n, k, bn, reps=20):
    a = (torch.randn((m, k), device="cuda") * 0.5).half()
    w = (torch.randn((n, k), device="cuda") * 0.05).half()
    o = torch.empty((m, n), device="cuda", dtype=torch.half)
    b = torch.zeros(n, device="cuda")
    for _ in range(3):
        ops.linear(a, w, o, bias=b, block_n=bn)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.linear(a, w, o, bias=b, block_n=bn)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for m, n, k in ((262144, 1152, 320), (262144, 384, 320), (65536, 2304, 640), (65536, 768, 640), (16384, 4224, 1280), (16384, 1408, 1280),
                (262144, 320, 320), (65536, 640, 640), (16384, 1280, 1280)):
    cands = [bn for bn in (256, 192, 160, 128, 96, 64) if n % bn == 0]
    res = {bn: t(m, n, k, bn) for bn in cands}
    pick = ops.pick_block_n(n, False, m)
    best = min(res, key=res.get)
    print(f"M{m} N{n} K{k}: " + "  ".join(f"bn{bn} {us:7.1f}us" for bn, us in res.items()) + f"   heuristic {pick}  best {best}  ({res[pick] / res[best]:.3f}x)")
