"""Per-tile timeline of CTA 0 of the GEMM / conv kernel (clock64 stamps written by the kernel's debug hook): where does a
tile's time go between the TMA producer, the MMA thread and the two epilogue groups?  Needs a trace-enabled library build:
    B200SD_NVCC_EXTRA="-DB200SD_GEMM_TRACE_ENABLE=1" python stable-diffusion-webui-distributed_b200/b200sd/build.py --force

events (per tile processed by CTA 0):
  producer lane   0 tile start                 1 last k-block's loads issued
  MMA lane        2 wants the accumulator      3 accumulator free    4 first operands landed    5 last MMA + commit issued
  epilogue group g (thread 0 of the group), base e = 8 + 12 g:
                  e+0 waiting for accumulator  e+1 accumulator complete   e+2+2i chunk i in registers   e+3+2i chunk i staged,
                  group barrier passed         e+10 accumulator released
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-webui-distributed_b200"))
import torch  # noqa: E402

from b200sd import _lib, ops  # noqa: E402


def trace(name, fn, n_chunks_per_group):
    for _ in range(3):
        fn()
    buf = torch.zeros((32, 64), device="cuda", dtype=torch.int64)
    _lib.lib().b200sd_debug_gemm_trace(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    _lib.lib().b200sd_debug_gemm_trace(ctypes.c_void_p(0))
    t = buf.cpu().tolist()
    js = list(range(8, 28))

    def avg(f):
        v = [f(j) for j in js]
        return sum(v) / len(v)

    print(f"\n== {name}: {e0.elapsed_time(e1) * 1e3:.1f} us; steady state over tiles 8..27 of CTA 0, cycles")
    if t[5][10] == 0:
        print("   (no stamps: library not built with -DB200SD_GEMM_TRACE_ENABLE=1, or fewer than 28 tiles per CTA)")
        return
    print(f"  tile period (MMA commit to commit)            {avg(lambda j: t[5][j + 1] - t[5][j]):8.0f}")
    print(f"  producer: loads of a tile (0->1)              {avg(lambda j: t[1][j] - t[0][j]):8.0f}")
    print(f"  producer: runs ahead of the MMA lane by       {avg(lambda j: t[4][j] - t[0][j]):8.0f}  (tile start -> its first operands used)")
    print(f"  MMA: wait for a free accumulator (2->3)       {avg(lambda j: t[3][j] - t[2][j]):8.0f}")
    print(f"  MMA: wait for the first operands (3->4)       {avg(lambda j: t[4][j] - t[3][j]):8.0f}")
    print(f"  MMA: k loop: waits + issue (4->5)             {avg(lambda j: t[5][j] - t[4][j]):8.0f}")
    for g in (0, 1):
        e = 8 + 12 * g
        print(f"  epilogue group {g}: wait for accumulator        {avg(lambda j: t[e + 1][j] - t[e][j]):8.0f}")
        print(f"  epilogue group {g}: commit issued -> seen        {avg(lambda j: t[e + 1][j] - t[5][j]):8.0f}")
        prev = e + 1
        for i in range(n_chunks_per_group):
            print(f"     chunk {i}: TMEM load + wait                   {avg(lambda j: t[e + 2 + 2 * i][j] - t[prev][j]):8.0f}")
            print(f"     chunk {i}: math, staging, fence, barrier      {avg(lambda j: t[e + 3 + 2 * i][j] - t[e + 2 + 2 * i][j]):8.0f}")
            prev = e + 3 + 2 * i
        print(f"  epilogue group {g}: last barrier -> released     {avg(lambda j: t[e + 10][j] - t[prev][j]):8.0f}")
        print(f"  epilogue group {g}: whole tile (e+0 -> e+10)     {avg(lambda j: t[e + 10][j] - t[e][j]):8.0f}")


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    rnd = lambda *s: (torch.randn(s, generator=g, device="cuda") * 0.5).half()  # noqa: E731
    m = 64 * 4096
    a320 = rnd(m, 320)
    for name, n, k, kw in (("qkv  N1152 K320", 1152, 320, {}), ("qkv  N1536 K320", 1536, 320, {}),
                           ("proj N320 K320 +bias +res", 320, 320, {"res": True}), ("q    N384 K320", 384, 320, {})):
        a = a320 if k == 320 else rnd(m, k)
        w = rnd(n, k)
        out = torch.empty((m, n), device="cuda", dtype=torch.half)
        bias = torch.randn(n, device="cuda")
        res = rnd(m, n) if kw.get("res") else None
        bn = ops.pick_block_n(n, False, m)
        trace(f"linear {name} bn{bn}", lambda: ops.linear(a, w, out, bias=bias, residual=res), (bn // 32 + 1) // 2)
    w = rnd(2560, 320)
    from b200sd.weights import pack_geglu
    wp, bp = pack_geglu(w.float(), torch.randn(2560), 256)
    out = torch.empty((m, 1280), device="cuda", dtype=torch.half)
    wp, bp = wp.half().cuda(), bp.cuda()
    trace("geglu N2560 K320 bn256", lambda: ops.linear(a320, wp, out, bias=bp, flags=ops.EPI_GEGLU), 2)
    a1280 = rnd(m, 1280)
    w = rnd(320, 1280)
    out = torch.empty((m, 320), device="cuda", dtype=torch.half)
    res = rnd(m, 320)
    bias = torch.randn(320, device="cuda")
    trace("ff2  N320 K1280 +res bn160", lambda: ops.linear(a1280, w, out, bias=bias, residual=res), 3)
    x = rnd(64, 64, 64, 320)
    w = rnd(320, 9 * 320)
    out = torch.empty((64 * 4096, 320), device="cuda", dtype=torch.half)
    trace("conv3x3 320->320 @64x64 bn160", lambda: ops.conv2d(x, w, out, ksize=3, bias=bias), 3)


if __name__ == "__main__":
    main()
