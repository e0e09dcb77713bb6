"""Per-tile timeline of one attention CTA (clock64 stamps written by the kernel's debug hook): where does a kv tile's
time go between the softmax warps and the MMA thread?  Needs a trace-enabled library build:
    B200SD_NVCC_EXTRA="-DB200SD_ATTN_TRACE_ENABLE=1" python stable-diffusion-webui-distributed_b200/b200sd/build.py --force

events (per kv tile j):
  softmax warp 2, lane 0:  0 loop top   1 S ready (s_full passed)   2 S in registers   3 exp/pack issued
                           5 P stored   6 proxy fence + syncwarp done (arrive next)
  MMA thread (ring mode):  7 waiting for p_full of tile j (P_j, K_(j+2), V_j)   8 complete   13 Q.K^T of tile j issued
                           (stamped at index j)   10 P.V of tile j issued + committed
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-webui-distributed_b200"))
import torch  # noqa: E402

from b200sd import _lib, ops  # noqa: E402


def main():
    nb = 16
    qkv = torch.zeros((nb, 4096, 3, 8, 64), device="cuda", dtype=torch.half)
    qkv[..., :40] = torch.randn((nb, 4096, 3, 8, 40), device="cuda").half()
    qkv[:, :, 2, :, 40] = 1.0
    flat = qkv.reshape(nb, 4096, 1536)
    q, k, v = flat[..., :512], flat[..., 512:1024], flat[..., 1024:]
    o = torch.empty((nb, 4096, 320), device="cuda", dtype=torch.half)
    for _ in range(3):
        ops.attention(q, k, v, o, 8, 40, 64, 40 ** -0.5, v_ones_col=True)
    buf = torch.zeros((16, 64), device="cuda", dtype=torch.int64)
    _lib.lib().b200sd_debug_attention_trace(ctypes.c_void_p(buf.data_ptr()))
    ops.attention(q, k, v, o, 8, 40, 64, 40 ** -0.5, v_ones_col=True)
    torch.cuda.synchronize()
    _lib.lib().b200sd_debug_attention_trace(ctypes.c_void_p(0))
    t = buf.cpu()
    t0 = int(t[0, 0])
    rel = (t - t0).tolist()
    print("tile |  softmax: top  Srdy  Sreg  exp  Pstor fence | MMA thread: waitP  Pfull  QK(j+2) issued  PV(j) issued")
    for j in list(range(0, 6)) + list(range(30, 36)) + list(range(58, 62)):
        r = [rel[e][j] for e in range(14)]
        print(f"{j:4d} | " + " ".join(f"{r[e]:7d}" for e in (0, 1, 2, 3, 5, 6)) + " | " +
              f"{r[7]:7d} {r[8]:7d} {rel[13][(j + 2) & 63]:7d} {r[10]:7d}")
    js = range(16, 56)
    def avg(f):
        vals = [f(j) for j in js]
        return sum(vals) / len(vals)
    print("\nsteady state (tiles 16..55), cycles:")
    print(f"  tile period (softmax loop top to top)        {avg(lambda j: rel[0][j + 1] - rel[0][j]):8.1f}")
    print(f"  softmax: wait for S (0->1)                   {avg(lambda j: rel[1][j] - rel[0][j]):8.1f}")
    print(f"  softmax: TMEM load of 32 columns (1->2)      {avg(lambda j: rel[2][j] - rel[1][j]):8.1f}")
    print(f"  softmax: FFMA2/ex2/pack/max (2->3)           {avg(lambda j: rel[3][j] - rel[2][j]):8.1f}")
    print(f"  softmax: any_sync + P stores (3->5)          {avg(lambda j: rel[5][j] - rel[3][j]):8.1f}")
    print(f"  softmax: proxy fence + syncwarp (5->6)       {avg(lambda j: rel[6][j] - rel[5][j]):8.1f}")
    print(f"  softmax: arrive + loop (6->next top)         {avg(lambda j: rel[0][j + 1] - rel[6][j]):8.1f}")
    print(f"  MMA: loop period (7->7)                      {avg(lambda j: rel[7][j + 1] - rel[7][j]):8.1f}")
    print(f"  MMA: wait for P_j + K_(j+2) + V_j (7->8)     {avg(lambda j: rel[8][j] - rel[7][j]):8.1f}")
    print(f"  MMA: traced warp's arrive -> P_j complete    {avg(lambda j: rel[8][j] - rel[6][j]):8.1f}")
    print(f"  MMA: issue Q.K(j+2) + commit (8->13)         {avg(lambda j: rel[13][j + 2] - rel[8][j]):8.1f}")
    print(f"  MMA: issue P.V(j) + commit (13->10)          {avg(lambda j: rel[10][j] - rel[13][j + 2]):8.1f}")
    print(f"  MMA: P.V issued -> next loop top (10->7)     {avg(lambda j: rel[7][j + 1] - rel[10][j]):8.1f}")
    print(f"  Q.K(j+2) issued -> softmax sees S(j+2) ready {avg(lambda j: rel[1][j + 2] - rel[13][j + 2]):8.1f}")
    print(f"  Q.K(j+2) issued -> softmax WANTS S(j+2)      {avg(lambda j: rel[0][j + 2] - rel[13][j + 2]):8.1f}  (< 0: the softmax warp waited for the issuer)")


if __name__ == "__main__":
    main()
