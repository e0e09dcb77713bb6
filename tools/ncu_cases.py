"""Representative hot-path launches for Nsight Compute (run under `ncu --profile-from-start off ...`).

    python tools/ncu_cases.py unet            # one eager SD1.5 UNet evaluation at UNet batch 16 (launch list)
    python tools/ncu_cases.py conv|geglu|proj|qkv|attn|gn|ln   # a single op at its UNet-batch-16 shape (--set full capture)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-webui-distributed_b200"))
import torch  # noqa: E402

from b200sd import ops  # noqa: E402


def rnd(*shape, scale=1.0):
    return (torch.randn(shape, device="cuda") * scale).half()


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "unet"
    nb = int(os.environ.get("NCU_NB", "16"))  # UNet batch (2 x images); bench.py's per-GPU batch 32 is NCU_NB=64
    if case == "unet":
        from b200sd import config as C, engine as E, synth
        cfgs = (C.SD15_UNET, C.SD15_VAE, C.SD15_CLIP)
        eng = E.SDEngine(synth.make_state_dict(*cfgs, seed=0), *cfgs, device="cuda:0", use_graphs=False)
        plan = eng.plan(nb // 2, 64, 64)
        plan.unet.set_context(rnd(nb, 77, 768))
        plan.table[:1].copy_(eng.temb.table(torch.tensor([651.0])))
        ops.pack_unet_input(torch.randn((nb // 2, 4096, 4), device="cuda"), plan.unet.xin, 1.0)
        fn = lambda: plan.step_ddim(7.0)  # noqa: E731
    elif case == "vae":
        from b200sd import config as C, engine as E, synth
        cfgs = (C.SD15_UNET, C.SD15_VAE, C.SD15_CLIP)
        eng = E.SDEngine(synth.make_state_dict(*cfgs, seed=0), *cfgs, device="cuda:0", use_graphs=False)
        lat = torch.randn((nb // 2, 4096, 4), device="cuda") * 0.18215
        fn = lambda: eng.decode(lat, 64, 64)  # noqa: E731
    elif case == "conv":
        x, w, o = rnd(nb, 64, 64, 320), rnd(320, 2880, scale=0.02), torch.empty((nb * 4096, 320), device="cuda", dtype=torch.half)
        b = torch.randn(320, device="cuda")
        fn = lambda: ops.conv2d(x, w, o, ksize=3, bias=b)  # noqa: E731
    elif case == "geglu":
        x, w, o = rnd(nb * 4096, 320), rnd(2560, 320, scale=0.05), torch.empty((nb * 4096, 1280), device="cuda", dtype=torch.half)
        b = torch.randn(2560, device="cuda")
        fn = lambda: ops.linear(x, w, o, bias=b, flags=ops.EPI_GEGLU)  # noqa: E731
    elif case == "proj":
        x, w, o, r = rnd(nb * 4096, 320), rnd(320, 320, scale=0.05), torch.empty((nb * 4096, 320), device="cuda", dtype=torch.half), rnd(nb * 4096, 320)
        b = torch.randn(320, device="cuda")
        fn = lambda: ops.linear(x, w, o, bias=b, residual=r)  # noqa: E731
    elif case == "attn":
        dp = 48  # head pitch round16(40 + 1), as the UNet's fused q|k|v buffer has it since round 2
        qkv = torch.zeros((nb, 4096, 3, 8, dp), device="cuda", dtype=torch.half)
        qkv[..., :40] = rnd(nb, 4096, 3, 8, 40)
        qkv[:, :, 2, :, 40] = 1.0  # ones column of V, as the UNet's qkv bias produces it
        flat = qkv.reshape(nb, 4096, 3 * 8 * dp)
        q, k, v = flat[..., :8 * dp], flat[..., 8 * dp:16 * dp], flat[..., 16 * dp:]
        o = torch.empty((nb, 4096, 320), device="cuda", dtype=torch.half)
        fn = lambda: ops.attention(q, k, v, o, 8, 40, dp, 40 ** -0.5, v_ones_col=True)  # noqa: E731
    elif case == "qkv":
        x, w, o = rnd(nb * 4096, 320), rnd(1152, 320, scale=0.05), torch.empty((nb * 4096, 1152), device="cuda", dtype=torch.half)
        b = torch.zeros(1152, device="cuda")
        fn = lambda: ops.linear(x, w, o, bias=b)  # noqa: E731
    elif case == "gn":
        x, o = rnd(nb, 4096, 320), torch.empty((nb, 4096, 320), device="cuda", dtype=torch.half)
        st = torch.zeros((ops.groupnorm_stats_floats(nb, 4096, 320, 32),), device="cuda")
        g, b = torch.ones(320, device="cuda"), torch.zeros(320, device="cuda")
        fn = lambda: ops.groupnorm(x, o, st, g, b, 32, 1e-5, True)  # noqa: E731
    elif case == "ln":
        x, o = rnd(nb * 4096, 320), torch.empty((nb * 4096, 320), device="cuda", dtype=torch.half)
        g, b = torch.ones(320, device="cuda"), torch.zeros(320, device="cuda")
        fn = lambda: ops.layernorm(x, o, g, b)  # noqa: E731
    else:
        raise SystemExit(f"unknown case {case}")
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
