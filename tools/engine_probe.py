"""GPU probe: end-to-end timing of the SD1.5 engine + per-op-category time breakdown of one UNet evaluation.
Writes gpurun_out/engine_probe.json.  Usage: python tools/engine_probe.py [--b 4] [--steps 20]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-webui-distributed_b200"))
import torch  # noqa: E402

from b200sd import config as C, engine as E, ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, nargs="+", default=[4])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--size", default="sd15")
    ap.add_argument("--hw", type=int, default=64)
    args = ap.parse_args()
    cfgs = (C.SD15_UNET, C.SD15_VAE, C.SD15_CLIP) if args.size == "sd15" else (C.TINY_UNET, C.TINY_VAE, C.TINY_CLIP)
    t0 = time.time()
    sd = synth.make_state_dict(*cfgs, seed=0)
    t1 = time.time()
    eng = E.SDEngine(sd, *cfgs, device="cuda:0", use_graphs=True)
    torch.cuda.synchronize()
    t2 = time.time()
    out = {"gen_weights_s": t1 - t0, "pack_weights_s": t2 - t1, "runs": []}
    vocab = cfgs[2].vocab
    for b in args.b:
        g = torch.Generator().manual_seed(1234)
        tok = torch.randint(0, vocab - 3, (b, 77), generator=g)
        neg = torch.full((b, 77), vocab - 1)
        hw = args.hw
        # warm-up (builds plans + graphs)
        eng.txt2img(tok, neg, 1000, args.steps, 7.0, hw * 8, hw * 8)
        torch.cuda.synchronize()
        rec = {"b": b, "hw": hw, "steps": args.steps}
        for rep in range(2):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            torch.cuda.synchronize()
            w0 = time.time()
            ev[0].record()
            cond = eng.encode_prompts(tok)
            unc = eng.encode_prompts(neg)
            nz = E.per_image_noise(1000, b, (4, hw, hw))
            ev[1].record()
            lat = eng.sample(cond, unc, nz[0], args.steps, 7.0, "DDIM")
            ev[2].record()
            u8 = eng.decode(lat, hw, hw)
            ev[3].record()
            torch.cuda.synchronize()
            wall = time.time() - w0
            rec[f"rep{rep}"] = {"wall_s": wall, "prep_ms": ev[0].elapsed_time(ev[1]), "sample_ms": ev[1].elapsed_time(ev[2]),
                                "decode_ms": ev[2].elapsed_time(ev[3]), "img_per_s": b / wall,
                                "ms_per_unet_eval": ev[1].elapsed_time(ev[2]) / max(1, eng.last_unet_evals)}
        # per-op breakdown of one eager UNet evaluation
        plan = eng.plan(b, hw, hw)
        cats = {}
        evs = []
        for fn, a, k in plan.unet.ops:
            name = getattr(fn, "__name__", "op")
            if name == "<lambda>":
                name = "groupnorm"
            if name in ("linear", "conv2d"):
                shp = a[2].shape[-1] if name == "linear" else a[1].shape[0]
                name = f"{name}"
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*a, **k)
            e1.record()
            evs.append((name, e0, e1, a))
        torch.cuda.synchronize()
        detail = {}
        for name, e0, e1, a in evs:
            ms = e0.elapsed_time(e1)
            cats[name] = cats.get(name, 0.0) + ms
            if name in ("linear", "conv2d", "attention"):
                if name == "attention":
                    key = f"attention Sq{a[0].shape[1]} Skv{a[1].shape[1]} d{a[5]}"
                elif name == "conv2d":
                    key = f"conv2d {tuple(a[0].shape)}->{a[1].shape[0]}"
                else:
                    key = f"linear M{a[0].numel() // a[0].shape[-1]} N{a[1].shape[0]} K{a[1].shape[1]}"
                d = detail.setdefault(key, [0, 0.0])
                d[0] += 1
                d[1] += ms
        rec["unet_eager_ms_by_op"] = cats
        rec["unet_eager_total_ms"] = sum(cats.values())
        rec["unet_n_ops"] = len(plan.unet.ops)
        rec["top_shapes"] = sorted(([k, v[0], round(v[1], 3)] for k, v in detail.items()), key=lambda r: -r[2])[:40]
        rec["pool_bytes_unet"] = plan.unet.pool.bytes
        rec["mem_allocated_gb"] = torch.cuda.memory_allocated() / 2**30
        out["runs"].append(rec)
        print(json.dumps(rec)[:3000])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "engine_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
