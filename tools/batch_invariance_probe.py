"""Which images of a batch differ when the batch is generated whole vs in shards on ONE GPU?  (debug probe for the
multi-GPU bit-identity test)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stable-diffusion-webui-distributed_b200")]
import torch  # noqa: E402

from b200sd import factory  # noqa: E402
from b200sd.factory import synthetic_tokens  # noqa: E402


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 17
    shards = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [3] + [2] * 7
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    eng = factory.default_engine_factory("cuda:0", "sd15")
    g = torch.Generator().manual_seed(11)
    tok1 = torch.randint(0, 49000, (1, 77), generator=g)
    neg1 = synthetic_tokens([""], eng.clip_cfg.vocab)
    parts, off = [], 0
    for b in shards:
        parts.append(eng.txt2img(tok1.expand(b, -1).contiguous(), neg1.expand(b, -1).contiguous(), 7000 + off, steps=steps,
                                 cfg_scale=7.0, height=512, width=512, sampler="DDIM").cpu())
        off += b
    parts = torch.cat(parts)
    for rep in range(2):
        full = eng.txt2img(tok1.expand(total, -1).contiguous(), neg1.expand(total, -1).contiguous(), 7000, steps=steps, cfg_scale=7.0,
                           height=512, width=512, sampler="DDIM").cpu()
        lat = eng.plan(total, 64, 64).x.clone()
        diff = [(k, int((full[k].int() - parts[k].int()).abs().max()), int((full[k] != parts[k]).sum())) for k in range(total)
                if not torch.equal(full[k], parts[k])]
        print(f"PDL={os.environ.get('B200SD_PDL', 'default')} total {total} shards {shards} rep {rep}: {len(diff)} images differ", diff[:20])


if __name__ == "__main__":
    main()
