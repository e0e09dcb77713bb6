"""LocalGPUWorker.request() — the drop-in boundary (reference scripts/spartan/worker.py:288-504) — on the CPU, with
b200sd.ops emulated (tests/ops_emulator.py): payload fields the reference forwards reach the executor with sdwui's
semantics.  The GPU twin of this file is tests/test_plugin_gpu.py."""
import json
import logging

import pytest
import torch

import ops_emulator


@pytest.fixture()
def env(monkeypatch):
    from b200sd import config as C, engine as E, ops, synth
    from scripts.spartan import pmodels, shared as sh
    from scripts.spartan.local_worker import LocalGPUWorker
    logging.getLogger("distributed").setLevel(logging.ERROR)
    ops_emulator.install(monkeypatch, ops)
    monkeypatch.setattr(E.SDEngine, "_require_cuda", False)
    cfgs = (C.TINY_UNET, C.TINY_VAE, C.TINY_CLIP)
    eng = E.SDEngine(synth.make_state_dict(*cfgs, seed=0), *cfgs, device="cpu", dtype=torch.float32, use_graphs=False, vae_chunk=2)
    sh.benchmark_payload = pmodels.Benchmark_Payload()
    wk = LocalGPUWorker(0, lambda d: eng, avg_ipm=600.0)
    return wk, eng, E


def _payload(**kw):
    p = {"prompt": "a b", "negative_prompt": "", "seed": 30, "subseed": 4, "subseed_strength": 0, "batch_size": 2, "n_iter": 1,
         "steps": 4, "width": 64, "height": 64, "sampler_name": "DDIM", "cfg_scale": 7.0}
    p.update(kw)
    return p


@pytest.mark.parametrize("name", ["Heun", "DPM++ SDE Karras", "PLMS", "LMS", "DPM fast"])
def test_reference_sampler_names_run_without_fallback(env, name):
    wk, eng, E = env
    from b200sd.factory import synthetic_tokens
    wk.request(_payload(sampler_name=name), None, False)
    r = wk.response
    assert r is not None and r["parameters"]["sampler_name"] == name
    v = eng.clip_cfg.vocab
    direct = eng.txt2img(synthetic_tokens(["a b"] * 2, v), synthetic_tokens([""] * 2, v), 30, steps=4, cfg_scale=7.0, height=64,
                         width=64, sampler=name)
    assert torch.equal(r["tensors"], direct.to(torch.uint8))


def test_variation_seeds_follow_sdwui(env):
    """subseed_strength != 0: every image of the job keeps the base seed, subseeds advance (processing.py all_seeds /
    all_subseeds; the reference does not offset `seed` per job in that case, scripts/distributed.py:297-305)"""
    wk, eng, E = env
    wk.request(_payload(batch_size=3, subseed_strength=0.4, n_iter=2), None, False)
    info = json.loads(wk.response["info"])
    assert info["all_seeds"] == [30] * 6 and info["all_subseeds"] == [4, 5, 6, 7, 8, 9]
    t = wk.response["tensors"]
    assert t.shape[0] == 6 and not torch.equal(t[0], t[1]) and not torch.equal(t[0], t[3])
    # image 4 (iteration 1, k = 1) = base noise(seed) slerped with noise(subseed + 3 + 1)
    from b200sd.factory import synthetic_tokens
    v = eng.clip_cfg.vocab
    eng.variation = (4 + 3, 0.4)
    direct = eng.txt2img(synthetic_tokens(["a b"] * 3, v), synthetic_tokens([""] * 3, v), 30, steps=4, cfg_scale=7.0, height=64,
                         width=64, sampler="DDIM")
    eng.variation = (None, 0.0)
    assert torch.equal(t[3:], direct.to(torch.uint8))


def test_only_masked_inpainting_request(env):
    """inpaint_full_res: a 96x80 init picture, a 64x64 processing size; the reply has the init picture's size and is
    untouched away from the mask; a k-diffusion sampler carries the mask"""
    from PIL import Image, ImageDraw
    wk, eng, E = env
    g = torch.Generator().manual_seed(8)
    arr = torch.randint(0, 256, (80, 96, 3), generator=g, dtype=torch.uint8)
    mask = Image.new("L", (96, 80), 0)
    ImageDraw.Draw(mask).rectangle((40, 30, 60, 50), fill=255)
    wk.request(_payload(sampler_name="Euler a", init_images=[Image.fromarray(arr.numpy())], image_mask=mask, mask_blur=2,
                        inpaint_full_res=True, inpaint_full_res_padding=8, denoising_strength=0.8, steps=6), None, False)
    r = wk.response
    assert r is not None and tuple(r["tensors"].shape) == (2, 80, 96, 3)
    assert torch.equal(r["tensors"][:, :20], arr[None, :20].expand(2, -1, -1, -1))
    assert not torch.equal(r["tensors"][:, 35:45, 45:55], arr[None, 35:45, 45:55].expand(2, -1, -1, -1))


def test_explicit_zero_denoising_strength_is_kept(env):
    wk, eng, E = env
    from PIL import Image
    g = torch.Generator().manual_seed(9)
    arr = torch.randint(0, 256, (64, 64, 3), generator=g, dtype=torch.uint8)
    wk.request(_payload(init_images=[Image.fromarray(arr.numpy())], denoising_strength=0, steps=8), None, False)
    assert wk.response is not None and eng.last_unet_evals == 0     # DDIM on timesteps[:1]: nothing to evaluate
    wk.request(_payload(init_images=[Image.fromarray(arr.numpy())], denoising_strength=0.05, steps=20), None, False)
    assert wk.response is not None and eng.last_unet_evals == 0     # used to raise IndexError
