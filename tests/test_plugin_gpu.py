"""The drop-in boundary on a GPU: DistributedScript hooks -> LocalGPUWorker.request() -> SDEngine, thin-client world.
Uses the reduced-width model (same topology) so the whole file runs in seconds."""
import logging

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import modules.processing as processing
    import modules.scripts as mscripts
    from b200sd import config as C, engine as E, synth
    from scripts.distributed import DistributedScript
    from scripts.spartan import pmodels, shared as sh
    from scripts.spartan.worker import State
    logging.getLogger("distributed").setLevel(logging.ERROR)
    cfgs = (C.TINY_UNET, C.TINY_VAE, C.TINY_CLIP)
    sd = synth.make_state_dict(*cfgs, seed=0)
    eng = E.SDEngine(sd, *cfgs, device="cuda:0")
    sh.benchmark_payload = pmodels.Benchmark_Payload()
    return processing, mscripts, DistributedScript, eng, State, sh


def _fresh_world(DistributedScript, factory, ipm=600.0):
    from scripts.spartan.world import World
    w = World(verify_remotes=False)
    DistributedScript.world = w
    wk = w.add_local_gpus(factory, devices=[0], avg_ipm=ipm)[0]
    wk.benchmarked = True
    w.thin_client_mode = True
    w.benchmark = lambda *a, **k: None
    return w, wk


def _request(processing, mscripts, script, batch, tokens, steps=6, size=128, sampler="DDIM"):
    p = processing.StableDiffusionProcessingTxt2Img(
        prompt="a synthetic prompt", negative_prompt="", seed=1000, subseed=5, subseed_strength=0, batch_size=batch,
        n_iter=1, steps=steps, width=size, height=size, sampler_name=sampler, cfg_scale=7.0,
        scripts=mscripts.ScriptRunner([script]), script_args=[])
    p.prompt_tokens = tokens.tolist()
    return p, processing.process_images(p)


def test_plugin_path_matches_direct_engine_call(env):
    processing, mscripts, DistributedScript, eng, State, sh = env
    w, wk = _fresh_world(DistributedScript, lambda d: eng)
    script = DistributedScript()
    script.args_from = script.args_to = 0
    g = torch.Generator().manual_seed(3)
    tokens = torch.randint(0, 990, (4, 77), generator=g)
    p, out = _request(processing, mscripts, script, 4, tokens)
    assert len(out.images) == 4 and p.seeds == [1000, 1001, 1002, 1003] and p.subseeds == [5, 6, 7, 8]
    assert all("Worker Label: gpu0" in t for t in out.infotexts[:4])
    assert wk.state == State.IDLE and wk.jobs_requested == 1 and wk.response is None  # cleared by postprocess
    from b200sd.factory import synthetic_tokens
    direct = eng.txt2img(tokens, synthetic_tokens([""] * 4, eng.clip_cfg.vocab), 1000, steps=6, cfg_scale=7.0, height=128,
                         width=128, sampler="DDIM").cpu()
    import numpy as np
    got = torch.stack([torch.from_numpy(np.asarray(im)) for im in out.images])
    # same kernels, same seeds through the hook chain: every kernel is run-to-run deterministic (GroupNorm statistics
    # are reduced in a fixed order, nothing uses floating-point atomics) and the float<->uint8 lane is lossless, so the
    # images are bit-identical.  A mismatch here means a race.
    assert torch.equal(got, direct)


def test_worker_reply_schema_and_png_lane(env):
    processing, mscripts, DistributedScript, eng, State, sh = env
    w, wk = _fresh_world(DistributedScript, lambda d: eng)
    wk.png_images = True
    payload = {"prompt": "x y", "negative_prompt": "", "seed": 7, "subseed": 9, "subseed_strength": 0, "batch_size": 2,
               "n_iter": 1, "steps": 4, "width": 64, "height": 64, "sampler_name": "Euler a", "cfg_scale": 5.0}
    wk.request(dict(payload), {"sd_model_checkpoint": "m", "sd_vae": None}, True)
    r = wk.response
    import base64, io, json
    from PIL import Image
    import numpy as np
    assert set(r) >= {"images", "parameters", "info", "tensors"} and len(r["images"]) == 2
    info = json.loads(r["info"])
    assert info["all_seeds"] == [7, 8] and info["all_subseeds"] == [9, 10] and len(info["infotexts"]) == 2
    png = np.asarray(Image.open(io.BytesIO(base64.b64decode(r["images"][1]))))
    assert np.array_equal(png, r["tensors"][1].numpy())
    assert wk.loaded_model == "m" and wk.response_time is not None and len(wk.eta_percent_error) <= 1


@pytest.mark.parametrize("name,sched,direct_sched", [("DPM++ 2M Karras", None, None), ("DPM++ 2M", "Uniform", "Uniform"),
                                                     ("DPM++ 2M", "Align Your Steps", None), ("Euler", None, None),
                                                     ("Euler a", "Exponential", "Exponential")])
def test_worker_sampler_names(env, name, sched, direct_sched):
    """the API's sampler / scheduler labels reach the executor (an unknown scheduler falls back to the sampler's default)"""
    processing, mscripts, DistributedScript, eng, State, sh = env
    w, wk = _fresh_world(DistributedScript, lambda d: eng)
    payload = {"prompt": "p q", "negative_prompt": "", "seed": 21, "subseed": 1, "subseed_strength": 0, "batch_size": 2,
               "n_iter": 1, "steps": 5, "width": 64, "height": 64, "sampler_name": name, "cfg_scale": 6.0}
    if sched is not None:
        payload["scheduler"] = sched
    wk.request(dict(payload), None, False)
    r = wk.response
    assert r is not None and r["parameters"]["sampler_name"] == name and wk.state == State.IDLE
    from b200sd.factory import synthetic_tokens
    vocab = eng.clip_cfg.vocab
    direct = eng.txt2img(synthetic_tokens(["p q"] * 2, vocab), synthetic_tokens([""] * 2, vocab), 21, steps=5, cfg_scale=6.0,
                         height=64, width=64, sampler=name, scheduler=direct_sched).cpu()
    assert torch.equal(r["tensors"], direct)


def test_worker_img2img_request(env):
    """img2img payload as the reference sends it: init_images are PIL images in p.__dict__ (worker.py:365-373)."""
    processing, mscripts, DistributedScript, eng, State, sh = env
    w, wk = _fresh_world(DistributedScript, lambda d: eng)
    import numpy as np
    from PIL import Image
    g = torch.Generator().manual_seed(5)
    arr = torch.randint(0, 256, (64, 64, 3), generator=g, dtype=torch.uint8)
    payload = {"prompt": "a b", "negative_prompt": "", "seed": 11, "subseed": 1, "subseed_strength": 0, "batch_size": 2,
               "n_iter": 1, "steps": 8, "width": 64, "height": 64, "sampler_name": "DDIM", "cfg_scale": 7.0,
               "denoising_strength": 0.75, "init_images": [Image.fromarray(arr.numpy())]}
    wk.request(dict(payload), None, False)
    r = wk.response
    assert r is not None and tuple(r["tensors"].shape) == (2, 64, 64, 3) and wk.state == State.IDLE
    from b200sd.factory import synthetic_tokens
    vocab = eng.clip_cfg.vocab
    direct = eng.img2img(synthetic_tokens(["a b"] * 2, vocab), synthetic_tokens([""] * 2, vocab), 11,
                         arr[None].expand(2, -1, -1, -1).contiguous(), 0.75, steps=8, cfg_scale=7.0).cpu()
    assert torch.equal(r["tensors"], direct)  # two runs of the same request are bit-identical (deterministic kernels)


def test_worker_inpainting_request(env):
    """img2img with a mask as the reference forwards it (`image_mask` in p.__dict__ -> API field `mask`,
    worker.py:365-373, 406-410): mask pipeline + per-step latent blend + overlay; equals the direct engine call"""
    processing, mscripts, DistributedScript, eng, State, sh = env
    w, wk = _fresh_world(DistributedScript, lambda d: eng)
    from PIL import Image, ImageDraw
    from b200sd import inpaint as inp
    from b200sd.factory import synthetic_tokens
    g = torch.Generator().manual_seed(6)
    arr = torch.randint(0, 256, (64, 64, 3), generator=g, dtype=torch.uint8)
    mask = Image.new("L", (64, 64), 0)
    ImageDraw.Draw(mask).rectangle((16, 20, 47, 50), fill=255)
    payload = {"prompt": "a b", "negative_prompt": "", "seed": 12, "subseed": 1, "subseed_strength": 0, "batch_size": 2,
               "n_iter": 1, "steps": 8, "width": 64, "height": 64, "sampler_name": "DDIM", "cfg_scale": 7.0,
               "denoising_strength": 0.75, "init_images": [Image.fromarray(arr.numpy())], "image_mask": mask, "mask_blur": 2,
               "inpainting_fill": 1, "inpaint_full_res": False, "inpainting_mask_invert": 0}
    wk.request(dict(payload), None, False)
    r = wk.response
    assert r is not None and tuple(r["tensors"].shape) == (2, 64, 64, 3) and wk.state == State.IDLE
    vocab = eng.clip_cfg.vocab
    init = arr[None].expand(2, -1, -1, -1).contiguous()
    down = 2 ** (len(eng.vae_cfg.ch_mult) - 1)
    m = inp.prepare_mask(mask, 64, 64, 64 // down, 64 // down, mask_blur=2)
    direct = eng.img2img(synthetic_tokens(["a b"] * 2, vocab), synthetic_tokens([""] * 2, vocab), 12, init, 0.75, steps=8,
                         cfg_scale=7.0, latmask=m.latmask).cpu()
    direct = inp.apply_overlays(direct, inp.overlays_for(init, m))
    assert torch.equal(r["tensors"], direct)
    # outside the (blurred) mask the request returns the init image
    assert torch.equal(r["tensors"][:, :4], init[:, :4])
    plain = eng.img2img(synthetic_tokens(["a b"] * 2, vocab), synthetic_tokens([""] * 2, vocab), 12, init, 0.75, steps=8,
                        cfg_scale=7.0).cpu()
    assert not torch.equal(plain[:, :4], init[:, :4])


def test_rest_worker_serves_the_executor(env):
    """SURVEY §8 f1 on a GPU: POST /sdapi/v1/txt2img -> LocalGPUWorker -> SDEngine; the PNGs decode to exactly the images
    a direct engine call produces (PNG is lossless, the kernels are deterministic)."""
    processing, mscripts, DistributedScript, eng, State, sh = env
    import base64, io
    import numpy as np
    from fastapi.testclient import TestClient
    from PIL import Image
    from b200sd.factory import synthetic_tokens
    from server.sdapi import create_app
    c = TestClient(create_app(lambda device: eng, [0]))
    mem = c.get("/sdapi/v1/memory").json()
    assert mem["cuda"]["system"]["total"] > mem["cuda"]["system"]["free"] > 0
    payload = {"prompt": "rest probe", "negative_prompt": "", "seed": 5, "subseed": 1, "batch_size": 2, "n_iter": 1, "steps": 5,
               "width": 64, "height": 64, "sampler_name": "DDIM", "cfg_scale": 6.0}
    r = c.post("/sdapi/v1/txt2img", json=payload)
    assert r.status_code == 200
    got = torch.stack([torch.from_numpy(np.array(Image.open(io.BytesIO(base64.b64decode(s))))) for s in r.json()["images"]])
    vocab = eng.clip_cfg.vocab
    direct = eng.txt2img(synthetic_tokens(["rest probe"] * 2, vocab), synthetic_tokens([""] * 2, vocab), 5, steps=5,
                         cfg_scale=6.0, height=64, width=64, sampler="DDIM").cpu()
    assert torch.equal(got, direct)


def test_device_failure_marks_worker_unavailable(env):
    processing, mscripts, DistributedScript, eng, State, sh = env

    def broken(dev):
        raise RuntimeError("CUDA error: an illegal memory access was encountered (injected)")

    w, wk = _fresh_world(DistributedScript, broken)
    script = DistributedScript()
    script.args_from = script.args_to = 0
    tokens = torch.zeros((2, 77), dtype=torch.long)
    p, out = _request(processing, mscripts, script, 2, tokens)
    assert wk.state == State.UNAVAILABLE and len(out.images) == 0


def test_local_benchmark_sets_ipm(env):
    processing, mscripts, DistributedScript, eng, State, sh = env
    w, wk = _fresh_world(DistributedScript, lambda d: eng, ipm=0.0)
    wk.benchmarked = False
    ipm = wk.benchmark()
    assert ipm > 0 and wk.avg_ipm == ipm and wk.benchmarked and wk.state == State.IDLE
