"""SURVEY §8(e) on the plugin path: one request's batch sharded over the GPUs of the box by World.optimize_jobs, one
LocalGPUWorker thread per device, thin-client master.  Needs >= 2 GPUs (skipped on the 1-GPU boxes; run with
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`)."""
import logging

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_batch_sharded_over_two_gpus_equals_one_gpu():
    import modules.processing as processing
    import modules.scripts as mscripts
    from b200sd import config as C, engine as E, synth
    from b200sd.factory import synthetic_tokens
    from scripts.distributed import DistributedScript
    from scripts.spartan import pmodels, shared as sh
    from scripts.spartan.worker import State
    from scripts.spartan.world import World
    logging.getLogger("distributed").setLevel(logging.ERROR)
    cfgs = (C.TINY_UNET, C.TINY_VAE, C.TINY_CLIP)
    sd = synth.make_state_dict(*cfgs, seed=0)
    engines = {}

    def factory(device):
        if device not in engines:
            engines[device] = E.SDEngine(sd, *cfgs, device=device)
        return engines[device]

    sh.benchmark_payload = pmodels.Benchmark_Payload()
    w = World(verify_remotes=False)
    DistributedScript.world = w
    workers = w.add_local_gpus(factory, devices=[0, 1], avg_ipm=600.0)
    for wk in workers:
        wk.benchmarked = True
    w.thin_client_mode = True
    w.benchmark = lambda *a, **k: None
    script = DistributedScript()
    script.args_from = script.args_to = 0
    batch = 7  # uneven shards: 4 + 3
    g = torch.Generator().manual_seed(9)
    tokens = torch.randint(0, 990, (1, 77), generator=g).expand(batch, -1).contiguous()  # one prompt per request
    p = processing.StableDiffusionProcessingTxt2Img(
        prompt="a synthetic prompt", negative_prompt="", seed=4000, subseed=50, subseed_strength=0, batch_size=batch, n_iter=1,
        steps=6, width=128, height=128, sampler_name="DDIM", cfg_scale=7.0, scripts=mscripts.ScriptRunner([script]),
        script_args=[])
    p.prompt_tokens = tokens.tolist()
    out = processing.process_images(p)
    assert len(out.images) == batch and p.seeds == list(range(4000, 4000 + batch))
    assert sorted(wk.jobs_requested for wk in workers) == [1, 1] and all(wk.state == State.IDLE for wk in workers)
    assert set(engines) == {"cuda:0", "cuda:1"}
    got = torch.stack([torch.from_numpy(np.array(im)) for im in out.images])
    # Every kernel is batch-invariant (an image's reduction orders do not depend on which other images share its
    # launch), so the sharded result is bit-identical to generating the whole batch on one GPU.
    direct = engines["cuda:0"].txt2img(tokens, synthetic_tokens([""] * batch, cfgs[2].vocab), 4000, steps=6, cfg_scale=7.0,
                                       height=128, width=128, sampler="DDIM").cpu()
    assert torch.equal(got, direct)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sd15_batch_sharded_over_all_gpus_equals_one_gpu():
    """The same property at the BASELINE model size (SD1.5, 512x512, 20 DDIM timesteps): ONE DistributedScript request
    sharded by World.optimize_jobs over every GPU of the box (uneven shards), bit-identical to one GPU making the whole
    batch, seeds offset per job as reference scripts/distributed.py:297-305."""
    import modules.processing as processing
    import modules.scripts as mscripts
    from b200sd import factory
    from b200sd.factory import synthetic_tokens
    from scripts.distributed import DistributedScript
    from scripts.spartan import pmodels, shared as sh
    from scripts.spartan.worker import State
    from scripts.spartan.world import World
    logging.getLogger("distributed").setLevel(logging.ERROR)
    ngpu = torch.cuda.device_count()
    sh.benchmark_payload = pmodels.Benchmark_Payload()
    w = World(verify_remotes=False)
    DistributedScript.world = w
    workers = w.add_local_gpus(lambda d: factory.default_engine_factory(d, "sd15"), devices=list(range(ngpu)), avg_ipm=600.0)
    for wk in workers:
        wk.benchmarked = True
    w.thin_client_mode = True
    w.benchmark = lambda *a, **k: None
    script = DistributedScript()
    script.args_from = script.args_to = 0
    batch = 2 * ngpu + 1  # uneven shards
    g = torch.Generator().manual_seed(11)
    tokens = torch.randint(0, 49000, (1, 77), generator=g).expand(batch, -1).contiguous()
    p = processing.StableDiffusionProcessingTxt2Img(
        prompt="a synthetic prompt", negative_prompt="", seed=7000, subseed=3, subseed_strength=0, batch_size=batch, n_iter=1,
        steps=20, width=512, height=512, sampler_name="DDIM", cfg_scale=7.0, scripts=mscripts.ScriptRunner([script]),
        script_args=[])
    p.prompt_tokens = tokens.tolist()
    out = processing.process_images(p)
    assert len(out.images) == batch and p.seeds == list(range(7000, 7000 + batch))
    assert [wk.jobs_requested for wk in workers] == [1] * ngpu and all(wk.state == State.IDLE for wk in workers)
    got = torch.stack([torch.from_numpy(np.array(im)) for im in out.images])
    eng0 = factory.default_engine_factory("cuda:0", "sd15")
    direct = eng0.txt2img(tokens, synthetic_tokens([""] * batch, eng0.clip_cfg.vocab), 7000, steps=20, cfg_scale=7.0,
                          height=512, width=512, sampler="DDIM").cpu()
    assert torch.equal(got, direct)
    print(f"sd15 {batch} images over {ngpu} GPUs == 1 GPU, bit-identical")
