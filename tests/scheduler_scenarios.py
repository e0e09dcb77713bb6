"""Scheduler / dispatch scenarios shared by the golden-vector generator (runs the REFERENCE code, here only)
and by tests/test_scheduler_parity.py (runs OUR drop-in).  TEST INFRASTRUCTURE.

Both sides expose the same module surface (`scripts.spartan.world`, `.worker`, `.shared`, `.pmodels`), so the
scenario code is written once against that surface and returns plain JSON-able results.
"""
import copy
import json
import random
import types


def _mk_world(mods, ipms, *, job_timeout=3, complement=True, step_scaling=False, thin=False, pixel_caps=None,
              mpe=None, states=None):
    world_mod, worker_mod, sh, pmodels = mods
    sh.benchmark_payload = pmodels.Benchmark_Payload()
    w = world_mod.World(verify_remotes=False)
    w.job_timeout = job_timeout
    w.complement_production = complement
    w.step_scaling = step_scaling
    w.thin_client_mode = thin
    m = w.master()
    m.avg_ipm = ipms[0]
    m.benchmarked = True
    for i, ipm in enumerate(ipms[1:], start=1):
        nw = w.add_worker(label=f"w{i}", address=f"10.0.0.{i}", port=7860 + i, avg_ipm=ipm, master=False)
        nw.benchmarked = True
    for i, wk in enumerate(w._workers):
        if pixel_caps and pixel_caps[i] is not None:
            wk.pixel_cap = pixel_caps[i]
        if mpe and mpe[i]:
            wk.eta_percent_error = list(mpe[i])
        if states and states[i] is not None:
            wk.state = worker_mod.State(states[i])
    return w


def _jobs_view(w):
    return [[j.worker.label, int(j.batch_size), bool(j.complementary),
             None if j.step_override is None else float(j.step_override)] for j in w.jobs]


def run_optimize(mods, spec):
    """spec: dict(ipms, batch, n_iter, payload, job_timeout, complement, step_scaling, thin, pixel_caps, mpe, states)"""
    import modules.processing as processing
    original_inner = processing.process_images_inner
    w = _mk_world(mods, spec["ipms"], job_timeout=spec.get("job_timeout", 3), complement=spec.get("complement", True),
                  step_scaling=spec.get("step_scaling", False), thin=spec.get("thin", False),
                  pixel_caps=spec.get("pixel_caps"), mpe=spec.get("mpe"), states=spec.get("states"))
    w.p = types.SimpleNamespace(batch_size=spec["batch"], n_iter=spec.get("n_iter", 1),
                                scripts=types.SimpleNamespace(postprocess_batch_list=lambda *a, **k: None,
                                                              postprocess=lambda *a, **k: None))
    out = {"size": w.size(), "default_batch_size": None, "error": None}
    try:
        w.make_jobs()
        out["default_batch_size"] = w.default_batch_size()
        out["jobs_initial"] = _jobs_view(w)
        payload = dict(spec["payload"])
        payload["batch_size"] = w.default_batch_size()
        w.optimize_jobs(payload)
        out["jobs"] = _jobs_view(w)
        out["num_requested"] = w.num_requested()
        out["bypass"] = processing.process_images_inner is not original_inner
        out["distro_summary"] = w.distro_summary()
    except Exception as e:  # the reference raises on some degenerate worlds; parity includes the exception type
        out["error"] = type(e).__name__
    finally:
        processing.process_images_inner = original_inner
    return out


def optimize_specs():
    base_payload = {"steps": 20, "width": 512, "height": 512, "sampler_name": "Euler a"}
    specs = [
        dict(name="equal8x4", ipms=[10, 10, 10, 10], batch=8),
        dict(name="equal32x8", ipms=[60] * 8, batch=32),
        dict(name="remainder", ipms=[10, 10, 10], batch=8),
        dict(name="b2_world3", ipms=[10, 10, 10], batch=2),
        dict(name="slow_worker", ipms=[20, 20, 2], batch=9),
        dict(name="slow_worker_stepscale", ipms=[20, 20, 0.5], batch=3, step_scaling=True),
        dict(name="slow_master", ipms=[1, 30, 30], batch=8),
        dict(name="pixel_cap", ipms=[10, 10], batch=8, pixel_caps=[None, 2 * 512 * 512]),
        dict(name="ddim_hr", ipms=[10, 5], batch=4,
             payload={"steps": 20, "width": 512, "height": 512, "sampler_name": "DDIM", "enable_hr": True,
                      "hr_scale": 2, "hr_second_pass_steps": 0}),
        dict(name="thin", ipms=[10, 10, 10], batch=4, thin=True),
        dict(name="b200_box_thin8", ipms=[1, 2400, 2400, 2400, 2400, 2400, 2400, 2400, 2400], batch=32, thin=True),
        dict(name="b200_box_uneven", ipms=[1, 2400, 2400, 2000, 2400, 1800, 2400, 2400, 2400], batch=30, thin=True),
        dict(name="no_complement", ipms=[10, 10, 10], batch=2, complement=False),
        dict(name="single", ipms=[10], batch=4),
        dict(name="unavailable_worker", ipms=[10, 10, 10], batch=6, states=[None, 4, None]),
        dict(name="disabled_worker", ipms=[10, 10, 10], batch=6, states=[None, None, 5]),
        dict(name="mpe_correction", ipms=[10, 10], batch=4, mpe=[[10, -5, 20], None]),
        dict(name="n_iter2", ipms=[10, 10, 10], batch=4, n_iter=2),
    ]
    rng = random.Random(20260921)
    samplers = ["Euler a", "DDIM", "Heun", "DPM++ 2M Karras", "DPM adaptive", "LMS", "UniPC"]
    for i in range(220):
        n = rng.randint(1, 8)
        fast = rng.choice([5, 10, 30, 120, 2400])
        ipms = [round(fast * rng.choice([1, 1, 1, 0.9, 0.5, 0.2, 0.05, 2.0]), 3) for _ in range(n)]
        payload = {"steps": rng.choice([10, 20, 30, 50]), "width": rng.choice([512, 768, 1024]),
                   "height": rng.choice([512, 768]), "sampler_name": rng.choice(samplers)}
        if rng.random() < 0.15:
            payload.update(enable_hr=True, hr_scale=rng.choice([1.5, 2.0]), hr_second_pass_steps=rng.choice([0, 10]))
        caps = None
        if rng.random() < 0.3:
            caps = [rng.choice([None, None, 512 * 512, 2 * 512 * 512, 4 * 768 * 768]) for _ in range(n)]
        mpe = None
        if rng.random() < 0.2:
            mpe = [[rng.uniform(-30, 30) for _ in range(rng.randint(1, 5))] if rng.random() < 0.5 else None
                   for _ in range(n)]
        specs.append(dict(name=f"rand{i}", ipms=ipms, batch=rng.randint(1, 40), n_iter=rng.choice([1, 1, 2]),
                          payload=payload, job_timeout=rng.choice([1, 3, 3, 10]), complement=rng.random() < 0.8,
                          step_scaling=rng.random() < 0.3, thin=(n > 1 and rng.random() < 0.25), pixel_caps=caps,
                          mpe=mpe))
    for s in specs:
        s.setdefault("payload", dict(base_payload))
    return specs


def run_eta(mods):
    world_mod, worker_mod, sh, pmodels = mods
    sh.benchmark_payload = pmodels.Benchmark_Payload()
    out = []
    cases = [
        dict(ipm=12, payload={"batch_size": 4, "steps": 20, "width": 512, "height": 512}),
        dict(ipm=12, payload={"batch_size": 4, "steps": 40, "width": 768, "height": 768, "sampler_name": "DPM++ 2M Karras"}),
        dict(ipm=12, payload={"batch_size": 1, "steps": 20, "width": 512, "height": 512, "sampler_name": "Heun"}),
        dict(ipm=12, payload={"batch_size": 2, "steps": 20, "width": 512, "height": 512, "enable_hr": True,
                              "hr_scale": 2.0, "hr_second_pass_steps": 10}),
        dict(ipm=12, payload={"batch_size": 4, "steps": 20, "width": 512, "height": 512}, mpe=[10, -5, 20]),
        dict(ipm=2400, payload={"batch_size": 4, "steps": 20, "width": 512, "height": 512, "sampler_name": "DDIM"}),
        dict(ipm=7.5, payload={"batch_size": 3, "steps": 33, "width": 640, "height": 960, "sampler_name": "NoSuchSampler"}),
        dict(ipm=7.5, payload={"batch_size": 3, "steps": 33, "width": 640, "height": 960}, batch_size=5, samples=1),
    ]
    rng = random.Random(7)
    names = list(worker_mod.Worker.other_to_euler_a.keys()) + ["Euler a"]
    for _ in range(60):
        p = {"batch_size": rng.randint(1, 16), "steps": rng.randint(1, 80), "width": rng.choice([256, 512, 768, 1024]),
             "height": rng.choice([256, 512, 768, 1024]), "sampler_name": rng.choice(names)}
        if rng.random() < 0.3:
            p.update(enable_hr=True, hr_scale=rng.choice([1.25, 2.0]), hr_second_pass_steps=rng.choice([0, 7, 20]))
        cases.append(dict(ipm=round(rng.uniform(0.5, 3000), 3), payload=p,
                          mpe=[rng.uniform(-50, 50) for _ in range(rng.randint(0, 5))] or None))
    for c in cases:
        wk = worker_mod.Worker(address="h", port=1, label="x", avg_ipm=c["ipm"])
        if c.get("mpe"):
            wk.eta_percent_error = list(c["mpe"])
        kw = {}
        if "batch_size" in c:
            kw["batch_size"] = c["batch_size"]
        if "samples" in c:
            kw["samples"] = c["samples"]
        out.append({"case": c, "eta": wk.eta(payload=dict(c["payload"]), quiet=True, **kw), "mpe": wk.eta_mpe()})
    return out


def run_fsm(mods):
    world_mod, worker_mod, sh, pmodels = mods
    S = worker_mod.State
    seqs = [
        [S.WORKING, S.INTERRUPTED, S.IDLE, S.WORKING, S.IDLE, S.UNAVAILABLE, S.WORKING, S.IDLE, S.DISABLED],
        [S.UNAVAILABLE, S.UNAVAILABLE, S.IDLE, S.INTERRUPTED, S.WORKING, S.WORKING, S.INTERRUPTED, S.UNAVAILABLE],
        [S.IDLE, S.IDLE, S.WORKING, S.IDLE],
    ]
    rng = random.Random(3)
    for _ in range(20):
        seqs.append([rng.choice(list(S)) for _ in range(12)])
    out = []
    for start in (S.IDLE, S.DISABLED, S.UNAVAILABLE):
        for seq in seqs:
            for cyc in (False, True):
                wk = worker_mod.Worker(address="h", port=1, label="x", avg_ipm=1.0, state=start)
                trace = []
                for s in seq:
                    wk.set_state(s, expect_cycle=cyc)
                    trace.append(wk.state.name)
                out.append({"start": start.name, "seq": [s.name for s in seq], "expect_cycle": cyc, "trace": trace})
    return out


def run_misc(mods):
    world_mod, worker_mod, sh, pmodels = mods
    out = {}
    out["full_url"] = worker_mod.Worker(address="h", port=1, label="x").full_url("txt2img")
    out["full_url_tls"] = worker_mod.Worker(address="h.example", port=8443, label="y", tls=True).full_url("memory")
    wk = worker_mod.Worker(address="h", port=1, label="x", avg_ipm=3.0)
    out["str"] = str(wk)
    out["eq"] = [wk == worker_mod.Worker(address="q", port=2, label="x"), wk == worker_mod.Worker(address="h", port=1, label="z")]
    j = world_mod.Job(worker=wk, batch_size=2)
    wk.pixel_cap = 3 * 512 * 512
    out["add_work"] = [j.add_work({"width": 512, "height": 512}, 1), j.batch_size,
                       j.add_work({"width": 512, "height": 512}, 1), j.batch_size]
    return out


def run_dispatch(mods, script_cls, specs=None):
    """Drive the real plugin hooks (before_process -> host generation -> postprocess_batch_list -> postprocess) with
    Worker.request replaced by a recorder that answers like an sdwui API worker."""
    import base64
    import io

    import modules.processing as processing
    import modules.scripts as mscripts
    import torch
    from PIL import Image

    world_mod, worker_mod, sh, pmodels = mods
    results = []
    specs = specs or [
        dict(name="t2i_3workers", ipms=[10, 10, 10], batch=6, seed=1000, subseed=77, subseed_strength=0),
        dict(name="t2i_subseed_strength", ipms=[10, 10], batch=4, seed=5, subseed=9, subseed_strength=0.3),
        dict(name="t2i_remainder", ipms=[10, 10, 10], batch=8, seed=42, subseed=43, subseed_strength=0),
        dict(name="t2i_complementary", ipms=[10, 10, 10], batch=2, seed=1, subseed=2, subseed_strength=0),
        dict(name="t2i_thin", ipms=[10, 10, 10], batch=4, seed=10, subseed=20, subseed_strength=0, thin=True),
        dict(name="t2i_slow_master", ipms=[1, 30, 30], batch=8, seed=100, subseed=200, subseed_strength=0),
        dict(name="t2i_n_iter2", ipms=[10, 10], batch=4, seed=7, subseed=8, subseed_strength=0, n_iter=2),
        dict(name="t2i_stepscale", ipms=[20, 20, 0.5], batch=3, seed=3, subseed=4, subseed_strength=0, step_scaling=True),
        dict(name="t2i_single_image_reply", ipms=[10, 10, 10], batch=3, seed=11, subseed=12, subseed_strength=0),
    ]

    def png_b64(val):
        img = Image.new("RGB", (8, 8), (val % 256, (val * 3) % 256, (val * 7) % 256))
        buf = io.BytesIO()
        img.save(buf, format="PNG")
        return base64.b64encode(buf.getvalue()).decode()

    for spec in specs:
        rec = {"name": spec["name"], "requests": []}
        script = script_cls()
        w = _mk_world(mods, spec["ipms"], thin=spec.get("thin", False), step_scaling=spec.get("step_scaling", False))
        script_cls.world = w
        script.world = w
        # keep update() from reloading config / benchmarking: workers are already benchmarked
        w.benchmark = lambda *a, **k: None

        def fake_request(self, payload, option_payload, sync_options, _rec=rec):
            n = payload["batch_size"] * payload["n_iter"]
            _rec["requests"].append({"label": self.label, "batch_size": payload["batch_size"], "seed": payload["seed"],
                                     "subseed": payload["subseed"], "steps": payload["steps"],
                                     "n_iter": payload["n_iter"], "sync": bool(sync_options),
                                     "option_payload": option_payload, "has_scripts_value": "scripts_value" in payload,
                                     "alwayson_scripts": payload.get("alwayson_scripts")})
            seeds = [payload["seed"] + i for i in range(n)]
            subseeds = [payload["subseed"] + i for i in range(n)]
            info = {"all_seeds": seeds, "all_subseeds": subseeds, "all_prompts": [payload["prompt"]] * n,
                    "all_negative_prompts": [payload["negative_prompt"]] * n,
                    "infotexts": [f"seed {s}" for s in seeds], "seed": seeds[0], "subseed": subseeds[0],
                    "prompt": payload["prompt"], "negative_prompt": payload["negative_prompt"]}
            self.response = {"images": [png_b64(s) for s in seeds],
                             "parameters": {"batch_size": payload["batch_size"], "n_iter": payload["n_iter"]},
                             "info": json.dumps(info)}
            self.jobs_requested += 1

        for wk in w._workers:
            wk.request = types.MethodType(fake_request, wk)

        def master_gen(p, n):
            return [torch.full((3, 8, 8), 0.5) for _ in range(p.batch_size)]

        processing.MASTER_GENERATOR = master_gen
        original_inner = processing.process_images_inner
        script.args_from, script.args_to = 0, 0
        p = processing.StableDiffusionProcessingTxt2Img(
            prompt="a prompt", negative_prompt="neg", seed=spec["seed"], subseed=spec["subseed"],
            subseed_strength=spec["subseed_strength"], batch_size=spec["batch"], n_iter=spec.get("n_iter", 1), steps=20,
            sampler_name="Euler a", scripts=mscripts.ScriptRunner([script]), script_args=[])
        try:
            processed = processing.process_images(p)
            rec["p_batch_size_after"] = p.batch_size
            rec["n_images"] = len(processed.images)
            rec["image_kinds"] = sorted({type(im).__name__ for im in processed.images})
            rec["seeds"] = list(p.seeds)
            rec["subseeds"] = list(p.subseeds)
            rec["n_prompts"] = len(p.prompts)
            rec["infotexts"] = list(processed.infotexts)
            rec["jobs"] = _jobs_view(w)
            rec["gallery_maps"] = [list(j.gallery_map) for j in w.jobs]
            rec["responses_cleared"] = all(wk.response is None for wk in w.get_workers())
            rec["inner_restored"] = processing.process_images_inner is original_inner
            rec["requests"].sort(key=lambda r: r["label"])
        except Exception as e:
            rec["error"] = f"{type(e).__name__}: {e}"
        finally:
            processing.process_images_inner = original_inner
        results.append(rec)
    return results


def run_all(mods, script_cls):
    return {
        "optimize": [dict(spec=copy.deepcopy(s), result=run_optimize(mods, s)) for s in optimize_specs()],
        "eta": run_eta(mods),
        "fsm": run_fsm(mods),
        "misc": run_misc(mods),
        "dispatch": run_dispatch(mods, script_cls),
    }
