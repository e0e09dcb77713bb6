"""World / Job — the scheduler that shards one generation request's batch over the workers.

Drop-in for the reference's scripts/spartan/world.py (`Job` :37-72, `World` :75, `benchmark` :199-278,
`make_jobs` :378, `update` :394, `get_workers` :405, `optimize_jobs` :418-601, `config/load_config/save_config`
:616-722, `ping_remotes` :724): same public methods, attributes and results — tests/test_scheduler_parity.py pins
238 scenarios against vectors produced by executing the reference itself.

What is new: `add_local_gpus()` registers one `LocalGPUWorker` per CUDA device, and `benchmark()` for those measures
the local executor directly (it/s -> ipm) instead of timing HTTP round trips, so `eta()` / `optimize_jobs()` balance
over the GPUs of the box with the reference's own arithmetic.
"""
import concurrent.futures
import json
import os
import time
from threading import Thread
from typing import List, Optional

import gradio
import modules.shared as shared
from modules import processing, progress
from modules.call_queue import queue_lock
from modules.images import image_grid
from modules.processing import StableDiffusionProcessingTxt2Img, process_images
from modules.scripts import PostprocessBatchListArgs

from . import shared as sh
from .pmodels import Benchmark_Payload, ConfigModel
from .shared import extension_path, logger
from .worker import State, Worker


class NotBenchmarked(Exception):
    """an operation needed benchmark statistics that do not exist yet"""


class Job:
    """How many images one worker contributes to the current request."""

    def __init__(self, worker: Worker, batch_size: int):
        self.worker: Worker = worker
        self.batch_size: int = batch_size
        self.complementary: bool = False
        self.step_override = None
        self.thread = None
        self.gallery_map: List[int] = []

    def __str__(self):
        head = "(complementary) " if self.complementary else ""
        return f"{head}Job: {self.batch_size} image(s) owned by '{self.worker.label}'. Rate: {self.worker.avg_ipm:0.2f} ipm"

    def add_work(self, payload: dict, batch_size: int = 1) -> bool:
        """grow the job unless that would exceed the worker's pixel cap"""
        cap = self.worker.pixel_cap
        if cap != -1:
            wanted = (self.batch_size + batch_size) * (payload["width"] * payload["height"])
            if wanted > cap:
                logger.debug(f"worker {self.worker.label} hit pixel cap ({wanted} > cap: {cap})")
                return False
        self.batch_size += batch_size
        return True


class World:
    """All workers (master included) plus the jobs of the request being processed."""

    config_path = shared.cmd_opts.distributed_config
    old_config_path = worker_info_path = extension_path.joinpath("workers.json")

    def __init__(self, verify_remotes: bool = True):
        self.p = None
        self.master_worker = Worker(master=True)
        self._workers: List[Worker] = [self.master_worker]
        self.jobs: List[Job] = []
        self.job_timeout: int = 3
        self.initialized: bool = False
        self.verify_remotes = verify_remotes
        self.thin_client_mode = False
        self.enabled = True
        self.enabled_i2i = True
        self.is_dropdown_handler_injected = False
        self.complement_production = True
        self.step_scaling = False

    # ------------------------------------------------------------------ registry
    def __getitem__(self, label: str) -> Optional[Worker]:
        return next((w for w in self._workers if w.label == label), None)

    def __repr__(self):
        return f"{len(self._workers)} workers"

    def __str__(self):
        return "".join(f"{job}\n" for job in self.jobs)

    def master(self) -> Worker:
        return self.master_worker

    def master_job(self) -> Job:
        for job in self.jobs:
            if job.worker.master:
                return job
        raise Exception("Master job not found")

    def add_worker(self, **kwargs):
        """register (or update, by label) a worker; refuses a socket identical to the master's"""
        if not kwargs.get("master"):
            m = self.master()
            if kwargs["address"] == m.address and kwargs["port"] == m.port:
                logger.error(f"refusing to add worker {kwargs['label']} as its socket definition({m.address}:{m.port}) matches master")
                return None
        existing = self[kwargs["label"]]
        if existing is None:
            existing = Worker(**kwargs)
            self._workers.append(existing)
        else:
            for name, value in kwargs.items():
                setattr(existing, name, value)
        return existing

    def add_local_gpus(self, engine_factory, devices=None, avg_ipm: float = 0.0):
        """one LocalGPUWorker per CUDA device of this box (labels gpu0..gpuN-1)"""
        from .local_worker import LocalGPUWorker
        import torch
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        added = []
        for idx in devices:
            label = f"gpu{idx}"
            w = self[label]
            if w is None:
                w = LocalGPUWorker(device_index=idx, engine_factory=engine_factory, label=label, avg_ipm=avg_ipm)
                self._workers.append(w)
            added.append(w)
        return added

    def get_workers(self) -> List[Worker]:
        usable = []
        for w in self._workers:
            if w.avg_ipm is not None and w.avg_ipm <= 0:
                logger.warning(f"config reports invalid speed (0 ipm) for worker '{w.label}'\nplease re-benchmark")
                continue
            if w.master and self.thin_client_mode:
                continue
            if w.state not in (State.UNAVAILABLE, State.DISABLED):
                usable.append(w)
        return usable

    def size(self) -> int:
        return len(self.get_workers())

    def default_batch_size(self) -> int:
        """images per worker under a perfectly even split"""
        return self.p.batch_size // self.size()

    def interrupt_remotes(self):
        for w in self.get_workers():
            if not w.master:
                Thread(target=w.interrupt, args=()).start()

    def refresh_checkpoints(self):
        for w in self.get_workers():
            if not w.master:
                Thread(target=w.refresh_checkpoints, args=()).start()

    def restart_all(self):
        for w in self._workers:
            w.restart()

    # ------------------------------------------------------------------ benchmark
    def sample_master(self) -> float:
        """seconds the host itself needs for one benchmark payload"""
        p = StableDiffusionProcessingTxt2Img()
        for key, value in sh.benchmark_payload.dict().items():
            setattr(p, key, value)
        p.do_not_save_samples = True
        begin = time.time()
        process_images(p)
        return time.time() - begin

    def benchmark(self, rebenchmark: bool = False):
        """benchmark every worker without a valid avg_ipm (all of them when `rebenchmark`)"""
        task_id = "task(distributed_bench)"
        if rebenchmark:
            for w in self._workers:
                w.benchmarked = False
            todo = list(self._workers)
        else:
            self.load_config()
            todo = []
            for w in self._workers:
                if w.avg_ipm is None or w.avg_ipm <= 0:
                    logger.debug(f"recorded speed for worker '{w.label}' is invalid")
                    todo.append(w)
                else:
                    w.benchmarked = True
        with concurrent.futures.ThreadPoolExecutor(thread_name_prefix="distributed_benchmark") as pool:
            loads = [pool.submit(w.load_options, model=shared.opts.sd_model_checkpoint, vae=shared.opts.sd_vae)
                     for w in todo if not w.master and w.state not in (State.DISABLED, State.UNAVAILABLE)]
            for fut in concurrent.futures.as_completed(loads):
                w = fut.result()
                if w is not None and getattr(w.response, "status_code", 200) != 200:
                    logger.error(f"refusing to benchmark worker '{w.label}' as it failed to load the selected model "
                                 f"'{shared.opts.sd_model_checkpoint}'")
                    todo = [x for x in todo if x != w]
            if not todo:
                return
            queue_lock.acquire()
            try:
                gradio.Info("Distributed: benchmarking in progress, please wait")
                runs = []
                for w in todo:
                    if w.state in (State.DISABLED, State.UNAVAILABLE):
                        logger.debug(f"worker '{w.label}' is {w.state.name}, refusing to benchmark")
                        continue
                    if w.model_override is not None:
                        logger.warning(f"model override is enabled for worker '{w.label}' which may result in poor optimization")
                    logger.info(f"benchmarking worker '{w.label}'")
                    if w.master:
                        if progress.current_task is None:
                            progress.add_task_to_queue(task_id)
                            progress.start_task(task_id)
                            shared.state.begin(job=task_id)
                            shared.state.job_count = sh.warmup_samples + sh.samples
                        w.benchmark(sample_function=self.sample_master)  # the host generates on the caller's thread
                    else:
                        runs.append(pool.submit(w.benchmark))
                concurrent.futures.wait(runs)
            finally:
                if progress.current_task == task_id:
                    shared.state.end()
                    progress.finish_task(task_id)
                queue_lock.release()
            logger.info("benchmarking finished")
            logger.info(self.speed_summary())
            gradio.Info("Distributed: benchmarking complete!")
            self.save_config()

    def speed_summary(self) -> str:
        ranked = sorted(self._workers, key=lambda w: w.avg_ipm, reverse=True)
        lines = ["World composition:"]
        for rank, w in enumerate(ranked, start=1):
            lines.append(f"{rank}. '{w.label}'({w}) - {w.avg_ipm:.2f} ipm")
        lines.append(f"total: ~{sum(w.avg_ipm for w in ranked):.2f} ipm")
        return "\n".join(lines)

    # ------------------------------------------------------------------ job bookkeeping
    def num_requested(self) -> int:
        return sum(job.batch_size for job in self.jobs)

    def num_gallery(self) -> int:
        return self.num_requested() * self.p.n_iter + shared.opts.return_grid

    def realtime_jobs(self) -> List[Job]:
        return [j for j in self.jobs
                if j.worker.benchmarked is not False and j.worker.avg_ipm is not None and j.complementary is False]

    def slowest_realtime_job(self) -> Job:
        return min(self.realtime_jobs(), key=lambda j: j.worker.avg_ipm)

    def fastest_realtime_job(self) -> Job:
        # first of the fastest (stable), like sorted(..., reverse=True)[0]
        return sorted(self.realtime_jobs(), key=lambda j: j.worker.avg_ipm, reverse=True)[0]

    def job_stall(self, worker: Worker, payload: dict, batch_size: int = None) -> float:
        """seconds the gallery would wait for `worker` beyond the fastest realtime worker"""
        fastest = self.fastest_realtime_job().worker
        if worker == fastest:
            return 0
        return worker.eta(payload=payload, quiet=True, batch_size=batch_size) - \
            fastest.eta(payload=payload, quiet=True, batch_size=batch_size)

    def make_jobs(self):
        """one job of default_batch_size() images per usable worker"""
        self.jobs = []
        share = self.default_batch_size()
        for w in self.get_workers():
            if w.state in (State.DISABLED, State.UNAVAILABLE):
                continue
            if w.avg_ipm is None or w.avg_ipm <= 0:
                logger.debug(f"No recorded speed for worker '{w.label}, benchmarking'")
                w.benchmark()
            self.jobs.append(Job(worker=w, batch_size=share))

    def update(self, p):
        """prepare for the next request"""
        self.p = p
        self.benchmark()
        self.make_jobs()
        self.initialized = True

    # ------------------------------------------------------------------ optimize_jobs and its stages
    def optimize_jobs(self, payload: dict):
        """Final image count per job.  payload['batch_size'] must be default_batch_size()."""
        deferred = self._classify(payload)
        if deferred > 0:
            self._place_deferred(payload, deferred)
        self._place_remainder(payload)
        if self.complement_production:
            self._fill_complementary(payload)
        else:
            logger.debug("complementary image production is disabled")
        logger.info(self.distro_summary())
        if self.thin_client_mode is True or self.master_job().batch_size == 0:
            self._bypass_local_generation()
        # drop empty jobs, but never index 0
        for idx in range(len(self.jobs) - 1, 0, -1):
            if self.jobs[idx].batch_size < 1:
                del self.jobs[idx]

    def _classify(self, payload: dict) -> int:
        """realtime jobs keep the even share; laggards become complementary and their share is deferred"""
        share = payload["batch_size"]
        deferred = checked = 0
        for job in self.jobs:
            lag = self.job_stall(job.worker, payload=payload)
            if lag < self.job_timeout or lag == 0:
                job.batch_size = share
                checked += share
                continue
            logger.debug(f"worker '{job.worker.label}' would stall the image gallery by ~{lag:.2f}s\n")
            job.complementary = True
            if deferred + checked + share <= self.p.batch_size:
                deferred += share
            job.batch_size = 0
        return deferred

    def _place_deferred(self, payload: dict, deferred: int):
        """hand deferred images out one at a time, round-robin, to jobs that stay within the stall budget"""
        saturated = []
        idx = 0
        while deferred > 0:
            if len(saturated) == len(self.jobs):
                logger.critical(f"all workers saturated, cannot distribute {deferred} remaining deferred image(s)")
                break
            job = self.jobs[idx]
            if self.job_stall(worker=job.worker, payload=payload, batch_size=job.batch_size + 1) < self.job_timeout:
                if job.add_work(payload, batch_size=1):
                    deferred -= 1
                else:
                    saturated.append(job)
            idx = idx + 1 if idx < len(self.jobs) - 1 else 0

    def _place_remainder(self, payload: dict):
        """images lost to the integer split go round-robin to the smallest realtime jobs"""
        left = self.p.batch_size - self.num_requested()
        if left < 1:
            return
        targets = sorted(self.realtime_jobs(), key=lambda j: j.batch_size)
        saturated = []
        while left >= 1:
            if len(saturated) >= len(self.jobs):
                logger.critical("all workers saturated, cannot fully distribute remainder of request")
                break
            for job in targets:
                if left < 1:
                    break
                if job.add_work(payload):
                    left -= 1
                else:
                    saturated.append(job)
        for job in self.jobs:  # e.g. batch 2 on 3 workers: the third still gets a (bonus) job
            if job.batch_size == 0:
                job.complementary = True

    def _fill_complementary(self, payload: dict):
        """bonus images a slow worker can finish inside the fastest worker's time + job_timeout"""
        for job in self.jobs:
            if job.complementary is False:
                continue
            fastest = self.fastest_realtime_job().worker
            for other in self.jobs:
                if other.worker.label == fastest.label:
                    slack = fastest.eta(payload=payload, batch_size=other.batch_size) + self.job_timeout
            per_image = job.worker.eta(payload=payload, batch_size=1)
            bonus = int(slack / per_image)
            logger.debug(f"worker '{job.worker.label}': {bonus} complementary image(s) = {slack:.2f}s slack "
                         f"/ {per_image:.2f}s per requested image")
            if not job.add_work(payload, batch_size=bonus):
                job.add_work(payload, batch_size=job.worker.pixel_cap // (payload["width"] * payload["height"]))
            if bonus == 0 and self.step_scaling:
                per_sample = job.worker.eta(payload=payload, batch_size=1, samples=1)
                job.add_work(payload=payload, batch_size=1)
                job.step_override = slack // per_sample
                logger.debug(f"job for '{job.worker.label}' downscaled to {job.step_override:.0f} samples "
                             f"(step reduction: {payload['steps']} -> {job.step_override:.0f})")

    def _bypass_local_generation(self):
        """thin client / master got no images: the host's inner loop is replaced by one that only collects"""
        logger.debug("bypassing local generation completely")
        world = self
        for w in self._workers:   # local-GPU workers build the PIL images this collector returns in their own job threads
            if getattr(w, "is_local_gpu", False):
                w.make_pil = True

        def process_images_inner_bypass(p) -> processing.Processed:
            from torchvision.transforms import ToPILImage
            p.seeds, p.subseeds, p.negative_prompts, p.prompts = [], [], [], []
            pp = PostprocessBatchListArgs(images=[])
            world.p.scripts.postprocess_batch_list(p, pp, batch_number=p.n_iter - 1)
            processed = processing.Processed(p, [], p.seed, info="")
            processed.all_prompts, processed.all_seeds = p.prompts, p.seeds
            processed.all_subseeds, processed.all_negative_prompts = p.subseeds, p.negative_prompts
            processed.infotexts = [""] * world.num_requested()
            to_pil = ToPILImage()

            def as_pil(im):
                # local-GPU fast lane: the uint8 HWC bytes came along with the float tensor (scripts/distributed.py
                # api_to_internal); a script that replaced pp.images[i] in postprocess_batch_list loses the attribute and
                # takes the reference's ToPILImage path
                ready = getattr(im, "b200sd_pil", None)
                if ready is not None:
                    return ready
                u8 = getattr(im, "b200sd_u8", None)
                if u8 is not None:
                    from PIL import Image
                    return Image.fromarray(u8.numpy())   # ("RGB" cannot be memory-mapped by PIL: one 0.75 ms copy per image)
                return to_pil(im)

            processed.images = [as_pil(im) for im in pp.images]
            world.p.scripts.postprocess(p, processed)
            if shared.opts.return_grid and len(processed.images) > 1:
                processed.images.insert(0, image_grid(processed.images, len(processed.images)))
                processed.infotexts.insert(0, processed.infotexts[0])
            return processed

        processing.process_images_inner = process_images_inner_bypass

    def distro_summary(self) -> str:
        total = self.num_requested()
        extra = total - self.p.batch_size
        text = f"Job distribution:\n{self.p.batch_size} * {self.p.n_iter} iteration(s)"
        if extra > 0:
            text += f" + {extra} complementary"
        text += f": {total * self.p.n_iter} images total\n"
        for job in self.jobs:
            text += f"'{job.worker.label}' - {job.batch_size * self.p.n_iter} image(s) @ {job.worker.avg_ipm:.2f} ipm\n"
        return text

    # ------------------------------------------------------------------ config persistence
    def config(self) -> dict:
        """parsed distributed-config.json; translates a legacy workers.json; creates an empty file if none"""
        if not os.path.exists(self.config_path):
            msg = f"Config was not found at '{self.config_path}'"
            logger.error(msg)
            gradio.Warning("Distributed: " + msg)
            if os.path.exists(self.old_config_path):
                with open(self.old_config_path) as f:
                    legacy = json.load(f)
                translated = {"workers": [], "benchmark_payload": legacy.pop("benchmark_payload", None)}
                for label, fields in legacy.items():
                    fields["address"] = "localhost"
                    translated["workers"].append({label: fields})
                logger.info("translated legacy config")
                return translated
            open(self.config_path, "w").close()
            logger.info(f"Generated new config file at '{self.config_path}'")
        with open(self.config_path, "r") as f:
            try:
                return json.load(f)
            except json.decoder.JSONDecodeError:
                logger.error("config is corrupt or invalid JSON, unable to load")

    def load_config(self):
        raw = self.config()
        if raw is None:
            logger.debug("cannot parse null config (present but empty config file?)\ngenerating defaults for config")
            sh.benchmark_payload = Benchmark_Payload()
            self.save_config()
            return
        if raw.get("benchmark_payload") is None:
            raw = {k: v for k, v in raw.items() if k != "benchmark_payload"}
        cfg = ConfigModel(**raw)
        for entry in cfg.workers:
            label, model = next(iter(entry.items()))
            fields = dict(model.dict())
            fields["label"] = label
            fields["verify_remotes"] = self.verify_remotes
            state = State(fields["state"])
            fields["state"] = state if state in (State.DISABLED, State.UNAVAILABLE) else State.IDLE
            existing = self[label]
            if existing is not None and getattr(existing, "is_local_gpu", False):
                for name in ("avg_ipm", "eta_percent_error", "pixel_cap", "state"):
                    setattr(existing, name, fields[name])
                continue
            self.add_worker(**fields)
        sh.benchmark_payload = Benchmark_Payload(**cfg.benchmark_payload.dict())
        self.job_timeout = cfg.job_timeout
        self.enabled = cfg.enabled
        self.enabled_i2i = cfg.enabled_i2i
        self.complement_production = cfg.complement_production
        self.step_scaling = cfg.step_scaling
        logger.debug(f"config loaded from '{os.path.abspath(self.config_path)}'")

    def save_config(self):
        cfg = ConfigModel(
            workers=[{w.label: w.model.dict()} for w in self._workers],
            benchmark_payload=sh.benchmark_payload if sh.benchmark_payload is not None else Benchmark_Payload(),
            job_timeout=self.job_timeout, enabled=self.enabled, enabled_i2i=self.enabled_i2i,
            complement_production=self.complement_production, step_scaling=self.step_scaling)
        with open(self.config_path, "w+") as f:
            f.write(cfg.json(indent=3))
        logger.debug("config saved")

    # ------------------------------------------------------------------ reachability / model sync
    def ping_remotes(self, indiscriminate: bool = False):
        """mark unreachable workers UNAVAILABLE, reachable ones IDLE (and learn their alwayson scripts)"""
        for w in self._workers:
            if w.master:
                continue
            if w.state == State.DISABLED:
                logger.debug(f"refusing to ping disabled worker '{w.label}'")
                continue
            if w.state != State.UNAVAILABLE and not indiscriminate:
                continue
            logger.debug(f"checking if worker '{w.label}' is reachable...")
            if w.reachable():
                if w.queried and w.state == State.IDLE:
                    continue
                w.supported_scripts = w.query_scripts() if hasattr(w, "query_scripts") else self._script_info(w)
                logger.info(f"worker '{w.label}' is online")
                gradio.Info(f"Distributed: worker '{w.label}' is online")
                w.set_state(State.IDLE, expect_cycle=True)
            else:
                msg = f"worker '{w.label}' is unreachable"
                code = getattr(w.response, "status_code", None)
                if code is not None:
                    msg += f" <{code}>"
                logger.info(msg)
                gradio.Warning("Distributed: " + msg)
                w.set_state(State.UNAVAILABLE)
            self.save_config()

    @staticmethod
    def _script_info(w: Worker) -> dict:
        found = {"txt2img": [], "img2img": []}
        r = w.session.get(url=w.full_url("script-info"))
        if r.status_code != 200:
            logger.error(f"failed to query script-info for worker '{w.label}': {r}")
            return found
        for entry in r.json():
            name = entry.get("name", None)
            if name is not None and entry.get("is_alwayson", False):
                found["img2img" if entry.get("is_img2img", False) else "txt2img"].append(name)
        return found

    def inject_model_dropdown_handler(self):
        """wrap the checkpoint dropdown's onchange so model switches propagate to the workers"""
        if (self.config() or {}).get("enabled", False):
            return
        if self.is_dropdown_handler_injected:
            return
        dropdown = shared.opts.data_labels.get("sd_model_checkpoint")
        original = dropdown.onchange

        def on_change():
            for w in self.get_workers():
                if w.master or w.model_override is not None:
                    continue
                Thread(target=w.load_options, args=(shared.opts.sd_model_checkpoint,),
                       name=f"{w.label}_on_dropdown_model_load").start()
            original()

        dropdown.onchange = on_change
        self.is_dropdown_handler_injected = True
