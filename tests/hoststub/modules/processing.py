import random
import sys

from modules import scripts as _scripts


def fix_seed(p):
    if p.seed is None or p.seed == -1 or p.seed == "":
        p.seed = int(random.randrange(4294967294))
    if p.subseed is None or p.subseed == -1 or p.subseed == "":
        p.subseed = int(random.randrange(4294967294))


class StableDiffusionProcessing:
    def __init__(self, **kw):
        self.prompt = ""
        self.negative_prompt = ""
        self.seed = -1
        self.subseed = -1
        self.subseed_strength = 0
        self.batch_size = 1
        self.n_iter = 1
        self.steps = 20
        self.width = 512
        self.height = 512
        self.sampler_name = "Euler a"
        self.cfg_scale = 7.0
        self.s_tmax = float("inf")
        self.do_not_save_samples = True
        self.scripts = _scripts.ScriptRunner([])
        self.script_args = []
        self.scripts_value = None
        self.seeds, self.subseeds, self.prompts, self.negative_prompts = [], [], [], []
        for k, v in kw.items():
            setattr(self, k, v)


class StableDiffusionProcessingTxt2Img(StableDiffusionProcessing):
    pass


class StableDiffusionProcessingImg2Img(StableDiffusionProcessing):
    def __init__(self, init_images=None, denoising_strength=0.75, **kw):
        super().__init__(**kw)
        self.init_images = init_images or []
        self.denoising_strength = denoising_strength


class Processed:
    def __init__(self, p, images_list, seed=-1, info="", subseed=None, all_prompts=None, all_negative_prompts=None,
                 all_seeds=None, all_subseeds=None, index_of_first_image=0, infotexts=None, comments=""):
        self.images = images_list
        self.prompt = p.prompt
        self.negative_prompt = p.negative_prompt
        self.seed = seed
        self.subseed = subseed
        self.info = info
        self.all_prompts = all_prompts or []
        self.all_negative_prompts = all_negative_prompts or []
        self.all_seeds = all_seeds or []
        self.all_subseeds = all_subseeds or []
        self.infotexts = infotexts or []


# tests / bench plug the master's generator in here: f(p, batch_number) -> list of CHW float tensors in [0,1]
MASTER_GENERATOR = None


def process_images_inner(p) -> Processed:
    """The host's own generation loop for the master's share (stand-in for upstream process_images_inner)."""
    fix_seed(p)
    p.seeds = [p.seed + i for i in range(p.batch_size)]
    p.subseeds = [p.subseed + i for i in range(p.batch_size)]
    p.prompts = [p.prompt] * p.batch_size
    p.negative_prompts = [p.negative_prompt] * p.batch_size
    out_images, infotexts = [], []
    all_seeds, all_subseeds, all_prompts, all_neg = [], [], [], []
    for n in range(p.n_iter):
        if MASTER_GENERATOR is None:
            raise RuntimeError("hoststub: modules.processing.MASTER_GENERATOR is not set")
        images = list(MASTER_GENERATOR(p, n)) if p.batch_size > 0 else []
        pp = _scripts.PostprocessBatchListArgs(images)
        p.scripts.postprocess_batch_list(p, pp, batch_number=n)
        out_images.extend(pp.images)
        all_seeds, all_subseeds, all_prompts, all_neg = p.seeds, p.subseeds, p.prompts, p.negative_prompts
    infotexts = [f"{p.prompt}\nSteps: {p.steps}, Sampler: {p.sampler_name}, Seed: {s}" for s in all_seeds]
    while len(infotexts) < len(out_images):
        infotexts.append("")
    res = Processed(p, out_images, p.seed, info="", subseed=p.subseed, all_prompts=list(all_prompts),
                    all_negative_prompts=list(all_neg), all_seeds=list(all_seeds), all_subseeds=list(all_subseeds),
                    infotexts=infotexts)
    p.scripts.postprocess(p, res)
    return res


def process_images(p) -> Processed:
    p.scripts.before_process(p)
    # resolved at call time: the extension monkey-patches modules.processing.process_images_inner
    return sys.modules[__name__].process_images_inner(p)
