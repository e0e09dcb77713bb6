import types


class _AlwaysVisible:
    pass


AlwaysVisible = _AlwaysVisible()


class PostprocessBatchListArgs:
    def __init__(self, images):
        self.images = images


class Script:
    alwayson = True
    args_from = 0
    args_to = 0

    def title(self):
        raise NotImplementedError

    def show(self, is_img2img):
        return True

    def ui(self, is_img2img):
        return []


class ScriptRunner:
    """The subset of sdwui's ScriptRunner the extension touches."""

    def __init__(self, scripts=None):
        self.scripts = list(scripts or [])
        self.alwayson_scripts = self.scripts

    def _args(self, p, s):
        return list(p.script_args[s.args_from:s.args_to])

    def before_process(self, p):
        for s in self.scripts:
            if hasattr(s, "before_process"):
                s.before_process(p, *self._args(p, s))

    def postprocess_batch_list(self, p, pp, **kwargs):
        for s in self.scripts:
            if hasattr(s, "postprocess_batch_list"):
                s.postprocess_batch_list(p, pp, *self._args(p, s), **kwargs)

    def postprocess(self, p, processed):
        for s in self.scripts:
            if hasattr(s, "postprocess"):
                s.postprocess(p, processed, *self._args(p, s))
