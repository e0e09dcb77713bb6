import os
import tempfile
import types

_cfg_dir = os.environ.get("HOSTSTUB_CONFIG_DIR") or tempfile.mkdtemp(prefix="hoststub_")

cmd_opts = types.SimpleNamespace(
    distributed_skip_verify_remotes=False,
    distributed_debug=False,
    distributed_config=os.path.join(_cfg_dir, "distributed-config.json"),
    distributed_remotes_autosave=False,
    distributed_remotes=None,
    port=None,
)


class _Label:
    def __init__(self):
        self.onchange = lambda: None


class _Opts:
    def __init__(self):
        self.return_grid = False
        self.sd_model_checkpoint = "synthetic-sd15.safetensors [0000000000]"
        self.sd_vae = None
        self.data = {"sd_model_checkpoint": self.sd_model_checkpoint, "sd_vae": None}
        self.data_labels = {"sd_model_checkpoint": _Label()}


opts = _Opts()


class _State:
    def __init__(self):
        self.interrupted = False
        self.textinfo = ""
        self.job_count = 0
        self.job = ""

    def begin(self, job=""):
        self.job = job
        self.interrupted = False

    def end(self):
        self.job = ""


state = _State()
device = "cpu"
