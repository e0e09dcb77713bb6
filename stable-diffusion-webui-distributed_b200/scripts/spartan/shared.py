"""Logging and benchmark constants shared by the scheduler modules.

Mirrors the surface of the reference's scripts/spartan/shared.py (logger "distributed", GUI ring buffer of the 16
latest messages :43-56, warmup_samples/samples :63-64, module-level `benchmark_payload` :77) without depending on
`rich`; the rotating file log goes next to the config file instead of into the extension directory.
"""
import logging
import os
from logging.handlers import RotatingFileHandler
from pathlib import Path

from modules.shared import cmd_opts

extension_path = Path(__file__).resolve().parent.parent.parent

LOG_LEVEL = "DEBUG" if getattr(cmd_opts, "distributed_debug", False) else "INFO"
logger = logging.getLogger("distributed")
logger.propagate = False
logger.setLevel(LOG_LEVEL)


class GuiHandler(logging.Handler):
    """Keeps the newest messages for the status tab (reference: 16-entry ring, newest first on dump)."""
    capacity = 16

    def __init__(self):
        super().__init__()
        self.messages = []
        self.setFormatter(logging.Formatter("%(levelname)s - %(message)s"))

    def emit(self, record):
        self.messages.append(self.format(record))
        while len(self.messages) >= self.capacity:
            self.messages.pop(0)

    def dump(self) -> str:
        return "".join(f"{m}\n" for m in reversed(self.messages))


def _install_handlers():
    if getattr(logger, "_b200_handlers", False):
        return
    console = logging.StreamHandler()
    console.setFormatter(logging.Formatter("DISTRIBUTED | %(levelname)-8s %(message)s"))
    logger.addHandler(console)
    try:
        log_dir = os.path.dirname(str(getattr(cmd_opts, "distributed_config", "") or "")) or str(extension_path)
        fh = RotatingFileHandler(os.path.join(log_dir, "distributed.log"), maxBytes=10_000_000, backupCount=1)
        fh.setFormatter(logging.Formatter("%(asctime)s - %(levelname)s - %(message)s"))
        logger.addHandler(fh)
    except OSError:
        pass  # read-only install location: console + GUI ring only
    logger._b200_handlers = True


gui_handler = GuiHandler()
_install_handlers()
logger.addHandler(gui_handler)

warmup_samples = 2  # un-timed generations before a benchmark sample counts
samples = 3         # timed benchmark generations per worker

benchmark_payload = None  # set by World.load_config(): a pmodels.Benchmark_Payload
