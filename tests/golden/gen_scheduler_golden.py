"""Generates tests/golden/scheduler_golden.json by executing the UNMODIFIED reference scheduler/dispatcher
(/root/reference/scripts) under the host stub.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/gen_scheduler_golden.py

The reference needs: pydantic v1 API (aliased to pydantic.v1), a `modules` package + gradio (tests/hoststub), and its
rotating log file redirected off the read-only tree (scripts/spartan/shared.py:33-34).
"""
import json
import logging.handlers
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
REF = os.environ.get("REFERENCE_DIR", "/root/reference")

tmp = tempfile.mkdtemp(prefix="refgold_")
os.environ["HOSTSTUB_CONFIG_DIR"] = tmp
sys.path[:0] = [os.path.join(TESTS, "hoststub"), REF, TESTS]

import pydantic.v1  # noqa: E402

sys.modules["pydantic"] = pydantic.v1

_Orig = logging.handlers.RotatingFileHandler


class _Redirected(_Orig):
    def __init__(self, filename, *a, **k):
        super().__init__(os.path.join(tmp, os.path.basename(str(filename))), *a, **k)


logging.handlers.RotatingFileHandler = _Redirected

from scripts.spartan import pmodels, shared, worker, world  # noqa: E402
from scripts.distributed import DistributedScript  # noqa: E402

import scheduler_scenarios as S  # noqa: E402

logging.getLogger("distributed").setLevel(logging.CRITICAL + 1)
res = S.run_all((world, worker, shared, pmodels), DistributedScript)
res["_meta"] = {"reference": "papuSpartan/stable-diffusion-webui-distributed @ 8fd65ebd", "generator": "tests/golden/gen_scheduler_golden.py"}
out = os.path.join(HERE, "scheduler_golden.json")
with open(out, "w") as f:
    json.dump(res, f, indent=0, sort_keys=True)
print("wrote", out, {k: (len(v) if hasattr(v, "__len__") else v) for k, v in res.items()})
errs = [r["spec"]["name"] for r in res["optimize"] if r["result"]["error"]]
print("optimize errors:", errs[:20], len(errs))
print("dispatch errors:", [(r["name"], r.get("error")) for r in res["dispatch"] if "error" in r])
