"""On-disk schema of distributed-config.json.

Same JSON as the reference's pydantic-v1 models (scripts/spartan/pmodels.py:4-46: `workers` = list of
{label: Worker_Model}, `benchmark_payload`, `job_timeout`, `enabled`, `enabled_i2i`, `complement_production`,
`step_scaling`), written for the installed pydantic 2.x, with the v1 spellings `.dict()` / `.json(indent=)` kept
because the scheduler API above it uses them (world.py:190, :711, :721).
"""
import json
from typing import Any, Dict, List, Optional

from pydantic import BaseModel, Field


class _Compat(BaseModel):
    def dict(self, **kw):  # noqa: A003 - v1 spelling
        return self.model_dump(**kw)

    def json(self, indent=None, **kw):  # v1 spelling; v1 separators so files diff cleanly against the reference's
        return json.dumps(self.model_dump(mode="json"), indent=indent)

    def keys(self):  # the reference does dict(benchmark_payload) (worker.py:547)
        return self.model_dump().keys()

    def __getitem__(self, k):
        return getattr(self, k)


class Benchmark_Payload(_Compat):
    prompt: str = Field(default="A herd of cows grazing at the bottom of a sunny valley")
    negative_prompt: str = Field(default="")
    steps: int = Field(default=20)
    width: int = Field(default=512)
    height: int = Field(default=512)
    batch_size: int = Field(default=1)


class Worker_Model(_Compat):
    avg_ipm: Optional[float] = Field(default=None, ge=0, title="Average Speed",
                                     description="the speed of a device measured in ipm(images per minute)")
    master: bool = Field(default=False, description="whether or not an instance is the master(local) node")
    address: Optional[str] = Field(default="localhost")
    port: Optional[int] = Field(default=7860, ge=0, le=65535)
    eta_percent_error: Optional[List[float]] = Field(default_factory=list, title="ETA Percent Error")
    tls: Optional[bool] = Field(default=False, title="Transport Layer Security")
    state: Optional[Any] = Field(default=1, description="The last known state of this worker")
    user: Optional[str] = Field(default=None)
    password: Optional[str] = Field(default=None)
    pixel_cap: Optional[int] = Field(default=-1, description="Max pixels one worker handles at once; -1 = no limit")

    @classmethod
    def from_worker(cls, worker) -> "Worker_Model":
        src = worker.__dict__
        vals = {name: src[name] for name in cls.model_fields if name in src}
        st = vals.get("state")
        if st is not None and hasattr(st, "value"):
            vals["state"] = st.value
        return cls(**vals)


class ConfigModel(_Compat):
    workers: List[Dict[str, Worker_Model]]
    benchmark_payload: Benchmark_Payload = Field(default_factory=Benchmark_Payload,
                                                 description="the payload used when benchmarking a node")
    job_timeout: Optional[int] = Field(default=3)
    enabled: Optional[bool] = Field(default=True)
    enabled_i2i: Optional[bool] = Field(default=True)
    complement_production: Optional[bool] = Field(default=True)
    step_scaling: Optional[bool] = Field(default=False)
