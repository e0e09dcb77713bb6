"""Scheduler / dispatch parity against golden vectors produced by EXECUTING THE REFERENCE
(tests/golden/gen_scheduler_golden.py -> tests/golden/scheduler_golden.json): 238 optimize_jobs worlds, 68 ETA
cases, 138 state-machine traces, and the full before_process -> postprocess_batch_list -> postprocess hook chain
with recorded worker payloads (seed offsets, option payload, script args, gallery maps, infotexts).

Integer / structural results must be identical; floats agree to 1e-12 relative (same arithmetic, same order).
"""
import json
import math
import os

import pytest

import scheduler_scenarios as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scheduler_golden.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def mods():
    import logging
    from scripts.spartan import pmodels, shared, worker, world
    logging.getLogger("distributed").setLevel(logging.CRITICAL + 1)
    return world, worker, shared, pmodels


def _close(a, b):
    if isinstance(a, float) or isinstance(b, float):
        if a is None or b is None:
            return a is b
        return math.isclose(a, b, rel_tol=1e-12, abs_tol=1e-12)
    if isinstance(a, list) and isinstance(b, list):
        return len(a) == len(b) and all(_close(x, y) for x, y in zip(a, b))
    if isinstance(a, dict) and isinstance(b, dict):
        return a.keys() == b.keys() and all(_close(a[k], b[k]) for k in a)
    return a == b


def test_optimize_jobs_matches_reference(mods, golden):
    bad = []
    for entry in golden["optimize"]:
        got = S.run_optimize(mods, entry["spec"])
        want = entry["result"]
        got = json.loads(json.dumps(got))
        if not _close(got, want):
            bad.append((entry["spec"]["name"], got, want))
    assert not bad, f"{len(bad)} of {len(golden['optimize'])} scenarios differ; first: {bad[0]}"
    assert len(golden["optimize"]) >= 230


def test_appendix_a_vectors(mods):
    """SURVEY.md App. A, spelled out (independent of the JSON file)."""
    want = {
        "equal8x4": [["master", 2, False, None], ["w1", 2, False, None], ["w2", 2, False, None], ["w3", 2, False, None]],
        "remainder": [["master", 3, False, None], ["w1", 3, False, None], ["w2", 2, False, None]],
        "b2_world3": [["master", 1, False, None], ["w1", 1, False, None], ["w2", 1, True, None]],
        "slow_worker": [["master", 5, False, None], ["w1", 4, False, None]],
        "slow_worker_stepscale": [["master", 2, False, None], ["w1", 1, False, None], ["w2", 1, True, 1.0]],
        "slow_master": [["master", 0, True, None], ["w1", 4, False, None], ["w2", 4, False, None]],
        "pixel_cap": [["master", 4, False, None], ["w1", 4, False, None]],
        "ddim_hr": [["master", 4, False, None]],
        "thin": [["w1", 2, False, None], ["w2", 2, False, None]],
    }
    specs = {s["name"]: s for s in S.optimize_specs()}
    for name, jobs in want.items():
        res = S.run_optimize(mods, specs[name])
        assert res["error"] is None and res["jobs"] == jobs, (name, res)
    assert S.run_optimize(mods, specs["slow_master"])["bypass"] is True
    assert S.run_optimize(mods, specs["equal32x8"])["jobs"] == [[lbl, 4, False, None] for lbl in
                                                                 ["master"] + [f"w{i}" for i in range(1, 8)]]


def test_eta_matches_reference(mods, golden):
    got = json.loads(json.dumps(S.run_eta(mods)))
    assert len(got) == len(golden["eta"])
    for g, w in zip(got, golden["eta"]):
        assert _close(g["eta"], w["eta"]) and _close(g["mpe"], w["mpe"]), (g, w)


def test_state_machine_matches_reference(mods, golden):
    got = S.run_fsm(mods)
    assert got == golden["fsm"]


def test_misc_matches_reference(mods, golden):
    assert json.loads(json.dumps(S.run_misc(mods))) == golden["misc"]


def test_dispatch_hooks_match_reference(mods, golden):
    from scripts.distributed import DistributedScript
    got = json.loads(json.dumps(S.run_dispatch(mods, DistributedScript)))
    for g, w in zip(got, golden["dispatch"]):
        assert g["name"] == w["name"]
        if "error" in w:
            # reference bug: the zero-master bypass calls the hook without batch_number -> KeyError (world.py:570 vs
            # distributed.py:330).  Ours completes and returns exactly the remote images.
            assert w["name"] == "t2i_slow_master" and "error" not in g
            assert g["n_images"] == 8 and g["seeds"] == list(range(100, 108))
            continue
        assert "error" not in g, g
        for key in ("requests", "seeds", "subseeds", "jobs", "gallery_maps", "n_images", "n_prompts",
                    "p_batch_size_after", "image_kinds", "responses_cleared", "inner_restored", "infotexts"):
            assert g[key] == w[key], (g["name"], key, g[key], w[key])


def test_config_roundtrip_same_json_schema(mods, tmp_path):
    """distributed-config.json written by save_config has the reference's keys (App. A) and loads back."""
    world_mod, worker_mod, sh, pmodels = mods
    w = S._mk_world(mods, [12.5, 30.0])
    w.config_path = str(tmp_path / "distributed-config.json")
    w["w1"].eta_percent_error = [1.5, -2.0]
    w["w1"].pixel_cap = 1048576
    w.job_timeout = 7
    w.step_scaling = True
    w.save_config()
    raw = json.loads(open(w.config_path).read())
    assert list(raw.keys()) == ["workers", "benchmark_payload", "job_timeout", "enabled", "enabled_i2i",
                                "complement_production", "step_scaling"]
    assert list(raw["workers"][1]["w1"].keys()) == ["avg_ipm", "master", "address", "port", "eta_percent_error", "tls",
                                                    "state", "user", "password", "pixel_cap"]
    assert raw["workers"][1]["w1"]["state"] == 1 and raw["benchmark_payload"]["steps"] == 20
    w2 = world_mod.World(verify_remotes=False)
    w2.config_path = w.config_path
    w2.load_config()
    assert w2["w1"].avg_ipm == 30.0 and w2["w1"].eta_percent_error == [1.5, -2.0] and w2["w1"].pixel_cap == 1048576
    assert w2.job_timeout == 7 and w2.step_scaling is True and w2.master().avg_ipm == 12.5
    # legacy workers.json translation (reference world.py:632-649)
    w3 = world_mod.World(verify_remotes=False)
    w3.config_path = str(tmp_path / "missing.json")
    w3.old_config_path = str(tmp_path / "workers.json")
    open(w3.old_config_path, "w").write(json.dumps({"old1": {"avg_ipm": 5.0, "port": 7861}}))
    cfg = w3.config()
    assert cfg["workers"] == [{"old1": {"avg_ipm": 5.0, "port": 7861, "address": "localhost"}}]
