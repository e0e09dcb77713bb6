"""Helpers for the GPU kernel parity tests: error summaries that localise a broken tile/descriptor."""
import json
import os

import torch

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def err_report(name, got, ref, atol, rtol, block=(32, 32)):
    got = got.float()
    ref = ref.float()
    diff = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = diff > tol
    nbad = int(bad.sum())
    rep = {
        "name": name, "shape": list(ref.shape), "max_abs_err": float(diff.max()), "ref_absmax": float(ref.abs().max()),
        "n_bad": nbad, "frac_bad": nbad / max(1, ref.numel()), "has_nan": bool(torch.isnan(got).any()),
    }
    if nbad:
        d2 = diff.reshape(-1, diff.shape[-1])
        r, c = d2.shape
        br, bc = block
        rr, cc = (r + br - 1) // br, (c + bc - 1) // bc
        pad = torch.zeros(rr * br, cc * bc, device=d2.device)
        pad[:r, :c] = (d2 > tol.reshape(-1, diff.shape[-1])).float()
        grid = pad.reshape(rr, br, cc, bc).sum(dim=(1, 3))
        rep["bad_rows_first"] = [int(x) for x in torch.nonzero(pad.sum(1))[:16].flatten().tolist()]
        rep["bad_cols_first"] = [int(x) for x in torch.nonzero(pad.sum(0))[:16].flatten().tolist()]
        rep["bad_block_grid_shape"] = [rr, cc]
        rep["bad_blocks_first"] = [[int(a), int(b), int(grid[a, b])] for a, b in torch.nonzero(grid)[:24].tolist()]
        idx = torch.nonzero(d2 > tol.reshape(-1, diff.shape[-1]))[:6]
        g2, r2 = got.reshape(-1, diff.shape[-1]), ref.reshape(-1, diff.shape[-1])
        rep["samples"] = [[int(i), int(j), float(g2[i, j]), float(r2[i, j])] for i, j in idx.tolist()]
    try:
        os.makedirs(OUT_DIR, exist_ok=True)
        with open(os.path.join(OUT_DIR, "kernel_parity.jsonl"), "a") as f:
            f.write(json.dumps(rep) + "\n")
    except OSError:
        pass
    return rep


def assert_close(name, got, ref, atol, rtol):
    rep = err_report(name, got, ref, atol, rtol)
    assert rep["n_bad"] == 0 and not rep["has_nan"], json.dumps(rep)
