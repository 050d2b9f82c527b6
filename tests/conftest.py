import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_ERRORS = {}          # test id -> list of {what, rel_err, rel_err_elem, tol, ...}: written to gpurun_out/ at session end


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny): the relative-error metric every parity test states.  Every evaluation made inside a
    GPU test is recorded with its call site (gpurun_out/gpu_test_errors.json)."""
    e = _rel(a, b)
    test = os.environ.get("PYTEST_CURRENT_TEST", "")
    if test and (getattr(a, "is_cuda", False) or getattr(b, "is_cuda", False)):
        f = sys._getframe(1)
        _ERRORS.setdefault(test.split(" ")[0], []).append(
            {"at": f"{os.path.basename(f.f_code.co_filename)}:{f.f_lineno}", "rel_err": e})
    return e


def rel_err_elem(a, b, floor=1e-2):
    """element-wise companion of rel_err: max_i |a_i - b_i| / (|b_i| + floor * max|b|).  Reported next to rel_err so that a
    small element hiding behind a large one shows up; the pass / fail bars are stated on rel_err."""
    import torch
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0
    return float(((a - b).abs() / (b.abs() + floor * b.abs().max().clamp_min(1e-30))).max())


def record_err(what, got, ref, tol, **extra):
    """rel_err(got, ref), recorded under the running test's id (achieved errors of the GPU suite are committed under
    profiles/ every round)."""
    e = _rel(got, ref)
    test = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
    row = {"what": what, "rel_err": e, "rel_err_elem": rel_err_elem(got, ref), "tol": tol}
    row.update(extra)
    _ERRORS.setdefault(test, []).append(row)
    return e


def pytest_sessionfinish(session, exitstatus):
    if not _ERRORS:
        return
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "gpu_test_errors.json")
        merged = {}
        if os.path.exists(path):                     # several pytest sessions of one GPU job (suite, sanitizer passes) add up
            try:
                merged = json.load(open(path))
            except ValueError:
                merged = {}
        merged.update(_ERRORS)
        with open(path, "w") as f:
            json.dump(merged, f, indent=1, sort_keys=True)
    except OSError:
        pass
