import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# measured parity errors of the GPU tests: {test id: {label: {"err": ..., "tol": ...}}}, dumped at session end to
# gpurun_out/parity_errors.json (copied to profiles/r2_parity_errors.json by the builder)
_PARITY = {}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.lib()
    return orc


@pytest.fixture
def perr(request):
    """perr(label, err, tol): log the measured error of one comparison, then assert err <= tol"""
    def record(label, err, tol):
        err, tol = float(err), float(tol)
        _PARITY.setdefault(request.node.nodeid, {})[label] = {"err": err, "tol": tol}
        assert err <= tol, (request.node.nodeid, label, err, tol)
    return record


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY:
        return
    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        worst = {}
        for tid, d in _PARITY.items():
            for label, e in d.items():
                key = label.split("@")[0]
                if key not in worst or e["err"] / e["tol"] > worst[key]["err"] / worst[key]["tol"]:
                    worst[key] = dict(e, test=tid)
        with open(os.path.join(out, "parity_errors.json"), "w") as f:
            json.dump({"tolerance_north_star": 1e-12, "worst_by_quantity": worst, "by_test": _PARITY}, f, indent=1, sort_keys=True)
    except OSError:
        pass
