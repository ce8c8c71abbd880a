import os
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line(
        "markers", "first_hw_run: a GPU test written after the round's GPU budget was spent -- green on the SIMT emulator "
        "(tests/test_simt_*.py), never yet run on a B200.  On a CUDA box it runs LAST and as a non-strict xfail, so a "
        "surprise on its first hardware run is reported (xfail / XPASS) without costing the tests already seen green "
        "under `pytest -x`.  Drop the marker once the test has been seen on hardware.")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        first = [it for it in items if "first_hw_run" in it.keywords]
        if first:
            seen = [it for it in items if "first_hw_run" not in it.keywords]
            for it in first:
                it.add_marker(pytest.mark.xfail(strict=False, reason="first hardware run of this test"))
            items[:] = seen + first
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(GOLDEN / f"{name}.pt", weights_only=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """One line per `first_hw_run` test that did not pass on its first hardware run (they are non-strict xfails, so the
    run itself stays green): which one and the last line of its failure, for whoever reads the log of the GPU tier."""
    xf = terminalreporter.stats.get("xfailed", [])
    xp = terminalreporter.stats.get("xpassed", [])
    if not (xf or xp):
        return
    terminalreporter.write_line(f"first_hw_run: {len(xp)} test(s) green on their first hardware run, {len(xf)} not")
    for rep in xf:
        text = getattr(rep, "longreprtext", "") or ""
        lines = [ln.strip() for ln in text.splitlines() if ln.strip()]
        last = next((ln[1:].strip() for ln in lines if ln.startswith("E ")), lines[-1] if lines else "")
        terminalreporter.write_line(f"  first_hw_run NOT GREEN: {rep.nodeid}: {last[:300]}")
