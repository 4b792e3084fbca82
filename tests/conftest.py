import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


# ---- measured parity numbers ---------------------------------------------------------------------------------------
# Every floating-point comparison against the reference's goldens records what it measured next to its bar; the table is
# printed in the terminal summary (so the GPU test log carries the margins, not just "passed") and written to
# gpurun_out/parity_report.json on the GPU box.
_PARITY = []


@pytest.fixture
def parity(request):
    def record(what, measured, bar, unit="rel-L2"):
        _PARITY.append({"test": request.node.nodeid, "what": what, "measured": float(measured), "bar": float(bar),
                        "unit": unit})
        return float(measured)
    return record


def pytest_terminal_summary(terminalreporter):
    if not _PARITY:
        return
    terminalreporter.section("measured parity vs the reference (measured / bar)")
    for r in _PARITY:
        terminalreporter.write_line(f"{r['measured']:.3e} / {r['bar']:.1e} {r['unit']:8s} {r['test']} :: {r['what']}")
    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        import json

        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "parity_report.json"), "w") as f:
            json.dump(_PARITY, f, indent=1)
    except OSError:
        pass
