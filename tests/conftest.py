import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def pytest_terminal_summary(terminalreporter):
    """Parity at a glance (also with -q): how many tensors were compared through tests/_util.check_close, the worst error
    against the reference's fp32 run, and the worst ratio of (error vs the fp64 run) / (the reference's own fp32 floor)."""
    try:
        from tests._util import PARITY_LOG, RTOL
    except Exception:
        return
    if not PARITY_LOG:
        return
    n = len(PARITY_LOG)
    within = sum(1 for _, e32, _, _ in PARITY_LOG if e32 <= RTOL)
    well = [r for r in PARITY_LOG if r[3] <= 1e-4] or PARITY_LOG     # the reference's own fp32 run is within 1e-4 of its fp64 run
    w32 = max(well, key=lambda r: r[1])
    stress = n - len(well)
    ratio = lambda r: r[2] / (r[3] + 8e-7)
    wr = max((r for r in PARITY_LOG if r[1] > RTOL), key=ratio, default=None)
    tr = terminalreporter
    tr.write_line(f"parity: {n} tensors compared with the reference; {within} within {RTOL:g} relative of its fp32 run; "
                  f"worst vs fp32 run where the reference's own fp32-vs-fp64 error is <= 1e-4: {w32[0]} e32={w32[1]:.1e} "
                  f"e64={w32[2]:.1e} reference-floor={w32[3]:.1e}; {stress} tensors belong to ill-conditioned stress cases "
                  f"(floor > 1e-4: jitter ladder, fractional masks)")
    if wr is not None:
        tr.write_line(f"parity: of the {n - within} above {RTOL:g}, worst (error vs fp64 run) / (reference's own fp32-vs-fp64 error): "
                      f"{wr[0]} e64={wr[2]:.1e} floor={wr[3]:.1e} ratio={wr[2] / max(wr[3], 1e-30):.2f} (allowed 2.5 + 2e-6)")
