import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "tests")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


# Observed parity errors: every GPU parity test reports its measured distances through
# mvptr_parity_utils.report(); they are printed at the end of the session (also under -q) and written to
# gpurun_out/parity_report.txt so the numbers behind each assert can be read, not just "passed".
PARITY_LINES = []


def pytest_terminal_summary(terminalreporter):
    if not PARITY_LINES:
        return
    terminalreporter.write_sep("-", "observed parity errors (CUDA path vs oracle)")
    for line in PARITY_LINES:
        terminalreporter.write_line(line)
    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.txt"), "w") as f:
            f.write("\n".join(PARITY_LINES) + "\n")
    except OSError:
        pass
