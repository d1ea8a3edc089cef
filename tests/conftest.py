import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
        return cache[name]

    return load


# PARESIS_REPORT_L2=<file>: resolved here, against the directory pytest was started from -- the workspace fixtures chdir
_REPORT_L2 = os.path.abspath(os.environ["PARESIS_REPORT_L2"]) if os.environ.get("PARESIS_REPORT_L2") else None


def rel_l2(a, b):
    import numpy as np

    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b.ravel())
    val = float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))
    if _REPORT_L2:       # parity report: every measured distance, with its call site
        fr = sys._getframe(1)
        with open(_REPORT_L2, "a") as fh:
            fh.write("%s:%d %s %.3e\n" % (os.path.basename(fr.f_code.co_filename), fr.f_lineno, fr.f_code.co_name, val))
    return val
