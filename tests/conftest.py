import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    """Golden vectors produced by the unmodified reference on CPU (oracle/gen_golden.py)."""
    import json

    import numpy as np

    gdir = ROOT / "tests" / "golden"
    meta = json.loads((gdir / "w8a8_golden.json").read_text())
    arrays = np.load(gdir / "w8a8_golden.npz")
    cases = []
    for entry in meta["cases"]:
        case = dict(entry)
        prefix = entry["id"] + "."
        for key in arrays.files:
            if key.startswith(prefix):
                case[key[len(prefix):]] = arrays[key]
        cases.append(case)
    return cases
