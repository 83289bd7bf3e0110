"""pytest configuration: registers the `gpu` marker and makes the repo importable.

`-m "not gpu"`: oracle vs golden vectors / reference KATs, host logic, C-ABI symbol export.
`-m gpu`      : the parity tests proper -- CUDA path (through the C-ABI) vs the oracle.
"""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "aztec-2.0_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libbbref.so (the compiled reference)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    from oracle import pyoracle
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle import pyoracle
    if not pyoracle.Ref.available():
        pytest.skip("oracle/_ref/libbbref.so not built (reference tree absent)")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def srs_mini(orc):
    """(points[4096], table[2*4096+slack]) decoded from tests/golden/srs_mini by the oracle's loader."""
    import inputs
    pts = orc.read_transcript_g1(inputs.SRS_MINI_POINTS, inputs.SRS_MINI_DIR)
    return pts, orc.point_table(pts)
