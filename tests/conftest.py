import json
import os
import sys
import tarfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _extract(name, dest):
    with tarfile.open(os.path.join(GOLDEN, name + ".tar.gz")) as tf:
        tf.extractall(dest, filter="data")
    return dest


@pytest.fixture
def golden(tmp_path):
    """golden("syn_pair") -> (directory with inputs + reference outputs, cases list)."""

    def get(name):
        d = tmp_path / name
        d.mkdir()
        _extract(name, str(d))
        cj = d / "cases.json"
        cases = json.load(open(cj)) if cj.exists() else []
        return str(d), cases

    return get


def read(path):
    with open(path, "rb") as fh:
        return fh.read()
