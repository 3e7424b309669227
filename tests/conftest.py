"""pytest configuration: markers, import paths and shared fixtures.

`-m "not gpu"` : CPU suite (oracle vs golden vectors, host logic, C-ABI symbol checks).
`-m gpu`       : parity tests proper; they drive libjpegb200.so through its C ABI on cuda:0.
"""
import hashlib
import importlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def frames():
    return importlib.import_module("jpeg-encoder-decoder_b200.frames")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    import cpu_checkers

    return cpu_checkers.Oracle()


@pytest.fixture(scope="session")
def ref():
    import cpu_checkers

    if not cpu_checkers.Ref.available():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference)")
    return cpu_checkers.Ref()


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()
