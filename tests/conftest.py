import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    # a fresh checkout has no built artefacts (they are git-ignored): build once, nvcc cross-compiles without a GPU
    lib = os.path.join(ROOT, "whisper-rust-ort_b200", "libwhisper_b200.so")
    cli = os.path.join(ROOT, "whisper-rust-ort_b200", "whisper_b200_cli")
    if not (os.path.exists(lib) and os.path.exists(cli)):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def wb():
    import wb200
    return wb200


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
