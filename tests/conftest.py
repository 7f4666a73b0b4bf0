import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle_built():
    from oracle import loader
    loader.build(ref=os.path.isdir("/root/reference/nGramSearch"))
    return loader


SEVEN = [b"LWMS", b"LWM", b"LWMA", b"LWYY", b"L", b"I", b"GHRSDGSDGS Egdsrtg g"]   # SearchTest/test.cpp:7-9
