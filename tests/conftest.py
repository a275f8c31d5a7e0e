import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (authoring container only)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def oracle_sd():
    from oracle import ladiff_oracle as O
    return O.make_state_dict(1234, 263, perturb=True)


@pytest.fixture(scope="session")
def engine(oracle_sd):
    """CUDA engine with the synthetic weights of the golden fixtures loaded (gpu tests only)."""
    import torch
    from ladiff_b200._lib import Engine
    from oracle import ladiff_oracle as O
    eng = Engine(nfeats=263)
    eng.set_weights({k: v.cuda() for k, v in O.sub(oracle_sd, "denoiser.").items()}, "denoiser.")
    eng.set_weights({k: v.cuda() for k, v in O.sub(oracle_sd, "vae.").items()}, "vae.")
    eng.finalize(3)
    return eng
