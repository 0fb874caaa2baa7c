import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_bigvgan():
    return dict(np.load(os.path.join(GOLDEN, "bigvgan_ref.npz")))


@pytest.fixture(scope="session")
def engine():
    """One libb200tts engine on cuda:0 for the whole GPU test session (fails loudly without a GPU/.so)."""
    import b200tts  # noqa: F401
    from b200tts import capi
    return capi.Engine(0)


@pytest.fixture(scope="session")
def bigvgan_engine(engine):
    import b200tts  # noqa: F401
    from b200tts import synth, weights
    engine.load_state("bigvgan", weights.bigvgan_engine_tensors(synth.bigvgan_state(1234)))
    engine.bigvgan_build()
    return engine


def snr_db(ref, x):
    ref = np.asarray(ref, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    return 10.0 * np.log10((ref ** 2).sum() / max(((ref - x) ** 2).sum(), 1e-30))
