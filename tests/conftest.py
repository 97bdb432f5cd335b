import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cfg():
    import yaml
    with open(os.path.join(ROOT, "configs", "prior.yaml")) as f:
        prior = yaml.safe_load(f)
    with open(os.path.join(ROOT, "configs", "prob.yaml")) as f:
        prob = yaml.safe_load(f)
    return {"prior_generator": prior, "prob_generator": prob}


@pytest.fixture(scope="session")
def flamed_sd(cfg):
    from oracle import weights as W
    return W.make_flamed_state_dict(cfg["prior_generator"], cfg["prob_generator"], 0)


@pytest.fixture(scope="session")
def codec_dec_sd():
    from oracle import weights as W
    return W.make_codec_decoder_state_dict(0)


@pytest.fixture(scope="session")
def codec_enc_sd():
    from oracle import weights as W
    return W.make_codec_encoder_state_dict(0)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
