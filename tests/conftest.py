import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CC_FIXED = dict(type="OSC_POSE", input_max=1, input_min=-1, output_max=[0.05] * 3 + [0.5] * 3, output_min=[-0.05] * 3 + [-0.5] * 3,
                kp=300, damping_ratio=1, impedance_mode="fixed", kp_limits=[0, 500], kp_input_max=1, kp_input_min=0,
                damping_ratio_limits=[0, 2], uncouple_pos_ori=True, control_delta=True)  # src/main.py:25-44
CC_TRACK = dict(CC_FIXED, impedance_mode="tracking")  # src/rl_config.yaml:33-51


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "task_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def art():
    with open(os.path.join(ROOT, "tests", "golden", "art_stats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def soft_model():
    from rui_b200.env import packed_model
    return packed_model(True)


@pytest.fixture(scope="session")
def rigid_model():
    from rui_b200.env import packed_model
    return packed_model(False)


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    oracle.lib()
    return oracle
