import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu on the GPU box")


def load_cases(name):
    """Golden .npz -> list of per-case dicts (keys stripped of their c<i>_ prefix)."""
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    n = int(z["num_cases"])
    cases = []
    for i in range(n):
        pre = f"c{i}_"
        cases.append({k[len(pre):]: z[k] for k in z.files if k.startswith(pre)})
    return cases


def nerr(a, b):
    """Normalised max error max|a-b| / max|b| (outputs can be near zero; SURVEY §7 'Numerics')."""
    import torch
    a = a.detach().double().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a)).double()
    b = b.detach().double().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b)).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


@pytest.fixture(scope="session")
def R():
    import robot_gnn_b200
    return robot_gnn_b200
