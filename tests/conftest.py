import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def split_golden(z):
    """-> (weights, inputs, outputs, rest) as torch tensors, keyed without the prefix."""
    def pick(p):
        return {k[len(p):]: torch.from_numpy(np.ascontiguousarray(v)) for k, v in z.items() if k.startswith(p)}
    rest = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in z.items()
            if not k.startswith(("w.", "in.", "out."))}
    return pick("w."), pick("in."), pick("out."), rest


@pytest.fixture(scope="session")
def has_cuda():
    return torch.cuda.is_available()
