import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
# The library's default precision is "fp32_tc" (tensor cores, fp32 tolerance).  Tests that do not pick a mode themselves run the
# SIMT "fp32" graph -- the independent arithmetic the tensor-core modes are also compared with; the tensor-core modes are selected
# explicitly (set_precision / parametrised fixtures).
os.environ.setdefault("UCDIR_PRECISION", "fp32")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def sid_weights():
    """Seeded full-size sid weights (seed 1234) drawn through ucdir_b200's mirror constructors; the
    state_dict is bit-identical to the reference's (guarded by the golden sha256)."""
    import hashlib
    import numpy as np
    import torch
    import ucdir_b200
    from ucdir_b200.model.networks import define_G
    torch.manual_seed(1234)
    net = define_G({"model": ucdir_b200.SID_MODEL_OPT})
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.cpu().contiguous().numpy().tobytes())
    want = str(np.load(os.path.join(GOLDEN, "unet.npz"))["digest"])
    assert h.hexdigest() == want, "seeded weights differ from the reference's (construction order / torch RNG changed)"
    return net, sd
