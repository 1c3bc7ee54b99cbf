"""Test-only import shims so that the REFERENCE's own caller code (model/model.py `DDPM`, core/logger.py, sr.py) can be imported
in a minimal environment and driven against `ucdir_b200.model.networks.define_G` (SURVEY 8c "What does NOT import here").

Nothing here re-implements reference behaviour on the hot path: the shims replace modules that are *absent from the image*
(`omegaconf`, `tensorboardX`, `lmdb`, `lpips`, `torchvision.transforms.functional_tensor`, removed upstream) with the
thinnest stand-in the reference's import statements accept.  TEST INFRASTRUCTURE ONLY -- never imported by ucdir_b200/.

The reference tree itself is looked up at /root/reference (build container) or baseline/_ref/ucdir_reference (a git-ignored
copy made by `scripts/vendor_reference.py`, which travels to the GPU box)."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CANDIDATES = [os.environ.get("UCDIR_REFERENCE", ""), "/root/reference", os.path.join(ROOT, "baseline", "_ref", "ucdir_reference")]


def reference_root():
    for c in CANDIDATES:
        if c and os.path.exists(os.path.join(c, "model", "model.py")):
            return c
    return None


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Insert the stand-ins (idempotent) and put the reference tree on sys.path.  Returns the reference root or None."""
    ref = reference_root()
    if ref is None:
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    if "lpips" not in sys.modules:
        _module("lpips")                                   # model/diffusion.py:12; only PerceptualGaussianDiffusion uses it
    if "lmdb" not in sys.modules:
        _module("lmdb")                                    # data/LRHR_dataset.py:3; only the lmdb dataset type opens one
    try:
        import torchvision.transforms.functional_tensor   # noqa: F401  (removed in torchvision >= 0.17)
    except Exception:
        import torchvision.transforms.functional as TF
        _module("torchvision.transforms.functional_tensor", rgb_to_grayscale=TF.rgb_to_grayscale)
    if "tensorboardX" not in sys.modules:
        class SummaryWriter:                               # sr.py:360: scalars / images for the training dashboard
            def __init__(self, *a, **k): pass
            def __getattr__(self, name): return lambda *a, **k: None
        _module("tensorboardX", SummaryWriter=SummaryWriter)
    if "omegaconf" not in sys.modules:
        import yaml

        class OmegaConf:                                   # core/logger.py:35: OmegaConf.load(path) -> mapping; to_container
            @staticmethod
            def load(path):
                with open(path) as f:
                    return yaml.safe_load(f)

            @staticmethod
            def create(obj=None):
                return obj if obj is not None else {}

            @staticmethod
            def to_container(cfg, resolve=True):
                return cfg

            @staticmethod
            def to_yaml(cfg):
                return yaml.safe_dump(cfg)
        _module("omegaconf", OmegaConf=OmegaConf)
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            _module("wandb")
    return ref
