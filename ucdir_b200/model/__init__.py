"""Mirror of the reference's `model` package surface for the hot path (networks / ucdir / diffusion)."""
from . import networks  # noqa: F401
