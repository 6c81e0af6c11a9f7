"""Test-only stub of the mmcv-full 1.x symbols the reference imports (see ../README.md)."""
__version__ = "1.6.2-stub"
from . import utils, cnn, runner, ops  # noqa: F401
