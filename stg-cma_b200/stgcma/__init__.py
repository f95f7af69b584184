"""stgcma -- host side of the B200-native STG-CMA adapter train step (kernels: ../csrc)."""
from . import lib  # noqa: F401
from .config import SwinAVEConfig, PRESETS  # noqa: F401
