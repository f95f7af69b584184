"""`model` package with the reference's module paths (the reference forgot its __init__.py, so
`import model; model.Swin_AVE` raises AttributeError there -- run_adapt_ave29.py:12,156)."""
from . import Swin_AVE  # noqa: F401
from . import CLIP_AVE  # noqa: F401
