"""Stand-in for torch_geometric 2.5.2 (absent in this image): only the symbols the
reference hot path touches.  TEST INFRASTRUCTURE ONLY."""
from . import nn, data, transforms, utils, typing  # noqa: F401
