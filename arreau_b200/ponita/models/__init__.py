from .ponita import PonitaFiberBundle  # noqa: F401
