from . import linalg, stats  # noqa: F401
