from .dice import MDiceLoss  # noqa: F401
