"""Test stub: matplotlib is not installed in this image."""
from . import cm  # noqa: F401
