"""Minimal torch_geometric surface for the reference's callers (see tests/shims/README.md)."""
from . import nn, typing, utils  # noqa: F401
