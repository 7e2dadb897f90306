"""Drop-in for the reference package ``emd`` (external/emd/emd/__init__.py:1-4)."""
from .emd_module import emdModule

__all__ = ["emdModule"]
