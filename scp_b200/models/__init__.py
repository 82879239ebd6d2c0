"""Module interfaces of the reference's ``models`` package (models/__init__.py:1-6)."""
from .ehem import EHEM
from .oct_attention import OctAttention

__all__ = ["EHEM", "OctAttention"]
