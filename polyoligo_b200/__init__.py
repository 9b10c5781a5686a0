"""polyoligo_b200 -- B200-native batch thermodynamic scoring for PolyOligo's hot path.

See DESIGN.md for the scope and include/polyoligo_b200.h
for the C ABI underneath this package.
"""
from ._native import (Context, pack, unpack, THAL_ANY, THAL_END1, THAL_END2, THAL_HAIRPIN, ITEM_STRUCTURE_FOUND, ITEM_SKIPPED,  # noqa: F401
                      PRIMER_VALID, ValidParams, OLIGO_DTYPE, THAL_OUT_DTYPE, PROPS_DTYPE)

__version__ = "0.1.0"
