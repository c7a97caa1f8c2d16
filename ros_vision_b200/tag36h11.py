"""tag36h11 family data: a view of the generated table (tools/gen_families.py)."""
from .tag_families import FAMILIES

_F = FAMILIES["tag36h11"]
BIT_X, BIT_Y, CODES = _F["bit_x"], _F["bit_y"], _F["codes"]
NBITS, WIDTH_AT_BORDER, TOTAL_WIDTH, MIN_HAMMING = _F["nbits"], _F["width_at_border"], _F["total_width"], _F["min_hamming"]
