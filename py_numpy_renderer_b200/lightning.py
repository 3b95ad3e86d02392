"""Light kinds (reference: obj/lightning.py:4-7).  One class object is shared by every import path of this
package (in the reference, importing `Lightning` through two module paths yields two
unequal Enums and silently changes behaviour; `compat/lightning.py` re-exports *this* class)."""
from enum import Enum


class Lightning(Enum):
    DIRECTIONAL_LIGHTNING = 0
    POINT_LIGHTNING = 1
    SPOT_LIGHTNING = 2
