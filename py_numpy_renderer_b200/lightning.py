"""Light kinds (reference: obj/lightning.py:4-7).  One class object is shared by every import path of this
package (in the reference, importing `Lightning` through two module paths yields two
unequal Enums and silently changes behaviour; `compat/lightning.py` re-exports *this* class).

The member values double as the light-type codes of the C ABI (`B2R_LIGHT_*` in include/b2r.h), so
`_abi.pack_light` passes `light_type.value` straight through."""
import enum

# name in the reference's API -> B2R_LIGHT_* code
_KINDS = {"DIRECTIONAL_LIGHTNING": 0, "POINT_LIGHTNING": 1, "SPOT_LIGHTNING": 2}

Lightning = enum.Enum("Lightning", _KINDS, module=__name__)
