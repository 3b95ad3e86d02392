"""MTL attribute bag (reference: obj/materials.py:4-77).

Data contract kept: class-level defaults `Kd=.8`, `Ks=1`, `Ns=64` (materials.py:47-55); assignment coerces a
1-element sequence to `float` (or keeps the string) and longer sequences to float32 arrays (materials.py:57-64).
Texture maps (`map_Kd`, `map_Ks`, `norm`) are attached as attributes by `TextureMaps.register`; in this package
they are `Texture` objects that keep the decoded uint8 texels (the kernels rebuild the reference's float32 texel
`u8/255` or `u8/255*2-1` exactly from a 256-entry table).
"""
import numpy as np


class Texture:
    """uint8 RGB texels + how the reference would have turned them into float32 (core.py:90-105)."""
    __slots__ = ("texels", "signed", "tangent")

    def __init__(self, texels: np.ndarray, signed: bool, tangent: bool):
        texels = np.ascontiguousarray(texels, dtype=np.uint8)
        assert texels.ndim == 3 and texels.shape[2] == 3
        self.texels = texels
        self.signed = bool(signed)      # True: u8/255*2-1 (core.py:96-97), False: u8/255
        self.tangent = bool(tangent)    # dtype metadata 'tangent' in the reference (core.py:94)

    @property
    def shape(self):
        return self.texels.shape

    def as_float32(self) -> np.ndarray:
        """The array the reference stores (for tests / host-side inspection)."""
        t = self.texels / 255
        if self.signed:
            t = t * 2 - 1
        return np.array(t, dtype=np.float32)


class Material:
    Pm = 0.5
    Pr = 0.5
    Ka = np.array((0.3, 0, 0))
    Kd = np.array((0.8, 0.8, 0.8))
    Ks = np.array((1., 1., 1.))
    d = 1.0
    Tr = 0
    Ns = 64
    illum = 1

    def __setattr__(self, key, value):
        if isinstance(value, Texture):
            object.__setattr__(self, key, value)
        elif len(value) == 1:
            try:
                object.__setattr__(self, key, float(value[0]))
            except ValueError:
                object.__setattr__(self, key, value[0])
        else:
            object.__setattr__(self, key, np.array(value, dtype=np.float32))
