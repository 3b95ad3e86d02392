"""Cubemap skybox container (reference: obj/cube_map.py:7-80).

Holds the six faces as uint8 (6, S, S, 3) in the reference's slot order right, left, top, bottom, front, back
after the same flips / rotations / transposes its constructor applies (cube_map.py:24-44); the lookup itself
(`CubeMap.__getitem__`, cube_map.py:63-80) and the full-screen fill (`fill_frame_from_skybox`, cube_map.py:83-101)
run on the device (csrc: skybox stage of the shade kernel)."""
import numpy as np
from PIL import Image


class CubeMap:
    def __init__(self, left, right, top, bottom, front, back, normalize_input=True):
        tex = self.load_texels
        if normalize_input:
            faces = [np.flip(tex(right), axis=[0, 1]),
                     np.rot90(tex(left).transpose((1, 0, 2)), -1),
                     tex(top).transpose((1, 0, 2)),
                     np.rot90(tex(bottom)),
                     np.rot90(tex(front), -1),
                     tex(back).transpose((1, 0, 2))]
        else:
            faces = [tex(right), tex(left), tex(top), tex(bottom), tex(front), tex(back)]
        self.texels = np.ascontiguousarray(np.array(faces), dtype=np.uint8)
        assert self.texels.ndim == 4 and self.texels.shape[1] == self.texels.shape[2], "cubemap faces must be square"

    @staticmethod
    def load_texels(name) -> np.ndarray:
        """uint8 RGB (alpha dropped) -- the reference divides by 255 here (cube_map.py:56-61)."""
        return np.asarray(Image.open(name))[..., :3].copy()

    @staticmethod
    def load_texture(name) -> np.ndarray:
        """The reference's loader name and value: float64 RGB in [0, 1] (cube_map.py:56-61)."""
        return CubeMap.load_texels(name) / 255

    @property
    def textures(self) -> np.ndarray:
        """The reference's float64 (6,S,S,3) view, for host-side inspection / tests."""
        return self.texels / 255

    def __getitem__(self, vectors):
        """Host restatement of the direction -> texel lookup (cube_map.py:63-80); the product path is on the
        device, this exists so code that indexed a CubeMap directly keeps working."""
        vectors = np.asarray(vectors, dtype=np.float64)
        rows = np.arange(vectors.shape[0])
        axis = np.abs(vectors).argmax(axis=1)
        amp = vectors[rows, axis, None]
        keep = np.ones(vectors.shape, bool)
        keep[rows, axis] = False
        st = (vectors[keep].reshape(vectors.shape[0], -1) / amp + 1) / 2
        side = (amp < 0).ravel() + axis * 2
        ij = (st.T * self.texels.shape[1] - 1).astype(int)
        return self.textures[side.astype(int), ij[0], ij[1]]
