"""Boundary checks against the UNMODIFIED reference's own classes (build container only: needs /root/reference).

* f2: the native OBJ tokenizer and the Python parser produce the arrays of the reference's `Model.load_model`
  (obj/core.py:257-318) on cube.obj (+MTL, usemtl groups) and diablo3_pose.obj;
* INTEGRATION.md section 2 for real: the reference's `Model` / `Camera` / `Light` / `Scene` objects go through
  `_abi.PackedScene` / `pack_view` / `pack_frame_params` and must pack to the same bytes as this package's classes
  (obj/core.py:231-256, 373-429);
* a14 fence: a texture lookup below -size raises IndexError in the reference (core.py:138-143, 162-173), in the
  oracle, and (GPU test) in the CUDA path.
"""
import ctypes
import os

import numpy as np
import pytest

import py_numpy_renderer_b200 as b2r
import scenes
from py_numpy_renderer_b200 import _abi


def _ref():
    import refboot
    if not refboot.available():
        pytest.skip("the Python reference is only present in the build container")
    return refboot.boot(), refboot


def _same_arrays(mine, ref_model):
    assert mine.vertices.dtype == ref_model.vertices.dtype and np.array_equal(mine.vertices, ref_model.vertices)
    for x, y in ((mine.uv, ref_model.uv), (mine.normals, ref_model.normals)):
        assert (x is None) == (y is None) and (x is None or (x.dtype == y.dtype and np.array_equal(x, y)))
    assert np.array_equal(mine._faces, np.asarray(ref_model._faces))
    assert list(mine.material_group) == list(ref_model.material_group)
    assert sorted(mine.materials) == sorted(ref_model.materials)


@pytest.mark.parametrize("rel", ["obj_loader_test/cube.obj", "diablo3_pose/diablo3_pose.obj"])
@pytest.mark.parametrize("native", [True, False])
def test_loader_equals_reference_loader(rel, native):
    ref, refboot = _ref()
    path = os.path.join(refboot.ASSETS, rel)
    want = ref.Model.load_model(path)
    got = b2r.Model.load_model(path, native=native)
    _same_arrays(got, want)
    for name, mat in want.materials.items():                       # MTL scalars and which maps were found
        mine = got.materials[name]
        for key in ("Kd", "Ks", "Ns"):
            assert np.allclose(np.asarray(getattr(mine, key), float).ravel(), np.asarray(getattr(mat, key), float).ravel())
        for key in ("map_Kd", "map_Ks", "norm"):
            assert hasattr(mine, key) == hasattr(mat, key), (name, key)
            if hasattr(mat, key):                                  # same texels: reference stores f32(u8 / 255)
                assert np.array_equal(np.asarray(getattr(mat, key)), np.array(getattr(mine, key).texels / 255, np.float32))


def test_malformed_face_raises_value_error_like_the_reference(tmp_path):
    ref, _ = _ref()
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 x 3\n")
    with pytest.raises(ValueError):
        ref.Model.load_model(str(bad))
    for native in (True, False):
        with pytest.raises(ValueError):
            b2r.Model.load_model(str(bad), native=native)


def _struct_bytes(s):
    return bytes(ctypes.string_at(ctypes.addressof(s), ctypes.sizeof(s)))


def test_reference_objects_pack_to_the_same_abi_bytes():
    """The stub of INTEGRATION.md section 2 hands the reference's own objects to the packers: same bytes as ours."""
    ref, refboot = _ref()
    A = refboot.ASSETS

    def build(api, own):
        d = api.Model.load_model(os.path.join(A, "diablo3_pose/diablo3_pose.obj"))
        d = d @ api.scale(0.9) @ api.translation((0.1, 0.0, -0.2)) @ api.rotate_xyz((0, 25, 0))
        fl = api.Model(scenes.FLOOR_V.copy(), scenes.FLOOR_UV.copy(), scenes.FLOOR_N.copy(), scenes.FLOOR_F.copy(),
                       clip=False)
        cam = api.Camera((0.5, 1.5, 3), center=np.array((0, 0, 0)), fovy=60, near=0.1, far=10, backface_culling=True)
        dcam = api.Camera((0.5, 1.5, 3), center=np.array((0, 0, 0)), fovy=90, near=0.05, far=20, backface_culling=True)
        light = api.Light((2, 3, 3), light_type=api.Lightning.SPOT_LIGHTNING, center=(0.1, 0.2, 0), ambient_strength=0.1,
                          linear=0.01, quadratic=0.001)
        sc = api.Scene(cam, light, debug_camera=dcam, resolution=(270, 480), system=api.SYSTEM.LH,
                       subsystem=api.SUBSYSTEM.OPENGL, skymap=(0.1, 0.2, 0.3))
        sc.add_model(d)
        sc.add_model(fl)
        return sc

    mine, theirs = build(b2r, True), build(ref, False)
    pm, pt = _abi.PackedScene(mine.models, None), _abi.PackedScene(theirs.models, None)
    assert pm.n_models == pt.n_models and pm.total_faces == pt.total_faces
    for i in range(pm.n_models):
        a, b = pm.models[i], pt.models[i]
        for field in ("n_vertices", "n_uv", "n_normals", "n_faces", "n_materials", "vertex_dtype", "uv_dtype",
                      "normal_dtype", "clip", "depth_test"):
            assert getattr(a, field) == getattr(b, field), field
        nbytes = a.n_vertices * 4 * (4 if a.vertex_dtype == _abi.B2R_F32 else 8)
        assert ctypes.string_at(a.vertices, nbytes) == ctypes.string_at(b.vertices, nbytes)
        assert ctypes.string_at(a.faces, a.n_faces * 48) == ctypes.string_at(b.faces, b.n_faces * 48)
    va = _abi.pack_view(mine.camera, mine.debug_camera, mine.system, False)
    vb = _abi.pack_view(theirs.camera, theirs.debug_camera, theirs.system, False)
    assert _struct_bytes(va) == _struct_bytes(vb)
    bg = (_abi.B2R_BG_COLOR, tuple(np.array((0.1, 0.2, 0.3), np.float32).tolist()))
    fa = _abi.pack_frame_params(mine.light, mine.resolution, bg, True)
    fb = _abi.pack_frame_params(theirs.light, theirs.resolution, bg, True)
    assert _struct_bytes(fa) == _struct_bytes(fb)


def _uv_out_of_range_scene(api, own=True, texture_factory=None):
    """Floor whose u runs from -2 to 0: columns below -width for most of it."""
    uv = scenes.FLOOR_UV.copy()
    uv[:, 0] = uv[:, 0] * 2 - 2
    fl = api.Model(scenes.FLOOR_V.copy(), uv, scenes.FLOOR_N.copy(), scenes.FLOOR_F.copy())
    tex = dict(texels=scenes.procedural_texture(32, 7), signed=False, tangent=False)
    if own:
        from py_numpy_renderer_b200.materials import Texture
        fl.materials['default'].map_Kd = Texture(tex['texels'], signed=False, tangent=False)
    else:
        object.__setattr__(fl.materials['default'], 'map_Kd', texture_factory(tex))
    cam = api.Camera((0.5, 1.5, 3), center=np.array((0, 0, 0)), fovy=60, near=0.1, far=10, backface_culling=True)
    dcam = api.Camera((0.5, 1.5, 3), center=np.array((0, 0, 0)), fovy=90, near=0.05, far=20, backface_culling=True)
    light = api.Light((2, 3, 3), ambient_strength=0.1, linear=0.01, quadratic=0.001)
    sc = api.Scene(cam, light, debug_camera=dcam, resolution=(60, 80), system=api.SYSTEM.LH,
                   subsystem=api.SUBSYSTEM.OPENGL)
    sc.add_model(fl)
    return sc


def test_uv_below_minus_one_raises_index_error_in_reference_and_oracle(oracle):
    ref, _ = _ref()

    def ref_texture(t):
        return np.array(t['texels'] / 255, dtype=np.dtype(np.float32, metadata={'tangent': False}))
    import contextlib
    import io
    with pytest.raises(IndexError), contextlib.redirect_stdout(io.StringIO()):
        _uv_out_of_range_scene(ref, own=False, texture_factory=ref_texture).render()
    with pytest.raises(IndexError):
        oracle.render_scene(_uv_out_of_range_scene(b2r))


@pytest.mark.gpu
def test_uv_below_minus_one_raises_index_error_on_the_gpu():
    sc = _uv_out_of_range_scene(b2r)
    sc.verbose = False
    with pytest.raises(IndexError):
        sc.render()
    with pytest.raises(IndexError):
        sc.render_batch([sc.camera])
    # asynchronous host render: reported by result(); device-resident render: reported by the next sync
    from py_numpy_renderer_b200 import _native
    with pytest.raises(IndexError):
        sc.render_batch_async([sc.camera]).result()
    import torch
    out = torch.empty((1, 60, 80, 3), dtype=torch.uint8, device="cuda:0")
    sc.render_batch([sc.camera], out=out)
    with pytest.raises(IndexError):
        _native.sync()
    _native.sync()                                   # the flag is cleared once reported
    # a scene that stays inside its maps is unaffected
    ok = scenes.c3_synthetic((60, 80), tex=64, nu=16, nv=8)
    ok.render()


@pytest.mark.parametrize("name", ["g7_cube_mtl_rh_directx", "g13_depth_test_false_rh"])
def test_product_scene_converted_to_reference_objects_reproduces_the_fixture(name):
    """oracle/refboot.to_reference_scene (what bench.py's NumPy-reference arm renders) is the same scene: the unmodified
    reference renders the committed fixture bit for bit from the product-side Scene description."""
    _, refboot = _ref()
    import contextlib
    import io
    import golden_util as gu
    scene, exp, _ = gu.load(name)
    with contextlib.redirect_stdout(io.StringIO()):
        rgb = refboot.to_reference_scene(scene).render()
    assert np.array_equal(rgb, exp['rgb'])
