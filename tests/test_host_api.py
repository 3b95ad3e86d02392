"""Host-side mirror of the reference API: loader, transform chain, matrix builders, error behaviour and the
documented quirks.  No GPU needed."""
import os

import numpy as np
import pytest

import py_numpy_renderer_b200 as b2r
from py_numpy_renderer_b200 import _abi, transformation as T
from py_numpy_renderer_b200.materials import Material, Texture

OBJ = """mtllib t.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0.5
vt 0 0
vt 1 0 0.25
vt 1 1
vn 0 0 1
usemtl red
f 1/1/1 2/2/1 3/3/1 4/1/1
usemtl blue
f 1/1/1 3/3/1 4//1
"""
MTL = """# comment
newmtl red
Ns 32.0
Kd 0.5 0.25 0.125
Ks 1 1 1
illum 2
newmtl blue
Kd 0 0 1
"""


@pytest.fixture
def obj_file(tmp_path):
    (tmp_path / "t.obj").write_text(OBJ)
    (tmp_path / "t.mtl").write_text(MTL)
    return str(tmp_path / "t.obj")


def test_loader_arrays(obj_file):
    m = b2r.Model.load_model(obj_file)
    assert m.vertices.dtype == np.float32 and m.vertices.shape == (4, 4) and (m.vertices[:, 3] == 1).all()
    assert m.uv.dtype == np.float32 and m.uv.shape == (3, 3) and m.uv[1, 2] == 0.25 and m.uv[0, 2] == 0
    assert m.normals.shape == (1, 3)
    assert m._faces.dtype == np.int32 and m._faces.shape == (3, 3, 4)       # quad fan-triangulated + 1 triangle
    assert m._faces[0, :, 0].tolist() == [0, 1, 2] and m._faces[1, :, 0].tolist() == [0, 2, 3]
    assert m._faces[2, 2].tolist() == [3, -1, 0, 2]                          # missing vt stays -1; slot 2 = 'blue'
    assert m.material_group == ['default', 'red', 'blue']
    assert m.materials['red'].Ns == 32.0 and m.materials['red'].Kd.dtype == np.float32
    assert m.materials['red'].illum == 2.0
    packed = _abi.PackedScene([m])
    mats = packed.models[0].materials
    assert [mats[1].Kd[k] for k in range(3)] == [0.5, 0.25, 0.125] and mats[1].Ns == 32.0
    assert [mats[0].Kd[k] for k in range(3)] == [0.8, 0.8, 0.8] and mats[0].Ns == 64.0   # materials.py defaults


def test_transform_chain_dtypes_and_values(obj_file):
    m = b2r.Model.load_model(obj_file)
    before = m.vertices.copy()
    assert b2r.scale(2).dtype == np.int64 and b2r.scale(0.5).dtype == np.float64
    assert b2r.rotate_xyz((10, 20, 30)).dtype == np.float32 and b2r.rotate is b2r.rotate_xyz
    out = m @ b2r.scale(2) @ b2r.translation((1, 2, 3))
    assert out is m and m.vertices.dtype == np.float64                      # in place, escalates to float64
    assert np.array_equal(m.vertices, before.astype(np.float64) * [2, 2, 2, 1] + [1, 2, 3, 0])
    # row-vector convention: the offset sits in row 3
    assert b2r.translation((1, 2, 3))[3].tolist() == [1, 2, 3, 1]
    # rotate_xyz: the matrix built from angle y rotates about x (naming swap of transformation.py:230-251)
    r = b2r.rotate_xyz((0, 90, 0))
    assert np.allclose(np.array([0, 1, 0, 1]) @ r, [0, 0, 1, 1], atol=1e-6)


def test_material_coercion_and_texture_registration(tmp_path):
    mat = Material()
    mat.Ns = ['12.5']
    mat.name = ['shiny']
    mat.Kd = ['0.1', '0.2', '0.3']
    assert mat.Ns == 12.5 and mat.name == 'shiny' and mat.Kd.dtype == np.float32
    from PIL import Image
    img = np.random.default_rng(0).integers(0, 255, (8, 4, 3), dtype=np.uint8)
    path = str(tmp_path / "t.png")
    Image.fromarray(img).save(path)
    m = b2r.Model(np.zeros((3, 4), np.float32), np.zeros((3, 3), np.float32), None, np.zeros((1, 3, 4), np.int32))
    m.textures.register('normals', path, tangent=True)
    m.textures.register('diffuse', path, normalize=False)
    nm, kd = m.materials['default'].norm, m.materials['default'].map_Kd
    assert isinstance(nm, Texture) and nm.signed and nm.tangent and not kd.signed
    assert np.array_equal(kd.as_float32(), np.array(img / 255, dtype=np.float32))            # core.py:104
    assert np.array_equal(nm.as_float32(), np.array(img / 255 * 2 - 1, dtype=np.float32))    # core.py:97
    with pytest.raises(ValueError):
        m.textures.register('glow', path)


def test_scene_quirks():
    cam = b2r.Camera((0, 0, 3), center=np.array((0, 0, 0)))
    with pytest.raises(AttributeError):                                      # Appendix B-1
        b2r.Scene(cam, b2r.Light((1, 1, 1)))
    dcam = b2r.Camera((0, 0, 3), center=np.array((0, 0, 0)))
    sc = b2r.Scene(cam, b2r.Light((1, 1, 1)), debug_camera=dcam, resolution=(100, 200), system=b2r.SYSTEM.RH,
                   subsystem=b2r.SUBSYSTEM.DIRECTX)
    mvp = cam.MVP
    cam.position = np.array((5, 5, 5))
    assert cam.MVP is mvp                                                    # cached_property: moving has no effect
    ortho = b2r.Camera((0, 0, 3), center=np.array((0, 0, 0)), projection_type=b2r.PROJECTION_TYPE.ORTHOGRAPHIC)
    assert ortho.near == 3.0                                                 # near := |position| (core.py:387)
    sc2 = b2r.Scene(ortho, b2r.Light((1, 1, 1)), debug_camera=dcam, resolution=(10, 10), system=b2r.SYSTEM.RH,
                    subsystem=b2r.SUBSYSTEM.OPENGL)
    with pytest.raises(KeyError):                                            # only OPENGL/LH has an ortho builder
        ortho.projection
    assert sc.resolution == (100, 200) and sc2.models == []


def test_matrix_builders_against_closed_forms():
    vp = T.ViewPort((1080, 1920), 10, 0.1, x_offset=3, y_offset=-2)
    assert vp[0, 0] == 960 and vp[1, 1] == 540 and vp[3].tolist() == [963, 538, (10 - 0.1) / 2, 1]
    p = T.opengl_perspectiveLH(60, 16 / 9, 0.1, 10)
    f = 1.0 / np.tan(np.radians(60) / 2.0)
    assert p[0, 0] == f / (16 / 9) and p[1, 1] == f and p[2, 3] == 1.0 and p[3, 2] == 2.0 * 10 * 0.1 / (10 - 0.1)
    assert T.opengl_orthographicLH(60, 1.5, 1, 5).dtype == np.float32
    planes = T.extract_frustum_planes(np.eye(4))
    assert np.allclose(np.linalg.norm(planes, axis=1), 1)
    assert np.array_equal(T.normalize(np.zeros(3)), np.zeros((1, 3)).squeeze()[None] * 0)
    assert T.bound_box(np.array([[-5., -5.], [-1., -2.], [-3., -4.]]), 10, 10) is None
    assert T.bound_box(np.array([[0.2, 0.5], [3.1, 9.9], [12., 4.]]), 10, 8).tolist() == [1, 8, 1, 10]
    assert T.barycentric(np.zeros(2), np.array([1., 1.]), np.array([2., 2.]), np.array([[0, 0]])) is None


def test_light_defaults_and_packing():
    light = b2r.Light((2, 3, 3), ambient_strength=0.1)
    assert light.light_type is b2r.Lightning.POINT_LIGHTNING and light.linear == 0.14 and light.quadratic == 0.07
    ld = _abi.pack_light(light)
    assert ld.type == 1 and [ld.ambient[k] for k in range(3)] == [0.1, 0.1, 0.1]
    assert np.allclose([ld.direction[k] for k in range(3)], np.array([2, 3, 3]) / np.sqrt(22))
    assert ld.spot_cos_outer == np.cos(np.deg2rad(20)) and ld.spot_cos_inner == np.cos(np.deg2rad(10))


def test_depth_test_false_reaches_the_abi():
    """Model(depth_test=False) (reference core.py:232-236, triangular.py:117) is passed through, not fenced."""
    m = b2r.Model(np.zeros((3, 4), np.float32), None, None, np.zeros((1, 3, 4), np.int32), depth_test=False)
    packed = _abi.PackedScene([m, b2r.Model(np.zeros((3, 4), np.float32), None, None, np.zeros((1, 3, 4), np.int32))])
    assert packed.models[0].depth_test == 0 and packed.models[1].depth_test == 1


def test_fast_look_at_equals_numpy_formulation():
    """transformation._look_at_basis evaluates the reference's look-at on Python floats; it must be bit-identical
    to the NumPy expressions of obj/transformation.py:83-98 (restated here) for arbitrary cameras."""
    rng = np.random.default_rng(5)

    def numpy_rotate(eye, center, up, sign):
        forward = T.normalize(center - eye).ravel()
        right = T.normalize(np.cross(up, forward)).ravel()
        new_up = np.cross(forward, right)
        rot = np.eye(4)
        rot[:3, :3] = np.column_stack((right, new_up, sign * forward))
        return rot

    for k in range(300):
        eye = rng.standard_normal(3) * rng.choice([0.1, 1, 10, 1000])
        center = rng.standard_normal(3) if k % 3 else np.array([0, 0, 0])
        up = np.array([0, 1, 0]) if k % 2 else rng.standard_normal(3)
        assert np.array_equal(T.look_at_rotate_lh(eye, center, up), numpy_rotate(eye, center, up, -1))
        assert np.array_equal(T.look_at_rotate_rh(eye, center, up), numpy_rotate(eye, center, up, 1))
    m = rng.standard_normal((4, 4))
    planes = np.zeros((6, 4))
    for i, p in enumerate((m[:, 3] + m[:, 0], m[:, 3] - m[:, 0], m[:, 3] + m[:, 1], m[:, 3] - m[:, 1],
                           m[:, 3] + m[:, 2], m[:, 3] - m[:, 2])):
        planes[i] = p / np.linalg.norm(p)
    assert np.array_equal(T.extract_frustum_planes(m), planes)


def _same_model(a, b):
    assert a.vertices.dtype == b.vertices.dtype and np.array_equal(a.vertices, b.vertices)
    for x, y in ((a.uv, b.uv), (a.normals, b.normals)):
        assert (x is None) == (y is None) and (x is None or (x.dtype == y.dtype and np.array_equal(x, y)))
    assert a._faces.dtype == b._faces.dtype == np.int32 and np.array_equal(a._faces, b._faces)
    assert a.material_group == b.material_group and sorted(a.materials) == sorted(b.materials)


def test_native_obj_tokenizer_equals_python_parser(obj_file, tmp_path):
    """the C++ tokenizer must hand back exactly the arrays of the Python (= reference) parser."""
    _same_model(b2r.Model.load_model(obj_file, native=True), b2r.Model.load_model(obj_file, native=False))
    rng = np.random.default_rng(3)
    lines = ["# random mesh", "o thing"]
    nv = 60
    for _ in range(nv):
        lines.append("v " + " ".join(f"{x:.7g}" for x in rng.standard_normal(3) * rng.choice([1e-3, 1, 1e3])))
    for _ in range(40):
        lines.append("vt " + " ".join(f"{x:.6f}" for x in rng.random(rng.choice([2, 3]))) if False else
                     "vt " + " ".join(f"{x:.6f}" for x in rng.random(2)))
    for _ in range(30):
        lines.append("vn " + " ".join(repr(float(x)) for x in rng.standard_normal(3)))
    for k in range(50):
        if k % 17 == 0:
            lines.append(f"usemtl mat{k % 3}")
        n = int(rng.integers(3, 6))
        corners = []
        for _ in range(n):
            a, b_, c = rng.integers(1, nv + 1), rng.integers(1, 41), rng.integers(1, 31)
            corners.append(rng.choice([f"{a}/{b_}/{c}", f"{a}//{c}", f"{a}/{b_}/{c}"]))
        lines.append("f " + "  ".join(corners))
    path = tmp_path / "rand.obj"
    path.write_text("\n".join(lines) + "\n")
    _same_model(b2r.Model.load_model(str(path), native=True), b2r.Model.load_model(str(path), native=False))
    with pytest.raises(FileNotFoundError):
        b2r.Model.load_model(str(tmp_path / "missing.obj"), native=True)


@pytest.mark.skipif(not os.path.isfile("/root/reference/obj/diablo3_pose/diablo3_pose.obj"), reason="asset not present")
def test_native_obj_tokenizer_on_diablo():
    p = "/root/reference/obj/diablo3_pose/diablo3_pose.obj"
    a, b = b2r.Model.load_model(p, native=True), b2r.Model.load_model(p, native=False)
    _same_model(a, b)
    assert a._faces.shape == (5022, 3, 4) and a.vertices.shape == (2519, 4)


def test_bind_host_to_gpu_is_harmless_without_nvml():
    """No GPU / no NVML in the build container: the affinity helper changes nothing and says so."""
    import os
    from py_numpy_renderer_b200 import _native
    before = os.sched_getaffinity(0)
    got = _native.bind_host_to_gpu(0)
    assert got is None or got == len(os.sched_getaffinity(0))
    if got is None:
        assert os.sched_getaffinity(0) == before


def test_window_views_describe_the_blocks():
    """`parallel._DevView` (what FrameWindow hands to torch / CuPy): CUDA array interface of a raw device range."""
    from py_numpy_renderer_b200 import parallel
    v = parallel._DevView(0x7f0000000000, (2, 4, 5, 3))
    cai = v.__cuda_array_interface__
    assert cai["shape"] == (2, 4, 5, 3) and cai["typestr"] == "|u1" and cai["data"] == (0x7f0000000000, False)
    assert cai["strides"] is None and cai["version"] == 3


def test_model_faces_views_and_light_reflect():
    """`model.faces` yields per-triangle views with the reference's attributes (core.py:108-136, 253-255); `Light.reflect`
    (core.py:493-495)."""
    import py_numpy_renderer_b200 as b2r
    import scenes
    model = scenes.floor_model()
    faces = list(model.faces)
    assert len(faces) == len(model._faces) == 2
    f = faces[0]
    assert f.vertices.shape == (3, 4) and f.uv.shape == (3, 3) and f.normals.shape == (3, 3)
    assert f.material is model.materials['default'] and f.textures is model.textures and f.model is model
    n = f.unit_normal_world_space
    assert n.shape == (3,) and abs(float(np.linalg.norm(n)) - 1) < 1e-6 and abs(abs(float(n[1])) - 1) < 1e-6  # plane y = -1
    assert np.array_equal(f.unit_normal_current_space, n)
    out = b2r.Light.reflect(np.array([[1.0, -1.0, 0.0]]), np.array([[0.0, 1.0, 0.0]]))
    assert np.allclose(out, np.array([[1.0, 1.0, 0.0]]) / np.sqrt(2))


def test_model_faces_match_the_reference_when_it_is_importable():
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import refboot
    if not refboot.available():
        pytest.skip("reference sources not mounted")
    import scenes
    ref = refboot.boot()
    mine = scenes.floor_model()
    theirs = ref.Model(mine.vertices.copy(), mine.uv.copy(), mine.normals.copy(), mine._faces.copy())
    for a, b in zip(mine.faces, theirs.faces):
        assert np.array_equal(a.vertices, b.vertices) and np.array_equal(a.uv, b.uv) and np.array_equal(a.normals, b.normals)
        assert np.array_equal(a.unit_normal_world_space, b.unit_normal_world_space)
    I, N = np.array([[0.3, -0.8, 0.5], [1.0, 2.0, 3.0]]), np.array([[0.0, 1.0, 0.0], [0.6, 0.0, 0.8]])
    import py_numpy_renderer_b200 as b2r
    assert np.array_equal(b2r.Light.reflect(I, N), ref.Light.reflect(I, N))


def test_texel_decode_closed_forms():
    """csrc texel_decode(): (float)((double)u8 * (1/255)) and (float)((double)u8 * (2/255) - 1) are bit-identical to the
    reference's f32(u8 / 255) and f32(u8 / 255 * 2 - 1) (core.py:96-104) for every uint8 value."""
    i = np.arange(256)
    assert np.array_equal((i.astype(np.float64) * (np.float64(1) / 255)).astype(np.float32), (i / 255).astype(np.float32))
    assert np.array_equal((i.astype(np.float64) * (np.float64(2) / 255) - 1).astype(np.float32),
                          (i / 255 * 2 - 1).astype(np.float32))


def test_overlay_fast_exit_never_hides_a_segment():
    """overlay.segments() returns [] early when the camera frustum lies strictly inside the debug frustum; whenever it
    does, the full clipping sequence of the reference (frustums.py:46-75) must come out empty as well."""
    import scenes
    from py_numpy_renderer_b200 import overlay
    rng = np.random.default_rng(11)
    sc = scenes.c3_synthetic((90, 160), tex=32, nu=8, nv=4)
    fast_hits = 0
    for k in range(150):
        pos = rng.uniform(-3, 3, 3)
        pos[1] = abs(pos[1]) + 0.2
        center = rng.uniform(-0.5, 0.5, 3)
        fovy, near, far = rng.uniform(30, 80), rng.uniform(0.05, 0.5), rng.uniform(4, 15)
        cam = b2r.Camera(tuple(pos), center=center, fovy=fovy, near=near, far=far)
        if k % 3 == 0:      # containing debug frustum (how every throughput scene is set up)
            dcam = b2r.Camera(tuple(pos), center=center, fovy=fovy + rng.uniform(5, 40), near=near / 2, far=far * 2)
        elif k % 3 == 1:    # unrelated debug camera
            dcam = b2r.Camera(tuple(rng.uniform(-3, 3, 3)), center=rng.uniform(-0.5, 0.5, 3), fovy=rng.uniform(30, 100),
                              near=rng.uniform(0.05, 1), far=rng.uniform(2, 20))
        else:               # nearly identical frusta: the margin decides
            dcam = b2r.Camera(tuple(pos), center=center, fovy=fovy + rng.uniform(-1e-3, 1e-3), near=near, far=far)
        cam.scene = dcam.scene = sc
        if overlay.camera_frustum_inside_debug_frustum(cam, dcam):
            fast_hits += 1
            assert overlay.segments_full(cam, dcam) == []
        else:
            assert len(overlay.segments(cam, dcam)) == len(overlay.segments_full(cam, dcam))
    assert fast_hits >= 40


def test_scene_shading_modes_reach_the_abi():
    """Scene.shading: 'general' = what the reference's render() runs; 'flat' / 'gouraud' / 'pbr' = the
    calls commented out at triangular.py:128-130; anything else is refused."""
    light = b2r.Light((2, 3, 3))
    for name, code in _abi.SHADING.items():
        fp = _abi.pack_frame_params(light, (4, 4), (_abi.B2R_BG_COLOR, (0, 0, 0)), False, shading=name)
        assert fp.shading == code
    with pytest.raises(ValueError):
        _abi.pack_frame_params(light, (4, 4), (_abi.B2R_BG_COLOR, (0, 0, 0)), False, shading='toon')
    m = b2r.Model(np.zeros((3, 4), np.float32), None, None, np.zeros((1, 3, 4), np.int32))
    packed = _abi.PackedScene([m])
    mat = packed.models[0].materials[0]
    assert (mat.Pm, mat.Pr) == (0.5, 0.5) and [mat.Ka[k] for k in range(3)] == [0.3, 0.0, 0.0]   # materials.py:47-49
