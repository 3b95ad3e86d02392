"""The C-ABI library builds, loads and exports every symbol include/b2r.h declares; the ctypes mirror of its structs
has the C layout; and without a GPU the product fails loudly instead of falling back to anything."""
import ctypes
import os
import re
import subprocess
import tempfile

import pytest

from py_numpy_renderer_b200 import _abi, _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b2r.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2r_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load_library()
    names = declared_functions()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), name
        assert name in _native._SYMBOLS, f"{name} is declared in b2r.h but not bound in _native.py"
    assert lib.b2r_abi_version() == _abi.B2R_ABI_VERSION


def test_ctypes_structs_match_c_layout():
    structs = {"b2r_texture_desc": _abi.TextureDesc, "b2r_material": _abi.MaterialDesc, "b2r_model_desc": _abi.ModelDesc,
               "b2r_cubemap_desc": _abi.CubemapDesc, "b2r_view": _abi.View, "b2r_light": _abi.LightDesc,
               "b2r_frame_params": _abi.FrameParams, "b2r_debug_out": _abi.DebugOut}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append('return 0;}')
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "l.c"), os.path.join(d, "l")
        open(src, "w").write("\n".join(lines))
        subprocess.check_call(["gcc", "-o", exe, src])
        out = dict(l.split() for l in subprocess.check_output([exe], text=True).splitlines())
    for cname, ct in structs.items():
        assert int(out[cname]) == ctypes.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    import golden_util as gu
    lib = _native.load_library()
    assert lib.b2r_init(0) != 0
    assert b"no CPU fallback" in lib.b2r_last_error()
    scene, _, _ = gu.load("g7_cube_mtl_rh_directx")
    with pytest.raises(RuntimeError, match="b2r"):
        scene.render()


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/ (the product has no CPU path)."""
    pkg = os.path.join(ROOT, "py_numpy_renderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "import oracle" not in text and "orc_render" not in text, f
