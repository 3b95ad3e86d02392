"""Seeded random scenes (random closed meshes, triangle soups with degenerate faces, planes; random textures,
tangent / object-space normal maps, specular maps; all light kinds; all system x subsystem combinations; float32 and
`@`-transformed float64 vertices; clip on/off; culling on/off):

* CPU, build container only (needs /root/reference): oracle == UNMODIFIED reference, bit-exact z / stencil / winners;
* GPU: CUDA == oracle on a larger set of seeds."""
import numpy as np
import pytest

import golden_util as gu
import scenes

CPU_SEEDS = list(range(100, 112))
GPU_SEEDS = list(range(100, 140))


def _ref_available():
    import refboot
    return refboot.available()


@pytest.mark.skipif(not _ref_available(), reason="the Python reference is only present in the build container")
@pytest.mark.parametrize("seed", CPU_SEEDS)
def test_oracle_equals_reference_on_random_scene(seed, oracle):
    import refboot
    ref = refboot.boot()
    spec = scenes.random_scene_spec(seed)

    def ref_texture(t):   # what TextureMaps.register would have stored (core.py:90-105)
        arr = t['texels'] / 255
        if t['signed']:
            arr = arr * 2 - 1
        return np.array(arr, dtype=np.dtype(np.float32, metadata={'tangent': t['tangent']}))

    ref_scene = scenes.build_random_scene(spec, api=ref, texture_factory=ref_texture)
    want = refboot.instrumented_render(ref_scene)
    want['winner'] = np.where(want['winner3'] >= 0, want['winner3'], want['winner1'])
    got = oracle.render_scene(scenes.build_random_scene(spec), extra=True)
    rep = gu.compare_planes(got, want)
    assert rep['z_mismatch'] == 0 and rep['stencil_mismatch'] == 0 and rep['winner_mismatch'] == 0, (seed, rep)
    assert rep['rgb_px_gt1'] == 0 and rep['rgb_px_diff'] * 1000 <= max(rep['pixels'], 1000), (seed, rep)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", GPU_SEEDS)
def test_cuda_equals_oracle_on_random_scene(seed, oracle):
    spec = scenes.random_scene_spec(seed)
    scene = scenes.build_random_scene(spec)
    want = {k: v[0] for k, v in oracle.render_scene(scene).items()}
    dbg = {}
    rgb = scene.render(debug=dbg)
    assert np.array_equal(scene.render(), rgb)          # production path (stencil shortcuts) == debug path
    got = dict(rgb=rgb, z=dbg['z'], stencil=dbg['stencil'], winner=dbg['winner'])
    rep = gu.compare_planes(got, want)
    assert rep['z_mismatch'] == 0 and rep['stencil_mismatch'] == 0 and rep['winner_mismatch'] == 0, (seed, rep)
    assert rep['rgb_px_gt1'] == 0 and rep['rgb_px_diff'] * 1000 <= max(rep['pixels'], 1000), (seed, rep)
    assert np.array_equal(dbg['face_status'], want['face_status']), seed
