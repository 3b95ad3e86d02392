"""Pins the CPU oracle (oracle/oracle.c) to the UNMODIFIED Python reference: fixtures in tests/golden/ were produced
by running the reference (oracle/make_golden.py); the oracle must reproduce its z-buffer, stencil buffer and
winner faces bit for bit and its uint8 frame within 1 LSB on >= 99.9 % of the pixels (BASELINE.json north star)."""
import numpy as np
import pytest

import golden_util as gu
from py_numpy_renderer_b200 import _abi

NAMES = gu.fixture_names()


def test_fixtures_present():
    assert len(NAMES) >= 13


@pytest.mark.parametrize("name", NAMES)
def test_host_matrices_equal_reference(name):
    """The host API must hand the kernels the very float64 matrices the reference computes (core.py:394-429)."""
    scene, exp, _ = gu.load(name)
    v = _abi.pack_view(scene.camera, scene.debug_camera, scene.system, False)
    assert np.array_equal(np.array(v.mvp).reshape(4, 4), exp['ref_mvp'])
    assert np.array_equal(np.array(v.mvp_dbg).reshape(4, 4), exp['ref_mvp_dbg'])
    assert np.array_equal(np.array(v.viewport).reshape(4, 4), exp['ref_viewport'])
    assert np.array_equal(np.array(v.planes).reshape(6, 4), exp['ref_planes'])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference(name, oracle):
    scene, exp, meta = gu.load(name)
    got = gu.oracle_frame(oracle, scene)   # + the host-side frustum overlay when the debug frustum is visible
    rep = gu.compare_planes(got, exp)
    assert rep['z_mismatch'] == 0, rep          # float64 z-buffer, bit exact
    assert rep['stencil_mismatch'] == 0, rep    # shadow-volume counts, exact
    assert rep['winner_mismatch'] == 0, rep     # z-test winners, exact
    assert np.array_equal(got['winner1'], exp['winner1'])
    assert rep['rgb_px_gt1'] == 0, rep                              # never more than 1 LSB
    assert rep['rgb_px_diff'] <= rep['pixels'] // 1000, rep         # >= 99.9 % identical
    assert list(got['n_silhouette']) == meta['n_silhouette']


def test_oracle_threads_deterministic(oracle):
    """The multi-threaded driver used by the CPU baseline renders the same bytes as a single thread."""
    scene, _, _ = gu.load("g7_cube_mtl_rh_directx")
    cams = [scene.camera] * 3
    a = oracle.render_scene(scene, cameras=cams, threads=1)
    b = oracle.render_scene(scene, cameras=cams, threads=3)
    for k in ('rgb', 'z', 'stencil', 'winner'):
        assert np.array_equal(a[k], b[k], equal_nan=(k == 'z'))
    assert np.array_equal(a['rgb'][0], a['rgb'][2])
