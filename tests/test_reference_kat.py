"""Known-answer tests from SURVEY.md Appendix C: the oracle, fed the reference's own assets at the reference's
own resolutions, must reproduce the SHA-256 of the frames the unmodified reference renders.  Needs the assets
(/root/reference in the build container, or baseline/_ref/assets on the GPU box); skipped elsewhere."""
import hashlib
import os

import numpy as np
import pytest

import scenes

ASSETS = scenes.asset_root()
pytestmark = pytest.mark.skipif(ASSETS is None, reason="reference assets not available")


def test_kat1_diablo_800(oracle):
    scene = scenes.kat1(ASSETS)
    rgb = oracle.render_scene(scene, planes=False)['rgb'][0]
    assert hashlib.sha256(rgb.tobytes()).hexdigest()[:16] == "d8208bb3d5da0422"
    assert int((rgb != rgb[0, 0]).any(axis=-1).sum()) == 42243  # non-background pixels


def test_kat2_diablo_floor_1080p(oracle):
    scene = scenes.kat2(ASSETS)
    out = oracle.render_scene(scene, planes=False)
    rgb = out['rgb'][0]
    assert list(out['n_silhouette'][0]) == [1381, 4]
    digest = hashlib.sha256(rgb.tobytes()).hexdigest()[:16]
    if digest != "7fd0ce3d94f34a5d":
        # float32 pow / LAPACK-level differences may flip single LSBs; everything structural must still agree
        assert abs(rgb.reshape(-1, 3).mean(0) - np.array([86.948, 139.588, 179.666])).max() < 0.01
    assert int((rgb != np.array([84, 146, 208], np.uint8)).any(axis=-1).sum()) == 619692
