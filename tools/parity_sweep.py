import sys, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, R+'/oracle', R+'/tests']
import numpy as np
import golden_util as gu, oracle as orc
for name in gu.fixture_names():
    scene, exp, meta = gu.load(name)
    ref = gu.oracle_frame(orc, scene)   # oracle + the host overlay pass where the debug frustum is visible (g10, g11)
    dbg = {}
    scene.persist_silhouette = False
    rgb = scene.render(debug=dbg)
    got = dict(rgb=rgb, z=dbg['z'], stencil=dbg['stencil'], winner=dbg['winner'])
    print(name, 'vs oracle', gu.compare_planes(got, ref), 'status mism', int((dbg['face_status']!=ref['face_status']).sum()), 'nsil', dbg['n_silhouette'], ref['n_silhouette'])
    print('   vs reference', gu.compare_planes(got, exp))
