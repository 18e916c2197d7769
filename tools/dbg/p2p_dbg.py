import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ope_pkg; ope_pkg.load()
from ope_b200 import cuda_lib, synth
import orc_py as orc
T = cuda_lib.T
ctx = cuda_lib.Context(0)
model = synth.make_model(20000)
views = synth.turntable_views(model, 36, first=2)
(sp, A), (tp, B) = views
sn, tn = orc.normals_knn(sp, 12), orc.normals_knn(tp, 12)
for iters in (1, 2, 3, 4, 5, 40):
    kw = dict(max_iterations=iters, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8, estimator=T.EST_NORMAL_SHOOTING,
              k_search=20, rejectors=[(T.REJ_SURFACE_NORMAL, 0.7)], with_normals=1, transformation=T.TE_POINT_TO_PLANE)
    cs, ct = ctx.upload(sp, sn), ctx.upload(tp, tn)
    g = ctx.icp(cs, ct, cuda_lib.icp_params(**kw))
    o = orc.icp(sp, tp, orc.icp_params(**kw), src_normals=sn, tgt_normals=tn)
    print(iters, "gpu", g.iterations, g.state, g.n_correspondences, "orc", o.iterations, o.state, o.n_correspondences,
          "maxabs", np.abs(T.mat4(g.T) - T.mat4(o.T)).max(), synth.pose_error(T.mat4(g.T), T.mat4(o.T)))
# single solves on the first iteration's correspondences
prm = cuda_lib.icp_params(estimator=T.EST_NORMAL_SHOOTING, k_search=20, rejectors=[(T.REJ_SURFACE_NORMAL, 0.7)])
cs, ct = ctx.upload(sp, sn), ctx.upload(tp, tn)
q, m, d = ctx.correspondences(cs, ct, prm)
gT, gi = ctx.point_to_plane(cs, ct, q, m, want_info=True)
oT, oi = orc.point_to_plane(sp, tp, tn, q, m, want_info=True)
print("single", gi, oi, np.abs(gT - oT).max())
print(gT - oT)
