"""Why does scene 861 of tools/pose_agreement.py disagree (2 mrad / 2.7 mm) although SAC-IA is bit-identical? Replays the fine stage
(D&L/src/poseestimator.cpp:161-379) by hand: same clouds on both sides, then ICP on the device once with the DEVICE's normals and
once with the ORACLE's normals."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ope_pkg; ope_pkg.load()
from ope_b200 import cuda_lib, synth
import orc_py as orc
T = cuda_lib.T
f = int(sys.argv[1]) if len(sys.argv) > 1 else 861
model = synth.make_model()
cl, _, pose = synth.make_frame(model, 1000 + f)
pe = orc.PoseEstimator(); src = model.copy(); orc.srand(1)
p = pe.estimate_final(src, cl)
coarse = T.mat4(p.coarse_pose)
aligned = orc.transform(model, coarse)
sp = aligned[orc.uniform_sample(aligned, 0.008)]
tp = cl[orc.uniform_sample(cl, 0.008)]
sn_o, tn_o = orc.normals_knn(sp, 30), orc.normals_knn(tp, 30)
ctx = cuda_lib.Context(0)
cs, ct = ctx.upload(sp), ctx.upload(tp)
sn_g, tn_g = ctx.normals_knn(cs, 30), ctx.normals_knn(ct, 30)
print("normals max |device - oracle|: source %.3g target %.3g; differing entries: %d / %d" %
      (np.nanmax(np.abs(sn_g - sn_o)), np.nanmax(np.abs(tn_g - tn_o)), int((sn_g != sn_o).sum() + (tn_g != tn_o).sum()), sn_g.size + tn_g.size))
kw = dict(max_iterations=100, transformation_epsilon=1e-8, euclidean_fitness_epsilon=1e-8, estimator=T.EST_NORMAL_SHOOTING,
          k_search=20, rejectors=[(T.REJ_SURFACE_NORMAL, 0.7), (T.REJ_SELF_OCCLUDED_NORMAL, 0.6)], with_normals=1)
o = orc.icp(sp, tp, orc.icp_params(**kw), src_normals=sn_o, tgt_normals=tn_o)
for name, (a, b) in (("device normals", (sn_g, tn_g)), ("oracle normals", (sn_o, tn_o))):
    g = ctx.icp(ctx.upload(sp, a), ctx.upload(tp, b), cuda_lib.icp_params(**kw))
    r, t = synth.pose_error(T.mat4(g.T), T.mat4(o.T))
    print("device ICP with %s vs oracle ICP (oracle normals): rot %.3g rad, trans %.3g m, iterations %d/%d, correspondences %d/%d"
          % (name, r, t, g.iterations, o.iterations, g.n_correspondences, o.n_correspondences))

# first iteration at which the two sides differ (same normals on both sides), and what differs there
first = None
for it in range(1, 101):
    kw["max_iterations"] = it
    o = orc.icp(sp, tp, orc.icp_params(**kw), src_normals=sn_o, tgt_normals=tn_o, want_corr=True)
    g = ctx.icp(ctx.upload(sp, sn_o), ctx.upload(tp, tn_o), cuda_lib.icp_params(**kw), want_corr=True)
    same_T = np.array_equal(np.array(list(g[0].T)), np.array(list(o[0].T)))
    gq, gm, gd = g[1]; oq, om, od = o[1]
    same_c = len(gq) == len(oq) and np.array_equal(gq, oq) and np.array_equal(gm, om)
    if not (same_T and same_c):
        first = it
        print("first difference at iteration", it, "| T identical:", same_T, "| correspondences identical:", same_c, len(gq), len(oq))
        sg, so = set(zip(gq.tolist(), gm.tolist())), set(zip(oq.tolist(), om.tolist()))
        print("only device:", sorted(sg - so)[:5], "only oracle:", sorted(so - sg)[:5])
        break
if first is None:
    print("no difference in 100 iterations")
