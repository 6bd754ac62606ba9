#!/usr/bin/env python3
"""Parity report against the REAL reference GPU-HC++ kernels (oracle/_ref/libref_gpuhc.so, unmodified sources built for sm_100a)
on the same GPU and inputs: flag agreement and end-point agreement, raw and after double-precision Newton refinement of both end
points against the target system (north_star: "converged solutions within 1e-4 relative after Newton refinement").
Lives under tests/ because it uses the oracle's refiner; not a pytest module.  GPU box only.
    python tests/parity_report.py [n_hyp] > profiles/parity_r1.txt"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyoracle import Oracle, ReferenceGPU
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc

H = int(sys.argv[1]) if len(sys.argv) > 1 else 100
prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
orc = Oracle(prob)
picked = hc.sample_hypotheses(0, H, rs["locations"].shape[0])
target, diff = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])

ref = ReferenceGPU(prob)
ref.setup(target, diff, rs["locations"], rs["K"])
ref.track()
tr_r, cv_r, inf_r = ref.results()
trk = hc.Tracker(problem=prob, stats=True)
trk.upload_params(target, diff)
trk.track(H, prune=True)
tr, cv, inf, st = trk.results(H)

P = H * 312
print("# parity report: this library vs the reference GPU-HC++ kernels, seed 0, %d hypotheses (%d paths), pruning on" % (H, P))
print("converged:  ours %d  reference %d   both %d  only ours %d  only reference %d   agreement %.3f %%"
      % (cv.sum(), cv_r.sum(), (cv & cv_r).sum(), (cv & ~cv_r & 1).sum(), (~cv & cv_r & 1).sum(), 100.0 * (cv == cv_r).mean()))
print("infinity:   ours %d  reference %d   both %d  only ours %d  only reference %d   agreement %.3f %%"
      % (inf.sum(), inf_r.sum(), (inf & inf_r).sum(), (inf & ~inf_r & 1).sum(), (~inf & inf_r & 1).sum(), 100.0 * (inf == inf_r).mean()))
mine, theirs = hc.count_solutions(tr, cv, inf, H), hc.count_solutions(tr_r, cv_r, inf_r, H)
print("per-hypothesis counts (converged, infinity, real): identical in %d / %d / %d of %d hypotheses; max |diff| %d / %d / %d"
      % tuple([(mine[:, k] == theirs[:, k]).sum() for k in range(3)] + [H] + [np.abs(mine[:, k] - theirs[:, k]).max() for k in range(3)]))
both = np.nonzero((cv == 1) & (cv_r == 1))[0]
rel = np.array([np.abs(tr[b, :30] - tr_r[b, :30]).max() / max(np.abs(tr_r[b, :30]).max(), 1e-30) for b in both])
q = lambda a: "median %.2e  90%% %.2e  99%% %.2e  max %.2e" % (np.median(a), np.quantile(a, 0.9), np.quantile(a, 0.99), a.max())
print("end points of the %d paths converged in both, max-norm relative difference, RAW: %s" % (len(both), q(rel)))
print("   within 1e-4: %.2f %%   within 1e-3: %.2f %%" % (100.0 * (rel < 1e-4).mean(), 100.0 * (rel < 1e-3).mean()))
ref_rel, res_a, res_b, move_a, move_b = [], [], [], [], []
for b in both:
    h = b // 312
    xa, ra = orc.newton_refine(target[h], tr[b], iters=8)
    xb, rb = orc.newton_refine(target[h], tr_r[b], iters=8)
    nrm = max(1.0, np.abs(xb).max())
    ref_rel.append(np.abs(xa - xb).max() / nrm)
    move_a.append(np.abs(xa - tr[b, :30]).max() / nrm)        # how far the polish moved each end point
    move_b.append(np.abs(xb - tr_r[b, :30]).max() / nrm)
    res_a.append(ra); res_b.append(rb)
ref_rel, res_a, res_b, move_a, move_b = map(np.array, (ref_rel, res_a, res_b, move_a, move_b))
fin = np.isfinite(ref_rel) & np.isfinite(res_a) & np.isfinite(res_b)
# an end point is REGULAR when Newton polishes it to a root without leaving its neighbourhood (an isolated, non-singular solution
# of the target system); the rest are end points at singular / positive-dimensional solutions where Newton wanders
reg = fin & (res_a < 1e-9) & (res_b < 1e-9) & (move_a < 1e-2) & (move_b < 1e-2)
print("after 8 double-precision Newton iterations on the target system (difference / max(1, |x|)):")
print("   regular end points in BOTH trackers (polish converges, residual < 1e-9, moves < 1e-2): %d of %d" % (reg.sum(), len(both)))
print("      %s" % q(ref_rel[reg]))
print("      within 1e-4: %d of %d (%.3f %%)" % ((ref_rel[reg] < 1e-4).sum(), reg.sum(), 100.0 * (ref_rel[reg] < 1e-4).mean()))
sing = fin & ~reg
print("   singular / ill-conditioned end points (polish does not settle nearby in at least one tracker): %d; RAW difference there: %s"
      % (sing.sum(), q(rel[sing]) if sing.any() else "-"))
print("   RAW difference on the regular ones: %s" % q(rel[reg]))
# the pose RANSAC returns
sup, best = None, None
print("hypothesis 0 / track 104 (ground-truth pose): converged ours %d reference %d, raw relative difference %.2e"
      % (cv[104], cv_r[104], np.abs(tr[104, :30] - tr_r[104, :30]).max() / np.abs(tr_r[104, :30]).max()))
