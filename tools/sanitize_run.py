#!/usr/bin/env python3
"""Small workload that touches every kernel of libhcb200.so, meant to run under compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
Tracker with and without split (parking at step 2, so every path is parked and resumed), early abort, refinement, scoring, statistics, pose
records, device target parameters, and a compiled non-trifocal problem.  Results are checked against the oracle as usual.  (compute-sanitizer is closed on the round-2 GPU pool; the workload itself runs clean and
bit-exact there, and doubles as a one-command smoke of every kernel.)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle.pyoracle import Oracle
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc, problem as pm

H = int(sys.argv[1]) if len(sys.argv) > 1 else 12
prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
orc = Oracle(prob)
picked = hc.sample_hypotheses(0, H, rs["locations"].shape[0])
tgt, dif = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
tr_o, cv_o, inf_o, st_o = orc.track(tgt, dif, True)
for split, cut in ((False, 0), (True, 2), (True, 0)):
    trk = hc.Tracker(problem=prob, stats=True, split=split)
    trk.suspend_step = cut
    trk.set_edgels(rs["locations"], rs["K"])
    trk.upload_params(tgt, dif)
    trk.track(H, prune=True)
    tr, cv, inf, st = trk.results(H)
    assert np.array_equal(cv, cv_o) and np.array_equal(inf, inf_o)
    a, b = np.ascontiguousarray(tr[:, :30]), np.ascontiguousarray(tr_o[:, :30])
    assert bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))
    print("track split=%s cut=%d ok" % (split, cut), flush=True)
trk.score_tracks(H)
trk.refine_tracks(H, iters=2)
trk.count_solutions_device(H) if hasattr(trk, "count_solutions_device") else None
trk.track_abort(H, prune=True)
torch.cuda.synchronize()
print("abort / score / refine / statistics ok", flush=True)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_synthetic_problem as msp
for name in msp.NAMES:
    msp.select(name)
    pdir = os.path.join(ROOT, "problems", name)
    p2 = pm.read_problem(pdir)
    t2 = msp.target_params(8)
    pt = pm.ProblemTracker(pdir, problem=p2, stats=True)
    pt.upload_params(t2)
    pt.track(8, prune=False)
    tr, cv, inf, st = pt.results(8)
    o2 = Oracle(p2, problem_dir=pdir)
    tr_o2, cv_o2, inf_o2, _ = o2.track(t2, pt.diff_params(t2), False)
    assert np.array_equal(cv, cv_o2) and np.array_equal(inf, inf_o2)
    print("compiled problem %s ok" % name, flush=True)
print("done")
