#!/usr/bin/env python3
"""Small end-to-end invocation of every device entry point (for compute-sanitizer runs)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc
prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
H = 2
picked = hc.sample_hypotheses(0, H, rs["locations"].shape[0])
target, diff = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
trk = hc.Tracker(problem=prob, stats=True)
trk.set_edgels(rs["locations"], rs["K"])
trk.upload_params(target, diff)
trk.track(H, prune=True); trk.results(H)
trk.track(H, prune=False); tr, cv, inf, st = trk.results(H)
sup, best = trk.score_tracks(H)
trk.track_abort(H, prune=True); trk.results(H)
d_picked = torch.from_numpy(picked).to(trk.device)
d_tan = torch.from_numpy(np.ascontiguousarray(rs["tangents"])).to(trk.device)
trk.build_target_params(d_picked, d_tan, H)
torch.cuda.synchronize()
print("ok", int(cv.sum()), best[:5].tolist(), int(trk.d_found.cpu()[0]))
