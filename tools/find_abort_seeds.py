#!/usr/bin/env python3
"""For sampler seeds 0..N-1: where in the 100-hypothesis round does the first passing pose sit, and how long does the
early-abort launch take?  (SURVEY.md §8d config 3 asks for a seed whose first hit is late.)  Runs on the GPU box.
Usage: python tools/find_abort_seeds.py [n_seeds] [n_hyp]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 100
prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
trk = hc.Tracker(problem=prob, stats=True)
trk.set_edgels(rs["locations"], rs["K"])
for seed in range(n_seeds):
    picked = hc.sample_hypotheses(seed, H, rs["locations"].shape[0])
    target, diff = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
    trk.upload_params(target, diff)
    # full round, device scoring: the deterministic answer
    trk.track(H, prune=True)
    sup, best_full = trk.score_tracks(H)
    tr, cv, inf, st = trk.results(H)
    E = rs["locations"].shape[0]
    passing = np.nonzero((sup[:, 0] >= 0.9 * E) & (sup[:, 1] >= 0.9 * E))[0]
    first = int(passing.min()) if len(passing) else -1
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); trk.track_abort(H, prune=True); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    best = trk.d_best.cpu().numpy()
    print("seed %2d: passing paths in the full round %3d, first at hypothesis %3d (path %5d); abort launch %.2f ms, best record %s"
          % (seed, len(passing), first // 312 if first >= 0 else -1, first, min(ts), best[:5].tolist()), flush=True)
