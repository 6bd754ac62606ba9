#!/usr/bin/env python3
"""Launch the three kernels besides the plain tracker once each on a realistic input, so that ncu can capture them
(`-k regex:hc_track_kernelILb1|hc_refine|hc_score_tracks`): the early-abort tracker on a late-hit round (sampler seed 13: the
first passing pose sits in hypothesis 30 of 100), the Newton refinement (3 iterations) and the final scoring of the default round.
Prints CUDA-event times.  GPU box only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc

prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
H = 100
trk = hc.Tracker(problem=prob, stats=True)
trk.set_edgels(rs["locations"], rs["K"])


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


for seed in (13, 0):
    picked = hc.sample_hypotheses(seed, H, rs["locations"].shape[0])
    target, diff = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
    trk.upload_params(target, diff)
    ms = timed(lambda: trk.track_abort(H, prune=True))
    best = trk.d_best.cpu().numpy()
    tr, cv, inf, st = trk.results(H)
    ran = ((st[:, 3] >> 16) < 4).sum()
    print("abort tracker, sampler seed %d: %.3f ms, first passing path %d (hypothesis %d), %d of %d paths ran, %d converged paths scored in-kernel"
          % (seed, ms, best[1], best[1] // 312, ran, H * 312, int(cv.sum())))
trk.track(H, prune=True)
torch.cuda.synchronize()
tr, cv, inf, st = trk.results(H)
ms = timed(lambda: trk.score_tracks_async(H))
print("final scoring of the default round: %.3f ms for %d paths, %d converged, %d edgel triplets -> %.2f us per converged path"
      % (ms, H * 312, int(cv.sum()), rs["locations"].shape[0], ms * 1e3 / max(int(cv.sum()), 1)))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); trk.refine_tracks(H, iters=3); b.record(); b.synchronize()
print("Newton refinement, 3 iterations on %d converged end points: %.3f ms" % (int(cv.sum()), a.elapsed_time(b)))
