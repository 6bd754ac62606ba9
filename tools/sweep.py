#!/usr/bin/env python3
"""Throughput sweep (BASELINE.json configs[4]): H hypotheses x 312 paths on ONE GPU, hypotheses drawn by the reference sampler
(seed 0) from dataset file 000; device-timed with CUDA events.  Usage: python tools/sweep.py 100 1000 10000 [--no-prune] [--abort]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc

args = [a for a in sys.argv[1:] if not a.startswith("--")]
prune = "--no-prune" not in sys.argv
abort = "--abort" in sys.argv
prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
trk = hc.Tracker(problem=prob, stats=True)
trk.set_edgels(rs["locations"], rs["K"])
for H in [int(a) for a in args] or [100, 1000]:
    picked = hc.sample_hypotheses(0, H, rs["locations"].shape[0])
    target, diff = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
    trk.upload_params(target, diff)
    run = (lambda: trk.track_abort(H, prune=prune)) if abort else (lambda: trk.track(H, prune=prune))
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    tr, cv, inf, st = trk.results(H)
    stages = float(st[:, 1].sum() + st[:, 2].sum())
    flops = 91312.0 * float(st[:, 1].sum()) + 90874.0 * float(st[:, 2].sum())
    ms = min(ts)
    print("H=%6d prune=%s abort=%s  %9.2f ms  %8.0f hyp/s  %10.0f paths/s  %.1f TFLOP/s(model)  conv=%d inf=%d stages/path=%.1f"
          % (H, prune, abort, ms, H / ms * 1e3, H * 312 / ms * 1e3, flops / ms / 1e9, int(cv.sum()), int(inf.sum()), stages / (H * 312)))
