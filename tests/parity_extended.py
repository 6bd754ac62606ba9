#!/usr/bin/env python3
"""Extended bit-exactness run (not a pytest module; uses the oracle, hence under tests/): GPU tracker vs CPU oracle on hypotheses the
committed goldens do not cover — other sampler seeds, other dataset files, both pruning modes.  GPU box only (the oracle runs on its
host cores).    python tests/parity_extended.py [hypotheses_per_case] > profiles/parity_extended_r1.txt"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyoracle import Oracle
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc

H = int(sys.argv[1]) if len(sys.argv) > 1 else 60
prob = fixtures.load_problem()
orc = Oracle(prob)
trk = hc.Tracker(problem=prob, stats=True)
total_paths = total_bad = 0
print("# GPU tracker vs CPU oracle, bit for bit (flags, step/stage counters, end reason, every end-point component; NaN == NaN)")
for dataset, seed, prune in [(0, 11, True), (0, 12, False), (1, 21, True), (2, 22, False), (3, 23, True), (3, 24, False)]:
    rs = fixtures.load_ransac(dataset)
    picked = hc.sample_hypotheses(seed, H, rs["locations"].shape[0])
    target, diff = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
    t0 = time.time()
    tr_o, cv_o, inf_o, st_o = orc.track(target, diff, prune)
    t_or = time.time() - t0
    trk.upload_params(target, diff)
    trk.track(H, prune=prune)
    tr_g, cv_g, inf_g, st_g = trk.results(H)
    a, b = np.ascontiguousarray(tr_g[:, :30]).view(np.uint32).reshape(H * 312, -1), np.ascontiguousarray(tr_o[:, :30]).view(np.uint32).reshape(H * 312, -1)
    nan_both = np.isnan(np.ascontiguousarray(tr_g[:, :30]).view(np.float32).reshape(H * 312, -1)) & np.isnan(np.ascontiguousarray(tr_o[:, :30]).view(np.float32).reshape(H * 312, -1))
    same_x = ((a == b) | nan_both).all(axis=1)
    same = same_x & (cv_g == cv_o) & (inf_g == inf_o) & (st_g[:, :3] == st_o[:, :3]).all(axis=1) & ((st_g[:, 3] & 0xffff) == st_o[:, 3]) & ((st_g[:, 3] >> 16) == st_o[:, 4])
    total_paths += H * 312
    total_bad += int((~same).sum())
    print("dataset %03d seed %2d prune %-5s: %6d paths, %5d converged, %5d infinity, %7.1f stages/path  -> identical paths %d / %d   (oracle %.0f s)"
          % (dataset, seed, prune, H * 312, cv_g.sum(), inf_g.sum(), (st_g[:, 1] + st_g[:, 2]).mean(), same.sum(), H * 312, t_or), flush=True)
print("TOTAL: %d of %d paths identical" % (total_paths - total_bad, total_paths))
sys.exit(1 if total_bad else 0)
