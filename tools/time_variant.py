#!/usr/bin/env python3
"""Time one build of libhcb200.so on the default RANSAC round (100 hyp x 312 paths, pruning on) and check one hypothesis
against the oracle.  Usage: python tools/time_variant.py <lib.so> [n_hyp]   (development helper for kernel variants)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.pyoracle import Oracle
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc

lib = sys.argv[1]
H = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hc.load_library(os.path.abspath(lib))
prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
picked = hc.sample_hypotheses(0, H, rs["locations"].shape[0])
target, diff = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
trk = hc.Tracker(problem=prob, stats=True)
trk.upload_params(target, diff)
for _ in range(3):
    trk.track(H, prune=True)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); trk.track(H, prune=True); b.record(); b.synchronize()
    ts.append(a.elapsed_time(b))
tr, cv, inf, st = trk.results(H)
orc = Oracle(prob)
tr_o, cv_o, inf_o, st_o = orc.track(target[:1], diff[:1], True)
same = bool(np.array_equal(cv[:312], cv_o) and np.array_equal(inf[:312], inf_o) and
            np.all((tr[:312, :30].view(np.uint64) == tr_o[:, :30].view(np.uint64)) | (np.isnan(tr[:312, :30]) & np.isnan(tr_o[:, :30]))))
print("%s: H=%d  min %.2f ms  median %.2f ms  -> %.0f hyp/s  conv=%d inf=%d  oracle-bit-exact(hyp0)=%s  info=%s"
      % (os.path.basename(lib), H, min(ts), float(np.median(ts)), H / (min(ts) * 1e-3), int(cv.sum()), int(inf.sum()), same,
         trk.kernel_info()))
