#!/usr/bin/env python3
"""Experiment (GPU box): how much of the default round's time is queue drain?  Runs a -DHC_DEBUG_ORDER build with the path queue in natural
order, longest-path-first (lengths from the committed oracle golden: an upper bound no real scheduler has), shortest-first and random."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc
lib = hc.load_library(os.path.abspath(sys.argv[1]))
prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
gold = np.load(os.path.join(ROOT, "tests", "golden", "oracle_seed0_h100_prune.npz"))
H = 100
picked = hc.sample_hypotheses(0, H, rs["locations"].shape[0])
tgt, dif = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
trk = hc.Tracker(problem=prob, stats=True)
trk.upload_params(tgt, dif)
steps = gold["steps"].astype(np.int64)
orders = {"natural": None, "longest first": np.argsort(-steps, kind="stable"), "shortest first": np.argsort(steps, kind="stable"),
          "random": np.random.RandomState(0).permutation(len(steps))}
lib.hcb200_debug_set_order.argtypes = [ctypes.c_void_p]
for name, o in orders.items():
    d = torch.from_numpy(o.astype(np.int32)).cuda() if o is not None else None
    assert lib.hcb200_debug_set_order(ctypes.c_void_p(d.data_ptr()) if d is not None else None) == 0
    for _ in range(3):
        trk.track(H, prune=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(9):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); trk.track(H, prune=True); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    tr, cv, inf, st = trk.results(H)
    ok = np.array_equal(np.packbits(cv), gold["converged_bits"])
    print("%-16s %.2f ms (median %.2f)  flags==golden %s" % (name, min(ts), float(np.median(ts)), ok), flush=True)
