#!/usr/bin/env python3
"""Timing harness (GPU box): what does each component of a stage cost IN PLACE?  Builds of csrc/hc_tracker.cu with -DHC_HARNESS=<bits> discard every
stage's result (so all builds run the same, deterministic number of stages) and remove one component each; see the macro's comment.
    for b in 0 1 2 3 4 7 8 24 32 63; do nvcc ... -DHC_HARNESS=$b -shared -o lib/variants/harness_$b.so csrc/hc_tracker.cu; done
    python tools/exp_harness.py trifocal_pose_estimation_using_improved_gpuhc_b200/lib/variants"""
import glob, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc
NAMES = {0: "everything (result discarded)", 1: "- pivot-row stores of the 4 super-steps", 2: "- pivot-row stores of the 12 warp-wide steps", 3: "- all pivot-row stores",
         4: "- pivot-row loads", 7: "- all pivot-row stores and loads", 8: "- evaluator operand gathers", 24: "- evaluator gathers and operand words",
         32: "- x-product / coefficient table builds", 63: "- all of the above (arithmetic, shuffles, control only)"}
prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
H = 1000
picked = hc.sample_hypotheses(0, H, rs["locations"].shape[0])
tgt, dif = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
base = None
for lib in sorted(glob.glob(os.path.join(sys.argv[1], "harness_*.so")), key=lambda p: int(p.split("_")[-1][:-3])):
    bits = int(lib.split("_")[-1][:-3])
    hc.load_library(os.path.abspath(lib))
    trk = hc.Tracker(problem=prob, stats=True, split=False)
    trk.upload_params(tgt, dif)
    for _ in range(2):
        trk.track(H, prune=False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); trk.track(H, prune=False); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    st = trk.results(H)[3]
    stages = int(st[:, 1].sum() + st[:, 2].sum())
    ms = min(ts)
    cyc = ms * 1e-3 * 1.965e9 * 148 / stages          # SM-cycles per stage (all 20 resident warps of an SM together)
    if base is None:
        base = cyc
    print("%-62s %8.2f ms  %9d stages  %6.1f SM-cycles/stage  (%+6.1f = %+5.1f %%)" % ("%2d %s" % (bits, NAMES.get(bits, "")), ms, stages, cyc, cyc - base, 100 * (cyc - base) / base), flush=True)
    del trk
