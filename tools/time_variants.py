#!/usr/bin/env python3
"""Time several builds of libhcb200.so (kernel variants compiled with different -D switches) in ONE process on the default RANSAC
round and on 1000 hypotheses, and check every variant bit for bit against the committed oracle golden of the default round.
Usage: python tools/time_variants.py <dir-or-.so> ... [--hyp 100,1000]   (development helper; GPU box)"""
import glob
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc

args = [a for a in sys.argv[1:] if not a.startswith("--")]
hyps = [100, 1000]
for a in sys.argv[1:]:
    if a.startswith("--hyp="):
        hyps = [int(v) for v in a[6:].split(",")]
libs = []
for a in args:
    libs += sorted(glob.glob(os.path.join(a, "*.so"))) if os.path.isdir(a) else [a]
prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
gold = np.load(os.path.join(ROOT, "tests", "golden", "oracle_seed0_h100_prune.npz"))
data = {}
for H in hyps:
    picked = hc.sample_hypotheses(0, H, rs["locations"].shape[0])
    data[H] = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
for lib in libs:
    hc.load_library(os.path.abspath(lib))
    line = "%-22s" % os.path.basename(lib)
    for H in hyps:
        trk = hc.Tracker(problem=prob, stats=True)
        trk.upload_params(*data[H])
        for _ in range(3):
            trk.track(H, prune=True)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); trk.track(H, prune=True); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        line += "  H=%d: %.2f ms (med %.2f) %.0f hyp/s" % (H, min(ts), float(np.median(ts)), H / (min(ts) * 1e-3))
        if H == 100:
            tr, cv, inf, st = trk.results(H)
            ok = np.array_equal(np.packbits(cv), gold["converged_bits"]) and np.array_equal(np.packbits(inf), gold["infinity_bits"])
            line += "  golden-flags=%s" % ok
            bad = 0
            for h in range(100):
                a = np.ascontiguousarray(tr[h * 312:(h + 1) * 312, :30]).view(np.float32).copy()
                a[np.isnan(a)] = np.float32(np.nan)
                bad += hashlib.sha256(a.view(np.uint32).tobytes()).hexdigest() != str(gold["digests"][h])
            line += " end-point-digests=%s" % ("ok" if bad == 0 else "%d BAD" % bad)
        info = trk.kernel_info()
        del trk
    print(line + "  " + str(info), flush=True)
