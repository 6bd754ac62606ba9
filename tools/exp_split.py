#!/usr/bin/env python3
"""Experiment (GPU box): HCB200_FLAG_SPLIT_LONG_PATHS on / off and the step at which long paths are parked; every run is checked bit for
bit against the committed oracle golden of the default round."""
import hashlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc
prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
gold = np.load(os.path.join(ROOT, "tests", "golden", "oracle_seed0_h100_prune.npz"))
hyps = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "100,1000").split(",")]
data = {}
for H in hyps:
    picked = hc.sample_hypotheses(0, H, rs["locations"].shape[0])
    data[H] = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
configs = [("unsplit", False, 0)] + [("split at step %d" % k, True, k) for k in (40, 48, 56, 60, 64, 68, 72)] + [("split (default 4/5)", True, 0)]
for name, split, k in configs:
    line = "%-22s" % name
    for H in hyps:
        trk = hc.Tracker(problem=prob, stats=True, split=split)
        trk.suspend_step = k
        trk.upload_params(*data[H])
        for _ in range(3):
            trk.track(H, prune=True)
        torch.cuda.synchronize()
        ts = []
        for _ in range(9):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); trk.track(H, prune=True); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        line += "  H=%d: %.2f ms (med %.2f) %.0f hyp/s" % (H, min(ts), float(np.median(ts)), H / (min(ts) * 1e-3))
        if H == 100:
            tr, cv, inf, st = trk.results(H)
            ok = np.array_equal(np.packbits(cv), gold["converged_bits"]) and np.array_equal(np.packbits(inf), gold["infinity_bits"])
            bad = 0
            for h in range(100):
                a = np.ascontiguousarray(tr[h * 312:(h + 1) * 312, :30]).view(np.float32).copy()
                a[np.isnan(a)] = np.float32(np.nan)
                bad += hashlib.sha256(a.view(np.uint32).tobytes()).hexdigest() != str(gold["digests"][h])
            ok_st = np.array_equal(st[:, 0].astype(np.uint8), gold["steps"]) and np.array_equal(st.sum(0)[:3].astype(np.int64), gold["stats_sum"][:3])
            line += "  flags=%s digests=%s stats=%s" % (ok, "ok" if bad == 0 else "%d BAD" % bad, ok_st)
        del trk
    print(line, flush=True)
