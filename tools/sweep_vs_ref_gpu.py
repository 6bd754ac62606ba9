#!/usr/bin/env python3
"""BASELINE.json configs[4] on one GPU (GPU box): round sizes 100 / 1 000 / 10 000 hypotheses, this library against the reference's own
GPU-HC++ kernel (oracle/_ref/libref_gpuhc.so, unmodified sources for sm_100a) on the same inputs; CUDA events, launch -> sync.
TEST INFRASTRUCTURE (drives oracle/_ref).   python tools/sweep_vs_ref_gpu.py [100,1000,10000]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle.pyoracle import ReferenceGPU
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc

prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
hyps = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "100,1000,10000").split(",")]
print("| hypotheses | this library (ms) | hyp/s | reference GPU-HC++ kernel (ms) | hyp/s | ratio | converged ours / reference |")
print("|---:|---:|---:|---:|---:|---:|---|")
for H in hyps:
    picked = hc.sample_hypotheses(0, H, rs["locations"].shape[0])
    tgt, dif = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
    trk = hc.Tracker(problem=prob)
    trk.upload_params(tgt, dif)
    ts = []
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); trk.track(H, prune=True); b.record(); b.synchronize()
        if i:
            ts.append(a.elapsed_time(b))
    cv = trk.results(H)[1]
    ours = min(ts)
    del trk
    ref = ReferenceGPU(prob)
    ref.setup(tgt, dif, rs["locations"], rs["K"])
    tr = []
    for i in range(2 if H >= 10000 else 3):
        ref.reload()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ref.track(); b.record(); b.synchronize()
        if i:
            tr.append(a.elapsed_time(b))
    cv_r = ref.results()[1]
    r = min(tr)
    del ref
    torch.cuda.empty_cache()
    print("| %d | %.2f | %.0f | %.1f | %.1f | %.1fx | %d / %d |" % (H, ours, H / ours * 1e3, r, H / r * 1e3, r / ours, cv.sum(), cv_r.sum()), flush=True)
