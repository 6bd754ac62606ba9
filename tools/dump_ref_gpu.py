#!/usr/bin/env python3
"""Run the UNMODIFIED reference GPU-HC++ kernels (oracle/_ref/libref_gpuhc.so) and this library on the same inputs (dataset 000,
sampler seed 0) and save the per-path results, so that the per-path parity analysis (tools/parity_envelope.py) can be
developed and re-run without a GPU.  TEST INFRASTRUCTURE (drives oracle/_ref).  GPU box only.
    python tools/dump_ref_gpu.py gpurun_out/refgpu 100 1000"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import ReferenceGPU, REF_GPU_DEBUG_SO
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc

prefix = sys.argv[1]
prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
for H in [int(a) for a in sys.argv[2:]] or [100]:
    picked = hc.sample_hypotheses(0, H, rs["locations"].shape[0])
    target, diff = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
    dbg = None
    if os.path.exists(REF_GPU_DEBUG_SO):      # the reference's own GPU_DEBUG output: (t0, delta_t) of every path that did not converge
        rd = ReferenceGPU(prob, debug=True)
        rd.setup(target, diff, rs["locations"], rs["K"])
        rd.track()
        dbg = rd.debug_t0_dt()
        del rd
    ref = ReferenceGPU(prob)
    ref.setup(target, diff, rs["locations"], rs["K"])
    runs = []
    for rep in range(3):                      # is the reference deterministic on this GPU?
        ref.reload()
        ref.track()
        runs.append(ref.results())
    tr_r, cv_r, inf_r = runs[0]
    det = all(np.array_equal(cv_r, r[1]) and np.array_equal(inf_r, r[2]) and
              np.array_equal(np.nan_to_num(tr_r.view(np.float32)), np.nan_to_num(r[0].view(np.float32))) for r in runs[1:])
    trk = hc.Tracker(problem=prob, stats=True)
    trk.upload_params(target, diff)
    trk.track(H, prune=True)
    tr, cv, inf, st = trk.results(H)
    out = dict(picked=picked, ref_conv=np.packbits(cv_r), ref_inf=np.packbits(inf_r), our_conv=np.packbits(cv), our_inf=np.packbits(inf),
               our_stats=st.astype(np.int32), ref_deterministic=np.array([det]))
    if dbg is not None:
        out["ref_debug_t0_dt"] = dbg.astype(np.float32)
    if H <= 100:
        out.update(ref_tracks=tr_r[:, :30], our_tracks=tr[:, :30])
    else:        # end points only where either side converged, as float32 pairs (keeps the file small)
        idx = np.nonzero((cv_r | cv) != 0)[0].astype(np.int32)
        out.update(conv_idx=idx, ref_tracks=tr_r[idx, :30], our_tracks=tr[idx, :30])
    np.savez_compressed("%s_seed0_h%d.npz" % (prefix, H), **out)
    print("H=%d: reference deterministic over 3 launches: %s; converged ours %d ref %d; infinity ours %d ref %d; flags equal: conv %.4f inf %.4f"
          % (H, det, cv.sum(), cv_r.sum(), inf.sum(), inf_r.sum(), (cv == cv_r).mean(), (inf == inf_r).mean()), flush=True)
