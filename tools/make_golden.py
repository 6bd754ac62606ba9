#!/usr/bin/env python3
"""Generate the committed golden vectors under tests/golden/ (run in the build container, after `make ref oracle`).

  ref_cpuhc_seed0_first6.npz     UNMODIFIED reference CPU-HC (oracle/_ref/libref_cpuhc.so): hypotheses 0..5 of the default
                                 run (seed 0, dataset 000): target params, per-path flags and end points
  ref_cpuhc_seed0_h100.npz       same binary, the full default run: per-path flags (bit-packed) + per-hypothesis counts
  ref_cpuhc_pruned_seed0_h100.npz  the reference CPU-HC with the GPU kernels' path pruning patched in (oracle/_ref/libref_cpuhc_pruned.so,
                                 SURVEY.md App. D.2): per-path flags of the default run — the reference's own CPU arithmetic under the
                                 GPU kernels' control flow, one arm of the parity envelope
  ref_util_support.npz           the reference's MVG helpers (util.hpp:29-209, through oracle/_ref/libref_cpuhc.so): candidate gate, inlier
                                 counts and normalised pose of every converged end point of the default round (oracle tracks, pruning on),
                                 plus value-level (rho, reprojection error) vectors — pins hcb200_score_tracks / host/mvg.hpp to the reference
  ref_eval_vectors.npz           the reference's own evaluators / LAPACK cgesv on fixed inputs (Hx, H, Ht, solve)
  oracle_seed0_h100_{prune,noprune}.npz
                                 oracle/hc_oracle.c on the full default run: flags, counters, per-hypothesis counts and a
                                 SHA-256 of every hypothesis' end points (so the GPU can be checked bit for bit at full size
                                 without re-running the oracle on the GPU box)
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle, ReferenceCPU  # noqa: E402
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def endpoint_digest(tracks_h):
    """SHA-256 over the float32 bytes of [312][30] complex end points, NaNs canonicalised."""
    a = np.ascontiguousarray(tracks_h[:, :30]).view(np.float32).copy()
    a[np.isnan(a)] = np.float32(np.nan)
    return hashlib.sha256(a.view(np.uint32).tobytes()).hexdigest()


def main():
    prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
    orc = Oracle(prob)
    ref = ReferenceCPU()
    what = sys.argv[1:] or ["ref6", "eval", "oracle", "ref100", "refpruned100", "refutil"]

    with tempfile.TemporaryDirectory() as tmp:
        bindir = fixtures.materialize_tree(tmp, files=[0])
        if "ref6" in what:
            tr, cv, inf, tp, sec = ref.run(bindir, 6, seed=0, dataset_index=0)
            np.savez_compressed(os.path.join(OUT, "ref_cpuhc_seed0_first6.npz"), tracks=tr, converged=cv, infinity=inf, target_params=tp)
            print("ref6: %.1f s, counts" % sec, hc.count_solutions(tr, cv, inf, 6).tolist())
        if "ref100" in what:
            tr, cv, inf, tp, sec = ref.run(bindir, 100, seed=0, dataset_index=0)
            counts = hc.count_solutions(tr, cv, inf, 100)
            np.savez_compressed(os.path.join(OUT, "ref_cpuhc_seed0_h100.npz"), converged_bits=np.packbits(cv), infinity_bits=np.packbits(inf),
                                counts=counts.astype(np.int32), seconds=np.float64(sec), cores=np.int32(os.cpu_count()))
            print("ref100: %.1f s, totals conv/inf/real" % sec, counts.sum(0).tolist())

    if "refpruned100" in what:
        with tempfile.TemporaryDirectory() as tmp:
            bindir = fixtures.materialize_tree(tmp, files=[0])
            tr, cv, inf, tp, sec = ReferenceCPU(pruned=True).run(bindir, 100, seed=0, dataset_index=0)
            counts = hc.count_solutions(tr, cv, inf, 100)
            real = ((cv != 0) & np.all(np.abs(tr[:, :30].imag).astype(np.float64) <= 1e-4, axis=1)).astype(np.uint8)
            np.savez_compressed(os.path.join(OUT, "ref_cpuhc_pruned_seed0_h100.npz"), converged_bits=np.packbits(cv), infinity_bits=np.packbits(inf),
                                real_bits=np.packbits(real), counts=counts.astype(np.int32), seconds=np.float64(sec), cores=np.int32(os.cpu_count()))
            print("refpruned100: %.1f s, totals conv/inf/real" % sec, counts.sum(0).tolist())

    if "problem" in what:
        # OTHER minimal problems (problems/<name>/, tools/make_synthetic_problem.py): the reference's own generic CPU-HC and the oracle built
        # for each problem's sizes, on 16 hypotheses
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import make_synthetic_problem as msp
        for name in msp.NAMES:
            msp.select(name)
            pdir = os.path.join(ROOT, "problems", name)
            p2 = msp.load_folder(pdir)
            n = p2["spec"]["n_vars"]
            tgt = msp.target_params(16)
            tr_r, cv_r, inf_r, sec = ref.run_problem(pdir, tgt)
            o2 = Oracle(p2, problem_dir=pdir)
            sp1 = np.concatenate([p2["start_params"], [1.0]]).astype(np.complex64)
            dif = np.empty_like(tgt)
            dif.real, dif.imag = tgt.real - sp1.real, tgt.imag - sp1.imag
            out = dict(target=tgt, ref_tracks=tr_r[:, :n], ref_converged=np.packbits(cv_r), ref_infinity=np.packbits(inf_r))
            for prune in (False, True):
                tr, cv, inf, st = o2.track(tgt, dif, prune)
                k = "prune" if prune else "noprune"
                a = np.ascontiguousarray(tr[:, :n]).view(np.float32).copy()
                a[np.isnan(a)] = np.float32(np.nan)
                out.update({"oracle_converged_" + k: np.packbits(cv), "oracle_infinity_" + k: np.packbits(inf), "oracle_steps_" + k: st[:, 0].astype(np.uint8),
                            "oracle_digest_" + k: np.array(hashlib.sha256(a.view(np.uint32).tobytes()).hexdigest())})
                if not prune:
                    out["oracle_tracks_h0"] = tr[:p2["spec"]["n_tracks"], :n]
                print("problem %s, pruning %s: oracle converged %d infinity %d of %d | reference CPU-HC converged %d infinity %d (%.2f s)"
                      % (name, prune, cv.sum(), inf.sum(), len(cv), cv_r.sum(), inf_r.sum(), sec))
            np.savez_compressed(os.path.join(OUT, "problem_%s_h16.npz" % name), **out)

    if "refutil" in what:
        import ctypes
        from oracle.pyoracle import REF_CPU_SO, c2f
        lib = ctypes.CDLL(REF_CPU_SO)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        tgt, dif, picked = orc.prepare_target_params(0, 100, rs["locations"], rs["tangents"])
        tr, cv, inf, st = orc.track(tgt, dif, True)
        loc = np.ascontiguousarray(rs["locations"], np.float32)
        K = np.ascontiguousarray(rs["K"], np.float32).reshape(-1)
        support = np.full((31200, 2), -1, np.int32)
        poses, cand = [], []
        for pth in np.nonzero(cv)[0]:
            a, b, pose = ctypes.c_int(), ctypes.c_int(), np.zeros(24, np.float32)
            if lib.ref_util_support(vp(c2f(tr[pth])), vp(loc), loc.shape[0], vp(K), ctypes.byref(a), ctypes.byref(b), vp(pose)):
                support[pth] = (a.value, b.value)
                cand.append(pth); poses.append(pose)
        rng = np.random.default_rng(7)
        g1, g2 = rng.normal(0, 0.3, (64, 2)).astype(np.float32), rng.normal(0, 0.3, (64, 2)).astype(np.float32)
        pairs = np.zeros((64, 2), np.float32)
        Rs, Ts = [], []
        for k in range(64):
            pose = poses[k % len(poses)]
            R, T = np.ascontiguousarray(pose[0:9]), np.ascontiguousarray(pose[9:12])
            lib.ref_util_pair(vp(g1[k]), vp(g2[k]), vp(R), vp(T), vp(K), vp(pairs[k]))
            Rs.append(R.copy()); Ts.append(T.copy())
        np.savez_compressed(os.path.join(OUT, "ref_util_support.npz"), support=support, candidates=np.array(cand, np.int32), poses=np.array(poses),
                            converged_bits=np.packbits(cv), g1=g1, g2=g2, R=np.array(Rs), T=np.array(Ts), K=K, rho_err=pairs)
        best = max(cand, key=lambda q: (min(support[q]), -q))
        print("refutil: %d candidates among %d converged paths; best path %d support %s" % (len(cand), int(cv.sum()), best, support[best].tolist()))

    if "eval" in what:
        rng = np.random.default_rng(12345)
        xs, ps, dps, Hx, H, Ht, sol = [], [], [], [], [], [], []
        tgt, dif, _ = orc.prepare_target_params(0, 3, rs["locations"], rs["tangents"])
        for k, (si, t) in enumerate([(0, 0.0), (5, 0.3), (104, 0.77), (311, 1.0)]):
            x = np.concatenate([prob["start_sols"][si] * (1 + 0.05 * (rng.normal(size=30) + 1j * rng.normal(size=30))), [1]]).astype(np.complex64)
            p = orc.param_homotopy(t, tgt[k % 3])
            dp = dif[k % 3]
            A = ref.eval_Hx(orc.hx, x, p)
            b = ref.eval_H(orc.ht, x, p)
            bt = ref.eval_Ht(orc.ht, x, p, dp)
            s, info = ref.cgesv(A, bt)
            xs.append(x); ps.append(p); dps.append(dp); Hx.append(A); H.append(b); Ht.append(bt); sol.append(s)
        np.savez_compressed(os.path.join(OUT, "ref_eval_vectors.npz"), x=np.array(xs), p=np.array(ps), dp=np.array(dps),
                            Hx=np.array(Hx), H=np.array(H), Ht=np.array(Ht), solve=np.array(sol))
        print("eval vectors written")

    if "oracle" in what:
        tgt, dif, picked = orc.prepare_target_params(0, 100, rs["locations"], rs["tangents"])
        for prune in (True, False):
            tr, cv, inf, st = orc.track(tgt, dif, prune)
            counts = hc.count_solutions(tr, cv, inf, 100)
            digests = np.array([endpoint_digest(tr[h * 312:(h + 1) * 312]) for h in range(100)])
            name = "oracle_seed0_h100_%s.npz" % ("prune" if prune else "noprune")
            np.savez_compressed(os.path.join(OUT, name), converged_bits=np.packbits(cv), infinity_bits=np.packbits(inf),
                                counts=counts.astype(np.int32), stats_sum=st.sum(0).astype(np.int64), steps=st[:, 0].astype(np.uint8),
                                end_reason=st[:, 4].astype(np.uint8), digests=digests, picked=picked,
                                track104_h0=tr[104], gt_real_tracks_h0=np.nonzero(cv[:312] & (np.abs(tr[:312, :30].imag).max(1) <= 1e-4))[0])
            print(name, "totals conv/inf/real", counts.sum(0).tolist(), "stage sums", st.sum(0).tolist())


if __name__ == "__main__":
    main()
