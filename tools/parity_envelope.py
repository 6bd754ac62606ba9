#!/usr/bin/env python3
"""Noise envelope of the tracker's integer results (CPU only; TEST INFRASTRUCTURE: drives oracle/).

Every rounding decision moves a few knife-edge paths across `converged` / `infinity` (the reference differs from itself the same
way: CPU-HC vs GPU-HC++, LAPACK build vs LAPACK build).  This tool tracks one round under MANY arithmetic variants of the same
algorithm — the spec, the reference's own choices at each point where the spec departs from it (literal LU + back substitution
with and without FMA contraction, exact-maximum pivot rule, cuCdivf reciprocal, left-to-right term products, sequential norm
sums, the round-1 RK constant) and K stochastic-arithmetic seeds (every linear-solve result moved by -1/0/+1 ulp) — and marks
every path whose flags differ between any two variants as UNSTABLE.  tests/test_parity_envelope.py then requires that every
path on which this library differs from the reference GPU-HC++ kernels is in that set, i.e. stable paths agree 100 %.

    python tools/parity_envelope.py --hyp 100 --seeds 64 --out tests/golden/envelope_seed0_h100_prune.npz [--no-prune]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc

STRUCTURED = [
    ("spec", {}),
    ("rk_final_mul (round-1 spec)", dict(rk_final_mul=1)),
    ("literal reference LU", dict(solver=1)),
    ("literal reference LU, FMA-contracted", dict(solver=2)),
    ("exact-maximum pivot rule", dict(solver=3)),
    ("cuCdivf reciprocal", dict(solver=4)),
    ("exact pivot + cuCdivf", dict(solver=5)),
    ("left-to-right term products", dict(term_order=1)),
    ("left-to-right, FMA-contracted", dict(term_order=1, contract=1)),
    ("reference-GPU-like (LU + left-to-right, contracted)", dict(solver=2, term_order=1, contract=1)),
    ("sequential norm sums (reference CPU)", dict(sum_order=1)),
]


def merge(base_path, ext_path, ext_raw_path, out):
    """Union of an envelope file and an extension run (--structured 0 --seed-start K --raw …) of the same round: variant 0 of the extension
    is the spec again and is dropped; the growth curve continues exactly (from the extension's raw flags)."""
    b, e, r = np.load(base_path), np.load(ext_path), np.load(ext_raw_path)
    assert np.array_equal(b["picked"], e["picked"]) and bool(b["prune"][0]) == bool(e["prune"][0])
    assert np.array_equal(b["spec_conv"], e["spec_conv"]) and np.array_equal(b["spec_inf"], e["spec_inf"]) and np.array_equal(b["spec_real"], e["spec_real"])
    P = len(b["flips"])
    unb = lambda a: np.unpackbits(a, axis=-1)[..., :P].astype(bool)
    conv, inf, real = unb(r["conv"]), unb(r["inf"]), unb(r["real"])
    acc = unb(b["unstable"]).copy()
    growth = list(b["growth"])
    for k in range(1, conv.shape[0]):
        acc |= (conv[k] != conv[0]) | (inf[k] != inf[0]) | (real[k] != real[0])
        growth.append(int(acc.sum()))
    unstable = unb(b["unstable"]) | unb(e["unstable"])
    assert np.array_equal(acc, unstable)
    cat = lambda k: np.concatenate([b[k], e[k][1:]])
    np.savez_compressed(out, names=cat("names"), picked=b["picked"], prune=b["prune"], spec_conv=b["spec_conv"], spec_inf=b["spec_inf"], spec_real=b["spec_real"],
                        unstable=np.packbits(unstable), flips=(b["flips"].astype(np.uint32) + e["flips"]).astype(np.uint16),
                        head_conv=cat("head_conv"), head_inf=cat("head_inf"), variant_fields=cat("variant_fields"),
                        seq_unstable=np.packbits(unb(b["seq_unstable"]) | unb(e["seq_unstable"])), growth=np.array(growth, np.int32),
                        variant_counts=cat("variant_counts"), variant_flips=cat("variant_flips"))
    print("merged %d + %d variants -> %s: unstable %d -> %d of %d paths" % (len(b["names"]), len(e["names"]) - 1, out, unb(b["unstable"]).sum(), unstable.sum(), P))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--merge":         # --merge BASE EXT EXT_RAW OUT
        merge(*sys.argv[2:6])
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--hyp", type=int, default=100)
    ap.add_argument("--seeds", type=int, default=64, help="number of stochastic-arithmetic variants")
    ap.add_argument("--seed-start", type=int, default=1, help="first stochastic-arithmetic seed (to extend an earlier run)")
    ap.add_argument("--structured", type=int, default=1, help="0: skip the structured variants except the spec (extension runs)")
    ap.add_argument("--sampler-seed", type=int, default=0)
    ap.add_argument("--dataset", type=int, default=0)
    ap.add_argument("--no-prune", action="store_true")
    ap.add_argument("--out", required=True)
    ap.add_argument("--raw", default=None, help="also save every variant's flags / step counts here (large)")
    a = ap.parse_args()
    prob, rs = fixtures.load_problem(), fixtures.load_ransac(a.dataset)
    orc = Oracle(prob)
    H, P = a.hyp, a.hyp * 312
    picked = hc.sample_hypotheses(a.sampler_seed, H, rs["locations"].shape[0])
    target, diff = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])
    variants = (list(STRUCTURED) if a.structured else STRUCTURED[:1]) + \
        [("ulp perturbation seed %d" % s, dict(perturb_seed=s)) for s in range(a.seed_start, a.seed_start + a.seeds)]
    conv, inf, real, steps = [], [], [], []
    t0 = time.time()
    for name, v in variants:
        tr, cv, fl, st = orc.track(target, diff, prune=not a.no_prune, variant=v or None)
        # "real" as Evaluations::Evaluate_HC_Sols counts it (Evaluations.cpp:145-167): converged and every |imag| <= 1e-4
        rl = ((cv != 0) & np.all(np.abs(tr[:, :30].imag).astype(np.float64) <= 1e-4, axis=1)).astype(np.uint8)
        conv.append(cv); inf.append(fl); real.append(rl); steps.append(st[:, 0].astype(np.uint8))
        print("%-52s conv %6d inf %6d real %5d | vs spec: conv %4d inf %4d real %3d flips  [%.0f s]"
              % (name, cv.sum(), fl.sum(), rl.sum(), (cv != conv[0]).sum(), (fl != inf[0]).sum(), (rl != real[0]).sum(), time.time() - t0), flush=True)
    conv, inf, real, steps = np.stack(conv), np.stack(inf), np.stack(real), np.stack(steps)
    flips = ((conv != conv[0]) | (inf != inf[0]) | (real != real[0])).sum(0).astype(np.uint16)   # in how many variants a path's flags differ from the spec's
    unstable = flips > 0
    seq_unstable = (steps != steps[0]).any(0)                                         # weaker: the step count differs somewhere
    # growth of the unstable set with the number of variants (is it saturating?)
    growth = [int((((conv[:k] != conv[0]) | (inf[:k] != inf[0]) | (real[:k] != real[0])).any(0)).sum()) for k in range(1, len(variants) + 1)]
    print("unstable paths: %d of %d (%.2f %%); step-count-unstable: %d (%.2f %%)" % (unstable.sum(), P, 100.0 * unstable.mean(), seq_unstable.sum(), 100.0 * seq_unstable.mean()))
    print("growth of the unstable set:", growth)
    np.savez_compressed(a.out, names=np.array([n for n, _ in variants]), picked=picked, prune=np.array([not a.no_prune]),
                        spec_conv=np.packbits(conv[0]), spec_inf=np.packbits(inf[0]), spec_real=np.packbits(real[0]),
                        unstable=np.packbits(unstable), flips=flips,
                        head_conv=np.packbits(conv[:, :936], axis=1), head_inf=np.packbits(inf[:, :936], axis=1),     # every variant, hypotheses 0..2
                        variant_fields=np.array([repr(sorted(v.items())) for _, v in variants]),
                        seq_unstable=np.packbits(seq_unstable), growth=np.array(growth, np.int32),
                        variant_counts=np.stack([conv.sum(1), inf.sum(1), real.sum(1)], 1).astype(np.int32),
                        variant_flips=np.stack([(conv != conv[0]).sum(1), (inf != inf[0]).sum(1), (real != real[0]).sum(1)], 1).astype(np.int32))
    if a.raw:
        np.savez_compressed(a.raw, conv=np.packbits(conv, axis=1), inf=np.packbits(inf, axis=1), real=np.packbits(real, axis=1), steps=steps)


if __name__ == "__main__":
    main()
