#!/usr/bin/env python3
"""Per-path parity report (CPU only; TEST INFRASTRUCTURE): writes profiles/parity_envelope_r2.md from

* tests/golden/envelope_seed0_h{100,1000}_{prune,noprune}.npz  — the noise envelope (tools/parity_envelope.py),
* tests/golden/ref_gpuhc_seed0_h{100,1000}.npz                 — flags of the UNMODIFIED reference GPU-HC++ kernels run on a B200
                                                                  (made here from the dumps of tools/dump_ref_gpu.py: --import-dumps),
* tests/golden/ref_cpuhc_seed0_h100.npz, ref_cpuhc_pruned_seed0_h100.npz — flags of the reference CPU-HC.

    python tools/parity_envelope_report.py --import-dumps gpurun_out/refgpu2     # once, after a GPU run
    python tools/parity_envelope_report.py                                       # the report
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
TR = 312


def bits(a, n):
    return np.unpackbits(a)[:n].astype(bool)


def real_flags(tracks, conv):
    return conv & np.all(np.abs(tracks.imag).astype(np.float64) <= 1e-4, axis=1)


def import_dumps(prefix):
    """gpurun_out dump (per-path end points, 10-100 MB) -> compact golden with the reference GPU kernels' flags only."""
    for H in (100, 1000):
        src = "%s_seed0_h%d.npz" % (prefix, H)
        if not os.path.exists(src):
            print("missing", src)
            continue
        z = np.load(src)
        P = H * TR
        cv_r, cv_o = bits(z["ref_conv"], P), bits(z["our_conv"], P)
        if "conv_idx" in z.files:
            idx = z["conv_idx"]
            real_r, real_o = np.zeros(P, bool), np.zeros(P, bool)
            real_r[idx] = real_flags(z["ref_tracks"], cv_r[idx])
            real_o[idx] = real_flags(z["our_tracks"], cv_o[idx])
        else:
            real_r, real_o = real_flags(z["ref_tracks"], cv_r), real_flags(z["our_tracks"], cv_o)
        out = dict(picked=z["picked"], converged_bits=z["ref_conv"], infinity_bits=z["ref_inf"], real_bits=np.packbits(real_r),
                   our_converged_bits=z["our_conv"], our_infinity_bits=z["our_inf"], our_real_bits=np.packbits(real_o),
                   deterministic=z["ref_deterministic"])
        # the reference's own GPU_DEBUG record (t0, delta_t) and this library's step counters, kept only where the two differ
        inf_r, inf_o = bits(z["ref_inf"], P), bits(z["our_inf"], P)
        d = np.nonzero((cv_r != cv_o) | (inf_r != inf_o) | (real_r != real_o))[0].astype(np.int32)
        out["diff_idx"] = d
        if "ref_debug_t0_dt" in z.files:
            out["diff_ref_t0_dt"] = z["ref_debug_t0_dt"][d]
        out["diff_our_stats"] = z["our_stats"][d]
        dst = os.path.join(GOLD, "ref_gpuhc_seed0_h%d.npz" % H)
        np.savez_compressed(dst, **out)
        print("wrote", dst, os.path.getsize(dst), "bytes;", len(d), "differing paths")


def compare(e, ref, P, with_real=True, ours=None):
    unstable = bits(e["unstable"], P)
    oc, oi, orl = bits(e["spec_conv"], P), bits(e["spec_inf"], P), bits(e["spec_real"], P)
    rc, ri = bits(ref["converged_bits"], P), bits(ref["infinity_bits"], P)
    diff = (oc != rc) | (oi != ri)
    rr = None
    if with_real and "real_bits" in ref.files:
        rr = bits(ref["real_bits"], P)
        diff |= orl != rr
    return dict(unstable=unstable, diff=diff, oc=oc, oi=oi, orl=orl, rc=rc, ri=ri, rr=rr)


def section(w, title, e, ref, P, extra=None):
    H = P // TR
    c = compare(e, ref, P)
    u, d = c["unstable"], c["diff"]
    outside = np.nonzero(d & ~u)[0]
    stable = ~u
    w("### %s" % title)
    w("")
    w("| | paths | share |")
    w("|---|---:|---:|")
    w("| all paths | %d | |" % P)
    w("| unstable (flags differ between two arithmetic variants) | %d | %.2f %% |" % (u.sum(), 100.0 * u.mean()))
    w("| differ from the reference (converged, infinity or real flag) | %d | %.3f %% |" % (d.sum(), 100.0 * d.mean()))
    w("| … of these inside the unstable set | %d | %.1f %% of the differences |" % ((d & u).sum(), 100.0 * (d & u).sum() / max(1, d.sum())))
    w("| … outside (stragglers) | %d | agreement on stable paths %.5f |" % (len(outside), 1.0 - len(outside) / float(stable.sum())))
    w("")
    # which flags
    w("Differences by flag: converged %d, infinity %d%s.  Totals ours / reference: converged %d / %d, infinity %d / %d%s."
      % ((c["oc"] != c["rc"]).sum(), (c["oi"] != c["ri"]).sum(), "" if c["rr"] is None else ", real %d" % (c["orl"] != c["rr"]).sum(),
         c["oc"].sum(), c["rc"].sum(), c["oi"].sum(), c["ri"].sum(), "" if c["rr"] is None else ", real %d / %d" % (c["orl"].sum(), c["rr"].sum())))
    w("")
    # per-hypothesis deltas, all paths and stable paths only
    def per_h(a, mask):
        return (a & mask).reshape(H, TR).sum(1).astype(int)
    allm = np.ones(P, bool)
    rows = []
    for name, a, b in (("converged", c["oc"], c["rc"]), ("infinity", c["oi"], c["ri"])) + ((("real", c["orl"], c["rr"]),) if c["rr"] is not None else ()):
        da, ds = per_h(a, allm) - per_h(b, allm), per_h(a, stable) - per_h(b, stable)
        rows.append((name, (da != 0).sum(), np.abs(da).max(), int(da.sum()), (ds != 0).sum(), np.abs(ds).max() if len(ds) else 0, int(ds.sum())))
    w("Per-hypothesis count deltas (ours − reference), over %d hypotheses:" % H)
    w("")
    w("| flag | hyps with Δ≠0, all paths | max abs Δ | ΣΔ | hyps with Δ≠0, stable paths only | max abs Δ | ΣΔ |")
    w("|---|---:|---:|---:|---:|---:|---:|")
    for r in rows:
        w("| %s | %d | %d | %+d | %d | %d | %+d |" % r)
    w("")
    if len(outside):
        w("Stragglers (differ from the reference, flipped in none of the %d variants):" % len(e["names"]))
        w("")
        have_dbg = extra is not None and "diff_idx" in extra.files
        w("| path | hypothesis | track | ours conv/inf | reference conv/inf |" + (" ours: steps / rejected / end | reference GPU_DEBUG (t0, Δt) of a path that did not converge |" if have_dbg else ""))
        w("|---:|---:|---:|---|---|" + ("---|---|" if have_dbg else ""))
        for p in outside[:40]:
            line = "| %d | %d | %d | %d/%d | %d/%d |" % (p, p // TR, p % TR, c["oc"][p], c["oi"][p], c["rc"][p], c["ri"][p])
            if have_dbg:
                k = np.nonzero(extra["diff_idx"] == p)[0]
                if len(k):
                    st = extra["diff_our_stats"][k[0]]
                    line += " %d / %d / %s |" % (int(st[0]), int(st[3]) & 0xffff, ("converged", "infinity", "pruned", "step cap", "skipped")[int(st[3]) >> 16])
                    line += (" (%.6f, %.3g) |" % tuple(extra["diff_ref_t0_dt"][k[0]])) if "diff_ref_t0_dt" in extra.files else " |"
                else:
                    line += " | |"
            w(line)
        if len(outside) > 40:
            w("| … %d more | | | | |" % (len(outside) - 40))
        w("")
    return dict(n_diff=int(d.sum()), outside=len(outside), unstable=int(u.sum()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--import-dumps", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "parity_envelope_r2.md"))
    a = ap.parse_args()
    if a.import_dumps:
        import_dumps(a.import_dumps)
        return
    L = []
    w = L.append
    w("# Per-path parity against the reference (round 2)")
    w("")
    w("Generated by `tools/parity_envelope_report.py` from the committed goldens under `tests/golden/`; gated by `tests/test_parity_envelope.py`.")
    w("")
    w("**Statement.**  This library's flags (== the oracle spec, bit for bit: `tests/test_gpu_full.py`) are compared, path by path, with the")
    w("reference's (GPU-HC++ kernels run unmodified on the same B200; CPU-HC with LAPACK `cgesv`; CPU-HC with the GPU kernels' pruning patched in).")
    w("A path is *unstable* when its flags differ between any two of the CPU arithmetic variants of `tools/parity_envelope.py` — a set computed")
    w("without looking at any reference result.  Every difference from the reference should be an unstable path; stable paths should agree 100 %.")
    w("The handful of *stragglers* below are rarely-flipping paths that a finite number of variants has not caught yet (the unstable set still")
    w("grows by ~1 path per extra stochastic seed, see the growth table) — each is listed with both sides' record of where the path stopped.")
    w("")
    summary = []
    for H, prune in ((100, "prune"), (100, "noprune"), (1000, "prune")):
        f = os.path.join(GOLD, "envelope_seed0_h%d_%s.npz" % (H, prune))
        if not os.path.exists(f):
            continue
        e = np.load(f)
        w("## Envelope, %d hypotheses, pruning %s: %d variants, %d unstable paths (%.2f %%)" % (H, "on" if prune == "prune" else "off", len(e["names"]), bits(e["unstable"], H * TR).sum(), 100.0 * bits(e["unstable"], H * TR).mean()))
        w("")
        w("| variant | converged | infinity | real | conv flips vs spec | inf flips | real flips | unstable set after this variant |")
        w("|---|---:|---:|---:|---:|---:|---:|---:|")
        names = [str(n) for n in e["names"]]
        seeds = [k for k, n in enumerate(names) if n.startswith("ulp perturbation")]
        for k, n in enumerate(names):
            if k in seeds[3:]:
                continue
            w("| %s | %d | %d | %d | %d | %d | %d | %d |" % ((n,) + tuple(e["variant_counts"][k]) + tuple(e["variant_flips"][k]) + (e["growth"][k],)))
        vf = e["variant_flips"][seeds]
        vc = e["variant_counts"][seeds]
        if len(seeds) > 3:
            w("| ulp perturbation seeds %d..%d (min–max) | %d–%d | %d–%d | %d–%d | %d–%d | %d–%d | %d–%d | %d |" % (
                4, len(seeds), vc[:, 0].min(), vc[:, 0].max(), vc[:, 1].min(), vc[:, 1].max(), vc[:, 2].min(), vc[:, 2].max(),
                vf[:, 0].min(), vf[:, 0].max(), vf[:, 1].min(), vf[:, 1].max(), vf[:, 2].min(), vf[:, 2].max(), e["growth"][-1]))
        w("")
        g = e["growth"]
        tail = [int(g[k] - g[k - 1]) for k in range(max(1, len(g) - 10), len(g))]
        w("Growth of the unstable set over the last variants: +%s paths — the tail of rarely-flipping paths is not exhausted, which is" % ", +".join(str(t) for t in tail))
        w("where the stragglers come from.  Flip-count histogram of the unstable paths (in how many of the %d variants a path differs from the spec):" % (len(names) - 1))
        fl = e["flips"][e["flips"] > 0]
        hist = [(1, 1), (2, 3), (4, 9), (10, 29), (30, 10 ** 6)]
        w("")
        w("| flips | " + " | ".join("%d" % lo if lo == hi else ("%d–%d" % (lo, hi) if hi < 10 ** 6 else "≥%d" % lo) for lo, hi in hist) + " |")
        w("|---|" + "---:|" * len(hist))
        w("| paths | " + " | ".join(str(int(((fl >= lo) & (fl <= hi)).sum())) for lo, hi in hist) + " |")
        w("")
    w("## Differences from the reference")
    w("")
    for H in (100, 1000):
        fe, fr = os.path.join(GOLD, "envelope_seed0_h%d_prune.npz" % H), os.path.join(GOLD, "ref_gpuhc_seed0_h%d.npz" % H)
        if os.path.exists(fe) and os.path.exists(fr):
            e, r = np.load(fe), np.load(fr)
            assert np.array_equal(e["picked"], r["picked"])
            assert np.array_equal(e["spec_conv"], r["our_converged_bits"]) and np.array_equal(e["spec_inf"], r["our_infinity_bits"]), "GPU run != spec"
            s = section(w, "Reference GPU-HC++ kernels on a B200 (unmodified, deterministic over 3 launches: %s), %d hypotheses, %d variants" % (bool(r["deterministic"][0]), H, len(e["names"])), e, r, H * TR, extra=r)
            summary.append(("reference GPU kernels, H=%d" % H, s))
    fe = os.path.join(GOLD, "envelope_seed0_h100_prune.npz")
    fr = os.path.join(GOLD, "ref_cpuhc_pruned_seed0_h100.npz")
    if os.path.exists(fe) and os.path.exists(fr):
        summary.append(("reference CPU-HC + pruning, H=100", section(w, "Reference CPU-HC with the GPU kernels' pruning patched in (oracle/ref_build/make_pruned_cpuhc.py), 100 hypotheses", np.load(fe), np.load(fr), 31200)))
    fe = os.path.join(GOLD, "envelope_seed0_h100_noprune.npz")
    fr = os.path.join(GOLD, "ref_cpuhc_seed0_h100.npz")
    if os.path.exists(fe) and os.path.exists(fr):
        summary.append(("reference CPU-HC (unmodified), pruning off, H=100", section(w, "Unmodified reference CPU-HC (LAPACK cgesv), pruning off, 100 hypotheses", np.load(fe), np.load(fr), 31200)))
    w("## Summary")
    w("")
    w("| comparison | differing paths | inside the unstable set | stragglers | unstable set |")
    w("|---|---:|---:|---:|---:|")
    for name, s in summary:
        w("| %s | %d | %d | %d | %d |" % (name, s["n_diff"], s["n_diff"] - s["outside"], s["outside"], s["unstable"]))
    w("")
    with open(a.out, "w") as f:
        f.write("\n".join(L) + "\n")
    print("wrote", a.out)
    for name, s in summary:
        print("  %-55s diff %5d outside %4d unstable %6d" % (name, s["n_diff"], s["outside"], s["unstable"]))


if __name__ == "__main__":
    main()
