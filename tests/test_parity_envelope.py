"""Per-path parity against the reference: every difference is a borderline path.

The integer results of the tracker (converged / infinity / real flags) are decided by knife-edge floating-point comparisons, and the
reference is not bit-stable against itself (CPU-HC vs GPU-HC++, LAPACK build vs LAPACK build).  Instead of a
statistical gate, this module makes a PER-PATH statement:

* `tools/parity_envelope.py` tracks a round on the CPU under many arithmetic variants of the same algorithm — the spec, the reference's own
  choice at every point where the spec departs from it (literal LU + back substitution with and without FMA contraction, exact-maximum
  pivot rule, cuCdivf reciprocal, left-to-right term products, sequential norm sums, the round-1 RK constant) and dozens of
  stochastic-arithmetic seeds (every linear-solve result moved by -1/0/+1 ulp).  A path whose flags differ between any two variants
  is UNSTABLE; the set is committed under tests/golden/envelope_*.npz.  It is computed WITHOUT looking at any reference result.
* The tests then require that the paths on which this library (== the oracle spec, bit for bit: tests/test_gpu_full.py) differs from
  - the reference GPU-HC++ kernels run live on the same GPU (pruning on, 100 and 1000 hypotheses; early abort),
  - the unmodified reference CPU-HC (pruning off; committed golden of the real thing),
  - the reference CPU-HC with the GPU kernels' pruning patched in (pruning on; committed golden),
  lie in the unstable set, and that STABLE paths agree.  Flag flips have a long tail of rarely-flipping paths, so a finite number of
  variants cannot catch every one: the gates below allow the handful of stragglers that were observed (1.5x head-room) and list them.
  Observed numbers and the per-hypothesis tables: profiles/parity_envelope_r2.md.
"""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _bits(a, n):
    return np.unpackbits(a)[:n].astype(bool)


def _load_env(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(name + " not generated yet (tools/parity_envelope.py)")
    return np.load(path)


def _check(diff, unstable, what, max_outside, min_coverage, min_stable_agreement=0.9998):
    n_diff, outside = int(diff.sum()), np.nonzero(diff & ~unstable)[0]
    stable = ~unstable
    agreement = 1.0 - len(outside) / float(stable.sum())
    msg = "%s: %d differing paths, %d outside the unstable set (%d paths = %.2f %% of all): %s" % (
        what, n_diff, len(outside), int(unstable.sum()), 100.0 * unstable.mean(), outside.tolist())
    assert len(outside) <= max_outside, msg
    assert n_diff == 0 or (n_diff - len(outside)) / n_diff >= min_coverage, msg
    assert agreement >= min_stable_agreement, msg
    return outside


# ---------------------------------------------------------------------------------------------------------------------
# CPU tests: committed goldens only

def test_envelope_goldens_are_consistent_with_the_oracle_goldens():
    """The spec column of every envelope file is the oracle golden of the same round, and the unstable set is what the stored flip
    counts say; the set stays a small minority of the paths."""
    for prune in ("prune", "noprune"):
        e = _load_env("envelope_seed0_h100_%s.npz" % prune)
        g = np.load(os.path.join(GOLD, "oracle_seed0_h100_%s.npz" % prune))
        assert np.array_equal(e["spec_conv"], g["converged_bits"]) and np.array_equal(e["spec_inf"], g["infinity_bits"])
        assert np.array_equal(e["picked"], g["picked"])
        unstable = _bits(e["unstable"], 31200)
        assert np.array_equal(unstable, e["flips"] > 0)
        assert len(e["names"]) >= 60 and unstable.mean() < (0.03 if prune == "prune" else 0.08)     # observed 2.25 % / 6.93 % (235 variants each)
        assert np.all(np.diff(e["growth"]) >= 0) and e["growth"][-1] == unstable.sum()


def test_envelope_head_is_reproduced_by_the_oracle(oracle, ransac0, problem):
    """The committed file was made by THIS oracle: re-track hypotheses 0..2 under a structured variant and two stochastic seeds and compare
    with the stored per-variant flags."""
    import ast
    from trifocal_pose_estimation_using_improved_gpuhc_b200 import hc
    e = _load_env("envelope_seed0_h100_prune.npz")
    target, diff = hc.target_params_from_picks(e["picked"][:3], ransac0["locations"], ransac0["tangents"], problem["start_params"])
    for k in (0, 3, 9, 11, 40):
        fields = dict(ast.literal_eval(str(e["variant_fields"][k])))
        tr, cv, inf, st = oracle.track(target, diff, prune=True, variant=fields or None)
        assert np.array_equal(np.packbits(cv), e["head_conv"][k]) and np.array_equal(np.packbits(inf), e["head_inf"][k]), e["names"][k]


def test_every_deviation_from_the_reference_sits_at_the_same_noise_floor():
    """what does each departure of the spec from the reference cost in agreement?  Nothing measurable: every variant —
    including the ones that restate the reference literally — flips the same ~0.3 % of the paths against the spec as a one-ulp nudge of
    the solve results does; the exact-maximum pivot rule alone changes next to nothing."""
    e = _load_env("envelope_seed0_h100_prune.npz")
    names = [str(n) for n in e["names"]]
    flips = e["variant_flips"][:, 0] + e["variant_flips"][:, 1]          # converged + infinity flips against the spec
    seeds = np.array([f for n, f in zip(names, flips) if n.startswith("ulp perturbation")])
    floor_lo, floor_hi = seeds.min(), seeds.max()
    assert 80 <= floor_lo and floor_hi <= 250                              # ~0.3-0.6 % of 31 200 paths
    for n, f in zip(names, flips):
        if n == "spec" or n.startswith("ulp perturbation"):
            continue
        if "exact-maximum pivot rule" in n or n.startswith("sequential norm sums"):
            assert f <= 10, (n, f)                                         # ties within 2^-18 are rare; sum order changes nothing
        else:
            assert 0.6 * floor_lo <= f <= 1.4 * floor_hi, (n, f, floor_lo, floor_hi)


def test_differences_from_the_reference_cpu_are_unstable_paths_pruning_off():
    """Oracle spec (== the GPU, bit for bit) vs the UNMODIFIED reference CPU-HC with LAPACK cgesv, full default round, no pruning."""
    e = _load_env("envelope_seed0_h100_noprune.npz")
    r = np.load(os.path.join(GOLD, "ref_cpuhc_seed0_h100.npz"))
    unstable = _bits(e["unstable"], 31200)
    diff = (_bits(e["spec_conv"], 31200) != _bits(r["converged_bits"], 31200)) | (_bits(e["spec_inf"], 31200) != _bits(r["infinity_bits"], 31200))
    _check(diff, unstable, "reference CPU-HC, pruning off", max_outside=MAX_OUTSIDE["cpu_noprune"], min_coverage=0.95)


def test_differences_from_the_pruned_reference_cpu_are_unstable_paths():
    """… vs the reference CPU-HC with the GPU kernels' path pruning patched in (oracle/ref_build/make_pruned_cpuhc.py)."""
    e = _load_env("envelope_seed0_h100_prune.npz")
    r = np.load(os.path.join(GOLD, "ref_cpuhc_pruned_seed0_h100.npz"))
    unstable = _bits(e["unstable"], 31200)
    diff = (_bits(e["spec_conv"], 31200) != _bits(r["converged_bits"], 31200)) | (_bits(e["spec_inf"], 31200) != _bits(r["infinity_bits"], 31200)) | \
           (_bits(e["spec_real"], 31200) != _bits(r["real_bits"], 31200))
    _check(diff, unstable, "reference CPU-HC + pruning", max_outside=MAX_OUTSIDE["cpu_prune"], min_coverage=0.95)


@pytest.mark.parametrize("H,key", [(100, "gpu_h100"), (1000, "gpu_h1000")])
def test_differences_from_the_committed_reference_gpu_flags_are_unstable_paths(H, key):
    """… vs the flags the UNMODIFIED reference GPU-HC++ kernels produced on a B200 (tests/golden/ref_gpuhc_*.npz, made by
    tools/dump_ref_gpu.py + tools/parity_envelope_report.py --import-dumps); the `-m gpu` test below repeats this with the kernels run live."""
    e = _load_env("envelope_seed0_h%d_prune.npz" % H)
    r = np.load(os.path.join(GOLD, "ref_gpuhc_seed0_h%d.npz" % H))
    P = H * 312
    assert np.array_equal(e["picked"], r["picked"]) and bool(r["deterministic"][0])
    assert np.array_equal(e["spec_conv"], r["our_converged_bits"]) and np.array_equal(e["spec_inf"], r["our_infinity_bits"])     # the GPU run WAS the spec
    unstable = _bits(e["unstable"], P)
    diff = (_bits(e["spec_conv"], P) != _bits(r["converged_bits"], P)) | (_bits(e["spec_inf"], P) != _bits(r["infinity_bits"], P)) | \
           (_bits(e["spec_real"], P) != _bits(r["real_bits"], P))
    _check(diff, unstable, "reference GPU-HC++ kernels (committed flags), %d hypotheses" % H, max_outside=MAX_OUTSIDE[key],
           min_coverage=MIN_COVERAGE[key], min_stable_agreement=MIN_STABLE[key])


# observed stragglers (paths that differ from a reference implementation but flip in none of the variants), with 1.5x head-room
MAX_OUTSIDE = {"gpu_h100": 2, "gpu_h1000": 20, "cpu_noprune": 3, "cpu_prune": 1}     # observed: 1 (235 variants), 12 (74 variants), 2 (235), 0 (235)
# 1000 hypotheses: the committed envelope holds 74 variants (4-5 CPU-minutes each), so its unstable set is less saturated (1.87 % of the paths
# against 2.25 % with 235 variants at 100 hypotheses): 99.1 % of the 1 367 differences fall inside it (94.4 % with the first 16 variants, 98.0 %
# with 36, 99.0 % with 60), stable paths agree to 0.99996.
MIN_COVERAGE = {"gpu_h100": 0.95, "gpu_h1000": 0.97}
MIN_STABLE = {"gpu_h100": 0.9998, "gpu_h1000": 0.9999}


# ---------------------------------------------------------------------------------------------------------------------
# GPU tests: the reference GPU-HC++ kernels live, on the same GPU and inputs

def _run_both(problem, ransac0, H):
    from oracle.pyoracle import ReferenceGPU, REF_GPU_SO
    from trifocal_pose_estimation_using_improved_gpuhc_b200 import hc
    if not os.path.exists(REF_GPU_SO):
        pytest.skip("oracle/_ref/libref_gpuhc.so not built")
    picked = hc.sample_hypotheses(0, H, ransac0["locations"].shape[0])
    target, diff = hc.target_params_from_picks(picked, ransac0["locations"], ransac0["tangents"], problem["start_params"])
    ref = ReferenceGPU(problem)
    ref.setup(target, diff, ransac0["locations"], ransac0["K"])
    ref.track()
    tr_r, cv_r, inf_r = ref.results()
    trk = hc.Tracker(problem=problem, stats=True)
    trk.set_edgels(ransac0["locations"], ransac0["K"])
    trk.upload_params(target, diff)
    trk.track(H, prune=True)
    tr, cv, inf, st = trk.results(H)
    return picked, (tr, cv, inf, st), (tr_r, cv_r, inf_r), trk, ref


def _real(tr, cv):
    return (cv != 0) & np.all(np.abs(tr[:, :30].imag).astype(np.float64) <= 1e-4, axis=1)


@pytest.mark.gpu
@pytest.mark.parametrize("H,key", [(100, "gpu_h100"), (1000, "gpu_h1000")])
def test_differences_from_the_reference_gpu_kernels_are_unstable_paths(problem, ransac0, H, key):
    e = _load_env("envelope_seed0_h%d_prune.npz" % H)
    P = H * 312
    picked, (tr, cv, inf, st), (tr_r, cv_r, inf_r), trk, ref = _run_both(problem, ransac0, H)
    assert np.array_equal(picked, e["picked"])
    # this library IS the spec column of the envelope (bit-exact oracle parity, here re-checked on the flags)
    assert np.array_equal(np.packbits(cv), e["spec_conv"]) and np.array_equal(np.packbits(inf), e["spec_inf"])
    unstable = _bits(e["unstable"], P)
    diff = (cv != cv_r) | (inf != inf_r) | (_real(tr, cv) != _real(tr_r, cv_r))
    _check(diff, unstable, "reference GPU-HC++ kernels, %d hypotheses" % H, max_outside=MAX_OUTSIDE[key],
           min_coverage=MIN_COVERAGE[key], min_stable_agreement=MIN_STABLE[key])
    # per hypothesis: counts over the STABLE paths are identical in every hypothesis but the stragglers'
    s = ~unstable
    for flag_a, flag_b in ((cv, cv_r), (inf, inf_r), (_real(tr, cv), _real(tr_r, cv_r))):
        a = (flag_a.astype(bool) & s).reshape(H, 312).sum(1)
        b = (flag_b.astype(bool) & s).reshape(H, 312).sum(1)
        assert (a != b).sum() <= MAX_OUTSIDE[key]


@pytest.mark.gpu
def test_early_abort_results_are_the_no_abort_results_of_the_paths_that_ran(problem, ransac0):
    """Abort mode adds no new kind of difference: a path that ran to completion before the flag went up has exactly its no-abort result
    (ours: bit for bit; reference: its own no-abort flags), so per-path parity in abort mode is inherited from the no-abort round."""
    H = 100
    picked, (tr, cv, inf, st), (tr_r, cv_r, inf_r), trk, ref = _run_both(problem, ransac0, H)
    trk.track_abort(H, prune=True)
    tr_a, cv_a, inf_a, st_a = trk.results(H)
    ran = (st_a[:, 3] >> 16) < 4                       # not skipped / cut by the flag
    assert ran.sum() > 100 and (~ran).sum() > 20000
    assert np.array_equal(cv_a[ran], cv[ran]) and np.array_equal(inf_a[ran], inf[ran])
    a, b = np.ascontiguousarray(tr_a[ran][:, :30]), np.ascontiguousarray(tr[ran][:, :30])
    assert bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))
    ref.reload()
    ref.track_abort()
    _, cv_ra, _ = ref.results()
    assert np.all(cv_r[cv_ra != 0] != 0)               # every path the reference reports converged under abort converged without it
    idx = ref.d_found_index.cpu().numpy()
    best = trk.d_best.cpu().numpy()
    assert best[0] == 1 and best[1] == 104 and 104 in idx[idx >= 0]      # both stop on hypothesis 0 / track 104, the ground-truth pose
