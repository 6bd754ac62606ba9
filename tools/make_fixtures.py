#!/usr/bin/env python3
"""Pack the reference's problem definition and RANSAC dataset into compact binary fixtures.

Run in the build container (where /root/reference exists):

    python tools/make_fixtures.py [--ref /root/reference] [--files 0 1 2 3]

Writes (committed, small):
    <pkg>/data/problem_trifocal_2op1p_30x30.npz   start sols/params, index tables, yaml text
    <pkg>/data/ransac_synthetic_XXX.npz            edgel triplets + GT poses + K of dataset file XXX

Every decimal token is converted straight to float32 with libc `strtof`, i.e. exactly what the
reference's `std::istream >> float` does (Data_Reader.cpp:37-60, 86-121, 273-338), so a fixture
materialised back to text with 9 significant digits parses to bit-identical floats.
The text formats themselves are documented in SURVEY.md App. A.3; `fixtures.py` in the package
re-creates a reference-layout directory tree from these files.
"""
import argparse
import ctypes
import os
import sys

import numpy as np

PKG = "trifocal_pose_estimation_using_improved_gpuhc_b200"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_libc = ctypes.CDLL(None)
_libc.strtof.restype = ctypes.c_float
_libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]


def read_floats(path):
    with open(path, "rb") as f:
        toks = f.read().split()
    return np.array([_libc.strtof(t, None) for t in toks], dtype=np.float32)


def read_ints(path):
    with open(path, "rb") as f:
        return np.array([int(t) for t in f.read().split()], dtype=np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--files", type=int, nargs="*", default=[0, 1, 2, 3])
    args = ap.parse_args()

    prob = os.path.join(args.ref, "problems", "trifocal_2op1p_30x30")
    out_dir = os.path.join(ROOT, PKG, "data")
    os.makedirs(out_dir, exist_ok=True)

    ss = read_floats(os.path.join(prob, "start_sols.txt")).reshape(312, 30, 2)
    sp = read_floats(os.path.join(prob, "start_params.txt")).reshape(33, 2)
    tp = read_floats(os.path.join(prob, "target_params.txt")).reshape(-1, 2)
    hx = read_ints(os.path.join(prob, "dHdx_indx.txt"))
    ht = read_ints(os.path.join(prob, "dHdt_indx.txt"))
    assert hx.size == 30 * 30 * 8 * 5 and ht.size == 30 * 16 * 6
    assert hx.min() >= -128 and hx.max() <= 127 and ht.min() >= -128 and ht.max() <= 127
    with open(os.path.join(prob, "gpuhc_settings.yaml")) as f:
        yaml_text = f.read()
    np.savez_compressed(
        os.path.join(out_dir, "problem_trifocal_2op1p_30x30.npz"),
        start_sols=ss, start_params=sp, target_params_file=tp,
        dHdx_indx=hx.astype(np.int8), dHdt_indx=ht.astype(np.int8),
        settings_yaml=np.frombuffer(yaml_text.encode(), dtype=np.uint8))

    ds = os.path.join(args.ref, "RANSAC_Data", "trifocal_2op1p_30x30", "Synthetic")
    K = read_floats(os.path.join(ds, "Intrinsic_Matrix.txt")).reshape(3, 3)
    for i in args.files:
        e = read_floats(os.path.join(ds, "Triplet_Edgels", "Triplet_Edgels_%03d.txt" % i)).reshape(-1, 12)
        p21 = read_floats(os.path.join(ds, "GT_Poses21", "GT_Poses21_%03d.txt" % i)).reshape(4, 3)
        p31 = read_floats(os.path.join(ds, "GT_Poses31", "GT_Poses31_%03d.txt" % i)).reshape(4, 3)
        np.savez_compressed(os.path.join(out_dir, "ransac_synthetic_%03d.npz" % i),
                            triplet_edgels=e, gt_pose21=p21, gt_pose31=p31, K=K)
        print("dataset %03d: %d edgel triplets" % (i, e.shape[0]))
    print("fixtures written to", out_dir)


if __name__ == "__main__":
    sys.exit(main())
