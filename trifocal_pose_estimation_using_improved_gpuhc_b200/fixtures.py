"""Problem-definition and RANSAC-dataset fixtures.

The reference reads its inputs from a directory tree of text files (SURVEY.md App. A.3;
reference `magmaHC/Data_Reader.cpp:37-338`).  The GPU box has no `/root/reference`, so the package
ships the same numbers as compact `.npz` fixtures (made by `tools/make_fixtures.py`) and this module

* loads them as numpy arrays (`load_problem`, `load_ransac`), and
* re-creates a reference-layout tree (`materialize_tree`) so the C++ host classes
  (`Data_Reader`, `GPU_HC_Solver`) and the reference build under `oracle/_ref` read the very same
  text formats the reference ships.

Values are float32 exactly as `std::istream >> float` parses the reference's decimals; they are written
back with 9 significant digits, which round-trips float32 bit-exactly.
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

PROBLEM_NAME = "trifocal_2op1p_30x30"
NUM_VARS = 30
NUM_PARAMS = 33
NUM_TRACKS = 312


def load_problem():
    """Return dict with start_sols[312,30] c64, start_params[33] c64, dHdx_indx, dHdt_indx (flat int32)."""
    z = np.load(os.path.join(_DATA, "problem_%s.npz" % PROBLEM_NAME))
    ss = z["start_sols"]
    sp = z["start_params"]
    return {
        "start_sols": (ss[..., 0] + 1j * ss[..., 1]).astype(np.complex64),
        "start_params": (sp[:, 0] + 1j * sp[:, 1]).astype(np.complex64),
        "dHdx_indx": z["dHdx_indx"].astype(np.int32),
        "dHdt_indx": z["dHdt_indx"].astype(np.int32),
        "settings_yaml": bytes(z["settings_yaml"]).decode(),
        "_raw": z,
    }


def load_ransac(index=0):
    """Return dict with triplet_edgels[E,12] f32, locations[E,6], tangents[E,6], gt_pose21/31[4,3], K[3,3]."""
    z = np.load(os.path.join(_DATA, "ransac_synthetic_%03d.npz" % index))
    e = z["triplet_edgels"]
    # Data_Reader.cpp:287-323: line = x1 y1 tx1 ty1 x2 y2 tx2 ty2 x3 y3 tx3 ty3
    loc = np.ascontiguousarray(e[:, [0, 1, 4, 5, 8, 9]])
    tan = np.ascontiguousarray(e[:, [2, 3, 6, 7, 10, 11]])
    return {"triplet_edgels": e, "locations": loc, "tangents": tan,
            "gt_pose21": z["gt_pose21"], "gt_pose31": z["gt_pose31"], "K": z["K"]}


def available_ransac_files():
    out = []
    for f in sorted(os.listdir(_DATA)):
        if f.startswith("ransac_synthetic_") and f.endswith(".npz"):
            out.append(int(f[len("ransac_synthetic_"):-4]))
    return out


def _fmt(v):
    return "%.9g" % float(v)


def _write_rows(path, arr):
    with open(path, "w") as f:
        for row in arr:
            f.write("\t".join(_fmt(v) for v in row) + "\n")


def materialize_tree(root, settings_overrides=None, files=None):
    """Create `<root>/problems/<name>/*`, `<root>/RANSAC_Data/<name>/Synthetic/*`, `<root>/Output_Write_Files/`
    and `<root>/build/bin/` exactly as the reference expects them relative to its working directory
    (`GPU_HC_Solver.cpp:125-127`: "../../problems/", "../../RANSAC_Data/", "../../Output_Write_Files/").
    Returns the path of `<root>/build/bin`."""
    prob = load_problem()
    z = prob["_raw"]
    pdir = os.path.join(root, "problems", PROBLEM_NAME)
    os.makedirs(pdir, exist_ok=True)
    _write_rows(os.path.join(pdir, "start_sols.txt"), z["start_sols"].reshape(-1, 2))
    _write_rows(os.path.join(pdir, "start_params.txt"), z["start_params"])
    _write_rows(os.path.join(pdir, "target_params.txt"), z["target_params_file"])
    for name, key in (("dHdx_indx.txt", "dHdx_indx"), ("dHdt_indx.txt", "dHdt_indx")):
        with open(os.path.join(pdir, name), "w") as f:
            for row in z[key].astype(np.int64).reshape(-1, NUM_VARS):
                f.write("\t".join(str(int(v)) for v in row) + "\t\n")
    text = prob["settings_yaml"]
    if settings_overrides:
        lines = []
        seen = set()
        for ln in text.splitlines():
            key = ln.split(":", 1)[0].strip() if ":" in ln and not ln.lstrip().startswith("#") else None
            if key in settings_overrides:
                lines.append("%s: %s" % (key, settings_overrides[key]))
                seen.add(key)
            else:
                lines.append(ln)
        for k, v in settings_overrides.items():
            if k not in seen:
                lines.append("%s: %s" % (k, v))
        text = "\n".join(lines) + "\n"
    with open(os.path.join(pdir, "gpuhc_settings.yaml"), "w") as f:
        f.write(text)

    ddir = os.path.join(root, "RANSAC_Data", PROBLEM_NAME, "Synthetic")
    for sub in ("Triplet_Edgels", "GT_Poses21", "GT_Poses31"):
        os.makedirs(os.path.join(ddir, sub), exist_ok=True)
    files = available_ransac_files() if files is None else files
    for i in files:
        r = load_ransac(i)
        _write_rows(os.path.join(ddir, "Triplet_Edgels", "Triplet_Edgels_%03d.txt" % i), r["triplet_edgels"])
        _write_rows(os.path.join(ddir, "GT_Poses21", "GT_Poses21_%03d.txt" % i), r["gt_pose21"])
        _write_rows(os.path.join(ddir, "GT_Poses31", "GT_Poses31_%03d.txt" % i), r["gt_pose31"])
        _write_rows(os.path.join(ddir, "Intrinsic_Matrix.txt"), r["K"])
    os.makedirs(os.path.join(root, "Output_Write_Files"), exist_ok=True)
    bindir = os.path.join(root, "build", "bin")
    os.makedirs(bindir, exist_ok=True)
    return bindir
