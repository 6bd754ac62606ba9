#!/usr/bin/env python3
"""Run the UNMODIFIED reference host layer (GPU_HC_Solver.cpp + Data_Reader.cpp + Evaluations.cpp, built by `make -C oracle
ref_dropin`) twice: linked against the reference's own kernels, and against integration/hcb200_shim.cpp + libhcb200.so.  Prints
what the reference itself reports: `GPU Computation Time` (its multi_GPUs_time, GPU_HC_Solver.cpp:384-446 — one round, the
first launch of the process, so module load is inside) and the statistics file.  GPU box only.
Usage: python tools/dropin_compare.py [n_hyp] [abort]"""
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n_hyp = sys.argv[1] if len(sys.argv) > 1 else "100"
abort = len(sys.argv) > 2 and sys.argv[2] == "abort"
with tempfile.TemporaryDirectory() as root:
    fixtures.materialize_tree(root, files=[0], settings_overrides={"Abort_RANSAC_by_Good_Sol": "true"} if abort else None)
    for exe in ("ref_gpuhc_on_refkernels", "ref_gpuhc_on_hcb200"):
        path = os.path.join(ROOT, "oracle", "_ref", exe)
        if not os.path.exists(path):
            print(exe, "not built")
            continue
        for rep in range(3):
            out = subprocess.run([path, "trifocal_2op1p_30x30", n_hyp], cwd=os.path.join(root, "build", "bin"), capture_output=True, text=True, timeout=900)
            ms = open(os.path.join(root, "Output_Write_Files", "GPU_Timings.txt")).read().split()
            st = open(os.path.join(root, "Output_Write_Files", "GPU_Sols_Statistics.txt")).read().split()
            print("%-26s n_hyp=%s abort=%s run %d: rc=%d  GPU Computation Time %s ms   converged/real/infinity = %s" % (exe, n_hyp, abort, rep, out.returncode, ms, st), flush=True)
