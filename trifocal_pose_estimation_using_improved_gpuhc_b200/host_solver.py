"""ctypes front-end of the C++ host class `GPU_HC_Solver` (lib/libhcb200_host.so, include/hcb200_host.h).

It drives the object the way the reference's own driver does (cmd/magmaHC-main.cpp:24-66 of the reference): create from a
gpuhc_settings.yaml, Allocate_Arrays, Read_Problem_Data, Read_RANSAC_Data, Prepare_Target_Params, Set_RANSAC_Abort_Arrays,
Data_Transfer_From_Host_To_Device, Solve_by_GPU_HC — one process, `Num_Of_GPUs` devices, one stream per device
(GPU_HC_Solver.cpp:85-88, 390-506 semantics).  Used by bench.py's `host_class` leg and by the tests."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "lib", "libhcb200_host.so")


class HostSolver:
    def __init__(self, tree, overrides=""):
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError("%s is missing: run `make` first" % HOST_LIB_PATH)
        self.lib = ctypes.CDLL(HOST_LIB_PATH)
        self.lib.hcb200_solver_create.restype = ctypes.c_void_p
        self.lib.hcb200_solver_kernel_seconds.restype = ctypes.c_double
        for n in ("destroy", "allocate", "read_problem", "read_ransac", "prepare", "set_abort_arrays", "h2d", "solve", "free_round",
                  "num_hypotheses", "kernel_seconds", "totals", "per_hypothesis", "copy_results", "copy_target_params", "best", "set_pruning"):
            getattr(self.lib, "hcb200_solver_" + n).argtypes = [ctypes.c_void_p] + ([ctypes.c_void_p] * 3 if n in ("copy_results", "best") else
                                                                              [ctypes.c_void_p] if n in ("totals", "per_hypothesis", "copy_target_params") else
                                                                              [ctypes.c_int] if n in ("read_ransac", "set_pruning") else
                                                                              [ctypes.c_uint] if n == "prepare" else [])
        self.lib.hcb200_solver_selected.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        yaml = os.path.join(tree, "problems", "trifocal_2op1p_30x30", "gpuhc_settings.yaml")
        ov = "Repo_Root=%s/;Verbose=false;%s" % (tree, overrides)
        self.h = self.lib.hcb200_solver_create(yaml.encode(), ov.encode())
        if not self.h:
            raise RuntimeError("hcb200_solver_create failed for %s (%s)" % (yaml, ov))

    def round(self, dataset=0, seed=0, prune=True, fetch=True):
        L, h = self.lib, self.h
        assert L.hcb200_solver_allocate(h) == 0
        assert L.hcb200_solver_read_problem(h) == 0 and L.hcb200_solver_read_ransac(h, dataset) == 0
        L.hcb200_solver_set_pruning(h, 1 if prune else 0)
        L.hcb200_solver_prepare(h, seed)
        L.hcb200_solver_set_abort_arrays(h)
        L.hcb200_solver_h2d(h)
        L.hcb200_solver_solve(h)
        H = L.hcb200_solver_num_hypotheses(h)
        n = H * 312
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        out = dict(H=H, seconds=L.hcb200_solver_kernel_seconds(h))
        tot = np.zeros(3, np.uint32)
        L.hcb200_solver_totals(h, vp(tot))
        per = np.zeros((H, 3), np.uint32)
        L.hcb200_solver_per_hypothesis(h, vp(per))
        best = np.zeros(16, np.int32)
        found = ctypes.c_int()
        res = np.zeros(4, np.float32)
        L.hcb200_solver_best(h, vp(best), ctypes.cast(ctypes.byref(found), ctypes.c_void_p), vp(res))
        sel_path = ctypes.c_int()
        sel_sup = np.zeros(2, np.uint32)
        L.hcb200_solver_selected(h, ctypes.cast(ctypes.byref(sel_path), ctypes.c_void_p), vp(sel_sup))
        out.update(totals=tot, per=per, best=best, pose_found=found.value, residuals=res, selected_path=sel_path.value,
                   selected_support=sel_sup.tolist())
        if fetch:
            tr = np.zeros((n, 31, 2), np.float32)
            cv, inf = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
            L.hcb200_solver_copy_results(h, vp(tr), vp(cv), vp(inf))
            out.update(tracks=tr[..., 0] + 1j * tr[..., 1], conv=cv, inf=inf)
        L.hcb200_solver_free_round(h)
        return out

    def close(self):
        if self.h:
            self.lib.hcb200_solver_destroy(self.h)
            self.h = None
