"""Other minimal problems: a problem folder in the reference's layout -> a tracker library -> paths tracked on the GPU.

The reference defines a minimal problem as DATA — `problems/<name>/gpuhc_settings.yaml` (Num_Of_Vars, Num_Of_Params, Num_Of_Tracks,
dHdx_Max_Terms, dHdt_Max_Terms …), `start_sols.txt`, `start_params.txt` and the two evaluation-index tables `dHdx_indx.txt`, `dHdt_indx.txt`
(Data_Reader.cpp:37-189; gpu-idx-evals/*.cuh walk the tables on the device) — but ships GPU kernels for one problem only.  Here the tables
are COMPILED: `codegen/gen_eval.py --problem-dir` writes a header, `make problem PROBLEM_DIR=…` builds csrc/hc_tracker.cu against it into
`lib/libhcb200_<name>.so`, and that library exports the same C ABI (`hcb200_track`, `hcb200_refine_tracks`, `hcb200_count_solutions`,
`hcb200_problem_info`) with the problem's own array sizes.  There is NO CPU fallback: a missing library or GPU raises.
"""
import ctypes
import os
import subprocess

import numpy as np

from .hc import FLAG_PRUNE_PATHS, FLAG_SPLIT_LONG_PATHS, SPLIT_MAX_HYPOTHESES, HCB200Error

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


def read_problem(problem_dir):
    """Numpy view of a problem folder: dict(spec, start_sols [T,N] c64, start_params [NP] c64, dHdx_indx, dHdt_indx flat int32)."""
    from .codegen import gen_eval
    spec, hx, ht = gen_eval.read_problem_dir(problem_dir)
    sp = np.loadtxt(os.path.join(problem_dir, "start_params.txt"), dtype=np.float32).reshape(-1, 2)
    ss = np.loadtxt(os.path.join(problem_dir, "start_sols.txt"), dtype=np.float32).reshape(spec["n_tracks"], spec["n_vars"], 2)
    if sp.shape[0] != spec["n_params"]:
        raise HCB200Error("%s: start_params.txt holds %d parameters, gpuhc_settings.yaml says %d" % (problem_dir, sp.shape[0], spec["n_params"]))
    return dict(spec=spec, start_params=(sp[:, 0] + 1j * sp[:, 1]).astype(np.complex64),
                start_sols=(ss[..., 0] + 1j * ss[..., 1]).astype(np.complex64), dHdx_indx=hx.astype(np.int32), dHdt_indx=ht.astype(np.int32))


def library_path(name):
    return os.path.join(_HERE, "lib", "libhcb200_%s.so" % name)


def build_library(problem_dir):
    """Compile the problem (generator + nvcc, sm_100a) unless its library is already there; returns the library path."""
    name = os.path.basename(os.path.normpath(problem_dir))
    subprocess.check_call(["make", "-C", _ROOT, "problem", "PROBLEM_DIR=" + os.path.relpath(os.path.abspath(problem_dir), _ROOT)], stdout=subprocess.DEVNULL)
    return library_path(name)


def load_problem_library(path):
    if not os.path.exists(path):
        raise HCB200Error("CUDA extension %s is missing: run `make problem PROBLEM_DIR=…` first" % path)
    lib = ctypes.CDLL(path)
    vp, i32, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint
    lib.hcb200_workspace_bytes.restype = ctypes.c_size_t
    lib.hcb200_workspace_bytes_for.restype = ctypes.c_size_t
    lib.hcb200_workspace_bytes_for.argtypes = [i32]
    lib.hcb200_error_string.restype = ctypes.c_char_p
    lib.hcb200_error_string.argtypes = [i32]
    lib.hcb200_problem_info.restype = i32
    lib.hcb200_problem_info.argtypes = [ctypes.POINTER(i32)] * 4 + [ctypes.POINTER(ctypes.c_char_p)]
    lib.hcb200_track.restype = i32
    lib.hcb200_track.argtypes = [vp, i32, i32, i32, i32, u32] + [vp] * 9
    lib.hcb200_track_abort.restype = i32
    lib.hcb200_track_abort.argtypes = [vp, i32, i32, i32, i32, i32, u32] + [vp] * 14
    lib.hcb200_refine_tracks.restype = i32
    lib.hcb200_refine_tracks.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    lib.hcb200_count_solutions.restype = i32
    lib.hcb200_count_solutions.argtypes = [vp, i32, vp, vp, vp, vp]
    return lib


def problem_info(lib):
    v = [ctypes.c_int() for _ in range(4)]
    name = ctypes.c_char_p()
    lib.hcb200_problem_info(*[ctypes.byref(x) for x in v], ctypes.byref(name))
    return dict(n_vars=v[0].value, n_params=v[1].value, n_tracks=v[2].value, trifocal=v[3].value, name=name.value.decode())


class ProblemTracker:
    """Device state and launches for one compiled problem: every hypothesis is one set of target parameters, every hypothesis tracks all
    Num_Of_Tracks start solutions (the same batch layout as the trifocal path: tracks [H*T][N+1], flags [H*T])."""

    def __init__(self, problem_dir, problem=None, device=None, max_steps=80, max_corr=3, dt_inc=4, stats=False, split=True):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise HCB200Error("no CUDA device: the tracker has no CPU path")
        name = os.path.basename(os.path.normpath(problem_dir))
        self.lib = load_problem_library(library_path(name))
        self.problem = problem or read_problem(problem_dir)
        spec, info = self.problem["spec"], problem_info(self.lib)
        if (info["name"], info["n_vars"], info["n_params"], info["n_tracks"]) != (spec["name"], spec["n_vars"], spec["n_params"], spec["n_tracks"]):
            raise HCB200Error("library %s was compiled for %r, the folder describes %r" % (library_path(name), info, spec))
        self.N, self.NP1, self.T = spec["n_vars"], spec["n_params"] + 1, spec["n_tracks"]
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_steps, self.max_corr, self.dt_inc = max_steps, max_corr, dt_inc
        ss = np.ones((self.T, self.N + 1), np.complex64)
        ss[:, :self.N] = self.problem["start_sols"]
        self.start_params_h = np.concatenate([self.problem["start_params"], [1.0]]).astype(np.complex64)
        with torch.cuda.device(self.device):
            self.d_start_sols = torch.view_as_real(torch.from_numpy(ss)).contiguous().to(self.device)
            self.d_start_params = torch.view_as_real(torch.from_numpy(self.start_params_h)).contiguous().to(self.device)
            self.d_ws = torch.zeros(int(self.lib.hcb200_workspace_bytes()), dtype=torch.uint8, device=self.device)
        self.want_stats = stats
        self.split = bool(split)
        self.capacity = 0

    def _check(self, code, what):
        if code != 0:
            msg = self.lib.hcb200_error_string(code)
            raise HCB200Error("%s failed: cudaError %d (%s)" % (what, code, msg.decode() if msg else "?"))

    def diff_params(self, target):
        """target - start, component-wise in float32 like the reference (GPU_HC_Solver.cpp:298-299)."""
        diff = np.empty_like(target)
        diff.real = target.real - self.start_params_h.real[None, :]
        diff.imag = target.imag - self.start_params_h.imag[None, :]
        return diff

    def upload_params(self, target):
        """target: complex64 [H][NP+1] (index NP is the constant-one pad)."""
        torch = self.torch
        target = np.ascontiguousarray(target, np.complex64)
        H = target.shape[0]
        if H > self.capacity:
            P = H * self.T
            dev = self.device
            self.d_target = torch.empty((H, self.NP1, 2), dtype=torch.float32, device=dev)
            self.d_diff = torch.empty_like(self.d_target)
            self.d_tracks = torch.empty((P, self.N + 1, 2), dtype=torch.float32, device=dev)
            self.d_conv = torch.empty(P, dtype=torch.uint8, device=dev)
            self.d_inf = torch.empty(P, dtype=torch.uint8, device=dev)
            self.d_stats = torch.empty((P, 4), dtype=torch.int32, device=dev) if self.want_stats else None
            self.d_counts = torch.empty((H, 3), dtype=torch.int32, device=dev)
            self.d_sums = torch.empty((P, 2), dtype=torch.float32, device=dev)
            if self.split:
                self.d_ws = torch.zeros(int(self.lib.hcb200_workspace_bytes_for(min(H, SPLIT_MAX_HYPOTHESES))), dtype=torch.uint8, device=dev)
            self.capacity = H
        self.d_target[:H].copy_(torch.view_as_real(torch.from_numpy(target)))
        self.d_diff[:H].copy_(torch.view_as_real(torch.from_numpy(self.diff_params(target))))
        return H

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def track(self, n_hyp, prune=False):
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        with self.torch.cuda.device(self.device):
            rc = self.lib.hcb200_track(self._stream(), n_hyp, self.max_steps, self.max_corr, self.dt_inc,
                                       (FLAG_PRUNE_PATHS if prune else 0) | (FLAG_SPLIT_LONG_PATHS if (self.split and n_hyp <= SPLIT_MAX_HYPOTHESES) else 0),
                                       p(self.d_start_sols), p(self.d_start_params), p(self.d_target), p(self.d_diff),
                                       p(self.d_tracks), p(self.d_conv), p(self.d_inf), p(self.d_stats), p(self.d_ws))
        self._check(rc, "hcb200_track")

    def refine_tracks(self, n_hyp, iters=3):
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        n = n_hyp * self.T
        with self.torch.cuda.device(self.device):
            rc = self.lib.hcb200_refine_tracks(self._stream(), n, int(iters), p(self.d_target), p(self.d_conv), p(self.d_tracks), p(self.d_sums), p(self.d_ws))
        self._check(rc, "hcb200_refine_tracks")
        self.torch.cuda.synchronize(self.device)
        return self.d_sums[:n].cpu().numpy()

    def count_solutions(self, n_hyp):
        """Per hypothesis (converged, infinity, real) on the device."""
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with self.torch.cuda.device(self.device):
            rc = self.lib.hcb200_count_solutions(self._stream(), n_hyp, p(self.d_tracks), p(self.d_conv), p(self.d_inf), p(self.d_counts))
        self._check(rc, "hcb200_count_solutions")
        self.torch.cuda.synchronize(self.device)
        return self.d_counts[:n_hyp].cpu().numpy()

    def results(self, n_hyp):
        self.torch.cuda.synchronize(self.device)
        n = n_hyp * self.T
        tr = self.d_tracks[:n].cpu().numpy()
        tracks = (tr[..., 0] + 1j * tr[..., 1]).astype(np.complex64)
        stats = self.d_stats[:n].cpu().numpy() if self.d_stats is not None else None
        return tracks, self.d_conv[:n].cpu().numpy(), self.d_inf[:n].cpu().numpy(), stats
