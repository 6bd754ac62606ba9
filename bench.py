#!/usr/bin/env python3
"""Benchmark of the one hot path: HC path tracking of the trifocal_2op1p_30x30 RANSAC round.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--hyp H] [--abort]

One "step" = one RANSAC round: H hypotheses x 312 homotopy paths tracked in one launch per GPU
(BASELINE.json configs[1]: the full default run — H = 100 = NUM_OF_RANSAC_ITERATIONS, 80 max steps, 3 corrections,
no early abort, positive-depth pruning on as in the reference GPU kernels; dataset file 000, sampler seed 0).
N > 1 (torchrun, one process per GPU): every rank tracks its own 100 hypotheses of a 100*N-hypothesis round,
sharded contiguously exactly like sub_RANSAC_iters (reference GPU_HC_Solver.cpp:85-88) -> weak scaling; the only
exchange is an all_gather of a 64-byte per-rank result record.

Prints ONE JSON line (rank 0).  `value` = hypotheses/s with inputs resident in HBM, timed with CUDA events on the
launch stream; `e2e` = the same through host buffers (H2D of the parameters + launch + D2H of every end point and flag,
what GPU_HC_Solver::Solve_by_GPU_HC does, reference GPU_HC_Solver.cpp:335-362,449-460).
`--impl reference` times the reference's own CPU-HC (oracle/_ref, built from /root/reference) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PRED_STAGE = 91312.0     # lerp + Hx + Ht + 30x30 complex solve + update
FLOPS_CORR_STAGE = 90874.0     #                   Hx + H + solve + update
NOMINAL_FP32_TFLOPS = 74.4     # 148 SMs x 128 lanes x 2 flop x 1.965 GHz


def _rank_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


def reference_arm(args):
    """Reference CPU-HC (unmodified sources, oracle/_ref) on all host cores; bounded sample per step."""
    rank, _, world = _rank_env()
    if rank != 0:
        return 0
    from oracle import pyoracle
    from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures
    cores = os.cpu_count() or 1
    sample = args.ref_hyp
    line = {"impl": "reference", "metric": "RANSAC hypotheses/s (312 HC paths each)", "unit": "hypotheses/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "c64", "data": "synthetic",
            "config": {"workload": "trifocal_2op1p_30x30 default RANSAC round: 100 hypotheses x 312 paths per GPU, 80 max steps, "
                                   "3 corrections, pruning on, early abort off, dataset Synthetic/000, seed 0",
                       "sample_hypotheses_per_step": sample,
                       "note": "bounded sample of that round; the reference CPU-HC has no path pruning (SURVEY.md App. E-8)"}}
    try:
        ref = pyoracle.ReferenceCPU()
        kind = "reference"
    except (FileNotFoundError, OSError):
        ref = None
        kind = "port"
    with tempfile.TemporaryDirectory() as tmp:
        times = []
        if ref is not None:
            bindir = fixtures.materialize_tree(tmp, files=[0])
            for i in range(args.warmup + args.steps):
                _, _, _, _, sec = ref.run(bindir, sample, seed=0, dataset_index=0, n_cores=cores)
                if i >= args.warmup:
                    times.append(sec)
        else:
            prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
            orc = pyoracle.Oracle(prob)
            tgt, dif, _ = orc.prepare_target_params(0, sample, rs["locations"], rs["tangents"])
            for i in range(args.warmup + args.steps):
                t0 = time.perf_counter()
                orc.track(tgt, dif, prune=False, n_threads=cores)
                if i >= args.warmup:
                    times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    v = sample / t
    line.update({"value": v, "ms_per_step": t * 1e3, "paths_per_s": v * 312,
                 "cpu_baseline": {"value": v, "unit": "hypotheses/s", "cores": cores, "kind": kind,
                                  "sample": "first %d hypotheses of the seed-0 round, OpenMP over paths, %d threads" % (sample, cores)},
                 "e2e": {"value": v, "unit": "hypotheses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(line))
    return 0


def cpu_baseline_leg(sample_hyp):
    from oracle import pyoracle
    from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures
    cores = os.cpu_count() or 1
    try:
        ref = pyoracle.ReferenceCPU()
        with tempfile.TemporaryDirectory() as tmp:
            bindir = fixtures.materialize_tree(tmp, files=[0])
            _, _, _, _, sec = ref.run(bindir, sample_hyp, seed=0, dataset_index=0, n_cores=cores)
        kind = "reference"
    except (FileNotFoundError, OSError):
        prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
        orc = pyoracle.Oracle(prob)
        tgt, dif, _ = orc.prepare_target_params(0, sample_hyp, rs["locations"], rs["tangents"])
        t0 = time.perf_counter()
        orc.track(tgt, dif, prune=False, n_threads=cores)
        sec = time.perf_counter() - t0
        kind = "port"
    return {"value": sample_hyp / sec, "unit": "hypotheses/s", "cores": cores, "kind": kind, "seconds": sec,
            "sample": "first %d hypotheses of the seed-0 round (%d paths), no pruning (reference CPU-HC), %d threads"
                      % (sample_hyp, sample_hyp * 312, cores)}


def ref_gpu_leg(prob, rs, target, diff, H, abort):
    """The reference's own GPU-HC++ kernel (unmodified sources compiled for sm_100a with a MAGMA shim) on the
    same round, same GPU, timed launch->sync with CUDA events like multi_GPUs_time; tracks are re-loaded outside the timing."""
    import torch
    from oracle import pyoracle
    if not os.path.exists(pyoracle.REF_GPU_SO):
        return None
    ref = pyoracle.ReferenceGPU(prob, device="cuda:%d" % torch.cuda.current_device())
    ref.setup(target, diff, rs["locations"], rs["K"])
    ts = []
    for i in range(5):
        ref.reload()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ref.track_abort() if abort else ref.track()
        b.record()
        b.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    tr, cv, inf = ref.results()
    ms = float(np.mean(ts))
    return {"what": "reference GPU-HC++ kernel (…TrunPaths%s.cu) built for sm_100a, same inputs" % ("_TrunRANSAC" if abort else ""),
            "ms_per_step": ms, "value": H / (ms * 1e-3), "unit": "hypotheses/s",
            "result": {"converged": int(cv.sum()), "infinity": int(inf.sum())}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--hyp", type=int, default=100, help="hypotheses per GPU per step (100 = default RANSAC round)")
    ap.add_argument("--abort", action="store_true", help="Abort_RANSAC_by_Good_Sol = true (configs[2])")
    ap.add_argument("--no-prune", action="store_true")
    ap.add_argument("--ref-hyp", type=int, default=8, help="hypotheses per step of the reference CPU arm (--impl reference)")
    ap.add_argument("--cpu-baseline-hyp", type=int, default=40, help="hypotheses of the cpu_baseline sample (about 15 s on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip timing the reference GPU-HC++ kernels (oracle/_ref/libref_gpuhc.so)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc

    rank, local_rank, world = _rank_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    prob = fixtures.load_problem()
    rs = fixtures.load_ransac(0)
    H = args.hyp
    prune = not args.no_prune
    # one rand() stream for the whole (multi-GPU) round, consumed in GPU-major order (GPU_HC_Solver.cpp:263-271)
    picked_all = hc.sample_hypotheses(0, H * world, rs["locations"].shape[0])
    offs = hc.shard_offsets(H * world, world)
    picked = picked_all[offs[rank]:offs[rank + 1]]
    target, diff = hc.target_params_from_picks(picked, rs["locations"], rs["tangents"], prob["start_params"])

    trk = hc.Tracker(device=dev, problem=prob, stats=True)
    trk.reserve(H)
    trk.set_edgels(rs["locations"], rs["K"])
    n_paths = H * hc.NUM_TRACKS

    # pinned host staging for the e2e leg (what GPU_HC_Solver keeps in h_Target_Params / h_GPU_HC_Track_Sols)
    h_target = torch.view_as_real(torch.from_numpy(target)).contiguous().pin_memory()
    h_diff = torch.view_as_real(torch.from_numpy(diff)).contiguous().pin_memory()
    h_tracks = torch.empty((n_paths, 31, 2), dtype=torch.float32).pin_memory()
    h_conv = torch.empty(n_paths, dtype=torch.uint8).pin_memory()
    h_inf = torch.empty(n_paths, dtype=torch.uint8).pin_memory()
    h_rec = torch.zeros(16, dtype=torch.int32).pin_memory()
    d_rec = torch.zeros(16, dtype=torch.int32, device=dev)
    gathered = [torch.zeros(16, dtype=torch.int32, device=dev) for _ in range(world)] if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def launch():
        if args.abort:
            trk.track_abort(H, prune=prune)
        else:
            trk.track(H, prune=prune)

    def step_device():
        launch()
        if world > 1:   # tiny result gather: found flag / best path / counts
            d_rec.copy_(trk.d_best)
            dist.all_gather(gathered, d_rec)

    trk.upload_params(target, diff)
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()

    # ---- device-timed leg ------------------------------------------------------------------------------------
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = trk.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.fill_(1.0)            # L2 flush between timed iterations (outside the event pair)
        a.record()
        step_device()
        b.record()
    torch.cuda.synchronize()
    launches = trk.launches - l0
    if world > 1:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    dev_ms = float(t_ms.item())
    ms_per_step = dev_ms / args.steps

    # per-path stage counters of the last step -> algorithmic flops of one launch
    tracks, conv, inf, stats = trk.results(H)
    flops_launch = FLOPS_PRED_STAGE * float(stats[:, 1].sum()) + FLOPS_CORR_STAGE * float(stats[:, 2].sum())
    counts = hc.count_solutions(tracks, conv, inf, H)

    # ---- e2e leg: host buffers in, host buffers out, every step --------------------------------------------
    def step_e2e():
        trk.d_target[:H].copy_(h_target, non_blocking=True)
        trk.d_diff[:H].copy_(h_diff, non_blocking=True)
        launch()
        h_tracks.copy_(trk.d_tracks[:n_paths], non_blocking=True)
        h_conv.copy_(trk.d_conv[:n_paths], non_blocking=True)
        h_inf.copy_(trk.d_inf[:n_paths], non_blocking=True)
        if args.abort:
            h_rec.copy_(trk.d_best, non_blocking=True)
        torch.cuda.synchronize()
        return int(h_conv.sum())     # device -> host read of the step's result

    for _ in range(2):
        step_e2e()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e.item())
    h2d = h_target.numel() * 4 + h_diff.numel() * 4
    d2h = h_tracks.numel() * 4 + h_conv.numel() + h_inf.numel() + (64 if args.abort else 0)

    if rank == 0:
        fp32_peak = trk.ffma_probe()
        hyp_per_s = H * world / (ms_per_step * 1e-3)
        achieved = flops_launch / (ms_per_step * 1e-3) / 1e12
        info = trk.kernel_info(abort=args.abort)
        line = {
            "metric": "RANSAC hypotheses/s (312 HC paths each)", "value": hyp_per_s, "unit": "hypotheses/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c64", "data": "synthetic",
            "paths_per_s": hyp_per_s * hc.NUM_TRACKS,
            "config": {"workload": "trifocal_2op1p_30x30 default RANSAC round: %d hypotheses x 312 paths per GPU, 80 max steps, "
                                   "3 corrections, pruning %s, early abort %s, dataset Synthetic/000, seed 0"
                                   % (H, "on" if prune else "off", "on" if args.abort else "off"),
                       "hypotheses_per_gpu": H, "paths_per_gpu": n_paths, "sharding": "contiguous hypotheses (sub_RANSAC_iters)",
                       "l2": "256 MB L2 flush between timed iterations (outside the event pair)",
                       "kernel": info},
            "e2e": {"value": H * world / (e2e_s / args.steps), "unit": "hypotheses/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp32_peak if fp32_peak else None,
                         "traffic": 2267392, "traffic_source": "dram__bytes_read.sum (2.03 MB) + dram__bytes_write.sum (0.24 MB) of the tracker launch, "
                                                              "ncu --set full capture v11 in profiles/ncu_r1.md (0.2-2 MB across captures); the 7.8 MB of results stay in the 126 MB L2",
                         "hbm": {"algorithmic_bytes_per_launch": H * (2 * 34 * 8) + n_paths * (31 * 8 + 2),
                                 "achieved_gbs": (H * (2 * 34 * 8) + n_paths * (31 * 8 + 2)) / (ms_per_step * 1e-3) / 1e9,
                                 "peak_gbs": 6553.0, "peak_source": "MEASURED_PEAKS.json hbm_gbs",
                                 "note": "the path is not HBM-bound: 78 KB and ~8 Gflop (model) per hypothesis"},
                         "peak_source": "FFMA probe kernel timed live on this GPU (MEASURED_PEAKS.json has no FP32 entry); "
                                        "nominal 148x128x2x1.965GHz = %.1f" % NOMINAL_FP32_TFLOPS,
                         "frac_of_nominal": achieved / NOMINAL_FP32_TFLOPS,
                         "flops_per_launch": flops_launch, "model": "91312*pred_stages + 90874*corr_stages (dense-LU model, SURVEY §8d)",
                         "stages_per_path": float(stats[:, 1].sum() + stats[:, 2].sum()) / n_paths},
            "result": {"converged": int(counts[:, 0].sum()), "infinity": int(counts[:, 1].sum()), "real": int(counts[:, 2].sum())},
        }
        if world == 1 and not args.no_ref_gpu:
            rg = ref_gpu_leg(prob, rs, target, diff, H, args.abort)
            if rg:
                rg["speedup_ours_vs_ref_gpu"] = rg["ms_per_step"] / ms_per_step
                line["ref_gpu"] = rg
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_leg(args.cpu_baseline_hyp)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
