"""world_size-2 gloo tests of the N > 1 host logic (no GPU): every rank takes its contiguous hypothesis shard
(sub_RANSAC_iters, reference GPU_HC_Solver.cpp:85-88) of ONE rand() stream, "tracks" it with the CPU oracle standing in
for the device, and the ranks exchange only a small result record through all_gather — the single collective of the multi-GPU
path.  The reduced result must equal the single-rank result.  Two record formats are covered: the 16-int early-abort record and the
128-byte best-pose record bench.py gathers at N > 1 (hcb200_pose_record; reduction mirrored by hc.reduce_pose_records_host)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def reduce_records(records, offsets_paths):
    """Host reduction of the per-rank records: found = any, best = smallest GLOBAL path id, counts summed."""
    found, best, n21, n31 = 0, -1, 0, 0
    conv = inf = real = 0
    for r, rec in enumerate(records):
        conv += int(rec[5]); inf += int(rec[6]); real += int(rec[7])
        if int(rec[0]):
            gid = offsets_paths[r] + int(rec[1])
            if not found or gid < best:
                best, n21, n31 = gid, int(rec[2]), int(rec[3])
            found = 1
    return found, best, n21, n31, conv, inf, real


def _rank_record(orc, hc, rs, target, diff):
    tr, cv, inf, st = orc.track(target, diff, prune=True, n_threads=2)
    rec = np.zeros(16, np.int32)
    rec[1] = -1
    hits = []
    for pth in np.nonzero(cv)[0]:
        ok, n21, n31, _ = orc.score(tr[pth], rs["locations"], rs["K"])
        if ok:
            hits.append((int(pth), n21, n31))
    if hits:
        rec[0], rec[1], rec[2], rec[3], rec[4] = 1, hits[0][0], hits[0][1], hits[0][2], len(hits)
    c = hc.count_solutions(tr, cv, inf, target.shape[0]).sum(0)
    rec[5:8] = c
    return rec


def _worker(rank, world, port, n_hyp, out_dir):
    sys.path.insert(0, ROOT)
    from oracle.pyoracle import Oracle
    from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
    orc = Oracle(prob)
    picked = hc.sample_hypotheses(0, n_hyp, rs["locations"].shape[0])
    offs = hc.shard_offsets(n_hyp, world)
    mine = picked[offs[rank]:offs[rank + 1]]
    target, diff = hc.target_params_from_picks(mine, rs["locations"], rs["tangents"], prob["start_params"])
    rec = torch.from_numpy(_rank_record(orc, hc, rs, target, diff))
    gathered = [torch.zeros(16, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, rec)
    dist.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), torch.stack(gathered).numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shards_reduce_to_single_rank_result(tmp_path, oracle, ransac0, problem):
    from trifocal_pose_estimation_using_improved_gpuhc_b200 import hc
    n_hyp, world = 3, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_hyp, str(tmp_path)), nprocs=world, join=True)
    gathered = np.load(tmp_path / "gathered.npy")
    offs = hc.shard_offsets(n_hyp, world)
    multi = reduce_records(gathered, [o * 312 for o in offs[:-1]])

    picked = hc.sample_hypotheses(0, n_hyp, 5117)
    target, diff = hc.target_params_from_picks(picked, ransac0["locations"], ransac0["tangents"], problem["start_params"])
    single = reduce_records([_rank_record(oracle, hc, ransac0, target, diff)], [0])
    assert multi == single
    assert multi[0] == 1 and multi[1] == 104 and multi[2:4] == (5117, 5117)     # hypothesis 0 / track 104 is the GT pose


# ---------------------------------------------------------------------------------------------------------------------
# the 128-byte best-pose record of bench.py's N > 1 step: score -> record -> all_gather_into_tensor -> arg-max

def _pose_record(orc, hc, rs, target, diff, path_offset, rank):
    """What hcb200_score_tracks + hcb200_make_pose_record produce on one GPU, restated with the CPU oracle and the host scoring."""
    import ctypes
    host = ctypes.CDLL(os.path.join(ROOT, "trifocal_pose_estimation_using_improved_gpuhc_b200", "lib", "libhcb200_host.so"))
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    loc = np.ascontiguousarray(rs["locations"], np.float32)
    K = np.ascontiguousarray(rs["K"], np.float32).reshape(-1)
    tr, cv, inf, st = orc.track(target, diff, prune=True, n_threads=2)
    best, n_cand = None, 0
    for pth in np.nonzero(cv)[0]:
        x = np.ascontiguousarray(np.stack([tr[pth].real, tr[pth].imag], -1).astype(np.float32))
        n21, n31 = ctypes.c_int(), ctypes.c_int()
        if host.hcb200_host_score_track(vp(x), vp(loc), loc.shape[0], vp(K), ctypes.byref(n21), ctypes.byref(n31)):
            n_cand += 1
            key = (min(n21.value, n31.value), -int(pth))
            if best is None or key > best[0]:
                best = (key, int(pth), n21.value, n31.value)
    if best is None:
        return hc.encode_pose_record(0, 0, 0, 0, 0, rank, -1)
    return hc.encode_pose_record(1, best[2], best[3], n_cand, 0, rank, path_offset + best[1])


def _pose_worker(rank, world, port, n_hyp, out_dir):
    sys.path.insert(0, ROOT)
    from oracle.pyoracle import Oracle
    from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures, hc
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
    orc = Oracle(prob)
    picked = hc.sample_hypotheses(0, n_hyp, rs["locations"].shape[0])
    offs = hc.shard_offsets(n_hyp, world)
    mine = picked[offs[rank]:offs[rank + 1]]
    target, diff = hc.target_params_from_picks(mine, rs["locations"], rs["tangents"], prob["start_params"])
    rec = torch.from_numpy(_pose_record(orc, hc, rs, target, diff, offs[rank] * 312, rank))
    gathered = torch.zeros(world * 32, dtype=torch.float32)
    dist.all_gather_into_tensor(gathered, rec)
    if rank == 0:
        np.save(os.path.join(out_dir, "pose_records.npy"), gathered.numpy().reshape(world, 32))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_best_pose_records_reduce_to_single_rank_result(tmp_path, oracle, ransac0, problem):
    from trifocal_pose_estimation_using_improved_gpuhc_b200 import hc
    n_hyp, world = 4, 2
    port = _free_port()
    mp.spawn(_pose_worker, args=(world, port, n_hyp, str(tmp_path)), nprocs=world, join=True)
    multi = hc.reduce_pose_records_host(np.load(tmp_path / "pose_records.npy"))
    picked = hc.sample_hypotheses(0, n_hyp, 5117)
    target, diff = hc.target_params_from_picks(picked, ransac0["locations"], ransac0["tangents"], problem["start_params"])
    single = hc.reduce_pose_records_host(_pose_record(oracle, hc, ransac0, target, diff, 0, 0)[None, :])
    for k in ("found", "path_id", "inliers21", "inliers31", "n_candidates", "abort_flag"):
        assert multi[k] == single[k], k
    assert multi["found"] == 1 and multi["path_id"] == 104 and (multi["inliers21"], multi["inliers31"]) == (5117, 5117) and multi["rank"] == 0
    # a record round-trips through its 128 bytes, and the reduction prefers support over path id, then the lower path id
    a = hc.encode_pose_record(1, 10, 20, 3, 0, 0, 999, np.arange(24))
    b = hc.encode_pose_record(1, 12, 11, 2, 1, 1, 5)
    c = hc.encode_pose_record(1, 30, 11, 1, 0, 2, 4)
    assert a.nbytes == 128 and hc.decode_pose_record(a)["path_id"] == 999 and hc.decode_pose_record(a)["t21"].tolist() == [9, 10, 11]
    r = hc.reduce_pose_records_host(np.stack([a, b, c]))
    assert r["path_id"] == 4 and r["n_candidates"] == 6 and r["abort_flag"] == 1
