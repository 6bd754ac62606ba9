#!/usr/bin/env python3
"""Queue-drain (tail) model of the default round (CPU only; TEST INFRASTRUCTURE: drives oracle/).

A path is a sequential chain of up to 549 stages (mean 282) and the default round gives every resident warp (148 SMs x 20) only 10.5 paths,
so the launch ends with warps idling while the last long paths finish.  This script takes the per-path stage counts of the default round
from the oracle, list-schedules them on 2 960 warp slots in several queue orders and prints the makespan over the ideal — to show which part
of the distance between the default round (4 270 hypotheses/s) and the large-round rate (4 600) is drain, and that no a-priori ordering
(by track, by hypothesis) removes it: only knowing the path lengths (LPT) would.
    python tools/tail_model.py
"""
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures

prob, rs = fixtures.load_problem(), fixtures.load_ransac(0)
orc = Oracle(prob)
tgt, dif, picked = orc.prepare_target_params(0, 100, rs["locations"], rs["tangents"])
tr, cv, inf, st = orc.track(tgt, dif, True)
cost = (st[:, 1] + st[:, 2]).astype(float)          # stages per path
C = cost.reshape(100, 312)
W = 148 * 20
ideal = cost.sum() / W


def makespan(order):
    h = [0.0] * W
    heapq.heapify(h)
    for p in order:
        heapq.heappush(h, heapq.heappop(h) + cost[p])
    return max(h) / ideal


print("stages per path: mean %.0f, median %.0f, p90 %.0f, max %.0f; paths above 450 stages: %.1f %%; paths per warp slot %.1f"
      % (cost.mean(), np.median(cost), np.percentile(cost, 90), cost.max(), 100.0 * (cost > 450).mean(), len(cost) / W))
print("makespan / ideal, list scheduling on %d warp slots:" % W)
print("  queue order (hypothesis-major, as launched)   %.3f" % makespan(range(len(cost))))
print("  random order                                  %.3f" % makespan(np.random.RandomState(0).permutation(len(cost))))
for name, key in (("per-track mean", C.mean(0)), ("per-track share of long paths", (C > 450).mean(0)), ("per-track maximum", C.max(0))):
    print("  tracks sorted by %-30s %.3f" % (name + " (in-sample)", makespan(np.argsort(-np.tile(key, 100), kind="stable"))))
print("  longest path first (needs the answer)         %.3f" % makespan(np.argsort(-cost)))
print("correlation of per-track mean length between hypotheses 0-49 and 50-99: %.3f" % np.corrcoef(C[:50].mean(0), C[50:].mean(0))[0, 1])
