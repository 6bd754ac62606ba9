"""The problem compiler (codegen/gen_eval.py): its schedule must contain every non-zero term of the reference's index tables
exactly once, in table order, and the committed header must be what the generator produces."""
import os

import numpy as np

from trifocal_pose_estimation_using_improved_gpuhc_b200.codegen import gen_eval

PKG = os.path.dirname(os.path.dirname(os.path.abspath(gen_eval.__file__)))


def test_generated_header_is_current(tmp_path):
    g = gen_eval.build()
    text = gen_eval.emit(g, str(tmp_path / "gen.h"))
    assert text == open(os.path.join(PKG, "csrc", "hc_problem_gen.h")).read()


def test_table_statistics_match_survey():
    g = gen_eval.build()
    assert len(g["hx_terms"]) == 170                                   # 170 of 900 entries
    assert sum(len(v) for v in g["hx_terms"].values()) == 558          # 558 of 7200 padded terms
    assert sum(len(v) for v in g["h_terms"].values()) == 360
    ht = sum(1 for v in g["h_terms"].values() for (c, a, b, xs) in v if not (a == 33 and b == 33))
    assert ht == 300                                                   # 60 parameter-free terms vanish in d/dt
    assert len(g["hx_slots"]) == 25 and len(g["h_slots"]) == 16 and len(g["ht_slots"]) == 16


def _unpack(word):
    return word & 0x3ff, [(word >> 10) & 31, (word >> 15) & 31, (word >> 20) & 31]


def test_schedule_covers_every_term_once_in_table_order():
    g = gen_eval.build()
    # Hx: per (row, class) the slot sequence must equal the row's entry in that class
    for lane in range(30):
        for ci, cols in enumerate(g["classes"]):
            mine = [c for c in cols if (lane, c) in g["hx_terms"]]
            assert len(mine) <= 1
            want = g["hx_terms"][(lane, mine[0])] if mine else []
            got = []
            for cls, row in g["hx_slots"]:
                if cls != ci:
                    continue
                off, xs = _unpack(gen_eval.pack_word(row[lane]))
                if off == 0:
                    assert xs == [30, 30, 30]
                    continue
                c, a, b = g["cq_list"][off // 8]
                got.append((c, a, b, [x for x in xs[:2] if x != 30]))
            assert got == want
    for name, lst, drop in (("h_slots", g["cq_list"], False), ("ht_slots", g["dq_list"], True)):
        for lane in range(30):
            want = [(c, a, b, xs) for (c, a, b, xs) in g["h_terms"][lane] if not (drop and a == 33 and b == 33)]
            got = []
            for row in g[name]:
                off, xs = _unpack(gen_eval.pack_word(row[lane]))
                if off == 0:
                    continue
                c, a, b = lst[off // 8]
                got.append((c, a, b, [x for x in xs if x != 30]))
            assert got == want
    # lanes 30, 31 never hold a term
    for _, row in g["hx_slots"]:
        assert row[30] is None and row[31] is None


def test_column_classes_partition_the_nonzero_columns():
    g = gen_eval.build()
    cols = sorted(c for cl in g["classes"] for c in cl)
    assert cols == list(range(30))
    for cl in g["classes"]:
        for lane in range(30):
            assert sum((lane, c) in g["hx_terms"] for c in cl) <= 1
