"""CPU: `.dat` hierarchy reader / writer (reference include/lexls/tools.h:273-453, export_hierarchy.m)."""
import os

import numpy as np


def test_read_test01(golden_dir):
    from lexls_b200 import datfile
    h = datfile.read_hierarchy(os.path.join(golden_dir, "test_01.dat"))
    assert h.hier_type == 210 and h.nVar == 88 and h.ObjDim == [74, 33, 3, 2, 97]
    assert h.objectives[0].shape == (74, 2) and h.objectives[1].shape == (33, 90)
    assert h.var_index.min() == 0 and h.var_index.max() == 87 and len(set(h.var_index.tolist())) == 74
    assert h.var_index[0] == 84  # the file says 85 (1-based)
    counts = [np.bincount(g, minlength=4).tolist() for g in h.active_set_guess]
    assert counts[0] == [66, 6, 0, 2] and counts[4] == [0, 0, 0, 97]
    assert h.solution.shape == (88,) and h.solution_guess.shape == (88,)


def test_roundtrip(tmp_path, golden_dir):
    from lexls_b200 import datfile
    h = datfile.read_hierarchy(os.path.join(golden_dir, "test_01.dat"))
    p = tmp_path / "copy.dat"
    datfile.write_hierarchy(str(p), h)
    g = datfile.read_hierarchy(str(p))
    assert g.ObjDim == h.ObjDim and g.ObjType == h.ObjType
    for a, b in zip(h.objectives, g.objectives):
        assert np.array_equal(a, b)
    assert np.array_equal(h.var_index, g.var_index)
    assert np.array_equal(h.solution, g.solution)
    for a, b in zip(h.active_set_guess, g.active_set_guess):
        assert np.array_equal(a, b)
