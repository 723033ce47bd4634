"""Pins the oracle's grid_map substrate against the reference's own known-answer tests.

Each test restates a gtest from grid_map-master/grid_map_core/test (cited per test); the expected values are the
reference's.  CPU only.
"""
import sys

import numpy as np
import pytest

from oracle import oracle as O

EPS = sys.float_info.epsilon


def geom(lx, ly, res, px, py, rows, cols, start=(0, 0)):
    g = O.make_geom(lx, ly, res, px, py, start)
    # The math tests pass buffer size / length explicitly instead of going through setGeometry.
    g.rows, g.cols, g.len_x, g.len_y = rows, cols, lx, ly
    return g


def test_position_from_index_simple():
    # GridMapMathTest.cpp:26-51
    g = geom(3.0, 2.0, 1.0, -1.0, 2.0, 3, 2)
    assert O.position_from_index(g, 0, 0) == pytest.approx((1.0 - 1.0, 0.5 + 2.0), abs=1e-15)
    assert O.position_from_index(g, 1, 0) == pytest.approx((0.0 - 1.0, 0.5 + 2.0), abs=1e-15)
    assert O.position_from_index(g, 1, 1) == pytest.approx((0.0 - 1.0, -0.5 + 2.0), abs=1e-15)
    assert O.position_from_index(g, 2, 1) == pytest.approx((-1.0 - 1.0, -0.5 + 2.0), abs=1e-15)
    assert O.position_from_index(g, 3, 1) is None


def test_position_from_index_circular_buffer():
    # GridMapMathTest.cpp:53-83
    g = geom(0.5, 0.4, 0.1, -0.1, 13.4, 5, 4, start=(3, 1))
    exp = {(3, 1): (0.2, 0.15), (4, 2): (0.1, 0.05), (2, 0): (-0.2, -0.15), (0, 0): (0.0, -0.15), (4, 3): (0.1, -0.05)}
    for (r, c), (x, y) in exp.items():
        got = O.position_from_index(g, r, c)
        assert got == pytest.approx((x - 0.1, y + 13.4), rel=1e-14, abs=1e-14)
    assert O.position_from_index(g, 5, 3) is None


def test_index_from_position_simple():
    # GridMapMathTest.cpp:85-114
    mp = (-12.4, -7.1)
    g = geom(3.0, 2.0, 1.0, mp[0], mp[1], 3, 2)
    cases = [((1.0, 0.5), (0, 0)), ((-1.0, -0.5), (2, 1)), ((0.6, 0.1), (0, 0)), ((0.4, -0.1), (1, 1)),
             ((0.4, 0.1), (1, 0))]
    for (x, y), rc in cases:
        assert O.index_from_position(g, x + mp[0], y + mp[1]) == rc
    assert O.index_from_position(g, 4.0 + mp[0], 0.5 + mp[1]) is None


def test_index_from_position_edge_cases():
    # GridMapMathTest.cpp:116-137
    g = geom(3.0, 2.0, 1.0, 0.0, 0.0, 3, 2)
    assert O.index_from_position(g, 0.0, EPS) == (1, 0)
    assert O.index_from_position(g, 0.5 - EPS, -EPS) == (1, 1)
    assert O.index_from_position(g, -0.5 - EPS, -EPS) == (2, 1)
    assert O.index_from_position(g, -1.5, 1.0) is None


def test_index_from_position_circular_buffer():
    # GridMapMathTest.cpp:139-155
    mp = (0.4, -0.9)
    g = geom(0.5, 0.4, 0.1, mp[0], mp[1], 5, 4, start=(3, 1))
    assert O.index_from_position(g, 0.2 + mp[0], 0.15 + mp[1]) == (3, 1)
    assert O.index_from_position(g, 0.03 + mp[0], -0.17 + mp[1]) == (0, 0)


def test_check_if_position_within_map():
    # GridMapMathTest.cpp:157-193
    g = geom(50.0, 25.0, 1.0, 11.4, 0.0, 50, 25)
    for x, y in [(0, 0), (5, 5), (20, 10), (20, -10), (-20, 10), (-20, -10)]:
        assert O.is_inside(g, x + 11.4, y + 0.0)
    g = geom(10.0, 5.0, 1.0, -3.0, 145.2, 10, 5)
    for x, y in [(5.5, 0.0), (-5.5, 0.0), (-5.5, 3.0), (-5.5, -3.0), (3.0, 3.0)]:
        assert not O.is_inside(g, x - 3.0, y + 145.2)
    g = geom(2.0, 3.0, 1.0, 0.0, 0.0, 2, 3)
    assert not O.is_inside(g, 1.0, -1.5)
    assert not O.is_inside(g, -1.0, 1.5)
    assert not O.is_inside(g, 1.0 + EPS, 1.0)
    assert O.is_inside(g, (2.0 + EPS) / 2.0, 1.0)
    assert not O.is_inside(g, 0.5, -1.5 - (2.0 * EPS))
    assert O.is_inside(g, -0.5, (3.0 + EPS) / 2.0)


def test_index_shift_from_position_shift():
    # GridMapMathTest.cpp:195-219
    import ctypes as C
    def shift(dx, dy, res):
        a, b = C.c_int(), C.c_int()
        O.lib().oracle_index_shift_from_position_shift(dx, dy, res, C.byref(a), C.byref(b))
        return a.value, b.value
    assert shift(0.0, 0.0, 1.0) == (0, 0)
    assert shift(0.35, -0.45, 1.0) == (0, 0)
    assert shift(0.55, -0.45, 1.0) == (-1, 0)
    assert shift(-1.3, -2.65, 1.0) == (1, 3)
    assert shift(-0.4, 0.09, 0.2) == (2, 0)


def test_line_iterator_end_outside_map():
    # LineIteratorTest.cpp:45-72: 8x5 map, res 1, (0,0)->(9,6): (4,2),(3,1),(2,1) then two more cells.
    g = O.make_geom(8.0, 5.0, 1.0, 0.0, 0.0)
    cells = O.line_cells(g, 0.0, 0.0, 9.0, 6.0)
    assert len(cells) == 5
    assert cells[:3].tolist() == [[4, 2], [3, 1], [2, 1]]


def test_line_iterator_start_and_end_outside_map():
    # LineIteratorTest.cpp:74-99
    g = O.make_geom(8.0, 5.0, 1.0, 0.0, 0.0)
    cells = O.line_cells(g, -7.0, -9.0, 8.0, 8.0)
    # three checked cells, then three more increments reach isPastEnd => nCells <= 5
    assert cells.tolist() == [[5, 4], [4, 3], [3, 2], [2, 1], [1, 0]]


def test_line_iterator_without_intersecting_map():
    # LineIteratorTest.cpp:101-109 (reference leaves the iterator uninitialised; defined as empty)
    g = O.make_geom(8.0, 5.0, 1.0, 0.0, 0.0)
    assert len(O.line_cells(g, -8.0, 8.0, 8.0, 8.0)) == 0


def test_line_iterator_start_outside_map_upstream_form():
    # Upstream form of LineIteratorTest.cpp:20-43 (the fork edited the geometry so its own expectations cannot
    # hold): 8x5 map, (2,2)->(0,0) => (2,0),(3,1),(4,2).
    g = O.make_geom(8.0, 5.0, 1.0, 0.0, 0.0)
    cells = O.line_cells(g, 2.0, 2.0, 0.0, 0.0)
    assert cells.tolist() == [[2, 0], [3, 1], [4, 2]]


def test_move_start_index_and_cleared_regions():
    # GridMapTest.cpp:57-85 (Move): 8.1 x 5.1 @1.0 -> 8x5; move to (-3,-2): startIndex (3,2); rows 0..2 and
    # cols 0..1 are reset (regions {(0,0),3x5} and {(0,0),8x2}); (3,2) and (7,4) stay valid.
    g = O.make_geom(8.1, 5.1, 1.0, 0.0, 0.0)
    assert (g.rows, g.cols) == (8, 5)
    layer = np.zeros((g.cols, g.rows), dtype=np.float32)
    assert O.move(g, [layer], -3.0, -2.0)
    assert (g.start0, g.start1) == (3, 2)
    lay = layer.T  # (row, col)
    valid = ~np.isnan(lay)
    assert not valid[0, 0]
    assert valid[3, 2]
    assert not valid[2, 2]
    assert not valid[3, 1]
    assert valid[7, 4]
    expect = np.ones((8, 5), dtype=bool)
    expect[0:3, :] = False
    expect[:, 0:2] = False
    assert (valid == expect).all()
    assert g.pos_x == -3.0 and g.pos_y == -2.0
