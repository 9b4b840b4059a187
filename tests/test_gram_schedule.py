"""The work-unit schedule of the warp-specialised Gram kernel, walked on the host (ob_debug_gram_schedule: no device):
for every problem shape each (group, panel, column tile, row segment) must be computed exactly once, the row segments
must tile the padded rows, CTAs must get equal shares (+-1) of every cost class, and a partly filled last panel must be
the only one with fewer than 16 slot groups.  Also under row sharding (world > 1): the ranks' segments tile the group."""
import ctypes as C
from collections import Counter

import numpy as np
import pytest


def schedule(K, na, nb, slots, world=1, rank=0, grid=148):
    from oaxaca_blinder_rs_b200 import _native
    _native.build()
    L = _native.lib()
    n = L.ob_debug_gram_schedule(K, na, nb, slots, world, rank, grid, None, 0)
    assert n >= 0
    out = np.zeros((max(n, 1), 8), dtype=np.int64)
    m = L.ob_debug_gram_schedule(K, na, nb, slots, world, rank, grid, out.ctypes.data_as(C.POINTER(C.c_int64)), n)
    assert m == n
    return out[:n]


SHAPES = [  # K, n_a, n_b, slots
    (51, 5_000_000, 5_000_000, 2001),      # config 3
    (21, 500_123, 499_877, 1001),          # config 2
    (31, 2_500_000, 2_500_000, 1001),      # config 4: 4 wide tiles + a one-quantum tail tile
    (17, 50_000_000, 50_000_000, 10001),   # config 5
    (6, 5000, 5000, 501), (2, 3, 4, 1), (3, 40, 33, 129), (9, 100_000, 17, 128), (90, 70_000, 70_001, 257),
]


@pytest.mark.parametrize("K,na,nb,slots", SHAPES)
def test_every_unit_once_and_balanced(K, na, nb, slots):
    u = schedule(K, na, nb, slots)
    V = K + 1
    pairs = V * (V + 1) // 2
    quanta = -(-pairs // 32)                 # column quantum: 32 columns = one DMMA sub-tile per scheduler
    nfull, tail_q = divmod(quanta, 4)        # tiles of 128 columns + a tail tile of one to three quanta
    ntiles = nfull + (tail_q > 0)
    panels = -(-slots // 128)
    keys = Counter(map(tuple, u[:, 1:5]))
    assert all(c == 1 for c in keys.values())
    segs = [int(u[u[:, 1] == g][:, 4].max()) + 1 if (u[:, 1] == g).any() else 0 for g in (0, 1)]
    assert len(keys) == (segs[0] + segs[1]) * panels * ntiles
    # segments tile the padded rows of each group: sum of stages * 32 over the segments of one (panel, tile)
    for g, n in ((0, na), (1, nb)):
        sel = u[(u[:, 1] == g) & (u[:, 2] == 0) & (u[:, 3] == 0)]
        assert sel[:, 5].sum() * 32 == max(32, -(-n // 32) * 32) and (sel[:, 5] > 0).all()
    # tail panel: only the last panel may have fewer slot groups, rounded up to 4 groups
    last = slots - (panels - 1) * 128
    mi = min(16, -(-(-(-last // 8)) // 4) * 4)
    assert set(u[u[:, 2] == panels - 1][:, 6]) == {mi} and (panels == 1 or set(u[u[:, 2] < panels - 1][:, 6]) == {16})
    assert set(u[u[:, 3] == nfull][:, 7]) <= {tail_q} and set(u[u[:, 3] < nfull][:, 7]) <= {0}
    # equal shares per cost class (tail tile?, tail panel?) across CTAs
    grid = min(148, len(u))
    for half in (0, 1):
        for tail in (0, 1):
            cls = u[((u[:, 7] > 0) == bool(half)) & ((u[:, 6] < 16) == bool(tail))]
            per = np.bincount(cls[:, 0], minlength=grid)
            assert per.max() - per.min() <= 2, (half, tail, per.max(), per.min())     # +-1, plus the class boundary


def test_row_shards_tile_the_group():
    K, na, nb, slots = 17, 1_000_003, 999_999, 300
    one = schedule(K, na, nb, slots)
    tot = {g: one[(one[:, 1] == g) & (one[:, 2] == 0) & (one[:, 3] == 0)][:, 5].sum() for g in (0, 1)}
    for world in (2, 4, 8):
        parts = [schedule(K, na, nb, slots, world, r) for r in range(world)]
        for g in (0, 1):
            got = sum(p[(p[:, 1] == g) & (p[:, 2] == 0) & (p[:, 3] == 0)][:, 5].sum() for p in parts)
            assert got == tot[g], (world, g)


def columns(K, T, n_cont, cat_levels):
    from oaxaca_blinder_rs_b200 import _native
    _native.build()
    L = _native.lib()
    lv = np.asarray(cat_levels, dtype=np.int32)
    lvp = lv.ctypes.data_as(C.POINTER(C.c_int32)) if len(lv) else None
    L.ob_debug_gram_columns.restype = C.c_int64
    L.ob_debug_gram_columns.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_uint16),
                                        C.POINTER(C.c_int32), C.c_int64, C.POINTER(C.c_int32)]
    til = np.zeros(3, dtype=np.int32)
    n = L.ob_debug_gram_columns(K, T, n_cont, lvp, len(lv), None, None, 0, til.ctypes.data_as(C.POINTER(C.c_int32)))
    assert n > 0
    pairs = np.zeros((n, 2), dtype=np.uint16)
    cmap = np.zeros(n, dtype=np.int32)
    m = L.ob_debug_gram_columns(K, T, n_cont, lvp, len(lv), pairs.ctypes.data_as(C.POINTER(C.c_uint16)),
                                cmap.ctypes.data_as(C.POINTER(C.c_int32)), n, None)
    assert m == n
    return pairs, cmap, til


def upper_triangle_index(K, T, j, l):
    """Row-major upper triangle of [x|y][x|y]^T as the solve reads it (internal.h: pair_base)."""
    return j * (K + T) - j * (j - 1) // 2 + (l - j)


@pytest.mark.parametrize("K,T,n_cont,levels", [
    (51, 1, 44, (4, 4)),          # config 3: six structural zeros -> 1372 columns = 10 tiles + 3 quanta
    (51, 1, 50, ()),              # no categoricals: nothing dropped, 11 tiles
    (31, 1, 30, ()),              # config 4: 4 tiles + 1 quantum
    (31, 3, 30, ()),              # config 4, three RIF outcomes in one pass
    (17, 1, 16, ()),              # config 5: 1 tile + 2 quanta
    (6, 1, 2, (4,)), (12, 2, 3, (3, 2, 6)), (9, 1, 8, ()), (2, 1, 1, ()), (5, 1, 1, (4,)),
    (8, 1, 3, (4,)),              # K does not match 1 + n_cont + dummies: the structure is ignored
])
def test_gram_columns_cover_the_upper_triangle(K, T, n_cont, levels):
    pairs, cmap, til = columns(K, T, n_cont, levels)
    known = 1 + n_cont + sum(m - 1 for m in levels) == K
    block = {}
    if known:
        c = 1 + n_cont
        for q, m in enumerate(levels):
            for _ in range(m - 1):
                block[c] = q
                c += 1
    zeros = {(j, l) for j in block for l in block if j < l and block[j] == block[l]}
    P1 = K * (K + 1) // 2 + K * T + 1
    expect = {}
    for j in range(K):
        for l in range(j, K + T):
            if (j, l) not in zeros:
                expect[upper_triangle_index(K, T, j, l)] = (j, l)
    expect[P1 - 1] = (K, K)
    real = cmap >= 0
    assert int(real.sum()) == len(expect) == til[0] == P1 - len(zeros)
    assert len(set(cmap[real].tolist())) == int(real.sum())                      # every cell at most once
    for c in np.nonzero(real)[0]:
        assert expect[int(cmap[c])] == (int(pairs[c, 0]), int(pairs[c, 1])), c     # the product feeds the right cell
    assert real[:til[0]].all() and not real[til[0]:].any()                       # computed columns first, then padding
    assert (pairs[~real] == K).all()                                             # padding = harmless (y_0, y_0)
    quanta = -(-int(til[0]) // 32)
    assert (til[1], til[2]) == divmod(quanta, 4) and len(cmap) == (til[1] + (til[2] > 0)) * 128
    if (K, levels) == (51, (4, 4)):
        assert tuple(til) == (1372, 10, 3)
