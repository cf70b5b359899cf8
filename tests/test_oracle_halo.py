"""Pins the oracle's POP_HaloUpdate against the reference's own halo unit tests:
test/unit/halo/POP.F90Dipole:136-147,277-292 and POP.F90Tripole:295-346 (Center), :565-612 (EFace),
:818-874 (NFace), :1079-1148 (NECorner), :1448-1492 / :1702-1751 (vector sign flip), on the
fixture test/unit/halo/POP_DomainSizeMod.F90:29-36 (30x31x3, nt=2, 4x4 blocks).
The expected-value rules below are restated from those test programs."""
import numpy as np
import pytest

from oracle.oracle import Oracle, P

c = P.config
NX, NY, KM, NT, BS = 30, 31, 3, 2, 4
NXB = BS + 4


def global_pattern(scale):
    """i4ArrayG: (i+j)*scale with a rectangular land hole of zeros (POP.F90Dipole:136-147)."""
    i = np.arange(1, NX + 1)[None, :]
    j = np.arange(1, NY + 1)[:, None]
    ocean = (i > NX // 2 + NXB) | (i < NX // 2 - NXB) | (j > NY // 2 + NXB) | (j < NY // 2 - NXB)
    return np.where(ocean, (i + j) * scale, 0).astype(np.float64)


def make(ns):
    cfg = c.make_config(nx_global=NX, ny_global=NY, km=KM, nt=NT, block_size_x=BS, block_size_y=BS,
                        ew_boundary_type=c.BNDY_CYCLIC, ns_boundary_type=ns)
    o = Oracle(cfg)
    G = global_pattern(100)
    # work per block; land blocks are eliminated as the rake distribution does (Dipole:152-191)
    active = []
    info = []
    for b in range(o.nblocks):
        i8, ig, jg = o.block_info(b)
        ib, ie, jb, je = i8[2:6]
        w = 0
        for j in range(jb, je + 1):
            for i in range(ib, ie + 1):
                if ig[i - 1] > 0 and jg[j - 1] > 0 and G[jg[j - 1] - 1, ig[i - 1] - 1] != 0:
                    w += 1
        active.append(1 if w > 0 else 0)
        info.append((i8, ig, jg))
    assert 0 in active, "fixture must contain at least one eliminated land block"
    o.set_active(active)
    return o, G, active, info


def orig_blocks(o, G, active, info):
    """i4Orig: global values looked up through iGlobal/jGlobal, 0 outside the domain."""
    A = np.zeros((o.nblocks, o.nyb, o.nxb))
    for b, (i8, ig, jg) in enumerate(info):
        if not active[b]:
            continue
        for j in range(o.nyb):
            for i in range(o.nxb):
                if ig[i] > 0 and jg[j] > 0:
                    A[b, j, i] = G[jg[j] - 1, ig[i] - 1]
    return A


def scattered(o, G, active, info):
    """what POP_RedistributeScatter gives: physical cells, ghost cells zero."""
    A = np.zeros((o.nblocks, o.nyb, o.nxb))
    for b, (i8, ig, jg) in enumerate(info):
        if not active[b]:
            continue
        ib, ie, jb, je = i8[2:6]
        for j in range(jb, je + 1):
            for i in range(ib, ie + 1):
                A[b, j - 1, i - 1] = G[jg[j - 1] - 1, ig[i - 1] - 1]
    return A


def test_dipole_2d_3d_4d():
    o, G, active, info = make(c.BNDY_CLOSED)
    exp = orig_blocks(o, G, active, info)
    A = scattered(o, G, active, info)
    o.halo_array(A, c.LOC_CENTER, c.KIND_SCALAR, 0.0)
    act = np.array(active, bool)
    assert np.array_equal(A[act], exp[act])
    # 3-d / 4-d: value*k, value*k*l (Dipole:294-300)
    A3 = np.ascontiguousarray(scattered(o, G, active, info)[:, None] * np.arange(1, KM + 1)[None, :, None, None])
    o.halo_array(A3, c.LOC_CENTER, c.KIND_SCALAR, 0.0)
    assert np.array_equal(A3[act], (exp[:, None] * np.arange(1, KM + 1)[None, :, None, None])[act])
    kl = (np.arange(1, NT + 1)[:, None] * np.arange(1, KM + 1)[None, :])[None, :, :, None, None]
    A4 = np.ascontiguousarray(scattered(o, G, active, info)[:, None, None] * kl)
    o.halo_array(A4, c.LOC_CENTER, c.KIND_SCALAR, 0.0)
    assert np.array_equal(A4[act], (exp[:, None, None] * kl)[act])
    # integer variant
    Ai = scattered(o, G, active, info).astype(np.int32)
    o.halo_array(Ai, c.LOC_CENTER, c.KIND_SCALAR, 0)
    assert np.array_equal(Ai[act], exp[act].astype(np.int32))


def tripole_expect(o, G, active, info, loc, kind):
    exp = orig_blocks(o, G, active, info)
    sgn = 1.0 if kind == c.KIND_SCALAR else -1.0
    for b, (i8, ig, jg) in enumerate(info):
        if not active[b]:
            continue
        je = i8[5]
        if not (je + 1 <= o.nyb and jg[je] < 0):
            continue
        for i in range(1, o.nxb + 1):
            g = ig[i - 1]
            for j in (1, 2):
                if loc == c.LOC_CENTER:                      # Tripole:333-342 / 1474-1486
                    jgl = NY + 1 - j
                    if 0 < g <= NX:
                        exp[b, je + j - 1, i - 1] = sgn * G[jgl - 1, NX - g + 1 - 1]
                elif loc == c.LOC_EFACE:                     # :595-610 / 1733-1748
                    jgl = NY + 1 - j
                    if 0 < g < NX and g != NX // 2:
                        exp[b, je + j - 1, i - 1] = sgn * G[jgl - 1, NX - g - 1]
                    elif g == NX // 2:
                        exp[b, je + j - 1, i - 1] = sgn * G[jgl - 1, NX // 2 - 1]
                    elif g == NX:
                        exp[b, je + j - 1, i - 1] = sgn * G[jgl - 1, NX - 1]
                    else:
                        exp[b, je + j - 1, i - 1] = 0
                elif loc == c.LOC_NFACE:                     # :855-862
                    jgl = NY - j
                    if g != 0:
                        exp[b, je + j - 1, i - 1] = sgn * G[jgl - 1, NX - g + 1 - 1]
                elif loc == c.LOC_NECORNER:                  # :1112-1124
                    jgl = NY - j
                    if g > 0 and g != NX and g != NX // 2:
                        exp[b, je + j - 1, i - 1] = sgn * G[jgl - 1, NX - g - 1]
                    elif g == NX // 2:
                        exp[b, je + j - 1, i - 1] = sgn * G[jgl - 1, NX // 2 - 1]
                    elif g == NX:
                        exp[b, je + j - 1, i - 1] = sgn * G[jgl - 1, NX - 1]
                    else:
                        exp[b, je + j - 1, i - 1] = 0
            # symmetry enforcement on the top physical row (scalar cases pinned by the reference)
            if kind == c.KIND_SCALAR:
                if loc == c.LOC_NFACE and g != 0:            # :868-873
                    exp[b, je - 1, i - 1] = 0.5 * (G[NY - 1, g - 1] + G[NY - 1, NX - g + 1 - 1])
                if loc == c.LOC_NECORNER:                    # :1131-1146
                    if 0 < g < NX and g != NX // 2:
                        exp[b, je - 1, i - 1] = 0.5 * (G[NY - 1, g - 1] + G[NY - 1, NX - g - 1])
                    elif g == NX or g == NX // 2:
                        exp[b, je - 1, i - 1] = G[NY - 1, g - 1]
                    else:
                        exp[b, je - 1, i - 1] = 0
    return exp


@pytest.mark.parametrize("loc,kind", [
    (c.LOC_CENTER, c.KIND_SCALAR), (c.LOC_EFACE, c.KIND_SCALAR), (c.LOC_NFACE, c.KIND_SCALAR),
    (c.LOC_NECORNER, c.KIND_SCALAR), (c.LOC_CENTER, c.KIND_VECTOR), (c.LOC_EFACE, c.KIND_VECTOR)])
def test_tripole(loc, kind):
    o, G, active, info = make(c.BNDY_TRIPOLE)
    exp = tripole_expect(o, G, active, info, loc, kind)
    A = orig_blocks(o, G, active, info)     # the reference test starts from i4Orig
    o.halo_array(A, loc, kind, 0.0)
    act = np.array(active, bool)
    # rows/cols inside ie+2 / je+2 only (padding is never addressed)
    for b in np.nonzero(act)[0]:
        i8 = info[b][0]
        ie, je = i8[3], i8[5]
        assert np.array_equal(A[b, : je + 2, : ie + 2], exp[b, : je + 2, : ie + 2]), (b, loc, kind)
    A3 = np.ascontiguousarray(orig_blocks(o, G, active, info)[:, None] * np.arange(1, KM + 1)[None, :, None, None])
    o.halo_array(A3, loc, kind, 0.0)
    for b in np.nonzero(act)[0]:
        ie, je = info[b][0][3], info[b][0][5]
        for k in range(KM):
            assert np.array_equal(A3[b, k, : je + 2, : ie + 2], exp[b, : je + 2, : ie + 2] * (k + 1))


def test_halo_idempotent():
    o, G, active, info = make(c.BNDY_TRIPOLE)
    rng = np.random.default_rng(1)
    A = rng.standard_normal((o.nblocks, o.nyb, o.nxb))
    o.halo_array(A, c.LOC_NECORNER, c.KIND_VECTOR, 0.0)
    B = A.copy()
    o.halo_array(B, c.LOC_NECORNER, c.KIND_VECTOR, 0.0)
    # a second update only re-applies the (idempotent after one pass) top-row symmetrisation
    o.halo_array(A, c.LOC_NECORNER, c.KIND_VECTOR, 0.0)
    assert np.array_equal(A, B)
