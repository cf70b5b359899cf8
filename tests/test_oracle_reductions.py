"""Pins the oracle's POP_GlobalSum / distribution / EOS against what the reference's own tests
and comments hold: test/unit/reduction/POP.F90:120-136,226-260 (34x34, 4x4 blocks, (i+j)*1000
pattern with a land hole, exact equality with the serial sum of the global array);
test/unit/blockDistribution (320x384, 32x24 blocks, 153 procs, cartesian); the MWJF known answer
source/state_mod.F90:785-787."""
import ctypes as C

import numpy as np

from oracle.oracle import Oracle, P, _p

c = P.config


def _mk(nx, ny, bs, ns=c.BNDY_CLOSED, km=3):
    cfg = c.make_config(nx_global=nx, ny_global=ny, km=km, nt=2, block_size_x=bs, block_size_y=bs,
                        ew_boundary_type=c.BNDY_CYCLIC, ns_boundary_type=ns)
    return Oracle(cfg)


def _pattern(nx, ny, nxb):
    i = np.arange(1, nx + 1)[None, :]
    j = np.arange(1, ny + 1)[:, None]
    ocean = (i > nx // 2 + nxb) | (i < nx // 2 - nxb) | (j > ny // 2 + nxb) | (j < ny // 2 - nxb)
    return np.where(ocean, (i + j) * 1000.0, 0.0)


def _scatter(o, G):
    A = np.zeros((o.nblocks, o.nyb, o.nxb))
    for b in range(o.nblocks):
        i8, ig, jg = o.block_info(b)
        for j in range(i8[4], i8[5] + 1):
            for i in range(i8[2], i8[3] + 1):
                A[b, j - 1, i - 1] = G[jg[j - 1] - 1, ig[i - 1] - 1]
    return A


def test_global_sum_matches_serial_sum():
    o = _mk(34, 34, 4)
    G = _pattern(34, 34, 8)
    A = _scatter(o, G)
    # ghost cells must not contribute: poison them
    B = A.copy()
    for b in range(o.nblocks):
        i8 = o.block_info(b)[0]
        m = np.ones((o.nyb, o.nxb), bool)
        m[i8[4] - 1:i8[5], i8[2] - 1:i8[3]] = False
        B[b][m] = 7777.0
    assert o.global_sum(B, c.LOC_CENTER) == G.sum()
    mask = (A != 0).astype(np.float64)                 # mask = (i4Array2D /= 0), POP.F90:218
    assert o.global_sum(B, c.LOC_CENTER, mask) == G.sum()
    half = (np.arange(o.nxb)[None, None, :] % 2 == 0) * np.ones_like(A)
    exp = sum(A[b][i8[4] - 1:i8[5], i8[2] - 1:i8[3]][:, :].__mul__(half[b][i8[4] - 1:i8[5], i8[2] - 1:i8[3]]).sum()
              for b in range(o.nblocks) for i8 in [o.block_info(b)[0]])
    assert o.global_sum(B, c.LOC_CENTER, half) == exp


def test_global_sum_tripole_dedup():
    """Nface/NEcorner fields on tripole blocks: the redundant half (iGlobal > nx/2) of the top
    physical row is subtracted (mpi/POP_ReductionsMod.F90:312-341)."""
    nx, ny = 32, 24
    o = _mk(nx, ny, 8, ns=c.BNDY_TRIPOLE)
    rng = np.random.default_rng(0)
    G = np.rint(rng.uniform(0, 100, (ny, nx)))
    A = _scatter(o, G)
    assert o.global_sum(A, c.LOC_CENTER) == G.sum()
    assert o.global_sum(A, c.LOC_NFACE) == G.sum() - G[-1, nx // 2:].sum()
    assert o.global_sum(A, c.LOC_NECORNER) == G.sum() - G[-1, nx // 2:].sum()


def test_cartesian_distribution_153_procs():
    """POP_DistributionCreateCartesian on the blockDistribution fixture (10 x 16 blocks, 153 procs):
    nint(sqrt(153))=12 -> first factor pair found walking down is 9 x 17; neither orientation
    divides the 10 x 16 block grid, so 9 x 17 is kept (POP_DistributionMod.F90:1655-1720)."""
    from oracle.oracle import lib
    L = lib()
    nbx, nby = 320 // 32, 384 // 24
    work = np.ones(nbx * nby, np.int32)
    work[5] = 0                                            # one land block -> dropped
    loc = np.zeros(nbx * nby, np.int32)
    rc = L.oracle_distribution_cartesian(153, nbx, nby, _p(work), _p(loc))
    px, py = rc // 1000, rc % 1000
    assert (px, py) == (9, 17)
    assert loc[5] == 0
    bxp, byp = (nbx - 1) // px + 1, (nby - 1) // py + 1
    for jb in range(nby):
        for ib in range(nbx):
            n = jb * nbx + ib
            if n == 5:
                continue
            assert loc[n] == (jb // byp) * px + (ib // bxp) + 1
    # 1 x P strips used by the product: nbx=1 blocks, P procs -> P x 1? must be 1 x P
    for Pn in (1, 2, 4, 8):
        w = np.ones(Pn, np.int32)
        l2 = np.zeros(Pn, np.int32)
        rc = L.oracle_distribution_cartesian(Pn, 1, Pn, _p(w), _p(l2))
        assert (rc // 1000, rc % 1000) == (1, Pn)
        assert list(l2) == list(range(1, Pn + 1))


def test_mwjf_known_answer():
    """MWJF test value rho = 1.033213387 for S = 35.0 PSU, theta = 20.0, pressz = 200.0
    (source/state_mod.F90:785-787, the published McDougall et al. 2003 check value).  The older
    comment at state_mod.F90:413-414 quotes 1.033213242, which the coefficients tabulated in the
    same file (:136-165) do not reproduce (they give 1.0332133866); the :786 value is the pin."""
    o = _mk(8, 8, 8, km=3)
    nx = ny = 8
    S = P.synthetic
    o.set_grid(S.horiz_grid(nx, ny), np.full((ny, nx), 3, np.int32), S.vert_grid("uniform", 3))
    pz = o.vec("pressz", 5)
    pz[2] = 200.0                                          # pressz(kk=2) = 200 bar
    T = np.full((o.nyb, o.nxb), 20.0)
    Sl = np.full((o.nyb, o.nxb), 0.035)
    R = np.zeros_like(T)
    o.L.o_state.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 4
    o.L.o_state(2, 2, _p(T), _p(Sl), 0, None, _p(R), None, None)
    assert abs(R[0, 0] - 1.033213387) < 5e-10
    # derivatives against centred finite differences
    DT, DS = np.zeros_like(T), np.zeros_like(T)
    o.L.o_state(2, 2, _p(T), _p(Sl), 0, None, None, _p(DT), _p(DS))
    h = 1e-4
    Rp, Rm = np.zeros_like(T), np.zeros_like(T)
    o.L.o_state(2, 2, _p(T + h), _p(Sl), 0, None, _p(Rp), None, None)
    o.L.o_state(2, 2, _p(T - h), _p(Sl), 0, None, _p(Rm), None, None)
    assert abs((Rp[0, 0] - Rm[0, 0]) / (2 * h) - DT[0, 0]) < 1e-9
    hs = 1e-7
    o.L.o_state(2, 2, _p(T), _p(Sl + hs), 0, None, _p(Rp), None, None)
    o.L.o_state(2, 2, _p(T), _p(Sl - hs), 0, None, _p(Rm), None, None)
    assert abs((Rp[0, 0] - Rm[0, 0]) / (2 * hs) - DS[0, 0]) < 1e-6
