"""CPU checks of the oracle's EVP block preconditioner against what the reference itself pins:
(vi) of SURVEY 8c -- the self-test bound of ExplicitBlockEvpPre (POP_SolversMod.F90:2594-2614: max |rinv*rin - I| <=
1e-8) -- the sub-block partition rule (EvpBlockPartition :2992-3040), and the defining property of the method: on a
sub-block without land it solves the five-point system (centre + corner weights, zero frame) exactly."""
import ctypes as C

import numpy as np

from parity import *  # noqa: F401,F403


def _partition(m, mm=8):
    """EvpBlockPartition restated from its description: mb = (m-3)/mm + 1 sub-blocks, regular starts 2 + i*mm, the last
    two share the remainder"""
    mb = (m - 3) // mm + 1
    mdi = [0] * (mb + 2)
    mdi[1] = 2
    if mb == 1:
        mdi[mb + 1] = m
    else:
        for i in range(1, mb - 1):
            mdi[i + 1] = 2 + i * mm
        mdi[mb] = (mdi[mb - 1] + m) // 2
        mdi[mb + 1] = m
    return mdi[1:]


def test_partition_rule_examples():
    # tx0.1v3 single block: nx_block - 2 = 3602 -> 450 sub-blocks of 8 columns, the last two 3586|3594|3602
    p = _partition(3602)
    assert len(p) - 1 == 450 and p[-3:] == [3586, 3594, 3602]
    assert all(b - a == 8 for a, b in zip(p[:-1], p[1:]))
    # a 20-cell block: 8 + 6 + 6
    assert _partition(22) == [2, 10, 16, 22]
    assert _partition(10) == [2, 10]
    # no sub-block is wider than 8 or narrower than 4 for any block size
    for m in range(7, 400):
        w = np.diff(_partition(m))
        assert w.max() <= 8 and (w.min() >= 4 or len(w) == 1), (m, w)


def test_selfcheck_bound_and_exact_inversion():
    cs = make_case(100, 52, 6, seed=81, ns=c.BNDY_TRIPOLE, given_vmix=True, solver_choice=c.SOLVER_CHRONGEAR,
                   dtt=7200.0, preconditioner_choice=c.PRECOND_EVP)
    o = load_oracle(cs, block_size=(50, 26))
    assert o.solvers_prep() == 0
    o.L.oracle_evp_selfcheck.restype = C.c_double
    assert 0.0 < o.L.oracle_evp_selfcheck() <= 1.0e-8          # the reference's own acceptance test
    ns, nl = C.c_int(), C.c_int()
    assert o.L.oracle_evp_counts(C.byref(ns), C.byref(nl)) == 0
    assert 0 < nl.value < ns.value
    # P = (PC) R solves  cc*P(i,j) + ne(i,j)*P(i+1,j+1) + ne(i,j-1)*P(i+1,j-1) + ne(i-1,j)*P(i-1,j+1) + ne(i-1,j-1)*P(i-1,j-1)
    # = R on every sub-block without land (zero outside the sub-block), and is the diagonal elsewhere
    rng = np.random.default_rng(3)
    R = np.ascontiguousarray(rng.standard_normal((o.nblocks, o.nyb, o.nxb)))
    PX = np.zeros_like(R)
    f = osig(o.L, "o_preconditioner", [C.c_void_p, C.c_void_p, C.c_int])
    for b in range(o.nblocks):
        f(op(PX), op(R), b)
    cc, ne = o.view("btropWgtCenter", 1), o.view("btropWgtNE", 1)
    xs, ys = _partition(o.nxb - 2), _partition(o.nyb - 2)
    checked = 0
    for b in range(o.nblocks):
        for j0, j1 in zip(ys[:-1], ys[1:]):
            for i0, i1 in zip(xs[:-1], xs[1:]):
                # interior array cells (0-based): rows j0 .. j1-1, columns i0 .. i1-1
                J, I = slice(j0, j1), slice(i0, i1)
                if np.any(ne[b, J, I] == 0.0):
                    with np.errstate(divide="ignore"):
                        icc = np.where(cc[b, J, I] != 0, 1.0 / cc[b, J, I], 0.0)
                    np.testing.assert_array_equal(PX[b, J, I], R[b, J, I] * icc)
                    continue
                Pz = np.zeros((o.nyb, o.nxb))
                Pz[J, I] = PX[b, J, I]
                lhs = (cc[b, J, I] * Pz[J, I] + ne[b, J, I] * Pz[j0 + 1:j1 + 1, i0 + 1:i1 + 1]
                       + ne[b, j0 - 1:j1 - 1, i0:i1] * Pz[j0 - 1:j1 - 1, i0 + 1:i1 + 1]
                       + ne[b, j0:j1, i0 - 1:i1 - 1] * Pz[j0 + 1:j1 + 1, i0 - 1:i1 - 1]
                       + ne[b, j0 - 1:j1 - 1, i0 - 1:i1 - 1] * Pz[j0 - 1:j1 - 1, i0 - 1:i1 - 1])
                assert np.max(np.abs(lhs - R[b, J, I])) <= 1.0e-7 * np.max(np.abs(R[b, J, I])), (b, i0, j0)
                checked += 1
    assert checked > 10
    # ghost cells and everything outside the sub-block interiors stay zero
    assert not PX[:, :2, :].any() and not PX[:, :, :2].any()
