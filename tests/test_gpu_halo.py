"""GPU parity of POP_HaloUpdate / block geometry against the oracle (itself pinned against the
reference's halo unit tests, tests/test_oracle_halo.py): exact equality for every field location,
field kind, rank (2-d/3-d/4-d) and type (R8, R4, I4: the nine members of the generic interface,
mpi/POP_HaloMod.F90:79-89), on closed, cyclic and tripole boundaries."""
import numpy as np
import pytest

from parity import *  # noqa: F401,F403

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ns", [c.BNDY_CLOSED, c.BNDY_CYCLIC, c.BNDY_TRIPOLE])
@pytest.mark.parametrize("ew", [c.BNDY_CYCLIC, c.BNDY_CLOSED])
def test_halo_all_locations_kinds_ranks(ns, ew):
    if ns == c.BNDY_TRIPOLE and ew == c.BNDY_CLOSED:
        pytest.skip("tripole grids are east-west cyclic")
    NX, NY, KM, NT = 30, 31, 3, 2      # the reference fixture test/unit/halo/POP_DomainSizeMod.F90:29-36
    cfg = c.make_config(nx_global=NX, ny_global=NY, km=KM, nt=NT, ew_boundary_type=ew, ns_boundary_type=ns)
    o = Oracle(cfg)
    p = P.api.Pop(cfg)
    try:
        rng = np.random.default_rng(3)
        blk = P.config.PopBlock()
        i8, ig, jg = o.block_info(0)
        assert [p.blk.ib, p.blk.ie, p.blk.jb, p.blk.je] == i8[2:6]
        assert np.array_equal(np.ctypeslib.as_array(p.blk.i_glob, (p.nxb,)), ig)
        assert np.array_equal(np.ctypeslib.as_array(p.blk.j_glob, (p.nyb,)), jg)
        for loc in (c.LOC_CENTER, c.LOC_NECORNER, c.LOC_NFACE, c.LOC_EFACE):
            for kind in (c.KIND_SCALAR, c.KIND_VECTOR, c.KIND_ANGLE):
                for shape in ((), (KM,), (NT, KM)):
                    A = rng.standard_normal(shape + (o.nyb, o.nxb)) * 100.0
                    Ao = np.ascontiguousarray(A[None])
                    Ap = np.ascontiguousarray(A)
                    o.halo_array(Ao, loc, kind, 0.0)
                    p.halo_update(Ap, loc, kind, 0.0)
                    assert np.array_equal(Ao[0], Ap), (loc, kind, shape)
                I = rng.integers(-50, 50, (o.nyb, o.nxb)).astype(np.int32)
                Io, Ip = np.ascontiguousarray(I[None]), I.copy()
                o.halo_array(Io, loc, kind, 0)
                p.halo_update(Ip, loc, kind, 0)
                assert np.array_equal(Io[0], Ip), (loc, kind, "i4")
                # R4: a halo update only moves values (and on the tripole seam negates / averages pairs of them, exact
                # for these multiples of 1/8), so the R8 oracle on the same numbers is the checker
                for shape in ((), (KM,), (NT, KM)):
                    F = (rng.integers(-2000, 2000, shape + (o.nyb, o.nxb)) / 8.0).astype(np.float32)
                    Fo = np.ascontiguousarray(F.astype(np.float64)[None])
                    Fp = F.copy()
                    o.halo_array(Fo, loc, kind, 0.0)
                    p.halo_update(Fp, loc, kind, 0.0)
                    assert Fp.dtype == np.float32 and np.array_equal(Fo[0], Fp.astype(np.float64)), (loc, kind, shape, "r4")
                    if shape:
                        # integers average differently on the tripole seam: level by level through the 2-d I4 oracle
                        J = rng.integers(-50, 50, shape + (o.nyb, o.nxb)).astype(np.int32)
                        Jp = J.copy()
                        p.halo_update(Jp, loc, kind, 0)
                        for lev, got in zip(J.reshape(-1, o.nyb, o.nxb), Jp.reshape(-1, o.nyb, o.nxb)):
                            Jo = np.ascontiguousarray(lev[None])
                            o.halo_array(Jo, loc, kind, 0)
                            assert np.array_equal(Jo[0], got), (loc, kind, shape, "i4")
        # idempotence
        A = np.ascontiguousarray(rng.standard_normal((o.nyb, o.nxb)))
        p.halo_update(A, c.LOC_CENTER, c.KIND_SCALAR)
        B = A.copy()
        p.halo_update(B, c.LOC_CENTER, c.KIND_SCALAR)
        assert np.array_equal(A, B)
    finally:
        p.finalize()
