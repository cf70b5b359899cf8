"""Worker of tests/test_gpu_multirank.py (one process per GPU, launched with torch.distributed.run):
the same seeded case is advanced with one strip per rank (NCCL / peer-memory halos, rank-ordered
double-double sums) and, on rank 0, with a single strip; the gathered fields must be bitwise equal and
the solver iteration counts identical (decomposition independence)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity import *  # noqa: F401,F403,E402

FIELDS_CMP = ("TRACER", "UVEL", "VVEL", "RHO", "PSURF", "UBTROP", "VBTROP", "GRADPX", "GRADPY")


def run(cs, cfg, comm_id, steps, coupled=False):
    p = load_pop(cs, cfg, comm_id)
    try:
        its = []
        rng = np.random.default_rng(17)
        for ts in steps:
            if coupled:
                # the coupler's entry point: every rank hands over the physical cells of its strip of the same global forcing
                r = p.rows()
                ocean, oceanu = (cs.kmt > 0), (cs.kmu > 0)
                f = [np.ascontiguousarray((1.0e-4 * rng.standard_normal((cs.nt, cs.ny, cs.nx)) * ocean)[:, r, :]),
                     np.ascontiguousarray((0.5 * rng.standard_normal((2, cs.ny, cs.nx)) * oceanu)[:, r, :]),
                     np.ascontiguousarray((1.0e-3 * rng.standard_normal((cs.ny, cs.nx)) * ocean)[r, :]),
                     np.ascontiguousarray((1.0e-6 * rng.standard_normal((cs.ny, cs.nx)) * ocean)[r, :])]
                # pinned buffers: STF is read and U1/V1 are stored by the kernels directly (the path bench.py times)
                import torch
                pin = [torch.from_numpy(x).pin_memory() for x in f]
                out_t = torch.zeros((5, r.stop - r.start, cs.nx), dtype=torch.float64).pin_memory()
                p.step_coupled(ts, pin[0].numpy(), pin[1].numpy(), pin[2].numpy(), pin[3].numpy(), out_t.numpy())
                sfc = out_t.numpy().copy()
            else:
                p.step(ts)
            its.append(p.solvers_get_diagnostics()[0])
        out = {n: p.gather(n, c.TIME_CUR) for n in FIELDS_CMP}
        if coupled:
            out["SFC_OUT"] = sfc   # what the coupler got back from the last step: SST, SSS, PSURF, U1, V1
            km = cs.km
            for idx, (name, lev) in enumerate((("TRACER", 0), ("TRACER", km), ("PSURF", 0), ("UVEL", 0), ("VVEL", 0))):
                if not np.array_equal(sfc[idx], out[name][lev]):
                    print("FAIL rank %d: surface output %d differs from the resident %s" % (cfg.rank, idx, name))
                    out["SFC_OUT"] = sfc * np.nan
        return its, out, p.rows()
    finally:
        p.finalize()


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    which = sys.argv[1] if len(sys.argv) > 1 else "pcsi"
    # (8 strips: 16 rows each, so that the deep-strip layout of the P-CSI passes -- 12 ghost rows -- is in play there too)
    kw = dict(nx=96, ny=64 if world < 8 else 128, km=6, seed=81)
    coupled = (which == "coupled")   # pop_step_coupled with host forcing strips (FW needs its ghost cells across the strip cut)
    if coupled:
        which = "pcsi"
    if which in ("pcsi22", "pcsi_plain"):
        # the P-CSI passes run on deep strips by default (12 ghost rows); also 22 rows deep, and the plain layout
        os.environ.update({"POP_B200_DEEP_HALO": "22"} if which == "pcsi22" else {"POP_B200_NO_DEEP_HALO": "1"})
        which = "pcsi"
    if which == "pcsi":
        kw.update(ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_DEL4, hmix_momentum_itype=c.HMIX_DEL4,
                  lvariable_hmixt=1, lvariable_hmixu=1, ah=-3.0e21, am=-27.0e21, given_vmix=True,
                  solver_choice=c.SOLVER_PCSI, dtt=600.0)
    elif which == "chrongear":
        kw.update(convergence_criterion=1e-12)
    elif which == "cyclic_pcg":
        kw.update(ns=c.BNDY_CYCLIC, solver_choice=c.SOLVER_PCG, tadvect=c.TADVECT_UPWIND3, nt=3)
    elif which == "gm":
        kw.update(ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_GM, ah_bolus=0.5e7, slm_b=0.2, given_vmix=True, nt=3,
                  solver_choice=c.SOLVER_PCSI, dtt=1800.0)
    elif which == "pbc":   # partial bottom cells, the production variant of config 4
        kw.update(ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_DEL4, hmix_momentum_itype=c.HMIX_DEL4,
                  lvariable_hmixt=1, lvariable_hmixu=1, ah=-3.0e21, am=-27.0e21, given_vmix=True,
                  solver_choice=c.SOLVER_PCSI, dtt=600.0, partial_bottom_cells=1)
    elif which == "lwlim":  # lw_lim advection: flux velocities two cells out cross the strip boundaries (comp_flux_vel_ghost)
        kw.update(ns=c.BNDY_TRIPOLE, nt=4, given_vmix=True, solver_choice=c.SOLVER_PCSI, dtt=1800.0,
                  tadvect=[c.TADVECT_CENTERED, c.TADVECT_CENTERED, c.TADVECT_LW_LIM, c.TADVECT_LW_LIM])
    elif which == "evp":   # EVP is block-local: the strips are compared with the oracle run on the same 1 x P blocks
        kw.update(ns=c.BNDY_TRIPOLE, given_vmix=True, solver_choice=c.SOLVER_PCSI, dtt=7200.0,
                  preconditioner_choice=c.PRECOND_EVP, max_lanczos_step=100, lanczos_convergence_criterion=0.15)
    cs = make_case(kw.pop("nx"), kw.pop("ny"), kw.pop("km"), **kw)
    steps = [c.TS_EULER, c.TS_LEAPFROG, c.TS_AVG, c.TS_LEAPFROG, c.TS_ROBERT]
    if which == "evp":
        steps = [c.TS_EULER, c.TS_LEAPFROG, c.TS_LEAPFROG]
    ref = None
    if rank == 0 and which == "evp":
        # reference = the oracle in its REPRODUCIBLE (r16-sum) build on the same decomposition: bit-identical expected
        o = load_oracle(cs, block_size=(cs.nx, cs.ny // world), reproducible=True)
        its = []
        for ts in steps:
            assert o.step(ts) == 0
            its.append(o.solver_diag()[0])
        ref = (its, {n: oracle_global(o, n, c.TIME_CUR) for n in FIELDS_CMP}, None)
    elif rank == 0:
        ref = run(cs, c.copy_config(cs.cfg, rank=0, nranks=1, device=0), None, steps, coupled)
    dist.barrier()
    obj = [P.api.Pop.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    its, out, rows = run(cs, c.copy_config(cs.cfg, rank=rank, nranks=world, device=int(os.environ["LOCAL_RANK"])),
                         obj[0], steps, coupled)
    parts = [None] * world if rank == 0 else None
    dist.gather_object((its, out, (rows.start, rows.stop)), parts, dst=0)
    ok = True
    if rank == 0:
        its1, out1, _ = ref
        for r, (itr, _, _) in enumerate(parts):
            if itr != its1:
                print("FAIL rank %d solver iterations %s vs single-strip %s" % (r, itr, its1))
                ok = False
        for n in FIELDS_CMP + (("SFC_OUT",) if coupled else ()):
            full = np.concatenate([p_[1][n] for p_ in parts], axis=1)
            if not np.array_equal(full, out1[n]):
                d = np.max(np.abs(full - out1[n]))
                s = np.max(np.abs(out1[n]))
                print("FAIL %s differs from the single-strip run: max abs %.3e (rel %.3e)" % (n, d, d / max(s, 1e-300)))
                ok = False
        print("MULTIRANK %s world=%d iterations=%s %s" % (which, world, its1, "OK" if ok else "FAILED"))
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
