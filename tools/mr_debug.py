"""Debug aid: compare the padded arrays (ghost cells included) of a 2-strip run with the single-strip
run after every phase of the first step.  torchrun --nproc-per-node 2 tools/mr_debug.py <case>"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity import *  # noqa
import torch, torch.distributed as dist

GRID = ["KMT", "KMU", "DXU", "DYU", "DXT", "DYT", "TAREA", "TAREA_R", "UAREA_R", "HU", "HUR", "FCOR", "RCALCT", "RCALCU",
        "AU0", "AUNE", "KXU", "KYU", "DTN", "DTS", "DTE", "DTW", "AHF", "DUC", "DUN", "DUS", "DUE", "DUW", "DMC", "DMN",
        "DME", "DUM", "AMF", "btropWgtNorth", "btropWgtEast", "btropWgtNE", "centerWgtClinicIndep", "mMaskTropic",
        "CHECKER", "CONSTNT"]
PROG = ["TRACER", "UVEL", "VVEL", "RHO", "PSURF", "UBTROP", "VBTROP", "GRADPX", "GRADPY"]
WORK = ["DH", "DHU", "ZX", "ZY", "PGUESS", "RHS_BT"]


def snap(p, names, tlevs=(c.TIME_CUR,)):
    out = {}
    for n in names:
        isint = n in ("KMT", "KMU", "CHECKER", "CONSTNT")
        for t in tlevs:
            try:
                out[(n, t)] = p.get_padded(n, t, np.int32 if isint else np.float64)
            except Exception as e:
                pass
    return out


def phases(p, ts):
    res = []
    res.append(("load", snap(p, GRID) | snap(p, PROG, (c.TIME_OLD, c.TIME_CUR))))
    p.set_timestep(ts); p.dhdt(); res.append(("dhdt", snap(p, WORK)))
    p.baroclinic_driver(); res.append(("baroclinic", snap(p, PROG + WORK, (c.TIME_NEW,))))
    p.halo_field("ZX", 0, c.LOC_NECORNER, c.KIND_VECTOR); p.halo_field("ZY", 0, c.LOC_NECORNER, c.KIND_VECTOR)
    p.barotropic_driver(); res.append(("barotropic", snap(p, PROG + WORK, (c.TIME_NEW,))))
    p.baroclinic_correct_adjust(); res.append(("correct", snap(p, PROG, (c.TIME_NEW,))))
    return res


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    which = sys.argv[1]
    kw = dict(nx=96, ny=64, km=6, seed=81)
    if which == "pcsi":
        kw.update(ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_DEL4, hmix_momentum_itype=c.HMIX_DEL4,
                  lvariable_hmixt=1, lvariable_hmixu=1, ah=-3.0e21, am=-27.0e21, given_vmix=True,
                  solver_choice=c.SOLVER_PCSI, dtt=600.0)
    elif which == "chrongear":
        kw.update(convergence_criterion=1e-12)
    elif which == "cyclic_pcg":
        kw.update(ns=c.BNDY_CYCLIC, solver_choice=c.SOLVER_PCG, tadvect=c.TADVECT_UPWIND3, nt=3)
    cs = make_case(kw.pop("nx"), kw.pop("ny"), kw.pop("km"), **kw)
    ref = None
    if rank == 0:
        p = load_pop(cs, c.copy_config(cs.cfg, rank=0, nranks=1, device=0), None)
        ref = phases(p, c.TS_EULER); p.finalize()
    dist.barrier()
    obj = [P.api.Pop.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    p = load_pop(cs, c.copy_config(cs.cfg, rank=rank, nranks=world, device=int(os.environ["LOCAL_RANK"])), obj[0])
    mine = phases(p, c.TS_EULER)
    j0, nyl = p.j0, p.ny_local
    p.finalize()
    parts = [None] * world if rank == 0 else None
    dist.gather_object((mine, j0, nyl), parts, dst=0)
    if rank == 0:
        for ph, (name, rsnap) in enumerate(ref):
            for key, A in rsnap.items():
                for r, (m, j0r, nylr) in enumerate(parts):
                    Bm = m[ph][1].get(key)
                    if Bm is None: continue
                    # strip rows (padded) j = 0..nylr+3 correspond to global padded rows j0r-1 .. 
                    Aref = A[:, j0r - 1: j0r - 1 + nylr + 4, :]
                    if not np.array_equal(Aref, Bm):
                        d = np.argwhere(Aref != Bm)
                        rows = sorted(set(d[:, 1].tolist())); cols = sorted(set(d[:, 2].tolist()))
                        print("%-10s %-22s t=%d rank %d: %6d cells differ, local rows %s cols %s.. maxabs %.3e" % (
                            name, key[0], key[1], r, len(d), rows[:8], cols[:6], np.max(np.abs(Aref.astype(float) - Bm))))
        print("done")
    dist.barrier(); dist.destroy_process_group()

main()
