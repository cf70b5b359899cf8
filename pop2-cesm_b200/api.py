"""ctypes binding of the C ABI (include/pop_b200.h) -- the host-side mirror of the reference's
subroutine interface for this path.  Method names and argument order follow the Fortran routines
(advt, hdifft, vdifft, advu, hdiffu, gradp, vdiffu, impvmixt, impvmixt_correct, impvmixu, state,
POP_SolversRun, POP_HaloUpdate, POP_GlobalSum, step); errors follow the reference's errorCode
convention and are raised as PopError with the library's message.

There is no CPU implementation behind this module: loading fails loudly when
csrc/libpop_b200.so has not been built, and every compute call fails when no CUDA device exists.
"""
import ctypes as C
import os

import numpy as np

from . import config as cfgmod

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libpop_b200.so")
_LIB = None

# every symbol include/pop_b200.h declares (checked by tests/test_abi_symbols.py)
SYMBOLS = [
    "pop_last_error", "pop_config_defaults", "pop_init", "pop_finalize", "pop_is_initialized",
    "pop_comm_unique_id", "pop_comm_init", "pop_get_block", "pop_local_shape", "pop_set_bottom_cells", "pop_set_grid",
    "pop_device_ptr", "pop_field_size", "pop_set_field", "pop_get_field", "pop_scatter_field",
    "pop_gather_field", "pop_scatter_field_levels", "pop_set_timestep", "pop_advt", "pop_comp_flux_vel_ghost", "pop_advu", "pop_hdifft", "pop_hdiffu",
    "pop_gradp", "pop_grad", "pop_div", "pop_vdifft", "pop_vdiffu", "pop_impvmixt",
    "pop_impvmixt_correct", "pop_impvmixu", "pop_vmix_coeffs", "pop_state", "pop_solvers_run",
    "pop_solvers_diagonal", "pop_solvers_get_diagnostics", "pop_btrop_operator", "pop_solvers_prep",
    "pop_solvers_get_eigs", "pop_solvers_get_evp_diagnostics", "pop_halo_update_2d_r8", "pop_halo_update_3d_r8",
    "pop_halo_update_4d_r8", "pop_halo_update_2d_i4", "pop_halo_update_3d_i4", "pop_halo_update_4d_i4",
    "pop_halo_update_2d_r4", "pop_halo_update_3d_r4", "pop_halo_update_4d_r4", "pop_global_sum_2d_r8",
    "pop_global_sum_nfields_2d_r8", "pop_dhdt", "pop_baroclinic_driver", "pop_barotropic_driver",
    "pop_baroclinic_correct_adjust", "pop_step", "pop_step_coupled", "pop_kernel_launch_count",
    "pop_timer_get", "pop_timers_reset", "pop_timers_enable", "pop_sync", "pop_stream",
]


class PopError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise PopError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(needs nvcc). The pop_b200 hot path has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.pop_last_error.restype = C.c_char_p
        L.pop_device_ptr.restype = C.c_void_p
        L.pop_device_ptr.argtypes = [C.c_char_p, C.c_int]
        L.pop_field_size.restype = C.c_long
        L.pop_field_size.argtypes = [C.c_char_p]
        L.pop_kernel_launch_count.restype = C.c_long
        L.pop_stream.restype = C.c_void_p
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        L.pop_init.argtypes = [C.POINTER(cfgmod.PopConfig)]
        L.pop_set_grid.argtypes = [vp] * 11
        for f in ("pop_set_field", "pop_get_field", "pop_scatter_field", "pop_gather_field"):
            getattr(L, f).argtypes = [C.c_char_p, ci, vp]
        L.pop_scatter_field_levels.argtypes = [C.c_char_p, ci, ci, ci, vp]
        L.pop_get_block.argtypes = [C.POINTER(cfgmod.PopBlock)]
        L.pop_local_shape.argtypes = [C.POINTER(ci)] * 4
        L.pop_comm_unique_id.argtypes = [C.c_char_p]
        L.pop_comm_init.argtypes = [ci, ci, C.c_char_p]
        bp = C.POINTER(cfgmod.PopBlock)
        L.pop_advt.argtypes = [ci] + [vp] * 6 + [bp]
        L.pop_comp_flux_vel_ghost.argtypes = [vp]
        L.pop_advu.argtypes = [ci] + [vp] * 5 + [bp]
        L.pop_hdifft.argtypes = [ci] + [vp] * 4 + [bp]
        L.pop_hdiffu.argtypes = [ci] + [vp] * 4 + [bp]
        L.pop_gradp.argtypes = [ci] + [vp] * 5 + [bp]
        L.pop_grad.argtypes = [ci] + [vp] * 3 + [bp]
        L.pop_div.argtypes = [ci] + [vp] * 3 + [bp]
        L.pop_vdifft.argtypes = [ci] + [vp] * 3 + [bp]
        L.pop_vdiffu.argtypes = [ci] + [vp] * 5 + [bp]
        L.pop_impvmixt.argtypes = [vp, vp, vp, ci, ci, bp]
        L.pop_impvmixt_correct.argtypes = [vp, vp, vp, ci, ci, bp]
        L.pop_impvmixu.argtypes = [vp, vp, bp]
        L.pop_vmix_coeffs.argtypes = [ci] + [vp] * 4 + [bp]
        L.pop_state.argtypes = [ci, ci, vp, vp, bp, vp, vp, vp, vp]
        L.pop_solvers_run.argtypes = [vp, vp]
        L.pop_solvers_diagonal.argtypes = [vp, ci]
        L.pop_solvers_get_diagnostics.argtypes = [C.POINTER(ci), C.POINTER(cd)]
        L.pop_btrop_operator.argtypes = [vp, vp, ci]
        L.pop_solvers_get_eigs.argtypes = [C.POINTER(cd), C.POINTER(cd)]
        L.pop_solvers_get_evp_diagnostics.argtypes = [C.POINTER(ci), C.POINTER(ci), C.POINTER(cd)]
        L.pop_halo_update_2d_r8.argtypes = [vp, ci, ci, cd]
        L.pop_halo_update_3d_r8.argtypes = [vp, ci, ci, ci, cd]
        L.pop_halo_update_4d_r8.argtypes = [vp, ci, ci, ci, ci, cd]
        L.pop_halo_update_2d_i4.argtypes = [vp, ci, ci, ci]
        L.pop_halo_update_3d_i4.argtypes = [vp, ci, ci, ci, ci]
        L.pop_halo_update_4d_i4.argtypes = [vp, ci, ci, ci, ci, ci]
        L.pop_halo_update_2d_r4.argtypes = [vp, ci, ci, C.c_float]
        L.pop_halo_update_3d_r4.argtypes = [vp, ci, ci, ci, C.c_float]
        L.pop_halo_update_4d_r4.argtypes = [vp, ci, ci, ci, ci, C.c_float]
        L.pop_global_sum_2d_r8.argtypes = [vp, ci, vp, C.POINTER(cd)]
        L.pop_global_sum_nfields_2d_r8.argtypes = [vp, ci, ci, vp, vp]
        L.pop_step.argtypes = [ci]
        L.pop_set_timestep.argtypes = [ci]
        L.pop_step_coupled.argtypes = [ci, vp, vp, vp, vp, vp]
        L.pop_timer_get.argtypes = [C.c_char_p, C.POINTER(cd), C.POINTER(C.c_long)]
        L.pop_timers_enable.argtypes = [ci]
        _LIB = L
    return _LIB


def _p(a):
    """numpy array -> host pointer; int -> device pointer; None -> NULL."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    assert a.flags.c_contiguous, "arrays handed to the C ABI must be contiguous"
    return a.ctypes.data_as(C.c_void_p)


class Pop:
    """One library instance per process / GPU (the C side is a singleton, like the Fortran modules)."""

    def __init__(self, cfg, comm_id=None):
        self.L = lib()
        self.cfg = cfg
        if cfg.nranks > 1:
            assert comm_id is not None, "multi-rank init needs the broadcast NCCL id"
            self._ck(self.L.pop_comm_init(cfg.rank, cfg.nranks, comm_id))
        self._ck(self.L.pop_init(C.byref(cfg)))
        nxb, nyb, j0, nyl = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.pop_local_shape(C.byref(nxb), C.byref(nyb), C.byref(j0), C.byref(nyl)))
        self.nxb, self.nyb, self.j0, self.ny_local = nxb.value, nyb.value, j0.value, nyl.value
        self.km, self.nt, self.nx = cfg.km, cfg.nt, cfg.nx_global
        self.blk = cfgmod.PopBlock()
        self._ck(self.L.pop_get_block(C.byref(self.blk)))
        self.this_block = C.byref(self.blk)

    # ---- plumbing
    def _ck(self, rc):
        if rc != 0:
            raise PopError(self.L.pop_last_error().decode())

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        if lib().pop_comm_unique_id(buf) != 0:
            raise PopError(lib().pop_last_error().decode())
        return buf.raw

    def finalize(self):
        self.L.pop_finalize()

    def sync(self):
        self._ck(self.L.pop_sync())

    def launches(self):
        return self.L.pop_kernel_launch_count()

    def timers(self, on):
        self.L.pop_timers_enable(1 if on else 0)

    def timer(self, name):
        ms, calls = C.c_double(), C.c_long()
        self.L.pop_timer_get(name.encode(), C.byref(ms), C.byref(calls))
        return ms.value, calls.value

    # ---- geometry / fields
    def rows(self):
        """global row slice [j0-1, j0-1+ny_local) owned by this rank."""
        return slice(self.j0 - 1, self.j0 - 1 + self.ny_local)

    def set_grid(self, grid, kmt, dz):
        """grid: dict of GLOBAL (ny,nx) arrays; each rank hands over its own rows."""
        r = self.rows()
        arrs = [np.ascontiguousarray(grid[k][r], dtype=np.float64) for k in
                ("ULAT", "HTN", "HTE", "HUS", "HUW", "DXU", "DYU", "DXT", "DYT")]
        kmt_s = np.ascontiguousarray(kmt[r], dtype=np.int32)
        dz = np.ascontiguousarray(dz, dtype=np.float64)
        self._ck(self.L.pop_set_grid(*[_p(a) for a in arrs], _p(kmt_s), _p(dz)))

    def set_bottom_cells(self, dzbc):
        """partial bottom cells: GLOBAL (ny, nx) thickness of the bottom cell; this rank's rows are passed on"""
        a = np.ascontiguousarray(np.asarray(dzbc, dtype=np.float64)[self.rows(), :])
        self.L.pop_set_bottom_cells.argtypes = [C.c_void_p]
        self._ck(self.L.pop_set_bottom_cells(_p(a)))

    def field_levels(self, name):
        return self.L.pop_field_size(name.encode()) // (self.nxb * self.nyb)

    def dptr(self, name, tlev=cfgmod.TIME_CUR):
        p = self.L.pop_device_ptr(name.encode(), tlev)
        if not p:
            raise PopError("unknown field " + name)
        return p

    def scatter(self, name, tlev, glob):
        """glob: GLOBAL array (..., ny, nx); this rank's rows go to the physical cells."""
        nz = self.field_levels(name)
        g = np.asarray(glob).reshape(nz, -1, self.nx)[:, self.rows(), :]
        isint = g.dtype.kind == "i"
        g = np.ascontiguousarray(g, dtype=np.int32 if isint else np.float64)
        self._ck(self.L.pop_scatter_field(name.encode(), tlev, _p(g)))

    def scatter_levels(self, name, tlev, z0, nz, strip):
        """strip: (nz, ny_local, nx) numpy array or an int device pointer to such a strip."""
        self._ck(self.L.pop_scatter_field_levels(name.encode(), tlev, z0, nz, _p(strip)))

    def gather(self, name, tlev=cfgmod.TIME_CUR, dtype=np.float64):
        """this rank's physical strip as (nz, ny_local, nx)."""
        nz = self.field_levels(name)
        out = np.zeros((nz, self.ny_local, self.nx), dtype=dtype)
        self._ck(self.L.pop_gather_field(name.encode(), tlev, _p(out)))
        return out

    def get_padded(self, name, tlev=cfgmod.TIME_CUR, dtype=np.float64):
        nz = self.field_levels(name)
        out = np.zeros((nz, self.nyb, self.nxb), dtype=dtype)
        self._ck(self.L.pop_get_field(name.encode(), tlev, _p(out)))
        return out

    def set_padded(self, name, tlev, arr):
        a = np.ascontiguousarray(arr)
        self._ck(self.L.pop_set_field(name.encode(), tlev, _p(a)))

    def halo_field(self, name, tlev, loc, kind):
        nz = self.field_levels(name)
        self._ck(self.L.pop_halo_update_3d_r8(self.dptr(name, tlev), nz, loc, kind, 0.0))

    # ---- reference-signature operators (arrays: numpy host arrays or int device pointers)
    def state(self, k, kk, TEMPK, SALTK, RHOOUT=None, RHOFULL=None, DRHODT=None, DRHODS=None):
        self._ck(self.L.pop_state(k, kk, _p(TEMPK), _p(SALTK), self.this_block, _p(RHOOUT), _p(RHOFULL),
                                  _p(DRHODT), _p(DRHODS)))

    def advt(self, k, LTK, WTK, TMIX, TRCR, UUU, VVV):
        self._ck(self.L.pop_advt(k, _p(LTK), _p(WTK), _p(TMIX), _p(TRCR), _p(UUU), _p(VVV), self.this_block))

    def comp_flux_vel_ghost(self, DH):
        """advection.F90:1014: once per step before the first advt level when a tracer uses lw_lim."""
        self._ck(self.L.pop_comp_flux_vel_ghost(_p(DH)))

    def advu(self, k, LUK, LVK, WUK, UUU, VVV):
        self._ck(self.L.pop_advu(k, _p(LUK), _p(LVK), _p(WUK), _p(UUU), _p(VVV), self.this_block))

    def hdifft(self, k, HDTK, TMIX, UMIX=None, VMIX=None):
        self._ck(self.L.pop_hdifft(k, _p(HDTK), _p(TMIX), _p(UMIX), _p(VMIX), self.this_block))

    def hdiffu(self, k, HDUK, HDVK, UMIXK, VMIXK):
        self._ck(self.L.pop_hdiffu(k, _p(HDUK), _p(HDVK), _p(UMIXK), _p(VMIXK), self.this_block))

    def gradp(self, k, PKX, PKY, RHOK_OLD, RHOK_CUR, RHOK_NEW):
        self._ck(self.L.pop_gradp(k, _p(PKX), _p(PKY), _p(RHOK_OLD), _p(RHOK_CUR), _p(RHOK_NEW), self.this_block))

    def grad(self, k, GX, GY, F):
        self._ck(self.L.pop_grad(k, _p(GX), _p(GY), _p(F), self.this_block))

    def div(self, k, D, UX, UY):
        self._ck(self.L.pop_div(k, _p(D), _p(UX), _p(UY), self.this_block))

    def vdifft(self, k, VDTK, TOLD, STF):
        self._ck(self.L.pop_vdifft(k, _p(VDTK), _p(TOLD), _p(STF), self.this_block))

    def vdiffu(self, k, VDUK, VDVK, UOLD, VOLD, SMF):
        self._ck(self.L.pop_vdiffu(k, _p(VDUK), _p(VDVK), _p(UOLD), _p(VOLD), _p(SMF), self.this_block))

    def impvmixt(self, TNEW, TOLD, PSFC, nfirst, nlast):
        self._ck(self.L.pop_impvmixt(_p(TNEW), _p(TOLD), _p(PSFC), nfirst, nlast, self.this_block))

    def impvmixt_correct(self, TNEW, PSFC, RHS, nfirst, nlast):
        self._ck(self.L.pop_impvmixt_correct(_p(TNEW), _p(PSFC), _p(RHS), nfirst, nlast, self.this_block))

    def impvmixu(self, UNEW, VNEW):
        self._ck(self.L.pop_impvmixu(_p(UNEW), _p(VNEW), self.this_block))

    def vmix_coeffs(self, k, TMIX, UMIX, VMIX, RHOMIX):
        self._ck(self.L.pop_vmix_coeffs(k, _p(TMIX), _p(UMIX), _p(VMIX), _p(RHOMIX), self.this_block))

    def solvers_run(self, sfcPressure, rhsClinic):
        self._ck(self.L.pop_solvers_run(_p(sfcPressure), _p(rhsClinic)))

    def solvers_diagonal(self, diagonalCorrection, blockIndx=1):
        self._ck(self.L.pop_solvers_diagonal(_p(diagonalCorrection), blockIndx))

    def solvers_prep(self):
        self._ck(self.L.pop_solvers_prep())

    def solvers_get_diagnostics(self):
        it, res = C.c_int(), C.c_double()
        self._ck(self.L.pop_solvers_get_diagnostics(C.byref(it), C.byref(res)))
        return it.value, res.value

    def solvers_get_eigs(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.pop_solvers_get_eigs(C.byref(a), C.byref(b)))
        return a.value, b.value

    def solvers_get_evp_diagnostics(self):
        """(sub-blocks, sub-blocks on the diagonal fallback, max |rinv*rin - I|) of the EVP set-up"""
        n, nl, e = C.c_int(), C.c_int(), C.c_double()
        self._ck(self.L.pop_solvers_get_evp_diagnostics(C.byref(n), C.byref(nl), C.byref(e)))
        return n.value, nl.value, e.value

    def btrop_operator(self, AX, X):
        self._ck(self.L.pop_btrop_operator(_p(AX), _p(X), 1))

    def halo_update(self, array, fieldLoc, fieldKind, fillValue=0.0):
        """POP_HaloUpdate on a padded local array (nyb,nxb) / (nz,nyb,nxb) / (nt,nz,nyb,nxb)."""
        a = array
        if a.dtype in (np.int32, np.float32):      # POP_HaloUpdate{2D,3D,4D}{I4,R4}
            sfx, fv = ("i4", int(fillValue)) if a.dtype == np.int32 else ("r4", float(fillValue))
            if a.ndim == 2:
                self._ck(getattr(self.L, "pop_halo_update_2d_" + sfx)(_p(a), fieldLoc, fieldKind, fv))
            elif a.ndim == 3:
                self._ck(getattr(self.L, "pop_halo_update_3d_" + sfx)(_p(a), a.shape[0], fieldLoc, fieldKind, fv))
            else:
                self._ck(getattr(self.L, "pop_halo_update_4d_" + sfx)(_p(a), a.shape[1], a.shape[0], fieldLoc, fieldKind, fv))
        elif a.ndim == 2:
            self._ck(self.L.pop_halo_update_2d_r8(_p(a), fieldLoc, fieldKind, float(fillValue)))
        elif a.ndim == 3:
            self._ck(self.L.pop_halo_update_3d_r8(_p(a), a.shape[0], fieldLoc, fieldKind, float(fillValue)))
        else:
            self._ck(self.L.pop_halo_update_4d_r8(_p(a), a.shape[1], a.shape[0], fieldLoc, fieldKind, float(fillValue)))

    def global_sum(self, array, fieldLoc, mMask=None):
        if isinstance(array, (int, np.integer)):   # device pointer of one padded 2-d field
            s = C.c_double()
            self._ck(self.L.pop_global_sum_2d_r8(_p(array), fieldLoc, _p(mMask), C.byref(s)))
            return s.value
        a = np.ascontiguousarray(array, dtype=np.float64)
        if a.ndim == 2:
            s = C.c_double()
            self._ck(self.L.pop_global_sum_2d_r8(_p(a), fieldLoc, _p(mMask), C.byref(s)))
            return s.value
        out = np.zeros(a.shape[0])
        self._ck(self.L.pop_global_sum_nfields_2d_r8(_p(a), a.shape[0], fieldLoc, _p(mMask), _p(out)))
        return out

    # ---- drivers
    def set_timestep(self, ts):
        self._ck(self.L.pop_set_timestep(ts))

    def dhdt(self):
        self._ck(self.L.pop_dhdt())

    def baroclinic_driver(self):
        self._ck(self.L.pop_baroclinic_driver())

    def barotropic_driver(self):
        self._ck(self.L.pop_barotropic_driver())

    def baroclinic_correct_adjust(self):
        self._ck(self.L.pop_baroclinic_correct_adjust())

    def step(self, ts):
        self._ck(self.L.pop_step(ts))

    def step_coupled(self, ts, STF, SMF, SHF_QSW, FW, sfc_out):
        self._ck(self.L.pop_step_coupled(ts, _p(STF), _p(SMF), _p(SHF_QSW), _p(FW), _p(sfc_out)))
