"""ctypes wrapper of the CPU ORACLE (oracle/libpop_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product package."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import pop_pkg  # noqa: E402

P = pop_pkg.load()
cfgmod = P.config
_LIB = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def build():
    subprocess.check_call(["make", "-s", "-C", HERE], stdout=subprocess.DEVNULL)


def build_fast():
    """TIMING copy (bench.py only): -O3 -march=native, built on the host it runs on; the file name carries a
    hash of that host's CPU model and flags so that a copy built elsewhere is never loaded."""
    import hashlib
    try:
        info = [l for l in open("/proc/cpuinfo") if l.startswith(("model name", "flags"))][:2]
    except OSError:
        info = [os.uname().machine]
    tag = hashlib.sha1("".join(info).encode()).hexdigest()[:10]
    so = os.path.join(HERE, "libpop_oracle_fast_%s.so" % tag)
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h"))]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE, "fast", "FAST_SO=" + so], stdout=subprocess.DEVNULL)
    return so


_LIBS = {}


def lib(so=None):
    """the checker build by default; bench.py passes the path of its timing copy (each shared object has its own
    model state, so both can live in one process)"""
    global _LIB
    so = so or os.path.join(HERE, "libpop_oracle.so")
    if so not in _LIBS:
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.oracle_get_max_threads.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_field.restype = C.c_void_p
        L.oracle_field.argtypes = [C.c_char_p, C.c_int]
        L.oracle_global_sum.restype = C.c_double
        L.oracle_global_sum.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_timer.restype = C.c_double
        for f in ("oracle_halo_2d",):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.oracle_halo_3d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double]
        L.oracle_halo_4d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.oracle_halo_2d_i4.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_scatter.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_gather_i4.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_set_grid.argtypes = [C.c_void_p] * 11
        _LIBS[so] = L
        if so.endswith("libpop_oracle.so"):
            _LIB = L
    return _LIBS[so]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """One oracle model instance (the C side is a singleton, like the Fortran modules)."""

    def __init__(self, cfg, so=None):
        self.L = lib(so)
        self.cfg = cfg
        assert self.L.oracle_init(C.byref(cfg)) == 0
        nb = C.c_int.in_dll(self.L, "M")  # not used; sizes come from cfg
        self.nxb = cfg.block_size_x + 4
        self.nyb = cfg.block_size_y + 4
        self.nbx = (cfg.nx_global - 1) // cfg.block_size_x + 1
        self.nby = (cfg.ny_global - 1) // cfg.block_size_y + 1
        self.nblocks = self.nbx * self.nby
        self.km, self.nt = cfg.km, cfg.nt

    # ---- geometry
    def block_info(self, b):
        out8 = (C.c_int * 8)()
        ig = (C.c_int * self.nxb)()
        jg = (C.c_int * self.nyb)()
        assert self.L.oracle_block_info(b, out8, ig, jg) == 0
        return list(out8), np.array(ig), np.array(jg)

    def set_active(self, active):
        a = np.ascontiguousarray(active, dtype=np.int32)
        self.L.oracle_set_active(_p(a))

    # ---- raw block-major views
    def view(self, name, tlev=1, inner=(), dtype=np.float64):
        ptr = self.L.oracle_field(name.encode(), tlev)
        if not ptr:
            raise KeyError(name)
        shape = (self.nblocks,) + tuple(inner) + (self.nyb, self.nxb)
        n = int(np.prod(shape))
        ct = C.c_double if dtype == np.float64 else C.c_int
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).reshape(shape)

    def vec(self, name, n):
        ptr = self.L.oracle_field(name.encode(), 1)
        return np.ctypeslib.as_array(C.cast(ptr, dp), shape=(n,))

    def inner_shape(self, name):
        km, nt = self.km, self.nt
        if name == "TRACER":
            return (nt, km)
        if name in ("UVEL", "VVEL", "RHO"):
            return (km,)
        if name in ("STF", "TFW"):
            return (nt,)
        if name == "SMF":
            return (2,)
        if name == "VDC":
            c = self.cfg
            nk = km + 2 if (c.vmix_itype == cfgmod.VMIX_GIVEN and c.vdc_kdim_halo) else km
            nd = c.vdc_ndim if c.vmix_itype == cfgmod.VMIX_GIVEN else 1
            return (nd, nk)
        if name == "VVC":
            return (km,)
        return ()

    # ---- global <-> blocks
    def scatter(self, name, tlev, glob):
        inner = self.inner_shape(name)
        nz = int(np.prod(inner)) if inner else 1
        g = np.ascontiguousarray(glob, dtype=np.float64)
        assert g.size == nz * self.cfg.nx_global * self.cfg.ny_global, (name, g.shape)
        # oracle_scatter expects dst [b][nz][..], glob [nz][ny][nx]
        self.L.oracle_scatter(self.L.oracle_field(name.encode(), tlev), _p(g), nz)

    def gather(self, name, tlev=1):
        inner = self.inner_shape(name)
        nz = int(np.prod(inner)) if inner else 1
        g = np.zeros(inner + (self.cfg.ny_global, self.cfg.nx_global))
        self.L.oracle_gather(_p(g), self.L.oracle_field(name.encode(), tlev), nz)
        return g

    def gather_i4(self, name):
        g = np.zeros((self.cfg.ny_global, self.cfg.nx_global), dtype=np.int32)
        self.L.oracle_gather_i4(_p(g), self.L.oracle_field(name.encode(), 1))
        return g

    def halo(self, name, tlev, loc, kind, fill=0.0):
        inner = self.inner_shape(name)
        ptr = self.L.oracle_field(name.encode(), tlev)
        if len(inner) == 0:
            self.L.oracle_halo_2d(ptr, loc, kind, fill)
        elif len(inner) == 1:
            self.L.oracle_halo_3d(ptr, inner[0], loc, kind, fill)
        else:
            self.L.oracle_halo_4d(ptr, inner[1], inner[0], loc, kind, fill)

    def halo_array(self, arr, loc, kind, fill=0.0):
        """arr: block-major (nblocks, [nz,] nyb, nxb) float64 or int32, updated in place."""
        a = arr
        assert a.flags.c_contiguous
        if a.dtype == np.int32:
            assert a.ndim == 3
            self.L.oracle_halo_2d_i4(_p(a), loc, kind, int(fill))
        elif a.ndim == 3:
            self.L.oracle_halo_2d(_p(a), loc, kind, float(fill))
        elif a.ndim == 4:
            self.L.oracle_halo_3d(_p(a), a.shape[1], loc, kind, float(fill))
        else:
            self.L.oracle_halo_4d(_p(a), a.shape[2], a.shape[1], loc, kind, float(fill))

    def global_sum(self, arr, loc, mask=None):
        return self.L.oracle_global_sum(_p(arr), loc, _p(mask))

    # ---- model setup
    def set_bottom_cells(self, dzbc):
        """partial bottom cells: global (ny, nx) thickness of the bottom cell of every column; before set_grid"""
        a = np.ascontiguousarray(dzbc, dtype=np.float64)
        assert self.L.oracle_set_bottom_cells(_p(a)) == 0

    def set_grid(self, grid, kmt, dz):
        args = [np.ascontiguousarray(grid[k], dtype=np.float64) for k in
                ("ULAT", "HTN", "HTE", "HUS", "HUW", "DXU", "DYU", "DXT", "DYT")]
        kmt = np.ascontiguousarray(kmt, dtype=np.int32)
        dz = np.ascontiguousarray(dz, dtype=np.float64)
        rc = self.L.oracle_set_grid(*[_p(a) for a in args], _p(kmt), _p(dz))
        assert rc == 0

    def state_all(self, tlev):
        """RHO(tlev) = state(T,S) for all blocks and levels."""
        T = self.view("TRACER", tlev, (self.nt, self.km))
        R = self.view("RHO", tlev, (self.km,))
        self.L.o_state.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 4
        for b in range(self.nblocks):
            for k in range(1, self.km + 1):
                self.L.o_state(k, k, _p(T[b, 0, k - 1]), _p(T[b, 1, k - 1]), b, _p(R[b, k - 1]),
                               None, None, None)

    def grad_psurf(self, tlev):
        Pv, GX, GY = self.view("PSURF", tlev), self.view("GRADPX", tlev), self.view("GRADPY", tlev)
        self.L.o_grad.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        for b in range(self.nblocks):
            self.L.o_grad(1, _p(GX[b]), _p(GY[b]), _p(Pv[b]), b)
        self.halo("GRADPX", tlev, cfgmod.LOC_NECORNER, cfgmod.KIND_VECTOR)
        self.halo("GRADPY", tlev, cfgmod.LOC_NECORNER, cfgmod.KIND_VECTOR)

    def solvers_prep(self):
        return self.L.o_solvers_prep()

    def set_timestep(self, ts):
        self.L.oracle_set_timestep(ts)

    def step(self, ts):
        return self.L.oracle_step(ts)

    def solver_diag(self):
        # numIterations, rmsResidual live in M: expose through tiny accessors
        self.L.oracle_num_iterations.restype = C.c_int
        self.L.oracle_rms_residual.restype = C.c_double
        return self.L.oracle_num_iterations(), self.L.oracle_rms_residual()

    def timer(self, i):
        return self.L.oracle_timer(i)

    def set_reproducible(self, on=True):
        """the reference's REPRODUCIBLE build: global sums accumulated in real(r16) and rounded once"""
        self.L.oracle_set_reproducible(1 if on else 0)

    def set_threads(self, n):
        """OpenMP team size of the block loops; returns the size actually in effect"""
        self.L.oracle_set_threads(int(n))
        return self.L.oracle_get_max_threads()
