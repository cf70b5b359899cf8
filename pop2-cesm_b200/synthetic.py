"""Seeded synthetic grids, bathymetry and prognostic fields, as GLOBAL numpy arrays
laid out (ny, nx) / (km, ny, nx) / (nt, km, ny, nx), i fastest -- i.e. the reference's
(nx_global, ny_global[, km[, nt]]) Fortran order.

Everything here is product-side input generation (used by tests, bench.py and smoke()); nothing
is computed by the oracle.  Units are cgs as in POP (cm, s, g; salinity g/g).
"""
import numpy as np

RADIUS = 6370.0e5          # cm, source/pop_constants.F90:239
GRAV = 980.6               # cm/s^2, source/pop_constants.F90:235

# dz tables (cm): input_templates/gx3v7_vert_grid (= gx1v7_vert_grid) and tx0.1v3_vert_grid
_DZ_GX = [1000.0] * 16 + [
    1019.6808, 1056.4484, 1105.9951, 1167.8070, 1242.4133, 1330.9678, 1435.1410, 1557.1259,
    1699.6796, 1866.2124, 2060.9023, 2288.8521, 2556.2471, 2870.5750, 3240.8372, 3677.7725,
    4194.0308, 4804.2236, 5524.7544, 6373.1919, 7366.9448, 8520.8926, 9843.6582, 11332.4658,
    12967.1992, 14705.3438, 16480.7090, 18209.1348, 19802.2344, 21185.9570, 22316.5098,
    23186.4941, 23819.4492, 24257.2168, 24546.7793, 24731.0137, 24844.3281, 24911.9746,
    24951.2910, 24973.5938, 24985.9609, 24992.6738, 24996.2441, 24998.1094]
_DZ_TX01V3 = _DZ_GX + [25000.0, 25000.0]


def vert_grid(name, km=None):
    if name in ("gx3v7", "gx1v7"):
        dz = np.array(_DZ_GX)
    elif name == "tx0.1v3":
        dz = np.array(_DZ_TX01V3)
    elif name == "uniform":
        dz = np.full(km, 5.0e5 / km)
    elif name == "stretched":   # smooth 10 m .. 250 m profile for small test grids
        k = np.arange(km)
        dz = 1000.0 + 24000.0 * (k / max(km - 1, 1)) ** 2
    else:
        raise ValueError(name)
    if km is not None and len(dz) != km:
        dz = dz[:km] if len(dz) > km else np.concatenate([dz, np.full(km - len(dz), dz[-1])])
    return np.ascontiguousarray(dz, dtype=np.float64)


def horiz_grid(nx, ny, tripole=False, lat0=-78.0, lat1=87.0):
    """Analytic lat-lon metrics; DXU/DYU/DXT/DYT by the averaging rules of read_horiz_grid
    (source/grid.F90:1433-1497).  Returns dict of (ny,nx) float64 arrays."""
    rad = np.pi / 180.0
    dlon = 360.0 / nx
    dlat = (lat1 - lat0) / ny
    j = np.arange(1, ny + 1, dtype=np.float64)
    ulat_deg = lat0 + j * dlat                  # U points
    tlat_deg = lat0 + (j - 0.5) * dlat          # T rows
    one = np.ones((ny, nx))
    ULAT = (ulat_deg * rad)[:, None] * one
    TLAT = (tlat_deg * rad)[:, None] * one      # only GM reads it (Coriolis parameter at T points)
    HTN = RADIUS * dlon * rad * np.cos(ulat_deg * rad)[:, None] * one
    HTE = RADIUS * dlat * rad * one
    HUS = RADIUS * dlon * rad * np.cos(tlat_deg * rad)[:, None] * one
    HUW = RADIUS * dlat * rad * one
    DXU = 0.5 * (HTN + np.roll(HTN, -1, axis=1))
    DXT = 0.5 * (HTN + np.roll(HTN, 1, axis=0))          # j=1 wraps to ny ("assume cyclic")
    DYT = 0.5 * (HTE + np.roll(HTE, 1, axis=1))
    DYU = 0.5 * (HTE + np.roll(HTE, -1, axis=0))
    if tripole:
        DYU[-1, :] = HTE[-1, :]
    return {k: np.ascontiguousarray(v) for k, v in dict(
        TLAT=TLAT, ULAT=ULAT, HTN=HTN, HTE=HTE, HUS=HUS, HUW=HUW, DXU=DXU, DYU=DYU, DXT=DXT, DYT=DYT).items()}


def _lowwave(nx, ny, rng, nmodes=6, periodic_j=False):
    i = np.arange(nx)[None, :] / nx
    j = np.arange(ny)[:, None] / ny
    s = np.zeros((ny, nx))
    for _ in range(nmodes):
        kx = rng.integers(1, 5)
        ky = rng.integers(1, 4)
        ph1, ph2 = rng.uniform(0, 2 * np.pi, 2)
        amp = rng.uniform(0.5, 1.0)
        s += amp * np.sin(2 * np.pi * kx * i + ph1) * np.cos(np.pi * ky * j + ph2)
    s -= s.min()
    s /= s.max()
    return s


def bathymetry(nx, ny, km, seed, land_frac_thresh=0.12, south_land_rows=2, flat=False):
    """KMT(ny,nx) int32: ~25% land, closed southern rows."""
    rng = np.random.default_rng(seed)
    s = _lowwave(nx, ny, rng)
    kmt = np.clip(np.rint(km * (0.25 + 0.75 * s)), 0, km).astype(np.int32)
    if flat:
        kmt[:] = km
    kmt[s < land_frac_thresh] = 0
    kmt[:south_land_rows, :] = 0
    return np.ascontiguousarray(kmt)


def bottom_cells(kmt, dz, seed, fmin=0.25):
    """DZBC(ny,nx): thickness of the bottom cell of every column for partial bottom cells (grid.F90:917-940): a seeded
    fraction fmin..1 of the full thickness of level KMT; land columns get dz(1) (never used: KMT == k is false)."""
    rng = np.random.default_rng(seed)
    frac = fmin + (1.0 - fmin) * rng.uniform(0.0, 1.0, kmt.shape)
    full = np.where(kmt > 0, dz[np.maximum(kmt, 1) - 1], dz[0])
    return np.ascontiguousarray(np.where(kmt > 0, frac * full, dz[0]))


def kmu_from_kmt(kmt, ew_cyclic=True, ns_type=0):
    """KMU = min of the 4 surrounding KMT (source/grid.F90:978-985) on the global grid."""
    ny, nx = kmt.shape
    e = np.roll(kmt, -1, axis=1)
    if not ew_cyclic:
        e[:, -1] = 0
    if ns_type == 2:      # tripole: T row ny+1 is row ny mirrored, (i) <- (nx-i+1)
        top = kmt[-1, ::-1][None, :]
    elif ns_type == 1:
        top = kmt[0:1, :]
    else:
        top = np.zeros((1, nx), kmt.dtype)
    n = np.concatenate([kmt[1:], top], axis=0)
    ne = np.roll(n, -1, axis=1)
    if not ew_cyclic:
        ne[:, -1] = 0
    return np.minimum(np.minimum(kmt, e), np.minimum(n, ne))


def state(nx, ny, km, nt, dz, kmt, kmu, seed, amp_noise=1.0):
    """T,S (+ passive), U,V at cur and old levels, PSURF at cur/old (global arrays).
    T = 2+23 exp(-z/800m) + noise, S = (34.7+noise)/1000, passive U(0,1); U,V from a random
    stream function (<~50 cm/s, e-fold 1000 m) + 1 cm/s noise, masked by KMU; land = 0."""
    rng = np.random.default_rng(seed)
    zt = (np.cumsum(dz) - 0.5 * dz) * 0.01      # m
    out = {}
    maskT = (np.arange(1, km + 1)[:, None, None] <= kmt[None]).astype(np.float64)
    maskU = (np.arange(1, km + 1)[:, None, None] <= kmu[None]).astype(np.float64)
    i = np.arange(nx)[None, :] / nx
    j = np.arange(ny)[:, None] / ny
    psi = np.zeros((ny, nx))
    for _ in range(5):
        kx, ky = rng.integers(1, 4), rng.integers(1, 4)
        p1, p2 = rng.uniform(0, 2 * np.pi, 2)
        psi += rng.uniform(0.3, 1.0) * np.sin(2 * np.pi * kx * i + p1) * np.sin(np.pi * ky * j + p2)
    u0 = -(np.roll(psi, -1, axis=0) - np.roll(psi, 1, axis=0))
    v0 = (np.roll(psi, -1, axis=1) - np.roll(psi, 1, axis=1))
    sc = 40.0 / max(np.abs(u0).max(), np.abs(v0).max(), 1e-30)
    prof = np.exp(-zt / 1000.0)[:, None, None]
    for lev, eps in (("cur", 0.0), ("old", 1.0)):
        T = np.empty((nt, km, ny, nx))
        T[0] = 2.0 + 23.0 * np.exp(-zt / 800.0)[:, None, None] \
            + amp_noise * 0.5 * rng.standard_normal((km, ny, nx)) * np.exp(-zt / 500.0)[:, None, None]
        T[1] = (34.7 + amp_noise * 0.2 * rng.standard_normal((km, ny, nx))) / 1000.0
        for n in range(2, nt):
            T[n] = rng.uniform(0.0, 1.0, (km, ny, nx))
        T *= maskT[None]
        U = (sc * u0[None] * prof + amp_noise * rng.standard_normal((km, ny, nx))) * maskU
        V = (sc * v0[None] * prof + amp_noise * rng.standard_normal((km, ny, nx))) * maskU
        eta = 50.0 * np.sin(2 * np.pi * (2 * i + 0.3 * eps)) * np.cos(np.pi * j) * (kmt > 0)
        out["TRACER_" + lev] = np.ascontiguousarray(T)
        out["UVEL_" + lev] = np.ascontiguousarray(U)
        out["VVEL_" + lev] = np.ascontiguousarray(V)
        out["PSURF_" + lev] = np.ascontiguousarray(GRAV * eta + eps * 10.0 * rng.standard_normal((ny, nx)) * (kmt > 0))
    return out


def kpp_shaped_vdc(nx, ny, km, kmt, seed, ndim=2, halo=True):
    """Synthetic 'KPP-shaped' VDC(ndim, 0:km+1 or 1:km, ny, nx) and VVC(km, ny, nx) (cm^2/s):
    large in a surface boundary layer, background 0.1-1 below (vertical_mix.F90:418-419)."""
    rng = np.random.default_rng(seed)
    nk = km + 2 if halo else km
    k = (np.arange(nk) - (1 if halo else 0))[:, None, None]
    hbl = rng.uniform(3, 10, (ny, nx))[None]
    shape = np.clip(1.0 - k / hbl, 0.0, 1.0)
    vdc = np.empty((ndim, nk, ny, nx))
    for d in range(ndim):
        vdc[d] = 0.1 + 500.0 * shape ** 2 * (1.0 + 0.1 * d) + 0.05 * rng.uniform(0, 1, (nk, ny, nx))
    kv = np.arange(1, km + 1)[:, None, None]
    vvc = 1.0 + 800.0 * np.clip(1.0 - kv / hbl, 0.0, 1.0) ** 2 + 0.1 * rng.uniform(0, 1, (km, ny, nx))
    return np.ascontiguousarray(vdc), np.ascontiguousarray(vvc)
