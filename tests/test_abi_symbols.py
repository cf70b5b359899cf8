"""CPU-side checks of the drop-in boundary: the built library exports every symbol that
include/pop_b200.h declares, the Python mirror of pop_config matches the C struct, and the product
fails loudly (no CPU fallback) when no CUDA device is present.  No compute calls are made."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    h = open(os.path.join(ROOT, "include", "pop_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(pop_[a-z0-9_]+)\s*\(", h)))


@pytest.fixture(scope="module")
def built(pkg):
    return pkg.build.build()


def test_header_and_python_symbol_lists_agree(pkg):
    assert header_functions() == sorted(pkg.api.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    missing = [s for s in header_functions() if not hasattr(lib, s)]
    assert not missing, missing


def test_config_struct_layout_matches(built, pkg):
    lib = ctypes.CDLL(built)
    cfg = pkg.config.PopConfig()
    lib.pop_config_defaults(ctypes.byref(cfg))
    # values written by the C side land in the right Python fields only if the layouts agree
    assert cfg.nt == 2 and cfg.max_iterations == 1000 and cfg.convergence_check_start == 60
    assert cfg.convergence_criterion == 1.0e-13 and cfg.slm_b == 0.3 and cfg.dtt == 3600.0
    assert cfg.tadvect_itype[63] == pkg.config.TADVECT_CENTERED and cfg.nranks == 1


def test_no_cpu_fallback(built, pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = pkg.config.make_config(nx_global=16, ny_global=16, km=4)
    with pytest.raises(pkg.api.PopError) as e:
        pkg.api.Pop(cfg)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_fortran_binding_names_are_header_symbols():
    """pop2-cesm_b200/fortran/pop_b200_bind.F90 cannot be compiled here (no Fortran compiler): keep
    its bind(C) names and the field order of its pop_config mirror in sync with the header."""
    f90 = open(os.path.join(ROOT, "pop2-cesm_b200", "fortran", "pop_b200_bind.F90")).read()
    names = re.findall(r"bind\(C,\s*name='(pop_[a-z0-9_]+)'\)", f90)
    assert len(names) >= 35
    hdr = set(header_functions())
    assert not [n for n in names if n not in hdr]
    # every reference-signature slab operator and collective is bound
    for must in ("pop_advt", "pop_advu", "pop_hdifft", "pop_hdiffu", "pop_gradp", "pop_vdifft", "pop_vdiffu",
                 "pop_impvmixt", "pop_impvmixt_correct", "pop_impvmixu", "pop_state", "pop_solvers_run",
                 "pop_solvers_diagonal", "pop_solvers_get_diagnostics", "pop_halo_update_2d_r8",
                 "pop_halo_update_3d_r8", "pop_halo_update_4d_r8", "pop_global_sum_2d_r8", "pop_step"):
        assert must in names, must
    # struct field order: C header vs Fortran derived type
    h = open(os.path.join(ROOT, "include", "pop_b200.h")).read()
    body = re.search(r"typedef struct pop_config \{(.*?)\} pop_config;", h, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    cf = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            cf += [re.sub(r"\[.*?\]", "", x).strip() for x in decl.split(None, 1)[1].split(",")]
    ft = re.search(r"type, bind\(C\) :: pop_config(.*?)end type pop_config", f90, flags=re.S).group(1)
    ff = []
    for line in ft.splitlines():
        if "::" in line:
            ff += [re.sub(r"\(.*?\)", "", x).strip() for x in line.split("::")[1].split(",")]
    assert cf == ff
