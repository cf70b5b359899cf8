!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
!
! pop_b200_bind -- ISO_C_BINDING interface blocks for libpop_b200.so (include/pop_b200.h).
!
! This module is the whole Fortran side of the drop-in: the bodies of the reference's hot-path
! procedures are replaced by one call each into the C ABI below, the callers in step_mod.F90,
! baroclinic.F90 and barotropic.F90 stay unchanged (INTEGRATION.md shows the replaced bodies).
!
! There is no Fortran compiler in the build image of this repository, so this file is kept in sync
! with include/pop_b200.h by tests/test_abi_symbols.py (every bind(C) name below must be a symbol
! declared in the header and exported by the library) rather than by compiling it.
!
! Conventions: arrays are passed by reference exactly as the reference passes them (explicit-shape,
! column-major, contiguous because only trailing dimensions are sliced); k, nfirst, nlast, blockIndx
! stay 1-based; every function returns POP_Success (0) / POP_Fail (-1) like POP_ErrorMod.F90:45-47.
!
!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||

module pop_b200_bind

   use, intrinsic :: iso_c_binding

   implicit none
   public

   integer (c_int), parameter :: POP_B200_MAX_NT = 64

   ! field location / kind ids (POP_GridHorzMod.F90:52-57, POP_FieldMod.F90:106-108)
   integer (c_int), parameter :: POP_B200_LOC_CENTER = 1, POP_B200_LOC_NECORNER = 2, &
                                 POP_B200_LOC_NFACE  = 3, POP_B200_LOC_EFACE    = 4
   integer (c_int), parameter :: POP_B200_KIND_SCALAR = 1, POP_B200_KIND_VECTOR = 2, &
                                 POP_B200_KIND_ANGLE  = 3
   integer (c_int), parameter :: POP_B200_TS_LEAPFROG = 1, POP_B200_TS_EULER = 2, POP_B200_TS_AVG = 3, &
                                 POP_B200_TS_ROBERT = 4
   integer (c_int), parameter :: POP_B200_TIME_OLD = 0, POP_B200_TIME_CUR = 1, POP_B200_TIME_NEW = 2

   ! mirrors struct pop_config (include/pop_b200.h); filled from the namelist variables the
   ! reference modules already hold (domain_size, advect_nml, hmix_*_nml, vertical_mix_nml, grid_nml,
   ! pressure_grad_nml, state_nml, solvers, time_manager_nml)
   type, bind(C) :: pop_config
      integer (c_int) :: nx_global, ny_global, km, nt
      integer (c_int) :: ew_boundary_type, ns_boundary_type
      integer (c_int) :: block_size_x, block_size_y
      integer (c_int) :: tadvect_itype(POP_B200_MAX_NT)
      integer (c_int) :: hmix_tracer_itype, hmix_momentum_itype
      real (c_double) :: ah, am
      integer (c_int) :: lvariable_hmixt, lvariable_hmixu
      integer (c_int) :: lauto_hmixt, lauto_hmixu
      real (c_double) :: ah_gm, ah_bolus, ah_bkg_srfbl, slm_r, slm_b
      integer (c_int) :: vmix_itype, implicit_vertical_mix
      integer (c_int) :: vdc_kdim_halo
      integer (c_int) :: vdc_ndim
      real (c_double) :: aidif, bottom_drag
      real (c_double) :: const_vdc, const_vvc
      real (c_double) :: bckgrnd_vdc, bckgrnd_vvc, rich_mix
      integer (c_int) :: convection_diff
      real (c_double) :: convect_diff, convect_visc
      integer (c_int) :: sfc_layer_type, partial_bottom_cells
      integer (c_int) :: lpressure_avg, lbouss_correct, impcor
      integer (c_int) :: state_itype, state_range_iopt
      integer (c_int) :: solver_choice, max_iterations, convergence_check_freq, convergence_check_start
      integer (c_int) :: max_lanczos_step
      real (c_double) :: convergence_criterion, lanczos_convergence_criterion
      real (c_double) :: dtt
      integer (c_int) :: rank, nranks, device
      real (c_double) :: robert_alpha, robert_nu
      integer (c_int) :: nconvad
      integer (c_int) :: preconditioner_choice
   end type pop_config

   ! mirrors struct pop_block = `type block` of blocks.F90:30-39
   type, bind(C) :: pop_block
      integer (c_int) :: block_id, local_id
      integer (c_int) :: ib, ie, jb, je
      integer (c_int) :: iblock, jblock
      type (c_ptr)    :: i_glob, j_glob
   end type pop_block

   interface

      ! ---------------------------------------------------------------- lifecycle
      subroutine pop_config_defaults(cfg) bind(C, name='pop_config_defaults')
         import :: pop_config
         type (pop_config), intent(out) :: cfg
      end subroutine

      function pop_init(cfg) bind(C, name='pop_init') result(ierr)
         import :: pop_config, c_int
         type (pop_config), intent(in) :: cfg
         integer (c_int) :: ierr
      end function

      function pop_finalize() bind(C, name='pop_finalize') result(ierr)
         import :: c_int
         integer (c_int) :: ierr
      end function

      function pop_comm_unique_id(id128) bind(C, name='pop_comm_unique_id') result(ierr)
         import :: c_int, c_char
         character (kind=c_char), intent(out) :: id128(128)
         integer (c_int) :: ierr
      end function

      function pop_comm_init(rank, nranks, id128) bind(C, name='pop_comm_init') result(ierr)
         import :: c_int, c_char
         integer (c_int), value :: rank, nranks
         character (kind=c_char), intent(in) :: id128(128)
         integer (c_int) :: ierr
      end function

      function pop_get_block(blk) bind(C, name='pop_get_block') result(ierr)
         import :: pop_block, c_int
         type (pop_block), intent(out) :: blk
         integer (c_int) :: ierr
      end function

      function pop_last_error() bind(C, name='pop_last_error') result(msg)
         import :: c_ptr
         type (c_ptr) :: msg
      end function

      ! ---------------------------------------------------------------- grid + fields
      ! grid.F90:2116 read_bottom_cell -> DZBC (physical strip); before pop_set_grid when partial_bottom_cells
      function pop_set_bottom_cells(DZBC) bind(C, name='pop_set_bottom_cells') result(ierr)
         import :: c_int, c_double
         real (c_double), intent(in) :: DZBC(*)
         integer (c_int) :: ierr
      end function

      function pop_set_grid(ULAT, HTN, HTE, HUS, HUW, DXU, DYU, DXT, DYT, KMT, dz) &
               bind(C, name='pop_set_grid') result(ierr)
         import :: c_int, c_double
         real (c_double), intent(in) :: ULAT(*), HTN(*), HTE(*), HUS(*), HUW(*), DXU(*), DYU(*), &
                                        DXT(*), DYT(*), dz(*)
         integer (c_int), intent(in) :: KMT(*)
         integer (c_int) :: ierr
      end function

      function pop_set_field(name, tlev, host) bind(C, name='pop_set_field') result(ierr)
         import :: c_int, c_char, c_double
         character (kind=c_char), intent(in) :: name(*)
         integer (c_int), value :: tlev
         real (c_double), intent(in) :: host(*)
         integer (c_int) :: ierr
      end function

      function pop_get_field(name, tlev, host) bind(C, name='pop_get_field') result(ierr)
         import :: c_int, c_char, c_double
         character (kind=c_char), intent(in) :: name(*)
         integer (c_int), value :: tlev
         real (c_double), intent(out) :: host(*)
         integer (c_int) :: ierr
      end function

      function pop_set_timestep(ts_type) bind(C, name='pop_set_timestep') result(ierr)
         import :: c_int
         integer (c_int), value :: ts_type
         integer (c_int) :: ierr
      end function

      ! ---------------------------------------------------------------- slab operators
      ! advection.F90:1577   advt(k,LTK,WTK,TMIX,TRCR,UUU,VVV,this_block)
      function pop_advt(k, LTK, WTK, TMIX, TRCR, UUU, VVV, blk) bind(C, name='pop_advt') result(ierr)
         import :: c_int, c_double, pop_block
         integer (c_int), value :: k
         real (c_double), intent(out)   :: LTK(*)
         real (c_double), intent(inout) :: WTK(*)
         real (c_double), intent(in)    :: TMIX(*), TRCR(*), UUU(*), VVV(*)
         type (pop_block), intent(in)   :: blk
         integer (c_int) :: ierr
      end function

      ! advection.F90:1014   comp_flux_vel_ghost(DH, errorCode)
      function pop_comp_flux_vel_ghost(DH) bind(C, name='pop_comp_flux_vel_ghost') result(ierr)
         import :: c_int, c_double
         real (c_double), intent(in) :: DH(*)
         integer (c_int) :: ierr
      end function

      ! advection.F90:1127   advu(k,LUK,LVK,WUK,UUU,VVV,this_block)
      function pop_advu(k, LUK, LVK, WUK, UUU, VVV, blk) bind(C, name='pop_advu') result(ierr)
         import :: c_int, c_double, pop_block
         integer (c_int), value :: k
         real (c_double), intent(out)   :: LUK(*), LVK(*)
         real (c_double), intent(inout) :: WUK(*)
         real (c_double), intent(in)    :: UUU(*), VVV(*)
         type (pop_block), intent(in)   :: blk
         integer (c_int) :: ierr
      end function

      ! horizontal_mix.F90:486   hdifft(k,HDTK,TMIX,UMIX,VMIX,this_block)
      function pop_hdifft(k, HDTK, TMIX, UMIX, VMIX, blk) bind(C, name='pop_hdifft') result(ierr)
         import :: c_int, c_double, pop_block
         integer (c_int), value :: k
         real (c_double), intent(out) :: HDTK(*)
         real (c_double), intent(in)  :: TMIX(*), UMIX(*), VMIX(*)
         type (pop_block), intent(in) :: blk
         integer (c_int) :: ierr
      end function

      ! horizontal_mix.F90:427   hdiffu(k,HDUK,HDVK,UMIXK,VMIXK,this_block)
      function pop_hdiffu(k, HDUK, HDVK, UMIXK, VMIXK, blk) bind(C, name='pop_hdiffu') result(ierr)
         import :: c_int, c_double, pop_block
         integer (c_int), value :: k
         real (c_double), intent(out) :: HDUK(*), HDVK(*)
         real (c_double), intent(in)  :: UMIXK(*), VMIXK(*)
         type (pop_block), intent(in) :: blk
         integer (c_int) :: ierr
      end function

      ! pressure_grad.F90:187   gradp(k,PKX,PKY,RHOK_OLD,RHOK_CUR,RHOK_NEW,this_block)
      function pop_gradp(k, PKX, PKY, RHOK_OLD, RHOK_CUR, RHOK_NEW, blk) &
               bind(C, name='pop_gradp') result(ierr)
         import :: c_int, c_double, pop_block
         integer (c_int), value :: k
         real (c_double), intent(out) :: PKX(*), PKY(*)
         real (c_double), intent(in)  :: RHOK_OLD(*), RHOK_CUR(*), RHOK_NEW(*)
         type (pop_block), intent(in) :: blk
         integer (c_int) :: ierr
      end function

      ! operators.F90:126 / :49
      function pop_grad(k, GRADX, GRADY, F, blk) bind(C, name='pop_grad') result(ierr)
         import :: c_int, c_double, pop_block
         integer (c_int), value :: k
         real (c_double), intent(out) :: GRADX(*), GRADY(*)
         real (c_double), intent(in)  :: F(*)
         type (pop_block), intent(in) :: blk
         integer (c_int) :: ierr
      end function

      function pop_div(k, DIV_OUT, UX, UY, blk) bind(C, name='pop_div') result(ierr)
         import :: c_int, c_double, pop_block
         integer (c_int), value :: k
         real (c_double), intent(out) :: DIV_OUT(*)
         real (c_double), intent(in)  :: UX(*), UY(*)
         type (pop_block), intent(in) :: blk
         integer (c_int) :: ierr
      end function

      ! vertical_mix.F90:691   vdifft(k,VDTK,TOLD,STF,this_block)
      function pop_vdifft(k, VDTK, TOLD, STF, blk) bind(C, name='pop_vdifft') result(ierr)
         import :: c_int, c_double, pop_block
         integer (c_int), value :: k
         real (c_double), intent(out) :: VDTK(*)
         real (c_double), intent(in)  :: TOLD(*), STF(*)
         type (pop_block), intent(in) :: blk
         integer (c_int) :: ierr
      end function

      ! vertical_mix.F90:853   vdiffu(k,VDUK,VDVK,UOLD,VOLD,SMF,this_block)
      function pop_vdiffu(k, VDUK, VDVK, UOLD, VOLD, SMF, blk) bind(C, name='pop_vdiffu') result(ierr)
         import :: c_int, c_double, pop_block
         integer (c_int), value :: k
         real (c_double), intent(out) :: VDUK(*), VDVK(*)
         real (c_double), intent(in)  :: UOLD(*), VOLD(*), SMF(*)
         type (pop_block), intent(in) :: blk
         integer (c_int) :: ierr
      end function

      ! vertical_mix.F90:1164   impvmixt(TNEW,TOLD,PSFC,nfirst,nlast,this_block)
      function pop_impvmixt(TNEW, TOLD, PSFC, nfirst, nlast, blk) bind(C, name='pop_impvmixt') result(ierr)
         import :: c_int, c_double, pop_block
         real (c_double), intent(inout) :: TNEW(*)
         real (c_double), intent(in)    :: TOLD(*), PSFC(*)
         integer (c_int), value :: nfirst, nlast
         type (pop_block), intent(in)   :: blk
         integer (c_int) :: ierr
      end function

      ! vertical_mix.F90:1460   impvmixt_correct(TNEW,PSFC,RHS,nfirst,nlast,this_block)
      function pop_impvmixt_correct(TNEW, PSFC, RHS, nfirst, nlast, blk) &
               bind(C, name='pop_impvmixt_correct') result(ierr)
         import :: c_int, c_double, pop_block
         real (c_double), intent(inout) :: TNEW(*)
         real (c_double), intent(in)    :: PSFC(*), RHS(*)
         integer (c_int), value :: nfirst, nlast
         type (pop_block), intent(in)   :: blk
         integer (c_int) :: ierr
      end function

      ! vertical_mix.F90:1679   impvmixu(UNEW,VNEW,this_block)
      function pop_impvmixu(UNEW, VNEW, blk) bind(C, name='pop_impvmixu') result(ierr)
         import :: c_int, c_double, pop_block
         real (c_double), intent(inout) :: UNEW(*), VNEW(*)
         type (pop_block), intent(in)   :: blk
         integer (c_int) :: ierr
      end function

      ! vertical_mix.F90:518   vmix_coeffs(k,TMIX,UMIX,VMIX,RHOMIX,this_block)
      function pop_vmix_coeffs(k, TMIX, UMIX, VMIX, RHOMIX, blk) bind(C, name='pop_vmix_coeffs') result(ierr)
         import :: c_int, c_double, pop_block
         integer (c_int), value :: k
         real (c_double), intent(in)  :: TMIX(*), UMIX(*), VMIX(*), RHOMIX(*)
         type (pop_block), intent(in) :: blk
         integer (c_int) :: ierr
      end function

      ! state_mod.F90:258   state(k,kk,TEMPK,SALTK,this_block,RHOOUT,RHOFULL,DRHODT,DRHODS)
      ! optional outputs of the reference become nullable pointers: pass c_null_ptr when absent
      function pop_state(k, kk, TEMPK, SALTK, blk, RHOOUT, RHOFULL, DRHODT, DRHODS) &
               bind(C, name='pop_state') result(ierr)
         import :: c_int, c_double, c_ptr, pop_block
         integer (c_int), value :: k, kk
         real (c_double), intent(in)  :: TEMPK(*), SALTK(*)
         type (pop_block), intent(in) :: blk
         type (c_ptr), value :: RHOOUT, RHOFULL, DRHODT, DRHODS
         integer (c_int) :: ierr
      end function

      ! ---------------------------------------------------------------- barotropic solver
      ! POP_SolversMod.F90:327   POP_SolversRun(sfcPressure,rhsClinic,errorCode)
      function pop_solvers_run(sfcPressure, rhsClinic) bind(C, name='pop_solvers_run') result(ierr)
         import :: c_int, c_double
         real (c_double), intent(inout) :: sfcPressure(*)
         real (c_double), intent(in)    :: rhsClinic(*)
         integer (c_int) :: ierr
      end function

      ! POP_SolversMod.F90:1110   POP_SolversDiagonal(diagonalCorrection,blockIndx,errorCode)
      function pop_solvers_diagonal(diagonalCorrection, blockIndx) &
               bind(C, name='pop_solvers_diagonal') result(ierr)
         import :: c_int, c_double
         real (c_double), intent(in) :: diagonalCorrection(*)
         integer (c_int), value :: blockIndx
         integer (c_int) :: ierr
      end function

      ! POP_SolversMod.F90:1158   POP_SolversGetDiagnostics(iterationCount,residual,errorCode)
      function pop_solvers_get_diagnostics(iterationCount, residual) &
               bind(C, name='pop_solvers_get_diagnostics') result(ierr)
         import :: c_int, c_double
         integer (c_int), intent(out) :: iterationCount
         real (c_double), intent(out) :: residual
         integer (c_int) :: ierr
      end function

      function pop_solvers_prep() bind(C, name='pop_solvers_prep') result(ierr)
         import :: c_int
         integer (c_int) :: ierr
      end function

      ! EVP preconditioner set-up diagnostics (EvpPre, POP_SolversMod.F90:2434-2506; self-check :2594-2614)
      function pop_solvers_get_evp_diagnostics(subBlocks, landSubBlocks, maxInverseError) &
               bind(C, name='pop_solvers_get_evp_diagnostics') result(ierr)
         import :: c_int, c_double
         integer (c_int), intent(out) :: subBlocks, landSubBlocks
         real (c_double), intent(out) :: maxInverseError
         integer (c_int) :: ierr
      end function

      ! ---------------------------------------------------------------- communication
      ! mpi/POP_HaloMod.F90:1732,2766,4122   POP_HaloUpdate(array,halo,fieldLoc,fieldKind,errorCode,fillValue)
      function pop_halo_update_2d_r8(array, fieldLoc, fieldKind, fillValue) &
               bind(C, name='pop_halo_update_2d_r8') result(ierr)
         import :: c_int, c_double
         real (c_double), intent(inout) :: array(*)
         integer (c_int), value :: fieldLoc, fieldKind
         real (c_double), value :: fillValue
         integer (c_int) :: ierr
      end function

      function pop_halo_update_3d_r8(array, nz, fieldLoc, fieldKind, fillValue) &
               bind(C, name='pop_halo_update_3d_r8') result(ierr)
         import :: c_int, c_double
         real (c_double), intent(inout) :: array(*)
         integer (c_int), value :: nz, fieldLoc, fieldKind
         real (c_double), value :: fillValue
         integer (c_int) :: ierr
      end function

      function pop_halo_update_4d_r8(array, nz, nt, fieldLoc, fieldKind, fillValue) &
               bind(C, name='pop_halo_update_4d_r8') result(ierr)
         import :: c_int, c_double
         real (c_double), intent(inout) :: array(*)
         integer (c_int), value :: nz, nt, fieldLoc, fieldKind
         real (c_double), value :: fillValue
         integer (c_int) :: ierr
      end function

      function pop_halo_update_2d_i4(array, fieldLoc, fieldKind, fillValue) &
               bind(C, name='pop_halo_update_2d_i4') result(ierr)
         import :: c_int
         integer (c_int), intent(inout) :: array(*)
         integer (c_int), value :: fieldLoc, fieldKind, fillValue
         integer (c_int) :: ierr
      end function

      ! mpi/POP_HaloMod.F90:3670,5062   POP_HaloUpdate3DI4, POP_HaloUpdate4DI4
      function pop_halo_update_3d_i4(array, nz, fieldLoc, fieldKind, fillValue) &
               bind(C, name='pop_halo_update_3d_i4') result(ierr)
         import :: c_int
         integer (c_int), intent(inout) :: array(*)
         integer (c_int), value :: nz, fieldLoc, fieldKind, fillValue
         integer (c_int) :: ierr
      end function

      function pop_halo_update_4d_i4(array, nz, nt, fieldLoc, fieldKind, fillValue) &
               bind(C, name='pop_halo_update_4d_i4') result(ierr)
         import :: c_int
         integer (c_int), intent(inout) :: array(*)
         integer (c_int), value :: nz, nt, fieldLoc, fieldKind, fillValue
         integer (c_int) :: ierr
      end function

      ! mpi/POP_HaloMod.F90:2078,3218,4592   POP_HaloUpdate2DR4, POP_HaloUpdate3DR4, POP_HaloUpdate4DR4
      function pop_halo_update_2d_r4(array, fieldLoc, fieldKind, fillValue) &
               bind(C, name='pop_halo_update_2d_r4') result(ierr)
         import :: c_int, c_float
         real (c_float), intent(inout) :: array(*)
         integer (c_int), value :: fieldLoc, fieldKind
         real (c_float), value :: fillValue
         integer (c_int) :: ierr
      end function

      function pop_halo_update_3d_r4(array, nz, fieldLoc, fieldKind, fillValue) &
               bind(C, name='pop_halo_update_3d_r4') result(ierr)
         import :: c_int, c_float
         real (c_float), intent(inout) :: array(*)
         integer (c_int), value :: nz, fieldLoc, fieldKind
         real (c_float), value :: fillValue
         integer (c_int) :: ierr
      end function

      function pop_halo_update_4d_r4(array, nz, nt, fieldLoc, fieldKind, fillValue) &
               bind(C, name='pop_halo_update_4d_r4') result(ierr)
         import :: c_int, c_float
         real (c_float), intent(inout) :: array(*)
         integer (c_int), value :: nz, nt, fieldLoc, fieldKind
         real (c_float), value :: fillValue
         integer (c_int) :: ierr
      end function

      ! mpi/POP_ReductionsMod.F90:144   POP_GlobalSum(array,dist,fieldLoc,errorCode,mMask)
      function pop_global_sum_2d_r8(array, fieldLoc, mMask, total) &
               bind(C, name='pop_global_sum_2d_r8') result(ierr)
         import :: c_int, c_double, c_ptr
         real (c_double), intent(in) :: array(*)
         integer (c_int), value :: fieldLoc
         type (c_ptr), value :: mMask            ! c_null_ptr when the optional mask is absent
         real (c_double), intent(out) :: total
         integer (c_int) :: ierr
      end function

      ! ---------------------------------------------------------------- fused drivers
      function pop_dhdt() bind(C, name='pop_dhdt') result(ierr)
         import :: c_int
         integer (c_int) :: ierr
      end function

      function pop_baroclinic_driver() bind(C, name='pop_baroclinic_driver') result(ierr)
         import :: c_int
         integer (c_int) :: ierr
      end function

      function pop_barotropic_driver() bind(C, name='pop_barotropic_driver') result(ierr)
         import :: c_int
         integer (c_int) :: ierr
      end function

      function pop_baroclinic_correct_adjust() bind(C, name='pop_baroclinic_correct_adjust') result(ierr)
         import :: c_int
         integer (c_int) :: ierr
      end function

      function pop_step(ts_type) bind(C, name='pop_step') result(ierr)
         import :: c_int
         integer (c_int), value :: ts_type
         integer (c_int) :: ierr
      end function

      function pop_step_coupled(ts_type, STF, SMF, SHF_QSW, FW, sfc_out) &
               bind(C, name='pop_step_coupled') result(ierr)
         import :: c_int, c_double
         integer (c_int), value :: ts_type
         real (c_double), intent(in)  :: STF(*), SMF(*), SHF_QSW(*), FW(*)
         real (c_double), intent(out) :: sfc_out(*)
         integer (c_int) :: ierr
      end function

   end interface

end module pop_b200_bind
