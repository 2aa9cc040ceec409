! fortran/dml_cuda.F90 — ISO_C_BINDING shim between the reference's Fortran host code and libdml.so.
!
! NOTE: this image has no Fortran compiler (gfortran/flang/nvfortran are absent), so this file has only been
! reviewed, never compiled here.  The tested stand-ins that issue the identical call sequence through the same C ABI
! are the ctypes binding (din_mol_li_b200/dml.py, used by tests/) and tools/dana_host.cpp.
!
! How a maintainer wires it in (see INTEGRATION.md):
!   * dana.F90 keeps its loop and its names; the bodies of fuerza / ermak_a / ermak_b / cbrownian_hs /
!     overlap_moveback / gcmc_run / calc_rho / maxz / bloques and gems_neighbor::test_update call the wrappers below.
!   * atom / group / igroup objects stay on the host; fortran/dml_sync.F90 (pack_atoms / unpack_atoms /
!     apply_membership_changes / shift_chunk_template) moves them to and from the flat, hs-slot indexed arrays of include/dml.h
!     whenever membership or salida() needs them; fortran/Neighbor_gpu.F90 is the gems_neighbor module over this shim.
module dml_cuda
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: dml_config, dml_scalars, dml_counters, dmlh_rng, dml_ctx_t, dmlf_create, dmlf_check, gpu

  integer(c_int), parameter, public :: DML_F_REF = 1, DML_F_GCMC = 2, DML_F_SKIP = 4, DML_F_LIMBO = 8

  ! mirrors struct dml_config of include/dml.h field by field
  type, bind(C) :: dml_config
    integer(c_int32_t) :: device, capacity
    real(c_double)     :: box(3)
    integer(c_int32_t) :: pbc(3)
    real(c_double)     :: rcut, nb_dcut
    real(c_double)     :: eps(9), r0(9)
    real(c_double)     :: mass(3)
    real(c_double)     :: h, gama, Tsist, kB_ui, kB_ui_gcmc
    real(c_double)     :: dif_sc, dif_sei, z_sei
    real(c_double)     :: prob, z0, z1, zmax, tau
    real(c_double)     :: act
    integer(c_int32_t) :: nadj, integrador, reservoir, rng_mode
    integer(c_int64_t) :: seed
    integer(c_int32_t) :: strict_order
  end type

  ! mirrors struct dml_scalars / dml_counters of include/dml.h and dmlh_rng of include/dml_host.h
  type, bind(C) :: dml_scalars
    real(c_double)     :: box(3), z0, z1, zmax, rho, rho0, t
    integer(c_int64_t) :: step
  end type
  type, bind(C) :: dml_counters
    integer(c_int64_t) :: nupd_vlist, try_, depo, choques, choques2, choques3, list_entries
    integer(c_int64_t) :: overlap_passes, gcmc_created, gcmc_destroyed, row_overflow
    real(c_double)     :: max_vel, msd_t, msd_max
    integer(c_int32_t) :: n_slots, nat_sys, nat_ref, nat_gcmc, ncells(3)
    real(c_double)     :: cell(3)
    integer(c_int32_t) :: tessellated, listed, rows_asym
  end type
  type, bind(C) :: dmlh_rng
    integer(c_int32_t) :: idum, ix, iy, stored
    real(c_double)     :: g
    integer(c_int64_t) :: calls
  end type

  type :: dml_ctx_t
    type(c_ptr) :: h = c_null_ptr
  end type

  ! the one device context of a dana run (created by dmlf_create in the main program, used by gems_neighbor and dml_sync)
  type(dml_ctx_t), save :: gpu

  interface
    integer(c_int) function dml_create(ctx, cfg) bind(C, name='dml_create')
      import; type(c_ptr), intent(out) :: ctx; type(dml_config), intent(in) :: cfg
    end function
    subroutine dml_destroy(ctx) bind(C, name='dml_destroy')
      import; type(c_ptr), value :: ctx
    end subroutine
    type(c_ptr) function dml_last_error(ctx) bind(C, name='dml_last_error')
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function dml_upload(ctx, n, pos, vel, acel, pos_old, old_cg, z, flags, uid, slot_b) bind(C, name='dml_upload')
      import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: n
      real(c_double), intent(in) :: pos(3,*), vel(3,*), acel(3,*), pos_old(3,*), old_cg(3,*)
      integer(c_int32_t), intent(in) :: z(*), flags(*), uid(*), slot_b(*)
    end function
    integer(c_int) function dml_download(ctx, n, pos, vel, acel, force, epot, pos_old, old_cg, z, flags, uid, slot_b) &
        bind(C, name='dml_download')
      import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: n
      real(c_double), intent(out) :: pos(3,*), vel(3,*), acel(3,*), force(3,*), epot(*), pos_old(3,*), old_cg(3,*)
      integer(c_int32_t), intent(out) :: z(*), flags(*), uid(*), slot_b(*)
    end function
    ! one entry per preserved call site of dana's loop (src/dana.F90:173-265)
    integer(c_int) function dml_test_update(ctx) bind(C, name='dml_test_update')
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function dml_fuerza(ctx) bind(C, name='dml_fuerza')
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function dml_ermak_a(ctx) bind(C, name='dml_ermak_a')
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function dml_ermak_b(ctx) bind(C, name='dml_ermak_b')
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function dml_cbrownian_hs(ctx) bind(C, name='dml_cbrownian_hs')
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function dml_overlap_moveback(ctx) bind(C, name='dml_overlap_moveback')
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function dml_msd_book(ctx) bind(C, name='dml_msd_book')
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function dml_promote(ctx) bind(C, name='dml_promote')
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function dml_gcmc_run(ctx) bind(C, name='dml_gcmc_run')
      import; type(c_ptr), value :: ctx
    end function
    integer(c_int) function dml_calc_rho(ctx, rho) bind(C, name='dml_calc_rho')
      import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: rho
    end function
    integer(c_int) function dml_maxz(ctx, zmax) bind(C, name='dml_maxz')
      import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: zmax
    end function
    integer(c_int) function dml_bloques(ctx, nchunk, cpos, cpos_old, dist, rhomedia, fired) bind(C, name='dml_bloques')
      import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: nchunk
      real(c_double), intent(in) :: cpos(3,*), cpos_old(3,*); real(c_double), value :: dist, rhomedia
      integer(c_int32_t), intent(out) :: fired
    end function
    integer(c_int) function dml_step(ctx, nsteps) bind(C, name='dml_step')
      import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: nsteps
    end function
    integer(c_int) function dml_reset_try_depo(ctx) bind(C, name='dml_reset_try_depo')
      import; type(c_ptr), value :: ctx
    end function
    ! salida()/kion() sums on the device (dana.F90:1155-1163, 1342-1376): no frame download for E.dat / T.dat
    integer(c_int) function dml_salida_sums(ctx, energia, energia_ref, temp, n_mobile) bind(C, name='dml_salida_sums')
      import; type(c_ptr), value :: ctx; real(c_double), intent(out) :: energia, energia_ref, temp
      integer(c_int32_t), intent(out) :: n_mobile
    end function
    ! observables: counts(nbins) are integer(8); type_mask bit z selects element z (1 Li, 2 CG, 3 F)
    integer(c_int) function dml_density_profile(ctx, zlo, zhi, nbins, type_mask, counts) bind(C, name='dml_density_profile')
      import; type(c_ptr), value :: ctx; real(c_double), value :: zlo, zhi; integer(c_int32_t), value :: nbins, type_mask
      integer(c_int64_t), intent(out) :: counts(*)
    end function
    ! host object model sync: slots whose occupant / element / membership changed since the previous call (kind bits: 1 new atom,
    ! 2 previous atom gone, 4 element changed, 8 left hs%ref, 16 left gcmc) -> attach / detach_all / setz exactly those atoms
    integer(c_int) function dml_membership_changes(ctx, max_changes, slot, kind, uid_now, z_now, n_changes) &
        bind(C, name='dml_membership_changes')
      import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: max_changes
      integer(c_int32_t), intent(out) :: slot(*), kind(*), uid_now(*), z_now(*), n_changes
    end function
    integer(c_int) function dml_set_scalars(ctx, s) bind(C, name='dml_set_scalars')
      import; type(c_ptr), value :: ctx; type(dml_scalars), intent(in) :: s
    end function
    integer(c_int) function dml_get_scalars(ctx, s) bind(C, name='dml_get_scalars')
      import; type(c_ptr), value :: ctx; type(dml_scalars), intent(out) :: s
    end function
    integer(c_int) function dml_get_counters(ctx, c) bind(C, name='dml_get_counters')
      import; type(c_ptr), value :: ctx; type(dml_counters), intent(out) :: c
    end function
    integer(c_int) function dml_set_chunk_template(ctx, nchunk, cpos, cpos_old, dist, rhomedia) bind(C, name='dml_set_chunk_template')
      import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: nchunk
      real(c_double), intent(in) :: cpos(3,*), cpos_old(3,*); real(c_double), value :: dist, rhomedia
    end function
    ! rows(width,n): row i holds the hs indices MINUS ONE of the nn(i) neighbours of hs%a(i) (C order [n][width]); rc = 1: width too small
    integer(c_int) function dml_get_neighbors(ctx, n, width, nn, rows) bind(C, name='dml_get_neighbors')
      import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: n, width
      integer(c_int32_t), intent(out) :: nn(*), rows(width,*)
    end function
    integer(c_int) function dml_upload_positions(ctx, n, pos, pos_old) bind(C, name='dml_upload_positions')
      import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: n
      real(c_double), intent(in) :: pos(3,*); type(c_ptr), value :: pos_old      ! c_null_ptr: pos_old stays resident
    end function
    integer(c_int) function dml_download_frame(ctx, n, pos, z) bind(C, name='dml_download_frame')
      import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: n
      real(c_double), intent(out) :: pos(3,*); integer(c_int32_t), intent(out) :: z(*)
    end function
    ! DML_RNG_REFERENCE (rng_mode = 2): hand the state of ran / gasdev over to the device (and take it back)
    integer(c_int) function dml_set_rng_state(ctx, r) bind(C, name='dml_set_rng_state')
      import; type(c_ptr), value :: ctx; type(dmlh_rng), intent(in) :: r
    end function
    integer(c_int) function dml_get_rng_state(ctx, r) bind(C, name='dml_get_rng_state')
      import; type(c_ptr), value :: ctx; type(dmlh_rng), intent(out) :: r
    end function
    ! multi-GPU: z-slab decomposition of one box (one image / MPI rank per GPU; id128 = 128 bytes from dml_comm_unique_id on rank 0)
    integer(c_int) function dml_comm_unique_id(id128) bind(C, name='dml_comm_unique_id')
      import; character(kind=c_char), intent(out) :: id128(128)
    end function
    integer(c_int) function dml_comm_init(ctx, id128, rank, nranks) bind(C, name='dml_comm_init')
      import; type(c_ptr), value :: ctx; character(kind=c_char), intent(in) :: id128(128); integer(c_int32_t), value :: rank, nranks
    end function
    integer(c_int) function dml_slab_plan(n, z, nranks, lo, hi, cuts) bind(C, name='dml_slab_plan')
      import; integer(c_int32_t), value :: n, nranks; real(c_double), intent(in) :: z(*); real(c_double), value :: lo, hi
      real(c_double), intent(out) :: cuts(*)
    end function
    integer(c_int) function dml_slab_setup(ctx, zlo, zhi) bind(C, name='dml_slab_setup')
      import; type(c_ptr), value :: ctx; real(c_double), value :: zlo, zhi
    end function
    integer(c_int) function dml_slab_step(ctx, nsteps) bind(C, name='dml_slab_step')
      import; type(c_ptr), value :: ctx; integer(c_int32_t), value :: nsteps
    end function
    integer(c_int) function dml_gr(ctx, rmax, nbins, type_mask, counts, n_selected) bind(C, name='dml_gr')
      import; type(c_ptr), value :: ctx; real(c_double), value :: rmax; integer(c_int32_t), value :: nbins, type_mask
      integer(c_int64_t), intent(out) :: counts(*); integer(c_int32_t), intent(out) :: n_selected
    end function
  end interface
  public :: dml_create, dml_destroy, dml_last_error, dml_upload, dml_download, dml_test_update, dml_fuerza, dml_ermak_a, &
            dml_ermak_b, dml_cbrownian_hs, dml_overlap_moveback, dml_msd_book, dml_promote, dml_gcmc_run, dml_calc_rho, &
            dml_maxz, dml_bloques, dml_step, dml_reset_try_depo, dml_salida_sums, dml_density_profile, dml_gr, dml_membership_changes, &
            dml_set_scalars, dml_get_scalars, dml_get_counters, dml_set_chunk_template, dml_get_neighbors, dml_upload_positions, &
            dml_download_frame, dml_set_rng_state, dml_get_rng_state, dml_comm_unique_id, dml_comm_init, dml_slab_plan, dml_slab_setup, &
            dml_slab_step

contains

  ! Fills dml_config from the variables dana reads in entrada()/config_run() (src/dana.F90:309-327,399-427) and creates the ctx.
  subroutine dmlf_create(ctx, capacity, box, h, nb_dcut, z0, z1, zmax, prob, dif_sc, dif_sei, integrador, reservoir, &
                         act, nadj, eps, r0, seed)
    type(dml_ctx_t), intent(out) :: ctx
    integer, intent(in)          :: capacity, reservoir, nadj, seed
    real(c_double), intent(in)   :: box(3), h, nb_dcut, z0, z1, zmax, prob, dif_sc, dif_sei, act, eps(3,3), r0(3,3)
    logical, intent(in)          :: integrador
    type(dml_config) :: c
    integer :: k, m
    c%device = 0; c%capacity = capacity; c%box = box; c%pbc = [1, 1, 0]
    c%rcut = 3.2_c_double; c%nb_dcut = nb_dcut
    do k = 1, 3; do m = 1, 3; c%eps((k-1)*3+m) = eps(k,m); c%r0((k-1)*3+m) = r0(k,m); end do; end do
    c%mass = 6.94_c_double
    c%h = h; c%gama = 1._c_double; c%Tsist = 300._c_double
    c%kB_ui = 8.617330350e-5_c_double*(96.485_c_double*100._c_double)      ! src/dana.F90:22-24
    c%kB_ui_gcmc = c%kB_ui                                                  ! replace by gems_constants::kB_ui (src/dana.F90:594)
    c%dif_sc = dif_sc; c%dif_sei = dif_sei; c%z_sei = 80._c_double
    c%prob = prob; c%z0 = z0; c%z1 = z1; c%zmax = zmax; c%tau = 0.1_c_double
    c%act = act; c%nadj = nadj
    c%integrador = merge(1, 0, integrador); c%reservoir = reservoir
    c%rng_mode = 0; c%seed = int(seed, c_int64_t); c%strict_order = 0
    call dmlf_check(ctx, dml_create(ctx%h, c))
  end subroutine

  ! C error convention -> gems_errors convention: rc/=0 ends the run through werr (src/Errors.f90:59-87)
  subroutine dmlf_check(ctx, rc)
    use gems_errors, only: werr
    type(dml_ctx_t), intent(in) :: ctx
    integer(c_int), intent(in)  :: rc
    character(kind=c_char), pointer :: p(:)
    character(:), allocatable :: msg
    integer :: i
    if (rc == 0) return
    msg = 'libdml error'
    if (c_associated(ctx%h)) then
      call c_f_pointer(dml_last_error(ctx%h), p, [256])
      msg = ''
      do i = 1, 256
        if (p(i) == c_null_char) exit
        msg = msg//p(i)
      end do
    end if
    call werr(msg, .true.)
  end subroutine

end module dml_cuda
