! fortran/dml_sync.F90 — the host object model <-> the flat arrays of include/dml.h.
!
! NOTE: no Fortran compiler exists in this image: reviewed against src/Groups.F90 / src/Neighbor.F90 of the reference, never compiled.
! tools/dana_host.cpp issues the same calls with the same array contents and is what CI runs (tools/test_cases.sh).
!
! Identity model (DESIGN.md section 2): the device indexes everything by "slot" = index in hs%a(:) minus one; two more integers
! per atom carry the orderings the reference takes from its pointer lists: uid = position in sys%alist minus one (creation rank:
! every "in list order" loop of dana.F90) and slot_b = index in hs%b%a(:) minus one (chain order inside a cell, Cells.F90:267-302).
module dml_sync
  use, intrinsic :: iso_c_binding
  use gems_constants, only: dp
  use gems_groups,    only: atom, atom_dclist, group, igroup, sys
  use gems_neighbor,  only: ngroup
  use dml_cuda
  implicit none
  private
  public :: pack_atoms, unpack_atoms, apply_membership_changes, shift_chunk_template, gpu_upload, gpu_download

  ! staging arrays, (3,n) so that column i is atom slot i-1 in C order [n][3]
  real(c_double), allocatable, public     :: xpos(:,:), xvel(:,:), xacel(:,:), xforce(:,:), xpos_old(:,:), xold_cg(:,:), xepot(:)
  integer(c_int32_t), allocatable, public :: xz(:), xflags(:), xuid(:), xslot_b(:)

contains

  subroutine ensure_stage(n)
    integer, intent(in) :: n
    if (allocated(xz)) then
      if (size(xz) >= n) return
      deallocate(xpos, xvel, xacel, xforce, xpos_old, xold_cg, xepot, xz, xflags, xuid, xslot_b)
    end if
    allocate(xpos(3,n), xvel(3,n), xacel(3,n), xforce(3,n), xpos_old(3,n), xold_cg(3,n), xepot(n), xz(n), xflags(n), xuid(n), xslot_b(n))
  end subroutine

  ! hs%a(1:hs%amax) -> flat arrays.  gcmc may be an unattached group (no grand-canonical reservoir): its id is then 0 and no atom has it.
  subroutine pack_atoms(hs, gcmc, n)
    type(ngroup), intent(inout), target :: hs
    type(group), intent(in), target     :: gcmc
    integer, intent(out)                :: n
    type(atom_dclist), pointer :: la
    type(atom), pointer        :: o
    integer :: i, j
    n = hs%amax
    call ensure_stage(max(n, 1))
    xz(1:n) = 0; xflags(1:n) = 0; xuid(1:n) = -1; xslot_b(1:n) = 0
    xpos(:,1:n) = 0._dp; xvel(:,1:n) = 0._dp; xacel(:,1:n) = 0._dp; xpos_old(:,1:n) = 0._dp; xold_cg(:,1:n) = 1.e8_dp
    do i = 1, n
      if (.not. associated(hs%a(i)%o)) cycle                          ! empty slot (Groups.F90:1118)
      if (associated(hs%a(i)%o, target=hs%limbo)) then                 ! parked on the limbo sentinel (Neighbor.F90:262-267)
        xflags(i) = DML_F_LIMBO
        cycle
      end if
      o => hs%a(i)%o
      xpos(:,i) = o%pos(:); xvel(:,i) = o%vel(:); xacel(:,i) = o%acel(:)
      xpos_old(:,i) = o%pos_old(:); xold_cg(:,i) = o%old_cg(:)
      xz(i) = o%z
      if (o%gri(hs%ref) /= 0) xflags(i) = ior(xflags(i), DML_F_REF)
      if (gcmc%id /= 0) then
        if (o%gri(gcmc) /= 0) xflags(i) = ior(xflags(i), DML_F_GCMC)
      end if
      if (o%skip) xflags(i) = ior(xflags(i), DML_F_SKIP)
      xslot_b(i) = o%gid(hs%b) - 1
    end do
    ! creation rank = position in sys%alist (every list appends at its tail: lib/fpt/include/cdlist_body.inc:43-58)
    la => sys%alist
    do j = 1, sys%nat
      la => la%next
      i = la%o%gid(hs)
      if (i > 0) xuid(i) = j - 1
    end do
  end subroutine

  ! flat arrays -> the atoms of hs%a(1:n) (positions, velocities, forces, element, skip flag); membership changes are NOT applied
  ! here: apply_membership_changes does that from the device's change report
  subroutine unpack_atoms(hs, n)
    type(ngroup), intent(inout), target :: hs
    integer, intent(in)                 :: n
    type(atom), pointer :: o
    integer :: i
    do i = 1, min(n, hs%amax)
      if (.not. associated(hs%a(i)%o)) cycle
      if (associated(hs%a(i)%o, target=hs%limbo)) cycle
      o => hs%a(i)%o
      o%pos(:) = xpos(:,i); o%vel(:) = xvel(:,i); o%acel(:) = xacel(:,i); o%force(:) = xforce(:,i)
      o%pos_old(:) = xpos_old(:,i); o%old_cg(:) = xold_cg(:,i); o%epot = xepot(i)
      if (o%z /= xz(i) .and. xz(i) > 0) call o%setz(int(xz(i)))
      o%skip = iand(xflags(i), DML_F_SKIP) /= 0
    end do
  end subroutine

  subroutine gpu_upload(hs, gcmc)
    type(ngroup), intent(inout), target :: hs
    type(group), intent(in), target     :: gcmc
    integer :: n
    call pack_atoms(hs, gcmc, n)
    call dmlf_check(gpu, dml_upload(gpu%h, int(n, c_int32_t), xpos, xvel, xacel, xpos_old, xold_cg, xz, xflags, xuid, xslot_b))
  end subroutine

  subroutine gpu_download(hs)
    type(ngroup), intent(inout), target :: hs
    type(dml_counters) :: c
    call dmlf_check(gpu, dml_get_counters(gpu%h, c))
    call ensure_stage(int(c%n_slots))
    call dmlf_check(gpu, dml_download(gpu%h, c%n_slots, xpos, xvel, xacel, xforce, xepot, xpos_old, xold_cg, xz, xflags, xuid, xslot_b))
    call unpack_atoms(hs, int(c%n_slots))
  end subroutine

  ! The device changes membership in gcmc_run (insert dana.F90:655-674, delete 703-706), bloques (716-773), atom_pbc /
  ! overlap_moveback (Li -> F) and the promotion loop (F -> CG, 228-236).  dml_membership_changes reports exactly the slots
  ! that changed since the last call; the host lists are patched in ascending slot order, deletions first (so that a slot that
  ! was vacated and re-occupied in the same interval is handled as "gone" then "new").
  !   kind bits: 1 new atom in the slot, 2 previous atom gone, 4 element changed, 8 left hs%ref, 16 left gcmc
  subroutine apply_membership_changes(hs, gcmc, template)
    type(ngroup), intent(inout), target :: hs
    type(group), intent(inout), target  :: gcmc
    type(atom), intent(in), target      :: template              ! an atom whose groups a created atom joins (dana.F90:655-674)
    integer(c_int32_t), allocatable :: slot(:), kind(:), uid_now(:), z_now(:)
    integer(c_int32_t) :: nch
    type(atom), pointer :: o
    integer :: k, i, maxc
    maxc = 65536
    allocate(slot(maxc), kind(maxc), uid_now(maxc), z_now(maxc))
    do
      call dmlf_check(gpu, dml_membership_changes(gpu%h, int(maxc, c_int32_t), slot, kind, uid_now, z_now, nch))
      do k = 1, min(int(nch), maxc)
        i = slot(k) + 1
        if (iand(kind(k), 2) /= 0) then                               ! the previous occupant was destroyed on the device
          if (i <= hs%amax) then
            if (associated(hs%a(i)%o)) then
              if (.not. associated(hs%a(i)%o, target=hs%limbo)) then
                o => hs%a(i)%o
                call o%dest()                                         ! detaches from gcmc, hs, sys (Groups.F90:433-467)
                deallocate(o)
              end if
            end if
          end if
        end if
      end do
      do k = 1, min(int(nch), maxc)
        i = slot(k) + 1
        if (iand(kind(k), 1) /= 0) then                               ! a new atom occupies the slot: same attach order as dana.F90:671-674
          allocate(o); call o%init()
          call o%setz(int(z_now(k)))
          o%pbc(:) = template%pbc(:)
          call sys%attach(o)
          if (z_now(k) /= 2) call hs%ref%attach(o)
          call hs%b%attach(o)
          call hs%attach(o)                                           ! must come last (dana.F90:471-480); takes the lowest free index = slot
          if (gcmc%id /= 0 .and. z_now(k) /= 2) call gcmc%attach(o)
          if (o%gid(hs) /= i) call dmlf_host_error('host and device disagree on the index of a created atom')
          o => null()
        else if (i <= hs%amax) then
          if (.not. associated(hs%a(i)%o)) cycle
          o => hs%a(i)%o
          if (iand(kind(k), 4) /= 0) call o%setz(int(z_now(k)))       ! Li -> F (atom_pbc / overlap_moveback) or F -> CG
          if (iand(kind(k), 8) /= 0) call hs%ref%detach(o)            ! promotion loop, dana.F90:228-236
          if (iand(kind(k), 16) /= 0 .and. gcmc%id /= 0) call gcmc%detach(o)
        end if
      end do
      if (nch <= maxc) exit
    end do
    deallocate(slot, kind, uid_now, z_now)
  end subroutine

  subroutine dmlf_host_error(msg)
    use gems_errors, only: werr
    character(*), intent(in) :: msg
    call werr(msg, .true.)
  end subroutine

  ! bloques fired: the chunk template moves up by dist (dana.F90:762-763)
  subroutine shift_chunk_template(chunk, dist)
    type(group), intent(inout) :: chunk
    real(dp), intent(in)       :: dist
    type(atom_dclist), pointer :: la
    integer :: j
    la => chunk%alist
    do j = 1, chunk%nat
      la => la%next
      la%o%pos(3) = la%o%pos(3) + dist
      la%o%pos_old(3) = la%o%pos_old(3) + dist
    end do
  end subroutine

end module dml_sync
