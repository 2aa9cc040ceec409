! fortran/Neighbor_gpu.F90 — module gems_neighbor over libdml.so: replaces src/Neighbor.F90 of the reference in the build
! (src/Makefile.am lists Neighbor.F90; list this file instead).  Every public name dana.F90 and the rest of GEMS use is kept
! (src/Neighbor.F90:32-112): type ngroup with ref / b / rcut / rcut2 / mnb / nn / list / listed and init / dest / attach_atom /
! detach_atom / setrc, the ngroup_dl / ngroup_vop containers, ngindex, nb_dcut, nupd_vlist, nn_vlist, test_update, update.
!
! What changes: the membership bookkeeping (who is in ref, b and the ngroup itself, and under which index) stays on the host,
! exactly as in the reference, because the index an atom gets in hs%a(:) IS the device slot; the search itself — do_pbc,
! tessellate + sort of the cells, the displacement test, ngroup_cells / ngroup_verlet, the incremental list upkeep of an attach or
! detach — runs on the device behind dml_test_update / dml_gcmc_run.  nn(:) and list(:,:) are no longer maintained on the host;
! ngroup_pull_list fills them from the device for a caller that wants to look at them (mnb bounds the row length it can take).
!
! NOTE: no Fortran compiler exists in this image: written against the reference sources, never compiled.
module gems_neighbor
use, intrinsic :: iso_c_binding
use gems_program_types, only: boxed, box, mic
use gems_groups,        only: vdistance, sys, igroup, group, atom, atom_dclist
use gems_constants,     only: dp,cdm,dm,ui_ev
use gems_cells,         only: cgroup, map, n1cells, cell_pbc
use gems_errors
use dml_cuda

implicit none

type, extends(igroup), public :: ngroup
  type(group)   :: ref                                  ! The reference group
  type(cgroup)  :: b                                    ! The group of possible neighbors (cells are sorted on the device)
  real(dp)               :: rcut=1.e10_dp,rcut2=1.e10_dp
  integer                :: mnb=10000
  integer,allocatable    :: nn(:)                       ! filled by ngroup_pull_list only
  integer,allocatable    :: list(:,:)
  logical :: autoswitch=.true.
  logical :: listed=.false.
  contains
    procedure :: ngroup_construct
    procedure :: ngroup_attach_atom
    procedure :: ngroup_detach_atom
    procedure :: init => ngroup_construct
    procedure :: dest => ngroup_destroy
    procedure :: attach_atom => ngroup_attach_atom
    procedure :: detach_atom => ngroup_detach_atom
    procedure :: ngroup_init => ngroup_construct
    procedure :: ngroup_dest => ngroup_destroy
    procedure :: setrc => ngroup_setrc
    procedure :: pull_list => ngroup_pull_list
endtype

#define SOFT
#define _NODE ngroup_dl
#define _CLASS class(ngroup)
#include "dlist_header.inc"

#define _NODE ngroup_aop
#define _CLASS class(ngroup)
#include "arrayofptrs_header.inc"

#define _NODE ngroup_vop
#define _TYPE type(ngroup_aop)
#include "vector_header.inc"

type(ngroup_vop),public :: ngindex

real(dp),target,public  :: nb_dcut=1._dp      ! The shell length for verlet update criteria (passed to the device in dml_config)
integer , public        :: nupd_vlist = 0,&   ! refreshed from the device after every test_update
                           nn_vlist =0
public :: test_update, update

contains

#define SOFT
#define _NODE ngroup_dl
#define _CLASS class(ngroup)
#include "dlist_body.inc"

#define _NODE ngroup_vop
#define _TYPE type(ngroup_aop)
#include "vector_body.inc"

! ngroup events (host bookkeeping only: same order of operations as src/Neighbor.F90:128-269)
! =============

subroutine ngroup_construct(g)
class(ngroup),target  :: g
call g%igroup_construct()
call g%ref%init()
call g%b%init()
call ngindex%append()
ngindex%o(ngindex%size)%o=>g
end subroutine ngroup_construct

subroutine ngroup_destroy(g)
class(ngroup)  :: g
call g%igroup%dest()
call g%ref%dest()
call g%b%dest()
if (allocated(g%nn)) deallocate(g%nn)
if (allocated(g%list)) deallocate(g%list)
g%listed=.false.
end subroutine ngroup_destroy

subroutine ngroup_attach_atom(g,a)
! The atom takes the lowest free index of g%a(:) (Groups.F90:1083-1093): that index is its device slot.  While a list exists the
! device inserts the atom itself (gcmc_run: dml_gcmc_run); an attach from host code after the list was built is made known to the
! device by the next gpu_upload (dml_sync), which also marks the list as not built.
class(ngroup),target  :: g
class(atom),target    :: a
call g%igroup_attach_atom(a)
if (g%listed) g%listed=.false.
end subroutine ngroup_attach_atom

subroutine ngroup_detach_atom(g,a)
class(ngroup)        :: g
class(atom),target   :: a
call g%ref%detach(a)
call g%b%detach(a)
call g%igroup_detach_atom(a)
if (g%listed) g%listed=.false.
end subroutine ngroup_detach_atom

subroutine ngroup_setrc(g,rc)
class(ngroup)         :: g
real(dp),intent(in)   :: rc
g%rcut=rc
g%rcut2=rc*rc
g%b%rcut=rc+nb_dcut                                   ! src/Neighbor.F90:323-334; the device takes rcut and nb_dcut from dml_config
end subroutine ngroup_setrc

! The rows of the device as nn(:) / list(:,:) of the reference (hs indices, row order = stencil order x chain order)
subroutine ngroup_pull_list(g)
class(ngroup)  :: g
integer(c_int32_t), allocatable :: cnn(:), rows(:,:)
integer :: n, w, i, rc
n = g%amax
w = 64
do
  if (allocated(cnn)) deallocate(cnn, rows)
  allocate(cnn(n), rows(w,n))
  rc = dml_get_neighbors(gpu%h, int(n,c_int32_t), int(w,c_int32_t), cnn, rows)
  if (rc /= 1) exit
  w = 4*w
end do
call dmlf_check(gpu, int(rc, c_int))
call werr('neighbour row longer than mnb', w > g%mnb .and. maxval(cnn) > g%mnb)
if (allocated(g%nn)) deallocate(g%nn, g%list)
allocate(g%nn(n), g%list(n, max(maxval(cnn),1)))
g%nn(:) = cnn(:)
do i = 1, n
  g%list(i,1:cnn(i)) = rows(1:cnn(i),i) + 1
end do
end subroutine ngroup_pull_list

! Search (device)
! ===============

subroutine test_update()
! src/Neighbor.F90:668-713: do_pbc, cell sort, displacement test, rebuild when needed — one device call
type(dml_counters) :: c
integer :: i
call dmlf_check(gpu, dml_test_update(gpu%h))
call dmlf_check(gpu, dml_get_counters(gpu%h, c))
nupd_vlist = int(c%nupd_vlist)
do i = 1, ngindex%size
  ngindex%o(i)%o%listed = c%listed /= 0
end do
end subroutine test_update

subroutine update()
! src/Neighbor.F90:608-633: an unconditional rebuild is a test_update on a list marked as not built
integer :: i
do i = 1, ngindex%size
  ngindex%o(i)%o%listed = .false.
end do
call test_update()
end subroutine update

end module gems_neighbor
