// dml_slab.cuh — z-slab domain decomposition of one large box over several GPUs (BASELINE config 4).
//
// z is the non-periodic axis (dana.F90:483-484), so 1-D slabs along z have two faces and no wrap.  Rank k owns the
// particles with zlo <= z < zhi and keeps read-only copies ("ghosts") of its neighbours' particles within one list
// radius (rcut + nb_dcut) of its faces.  Ghosts are candidates for the neighbour rows of owned particles and sources
// of force on them; they have no rows of their own and are never integrated.  Face data moves with grouped
// ncclSend/ncclRecv on the ctx stream (NVLink 5 / NVSwitch: every peer is equidistant, so slab k <-> GPU k is arbitrary).
// NCCL is resolved with dlopen at the first comm call: the library itself has no link-time NCCL dependency.
//
// Round-1 scope: set-up (ghost selection + exchange), per-step halo refresh, and the list build / pair force on top of
// them — validated against the single-GPU result (same pair sets, forces within 1e-12).  Particle migration at
// rebuild, the global rebuild decision and the global piston are the next step.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace dml {

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi *nccl_api() {
  static NcclApi a;
  if (a.h) return &a;
  a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!a.h) return nullptr;
#define SYM(f, name) *(void **)(&a.f) = dlsym(a.h, name)
  SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  if (!a.GetUniqueId || !a.CommInitRank || !a.Send || !a.Recv || !a.GroupStart || !a.GroupEnd) { a.h = nullptr; return nullptr; }
  return &a;
}

// face selection: owned particles within w of the upper / lower face (order fixed by a scan-free atomic cursor; the
// same lists are reused by every halo refresh until the next set-up, so sender and receiver stay aligned)
__global__ void k_slab_select(const double4 *__restrict__ posm, int *__restrict__ send_lo, int *__restrict__ send_hi,
                              int *__restrict__ counts, double zlo, double zhi, double w, int has_lo, int has_hi, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double4 p = ld_rec_nc(&posm[s]);
  if (!(meta_of(p) & MF_TYPE)) return;
  if (has_lo && p.z < zlo + w) send_lo[atomicAdd(&counts[0], 1)] = s;
  if (has_hi && p.z >= zhi - w) send_hi[atomicAdd(&counts[1], 1)] = s;
}
__global__ void k_slab_pack(const double4 *__restrict__ posm, const int *__restrict__ uid, const int *__restrict__ list, int cnt,
                            double4 *__restrict__ out_p, int *__restrict__ out_uid) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  int s = list[i];
  st_rec(&out_p[i], ld_rec_nc(&posm[s]));
  if (out_uid) out_uid[i] = uid[s];
}
// received records become ghosts: keep position and element, drop the owner's membership flags
__global__ void k_slab_mark(double4 *__restrict__ posm, int *__restrict__ slot_b, unsigned char *__restrict__ halo_of, int first, int cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  int s = first + i;
  double4 p = ld_rec(&posm[s]);
  long long m = meta_of(p);
  long long nm = (m & MF_TYPE) | MF_GHOST | ((m & (MF_REF | MF_GREF)) ? MF_GREF : 0);
  p.w = meta_as_double(with_disp(nm, DISP_INF));
  st_rec(&posm[s], p);
  if (slot_b) slot_b[s] = s;
  if (halo_of) halo_of[s] = 0;
}

} // namespace dml
