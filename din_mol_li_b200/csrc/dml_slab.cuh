// dml_slab.cuh — z-slab domain decomposition of one large box over several GPUs (BASELINE config 4).
//
// z is the non-periodic axis (dana.F90:483-484), so 1-D slabs along z have two faces and no wrap.  Rank k owns the
// particles with zlo <= z < zhi and keeps read-only copies ("ghosts") of its neighbours' particles within one list
// radius (rcut + nb_dcut) of its faces.  Ghosts are candidates for the neighbour rows of owned particles and sources
// of force on them; they have no rows of their own and are never integrated.  Face data moves with grouped
// ncclSend/ncclRecv on the ctx stream (NVLink 5 / NVSwitch: every peer is equidistant, so slab k <-> GPU k is arbitrary).
// NCCL is resolved with dlopen at the first comm call: the library itself has no link-time NCCL dependency.
//
// dml_slab_step runs dana's loop body (Ermak + piston, the configuration of BASELINE config 4) on the decomposed box:
//   - the rebuild decision is global: every rank reduces its own two largest squared displacements, the pairs are
//     all-gathered and every rank merges them to the same decision (top-2 merging is order independent);
//   - at a rebuild the particles that left the slab migrate to the neighbour (full state), holes are refilled, ghosts are
//     re-selected and the rows rebuilt;
//   - rho of the piston is global (all-gather of the per-rank counts), maxz is then the same arithmetic on every copy;
//   - overlap_moveback treats ghosts that are mobile on their owner as full participants of the conflict components, visited
//     in creation-rank order like everybody else (their visit acts on the owned atoms in range), so both sides of a face
//     replay the same pair decisions from the same inputs; what a rank cannot see (third atoms beyond its halo) makes the
//     multi-GPU trajectory an approximation of the single-GPU one — validated on observables.
// Noise is Philox keyed by (seed; creation rank, step), i.e. independent of the decomposition.
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
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
  SYM(GetErrorString, "ncclGetErrorString"); SYM(AllGather, "ncclAllGather");
#undef SYM
  if (!a.GetUniqueId || !a.CommInitRank || !a.Send || !a.Recv || !a.GroupStart || !a.GroupEnd || !a.AllGather) { a.h = nullptr; return nullptr; }
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
// both faces in one launch
__global__ void k_slab_pack2(const double4 *__restrict__ posm, const int *__restrict__ uid, const int *__restrict__ list_lo, int cnt_lo,
                             double4 *__restrict__ out_lo, int *__restrict__ uid_lo, const int *__restrict__ list_hi, int cnt_hi,
                             double4 *__restrict__ out_hi, int *__restrict__ uid_hi) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt_lo) {
    const int s = list_lo[i];
    st_rec(&out_lo[i], ld_rec_nc(&posm[s]));
    if (uid_lo) uid_lo[i] = uid[s];
  } else if (i < cnt_lo + cnt_hi) {
    const int k = i - cnt_lo, s = list_hi[k];
    st_rec(&out_hi[k], ld_rec_nc(&posm[s]));
    if (uid_hi) uid_hi[k] = uid[s];
  }
}
// received records become ghosts: keep position, element and skip flag, drop the owner's membership flags.  With old_cg
// (refresh inside a step) the move of the ghost since the step started is measured: it goes into the record (prefilter of
// k_ov_detect) and into step_disp_bits, because the gather-skip bound must cover the ghosts' moves as well as the local ones.
__global__ void k_slab_mark(double4 *__restrict__ posm, int *__restrict__ slot_b, unsigned char *__restrict__ halo_of, int first, int cnt,
                            const double *__restrict__ old_cg, DevScal *__restrict__ sc, Geo g) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float df = 0.0f;
  if (i < cnt) {
    int s = first + i;
    double4 p = ld_rec(&posm[s]);
    long long m = meta_of(p);
    long long nm = (m & (MF_TYPE | MF_SKIP)) | MF_GHOST | ((m & (MF_REF | MF_GREF)) ? MF_GREF : 0);
    unsigned int bits = DISP_INF;
    if (old_cg) {
      double dx = p.x - old_cg[3 * s], dy = p.y - old_cg[3 * s + 1], dz = p.z - old_cg[3 * s + 2];
      dx = dx - g.box[0] * round(dx * g.one_box[0]); dy = dy - g.box[1] * round(dy * g.one_box[1]);
      df = __double2float_ru(sqrt(dx * dx + dy * dy + dz * dz)) * 1.000001f;
      bits = (unsigned int)__float_as_int(df);
    }
    p.w = meta_as_double(with_disp(nm, bits));
    st_rec(&posm[s], p);
    if (slot_b) slot_b[s] = s;
    if (halo_of) halo_of[s] = 0;
  }
  if (old_cg) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) df = fmaxf(df, __shfl_xor_sync(0xffffffffu, df, o));
    if ((threadIdx.x & 31) == 0 && df > 0.0f) atomicMax(&sc->step_disp_bits, (unsigned int)__float_as_int(df));
  }
}

// ---- migration at a rebuild ------------------------------------------------------------------------------------
constexpr int MIG_D = 13;     // doubles per migrant: record (4), vel (3), acel (3), old_cg (3)
__global__ void k_slab_mig_select(const double4 *__restrict__ posm, int *__restrict__ send_lo, int *__restrict__ send_hi,
                                  int *__restrict__ counts, double zlo, double zhi, int has_lo, int has_hi, int n_owned) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_owned) return;
  double4 p = ld_rec_nc(&posm[s]);
  if (!(meta_of(p) & MF_TYPE)) return;
  if (has_lo && p.z < zlo) send_lo[atomicAdd(&counts[4], 1)] = s;
  else if (has_hi && p.z >= zhi) send_hi[atomicAdd(&counts[5], 1)] = s;
}
// full state of the leavers into the send buffers; their slots become holes
__global__ void k_slab_mig_pack(double4 *__restrict__ posm, const double *__restrict__ vel, const double *__restrict__ acel,
                                const double *__restrict__ old_cg, const int *__restrict__ uid, RowHead *__restrict__ rh,
                                const int *__restrict__ list, int cnt, double *__restrict__ out_d, int *__restrict__ out_i) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  const int s = list[i];
  const double4 p = ld_rec(&posm[s]);
  double *o = out_d + (size_t)i * MIG_D;
  o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.w;
  for (int k = 0; k < 3; ++k) { o[4 + k] = vel[3 * s + k]; o[7 + k] = acel[3 * s + k]; o[10 + k] = old_cg[3 * s + k]; }
  out_i[i] = uid[s];
  st_rec(&posm[s], make_double4(0.0, 0.0, 0.0, meta_as_double(0)));
  rh_store_plain(&rh[s], s * ROW_W, 0, 0, 255);
}
__global__ void k_slab_holes(const double4 *__restrict__ posm, int *__restrict__ holes, int *__restrict__ counts, int n_owned) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_owned) return;
  if (!(meta_of(ld_rec_nc(&posm[s])) & MF_TYPE)) holes[atomicAdd(&counts[6], 1)] = s;
}
// arrivals take the holes first, then the slots behind the owned region
__global__ void k_slab_mig_unpack(double4 *__restrict__ posm, double *__restrict__ vel, double *__restrict__ acel, double *__restrict__ pos_old,
                                  double *__restrict__ old_cg, int *__restrict__ uid, int *__restrict__ slot_b, double4 *__restrict__ fe,
                                  RowHead *__restrict__ rh, unsigned char *__restrict__ halo_of,
                                  const double *__restrict__ in_d, const int *__restrict__ in_i, int cnt, int first,
                                  const int *__restrict__ holes, int nholes, int n_owned) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  const int q = first + i;                                   // rank among all arrivals of this rebuild
  const int s = q < nholes ? holes[q] : n_owned + (q - nholes);
  const double *o = in_d + (size_t)i * MIG_D;
  const long long m = with_disp(__double_as_longlong(o[3]) & 0xffffffffll, DISP_INF);
  st_rec(&posm[s], make_double4(o[0], o[1], o[2], meta_as_double(m)));
  for (int k = 0; k < 3; ++k) { vel[3 * s + k] = o[4 + k]; acel[3 * s + k] = o[7 + k]; old_cg[3 * s + k] = o[10 + k]; pos_old[3 * s + k] = o[k]; }
  uid[s] = in_i[i]; slot_b[s] = s; halo_of[s] = 0;
  st_rec(&fe[s], make_double4(0.0, 0.0, 0.0, 0.0));
  rh_store_plain(&rh[s], s * ROW_W, 0, 0, 255);
}
// fresh ghosts: pos_old = old_cg = pos, no velocity, no row
__global__ void k_slab_ghost_init(const double4 *__restrict__ posm, double *__restrict__ pos_old, double *__restrict__ old_cg,
                                  double *__restrict__ vel, double *__restrict__ acel, RowHead *__restrict__ rh, int first, int cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  const int s = first + i;
  const double4 p = ld_rec_nc(&posm[s]);
  const double q[3] = {p.x, p.y, p.z};
  for (int k = 0; k < 3; ++k) { pos_old[3 * s + k] = q[k]; old_cg[3 * s + k] = q[k]; vel[3 * s + k] = 0.0; acel[3 * s + k] = 0.0; }
  rh_store_plain(&rh[s], s * ROW_W, 0, 0, 255);
}
// previous positions of the ghosts = where they are when the step starts (their owners' old_cg of this step)
__global__ void k_slab_ghost_save(const double4 *__restrict__ posm, double *__restrict__ old_cg, int first, int cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  const int s = first + i;
  const double4 p = ld_rec_nc(&posm[s]);
  old_cg[3 * s] = p.x; old_cg[3 * s + 1] = p.y; old_cg[3 * s + 2] = p.z;
}
// F -> CG promotion (dana.F90:228-236) and the census of calc_rho (521-549) over the OWNED particles; the counts of all
// ranks are gathered (with the displacements of the test_update in front of it) and k_slab_tu_final turns them into the same rho everywhere
__global__ void __launch_bounds__(TPB) k_slab_promote_count(double4 *__restrict__ posm, DevScal *__restrict__ sc, int *__restrict__ out, int n_owned) {
  const double z0 = sc->z0, zl = sc->zmax;
  int c = 0, dref = 0, nref = 0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_owned; s += gridDim.x * blockDim.x) {
    double4 p = ld_rec(&posm[s]);
    long long m = meta_of(p);
    if ((m & MF_REF) && (m & MF_TYPE) == 3) {
      dref++;
      m = (m & ~(MF_TYPE | MF_REF | MF_GCMC)) | 2;
      p.w = meta_as_double(m); st_rec(&posm[s], p);
    }
    if (m & MF_REF) nref++;
    if ((m & MF_TYPE) && p.z > z0 && p.z < zl) c++;
  }
  c = __reduce_add_sync(0xffffffffu, c); dref = __reduce_add_sync(0xffffffffu, dref); nref = __reduce_add_sync(0xffffffffu, nref);
  if ((threadIdx.x & 31) == 0) { if (c) atomicAdd(&out[0], c); if (dref) atomicAdd(&out[1], dref); if (nref) atomicAdd(&out[2], nref); }
}
// What one rank contributes to the merged all-gather of a test_update of the decomposed box: its two largest squared displacements
// and (second test_update of a step) the census of k_slab_promote_count.
struct SlabTU { double a1, a2; int cnt[4]; };
// test_update's decision from the gathered contributions (d_top2_final) and, with_rho, calc_rho + the msd bookkeeping from the
// gathered census
__global__ void k_slab_tu_final(const SlabTU *__restrict__ all, int nranks, SlabTU *__restrict__ own, DevScal *__restrict__ sc,
                                unsigned int *__restrict__ lay, Geo g, double nb_dcut, double rmax_f, double rmax_o, int with_rho, double area) {
  d_top2_final(reinterpret_cast<const double *>(all), nranks, sc, lay, g, nb_dcut, rmax_f, rmax_o, (int)(sizeof(SlabTU) / sizeof(double)));
  __syncthreads();
  if (with_rho && threadIdx.x == 0) {
    long long c = 0, dref = 0, nref = 0;
    for (int r = 0; r < nranks; ++r) { c += all[r].cnt[0]; dref += all[r].cnt[1]; nref += all[r].cnt[2]; }
    sc->msd_t = sc->msd_t / (double)(nref + dref);          // dana.F90:201-202 (before the promotion loop; local sum, global count)
    sc->msd_max = fmax(sc->msd_max, sc->msd_t);
    sc->nat_ref -= own->cnt[1];
    sc->rho = (double)c / (area * (sc->zmax - sc->z0));
    sc->step_disp_bits = 0u;
    own->cnt[0] = own->cnt[1] = own->cnt[2] = own->cnt[3] = 0;
  }
}

} // namespace dml
