// dml_coop.cuh — persistent cooperative kernels: one launch per call site of dana's loop.
//
// test_update and overlap_moveback are chains of data-dependent phases (bin -> scan -> scatter -> order -> count ->
// scan -> fill; init -> detect -> count -> alloc -> fill -> resolve -> apply).  As separate launches each phase costs
// a launch whether or not it has work (the rebuild decision is taken on the device).  Here the grid is sized to the
// machine (148 SMs x resident blocks), every phase is a grid-stride loop and phases are separated by grid.sync();
// a step that needs no rebuild leaves after the first phase.  Launched with cudaLaunchCooperativeKernel.
#pragma once
#include <cooperative_groups.h>
#include "dml_kernels.cuh"

namespace dml {
namespace cg = cooperative_groups;

__device__ __forceinline__ int block_sum_int(int v, int *sh /*>=33*/) {
  v = __reduce_add_sync(0xffffffffu, v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) { int x = lane < nw ? sh[lane] : 0; x = __reduce_add_sync(0xffffffffu, x); if (lane == 0) sh[32] = x; }
  __syncthreads();
  return sh[32];
}

// Exclusive scan of in[0..n) into out[0..n) by the whole grid; one grid.sync inside.  Each block owns a contiguous
// chunk (multiple of 1024).  total -> *total_out (written by block 0).  ZERO_IN clears in[] after reading.
template <bool ZERO_IN>
__device__ void coop_scan(cg::grid_group &grid, int *__restrict__ in, int *__restrict__ out, int n, int *__restrict__ sums,
                          int *__restrict__ total_out) {
  __shared__ int sh[34];
  __shared__ int wsum[8];
  const int nb = gridDim.x;
  const int chunk = (((n + nb - 1) / nb) + 1023) & ~1023;
  const int b0 = min(n, (int)blockIdx.x * chunk), b1 = min(n, b0 + chunk);
  int local = 0;
  for (int i = b0 + threadIdx.x; i < b1; i += blockDim.x) local += in[i];
  int tot = block_sum_int(local, sh);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
  grid.sync();
  int pre = 0, all = 0;
  for (int j = threadIdx.x; j < nb; j += blockDim.x) { int v = sums[j]; all += v; if (j < (int)blockIdx.x) pre += v; }
  int prefix = block_sum_int(pre, sh);
  int total = block_sum_int(all, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0 && total_out) *total_out = total;
  int carry = prefix;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = b0; base < b1; base += 1024) {
    int idx = base + threadIdx.x * 4;
    int v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = (idx + i < b1) ? in[idx + i] : 0; if (ZERO_IN && idx + i < b1) in[idx + i] = 0; }
    int t = v[0] + v[1] + v[2] + v[3], x = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    __syncthreads();
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      int q = lane < 8 ? wsum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, q, o); if (lane >= o) q += y; }
      if (lane < 8) wsum[lane] = q;
    }
    __syncthreads();
    int run = x - t + (w ? wsum[w - 1] : 0) + carry;
#pragma unroll
    for (int i = 0; i < 4; ++i) { if (idx + i < b1) out[idx + i] = run; run += v[i]; }
    carry += wsum[7];
  }
}

struct TUArgs {
  double4 *posm; double *pos_old; double *part; int *cell_of, *cell_cnt, *cell_start, *cell_cur, *sorted_slot, *sorted_raw, *sorted_cell; double4 *sorted_posm; float4 *sorted_posf;
  const int *slot_b; RowHead *rh; int *cols; unsigned char *bq, *halo_of; unsigned int *lay; int *sums; DevScal *sc; Geo g; int n, nct, force_sort, slack, lazy; double nb_dcut, rmax_f, rmax_o;
  // neighbours of the call inside dml_step folded into the first pass over the slots: fuse bit 0 = the initialisation of the
  // overlap_moveback that follows (k_ov_init), bit 1 = the write-back of the overlap_moveback that came before (k_ov_apply)
  int fuse; int *parent, *ovst, *comp_cnt, *ov_head; double *vel, *acel; const double *old_cg;
};

// test_update (Neighbor.F90:668-713) in one launch
__global__ void __launch_bounds__(TPB) k_test_update_coop(TUArgs A) {
  cg::grid_group grid = cg::this_grid();
  const int gsz = gridDim.x * blockDim.x, gt = blockIdx.x * blockDim.x + threadIdx.x;
  DevScal *sc = A.sc;
  // phase 0: do_pbc + per-block top-2 squared displacement
  if (gt == 0) sc->halo_flag = 0;
  __shared__ unsigned int s_lay[LAY_MAX];
  for (int i = threadIdx.x; i < A.g.nlay; i += blockDim.x) s_lay[i] = 0u;
  __syncthreads();
  const int lay_old = sc->lay_cur;
  const int was_listed = sc->listed;                             // read before the first grid.sync: block 0 rewrites it right after
  double a1 = -1.0, a2 = -1.0;
  if (gt == 0 && (A.fuse & 2)) {                                 // k_ov_apply's bookkeeping (dana.F90:939-941)
    if (sc->ch_later > sc->choques2) sc->choques2 = sc->ch_later;
    sc->overlap_passes += sc->any_active > 0 ? sc->any_active : 1;
  }
  if (gt == 0 && (A.fuse & 1)) { sc->again = 0; sc->n_roots = 0; sc->member_cursor = 0; sc->ch_later = 0; sc->any_active = 0; }
  for (int s = gt; s < A.n; s += gsz) {
    if (A.fuse & 2) d_ov_apply(A.posm, A.vel, A.acel, A.old_cg, A.ovst, s);
    double rd = d_pbc_disp(A.posm, A.pos_old, A.g, s);
    if (A.fuse & 1) {
      const long long m = meta_of(ld_rec(&A.posm[s]));
      A.parent[s] = s; A.comp_cnt[s] = 0; A.ov_head[s] = -1;
      A.ovst[s] = ((m & MF_SKIP) ? OV_SKIP : 0) | ((int)(m & MF_TYPE) << OV_TSHIFT);
    }
    if (rd >= 0.0) lay_note(s_lay, A.g, A.posm[s].z, rd);
    top2_merge(a1, a2, rd, -1.0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < A.g.nlay; i += blockDim.x) if (s_lay[i]) atomicMax(&A.lay[(lay_old ^ 1) * LAY_MAX + i], s_lay[i]);
  block_top2(a1, a2);
  if (threadIdx.x == 0) { A.part[2 * blockIdx.x] = a1; A.part[2 * blockIdx.x + 1] = a2; }
  grid.sync();
  // phase 1: every block reduces the partials to the same decision (top-2 merging is order independent)
  a1 = 1e-16; a2 = 1e-16;                                        // Neighbor.F90:643-644
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) top2_merge(a1, a2, A.part[2 * i], A.part[2 * i + 1]);
  block_top2(a1, a2);
  __shared__ int s_need;
  if (threadIdx.x == 0) {
    int need = (!was_listed) || (sqrt(a1) + sqrt(a2) > A.nb_dcut);   // Neighbor.F90:697-710
    s_need = need;
  }
  __syncthreads();
  const bool need = s_need != 0;
  if (gt == 0) {
    sc->d1 = a1; sc->d2 = a2; sc->need_rebuild = need ? 1 : 0;
    sc->dsum_tu = need ? 0.0 : sqrt(a1) + sqrt(a2); sc->maxz_disp = 0.0; sc->maxz_fac = 0.0;
    if (!need) sc->lay_cur = lay_old ^ 1;
    if (need) { sc->nupd++; sc->listed = 1; sc->nlimbo = 0; sc->hole_lo = 0; sc->rev_valid = 0; sc->rows_pending = 1; sc->cols_used = sc->cols_tail0; }
  }
  if (blockIdx.x == 0) {                                         // z-layer tables (see k_top2_final)
    if (need) { for (int i = threadIdx.x; i < 2 * LAY_MAX; i += blockDim.x) A.lay[i] = 0u; }
    else { for (int i = threadIdx.x; i < A.g.nlay; i += blockDim.x) A.lay[lay_old * LAY_MAX + i] = 0u; }
    __threadfence();
    __syncthreads();
    d_qtab(A.lay, sc, A.g, A.rmax_f, A.rmax_o);                  // skip tables for the consumers that follow
  }
  if (!(need || A.force_sort)) return;
  // phase 2: binning
  for (int s = gt; s < A.n; s += gsz) d_bin(A.posm, A.cell_of, A.cell_cnt, A.rh, A.halo_of, sc, A.g, need, s);
  grid.sync();
  coop_scan<true>(grid, A.cell_cnt, A.cell_start, A.nct, A.sums, A.cell_start + A.nct);
  grid.sync();
  for (int s = gt; s < A.n; s += gsz) d_scatter(A.posm, A.pos_old, A.cell_of, A.cell_start, A.cell_cur, A.sorted_raw, need, s);
  grid.sync();
  if (gt == 0 && need) sc->rows_asym = sc->halo_flag ? 1 : 0;
  {
    const int nsorted = A.cell_start[A.nct];
    for (int i = gt; i < nsorted; i += gsz)
      d_cell_rank(A.posm, A.slot_b, A.cell_of, A.cell_start, A.cell_cur, A.sorted_raw, A.sorted_slot, A.sorted_posm, A.sorted_posf, A.sorted_cell, i);
  }
  // the rows themselves (ngroup_cells, Neighbor.F90:465-548) are built from this snapshot by k_rows when a consumer first needs them
}

struct OVArgs {
  double4 *posm; double *vel, *acel; const double *old_cg; const RowHead *rh; const int *cols; const unsigned char *bq; const unsigned char *qmin; const unsigned int *lay; int *parent, *ovst, *comp_cnt, *comp_off,
      *members, *roots, *ov_head, *ov_next; const int *uid; OvRp rp_uovl; DevScal *sc; Geo g; Phys ph; unsigned int step; int n, guard_pass;
};

// overlap_moveback (dana.F90:849-943) in one launch (prob>=1)
__global__ void __launch_bounds__(TPB) k_overlap_coop(OVArgs A) {
  cg::grid_group grid = cg::this_grid();
  p_ov_init(A.posm, A.parent, A.ovst, A.comp_cnt, A.ov_head, A.sc, A.n);
  grid.sync();
  p_ov_detect(A.posm, A.old_cg, A.rh, A.cols, A.bq, A.lay, A.parent, A.ovst, A.sc, A.g, A.n, A.qmin);
  grid.sync();
  p_ov_link(A.parent, A.ovst, A.ov_head, A.ov_next, A.roots, A.sc, A.n);
  grid.sync();
  p_ov_resolve(A.posm, A.old_cg, A.rh, A.cols, A.bq, A.lay, A.ovst, A.roots, A.ov_head, A.ov_next, A.members, A.uid, A.rp_uovl,
               A.sc, A.g, A.ph, A.step, A.guard_pass);
  grid.sync();
  p_ov_apply(A.posm, A.vel, A.acel, A.old_cg, A.ovst, A.sc, A.n);
}

// Transposed rows (k_rev_count -> scan -> k_rev_fill -> k_rev_done) in one launch.  The multi-launch form costs four guarded
// launches in front of every pair-force call (measured at 1 M particles: 31 us per step, all of it idle outside rebuild steps);
// here an idle call is one launch whose blocks return on the guard, which nobody rewrites before the last phase.
struct RevArgs {
  const RowHead *rh; const int *cols; const double4 *posm; int *rev_start, *rev_len, *rev_cnt, *rev_cols;
  const unsigned char *bq; unsigned char *rev_bq; const unsigned char *halo_of; int halo_only; int *sums; DevScal *sc; int n; unsigned char *qmin;
};
__global__ void __launch_bounds__(TPB) k_rev_coop(RevArgs A) {
  REV_GUARD(A.sc);
  cg::grid_group grid = cg::this_grid();
  p_rev_count(A.rh, A.cols, A.posm, A.rev_len, A.rev_cnt, A.halo_of, A.halo_only, A.sc, A.n);
  grid.sync();
  coop_scan<true>(grid, A.rev_cnt, A.rev_start, A.n, A.sums, &A.sc->rev_used);
  grid.sync();
  p_rev_fill(A.rh, A.cols, A.posm, A.rev_start, A.rev_len, A.rev_cols, A.bq, A.rev_bq, A.halo_of, A.halo_only, A.sc, A.n, A.qmin);
  if (blockIdx.x == 0 && threadIdx.x == 0) A.sc->rev_valid = 1;   // every block read the guard before the first grid.sync
}

} // namespace dml
