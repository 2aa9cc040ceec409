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
  // overlap_moveback that follows (k_ov_init), bit 1 = the write-back of the overlap_moveback that came before (k_ov_apply),
  // bit 2 = the tail of the loop body: msd bookkeeping + F -> CG promotion (dana.F90:201-202,228-236) + calc_rho (521-549) and,
  // with the piston, maxz (776-794) after the rebuild decision
  int fuse; int *parent, *ovst, *comp_cnt, *ov_head; double *vel, *acel; const double *old_cg;
  double area, h_over_tau; int use_z1, piston;
  // defer = 1: a rebuild snapshots the positions (snap) and leaves the cell sort to whoever needs the cells first (sc->sort_pending):
  // the list rebuilt by the second test_update of a Brownian step is superseded by the next step's rebuild before anything reads it
  int defer; double4 *snap;
  // fuse bit 3 (k_test_update_coop<true>, Brownian integrator, Philox noise): cbrownian_hs + atom_pbc (dana.F90:798-846,1187-1250) of
  // slot s run in front of do_pbc of slot s in the same pass: the call site in front of the first test_update of a step
  const int *uid; double *ranv, *old_cg_w; Phys ph;
  unsigned char *dq;                                          // own displacement since the build, one byte per slot (dq_byte), or nullptr
};

// cgroup_sort (Cells.F90:267-302) by the whole grid: bin -> scan -> scatter -> order inside every cell; three grid-wide barriers.
// rebuild: the sort belongs to a list rebuild (halo flags, row heads of empty slots); snapshot: also update()'s pos_old = pos.
__device__ __forceinline__ void coop_sort_cells(cg::grid_group &grid, const TUArgs &A, const double4 *src, bool rebuild, bool snapshot) {
  const int gsz = gridDim.x * blockDim.x, gt = blockIdx.x * blockDim.x + threadIdx.x;
  DevScal *sc = A.sc;
  for (int s = gt; s < A.n; s += gsz) d_bin(src, A.cell_of, A.cell_cnt, A.rh, A.halo_of, sc, A.g, rebuild, s);
  grid.sync();
  coop_scan<true>(grid, A.cell_cnt, A.cell_start, A.nct, A.sums, A.cell_start + A.nct);
  grid.sync();
  for (int s = gt; s < A.n; s += gsz) d_scatter(A.posm, A.pos_old, src, A.cell_of, A.cell_start, A.cell_cur, A.sorted_raw, snapshot, s, A.dq);
  grid.sync();
  if (gt == 0 && rebuild) { sc->rows_asym = sc->halo_flag ? 1 : 0; sc->sort_pending = 0; }
  const int nsorted = A.cell_start[A.nct];
  for (int i = gt; i < nsorted; i += gsz)
    d_cell_rank(src, A.slot_b, A.cell_of, A.cell_start, A.cell_cur, A.sorted_raw, A.sorted_slot, A.sorted_posm, A.sorted_posf, A.sorted_cell, i);
}
// the cell sort a deferred rebuild left behind, for a consumer that needs the cells before the next test_update
__global__ void __launch_bounds__(TPB) k_sort_catchup(TUArgs A) {
  if (!((volatile const DevScal *)A.sc)->sort_pending) return;
  cg::grid_group grid = cg::this_grid();
  if (blockIdx.x == 0 && threadIdx.x == 0) A.sc->halo_flag = 0;
  grid.sync();
  coop_sort_cells(grid, A, A.snap, true, false);
}

// test_update (Neighbor.F90:668-713) in one launch
template <bool BI>
__global__ void __launch_bounds__(TPB) k_test_update_coop(TUArgs A) {
  cg::grid_group grid = cg::this_grid();
  const int gsz = gridDim.x * blockDim.x, gt = blockIdx.x * blockDim.x + threadIdx.x;
  DevScal *sc = A.sc;
  // phase 0: do_pbc + per-block top-2 squared displacement
  if (gt == 0) sc->halo_flag = 0;
  __shared__ unsigned int s_lay[LAY_MAX];
  __shared__ int s_red[34];
  for (int i = threadIdx.x; i < A.g.nlay; i += blockDim.x) s_lay[i] = 0u;
  __syncthreads();
  const int lay_old = sc->lay_cur;
  const int was_listed = sc->listed;                             // read before the first grid.sync: block 0 rewrites it right after
  const int pending_in = sc->sort_pending;
  const int par = sc->tu_par;
  const bool tail = (A.fuse & 4) != 0;
  const double z0 = sc->z0, zl = A.use_z1 ? sc->z1 : sc->zmax;
  const double pist_in = sc->pist_P, pist_c = 1.0 - 1.0 / pist_in;   // piston shift taken out of the recorded displacements (lay_note)
  double a1 = -1.0, a2 = -1.0;
  int c_rho = 0, c_dref = 0;
  if (gt == 0 && (A.fuse & 2)) {                                 // k_ov_apply's bookkeeping (dana.F90:939-941)
    if (sc->ch_later > sc->choques2) sc->choques2 = sc->ch_later;
    sc->overlap_passes += sc->any_active > 0 ? sc->any_active : 1;
  }
  if (gt == 0 && (A.fuse & 1)) { sc->again = 0; sc->n_roots = 0; sc->member_cursor = 0; sc->ch_later = 0; sc->any_active = 0; }
  BlockAcc iacc = {0, 0, 0.0, 0.0, 0.0f};
  const unsigned int istep = BI ? sc->istep : 0u;
  for (int s = gt; s < A.n; s += gsz) {
    if (BI) {                                                    // the Brownian step of this slot (same code as k_integrate<false>)
      const double4 p = ld_rec(&A.posm[s]);
      if (meta_of(p) & MF_REF) {
        const unsigned int id = (unsigned int)A.uid[s];
        double gs[6], v[3] = {0.0, 0.0, 0.0};
        const double a[3] = {0.0, 0.0, 0.0};
        Philox r;
        r.run(A.ph.seed, id, istep, RS_INTEG0, 0u); r.gauss4f(gs[0], gs[1], gs[2], gs[3]);
        RngSrc rs = {0, A.ph.seed, id, istep, nullptr, s, nullptr};
        integrate_one<false>(A.posm, A.vel, A.pos_old, A.old_cg_w, A.ranv, sc, A.g, A.ph, p, v, a, gs, rs, s, iacc);
      }
    }
    if (A.fuse & 2) d_ov_apply(A.posm, A.vel, A.acel, A.old_cg, A.ovst, s);
    double rel2, zn;
    double rd = d_pbc_disp(A.posm, A.pos_old, A.g, s, z0, pist_c, rel2, zn);
    if (A.fuse & 1) {
      const long long m = meta_of(ld_rec(&A.posm[s]));
      A.parent[s] = s; A.comp_cnt[s] = 0; A.ov_head[s] = -1;
      A.ovst[s] = ((m & MF_SKIP) ? OV_SKIP : 0) | ((int)(m & MF_TYPE) << OV_TSHIFT);
    }
    if (rd >= 0.0) lay_note(s_lay, A.g, zn, rel2);
    if (A.dq) A.dq[s] = dq_byte(A.g, rel2);
    top2_merge(a1, a2, rd, -1.0);
    if (tail) {                                                  // promotion loop + census of calc_rho (same pass as k_promote_rho)
      double4 p = ld_rec(&A.posm[s]);
      long long m = meta_of(p);
      if ((m & MF_REF) && (m & MF_TYPE) == 3) {
        c_dref++;
        m = (m & ~(MF_TYPE | MF_REF | MF_GCMC)) | 2;
        p.w = meta_as_double(m); st_rec(&A.posm[s], p);
      }
      if ((m & MF_TYPE) && p.z > z0 && p.z < zl) c_rho++;
    }
  }
  if (BI) block_flush(iacc, sc);                                 // counters and the largest move of the step (read by d_qtab behind the barrier)
  __syncthreads();
  for (int i = threadIdx.x; i < A.g.nlay; i += blockDim.x) if (s_lay[i]) atomicMax(&A.lay[(lay_old ^ 1) * LAY_MAX + i], s_lay[i]);
  block_top2(a1, a2);
  if (threadIdx.x == 0) { A.part[2 * blockIdx.x] = a1; A.part[2 * blockIdx.x + 1] = a2; }
  if (tail) {                                                    // one pair of atomics per block
    c_rho = block_sum_int(c_rho, s_red); c_dref = block_sum_int(c_dref, s_red);
    if (threadIdx.x == 0) { if (c_rho) atomicAdd(&sc->rho_cnt2[par], c_rho); if (c_dref) atomicAdd(&sc->dref_cnt2[par], c_dref); }
  }
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
  // fused tail: rho of this step (every thread computes it from the finished census), then the piston
  double lohi = 0.0;
  if (tail) {
    const int gct = *((volatile int *)&sc->rho_cnt2[par]);
    const double rho = gct / (A.area * (zl - z0));               // box(1)*box(2)*(z-z0), dana.F90:521-549
    if (A.piston) lohi = A.h_over_tau * ((sc->rho0 - rho) / rho);  // maxz, dana.F90:776-794
    if (gt == 0) {
      const int dr = *((volatile int *)&sc->dref_cnt2[par]);
      sc->msd_t = sc->msd_t / sc->nat_ref;                       // dana.F90:201-202 (before the promotion loop)
      sc->msd_max = fmax(sc->msd_max, sc->msd_t);
      sc->nat_ref -= dr;
      sc->rho = rho;
      sc->tu_par = par ^ 1; sc->rho_cnt2[par ^ 1] = 0; sc->dref_cnt2[par ^ 1] = 0;
    }
  }
  if (gt == 0) {
    sc->d1 = a1; sc->d2 = a2; sc->need_rebuild = need ? 1 : 0;
    sc->dsum_tu = need ? 0.0 : sqrt(a1) + sqrt(a2); sc->maxz_disp = 0.0; sc->maxz_fac = 0.0;
    sc->kappa_tu = need ? 0.0 : fabs(pist_c) * 1.01;
    if (!need) sc->lay_cur = lay_old ^ 1;
    if (need) { sc->nupd++; sc->listed = 1; sc->nlimbo = 0; sc->hole_lo = 0; sc->rev_valid = 0; sc->rows_pending = 1; sc->cols_used = sc->cols_tail0; sc->pist_P = 1.0; }
    if (tail && A.piston) {                                      // k_maxz's scalar part; the skip tables of this call are not read again inside dml_step
      sc->maxz_disp += fabs(lohi) * fmax(sc->zmax - z0, 0.0) * 1.01;
      sc->maxz_fac += fabs(lohi) * 1.01;
      sc->pist_P = (need ? 1.0 : sc->pist_P) * (1.0 - lohi);
      sc->zmax = sc->zmax - lohi * (sc->zmax - z0);
    }
  }
  if (blockIdx.x == 0) {                                         // z-layer tables (see k_top2_final)
    if (need) { for (int i = threadIdx.x; i < 2 * LAY_MAX; i += blockDim.x) A.lay[i] = 0u; }
    else { for (int i = threadIdx.x; i < A.g.nlay; i += blockDim.x) A.lay[lay_old * LAY_MAX + i] = 0u; }
    __threadfence();
    __syncthreads();
    if (!tail) d_qtab(A.lay, sc, A.g, A.rmax_f, A.rmax_o);       // skip tables for the consumers that follow (none after the tail of a step)
    else { __syncthreads(); if (threadIdx.x == 0) sc->step_disp_bits = 0u; }   // end of the step: the next integrator call records its own largest move
  }
  const bool defer = A.defer && !A.force_sort;
  const bool do_sort = !defer && (need || pending_in || A.force_sort);
  if (need && defer) {
    // update() without the cell sort: pos_old = pos (Neighbor.F90:620-624), igroup_clean (Groups.F90:1036-1053), snapshot for the deferred sort
    for (int s = gt; s < A.n; s += gsz) {
      double4 p = ld_rec(&A.posm[s]);
      const long long m = meta_of(p);
      if (m & MF_TYPE) { A.pos_old[3 * s] = p.x; A.pos_old[3 * s + 1] = p.y; A.pos_old[3 * s + 2] = p.z; }
      else if (m & MF_LIMBO) { p.w = meta_as_double(0); st_rec(&A.posm[s], p); }
      if (A.dq) A.dq[s] = 0;
      st_rec(&A.snap[s], p);
    }
    if (gt == 0) sc->sort_pending = 1;
  }
  if (tail && A.piston && !(do_sort)) {                          // maxz on the particles (after the snapshot; a sorting call does it at its end)
    for (int s = gt; s < A.n; s += gsz) {
      double4 p = ld_rec(&A.posm[s]);
      if ((meta_of(p) & MF_TYPE) && p.z > z0) { p.z = p.z - lohi * (p.z - z0); st_rec(&A.posm[s], p); }
    }
  }
  if (!do_sort) return;
  // phase 2: cells.  Source of the positions: the snapshot of a deferred rebuild that is being caught up with, else the records.
  const bool from_snap = !need && pending_in;
  coop_sort_cells(grid, A, from_snap ? A.snap : A.posm, need || pending_in, need);
  // the rows themselves (ngroup_cells, Neighbor.F90:465-548) are built from this snapshot by k_rows when a consumer first needs them
  if (tail && A.piston) {                                        // maxz on the particles, after everything that needed the positions of the rebuild
    grid.sync();
    for (int s = gt; s < A.n; s += gsz) {
      double4 p = ld_rec(&A.posm[s]);
      if ((meta_of(p) & MF_TYPE) && p.z > z0) { p.z = p.z - lohi * (p.z - z0); st_rec(&A.posm[s], p); }
    }
  }
}

struct OVArgs {
  double4 *posm; double *vel, *acel; const double *old_cg; const RowHead *rh; const int *cols; const unsigned char *bq; const unsigned int *lay; int *parent, *ovst, *comp_cnt, *comp_off,
      *members, *roots, *ov_head, *ov_next; const int *uid; OvRp rp_uovl; DevScal *sc; Geo g; Phys ph; unsigned int step; int n, guard_pass;
};

// overlap_moveback (dana.F90:849-943) in one launch (prob>=1)
__global__ void __launch_bounds__(TPB) k_overlap_coop(OVArgs A) {
  cg::grid_group grid = cg::this_grid();
  p_ov_init(A.posm, A.parent, A.ovst, A.comp_cnt, A.ov_head, A.sc, A.n);
  grid.sync();
  p_ov_detect<1>(A.posm, A.old_cg, A.rh, A.cols, A.bq, A.lay, A.parent, A.ovst, A.sc, A.g, A.n);
  grid.sync();
  p_ov_link(A.parent, A.ovst, A.ov_head, A.ov_next, A.roots, A.sc, A.n);
  grid.sync();
  p_ov_resolve<false>(A.posm, A.old_cg, A.rh, A.cols, A.bq, A.lay, A.ovst, A.roots, A.ov_head, A.ov_next, A.members, A.uid, A.rp_uovl,
               A.sc, A.g, A.ph, A.step, A.guard_pass, nullptr);
  grid.sync();
  p_ov_apply(A.posm, A.vel, A.acel, A.old_cg, A.ovst, A.sc, A.n);
}

// Transposed rows (k_rev_count -> scan -> k_rev_fill -> k_rev_done) in one launch.  The multi-launch form costs four guarded
// launches in front of every pair-force call (measured at 1 M particles: 31 us per step, all of it idle outside rebuild steps);
// here an idle call is one launch whose blocks return on the guard, which nobody rewrites before the last phase.
struct RevArgs {
  const RowHead *rh; const int *cols; const double4 *posm; int *rev_start, *rev_len, *rev_cnt, *rev_cols;
  const unsigned char *bq; unsigned char *rev_bq; const unsigned char *halo_of; int halo_only; int *sums; DevScal *sc; int n;
};
__global__ void __launch_bounds__(TPB) k_rev_coop(RevArgs A) {
  REV_GUARD(A.sc);
  cg::grid_group grid = cg::this_grid();
  p_rev_count(A.rh, A.cols, A.posm, A.rev_len, A.rev_cnt, A.halo_of, A.halo_only, A.sc, A.n);
  grid.sync();
  coop_scan<true>(grid, A.rev_cnt, A.rev_start, A.n, A.sums, &A.sc->rev_used);
  grid.sync();
  p_rev_fill(A.rh, A.cols, A.posm, A.rev_start, A.rev_len, A.rev_cols, A.bq, A.rev_bq, A.halo_of, A.halo_only, A.sc, A.n);
  if (blockIdx.x == 0 && threadIdx.x == 0) A.sc->rev_valid = 1;   // every block read the guard before the first grid.sync
}

} // namespace dml
