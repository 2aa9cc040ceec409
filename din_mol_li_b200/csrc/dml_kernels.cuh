// dml_kernels.cuh — the sm_100a kernels of the dana hot path.  One thread per particle slot unless noted.
// Reference lines cited per kernel are relative to the reference tree (src/...).
#pragma once
#include "dml_device.cuh"

namespace dml {

constexpr int TPB = 256;
constexpr unsigned int STEP_FROM_DEVICE = 0xffffffffu;   // kernels given this step word read DevScal::istep
// one integrator call = one tick of the Philox step word (dml_step enqueues it in front of the integrator)
__global__ void k_tick(DevScal *sc) { sc->istep = sc->istep + 1u; }
constexpr int ROW_W = 24;   // Poisson(5.8) neighbours in solution: P(n > 16) = 1.2e-4 left a dozen rows per 100 k particles on the serial ordered walk (a 28 us tail of the 70 us kernel, ncu: SMs active 57 % of the time); P(n > 24) = 3e-9

// Guard used by every kernel of the rebuild sequence: they are always launched (no host round trip) and return
// immediately unless test_update decided to rebuild (or the caller forces the cell sort, e.g. for gcmc).
#define REBUILD_GUARD(sc, force) if (!(((volatile const DevScal *)(sc))->need_rebuild | (force))) return

// ================================================================================================
// Exclusive scan of int32, single pass with decoupled look-back (1024 items per tile).
// state[tile] = epoch:30 | flag:2 | value:32; tiles are handed out by a ticket so a tile only ever waits for
// tiles whose blocks are already running.  The epoch makes re-zeroing of state[] unnecessary.
// ZERO_IN=true clears the input after reading it (cell histogram invariant: all zero between rebuilds).
// ================================================================================================
template <bool ZERO_IN>
__global__ void __launch_bounds__(TPB) k_scan(int *__restrict__ in, int *__restrict__ out, int n, unsigned long long *state,
                                              unsigned int *tickets, unsigned int epoch, int *total_out,
                                              const DevScal *sc, int force, int guard_mode) {
  if (guard_mode == 0) { REBUILD_GUARD(sc, force); }
  else if (guard_mode == 1) { if (!(((volatile const DevScal *)sc)->rows_asym && !((volatile const DevScal *)sc)->rev_valid)) return; }
  else if (guard_mode == 3) { if (!((volatile const DevScal *)sc)->rows_pending) return; }
  __shared__ int wsum[8];
  __shared__ int s_tile, s_prefix;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(&tickets[0], 1u);
  __syncthreads();
  const int tile = s_tile;
  int base = tile * 1024 + threadIdx.x * 4;
  int v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; if (ZERO_IN && base + i < n) in[base + i] = 0; }
  int t = v[0] + v[1] + v[2] + v[3];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) wsum[w] = x;
  __syncthreads();
  if (w == 0) {
    int q = lane < 8 ? wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, q, o); if (lane >= o) q += y; }
    if (lane < 8) wsum[lane] = q;
  }
  __syncthreads();
  const int T = wsum[7];
  if (w == 0) {
    // decoupled look-back by one warp: 32 predecessors per poll (a single polling thread made a tile wait for a chain of
    // dependent L2 round trips as long as its distance to the nearest finished predecessor)
    const unsigned long long ep = (unsigned long long)(epoch & 0x3fffffffu) << 34;
    volatile unsigned long long *vs = (volatile unsigned long long *)state;
    int prefix = 0;
    if (tile == 0) {
      if (lane == 0) vs[0] = ep | (2ull << 32) | (unsigned int)T;
    } else {
      if (lane == 0) { vs[tile] = ep | (1ull << 32) | (unsigned int)T; __threadfence(); }
      for (int j = tile - 1; j >= 0; j -= 32) {
        const int idx = j - lane;
        unsigned long long sv = 0ull;
        if (idx >= 0) { do { sv = vs[idx]; } while ((sv >> 34) != (ep >> 34) || ((sv >> 32) & 3ull) == 0); }
        const unsigned int incl = __ballot_sync(0xffffffffu, idx >= 0 && ((sv >> 32) & 3ull) == 2);
        const int first = incl ? __ffs(incl) - 1 : 32;          // nearest predecessor that already holds an inclusive prefix
        prefix += __reduce_add_sync(0xffffffffu, (idx >= 0 && lane <= first) ? (int)(unsigned int)sv : 0);
        if (incl) break;
      }
      if (lane == 0) vs[tile] = ep | (2ull << 32) | (unsigned int)(prefix + T);
    }
    if (lane == 0) {
      __threadfence();
      s_prefix = prefix;
      if (tile == (int)gridDim.x - 1 && total_out) *total_out = prefix + T;
    }
  }
  __syncthreads();
  int run = x - t + (w ? wsum[w - 1] : 0) + s_prefix;
#pragma unroll
  for (int i = 0; i < 4; ++i) { if (base + i < n) out[base + i] = run; run += v[i]; }
  if (tile == (int)gridDim.x - 1 && threadIdx.x == TPB - 1 && base + 4 >= n) { /* out[n] is written through total_out by the caller's choice */ }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int d = atomicAdd(&tickets[1], 1u);
    if (d == gridDim.x - 1) { tickets[0] = 0; tickets[1] = 0; __threadfence(); }
  }
}

// ================================================================================================
// K1  test_update (Neighbor.F90:668-713): do_pbc (Groups.F90:1440-1467) + the two largest squared displacements
//     (inq_dispmax, Neighbor.F90:635-666) in one streaming pass; k_top2_final takes the rebuild decision on the device.
// ================================================================================================
// rel2 (out): squared displacement with the piston's affine shift taken out (see lay_note); znow: the particle's height
__device__ __forceinline__ double d_pbc_disp(double4 *__restrict__ posm, double *__restrict__ pos_old, const Geo &g, int s,
                                             double z0, double pist_c, double &rel2, double &znow) {
  double4 p = ld_rec(&posm[s]);
  long long m = meta_of(p);
  rel2 = -1.0; znow = p.z;
  if (!(m & MF_TYPE)) return -1.0;
  double po[3] = {__ldcs(&pos_old[3 * s]), __ldcs(&pos_old[3 * s + 1]), __ldcs(&pos_old[3 * s + 2])};
  double q[3] = {p.x, p.y, p.z};
  bool ch = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) if (g.pbc[k]) {
    if (q[k] >= g.box[k]) { q[k] = q[k] - g.box[k]; po[k] = po[k] - g.box[k]; ch = true; }
    else if (q[k] < 0.0) { q[k] = q[k] + g.box[k]; po[k] = po[k] + g.box[k]; ch = true; }
  }
  if (ch) {
    p.x = q[0]; p.y = q[1]; p.z = q[2]; st_rec(&posm[s], p);
    pos_old[3 * s] = po[0]; pos_old[3 * s + 1] = po[1]; pos_old[3 * s + 2] = po[2];
  }
  double vx = q[0] - po[0], vy = q[1] - po[1], vz = q[2] - po[2];
  const double wz = vz - pist_c * fmax(q[2] - z0, 0.0);
  rel2 = (vx * vx + vy * vy) + wz * wz;
  return (vx * vx + vy * vy) + vz * vz;
}
// block-wide top-2 of per-thread (a1,a2); result valid in thread 0
__device__ __forceinline__ void block_top2(double &a1, double &a2) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double b1 = __shfl_xor_sync(0xffffffffu, a1, o), b2 = __shfl_xor_sync(0xffffffffu, a2, o);
    top2_merge(a1, a2, b1, b2);
  }
  __shared__ double s1[32], s2[32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) { s1[w] = a1; s2[w] = a2; }
  __syncthreads();
  if (w == 0) {
    a1 = lane < nw ? s1[lane] : -1.0; a2 = lane < nw ? s2[lane] : -1.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double b1 = __shfl_xor_sync(0xffffffffu, a1, o), b2 = __shfl_xor_sync(0xffffffffu, a2, o);
      top2_merge(a1, a2, b1, b2);
    }
  }
}
// z-layer table of the largest displacement since the rows were built (float bits, order preserving for x>=0):
// lets the pair-force kernel bound |move of i| + |move of j| locally instead of by the global top-2 (the piston shifts
// particles near the ceiling far more than those near the electrode).
constexpr int LAY_MAX = 1024;
__device__ __forceinline__ int layer_of(const Geo &g, double z) {
  // any monotone map of z onto slabs at least one list radius thick will do, as long as the table is filled and read with
  // the same one: the inverse cell height saves the fp64 division of the exact cell index
  int cz = (int)(z * g.inv_cell2) + 1;
  cz = cz < 0 ? 0 : (cz > g.nc[2] + 1 ? g.nc[2] + 1 : cz);
  return cz >> g.lay_shift;
}
// What the table records is NOT the displacement u = pos - pos_old itself but w = u - c(z), c(z) = (1 - 1/P) max(z - z0, 0) z^ being the
// shift the piston (maxz: z -= lohi (z - z0), P = product of (1 - lohi) since the rows were built) gave a particle now at height z.
// For a pair, u_i - u_j = (w_i - w_j) + (c(z_i) - c(z_j)) and |c(z_i) - c(z_j)| <= |1 - 1/P| |z_i - z_j| whatever c models, so
// |change of the pair separation| <= |w_i| + |w_j| + kappa |z_i - z_j|: an identity, rigorous for any P.  The piston moves a particle
// at the top of a 1 M box by ~1 A per step but a PAIR by ~0.01 A; with absolute displacements the skip bound of the upper layers
// was used up two steps after every rebuild (pair force 31 us after a rebuild, 56 us on average).
__device__ __forceinline__ void lay_note(unsigned int *s_lay, const Geo &g, double z, double rd) {
  if (rd < 0.0) return;
  float d = __double2float_ru(sqrt(rd)) * 1.000001f;
  atomicMax(&s_lay[layer_of(g, z)], (unsigned int)__float_as_int(d));
}
// The same bound tabulated per z-layer for overlap_moveback (table 1, rmax = rcut; table 0 holds the pair-force value).  The
// single block that finishes test_update refreshes it (d_top2_final), so the overlap kernels that follow read one byte where
// every thread would evaluate the formula; a stand-alone dml_overlap_moveback call refreshes it with k_qtab first. table 0 for the pair force (rmax = largest
// cut-off of the pair table), table 1 for overlap_moveback (rmax = rcut).  The bound of a layer uses the top of the layer where
// the per-particle formula used z, so it is never smaller.  The tables live behind the two displacement tables of lay[].
__device__ __forceinline__ int skip_qtab(const Geo &g, const unsigned int *__restrict__ lay, double z, int which) {
  const unsigned char *qt = reinterpret_cast<const unsigned char *>(lay + 2 * LAY_MAX) + which * LAY_MAX;
  return (int)__ldg(&qt[layer_of(g, z)]);
}
// Bound of the change of a pair separation since the rows were built: relative displacements recorded at the last test_update (mx, per
// layer, piston shift taken out: lay_note) or the global top-2 sum if smaller, the piston's relative shift before (kappa) and after
// (maxz_fac) that test_update over at most 2 list radii of height difference, and the integrator's move since (sdisp, each partner).
__device__ __forceinline__ double pair_shift_bound(const Geo &g, double mx, double dsum, double kappa, double maxz_fac, double sdisp) {
  const double R = 2.0 * sqrt(g.rc_list2);
  return fmin(2.0 * mx + kappa * R, dsum) + maxz_fac * R + 2.0 * sdisp;
}
// Per-particle refinement: a pair moved at most (own displacement) + (largest displacement of the partner's layers), where the
// layer formula charges the layer maximum twice.  The own displacement (relative, like the layer table's) is kept as one byte per
// slot in units of the build-distance bytes, rounded up (dq, written by every test_update, zero after a rebuild); the second set of
// tables holds the bound without the own share: qmax_i = min(layer entry, base entry + dq_i).  The layer maximum of a few ten
// thousand particles is ~4 sigma of the displacement distribution, the typical particle sits at ~1.6 sigma.
__device__ __forceinline__ unsigned char dq_byte(const Geo &g, double rel2) {
  if (rel2 < 0.0) return 0;
  return (unsigned char)fmin(255.0, ceil(sqrt(rel2) * 1.000001 * g.bq_scale) + 1.0);
}
__device__ __forceinline__ int qtab_base_entry(const unsigned int *lt, const Geo &g, int l, double thick, double maxz_fac, double z0, double zmax,
                                               double sdisp, double rmax, double kappa) {
  unsigned int mx = 0u;
#pragma unroll
  for (int d = -2; d <= 2; ++d) { int q = l + d; if (q >= 0 && q < g.nlay) mx = max(mx, __ldcg(&lt[q])); }
  double ztop = thick * (double)(l + 1);
  if (l == g.nlay - 1) ztop = fmax(ztop, zmax);
  const double since = maxz_fac * fmax(ztop + 2.0 * thick - z0, 0.0) + sdisp;
  if (since > g.cell[2]) return 255;
  const double R = 2.0 * sqrt(g.rc_list2);
  return (int)fmin(255.0, ceil((rmax * 1.000001 + (double)__int_as_float((int)mx) + kappa * R + maxz_fac * R + 2.0 * sdisp) * g.bq_scale) + 1.0);
}
__device__ __forceinline__ int skip_qmax(const Geo &g, const unsigned int *__restrict__ lay, double z, int which, const unsigned char *__restrict__ dq, int s) {
  const unsigned char *qt = reinterpret_cast<const unsigned char *>(lay + 2 * LAY_MAX);
  const int l = layer_of(g, z);
  const int cap = (int)__ldg(&qt[which * LAY_MAX + l]);
  if (!dq) return cap;
  return min(cap, min(255, (int)__ldg(&qt[(2 + which) * LAY_MAX + l]) + (int)__ldg(&dq[s])));
}
// One entry of the tables: the bound of layer l uses the top of the layer where the per-particle formula used z.
__device__ __forceinline__ int qtab_entry(const unsigned int *lt, const Geo &g, int l, double thick, double maxz_fac, double z0, double zmax,
                                          double dsum, double sdisp, double rmax, double kappa) {
  unsigned int mx = 0u;
#pragma unroll
  for (int d = -2; d <= 2; ++d) { int q = l + d; if (q >= 0 && q < g.nlay) mx = max(mx, __ldcg(&lt[q])); }
  double ztop = thick * (double)(l + 1);             // every particle of layer l sits below ...
  if (l == g.nlay - 1) ztop = fmax(ztop, zmax);      // ... except in the last one, which also takes what is above the box (up to the ceiling)
  const double since = maxz_fac * fmax(ztop + 2.0 * thick - z0, 0.0) + sdisp;     // absolute move of a particle of this layer since the tables were filled
  if (since > g.cell[2]) return 255;                 // particles may have changed layer: no skipping
  return (int)fmin(255.0, ceil((rmax * 1.000001 + pair_shift_bound(g, (double)__int_as_float((int)mx), dsum, kappa, maxz_fac, sdisp)) * g.bq_scale) + 1.0);
}
// Executed by one block.  sc is read around L1 (the block may hold a stale line of it from earlier in its kernel).
__device__ __forceinline__ void d_qtab(unsigned int *__restrict__ lay, const DevScal *sc, const Geo &g, double rmax_f, double rmax_o) {
  volatile const DevScal *v = sc;
  unsigned char *qt = reinterpret_cast<unsigned char *>(lay + 2 * LAY_MAX);
  const unsigned int *lt = lay + v->lay_cur * LAY_MAX;
  const double thick = g.cell[2] * (double)(1 << g.lay_shift);
  const double maxz_fac = v->maxz_fac, z0 = v->z0, zmax = v->zmax, dsum = v->dsum_tu;
  const double sdisp = (double)__int_as_float((int)v->step_disp_bits);
  const double kappa = v->kappa_tu;
  for (int l = threadIdx.x; l < g.nlay; l += blockDim.x) {
    qt[l] = (unsigned char)qtab_entry(lt, g, l, thick, maxz_fac, z0, zmax, dsum, sdisp, rmax_f, kappa);
    qt[LAY_MAX + l] = (unsigned char)qtab_entry(lt, g, l, thick, maxz_fac, z0, zmax, dsum, sdisp, rmax_o, kappa);
    qt[2 * LAY_MAX + l] = (unsigned char)qtab_base_entry(lt, g, l, thick, maxz_fac, z0, zmax, sdisp, rmax_f, kappa);
    qt[3 * LAY_MAX + l] = (unsigned char)qtab_base_entry(lt, g, l, thick, maxz_fac, z0, zmax, sdisp, rmax_o, kappa);
  }
}
__global__ void k_qtab(unsigned int *__restrict__ lay, const DevScal *__restrict__ sc, Geo g, double rmax_f, double rmax_o) { d_qtab(lay, sc, g, rmax_f, rmax_o); }
// Final step of test_update, run by ONE block: merges the per-block top-2 partials, takes the rebuild decision
// (Neighbor.F90:697-710) and rotates the z-layer tables.
__device__ __forceinline__ void d_top2_final(const double *part, int nb, DevScal *__restrict__ sc, unsigned int *__restrict__ lay, const Geo &g,
                                             double nb_dcut, double rmax_f, double rmax_o, int stride = 2) {
  const int nlay = g.nlay;
  double a1 = 1e-16, a2 = 1e-16;     // Neighbor.F90:643-644
  for (int i = threadIdx.x; i < nb; i += blockDim.x) top2_merge(a1, a2, __ldcg(&part[stride * i]), __ldcg(&part[stride * i + 1]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double b1 = __shfl_xor_sync(0xffffffffu, a1, o), b2 = __shfl_xor_sync(0xffffffffu, a2, o);
    top2_merge(a1, a2, b1, b2);
  }
  __shared__ double s1[32], s2[32];
  __shared__ int s_need_sh;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { s1[w] = a1; s2[w] = a2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) top2_merge(a1, a2, s1[i], s2[i]);
    sc->d1 = a1; sc->d2 = a2;
    int need = (!sc->listed) || (sqrt(a1) + sqrt(a2) > nb_dcut);     // Neighbor.F90:697-710
    sc->need_rebuild = need;
    sc->dsum_tu = need ? 0.0 : sqrt(a1) + sqrt(a2); sc->maxz_disp = 0.0; sc->maxz_fac = 0.0;
    sc->kappa_tu = need ? 0.0 : fabs(1.0 - 1.0 / sc->pist_P) * 1.01;     // the factor k_pbc_disp took the piston shift out with
    s_need_sh = need;
    if (need) { sc->nupd++; sc->listed = 1; sc->nlimbo = 0; sc->hole_lo = 0; sc->halo_flag = 0; sc->rev_valid = 0; sc->rows_pending = 1; sc->cols_used = sc->cols_tail0; sc->pist_P = 1.0; }
  }
  // z-layer tables: after a rebuild both are zero (displacements restart); otherwise the one just filled becomes current
  // and the previous one is cleared for the next test_update
  __syncthreads();
  const int cur = sc->lay_cur;
  if (s_need_sh) { for (int i = threadIdx.x; i < 2 * LAY_MAX; i += blockDim.x) lay[i] = 0u; }
  else {
    for (int i = threadIdx.x; i < nlay; i += blockDim.x) lay[cur * LAY_MAX + i] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) sc->lay_cur = cur ^ 1;
  }
  __threadfence();
  __syncthreads();
  d_qtab(lay, sc, g, rmax_f, rmax_o);                  // the skip tables the next consumers (overlap_moveback, pair force) will read
}
// finalize == 1: the last block to finish runs d_top2_final itself (single-GPU path, one launch less per test_update);
// finalize == 0: the partials are left in part[] for the caller;
// finalize == 2 (slab mode): the last block reduces the partials to this rank's two largest (own_out[0..1]), which are then merged
// across the ranks.
__global__ void __launch_bounds__(TPB) k_pbc_disp(double4 *__restrict__ posm, double *__restrict__ pos_old, double *part,
                                                  unsigned int *__restrict__ lay, DevScal *__restrict__ sc, Geo g, int n, int n_disp,
                                                  int finalize, double nb_dcut, double rmax_f, double rmax_o, double *__restrict__ own_out = nullptr,
                                                  unsigned char *__restrict__ dq = nullptr) {
  // persistent grid (a few blocks per SM, grid-stride): one flush of the per-block layer table per block
  __shared__ unsigned int s_lay[LAY_MAX];
  __shared__ int s_last;
  for (int i = threadIdx.x; i < g.nlay; i += blockDim.x) s_lay[i] = 0u;
  __syncthreads();
  const double z0p = sc->z0, pist_c = 1.0 - 1.0 / sc->pist_P;
  double a1 = -1.0, a2 = -1.0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    double rel2, zn;
    double rd = d_pbc_disp(posm, pos_old, g, s, z0p, pist_c, rel2, zn);
    if (rd >= 0.0) lay_note(s_lay, g, zn, rel2);
    if (dq) dq[s] = dq_byte(g, rel2);
    if (s < n_disp) top2_merge(a1, a2, rd, -1.0);        // slab mode: ghosts are measured by their owners
  }
  __syncthreads();
  unsigned int *dst = lay + (sc->lay_cur ^ 1) * LAY_MAX;
  for (int i = threadIdx.x; i < g.nlay; i += blockDim.x) if (s_lay[i]) atomicMax(&dst[i], s_lay[i]);
  block_top2(a1, a2);
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = a1; part[2 * blockIdx.x + 1] = a2; }
  if (!finalize) return;
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); s_last = (atomicAdd(&sc->ticket3, 1u) == gridDim.x - 1) ? 1 : 0; }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) sc->ticket3 = 0u;
  __threadfence();
  if (finalize == 2) {
    double b1 = -1.0, b2 = -1.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) top2_merge(b1, b2, __ldcg(&part[2 * i]), __ldcg(&part[2 * i + 1]));
    block_top2(b1, b2);
    if (threadIdx.x == 0) { own_out[0] = b1; own_out[1] = b2; }
    return;
  }
  d_top2_final(part, (int)gridDim.x, sc, lay, g, nb_dcut, rmax_f, rmax_o);
}
__global__ void k_top2_final(const double *part, int nb, DevScal *__restrict__ sc, unsigned int *__restrict__ lay, Geo g,
                             double nb_dcut, double rmax_f, double rmax_o) {
  d_top2_final(part, nb, sc, lay, g, nb_dcut, rmax_f, rmax_o);
}

// ================================================================================================
// K2  cell binning + counting sort (cgroup_sort, Cells.F90:267-302).  In-cell order must be DESCENDING b-slot
//     (head insertion of ascending slots): scatter with an atomic cursor, then order every cell segment.
//     k_scatter also performs update()'s pos_old=pos (Neighbor.F90:620-624) and igroup_clean (Groups.F90:1036-1053)
//     because both happen exactly when the list is rebuilt.  Invariant: cell_cnt and cell_cur are all zero
//     between rebuilds (the scan clears cell_cnt, k_cell_order clears cell_cur).
// ================================================================================================
__device__ __forceinline__ void d_bin(const double4 *__restrict__ posm, int *__restrict__ cell_of, int *__restrict__ cell_cnt,
                                      RowHead *__restrict__ rh, unsigned char *__restrict__ halo_of,
                                      DevScal *__restrict__ sc, const Geo &g, bool rebuild, int s) {
  double4 p = ld_rec_nc(&posm[s]);
  if (meta_of(p) & MF_TYPE) {
    int cx, cy, cz;
    if (cell_index(g, p.x, p.y, p.z, cx, cy, cz)) {
      int lin = cell_lin(g, cx, cy, cz);
      cell_of[s] = lin;
      atomicAdd(&cell_cnt[lin], 1);
      if (rebuild) {
        const bool halo = cx == 0 || cy == 0 || cz == 0 || cx == g.nc[0] + 1 || cy == g.nc[1] + 1 || cz == g.nc[2] + 1;
        halo_of[s] = halo ? 1 : 0;
        if (halo) sc->halo_flag = 1;
      }
    } else { cell_of[s] = -1; atomicCAS(&sc->err, 0, DML_E_OUT_OF_TESS); }
  } else {
    cell_of[s] = -1;
    if (rebuild) { rh_store_plain(&rh[s], s * ROW_W, 0, 0, 255); halo_of[s] = 0; }
  }
}
__global__ void __launch_bounds__(TPB) k_bin(const double4 *__restrict__ posm, int *__restrict__ cell_of, int *__restrict__ cell_cnt,
                                             RowHead *__restrict__ rh, unsigned char *__restrict__ halo_of,
                                             DevScal *__restrict__ sc, Geo g, int n, int force) {
  REBUILD_GUARD(sc, force);                               // small persistent grid: a launch that has nothing to do stays cheap
  const bool rebuild = ((volatile const DevScal *)sc)->need_rebuild != 0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
    d_bin(posm, cell_of, cell_cnt, rh, halo_of, sc, g, rebuild, s);
}
// src: where the binned positions came from (the records, or the snapshot of a deferred rebuild: dml_coop.cuh)
__device__ __forceinline__ void d_scatter(double4 *__restrict__ posm, double *__restrict__ pos_old, const double4 *src, const int *__restrict__ cell_of,
                                          const int *__restrict__ cell_start, int *__restrict__ cell_cur, int *__restrict__ sorted_slot,
                                          bool snapshot, int s, unsigned char *__restrict__ dq = nullptr) {
  double4 p = ld_rec(&src[s]);
  long long m = meta_of(p);
  if (snapshot && dq) dq[s] = 0;                          // pos_old = pos: nobody has moved since this build
  if (m & MF_TYPE) {
    int lin = cell_of[s];
    if (lin >= 0) { int idx = cell_start[lin] + atomicAdd(&cell_cur[lin], 1); sorted_slot[idx] = s; }
    if (snapshot) { pos_old[3 * s] = p.x; pos_old[3 * s + 1] = p.y; pos_old[3 * s + 2] = p.z; }
  } else if (snapshot && (m & MF_LIMBO)) { p.w = meta_as_double(0); st_rec(&posm[s], p); }
}
__global__ void __launch_bounds__(TPB) k_scatter(double4 *__restrict__ posm, double *__restrict__ pos_old, const int *__restrict__ cell_of,
                                                 const int *__restrict__ cell_start, int *__restrict__ cell_cur,
                                                 int *__restrict__ sorted_slot, const DevScal *__restrict__ sc, int n, int force,
                                                 unsigned char *__restrict__ dq = nullptr) {
  REBUILD_GUARD(sc, force);
  const bool snapshot = ((volatile const DevScal *)sc)->need_rebuild != 0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
    d_scatter(posm, pos_old, posm, cell_of, cell_start, cell_cur, sorted_slot, snapshot, s, dq);
}
// One thread per binned particle (raw = the scatter's output, in-cell order arbitrary): its place in the cell is the number of
// cell mates with a larger b index (chains are visited in descending b index, Cells.F90:267-302), found with independent loads;
// the thread then writes the slot, the record, its single-precision copy and the cell id at that place.  (One thread per CELL with
// an insertion sort was a serial chain per dense cell — 28 atoms per cell over a metal slab — and a thread per empty cell in boxes
// with more cells than particles: 1.15 M cells for 100 k particles at skin 2.)
__device__ __forceinline__ void d_cell_rank(const double4 *__restrict__ posm, const int *__restrict__ slot_b, const int *__restrict__ cell_of,
                                            const int *__restrict__ cell_start, int *__restrict__ cell_cur, const int *__restrict__ raw,
                                            int *__restrict__ sorted_slot, double4 *__restrict__ sorted_posm,
                                            float4 *__restrict__ sorted_posf, int *__restrict__ sorted_cell, int i) {
  const int sl = raw[i];
  const int c = cell_of[sl];
  const int b = cell_start[c], e = cell_start[c + 1];
  const int key = slot_b[sl];
  int rank = 0;
  for (int j = b; j < e; ++j) rank += slot_b[raw[j]] > key ? 1 : 0;
  if (rank == 0) cell_cur[c] = 0;                        // the scatter cursor of the cell is zero again between rebuilds
  const int pos = b + rank;
  const double4 p = ld_rec(&posm[sl]);
  sorted_slot[pos] = sl;
  st_rec(&sorted_posm[pos], p);
  // single-precision copy for the candidate scan of k_rows; w carries the slot
  sorted_posf[pos] = make_float4((float)p.x, (float)p.y, (float)p.z, __int_as_float(sl));
  sorted_cell[pos] = c;                                  // cell of the sorted particle (k_rows: no gather through cell_of[slot])
}
__global__ void k_cell_order(const double4 *__restrict__ posm, const int *__restrict__ slot_b, const int *__restrict__ cell_of,
                             const int *__restrict__ cell_start, int *__restrict__ cell_cur, const int *__restrict__ raw,
                             int *__restrict__ sorted_slot, double4 *__restrict__ sorted_posm,
                             float4 *__restrict__ sorted_posf, int *__restrict__ sorted_cell, DevScal *__restrict__ sc, int ncell, int force) {
  REBUILD_GUARD(sc, force);
  if (blockIdx.x == 0 && threadIdx.x == 0 && ((volatile const DevScal *)sc)->need_rebuild) sc->rows_asym = sc->halo_flag ? 1 : 0;
  const int nsorted = cell_start[ncell];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nsorted; i += gridDim.x * blockDim.x)
    d_cell_rank(posm, slot_b, cell_of, cell_start, cell_cur, raw, sorted_slot, sorted_posm, sorted_posf, sorted_cell, i);
}

// ================================================================================================
// K3  Verlet rows over linked cells      (ngroup_cells, Neighbor.F90:465-548; cell_pbc Cells.F90:378-404;
//     vdistance Groups.F90:995-1016).  One thread per cell-sorted ref particle (see d_rows): rows come out in the
//     reference's order (stencil order x chain order).  FILL=false counts, FILL=true writes.
// ================================================================================================
// Row storage: slot s owns ROW_W entries at cols[s*ROW_W ..) ("region A").  A row that fits (length + gcmc slack <= ROW_W;
// n̄n is about 6 in solution) lives there, so the build needs no count pass and no scan; longer rows (next to dense metal)
// get a segment of the tail region [cols_tail0, cols_cap) by an atomic bump and are written by a second walk of the same
// thread.  Where the row lives, its length and its first 16 build distances are the slot's RowHead (dml_device.cuh); the
// build distances of entries 16.. of a long row are in bq[], indexed like cols[].

// One thread per cell-sorted ref particle, two dense loops and no per-thread arrays (a queue or a table of cell ranges in local
// memory costs more L2/DRAM traffic than the whole list):
//  walk   one flat loop over the candidates in stencil order x chain order (which IS the reference's row order), single
//         precision; a candidate below the upper edge of the fp32 error band is parked in the row's own storage (its sorted
//         index, 4 bytes).  Lanes of a warp differ only in their total candidate count (~38 +- 6), and the next cell's range
//         is fetched when a lane runs out of candidates.
//  settle the parked candidates get the reference's exact fp64 test (vdistance, Groups.F90:995-1016; strict <,
//         Neighbor.F90:515), the build-distance byte, and are compacted in place as slot ids.
// About one candidate in seven is a hit: with the fp64 work inside the walk some lane of the warp would take the heavy path
// in nearly every iteration with ~5 lanes active.
// Ordered walk + settle of one particle into cols[dst .. dst+lim): the reference's enumeration (27 stencil cells in map order x
// chain order).  Returns the number of parked candidates; cnt / hb describe the finished row when npark <= lim.
struct RowOut { int npark, cnt; Near5 n5; };
__device__ __forceinline__ RowOut rows_ordered_into(const double4 *__restrict__ sorted_posm, const float4 *__restrict__ sorted_posf,
                                                    const int *__restrict__ sorted_slot, const int *__restrict__ cell_start,
                                                    int *__restrict__ cols, unsigned char *__restrict__ bq, const Geo &g,
                                                    const double4 &p, int t, int cx, int cy, int cz, int dst, int lim, bool pad) {
  const float hbx = g.pbc[0] ? 0.5f * (float)g.box[0] : 3.0e38f, hby = g.pbc[1] ? 0.5f * (float)g.box[1] : 3.0e38f;
  const float rc2hi = (float)g.rc_list2 + g.band2;
  const float pxf = (float)p.x, pyf = (float)p.y, pzf = (float)p.z;
  // ---- walk ----
  int npark = 0, nab = -1, u = 0, e = 0;
  for (;;) {
    while (u == e) {                                  // next stencil cell (empty ones are passed over)
      if (++nab == 27) break;
      int dx, dy, dz; map_of_lane(nab, dx, dy, dz);
      int nx = dx + cx - 1, ny = dy + cy - 1, nz = dz + cz - 1;     // cell_pbc wraps every axis, z included (Cells.F90:387-391)
      nx = (nx < 0 ? nx + g.nc[0] : (nx >= g.nc[0] ? nx - g.nc[0] : nx)) + 1;   // |offset| <= 2 < nc, one conditional add == mod
      ny = (ny < 0 ? ny + g.nc[1] : (ny >= g.nc[1] ? ny - g.nc[1] : ny)) + 1;
      nz = (nz < 0 ? nz + g.nc[2] : (nz >= g.nc[2] ? nz - g.nc[2] : nz)) + 1;
      const int nl = cell_lin(g, nx, ny, nz);
      u = __ldg(&cell_start[nl]); e = __ldg(&cell_start[nl + 1]);
    }
    if (nab == 27) break;
    if (u != t) {
      const float4 q = __ldg(&sorted_posf[u]);
      float vx = q.x - pxf, vy = q.y - pyf, vz = q.z - pzf;
      if (vx > hbx) vx -= 2.0f * hbx; else if (vx < -hbx) vx += 2.0f * hbx;
      if (vy > hby) vy -= 2.0f * hby; else if (vy < -hby) vy += 2.0f * hby;
      if (vx * vx + vy * vy + vz * vz <= rc2hi) { if (npark < lim) cols[dst + npark] = u; ++npark; }
    }
    ++u;
  }
  // ---- settle: the reference's exact fp64 test (vdistance, Groups.F90:995-1016; strict <, Neighbor.F90:515) ----
  RowOut o; o.npark = npark; o.cnt = 0; near5_init(o.n5);
  const int nset = min(npark, lim);
  for (int i = 0; i < nset; ++i) {
    const int uq = cols[dst + i];
    const double4 qd = ld_rec_nc(&sorted_posm[uq]);
    const double rd = dist2_idnint(g, qd.x, qd.y, qd.z, p.x, p.y, p.z);   // vdistance(vd,aj,ai)
    if (rd < g.rc_list2) {
      const int cnt = o.cnt;
      cols[dst + cnt] = sorted_slot[uq];               // cnt <= i: compaction in place
      // lower bound of the build-time distance in 1/255 of the list radius (feeds the gather skip of the consumers)
      const unsigned int qb = (unsigned int)min(255, (int)(sqrt(rd) * g.bq_scale * 0.999999999));
      bq[dst + cnt] = (unsigned char)qb;
      near5_add(o.n5, (qb << 16) | (unsigned int)min(cnt, 65535));
      o.cnt = cnt + 1;
    }
  }
  if (pad && npark > 0 && npark <= ROW_W) {            // leave no partly written sector behind: pad to the next 8 entries
    for (int i = npark; i < ((npark + 7) & ~7); ++i) cols[dst + i] = -1;
  }
  return o;
}

// inverse of the stencil map: nab_of(dx,dy,dz) = position of the offset in Cells.F90:28-36 (5 bits per entry, one word per dz)
constexpr int MAP27[27][3] = {
  {0,0,0},{1,0,0},{1,1,0},{0,1,0},{-1,1,0},{1,0,-1},{1,1,-1},{0,1,-1},{-1,1,-1},
  {1,0,1},{1,1,1},{0,1,1},{-1,1,1},{0,0,1},{-1,0,0},{-1,-1,0},{0,-1,0},{1,-1,0},
  {-1,0,1},{-1,-1,1},{0,-1,1},{1,-1,1},{-1,0,-1},{-1,-1,-1},{0,-1,-1},{1,-1,-1},{0,0,-1}};
__host__ __device__ constexpr unsigned long long inv_map_word(int dzp) {
  unsigned long long w = 0ull;
  for (int n = 0; n < 27; ++n)
    if (MAP27[n][2] + 1 == dzp) w |= (unsigned long long)n << (5 * ((MAP27[n][1] + 1) * 3 + (MAP27[n][0] + 1)));
  return w;
}
constexpr unsigned long long INV_MAP_W0 = inv_map_word(0), INV_MAP_W1 = inv_map_word(1), INV_MAP_W2 = inv_map_word(2);
__device__ __forceinline__ int nab_of(int dx, int dy, int dz) {
  constexpr unsigned long long w0 = INV_MAP_W0, w1 = INV_MAP_W1, w2 = INV_MAP_W2;
  const unsigned long long w = dz < 0 ? w0 : (dz == 0 ? w1 : w2);
  return (int)((w >> (5 * ((dy + 1) * 3 + dx + 1))) & 31ull);
}

// Long row (next to dense metal: an ion over a two-layer bcc electrode sees ~90 CG atoms) built by a whole warp: lane l < 27 owns
// stencil cell l of the map (Cells.F90:28-36), counts its fp32 candidates, an exclusive prefix over the lanes gives every cell's
// place in the row (stencil order x chain order), a second walk parks the candidates there, and the fp64 test + compaction run
// 32 entries at a time with ballots.  One thread doing this alone (what the thread-per-particle pass would do) takes ~10 k
// dependent instructions per row: measured 326 us instead of 163 us of test_update per step on the 1 M box with the CG slab.
__device__ __forceinline__ void rows_long_warp(const double4 *__restrict__ sorted_posm, const float4 *__restrict__ sorted_posf,
                                               const int *__restrict__ sorted_slot, const int *__restrict__ sorted_cell,
                                               const int *__restrict__ cell_start, RowHead *__restrict__ rh,
                                               int *__restrict__ cols, unsigned char *__restrict__ bq,
                                               DevScal *__restrict__ sc, const Geo &g, int slack, int t) {
  const unsigned int full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const double4 p = ld_rec_nc(&sorted_posm[t]);
  const int s = sorted_slot[t];
  const int lin = sorted_cell[t];
  const int cx = lin % g.hd[0], r = lin / g.hd[0], cy = r % g.hd[1], cz = r / g.hd[1];
  const float hbx = g.pbc[0] ? 0.5f * (float)g.box[0] : 3.0e38f, hby = g.pbc[1] ? 0.5f * (float)g.box[1] : 3.0e38f;
  const float rc2hi = (float)g.rc_list2 + g.band2;
  const float pxf = (float)p.x, pyf = (float)p.y, pzf = (float)p.z;
  int b = 0, e = 0;
  if (lane < 27) {
    int dx, dy, dz; map_of_lane(lane, dx, dy, dz);
    int nx = dx + cx - 1, ny = dy + cy - 1, nz = dz + cz - 1;     // cell_pbc wraps every axis, z included (Cells.F90:387-391)
    nx = (nx < 0 ? nx + g.nc[0] : (nx >= g.nc[0] ? nx - g.nc[0] : nx)) + 1;
    ny = (ny < 0 ? ny + g.nc[1] : (ny >= g.nc[1] ? ny - g.nc[1] : ny)) + 1;
    nz = (nz < 0 ? nz + g.nc[2] : (nz >= g.nc[2] ? nz - g.nc[2] : nz)) + 1;
    const int nl = cell_lin(g, nx, ny, nz);
    b = __ldg(&cell_start[nl]); e = __ldg(&cell_start[nl + 1]);
  }
  int tb = 0, npark = 0;
  for (int pass = 0; pass < 2; ++pass) {
    int c = 0;
    int off = 0;
    if (pass == 1) {
      // place of this lane's cell in the row
      int incl = npark;                                     // npark holds the lane's own count of pass 0 here
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(full, incl, o); if (lane >= o) incl += y; }
      off = incl - npark;
      npark = __shfl_sync(full, incl, 31);
      const int need = npark + slack;
      if (lane == 0) tb = atomicAdd(&sc->cols_used, need);
      tb = __shfl_sync(full, tb, 0);
      if (tb + need > sc->cols_cap) {
        if (lane == 0) { atomicCAS(&sc->err, 0, DML_E_COLS_OVERFLOW); rh_store_plain(&rh[s], s * ROW_W, 0, ROW_W, 255); }
        return;
      }
    }
    for (int u = b; u < e; ++u) {
      if (u == t) continue;
      const float4 q = __ldg(&sorted_posf[u]);
      float vx = q.x - pxf, vy = q.y - pyf, vz = q.z - pzf;
      if (vx > hbx) vx -= 2.0f * hbx; else if (vx < -hbx) vx += 2.0f * hbx;
      if (vy > hby) vy -= 2.0f * hby; else if (vy < -hby) vy += 2.0f * hby;
      if (vx * vx + vy * vy + vz * vz <= rc2hi) { if (pass == 1) cols[tb + off + c] = u; ++c; }
    }
    if (pass == 0) npark = c;
  }
  __syncwarp();
  // settle: the reference's exact test (vdistance, Groups.F90:995-1016; strict <, Neighbor.F90:515), ordered compaction in place
  int cnt = 0;
  for (int i0 = 0; i0 < npark; i0 += 32) {
    const int i = i0 + lane;
    bool hit = false; int slot = -1; unsigned int qb = 0u;
    if (i < npark) {
      const int uq = cols[tb + i];
      const double4 qd = ld_rec_nc(&sorted_posm[uq]);
      const double rd = dist2_idnint(g, qd.x, qd.y, qd.z, p.x, p.y, p.z);
      if (rd < g.rc_list2) { hit = true; slot = sorted_slot[uq]; qb = (unsigned int)min(255, (int)(sqrt(rd) * g.bq_scale * 0.999999999)); }
    }
    const unsigned int hm = __ballot_sync(full, hit);
    __syncwarp();                                           // every lane has read its parked entry before the row is overwritten
    if (hit) {
      const int pos = cnt + __popc(hm & ((1u << lane) - 1u));
      cols[tb + pos] = slot;
      bq[tb + pos] = (unsigned char)qb;
    }
    cnt += __popc(hm);
  }
  __syncwarp();
  // near list: the five smallest (distance, position) keys of the finished row, one warp-wide minimum per round
  Near5 n5; near5_init(n5);
  unsigned int last = 0u;
  for (int r5 = 0; r5 < 5; ++r5) {
    unsigned int best = 0xffffffffu;
    for (int i = lane; i < cnt; i += 32) {
      const unsigned int key = ((unsigned int)bq[tb + i] << 16) | (unsigned int)min(i, 65535);
      if ((r5 == 0 || key > last) && key < best) best = key;
    }
    best = __reduce_min_sync(full, best);
    if (best == 0xffffffffu) break;
    n5.k[r5] = best; last = best;
  }
  if (lane == 0) rh_store_near<0>(&rh[s], n5, tb, cnt, npark + slack, [&](int pos) { return cols[tb + pos]; });
}

// ================================================================================================
// K3b  Boxes with fewer than 4 cells on EVERY axis: the reference does not tessellate and builds the rows by the O(N^2) loop
//      ngroup_verlet (Neighbor.F90:358-424): for every ref atom, candidates in ascending index of hs%b, vdistance, entry when
//      rd <= (rcut+skin)^2 (inclusive here, strict in the cell walk).  Such a box holds a few hundred atoms at most.
//      k_verlet_prepare is update() for this path (pos_old = pos, limbo slots freed, Neighbor.F90:608-633) plus the b-index -> slot
//      table; k_rows_verlet builds one row per warp from the positions of the rebuild (pos_old), 32 candidates at a time.
// ================================================================================================
__global__ void __launch_bounds__(TPB) k_verlet_prepare(double4 *__restrict__ posm, double *__restrict__ pos_old, const int *__restrict__ slot_b,
                                                        int *__restrict__ b2slot, RowHead *__restrict__ rh, unsigned char *__restrict__ halo_of,
                                                        DevScal *__restrict__ sc, int n) {
  REBUILD_GUARD(sc, 0);
  if (blockIdx.x == 0 && threadIdx.x == 0) { sc->rows_asym = 0; sc->halo_flag = 0; }
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    double4 p = ld_rec(&posm[s]);
    const long long m = meta_of(p);
    halo_of[s] = 0;
    if (m & MF_TYPE) {
      b2slot[slot_b[s]] = s;
      pos_old[3 * s] = p.x; pos_old[3 * s + 1] = p.y; pos_old[3 * s + 2] = p.z;
    } else {
      if (m & MF_LIMBO) { p.w = meta_as_double(0); st_rec(&posm[s], p); }
      rh_store_plain(&rh[s], s * ROW_W, 0, 0, 255);
    }
  }
}
__global__ void __launch_bounds__(TPB) k_rows_verlet(const double4 *__restrict__ posm, const double *__restrict__ pos_old,
                                                     const int *__restrict__ b2slot, RowHead *__restrict__ rh,
                                                     int *__restrict__ cols, unsigned char *__restrict__ bq, DevScal *__restrict__ sc, Geo g, int n, int slack) {
  if (!((volatile const DevScal *)sc)->rows_pending) return;
  const unsigned int full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int b_amax = sc->b_amax;
  for (int s = wid; s < n; s += nw) {
    const long long m1 = meta_of(ld_rec_nc(&posm[s]));
    if (!(m1 & MF_TYPE)) continue;
    if (!(m1 & MF_REF)) { if (lane == 0) rh_store_plain(&rh[s], s * ROW_W, 0, 0, 255); continue; }   // rows exist only for ref atoms
    const double px = pos_old[3 * s], py = pos_old[3 * s + 1], pz = pos_old[3 * s + 2];
    int dst = s * ROW_W, lim = ROW_W, total = 0, cnt = 0;
    for (int pass = 0; pass < 2; ++pass) {
      cnt = 0;
      for (int j0 = 0; j0 < b_amax; j0 += 32) {
        const int j = j0 + lane;
        const int sj = j < b_amax ? b2slot[j] : -1;
        bool hit = false; double rd = 0.0;
        if (sj >= 0 && sj != s) {
          rd = dist2_idnint(g, px, py, pz, pos_old[3 * sj], pos_old[3 * sj + 1], pos_old[3 * sj + 2]);   // vdistance(vd,ai,aj)
          hit = !(rd > g.rc_list2);                          // "if (rd>rcut) cycle", Neighbor.F90:399
        }
        const unsigned int hm = __ballot_sync(full, hit);
        if (pass == 1 && hit) {
          const int pos = cnt + __popc(hm & ((1u << lane) - 1u));
          cols[dst + pos] = sj;
          bq[dst + pos] = (unsigned char)min(255, (int)(sqrt(rd) * g.bq_scale * 0.999999999));
        }
        cnt += __popc(hm);
      }
      if (pass == 0) {
        total = cnt;
        if (total + slack > ROW_W) {                        // long row: a segment of the tail region
          const int need = total + slack;
          int tb = 0;
          if (lane == 0) tb = atomicAdd(&sc->cols_used, need);
          tb = __shfl_sync(full, tb, 0);
          if (tb + need > sc->cols_cap) { if (lane == 0) atomicCAS(&sc->err, 0, DML_E_COLS_OVERFLOW); total = -1; break; }
          dst = tb; lim = need;
        }
      }
    }
    if (total < 0) { if (lane == 0) rh_store_plain(&rh[s], s * ROW_W, 0, ROW_W, 255); continue; }
    if (lane == 0) rh_store_plain(&rh[s], dst, total, lim, 0);   // O(N^2) rows of a tiny box: no near list, every entry is looked at
  }
  __syncthreads();                                      // the last block to finish marks the rows as materialised
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&sc->ticket2, 1u) == gridDim.x - 1) { sc->ticket2 = 0; sc->rows_pending = 0; }
  }
}

// ---- asynchronous bulk copies global -> shared (cp.async.bulk, completion on an mbarrier; SASS: UBLKCP / SYNCS) ----------------
__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 16-byte aligned source / destination, size a multiple of 16 bytes
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, unsigned int bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity) {
  unsigned int ok = 0;
  while (!ok)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ================================================================================================
// k_rows — ngroup_cells (Neighbor.F90:465-548) from the cell-sorted snapshot of the last rebuild, one thread per sorted ref particle,
// one block per RB consecutive sorted particles.
//  staging  The RB particles of a block sit in consecutive cells c0..c1 of the (halo-inclusive) linear cell order, so for each of
//           the nine (dy,dz) stencil rows the cells c0-1+shift .. c1+1+shift hold every candidate of every particle of the block
//           that does not come through a periodic wrap, and they are ONE contiguous range of the cell-sorted arrays.  The nine
//           ranges of the single-precision copy (16 B per candidate) are brought into shared memory by nine bulk copies
//           (cp.async.bulk + mbarrier: the one tile-movement kernel of the path), capped at WCAP candidates each; candidates
//           outside the staged ranges (periodic wrap cells, ranges beyond the cap next to dense metal) are read from global memory
//           through the same generic pointer.
//  walk     per stencil row the x-neighbours cx-1..cx+1 are one contiguous run of candidates (the wrap cell of a particle at the
//           periodic x edge is a second, short run): nine runs of ~4 candidates instead of 27 cells of ~1.4, each with a warp-uniform
//           trip count (redux max).  Candidates are screened in fp32 (one load, three subtractions, three fused multiply-adds, one
//           compare) with a rigorous error band; only the ones inside the band (~0.5 %) get the reference's fp64 test (vdistance,
//           Groups.F90:995-1016; strict <, Neighbor.F90:515).  The minimum image is a per-run constant folded into the particle's
//           coordinate: all candidates of a cell sit in the same periodic image relative to the particle, and it is the operation
//           the branchy form applies whenever the pair can be inside the list radius (cells are at least one list radius wide, at
//           least three per axis).
//  order    a hit knows its stencil position nab (Cells.F90:28-36) and its rank among the hits of that cell (cells are scanned in
//           chain order = ascending sorted index); 27 byte counters per thread in shared memory, one prefix pass in map order, and
//           every hit lands at prefix[nab] + rank: the reference's row order (stencil order x chain order) without sorting.
//  rows     written as the slot's own ROW_W entries + one 32-byte RowHead + the qmin byte.  Rows that do not fit (next to dense
//           metal) are rebuilt by whole warps (rows_long_warp); particles binned in an x/y halo cell and boxes with fewer than 3
//           cells on an axis (where the reference visits a cell twice) take the ordered walk of one thread (rows_ordered_into).
//  history  round 1 gathered every candidate through L1/L2 per thread and ordered the hits by key insertion: 57 us at 110 k
//           particles (24.8 M warp instructions, IPC 0.5); a 27-cell walk in map order over the staged windows needed no ordering
//           but 21.3 M instructions (69 us: 27 x (two dependent cell_start loads + a loop of ~4.6 trips with 45 % of the lanes idle)).
// ================================================================================================
constexpr int RB = 256;          // sorted particles per block
constexpr int WCAP = 384;        // staged candidates per stencil row
constexpr size_t ROWS_OFF_HIT = (size_t)9 * WCAP * sizeof(float4);
constexpr size_t ROWS_OFF_META = ROWS_OFF_HIT + (size_t)(ROW_W + 1) * RB * sizeof(int);   // + the spare row of the branch-free park
constexpr size_t ROWS_OFF_QB = ROWS_OFF_META + (size_t)ROW_W * RB * sizeof(unsigned short);
constexpr size_t ROWS_OFF_CNT = ROWS_OFF_QB + (size_t)ROW_W * RB;
constexpr size_t ROWS_OFF_LONG = ROWS_OFF_CNT + (size_t)28 * RB;
constexpr size_t ROWS_SMEM = ROWS_OFF_LONG + (size_t)RB * sizeof(int);

__global__ void __launch_bounds__(RB, 2) k_rows(const double4 *__restrict__ sorted_posm, const float4 *__restrict__ sorted_posf,
                                               const int *__restrict__ sorted_slot,
                                               const int *__restrict__ sorted_cell, const int *__restrict__ cell_start,
                                               RowHead *__restrict__ rh,
                                               int *__restrict__ cols, unsigned char *__restrict__ bq,
                                               DevScal *__restrict__ sc, const __grid_constant__ Geo g, int ncell, int slack) {
  if (!((volatile const DevScal *)sc)->rows_pending) return;
  extern __shared__ __align__(128) unsigned char rows_smem[];
  float4 *s_win = reinterpret_cast<float4 *>(rows_smem);                                              // [9][WCAP]
  int (*s_hit)[RB] = reinterpret_cast<int (*)[RB]>(rows_smem + ROWS_OFF_HIT);                         // slot of parked hit i
  unsigned short (*s_meta)[RB] = reinterpret_cast<unsigned short (*)[RB]>(rows_smem + ROWS_OFF_META);  // nab | rank in its cell << 5
  unsigned char (*s_qb)[RB] = reinterpret_cast<unsigned char (*)[RB]>(rows_smem + ROWS_OFF_QB);        // build-distance byte
  unsigned char (*s_cnt)[RB] = reinterpret_cast<unsigned char (*)[RB]>(rows_smem + ROWS_OFF_CNT);      // [27] hits per stencil cell, then their prefix
  int *s_long = reinterpret_cast<int *>(rows_smem + ROWS_OFF_LONG);
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int s_wlo[9], s_whi[9];
  __shared__ int s_nlong;
  __shared__ unsigned char s_nab[27];                   // stencil position of (run, x-neighbour): inverse of the map, Cells.F90:28-36
  const unsigned int full = 0xffffffffu;
  const int tid = threadIdx.x;
  if (tid < 27) s_nab[tid] = (unsigned char)nab_of(tid % 3 - 1, (tid / 3) % 3 - 1, tid / 9 - 1);   // index = run * 3 + x-neighbour
  const int nsorted = __ldg(&cell_start[ncell]);        // number of binned particles
  if (tid == 0) mbar_init(&s_bar, 1);
  unsigned int phase = 0u;
  // persistent grid (two blocks per SM), batch-stride: a launch that finds nothing pending costs one wave of empty blocks (a grid
  // of one block per batch paid 12 us per idle launch at 1 M particles for 13 waves of 106 KB shared-memory allocations)
  for (int t0 = blockIdx.x * RB; t0 < nsorted; t0 += gridDim.x * RB, phase ^= 1u) {
    __syncthreads();                                    // previous batch done with the windows, the hit tables and s_nlong
    const int tl = min(t0 + RB, nsorted) - 1;
    if (tid < 9) {
      int lo = 0, hi = 0;
      if (g.rows_fast) {
        const int c0 = __ldg(&sorted_cell[t0]), c1 = __ldg(&sorted_cell[tl]);
        const int sh = (tid % 3 - 1) * g.hd[0] + (tid / 3 - 1) * g.hd[0] * g.hd[1];
        const int ca = min(max(c0 - 1 + sh, 0), ncell), cb = min(max(c1 + 2 + sh, 0), ncell);
        if (ca < cb) { lo = __ldg(&cell_start[ca]); hi = min(__ldg(&cell_start[cb]), lo + WCAP); }
      }
      s_wlo[tid] = lo; s_whi[tid] = hi;
    }
    if (tid == 0) s_nlong = 0;
    {                                                    // hit counters of the block: 28 * RB bytes
      unsigned int *z = reinterpret_cast<unsigned int *>(rows_smem + ROWS_OFF_CNT);
#pragma unroll
      for (int i = 0; i < 7; ++i) z[i * RB + tid] = 0u;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int bytes = 0;
#pragma unroll
      for (int k = 0; k < 9; ++k) bytes += (unsigned int)(s_whi[k] - s_wlo[k]) * (unsigned int)sizeof(float4);
      mbar_expect_tx(&s_bar, bytes);
#pragma unroll
      for (int k = 0; k < 9; ++k)
        if (s_whi[k] > s_wlo[k]) bulk_g2s(s_win + k * WCAP, sorted_posf + s_wlo[k], (unsigned int)(s_whi[k] - s_wlo[k]) * (unsigned int)sizeof(float4), &s_bar);
    }
    // ---- own particle (while the copies are in flight) ----
    const int t = t0 + tid;
    int mode = 0;                                       // 0 nothing to build, 1 staged walk, 2 ordered walk of one thread
    double4 p = make_double4(0, 0, 0, 0);
    int s = 0, cx = 1, cy = 1, cz = 1;
    if (t < nsorted) {
      p = ld_rec_nc(&sorted_posm[t]);
      s = __ldg(&sorted_slot[t]);
      if (!(meta_of(p) & MF_REF)) rh_store_plain(&rh[s], s * ROW_W, 0, 0, 255);   // rows exist only for ref atoms
      else {
        const int lin = __ldg(&sorted_cell[t]);
        cx = lin % g.hd[0]; const int r = lin / g.hd[0]; cy = r % g.hd[1]; cz = r / g.hd[1];
        mode = (g.rows_fast && cx >= 1 && cx <= g.nc[0] && cy >= 1 && cy <= g.nc[1]) ? 1 : 2;
      }
    }
    const float pxf = (float)p.x, pyf = (float)p.y, pzf = (float)p.z;
    const float bxf = g.pbc[0] ? (float)g.box[0] : 0.0f, byf = g.pbc[1] ? (float)g.box[1] : 0.0f;
    // wrapped neighbour rows (cell_pbc wraps every axis, z included, Cells.F90:387-391) as offsets into cell_start; the image
    // shift of a wrapped y row is folded into the particle's y
    int ly0, ly1, ly2, lz0, lz1, lz2;
    float py0 = pyf, py2 = pyf;
    {
      const int ecz = (cz - 1 < 0 ? cz - 1 + g.nc[2] : (cz - 1 >= g.nc[2] ? cz - 1 - g.nc[2] : cz - 1)) + 1;   // a centre in the z halo is wrapped like any other cell
      int a = cy - 1, c = cy + 1;
      if (a < 1) { a += g.nc[1]; py0 = pyf + byf; }        // candidates of the wrapped row sit one box length up: vy = q.y - (py + by)
      if (c > g.nc[1]) { c -= g.nc[1]; py2 = pyf - byf; }
      ly0 = a * g.hd[0]; ly1 = cy * g.hd[0]; ly2 = c * g.hd[0];
      int d = ecz - 1, f = ecz + 1;
      if (d < 1) d += g.nc[2];
      if (f > g.nc[2]) f -= g.nc[2];
      const int hz = g.hd[0] * g.hd[1];
      lz0 = d * hz; lz1 = ecz * hz; lz2 = f * hz;
    }
    const int xa = max(cx - 1, 1), xb = min(cx + 1, g.nc[0]);
    const int xw = mode == 1 ? (cx == 1 ? g.nc[0] : (cx == g.nc[0] ? 1 : 0)) : 0;     // periodic wrap cell of a particle at the x edge
    const float pxw = cx == 1 ? pxf + bxf : pxf - bxf;
    const int dxw = cx == 1 ? -1 : 1;
    const float rc2hi = (float)g.rc_list2 + g.band2;
    const float rc2lo = __double2float_rd(g.rc_list2) - g.band2;     // below this the fp32 distance is inside the list radius for sure
    const float bqs = __double2float_rd(g.bq_scale * 0.999999);
    int npark = 0;
    // One candidate inside the loops: the fp32 screen and nothing else.  A candidate below the upper edge of the error band is
    // parked as one word (sorted index | run << 24 | x-neighbour << 28 | "inside the list radius for sure" << 30); everything a hit
    // needs beyond that (exact test, stencil position, rank, build-distance byte) is done afterwards with all lanes busy: inlined
    // here it ran in nearly every iteration for the two or three lanes that had a hit (20 M warp instructions, 4 of them useful).
#define ROWS_CAND(OK, U, Q, PX, PY, DXV, SEG) do {                                                                           \
      const float vx_ = (Q).x - (PX), vy_ = (Q).y - (PY), vz_ = (Q).z - pzf;                                                 \
      const float d2_ = __fmaf_rn(vx_, vx_, __fmaf_rn(vy_, vy_, vz_ * vz_));                                                 \
      const bool h_ = (OK) && d2_ <= rc2hi && ((SEG) != 4 || (U) != t);   /* the particle itself sits in the centre run */    \
      /* branch-free: a candidate that is not parked writes the spare row ROW_W (the branchy form spent a quarter of the     \
         walk's instructions on BSSY / BSYNC / BRA) */                                                                        \
      s_hit[h_ ? min(npark, ROW_W) : ROW_W][tid] = (U) | ((SEG) << 24) | ((DXV) << 28) | (d2_ < rc2lo ? (1 << 30) : 0);      \
      npark += h_ ? 1 : 0;                                                                                                   \
    } while (0)
    // the bounds of the nine runs are requested together, before the wait for the staged windows (36 independent loads in flight
    // instead of nine dependent batches: with 5 A cells at skin 2 the cell table is 4.6 MB and every batch is an L2 round trip)
    int ru0[9], rb1[9], rb2[9], rn[9];
#pragma unroll
    for (int seg = 0; seg < 9; ++seg) {
      const int dy = seg % 3 - 1, dz = seg / 3 - 1;
      const int row = (dy < 0 ? ly0 : (dy == 0 ? ly1 : ly2)) + (dz < 0 ? lz0 : (dz == 0 ? lz1 : lz2));
      ru0[seg] = rb1[seg] = rb2[seg] = rn[seg] = 0;
      if (mode == 1) {
        ru0[seg] = __ldg(&cell_start[row + xa]); rb1[seg] = __ldg(&cell_start[row + cx]); rb2[seg] = __ldg(&cell_start[row + cx + 1]);
        rn[seg] = __ldg(&cell_start[row + xb + 1]);
      }
    }
    mbar_wait(&s_bar, phase);
    if (__any_sync(full, mode == 1)) {
      const bool anyw = __any_sync(full, xw != 0);
#pragma unroll
      for (int seg = 0; seg < 9; ++seg) {
        const int dy = seg % 3 - 1, dz = seg / 3 - 1;
        const int row = (dy < 0 ? ly0 : (dy == 0 ? ly1 : ly2)) + (dz < 0 ? lz0 : (dz == 0 ? lz1 : lz2));
        const float pys = dy < 0 ? py0 : (dy == 0 ? pyf : py2);
        const int u0 = ru0[seg], b1 = rb1[seg], b2 = rb2[seg], n = rn[seg] - ru0[seg];
        int uw = 0, nw = 0;
        if (xw) { uw = __ldg(&cell_start[row + xw]); nw = __ldg(&cell_start[row + xw + 1]) - uw; }
        const int wlo = s_wlo[seg], whi = s_whi[seg];
        {
          const float4 *cp = n <= 0 ? s_win : ((u0 >= wlo && u0 + n <= whi) ? (s_win + seg * WCAP + (u0 - wlo)) : (sorted_posf + u0));
          const int nmax = __reduce_max_sync(full, n), last = max(n - 1, 0);
          for (int i = 0; i < nmax; i += 4) {              // four independent candidates per trip; lanes beyond their run re-read its last one
            float4 q[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) q[e] = cp[min(i + e, last)];
#pragma unroll
            for (int e = 0; e < 4; ++e) { const int u = u0 + i + e; ROWS_CAND(i + e < n, u, q[e], pxf, pys, (u >= b1 ? 1 : 0) + (u >= b2 ? 1 : 0), seg); }
          }
        }
        if (anyw) {                                       // wrap cell of the particles at the periodic x edge
          const float4 *cp = nw <= 0 ? s_win : ((uw >= wlo && uw + nw <= whi) ? (s_win + seg * WCAP + (uw - wlo)) : (sorted_posf + uw));
          const int nmax = __reduce_max_sync(full, nw), last = max(nw - 1, 0);
          for (int i = 0; i < nmax; ++i) {
            const float4 q = cp[min(i, last)];
            ROWS_CAND(i < nw, uw + i, q, pxw, pys, dxw + 1, seg);
          }
        }
      }
    }
#undef ROWS_CAND
    // ---- settle the parked candidates: exact test inside the band, stencil position, rank among the hits of its cell ----
    if (mode == 1 && npark + slack > ROW_W) mode = 3;      // cannot fit the slot's own storage whatever the exact tests say: long row
    {
      const int np = mode == 1 ? npark : 0;
      const int npmax = __reduce_max_sync(full, np);
      int nh = 0;
      for (int i = 0; i < npmax; ++i) {
        if (i < np) {
          const int wd = s_hit[i][tid];
          const int u = wd & 0xffffff, seg = (wd >> 24) & 15, dxv = (wd >> 28) & 3;
          const int wlo = s_wlo[seg];
          const float4 q = (u >= wlo && u < s_whi[seg]) ? s_win[seg * WCAP + (u - wlo)] : __ldg(&sorted_posf[u]);
          bool hit = (wd >> 30) & 1;
          if (!hit) {                                      // the reference's test (vdistance, Groups.F90:995-1016; strict <, Neighbor.F90:515)
            const double4 qd = ld_rec_nc(&sorted_posm[u]);
            hit = dist2_idnint(g, qd.x, qd.y, qd.z, p.x, p.y, p.z) < g.rc_list2;
          }
          if (hit) {
            const int dz3 = seg / 3, dy3 = seg - 3 * dz3;
            // the run's image (wrap cell: one box length in x; wrapped y row: one box length in y) for the build-distance byte
            const float pxs = (dxv - 1 == dxw && xw) ? pxw : pxf;
            const float pys = dy3 == 0 ? py0 : (dy3 == 1 ? pyf : py2);
            const float vx = q.x - pxs, vy = q.y - pys, vz = q.z - pzf;
            const float d2 = __fmaf_rn(vx, vx, __fmaf_rn(vy, vy, vz * vz));
            const int nab = s_nab[seg * 3 + dxv];
            const int c = s_cnt[nab][tid];
            s_cnt[nab][tid] = (unsigned char)(c + 1);
            s_hit[nh][tid] = __float_as_int(q.w);           // w carries the slot (nh <= i: compaction in place)
            s_meta[nh][tid] = (unsigned short)(nab | (c << 5));
            // lower bound of the build-time distance in 1/255 of the list radius (feeds the gather skip of the consumers)
            s_qb[nh][tid] = (unsigned char)min(255, (int)__fmul_rd(__fsqrt_rd(fmaxf(d2 - g.band2, 0.0f)), bqs));
            ++nh;
          }
        }
      }
      npark = mode == 1 ? nh : npark;
    }
    if (mode == 1) {
      if (npark + slack <= ROW_W) {
        // prefix of the per-cell hit counts in map order, then every hit goes to prefix[nab] + rank
        int run = 0;
#pragma unroll
        for (int nab = 0; nab < 27; ++nab) { const int c = s_cnt[nab][tid]; s_cnt[nab][tid] = (unsigned char)run; run += c; }
        const int dst = s * ROW_W;
        Near5 n5; near5_init(n5);
        for (int i = 0; i < npark; ++i) {
          const int mt = s_meta[i][tid];
          const int pos = (int)s_cnt[mt & 31][tid] + (mt >> 5);
          const unsigned int qb = s_qb[i][tid];
          cols[dst + pos] = s_hit[i][tid];
          bq[dst + pos] = (unsigned char)qb;
          near5_add(n5, (qb << 16) | ((unsigned int)pos << 8) | (unsigned int)i);   // position orders the near list, i finds the slot again
        }
        for (int i = npark; i < ((npark + 7) & ~7); ++i) cols[dst + i] = -1;   // leave no partly written sector behind
        rh_store_near<8>(&rh[s], n5, dst, npark, ROW_W, [&](int k16) { return s_hit[k16 & 255][tid]; });  // one full-sector store per row
      } else s_long[atomicAdd(&s_nlong, 1)] = t;             // long row (next to dense metal): built by a warp below
    } else if (mode == 3) s_long[atomicAdd(&s_nlong, 1)] = t;
    else if (mode == 2) {
      const RowOut o = rows_ordered_into(sorted_posm, sorted_posf, sorted_slot, cell_start, cols, bq, g, p, t, cx, cy, cz, s * ROW_W, ROW_W, true);
      if (o.npark + slack <= ROW_W) rh_store_near<0>(&rh[s], o.n5, s * ROW_W, o.cnt, ROW_W, [&](int pos) { return cols[s * ROW_W + pos]; });
      else s_long[atomicAdd(&s_nlong, 1)] = t;
    }
    __syncthreads();
    const int nlong = s_nlong;
    for (int i = tid >> 5; i < nlong; i += RB / 32)
      rows_long_warp(sorted_posm, sorted_posf, sorted_slot, sorted_cell, cell_start, rh, cols, bq, sc, g, slack, s_long[i]);
  }
  __syncthreads();                                      // the last block to finish marks the rows as materialised
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&sc->ticket2, 1u) == gridDim.x - 1) { sc->ticket2 = 0; sc->rows_pending = 0; }
  }
}
__global__ void k_sum_rowlen(const RowHead *__restrict__ rh, int n, long long *__restrict__ out) {
  long long a = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a += rh[i].len;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0 && a) atomicAdd((unsigned long long *)out, (unsigned long long)a);
}

// ================================================================================================
// K5  pair force                (fuerza, dana.F90:1055-1139).  Gather formulation of the reference's
//     scatter loop: F_i = sum over row(i) of w_ij f_ij with w=2 for j in ref (the reference visits a
//     ref-ref pair from both rows and applies Newton's third law each time), w=1 for CG.
//     STRICT=true adds the terms in the reference's global visiting order (ascending creation rank of
//     the row owner), which makes force/epot bit-identical to the reference.
// ================================================================================================
__device__ __forceinline__ bool pair_terms(const Geo &g, const Phys &ph, const double4 &p1, int k, const double4 &p2, int m,
                                           double f[3], double &u) {
  double vd[3] = {p1.x - p2.x, p1.y - p2.y, p1.z - p2.z};
#pragma unroll
  for (int l = 0; l < 2; ++l) {                       // branchy minimum image, x and y only (dana.F90:1098-1106)
    if (vd[l] > g.half_box[l]) vd[l] = vd[l] - g.box[l];
    else if (vd[l] < -g.half_box[l]) vd[l] = vd[l] + g.box[l];
  }
  if (k == 2 && m == 2) return false;
  double dr = (vd[0] * vd[0] + vd[1] * vd[1]) + vd[2] * vd[2];
  int km = (k - 1) * 3 + (m - 1);
  if (dr > ph.r0sq[km]) return false;
  dr = sqrt(dr);
  double b = ph.r0p6[km];
  double c = ph.eps[km] * 12.0 * b;
  c = c / pow7(dr);
  b = b / pow6(dr);
  double aux = c * (b - 1.0);
#pragma unroll
  for (int l = 0; l < 3; ++l) f[l] = aux * vd[l] / dr;
  aux = ph.eps[km] * b * (b - 2.0);
  aux = aux + ph.eps[km];
  u = aux * .5;
  return true;
}

// Transposed rows: rev(i) = { j : i in row(j) }.  Needed when rows can be asymmetric (a particle in a halo cell is
// never found as a candidate, Cells.F90:248 + cell_pbc wrap; incremental gcmc appends use <= instead of <).
#define REV_GUARD(sc) if (!(((volatile const DevScal *)(sc))->rows_asym && !((volatile const DevScal *)(sc))->rev_valid)) return
__device__ __forceinline__ void p_rev_count(const RowHead *__restrict__ rh, const int *__restrict__ cols,
                            const double4 *__restrict__ posm, int *__restrict__ rev_len, int *__restrict__ rev_cnt,
                            const unsigned char *__restrict__ halo_of, int halo_only, const DevScal *__restrict__ sc, int n) {
  const bool light = halo_only && sc->rows_asym == 1;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    rev_len[s] = 0;                                     // becomes the fill cursor of k_rev_fill
    if (light && !halo_of[s]) continue;                 // light mode: only rows of halo-cell particles are transposed
    if (!(meta_of(ld_rec_nc(&posm[s])) & MF_REF)) continue;
    const RowMeta m = rh_meta(&rh[s]);
    for (int jj = 0; jj < m.len; ++jj) atomicAdd(&rev_cnt[cols[m.start + jj]], 1);   // rev_cnt is all zero between builds (the scan clears it)
  }
}
__global__ void k_rev_count(const RowHead *__restrict__ rh, const int *__restrict__ cols,
                            const double4 *__restrict__ posm, int *__restrict__ rev_len, int *__restrict__ rev_cnt,
                            const unsigned char *__restrict__ halo_of, int halo_only, const DevScal *__restrict__ sc, int n) {
  REV_GUARD(sc);
  p_rev_count(rh, cols, posm, rev_len, rev_cnt, halo_of, halo_only, sc, n);
}
__device__ __forceinline__ void p_rev_fill(const RowHead *__restrict__ rh, const int *__restrict__ cols,
                           const double4 *__restrict__ posm, const int *__restrict__ rev_start, int *__restrict__ rev_len,
                           int *__restrict__ rev_cols, const unsigned char *__restrict__ bq, unsigned char *__restrict__ rev_bq,
                           const unsigned char *__restrict__ halo_of, int halo_only, const DevScal *__restrict__ sc, int n) {
  const bool light = halo_only && sc->rows_asym == 1;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    if (light && !halo_of[s]) continue;
    if (!(meta_of(ld_rec_nc(&posm[s])) & MF_REF)) continue;
    const RowMeta m = rh_meta(&rh[s]);
    // the scan cleared rev_len; it is rebuilt here as the fill cursor and ends as the row length
    for (int jj = 0; jj < m.len; ++jj) {
      int j = cols[m.start + jj];
      int w = rev_start[j] + atomicAdd(&rev_len[j], 1);
      rev_cols[w] = s; rev_bq[w] = bq[m.start + jj];
    }
  }
}
__global__ void k_rev_fill(const RowHead *__restrict__ rh, const int *__restrict__ cols,
                           const double4 *__restrict__ posm, const int *__restrict__ rev_start, int *__restrict__ rev_len,
                           int *__restrict__ rev_cols, const unsigned char *__restrict__ bq, unsigned char *__restrict__ rev_bq,
                           const unsigned char *__restrict__ halo_of, int halo_only, const DevScal *__restrict__ sc, int n) {
  REV_GUARD(sc);
  p_rev_fill(rh, cols, posm, rev_start, rev_len, rev_cols, bq, rev_bq, halo_of, halo_only, sc, n);
}
__global__ void k_rev_done(DevScal *sc) { if (sc->rows_asym && !sc->rev_valid) sc->rev_valid = 1; }
// Reverse-visit candidates of atom s: with symmetric rows they are the ref entries of its own row, otherwise rev(s).
template <bool STRICT>
__global__ void __launch_bounds__(TPB) k_fuerza(const double4 *__restrict__ posm, const RowHead *__restrict__ rh,
                                                const int *__restrict__ cols,
                                                const int *__restrict__ rev_start, const int *__restrict__ rev_len,
                                                const int *__restrict__ rev_cols, const DevScal *__restrict__ sc,
                                                const int *__restrict__ uid, double4 *__restrict__ fe,
                                                Geo g, Phys ph, int n, unsigned char *__restrict__ fnz, unsigned char *__restrict__ kb) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double4 p1 = ld_rec_nc(&posm[s]);
  long long m1 = meta_of(p1);
  kb[s] = (m1 & MF_REF) ? (unsigned char)(m1 & MF_TYPE) : (unsigned char)0;   // what ermak_b needs of the record (k_ermak_b_flat)
  if (!(m1 & MF_REF)) return;
  int k = (int)(m1 & MF_TYPE);
  const int asym = sc->rows_asym;
  const RowMeta rm = rh_meta(&rh[s]);
  int b = rm.start, len = rm.len;
  const int *rv = asym ? rev_cols + rev_start[s] : cols + b;
  int rvlen = asym ? rev_len[s] : len;
  double fx = 0.0, fy = 0.0, fz = 0.0, ep = 0.0;
  if (!STRICT) {
    for (int jj = 0; jj < len; ++jj) {
      int j = cols[b + jj];
      double4 p2 = ld_rec_nc(&posm[j]);
      long long m2 = meta_of(p2);
      int m = (int)(m2 & MF_TYPE);
      if (m == 0) continue;                           // limbo / removed (dana.F90:1090-1092)
      double f[3], u;
      if (!pair_terms(g, ph, p1, k, p2, m, f, u)) continue;
      fx = fx + f[0]; fy = fy + f[1]; fz = fz + f[2]; ep = ep + u;
      if (!asym && (m2 & MF_ANYREF)) { fx = fx + f[0]; fy = fy + f[1]; fz = fz + f[2]; ep = ep + u; }
    }
    if (asym) {
      for (int jj = 0; jj < rvlen; ++jj) {
        int j = rv[jj];
        double4 p2 = ld_rec_nc(&posm[j]);
        long long m2 = meta_of(p2);
        int m = (int)(m2 & MF_TYPE);
        if (m == 0 || !(m2 & MF_ANYREF)) continue;
        double f[3], u;
        if (!pair_terms(g, ph, p1, k, p2, m, f, u)) continue;
        fx = fx + f[0]; fy = fy + f[1]; fz = fz + f[2]; ep = ep + u;
      }
    }
  } else {
    // Visiting order of the reference for atom i: reverse visits by row owners j with rank_j < rank_i (ascending
    // rank; the term -f_ji equals +f_ij bit for bit), then i's own row in row order, then reverse visits with
    // rank_j > rank_i.  Only entries inside the cut-off contribute, so they are collected first and then ordered.
    constexpr int KMAX = 12;
    int myuid = uid[s];
    int cu[KMAX]; double cf[KMAX][4]; int nrev = 0; bool overflow = false;
    for (int jj = 0; jj < rvlen; ++jj) {
      int j = rv[jj];
      double4 p2 = ld_rec_nc(&posm[j]);
      long long m2 = meta_of(p2);
      int m = (int)(m2 & MF_TYPE);
      if (m == 0 || !(m2 & MF_ANYREF)) continue;
      double f[3], u;
      if (!pair_terms(g, ph, p1, k, p2, m, f, u)) continue;
      if (nrev == KMAX) { overflow = true; break; }
      cu[nrev] = uid[j]; cf[nrev][0] = f[0]; cf[nrev][1] = f[1]; cf[nrev][2] = f[2]; cf[nrev][3] = u; ++nrev;
    }
    for (int phase = 0; phase < 3; ++phase) {
      if (phase == 1) {
        for (int jj = 0; jj < len; ++jj) {
          int j = cols[b + jj];
          double4 p2 = ld_rec_nc(&posm[j]);
          int m = (int)(meta_of(p2) & MF_TYPE);
          if (m == 0) continue;
          double f[3], u;
          if (!pair_terms(g, ph, p1, k, p2, m, f, u)) continue;
          fx = fx + f[0]; fy = fy + f[1]; fz = fz + f[2]; ep = ep + u;
        }
      } else if (!overflow) {
        int last = phase == 0 ? -1 : myuid;
        for (;;) {
          int best = 0x7fffffff, bi = -1;
          for (int i = 0; i < nrev; ++i) {
            int uj = cu[i];
            if (uj <= last || uj >= best) continue;
            if (phase == 0 && uj >= myuid) continue;
            bi = i; best = uj;
          }
          if (bi < 0) break;
          last = best;
          fx = fx + cf[bi][0]; fy = fy + cf[bi][1]; fz = fz + cf[bi][2]; ep = ep + cf[bi][3];
        }
      } else {
        // dense neighbourhood: same ordering by repeated selection over the candidate list
        int last = phase == 0 ? -1 : myuid;
        for (;;) {
          int best = 0x7fffffff, bj = -1;
          for (int jj = 0; jj < rvlen; ++jj) {
            int j = rv[jj];
            int uj = uid[j];
            if (uj <= last || uj >= best) continue;
            if (phase == 0 && uj >= myuid) continue;
            bj = j; best = uj;
          }
          if (bj < 0) break;
          last = best;
          double4 p2 = ld_rec_nc(&posm[bj]);
          long long m2 = meta_of(p2);
          int m = (int)(m2 & MF_TYPE);
          if (m == 0 || !(m2 & MF_ANYREF)) continue;
          double f[3], u;
          if (!pair_terms(g, ph, p1, k, p2, m, f, u)) continue;
          fx = fx + f[0]; fy = fy + f[1]; fz = fz + f[2]; ep = ep + u;
        }
      }
    }
  }
  st_rec(&fe[s], make_double4(fx, fy, fz, ep));        // force(3) + epot in one 256-bit store
  fnz[s] = (fx != 0.0 || fy != 0.0 || fz != 0.0 || ep != 0.0) ? 1 : 0;   // see k_fuerza_sub
}

// Production pair force (row order instead of the reference's global visiting order: same terms, 1e-12 budget of the north star).
// the rarely taken heavy part (pair inside the cut-off) lives out of line so the streaming path stays lean in registers
__device__ __noinline__ double4 lj_terms(double vx, double vy, double vz, double dr2, double eps, double r0p6) {
  double dr = sqrt(dr2);
  double b = r0p6;
  double c = eps * 12.0 * b;
  c = c / pow7(dr);
  b = b / pow6(dr);
  double aux = c * (b - 1.0);
  double4 r;
  r.x = aux * vx / dr; r.y = aux * vy / dr; r.z = aux * vz / dr;
  aux = eps * b * (b - 2.0);
  aux = aux + eps;
  r.w = aux * .5;
  return r;
}
// Pair term of the production kernel once the partner's record is here: cheap cut-off tests first, heavy math out of line.
// pass 0 = own row, pass 1 = transposed row (reverse visits come from row owners only).
struct FAcc { double fx, fy, fz, ep; };
__device__ __forceinline__ void fuerza_pair(const Geo &g, const Phys &ph, const double4 &p1, int k3, const double4 &p2,
                                            int pass, int asym, bool i_halo, FAcc &a) {
  double vx = p1.x - p2.x, vy = p1.y - p2.y, vz = p1.z - p2.z;
  if (vx > g.half_box[0]) vx = vx - g.box[0]; else if (vx < -g.half_box[0]) vx = vx + g.box[0];   // dana.F90:1098-1106
  if (vy > g.half_box[1]) vy = vy - g.box[1]; else if (vy < -g.half_box[1]) vy = vy + g.box[1];
  const double dr2 = (vx * vx + vy * vy) + vz * vz;
  if (dr2 > ph.r0sq_max) return;
  const long long m2 = meta_of(p2);
  const int m = (int)(m2 & MF_TYPE);
  if (m == 0) return;                                          // limbo / removed
  if (pass == 1 && !(m2 & MF_ANYREF)) return;                  // reverse visits come from row owners only
  const int km = k3 + m - 1;
  if (dr2 > ph.r0sq[km]) return;
  const double4 t = lj_terms(vx, vy, vz, dr2, ph.eps[km], ph.r0p6[km]);
  // weight 2 = own visit + the reverse visit by j, when j is a row owner that sees i (exact doubling, one rounding per add)
  const double w = (pass == 0 && (m2 & MF_ANYREF) && (asym == 0 || (asym == 1 && !i_halo))) ? 2.0 : 1.0;
  a.fx += w * t.x; a.fy += w * t.y; a.fz += w * t.z; a.ep += w * t.w;
}
// Gather skip: an entry whose build-time distance D satisfies D - S > r0_max cannot be inside any cut-off, S being a
// bound of |move of i| + |move of j| since the rows were built: the largest displacement recorded for the z-layers
// around i at the last test_update (or the global top-2 sum if smaller), plus what maxz (z-dependent) and the
// integrator moved since.  qmax is the largest quantised D that still has to be looked at (d_qtab tabulates it per z-layer).
// Full row walk of one ref particle (taken when the near list of its head does not hold every entry that has to be looked at, or
// when rows are asymmetric): own row, then the transposed one; eight build-distance bytes per trip.
__device__ __noinline__ void fuerza_row(const double4 *__restrict__ posm, const int *__restrict__ cols,
                                        const int *__restrict__ rev_start, const int *__restrict__ rev_len, const int *__restrict__ rev_cols,
                                        const unsigned char *__restrict__ bq, const unsigned char *__restrict__ rev_bq,
                                        const Geo &g, const Phys &ph, const double4 &p1, int k3, int s, int start, int len0,
                                        int qmax, int asym, bool i_halo, FAcc &a) {
  const int npass = asym ? 2 : 1;
  for (int pass = 0; pass < npass; ++pass) {
    const int off = pass == 0 ? start : rev_start[s];
    const int *lst = (pass == 0 ? cols : rev_cols) + off;
    const unsigned char *lq = (pass == 0 ? bq : rev_bq) + off;
    const int len = pass == 0 ? len0 : rev_len[s];
    for (int j0 = 0; j0 < len; j0 += 8) {
      unsigned int need = 0u;                                      // eight build-distance bytes per trip, loads back to back
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int jj = j0 + q;
        const int b = jj < len ? (int)__ldg(&lq[jj]) : 1000;
        need |= (b <= qmax ? 1u : 0u) << q;
      }
      while (need) {
        const int q = __ffs(need) - 1; need &= need - 1;
        fuerza_pair(g, ph, p1, k3, ld_rec_nc(&posm[__ldg(&lst[j0 + q])]), pass, asym, i_halo, a);
      }
    }
  }
}

// One thread per slot; two dependent memory round trips for nearly every particle of a solution:
//  1. the record, the 32-byte row head (near list: the four nearest entries of the row as slot ids with their build distances, and
//     the fifth-nearest distance) and the "force is non-zero" byte, all addressed by the slot alone;
//  2. the records of the near-list partners whose build distance is within the skip bound of the particle's z-layer (the bound is
//     tabulated per layer by every block for itself while round trip 1 is in flight).
// Only when the fifth-nearest entry is within the bound too (dense neighbourhoods, long after a rebuild) or the rows are asymmetric
// does the thread walk the row itself (index and distance-byte loads: a third dependent round trip, out of line).  The terms are
// added in row order on both paths.  A particle with nothing inside its cut-offs whose stored force is already zero writes nothing.
// Round 1: record + head with sixteen distance bytes, then indices, then partners (41 us at 1 M particles); an intermediate form
// that skipped the head behind a one-byte "nearest distance" array moved fewer bytes and was slower (a fourth dependent trip).
// FUSEB: the thread that holds the finished force also does the particle's ermak_b update (dana.F90:1031-1052), same arithmetic
// as k_ermak_b.
template <bool FUSEB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) k_fuerza_sub(
    const double4 *__restrict__ posm, const RowHead *__restrict__ rh,
    const int *__restrict__ cols, const int *__restrict__ rev_start, const int *__restrict__ rev_len,
    const int *__restrict__ rev_cols, const unsigned char *__restrict__ bq, const unsigned char *__restrict__ rev_bq,
    const unsigned char *__restrict__ halo_of, const unsigned int *__restrict__ lay,
    const DevScal *__restrict__ sc, double4 *__restrict__ fe, const __grid_constant__ Geo g, const __grid_constant__ Phys ph, int n,
    double *__restrict__ vel, double *__restrict__ acel, const double *__restrict__ ranv, unsigned char *__restrict__ fnz,
    const unsigned char *__restrict__ dq, unsigned char *__restrict__ kb) {
  __shared__ unsigned int s_qt[LAY_MAX / 4], s_qb[LAY_MAX / 4];
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = s < n;
  // everything addressed by the slot alone is requested together
  const double4 p1 = in ? ld_rec_nc(&posm[s]) : make_double4(0, 0, 0, 0);
  const int4 nr = in ? rh_near(&rh[s]) : make_int4(-1, -1, -1, -1);
  RowMeta rm; rm.nbq = 0xffffffffu; rm.start = 0; rm.len = 0; rm.cap = 0; rm.q5 = 255;
  if (in) rm = rh_meta(&rh[s]);
  const int was_nz = in ? (int)fnz[s] : 0;
  const int dqi = (in && dq) ? (int)__ldg(&dq[s]) : 255;  // own displacement since the build (see dq_byte); 255: refinement off
  {
    // skip bound per z-layer for the largest cut-off of the pair table, from the displacement table of the last test_update and what
    // the integrator / maxz moved since (same formula as d_qtab; every block tabulates it for itself while its records travel)
    // (both displacement tables are requested before lay_cur is known: one round trip, in parallel with the record's)
    unsigned char *q8 = reinterpret_cast<unsigned char *>(s_qt), *qb8 = reinterpret_cast<unsigned char *>(s_qb);
    const double thick = g.cell[2] * (double)(1 << g.lay_shift);
    const double Rl = 2.0 * sqrt(g.rc_list2);
    for (int l = threadIdx.x; l < g.nlay; l += blockDim.x) {
      unsigned int m0 = 0u, m1_ = 0u;
#pragma unroll
      for (int d = -2; d <= 2; ++d) { const int q = l + d; if (q >= 0 && q < g.nlay) { m0 = max(m0, __ldg(&lay[q])); m1_ = max(m1_, __ldg(&lay[LAY_MAX + q])); } }
      const unsigned int mx = __ldg(&sc->lay_cur) ? m1_ : m0;
      const double maxz_fac = __ldg(&sc->maxz_fac), z0 = __ldg(&sc->z0), zmax = __ldg(&sc->zmax), dsum = __ldg(&sc->dsum_tu);
      const double sdisp = (double)__int_as_float((int)__ldg(&sc->step_disp_bits));
      double ztop = thick * (double)(l + 1);
      if (l == g.nlay - 1) ztop = fmax(ztop, zmax);
      const double since = maxz_fac * fmax(ztop + 2.0 * thick - z0, 0.0) + sdisp;
      const double kappa = __ldg(&sc->kappa_tu);
      const double S = pair_shift_bound(g, (double)__int_as_float((int)mx), dsum, kappa, maxz_fac, sdisp);
      q8[l] = (unsigned char)((since > g.cell[2]) ? 255 : (int)fmin(255.0, ceil((ph.r0_max * 1.000001 + S) * g.bq_scale) + 1.0));
      // the same without the particle's own share (qtab_base_entry): qmax = min(q8, qb8 + dq)
      const double Sb = (double)__int_as_float((int)mx) + kappa * Rl + maxz_fac * Rl + 2.0 * sdisp;
      qb8[l] = (unsigned char)((since > g.cell[2]) ? 255 : (int)fmin(255.0, ceil((ph.r0_max * 1.000001 + Sb) * g.bq_scale) + 1.0));
    }
  }
  __syncthreads();
  const long long m1 = meta_of(p1);
  if (in) kb[s] = (m1 & MF_REF) ? (unsigned char)(m1 & MF_TYPE) : (unsigned char)0;   // what ermak_b needs of the record (k_ermak_b_flat)
  if (!in || !(m1 & MF_REF)) return;
  FAcc a = {0.0, 0.0, 0.0, 0.0};
  const int lyr = layer_of(g, p1.z);
  const int qmax = min((int)reinterpret_cast<const unsigned char *>(s_qt)[lyr], min(255, (int)reinterpret_cast<const unsigned char *>(s_qb)[lyr] + dqi));
  const int asym = __ldg(&sc->rows_asym);                // 0 symmetric, 1 halo-only, 2 general
  const int k3 = ((int)(m1 & MF_TYPE) - 1) * 3;
  // asym == 1 (some particle sits in a halo cell above box(3): the usual state of a piston run): only the rows of those few
  // particles are transposed, so a particle outside the halo that nobody's transposed row mentions is as good as symmetric
  const bool plain = asym == 0 || (asym == 1 && halo_of[s] == 0 && rev_len[s] == 0);
  if (plain && rm.q5 > qmax) {
    // every entry that has to be looked at is in the near list (row order)
    // (usually none or one of them is within the bound: no need to keep four partner records in flight)
    const int nrs[4] = {nr.x, nr.y, nr.z, nr.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if ((int)((rm.nbq >> (8 * k)) & 255u) <= qmax && nrs[k] >= 0) fuerza_pair(g, ph, p1, k3, ld_rec_nc(&posm[nrs[k]]), 0, asym, false, a);
  } else {
    const bool i_halo = asym == 1 && halo_of[s] != 0;
    fuerza_row(posm, cols, rev_start, rev_len, rev_cols, bq, rev_bq, g, ph, p1, k3, s, rm.start, rm.len, qmax, asym, i_halo, a);
  }
  // fnz[s] == 0 guarantees that fe[s] already holds zeros: a particle with nothing inside its cut-offs (nearly all of them in
  // solution) then writes nothing at all
  const int nz = (a.fx != 0.0 || a.fy != 0.0 || a.fz != 0.0 || a.ep != 0.0) ? 1 : 0;
  if (nz | was_nz) st_rec(&fe[s], make_double4(a.fx, a.fy, a.fz, a.ep));
  if (nz != was_nz) fnz[s] = (unsigned char)nz;
  if (FUSEB) {
    const int zt = (int)(m1 & MF_TYPE);
    if (zt != 2) {
      const double mass = ph.mass[zt - 1];
      const double fv[3] = {a.fx, a.fy, a.fz};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double v = vel[3 * s + k], ac = acel[3 * s + k];
        vel[3 * s + k] = ph.cc0 * v + ph.cc1mcc2 * ac + ph.cc2 * fv[k] / mass + __ldg(&ranv[3 * s + k]);
        acel[3 * s + k] = fv[k] / mass;
      }
    }
  }
}

// ================================================================================================
// K4  integrators + boundary handling   (ermak_a dana.F90:974-1028, cbrownian_hs 798-846, atom_pbc 1187-1250)
// ================================================================================================
struct BlockAcc { long long tr, de; double msd, mv; float dmax; };

// One set of atomics per BLOCK: a per-warp flush put 60 k atomics per call on two addresses at 1 M particles.
__device__ __forceinline__ void block_flush(BlockAcc a, DevScal *sc) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a.tr += __shfl_xor_sync(0xffffffffu, a.tr, o); a.de += __shfl_xor_sync(0xffffffffu, a.de, o);
    a.msd += __shfl_xor_sync(0xffffffffu, a.msd, o); a.mv = fmax(a.mv, __shfl_xor_sync(0xffffffffu, a.mv, o));
    a.dmax = fmaxf(a.dmax, __shfl_xor_sync(0xffffffffu, a.dmax, o));
  }
  __shared__ long long s_tr[32], s_de[32];
  __shared__ double s_msd[32], s_mv[32];
  __shared__ float s_dm[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { s_tr[w] = a.tr; s_de[w] = a.de; s_msd[w] = a.msd; s_mv[w] = a.mv; s_dm[w] = a.dmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < nw; ++i) { a.tr += s_tr[i]; a.de += s_de[i]; a.msd += s_msd[i]; a.mv = fmax(a.mv, s_mv[i]); a.dmax = fmaxf(a.dmax, s_dm[i]); }
    if (a.tr) atomicAdd((unsigned long long *)&sc->try_, (unsigned long long)a.tr);
    if (a.de) atomicAdd((unsigned long long *)&sc->depo, (unsigned long long)a.de);
    if (a.msd != 0.0) atomicAdd(&sc->msd_t, a.msd);
    if (a.mv > 0.0) atomicMax((unsigned long long *)&sc->max_vel, (unsigned long long)__double_as_longlong(a.mv));
    if (a.dmax > 0.0f) atomicMax(&sc->step_disp_bits, (unsigned int)__float_as_int(a.dmax));
  }
}

// atom_pbc — dana.F90:1187-1250.  Returns depos.  The deposition uniform is drawn only when z<=0.
struct RngSrc { int mode; unsigned long long seed; unsigned int id, step; const double *rp; int slot; RefRng *rr; };
// pos_old is only touched when the particle crosses a periodic face (a few per thousand per step) and, under Ermak, vel only when
// it bounces off the ceiling: both are read / written on demand (pos_old_s points at the slot's three doubles, wrote_v reports the
// bounce), which takes 72 of the 236 bytes per particle out of the streaming pass.
__device__ __forceinline__ bool atom_pbc_dev(const Geo &g, const Phys &ph, double zmax, double q[3], double *__restrict__ pos_old_s, const double og[3],
                                             double v[3], long long &meta, const RngSrc &rs, BlockAcc &acc, bool &wrote_v) {
  bool depos = false;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (q[j] > g.box[j]) { q[j] = q[j] - g.box[j]; pos_old_s[j] = pos_old_s[j] - g.box[j]; }
    if (q[j] < 0.0) { q[j] = q[j] + g.box[j]; pos_old_s[j] = pos_old_s[j] + g.box[j]; }
  }
  if (q[2] > zmax) {
    if (ph.integrador) { q[2] = q[2] - 2 * (q[2] - zmax); v[2] = -v[2]; wrote_v = true; }
    else { q[0] = og[0]; q[1] = og[1]; q[2] = og[2]; }
  }
  acc.msd += v[0] * v[0] * ph.h * ph.h;
  if (q[2] <= 0.0) {
    acc.tr++;
    double ne;
    if (rs.mode == 1) ne = rs.rp[rs.slot];
    else if (rs.mode == 2) ne = ref_ran(rs.rr);             // the reference's stream, drawn in list order by the sequential kernel
    else { Philox r; r.run(rs.seed, rs.id, rs.step, RS_PBC, 0u); ne = r.u01(0); }
    if (ne < ph.prob) { acc.de++; meta = (meta & ~MF_TYPE) | 3; depos = true; }
    q[0] = og[0]; q[1] = og[1]; q[2] = og[2];
  }
  return depos;
}

// One ref particle through ermak_a / cbrownian_hs + atom_pbc, its Gaussians already drawn (gs: r1, r2 per axis for Ermak, one
// per axis for the Brownian step — the order the reference draws them in, dana.F90:1006-1015, 824-829).
template <bool ERMAK>
__device__ __forceinline__ void integrate_one(double4 *__restrict__ posm, double *__restrict__ vel, double *__restrict__ pos_old,
                                              double *__restrict__ old_cg, double *__restrict__ ranv, DevScal *__restrict__ sc, const Geo &g, const Phys &ph,
                                              double4 p, double v[3], const double a[3], const double gs[6], const RngSrc &rs, int s, BlockAcc &acc) {
  long long m = meta_of(p);
  double q[3] = {p.x, p.y, p.z}, og[3] = {p.x, p.y, p.z};
  // written once per step and read by few: streaming stores keep these 48 bytes per particle from sitting dirty in the L2 in front
  // of the pair force (which ran 66 us right behind this kernel and 31 us behind a clean L2)
  __stcs(&old_cg[3 * s], og[0]); __stcs(&old_cg[3 * s + 1], og[1]); __stcs(&old_cg[3 * s + 2], og[2]);
  int zt = (int)(m & MF_TYPE);
  if (ERMAK) {
    double sm = ph.sqrt_mass[zt - 1];
    double A = ph.skt / sm * ph.sdr, B = ph.skt / sm * ph.sdv;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double r1 = gs[2 * j], r2 = gs[2 * j + 1];
      double ranr = A * r1;
      q[j] = q[j] + ph.cc1 * v[j] + ph.cc2h * a[j] + ranr;
      __stcs(&ranv[3 * s + j], B * (ph.crv1 * r1 + ph.crv2 * r2));
    }
  } else {
    double fac1 = (q[2] > ph.z_sei) ? ph.fac_sc : ph.fac_sei;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double posold = q[j];
      q[j] = posold + gs[j] * fac1;
      v[j] = (q[j] - posold) / ph.h;
    }
  }
  double zmax = sc->zmax;
  bool wrote_v = !ERMAK;                               // the Brownian step always rewrites vel
  bool depos = atom_pbc_dev(g, ph, zmax, q, pos_old + 3 * (size_t)s, og, v, m, rs, acc, wrote_v);
  if (!depos) {
    if (!ERMAK) acc.mv = fmax(acc.mv, (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    m &= ~MF_SKIP;
  }
  {
    double dx = q[0] - og[0], dy = q[1] - og[1], dz = q[2] - og[2];
    dx = dx - g.box[0] * round(dx * g.one_box[0]); dy = dy - g.box[1] * round(dy * g.one_box[1]);
    float df = __double2float_ru(sqrt(dx * dx + dy * dy + dz * dz)) * 1.000001f;
    m = with_disp(m, (unsigned int)__float_as_int(df));
    acc.dmax = fmaxf(acc.dmax, df);
  }
  p.x = q[0]; p.y = q[1]; p.z = q[2]; p.w = meta_as_double(m);
  st_rec(&posm[s], p);
  if (wrote_v) { __stcs(&vel[3 * s], v[0]); __stcs(&vel[3 * s + 1], v[1]); __stcs(&vel[3 * s + 2], v[2]); }
}

template <bool ERMAK>
__global__ void __launch_bounds__(TPB, 4) k_integrate(double4 *__restrict__ posm, double *__restrict__ vel, const double *__restrict__ acel,
                                                   double *__restrict__ pos_old, double *__restrict__ old_cg, double *__restrict__ ranv,
                                                   const int *__restrict__ uid, const double *__restrict__ rp_gauss,
                                                   const double *__restrict__ rp_upbc, DevScal *__restrict__ sc, Geo g, Phys ph,
                                                   unsigned int step, int n) {
  if (step == STEP_FROM_DEVICE) step = sc->istep;
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  BlockAcc acc = {0, 0, 0.0, 0.0, 0.0f};
  if (s < n) {
    // everything addressed by the slot alone is requested together: one memory round trip in front of the arithmetic
    double4 p = ld_rec(&posm[s]);
    double v[3] = {0.0, 0.0, 0.0}, a[3] = {0.0, 0.0, 0.0};
    if (ERMAK) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { v[j] = __ldcs(&vel[3 * s + j]); a[j] = __ldcs(&acel[3 * s + j]); }
    }
    const unsigned int id = (unsigned int)uid[s];
    if (meta_of(p) & MF_REF) {
      double gs[6];
      if (ph.rng_mode == 1) {
#pragma unroll
        for (int i = 0; i < (ERMAK ? 6 : 3); ++i) gs[i] = rp_gauss[6 * s + i];
      } else {
        Philox r;
        double sp0, sp1;
        r.run(ph.seed, id, step, RS_INTEG0, 0u); r.gauss4f(gs[0], gs[1], gs[2], gs[3]);
        if (ERMAK) { r.run(ph.seed, id, step, RS_INTEG1, 0u); r.gauss4f(gs[4], gs[5], sp0, sp1); }
      }
      RngSrc rs = {ph.rng_mode, ph.seed, id, step, rp_upbc, s, nullptr};
      integrate_one<ERMAK>(posm, vel, pos_old, old_cg, ranv, sc, g, ph, p, v, a, gs, rs, s, acc);
    }
  }
  block_flush(acc, sc);
}

// DML_RNG_REFERENCE: the reference draws from ONE sequential stream while it walks hs%ref in list order (creation rank), and an
// atom that reaches z <= 0 takes one more uniform in the middle of it (atom_pbc, dana.F90:1236-1240): whether it does depends on
// the move it just made.  So one thread walks the ref atoms in creation-rank order (ord[rank] = slot, -1 = none), drawing and
// integrating as it goes.  This mode exists to make the GPU-backed binary reproduce the reference's own test cases (tests/test.sh)
// digit for digit; production runs use the counter-based generator.
__global__ void k_ord_scatter(const double4 *__restrict__ posm, const int *__restrict__ uid, int *__restrict__ ord, int n, int nord) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  if (!(meta_of(ld_rec_nc(&posm[s])) & MF_REF)) return;
  const int u = uid[s];
  if (u >= 0 && u < nord) ord[u] = s;
}
template <bool ERMAK>
__global__ void k_integrate_seq(double4 *__restrict__ posm, double *__restrict__ vel, const double *__restrict__ acel,
                                double *__restrict__ pos_old, double *__restrict__ old_cg, double *__restrict__ ranv,
                                const int *__restrict__ ord, int nord, DevScal *__restrict__ sc, Geo g, Phys ph) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  BlockAcc acc = {0, 0, 0.0, 0.0, 0.0f};
  RefRng rr = sc->rr;
  for (int u = 0; u < nord; ++u) {
    const int s = ord[u];
    if (s < 0) continue;
    double4 p = ld_rec(&posm[s]);
    double v[3] = {vel[3 * s], vel[3 * s + 1], vel[3 * s + 2]}, a[3] = {acel[3 * s], acel[3 * s + 1], acel[3 * s + 2]};
    double gs[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < (ERMAK ? 6 : 3); ++i) gs[i] = ref_gasdev(&rr);
    RngSrc rs = {2, 0ull, 0u, 0u, nullptr, s, &rr};
    integrate_one<ERMAK>(posm, vel, pos_old, old_cg, ranv, sc, g, ph, p, v, a, gs, rs, s, acc);
  }
  sc->rr = rr;
  if (acc.tr) sc->try_ += acc.tr;
  if (acc.de) sc->depo += acc.de;
  if (acc.msd != 0.0) sc->msd_t += acc.msd;
  if (acc.mv > 0.0) sc->max_vel = fmax(sc->max_vel, acc.mv);
  if (acc.dmax > 0.0f) sc->step_disp_bits = max(sc->step_disp_bits, (unsigned int)__float_as_int(acc.dmax));
}
// overlap_moveback in this mode: with prob >= 1 the value of a deposition uniform never matters, only how many were drawn
__global__ void k_rng_mark(DevScal *sc) { sc->rr_mark = sc->try_; }
__global__ void k_rng_advance(DevScal *sc) {
  RefRng rr = sc->rr;
  for (long long i = sc->rr_mark; i < sc->try_; ++i) (void)ref_ran(&rr);
  sc->rr = rr;
}

// ermak_b — dana.F90:1031-1052
__global__ void __launch_bounds__(TPB) k_ermak_b(const double4 *__restrict__ posm, double *__restrict__ vel, double *__restrict__ acel,
                                                 const double4 *__restrict__ fe, const double *__restrict__ ranv, Phys ph, int n,
                                                 const unsigned char *__restrict__ fnz) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  long long m = meta_of(ld_rec_nc(&posm[s]));
  const int nz = (int)__ldg(&fnz[s]);
  if (!(m & MF_REF)) return;
  int zt = (int)(m & MF_TYPE);
  if (zt == 2) return;
  double mass = ph.mass[zt - 1];
  const double4 f4 = nz ? ld_rec_nc(&fe[s]) : make_double4(0.0, 0.0, 0.0, 0.0);     // fnz == 0: fe[s] holds zeros (k_fuerza_sub)
  const double fv[3] = {f4.x, f4.y, f4.z};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double f = fv[k], a = __ldcs(&acel[3 * s + k]), v = __ldcs(&vel[3 * s + k]);
    __stcs(&vel[3 * s + k], ph.cc0 * v + ph.cc1mcc2 * a + ph.cc2 * f / mass + __ldcs(&ranv[3 * s + k]));
    __stcs(&acel[3 * s + k], f / mass);
  }
}

// ermak_b right behind a pair-force call, one thread per COMPONENT: the update is elementwise over vel / acel / ranv ([n][3]
// doubles read as flat arrays: 8 contiguous bytes per thread instead of three 24-byte-strided requests per array), the force
// component comes from fe only where fnz says it is non-zero, and element + ref membership come from the byte the pair-force kernel
// left per slot (kb) instead of the 32-byte record.  Same arithmetic as k_ermak_b.
__global__ void __launch_bounds__(TPB) k_ermak_b_flat(double *__restrict__ vel, double *__restrict__ acel, const double4 *__restrict__ fe,
                                                      const double *__restrict__ ranv, Phys ph, int n3, const unsigned char *__restrict__ fnz,
                                                      const unsigned char *__restrict__ kb) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n3) return;
  const int s = e / 3, k = e - 3 * s;
  const int zt = (int)__ldg(&kb[s]);
  if (zt == 0 || zt == 2) return;
  const double mass = ph.mass[zt - 1];
  const double f = __ldg(&fnz[s]) ? reinterpret_cast<const double *>(&fe[s])[k] : 0.0;
  const double a = __ldcs(&acel[e]), v = __ldcs(&vel[e]);
  __stcs(&vel[e], ph.cc0 * v + ph.cc1mcc2 * a + ph.cc2 * f / mass + __ldcs(&ranv[e]));
  __stcs(&acel[e], f / mass);
}

// ================================================================================================
// K7  overlap_moveback (dana.F90:849-943) — exact sequential semantics on a parallel machine.
//     During the resolution an atom is either at its moved position or at old_cg, so the mutable state
//     is one int per atom (ovst).  Pairs that can ever come within rcut (any of the 4 new/old
//     combinations) define a conflict graph; its connected components (CG atoms do not connect) are
//     independent, so each is resolved by one thread replaying the reference's passes over the
//     component's atoms in ascending creation rank (= order of hs%ref%alist) and row order.
// ================================================================================================
constexpr int OV_MOVED = 1, OV_SKIP = 2, OV_TSHIFT = 2, OV_INVOLVED = 16, OV_ZERO = 32;

__device__ __forceinline__ void p_ov_init(const double4 *__restrict__ posm, int *__restrict__ parent, int *__restrict__ ovst,
                          int *__restrict__ comp_cnt, int *__restrict__ ov_head, DevScal *__restrict__ sc, int n) {
  if (blockIdx.x == 0 && threadIdx.x == 0) { sc->again = 0; sc->n_roots = 0; sc->member_cursor = 0; sc->ch_later = 0; sc->any_active = 0; }
  const int s_end = n;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < s_end; s += gridDim.x * blockDim.x) {
    long long m = meta_of(ld_rec_nc(&posm[s]));
    parent[s] = s; comp_cnt[s] = 0; ov_head[s] = -1;
    ovst[s] = ((m & MF_SKIP) ? OV_SKIP : 0) | ((int)(m & MF_TYPE) << OV_TSHIFT);
  }
}
__global__ void k_ov_init(const double4 *__restrict__ posm, int *__restrict__ parent, int *__restrict__ ovst,
                          int *__restrict__ comp_cnt, int *__restrict__ ov_head, DevScal *__restrict__ sc, int n) { p_ov_init(posm, parent, ovst, comp_cnt, ov_head, sc, n); }
__device__ __forceinline__ int uf_find(int *parent, int x) {
  for (;;) {
    int y = ((volatile int *)parent)[x];
    if (y == x) return x;
    int z = ((volatile int *)parent)[y];
    if (z != y) parent[x] = z;       // path halving (benign race: only ever points higher up the same tree)
    x = y;
  }
}
__device__ __forceinline__ void uf_unite(int *parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }           // hook the larger root under the smaller: no cycles
    int old = atomicCAS(&parent[a], a, b);
    if (old == a) return;
  }
}
// L lanes share one particle (L = 1, 2 or 4 adjacent threads): lane l looks at entries l, l+L, ... of the row (or at near entry
// l, l+L, ...), so the dependent gathers of one row run side by side.  The conflict graph does not depend on the order the
// edges are found in (uf_unite hooks the larger root under the smaller one).
template <int L>
__device__ __forceinline__ void p_ov_detect(const double4 *__restrict__ posm, const double *__restrict__ old_cg,
                                                   const RowHead *__restrict__ rh,
                                                   const int *__restrict__ cols, const unsigned char *__restrict__ bq,
                                                   const unsigned int *__restrict__ lay, int *__restrict__ parent, int *__restrict__ ovst,
                                                   DevScal *__restrict__ sc, Geo g, int n, const unsigned char *__restrict__ dq = nullptr) {
  const int s_end = n;
  const double rcut = sqrt(g.rcut2);
  const int lane = L > 1 ? (int)(threadIdx.x % L) : 0;
  for (int s = (blockIdx.x * blockDim.x + threadIdx.x) / L; s < s_end; s += gridDim.x * blockDim.x / L) {
    // one round trip for everything addressed by the slot alone
    const double4 p1 = ld_rec_nc(&posm[s]);
    const int4 nr = rh_near(&rh[s]);
    const RowMeta rm = rh_meta(&rh[s]);
    const long long m1 = meta_of(p1);
    if (!(m1 & MF_REF)) continue;
    const float d1 = disp_of(m1);
    const int qmax = skip_qmax(g, lay, p1.z, 1, dq, s);   // same build-distance skip as the pair force (covers new and old positions)
    const int b = rm.start, len = rm.len;
    bool inv = false, have_o1 = false;
    double o1[3] = {0.0, 0.0, 0.0};
    // one entry of the row: conflict-graph edge when any new/old combination of the pair can be within rcut
    auto look = [&](int j) {
      const double4 p2 = ld_rec_nc(&posm[j]);
      const long long m2 = meta_of(p2);
      if (!(m2 & MF_TYPE)) return;
      // Exact-safe prefilter: by the triangle inequality no new/old combination can be within rcut when the current
      // separation exceeds rcut + |move of i| + |move of j| (bounds carried in the records, tiny relative margin).
      const double rd_nn = dist2_idnint(g, p1.x, p1.y, p1.z, p2.x, p2.y, p2.z);
      {
        // (slab mode: a ghost that is mobile on its owner, MF_GREF, carries an infinite bound and its old_cg was saved
        //  when the step started, so it joins the conflict graph like a ref atom)
        const double thr = (rcut + (double)d1 + ((m2 & MF_ANYREF) ? (double)disp_of(m2) : 0.0)) * 1.000001;
        if (rd_nn > thr * thr) return;
      }
      if (!have_o1) { o1[0] = old_cg[3 * s]; o1[1] = old_cg[3 * s + 1]; o1[2] = old_cg[3 * s + 2]; have_o1 = true; }
      bool hit = rd_nn <= g.rcut2 || dist2_idnint(g, o1[0], o1[1], o1[2], p2.x, p2.y, p2.z) <= g.rcut2;
      if (m2 & MF_ANYREF) {
        const double o2[3] = {old_cg[3 * j], old_cg[3 * j + 1], old_cg[3 * j + 2]};
        hit = hit || dist2_idnint(g, p1.x, p1.y, p1.z, o2[0], o2[1], o2[2]) <= g.rcut2 ||
              dist2_idnint(g, o1[0], o1[1], o1[2], o2[0], o2[1], o2[2]) <= g.rcut2;
        if (hit) { uf_unite(parent, s, j); atomicOr(&ovst[j], OV_INVOLVED); }
      }
      inv = inv || hit;
    };
    if (rm.q5 > qmax) {                                   // every entry that has to be looked at is in the near list of the head
      const int nrs[4] = {nr.x, nr.y, nr.z, nr.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) if ((L == 1 || (k % L) == lane) && (int)((rm.nbq >> (8 * k)) & 255u) <= qmax && nrs[k] >= 0) look(nrs[k]);
    } else if (L > 1) {
      for (int jj = lane; jj < len; jj += L)
        if ((int)__ldg(&bq[b + jj]) <= qmax) look(__ldg(&cols[b + jj]));
    } else {
      for (int j0 = 0; j0 < len; j0 += 16) {
        unsigned int need = 0u;
#pragma unroll
        for (int q = 0; q < 16; ++q) { const int jj = j0 + q; const int bb = jj < len ? (int)__ldg(&bq[b + jj]) : 1000; need |= (bb <= qmax ? 1u : 0u) << q; }
        while (need) {
          const int q = __ffs(need) - 1; need &= need - 1u;
          look(__ldg(&cols[b + j0 + q]));
        }
      }
    }
    if (inv) atomicOr(&ovst[s], OV_INVOLVED);
  }
}
template <int L>
__global__ void __launch_bounds__(TPB) k_ov_detect(const double4 *__restrict__ posm, const double *__restrict__ old_cg,
                                                   const RowHead *__restrict__ rh,
                                                   const int *__restrict__ cols, const unsigned char *__restrict__ bq,
                                                   const unsigned int *__restrict__ lay, int *__restrict__ parent, int *__restrict__ ovst,
                                                   DevScal *__restrict__ sc, Geo g, int n, const unsigned char *__restrict__ dq) {
  p_ov_detect<L>(posm, old_cg, rh, cols, bq, lay, parent, ovst, sc, g, n, dq);
}
__device__ __forceinline__ void p_ov_count(int *__restrict__ parent, const int *__restrict__ ovst, int *__restrict__ comp_cnt, int n) {

  const int s_end = n;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < s_end; s += gridDim.x * blockDim.x) {
    if (!(ovst[s] & OV_INVOLVED)) continue;
    int r = uf_find(parent, s);
    parent[s] = r;
    atomicAdd(&comp_cnt[r], 1);
  }
}
__global__ void k_ov_count(int *__restrict__ parent, const int *__restrict__ ovst, int *__restrict__ comp_cnt, int n) { p_ov_count(parent, ovst, comp_cnt, n); }
__device__ __forceinline__ void p_ov_alloc(const int *__restrict__ parent, const int *__restrict__ ovst, int *__restrict__ comp_cnt,
                           int *__restrict__ comp_off, int *__restrict__ roots, DevScal *__restrict__ sc, int n) {

  const int s_end = n;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < s_end; s += gridDim.x * blockDim.x) {
    if (!(ovst[s] & OV_INVOLVED) || parent[s] != s) continue;
    int c = comp_cnt[s];
    comp_off[s] = atomicAdd(&sc->member_cursor, c);
    comp_cnt[s] = 0;                                       // reused as the fill cursor
    roots[atomicAdd(&sc->n_roots, 1)] = s;
  }
}
__global__ void k_ov_alloc(const int *__restrict__ parent, const int *__restrict__ ovst, int *__restrict__ comp_cnt,
                           int *__restrict__ comp_off, int *__restrict__ roots, DevScal *__restrict__ sc, int n) { p_ov_alloc(parent, ovst, comp_cnt, comp_off, roots, sc, n); }
__device__ __forceinline__ void p_ov_fill(const int *__restrict__ parent, const int *__restrict__ ovst, int *__restrict__ comp_cnt,
                          const int *__restrict__ comp_off, int *__restrict__ members, int n) {

  const int s_end = n;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < s_end; s += gridDim.x * blockDim.x) {
    if (!(ovst[s] & OV_INVOLVED)) continue;
    int r = parent[s];
    members[comp_off[r] + atomicAdd(&comp_cnt[r], 1)] = s;
  }
}
__global__ void k_ov_fill(const int *__restrict__ parent, const int *__restrict__ ovst, int *__restrict__ comp_cnt,
                          const int *__restrict__ comp_off, int *__restrict__ members, int n) { p_ov_fill(parent, ovst, comp_cnt, comp_off, members, n); }
// order the members [b,e) of a component by creation rank: insertion sort for the usual handful, heap sort beyond
__device__ __forceinline__ void ov_sort_members(int *__restrict__ members, const int *__restrict__ uid, int b, int e) {
  const int n = e - b;
  if (n <= 32) {
    for (int i = b + 1; i < e; ++i) {
      int s = members[i], key = uid[s], j = i - 1;
      while (j >= b && uid[members[j]] > key) { members[j + 1] = members[j]; --j; }
      members[j + 1] = s;
    }
    return;
  }
  int *a = members + b;
  auto sift = [&](int start, int end) {
    int root = start;
    for (;;) {
      int child = 2 * root + 1;
      if (child > end) break;
      if (child + 1 <= end && uid[a[child]] < uid[a[child + 1]]) ++child;
      if (uid[a[root]] < uid[a[child]]) { int t = a[root]; a[root] = a[child]; a[child] = t; root = child; } else break;
    }
  };
  for (int st = (n - 2) / 2; st >= 0; --st) sift(st, n - 1);
  for (int end = n - 1; end > 0; --end) { int t = a[end]; a[end] = a[0]; a[0] = t; sift(0, end - 1); }
}
// one thread per component: order the members by creation rank (once)
__global__ void k_ov_sort(const int *__restrict__ roots, const int *__restrict__ comp_cnt, const int *__restrict__ comp_off,
                          int *__restrict__ members, const int *__restrict__ uid, const DevScal *__restrict__ sc) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= sc->n_roots) return;
  int root = roots[r], b = comp_off[root], e = b + comp_cnt[root];
  ov_sort_members(members, uid, b, e);
}
__device__ __forceinline__ void ov_pos(const double4 *posm, const double *old_cg, int a, int st, double q[3]) {
  if (st & OV_MOVED) { q[0] = old_cg[3 * a]; q[1] = old_cg[3 * a + 1]; q[2] = old_cg[3 * a + 2]; }
  else { double4 p = ld_rec_nc(&posm[a]); q[0] = p.x; q[1] = p.y; q[2] = p.z; }
}
// one reference pass (one recursion level) over the members [b,e) of one component; returns "again"
struct OvAcc { long long tr, de, ch, ch3; };
// Injected deposition uniforms of overlap_moveback (trace-replay parity mode): the k-th draw an atom makes inside one call reads
// vals[qstart[slot] + k] (the reference draws a fresh ran(idum) at every retry of a failed deposition, dana.F90:898-911); the
// per-slot draw counters are cleared by k_ov_init.  vals == nullptr: nothing injected (prob >= 1: the value never matters).
struct OvRp { const double *vals; const int *qstart; int *draws; };
__device__ __forceinline__ double ov_replay_draw(const OvRp &rp, int a1, bool lead, DevScal *sc, double prob) {
  if (!rp.vals) return 0.0;
  const int k = rp.draws[a1], b = rp.qstart[a1], cnt = rp.qstart[a1 + 1] - b;
  __syncwarp(__activemask());
  if (lead) rp.draws[a1] = k + 1;
  if (k < cnt) return rp.vals[b + k];
  if (prob < 1.0 && lead) atomicCAS(&sc->err, 0, DML_E_REPLAY_EXHAUSTED);
  return 0.0;
}
__device__ __forceinline__ bool ov_one_pass(const double4 *__restrict__ posm, const double *__restrict__ old_cg,
                                            const RowHead *__restrict__ rh,
                                            const int *__restrict__ cols, const unsigned char *__restrict__ bq,
                                            const unsigned int *__restrict__ lay,
                                            int *__restrict__ ovst, const int *__restrict__ members,
                                            const int *__restrict__ uid, const OvRp rp_uovl, DevScal *__restrict__ sc,
                                            const Geo &g, const Phys &ph, unsigned int step, int pass, int guard, double z0,
                                            int b, int e, OvAcc &acc) {
  bool again = false;
  const double rcut = sqrt(g.rcut2);
  for (int i = b; i < e; ++i) {
    int a1 = members[i];
    int st1 = ((volatile int *)ovst)[a1];
    if (st1 & OV_SKIP) continue;
    if (!((st1 >> OV_TSHIFT) & 3)) continue;
    st1 |= OV_SKIP;
    double q1[3]; ov_pos(posm, old_cg, a1, st1, q1);
    const RowMeta rm = rh_meta(&rh[a1]);
    const int rb = rm.start, rl = rm.len;
    // the build-distance skip of k_ov_detect (same bound, same record): an entry it skipped cannot be within rcut in any
    // new/old combination, so skipping it here changes nothing and saves the dependent gathers of the replay
    const int qmax = skip_qtab(g, lay, ld_rec_nc(&posm[a1]).z, 1);
    for (int jj = 0; jj < rl; ++jj) {
      if ((int)__ldg(&bq[rb + jj]) > qmax) continue;
      int a2 = cols[rb + jj];
      int st2 = ((volatile int *)ovst)[a2];
      int t2 = (st2 >> OV_TSHIFT) & 3;
      if (t2 == 0) continue;                                   // limbo (dana.F90:881-883)
      double q2[3]; ov_pos(posm, old_cg, a2, st2, q2);
      double dr = dist2_idnint(g, q1[0], q1[1], q1[2], q2[0], q2[1], q2[2]);
      if (dr > g.rcut2) continue;
      if (t2 == 2) {                                           // contact with metal: deposition attempt
        acc.tr++;
        double ne;
        if (ph.rng_mode == 1) ne = ov_replay_draw(rp_uovl, a1, true, sc, ph.prob);
        else if (ph.rng_mode == 2) ne = 0.0;                      // prob >= 1 in this mode: k_rng_advance consumes the stream afterwards
        else { Philox rr; rr.run(ph.seed, (unsigned int)uid[a1], step, RS_OVERLAP, (unsigned int)pass); ne = rr.u01(0); }
        if (ne < ph.prob) {
          acc.de++; st1 = (st1 & ~(3 << OV_TSHIFT)) | (3 << OV_TSHIFT);
          if (q1[2] > z0) atomicCAS(&sc->err, 0, DML_E_SUPERO_Z0);
        } else {
          st1 |= OV_MOVED; st1 &= ~OV_SKIP;
          ov_pos(posm, old_cg, a1, st1, q1);
        }
        break;
      }
      if (ph.piston || guard) {                                // unsolvable pair guard (dana.F90:920-927)
        double og2[3] = {old_cg[3 * a2], old_cg[3 * a2 + 1], old_cg[3 * a2 + 2]};
        if (q2[0] == og2[0] && q2[1] == og2[1] && q2[2] == og2[2]) {
          double og1[3] = {old_cg[3 * a1], old_cg[3 * a1 + 1], old_cg[3 * a1 + 2]};
          if (q1[0] == og1[0] && q1[1] == og1[1] && q1[2] == og1[2]) { acc.ch3++; continue; }
        }
      }
      // o2 goes back to its previous position; velocities are zeroed when the state is applied
      st2 = (st2 | OV_MOVED | OV_ZERO) & ~OV_SKIP;
      ovst[a2] = st2;
      acc.ch++; again = true;
    }
    ovst[a1] = st1;
  }
  return again;
}
__device__ __forceinline__ void ov_flush(const OvAcc &acc, DevScal *sc) {
  if (acc.tr) atomicAdd((unsigned long long *)&sc->try_, (unsigned long long)acc.tr);
  if (acc.de) atomicAdd((unsigned long long *)&sc->depo, (unsigned long long)acc.de);
  if (acc.ch) atomicAdd((unsigned long long *)&sc->choques, (unsigned long long)acc.ch);
  if (acc.ch3) atomicAdd((unsigned long long *)&sc->choques3, (unsigned long long)acc.ch3);
}
// Global-synchronous variant: one launch = one recursion level for every component (needed when prob<1, where a
// failed deposition leaves skip=.false. without requesting another pass, so the pass count couples components).
__global__ void k_ov_pass(const double4 *__restrict__ posm, const double *__restrict__ old_cg, const RowHead *__restrict__ rh,
                          const int *__restrict__ cols, const unsigned char *__restrict__ bq,
                          const unsigned int *__restrict__ lay, int *__restrict__ ovst,
                          const int *__restrict__ roots, const int *__restrict__ comp_cnt, const int *__restrict__ comp_off,
                          const int *__restrict__ members, const int *__restrict__ uid, const OvRp rp_uovl,
                          DevScal *__restrict__ sc, Geo g, Phys ph, unsigned int step, int pass, int guard) {
  if (step == STEP_FROM_DEVICE) step = sc->istep;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= sc->n_roots) return;
  int root = roots[r], b = comp_off[root], e = b + comp_cnt[root];
  OvAcc acc = {0, 0, 0, 0};
  bool again = ov_one_pass(posm, old_cg, rh, cols, bq, lay, ovst, members, uid, rp_uovl, sc, g, ph, step, pass, guard, sc->z0, b, e, acc);
  ov_flush(acc, sc);
  if (pass >= 1 && acc.ch) atomicAdd((unsigned long long *)&sc->ch_later, (unsigned long long)acc.ch);
  if (again) sc->again = 1;
}
// prob>=1 path (every metal contact deposits, nothing couples components): the members of a component are chained into a
// list hanging off its root (one kernel instead of count / allocate / fill), and one WARP per component replays all
// recursion levels locally.  No host round trip.
__device__ __forceinline__ void p_ov_link(int *__restrict__ parent, const int *__restrict__ ovst, int *__restrict__ ov_head,
                                          int *__restrict__ ov_next, int *__restrict__ roots, DevScal *__restrict__ sc, int n) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    if (!(ovst[s] & OV_INVOLVED)) continue;
    const int r = uf_find(parent, s);
    ov_next[s] = atomicExch(&ov_head[r], s);
    if (r == s) roots[atomicAdd(&sc->n_roots, 1)] = s;
  }
}
__global__ void k_ov_link(int *__restrict__ parent, const int *__restrict__ ovst, int *__restrict__ ov_head, int *__restrict__ ov_next,
                          int *__restrict__ roots, DevScal *__restrict__ sc, int n) { p_ov_link(parent, ovst, ov_head, ov_next, roots, sc, n); }

// One warp per component.  Members are visited in ascending creation rank (= order of hs%ref%alist) like the reference;
// inside a visit the lanes take one row entry each.  That is exact because the entries of one row are independent given the
// states at the start of the visit (they are distinct atoms, and an entry only changes its own state), except for the `exit`
// at the first metal contact, which is applied with a ballot: entries behind it are ignored, entries before it act.
//
// STAGED (k_ov_resolve): the replay of a component is a chain of dependent loads — state of the member, its record, its row, the
// state and record of every entry — repeated for every member at every recursion level; the kernel lasts as long as the longest
// chain.  So a component of up to 32 members is first staged into shared memory with all loads of one kind issued side by side:
// the members (both positions, state), then every row entry of every member at once (lanes over the concatenated rows), of which
// only the entries that can be within rcut in ANY new/old combination are kept, in (member, row position) order.  The levels are
// then replayed from shared memory.  Dropping the other entries is exact: the replay only ever tests those combinations.
// Entries that are members read their state from the staged copy; metal partners never change; a mobile partner that is not a
// member (not in ref: F atoms) keeps the global read/write of the unstaged form.
constexpr int OVS_ENT = 64;                                   // staged entries per component; more -> unstaged replay
struct OvStage {
  double pn[32][3], po[32][3];                                // members: moved position, previous position
  double ep[OVS_ENT][3];                                      // entries: partner's moved position
  int st[32], slot[32], incl[32], rb[32], qm[32], eb[33];
  int ea2[OVS_ENT], eloc[OVS_ENT], et2[OVS_ENT], em[OVS_ENT];
};
template <bool STAGED>
__device__ __forceinline__ void p_ov_resolve(const double4 *__restrict__ posm, const double *__restrict__ old_cg, const RowHead *__restrict__ rh,
                             const int *__restrict__ cols, const unsigned char *__restrict__ bq,
                             const unsigned int *__restrict__ lay, int *__restrict__ ovst,
                             const int *__restrict__ roots, const int *__restrict__ ov_head, const int *__restrict__ ov_next,
                             int *__restrict__ members, const int *__restrict__ uid, const OvRp rp_uovl,
                             DevScal *__restrict__ sc, Geo g, Phys ph, unsigned int step, int guard_pass, OvStage *stg_all) {
  if (step == STEP_FROM_DEVICE) step = sc->istep;
  const unsigned int FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int r_end = sc->n_roots;
  const double rcut = sqrt(g.rcut2), z0 = sc->z0;
  volatile int *vst = (volatile int *)ovst;
  for (int r = wid; r < r_end; r += nw) {
    const int root = roots[r];
    // ---- collect the members (list walk, uniform) and order them by creation rank ----
    int mine = -1, cnt = 0;
    for (int cur = ov_head[root]; cur >= 0; cur = ov_next[cur]) { if (cnt == lane) mine = cur; ++cnt; }
    const bool big = cnt > 32;
    int sorted = -1, base = 0;
    if (!big) {
      const int myuid = mine >= 0 ? uid[mine] : 0x7fffffff;
      int rank = 0;
      for (int i = 0; i < cnt; ++i) rank += (__shfl_sync(FULL, myuid, i) < myuid) ? 1 : 0;
      for (int i = 0; i < cnt; ++i) { const int rk = __shfl_sync(FULL, rank, i), mv = __shfl_sync(FULL, mine, i); if (rk == lane) sorted = mv; }
    } else {                                               // rare: spill to the scratch array, one lane sorts
      if (lane == 0) {
        base = atomicAdd(&sc->member_cursor, cnt);
        int i = 0;
        for (int cur = ov_head[root]; cur >= 0; cur = ov_next[cur]) members[base + i++] = cur;
        ov_sort_members(members, uid, base, base + cnt);
      }
      base = __shfl_sync(FULL, base, 0);
      __syncwarp();
    }
    OvAcc acc = {0, 0, 0, 0};
    long long later = 0;
    int pass = 0;
    bool staged = false;
    if (STAGED && !big) {
      OvStage &S = stg_all[threadIdx.x >> 5];
      __syncwarp();
      // ---- members, lane = rank in creation order ----
      int rl = 0;
      bool ghost = false;
      if (sorted >= 0) {
        const double4 rec = ld_rec_nc(&posm[sorted]);
        const RowMeta rm = rh_meta(&rh[sorted]);
        S.st[lane] = vst[sorted];
        S.pn[lane][0] = rec.x; S.pn[lane][1] = rec.y; S.pn[lane][2] = rec.z;
        S.po[lane][0] = old_cg[3 * sorted]; S.po[lane][1] = old_cg[3 * sorted + 1]; S.po[lane][2] = old_cg[3 * sorted + 2];
        S.slot[lane] = sorted; S.rb[lane] = rm.start;
        S.qm[lane] = skip_qtab(g, lay, rec.z, 1);
        ghost = (meta_of(rec) & MF_GHOST) != 0;
        rl = rm.len;
      }
      int incl = rl;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
      S.incl[lane] = incl;
      const int tot = __shfl_sync(FULL, incl, 31);
      const bool any_ghost = __ballot_sync(FULL, ghost) != 0u;   // slab mode: a ghost member has no row here -> unstaged form
      __syncwarp();
      if (!any_ghost) {
        // ---- every row entry of every member, lanes over the concatenated rows ----
        int nent = 0;
        for (int f0 = 0; f0 < tot; f0 += 32) {
          const int f = f0 + lane;
          bool rel = false;
          int m = 0, a2 = -1, loc2 = -1, t2 = 0;
          double p2[3] = {0.0, 0.0, 0.0};
          if (f < tot) {
            for (int i = 0; i < cnt; ++i) m += S.incl[i] <= f ? 1 : 0;
            const int jj = f - (m ? S.incl[m - 1] : 0), rb = S.rb[m];
            if ((int)__ldg(&bq[rb + jj]) <= S.qm[m]) {
              a2 = __ldg(&cols[rb + jj]);
              const int st2 = vst[a2];
              t2 = (st2 >> OV_TSHIFT) & 3;
              if (t2) {
                const double4 r2 = ld_rec_nc(&posm[a2]);
                p2[0] = r2.x; p2[1] = r2.y; p2[2] = r2.z;
                for (int i = 0; i < cnt; ++i) if (S.slot[i] == a2) loc2 = i;
                const double *n1 = S.pn[m], *o1 = S.po[m];
                rel = !(dist2_idnint(g, n1[0], n1[1], n1[2], p2[0], p2[1], p2[2]) > g.rcut2) ||
                      !(dist2_idnint(g, o1[0], o1[1], o1[2], p2[0], p2[1], p2[2]) > g.rcut2);
                if (t2 != 2) {                                     // a partner that can be moved back: its previous position counts too
                  double o2[3];
                  if (loc2 >= 0) { o2[0] = S.po[loc2][0]; o2[1] = S.po[loc2][1]; o2[2] = S.po[loc2][2]; }
                  else { o2[0] = old_cg[3 * a2]; o2[1] = old_cg[3 * a2 + 1]; o2[2] = old_cg[3 * a2 + 2]; loc2 = -2; }
                  rel = rel || !(dist2_idnint(g, n1[0], n1[1], n1[2], o2[0], o2[1], o2[2]) > g.rcut2) ||
                        !(dist2_idnint(g, o1[0], o1[1], o1[2], o2[0], o2[1], o2[2]) > g.rcut2);
                }
              }
            }
          }
          const unsigned int mk = __ballot_sync(FULL, rel);
          const int at = nent + __popc(mk & ((1u << lane) - 1u));
          if (rel && at < OVS_ENT) {
            S.ep[at][0] = p2[0]; S.ep[at][1] = p2[1]; S.ep[at][2] = p2[2];
            S.ea2[at] = a2; S.eloc[at] = loc2; S.et2[at] = t2; S.em[at] = m;
          }
          nent += __popc(mk);
        }
        __syncwarp();
        if (nent <= OVS_ENT) {
          staged = true;
          int eb = 0;
          for (int k = 0; k < nent; ++k) eb += S.em[k] < lane ? 1 : 0;
          S.eb[lane] = eb;
          if (lane == 0) S.eb[32] = nent;
          __syncwarp();
          volatile int *sst = (volatile int *)S.st;
          // ---- the recursion levels, from shared memory ----
          for (;; ++pass) {
            const int guard = (guard_pass > 0 && pass >= guard_pass) ? 1 : 0;
            const long long ch0 = acc.ch;
            bool again = false;
            for (int i = 0; i < cnt; ++i) {
              int st1 = sst[i];
              if (st1 & OV_SKIP) continue;
              if (!((st1 >> OV_TSHIFT) & 3)) continue;
              st1 |= OV_SKIP;
              const double *q1 = (st1 & OV_MOVED) ? S.po[i] : S.pn[i];
              const double q1x = q1[0], q1y = q1[1], q1z = q1[2];
              const bool at_old1 = q1x == S.po[i][0] && q1y == S.po[i][1] && q1z == S.po[i][2];
              const int a1 = S.slot[i];
              const int e0 = S.eb[i], e1 = i + 1 < 32 ? S.eb[i + 1] : nent;
              bool stop = false;
              for (int j0 = e0; j0 < e1 && !stop; j0 += 32) {
                const int k = j0 + lane;
                bool in = false;
                int a2 = -1, st2 = 0, t2 = 0, loc2 = -1;
                double q2[3] = {0.0, 0.0, 0.0};
                bool at_old2 = false;
                if (k < e1) {
                  a2 = S.ea2[k]; loc2 = S.eloc[k];
                  if (loc2 >= 0) {
                    st2 = sst[loc2]; t2 = (st2 >> OV_TSHIFT) & 3;
                    const double *q = (st2 & OV_MOVED) ? S.po[loc2] : S.pn[loc2];
                    q2[0] = q[0]; q2[1] = q[1]; q2[2] = q[2];
                    at_old2 = q2[0] == S.po[loc2][0] && q2[1] == S.po[loc2][1] && q2[2] == S.po[loc2][2];
                  } else if (loc2 == -1) {
                    t2 = S.et2[k]; q2[0] = S.ep[k][0]; q2[1] = S.ep[k][1]; q2[2] = S.ep[k][2];
                  } else {
                    st2 = vst[a2]; t2 = (st2 >> OV_TSHIFT) & 3;
                    if (t2) {
                      ov_pos(posm, old_cg, a2, st2, q2);
                      at_old2 = q2[0] == old_cg[3 * a2] && q2[1] == old_cg[3 * a2 + 1] && q2[2] == old_cg[3 * a2 + 2];
                    }
                  }
                  if (t2) in = !(dist2_idnint(g, q1x, q1y, q1z, q2[0], q2[1], q2[2]) > g.rcut2);
                }
                const unsigned int cgm = __ballot_sync(FULL, in && t2 == 2);
                const int first = cgm ? __ffs(cgm) - 1 : 32;
                const bool mob = in && t2 != 2 && lane < first;
                const bool unsolv = mob && (ph.piston || guard) && at_old2 && at_old1;     // dana.F90:920-927
                const bool mv = mob && !unsolv;
                if (mv) {
                  const int nst = (st2 | OV_MOVED | OV_ZERO) & ~OV_SKIP;
                  if (loc2 >= 0) sst[loc2] = nst; else ovst[a2] = nst;
                }
                acc.ch3 += __popc(__ballot_sync(FULL, unsolv));
                const int nm = __popc(__ballot_sync(FULL, mv));
                acc.ch += nm;
                if (nm) again = true;
                if (cgm) {
                  acc.tr++;
                  double ne;
                  if (ph.rng_mode == 1) ne = ov_replay_draw(rp_uovl, a1, lane == 0, sc, ph.prob);
                  else if (ph.rng_mode == 2) ne = 0.0;
                  else { Philox rr; rr.run(ph.seed, (unsigned int)uid[a1], step, RS_OVERLAP, (unsigned int)pass); ne = rr.u01(0); }
                  if (ne < ph.prob) {
                    acc.de++; st1 = (st1 & ~(3 << OV_TSHIFT)) | (3 << OV_TSHIFT);
                    if (q1z > z0 && lane == 0) atomicCAS(&sc->err, 0, DML_E_SUPERO_Z0);
                  } else { st1 |= OV_MOVED; st1 &= ~OV_SKIP; }
                  stop = true;
                }
              }
              __syncwarp();
              if (lane == 0) sst[i] = st1;
              __syncwarp();
            }
            if (pass >= 1) later += acc.ch - ch0;
            if (!again) break;
          }
          if (sorted >= 0) ovst[sorted] = sst[lane];
        }
      }
      __syncwarp();
    }
    if (!staged)
    for (;; ++pass) {
      const int guard = (guard_pass > 0 && pass >= guard_pass) ? 1 : 0;
      const long long ch0 = acc.ch;
      bool again = false;
      for (int i = 0; i < cnt; ++i) {
        const int a1 = big ? ((volatile int *)members)[base + i] : __shfl_sync(FULL, sorted, i);
        int st1 = vst[a1];
        if (st1 & OV_SKIP) continue;
        if (!((st1 >> OV_TSHIFT) & 3)) continue;
        st1 |= OV_SKIP;
        double q1[3]; ov_pos(posm, old_cg, a1, st1, q1);
        const double4 rec1 = ld_rec_nc(&posm[a1]);
        if (meta_of(rec1) & MF_GHOST) {
          // Slab mode: a1 is owned by the neighbouring slab and has no row here.  Its visit (it comes in creation-rank order
          // like everybody else) acts on the owned members of the component that are in range — the same pair decisions its
          // owner takes from the same inputs; its contacts with metal are the owner's business.
          for (int i0 = 0; i0 < cnt; i0 += 32) {
            const int mi = i0 + lane;
            const int a2 = mi < cnt ? (big ? ((volatile int *)members)[base + mi] : (i0 == 0 ? sorted : -1)) : -1;
            bool mv = false, unsolv = false;
            if (a2 >= 0 && a2 != a1 && !(meta_of(ld_rec_nc(&posm[a2])) & MF_GHOST)) {
              const int st2 = vst[a2];
              if ((st2 >> OV_TSHIFT) & 3) {
                double q2[3]; ov_pos(posm, old_cg, a2, st2, q2);
                if (!(dist2_idnint(g, q1[0], q1[1], q1[2], q2[0], q2[1], q2[2]) > g.rcut2)) {
                  if ((ph.piston || guard) &&
                      q2[0] == old_cg[3 * a2] && q2[1] == old_cg[3 * a2 + 1] && q2[2] == old_cg[3 * a2 + 2] &&
                      q1[0] == old_cg[3 * a1] && q1[1] == old_cg[3 * a1 + 1] && q1[2] == old_cg[3 * a1 + 2]) unsolv = true;
                  else { mv = true; ovst[a2] = (st2 | OV_MOVED | OV_ZERO) & ~OV_SKIP; }
                }
              }
            }
            acc.ch3 += __popc(__ballot_sync(FULL, unsolv));
            const int nm = __popc(__ballot_sync(FULL, mv));
            acc.ch += nm;
            if (nm) again = true;
          }
          __syncwarp();
          if (lane == 0) ovst[a1] = st1;
          __syncwarp();
          continue;
        }
        const RowMeta rm = rh_meta(&rh[a1]);
        const int rb = rm.start, rl = rm.len;
        // the build-distance skip of k_ov_detect (same bound, same record): an entry it skipped cannot be within rcut in any
        // new/old combination, so skipping it here changes nothing
        const int qmax = skip_qtab(g, lay, rec1.z, 1);
        bool stop = false;
        for (int j0 = 0; j0 < rl && !stop; j0 += 32) {
          const int jj = j0 + lane;
          bool in = false;
          int a2 = -1, st2 = 0, t2 = 0;
          double q2[3] = {0.0, 0.0, 0.0};
          if (jj < rl && (int)__ldg(&bq[rb + jj]) <= qmax) {
            a2 = __ldg(&cols[rb + jj]);
            st2 = vst[a2];
            t2 = (st2 >> OV_TSHIFT) & 3;                            // 0 = limbo (dana.F90:881-883)
            if (t2) { ov_pos(posm, old_cg, a2, st2, q2); in = !(dist2_idnint(g, q1[0], q1[1], q1[2], q2[0], q2[1], q2[2]) > g.rcut2); }
          }
          const unsigned int cgm = __ballot_sync(FULL, in && t2 == 2);
          const int first = cgm ? __ffs(cgm) - 1 : 32;              // first contact with metal: the reference leaves the row there
          const bool mob = in && t2 != 2 && lane < first;
          bool unsolv = false;
          if (mob && (ph.piston || guard)) {                        // unsolvable pair guard (dana.F90:920-927)
            if (q2[0] == old_cg[3 * a2] && q2[1] == old_cg[3 * a2 + 1] && q2[2] == old_cg[3 * a2 + 2] &&
                q1[0] == old_cg[3 * a1] && q1[1] == old_cg[3 * a1 + 1] && q1[2] == old_cg[3 * a1 + 2]) unsolv = true;
          }
          const bool mv = mob && !unsolv;
          // o2 goes back to its previous position; velocities are zeroed when the state is applied
          if (mv) ovst[a2] = (st2 | OV_MOVED | OV_ZERO) & ~OV_SKIP;
          acc.ch3 += __popc(__ballot_sync(FULL, mob && unsolv));
          const int nm = __popc(__ballot_sync(FULL, mv));
          acc.ch += nm;
          if (nm) again = true;
          if (cgm) {                                                 // deposition attempt (uniform over the warp)
            acc.tr++;
            double ne;
            if (ph.rng_mode == 1) ne = ov_replay_draw(rp_uovl, a1, lane == 0, sc, ph.prob);
            else if (ph.rng_mode == 2) ne = 0.0;
            else { Philox rr; rr.run(ph.seed, (unsigned int)uid[a1], step, RS_OVERLAP, (unsigned int)pass); ne = rr.u01(0); }
            if (ne < ph.prob) {
              acc.de++; st1 = (st1 & ~(3 << OV_TSHIFT)) | (3 << OV_TSHIFT);
              if (q1[2] > z0 && lane == 0) atomicCAS(&sc->err, 0, DML_E_SUPERO_Z0);
            } else { st1 |= OV_MOVED; st1 &= ~OV_SKIP; }
            stop = true;
          }
        }
        __syncwarp();
        if (lane == 0) ovst[a1] = st1;
        __syncwarp();
      }
      if (pass >= 1) later += acc.ch - ch0;
      if (!again) break;
    }
    if (lane == 0) {
      ov_flush(acc, sc);
      if (later) atomicAdd((unsigned long long *)&sc->ch_later, (unsigned long long)later);
      atomicMax(&sc->any_active, pass + 1);                        // deepest recursion level of this call
    }
  }
}
template <bool STAGED>
__global__ void __launch_bounds__(128) k_ov_resolve(const double4 *__restrict__ posm, const double *__restrict__ old_cg, const RowHead *__restrict__ rh,
                             const int *__restrict__ cols, const unsigned char *__restrict__ bq,
                             const unsigned int *__restrict__ lay, int *__restrict__ ovst,
                             const int *__restrict__ roots, const int *__restrict__ ov_head, const int *__restrict__ ov_next,
                             int *__restrict__ members, const int *__restrict__ uid, const OvRp rp_uovl,
                             DevScal *__restrict__ sc, Geo g, Phys ph, unsigned int step, int guard_pass) {
  __shared__ OvStage stg[STAGED ? 4 : 1];
  p_ov_resolve<STAGED>(posm, old_cg, rh, cols, bq, lay, ovst, roots, ov_head, ov_next, members, uid, rp_uovl, sc, g, ph, step, guard_pass, stg);
}
// write the resolved state back: positions, zeroed vel/acel of moved-back atoms, skip flags and new F atoms
__device__ __forceinline__ void d_ov_apply(double4 *__restrict__ posm, double *__restrict__ vel, double *__restrict__ acel,
                                           const double *__restrict__ old_cg, const int *__restrict__ ovst, int s) {
  double4 p = ld_rec(&posm[s]);
  long long m = meta_of(p);
  if (!(m & MF_REF)) return;
  int st = ovst[s];
  long long nm = m;
  if (st & OV_INVOLVED) {
    nm = (m & ~(MF_TYPE | MF_SKIP)) | (long long)((st >> OV_TSHIFT) & 3) | ((st & OV_SKIP) ? MF_SKIP : 0);
    if (st & OV_MOVED) { p.x = old_cg[3 * s]; p.y = old_cg[3 * s + 1]; p.z = old_cg[3 * s + 2]; }
    if (st & OV_ZERO) {
      vel[3 * s] = 0.0; vel[3 * s + 1] = 0.0; vel[3 * s + 2] = 0.0;
      acel[3 * s] = 0.0; acel[3 * s + 1] = 0.0; acel[3 * s + 2] = 0.0;
    }
  } else nm = m | MF_SKIP;                                     // processed in the first pass, nothing in range
  if (nm != m || (st & OV_MOVED)) { p.w = meta_as_double(nm); st_rec(&posm[s], p); }
}
__device__ __forceinline__ void p_ov_apply(double4 *__restrict__ posm, double *__restrict__ vel, double *__restrict__ acel,
                           const double *__restrict__ old_cg, const int *__restrict__ ovst, DevScal *__restrict__ sc, int n) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {                                                // choques2=max(choques2,choques-i), dana.F90:939-941
    if (sc->ch_later > sc->choques2) sc->choques2 = sc->ch_later;
    sc->overlap_passes += sc->any_active > 0 ? sc->any_active : 1;
  }
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) d_ov_apply(posm, vel, acel, old_cg, ovst, s);
}
__global__ void k_ov_apply(double4 *__restrict__ posm, double *__restrict__ vel, double *__restrict__ acel,
                           const double *__restrict__ old_cg, const int *__restrict__ ovst, DevScal *__restrict__ sc, int n) { p_ov_apply(posm, vel, acel, old_cg, ovst, sc, n); }

// ================================================================================================
// K8  F -> CG promotion (dana.F90:228-236), calc_rho (521-549), maxz (776-794)
// ================================================================================================
__global__ void k_promote(double4 *__restrict__ posm, DevScal *__restrict__ sc, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  int dref = 0, dg = 0;
  if (s < n) {
    double4 p = ld_rec(&posm[s]);
    long long m = meta_of(p);
    if ((m & MF_REF) && (m & MF_TYPE) == 3) {
      dref = 1; dg = (m & MF_GCMC) ? 1 : 0;
      m = (m & ~(MF_TYPE | MF_REF | MF_GCMC)) | 2;
      p.w = meta_as_double(m); st_rec(&posm[s], p);
    }
  }
  dref = __reduce_add_sync(0xffffffffu, dref); dg = __reduce_add_sync(0xffffffffu, dg);
  if ((threadIdx.x & 31) == 0) { if (dref) atomicSub(&sc->nat_ref, dref); if (dg) atomicSub(&sc->nat_gcmc, dg); }
}
// promotion + msd bookkeeping + calc_rho in one pass (reservoirs 1 and 2, where nothing runs between them)
__global__ void __launch_bounds__(TPB) k_promote_rho(double4 *__restrict__ posm, DevScal *__restrict__ sc, double area, int use_z1, int n) {
  double z0 = sc->z0, zl = use_z1 ? sc->z1 : sc->zmax;
  int c = 0, dref = 0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    double4 p = ld_rec(&posm[s]);
    long long m = meta_of(p);
    if ((m & MF_REF) && (m & MF_TYPE) == 3) {
      dref++;
      m = (m & ~(MF_TYPE | MF_REF | MF_GCMC)) | 2;
      p.w = meta_as_double(m); st_rec(&posm[s], p);
    }
    if ((m & MF_TYPE) && p.z > z0 && p.z < zl) c++;
  }
  c = __reduce_add_sync(0xffffffffu, c); dref = __reduce_add_sync(0xffffffffu, dref);
  __shared__ int last;
  if ((threadIdx.x & 31) == 0) { if (c) atomicAdd(&sc->rho_count, c); if (dref) atomicAdd(&sc->n_involved, dref); }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&sc->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    int gct = *((volatile int *)&sc->rho_count), dr = *((volatile int *)&sc->n_involved);
    sc->msd_t = sc->msd_t / sc->nat_ref;                 // dana.F90:201-202 (before the promotion loop)
    sc->msd_max = fmax(sc->msd_max, sc->msd_t);
    sc->nat_ref -= dr;
    double vol = area * (zl - z0);
    sc->rho = gct / vol;
    sc->rho_count = 0; sc->n_involved = 0; sc->ticket = 0;
    sc->step_disp_bits = 0u;                             // end of the step: the next integrator call records its own largest move
  }
}
__global__ void k_calc_rho(const double4 *__restrict__ posm, DevScal *__restrict__ sc, double area, int use_z1, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  double z0 = sc->z0, zl = use_z1 ? sc->z1 : sc->zmax;
  int c = 0;
  if (s < n) { double4 p = ld_rec_nc(&posm[s]); if ((meta_of(p) & MF_TYPE) && p.z > z0 && p.z < zl) c = 1; }
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int last;
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sc->rho_count, c);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&sc->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    int gct = *((volatile int *)&sc->rho_count);
    double vol = area * (zl - z0);                       // box(1)*box(2)*(z-z0)
    sc->rho = gct / vol;
    sc->rho_count = 0; sc->ticket = 0;
    sc->step_disp_bits = 0u;
  }
}
__global__ void k_maxz(double4 *__restrict__ posm, DevScal *__restrict__ sc, double h_over_tau, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  double rho = sc->rho, rho0 = sc->rho0, z0 = sc->z0;
  double lohi = (h_over_tau * ((rho0 - rho) / rho));
  if (s < n) {
    double4 p = ld_rec(&posm[s]);
    if ((meta_of(p) & MF_TYPE) && p.z > z0) { p.z = p.z - lohi * (p.z - z0); st_rec(&posm[s], p); }
  }
  if (s == 0) {
    sc->maxz_disp += fabs(lohi) * fmax(sc->zmax - z0, 0.0) * 1.01;
    sc->maxz_fac += fabs(lohi) * 1.01;   // |z shift| of any atom below the ceiling
    sc->pist_P *= (1.0 - lohi);
    sc->zmax = sc->zmax - lohi * (sc->zmax - z0);
  }
}

// ================================================================================================
// pack / unpack between the caller's [n][3] arrays and the device records
// ================================================================================================
__global__ void k_pack(double4 *__restrict__ posm, const double *__restrict__ pos, const int *__restrict__ z,
                       const int *__restrict__ flags, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  long long m = 0;
  int zz = z[s], f = flags[s];
  if (f & 8) m = MF_LIMBO;
  else if (zz >= 1 && zz <= 3) m = with_disp(zz | ((f & 1) ? MF_REF : 0) | ((f & 2) ? MF_GCMC : 0) | ((f & 4) ? MF_SKIP : 0), DISP_INF);
  double4 p = {pos[3 * s], pos[3 * s + 1], pos[3 * s + 2], meta_as_double(m)};
  st_rec(&posm[s], p);
}
// new coordinates for the same atoms (dml_upload_positions): element, membership and flags stay; the displacement bound is unknown
__global__ void k_repos(double4 *__restrict__ posm, const double *__restrict__ pos, int n, DevScal *__restrict__ sc) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  // the caller may have moved anything anywhere: no gather skipping until the next test_update has measured the displacements
  if (s == 0) atomicMax(&sc->step_disp_bits, 0x7f800000u);
  if (s >= n) return;
  double4 p = ld_rec(&posm[s]);
  const long long m = meta_of(p);
  if (!(m & MF_TYPE)) return;
  p.x = pos[3 * s]; p.y = pos[3 * s + 1]; p.z = pos[3 * s + 2]; p.w = meta_as_double(with_disp(m, DISP_INF));
  st_rec(&posm[s], p);
}
__global__ void k_unpack(const double4 *__restrict__ posm, double *__restrict__ pos, int *__restrict__ z, int *__restrict__ flags, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double4 p = ld_rec_nc(&posm[s]);
  long long m = meta_of(p);
  if (pos) { pos[3 * s] = p.x; pos[3 * s + 1] = p.y; pos[3 * s + 2] = p.z; }
  if (z) z[s] = (int)(m & MF_TYPE);
  if (flags) flags[s] = ((m & MF_REF) ? 1 : 0) | ((m & MF_GCMC) ? 2 : 0) | ((m & MF_SKIP) ? 4 : 0) | ((m & MF_LIMBO) ? 8 : 0);
}
__global__ void k_unpack_fe(const double4 *__restrict__ fe, double *__restrict__ force, double *__restrict__ epot, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double4 f = ld_rec_nc(&fe[s]);
  force[3 * s] = f.x; force[3 * s + 1] = f.y; force[3 * s + 2] = f.z; epot[s] = f.w;
}
__global__ void k_count_members(const double4 *__restrict__ posm, DevScal *__restrict__ sc, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  int a = 0, r = 0, gq = 0, l = 0;
  if (s < n) { long long m = meta_of(ld_rec_nc(&posm[s])); a = (m & MF_TYPE) ? 1 : 0; r = (m & MF_REF) ? 1 : 0; gq = (m & MF_GCMC) ? 1 : 0; l = (m & MF_LIMBO) ? 1 : 0; }
  a = __reduce_add_sync(0xffffffffu, a); r = __reduce_add_sync(0xffffffffu, r); gq = __reduce_add_sync(0xffffffffu, gq); l = __reduce_add_sync(0xffffffffu, l);
  if ((threadIdx.x & 31) == 0) { if (a) atomicAdd(&sc->nat_sys, a); if (r) atomicAdd(&sc->nat_ref, r); if (gq) atomicAdd(&sc->nat_gcmc, gq); if (l) atomicAdd(&sc->nlimbo, l); }
}
// dml_upload without a gcmc group: member counts (k_count_members), default creation ranks / b indices, occupancy of the b index
// (Groups.F90:1083-1093) and the running maxima hs%b%amax and next creation rank, in one pass over the uploaded slots.
__global__ void k_upload_book(const double4 *__restrict__ posm, int *__restrict__ uid, int *__restrict__ slot_b, int *__restrict__ b_occ,
                              DevScal *__restrict__ sc, int n, int cap, int gen_uid, int gen_sb) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  int a = 0, r = 0, l = 0, amax = 0, nuid = 0;
  if (s < n) {
    const long long m = meta_of(ld_rec_nc(&posm[s]));
    a = (m & MF_TYPE) ? 1 : 0; r = (m & MF_REF) ? 1 : 0; l = (m & MF_LIMBO) ? 1 : 0;
    if (gen_uid) uid[s] = s;
    if (gen_sb) slot_b[s] = s;
    nuid = (gen_uid ? s : uid[s]) + 1;
    if (a) {
      const int sb = gen_sb ? s : slot_b[s];
      if (sb < 0 || sb >= cap) atomicCAS(&sc->err, 0, DML_E_CAPACITY);
      else { b_occ[sb] = 1; amax = sb + 1; }
    }
  }
  a = __reduce_add_sync(0xffffffffu, a); r = __reduce_add_sync(0xffffffffu, r); l = __reduce_add_sync(0xffffffffu, l);
  amax = __reduce_max_sync(0xffffffffu, amax); nuid = __reduce_max_sync(0xffffffffu, nuid);
  if ((threadIdx.x & 31) == 0) {
    if (a) atomicAdd(&sc->nat_sys, a);
    if (r) atomicAdd(&sc->nat_ref, r);
    if (l) atomicAdd(&sc->nlimbo, l);
    if (amax) atomicMax(&sc->b_amax, amax);
    if (nuid) atomicMax(&sc->next_uid, nuid);
  }
}
// chain position of every particle inside its cell (inspection only)
__global__ void k_msd_book(DevScal *__restrict__ sc) {      // dana.F90:201-202
  sc->msd_t = sc->msd_t / sc->nat_ref;
  sc->msd_max = fmax(sc->msd_max, sc->msd_t);
}
__global__ void k_chain_pos(const int *__restrict__ cell_start, const int *__restrict__ sorted_slot, int *__restrict__ chain_pos, int ncell) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  for (int i = cell_start[c]; i < cell_start[c + 1]; ++i) chain_pos[sorted_slot[i]] = i - cell_start[c];
}

} // namespace dml
