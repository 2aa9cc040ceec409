// dml_kernels.cuh — the sm_100a kernels of the dana hot path.  One thread per particle slot unless noted.
// Reference lines cited per kernel are relative to the reference tree (src/...).
#pragma once
#include "dml_device.cuh"

namespace dml {

constexpr int TPB = 256;

// ================================================================================================
// Generic exclusive scan of int32 (3 small kernels; 1024 items per block)
// ================================================================================================
__global__ void k_scan_local(const int *__restrict__ in, int *__restrict__ out, int *__restrict__ sums, int n) {
  __shared__ int wsum[8];
  int base = blockIdx.x * 1024 + threadIdx.x * 4;
  int v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = (base + i < n) ? in[base + i] : 0;
  int t = v[0] + v[1] + v[2] + v[3];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) wsum[w] = x;
  __syncthreads();
  if (w == 0) {
    int s = lane < 8 ? wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    if (lane < 8) wsum[lane] = s;
  }
  __syncthreads();
  int excl = x - t + (w ? wsum[w - 1] : 0);
  int run = excl;
#pragma unroll
  for (int i = 0; i < 4; ++i) { if (base + i < n) out[base + i] = run; run += v[i]; }
  if (threadIdx.x == TPB - 1) sums[blockIdx.x] = run;
}
__global__ void k_scan_sums(int *__restrict__ sums, int nb, int *__restrict__ total_out, int *__restrict__ total_out2) {
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    int t = i < nb ? sums[i] : 0, x = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      int s = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
      wsum[lane] = s;
    }
    __syncthreads();
    int excl = x - t + (w ? wsum[w - 1] : 0) + carry;
    if (i < nb) sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + t;
    __syncthreads();
  }
  if (threadIdx.x == 0) { if (total_out) *total_out = carry; if (total_out2) *total_out2 = carry; }
}
__global__ void k_scan_add(int *__restrict__ out, const int *__restrict__ sums, int n) {
  int i = blockIdx.x * 1024 + threadIdx.x * 4;
  int add = sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < 4; ++k) if (i + k < n) out[i + k] += add;
}

// ================================================================================================
// K1  do_pbc + cell binning + displacement top-2        (test_update, Neighbor.F90:668-713;
//     do_pbc Groups.F90:1440-1467; cgroup_sort_atom index math Cells.F90:281-302; inq_dispmax 635-666)
// ================================================================================================
__global__ void k_pbc_bin(double4 *__restrict__ posm, double *__restrict__ pos_old, int *__restrict__ cell_of,
                          int *__restrict__ cell_cnt, double *__restrict__ part, DevScal *__restrict__ sc, Geo g, int n, int do_bin) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  double a1 = -1.0, a2 = -1.0;
  if (s < n) {
    double4 p = ld_rec(&posm[s]);
    long long m = meta_of(p);
    if (m & MF_TYPE) {
      double po[3] = {pos_old[3 * s], pos_old[3 * s + 1], pos_old[3 * s + 2]};
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
      if (do_bin) {
        int cx, cy, cz;
        if (cell_index(g, q[0], q[1], q[2], cx, cy, cz)) {
          int lin = cell_lin(g, cx, cy, cz);
          cell_of[s] = lin;
          atomicAdd(&cell_cnt[lin], 1);
          if (cx == 0 || cy == 0 || cz == 0 || cx == g.nc[0] + 1 || cy == g.nc[1] + 1 || cz == g.nc[2] + 1) sc->halo_flag = 1;
        } else { cell_of[s] = -1; atomicCAS(&sc->err, 0, DML_E_OUT_OF_TESS); }
      }
      double vx = q[0] - po[0], vy = q[1] - po[1], vz = q[2] - po[2];
      a1 = (vx * vx + vy * vy) + vz * vz;
    } else if (do_bin) cell_of[s] = -1;
  }
  // block top-2
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double b1 = __shfl_xor_sync(0xffffffffu, a1, o), b2 = __shfl_xor_sync(0xffffffffu, a2, o);
    top2_merge(a1, a2, b1, b2);
  }
  __shared__ double s1[TPB / 32], s2[TPB / 32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { s1[w] = a1; s2[w] = a2; }
  __syncthreads();
  if (w == 0) {
    a1 = lane < TPB / 32 ? s1[lane] : -1.0; a2 = lane < TPB / 32 ? s2[lane] : -1.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      double b1 = __shfl_xor_sync(0xffffffffu, a1, o), b2 = __shfl_xor_sync(0xffffffffu, a2, o);
      top2_merge(a1, a2, b1, b2);
    }
    if (lane == 0) { part[2 * blockIdx.x] = a1; part[2 * blockIdx.x + 1] = a2; }
  }
}
__global__ void k_top2_final(const double *__restrict__ part, int nb, DevScal *__restrict__ sc, int listed, double nb_dcut) {
  double a1 = 1e-16, a2 = 1e-16;     // Neighbor.F90:643-644
  for (int i = threadIdx.x; i < nb; i += blockDim.x) top2_merge(a1, a2, part[2 * i], part[2 * i + 1]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double b1 = __shfl_xor_sync(0xffffffffu, a1, o), b2 = __shfl_xor_sync(0xffffffffu, a2, o);
    top2_merge(a1, a2, b1, b2);
  }
  __shared__ double s1[32], s2[32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { s1[w] = a1; s2[w] = a2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) top2_merge(a1, a2, s1[i], s2[i]);
    sc->d1 = a1; sc->d2 = a2;
    sc->need_rebuild = (!listed) || (sqrt(a1) + sqrt(a2) > nb_dcut);   // Neighbor.F90:697-710
  }
}

// ================================================================================================
// K2  counting sort into cells.  In-cell order must be DESCENDING b-slot (head insertion of ascending
//     slots, Cells.F90:267-302): scatter with an atomic cursor, then order every cell segment.
//     k_scatter also performs update()'s pos_old=pos (Neighbor.F90:620-624) and igroup_clean
//     (Groups.F90:1036-1053) because both happen exactly when the list is rebuilt.
// ================================================================================================
__global__ void k_scatter(double4 *__restrict__ posm, double *__restrict__ pos_old, const int *__restrict__ cell_of,
                          const int *__restrict__ cell_start, int *__restrict__ cell_cur, int *__restrict__ sorted_slot,
                          int n, int snapshot) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double4 p = ld_rec(&posm[s]);
  long long m = meta_of(p);
  if (m & MF_TYPE) {
    int lin = cell_of[s];
    if (lin >= 0) { int idx = cell_start[lin] + atomicAdd(&cell_cur[lin], 1); sorted_slot[idx] = s; }
    if (snapshot) { pos_old[3 * s] = p.x; pos_old[3 * s + 1] = p.y; pos_old[3 * s + 2] = p.z; }
  } else if (snapshot && (m & MF_LIMBO)) { p.w = meta_as_double(0); st_rec(&posm[s], p); }
}
// one thread per cell: insertion sort of the segment by descending slot_b, then gather the records
__global__ void k_cell_order(const double4 *__restrict__ posm, const int *__restrict__ slot_b, const int *__restrict__ cell_start,
                             int *__restrict__ sorted_slot, double4 *__restrict__ sorted_posm, int ncell) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  int b = cell_start[c], e = cell_start[c + 1];
  for (int i = b + 1; i < e; ++i) {
    int s = sorted_slot[i], key = slot_b[s], j = i - 1;
    while (j >= b) { int sj = sorted_slot[j]; if (slot_b[sj] >= key) break; sorted_slot[j + 1] = sj; --j; }
    sorted_slot[j + 1] = s;
  }
  for (int i = b; i < e; ++i) { double4 p = ld_rec(&posm[sorted_slot[i]]); st_rec(&sorted_posm[i], p); }
}

// ================================================================================================
// K3  Verlet rows over linked cells      (ngroup_cells, Neighbor.F90:465-548; cell_pbc Cells.F90:378-404;
//     vdistance Groups.F90:995-1016).  One thread per cell-sorted particle; rows are written in the
//     reference's order (stencil order x chain order) because one thread walks them sequentially.
//     FILL=false counts, FILL=true writes cols[row_start[slot] ...].
// ================================================================================================
template <bool FILL>
__global__ void k_rows(const double4 *__restrict__ sorted_posm, const int *__restrict__ sorted_slot, const int *__restrict__ cell_of,
                       const int *__restrict__ cell_start, int *__restrict__ row_len, const int *__restrict__ row_start,
                       int *__restrict__ cols, Geo g, int ncell) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cell_start[ncell]) return;                 // number of binned particles
  double4 p = ld_rec_nc(&sorted_posm[t]);
  if (!(meta_of(p) & MF_REF)) return;                 // rows exist only for ref atoms
  int s = sorted_slot[t];
  int lin = cell_of[s];
  if (lin < 0) return;
  int cx = lin % g.hd[0], r = lin / g.hd[0], cy = r % g.hd[1], cz = r / g.hd[1];
  int cnt = 0;
  int base = FILL ? row_start[s] : 0;
#pragma unroll 1
  for (int nab = 0; nab < 27; ++nab) {
    int nx = (c_map[nab][0] + cx - 1 + g.nc[0]) % g.nc[0] + 1;       // wraps every axis, z included
    int ny = (c_map[nab][1] + cy - 1 + g.nc[1]) % g.nc[1] + 1;
    int nz = (c_map[nab][2] + cz - 1 + g.nc[2]) % g.nc[2] + 1;
    int nl = cell_lin(g, nx, ny, nz);
    int b = cell_start[nl], e = cell_start[nl + 1];
    for (int u = b; u < e; ++u) {
      if (u == t) continue;
      double4 q = ld_rec_nc(&sorted_posm[u]);
      double rd = dist2_idnint(g, q.x, q.y, q.z, p.x, p.y, p.z);   // vdistance(vd,aj,ai)
      if (rd < g.rc_list2) { if (FILL) cols[base + cnt] = sorted_slot[u]; ++cnt; }
    }
  }
  if (!FILL) row_len[s] = cnt;
}
// row capacity = length + slack (room for incremental gcmc appends, Neighbor.F90:309-312)
__global__ void k_row_caps(const int *__restrict__ row_len, int *__restrict__ row_cap, int n, int slack) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) row_cap[s] = row_len[s] + slack;
}
__global__ void k_sum_int(const int *__restrict__ v, int n, long long *__restrict__ out) {
  long long a = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a += v[i];
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
__global__ void k_rev_count(const int *__restrict__ row_start, const int *__restrict__ row_len, const int *__restrict__ cols,
                            int *__restrict__ rev_len, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int b = row_start[s], len = row_len[s];
  for (int jj = 0; jj < len; ++jj) atomicAdd(&rev_len[cols[b + jj]], 1);
}
__global__ void k_rev_fill(const int *__restrict__ row_start, const int *__restrict__ row_len, const int *__restrict__ cols,
                           const int *__restrict__ rev_start, int *__restrict__ rev_cur, int *__restrict__ rev_cols, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int b = row_start[s], len = row_len[s];
  for (int jj = 0; jj < len; ++jj) { int j = cols[b + jj]; rev_cols[rev_start[j] + atomicAdd(&rev_cur[j], 1)] = s; }
}

// Reverse-visit candidates of atom s: with symmetric rows they are the ref entries of its own row, otherwise rev(s).
template <bool STRICT>
__global__ void __launch_bounds__(TPB) k_fuerza(const double4 *__restrict__ posm, const int *__restrict__ row_start,
                                                const int *__restrict__ row_len, const int *__restrict__ cols,
                                                const int *__restrict__ rev_start, const int *__restrict__ rev_len,
                                                const int *__restrict__ rev_cols, int asym,
                                                const int *__restrict__ uid, double *__restrict__ force, double *__restrict__ epot,
                                                Geo g, Phys ph, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double4 p1 = ld_rec_nc(&posm[s]);
  long long m1 = meta_of(p1);
  if (!(m1 & MF_REF)) return;
  int k = (int)(m1 & MF_TYPE);
  int b = row_start[s], len = row_len[s];
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
      if (!asym && (m2 & MF_REF)) { fx = fx + f[0]; fy = fy + f[1]; fz = fz + f[2]; ep = ep + u; }
    }
    if (asym) {
      for (int jj = 0; jj < rvlen; ++jj) {
        int j = rv[jj];
        double4 p2 = ld_rec_nc(&posm[j]);
        long long m2 = meta_of(p2);
        int m = (int)(m2 & MF_TYPE);
        if (m == 0 || !(m2 & MF_REF)) continue;
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
      if (m == 0 || !(m2 & MF_REF)) continue;
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
          if (m == 0 || !(m2 & MF_REF)) continue;
          double f[3], u;
          if (!pair_terms(g, ph, p1, k, p2, m, f, u)) continue;
          fx = fx + f[0]; fy = fy + f[1]; fz = fz + f[2]; ep = ep + u;
        }
      }
    }
  }
  force[3 * s] = fx; force[3 * s + 1] = fy; force[3 * s + 2] = fz; epot[s] = ep;
}

// ================================================================================================
// K4  integrators + boundary handling   (ermak_a dana.F90:974-1028, cbrownian_hs 798-846, atom_pbc 1187-1250)
// ================================================================================================
struct BlockAcc { long long tr, de; double msd, mv; };

__device__ __forceinline__ void block_flush(BlockAcc a, DevScal *sc) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a.tr += __shfl_xor_sync(0xffffffffu, a.tr, o); a.de += __shfl_xor_sync(0xffffffffu, a.de, o);
    a.msd += __shfl_xor_sync(0xffffffffu, a.msd, o); a.mv = fmax(a.mv, __shfl_xor_sync(0xffffffffu, a.mv, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (a.tr) atomicAdd((unsigned long long *)&sc->try_, (unsigned long long)a.tr);
    if (a.de) atomicAdd((unsigned long long *)&sc->depo, (unsigned long long)a.de);
    if (a.msd != 0.0) atomicAdd(&sc->msd_t, a.msd);
    if (a.mv > 0.0) atomicMax((unsigned long long *)&sc->max_vel, (unsigned long long)__double_as_longlong(a.mv));
  }
}

// atom_pbc — dana.F90:1187-1250.  Returns depos.  The deposition uniform is drawn only when z<=0.
struct RngSrc { int mode; unsigned long long seed; unsigned int id, step; const double *rp; int slot; };
__device__ __forceinline__ bool atom_pbc_dev(const Geo &g, const Phys &ph, double zmax, double q[3], double po[3], const double og[3],
                                             double v[3], long long &meta, const RngSrc &rs, BlockAcc &acc) {
  bool depos = false;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (q[j] > g.box[j]) { q[j] = q[j] - g.box[j]; po[j] = po[j] - g.box[j]; }
    if (q[j] < 0.0) { q[j] = q[j] + g.box[j]; po[j] = po[j] + g.box[j]; }
  }
  if (q[2] > zmax) {
    if (ph.integrador) { q[2] = q[2] - 2 * (q[2] - zmax); v[2] = -v[2]; }
    else { q[0] = og[0]; q[1] = og[1]; q[2] = og[2]; }
  }
  acc.msd += v[0] * v[0] * ph.h * ph.h;
  if (q[2] <= 0.0) {
    acc.tr++;
    double ne;
    if (rs.mode == 1) ne = rs.rp[rs.slot];
    else { Philox r; r.run(rs.seed, rs.id, rs.step, RS_PBC, 0u); ne = r.u01(0); }
    if (ne < ph.prob) { acc.de++; meta = (meta & ~MF_TYPE) | 3; depos = true; }
    q[0] = og[0]; q[1] = og[1]; q[2] = og[2];
  }
  return depos;
}

template <bool ERMAK>
__global__ void __launch_bounds__(TPB) k_integrate(double4 *__restrict__ posm, double *__restrict__ vel, const double *__restrict__ acel,
                                                   double *__restrict__ pos_old, double *__restrict__ old_cg, double *__restrict__ ranv,
                                                   const int *__restrict__ uid, const double *__restrict__ rp_gauss,
                                                   const double *__restrict__ rp_upbc, DevScal *__restrict__ sc, Geo g, Phys ph,
                                                   unsigned int step, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  BlockAcc acc = {0, 0, 0.0, 0.0};
  if (s < n) {
    double4 p = ld_rec(&posm[s]);
    long long m = meta_of(p);
    if (m & MF_REF) {
      double q[3] = {p.x, p.y, p.z}, og[3] = {p.x, p.y, p.z};
      double v[3] = {vel[3 * s], vel[3 * s + 1], vel[3 * s + 2]};
      double po[3] = {pos_old[3 * s], pos_old[3 * s + 1], pos_old[3 * s + 2]};
      old_cg[3 * s] = og[0]; old_cg[3 * s + 1] = og[1]; old_cg[3 * s + 2] = og[2];
      int zt = (int)(m & MF_TYPE);
      double gs[6];
      if (ph.rng_mode == 1) {
#pragma unroll
        for (int i = 0; i < (ERMAK ? 6 : 3); ++i) gs[i] = rp_gauss[6 * s + i];
      } else {
        Philox r; unsigned int id = (unsigned int)uid[s];
#pragma unroll
        for (int i = 0; i < (ERMAK ? 3 : 2); ++i) { r.run(ph.seed, id, step, RS_INTEG0 + i, 0u); r.gauss2(gs[2 * i], gs[2 * i + 1]); }
      }
      if (ERMAK) {
        double a[3] = {acel[3 * s], acel[3 * s + 1], acel[3 * s + 2]};
        double sm = ph.sqrt_mass[zt - 1];
        double A = ph.skt / sm * ph.sdr, B = ph.skt / sm * ph.sdv;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double r1 = gs[2 * j], r2 = gs[2 * j + 1];
          double ranr = A * r1;
          q[j] = q[j] + ph.cc1 * v[j] + ph.cc2h * a[j] + ranr;
          ranv[3 * s + j] = B * (ph.crv1 * r1 + ph.crv2 * r2);
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
      RngSrc rs = {ph.rng_mode, ph.seed, (unsigned int)uid[s], step, rp_upbc, s};
      bool depos = atom_pbc_dev(g, ph, zmax, q, po, og, v, m, rs, acc);
      if (!depos) {
        if (!ERMAK) acc.mv = fmax(acc.mv, (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
        m &= ~MF_SKIP;
      }
      p.x = q[0]; p.y = q[1]; p.z = q[2]; p.w = meta_as_double(m);
      st_rec(&posm[s], p);
      vel[3 * s] = v[0]; vel[3 * s + 1] = v[1]; vel[3 * s + 2] = v[2];
      pos_old[3 * s] = po[0]; pos_old[3 * s + 1] = po[1]; pos_old[3 * s + 2] = po[2];
    }
  }
  block_flush(acc, sc);
}

// ermak_b — dana.F90:1031-1052
__global__ void __launch_bounds__(TPB) k_ermak_b(const double4 *__restrict__ posm, double *__restrict__ vel, double *__restrict__ acel,
                                                 const double *__restrict__ force, const double *__restrict__ ranv, Phys ph, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  long long m = meta_of(ld_rec_nc(&posm[s]));
  if (!(m & MF_REF)) return;
  int zt = (int)(m & MF_TYPE);
  if (zt == 2) return;
  double mass = ph.mass[zt - 1];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double f = force[3 * s + k], a = acel[3 * s + k], v = vel[3 * s + k];
    vel[3 * s + k] = ph.cc0 * v + ph.cc1mcc2 * a + ph.cc2 * f / mass + ranv[3 * s + k];
    acel[3 * s + k] = f / mass;
  }
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

__global__ void k_ov_init(const double4 *__restrict__ posm, int *__restrict__ parent, int *__restrict__ ovst,
                          int *__restrict__ comp_cnt, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  long long m = meta_of(ld_rec_nc(&posm[s]));
  parent[s] = s; comp_cnt[s] = 0;
  ovst[s] = ((m & MF_SKIP) ? OV_SKIP : 0) | ((int)(m & MF_TYPE) << OV_TSHIFT);
}
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
__global__ void __launch_bounds__(TPB) k_ov_detect(const double4 *__restrict__ posm, const double *__restrict__ old_cg,
                                                   const int *__restrict__ row_start, const int *__restrict__ row_len,
                                                   const int *__restrict__ cols, int *__restrict__ parent, int *__restrict__ ovst,
                                                   DevScal *__restrict__ sc, Geo g, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double4 p1 = ld_rec_nc(&posm[s]);
  long long m1 = meta_of(p1);
  if (!(m1 & MF_REF)) return;
  double o1[3] = {old_cg[3 * s], old_cg[3 * s + 1], old_cg[3 * s + 2]};
  int b = row_start[s], len = row_len[s];
  bool inv = false;
  for (int jj = 0; jj < len; ++jj) {
    int j = cols[b + jj];
    double4 p2 = ld_rec_nc(&posm[j]);
    long long m2 = meta_of(p2);
    if (!(m2 & MF_TYPE)) continue;
    bool hit = dist2_idnint(g, p1.x, p1.y, p1.z, p2.x, p2.y, p2.z) <= g.rcut2 ||
               dist2_idnint(g, o1[0], o1[1], o1[2], p2.x, p2.y, p2.z) <= g.rcut2;
    if (m2 & MF_REF) {
      double o2[3] = {old_cg[3 * j], old_cg[3 * j + 1], old_cg[3 * j + 2]};
      hit = hit || dist2_idnint(g, p1.x, p1.y, p1.z, o2[0], o2[1], o2[2]) <= g.rcut2 ||
            dist2_idnint(g, o1[0], o1[1], o1[2], o2[0], o2[1], o2[2]) <= g.rcut2;
      if (hit) { uf_unite(parent, s, j); atomicOr(&ovst[j], OV_INVOLVED); }
    }
    inv = inv || hit;
  }
  if (inv) atomicOr(&ovst[s], OV_INVOLVED);
}
__global__ void k_ov_count(int *__restrict__ parent, const int *__restrict__ ovst, int *__restrict__ comp_cnt, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  if (!(ovst[s] & OV_INVOLVED)) return;
  int r = uf_find(parent, s);
  parent[s] = r;
  atomicAdd(&comp_cnt[r], 1);
}
__global__ void k_ov_alloc(const int *__restrict__ parent, const int *__restrict__ ovst, int *__restrict__ comp_cnt,
                           int *__restrict__ comp_off, int *__restrict__ roots, DevScal *__restrict__ sc, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  if (!(ovst[s] & OV_INVOLVED) || parent[s] != s) return;
  int c = comp_cnt[s];
  comp_off[s] = atomicAdd(&sc->member_cursor, c);
  comp_cnt[s] = 0;                                       // reused as the fill cursor
  roots[atomicAdd(&sc->n_roots, 1)] = s;
}
__global__ void k_ov_fill(const int *__restrict__ parent, const int *__restrict__ ovst, int *__restrict__ comp_cnt,
                          const int *__restrict__ comp_off, int *__restrict__ members, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  if (!(ovst[s] & OV_INVOLVED)) return;
  int r = parent[s];
  members[comp_off[r] + atomicAdd(&comp_cnt[r], 1)] = s;
}
// one thread per component: order the members by creation rank (once)
__global__ void k_ov_sort(const int *__restrict__ roots, const int *__restrict__ comp_cnt, const int *__restrict__ comp_off,
                          int *__restrict__ members, const int *__restrict__ uid, const DevScal *__restrict__ sc) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= sc->n_roots) return;
  int root = roots[r], b = comp_off[root], e = b + comp_cnt[root];
  for (int i = b + 1; i < e; ++i) {
    int s = members[i], key = uid[s], j = i - 1;
    while (j >= b && uid[members[j]] > key) { members[j + 1] = members[j]; --j; }
    members[j + 1] = s;
  }
}
__device__ __forceinline__ void ov_pos(const double4 *posm, const double *old_cg, int a, int st, double q[3]) {
  if (st & OV_MOVED) { q[0] = old_cg[3 * a]; q[1] = old_cg[3 * a + 1]; q[2] = old_cg[3 * a + 2]; }
  else { double4 p = ld_rec_nc(&posm[a]); q[0] = p.x; q[1] = p.y; q[2] = p.z; }
}
// one reference pass (one recursion level) for every component
__global__ void k_ov_pass(const double4 *__restrict__ posm, const double *__restrict__ old_cg, const int *__restrict__ row_start,
                          const int *__restrict__ row_len, const int *__restrict__ cols, int *__restrict__ ovst,
                          const int *__restrict__ roots, const int *__restrict__ comp_cnt, const int *__restrict__ comp_off,
                          const int *__restrict__ members, const int *__restrict__ uid, const double *__restrict__ rp_uovl,
                          DevScal *__restrict__ sc, Geo g, Phys ph, unsigned int step, int pass, int guard) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= sc->n_roots) return;
  int root = roots[r], b = comp_off[root], e = b + comp_cnt[root];
  long long tr = 0, de = 0, ch = 0, ch3 = 0; bool again = false;
  double z0 = sc->z0;
  for (int i = b; i < e; ++i) {
    int a1 = members[i];
    int st1 = ((volatile int *)ovst)[a1];
    if (st1 & OV_SKIP) continue;
    if (!((st1 >> OV_TSHIFT) & 3)) continue;
    st1 |= OV_SKIP;
    double q1[3]; ov_pos(posm, old_cg, a1, st1, q1);
    int rb = row_start[a1], rl = row_len[a1];
    for (int jj = 0; jj < rl; ++jj) {
      int a2 = cols[rb + jj];
      int st2 = ((volatile int *)ovst)[a2];
      int t2 = (st2 >> OV_TSHIFT) & 3;
      if (t2 == 0) continue;                                   // limbo (dana.F90:881-883)
      double q2[3]; ov_pos(posm, old_cg, a2, st2, q2);
      double dr = dist2_idnint(g, q1[0], q1[1], q1[2], q2[0], q2[1], q2[2]);
      if (dr > g.rcut2) continue;
      if (t2 == 2) {                                           // contact with metal: deposition attempt
        tr++;
        double ne;
        if (ph.rng_mode == 1) ne = rp_uovl ? rp_uovl[a1] : 0.0;
        else { Philox rr; rr.run(ph.seed, (unsigned int)uid[a1], step, RS_OVERLAP, (unsigned int)pass); ne = rr.u01(0); }
        if (ne < ph.prob) {
          de++; st1 = (st1 & ~(3 << OV_TSHIFT)) | (3 << OV_TSHIFT);
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
          if (q1[0] == og1[0] && q1[1] == og1[1] && q1[2] == og1[2]) { ch3++; continue; }
        }
      }
      // o2 goes back to its previous position; velocities are zeroed when the state is applied
      st2 = (st2 | OV_MOVED | OV_ZERO) & ~OV_SKIP;
      ovst[a2] = st2;
      ch++; again = true;
    }
    ovst[a1] = st1;
  }
  if (tr) atomicAdd((unsigned long long *)&sc->try_, (unsigned long long)tr);
  if (de) atomicAdd((unsigned long long *)&sc->depo, (unsigned long long)de);
  if (ch) atomicAdd((unsigned long long *)&sc->choques, (unsigned long long)ch);
  if (ch3) atomicAdd((unsigned long long *)&sc->choques3, (unsigned long long)ch3);
  if (again) sc->again = 1;
}
// write the resolved state back: positions, zeroed vel/acel of moved-back atoms, skip flags and new F atoms
__global__ void k_ov_apply(double4 *__restrict__ posm, double *__restrict__ vel, double *__restrict__ acel,
                           const double *__restrict__ old_cg, const int *__restrict__ ovst, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
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
  if (s == 0) sc->zmax = sc->zmax - lohi * (sc->zmax - z0);
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
  else if (zz >= 1 && zz <= 3) m = zz | ((f & 1) ? MF_REF : 0) | ((f & 2) ? MF_GCMC : 0) | ((f & 4) ? MF_SKIP : 0);
  double4 p = {pos[3 * s], pos[3 * s + 1], pos[3 * s + 2], meta_as_double(m)};
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
__global__ void k_count_members(const double4 *__restrict__ posm, DevScal *__restrict__ sc, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  int a = 0, r = 0, gq = 0, l = 0;
  if (s < n) { long long m = meta_of(ld_rec_nc(&posm[s])); a = (m & MF_TYPE) ? 1 : 0; r = (m & MF_REF) ? 1 : 0; gq = (m & MF_GCMC) ? 1 : 0; l = (m & MF_LIMBO) ? 1 : 0; }
  a = __reduce_add_sync(0xffffffffu, a); r = __reduce_add_sync(0xffffffffu, r); gq = __reduce_add_sync(0xffffffffu, gq); l = __reduce_add_sync(0xffffffffu, l);
  if ((threadIdx.x & 31) == 0) { if (a) atomicAdd(&sc->nat_sys, a); if (r) atomicAdd(&sc->nat_ref, r); if (gq) atomicAdd(&sc->nat_gcmc, gq); if (l) atomicAdd(&sc->nlimbo, l); }
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
