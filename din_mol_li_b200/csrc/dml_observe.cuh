// dml_observe.cuh — device-side output reductions and observables.
//   k_salida_sums       energia / kion sums of salida()      src/dana.F90:1143-1183, 1342-1376
//   k_density_profile   Li density profile rho(z)            (analysis tool; the reference only has the scalar calc_rho, dana.F90:521-549)
//   k_gr_*              pair-distance histogram g(r) on a private cell grid, distances = vdistance (Groups.F90:995-1016)
// The frame of a 1 M-particle box is 56 MB; these kernels turn what salida needs into a few doubles on the device so the
// per-frame download disappears from the step loop.
#pragma once
#include "dml_kernels.cuh"

namespace dml {

// ---- salida(): energia = sum of epot over sys (dana.F90:1160); kion(): vdac = sum of m*v^2 over the non-CG atoms, j = their
// count (dana.F90:1356-1372).  The reference adds in sys%alist order; here every thread adds its slots in ascending order, the
// block combines with a fixed shuffle tree and the last block adds the per-block partials in block order: the result is the
// same from run to run and differs from the reference's serial sum only by re-association (1e-12 relative in the tests).
// out[0] = energia (all of sys), out[1] = energia of hs%ref only (the parity scope of SURVEY.md Q2), out[2] = vdac, out[3] = j.
constexpr int OBS_TPB = 256;
__global__ void __launch_bounds__(OBS_TPB) k_salida_sums(const double4 *__restrict__ posm, const double4 *__restrict__ fe,
                                                         const double *__restrict__ vel, Phys ph, int n, double *__restrict__ part,
                                                         unsigned int *__restrict__ ticket, double *__restrict__ out) {
  double e = 0.0, er = 0.0, vd = 0.0, jm = 0.0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    const long long m = meta_of(ld_rec_nc(&posm[s]));
    const int zt = (int)(m & MF_TYPE);
    if (zt == 0 || (m & MF_GHOST)) continue;
    const double ep = ld_rec_nc(&fe[s]).w;
    e = e + ep;
    if (m & MF_REF) er = er + ep;
    if (zt != 2) {
      const double v0 = vel[3 * s], v1 = vel[3 * s + 1], v2 = vel[3 * s + 2];
      double q = (v0 * v0 + v1 * v1) + v2 * v2;               // dot_product(vel,vel)
      q = q * ph.mass[zt - 1];
      vd = vd + q; jm = jm + 1.0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e = e + __shfl_down_sync(0xffffffffu, e, o); er = er + __shfl_down_sync(0xffffffffu, er, o);
    vd = vd + __shfl_down_sync(0xffffffffu, vd, o); jm = jm + __shfl_down_sync(0xffffffffu, jm, o);
  }
  __shared__ double sh[OBS_TPB / 32][4];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[w][0] = e; sh[w][1] = er; sh[w][2] = vd; sh[w][3] = jm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < OBS_TPB / 32; ++i) for (int k = 0; k < 4; ++k) a[k] = a[k] + sh[i][k];
    for (int k = 0; k < 4; ++k) part[4 * blockIdx.x + k] = a[k];
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x < 4) {
    __threadfence();
    double a = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) a = a + ((volatile double *)part)[4 * b + threadIdx.x];
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

// ---- rho(z): counts[b] = #{particles of the selected elements with zlo + b*dz <= z < zlo + (b+1)*dz}; bin = int((z-zlo)/dz) with
// an fp64 division, so a numpy restatement with the same two operations gives identical integers.
constexpr int OBS_MAX_BINS = 8192;
__global__ void __launch_bounds__(OBS_TPB) k_density_profile(const double4 *__restrict__ posm, int n, double zlo, double dz, int nbins,
                                                             int type_mask, unsigned long long *__restrict__ counts) {
  extern __shared__ unsigned int hist[];
  for (int b = threadIdx.x; b < nbins; b += blockDim.x) hist[b] = 0u;
  __syncthreads();
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    const double4 p = ld_rec_nc(&posm[s]);
    const long long m = meta_of(p);
    const int zt = (int)(m & MF_TYPE);
    if (zt == 0 || (m & MF_GHOST) || !((type_mask >> zt) & 1)) continue;
    const double q = (p.z - zlo) / dz;
    if (!(q >= 0.0)) continue;
    if (q >= (double)nbins) continue;
    const int b = (int)q;
    if (b < nbins) atomicAdd(&hist[b], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += blockDim.x) if (hist[b]) atomicAdd(&counts[b], (unsigned long long)hist[b]);
}

// ---- g(r) on a private grid (cells at least rmax wide; x,y periodic, z clamped).  The neighbour-list cells are not reused: they
// hold the snapshot of the last rebuild, and re-sorting them here would invalidate rows that are still pending (lazy build).
struct GrGrid { int nc[3]; double cell[3]; };
__device__ __forceinline__ int gr_axis(double x, double cell, int nc) {
  int c = (int)(x / cell);
  return c < 0 ? 0 : (c >= nc ? nc - 1 : c);              // monotone clamp keeps adjacency (cells are >= rmax wide)
}
__global__ void __launch_bounds__(OBS_TPB) k_gr_bin(const double4 *__restrict__ posm, int n, GrGrid gg, int type_mask,
                                                    int *__restrict__ cell_of, int *__restrict__ cell_cnt, int *__restrict__ nsel) {
  int mine = 0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    const double4 p = ld_rec_nc(&posm[s]);
    const long long m = meta_of(p);
    const int zt = (int)(m & MF_TYPE);
    int c = -1;
    if (zt != 0 && !(m & MF_GHOST) && ((type_mask >> zt) & 1)) {
      c = (gr_axis(p.z, gg.cell[2], gg.nc[2]) * gg.nc[1] + gr_axis(p.y, gg.cell[1], gg.nc[1])) * gg.nc[0] + gr_axis(p.x, gg.cell[0], gg.nc[0]);
      atomicAdd(&cell_cnt[c], 1);
      ++mine;
    }
    cell_of[s] = c;
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(nsel, mine);
}
__global__ void __launch_bounds__(OBS_TPB) k_gr_scatter(const double4 *__restrict__ posm, int n, const int *__restrict__ cell_of,
                                                        const int *__restrict__ cell_start, int *__restrict__ cell_cur,
                                                        double4 *__restrict__ sorted) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
    const int c = cell_of[s];
    if (c < 0) continue;
    const int k = atomicAdd(&cell_cur[c], 1);
    st_rec(&sorted[cell_start[c] + k], ld_rec_nc(&posm[s]));
  }
}
// one thread per selected particle a (cell-sorted index); a pair {a,b} is counted by the side with the smaller sorted index
__global__ void __launch_bounds__(OBS_TPB) k_gr_pairs(const double4 *__restrict__ sorted, int nsel, const int *__restrict__ cell_start,
                                                      GrGrid gg, Geo g, double rmax2, double dr_bin, int nbins,
                                                      unsigned long long *__restrict__ counts) {
  extern __shared__ unsigned int hist[];
  for (int b = threadIdx.x; b < nbins; b += blockDim.x) hist[b] = 0u;
  __syncthreads();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < nsel) {
    const double4 p = ld_rec_nc(&sorted[a]);
    const int cx = gr_axis(p.x, gg.cell[0], gg.nc[0]), cy = gr_axis(p.y, gg.cell[1], gg.nc[1]), cz = gr_axis(p.z, gg.cell[2], gg.nc[2]);
    for (int dz = -1; dz <= 1; ++dz) {
      const int z = cz + dz;
      if (z < 0 || z >= gg.nc[2]) continue;                 // z is not periodic (dana.F90:483-484)
      for (int dy = -1; dy <= 1; ++dy) {
        const int y = (cy + dy + gg.nc[1]) % gg.nc[1];
        for (int dx = -1; dx <= 1; ++dx) {
          const int x = (cx + dx + gg.nc[0]) % gg.nc[0];
          const int c = (z * gg.nc[1] + y) * gg.nc[0] + x;
          const int e = cell_start[c + 1];
          for (int b = max(cell_start[c], a + 1); b < e; ++b) {
            const double4 q = ld_rec_nc(&sorted[b]);
            const double dr2 = dist2_idnint(g, p.x, p.y, p.z, q.x, q.y, q.z);
            if (!(dr2 < rmax2)) continue;
            const int bin = (int)(sqrt(dr2) / dr_bin);
            if (bin < nbins) atomicAdd(&hist[bin], 1u);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += blockDim.x) if (hist[b]) atomicAdd(&counts[b], (unsigned long long)hist[b]);
}


// ---- host object model sync: what changed in sys / hs%ref / gcmc membership since the last snapshot -----------
// The reference keeps membership in pointer lists (Groups.F90 atom / group / igroup); the device changes it in four places
// (gcmc insertion and deletion dana.F90:655-706, chunk blocks 716-773, Li -> F on deposition 1236-1240, F -> CG promotion 228-236).
// Instead of instrumenting each of them, the slot's occupant (creation rank) and a membership byte are snapshotted and compared.
constexpr int MC_NEW = 1, MC_GONE = 2, MC_ELEMENT = 4, MC_LEFT_REF = 8, MC_LEFT_GCMC = 16;
__device__ __forceinline__ int member_byte(long long m) {       // element | ref << 2 | gcmc << 3; 0 = empty slot (or parked on limbo)
  const int zt = (int)(m & MF_TYPE);
  if (zt == 0 || (m & MF_GHOST)) return 0;
  return zt | ((m & MF_REF) ? 4 : 0) | ((m & MF_GCMC) ? 8 : 0);
}
__global__ void __launch_bounds__(OBS_TPB) k_member_snap(const double4 *__restrict__ posm, const int *__restrict__ uid, int n, int cap,
                                                        int *__restrict__ snap_uid, unsigned char *__restrict__ snap_mb) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += gridDim.x * blockDim.x) {
    const int mb = s < n ? member_byte(meta_of(ld_rec_nc(&posm[s]))) : 0;
    snap_mb[s] = (unsigned char)mb; snap_uid[s] = mb ? uid[s] : -1;
  }
}
__global__ void __launch_bounds__(OBS_TPB) k_member_diff(const double4 *__restrict__ posm, const int *__restrict__ uid, int n, int cap,
                                                        int *__restrict__ snap_uid, unsigned char *__restrict__ snap_mb,
                                                        int max_out, int4 *__restrict__ out, int *__restrict__ count) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += gridDim.x * blockDim.x) {
    const int mb = s < n ? member_byte(meta_of(ld_rec_nc(&posm[s]))) : 0;
    const int u = mb ? uid[s] : -1;
    const int omb = snap_mb[s], ou = snap_uid[s];
    int kind = 0;
    if (mb && (!omb || ou != u)) kind |= MC_NEW;
    if (omb && (!mb || ou != u)) kind |= MC_GONE;
    if (mb && omb && ou == u) {
      if ((mb & 3) != (omb & 3)) kind |= MC_ELEMENT;
      if ((omb & 4) && !(mb & 4)) kind |= MC_LEFT_REF;
      if ((omb & 8) && !(mb & 8)) kind |= MC_LEFT_GCMC;
    }
    if (kind) {
      const int k = atomicAdd(count, 1);
      if (k < max_out) { out[k] = make_int4(s, kind, u, mb & 3); snap_mb[s] = (unsigned char)mb; snap_uid[s] = u; }   // unreported slots stay pending
    }
  }
}

} // namespace dml
