// dml_device.cuh — device-side building blocks shared by the kernels of libdml.so.
//
// Bit-exactness notes (SURVEY.md §7 "Hard parts"): the whole library is compiled with -fmad=false so no
// multiply-add is contracted; fp64 division and sqrt are IEEE round-to-nearest on the device; idnint is
// round() (half away from zero); int(x) is a C cast (truncation).  With those four rules the distance tests
// and the force arithmetic below round exactly like the reference's -O0 x86-64 build.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dml {

// ---- particle record: double4 {x, y, z, meta}; meta is an int64 bit-cast into the 4th lane -------------
constexpr long long MF_TYPE = 3;    // bits 0-1: element id (0 empty, 1 Li, 2 CG, 3 F)  dana.F90:82-84
constexpr long long MF_REF = 4;     // member of hs%ref
constexpr long long MF_GCMC = 8;    // member of gcmc
constexpr long long MF_SKIP = 16;   // atom%skip
constexpr long long MF_LIMBO = 32;  // slot parked on hs%limbo until the next full build
constexpr long long MF_GHOST = 64;  // slab mode: copy of a particle owned by a neighbouring slab (no row, never integrated)
constexpr long long MF_GREF = 128;  // slab mode: ghost that is a member of hs%ref on its owner (reverse visits exist there)
constexpr long long MF_ANYREF = MF_REF | MF_GREF;

// bits 32-63: float32 upper bound of |pos - old_cg| (minimum image) since the last integrator call, +inf when unknown.
// It only feeds the exact-safe prefilter of k_ov_detect; all flag tests mask the low bits.
constexpr unsigned int DISP_INF = 0x7f800000u;
__device__ __forceinline__ long long meta_of(const double4 &p) { return __double_as_longlong(p.w); }
__device__ __forceinline__ float disp_of(long long m) { return __int_as_float((int)((unsigned long long)m >> 32)); }
__device__ __forceinline__ long long with_disp(long long m, unsigned int bits) { return (m & 0xffffffffll) | ((long long)bits << 32); }
__device__ __forceinline__ double meta_as_double(long long m) { return __longlong_as_double(m); }

// One 256-bit load of a particle record (sm_100: ld.global.v4.f64 needs 32-byte alignment; cudaMalloc gives 256).
__device__ __forceinline__ double4 ld_rec(const double4 *p) {
  double4 r;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double4 ld_rec_nc(const double4 *p) {   // read-only path (record not written in this kernel)
  double4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_rec(double4 *p, const double4 &r) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.x), "d"(r.y), "d"(r.z), "d"(r.w) : "memory");
}

// ---- row head: one 32-byte sector per slot ---------------------------------------------------------------------
// Everything a consumer needs before it touches a row, fetched together with the particle record in ONE memory round trip and
// written by the list build as ONE full-sector store: where the row lives, and the NEAR LIST — the (up to) four entries of the row
// with the smallest quantised build-time distance (see k_rows), as slot ids in row order, plus the fifth-smallest distance.
// A consumer that has to look at every entry with build distance <= qmax (pair force, overlap detection: see skip_qmax) finds
// all of them in the near list whenever q5 > qmax — in solution that is nearly always — and goes from the head straight to the
// partners' records; round 1 kept sixteen distance bytes here and paid a third dependent round trip for the indices.
struct __align__(32) RowHead {
  int near[4];            // -1 = none
  unsigned char nbq[4];   // quantised build distances of near[]; 255 = none
  int start;              // first entry in cols[] / bq[] (slot*ROW_W, or a segment of the tail region for long rows)
  unsigned short len;     // nn(i)   (Neighbor.F90:45)
  unsigned short cap;     // entries that fit at start (gcmc appends, Neighbor.F90:295-314)
  unsigned char q5;       // fifth-smallest build distance of the row (255 = fewer than five entries; 0 = near list not maintained)
  unsigned char pad[3];
};
struct RowMeta { unsigned int nbq; int start, len, cap, q5; };
__device__ __forceinline__ int4 rh_near(const RowHead *p) { return __ldg(reinterpret_cast<const int4 *>(p)); }
__device__ __forceinline__ RowMeta rh_meta(const RowHead *p) {
  const int4 v = __ldg(reinterpret_cast<const int4 *>(p) + 1);
  RowMeta m; m.nbq = (unsigned int)v.x; m.start = v.y; m.len = v.z & 0xffff; m.cap = (v.z >> 16) & 0xffff; m.q5 = v.w & 255;
  return m;
}
__device__ __forceinline__ void rh_store(RowHead *p, const int4 &nr, unsigned int nbq, int start, int len, int cap, int q5) {
  reinterpret_cast<int4 *>(p)[0] = nr;
  reinterpret_cast<int4 *>(p)[1] = make_int4((int)nbq, start, (len & 0xffff) | (min(cap, 65535) << 16), q5 & 255);
}
// a row without a near list (empty, or one whose every entry has to be looked at): q5 = 255 means "nothing beyond the near list"
__device__ __forceinline__ void rh_store_plain(RowHead *p, int start, int len, int cap, int q5) {
  rh_store(p, make_int4(-1, -1, -1, -1), 0xffffffffu, start, len, cap, q5);
}
// The five smallest keys (build distance << 16 | row position) of a row, ascending: a compare-exchange chain in registers.
struct Near5 { unsigned int k[5]; };
__device__ __forceinline__ void near5_init(Near5 &n) {
#pragma unroll
  for (int i = 0; i < 5; ++i) n.k[i] = 0xffffffffu;
}
__device__ __forceinline__ void near5_add(Near5 &n, unsigned int key) {
#pragma unroll
  for (int i = 0; i < 5; ++i) { const unsigned int lo = min(n.k[i], key); key = max(n.k[i], key); n.k[i] = lo; }
}
// head of a finished row from its five smallest keys: the four nearest in ROW order.  Key layout: distance << 16 | position
// (NEAR_POS_BITS wide) [| index into the caller's own table of slot ids]; slot_of(key) returns the slot id of the entry.
template <int POS_SHIFT, typename SlotOf>
__device__ __forceinline__ void rh_store_near(RowHead *p, const Near5 &n, int start, int len, int cap, SlotOf slot_of) {
  unsigned int e[4];                                     // low 16 key bits << 8 | distance: sorting orders by row position
#pragma unroll
  for (int i = 0; i < 4; ++i) e[i] = n.k[i] == 0xffffffffu ? 0xffffffffu : (((n.k[i] & 0xffffu) << 8) | (n.k[i] >> 16));
#define CE_(a, b) { const unsigned int lo_ = min(e[a], e[b]), hi_ = max(e[a], e[b]); e[a] = lo_; e[b] = hi_; }
  CE_(0, 1) CE_(2, 3) CE_(0, 2) CE_(1, 3) CE_(1, 2)
#undef CE_
  int nr[4]; unsigned int nb = 0u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool on = e[i] != 0xffffffffu;
    nr[i] = on ? slot_of((int)(e[i] >> 8)) : -1;
    nb |= (on ? (e[i] & 255u) : 255u) << (8 * i);
  }
  const int q5 = n.k[4] == 0xffffffffu ? 255 : (int)(n.k[4] >> 16);
  rh_store(p, make_int4(nr[0], nr[1], nr[2], nr[3]), nb, start, len, cap, q5);
}

// ---- the reference's own generator on the device (DML_RNG_REFERENCE) --------------------------------------------------------
// ran (src/dana.F90:1407-1428): Park-Miller with Schrage's trick combined with a 13/17/5 xorshift; gasdev (1379-1404): Marsaglia's
// polar method whose squared radius, logarithm and factor are real(sp).  The logarithm is glibc's logf, restated here (its table
// and polynomial are public: sysdeps/ieee754/flt-32/e_logf.c, logf_data.c of glibc 2.28+): 16-entry table, degree-3 polynomial in
// double, one rounding to float.  tests/ checks the restatement against the host libm over every float in (0, 1].
// One stream, one spare deviate, shared by every caller in call order: all draws are made by ONE thread.
struct RefRng { int idum, ix, iy, stored; double g; unsigned long long calls; };
__device__ __forceinline__ float ref_logf(float x) {
  const double T[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2}, {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},
    {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3}, {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4}, {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5},
    {0x1p+0, 0x0p+0}, {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5}, {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3}, {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3}, {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},
    {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
  const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2, Ln2 = 0x1.62e42fefa39efp-1;
  const unsigned int ix = (unsigned int)__float_as_int(x);          // callers pass 0 < x < 1 (normal numbers)
  if (ix == 0x3f800000u) return 0.0f;
  const unsigned int tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const int k = (int)tmp >> 23;
  const unsigned int iz = ix - (tmp & 0xff800000u);
  const double z = (double)__int_as_float((int)iz);
  const double r = z * T[i][0] - 1.0;
  const double y0 = T[i][1] + (double)k * Ln2;
  const double r2 = r * r;
  double y = A1 * r + A2;
  y = A0 * r2 + y;
  y = y * r2 + (y0 + r);
  return (float)y;
}
__device__ __forceinline__ double ref_ran(RefRng *r) {
  const double am = 0x1.fffffep-1 / 2147483647.0;                   // nearest(1.0,-1.0)/real(im,dp)
  r->calls++;
  if (r->idum <= 0 || r->iy < 0) {
    const int a = abs(r->idum);
    r->iy = (888889999 ^ a) | 1;
    r->ix = 777755555 ^ a;
    r->idum = a + 1;
  }
  unsigned int x = (unsigned int)r->ix;
  x ^= x << 13; x ^= x >> 17; x ^= x << 5;
  r->ix = (int)x;
  const int k = r->iy / 127773;
  r->iy = 16807 * (r->iy - k * 127773) - 2836 * k;
  if (r->iy < 0) r->iy += 2147483647;
  return am * (double)((2147483647 & (r->ix ^ r->iy)) | 1);
}
__device__ __forceinline__ double ref_gasdev(RefRng *r) {
  if (r->stored) { r->stored = 0; return r->g; }
  double a, b; float s;
  do {
    a = 2.0 * ref_ran(r) - 1.0;
    b = 2.0 * ref_ran(r) - 1.0;
    s = (float)(a * a + b * b);
  } while (!((double)s > 0.0 && (double)s < 1.0));
  const float f = (float)sqrt(-2.0 * (double)ref_logf(s) / (double)s);
  r->g = b * (double)f; r->stored = 1;
  return a * (double)f;
}

// ---- device-resident scalars (one struct in global memory; the host mirrors it on demand) --------------
struct DevScal {
  double z0, z1, zmax, rho, rho0;
  double d1, d2;                 // two largest squared displacements (inq_dispmax, Neighbor.F90:635-666)
  double max_vel, msd_t, msd_max;
  long long try_, depo, choques, choques3;
  long long list_entries;
  long long gcmc_created, gcmc_destroyed, row_overflow;
  int rho_count;
  int need_rebuild;
  int again;                     // overlap_moveback: another pass needed
  int any_active;
  int err;                       // first device-side error (DML_E_*)
  int n_involved, n_roots, member_cursor;
  int n_slots;                   // hs%amax (device copy; gcmc may grow it)
  int nat_sys, nat_ref, nat_gcmc, nlimbo;
  int cols_used;                 // bump pointer into the tail region of cols[] (rows longer than ROW_W, gcmc rows)
  int cols_tail0;                // first entry of the tail region = capacity * ROW_W
  int next_uid;
  unsigned int ticket;
  int halo_flag;                 // some particle sits in a halo cell (rows may be asymmetric, see k_fuerza)
  int rev_used;
  int glen, ghead, gtomb, b_amax;   // gcmc membership array (list order) and hs%b%amax
  double maxz_fac;                  // sum of |lohi| applied by maxz since the last test_update (z-dependent piston shift bound)
  int lay_cur;                      // which of the two z-layer displacement tables is current
  double dsum_tu;                   // sqrt(d1)+sqrt(d2) of the last test_update (0 right after a rebuild)
  double maxz_disp;                 // bound of the displacement applied by maxz since the last test_update
  unsigned int step_disp_bits;      // float bits: largest |pos-old_cg| of the last integrator call
  int rows_pending;                 // rows of the last rebuild not materialised yet (lazy build)
  unsigned int ticket2;
  int rows_asym, rev_valid;         // 0 symmetric rows, 1 asymmetric only through halo-cell particles, 2 general; transposed rows current
  int listed;                       // hs%listed (Neighbor.F90:53)
  int cols_cap;                     // capacity of cols[] / rev_cols[]
  long long nupd;                   // nupd_vlist (Neighbor.F90:110)
  long long choques2, ch_later;     // dana.F90:941 bookkeeping
  long long overlap_passes;
  unsigned int ticket3;             // last-block election of k_pbc_disp
  unsigned int ticket4;             // last-block election of k_integrate
  double kappa_tu;                  // |1 - 1/pist_P| the displacement tables of the last test_update were recorded with (0 after a rebuild)
  double pist_P;                    // product of (1 - lohi) of the maxz calls since the rows were built (1 without a piston): see lay_note
  RefRng rr; long long rr_mark;     // DML_RNG_REFERENCE: the reference's generator state; try_ when the running overlap_moveback started
  unsigned int istep;               // integrator calls so far: the step word of the Philox counters (kernels read it here so that a captured
                                    // CUDA graph of the loop body stays valid from step to step)
  int sort_pending;                 // a rebuild snapshotted the positions but the cell sort was left to whoever needs it first (dml_coop.cuh)
  int tu_par, rho_cnt2[2], dref_cnt2[2];   // census / promotion counters of the fused step tail, double-buffered by call parity
  int hole_lo, bhole_lo;            // gcmc index reuse: no empty hs slot / free b index below these (reset when a rebuild frees the limbo slots)
};

enum { DML_E_OUT_OF_TESS = 1, DML_E_SUPERO_Z0 = 2, DML_E_ROW_OVERFLOW = 3, DML_E_CAPACITY = 4, DML_E_NO_PARTICLES = 5,
       DML_E_COLS_OVERFLOW = 6, DML_E_GCMC_CHOSEN = 7, DML_E_REPLAY_EXHAUSTED = 8 };

// ---- geometry parameters passed by value ----------------------------------------------------------------
struct Geo {
  double box[3], one_box[3], half_box[3];
  double cell[3];
  int nc[3];      // ncells
  int hd[3];      // nc+2 (halo-inclusive extents, Cells.F90:248)
  int pbc[3];
  double rc_list2;   // (rcut+nb_dcut)^2
  float band2;                       // half-width (in distance^2) of the band around rc_list^2 where fp32 cannot decide (k_rows)
  int lay_shift, nlay;               // z-layer displacement table: layer = cell_z >> lay_shift
  double bq_scale;   // 255/(rcut+nb_dcut): quantisation of build-time distances (dml_kernels.cuh, k_rows)
  double rcut2;      // rcut^2
  double inv_cell2;  // 1/cell[2] (z-layer lookup only, see layer_of)
  int rows_fast;     // every axis has >= 3 cells: d_rows may merge the x-neighbours of a stencil row (dml_kernels.cuh)
};

// idnint(x) for the minimum image, bit-identical to round(): particles live inside the box, so |x| = |vd/box| < 1.5 and
// the result is one of -1, 0, +1, decided by two comparisons (half away from zero: x>=0.5 -> 1, x<=-0.5 -> -1).
// The software fp64 round() is only taken outside that range (never for in-box pairs), keeping the result exact in general.
__device__ __forceinline__ double mic_axis(double v, double box, double one_box) {
  const double x = v * one_box;
  double k = (x >= 0.5 ? 1.0 : 0.0) - (x <= -0.5 ? 1.0 : 0.0);
  if (fabs(x) >= 1.5) k = round(x);
  return v - box * k;
}
// vdistance (Groups.F90:995-1016): a minus b, idnint minimum image on periodic axes, |.|^2 = (x²+y²)+z²
__device__ __forceinline__ double dist2_idnint(const Geo &g, double ax, double ay, double az, double bx, double by, double bz) {
  double vx = ax - bx, vy = ay - by, vz = az - bz;
  if (g.pbc[0]) vx = mic_axis(vx, g.box[0], g.one_box[0]);
  if (g.pbc[1]) vy = mic_axis(vy, g.box[1], g.one_box[1]);
  if (g.pbc[2]) vz = mic_axis(vz, g.box[2], g.one_box[2]);
  return (vx * vx + vy * vy) + vz * vz;
}

// cell index triple int(pos/cell)+1 (Cells.F90:289); returns false when outside 0..nc+1
__device__ __forceinline__ bool cell_index(const Geo &g, double x, double y, double z, int &cx, int &cy, int &cz) {
  cx = (int)(x / g.cell[0]) + 1; cy = (int)(y / g.cell[1]) + 1; cz = (int)(z / g.cell[2]) + 1;
  return !(cx < 0 || cy < 0 || cz < 0 || cx > g.nc[0] + 1 || cy > g.nc[1] + 1 || cz > g.nc[2] + 1);
}
__device__ __forceinline__ int cell_lin(const Geo &g, int cx, int cy, int cz) { return cx + g.hd[0] * (cy + g.hd[1] * cz); }

// ---- libgcc __powidf2 order for x**6, x**7 (gfortran -O0) ---------------------------------
__device__ __forceinline__ double pow6(double x) { double x2 = x * x; double x4 = x2 * x2; return x2 * x4; }
__device__ __forceinline__ double pow7(double x) { double x2 = x * x; double y = x * x2; double x4 = x2 * x2; return y * x4; }

// ---- Philox4x32-10 counter-based RNG ----------------------------------------------------------------------
struct Philox {
  uint32_t c[4], k[2];
  __device__ __forceinline__ void round_() {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ __forceinline__ void run(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    c[0] = c0; c[1] = c1; c[2] = c2; c[3] = c3; k[0] = (uint32_t)seed; k[1] = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) { round_(); k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
  }
  // two uniforms in (0,1) with 53-bit resolution
  __device__ __forceinline__ double u01(int i) const {
    uint64_t b = ((uint64_t)c[2 * i] << 32) | c[2 * i + 1];
    return ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  // two independent standard normals in fp64 (gcmc_run's velocity draw: three per accepted insertion)
  __device__ __forceinline__ void gauss2(double &g0, double &g1) const {
    double u = u01(0), v = u01(1);
    double r = sqrt(-2.0 * log(u));
    double s, c_; sincospi(2.0 * v, &s, &c_);
    g0 = r * c_; g1 = r * s;
  }
  // Four independent standard normals from the four 32-bit words (Box-Muller in single precision, results widened to fp64).
  // The reference's own gasdev works at this resolution (its radius, logarithm and factor are real(sp): logf and a float sqrt,
  // dana.F90:1406-1428).  The fp64 form (log + sincospi in double, one Philox block per pair) made the Ermak half-step
  // instruction-bound: 33 M warp instructions per million particles, 69 us where its 164 bytes per particle stream in ~30 us.
  __device__ __forceinline__ void gauss4f(double &g0, double &g1, double &g2, double &g3) const {
    const float k = 2.3283064365386963e-10f;                 // 2^-32
    const float u0 = __fmaf_rn(__uint2float_rz(c[0]), k, 1.1641532182693481e-10f), v0 = __uint2float_rz(c[1]) * k;   // u in (0,1], v in [0,1)
    const float u1 = __fmaf_rn(__uint2float_rz(c[2]), k, 1.1641532182693481e-10f), v1 = __uint2float_rz(c[3]) * k;
    const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u1));
    float s0, c0, s1, c1;
    sincospif(2.0f * v0, &s0, &c0); sincospif(2.0f * v1, &s1, &c1);
    g0 = (double)(r0 * c0); g1 = (double)(r0 * s0); g2 = (double)(r1 * c1); g3 = (double)(r1 * s1);
  }
};

// streams of the counter's third word
enum { RS_INTEG0 = 0, RS_INTEG1 = 1, RS_INTEG2 = 2, RS_PBC = 3, RS_OVERLAP = 4, RS_GCMC = 5 };

// ---- top-2 merge (order independent) -----------------------------------------------------------------------
__device__ __forceinline__ void top2_merge(double &a1, double &a2, double b1, double b2) {
  double m1 = fmax(a1, b1);
  double m2 = fmax(fmin(a1, b1), fmax(a2, b2));
  a1 = m1; a2 = m2;
}

__constant__ int c_map[27][3] = {   // Cells.F90:28-36, stencil order fixes the order of every row
  {0,0,0},{1,0,0},{1,1,0},{0,1,0},{-1,1,0},{1,0,-1},{1,1,-1},{0,1,-1},{-1,1,-1},
  {1,0,1},{1,1,1},{0,1,1},{-1,1,1},{0,0,1},{-1,0,0},{-1,-1,0},{0,-1,0},{1,-1,0},
  {-1,0,1},{-1,-1,1},{0,-1,1},{1,-1,1},{-1,0,-1},{-1,-1,-1},{0,-1,-1},{1,-1,-1},{0,0,-1}};

// The same stencil for "one lane per stencil cell" code: constant memory would serialise a warp whose lanes index
// different entries, so lane l decodes its own offset from two 64-bit literals (2 bits per component, value+1).
__device__ __forceinline__ void map_of_lane(int l, int &dx, int &dy, int &dz) {
  // entry e = (dx+1) | (dy+1)<<2 | (dz+1)<<4, 6 bits each; entries 0-9 in lo, 10-19 in mid, 20-26 in hi
  const unsigned long long lo = 0ull
    | (unsigned long long)(1 | 1 << 2 | 1 << 4) << 0   | (unsigned long long)(2 | 1 << 2 | 1 << 4) << 6
    | (unsigned long long)(2 | 2 << 2 | 1 << 4) << 12  | (unsigned long long)(1 | 2 << 2 | 1 << 4) << 18
    | (unsigned long long)(0 | 2 << 2 | 1 << 4) << 24  | (unsigned long long)(2 | 1 << 2 | 0 << 4) << 30
    | (unsigned long long)(2 | 2 << 2 | 0 << 4) << 36  | (unsigned long long)(1 | 2 << 2 | 0 << 4) << 42
    | (unsigned long long)(0 | 2 << 2 | 0 << 4) << 48  | (unsigned long long)(2 | 1 << 2 | 2 << 4) << 54;
  const unsigned long long mid = 0ull
    | (unsigned long long)(2 | 2 << 2 | 2 << 4) << 0   | (unsigned long long)(1 | 2 << 2 | 2 << 4) << 6
    | (unsigned long long)(0 | 2 << 2 | 2 << 4) << 12  | (unsigned long long)(1 | 1 << 2 | 2 << 4) << 18
    | (unsigned long long)(0 | 1 << 2 | 1 << 4) << 24  | (unsigned long long)(0 | 0 << 2 | 1 << 4) << 30
    | (unsigned long long)(1 | 0 << 2 | 1 << 4) << 36  | (unsigned long long)(2 | 0 << 2 | 1 << 4) << 42
    | (unsigned long long)(0 | 1 << 2 | 2 << 4) << 48  | (unsigned long long)(0 | 0 << 2 | 2 << 4) << 54;
  const unsigned long long hi = 0ull
    | (unsigned long long)(1 | 0 << 2 | 2 << 4) << 0   | (unsigned long long)(2 | 0 << 2 | 2 << 4) << 6
    | (unsigned long long)(0 | 1 << 2 | 0 << 4) << 12  | (unsigned long long)(0 | 0 << 2 | 0 << 4) << 18
    | (unsigned long long)(1 | 0 << 2 | 0 << 4) << 24  | (unsigned long long)(2 | 0 << 2 | 0 << 4) << 30
    | (unsigned long long)(1 | 1 << 2 | 0 << 4) << 36;
  unsigned long long w = l < 10 ? lo : (l < 20 ? mid : hi);
  int sh = (l < 10 ? l : (l < 20 ? l - 10 : l - 20)) * 6;
  int e = (int)(w >> sh) & 63;
  dx = (e & 3) - 1; dy = ((e >> 2) & 3) - 1; dz = ((e >> 4) & 3) - 1;
}

// pair tables (dana.F90:87-100) and integrator constants, set per ctx before launches
struct Phys {
  double eps[9], r0[9], r0sq[9], r0p6[9];
  double r0sq_max, r0_max;                            // largest cut-off of the table, squared and plain (cheap first test, gather skip)
  double mass[3], sqrt_mass[3];
  double h, prob, tau;
  double cc0, cc1, cc2, sdr, sdv, crv1, crv2, skt;    // set_ermak, dana.F90:947-971
  double cc1mcc2, cc2h;
  double dif_sc, dif_sei, z_sei, fac_sc, fac_sei;     // cbrownian_hs, dana.F90:817-823
  int integrador, piston, chunks, rng_mode;
  unsigned long long seed;
};

} // namespace dml
