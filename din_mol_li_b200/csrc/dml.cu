// dml.cu — C-ABI entry points of libdml.so (declared in include/dml.h) and the host orchestration of
// the kernels in dml_kernels.cuh.  No CPU fallback: every entry point needs a CUDA device.
#include "../../include/dml.h"
#include "../../include/dml_host.h"
#include "dml_kernels.cuh"
#include "dml_coop.cuh"
#include "dml_slab.cuh"
#include "dml_observe.cuh"
namespace dml { __global__ void k_gcmc_tomb(const double4 *__restrict__ posm, int *__restrict__ gorder, const int *__restrict__ gpos, DevScal *__restrict__ sc, int n); }
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <functional>
#include <chrono>

using namespace dml;

namespace {

struct ProfEv { cudaEvent_t a, b; int cls; };

template <typename T>
struct DBuf {
  T *p = nullptr; size_t cap = 0;
  cudaError_t ensure(size_t n, cudaStream_t st, bool keep = false) {
    if (n <= cap) return cudaSuccess;
    size_t nc = std::max(n, cap + cap / 2);
    T *np_ = nullptr;
    cudaError_t e = cudaMalloc(&np_, nc * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep && p && cap) cudaMemcpyAsync(np_, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
    if (p) { cudaStreamSynchronize(st); cudaFree(p); }
    p = np_; cap = nc;
    return cudaSuccess;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

} // namespace

struct dml_ctx {
  dml_config cfg;
  std::string err;
  cudaStream_t st = nullptr;
  int cap = 0, n = 0;
  Geo geo; Phys ph;
  bool tessellated = false, listed = false, cells_sorted = false, binned = false;
  int nct = 0;
  int64_t nupd = 0, step = 0, choques2 = 0, overlap_passes = 0;
  double t = 0.0;
  int64_t launches = 0;
  bool profiling = false, capturing = false; int prof_only = -1; std::vector<ProfEv> evs; std::vector<ProfEv> pool;
  double prof_ms[32] = {0}; int64_t prof_n[32] = {0};
  // particle state
  DBuf<float4> sorted_posf;                            // single-precision copy of the cell-sorted records (k_rows prefilter)
  DBuf<double4> posm, sorted_posm, fe;                 // fe = {force(3), epot}
  DBuf<double> vel, acel, pos_old, old_cg, ranv;
  DBuf<int> uid, slot_b;
  // cells
  DBuf<int> b2slot;          // boxes without cell lists (ngroup_verlet): slot of every hs%b index
  DBuf<int> cell_of, cell_cnt, cell_start, cell_cur, sorted_slot, sorted_raw, sorted_cell, chain_pos;   // sorted_raw: scatter output (in-cell order arbitrary)
  // rows
  DBuf<RowHead> rh; DBuf<int> cols; DBuf<unsigned char> bq, rev_bq, halo_of, fnz, dq, kb; DBuf<unsigned int> lay;   // bq: quantised build-time distance per entry
  DBuf<int> rev_start, rev_len, rev_cur, rev_cols; bool rows_asym = false; bool rev_valid = false;
  // slab decomposition (dml_slab.cuh)
  ncclComm_t comm = nullptr; int rank = 0, nranks = 1, n_owned = 0;
  DBuf<int> send_lo, send_hi, slab_counts, pack_uid_lo, pack_uid_hi; DBuf<double4> pack_lo, pack_hi;
  double zlo = 0.0, zhi = 0.0; int slab_holes = 0; bool slab_ready = false;     // slab bounds, holes in the owned region
  DBuf<int> mig_list_lo, mig_list_hi, mig_rc, mig_holes, mig_si_lo, mig_si_hi, mig_ri;
  DBuf<double> mig_sd_lo, mig_sd_hi, mig_rd, top2_own, top2_all;
  int nsend_lo = 0, nsend_hi = 0, nrecv_lo = 0, nrecv_hi = 0, ghost_lo_first = 0, ghost_hi_first = 0;
  bool rows_legacy = false; // DML_ROWS_LEGACY=1: 27-cell ordered walk of one thread for every row (the staged walk needs >= 3 cells per axis)
  int coop_max_n = 65536;   // persistent cooperative kernels pay off while launch latency dominates (overlap_moveback)
  int coop_tu_max_n = 4194304;  // test_update is a chain of short data-dependent phases, most of them idle when no rebuild is due: the
                                // one-launch form wins at every size measured (100 k: 0.312 -> 0.300 ms/step, 1 M: 0.449 -> 0.405 ms/step)
  int tu_fused = 0;         // what the last test_update launch folded in (bit 0 k_ov_init, bit 1 k_ov_apply, bit 2 the tail of the step)
  // CUDA graph of part a of the loop body (enq_step_a): every launch is unconditional with device-side guards, so the captured
  // sequence stays valid until the slot count or the geometry changes; replaying it removes the ~6 us host / front-end gap in front
  // of each of the ~11 launches of a step
  std::vector<ProfEv> sg_evs; bool sg_ran = false;   // event pairs recorded inside the captured step (dml_profile(2 + kid)) and whether it ran since they were read
  cudaGraphExec_t step_graph = nullptr; int sg_n = -1, sg_nct = -1; Geo sg_geo; int64_t sg_launches = 0; bool use_graph = true;
  bool step_tail_done = false;       // enq_step_a folded the tail of the step into the second test_update
  bool sort_maybe_pending = false;   // a deferring test_update was enqueued since the last cell sort (k_sort_catchup is launched on demand)
  DBuf<double4> snap;       // positions of a rebuild whose cell sort was deferred
  bool no_flat_b = false;   // DML_NO_FLAT_B=1: ermak_b always reads the records (k_ermak_b)
  bool no_bi_fuse = false;  // DML_NO_BI_FUSE=1: the Brownian integrator stays a launch of its own inside dml_step
  bool no_tu_fuse = false;  // DML_NO_TU_FUSE=1: keep k_ov_init / k_ov_apply as launches of their own inside dml_step
  int l2_slots = 0; long long l2_max_persist = -1, l2_max_window = 0; bool no_l2_persist = false;   // slots covered by the persisting-L2 window (l2_window)
  bool use_coop = true; int coop_grid_tu = 0, coop_grid_tu_bi = 0, coop_grid_ov = 0, coop_grid_rev = 0; DBuf<int> coop_sums;   // persistent cooperative kernels (dml_coop.cuh)
  bool ov_unstaged = false; // DML_OV_UNSTAGED=1: k_ov_resolve replays from global memory (the form the cooperative kernel uses)
  // decomposed step: the two host read-backs of a step cut it into two segments, each captured as a CUDA graph (NCCL calls included)
  // and replayed until a rebuild changes the slab (slot counts, ghost lists): see slab_segment
  struct SlabGraph { cudaGraphExec_t exec = nullptr; bool warm = false; int64_t launches = 0, steps = 0; } sgA, sgB;
  bool slab_graph_on = true; Geo slab_geo; int slab_since = 0, slab_last_interval = 0;   // steps since the last rebuild, length of the interval before it
  bool kb_valid = false;    // kb[] (element / ref byte per slot) was written by the pair-force call that has just been enqueued
  bool use_dq = false;      // per-particle refinement of the gather-skip bound (dq_byte); DML_NO_DQ=1: layer bound only
  bool rows_eager = false;  // inside dml_slab_step: the consumers' guarded row-build launches are left out
  int ov_res_bpsm = 8;      // blocks of 4 warps per SM of k_ov_resolve (one warp per conflict component; DML_OV_RES_BPSM)
  int ov_lanes = 0;         // threads per particle of the overlap detection (DML_OV_LANES: 1, 2, 4; 0 = by integrator)
  int force_minb = 4;       // resident blocks per SM the production pair-force kernel is compiled for (DML_FORCE_MINB: 4, 6, 8)
  bool fuse_ermak_b = false; // DML_FUSE_ERMAK_B=1: dml_step applies ermak_b inside the production pair-force kernel
  int ov_guard_pass = 64;   // from this pass on, pairs that overlap at their previous positions are skipped in every mode
  DBuf<int> scan_sums; DBuf<unsigned long long> scan_state; unsigned int *scan_tickets = nullptr; unsigned int scan_epoch = 0;
  DBuf<int> rev_cnt;
  DBuf<double> part;
  // overlap
  DBuf<int> parent, ovst, comp_cnt, comp_off, members, roots, ov_head, ov_next;
  // gcmc
  DBuf<int> gorder, gpos, gcc, gpend, b_occ; int gorder_cap = 0;
  // replay
  DBuf<double> rp_gauss, rp_upbc, rp_uovl, rp_gu, rp_gg; bool have_rp = false, have_rp_ovl = false; int rp_nu = 0, rp_ng = 0;
  DBuf<int> ord;            // DML_RNG_REFERENCE: slot of every creation rank (hs%ref in list order)
  DBuf<int> rp_qstart, ov_draws;   // overlap_moveback replay queue: draws k = 0,1,.. of slot s read rp_uovl[rp_qstart[s] + k]
  // output reductions and observables (dml_observe.cuh)
  DBuf<double> obs_part, obs_out; DBuf<unsigned long long> obs_counts; DBuf<int> gr_cell_of, gr_cnt, gr_start; DBuf<double4> gr_sorted;
  unsigned int *obs_ticket = nullptr;
  DBuf<int> snap_uid; DBuf<unsigned char> snap_mb; DBuf<int4> mc_out; DBuf<int> mc_count; bool have_snap = false;   // dml_membership_changes
  // staging
  DBuf<double> stage_d, stage_f; DBuf<int> stage_i;
  DevScal *sc = nullptr; DevScal *hsc = nullptr;    // device / pinned host mirror
  // chunk template (reservoir 2)
  std::vector<double> ch_pos, ch_pos_old; double ch_dist = 0, ch_rhomedia = 0; bool have_chunk = false;
  int row_slack = 0;
};

#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return -1; } } while (0)
#define FAIL(msg) do { ctx->err = (msg); return -1; } while (0)
#define TRY(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)
// every public entry point runs on the ctx's own device, whatever the calling thread's current device is
#define ENTER(ctx) do { if (!(ctx)) return -1; cudaSetDevice((ctx)->cfg.device); } while (0)

static int gcmc_run_impl(dml_ctx *ctx);
static int enq_build_rev(dml_ctx *ctx);
static int enq_sort_cells(dml_ctx *ctx, int force);
static int enq_materialize_rows(dml_ctx *ctx, bool force = false);
static void fill_tu_args(dml_ctx *ctx, TUArgs &A, int force);
static int finish(dml_ctx *ctx);
static int pull_scal(dml_ctx *ctx);
static void slab_graphs_drop(dml_ctx *ctx) {
  for (auto *G : {&ctx->sgA, &ctx->sgB}) { if (G->exec) cudaGraphExecDestroy(G->exec); G->exec = nullptr; G->warm = false; }
}

enum { CLS_FORCE = 0, CLS_LIST = 1, CLS_INTEG = 2, CLS_OVERLAP = 3, CLS_ALL = 4, CLS_BIN = 5, CLS_OTHER = 6, CLS_GCMC = 7 };
// one id per kernel so bench.py can time each of them with CUDA events on the ctx stream
enum { K_SCAN = 0, K_PBC_BIN, K_TOP2, K_SCATTER, K_CELL_ORDER, K_ROWS_COUNT, K_ROWS_FILL, K_OV_LINK, K_FUERZA, K_INTEGRATE,
       K_ERMAK_B, K_OV_INIT, K_OV_DETECT, K_OV_COUNT, K_OV_ALLOC, K_OV_FILL, K_OV_SORT, K_OV_PASS, K_OV_APPLY, K_PROMOTE,
       K_CALC_RHO, K_MAXZ, K_PACK, K_MISC, K_GCMC, K_REV, K_BIN, K_TU_COOP, K_OV_COOP, K_NKERN };
static const char *const kern_name[K_NKERN] = {"scan", "pbc_disp", "top2_final", "scatter", "cell_order", "rows_count", "rows_build",
  "ov_link", "fuerza", "integrate", "ermak_b", "ov_init", "ov_detect", "ov_count", "ov_alloc", "ov_fill", "ov_sort", "ov_pass",
  "ov_apply", "promote", "calc_rho", "maxz", "pack", "misc", "gcmc", "rev_rows", "bin", "test_update_coop", "overlap_coop"};
static const int kern_cls[K_NKERN] = {CLS_LIST, CLS_BIN, CLS_BIN, CLS_LIST, CLS_LIST, CLS_LIST, CLS_LIST, CLS_OVERLAP, CLS_FORCE, CLS_INTEG,
  CLS_INTEG, CLS_OVERLAP, CLS_OVERLAP, CLS_OVERLAP, CLS_OVERLAP, CLS_OVERLAP, CLS_OVERLAP, CLS_OVERLAP, CLS_OVERLAP, CLS_OTHER,
  CLS_OTHER, CLS_OTHER, CLS_OTHER, CLS_OTHER, CLS_GCMC, CLS_LIST, CLS_LIST, CLS_LIST, CLS_OVERLAP};

static void prof_begin(dml_ctx *ctx, int cls) {
  ctx->launches++;
  if (!ctx->profiling || (ctx->prof_only >= 0 && ctx->prof_only != cls)) return;
  ProfEv ev;
  if (!ctx->pool.empty()) { ev = ctx->pool.back(); ctx->pool.pop_back(); }
  else { cudaEventCreate(&ev.a); cudaEventCreate(&ev.b); }
  ev.cls = cls;
  // inside a stream capture a plain record is only a dependency marker: an external record node takes the timestamp at replay
  if (ctx->capturing) cudaEventRecordWithFlags(ev.a, ctx->st, cudaEventRecordExternal); else cudaEventRecord(ev.a, ctx->st);
  ctx->evs.push_back(ev);
}
static void prof_end(dml_ctx *ctx, int cls) {
  if (!ctx->profiling || (ctx->prof_only >= 0 && ctx->prof_only != cls)) return;
  if (ctx->capturing) cudaEventRecordWithFlags(ctx->evs.back().b, ctx->st, cudaEventRecordExternal); else cudaEventRecord(ctx->evs.back().b, ctx->st);
}
static void prof_collect(dml_ctx *ctx) {
  if (ctx->evs.empty()) return;
  cudaStreamSynchronize(ctx->st);
  for (auto &ev : ctx->evs) {
    float ms = 0; cudaEventElapsedTime(&ms, ev.a, ev.b);
    ctx->prof_ms[ev.cls] += ms; ctx->prof_n[ev.cls]++;
    ctx->pool.push_back(ev);
  }
  ctx->evs.clear();
}
// a launch that the driver refuses (too many resources, bad configuration) must not pass silently
#define LAUNCH(cls, kern, grid, block, ...) do { prof_begin(ctx, cls); kern<<<(grid), (block), 0, ctx->st>>>(__VA_ARGS__); prof_end(ctx, cls); \
  cudaError_t le_ = cudaPeekAtLastError(); if (le_ != cudaSuccess) { cudaGetLastError(); ctx->err = std::string("launch of " #kern ": ") + cudaGetErrorString(le_); return -1; } } while (0)

#define LAUNCH_COOP(kid, kern, grid, argstruct) do { prof_begin(ctx, kid); void *a_[] = {(void *)&(argstruct)}; \
  cudaError_t e_ = cudaLaunchCooperativeKernel((void *)kern, dim3(grid), dim3(TPB), a_, 0, ctx->st); prof_end(ctx, kid); \
  if (e_ != cudaSuccess) { ctx->err = std::string("cooperative launch of " #kern ": ") + cudaGetErrorString(e_); return -1; } } while (0)

static inline int nblk(int n, int b = TPB) { return std::max(1, (n + b - 1) / b); }

static const char *dev_err_msg(int e) {
  switch (e) {
    case DML_E_OUT_OF_TESS: return "Particle out of tessellation";
    case DML_E_SUPERO_Z0: return "supero z0";
    case DML_E_ROW_OVERFLOW: return "neighbour row overflow";
    case DML_E_CAPACITY: return "slot capacity exhausted";
    case DML_E_NO_PARTICLES: return "No more particles";
    case DML_E_COLS_OVERFLOW: return "neighbour storage exhausted";
    case DML_E_GCMC_CHOSEN: return "Chosen particle does not exists";
    case DML_E_REPLAY_EXHAUSTED: return "replay stream exhausted";
  }
  return "unknown device error";
}

static int pull_scal(dml_ctx *ctx) {
  CKC(cudaMemcpyAsync(ctx->hsc, ctx->sc, sizeof(DevScal), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));
  if (ctx->hsc->err) { ctx->err = dev_err_msg(ctx->hsc->err); return -ctx->hsc->err - 100; }
  return 0;
}
static int push_scal(dml_ctx *ctx) {
  CKC(cudaMemcpyAsync(ctx->sc, ctx->hsc, sizeof(DevScal), cudaMemcpyHostToDevice, ctx->st));
  return 0;
}

static void set_box(dml_ctx *ctx, const double box[3]) {
  for (int k = 0; k < 3; ++k) { ctx->geo.box[k] = box[k]; ctx->geo.one_box[k] = 1.0 / box[k]; ctx->geo.half_box[k] = box[k] * .5; }
  {
    // fp32 error band of k_rows: coordinates up to L (+ one list radius outside the box) carry ulp(L)/2 each, the image shift
    // another ulp; |d(d^2)| <= 2*sqrt(3)*rc*err + rounding of the products.  A factor 4 of slack on top.
    double L = std::max(std::max(box[0], box[1]), box[2]) * 2.5 + 64.0;   // k_rows folds the image shift into the particle's coordinate: |p + box| < 2 box
    double ulp = std::ldexp(1.0, (int)std::ceil(std::log2(L)) - 23);
    double rl = ctx->cfg.rcut + ctx->cfg.nb_dcut;
    ctx->geo.band2 = (float)(4.0 * (2.0 * 1.7321 * (rl + 1.0) * 3.0 * ulp + 1e-5 * rl * rl));
  }
}

// Keep the 32-byte particle records resident in the 126 MB L2 across the kernels of a step: every gather of the pair-force /
// overlap / list kernels then hits L2 and HBM only sees the streaming arrays.  The window covers the slots in use
// (plus head-room), not the capacity: a window larger than the persisting carve-out lowers the hit ratio of every record
// (measured at 1 M particles: capacity 2.5 M slots -> pair force 41 -> 55 us, step 0.354 -> 0.428 ms).
static void l2_window(dml_ctx *ctx, int nslots) {
  if (ctx->no_l2_persist) return;
  nslots = std::min(std::max(nslots, 1024), ctx->cap);
  if (ctx->l2_max_persist < 0) {                         // device attributes are read once per context
    int dev = 0, a = 0, b = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&a, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&b, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    ctx->l2_max_persist = a; ctx->l2_max_window = b;
  }
  if (ctx->l2_max_persist <= 0) return;
  if (nslots <= ctx->l2_slots && nslots * 2 > ctx->l2_slots) return;      // the current window already fits
  struct { size_t persistingL2CacheMaxSize, accessPolicyMaxWindowSize; } prop = {(size_t)ctx->l2_max_persist, (size_t)ctx->l2_max_window};
  size_t bytes = (size_t)nslots * sizeof(double4);
  size_t want = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, bytes);
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
  cudaStreamAttrValue attr; memset(&attr, 0, sizeof attr);
  attr.accessPolicyWindow.base_ptr = ctx->posm.p;
  attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
  attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)attr.accessPolicyWindow.num_bytes);
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cudaStreamSetAttribute(ctx->st, cudaStreamAttributeAccessPolicyWindow, &attr);
  cudaGetLastError();
  ctx->l2_slots = nslots;
}

// cgroup_tessellate — Cells.F90:180-265 (host side: it depends only on the box and rcut+nb_dcut)
static void tessellate(dml_ctx *ctx) {
  Geo &g = ctx->geo;
  double rc = ctx->cfg.rcut + ctx->cfg.nb_dcut;
  if (ctx->tessellated) {
    bool ok1 = true, ok2 = true;
    for (int k = 0; k < 3; ++k) if (!((double)(g.nc[k] + 1) >= g.box[k] / rc)) ok1 = false;
    if (ok1) {
      for (int k = 0; k < 3; ++k) if (!(rc < g.box[k] / (double)g.nc[k])) ok2 = false;
      if (ok2) { for (int k = 0; k < 3; ++k) g.cell[k] = g.box[k] / (double)g.nc[k]; g.inv_cell2 = 1.0 / g.cell[2]; return; }
    }
  }
  int nc[3];
  for (int k = 0; k < 3; ++k) nc[k] = (int)(g.box[k] / rc);
  for (int k = 0; k < 3; ++k) g.nc[k] = nc[k];
  if (nc[0] < 4 && nc[1] < 4 && nc[2] < 4) {             // no cells: the reference falls back to the O(N^2) list (ngroup_verlet)
    // the z-layer displacement tables of the gather skip still want a layer thickness: one list radius
    for (int k = 0; k < 3; ++k) { g.nc[k] = std::max(nc[k], 1); g.cell[k] = g.box[k] / (double)g.nc[k]; g.hd[k] = g.nc[k] + 2; }
    g.inv_cell2 = 1.0 / g.cell[2];
    ctx->nct = g.hd[0] * g.hd[1] * g.hd[2];
    g.rows_fast = 0; g.lay_shift = 0; g.nlay = g.nc[2] + 2;
    ctx->tessellated = false;
    return;
  }
  for (int k = 0; k < 3; ++k) { g.cell[k] = g.box[k] / (double)nc[k]; g.hd[k] = nc[k] + 2; }
  g.inv_cell2 = 1.0 / g.cell[2];
  ctx->nct = g.hd[0] * g.hd[1] * g.hd[2];
  g.rows_fast = (nc[0] >= 3 && nc[1] >= 3 && nc[2] >= 3 && !ctx->rows_legacy) ? 1 : 0;
  g.lay_shift = 0; while (((g.nc[2] + 2) >> g.lay_shift) + 1 > LAY_MAX) g.lay_shift++;
  g.nlay = ((g.nc[2] + 1) >> g.lay_shift) + 1;
  ctx->tessellated = true;
}

// single-pass scan (decoupled look-back).  guard_mode: 0 rebuild guard (| force), 1 transposed-rows guard, 2 always
static int scan_excl(dml_ctx *ctx, int *in, int *out, int n, int *total_out, bool zero_in, int guard_mode, int force) {
  int nb = nblk(n, 1024);
  if ((size_t)nb + 8 > ctx->scan_state.cap) {
    // fresh (or grown) state must not carry epochs of an earlier context that owned the same memory
    CKC(ctx->scan_state.ensure((size_t)nb + 1024, ctx->st));
    CKC(cudaMemsetAsync(ctx->scan_state.p, 0, ctx->scan_state.cap * sizeof(unsigned long long), ctx->st));
  }
  if (!ctx->scan_tickets) { CKC(cudaMalloc(&ctx->scan_tickets, 2 * sizeof(unsigned int))); CKC(cudaMemsetAsync(ctx->scan_tickets, 0, 2 * sizeof(unsigned int), ctx->st)); }
  unsigned int ep = ++ctx->scan_epoch;
  if ((ep & 0x3fffffffu) == 0) ep = ++ctx->scan_epoch;
  if (zero_in) LAUNCH(K_SCAN, (k_scan<true>), nb, TPB, in, out, n, ctx->scan_state.p, ctx->scan_tickets, ep, total_out, ctx->sc, force, guard_mode);
  else LAUNCH(K_SCAN, (k_scan<false>), nb, TPB, in, out, n, ctx->scan_state.p, ctx->scan_tickets, ep, total_out, ctx->sc, force, guard_mode);
  return 0;
}

static int ensure_particles(dml_ctx *ctx, int n) {
  if (n > ctx->cap) FAIL("slot capacity exhausted (dml_config.capacity)");
  return 0;
}

// cell binning + counting sort; guarded on the device by need_rebuild | force
static int enq_sort_cells(dml_ctx *ctx, int force) {
  int n = ctx->n, nct = ctx->nct;
  LAUNCH(K_BIN, k_bin, std::min(nblk(n), 148 * 8), TPB, ctx->posm.p, ctx->cell_of.p, ctx->cell_cnt.p, ctx->rh.p, ctx->halo_of.p, ctx->sc, ctx->geo, n, force);
  TRY(scan_excl(ctx, ctx->cell_cnt.p, ctx->cell_start.p, nct, ctx->cell_start.p + nct, true, 0, force));
  LAUNCH(K_SCATTER, k_scatter, std::min(nblk(n), 148 * 8), TPB, ctx->posm.p, ctx->pos_old.p, ctx->cell_of.p, ctx->cell_start.p, ctx->cell_cur.p,
         ctx->sorted_raw.p, ctx->sc, n, force, ctx->use_dq ? ctx->dq.p : nullptr);
  LAUNCH(K_CELL_ORDER, k_cell_order, std::min(nblk(n), 148 * 8), TPB, ctx->posm.p, ctx->slot_b.p, ctx->cell_of.p, ctx->cell_start.p, ctx->cell_cur.p, ctx->sorted_raw.p,
         ctx->sorted_slot.p, ctx->sorted_posm.p, ctx->sorted_posf.p, ctx->sorted_cell.p, ctx->sc, nct, force);
  return 0;
}

// ngroup_cells (Neighbor.F90:465-548) from the cell-sorted snapshot of the last rebuild; no-op unless rows are pending.
// Rows are always built on demand (the first consumer after a rebuild): in Brownian mode the list rebuilt by the second
// test_update of a step is superseded by the next step's rebuild before anything reads it.
static int enq_materialize_rows(dml_ctx *ctx, bool force) {
  int n = ctx->n, nct = ctx->nct;
  if (ctx->rows_eager && !force) return 0;                // dml_slab_step builds the rows at the rebuild itself: nothing can be pending here
  if (ctx->sort_maybe_pending && ctx->tessellated) {      // the cell sort a deferring test_update left behind (no-op unless pending)
    TUArgs A;
    fill_tu_args(ctx, A, 0);
    LAUNCH_COOP(K_TU_COOP, k_sort_catchup, ctx->coop_grid_tu, A);
    ctx->sort_maybe_pending = false;
  }
  if (!ctx->tessellated) {                                // ngroup_verlet, one warp per row
    LAUNCH(K_ROWS_FILL, k_rows_verlet, std::min(nblk(n * 32), 148 * 8), TPB, ctx->posm.p, ctx->pos_old.p, ctx->b2slot.p, ctx->rh.p, ctx->cols.p, ctx->bq.p,
           ctx->sc, ctx->geo, n, ctx->row_slack);
    return 0;
  }
  prof_begin(ctx, K_ROWS_FILL);
  k_rows<<<std::min(nblk(n, RB), 148 * 2), RB, ROWS_SMEM, ctx->st>>>(ctx->sorted_posm.p, ctx->sorted_posf.p, ctx->sorted_slot.p, ctx->sorted_cell.p, ctx->cell_start.p,
                                                  ctx->rh.p, ctx->cols.p, ctx->bq.p, ctx->sc, ctx->geo, nct, ctx->row_slack);
  prof_end(ctx, K_ROWS_FILL);
  CKC(cudaPeekAtLastError());
  return 0;
}

// test_update (Neighbor.F90:668-713) enqueued without any host round trip: the rebuild decision is taken by
// k_top2_final on the device and the rebuild kernels (update + ngroup_cells, Neighbor.F90:608-633,465-548) are
// always launched but return immediately when no rebuild is due.
static void fill_tu_args(dml_ctx *ctx, TUArgs &A, int force) {
  int n = ctx->n, nct = ctx->nct;
  A.posm = ctx->posm.p; A.pos_old = ctx->pos_old.p; A.part = ctx->part.p; A.cell_of = ctx->cell_of.p; A.cell_cnt = ctx->cell_cnt.p;
  A.cell_start = ctx->cell_start.p; A.cell_cur = ctx->cell_cur.p; A.sorted_slot = ctx->sorted_slot.p; A.sorted_raw = ctx->sorted_raw.p; A.sorted_cell = ctx->sorted_cell.p; A.sorted_posm = ctx->sorted_posm.p; A.sorted_posf = ctx->sorted_posf.p;
  A.slot_b = ctx->slot_b.p; A.rh = ctx->rh.p; A.cols = ctx->cols.p; A.bq = ctx->bq.p; A.halo_of = ctx->halo_of.p; A.lay = ctx->lay.p;
  A.sums = ctx->coop_sums.p; A.sc = ctx->sc; A.g = ctx->geo; A.n = n; A.nct = nct; A.force_sort = force; A.slack = ctx->row_slack; A.lazy = 1;
  A.nb_dcut = ctx->cfg.nb_dcut; A.rmax_f = ctx->ph.r0_max; A.rmax_o = ctx->cfg.rcut;
  A.fuse = 0; A.parent = ctx->parent.p; A.ovst = ctx->ovst.p; A.comp_cnt = ctx->comp_cnt.p; A.ov_head = ctx->ov_head.p;
  A.vel = ctx->vel.p; A.acel = ctx->acel.p; A.old_cg = ctx->old_cg.p;
  A.area = ctx->geo.box[0] * ctx->geo.box[1]; A.h_over_tau = ctx->cfg.h / ctx->cfg.tau; A.use_z1 = ctx->cfg.reservoir == 2 ? 1 : 0; A.piston = ctx->cfg.reservoir == 1 ? 1 : 0;
  A.defer = 0; A.snap = ctx->snap.p;
  A.uid = ctx->uid.p; A.ranv = ctx->ranv.p; A.old_cg_w = ctx->old_cg.p; A.ph = ctx->ph;
  A.dq = ctx->use_dq ? ctx->dq.p : nullptr;
}
// fuse: see TUArgs (bits 0-1 overlap_moveback's first / last pass, bit 2 the tail of the loop body); defer: leave the cell sort of a
// rebuild to whoever needs the cells first
static int enq_test_update(dml_ctx *ctx, int fuse = 0, bool cells_wanted = true, bool defer = false) {
  ctx->tu_fused = 0;
  tessellate(ctx);
  int n = ctx->n, nct = ctx->nct;
  if (!ctx->tessellated) {
    // fewer than 4 cells on every axis: no cell lists (Cells.F90:231), rows by ngroup_verlet (Neighbor.F90:358-424)
    if (ctx->cfg.reservoir == 3) FAIL("gcmc_run needs the cell lists: box smaller than 4 cells in every direction");
    const int nbv = std::min(nblk(n), 148 * 6);
    if (fuse & 2) LAUNCH(K_OV_APPLY, k_ov_apply, nblk(n), TPB, ctx->posm.p, ctx->vel.p, ctx->acel.p, ctx->old_cg.p, ctx->ovst.p, ctx->sc, n);   // passes dml_step folds into the cooperative kernel
    CKC(ctx->part.ensure((size_t)2 * nbv, ctx->st)); CKC(ctx->b2slot.ensure(ctx->cap, ctx->st));
    LAUNCH(K_PBC_BIN, k_pbc_disp, nbv, TPB, ctx->posm.p, ctx->pos_old.p, ctx->part.p, ctx->lay.p, ctx->sc, ctx->geo, n, n, 1, ctx->cfg.nb_dcut, ctx->ph.r0_max, ctx->cfg.rcut);
    CKC(cudaMemsetAsync(ctx->b2slot.p, 0xff, (size_t)ctx->cap * sizeof(int), ctx->st));
    LAUNCH(K_BIN, k_verlet_prepare, std::min(nblk(n), 148 * 8), TPB, ctx->posm.p, ctx->pos_old.p, ctx->slot_b.p, ctx->b2slot.p, ctx->rh.p, ctx->halo_of.p, ctx->sc, n);
    if (fuse & 1) LAUNCH(K_OV_INIT, k_ov_init, nblk(n), TPB, ctx->posm.p, ctx->parent.p, ctx->ovst.p, ctx->comp_cnt.p, ctx->ov_head.p, ctx->sc, n);
    ctx->binned = true;
    return 0;
  }
  if ((size_t)nct + 2 > ctx->cell_start.cap) {
    CKC(ctx->cell_cnt.ensure(nct + 1, ctx->st)); CKC(ctx->cell_start.ensure(nct + 2, ctx->st)); CKC(ctx->cell_cur.ensure(nct + 1, ctx->st));
    CKC(cudaMemsetAsync(ctx->cell_cnt.p, 0, ctx->cell_cnt.cap * sizeof(int), ctx->st));
    CKC(cudaMemsetAsync(ctx->cell_cur.p, 0, ctx->cell_cur.cap * sizeof(int), ctx->st));
  }
  // gcmc_run needs the cells of the current positions every step; inside dml_step only the test_update right in front of it has to
  // provide them (the one in front of overlap_moveback sorts only when it rebuilds)
  int force = (ctx->cfg.reservoir == 3 && cells_wanted) ? 1 : 0;
  // a forced sort without a rebuild replaces the snapshot that pending rows would be built from: build them first (no-op otherwise)
  if (force) TRY(enq_materialize_rows(ctx));
  if (ctx->use_coop && n <= ctx->coop_tu_max_n) {
    TUArgs A;
    fill_tu_args(ctx, A, force);
    A.fuse = ctx->no_tu_fuse ? 0 : fuse;
    A.defer = (defer && !force) ? 1 : 0;
    if (A.defer) { CKC(ctx->snap.ensure(ctx->cap, ctx->st)); A.snap = ctx->snap.p; ctx->sort_maybe_pending = true; }
    else ctx->sort_maybe_pending = false;                 // a sorting call catches up with (or supersedes) a deferred sort
    ctx->tu_fused = A.fuse;
    if (A.fuse & 8) LAUNCH_COOP(K_TU_COOP, k_test_update_coop<true>, ctx->coop_grid_tu_bi, A);
    else LAUNCH_COOP(K_TU_COOP, k_test_update_coop<false>, ctx->coop_grid_tu, A);
    ctx->binned = true;
    return 0;
  }
  int nb = std::min(nblk(n), 148 * 6);
  CKC(ctx->part.ensure((size_t)2 * nb, ctx->st));
  LAUNCH(K_PBC_BIN, k_pbc_disp, nb, TPB, ctx->posm.p, ctx->pos_old.p, ctx->part.p, ctx->lay.p, ctx->sc, ctx->geo, n, n, 1, ctx->cfg.nb_dcut, ctx->ph.r0_max, ctx->cfg.rcut,
         (double *)nullptr, ctx->use_dq ? ctx->dq.p : nullptr);
  TRY(enq_sort_cells(ctx, force));
  ctx->binned = true;
  return 0;
}

static int enq_integrate(dml_ctx *ctx, bool ermak) {
  ctx->kb_valid = false;
  int n = ctx->n;
  ctx->step++;
  if (ctx->ph.rng_mode == DML_RNG_REPLAY && !ctx->have_rp) FAIL("replay mode: call dml_set_replay_integrator before the integrator");
  LAUNCH(K_MISC, k_tick, 1, 1, ctx->sc);                  // the Philox step word lives on the device (DevScal::istep)
  if (ctx->ph.rng_mode == DML_RNG_REFERENCE) {
    // the reference's sequential stream: one thread walks hs%ref in creation-rank order (k_integrate_seq)
    TRY(pull_scal(ctx));
    const int nord = std::max(ctx->hsc->next_uid, 1);
    CKC(ctx->ord.ensure((size_t)nord + 64, ctx->st));
    CKC(cudaMemsetAsync(ctx->ord.p, 0xff, (size_t)nord * sizeof(int), ctx->st));
    LAUNCH(K_MISC, k_ord_scatter, nblk(n), TPB, ctx->posm.p, ctx->uid.p, ctx->ord.p, n, nord);
    if (ermak) LAUNCH(K_INTEGRATE, (k_integrate_seq<true>), 1, 32, ctx->posm.p, ctx->vel.p, ctx->acel.p, ctx->pos_old.p, ctx->old_cg.p, ctx->ranv.p, ctx->ord.p, nord, ctx->sc, ctx->geo, ctx->ph);
    else LAUNCH(K_INTEGRATE, (k_integrate_seq<false>), 1, 32, ctx->posm.p, ctx->vel.p, ctx->acel.p, ctx->pos_old.p, ctx->old_cg.p, ctx->ranv.p, ctx->ord.p, nord, ctx->sc, ctx->geo, ctx->ph);
    return 0;
  }
  if (ermak)
    LAUNCH(K_INTEGRATE, (k_integrate<true>), nblk(n), TPB, ctx->posm.p, ctx->vel.p, ctx->acel.p, ctx->pos_old.p, ctx->old_cg.p, ctx->ranv.p,
           ctx->uid.p, ctx->rp_gauss.p, ctx->rp_upbc.p, ctx->sc, ctx->geo, ctx->ph, STEP_FROM_DEVICE, n);
  else
    LAUNCH(K_INTEGRATE, (k_integrate<false>), nblk(n), TPB, ctx->posm.p, ctx->vel.p, ctx->acel.p, ctx->pos_old.p, ctx->old_cg.p, ctx->ranv.p,
           ctx->uid.p, ctx->rp_gauss.p, ctx->rp_upbc.p, ctx->sc, ctx->geo, ctx->ph, STEP_FROM_DEVICE, n);
  ctx->have_rp = false;
  return 0;
}

// transposed rows, built on the device only when rows can be asymmetric (guarded launches, no-ops otherwise)
static int enq_build_rev(dml_ctx *ctx) {
  int n = ctx->n;
  if (ctx->use_coop && ctx->coop_grid_rev > 0) {
    RevArgs A;
    A.rh = ctx->rh.p; A.cols = ctx->cols.p; A.posm = ctx->posm.p; A.rev_start = ctx->rev_start.p; A.rev_len = ctx->rev_len.p; A.rev_cnt = ctx->rev_cnt.p;
    A.rev_cols = ctx->rev_cols.p; A.bq = ctx->bq.p; A.rev_bq = ctx->rev_bq.p; A.halo_of = ctx->halo_of.p; A.halo_only = ctx->cfg.strict_order ? 0 : 1;
    A.sums = ctx->coop_sums.p; A.sc = ctx->sc; A.n = n;
    LAUNCH_COOP(K_REV, k_rev_coop, ctx->coop_grid_rev, A);
    return 0;
  }
  LAUNCH(K_REV, k_rev_count, std::min(nblk(n), 148 * 8), TPB, ctx->rh.p, ctx->cols.p, ctx->posm.p, ctx->rev_len.p, ctx->rev_cnt.p,
         ctx->halo_of.p, ctx->cfg.strict_order ? 0 : 1, ctx->sc, n);
  TRY(scan_excl(ctx, ctx->rev_cnt.p, ctx->rev_start.p, n, &ctx->sc->rev_used, true, 1, 0));
  LAUNCH(K_REV, k_rev_fill, std::min(nblk(n), 148 * 8), TPB, ctx->rh.p, ctx->cols.p, ctx->posm.p, ctx->rev_start.p, ctx->rev_len.p,
         ctx->rev_cols.p, ctx->bq.p, ctx->rev_bq.p, ctx->halo_of.p, ctx->cfg.strict_order ? 0 : 1, ctx->sc, n);
  LAUNCH(K_REV, k_rev_done, 1, 1, ctx->sc);
  return 0;
}

static int enq_qtab(dml_ctx *ctx) {
  LAUNCH(K_MISC, k_qtab, 1, 256, ctx->lay.p, ctx->sc, ctx->geo, ctx->ph.r0_max, ctx->cfg.rcut);
  return 0;
}
// ermak_b (dana.F90:1031-1052): right behind a pair-force call the flat form (k_ermak_b_flat), else the form that reads the records
static int enq_ermak_b(dml_ctx *ctx) {
  const int n = ctx->n;
  if (ctx->kb_valid && !ctx->no_flat_b)
    LAUNCH(K_ERMAK_B, k_ermak_b_flat, nblk(3 * n), TPB, ctx->vel.p, ctx->acel.p, ctx->fe.p, ctx->ranv.p, ctx->ph, 3 * n, ctx->fnz.p, ctx->kb.p);
  else
    LAUNCH(K_ERMAK_B, k_ermak_b, nblk(n), TPB, ctx->posm.p, ctx->vel.p, ctx->acel.p, ctx->fe.p, ctx->ranv.p, ctx->ph, n, ctx->fnz.p);
  ctx->kb_valid = false;
  return 0;
}
// fused = called from the step sequence: the production kernel may also apply ermak_b (k_fuerza_sub<true, ..>)
static int enq_fuerza(dml_ctx *ctx, bool fused = false) {
  int n = ctx->n;
  TRY(enq_materialize_rows(ctx));
  TRY(enq_build_rev(ctx));                              // guarded on the device: no-ops unless rows are asymmetric and the transposed rows stale
  if (ctx->cfg.strict_order) {
    LAUNCH(K_FUERZA, (k_fuerza<true>), nblk(n), TPB, ctx->posm.p, ctx->rh.p, ctx->cols.p, ctx->rev_start.p,
           ctx->rev_len.p, ctx->rev_cols.p, ctx->sc, ctx->uid.p, ctx->fe.p, ctx->geo, ctx->ph, n, ctx->fnz.p, ctx->kb.p);
    ctx->kb_valid = true;
    return 0;
  }
#define FSUB(F, B) LAUNCH(K_FUERZA, (k_fuerza_sub<F, B>), nblk(n), TPB, ctx->posm.p, ctx->rh.p, ctx->cols.p, ctx->rev_start.p, \
                       ctx->rev_len.p, ctx->rev_cols.p, ctx->bq.p, ctx->rev_bq.p, ctx->halo_of.p, ctx->lay.p, ctx->sc, ctx->fe.p, ctx->geo, ctx->ph, n, \
                       ctx->vel.p, ctx->acel.p, ctx->ranv.p, ctx->fnz.p, ctx->use_dq ? (const unsigned char *)ctx->dq.p : (const unsigned char *)nullptr, ctx->kb.p)
  if (fused && ctx->fuse_ermak_b) { if (ctx->force_minb >= 6) FSUB(true, 6); else FSUB(true, 4); }
  else if (ctx->force_minb >= 8) FSUB(false, 8);
  else if (ctx->force_minb >= 6) FSUB(false, 6);
  else FSUB(false, 4);
#undef FSUB
  ctx->kb_valid = true;
  return 0;
}

// overlap_moveback (dana.F90:849-943)
// inside dml_step the two test_update launches around overlap_moveback can carry its first and last pass over the slots
static bool tu_can_fuse(const dml_ctx *ctx) { return ctx->use_coop && ctx->n <= ctx->coop_tu_max_n && !ctx->no_tu_fuse; }
static bool ov_is_multi_launch(const dml_ctx *ctx) { return !(ctx->use_coop && ctx->n <= ctx->coop_max_n && ctx->cfg.prob >= 1.0); }

static OvRp ov_replay(const dml_ctx *ctx) {
  OvRp r; r.vals = ctx->have_rp_ovl ? ctx->rp_uovl.p : nullptr; r.qstart = ctx->rp_qstart.p; r.draws = ctx->ov_draws.p;
  return r;
}
static int enq_overlap_impl(dml_ctx *ctx, bool fused, bool init_done, bool defer_apply);
static int enq_overlap(dml_ctx *ctx, bool fused = false, bool init_done = false, bool defer_apply = false) {
  if (ctx->ph.rng_mode != DML_RNG_REFERENCE) return enq_overlap_impl(ctx, fused, init_done, defer_apply);
  LAUNCH(K_MISC, k_rng_mark, 1, 1, ctx->sc);              // the reference draws one uniform per deposition attempt (dana.F90:898): consumed afterwards
  TRY(enq_overlap_impl(ctx, fused, init_done, defer_apply));
  LAUNCH(K_MISC, k_rng_advance, 1, 1, ctx->sc);
  return 0;
}
static int enq_overlap_impl(dml_ctx *ctx, bool fused, bool init_done, bool defer_apply) {
  ctx->kb_valid = false;
  int n = ctx->n;
  TRY(enq_materialize_rows(ctx));
  if (!fused) TRY(enq_qtab(ctx));
  if (ctx->use_coop && n <= ctx->coop_max_n && ctx->cfg.prob >= 1.0) {
    OVArgs A;
    A.posm = ctx->posm.p; A.vel = ctx->vel.p; A.acel = ctx->acel.p; A.old_cg = ctx->old_cg.p; A.rh = ctx->rh.p;
    A.cols = ctx->cols.p; A.bq = ctx->bq.p; A.lay = ctx->lay.p; A.parent = ctx->parent.p; A.ovst = ctx->ovst.p; A.comp_cnt = ctx->comp_cnt.p;
    A.comp_off = ctx->comp_off.p; A.members = ctx->members.p; A.roots = ctx->roots.p; A.ov_head = ctx->ov_head.p; A.ov_next = ctx->ov_next.p; A.uid = ctx->uid.p;
    A.rp_uovl = ov_replay(ctx); A.sc = ctx->sc; A.g = ctx->geo; A.ph = ctx->ph; A.step = STEP_FROM_DEVICE;
    A.n = n; A.guard_pass = ctx->ov_guard_pass;
    LAUNCH_COOP(K_OV_COOP, k_overlap_coop, ctx->coop_grid_ov, A);
    ctx->have_rp_ovl = false;
    return 0;
  }
  if (!init_done) LAUNCH(K_OV_INIT, k_ov_init, nblk(n), TPB, ctx->posm.p, ctx->parent.p, ctx->ovst.p, ctx->comp_cnt.p, ctx->ov_head.p, ctx->sc, n);
#define OVDET(L) LAUNCH(K_OV_DETECT, k_ov_detect<L>, nblk((long long)n * L), TPB, ctx->posm.p, ctx->old_cg.p, ctx->rh.p, ctx->cols.p, ctx->bq.p, \
                        ctx->lay.p, ctx->parent.p, ctx->ovst.p, ctx->sc, ctx->geo, n, ctx->use_dq ? (const unsigned char *)ctx->dq.p : (const unsigned char *)nullptr)
  if (ctx->ov_lanes >= 4) OVDET(4); else if (ctx->ov_lanes >= 2) OVDET(2); else OVDET(1);
#undef OVDET
  const OvRp uovl = ov_replay(ctx);
  if (ctx->cfg.prob >= 1.0) {
    LAUNCH(K_OV_LINK, k_ov_link, nblk(n), TPB, ctx->parent.p, ctx->ovst.p, ctx->ov_head.p, ctx->ov_next.p, ctx->roots.p, ctx->sc, n);
    if (ctx->ov_unstaged) {
      LAUNCH(K_OV_PASS, k_ov_resolve<false>, std::min(nblk(n, 4), 148 * ctx->ov_res_bpsm), 128, ctx->posm.p, ctx->old_cg.p, ctx->rh.p, ctx->cols.p, ctx->bq.p,
           ctx->lay.p, ctx->ovst.p, ctx->roots.p, ctx->ov_head.p, ctx->ov_next.p, ctx->members.p, ctx->uid.p, uovl, ctx->sc, ctx->geo, ctx->ph,
           STEP_FROM_DEVICE, ctx->ov_guard_pass);
    } else {
      LAUNCH(K_OV_PASS, k_ov_resolve<true>, std::min(nblk(n, 4), 148 * ctx->ov_res_bpsm), 128, ctx->posm.p, ctx->old_cg.p, ctx->rh.p, ctx->cols.p, ctx->bq.p,
           ctx->lay.p, ctx->ovst.p, ctx->roots.p, ctx->ov_head.p, ctx->ov_next.p, ctx->members.p, ctx->uid.p, uovl, ctx->sc, ctx->geo, ctx->ph,
           STEP_FROM_DEVICE, ctx->ov_guard_pass);
    }
  } else {
    LAUNCH(K_OV_COUNT, k_ov_count, nblk(n), TPB, ctx->parent.p, ctx->ovst.p, ctx->comp_cnt.p, n);
    LAUNCH(K_OV_ALLOC, k_ov_alloc, nblk(n), TPB, ctx->parent.p, ctx->ovst.p, ctx->comp_cnt.p, ctx->comp_off.p, ctx->roots.p, ctx->sc, n);
    LAUNCH(K_OV_FILL, k_ov_fill, nblk(n), TPB, ctx->parent.p, ctx->ovst.p, ctx->comp_cnt.p, ctx->comp_off.p, ctx->members.p, n);
    // prob<1: a failed deposition leaves skip=.false. without asking for another pass, so whether that atom is looked at
    // again depends on the other components: keep the reference's global recursion levels (one launch + one flag read each)
    TRY(pull_scal(ctx));
    int nroots = ctx->hsc->n_roots;
    if (nroots > 0) {
      LAUNCH(K_OV_SORT, k_ov_sort, nblk(nroots, 128), 128, ctx->roots.p, ctx->comp_cnt.p, ctx->comp_off.p, ctx->members.p, ctx->uid.p, ctx->sc);
      for (int pass = 0;; ++pass) {
        CKC(cudaMemsetAsync(&ctx->sc->again, 0, sizeof(int), ctx->st));
        LAUNCH(K_OV_PASS, k_ov_pass, nblk(nroots, 64), 64, ctx->posm.p, ctx->old_cg.p, ctx->rh.p, ctx->cols.p,
               ctx->bq.p, ctx->lay.p, ctx->ovst.p, ctx->roots.p, ctx->comp_cnt.p, ctx->comp_off.p, ctx->members.p, ctx->uid.p, uovl, ctx->sc, ctx->geo, ctx->ph,
               STEP_FROM_DEVICE, pass, (ctx->ov_guard_pass > 0 && pass >= ctx->ov_guard_pass) ? 1 : 0);
        TRY(pull_scal(ctx));
        if (!ctx->hsc->again) { CKC(cudaMemsetAsync(&ctx->sc->any_active, 0, sizeof(int), ctx->st)); ctx->hsc->any_active = pass + 1;
                                CKC(cudaMemcpyAsync(&ctx->sc->any_active, &ctx->hsc->any_active, sizeof(int), cudaMemcpyHostToDevice, ctx->st)); break; }
      }
    }
  }
  if (!defer_apply) LAUNCH(K_OV_APPLY, k_ov_apply, nblk(n), TPB, ctx->posm.p, ctx->vel.p, ctx->acel.p, ctx->old_cg.p, ctx->ovst.p, ctx->sc, n);
  ctx->have_rp_ovl = false;
  return 0;
}

static int enq_promote(dml_ctx *ctx) {
  ctx->kb_valid = false;
  if (ctx->cfg.reservoir == 3) LAUNCH(K_PROMOTE, k_gcmc_tomb, nblk(ctx->n), TPB, ctx->posm.p, ctx->gorder.p, ctx->gpos.p, ctx->sc, ctx->n);
  LAUNCH(K_PROMOTE, k_promote, nblk(ctx->n), TPB, ctx->posm.p, ctx->sc, ctx->n);
  return 0;
}
static int enq_calc_rho(dml_ctx *ctx) {
  LAUNCH(K_CALC_RHO, k_calc_rho, nblk(ctx->n), TPB, ctx->posm.p, ctx->sc, ctx->geo.box[0] * ctx->geo.box[1], ctx->cfg.reservoir == 2 ? 1 : 0, ctx->n);
  return 0;
}
static int enq_maxz(dml_ctx *ctx) { LAUNCH(K_MAXZ, k_maxz, nblk(ctx->n), TPB, ctx->posm.p, ctx->sc, ctx->cfg.h / ctx->cfg.tau, ctx->n); return 0; }

static int upload_d(dml_ctx *ctx, double *dst, const double *src, size_t cnt) {
  if (!src) return 0;
  CKC(cudaMemcpyAsync(dst, src, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  return 0;
}

// End of a public call: one device->host read of the scalar block; refresh the host mirrors, surface device-side
// errors, and grow the neighbour storage while there is still head-room.
static int finish(dml_ctx *ctx) {
  TRY(pull_scal(ctx));
  if (ctx->sg_ran && !ctx->sg_evs.empty()) {              // kernels timed inside the captured step (the stream is idle here)
    for (auto &ev : ctx->sg_evs) { float ms = 0; if (cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess) { ctx->prof_ms[ev.cls] += ms; ctx->prof_n[ev.cls]++; } else cudaGetLastError(); }
  }
  ctx->sg_ran = false;
  ctx->n = ctx->hsc->n_slots;
  if (ctx->n > ctx->l2_slots) l2_window(ctx, ctx->n + ctx->n / 4);
  const size_t tail0 = (size_t)ctx->hsc->cols_tail0;
  size_t used = (size_t)std::max(ctx->hsc->cols_used, ctx->hsc->rev_used);
  if (used > tail0 && (used - tail0) * 2 > ctx->cols.cap - tail0) {          // half of the tail region is in use
    size_t want = tail0 + (used - tail0) * 3 + 4096;
    CKC(ctx->cols.ensure(want, ctx->st, true)); CKC(ctx->rev_cols.ensure(ctx->cols.cap, ctx->st, false));
    CKC(ctx->bq.ensure(ctx->cols.cap, ctx->st, true)); CKC(ctx->rev_bq.ensure(ctx->cols.cap, ctx->st, false));
    ctx->hsc->cols_cap = (int)std::min<size_t>(ctx->cols.cap, 0x7fffffff);
    ctx->hsc->rev_valid = 0;
    ctx->sg_n = -1;                                         // the captured step holds the old pointers
    slab_graphs_drop(ctx);
    TRY(push_scal(ctx));
    CKC(cudaStreamSynchronize(ctx->st));
  }
  return 0;
}

// bloques — dana.F90:716-773 (host side: needs rho, appends the template block, changes the box)
static int do_bloques(dml_ctx *ctx, int nchunk, const double *cpos, const double *cpos_old, double dist, double rhomedia, int *fired) {
  ctx->kb_valid = false;
  TRY(pull_scal(ctx));
  double drho = ctx->hsc->rho - rhomedia;
  if (fired) *fired = 0;
  if (std::fabs(drho) < (rhomedia * (double)0.186f)) return 0;
  if (fired) *fired = 1;
  int n0 = ctx->n;
  TRY(ensure_particles(ctx, n0 + nchunk));
  ctx->hsc->z0 += dist; ctx->hsc->z1 += dist; ctx->hsc->zmax += dist;
  std::vector<double> zero3((size_t)nchunk * 3, 0.0), og((size_t)nchunk * 3, 1e8);
  std::vector<int> z(nchunk, 1), fl(nchunk, DML_F_REF), uid(nchunk), sb(nchunk), one(nchunk, 1);
  for (int i = 0; i < nchunk; ++i) { uid[i] = ctx->hsc->next_uid + i; sb[i] = n0 + i; }
  ctx->hsc->next_uid += nchunk; ctx->hsc->n_slots = n0 + nchunk; ctx->hsc->b_amax = n0 + nchunk;
  ctx->hsc->nat_sys += nchunk; ctx->hsc->nat_ref += nchunk;
  ctx->hsc->listed = 0;                                   // hs%listed=.false. (dana.F90:737)
  TRY(push_scal(ctx));
  CKC(ctx->stage_d.ensure((size_t)nchunk * 3, ctx->st)); CKC(ctx->stage_i.ensure((size_t)nchunk * 2, ctx->st));
  CKC(cudaMemcpyAsync(ctx->stage_d.p, cpos, (size_t)nchunk * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  CKC(cudaMemcpyAsync(ctx->stage_i.p, z.data(), nchunk * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  CKC(cudaMemcpyAsync(ctx->stage_i.p + nchunk, fl.data(), nchunk * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  LAUNCH(K_PACK, k_pack, nblk(nchunk), TPB, ctx->posm.p + n0, ctx->stage_d.p, ctx->stage_i.p, ctx->stage_i.p + nchunk, nchunk);
  TRY(upload_d(ctx, ctx->pos_old.p + (size_t)3 * n0, cpos_old, (size_t)nchunk * 3));
  TRY(upload_d(ctx, ctx->vel.p + (size_t)3 * n0, zero3.data(), (size_t)nchunk * 3));
  TRY(upload_d(ctx, ctx->acel.p + (size_t)3 * n0, zero3.data(), (size_t)nchunk * 3));
  CKC(cudaMemsetAsync(ctx->fe.p + n0, 0, (size_t)nchunk * sizeof(double4), ctx->st));
  TRY(upload_d(ctx, ctx->old_cg.p + (size_t)3 * n0, og.data(), (size_t)nchunk * 3));
  CKC(cudaMemcpyAsync(ctx->uid.p + n0, uid.data(), nchunk * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  CKC(cudaMemcpyAsync(ctx->slot_b.p + n0, sb.data(), nchunk * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  CKC(cudaMemcpyAsync(ctx->b_occ.p + n0, one.data(), nchunk * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));                     // host vectors go out of scope
  ctx->n = n0 + nchunk;
  double box[3] = {ctx->geo.box[0], ctx->geo.box[1], ctx->hsc->zmax};
  set_box(ctx, box);
  return enq_test_update(ctx);
}

// one iteration of dana's loop body (dana.F90:173-265), enqueue only.  Part a: everything up to calc_rho (no host round trip);
// part b: the reservoirs that need the host (reservoir 2 reads rho once per step) and the end of the step.
static int enq_step_a(dml_ctx *ctx) {
  int n = ctx->n;
  if (ctx->cfg.integrador) {
    TRY(enq_integrate(ctx, true)); TRY(enq_fuerza(ctx, true));
    if (ctx->cfg.strict_order || !ctx->fuse_ermak_b)
      TRY(enq_ermak_b(ctx));
  }
  // Brownian step with Philox noise: the integrator rides on the first pass of the test_update that follows it (one launch and one
  // pass over the records less); DML_NO_BI_FUSE=1 keeps the launch of its own
  tessellate(ctx);
  const bool bi = !ctx->cfg.integrador && tu_can_fuse(ctx) && ctx->tessellated && ctx->ph.rng_mode == DML_RNG_PHILOX && !ctx->no_bi_fuse && ctx->coop_grid_tu_bi > 0;
  if (!ctx->cfg.integrador) {
    if (bi) { ctx->step++; LAUNCH(K_MISC, k_tick, 1, 1, ctx->sc); ctx->have_rp = false; }
    else TRY(enq_integrate(ctx, false));
  }
  const bool fz = tu_can_fuse(ctx) && ov_is_multi_launch(ctx);   // k_ov_init rides on the first test_update, k_ov_apply on the second
  TRY(enq_test_update(ctx, (fz ? 1 : 0) | (bi ? 8 : 0), false));
  TRY(enq_overlap(ctx, true, fz, fz));
  // the second test_update also carries the tail of the loop body (msd bookkeeping, promotion, calc_rho, maxz) when it runs as
  // one cooperative launch, and in Brownian mode leaves the cell sort of its rebuild to whoever needs it (nobody, usually: Q11)
  const bool tail = ctx->cfg.reservoir != 3 && ctx->tessellated && ctx->use_coop && n <= ctx->coop_tu_max_n && !ctx->no_tu_fuse;
  TRY(enq_test_update(ctx, (fz ? 2 : 0) | (tail ? 4 : 0), true, tail && !ctx->cfg.integrador));
  const bool tail_done = (ctx->tu_fused & 4) != 0;
  if (ctx->cfg.reservoir == 3) {
    LAUNCH(K_MISC, k_msd_book, 1, 1, ctx->sc);
    TRY(enq_promote(ctx));
    TRY(gcmc_run_impl(ctx));
    TRY(enq_calc_rho(ctx));
  } else if (!tail_done) {
    LAUNCH(K_PROMOTE, k_promote_rho, std::min(nblk(n), 148 * 8), TPB, ctx->posm.p, ctx->sc, ctx->geo.box[0] * ctx->geo.box[1], ctx->cfg.reservoir == 2 ? 1 : 0, n);
  }
  ctx->step_tail_done = tail_done;
  return 0;
}
static int enq_step_b(dml_ctx *ctx) {
  const bool tail_done = ctx->step_tail_done;
  if (ctx->cfg.reservoir == 2) {
    if (!ctx->have_chunk) FAIL("reservoir 2: call dml_set_chunk_template before dml_step");
    int fired = 0;
    int nch = (int)(ctx->ch_pos.size() / 3);
    TRY(do_bloques(ctx, nch, ctx->ch_pos.data(), ctx->ch_pos_old.data(), ctx->ch_dist, ctx->ch_rhomedia, &fired));
    if (fired) for (int i = 0; i < nch; ++i) { ctx->ch_pos[3 * i + 2] += ctx->ch_dist; ctx->ch_pos_old[3 * i + 2] += ctx->ch_dist; }
  }
  if (ctx->cfg.reservoir == 1 && !tail_done) TRY(enq_maxz(ctx));
  ctx->t = ctx->t + ctx->cfg.h;
  return 0;
}
static bool graph_ok(const dml_ctx *ctx) {
  return ctx->use_graph && (!ctx->profiling || ctx->prof_only >= 0) && ctx->ph.rng_mode == DML_RNG_PHILOX && ctx->cfg.prob >= 1.0 && ctx->cfg.reservoir != 3 &&
         ctx->tessellated && ctx->use_coop && ctx->n <= ctx->coop_tu_max_n;
}
// part a of one step: replay of the captured graph when there is a valid one, capture + launch otherwise
static int launch_step_a(dml_ctx *ctx) {
  if (!graph_ok(ctx)) return enq_step_a(ctx);
  if ((size_t)ctx->nct + 2 > ctx->cell_start.cap) return enq_step_a(ctx);   // (re)allocation of the cell tables happens outside a capture
  if (!ctx->cfg.integrador) CKC(ctx->snap.ensure(ctx->cap, ctx->st));
  if (!ctx->step_graph || ctx->sg_n != ctx->n || ctx->sg_nct != ctx->nct || memcmp(&ctx->sg_geo, &ctx->geo, sizeof(Geo)) != 0) {
    if (ctx->step_graph) { cudaGraphExecDestroy(ctx->step_graph); ctx->step_graph = nullptr; }
    const int64_t step0 = ctx->step, l0 = ctx->launches;
    prof_collect(ctx);                                      // events of plain launches so far
    for (auto &ev : ctx->sg_evs) ctx->pool.push_back(ev);
    ctx->sg_evs.clear();
    cudaGraph_t g = nullptr;
    CKC(cudaStreamBeginCapture(ctx->st, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    const int rc = enq_step_a(ctx);
    ctx->capturing = false;
    cudaError_t e = cudaStreamEndCapture(ctx->st, &g);
    ctx->sg_launches = ctx->launches - l0;
    ctx->step = step0; ctx->launches = l0;
    ctx->sg_evs.swap(ctx->evs);                             // event-record nodes of the graph: read after every replay (finish)
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess || !g) { cudaGetLastError(); ctx->use_graph = false; return enq_step_a(ctx); }   // capture not possible here: plain launches from now on
    e = cudaGraphInstantiate(&ctx->step_graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { cudaGetLastError(); ctx->step_graph = nullptr; ctx->use_graph = false; return enq_step_a(ctx); }
    ctx->sg_n = ctx->n; ctx->sg_nct = ctx->nct; ctx->sg_geo = ctx->geo;
  }
  ctx->step++; ctx->launches += ctx->sg_launches;
  CKC(cudaGraphLaunch(ctx->step_graph, ctx->st));
  ctx->sg_ran = true;
  return 0;
}
static int enq_step(dml_ctx *ctx) { TRY(launch_step_a(ctx)); return enq_step_b(ctx); }

#include "dml_gcmc.cuh"

// ---------------------------------------------------------------------------------------------------
extern "C" {

const char *dml_version(void) { return "dml-b200 0.1 (sm_100a)"; }
const char *dml_last_error(dml_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int dml_create(dml_ctx **out, const dml_config *cfg) {
  if (!out || !cfg) return -1;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return -2;      // no CPU fallback
  if (cudaSetDevice(cfg->device) != cudaSuccess) return -3;
  dml_ctx *ctx = new dml_ctx;
  ctx->cfg = *cfg;
  *out = ctx;
  CKC(cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking));
  int cap = ctx->cap = std::max(cfg->capacity, 256);
  memset(&ctx->geo, 0, sizeof ctx->geo);
  set_box(ctx, cfg->box);
  for (int k = 0; k < 3; ++k) { ctx->geo.pbc[k] = cfg->pbc[k]; ctx->geo.nc[k] = 1; }
  double rl = cfg->rcut + cfg->nb_dcut;
  ctx->geo.rc_list2 = rl * rl; ctx->geo.rcut2 = cfg->rcut * cfg->rcut; ctx->geo.bq_scale = 255.0 / rl;
  Phys &ph = ctx->ph; memset(&ph, 0, sizeof ph);
  for (int i = 0; i < 9; ++i) {
    ph.eps[i] = cfg->eps[i]; ph.r0[i] = cfg->r0[i]; ph.r0sq[i] = cfg->r0[i] * cfg->r0[i];
    double x = cfg->r0[i], x2 = x * x, x4 = x2 * x2; ph.r0p6[i] = x2 * x4;
    ph.r0sq_max = std::max(ph.r0sq_max, ph.r0sq[i]);
  }
  ph.r0_max = std::sqrt(ph.r0sq_max);
  for (int i = 0; i < 3; ++i) { ph.mass[i] = cfg->mass[i]; ph.sqrt_mass[i] = std::sqrt(cfg->mass[i]); }
  ph.h = cfg->h; ph.prob = cfg->prob; ph.tau = cfg->tau;
  {                                                       // set_ermak — dana.F90:947-971
    double h = cfg->h, gama = cfg->gama;
    ph.cc0 = std::exp(-h * gama);
    ph.cc1 = (1.0 - ph.cc0) / gama;
    ph.cc2 = (1.0 - ph.cc1 / h) / gama;
    ph.sdr = std::sqrt(h / gama * (2.0 - (3.0 - 4.0 * ph.cc0 + ph.cc0 * ph.cc0) / (h * gama)));
    ph.sdv = std::sqrt(1.0 - ph.cc0 * ph.cc0);
    ph.crv1 = (1.0 - ph.cc0) * (1.0 - ph.cc0) / (gama * ph.sdr * ph.sdv);
    ph.crv2 = std::sqrt(1.0 - (ph.crv1 * ph.crv1));
    ph.skt = std::sqrt(cfg->kB_ui * cfg->Tsist);
    ph.cc1mcc2 = ph.cc1 - ph.cc2; ph.cc2h = ph.cc2 * h;
  }
  ph.dif_sc = cfg->dif_sc; ph.dif_sei = cfg->dif_sei; ph.z_sei = cfg->z_sei;
  ph.fac_sc = std::sqrt(2.0 * cfg->dif_sc * cfg->h); ph.fac_sei = std::sqrt(2.0 * cfg->dif_sei * cfg->h);
  ph.integrador = cfg->integrador; ph.piston = cfg->reservoir == 1; ph.chunks = cfg->reservoir == 2;
  ph.rng_mode = cfg->rng_mode; ph.seed = cfg->seed;
  if (cfg->rng_mode == DML_RNG_REFERENCE && cfg->prob < 1.0)
    FAIL("DML_RNG_REFERENCE needs prob = 1: the order of the deposition draws of overlap_moveback is only reproduced in number");
  ctx->row_slack = cfg->reservoir == 3 ? 8 : 0;
  if (const char *e = getenv("DML_COOP_MAX_N")) ctx->coop_max_n = ctx->coop_tu_max_n = atoi(e);
  if (const char *e = getenv("DML_COOP_TU_MAX_N")) ctx->coop_tu_max_n = atoi(e);
  if (getenv("DML_OV_UNSTAGED")) ctx->ov_unstaged = true;
  if (const char *e = getenv("DML_OV_RES_BPSM")) { int v = atoi(e); if (v >= 1 && v <= 16) ctx->ov_res_bpsm = v; }
  if (const char *e = getenv("DML_OV_LANES")) ctx->ov_lanes = atoi(e);
  if (ctx->ov_lanes <= 0) ctx->ov_lanes = ctx->cfg.integrador ? 1 : 4;   // measured: Brownian 100 k 36 -> 33 us with 4; Ermak 1 M (near-list path) 53 -> 91 us
  if (const char *e = getenv("DML_FORCE_MINB")) { int v = atoi(e); if (v >= 2 && v <= 8) ctx->force_minb = v; }
  if (getenv("DML_FUSE_ERMAK_B")) ctx->fuse_ermak_b = true;
  if (getenv("DML_ROWS_LEGACY")) ctx->rows_legacy = true;
  if (getenv("DML_NO_L2_PERSIST")) ctx->no_l2_persist = true;
  if (getenv("DML_NO_TU_FUSE")) ctx->no_tu_fuse = true;
  if (getenv("DML_NO_BI_FUSE")) ctx->no_bi_fuse = true;
  if (getenv("DML_NO_FLAT_B")) ctx->no_flat_b = true;
  if (getenv("DML_NO_SLAB_GRAPH") || getenv("DML_NO_GRAPH")) ctx->slab_graph_on = false;
  if (getenv("DML_NO_GRAPH")) ctx->use_graph = false;
  size_t c3 = (size_t)cap * 3;
  CKC(ctx->posm.ensure(cap, ctx->st)); CKC(ctx->sorted_posm.ensure(cap, ctx->st)); CKC(ctx->sorted_posf.ensure(cap, ctx->st));
  CKC(ctx->vel.ensure(c3, ctx->st)); CKC(ctx->acel.ensure(c3, ctx->st)); CKC(ctx->fe.ensure(cap, ctx->st));
  CKC(ctx->pos_old.ensure(c3, ctx->st)); CKC(ctx->old_cg.ensure(c3, ctx->st));
  CKC(ctx->ranv.ensure(c3, ctx->st)); CKC(ctx->uid.ensure(cap, ctx->st)); CKC(ctx->slot_b.ensure(cap, ctx->st));
  CKC(ctx->cell_of.ensure(cap, ctx->st)); CKC(ctx->sorted_slot.ensure(cap, ctx->st)); CKC(ctx->sorted_raw.ensure(cap, ctx->st)); CKC(ctx->chain_pos.ensure(cap, ctx->st));
  CKC(ctx->rh.ensure(cap, ctx->st)); CKC(cudaMemsetAsync(ctx->rh.p, 0, (size_t)cap * sizeof(RowHead), ctx->st));
  CKC(ctx->cols.ensure((size_t)cap * (ROW_W + 32) + 4096, ctx->st));   // region A (ROW_W per slot) + tail
  CKC(ctx->rev_cols.ensure(ctx->cols.cap, ctx->st));
  CKC(ctx->bq.ensure(ctx->cols.cap, ctx->st)); CKC(ctx->rev_bq.ensure(ctx->cols.cap, ctx->st));
  CKC(cudaMemsetAsync(ctx->bq.p, 0, ctx->bq.cap, ctx->st));
  CKC(ctx->sorted_cell.ensure(cap, ctx->st));
  CKC(ctx->halo_of.ensure(cap, ctx->st)); CKC(cudaMemsetAsync(ctx->halo_of.p, 0, cap, ctx->st));
  CKC(ctx->fnz.ensure(cap, ctx->st)); CKC(cudaMemsetAsync(ctx->fnz.p, 0, cap, ctx->st));
  CKC(ctx->dq.ensure(cap, ctx->st)); CKC(cudaMemsetAsync(ctx->dq.p, 0xff, cap, ctx->st));
  CKC(ctx->kb.ensure(cap, ctx->st)); CKC(cudaMemsetAsync(ctx->kb.p, 0, cap, ctx->st));
  ctx->use_dq = cfg->reservoir != 3 && !getenv("DML_NO_DQ");   // (gcmc reuses slots and appends rows between rebuilds: the layer bound alone)
  CKC(cudaFuncSetAttribute(k_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROWS_SMEM));
  CKC(ctx->lay.ensure(3 * LAY_MAX, ctx->st)); CKC(cudaMemsetAsync(ctx->lay.p, 0xff, 3 * LAY_MAX * sizeof(unsigned int), ctx->st));   // 2 displacement tables + the skip tables (k_qtab)
  CKC(cudaMemsetAsync(ctx->lay.p, 0, 2 * LAY_MAX * sizeof(unsigned int), ctx->st));
  CKC(ctx->rev_start.ensure(cap + 1, ctx->st)); CKC(ctx->rev_len.ensure(cap, ctx->st)); CKC(ctx->rev_cnt.ensure(cap, ctx->st));
  CKC(cudaMemsetAsync(ctx->rev_cnt.p, 0, (size_t)cap * sizeof(int), ctx->st));
  CKC(ctx->parent.ensure(cap, ctx->st)); CKC(ctx->ovst.ensure(cap, ctx->st)); CKC(ctx->comp_cnt.ensure(cap, ctx->st));
  CKC(ctx->comp_off.ensure(cap, ctx->st)); CKC(ctx->members.ensure(cap, ctx->st)); CKC(ctx->roots.ensure(cap, ctx->st));
  CKC(ctx->ov_head.ensure(cap, ctx->st)); CKC(ctx->ov_next.ensure(cap, ctx->st));
  ctx->gorder_cap = 2 * cap + 2048;
  CKC(ctx->gorder.ensure((size_t)2 * ctx->gorder_cap, ctx->st)); CKC(ctx->gpos.ensure(cap, ctx->st));
  CKC(ctx->gcc.ensure((size_t)ctx->gorder_cap / 1024 + 8, ctx->st)); CKC(ctx->b_occ.ensure(cap, ctx->st));
  CKC(ctx->rp_gauss.ensure((size_t)cap * 6, ctx->st)); CKC(ctx->rp_upbc.ensure(cap, ctx->st)); CKC(ctx->rp_uovl.ensure(cap, ctx->st));
  CKC(cudaMemsetAsync(ctx->posm.p, 0, (size_t)cap * sizeof(double4), ctx->st));
  CKC(cudaMemsetAsync(ctx->fe.p, 0, (size_t)cap * sizeof(double4), ctx->st));
  l2_window(ctx, cap);
  {
    int dev = 0, nsm = 0, coop = 0, b1 = 0, b2 = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    int b1i = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_test_update_coop<false>, TPB, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1i, k_test_update_coop<true>, TPB, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b2, k_overlap_coop, TPB, 0);
    ctx->coop_grid_tu = nsm * std::min(b1, 4); ctx->coop_grid_ov = nsm * b2;   // a larger grid only makes the grid-wide barriers slower
    ctx->coop_grid_tu_bi = nsm * std::min(b1i, 4);
    { int b3 = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b3, k_rev_coop, TPB, 0); ctx->coop_grid_rev = getenv("DML_NO_REV_COOP") ? 0 : nsm * std::min(b3, 4); }
    if (const char *e = getenv("DML_COOP_TU_BPSM")) { int v = atoi(e); if (v >= 1 && v <= b1) ctx->coop_grid_tu = nsm * v; if (v >= 1 && v <= b1i) ctx->coop_grid_tu_bi = nsm * v; }   // blocks per SM of k_test_update_coop
    if (const char *e = getenv("DML_COOP_OV_BPSM")) { int v = atoi(e); if (v >= 1 && v <= b2) ctx->coop_grid_ov = nsm * v; }
    ctx->use_coop = coop && b1 > 0 && b2 > 0 && !getenv("DML_NO_COOP");
    int gmax = std::max(std::max(std::max(std::max(ctx->coop_grid_tu, ctx->coop_grid_tu_bi), ctx->coop_grid_ov), ctx->coop_grid_rev), 1);
    CKC(ctx->coop_sums.ensure((size_t)gmax + 8, ctx->st));
    CKC(ctx->part.ensure((size_t)2 * std::max(gmax, nblk(cap)) + 8, ctx->st));
  }
  CKC(cudaMalloc(&ctx->sc, sizeof(DevScal)));
  CKC(cudaMallocHost(&ctx->hsc, sizeof(DevScal)));
  memset(ctx->hsc, 0, sizeof(DevScal));
  ctx->hsc->z0 = cfg->z0; ctx->hsc->z1 = cfg->z1; ctx->hsc->zmax = cfg->zmax;
  ctx->hsc->pist_P = 1.0;
  ctx->hsc->cols_cap = (int)std::min<size_t>(ctx->cols.cap, 0x7fffffff);
  ctx->hsc->cols_tail0 = ctx->hsc->cols_used = cap * ROW_W;
  TRY(push_scal(ctx));
  CKC(cudaStreamSynchronize(ctx->st));
  return 0;
}

void dml_destroy(dml_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  cudaStreamSynchronize(ctx->st);
  if (ctx->step_graph) cudaGraphExecDestroy(ctx->step_graph);
  slab_graphs_drop(ctx);
  prof_collect(ctx);
  for (auto &ev : ctx->pool) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  ctx->posm.release(); ctx->sorted_posm.release(); ctx->sorted_posf.release(); ctx->vel.release(); ctx->acel.release(); ctx->fe.release();
  ctx->pos_old.release(); ctx->old_cg.release(); ctx->ranv.release(); ctx->uid.release(); ctx->slot_b.release();
  ctx->snap.release(); ctx->cell_of.release(); ctx->cell_cnt.release(); ctx->cell_start.release(); ctx->cell_cur.release(); ctx->b2slot.release(); ctx->sorted_slot.release(); ctx->sorted_raw.release(); ctx->chain_pos.release();
  ctx->rh.release(); ctx->cols.release(); ctx->scan_sums.release(); ctx->part.release();
  ctx->parent.release(); ctx->ovst.release(); ctx->comp_cnt.release(); ctx->comp_off.release(); ctx->members.release(); ctx->roots.release(); ctx->ov_head.release(); ctx->ov_next.release();
  if (ctx->comm && nccl_api() && nccl_api()->CommDestroy) nccl_api()->CommDestroy(ctx->comm);
  ctx->send_lo.release(); ctx->send_hi.release(); ctx->slab_counts.release(); ctx->pack_uid_lo.release(); ctx->pack_uid_hi.release();
  ctx->pack_lo.release(); ctx->pack_hi.release();
  ctx->mig_list_lo.release(); ctx->mig_list_hi.release(); ctx->mig_rc.release(); ctx->mig_holes.release(); ctx->mig_si_lo.release(); ctx->mig_si_hi.release();
  ctx->mig_ri.release(); ctx->mig_sd_lo.release(); ctx->mig_sd_hi.release(); ctx->mig_rd.release();
  ctx->top2_own.release(); ctx->top2_all.release();
  ctx->bq.release(); ctx->rev_bq.release(); ctx->halo_of.release(); ctx->fnz.release(); ctx->dq.release(); ctx->kb.release(); ctx->sorted_cell.release(); ctx->lay.release();
  ctx->rev_start.release(); ctx->rev_len.release(); ctx->rev_cur.release(); ctx->rev_cols.release(); ctx->rev_cnt.release();
  ctx->coop_sums.release(); ctx->scan_state.release(); if (ctx->scan_tickets) cudaFree(ctx->scan_tickets);
  ctx->gorder.release(); ctx->gpos.release(); ctx->gcc.release(); ctx->gpend.release(); ctx->b_occ.release();
  ctx->obs_part.release(); ctx->obs_out.release(); ctx->obs_counts.release(); ctx->gr_cell_of.release(); ctx->gr_cnt.release(); ctx->gr_start.release();
  ctx->gr_sorted.release(); if (ctx->obs_ticket) cudaFree(ctx->obs_ticket);
  ctx->snap_uid.release(); ctx->snap_mb.release(); ctx->mc_out.release(); ctx->mc_count.release();
  ctx->rp_gauss.release(); ctx->rp_upbc.release(); ctx->rp_uovl.release(); ctx->rp_qstart.release(); ctx->ov_draws.release(); ctx->ord.release(); ctx->rp_gu.release(); ctx->rp_gg.release();
  ctx->stage_d.release(); ctx->stage_f.release(); ctx->stage_i.release();
  if (ctx->sc) cudaFree(ctx->sc);
  if (ctx->hsc) cudaFreeHost(ctx->hsc);
  if (ctx->st) cudaStreamDestroy(ctx->st);
  delete ctx;
}

// snapshot of who occupies every slot and of its group membership (dml_membership_changes reports against it)
static int member_snapshot(dml_ctx *ctx) {
  CKC(ctx->snap_uid.ensure(ctx->cap, ctx->st)); CKC(ctx->snap_mb.ensure(ctx->cap, ctx->st));
  LAUNCH(K_MISC, k_member_snap, std::min(nblk(ctx->cap, OBS_TPB), 148 * 8), OBS_TPB, ctx->posm.p, ctx->uid.p, ctx->n, ctx->cap, ctx->snap_uid.p, ctx->snap_mb.p);
  ctx->have_snap = true;
  return 0;
}

int dml_upload(dml_ctx *ctx, int32_t n, const double *pos, const double *vel, const double *acel, const double *pos_old,
               const double *old_cg, const int32_t *z, const int32_t *flags, const int32_t *uid, const int32_t *slot_b) { ENTER(ctx);
  TRY(ensure_particles(ctx, n));
  ctx->kb_valid = false;
  if (!pos || !z || !flags) FAIL("dml_upload: pos, z and flags are required");
  size_t n3 = (size_t)n * 3;
  CKC(ctx->stage_d.ensure(n3, ctx->st)); CKC(ctx->stage_i.ensure((size_t)n * 2, ctx->st));
  CKC(cudaMemcpyAsync(ctx->stage_d.p, pos, n3 * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  CKC(cudaMemcpyAsync(ctx->stage_i.p, z, n * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  CKC(cudaMemcpyAsync(ctx->stage_i.p + n, flags, n * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  CKC(cudaMemsetAsync(ctx->posm.p, 0, (size_t)ctx->cap * sizeof(double4), ctx->st));
  LAUNCH(K_PACK, k_pack, nblk(n), TPB, ctx->posm.p, ctx->stage_d.p, ctx->stage_i.p, ctx->stage_i.p + n, n);
  TRY(upload_d(ctx, ctx->vel.p, vel, n3)); TRY(upload_d(ctx, ctx->acel.p, acel, n3));
  TRY(upload_d(ctx, ctx->pos_old.p, pos_old ? pos_old : pos, n3));
  if (old_cg) TRY(upload_d(ctx, ctx->old_cg.p, old_cg, n3));
  if (!vel) CKC(cudaMemsetAsync(ctx->vel.p, 0, n3 * sizeof(double), ctx->st));
  if (!acel) CKC(cudaMemsetAsync(ctx->acel.p, 0, n3 * sizeof(double), ctx->st));
  ctx->n = n; ctx->binned = false;
  l2_window(ctx, n + n / 4);
  if (ctx->cfg.reservoir != 3) {
    // No gcmc group to keep in list order: creation ranks, b indices, their occupancy and the two running maxima are filled in
    // on the device (k_upload_book), so the call is copies + three small kernels and one synchronisation.
    if (uid) CKC(cudaMemcpyAsync(ctx->uid.p, uid, n * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
    if (slot_b) CKC(cudaMemcpyAsync(ctx->slot_b.p, slot_b, n * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
    CKC(cudaMemsetAsync(ctx->b_occ.p, 0, (size_t)ctx->cap * sizeof(int), ctx->st));
    TRY(pull_scal(ctx));
    ctx->hsc->glen = 0; ctx->hsc->ghead = 0; ctx->hsc->gtomb = 0; ctx->hsc->b_amax = 0;
    ctx->hsc->n_slots = n; ctx->hsc->next_uid = 0; ctx->hsc->listed = 0; ctx->hsc->rows_asym = 0; ctx->hsc->rev_valid = 0; ctx->hsc->need_rebuild = 0; ctx->hsc->rows_pending = 0; ctx->hsc->nat_sys = ctx->hsc->nat_ref = ctx->hsc->nat_gcmc = ctx->hsc->nlimbo = 0; ctx->hsc->hole_lo = ctx->hsc->bhole_lo = 0;
    TRY(push_scal(ctx));
    LAUNCH(K_MISC, k_upload_book, nblk(n), TPB, ctx->posm.p, ctx->uid.p, ctx->slot_b.p, ctx->b_occ.p, ctx->sc, n, ctx->cap, uid ? 0 : 1, slot_b ? 0 : 1);
    TRY(member_snapshot(ctx));
    TRY(pull_scal(ctx));                                  // surfaces an out-of-range slot_b (DML_E_CAPACITY) and refreshes the host mirror
    return 0;
  }
  std::vector<int> tmp;
  int mx = -1;
  if (uid) { CKC(cudaMemcpyAsync(ctx->uid.p, uid, n * sizeof(int), cudaMemcpyHostToDevice, ctx->st)); for (int i = 0; i < n; ++i) mx = std::max(mx, uid[i]); }
  else { tmp.resize(n); for (int i = 0; i < n; ++i) tmp[i] = i; mx = n - 1; CKC(cudaMemcpyAsync(ctx->uid.p, tmp.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->st)); CKC(cudaStreamSynchronize(ctx->st)); }
  if (slot_b) CKC(cudaMemcpyAsync(ctx->slot_b.p, slot_b, n * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  else { tmp.resize(n); for (int i = 0; i < n; ++i) tmp[i] = i; CKC(cudaMemcpyAsync(ctx->slot_b.p, tmp.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->st)); CKC(cudaStreamSynchronize(ctx->st)); }
  // gcmc membership in list order (= creation order) and occupancy of the b index (Groups.F90:1083-1093)
  std::vector<std::pair<int, int>> gm;
  std::vector<int> bocc(ctx->cap, 0), gpos(ctx->cap, 0), gord;
  int b_amax = 0;
  for (int i = 0; i < n; ++i) {
    bool alive = z[i] >= 1 && z[i] <= 3 && !(flags[i] & DML_F_LIMBO);
    if (!alive) continue;
    int sb = slot_b ? slot_b[i] : i;
    if (sb < 0 || sb >= ctx->cap) FAIL("dml_upload: slot_b out of range");
    bocc[sb] = 1; b_amax = std::max(b_amax, sb + 1);
    if (flags[i] & DML_F_GCMC) gm.push_back({uid ? uid[i] : i, i});
  }
  std::sort(gm.begin(), gm.end());
  for (size_t q = 0; q < gm.size(); ++q) { gord.push_back(gm[q].second); gpos[gm[q].second] = (int)q; }
  if ((int)gord.size() + 8 > ctx->gorder_cap) FAIL("gcmc membership exceeds capacity");
  if (!gord.empty()) CKC(cudaMemcpyAsync(ctx->gorder.p, gord.data(), gord.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  CKC(cudaMemcpyAsync(ctx->gpos.p, gpos.data(), (size_t)ctx->cap * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  CKC(cudaMemcpyAsync(ctx->b_occ.p, bocc.data(), (size_t)ctx->cap * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  TRY(pull_scal(ctx));
  ctx->hsc->glen = (int)gord.size(); ctx->hsc->ghead = 0; ctx->hsc->gtomb = 0; ctx->hsc->b_amax = b_amax;
  ctx->hsc->n_slots = n; ctx->hsc->next_uid = mx + 1; ctx->hsc->listed = 0; ctx->hsc->rows_asym = 0; ctx->hsc->rev_valid = 0; ctx->hsc->need_rebuild = 0; ctx->hsc->rows_pending = 0; ctx->hsc->nat_sys = ctx->hsc->nat_ref = ctx->hsc->nat_gcmc = ctx->hsc->nlimbo = 0; ctx->hsc->hole_lo = ctx->hsc->bhole_lo = 0;
  TRY(push_scal(ctx));
  LAUNCH(K_MISC, k_count_members, nblk(n), TPB, ctx->posm.p, ctx->sc, n);
  TRY(member_snapshot(ctx));
  CKC(cudaStreamSynchronize(ctx->st));
  return 0;
}

int dml_download(dml_ctx *ctx, int32_t n, double *pos, double *vel, double *acel, double *force, double *epot, double *pos_old,
                 double *old_cg, int32_t *z, int32_t *flags, int32_t *uid, int32_t *slot_b) { ENTER(ctx);
  if (n > ctx->n) FAIL("dml_download: n exceeds the number of slots");
  size_t n3 = (size_t)n * 3;
  CKC(ctx->stage_d.ensure(n3, ctx->st)); CKC(ctx->stage_i.ensure((size_t)n * 2, ctx->st));
  LAUNCH(K_PACK, k_unpack, nblk(n), TPB, ctx->posm.p, ctx->stage_d.p, ctx->stage_i.p, ctx->stage_i.p + n, n);
  if (pos) CKC(cudaMemcpyAsync(pos, ctx->stage_d.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
  if (z) CKC(cudaMemcpyAsync(z, ctx->stage_i.p, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
  if (flags) CKC(cudaMemcpyAsync(flags, ctx->stage_i.p + n, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
  if (vel) CKC(cudaMemcpyAsync(vel, ctx->vel.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
  if (acel) CKC(cudaMemcpyAsync(acel, ctx->acel.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
  if (force || epot) {
    CKC(ctx->stage_f.ensure(n3 + (size_t)n, ctx->st));
    LAUNCH(K_PACK, k_unpack_fe, nblk(n), TPB, ctx->fe.p, ctx->stage_f.p, ctx->stage_f.p + n3, n);
    if (force) CKC(cudaMemcpyAsync(force, ctx->stage_f.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    if (epot) CKC(cudaMemcpyAsync(epot, ctx->stage_f.p + n3, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
  }
  if (pos_old) CKC(cudaMemcpyAsync(pos_old, ctx->pos_old.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
  if (old_cg) CKC(cudaMemcpyAsync(old_cg, ctx->old_cg.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
  if (uid) CKC(cudaMemcpyAsync(uid, ctx->uid.p, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
  if (slot_b) CKC(cudaMemcpyAsync(slot_b, ctx->slot_b.p, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));
  return 0;
}

int dml_set_scalars(dml_ctx *ctx, const dml_scalars *s) { ENTER(ctx);
  TRY(pull_scal(ctx));
  ctx->hsc->z0 = s->z0; ctx->hsc->z1 = s->z1; ctx->hsc->zmax = s->zmax; ctx->hsc->rho = s->rho; ctx->hsc->rho0 = s->rho0;
  ctx->hsc->istep = (unsigned int)s->step;
  TRY(push_scal(ctx));
  set_box(ctx, s->box);
  ctx->t = s->t; ctx->step = s->step;
  CKC(cudaStreamSynchronize(ctx->st));
  return 0;
}
int dml_get_scalars(dml_ctx *ctx, dml_scalars *s) { ENTER(ctx);
  TRY(pull_scal(ctx));
  for (int k = 0; k < 3; ++k) s->box[k] = ctx->geo.box[k];
  s->z0 = ctx->hsc->z0; s->z1 = ctx->hsc->z1; s->zmax = ctx->hsc->zmax; s->rho = ctx->hsc->rho; s->rho0 = ctx->hsc->rho0;
  s->t = ctx->t; s->step = ctx->step;
  return 0;
}
int dml_set_rng_state(dml_ctx *ctx, const dmlh_rng *r) { ENTER(ctx);
  TRY(pull_scal(ctx));
  ctx->hsc->rr.idum = r->idum; ctx->hsc->rr.ix = r->ix; ctx->hsc->rr.iy = r->iy; ctx->hsc->rr.stored = r->stored; ctx->hsc->rr.g = r->g; ctx->hsc->rr.calls = r->calls;
  TRY(push_scal(ctx));
  CKC(cudaStreamSynchronize(ctx->st));
  return 0;
}
int dml_get_rng_state(dml_ctx *ctx, dmlh_rng *r) { ENTER(ctx);
  TRY(pull_scal(ctx));
  r->idum = ctx->hsc->rr.idum; r->ix = ctx->hsc->rr.ix; r->iy = ctx->hsc->rr.iy; r->stored = ctx->hsc->rr.stored; r->g = ctx->hsc->rr.g; r->calls = ctx->hsc->rr.calls;
  return 0;
}
int dml_get_counters(dml_ctx *ctx, dml_counters *c) { ENTER(ctx);
  TRY(pull_scal(ctx));
  memset(c, 0, sizeof *c);
  if (ctx->hsc->listed) {
    TRY(enq_materialize_rows(ctx));
    CKC(cudaMemsetAsync(&ctx->sc->list_entries, 0, sizeof(long long), ctx->st));
    LAUNCH(K_MISC, k_sum_rowlen, 64, TPB, ctx->rh.p, ctx->n, &ctx->sc->list_entries);
    TRY(pull_scal(ctx));
  }
  DevScal *h = ctx->hsc;
  c->nupd_vlist = h->nupd; c->try_ = h->try_; c->depo = h->depo; c->choques = h->choques; c->choques2 = h->choques2; c->choques3 = h->choques3;
  c->list_entries = h->listed ? h->list_entries : 0; c->overlap_passes = h->overlap_passes;
  c->gcmc_created = h->gcmc_created; c->gcmc_destroyed = h->gcmc_destroyed; c->row_overflow = h->row_overflow;
  c->max_vel = h->max_vel; c->msd_t = h->msd_t; c->msd_max = h->msd_max;
  c->n_slots = ctx->n; c->nat_sys = h->nat_sys; c->nat_ref = h->nat_ref; c->nat_gcmc = h->nat_gcmc;
  for (int k = 0; k < 3; ++k) { c->ncells[k] = ctx->geo.nc[k]; c->cell[k] = ctx->geo.cell[k]; }
  c->tessellated = ctx->tessellated; c->listed = h->listed; c->rows_asym = h->rows_asym;
  return 0;
}
int dml_reset_try_depo(dml_ctx *ctx) { ENTER(ctx);
  CKC(cudaMemsetAsync(&ctx->sc->try_, 0, sizeof(long long), ctx->st));
  CKC(cudaMemsetAsync(&ctx->sc->depo, 0, sizeof(long long), ctx->st));
  return 0;
}

int dml_test_update(dml_ctx *ctx) { ENTER(ctx); TRY(enq_test_update(ctx)); return finish(ctx); }
int dml_fuerza(dml_ctx *ctx) { ENTER(ctx);
  TRY(pull_scal(ctx));
  if (!ctx->hsc->listed) FAIL("fuerza called without a neighbour list");
  TRY(enq_fuerza(ctx)); return finish(ctx);
}
int dml_ermak_a(dml_ctx *ctx) { ENTER(ctx); TRY(enq_integrate(ctx, true)); return finish(ctx); }
int dml_ermak_b(dml_ctx *ctx) { ENTER(ctx);
  TRY(enq_ermak_b(ctx));
  return finish(ctx);
}
int dml_cbrownian_hs(dml_ctx *ctx) { ENTER(ctx); TRY(enq_integrate(ctx, false)); return finish(ctx); }
int dml_overlap_moveback(dml_ctx *ctx) { ENTER(ctx);
  TRY(pull_scal(ctx));
  if (!ctx->hsc->listed) FAIL("overlap_moveback called without a neighbour list");
  TRY(enq_overlap(ctx)); return finish(ctx);
}
int dml_msd_book(dml_ctx *ctx) { ENTER(ctx); LAUNCH(K_MISC, k_msd_book, 1, 1, ctx->sc); return 0; }
int dml_promote(dml_ctx *ctx) { ENTER(ctx); TRY(enq_promote(ctx)); return finish(ctx); }
int dml_gcmc_run(dml_ctx *ctx) { ENTER(ctx); TRY(gcmc_run_impl(ctx)); return finish(ctx); }
int dml_calc_rho(dml_ctx *ctx, double *rho) { ENTER(ctx);
  TRY(enq_calc_rho(ctx));
  TRY(finish(ctx));
  if (rho) *rho = ctx->hsc->rho;
  return 0;
}
int dml_maxz(dml_ctx *ctx, double *zmax) { ENTER(ctx);
  TRY(enq_maxz(ctx));
  TRY(finish(ctx));
  if (zmax) *zmax = ctx->hsc->zmax;
  return 0;
}
int dml_bloques(dml_ctx *ctx, int32_t nchunk, const double *chunk_pos, const double *chunk_pos_old, double dist, double rhomedia, int32_t *fired) { ENTER(ctx);
  TRY(do_bloques(ctx, nchunk, chunk_pos, chunk_pos_old, dist, rhomedia, fired));
  return finish(ctx);
}
int dml_set_chunk_template(dml_ctx *ctx, int32_t nchunk, const double *chunk_pos, const double *chunk_pos_old, double dist, double rhomedia) { ENTER(ctx);
  ctx->ch_pos.assign(chunk_pos, chunk_pos + (size_t)nchunk * 3);
  ctx->ch_pos_old.assign(chunk_pos_old, chunk_pos_old + (size_t)nchunk * 3);
  ctx->ch_dist = dist; ctx->ch_rhomedia = rhomedia; ctx->have_chunk = true;
  return 0;
}
int dml_step(dml_ctx *ctx, int32_t nsteps) { ENTER(ctx);
  for (int i = 0; i < nsteps; ++i) {
    TRY(enq_step(ctx));
    if ((i & 15) == 15) TRY(finish(ctx));              // periodic error check / storage growth; no other host round trip
  }
  return finish(ctx);
}

// Ensemble of independent replicas on one GPU (BASELINE config 5): every ctx has its own stream, so enqueueing the
// same step of all replicas from ONE host thread lets their (latency-bound, machine-underfilling) kernels overlap on the device.
// Part a of every replica is enqueued before the first part b (the only host round trip: reservoir 2 reads rho) is waited for.
int dml_ensemble_step(dml_ctx **ctxs, int32_t nctx, int32_t nsteps) {
  if (!ctxs || nctx <= 0) return -1;
  for (int i = 0; i < nsteps; ++i) {
    for (int r = 0; r < nctx; ++r) { dml_ctx *ctx = ctxs[r]; ENTER(ctx); TRY(launch_step_a(ctx)); }
    for (int r = 0; r < nctx; ++r) { dml_ctx *ctx = ctxs[r]; ENTER(ctx); TRY(enq_step_b(ctx)); if ((i & 15) == 15) TRY(finish(ctx)); }
  }
  for (int r = 0; r < nctx; ++r) { dml_ctx *ctx = ctxs[r]; ENTER(ctx); TRY(finish(ctx)); }
  return 0;
}
// a replica of an ensemble leaves room for the others: one block per SM for the cooperative kernels (a grid that fills the
// machine cannot overlap with another replica's; measured with 4 x 200 k boxes: 7.5e8 -> 9.6e8 particle-steps/s)
int dml_set_ensemble_member(dml_ctx *ctx, int32_t on) { ENTER(ctx);
  int dev = 0, nsm = 0;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  int b1 = 0, b1i = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, k_test_update_coop<false>, TPB, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1i, k_test_update_coop<true>, TPB, 0);
  ctx->coop_grid_tu = nsm * (on ? 1 : std::min(b1, 4));
  ctx->coop_grid_tu_bi = nsm * (on ? 1 : std::min(b1i, 4));
  return 0;
}

// Host-buffer path of a host that keeps the frame (what salida() reads: positions and element, src/dana.F90:1151-1157): new
// positions in, frame out; membership, velocities and the neighbour structures stay resident.
int dml_upload_positions(dml_ctx *ctx, int32_t n, const double *pos, const double *pos_old) { ENTER(ctx);
  if (n != ctx->n) FAIL("dml_upload_positions: n must be the current number of slots (use dml_upload for structural changes)");
  if (!pos) FAIL("dml_upload_positions: pos is required");
  const size_t n3 = (size_t)n * 3;
  CKC(ctx->stage_d.ensure(n3, ctx->st));
  CKC(cudaMemcpyAsync(ctx->stage_d.p, pos, n3 * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  if (pos_old) CKC(cudaMemcpyAsync(ctx->pos_old.p, pos_old, n3 * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  LAUNCH(K_PACK, k_repos, nblk(n), TPB, ctx->posm.p, ctx->stage_d.p, n, ctx->sc);
  return 0;                                               // stream-ordered: the next call on this ctx sees the new positions
}
int dml_download_frame(dml_ctx *ctx, int32_t n, double *pos, int32_t *z) { ENTER(ctx);
  if (n > ctx->n) FAIL("dml_download_frame: n exceeds the number of slots");
  const size_t n3 = (size_t)n * 3;
  CKC(ctx->stage_d.ensure(n3, ctx->st)); CKC(ctx->stage_i.ensure((size_t)n * 2, ctx->st));
  LAUNCH(K_PACK, k_unpack, nblk(n), TPB, ctx->posm.p, pos ? ctx->stage_d.p : nullptr, z ? ctx->stage_i.p : nullptr, (int *)nullptr, n);
  if (pos) CKC(cudaMemcpyAsync(pos, ctx->stage_d.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
  if (z) CKC(cudaMemcpyAsync(z, ctx->stage_i.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));
  return 0;
}

int dml_get_cells(dml_ctx *ctx, int32_t n, int32_t *cell_xyz, int32_t *chain_pos) { ENTER(ctx);
  if (!ctx->binned) FAIL("dml_get_cells: call dml_test_update first");
  if (!ctx->tessellated) FAIL("dml_get_cells: the box has no cell lists (fewer than 4 cells on every axis, Cells.F90:231)");
  TRY(enq_materialize_rows(ctx));
  TRY(enq_sort_cells(ctx, 1));
  CKC(cudaMemsetAsync(ctx->chain_pos.p, 0xff, (size_t)ctx->cap * sizeof(int), ctx->st));
  LAUNCH(K_MISC, k_chain_pos, nblk(ctx->nct, 128), 128, ctx->cell_start.p, ctx->sorted_slot.p, ctx->chain_pos.p, ctx->nct);
  std::vector<int> lin(n);
  CKC(cudaMemcpyAsync(lin.data(), ctx->cell_of.p, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaMemcpyAsync(chain_pos, ctx->chain_pos.p, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));
  for (int i = 0; i < n; ++i) {
    if (lin[i] < 0) { cell_xyz[3 * i] = cell_xyz[3 * i + 1] = cell_xyz[3 * i + 2] = -1; continue; }
    int r = lin[i] / ctx->geo.hd[0];
    cell_xyz[3 * i] = lin[i] % ctx->geo.hd[0]; cell_xyz[3 * i + 1] = r % ctx->geo.hd[1]; cell_xyz[3 * i + 2] = r / ctx->geo.hd[1];
  }
  return 0;
}

int dml_get_neighbors(dml_ctx *ctx, int32_t n, int32_t width, int32_t *nn, int32_t *rows) { ENTER(ctx);
  TRY(enq_materialize_rows(ctx));
  TRY(pull_scal(ctx));
  if (!ctx->hsc->listed) FAIL("no neighbour list");
  std::vector<RowHead> rh(n);
  std::vector<int> cols((size_t)std::max(ctx->hsc->cols_used, 1));
  CKC(cudaMemcpyAsync(rh.data(), ctx->rh.p, n * sizeof(RowHead), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaMemcpyAsync(cols.data(), ctx->cols.p, cols.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));
  int rc = 0;
  for (int i = 0; i < n; ++i) {
    nn[i] = rh[i].len;
    for (int m = 0; m < rh[i].len; ++m) { if (m >= width) { rc = 1; break; } rows[(size_t)i * width + m] = cols[rh[i].start + m]; }
  }
  return rc;
}
int dml_set_neighbors(dml_ctx *ctx, int32_t n, int32_t width, const int32_t *nn, const int32_t *rows) { ENTER(ctx);
  if (n > ctx->n) FAIL("dml_set_neighbors: n exceeds the number of slots");
  // same layout as the device build: a row that fits (with the gcmc slack) sits in its slot's ROW_W entries, longer ones in the tail
  const int tail0 = ctx->cap * ROW_W;
  std::vector<RowHead> rh(ctx->n);
  memset(rh.data(), 0, rh.size() * sizeof(RowHead));        // zero build distances, q5 = 0 (no near list): every entry is looked at
  for (auto &h : rh) { for (int k = 0; k < 4; ++k) { h.near[k] = -1; h.nbq[k] = 255; } }
  size_t off = (size_t)tail0;
  for (int i = 0; i < n; ++i) {
    if (nn[i] < 0 || nn[i] + ctx->row_slack > 65535) FAIL("dml_set_neighbors: row length out of range");
    rh[i].len = (unsigned short)nn[i];
    if (nn[i] + ctx->row_slack <= ROW_W) { rh[i].start = i * ROW_W; rh[i].cap = ROW_W; }
    else { rh[i].start = (int)off; rh[i].cap = (unsigned short)(nn[i] + ctx->row_slack); off += (size_t)rh[i].cap; }
  }
  CKC(ctx->cols.ensure(off + 4096, ctx->st, true)); CKC(ctx->rev_cols.ensure(ctx->cols.cap, ctx->st));
  CKC(ctx->bq.ensure(ctx->cols.cap, ctx->st)); CKC(ctx->rev_bq.ensure(ctx->cols.cap, ctx->st));
  std::vector<int> cols(off, -1);
  for (int i = 0; i < n; ++i) for (int m = 0; m < nn[i]; ++m) cols[(size_t)rh[i].start + m] = rows[(size_t)i * width + m];
  CKC(cudaMemsetAsync(ctx->bq.p, 0, ctx->bq.cap, ctx->st));   // caller's rows carry no build distances: never skip
  CKC(cudaMemcpyAsync(ctx->rh.p, rh.data(), ctx->n * sizeof(RowHead), cudaMemcpyHostToDevice, ctx->st));
  CKC(cudaMemcpyAsync(ctx->cols.p, cols.data(), off * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  TRY(pull_scal(ctx));
  ctx->hsc->cols_used = (int)off;
  ctx->hsc->rows_pending = 0;
  ctx->hsc->listed = 1; ctx->hsc->rows_asym = 2; ctx->hsc->rev_valid = 0;   // caller's rows: make no symmetry assumption
  ctx->hsc->cols_cap = (int)std::min<size_t>(ctx->cols.cap, 0x7fffffff);
  TRY(push_scal(ctx));
  CKC(cudaStreamSynchronize(ctx->st));
  return 0;
}

static int set_overlap_queue(dml_ctx *ctx, int n, const int *qstart, int nvals, const double *vals) {
  CKC(ctx->rp_qstart.ensure((size_t)ctx->cap + 1, ctx->st)); CKC(ctx->ov_draws.ensure(ctx->cap, ctx->st));
  CKC(ctx->rp_uovl.ensure((size_t)std::max(nvals, 1), ctx->st));
  // slots beyond n (none exist yet) get empty queues
  std::vector<int> qs((size_t)ctx->cap + 1, qstart[n]);
  std::copy(qstart, qstart + n + 1, qs.begin());
  CKC(cudaMemcpyAsync(ctx->rp_qstart.p, qs.data(), qs.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->st));
  if (nvals) CKC(cudaMemcpyAsync(ctx->rp_uovl.p, vals, (size_t)nvals * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  CKC(cudaMemsetAsync(ctx->ov_draws.p, 0, (size_t)ctx->cap * sizeof(int), ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));                      // qs goes out of scope
  ctx->have_rp_ovl = true;
  return 0;
}
int dml_set_replay_integrator(dml_ctx *ctx, int32_t n, const double *gauss, const double *unif_pbc, const double *unif_ovl) { ENTER(ctx);
  if (n > ctx->cap) FAIL("replay arrays exceed capacity");
  if (gauss) CKC(cudaMemcpyAsync(ctx->rp_gauss.p, gauss, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  if (unif_pbc) CKC(cudaMemcpyAsync(ctx->rp_upbc.p, unif_pbc, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  else CKC(cudaMemsetAsync(ctx->rp_upbc.p, 0, (size_t)n * sizeof(double), ctx->st));
  if (unif_ovl) {                                          // one value per slot: a queue of length 1 for every slot
    std::vector<int> qs((size_t)n + 1);
    for (int i = 0; i <= n; ++i) qs[i] = i;
    TRY(set_overlap_queue(ctx, n, qs.data(), n, unif_ovl));
  }
  CKC(cudaStreamSynchronize(ctx->st));
  ctx->have_rp = gauss != nullptr;
  return 0;
}
int dml_set_replay_overlap(dml_ctx *ctx, int32_t n, const int32_t *qstart, int32_t nvals, const double *vals) {
  ENTER(ctx);
  if (n > ctx->cap || !qstart || nvals < 0 || (nvals > 0 && !vals)) FAIL("dml_set_replay_overlap: bad arguments");
  TRY(set_overlap_queue(ctx, n, qstart, nvals, vals));
  CKC(cudaStreamSynchronize(ctx->st));
  return 0;
}
int dml_set_replay_gcmc(dml_ctx *ctx, int32_t nu, const double *unif, int32_t ng, const double *gauss) { ENTER(ctx);
  CKC(ctx->rp_gu.ensure((size_t)std::max(nu, 1), ctx->st)); CKC(ctx->rp_gg.ensure((size_t)std::max(ng, 1), ctx->st));
  if (nu) CKC(cudaMemcpyAsync(ctx->rp_gu.p, unif, (size_t)nu * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  if (ng) CKC(cudaMemcpyAsync(ctx->rp_gg.p, gauss, (size_t)ng * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));
  ctx->rp_nu = nu; ctx->rp_ng = ng;
  return 0;
}

// ---- slab decomposition over NCCL (dml_slab.cuh) -----------------------------------------------------------------
#define NCK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { ctx->err = std::string(#call) + ": " + (nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r_) : "nccl error"); return -1; } } while (0)

int dml_comm_unique_id(void *id128) {
  NcclApi *N = nccl_api();
  if (!N) return -1;
  ncclUniqueId id;
  if (N->GetUniqueId(&id) != ncclSuccess) return -2;
  memcpy(id128, &id, sizeof id);
  return 0;
}
int dml_comm_init(dml_ctx *ctx, const void *id128, int32_t rank, int32_t nranks) { ENTER(ctx);
  NcclApi *N = nccl_api();
  if (!N) FAIL("libnccl.so.2 not found");
  ncclUniqueId id; memcpy(&id, id128, sizeof id);
  NCK(N->CommInitRank(&ctx->comm, nranks, id, rank));
  ctx->rank = rank; ctx->nranks = nranks;
  return 0;
}
// host-side planning: z cuts that give every slab the same number of particles (nranks+1 values, cuts[0]=-inf side = lo)
int dml_slab_plan(int32_t n, const double *z, int32_t nranks, double lo, double hi, double *cuts) {
  std::vector<double> zs(z, z + n);
  std::sort(zs.begin(), zs.end());
  cuts[0] = lo; cuts[nranks] = hi;
  for (int k = 1; k < nranks; ++k) {
    size_t i = (size_t)((double)n * k / nranks);
    cuts[k] = n ? 0.5 * (zs[std::min<size_t>(i, n - 1)] + zs[i ? i - 1 : 0]) : lo + (hi - lo) * k / nranks;
  }
  return 0;
}
static int slab_exchange(dml_ctx *ctx, bool with_uid, bool measure = false) {
  NcclApi *N = nccl_api();
  const bool has_lo = ctx->rank > 0, has_hi = ctx->rank < ctx->nranks - 1;
  {
    const int nlo = has_lo ? ctx->nsend_lo : 0, nhi = has_hi ? ctx->nsend_hi : 0;
    if (nlo + nhi) LAUNCH(K_PACK, k_slab_pack2, nblk(nlo + nhi), TPB, ctx->posm.p, ctx->uid.p, ctx->send_lo.p, nlo, ctx->pack_lo.p, with_uid ? ctx->pack_uid_lo.p : nullptr,
                          ctx->send_hi.p, nhi, ctx->pack_hi.p, with_uid ? ctx->pack_uid_hi.p : nullptr);
  }
  NCK(N->GroupStart());
  if (has_hi) {
    if (ctx->nsend_hi) NCK(N->Send(ctx->pack_hi.p, (size_t)ctx->nsend_hi * 4, ncclDouble, ctx->rank + 1, ctx->comm, ctx->st));
    if (ctx->nrecv_hi) NCK(N->Recv(ctx->posm.p + ctx->ghost_hi_first, (size_t)ctx->nrecv_hi * 4, ncclDouble, ctx->rank + 1, ctx->comm, ctx->st));
    if (with_uid && ctx->nsend_hi) NCK(N->Send(ctx->pack_uid_hi.p, ctx->nsend_hi, ncclInt, ctx->rank + 1, ctx->comm, ctx->st));
    if (with_uid && ctx->nrecv_hi) NCK(N->Recv(ctx->uid.p + ctx->ghost_hi_first, ctx->nrecv_hi, ncclInt, ctx->rank + 1, ctx->comm, ctx->st));
  }
  if (has_lo) {
    if (ctx->nsend_lo) NCK(N->Send(ctx->pack_lo.p, (size_t)ctx->nsend_lo * 4, ncclDouble, ctx->rank - 1, ctx->comm, ctx->st));
    if (ctx->nrecv_lo) NCK(N->Recv(ctx->posm.p + ctx->ghost_lo_first, (size_t)ctx->nrecv_lo * 4, ncclDouble, ctx->rank - 1, ctx->comm, ctx->st));
    if (with_uid && ctx->nsend_lo) NCK(N->Send(ctx->pack_uid_lo.p, ctx->nsend_lo, ncclInt, ctx->rank - 1, ctx->comm, ctx->st));
    if (with_uid && ctx->nrecv_lo) NCK(N->Recv(ctx->uid.p + ctx->ghost_lo_first, ctx->nrecv_lo, ncclInt, ctx->rank - 1, ctx->comm, ctx->st));
  }
  NCK(N->GroupEnd());
  int ng = ctx->nrecv_lo + ctx->nrecv_hi;
  if (ng) LAUNCH(K_PACK, k_slab_mark, nblk(ng), TPB, ctx->posm.p, with_uid ? ctx->slot_b.p : nullptr, with_uid ? ctx->halo_of.p : nullptr, ctx->n_owned, ng,
                 measure ? ctx->old_cg.p : nullptr, ctx->sc, ctx->geo);
  return 0;
}
// Selects the owned particles within one list radius of each face, exchanges them with rank-1 / rank+1 and appends the
// received ones as ghost slots [n_owned, n_owned+n_ghost).  One host synchronisation (the four counts).
static int slab_refresh_ghosts(dml_ctx *ctx) {
  NcclApi *N = nccl_api();
  const bool has_lo = ctx->rank > 0, has_hi = ctx->rank < ctx->nranks - 1;
  const int n = ctx->n_owned;
  const double w = ctx->cfg.rcut + ctx->cfg.nb_dcut;
  CKC(ctx->send_lo.ensure(ctx->cap, ctx->st)); CKC(ctx->send_hi.ensure(ctx->cap, ctx->st)); CKC(ctx->slab_counts.ensure(8, ctx->st));
  CKC(cudaMemsetAsync(ctx->slab_counts.p, 0, 4 * sizeof(int), ctx->st));
  LAUNCH(K_PACK, k_slab_select, nblk(n), TPB, ctx->posm.p, ctx->send_lo.p, ctx->send_hi.p, ctx->slab_counts.p, ctx->zlo, ctx->zhi, w, has_lo ? 1 : 0, has_hi ? 1 : 0, n);
  NCK(N->GroupStart());
  if (has_hi) { NCK(N->Send(ctx->slab_counts.p + 1, 1, ncclInt, ctx->rank + 1, ctx->comm, ctx->st)); NCK(N->Recv(ctx->slab_counts.p + 3, 1, ncclInt, ctx->rank + 1, ctx->comm, ctx->st)); }
  if (has_lo) { NCK(N->Send(ctx->slab_counts.p + 0, 1, ncclInt, ctx->rank - 1, ctx->comm, ctx->st)); NCK(N->Recv(ctx->slab_counts.p + 2, 1, ncclInt, ctx->rank - 1, ctx->comm, ctx->st)); }
  NCK(N->GroupEnd());
  int hc[4];
  CKC(cudaMemcpyAsync(hc, ctx->slab_counts.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));
  ctx->nsend_lo = hc[0]; ctx->nsend_hi = hc[1]; ctx->nrecv_lo = has_lo ? hc[2] : 0; ctx->nrecv_hi = has_hi ? hc[3] : 0;
  if (n + ctx->nrecv_lo + ctx->nrecv_hi > ctx->cap) FAIL("slot capacity exhausted by ghost particles (dml_config.capacity)");
  ctx->ghost_lo_first = n; ctx->ghost_hi_first = n + ctx->nrecv_lo;
  CKC(ctx->pack_lo.ensure((size_t)std::max(ctx->nsend_lo, 1), ctx->st)); CKC(ctx->pack_hi.ensure((size_t)std::max(ctx->nsend_hi, 1), ctx->st));
  CKC(ctx->pack_uid_lo.ensure((size_t)std::max(ctx->nsend_lo, 1), ctx->st)); CKC(ctx->pack_uid_hi.ensure((size_t)std::max(ctx->nsend_hi, 1), ctx->st));
  TRY(slab_exchange(ctx, true));
  const int ng = ctx->nrecv_lo + ctx->nrecv_hi;
  // ghosts start with pos_old = old_cg = pos, no velocity and no row
  if (ng) LAUNCH(K_PACK, k_slab_ghost_init, nblk(ng), TPB, ctx->posm.p, ctx->pos_old.p, ctx->old_cg.p, ctx->vel.p, ctx->acel.p, ctx->rh.p, n, ng);
  ctx->n = n + ng;
  return 0;
}
int dml_slab_setup(dml_ctx *ctx, double zlo, double zhi) { ENTER(ctx);
  NcclApi *N = nccl_api();
  if (!ctx->comm || !N) FAIL("dml_slab_setup: call dml_comm_init first");
  TRY(finish(ctx));
  ctx->zlo = zlo; ctx->zhi = zhi;
  const int n = ctx->n_owned = ctx->n;
  {                                                        // holes of the uploaded arrays (z == 0 slots)
    CKC(ctx->slab_counts.ensure(8, ctx->st)); CKC(ctx->mig_holes.ensure(ctx->cap, ctx->st));
    CKC(cudaMemsetAsync(ctx->slab_counts.p, 0, 8 * sizeof(int), ctx->st));
    LAUNCH(K_PACK, k_slab_holes, nblk(n), TPB, ctx->posm.p, ctx->mig_holes.p, ctx->slab_counts.p, n);
    CKC(cudaMemcpyAsync(&ctx->slab_holes, ctx->slab_counts.p + 6, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CKC(cudaStreamSynchronize(ctx->st));
  }
  TRY(slab_refresh_ghosts(ctx));
  const int ng = ctx->n - n;
  TRY(pull_scal(ctx));
  ctx->hsc->n_slots = ctx->n; ctx->hsc->b_amax = ctx->n; ctx->hsc->nat_sys += ng; ctx->hsc->listed = 0; ctx->hsc->rows_pending = 0;
  TRY(push_scal(ctx));
  CKC(cudaStreamSynchronize(ctx->st));
  ctx->slab_ready = true;
  return 0;
}
// Particles that left [zlo, zhi) move to the neighbouring slab with their full state; arrivals refill the holes.
static int slab_migrate(dml_ctx *ctx) {
  NcclApi *N = nccl_api();
  const bool has_lo = ctx->rank > 0, has_hi = ctx->rank < ctx->nranks - 1;
  const int n = ctx->n_owned;
  CKC(ctx->mig_list_lo.ensure(ctx->cap, ctx->st)); CKC(ctx->mig_list_hi.ensure(ctx->cap, ctx->st)); CKC(ctx->mig_rc.ensure(2, ctx->st));
  CKC(ctx->mig_holes.ensure(ctx->cap, ctx->st));
  CKC(cudaMemsetAsync(ctx->slab_counts.p + 4, 0, 4 * sizeof(int), ctx->st));
  CKC(cudaMemsetAsync(ctx->mig_rc.p, 0, 2 * sizeof(int), ctx->st));
  LAUNCH(K_PACK, k_slab_mig_select, nblk(n), TPB, ctx->posm.p, ctx->mig_list_lo.p, ctx->mig_list_hi.p, ctx->slab_counts.p, ctx->zlo, ctx->zhi,
         has_lo ? 1 : 0, has_hi ? 1 : 0, n);
  NCK(N->GroupStart());
  if (has_hi) { NCK(N->Send(ctx->slab_counts.p + 5, 1, ncclInt, ctx->rank + 1, ctx->comm, ctx->st)); NCK(N->Recv(ctx->mig_rc.p + 1, 1, ncclInt, ctx->rank + 1, ctx->comm, ctx->st)); }
  if (has_lo) { NCK(N->Send(ctx->slab_counts.p + 4, 1, ncclInt, ctx->rank - 1, ctx->comm, ctx->st)); NCK(N->Recv(ctx->mig_rc.p + 0, 1, ncclInt, ctx->rank - 1, ctx->comm, ctx->st)); }
  NCK(N->GroupEnd());
  int hs[2], hr[2];
  CKC(cudaMemcpyAsync(hs, ctx->slab_counts.p + 4, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaMemcpyAsync(hr, ctx->mig_rc.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));
  const int nsl = hs[0], nsh = hs[1], nrl = has_lo ? hr[0] : 0, nrh = has_hi ? hr[1] : 0;
  if (nsl + nsh + nrl + nrh == 0) return 0;
  const int nholes = ctx->slab_holes + nsl + nsh, narr = nrl + nrh;
  if (n + std::max(0, narr - nholes) > ctx->cap) FAIL("slot capacity exhausted by migrating particles (dml_config.capacity)");
  CKC(ctx->mig_sd_lo.ensure((size_t)std::max(nsl, 1) * MIG_D, ctx->st)); CKC(ctx->mig_sd_hi.ensure((size_t)std::max(nsh, 1) * MIG_D, ctx->st));
  CKC(ctx->mig_si_lo.ensure((size_t)std::max(nsl, 1), ctx->st)); CKC(ctx->mig_si_hi.ensure((size_t)std::max(nsh, 1), ctx->st));
  CKC(ctx->mig_rd.ensure((size_t)std::max(narr, 1) * MIG_D, ctx->st)); CKC(ctx->mig_ri.ensure((size_t)std::max(narr, 1), ctx->st));
  if (nsl) LAUNCH(K_PACK, k_slab_mig_pack, nblk(nsl), TPB, ctx->posm.p, ctx->vel.p, ctx->acel.p, ctx->old_cg.p, ctx->uid.p, ctx->rh.p, ctx->mig_list_lo.p, nsl, ctx->mig_sd_lo.p, ctx->mig_si_lo.p);
  if (nsh) LAUNCH(K_PACK, k_slab_mig_pack, nblk(nsh), TPB, ctx->posm.p, ctx->vel.p, ctx->acel.p, ctx->old_cg.p, ctx->uid.p, ctx->rh.p, ctx->mig_list_hi.p, nsh, ctx->mig_sd_hi.p, ctx->mig_si_hi.p);
  LAUNCH(K_PACK, k_slab_holes, nblk(n), TPB, ctx->posm.p, ctx->mig_holes.p, ctx->slab_counts.p, n);
  NCK(N->GroupStart());
  if (nsh) { NCK(N->Send(ctx->mig_sd_hi.p, (size_t)nsh * MIG_D, ncclDouble, ctx->rank + 1, ctx->comm, ctx->st)); NCK(N->Send(ctx->mig_si_hi.p, nsh, ncclInt, ctx->rank + 1, ctx->comm, ctx->st)); }
  if (nrh) { NCK(N->Recv(ctx->mig_rd.p + (size_t)nrl * MIG_D, (size_t)nrh * MIG_D, ncclDouble, ctx->rank + 1, ctx->comm, ctx->st)); NCK(N->Recv(ctx->mig_ri.p + nrl, nrh, ncclInt, ctx->rank + 1, ctx->comm, ctx->st)); }
  if (nsl) { NCK(N->Send(ctx->mig_sd_lo.p, (size_t)nsl * MIG_D, ncclDouble, ctx->rank - 1, ctx->comm, ctx->st)); NCK(N->Send(ctx->mig_si_lo.p, nsl, ncclInt, ctx->rank - 1, ctx->comm, ctx->st)); }
  if (nrl) { NCK(N->Recv(ctx->mig_rd.p, (size_t)nrl * MIG_D, ncclDouble, ctx->rank - 1, ctx->comm, ctx->st)); NCK(N->Recv(ctx->mig_ri.p, nrl, ncclInt, ctx->rank - 1, ctx->comm, ctx->st)); }
  NCK(N->GroupEnd());
  if (narr) LAUNCH(K_PACK, k_slab_mig_unpack, nblk(narr), TPB, ctx->posm.p, ctx->vel.p, ctx->acel.p, ctx->pos_old.p, ctx->old_cg.p, ctx->uid.p, ctx->slot_b.p,
                   ctx->fe.p, ctx->rh.p, ctx->halo_of.p, ctx->mig_rd.p, ctx->mig_ri.p, narr, 0, ctx->mig_holes.p, nholes, n);
  ctx->slab_holes = std::max(0, nholes - narr);
  ctx->n_owned = n + std::max(0, narr - nholes);
  return 0;
}
// test_update (Neighbor.F90:668-713) on the decomposed box: global decision, migration + ghost re-selection + rows at a rebuild
// test_update of the decomposed box, device part: displacements, merged all-gather, decision (and calc_rho with_rho)
static int slab_tu_enqueue(dml_ctx *ctx, bool with_rho) {
  NcclApi *N = nccl_api();
  tessellate(ctx);
  if (!ctx->tessellated) FAIL("box smaller than 4 cells in every direction");
  int n = ctx->n, nct = ctx->nct;
  if ((size_t)nct + 2 > ctx->cell_start.cap) {
    CKC(ctx->cell_cnt.ensure(nct + 1, ctx->st)); CKC(ctx->cell_start.ensure(nct + 2, ctx->st)); CKC(ctx->cell_cur.ensure(nct + 1, ctx->st));
    CKC(cudaMemsetAsync(ctx->cell_cnt.p, 0, ctx->cell_cnt.cap * sizeof(int), ctx->st));
    CKC(cudaMemsetAsync(ctx->cell_cur.p, 0, ctx->cell_cur.cap * sizeof(int), ctx->st));
  }
  const int nb = std::min(nblk(n), 148 * 6);
  const size_t per = sizeof(SlabTU) / sizeof(double);
  CKC(ctx->part.ensure((size_t)2 * nb, ctx->st));
  if (!ctx->top2_own.p) { CKC(ctx->top2_own.ensure(per, ctx->st)); CKC(cudaMemsetAsync(ctx->top2_own.p, 0, sizeof(SlabTU), ctx->st)); }
  CKC(ctx->top2_all.ensure(per * ctx->nranks, ctx->st));
  SlabTU *own = reinterpret_cast<SlabTU *>(ctx->top2_own.p), *all = reinterpret_cast<SlabTU *>(ctx->top2_all.p);
  // second test_update of a step: F -> CG promotion and the census of calc_rho ride on the same all-gather (they come in front
  // of the rebuild decision like in the fused tail of the single-GPU step)
  if (with_rho) LAUNCH(K_PROMOTE, k_slab_promote_count, std::min(nblk(ctx->n_owned), 148 * 8), TPB, ctx->posm.p, ctx->sc, own->cnt, ctx->n_owned);
  LAUNCH(K_PBC_BIN, k_pbc_disp, nb, TPB, ctx->posm.p, ctx->pos_old.p, ctx->part.p, ctx->lay.p, ctx->sc, ctx->geo, n, ctx->n_owned, 2, ctx->cfg.nb_dcut, ctx->ph.r0_max, ctx->cfg.rcut,
         ctx->top2_own.p, ctx->use_dq ? ctx->dq.p : nullptr);
  NCK(N->AllGather(own, all, sizeof(SlabTU), ncclChar, ctx->comm, ctx->st));
  LAUNCH(K_TOP2, k_slab_tu_final, 1, 256, all, ctx->nranks, own, ctx->sc, ctx->lay.p, ctx->geo, ctx->cfg.nb_dcut, ctx->ph.r0_max, ctx->cfg.rcut, with_rho ? 1 : 0,
         ctx->geo.box[0] * ctx->geo.box[1]);
  return 0;
}
// host part: read the decision back; at a rebuild migrate, re-select the ghosts, sort the cells and build the rows
static int slab_tu_decide(dml_ctx *ctx) {
  TRY(pull_scal(ctx));
  if (ctx->hsc->need_rebuild) {
    slab_graphs_drop(ctx);                                // slot counts and ghost lists change: the captured segments are void
    ctx->slab_last_interval = ctx->slab_since; ctx->slab_since = 0;
    TRY(slab_migrate(ctx));
    TRY(slab_refresh_ghosts(ctx));
    TRY(pull_scal(ctx));
    ctx->hsc->n_slots = ctx->n; ctx->hsc->b_amax = ctx->n;
    TRY(push_scal(ctx));
    TRY(enq_sort_cells(ctx, 0));
    TRY(enq_materialize_rows(ctx, true));
  }
  ctx->binned = true;
  return 0;
}
// One segment of the decomposed step (everything between two host read-backs).  The first pass after the slab changed runs as
// plain launches (buffers grow, NCCL sets up its connections), the second is captured into a graph, later ones replay it: the
// segment's ~10 launches and its NCCL operations then cost one graph launch on the host and no launch gaps on the device.
static int slab_segment(dml_ctx *ctx, dml_ctx::SlabGraph &G, const std::function<int()> &enq) {
  if (!ctx->slab_graph_on || ctx->profiling) return enq();
  if (memcmp(&ctx->slab_geo, &ctx->geo, sizeof(Geo)) != 0) { slab_graphs_drop(ctx); ctx->slab_geo = ctx->geo; }
  if (!G.exec) {
    // capturing the two segments costs ~1 ms (measured: 0.4-0.7 ms each, the NCCL calls dominate) and a replayed step saves ~25 us:
    // only when the list is expected to live for more than ~40 steps (at 1 M particles and h = 1e-2 it lives 5-28: plain launches)
    const bool pays = ctx->slab_since >= (ctx->slab_last_interval >= 80 ? 1 : 40);
    if (!G.warm || !pays) { G.warm = true; return enq(); }
    const int64_t l0 = ctx->launches, s0 = ctx->step;
    cudaGraph_t g = nullptr;
    const auto t_cap0 = std::chrono::steady_clock::now();
    CKC(cudaStreamBeginCapture(ctx->st, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    const int rc = enq();
    ctx->capturing = false;
    cudaError_t e = cudaStreamEndCapture(ctx->st, &g);
    G.launches = ctx->launches - l0; G.steps = ctx->step - s0;
    ctx->launches = l0; ctx->step = s0;
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e == cudaSuccess && g) { e = cudaGraphInstantiate(&G.exec, g, 0); cudaGraphDestroy(g); }
    if (e != cudaSuccess || !G.exec) { cudaGetLastError(); G.exec = nullptr; ctx->slab_graph_on = false; return enq(); }   // not capturable here: plain launches from now on
    if (getenv("DML_SLAB_GRAPH_DEBUG"))
      fprintf(stderr, "[dml] rank %d: slab segment captured in %.1f us (%lld launches)\n", ctx->rank,
              std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_cap0).count(), (long long)G.launches);
  }
  ctx->launches += G.launches; ctx->step += G.steps;
  CKC(cudaGraphLaunch(G.exec, ctx->st));
  return 0;
}
// nsteps iterations of dana's loop body (dana.F90:173-265; Ermak integrator + piston) on the decomposed box
int dml_slab_step(dml_ctx *ctx, int32_t nsteps) { ENTER(ctx);
  if (!ctx->comm || !ctx->slab_ready) FAIL("dml_slab_step: call dml_comm_init and dml_slab_setup first");
  if (!ctx->cfg.integrador || ctx->cfg.reservoir != 1) FAIL("dml_slab_step: the decomposed box runs the Ermak integrator with the piston reservoir");
  ctx->rows_eager = true;
  struct Eager { dml_ctx *c; ~Eager() { c->rows_eager = false; } } eager_guard{ctx};
  for (int i = 0; i < nsteps; ++i) {
    // segment A: integrator, ghost refresh, pair force, ermak_b, first test_update up to the gathered decision
    TRY(slab_segment(ctx, ctx->sgA, [&]() -> int {
      const int ng = ctx->n - ctx->n_owned;
      if (ng) LAUNCH(K_PACK, k_slab_ghost_save, nblk(ng), TPB, ctx->posm.p, ctx->old_cg.p, ctx->n_owned, ng);
      TRY(enq_integrate(ctx, true));
      TRY(slab_exchange(ctx, false, true));               // ghosts at their new positions; their moves enter the skip bound
      TRY(enq_fuerza(ctx, true));
      if (ctx->cfg.strict_order || !ctx->fuse_ermak_b)
        TRY(enq_ermak_b(ctx));
      return slab_tu_enqueue(ctx, false);
    }));
    TRY(slab_tu_decide(ctx));                             // host read-back; migration and list rebuild when due
    // segment B: overlap_moveback, ghost refresh, second test_update (with promotion and the census of calc_rho)
    TRY(slab_segment(ctx, ctx->sgB, [&]() -> int {
      TRY(enq_overlap(ctx, true));
      TRY(slab_exchange(ctx, false, true));
      return slab_tu_enqueue(ctx, true);
    }));
    TRY(slab_tu_decide(ctx));
    TRY(enq_maxz(ctx));
    ctx->slab_since++;
    ctx->t = ctx->t + ctx->cfg.h;
  }
  return finish(ctx);
}
// per-step refresh of the ghost positions from their owners (same lists as the last dml_slab_setup)
int dml_slab_halo_exchange(dml_ctx *ctx) { ENTER(ctx);
  if (!ctx->comm) FAIL("dml_slab_halo_exchange: no communicator");
  return slab_exchange(ctx, false);
}
int dml_slab_info(dml_ctx *ctx, int32_t *n_owned, int32_t *n_ghost, int32_t *nsend_lo, int32_t *nsend_hi) { ENTER(ctx);
  if (n_owned) *n_owned = ctx->n_owned;
  if (n_ghost) *n_ghost = ctx->nrecv_lo + ctx->nrecv_hi;
  if (nsend_lo) *nsend_lo = ctx->nsend_lo;
  if (nsend_hi) *nsend_hi = ctx->nsend_hi;
  return 0;
}


// ---- output reductions and observables ---------------------------------------------------------------
int dml_salida_sums(dml_ctx *ctx, double *energia, double *energia_ref, double *temp, int32_t *n_mobile) { ENTER(ctx);
  const int n = ctx->n;
  const int nb = std::min(nblk(n, OBS_TPB), 148 * 4);
  CKC(ctx->obs_part.ensure((size_t)4 * 148 * 4, ctx->st)); CKC(ctx->obs_out.ensure(4, ctx->st));
  if (!ctx->obs_ticket) { CKC(cudaMalloc(&ctx->obs_ticket, sizeof(unsigned int))); CKC(cudaMemsetAsync(ctx->obs_ticket, 0, sizeof(unsigned int), ctx->st)); }
  LAUNCH(K_MISC, k_salida_sums, nb, OBS_TPB, ctx->posm.p, ctx->fe.p, ctx->vel.p, ctx->ph, n, ctx->obs_part.p, ctx->obs_ticket, ctx->obs_out.p);
  double out[4];
  CKC(cudaMemcpyAsync(out, ctx->obs_out.p, sizeof(out), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));
  if (energia) *energia = out[0];
  if (energia_ref) *energia_ref = out[1];
  if (temp) *temp = out[2] / (out[3] * 3.0 * ctx->cfg.kB_ui);          // kion, dana.F90:1374
  if (n_mobile) *n_mobile = (int32_t)out[3];
  return 0;
}

int dml_density_profile(dml_ctx *ctx, double zlo, double zhi, int32_t nbins, int32_t type_mask, int64_t *counts) { ENTER(ctx);
  if (nbins < 1 || nbins > OBS_MAX_BINS || !(zhi > zlo) || !counts) FAIL("dml_density_profile: need 1 <= nbins <= 8192, zhi > zlo and an output array");
  const int n = ctx->n;
  CKC(ctx->obs_counts.ensure(OBS_MAX_BINS, ctx->st));
  CKC(cudaMemsetAsync(ctx->obs_counts.p, 0, (size_t)nbins * sizeof(unsigned long long), ctx->st));
  const double dz = (zhi - zlo) / (double)nbins;
  prof_begin(ctx, K_MISC);
  k_density_profile<<<std::min(nblk(n, OBS_TPB), 148 * 2), OBS_TPB, (size_t)nbins * sizeof(unsigned int), ctx->st>>>(ctx->posm.p, n, zlo, dz, nbins, type_mask,
                                                                                                                  ctx->obs_counts.p);
  prof_end(ctx, K_MISC);
  CKC(cudaMemcpyAsync(counts, ctx->obs_counts.p, (size_t)nbins * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));
  CKC(cudaGetLastError());
  return 0;
}

int dml_gr(dml_ctx *ctx, double rmax, int32_t nbins, int32_t type_mask, int64_t *counts, int32_t *n_selected) { ENTER(ctx);
  if (nbins < 1 || nbins > OBS_MAX_BINS || !(rmax > 0.0) || !counts) FAIL("dml_gr: need 1 <= nbins <= 8192, rmax > 0 and an output array");
  const int n = ctx->n;
  TRY(pull_scal(ctx));
  GrGrid gg;
  const double span[3] = {ctx->geo.box[0], ctx->geo.box[1], std::max(ctx->geo.box[2], ctx->hsc->zmax)};
  for (int k = 0; k < 3; ++k) {
    gg.nc[k] = std::max(1, (int)(span[k] / rmax));
    gg.nc[k] = std::min(gg.nc[k], 512);
    gg.cell[k] = span[k] / (double)gg.nc[k];
  }
  for (int k = 0; k < 2; ++k)
    if (ctx->geo.pbc[k] && gg.nc[k] < 3) FAIL("dml_gr: rmax too large for the periodic box (need box >= 3 rmax in x and y)");
  const int nct = gg.nc[0] * gg.nc[1] * gg.nc[2];
  CKC(ctx->gr_cell_of.ensure(std::max(n, 1), ctx->st)); CKC(ctx->gr_sorted.ensure(std::max(n, 1), ctx->st));
  CKC(ctx->gr_cnt.ensure((size_t)nct + 2, ctx->st)); CKC(ctx->gr_start.ensure((size_t)nct + 2, ctx->st));
  CKC(ctx->obs_counts.ensure(OBS_MAX_BINS, ctx->st));
  CKC(cudaMemsetAsync(ctx->obs_counts.p, 0, (size_t)nbins * sizeof(unsigned long long), ctx->st));
  CKC(cudaMemsetAsync(ctx->gr_cnt.p, 0, ((size_t)nct + 2) * sizeof(int), ctx->st));
  int *nsel = ctx->gr_cnt.p + nct + 1;
  const int nb = std::min(nblk(n, OBS_TPB), 148 * 8);
  LAUNCH(K_MISC, k_gr_bin, nb, OBS_TPB, ctx->posm.p, n, gg, type_mask, ctx->gr_cell_of.p, ctx->gr_cnt.p, nsel);
  TRY(scan_excl(ctx, ctx->gr_cnt.p, ctx->gr_start.p, nct, ctx->gr_start.p + nct, true, 2, 1));   // clears gr_cnt: reused as the fill cursor
  LAUNCH(K_MISC, k_gr_scatter, nb, OBS_TPB, ctx->posm.p, n, ctx->gr_cell_of.p, ctx->gr_start.p, ctx->gr_cnt.p, ctx->gr_sorted.p);
  int hsel = 0;
  CKC(cudaMemcpyAsync(&hsel, nsel, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));
  if (hsel > 0) {
    const double dr_bin = rmax / (double)nbins;
    prof_begin(ctx, K_MISC);
    k_gr_pairs<<<nblk(hsel, OBS_TPB), OBS_TPB, (size_t)nbins * sizeof(unsigned int), ctx->st>>>(ctx->gr_sorted.p, hsel, ctx->gr_start.p, gg, ctx->geo,
                                                                                              rmax * rmax, dr_bin, nbins, ctx->obs_counts.p);
    prof_end(ctx, K_MISC);
  }
  CKC(cudaMemcpyAsync(counts, ctx->obs_counts.p, (size_t)nbins * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));
  CKC(cudaGetLastError());
  if (n_selected) *n_selected = hsel;
  return 0;
}

int dml_membership_changes(dml_ctx *ctx, int32_t max_changes, int32_t *slot, int32_t *kind, int32_t *uid_now, int32_t *z_now, int32_t *n_changes) { ENTER(ctx);
  if (!n_changes || max_changes < 0) FAIL("dml_membership_changes: n_changes is required");
  if (!ctx->have_snap) { TRY(member_snapshot(ctx)); *n_changes = 0; CKC(cudaStreamSynchronize(ctx->st)); return 0; }
  CKC(ctx->mc_out.ensure((size_t)std::max(max_changes, 1), ctx->st)); CKC(ctx->mc_count.ensure(1, ctx->st));
  CKC(cudaMemsetAsync(ctx->mc_count.p, 0, sizeof(int), ctx->st));
  LAUNCH(K_MISC, k_member_diff, std::min(nblk(ctx->cap, OBS_TPB), 148 * 8), OBS_TPB, ctx->posm.p, ctx->uid.p, ctx->n, ctx->cap, ctx->snap_uid.p, ctx->snap_mb.p,
         max_changes, ctx->mc_out.p, ctx->mc_count.p);
  int cnt = 0;
  CKC(cudaMemcpyAsync(&cnt, ctx->mc_count.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
  CKC(cudaStreamSynchronize(ctx->st));
  *n_changes = cnt;
  const int m = std::min(cnt, (int)max_changes);
  if (m > 0) {
    std::vector<int4> rec(m);
    CKC(cudaMemcpy(rec.data(), ctx->mc_out.p, (size_t)m * sizeof(int4), cudaMemcpyDeviceToHost));
    std::sort(rec.begin(), rec.end(), [](const int4 &a, const int4 &b) { return a.x < b.x; });   // ascending slot
    for (int i = 0; i < m; ++i) {
      if (slot) slot[i] = rec[i].x;
      if (kind) kind[i] = rec[i].y;
      if (uid_now) uid_now[i] = rec[i].z;
      if (z_now) z_now[i] = rec[i].w;
    }
  }
  return 0;
}

int dml_profile(dml_ctx *ctx, int32_t enable) { ENTER(ctx);
  prof_collect(ctx);
  ctx->sg_n = -1;                                         // the captured step is re-recorded with / without event nodes
  ctx->profiling = enable != 0;
  ctx->prof_only = enable >= 2 ? enable - 2 : -1;          // enable = 2 + kernel id: events around that kernel only (undisturbed pipeline)
  if (enable) while (ctx->pool.size() < 8192) { ProfEv ev; cudaEventCreate(&ev.a); cudaEventCreate(&ev.b); ev.cls = 0; ctx->pool.push_back(ev); }
  return 0;
}
int dml_profile_get(dml_ctx *ctx, int32_t cls, double *ms, int64_t *launches, int32_t reset) { ENTER(ctx);
  prof_collect(ctx);
  double m = 0; int64_t l = 0;
  for (int i = 0; i < K_NKERN; ++i) if (cls == CLS_ALL || kern_cls[i] == cls) { m += ctx->prof_ms[i]; l += ctx->prof_n[i]; }
  if (ms) *ms = m;
  if (launches) *launches = l;
  if (reset) for (int i = 0; i < 32; ++i) { ctx->prof_ms[i] = 0; ctx->prof_n[i] = 0; }
  return 0;
}
int dml_profile_kernel(dml_ctx *ctx, int32_t kid, const char **name, double *ms, int64_t *launches) { ENTER(ctx);
  prof_collect(ctx);
  if (kid < 0 || kid >= K_NKERN) return 1;
  if (name) *name = kern_name[kid];
  if (ms) *ms = ctx->prof_ms[kid];
  if (launches) *launches = ctx->prof_n[kid];
  return 0;
}
int32_t dml_n_slots(dml_ctx *ctx) { ENTER(ctx); return ctx->n; }
int dml_set_strict_order(dml_ctx *ctx, int32_t on) { ENTER(ctx);
  ctx->cfg.strict_order = on ? 1 : 0;
  CKC(cudaMemsetAsync(&ctx->sc->rev_valid, 0, sizeof(int), ctx->st));   // the two kernels use differently scoped transposed rows
  return 0;
}
int64_t dml_launch_count(dml_ctx *ctx) { ENTER(ctx); return ctx->launches; }
void *dml_stream(dml_ctx *ctx) { return (void *)ctx->st; }

} // extern "C"
