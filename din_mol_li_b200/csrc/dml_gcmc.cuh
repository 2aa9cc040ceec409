// dml_gcmc.cuh — grand-canonical insertion/deletion trials on the device (gcmc_run, dana.F90:590-713) with
// the incremental neighbour-list maintenance they trigger (ngroup_attach_atom/ngroup_sort_atom/ngroup_cells_atom
// Neighbor.F90:173-318,550-603; ngroup_detach_atom 226-269; cgroup_attach/unsort Cells.F90:105-144,304-352;
// igroup index reuse Groups.F90:1083-1093).  Included by dml.cu after dml_ctx is defined.
//
// The nadj attempts of one call are sequentially dependent (n, the inserted atoms and the deletions feed the next
// attempt), so one thread block runs them in order; inside an attempt the O(N) scans of the reference become
// block-parallel searches: the overlap test visits the 27 cells around the trial point plus the atoms inserted
// earlier in this call, the "m-th atom inside the control volume in list order" is a two-level prefix count over
// the gcmc membership array kept in list order, and the lowest-free-slot searches are block-wide minima.
#pragma once

namespace dml {

struct GcmcArgs {
  double4 *posm; double4 *fe; unsigned char *fnz; double *vel, *acel, *pos_old, *old_cg;
  int *uid, *slot_b, *b_occ;
  const int *cell_of_unused; const int *cell_start; const int *sorted_slot; const double4 *sorted_posm;
  RowHead *rh; int *cols; unsigned char *bq; int cols_cap;
  int *gorder, *gpos, *gcc; int gorder_cap;
  int *pend;                       // slots inserted during this call, in insertion order
  const double *rp_u, *rp_g; int rp_nu, rp_ng;
  DevScal *sc; Geo g; Phys ph;
  double act, beta_kT;             // beta = sqrt(kB_ui*Tsist/mass) uses beta_kT = kB_ui(gems_constants)*Tsist
  int nadj, cap, listed, row_slack; unsigned int step;
};

constexpr int GB = 1024;           // threads of the gcmc block

__device__ __forceinline__ int block_excl_scan(int v, int *total, int *sh /*33 ints*/) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  __syncthreads();
  if (lane == 31) sh[w] = x;
  __syncthreads();
  if (w == 0) {
    int s = sh[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
    sh[lane] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) sh[32] = sh[31];
  int excl = x - v + (w ? sh[w - 1] : 0);
  __syncthreads();
  *total = sh[32];
  return excl;
}

// one entry of the replay stream or a Philox draw; thread 0 only
struct GcmcRng {
  const GcmcArgs *A; int iu, ig, ctr;
  __device__ double unif() {
    if (A->ph.rng_mode == 1) { if (iu >= A->rp_nu) { atomicCAS(&A->sc->err, 0, DML_E_REPLAY_EXHAUSTED); return 0.5; } return A->rp_u[iu++]; }
    if (A->ph.rng_mode == 2) return ref_ran(&A->sc->rr);
    Philox r; r.run(A->ph.seed, (unsigned int)(ctr++), A->step == STEP_FROM_DEVICE ? A->sc->istep : A->step, RS_GCMC, 0u); return r.u01(0);
  }
  __device__ double gauss() {
    if (A->ph.rng_mode == 1) { if (ig >= A->rp_ng) { atomicCAS(&A->sc->err, 0, DML_E_REPLAY_EXHAUSTED); return 0.0; } return A->rp_g[ig++]; }
    if (A->ph.rng_mode == 2) return ref_gasdev(&A->sc->rr);
    Philox r; r.run(A->ph.seed, (unsigned int)(ctr++), A->step == STEP_FROM_DEVICE ? A->sc->istep : A->step, RS_GCMC, 1u); double a, b; r.gauss2(a, b); return a;
  }
};

__device__ __forceinline__ bool in_volume(const double4 &p, double z0, double zmax) { return !(p.z < z0 || p.z > zmax); }

// Control-volume census (dana.F90:608-615) by the whole grid: gcc[c] = members of chunk c (1024 entries of the membership array)
// inside [z0, zmax].  Inside k_gcmc the same loop is 98 dependent gather + barrier rounds of ONE block at 100 k members (~150 us of
// the 426 us the kernel took); here it is one pass at full width.  k_gcmc adds the chunk counts up (and redoes the census itself
// only when it had to compact the membership array first).
__global__ void __launch_bounds__(256) k_gcmc_census(const double4 *__restrict__ posm, const int *__restrict__ gorder, int *__restrict__ gcc,
                                                     const DevScal *__restrict__ sc) {
  const int glen = sc->glen;
  const double z0 = sc->z0, zmax = sc->zmax;
  const int nchunk = (glen + GB - 1) / GB;
  __shared__ int s_cnt[8];
  for (int c = blockIdx.x; c < nchunk; c += gridDim.x) {
    int in = 0;
#pragma unroll
    for (int k = 0; k < GB / 256; ++k) {
      const int i = c * GB + k * 256 + threadIdx.x;
      if (i < glen) { const int s = gorder[i]; if (s >= 0) in += in_volume(ld_rec(&posm[s]), z0, zmax) ? 1 : 0; }
    }
    in = __reduce_add_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = in;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_cnt[w]; gcc[c] = t; }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(GB) k_gcmc(GcmcArgs A) {
  __shared__ int sh_scan[33];
  __shared__ int s_i[16];
  __shared__ double s_d[8];
  DevScal *sc = A.sc;
  const Geo &g = A.g;
  const int tid = threadIdx.x;
  GcmcRng rng = {&A, 0, 0, 0};
  const double z0 = sc->z0, zmax = sc->zmax;
  const double v = g.box[0] * g.box[1] * (zmax - z0);
  const double rc2 = g.rcut2;

  // ---- compaction of the membership array when it carries many tombstones or is nearly full ----
  int glen = sc->glen;
  bool compacted = false;
  if (sc->gtomb * 4 > glen || glen + A.nadj + 8 > A.gorder_cap) {
    compacted = true;
    int *tmp = A.gorder + A.gorder_cap;                  // second half of the allocation is scratch
    int outn = 0;
    for (int base = 0; base < glen; base += GB) {
      int i = base + tid;
      int s = i < glen ? A.gorder[i] : -1;
      int tot; int ex = block_excl_scan(s >= 0 ? 1 : 0, &tot, sh_scan);
      if (s >= 0) tmp[outn + ex] = s;
      outn += tot;
    }
    __syncthreads();
    for (int i = tid; i < outn; i += GB) { int s = tmp[i]; A.gorder[i] = s; A.gpos[s] = i; }
    __syncthreads();
    glen = outn;
    if (tid == 0) { sc->glen = glen; sc->gtomb = 0; sc->ghead = 0; }
    __syncthreads();
  }
  // ---- control-volume census: n and the per-1024 chunk counts (dana.F90:608-615) ----
  int n = 0;
  if (compacted) {
    for (int base = 0, c = 0; base < glen; base += GB, ++c) {
      int i = base + tid, in = 0;
      if (i < glen) { int s = A.gorder[i]; if (s >= 0) in = in_volume(ld_rec(&A.posm[s]), z0, zmax) ? 1 : 0; }
      int cnt = __syncthreads_count(in);
      if (tid == 0) A.gcc[c] = cnt;
      n += cnt;
    }
  } else {                                               // chunk counts come from k_gcmc_census
    const int nchunk0 = (glen + GB - 1) / GB;
    int part = 0;
    for (int c = tid; c < nchunk0; c += GB) part += A.gcc[c];
    part = __reduce_add_sync(0xffffffffu, part);
    if ((tid & 31) == 0) sh_scan[tid >> 5] = part;
    __syncthreads();
    for (int w = 0; w < GB / 32; ++w) n += sh_scan[w];
    __syncthreads();
  }
  if (tid == 0) A.gcc[(glen + GB - 1) / GB] = 0;
  int npend = 0;
  __syncthreads();

  for (int att = 0; att < A.nadj; ++att) {
    // -- template atom = first member of the list (dana.F90:622-624) --
    if (tid == 0) {
      int h = sc->ghead;
      while (h < glen && A.gorder[h] < 0) ++h;
      sc->ghead = h;
      s_i[0] = h < glen ? A.gorder[h] : -1;
      if (h >= glen) atomicCAS(&sc->err, 0, DML_E_NO_PARTICLES);
      s_i[1] = (s_i[0] >= 0 && rng.unif() < 0.5) ? 1 : 0;          // creation or destruction
    }
    __syncthreads();
    const int tmpl = s_i[0];
    if (tmpl < 0) break;
    const int create = s_i[1];
    __syncthreads();
    if (create) {
      if (tid == 0) {
        int acc = !(A.act * v / (n + 1) < rng.unif());
        s_i[2] = acc;
        if (acc) {
          s_d[0] = rng.unif() * g.box[0];
          s_d[1] = rng.unif() * g.box[1];
          s_d[2] = rng.unif() * (zmax - z0) + z0;
        }
      }
      __syncthreads();
      if (!s_i[2]) { __syncthreads(); continue; }
      const double rx = s_d[0], ry = s_d[1], rz = s_d[2];
      int rcx, rcy, rcz;
      bool okc = cell_index(g, rx, ry, rz, rcx, rcy, rcz);
      // -- overlap with any member of the gcmc group (dana.F90:638-653): 27 cells + atoms inserted in this call --
      int found = 0;
      if (okc && tid < 27) {
        int nx = (c_map[tid][0] + rcx - 1 + g.nc[0]) % g.nc[0] + 1;
        int ny = (c_map[tid][1] + rcy - 1 + g.nc[1]) % g.nc[1] + 1;
        int nz = (c_map[tid][2] + rcz - 1 + g.nc[2]) % g.nc[2] + 1;
        int nl = cell_lin(g, nx, ny, nz);
        for (int u = A.cell_start[nl]; u < A.cell_start[nl + 1]; ++u) {
          int s = A.sorted_slot[u];
          double4 p = ld_rec(&A.posm[s]);
          if (!(meta_of(p) & MF_GCMC)) continue;
          // distance(o%pos, r, o%pbc): r minus pos
          if (dist2_idnint(g, rx, ry, rz, p.x, p.y, p.z) < rc2) { found = 1; break; }
        }
      }
      for (int q = tid; q < npend; q += GB) {
        int s = A.pend[q];
        double4 p = ld_rec(&A.posm[s]);
        if (!(meta_of(p) & MF_GCMC)) continue;
        if (dist2_idnint(g, rx, ry, rz, p.x, p.y, p.z) < rc2) found = 1;
      }
      if (__syncthreads_or(found)) continue;
      // -- accepted: index assignment (lowest hole or append, Groups.F90:1083-1093) --
      if (tid == 0) { s_i[3] = 0x7fffffff; s_i[4] = 0x7fffffff; }
      __syncthreads();
      const int amax = sc->n_slots, nat_new = sc->nat_sys + 1, nlimbo = sc->nlimbo, b_amax = sc->b_amax;
      const bool hs_hole = amax >= nat_new + nlimbo, b_hole = b_amax >= nat_new;
      // lowest empty index: four independent loads per thread and trip, the block leaves at the first trip that found one, and the
      // search starts at the cursor below which nothing is free (a per-thread loop with a data-dependent exit ran ~100 dependent
      // round trips per search at 100 k slots: most of the 426 us this kernel took)
      if (hs_hole) {
        for (int base = sc->hole_lo; base < amax; base += 4 * GB) {
          int best = 0x7fffffff;
#pragma unroll
          for (int k = 3; k >= 0; --k) { const int i = base + k * GB + tid; if (i < amax && meta_of(ld_rec(&A.posm[i])) == 0) best = i; }
          if (best != 0x7fffffff) atomicMin(&s_i[3], best);
          __syncthreads();
          if (s_i[3] != 0x7fffffff) break;
          __syncthreads();
        }
      }
      if (b_hole) {
        for (int base = sc->bhole_lo; base < b_amax; base += 4 * GB) {
          int best = 0x7fffffff;
#pragma unroll
          for (int k = 3; k >= 0; --k) { const int i = base + k * GB + tid; if (i < b_amax && A.b_occ[i] == 0) best = i; }
          if (best != 0x7fffffff) atomicMin(&s_i[4], best);
          __syncthreads();
          if (s_i[4] != 0x7fffffff) break;
          __syncthreads();
        }
      }
      __syncthreads();
      if (tid == 0) {
        int ns = hs_hole ? s_i[3] : amax;
        int nb = b_hole ? s_i[4] : b_amax;
        if (ns >= A.cap || nb >= A.cap || ns == 0x7fffffff || nb == 0x7fffffff) { atomicCAS(&sc->err, 0, DML_E_CAPACITY); s_i[5] = -1; }
        else {
          s_i[5] = ns;
          if (!hs_hole) sc->n_slots = amax + 1; else sc->hole_lo = ns + 1;
          if (!b_hole) sc->b_amax = b_amax + 1; else sc->bhole_lo = nb + 1;
          n = n + 1;
          // velocity from the Maxwell-Boltzmann distribution lands on the LAST atom of the list (dana.F90:665-668, Q6)
          int l = glen - 1;
          while (l >= 0 && A.gorder[l] < 0) --l;
          int last = A.gorder[l];
          double4 pt = ld_rec(&A.posm[tmpl]);
          int zt = (int)(meta_of(pt) & MF_TYPE);
          double beta = sqrt(A.beta_kT / A.ph.mass[zt - 1]);
          // new atom copies the template (atom_asign, Groups.F90:484-502) BEFORE the draw touches anybody's velocity
          for (int k = 0; k < 3; ++k) {
            A.vel[3 * ns + k] = A.vel[3 * tmpl + k]; A.acel[3 * ns + k] = A.acel[3 * tmpl + k];
            A.old_cg[3 * ns + k] = 1e8;
          }
          st_rec(&A.fe[ns], ld_rec(&A.fe[tmpl]));             // force and epot of the template (atom_asign)
          A.fnz[ns] = 1;
          A.pos_old[3 * ns] = rx; A.pos_old[3 * ns + 1] = ry; A.pos_old[3 * ns + 2] = rz;
          for (int k = 0; k < 3; ++k) A.vel[3 * last + k] = beta * rng.gauss();
          double4 pn = {rx, ry, rz, meta_as_double(with_disp((long long)zt | MF_REF | MF_GCMC, DISP_INF))};
          st_rec(&A.posm[ns], pn);
          A.uid[ns] = sc->next_uid++;
          A.slot_b[ns] = nb; A.b_occ[nb] = 1;
          sc->nat_sys++; sc->nat_ref++; sc->nat_gcmc++; sc->gcmc_created++;
          // membership array (list order = creation order)
          A.gorder[glen] = ns; A.gpos[ns] = glen; A.gcc[glen / GB] += 1;
          if ((glen + 1) % GB == 0) A.gcc[(glen + 1) / GB] = 0;
          A.pend[npend] = ns;
        }
      }
      __syncthreads();
      // -- ngroup_sort_atom (Neighbor.F90:271-318): own row (strict <, stencil x chain order) and append to the rows of every ref
      //    atom within the list radius (<=).  Warp 0, lane = stencil cell: every lane walks its own chain (atoms inserted in this
      //    call that fell into the cell, most recent first, then the sorted segment), an exclusive prefix over the lanes gives the
      //    cell's place in the new row; an atom sits in one cell only, so the appends to other rows never collide.  (One thread
      //    doing the 27 cells in turn paid ~60 dependent round trips per accepted insertion.) --
      if (s_i[5] >= 0 && tid < 32) {
        const int ns = s_i[5];
        if (sc->listed) {
          const int base = sc->cols_used;
          const int lane = tid;
          int cb = 0, ce = 0, nl = -1;
          if (lane < 27 && okc) {
            int nx = (c_map[lane][0] + rcx - 1 + g.nc[0]) % g.nc[0] + 1;
            int ny = (c_map[lane][1] + rcy - 1 + g.nc[1]) % g.nc[1] + 1;
            int nz = (c_map[lane][2] + rcz - 1 + g.nc[2]) % g.nc[2] + 1;
            nl = cell_lin(g, nx, ny, nz);
            cb = A.cell_start[nl]; ce = A.cell_start[nl + 1];
          }
          const int npe = npend + 1;                        // pend[] already holds the new atom (skipped below)
          int off = 0, total = 0;
          bool ovf = false;
          for (int pass = 0; pass < 2; ++pass) {
            int cnt = 0;
            if (nl >= 0) {
              for (int q = npe + (ce - cb) - 1; q >= 0; --q) {
                int s;
                if (q >= ce - cb) {
                  s = A.pend[q - (ce - cb)];
                  if (s == ns) continue;
                  double4 pp = ld_rec(&A.posm[s]);
                  if (!(meta_of(pp) & MF_TYPE)) continue;
                  int cx, cy, cz;
                  if (!cell_index(g, pp.x, pp.y, pp.z, cx, cy, cz) || cell_lin(g, cx, cy, cz) != nl) continue;
                } else s = A.sorted_slot[cb + (ce - cb - 1 - q)];
                double4 p = ld_rec(&A.posm[s]);
                long long m = meta_of(p);
                if (!(m & MF_TYPE)) continue;                       // removed from its chain (cgroup_unsort_atom)
                double rd = dist2_idnint(g, p.x, p.y, p.z, rx, ry, rz);
                if (rd < g.rc_list2) {
                  if (pass == 1) { if (base + off + cnt < A.cols_cap) { A.cols[base + off + cnt] = s; A.bq[base + off + cnt] = 0; } else ovf = true; }
                  ++cnt;
                }
                if (pass == 1 && (m & MF_REF) && !(rd > g.rc_list2)) {
                  RowHead *h = &A.rh[s];
                  const int len = h->len;
                  if (len < h->cap) {
                    A.cols[h->start + len] = ns; h->len = len + 1;
                    A.bq[h->start + len] = 0; h->q5 = 0;   // appended entries carry no build distance: never skipped, and the near list of the head is no longer complete
                  }
                  else { atomicAdd((unsigned long long *)&sc->row_overflow, 1ull); atomicCAS(&sc->err, 0, DML_E_ROW_OVERFLOW); }
                }
              }
            }
            if (pass == 0) {
              int incl = cnt;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
              off = incl - cnt;
              total = __shfl_sync(0xffffffffu, incl, 31);
            }
          }
          if (__any_sync(0xffffffffu, ovf) && lane == 0) atomicCAS(&sc->err, 0, DML_E_COLS_OVERFLOW);
          __syncwarp();
          if (lane == 0) {
            rh_store_plain(&A.rh[ns], base, total, total + A.row_slack, 0);   // zero build distances, no near list: nothing is ever skipped
            sc->cols_used = base + total + A.row_slack;
          }
        } else if (tid == 0) rh_store_plain(&A.rh[ns], ns * ROW_W, 0, 0, 255);
      }
      __syncthreads();
      if (s_i[5] >= 0) { glen += 1; npend += 1; if (tid != 0) n = n + 1; }
      __syncthreads();
    } else {
      // ---- destruction attempt (dana.F90:680-708) ----
      if (tid == 0) {
        int acc = !((double)n / (v * A.act) < rng.unif());
        s_i[2] = acc;
        if (acc) { int m = (int)floor(rng.unif() * n) + 1; if (m > n) m = n; s_i[6] = m; }
      }
      __syncthreads();
      if (!s_i[2]) { __syncthreads(); continue; }
      int m = s_i[6];
      if (tid == 0) s_i[9] = -1;
      if (m <= 0) { if (tid == 0) atomicCAS(&sc->err, 0, DML_E_GCMC_CHOSEN); __syncthreads(); continue; }
      // chunk holding the m-th in-volume member
      int nchunk = (glen + GB - 1) / GB, before = 0, chunk = -1;
      for (int base = 0; base < nchunk && chunk < 0; base += GB) {
        int c = base + tid;
        int cnt = c < nchunk ? A.gcc[c] : 0;
        int tot; int ex = block_excl_scan(cnt, &tot, sh_scan);
        if (tid == 0) s_i[7] = -1;
        __syncthreads();
        if (c < nchunk && before + ex < m && m <= before + ex + cnt) { s_i[7] = c; s_i[8] = before + ex; }
        __syncthreads();
        if (s_i[7] >= 0) { chunk = s_i[7]; before = s_i[8]; } else before += tot;
        __syncthreads();
      }
      if (chunk < 0) { if (tid == 0) atomicCAS(&sc->err, 0, DML_E_GCMC_CHOSEN); __syncthreads(); continue; }
      {
        int i = chunk * GB + tid, in = 0, s = -1;
        if (i < glen) { s = A.gorder[i]; if (s >= 0) in = in_volume(ld_rec(&A.posm[s]), z0, zmax) ? 1 : 0; }
        int tot; int ex = block_excl_scan(in, &tot, sh_scan);
        if (in && before + ex + 1 == m) s_i[9] = s;
        __syncthreads();
        (void)tot;
      }
      if (tid == 0 && s_i[9] < 0) atomicCAS(&sc->err, 0, DML_E_GCMC_CHOSEN);
      if (tid == 0 && s_i[9] >= 0) {
        int s = s_i[9];
        // atom%dest(): detach from gcmc, hs (row dropped, slot to limbo), ref, b (chain), sys — Groups.F90:433-467
        n = n - 1;
        A.gcc[A.gpos[s] / GB] -= 1;
        A.gorder[A.gpos[s]] = -1; sc->gtomb++;
        rh_store_plain(&A.rh[s], s * ROW_W, 0, 0, 255);
        A.b_occ[A.slot_b[s]] = 0;
        if (A.slot_b[s] < sc->bhole_lo) sc->bhole_lo = A.slot_b[s];
        if (!sc->listed && s < sc->hole_lo) sc->hole_lo = s;
        double4 p = ld_rec(&A.posm[s]);
        p.w = meta_as_double(sc->listed ? MF_LIMBO : 0);
        st_rec(&A.posm[s], p);
        if (sc->listed) sc->nlimbo++;
        sc->nat_sys--; sc->nat_ref--; sc->nat_gcmc--; sc->gcmc_destroyed++;
      }
      __syncthreads();
      if (tid != 0 && s_i[9] >= 0) n = n - 1;
      __syncthreads();
    }
  }
  if (tid == 0) { sc->glen = glen; if (npend > 0) { sc->rows_asym = 2; sc->rev_valid = 0; } }   // appended entries use <= (Neighbor.F90:307)
}

// k_promote variant that also drops promoted atoms from the gcmc membership array (dana.F90:235)
__global__ void k_gcmc_tomb(const double4 *__restrict__ posm, int *__restrict__ gorder, const int *__restrict__ gpos,
                            DevScal *__restrict__ sc, int n) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  long long m = meta_of(ld_rec_nc(&posm[s]));
  if ((m & MF_REF) && (m & MF_TYPE) == 3 && (m & MF_GCMC)) { gorder[gpos[s]] = -1; atomicAdd(&sc->gtomb, 1); }
}

} // namespace dml

static int gcmc_run_impl(dml_ctx *ctx) {
  ctx->kb_valid = false;
  using namespace dml;
  if (ctx->cfg.reservoir != 3) return 0;
  if (!ctx->binned || !ctx->tessellated) FAIL("gcmc_run: call test_update first");
  int nadj = ctx->cfg.nadj;
  if (ctx->n + nadj > ctx->cap) { TRY(finish(ctx)); if (ctx->n + nadj > ctx->cap) FAIL("slot capacity exhausted (dml_config.capacity)"); }
  CKC(ctx->gpend.ensure((size_t)nadj + 8, ctx->st));
  if (ctx->ph.rng_mode == DML_RNG_REPLAY && ctx->rp_nu == 0 && nadj > 0) FAIL("replay mode: call dml_set_replay_gcmc before gcmc_run");
  GcmcArgs A;
  A.posm = ctx->posm.p; A.vel = ctx->vel.p; A.acel = ctx->acel.p; A.fe = ctx->fe.p; A.fnz = ctx->fnz.p;
  A.pos_old = ctx->pos_old.p; A.old_cg = ctx->old_cg.p; A.uid = ctx->uid.p; A.slot_b = ctx->slot_b.p; A.b_occ = ctx->b_occ.p;
  A.cell_of_unused = nullptr; A.cell_start = ctx->cell_start.p; A.sorted_slot = ctx->sorted_slot.p; A.sorted_posm = ctx->sorted_posm.p;
  A.rh = ctx->rh.p; A.cols = ctx->cols.p; A.bq = ctx->bq.p; A.cols_cap = (int)ctx->cols.cap;
  A.gorder = ctx->gorder.p; A.gpos = ctx->gpos.p; A.gcc = ctx->gcc.p; A.gorder_cap = ctx->gorder_cap;
  A.pend = ctx->gpend.p; A.rp_u = ctx->rp_gu.p; A.rp_g = ctx->rp_gg.p; A.rp_nu = ctx->rp_nu; A.rp_ng = ctx->rp_ng;
  A.sc = ctx->sc; A.g = ctx->geo; A.ph = ctx->ph; A.act = ctx->cfg.act; A.beta_kT = ctx->cfg.kB_ui_gcmc * ctx->cfg.Tsist;
  A.nadj = nadj; A.cap = ctx->cap; A.listed = 1; A.row_slack = ctx->row_slack; A.step = STEP_FROM_DEVICE;
  TRY(enq_materialize_rows(ctx));                       // the incremental list upkeep of an accepted attempt edits the rows in place
  LAUNCH(K_GCMC, k_gcmc_census, 148 * 4, 256, ctx->posm.p, ctx->gorder.p, ctx->gcc.p, ctx->sc);
  LAUNCH(K_GCMC, k_gcmc, 1, GB, A);
  ctx->n = std::min(ctx->cap, ctx->n + nadj);          // upper bound of hs%amax until the next read-back (empty slots are skipped)
  ctx->rp_nu = ctx->rp_ng = 0;
  return 0;
}
