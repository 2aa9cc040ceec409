/* include/dml.h — C ABI of libdml.so, the B200 (sm_100a) hot path of din-mol-Li (`dana`).
 *
 * The reference has no FFI: its hot path is reached by plain Fortran calls from the main program
 * (src/dana.F90:173-265).  "Drop-in" therefore means one exported entry point per preserved call
 * site; the Fortran side binds them with ISO_C_BINDING (fortran/dml_cuda.F90).
 * Every entry point below cites the reference procedure whose body it replaces (paths relative to
 * the reference tree).
 *
 * Conventions
 *   - return 0 = ok, <0 = error; dml_last_error(ctx) gives the message.  Never exits the process.
 *     The Fortran shim maps rc/=0 onto `call werr(msg,.true.)` (src/Errors.f90:59-87).
 *   - plain pointers and sizes only; caller owns host arrays, ctx owns device memory and one stream.
 *   - not re-entrant per ctx (one host thread per ctx); any number of ctxs per process; no globals.  Every entry point
 *     selects the ctx's own device (cudaSetDevice) first, so ctxs on different GPUs can be driven from any host thread.
 *   - particle arrays are indexed by "slot" = index in the reference's hs%a(:) minus 1
 *     (src/Neighbor.F90:32, src/Groups.F90:179-219).  z[slot]==0 marks an empty slot.
 *   - there is NO CPU fallback: every call needs a CUDA device.
 */
#ifndef DML_H
#define DML_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dml_ctx dml_ctx;

/* flags[] bits in dml_upload / dml_download */
#define DML_F_REF   1   /* member of hs%ref  (src/Neighbor.F90:35)            */
#define DML_F_GCMC  2   /* member of the gcmc group (src/dana.F90:130-137)    */
#define DML_F_SKIP  4   /* atom%skip (src/Groups.F90:296)                     */
#define DML_F_LIMBO 8   /* slot parked on hs%limbo (src/Neighbor.F90:262-267) */

#define DML_RNG_PHILOX 0  /* counter-based Philox4x32-10, keyed (seed; uid, step, stream) */
#define DML_RNG_REPLAY 1  /* consume the numbers injected with dml_set_replay_*            */
#define DML_RNG_REFERENCE 2 /* the reference's own ran / gasdev stream (src/dana.F90:1379-1428) drawn on the device in the reference's
                             * order: one thread walks hs%ref; makes the reference's test cases reproducible digit for digit (needs prob = 1) */

/* Filled by the host from the variables dana reads in entrada()/config_run() plus its parameters
 * (src/dana.F90:12-25,87-100,309-327,399-427). */
typedef struct dml_config {
  int32_t device;        /* CUDA device ordinal */
  int32_t capacity;      /* maximum number of slots */
  double  box[3];        /* gems_program_types::box (src/Program_Types.F90:43-54) */
  int32_t pbc[3];        /* atom%pbc: 1,1,0 in dana (src/dana.F90:483-484) */
  double  rcut;          /* hs%rcut = 3.2 (src/dana.F90:113) */
  double  nb_dcut;       /* Verlet skin (src/Neighbor.F90:109) */
  double  eps[9], r0[9]; /* eps(k,m), r0(k,m) stored [(k-1)*3+(m-1)] (src/dana.F90:87-100) */
  double  mass[3];       /* element masses for z=1,2,3 (src/dana.F90:82-84) */
  double  h, gama, Tsist, kB_ui;   /* src/dana.F90:12-25 */
  double  kB_ui_gcmc;    /* gems_constants::kB_ui, used only in gcmc_run (src/dana.F90:594) */
  double  dif_sc, dif_sei, z_sei;  /* z_sei = 80 (src/dana.F90:817) */
  double  prob, z0, z1, zmax, tau; /* tau = 0.1 */
  double  act; int32_t nadj;       /* gcmc (src/dana.F90:421) */
  int32_t integrador;    /* 1 Ermak / 0 Brownian (atom_pbc needs it, src/dana.F90:1212) */
  int32_t reservoir;     /* 1 piston, 2 chunks, 3 gcmc */
  int32_t rng_mode;      /* DML_RNG_* */
  uint64_t seed;
  int32_t strict_order;  /* 1: force sums in the reference's visiting order (bit-exact), 0: row order */
} dml_config;

typedef struct dml_counters {
  int64_t nupd_vlist;    /* src/Neighbor.F90:110 */
  int64_t try_, depo;    /* src/dana.F90:50 (since the last dml_reset_try_depo) */
  int64_t choques, choques2, choques3;  /* src/dana.F90:33 */
  int64_t list_entries;  /* sum of row lengths */
  int64_t overlap_passes, gcmc_created, gcmc_destroyed, row_overflow;
  double  max_vel, msd_t, msd_max;
  int32_t n_slots;       /* hs%amax */
  int32_t nat_sys, nat_ref, nat_gcmc;
  int32_t ncells[3]; double cell[3]; int32_t tessellated, listed;
  int32_t rows_asym;     /* rows may be asymmetric (halo cells, gcmc appends): the pair force uses the transposed rows */
} dml_counters;

typedef struct dml_scalars { double box[3]; double z0, z1, zmax, rho, rho0, t; int64_t step; } dml_scalars;

int  dml_create(dml_ctx **out, const dml_config *cfg);           /* replaces hs%init/setrc set-up: src/dana.F90:112-113 */
void dml_destroy(dml_ctx *ctx);
const char *dml_last_error(dml_ctx *ctx);
const char *dml_version(void);

/* State transfer.  n = number of slots (hs%amax).  Arrays are [n][3] or [n]; NULL = leave/skip.
 * uid = creation rank (order in sys%alist, src/dana.F90:1153-1157); slot_b = index in hs%b%a(:) minus 1. */
int dml_upload(dml_ctx *ctx, int32_t n, const double *pos, const double *vel, const double *acel,
               const double *pos_old, const double *old_cg, const int32_t *z, const int32_t *flags,
               const int32_t *uid, const int32_t *slot_b);
int dml_download(dml_ctx *ctx, int32_t n, double *pos, double *vel, double *acel, double *force, double *epot,
                 double *pos_old, double *old_cg, int32_t *z, int32_t *flags, int32_t *uid, int32_t *slot_b);
int dml_set_scalars(dml_ctx *ctx, const dml_scalars *s);         /* z0,z1,zmax,rho,rho0,box after host-side changes */
int dml_get_scalars(dml_ctx *ctx, dml_scalars *s);
int dml_get_counters(dml_ctx *ctx, dml_counters *c);
/* DML_RNG_REFERENCE: state of ran / gasdev (include/dml_host.h), e.g. as pos_inic left it (src/dana.F90:330-396) */
struct dmlh_rng;
int dml_set_rng_state(dml_ctx *ctx, const struct dmlh_rng *r);
int dml_get_rng_state(dml_ctx *ctx, struct dmlh_rng *r);
int dml_reset_try_depo(dml_ctx *ctx);                             /* salida(): try=0; depo=0 (src/dana.F90:1171-1174) */

/* One entry per preserved call site of the loop body (src/dana.F90:173-265). */
int dml_test_update(dml_ctx *ctx);        /* gems_neighbor::test_update  src/Neighbor.F90:668-713 (+do_pbc Groups.F90:1440-1467,
                                             cgroup_tessellate/sort Cells.F90:180-302, update/ngroup_cells Neighbor.F90:465-633) */
int dml_fuerza(dml_ctx *ctx);             /* fuerza          src/dana.F90:1055-1139 */
int dml_ermak_a(dml_ctx *ctx);            /* ermak_a+atom_pbc src/dana.F90:974-1028,1187-1250 */
int dml_ermak_b(dml_ctx *ctx);            /* ermak_b         src/dana.F90:1031-1052 */
int dml_cbrownian_hs(dml_ctx *ctx);       /* cbrownian_hs+atom_pbc src/dana.F90:798-846 */
int dml_overlap_moveback(dml_ctx *ctx);   /* overlap_moveback src/dana.F90:849-943 */
int dml_msd_book(dml_ctx *ctx);           /* msd_t/msd_max bookkeeping src/dana.F90:201-202 */
int dml_promote(dml_ctx *ctx);            /* F -> CG promotion loop src/dana.F90:228-236 */
int dml_gcmc_run(dml_ctx *ctx);           /* gcmc_run        src/dana.F90:590-713 */
int dml_calc_rho(dml_ctx *ctx, double *rho);   /* calc_rho   src/dana.F90:521-549 */
int dml_maxz(dml_ctx *ctx, double *zmax);      /* maxz       src/dana.F90:776-794 */
/* bloques (src/dana.F90:716-773): the density test runs on the host side of this call; when it fires the
 * nchunk template atoms (pos, pos_old as currently held by the chunk group) are appended and the list rebuilt.
 * fired (out): 1 if the block was added.  The caller shifts its template afterwards like dana.F90:762-763. */
int dml_bloques(dml_ctx *ctx, int32_t nchunk, const double *chunk_pos, const double *chunk_pos_old,
                double dist, double rhomedia, int32_t *fired);
/* Stores the chunk template inside the ctx so that dml_step can run reservoir 2 without host round trips
 * (same semantics as calling dml_bloques every step and shifting the template by dist when it fires). */
int dml_set_chunk_template(dml_ctx *ctx, int32_t nchunk, const double *chunk_pos, const double *chunk_pos_old,
                           double dist, double rhomedia);
/* nsteps full loop iterations (src/dana.F90:173-265 minus salida/timer) without returning to the caller
 * between call sites (reservoir 2 uses the template given to dml_set_chunk_template). */
int dml_step(dml_ctx *ctx, int32_t nsteps);
/* The same for an ensemble of independent replicas on one GPU (BASELINE config 5): one host thread enqueues the
 * step of every replica on its own stream so their kernels overlap; no data-path communication.  dml_set_ensemble_member sizes
 * the cooperative kernels of a ctx so that the replicas can overlap (call it once per member before the loop). */
int dml_ensemble_step(dml_ctx **ctxs, int32_t nctx, int32_t nsteps);
int dml_set_ensemble_member(dml_ctx *ctx, int32_t on);
/* Frame path of a host that only follows the coordinates (what salida() reads, src/dana.F90:1151-1157): new positions of the same
 * atoms in (membership, velocities, lists stay resident; n must equal the current slot count; the copy is stream-ordered; the
 * library no longer knows how far anything moved, so the pair force and overlap_moveback look at every list entry until the next
 * test_update has measured the displacements, and that test_update rebuilds the list if they exceed the skin),
 * positions + element out (synchronous). */
int dml_upload_positions(dml_ctx *ctx, int32_t n, const double *pos, const double *pos_old);
int dml_download_frame(dml_ctx *ctx, int32_t n, double *pos, int32_t *z);

/* Output path on the device: what salida() needs without downloading the frame.
 * dml_salida_sums: energia = sum of epot over sys (src/dana.F90:1155-1163), temp = kion(sys) (src/dana.F90:1342-1376:
 * sum of m*v.v over the non-CG atoms / (j*3*kB_ui)), n_mobile = j.  energia_ref sums hs%ref only (CG atoms keep whatever
 * epot they had when they were promoted: the reference never zeroes them).  Sums are re-associated
 * (fixed tree): equal from run to run, within 1e-12 relative of the reference's serial loop.
 * dml_density_profile: counts[b] = particles of the selected elements (type_mask bit z: 1 Li, 2 CG, 3 F) with
 * int((z-zlo)/((zhi-zlo)/nbins)) == b.  dml_gr: counts[b] = unordered pairs of selected particles with
 * int(|vdistance|/(rmax/nbins)) == b and |vdistance|^2 < rmax^2 (vdistance: src/Groups.F90:995-1016).  Integer results. */
int dml_salida_sums(dml_ctx *ctx, double *energia, double *energia_ref, double *temp, int32_t *n_mobile);
int dml_density_profile(dml_ctx *ctx, double zlo, double zhi, int32_t nbins, int32_t type_mask, int64_t *counts /*[nbins]*/);
int dml_gr(dml_ctx *ctx, double rmax, int32_t nbins, int32_t type_mask, int64_t *counts /*[nbins]*/, int32_t *n_selected);

/* Host object model sync: the reference keeps membership in pointer lists (src/Groups.F90 atom/group/igroup);
 * the device changes it in gcmc_run (insert src/dana.F90:655-674, delete 703-706), bloques (716-773), atom_pbc /
 * overlap_moveback (Li -> F, 1236-1240, 905-910) and the promotion loop (F -> CG, 228-236).  This call reports every slot whose
 * occupant or membership differs from the previous call (the first one compares with dml_upload), in ascending slot order, so
 * the Fortran side can attach / detach / setz exactly those atoms instead of re-reading the whole system.
 * kind bits: 1 a new atom occupies the slot (uid_now = its creation rank, z_now its element), 2 the previous occupant is gone,
 * 4 element changed (z_now), 8 left hs%ref, 16 left the gcmc group.  *n_changes = records found; at most max_changes are returned
 * (the snapshot advances only for slots reported in full: if *n_changes > max_changes, call again for the rest). */
int dml_membership_changes(dml_ctx *ctx, int32_t max_changes, int32_t *slot, int32_t *kind, int32_t *uid_now, int32_t *z_now, int32_t *n_changes);

/* Parity / inspection */
int dml_get_cells(dml_ctx *ctx, int32_t n, int32_t *cell_xyz /*[n][3], halo-inclusive 0..nc+1*/, int32_t *chain_pos /*[n]*/);
int dml_get_neighbors(dml_ctx *ctx, int32_t n, int32_t width, int32_t *nn /*[n]*/, int32_t *rows /*[n][width], slot ids*/);
int dml_set_neighbors(dml_ctx *ctx, int32_t n, int32_t width, const int32_t *nn, const int32_t *rows);
/* Injected random numbers (trace-replay parity mode).  gauss: [n][6] (Ermak: r1,r2 per axis;
 * Brownian uses the first 3), unif_pbc: [n] uniform drawn by atom_pbc, unif_ovl: [n] uniform drawn by the
 * first CG contact of the slot in overlap_moveback.  Valid for the next integrator/overlap call. */
int dml_set_replay_integrator(dml_ctx *ctx, int32_t n, const double *gauss, const double *unif_pbc, const double *unif_ovl);
/* Deposition uniforms of the next dml_overlap_moveback as per-slot queues: the k-th draw of slot s inside the call (the reference
 * draws a fresh ran(idum) at every retry of a failed deposition, src/dana.F90:898-911) reads vals[qstart[s] + k]; qstart has n+1
 * entries.  With prob < 1 an exhausted queue is an error; with prob >= 1 the value never matters. */
int dml_set_replay_overlap(dml_ctx *ctx, int32_t n, const int32_t *qstart, int32_t nvals, const double *vals);
/* uniforms / gaussians consumed by the next dml_gcmc_run in order */
int dml_set_replay_gcmc(dml_ctx *ctx, int32_t nu, const double *unif, int32_t ng, const double *gauss);

/* Multi-GPU, one large box: z-slab decomposition with NCCL halo exchange (no counterpart in the reference, which is a
 * serial program).  One ctx per rank, every rank makes the same calls in the same order.
 * dml_slab_step = dml_step for the decomposed box (Ermak integrator + piston reservoir): global rebuild decision, particle
 * migration and ghost re-selection at a rebuild, global rho for the piston, cross-face overlap resolution.  Counters
 * (try, depo, choques) are per rank; nupd_vlist, rho and zmax are the same on every rank. */
int dml_comm_unique_id(void *id128);                                   /* ncclGetUniqueId (rank 0), 128 bytes */
int dml_comm_init(dml_ctx *ctx, const void *id128, int32_t rank, int32_t nranks);
int dml_slab_plan(int32_t n, const double *z, int32_t nranks, double lo, double hi, double *cuts /*[nranks+1]*/);  /* host only */
int dml_slab_setup(dml_ctx *ctx, double zlo, double zhi);              /* after dml_upload of the owned particles */
int dml_slab_halo_exchange(dml_ctx *ctx);                              /* refresh ghost positions (grouped ncclSend/ncclRecv) */
int dml_slab_step(dml_ctx *ctx, int32_t nsteps);                       /* dana.F90:173-265 on the decomposed box */
int dml_slab_info(dml_ctx *ctx, int32_t *n_owned, int32_t *n_ghost, int32_t *nsend_lo, int32_t *nsend_hi);

/* Timing helper for bench.py: device-side duration (ms) of the kernels of the named class accumulated since
 * the last call with reset!=0.  cls: 0 pair force, 1 list build, 2 integrator, 3 overlap, 4 all. */
int dml_profile(dml_ctx *ctx, int32_t enable);   /* 0 off, 1 every kernel, 2+kid only kernel kid (dml_profile_kernel order) */
int dml_profile_get(dml_ctx *ctx, int32_t cls, double *ms, int64_t *launches, int32_t reset);
/* per-kernel timing: kid = 0,1,2,... until the call returns 1; name points to a static string */
int dml_profile_kernel(dml_ctx *ctx, int32_t kid, const char **name, double *ms, int64_t *launches);
int32_t dml_n_slots(dml_ctx *ctx);
int dml_set_strict_order(dml_ctx *ctx, int32_t on);   /* switch dml_fuerza between reference summation order and production kernel */         /* hs%amax as known to the host side (no synchronisation) */
int64_t dml_launch_count(dml_ctx *ctx);   /* kernels launched by this ctx so far */
void *dml_stream(dml_ctx *ctx);           /* cudaStream_t used by every kernel of this ctx */

#ifdef __cplusplus
}
#endif
#endif
