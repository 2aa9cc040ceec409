/* oracle/dana_oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference's hot
 * path (pauvals/din-mol-Li: src/dana.F90, src/Neighbor.F90, src/Cells.F90 and
 * the parts of src/Groups.F90 / src/Program_Types.F90 they touch).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (libdml.so) never links or calls it.
 *
 * Parity status: PINNED — the oracle reproduces tests/{ermak,brown,gcmc}/ref.xyz
 * of the reference bit-for-bit (see tests/test_oracle_golden.py).
 */
#ifndef DANA_ORACLE_H
#define DANA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors entrada.ini (dana.F90:309-327) + movedor.ini (dana.F90:399-427). */
typedef struct orc_params {
  int32_t idum;        /* seed */
  double  prob;        /* deposition probability */
  double  h;           /* time step */
  int32_t nst, nwr;    /* steps, output cadence */
  double  xi, yi;      /* box x,y */
  double  dist;        /* chunk height increment */
  double  z0, zmax;    /* reservoir limits */
  double  dif_sc, dif_sei;
  double  nb_dcut;     /* Verlet skin */
  int32_t integrador;  /* 1 = Ermak, 0 = Brownian */
  int32_t reservoir;   /* 1 = piston, 2 = chunks, 3 = gcmc */
  double  act;         /* gcmc activity */
  int32_t nadj;        /* gcmc attempts per step */
  int32_t nchunk;      /* chunk template size (reservoir 2) */
  const double *chunk_xyz; /* nchunk*3, as read from chunk.xyz */
  int32_t mnb;         /* neighbour table width (reference: 10000) */
  int32_t fast_init;   /* 1: cell-accelerated pos_inic (same RNG stream, same result) */
  int32_t n_init;      /* >0: skip pos_inic, take n_init explicit atoms below   */
  const double  *init_xyz; /* n_init*3 */
  const int32_t *init_z;   /* n_init element ids (1 Li, 2 CG, 3 F) */
  int32_t ov_guard_pass;   /* 0 = faithful (default); >0: see overlap_moveback in dana_oracle.cpp */
} orc_params;

enum orc_op {
  ORC_ERMAK_A = 1, ORC_FUERZA = 2, ORC_ERMAK_B = 3, ORC_CBROWNIAN = 4,
  ORC_TEST_UPDATE = 5, ORC_OVERLAP = 6, ORC_PROMOTE = 7, ORC_GCMC = 8,
  ORC_CALC_RHO = 9, ORC_BLOQUES = 10, ORC_SALIDA = 11, ORC_MAXZ = 12,
  ORC_MSD = 13, ORC_STEP_END = 14
};

/* trace kinds (orc_trace_*) */
enum orc_trace_kind {
  ORC_TR_GAUSS_INTEG = 0,   /* gasdev() drawn by ermak_a / cbrownian_hs, uid = atom */
  ORC_TR_UNIF_PBC = 1,      /* ran() drawn by atom_pbc deposition attempt          */
  ORC_TR_UNIF_OVERLAP = 2,  /* ran() drawn by overlap_moveback CG contact          */
  ORC_TR_UNIF_GCMC = 3,     /* ran() drawn by gcmc_run                              */
  ORC_TR_GAUSS_GCMC = 4     /* gasdev() drawn by gcmc_run                           */
};

typedef struct orc_scalars {
  double  box[3];
  double  z0, z1, zmax, rho, rho0, t, h;
  double  cell[3];
  int32_t ncells[3];
  int32_t tessellated, listed;
  int32_t nat_sys, nat_ref, nat_b, nat_hs, nat_gcmc;
  int32_t hs_amax, b_amax;
  int64_t nupd, choques, choques2, choques3, try_, depo;
  double  max_vel, msd_t, msd_max;
  uint64_t ran_calls;
  int32_t step;
  /* Ermak constants (dana.F90:947-971) */
  double cc0, cc1, cc2, sdr, sdv, crv1, crv2, skt;
} orc_scalars;

void *orc_create(const orc_params *p);        /* everything dana does before its time loop */
void  orc_destroy(void *h);
const char *orc_last_error(void *h);
int   orc_step(void *h, int nsteps);          /* full loop iterations (dana.F90:173-265) */
int   orc_call(void *h, int op);              /* one call site of the loop body */
void  orc_get_scalars(void *h, orc_scalars *s);

/* Per-atom state in sys%alist order (creation order).  Any pointer may be NULL. */
int   orc_get_state(void *h, int64_t *uid, int32_t *z, double *pos, double *vel,
                    double *acel, double *force, double *epot, double *pos_old,
                    double *old_cg, int32_t *flags /*1 ref,2 gcmc,4 skip*/,
                    int32_t *slot_hs, int32_t *slot_b);
/* Neighbour rows: nn[i], rows[i*width .. ] for hs slot i+1 (i < hs_amax); entries are hs slots (1-based).
 * slot_uid[i] = uid of the atom in hs slot i+1, -1 for null, -2 for limbo. */
int   orc_get_rows(void *h, int32_t width, int32_t *nn, int32_t *rows, int64_t *slot_uid);
/* Cell chains: for every b slot i (< b_amax): cell index triple (with halo, 0..n+1) and chain order rank */
int   orc_get_cells(void *h, int32_t *cell_of_slot /*3 per slot*/, int32_t *chain_pos);

/* Last frame written by salida() (dana.F90:1143-1183), sys order */
int   orc_get_frame(void *h, int32_t *nat, double *zmax, int32_t *z, double *pos, double *scal /*t,E,T,rho,try,depo*/);

/* RNG trace of the calls made since the last orc_trace_clear() */
void  orc_trace_enable(void *h, int on);
void  orc_trace_clear(void *h);
int64_t orc_trace_size(void *h);
void  orc_trace_get(void *h, int32_t *kind, int64_t *uid, double *val);

int   orc_threads(void);
void  orc_set_threads(int n);                 /* overrides OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1) */                      /* threads the list build uses (OpenMP, like the reference) */

/* Stand-alone pieces (known-answer tests / generators) */
void  orc_rng_kat(int32_t idum, int n_ran, double *ran_out, int n_gas, double *gas_out);
/* pos_inic rule (dana.F90:330-396): random sequential insertion, returns n; xyz quantised through %.12f.
 * idum is updated like the Fortran argument; rng state is private to the call. */
int   orc_pos_inic(int32_t idum, double xi, double yi, double alto, int fast, double *xyz, int cap, uint64_t *ran_calls);

#ifdef __cplusplus
}
#endif
#endif
