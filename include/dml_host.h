/* include/dml_host.h — host-side pieces of `dana` that sit next to the hot path (C ABI, plain C++ inside libdml.so).
 * They restate the parts of the reference's main program a caller needs to set a run up without the Fortran
 * binary: the RNG (src/dana.F90:1379-1428) and the initial-configuration rule pos_inic (src/dana.F90:330-396).
 * No device work happens here. */
#ifndef DML_HOST_H
#define DML_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dmlh_rng { int32_t idum, ix, iy, stored; double g; uint64_t calls; } dmlh_rng;

void   dmlh_rng_init(dmlh_rng *r, int32_t idum);   /* state of `ran` before its first call (ix=iy=-1) */
double dmlh_ran(dmlh_rng *r);                      /* ran(idum)   src/dana.F90:1407-1428 */
double dmlh_gasdev(dmlh_rng *r);                   /* gasdev()    src/dana.F90:1379-1404 */

/* pos_inic (src/dana.F90:330-396): n = int(xi*yi*alto*6.022e-4) particles by random sequential insertion with
 * minimum separation 3.2 (x,y minimum image), drawn from r; coordinates are passed through the reference's
 * f25.12 text round trip.  The scan over earlier particles is cell-accelerated: same accept/reject decisions,
 * same RNG consumption, same result as the reference's O(N^2) loop.  Returns n, or -needed if cap is too small,
 * or -1 when 10000 attempts fail for one particle (the reference stops there). */
int32_t dmlh_pos_inic(dmlh_rng *r, double xi, double yi, double alto, double *xyz, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif
