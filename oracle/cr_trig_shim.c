/* oracle/cr_trig_shim.c -- TEST INFRASTRUCTURE.
 * Correctly-rounded sin/cos for the reference build: binary128 evaluation
 * (libquadmath) rounded once to double.  tests/orig of the reference were
 * produced with a correctly rounded libm; linking this
 * shim in front of libm makes the unmodified reference reproduce them byte for
 * byte with glibc 2.39 too.  The GPU engine's own sin/cos (double-double,
 * csrc/friq_trig.cuh) is validated against the same functions. */
#include <quadmath.h>
double sin(double x) { return (double)sinq((__float128)x); }
double cos(double x) { return (double)cosq((__float128)x); }
/* gcc -O2 merges sin(x) and cos(x) of one argument into a sincos() call
 * (cartpole.c:64-66), so that entry point must be covered as well. */
void sincos(double x, double *s, double *c) { *s = (double)sinq((__float128)x); *c = (double)cosq((__float128)x); }
