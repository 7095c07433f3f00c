/* tests/trig_host_shim.c -- compiles the PRODUCT's csrc/friq_trig.h for the host
 * so the exact device algorithm can be checked against binary128 on the CPU. */
#include "friq_trig.h"
#include <quadmath.h>
double shim_sin(double x) { return friq_sin(x); }
double shim_cos(double x) { return friq_cos(x); }
/* returns the number of points where the result is not the correctly rounded one */
long shim_check(const double *x, long n, double *worst_x) {
    long bad = 0;
    for (long i = 0; i < n; i++) {
        double s, c; friq_sincos(x[i], &s, &c);
        if (s != (double)sinq((__float128)x[i]) || c != (double)cosq((__float128)x[i])) { bad++; *worst_x = x[i]; }
    }
    return bad;
}
/* friq_sin / friq_cos (the single-function quick phase): correctly rounded, and the quick values are bit for bit the ones
 * friq_sincos_quick returns for that function; returns the number of violations */
long shim_check_one(const double *x, long n, double *worst_x) {
    long bad = 0;
    for (long i = 0; i < n; i++) {
        friq_dd s, c; friq_sincos_quick(x[i], &s, &c);
        const friq_dd s1 = friq_sincos_quick_one(x[i], 0), c1 = friq_sincos_quick_one(x[i], 1);
        const int same = s1.h == s.h && s1.l == s.l && c1.h == c.h && c1.l == c.l;
        if (!same || friq_sin(x[i]) != (double)sinq((__float128)x[i]) || friq_cos(x[i]) != (double)cosq((__float128)x[i])) { bad++; *worst_x = x[i]; }
    }
    return bad;
}
/* max relative error of the double-double result vs binary128, in units of 2^-90 */
double shim_dd_err(const double *x, long n) {
    double worst = 0;
    for (long i = 0; i < n; i++) {
        friq_dd s, c; friq_sincos_dd(x[i], &s, &c);
        __float128 es = sinq((__float128)x[i]), ec = cosq((__float128)x[i]);
        __float128 ds = ((__float128)s.h + s.l - es) / es, dc = ((__float128)c.h + c.l - ec) / ec;
        double a = fabs((double)ds) * 0x1p90, b = fabs((double)dc) * 0x1p90;
        if (a > worst) worst = a;
        if (b > worst) worst = b;
    }
    return worst;
}
/* quick phase: max relative error of (h + l) vs binary128 in units of 2^-64, and the share of points the rounding test
 * sends to the accurate phase */
double shim_quick_err(const double *x, long n, double *fallback_share) {
    double worst = 0; long fb = 0;
    for (long i = 0; i < n; i++) {
        friq_dd s, c; friq_sincos_quick(x[i], &s, &c);
        __float128 es = sinq((__float128)x[i]), ec = cosq((__float128)x[i]);
        __float128 ds = ((__float128)s.h + s.l - es) / es, dc = ((__float128)c.h + c.l - ec) / ec;
        double a = fabs((double)ds) * 0x1p64, b = fabs((double)dc) * 0x1p64;
        if (es != 0 && a > worst) worst = a;
        if (ec != 0 && b > worst) worst = b;
        if (!friq_quick_ok(s) || !friq_quick_ok(c)) fb++;
    }
    *fallback_share = (double)fb / (double)n;
    return worst;
}
