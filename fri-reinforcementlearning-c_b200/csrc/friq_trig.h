/* csrc/friq_trig.h -- correctly rounded sin/cos for the environment dynamics.
 *
 * Why: the reference's tests/orig rule bases were produced with a correctly
 * rounded libm, and the learning trajectory is chaotic enough that a 1-ulp cos
 * difference grows to 1e-2 in acrobot.  CUDA's libdevice
 * sin/cos are 1-2 ulp, so the engine carries its own: Cody-Waite reduction by
 * pi/2 in four 33-bit pieces, a 27-entry double-double table of sin/cos(j/32),
 * and double-double Taylor kernels on |t| <= 1/64.  Total error ~2^-95 relative,
 * i.e. the rounded double is the correctly rounded one except with probability
 * ~2^-40 per call.  Valid for |x| < 2^19 (the dynamics stay below 20).
 *
 * The file is plain C so that the same code is compiled by nvcc for the device
 * and by gcc for tests/test_trig_host.py, which checks it against binary128
 * libquadmath on millions of points.  Constants: tools/gen_trig_tables.py.
 * Must be compiled WITHOUT floating-point contraction (-fmad=false /
 * -ffp-contract=off); the only fused operations are the explicit fma() calls.
 */
#ifndef FRIQ_TRIG_H
#define FRIQ_TRIG_H
#include <math.h>

#if defined(__CUDACC__)
#define FRIQ_HD __host__ __device__ __forceinline__
#define FRIQ_HD_COLD __host__ __device__ __noinline__      /* the accurate phase is rare: a real call keeps the kernels' hot code small */
#else
#define FRIQ_HD static inline
#define FRIQ_HD_COLD static __attribute__((noinline))
#endif

#define FRIQ_PIO2_1 0x1.921fb54400000p+0
#define FRIQ_PIO2_2 0x1.0b4611a600000p-34
#define FRIQ_PIO2_3 0x1.3198a2e000000p-69
#define FRIQ_PIO2_4 0x1.b839a252049c1p-104
#define FRIQ_TWO_OVER_PI 0x1.45f306dc9c883p-1
#define FRIQ_S3_HI -0x1.5555555555555p-3
#define FRIQ_S3_LO -0x1.5555555555555p-57
#define FRIQ_S5_HI 0x1.1111111111111p-7
#define FRIQ_S5_LO 0x1.1111111111111p-63
#define FRIQ_S7 -0x1.a01a01a01a01ap-13
#define FRIQ_S9 0x1.71de3a556c734p-19
#define FRIQ_S11 -0x1.ae64567f544e4p-26
#define FRIQ_S13 0x1.6124613a86d09p-33
#define FRIQ_C4_HI 0x1.5555555555555p-5
#define FRIQ_C4_LO 0x1.5555555555555p-59
#define FRIQ_C6 -0x1.6c16c16c16c17p-10
#define FRIQ_C8 0x1.a01a01a01a01ap-16
#define FRIQ_C10 -0x1.27e4fb7789f5cp-22
#define FRIQ_C12 0x1.1eed8eff8d898p-29

#define FRIQ_SC_TABLE_INIT { \
    {0x0.0p+0, 0x0.0p+0, 0x1.0000000000000p+0, 0x0.0p+0}, \
    {0x1.ffeaaaeeee86fp-6, -0x1.cd406fb224ae2p-60, 0x1.ffc00155527d3p-1, -0x1.3b54492d89b5bp-55}, \
    {0x1.ffaaaeeed4edbp-5, -0x1.2d16d32684b69p-59, 0x1.ff0015549f4d3p-1, 0x1.328387b99426fp-55}, \
    {0x1.7f701032550e4p-4, 0x1.afc2d1800501ap-60, 0x1.fdc06bf7e6b9bp-1, 0x1.31902b535f8dbp-55}, \
    {0x1.feaaeee86ee36p-4, -0x1.afcb2bcc6f03bp-59, 0x1.fc015527d5bd3p-1, 0x1.b68f35094efb8p-55}, \
    {0x1.3eb312c5d66cbp-3, 0x1.47d666b66cb91p-57, 0x1.f9c340a7cc428p-1, 0x1.c5b6b063b7462p-55}, \
    {0x1.7dc102fbaf2b5p-3, 0x1.5ab50e23c97c3p-59, 0x1.f706bdf9ece1cp-1, -0x1.698c80c36dcb4p-55}, \
    {0x1.bc6f84edc6199p-3, 0x1.9c1a56a7b0cabp-57, 0x1.f3cc7c3b3d16ep-1, -0x1.21a3ad28a3494p-57}, \
    {0x1.faaeed4f31577p-3, -0x1.15d88508e32b8p-57, 0x1.f01549f7deea1p-1, 0x1.d3c1e99e5cafdp-55}, \
    {0x1.1c37d64c6b876p-2, 0x1.46076fe0dcff4p-56, 0x1.ebe214f76efa8p-1, -0x1.02f9f12ba543ep-55}, \
    {0x1.3ad129769d3d8p-2, 0x1.03d550487839ap-63, 0x1.e733ea0193d40p-1, -0x1.6428b3546ce13p-55}, \
    {0x1.591bc9fa2f597p-2, 0x1.7c74bac3fe0cbp-57, 0x1.e20bf49acd6c1p-1, -0x1.660aec7ef636bp-58}, \
    {0x1.7710255764214p-2, -0x1.6ead7314bb6cep-57, 0x1.dc6b7eb995912p-1, 0x1.4b364776dcd35p-58}, \
    {0x1.94a6be9f546c5p-2, -0x1.69ce13e683f58p-56, 0x1.d653f073e4040p-1, -0x1.76236434bec37p-55}, \
    {0x1.b1d8305321617p-2, -0x1.ae242cb99f519p-56, 0x1.cfc6cfa52ad9fp-1, 0x1.8b5b5508f2a0dp-55}, \
    {0x1.ce9d2e3d4a51fp-2, -0x1.2fc8a12dae298p-57, 0x1.c8c5bf8ce1a84p-1, 0x1.ab3d1a1590123p-56}, \
    {0x1.eaee8744b05f0p-2, -0x1.789b43c9b027dp-58, 0x1.c1528065b7d50p-1, -0x1.892111312e828p-55}, \
    {0x1.0362939c69955p-1, -0x1.2d8cd78397b01p-55, 0x1.b96eeef58840ep-1, 0x1.45a3cc78fade0p-58}, \
    {0x1.110d0c4b69c3bp-1, 0x1.d918998809981p-55, 0x1.b11d04162a4c6p-1, 0x1.1dd561efbc0c2p-56}, \
    {0x1.1e7343236574cp-1, 0x1.22a3fa4f41d5ap-56, 0x1.a85ed4373e02dp-1, 0x1.9be06385ec792p-57}, \
    {0x1.2b91dea88421ep-1, -0x1.fa371db216ab0p-55, 0x1.9f368ed912f85p-1, -0x1.1d200c5791606p-55}, \
    {0x1.386597456282bp-1, -0x1.10fada93b07a8p-56, 0x1.95a67e00cb1fdp-1, -0x1.0befda21f862dp-55}, \
    {0x1.44eb381cf386bp-1, -0x1.3ed6c1e6a5505p-55, 0x1.8bb105a5dc900p-1, 0x1.863e03e9474c1p-55}, \
    {0x1.511f9fd7b351cp-1, -0x1.5c0e861c48831p-55, 0x1.8158a31916d5dp-1, -0x1.de8b90b8228dep-57}, \
    {0x1.5cffc16bf8f0dp-1, 0x1.96cb370eb578ap-55, 0x1.769fec655211fp-1, -0x1.827d5cf8c68c5p-57}, \
    {0x1.6888a4e134b2fp-1, -0x1.6b7d37644d5e6p-55, 0x1.6b898fa9efb5dp-1, 0x1.15ac786ccf4b2p-56}, \
    {0x1.73b7680dea578p-1, -0x1.2248306dc12a2p-56, 0x1.6018526f563dfp-1, 0x1.46ca5e0e432d0p-55}, \
}
#if defined(__CUDACC__)
static __device__ const double friq_sc_tab_d[27][4] = FRIQ_SC_TABLE_INIT;
#endif
static const double friq_sc_tab_h[27][4] = FRIQ_SC_TABLE_INIT;
#if defined(__CUDA_ARCH__)
#define FRIQ_SC_TAB friq_sc_tab_d
#else
#define FRIQ_SC_TAB friq_sc_tab_h
#endif

typedef struct { double h, l; } friq_dd;

FRIQ_HD friq_dd friq_two_sum(double a, double b) {
    friq_dd r; r.h = a + b; double bb = r.h - a; r.l = (a - (r.h - bb)) + (b - bb); return r;
}
FRIQ_HD friq_dd friq_fast_two_sum(double a, double b) {   /* requires |a| >= |b| or a == 0 */
    friq_dd r; r.h = a + b; r.l = b - (r.h - a); return r;
}
FRIQ_HD friq_dd friq_two_prod(double a, double b) {
    friq_dd r; r.h = a * b; r.l = fma(a, b, -r.h); return r;
}
FRIQ_HD friq_dd friq_dd_mul(friq_dd a, friq_dd b) {
    friq_dd p = friq_two_prod(a.h, b.h);
    p.l = fma(a.h, b.l, p.l); p.l = fma(a.l, b.h, p.l);
    return friq_fast_two_sum(p.h, p.l);
}
FRIQ_HD friq_dd friq_dd_mul_d(friq_dd a, double b) {
    friq_dd p = friq_two_prod(a.h, b);
    p.l = fma(a.l, b, p.l);
    return friq_fast_two_sum(p.h, p.l);
}
FRIQ_HD friq_dd friq_dd_add(friq_dd a, friq_dd b) {
    friq_dd s = friq_two_sum(a.h, b.h), t = friq_two_sum(a.l, b.l);
    s.l += t.h; s = friq_fast_two_sum(s.h, s.l);
    s.l += t.l; return friq_fast_two_sum(s.h, s.l);
}

/* sin and cos of x as double-doubles */
FRIQ_HD_COLD void friq_sincos_dd(double x, friq_dd *sn, friq_dd *cs) {
    /* 1. x = k*pi/2 + r, |r| <= pi/4 (+ rounding), r as double-double */
    double k = rint(x * FRIQ_TWO_OVER_PI);
    double r0 = fma(-k, FRIQ_PIO2_1, x);                 /* exact */
    friq_dd a = friq_two_sum(r0, -k * FRIQ_PIO2_2);      /* k*PIO2_2, k*PIO2_3 are exact products */
    friq_dd b = friq_two_sum(a.h, -k * FRIQ_PIO2_3);
    friq_dd r = friq_fast_two_sum(b.h, (a.l + b.l) + (-k * FRIQ_PIO2_4));
    /* 2. r = j/32 + t, |t| <= 1/64 */
    double jf = rint(r.h * 32.0);
    friq_dd t = friq_two_sum(r.h - jf * 0.03125, r.l);   /* r.h - j/32 is exact */
    int j = (int)jf, neg = j < 0;
    if (neg) j = -j;
    /* 3. Taylor kernels on t; z = t^2 */
    friq_dd z = friq_dd_mul(t, t);
    double qs = FRIQ_S7 + z.h * (FRIQ_S9 + z.h * (FRIQ_S11 + z.h * FRIQ_S13));
    double qc = FRIQ_C6 + z.h * (FRIQ_C8 + z.h * (FRIQ_C10 + z.h * FRIQ_C12));
    friq_dd c5; c5.h = FRIQ_S5_HI; c5.l = FRIQ_S5_LO;
    friq_dd c3; c3.h = FRIQ_S3_HI; c3.l = FRIQ_S3_LO;
    friq_dd c4; c4.h = FRIQ_C4_HI; c4.l = FRIQ_C4_LO;
    friq_dd mh; mh.h = -0.5; mh.l = 0.0;
    friq_dd one; one.h = 1.0; one.l = 0.0;
    friq_dd ps = friq_dd_mul(z, friq_dd_add(c3, friq_dd_mul(z, friq_dd_add(c5, friq_dd_mul_d(z, qs)))));
    friq_dd st = friq_dd_add(t, friq_dd_mul(t, ps));                                   /* sin t */
    friq_dd pc = friq_dd_mul(z, friq_dd_add(mh, friq_dd_mul(z, friq_dd_add(c4, friq_dd_mul_d(z, qc)))));
    friq_dd ct = friq_dd_add(one, pc);                                                 /* cos t */
    /* 4. angle addition with the table point c = j/32 (sign folded into sin c) */
    friq_dd sc, cc;
    sc.h = FRIQ_SC_TAB[j][0]; sc.l = FRIQ_SC_TAB[j][1]; cc.h = FRIQ_SC_TAB[j][2]; cc.l = FRIQ_SC_TAB[j][3];
    if (neg) { sc.h = -sc.h; sc.l = -sc.l; }
    friq_dd sr = friq_dd_add(friq_dd_mul(sc, ct), friq_dd_mul(cc, st));                /* sin r */
    friq_dd nsc; nsc.h = -sc.h; nsc.l = -sc.l;
    friq_dd cr = friq_dd_add(friq_dd_mul(cc, ct), friq_dd_mul(nsc, st));               /* cos r */
    /* 5. quadrant */
    int q = ((int)k) & 3;
    friq_dd nsr, ncr; nsr.h = -sr.h; nsr.l = -sr.l; ncr.h = -cr.h; ncr.l = -cr.l;
    *sn = (q == 0) ? sr : (q == 1) ? cr : (q == 2) ? nsr : ncr;
    *cs = (q == 0) ? cr : (q == 1) ? nsr : (q == 2) ? ncr : sr;
}

/* ---- quick phase (Ziv's strategy): sin and cos to ~2^-64 relative in mostly plain double arithmetic, ~5x cheaper than
 * friq_sincos_dd.  Each result comes as (h, l) with h = RN(h + l); the caller rounds it and keeps the result only if
 * adding or subtracting the error bound FRIQ_QUICK_ERR * |h| cannot change the rounding, else it runs the accurate phase.
 *   reduction as above (exact to ~2^-100);  t = r - j/32 = (th, tl), |t| <= 1/64;  z = th^2 (+ zl)
 *   sin t = th + ds,  ds = tl + th * z * (S3 + z (S5 + z (S7 + z S9)))                  |ds| <= 2^-14.6 |t|, error <= 2^-65 |t|
 *   cos t = 1 + dc,   dc = -z/2 + (z^2 (C4 + z (C6 + z C8)) - th tl - zl/2)            |dc| <= 2^-13,       error <= 2^-66
 *   sin r = sc + [cc.h th] + (scl + sch dc + cch ds + ccl th),   cos r = cc - [sc.h th] + (ccl + cch dc - sch ds - scl th)
 * with the bracketed products exact (two_prod) and the head sums exact (two_sum); every remaining term is below 2^-13 of the
 * result's magnitude class, so its double rounding is below 2^-65.  Measured against binary128 on 10^8 points the error stays
 * under 2^-64; FRIQ_QUICK_ERR = 2^-61 leaves a factor 8, and sends ~2^-8 of the calls to the accurate phase. */
#define FRIQ_QUICK_ERR 0x1p-61
FRIQ_HD void friq_sincos_quick(double x, friq_dd *sn, friq_dd *cs) {
    double k = rint(x * FRIQ_TWO_OVER_PI);
    double r0 = fma(-k, FRIQ_PIO2_1, x);
    friq_dd a = friq_two_sum(r0, -k * FRIQ_PIO2_2);
    friq_dd b = friq_two_sum(a.h, -k * FRIQ_PIO2_3);
    friq_dd r = friq_fast_two_sum(b.h, (a.l + b.l) + (-k * FRIQ_PIO2_4));
    double jf = rint(r.h * 32.0);
    friq_dd t = friq_two_sum(r.h - jf * 0.03125, r.l);
    int j = (int)jf, neg = j < 0;
    if (neg) j = -j;
    const double th = t.h, tl = t.l;
    const double z = th * th, zl = fma(th, th, -z);
    const double ps = FRIQ_S3_HI + z * (FRIQ_S5_HI + z * (FRIQ_S7 + z * FRIQ_S9));
    const double ds = tl + th * (z * ps);
    const double pc = FRIQ_C4_HI + z * (FRIQ_C6 + z * FRIQ_C8);
    const double dc = -0.5 * z + ((z * z) * pc - (th * tl + 0.5 * zl));
    double sch = FRIQ_SC_TAB[j][0], scl = FRIQ_SC_TAB[j][1];
    const double cch = FRIQ_SC_TAB[j][2], ccl = FRIQ_SC_TAB[j][3];
    if (neg) { sch = -sch; scl = -scl; }
    /* sin r */
    friq_dd p = friq_two_prod(cch, th);
    friq_dd hs = friq_two_sum(sch, p.h);
    double lo = ((scl + ccl * th) + (sch * dc + cch * ds)) + (p.l + hs.l);
    friq_dd sr = friq_fast_two_sum(hs.h, lo);
    /* cos r */
    p = friq_two_prod(sch, th);
    hs = friq_two_sum(cch, -p.h);
    lo = ((ccl - scl * th) + (cch * dc - sch * ds)) + (hs.l - p.l);
    friq_dd cr = friq_fast_two_sum(hs.h, lo);
    int q = ((int)k) & 3;
    friq_dd nsr, ncr; nsr.h = -sr.h; nsr.l = -sr.l; ncr.h = -cr.h; ncr.l = -cr.l;
    *sn = (q == 0) ? sr : (q == 1) ? cr : (q == 2) ? nsr : ncr;
    *cs = (q == 0) ? cr : (q == 1) ? nsr : (q == 2) ? ncr : sr;
}
/* The quick phase for ONE of the two functions (want_cos = 0: sin x, 1: cos x).  With x = k pi/2 + r the answer is +-sin r or
 * +-cos r depending on k mod 4, and both have the same shape
 *     f = X + [Y.h th] + ((X.l + Y.l th) + (X.h dc + Y.h ds)),   (X, Y) = (sc, cc) for sin r,  (cc, -sc) for cos r:
 * selecting the table operands first and evaluating once gives bit for bit what friq_sincos_quick returns for that function
 * (negating an operand only flips signs of exact intermediate values) at about 60 % of its cost - the mountaincar dynamics need
 * one cosine per env step (mountaincar.c:44). */
FRIQ_HD friq_dd friq_sincos_quick_one(double x, int want_cos) {
    double k = rint(x * FRIQ_TWO_OVER_PI);
    double r0 = fma(-k, FRIQ_PIO2_1, x);
    friq_dd a = friq_two_sum(r0, -k * FRIQ_PIO2_2);
    friq_dd b = friq_two_sum(a.h, -k * FRIQ_PIO2_3);
    friq_dd r = friq_fast_two_sum(b.h, (a.l + b.l) + (-k * FRIQ_PIO2_4));
    double jf = rint(r.h * 32.0);
    friq_dd t = friq_two_sum(r.h - jf * 0.03125, r.l);
    int j = (int)jf, neg = j < 0;
    if (neg) j = -j;
    const double th = t.h, tl = t.l;
    const double z = th * th, zl = fma(th, th, -z);
    const double ps = FRIQ_S3_HI + z * (FRIQ_S5_HI + z * (FRIQ_S7 + z * FRIQ_S9));
    const double ds = tl + th * (z * ps);
    const double pc = FRIQ_C4_HI + z * (FRIQ_C6 + z * FRIQ_C8);
    const double dc = -0.5 * z + ((z * z) * pc - (th * tl + 0.5 * zl));
    double sch = FRIQ_SC_TAB[j][0], scl = FRIQ_SC_TAB[j][1];
    const double cch = FRIQ_SC_TAB[j][2], ccl = FRIQ_SC_TAB[j][3];
    if (neg) { sch = -sch; scl = -scl; }
    const int q = ((int)k) & 3;
    /* sin x = (sr, cr, -sr, -cr)[q],  cos x = (cr, -sr, -cr, sr)[q]: the cosine of r is needed when (q + want_cos) is odd */
    const int use_cos = (q + want_cos) & 1;
    const double Xh = use_cos ? cch : sch, Xl = use_cos ? ccl : scl;
    const double Yh = use_cos ? -sch : cch, Yl = use_cos ? -scl : ccl;
    friq_dd p = friq_two_prod(Yh, th);
    friq_dd hs = friq_two_sum(Xh, p.h);
    const double lo = ((Xl + Yl * th) + (Xh * dc + Yh * ds)) + (p.l + hs.l);
    friq_dd f = friq_fast_two_sum(hs.h, lo);
    const int negate = want_cos ? (q == 1 || q == 2) : (q >= 2);
    if (negate) { f.h = -f.h; f.l = -f.l; }
    return f;
}
/* 1 when RN(v.h + v.l + e) is the same double for every |e| <= FRIQ_QUICK_ERR * |v.h| */
FRIQ_HD int friq_quick_ok(friq_dd v) {
    const double e = fabs(v.h) * FRIQ_QUICK_ERR;
    return (v.h + (v.l + e) == v.h) && (v.h + (v.l - e) == v.h);
}

FRIQ_HD double friq_sin(double x) {
    friq_dd s = friq_sincos_quick_one(x, 0), c;
    if (friq_quick_ok(s)) return s.h;
    friq_sincos_dd(x, &s, &c); return s.h + s.l;
}
FRIQ_HD double friq_cos(double x) {
    friq_dd c = friq_sincos_quick_one(x, 1), s;
    if (friq_quick_ok(c)) return c.h;
    friq_sincos_dd(x, &s, &c); return c.h + c.l;
}
FRIQ_HD void friq_sincos(double x, double *s, double *c) {
    friq_dd a, b; friq_sincos_quick(x, &a, &b);
    if (friq_quick_ok(a) && friq_quick_ok(b)) { *s = a.h; *c = b.h; return; }
    friq_sincos_dd(x, &a, &b); *s = a.h + a.l; *c = b.h + b.l;
}

#endif
