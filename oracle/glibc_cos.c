/*
 * TEST INFRASTRUCTURE (oracle): CPU restatement in plain C of glibc's double-precision cos
 * (IBM Accurate Mathematical Library; glibc >= 2.28, sysdeps/ieee754/dbl-64/s_sin.c: __cos,
 * do_cos, do_sin, TAYLOR_SIN, reduce_sincos) on |x| < 105414350 — the function numpy.cos resolves
 * to on the hosts this project runs on, hence what the reference's node generator evaluates
 * (reference src/pytristan/grid.py:75).  It is the checker of the DEVICE cosine in
 * pytristan_b200/csrc/glibc_cos.cuh; only tests/ may load it.
 *
 * glibc is a third-party dependency of the reference (through NumPy) that is not in
 * /root/reference; version on this image: glibc 2.39 (Ubuntu 2.39-0ubuntu8.5).  The restatement is
 * pinned against the real libm of the host: tests/test_glibc_cos.py compares it with math.cos /
 * numpy.cos on every CGL argument j*pi/(n-1), n <= 4200, and on random arguments (4e8 were checked
 * when it was written, 0 mismatches, for both builds: the non-FMA build was reached with
 * GLIBC_TUNABLES=glibc.cpu.hwcaps=-FMA).
 *
 * use_fma != 0: the x86-64 -mfma -mavx2 build (__cos_fma), every a*b+c of the source contracted
 * (which products GCC fused was read off the disassembly of libm.so.6); use_fma == 0: the baseline
 * build (__cos_sse2/__cos_avx), every product rounded.  Compile with -ffp-contract=off.
 *
 * The table (sin/cos of k/128 as double-double, __sincostab) is data of the host libm, extracted by
 * tools/gen_glibc_table.py into pytristan_b200/csrc/glibc_sincos_table.inc.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static const double TAB[440] = {
#include "../pytristan_b200/csrc/glibc_sincos_table.inc"
};

static double mad(int f, double a, double b, double c) { return f ? fma(a, b, c) : a * b + c; }
static int lowword(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(uint32_t)u; }
static double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

static const double big = 0x1.8p45, hp0 = 1.5707963267948966, hp1 = 6.123233995736766e-17;
static const double sn3 = -1.66666666666664880952546298448555E-01, sn5 = 8.33333214285722277379541354343671E-03;
static const double cs2 = 4.99999999999999999999950396842453E-01, cs4 = -4.16666666666664434524222570944589E-02,
                    cs6 = 1.38888874007937613028114285595617E-03;
static const double s1 = -0x1.5555555555555p-3, s2 = 0x1.1111111110ECEp-7, s3 = -0x1.A01A019DB08B8p-13,
                    s4 = 0x1.71DE27B9A7ED9p-19, s5 = -0x1.ADDFFC2FCDF59p-26;
static const double hpinv = 0x1.45F306DC9C883p-1, toint = 0x1.8p52;

static double taylor_sin(int f, double xx, double x, double dx) {
  double p = mad(f, s5, xx, s4);
  p = mad(f, p, xx, s3);
  p = mad(f, p, xx, s2);
  p = mad(f, p, xx, s1);
  double t = mad(f, p, x, -(0.5 * dx));
  t = mad(f, t, xx, dx);
  return x + t;
}

static double do_cos(int f, double x, double dx) {
  if (x < 0) dx = -dx;
  double ux = big + fabs(x);
  x = fabs(x) - (ux - big) + dx;
  double xx = x * x;
  double s = mad(f, x * xx, mad(f, xx, sn5, sn3), x);
  double c = xx * mad(f, xx, mad(f, xx, cs6, cs4), cs2);
  int k = lowword(ux) * 4;
  double sn = TAB[k], ssn = TAB[k + 1], cs = TAB[k + 2], ccs = TAB[k + 3];
  double cor = mad(f, -s, ssn, ccs);
  cor = mad(f, -cs, c, cor);
  cor = mad(f, -sn, s, cor);
  return cs + cor;
}

static double do_sin(int f, double x, double dx) {
  double xold = x;
  if (fabs(x) < 0.126) return taylor_sin(f, x * x, x, dx);
  if (x <= 0) dx = -dx;
  double ux = big + fabs(x);
  x = fabs(x) - (ux - big);
  double xx = x * x;
  double s = x + mad(f, x * xx, mad(f, xx, sn5, sn3), dx);
  double q = mad(f, xx, mad(f, xx, cs6, cs4), cs2);
  double c = mad(f, x, dx, xx * q);
  int k = lowword(ux) * 4;
  double sn = TAB[k], ssn = TAB[k + 1], cs = TAB[k + 2], ccs = TAB[k + 3];
  double cor = mad(f, s, ccs, ssn);
  cor = mad(f, -sn, c, cor);
  cor = mad(f, cs, s, cor);
  return copysign(sn + cor, xold);
}

double oracle_glibc_cos(double x, int use_fma) {
  const int f = use_fma;
  uint64_t ux;
  memcpy(&ux, &x, 8);
  int k = (int)((ux >> 32) & 0x7fffffff);
  if (k < 0x3e400000) return 1.0;
  if (k < 0x3feb6000) return do_cos(f, x, 0);
  if (k < 0x400368fd) {
    double y = hp0 - fabs(x), a = y + hp1, da = (y - a) + hp1;
    return do_sin(f, a, da);
  }
  if (k < 0x419921FB) {
    const double mp1 = from_bits(0x3FF921FB58000000ull), mp2 = from_bits(0xBE4DDE973C000000ull),
                 pp3 = from_bits(0xBC8CB3B398000000ull), pp4 = from_bits(0xBACD747F23E32ED7ull);
    double t = mad(f, x, hpinv, toint);
    double xn = t - toint;
    double y = mad(f, -xn, mp2, mad(f, -xn, mp1, x));
    int n = (lowword(t) & 3) + 1;
    double t2 = mad(f, -xn, pp3, y);
    double db = mad(f, -xn, pp3, y - t2);
    double b = mad(f, -xn, pp4, t2);
    db += mad(f, -xn, pp4, t2 - b);
    double r = (n & 1) ? do_cos(f, b, db) : do_sin(f, b, db);
    return (n & 2) ? -r : r;
  }
  return cos(x);
}

/* vectorised entry for the tests: out[i] = oracle_glibc_cos(x[i]) */
void oracle_glibc_cos_array(const double* x, double* out, int64_t n, int use_fma) {
  for (int64_t i = 0; i < n; ++i) out[i] = oracle_glibc_cos(x[i], use_fma);
}

/* CGL nodes exactly as reference grid.py:75 evaluates them: cos(fl(fl(j*pi)/(n-1))) */
void oracle_cheb_nodes(double* out, int64_t n, int use_fma) {
  for (int64_t j = 0; j < n; ++j) out[j] = oracle_glibc_cos(((double)j * M_PI) / (double)(n - 1), use_fma);
}
