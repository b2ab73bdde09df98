// Device cosine that reproduces glibc's double-precision cos (IBM Accurate Mathematical Library,
// glibc 2.28+ sysdeps/ieee754/dbl-64/s_sin.c) bit for bit on the argument range of the
// Chebyshev-Gauss-Lobatto generator, |x| <= 105414350 (the reference computes its nodes as
// numpy.cos(j*pi/(n-1)), reference grid.py:75, and numpy.cos is the host libm's cos).
//
// glibc's result is NOT the correctly rounded cosine (max 0.52 ulp), so matching it requires the
// same algorithm, the same table (csrc/glibc_sincos_table.inc, tools/gen_glibc_table.py) and the
// same roundings.  On x86-64 glibc ships two builds of this code, selected at load time:
//   FMA  = true : the -mfma -mavx2 build (__cos_fma): GCC contracted every a*b+c of the source
//                 into one fused multiply-add (which products were fused was read off the
//                 disassembly and confirmed on 4e8 random arguments: oracle/glibc_cos.c);
//   FMA  = false: the baseline build (__cos_sse2 / __cos_avx): every product rounded separately.
// All arithmetic below uses the explicit-rounding intrinsics so that nvcc contracts nothing itself.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptglibc {

__device__ const double kSincosTab[440] = {
#include "glibc_sincos_table.inc"
};

template <bool FMA>
__device__ __forceinline__ double mad(double a, double b, double c) {
  return FMA ? __fma_rn(a, b, c) : __dadd_rn(__dmul_rn(a, b), c);
}
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }

constexpr double kBig = 0x1.8p45;                       // |x| + kBig leaves round(|x|*128) in the low word
constexpr double kSn3 = -1.66666666666664880952546298448555E-01;
constexpr double kSn5 = 8.33333214285722277379541354343671E-03;
constexpr double kCs2 = 4.99999999999999999999950396842453E-01;
constexpr double kCs4 = -4.16666666666664434524222570944589E-02;
constexpr double kCs6 = 1.38888874007937613028114285595617E-03;
constexpr double kS1 = -0x1.5555555555555p-3, kS2 = 0x1.1111111110ECEp-7, kS3 = -0x1.A01A019DB08B8p-13,
                 kS4 = 0x1.71DE27B9A7ED9p-19, kS5 = -0x1.ADDFFC2FCDF59p-26;
constexpr double kHp0 = 1.5707963267948966, kHp1 = 6.123233995736766e-17;
constexpr double kHpInv = 0x1.45F306DC9C883p-1, kToInt = 0x1.8p52;

// sin(x + dx) for |x| < 0.126 (TAYLOR_SIN)
template <bool FMA>
__device__ __forceinline__ double taylor_sin(double xx, double x, double dx) {
  double p = mad<FMA>(kS5, xx, kS4);
  p = mad<FMA>(p, xx, kS3);
  p = mad<FMA>(p, xx, kS2);
  p = mad<FMA>(p, xx, kS1);
  double t = mad<FMA>(p, x, -mul(0.5, dx));
  t = mad<FMA>(t, xx, dx);
  return add(x, t);
}

template <bool FMA>
__device__ __forceinline__ double do_cos(double x, double dx) {
  if (x < 0.0) dx = -dx;
  const double ux = add(kBig, fabs(x));
  x = add(sub(fabs(x), sub(ux, kBig)), dx);
  const double xx = mul(x, x);
  const double s = mad<FMA>(mul(x, xx), mad<FMA>(xx, kSn5, kSn3), x);
  const double c = mul(xx, mad<FMA>(xx, mad<FMA>(xx, kCs6, kCs4), kCs2));
  const int k = __double2loint(ux) * 4;
  const double sn = kSincosTab[k], ssn = kSincosTab[k + 1], cs = kSincosTab[k + 2], ccs = kSincosTab[k + 3];
  double cor = mad<FMA>(-s, ssn, ccs);
  cor = mad<FMA>(-cs, c, cor);
  cor = mad<FMA>(-sn, s, cor);
  return add(cs, cor);
}

template <bool FMA>
__device__ __forceinline__ double do_sin(double x, double dx) {
  const double xold = x;
  if (fabs(x) < 0.126) return taylor_sin<FMA>(mul(x, x), x, dx);
  if (x <= 0.0) dx = -dx;
  const double ux = add(kBig, fabs(x));
  x = sub(fabs(x), sub(ux, kBig));
  const double xx = mul(x, x);
  const double s = add(x, mad<FMA>(mul(x, xx), mad<FMA>(xx, kSn5, kSn3), dx));
  const double q = mad<FMA>(xx, mad<FMA>(xx, kCs6, kCs4), kCs2);
  const double c = mad<FMA>(x, dx, mul(xx, q));
  const int k = __double2loint(ux) * 4;
  const double sn = kSincosTab[k], ssn = kSincosTab[k + 1], cs = kSincosTab[k + 2], ccs = kSincosTab[k + 3];
  double cor = mad<FMA>(s, ccs, ssn);
  cor = mad<FMA>(-sn, c, cor);
  cor = mad<FMA>(cs, s, cor);
  return copysign(add(sn, cor), xold);
}

// cos(x) as glibc computes it; falls back to the CUDA cosine outside the table-driven ranges
// (|x| >= 105414350, inf, nan), which the node generator never reaches.
template <bool FMA>
__device__ double glibc_cos(double x) {
  const int k = __double2hiint(x) & 0x7fffffff;
  if (k < 0x3e400000) return 1.0;                       // |x| < 2^-27
  if (k < 0x3feb6000) return do_cos<FMA>(x, 0.0);       // |x| < 0.855469
  if (k < 0x400368fd) {                                 // |x| < 2.426265: cos x = sin(pi/2 - |x|)
    const double y = sub(kHp0, fabs(x));
    const double a = add(y, kHp1);
    const double da = add(sub(y, a), kHp1);
    return do_sin<FMA>(a, da);
  }
  if (k < 0x419921FB) {                                 // |x| < 105414350: reduce by multiples of pi/2
    const double mp1 = __longlong_as_double(0x3FF921FB58000000LL);
    const double mp2 = __longlong_as_double(0xBE4DDE973C000000LL);
    const double pp3 = __longlong_as_double(0xBC8CB3B398000000LL);
    const double pp4 = __longlong_as_double(0xBACD747F23E32ED7LL);
    const double t = mad<FMA>(x, kHpInv, kToInt);
    const double xn = sub(t, kToInt);
    const double y = mad<FMA>(-xn, mp2, mad<FMA>(-xn, mp1, x));
    const int n = (__double2loint(t) & 3) + 1;
    const double t2 = mad<FMA>(-xn, pp3, y);
    double db = mad<FMA>(-xn, pp3, sub(y, t2));
    const double b = mad<FMA>(-xn, pp4, t2);
    db = add(db, mad<FMA>(-xn, pp4, sub(t2, b)));
    const double r = (n & 1) ? do_cos<FMA>(b, db) : do_sin<FMA>(b, db);
    return (n & 2) ? -r : r;
  }
  return cos(x);
}

}  // namespace ptglibc
