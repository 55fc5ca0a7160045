// Correctly-rounded atan(x) in double-double arithmetic.
//
// Why: the template pixels of PatchFinder are the TRUNCATION of a double bilinear blend
// (jni/vision/ImageHandler.cpp:12-19) at positions that depend on ATANCamera::Project's atan
// (jni/ATANCamera.h:136-142).  CUDA's atan() is accurate to 2 ulp, glibc's to < 1 ulp: a last-bit
// difference moves a sample position by 1e-16 and can flip a truncated pixel in a flat image region.
// This routine returns the correctly rounded value, which is what glibc 2.39 returns for 99.99 % of
// arguments (measured in tests/test_oracle_cpu.py::test_glibc_atan_is_almost_correctly_rounded).
//
// Method: x > 1 -> pi/2 - atan(1/x); t in [0,1] -> atan(k/8) + atan(u), u = (t - k/8)/(1 + t k/8),
// |u| <= 1/16; atan(u) = u * sum_{j<16} (-1)^j u^(2j)/(2j+1) by Horner in double-double (error < 2^-100).
#pragma once

namespace ddm {
struct dd { double hi, lo; };
__device__ __forceinline__ dd two_sum(double a, double b) { const double s = a + b, bb = s - a; return {s, (a - (s - bb)) + (b - bb)}; }
__device__ __forceinline__ dd fast_two_sum(double a, double b) { const double s = a + b; return {s, b - (s - a)}; }
__device__ __forceinline__ dd two_prod(double a, double b) { const double p = a * b; return {p, __fma_rn(a, b, -p)}; }
__device__ __forceinline__ dd add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi); const dd t = two_sum(a.lo, b.lo);
  s.lo += t.hi; s = fast_two_sum(s.hi, s.lo); s.lo += t.lo; return fast_two_sum(s.hi, s.lo);
}
__device__ __forceinline__ dd mul(dd a, dd b) { dd p = two_prod(a.hi, b.hi); p.lo += a.hi * b.lo + a.lo * b.hi; return fast_two_sum(p.hi, p.lo); }
__device__ __forceinline__ dd neg(dd a) { return {-a.hi, -a.lo}; }
__device__ __forceinline__ dd div(dd a, dd b) {
  const double q1 = a.hi / b.hi;
  dd r = add(a, neg(mul(b, dd{q1, 0.0})));
  const double q2 = r.hi / b.hi;
  r = add(r, neg(mul(b, dd{q2, 0.0})));
  const double q3 = r.hi / b.hi;
  dd q = fast_two_sum(q1, q2);
  return add(q, dd{q3, 0.0});
}
}  // namespace ddm

static __device__ __noinline__ double atan_cr_dd(double x) {
  using namespace ddm;
  if (!(x == x)) return x;
  const bool negx = x < 0; if (negx) x = -x;
  if (x == 0.0) return negx ? -0.0 : 0.0;
  if (x > 1e18) return negx ? -1.5707963267948966 : 1.5707963267948966;
  if (x < 1e-9) return negx ? -x : x;   // atan(x) = x(1 - x^2/3 ...): rounds to x below 2^-27
  static const double kAtanHi[9] = {0x0.0p+0, 0x1.fd5ba9aac2f6ep-4, 0x1.f5b75f92c80ddp-3, 0x1.6f61941e4def1p-2, 0x1.dac670561bb4fp-2,
                                    0x1.1e00babdefeb4p-1, 0x1.4978fa3269ee1p-1, 0x1.700a7c5784634p-1, 0x1.921fb54442d18p-1};
  static const double kAtanLo[9] = {0x0.0p+0, -0x1.cd37686760c17p-59, 0x1.8ab6e3cf7afbdp-57, -0x1.c63aae6f6e918p-56, 0x1.a2b7f222f65e2p-56,
                                    -0x1.928df287a668fp-58, 0x1.2419a87f2a458p-56, -0x1.8c34d25aadef6p-56, 0x1.1a62633145c07p-55};
  static const double kCHi[16] = {0x1.0000000000000p+0, -0x1.5555555555555p-2, 0x1.999999999999ap-3, -0x1.2492492492492p-3, 0x1.c71c71c71c71cp-4,
                                  -0x1.745d1745d1746p-4, 0x1.3b13b13b13b14p-4, -0x1.1111111111111p-4, 0x1.e1e1e1e1e1e1ep-5, -0x1.af286bca1af28p-5,
                                  0x1.8618618618618p-5, -0x1.642c8590b2164p-5, 0x1.47ae147ae147bp-5, -0x1.2f684bda12f68p-5, 0x1.1a7b9611a7b96p-5,
                                  -0x1.0842108421084p-5};
  static const double kCLo[16] = {0x0.0p+0, -0x1.5555555555555p-56, -0x1.999999999999ap-57, -0x1.2492492492492p-57, 0x1.c71c71c71c71cp-58,
                                  0x1.745d1745d1746p-59, -0x1.3b13b13b13b14p-58, -0x1.1111111111111p-60, 0x1.e1e1e1e1e1e1ep-61, -0x1.af286bca1af28p-59,
                                  0x1.8618618618618p-59, -0x1.642c8590b2164p-60, -0x1.eb851eb851eb8p-61, -0x1.2f684bda12f68p-59, 0x1.1a7b9611a7b96p-61,
                                  -0x1.0842108421084p-60};
  const bool inv = x > 1.0;
  dd t = inv ? div(dd{1.0, 0.0}, dd{x, 0.0}) : dd{x, 0.0};
  const int k = (int)(t.hi * 8.0 + 0.5);
  dd u = t;
  if (k > 0) {
    const double c = k * 0.125;   // exact
    u = div(add(t, dd{-c, 0.0}), add(dd{1.0, 0.0}, mul(t, dd{c, 0.0})));
  }
  const dd z = mul(u, u);
  dd s = {kCHi[15], kCLo[15]};
#pragma unroll
  for (int j = 14; j >= 0; j--) s = add(mul(s, z), dd{kCHi[j], kCLo[j]});
  dd r = add(mul(u, s), dd{kAtanHi[k], kAtanLo[k]});
  if (inv) r = add(dd{0x1.921fb54442d18p+0, 0x1.1a62633145c07p-54}, neg(r));
  return negx ? -r.hi : r.hi;
}

// Fast path of atan_cr for |x| <= 1/16 (the FOV camera of the reference calls atan(r * 2 tan(w/2)) with |w| ~ 0.013, so every call of
// the tracker lands here): atan(x) = x - x^3/3 + x^5 (1/5 - z/7 + ... - z^7/19), z = x^2, with the cubic term in double-double and the
// rest in double -- absolute error below |x| 2^-68 -- followed by Ziv's rounding test: if adding and subtracting the error bound
// round to the same double, that double IS the correctly rounded result; otherwise (about one call in 10^4) the double-double
// routine above decides.  ~45 flops instead of ~800: the projection passes of k_project_lists / k_pose are FP64-pipe bound.
__device__ __forceinline__ double atan_cr(double x) {
  const double ax = fabs(x);
  if (ax <= 0.0625 && ax >= 1e-9) {
    const double zh = x * x, zl = __fma_rn(x, x, -zh);
    double r = -0x1.af286bca1af28p-5;                 // -1/19
    r = __fma_rn(r, zh, 0x1.e1e1e1e1e1e1ep-5);        //  1/17
    r = __fma_rn(r, zh, -0x1.1111111111111p-4);       // -1/15
    r = __fma_rn(r, zh, 0x1.3b13b13b13b14p-4);        //  1/13
    r = __fma_rn(r, zh, -0x1.745d1745d1746p-4);       // -1/11
    r = __fma_rn(r, zh, 0x1.c71c71c71c71cp-4);        //  1/9
    r = __fma_rn(r, zh, -0x1.2492492492492p-3);       // -1/7
    r = __fma_rn(r, zh, 0x1.999999999999ap-3);        //  1/5
    const double t5 = ((x * zh) * zh) * r;
    const double ph = x * zh, pl = __fma_rn(x, zh, -ph) + x * zl;                                                    // x^3
    const double th = -0x1.5555555555555p-2, tl = -0x1.5555555555555p-56;                                             // -1/3
    const double qh = ph * th, ql = (__fma_rn(ph, th, -qh) + ph * tl) + pl * th;                                      // -x^3/3
    const double sh = x + qh, bb = sh - x, sl = (x - (sh - bb)) + (qh - bb);                                          // two_sum(x, qh)
    const double low = sl + (ql + t5);
    const double res = sh + low, E = ax * 0x1.0p-67;
    if (sh + (low + E) == res && sh + (low - E) == res) return res;
  }
  return atan_cr_dd(x);
}
