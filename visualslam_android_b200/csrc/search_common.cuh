// Device functions shared by the patch-search kernels: k_search / k_epipolar (track.cu) and k_search_fast / k_subpix (search_fast.cu).
#pragma once
#include "track_dev.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// SearchForPoints, one warp per list entry.
constexpr int kCandCap = 96;           // ZMSSD candidates gathered per round of one warp
struct SearchSmem {
  union {   // the three phases of a warp never overlap: template generation (pos), candidate scoring (cand_*), sub-pixel (pos, jx, jy, prod2)
    struct { double pos[VS_MAXP * VS_MAXP * 2]; double jx[81], jy[81], prod2[81]; };   // template sample positions / sub-pixel products and gradients
    struct { uint32_t cand_cw[kCandCap]; int cand_idx[kCandCap]; };
  };
  uint32_t tmpl_w[VS_TMPL_BYTES / 4];  // template, one row = 3 zero-padded words (12 bytes): the dp4a operand layout
};

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// ZMSSDAtPoint (jni/PatchFinder.cc:352-380) of the `ncand` candidates filed in sm.cand_cw / cand_idx; returns min(best, keys) with
// key = ssd << 32 | cand_idx (ties: lowest index).  Eight lanes per candidate, four candidates per step: lane slot r works on template
// rows r and r + 8 (eight image words in flight per lane: the kernel waits on these loads, not on the dp4a pipe), adds its two rows in
// registers and the group's three sums -- packed into one 64-bit word, 23 + 23 + 15 bits hold the totals of an 11x11 patch -- are
// combined by a three-step shuffle butterfly.  (The first version accumulated the rows with shared-memory atomics: up to 11 lanes on one
// address, ten LSU wavefronts per instruction, a fifth of all LSU wavefronts of the kernel, whose LSU data pipe is 77 % busy.)
static_assert(VSLAM_MAX_PATCH <= 16 && VSLAM_MAX_PATCH * VSLAM_MAX_PATCH * 255 * 255 < (1 << 23), "score_candidates: two rows per lane slot, 23-bit packed sums");
template <int PT>
__device__ __forceinline__ unsigned long long score_candidates(SearchSmem& sm, int ncand, const uint8_t* __restrict__ img, int pitch, int lw, int lh, int P,
                                                                int tsum, int tsumsq, int maxSSD, unsigned long long best) {
  const int lane = threadIdx.x & 31, PP = P * P;
  const int b = P / 2, nwords = (P + 3) >> 2;
  const uint32_t lastmask = (P & 3) ? ((1u << (8 * (P & 3))) - 1u) : 0xffffffffu;
  const int slot = lane & 7, grp = lane >> 3;
  for (int c0 = 0; c0 < ncand; c0 += 4) {
    const int c = c0 + grp;
    const bool have = c < ncand;
    const uint32_t cw = sm.cand_cw[have ? c : 0];
    const int cx = cw & 0xffff, cy = cw >> 16;
    const bool inb = have && (cx >= b && cy >= b && cx < lw - b && cy < lh - b);
    uint32_t w[2][4]; unsigned shf[2]; bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int r = slot + 8 * u;
      ok[u] = inb && r < P;
      w[u][0] = w[u][1] = w[u][2] = w[u][3] = 0u; shf[u] = 0;
      if (ok[u]) {
        const uint8_t* rp = img + (size_t)(cy - b + r) * pitch + (cx - b);
        const unsigned a = (unsigned)((uintptr_t)rp & 3u);
        shf[u] = a * 8;
        const uint32_t* wp = (const uint32_t*)(rp - a);
        w[u][0] = __ldg(wp); w[u][1] = __ldg(wp + 1);
        if (a + P > 8) w[u][2] = __ldg(wp + 2);
        if (a + P > 12) w[u][3] = __ldg(wp + 3);
      }
    }
    unsigned long long acc = 0ull;
#pragma unroll
    for (int u = 0; u < 2; u++) {
      if (!ok[u]) continue;
      uint32_t n0 = __funnelshift_r(w[u][0], w[u][1], shf[u]), n1 = __funnelshift_r(w[u][1], w[u][2], shf[u]), n2 = __funnelshift_r(w[u][2], w[u][3], shf[u]);
      if (nwords == 3) n2 &= lastmask; else if (nwords == 2) { n1 &= lastmask; n2 = 0; } else { n0 &= lastmask; n1 = 0; n2 = 0; }
      const int r = slot + 8 * u;
      unsigned sum = __dp4a(n0, 0x01010101u, 0u), sumsq = __dp4a(n0, n0, 0u), cross = __dp4a(n0, sm.tmpl_w[3 * r], 0u);
      sum = __dp4a(n1, 0x01010101u, sum); sumsq = __dp4a(n1, n1, sumsq); cross = __dp4a(n1, sm.tmpl_w[3 * r + 1], cross);
      sum = __dp4a(n2, 0x01010101u, sum); sumsq = __dp4a(n2, n2, sumsq); cross = __dp4a(n2, sm.tmpl_w[3 * r + 2], cross);
      acc += (unsigned long long)cross | ((unsigned long long)sumsq << 23) | ((unsigned long long)sum << 46);
    }
#pragma unroll
    for (int d = 4; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (slot == 0 && have) {
      int ssd;
      if (!inb) ssd = maxSSD + 1;
      else {
        const int SA = tsum, SB = (int)(acc >> 46), sq = (int)((acc >> 23) & 0x7fffffull), cr = (int)(acc & 0x7fffffull);
        ssd = ((2 * SA * SB - SA * SA - SB * SB) / PP + sq + tsumsq - 2 * cr);
      }
      const unsigned long long key = ((unsigned long long)(unsigned)ssd << 32) | (unsigned)sm.cand_idx[c];   // ssd >= 0; ties -> lowest corner index
      best = key < best ? key : best;
    }
  }
  return best;
}

// MakeSubPixTemplate + IterateSubPixToConvergence (jni/PatchFinder.cc:242-350) around (coarse0, coarse1) (level-0 pixels) in the
// level image `img`; tmpl: the template in 12-byte rows.  Returns 1 if converged; (out0, out1) = mv2SubPixPos in any case.
__device__ __forceinline__ int subpix_refine(SearchSmem& sm, const uint8_t* tmpl, const uint8_t* __restrict__ img, int pitch, int lw, int lh, int level, int P,
                                             int subpix, double coarse0, double coarse1, double& out0, double& out1) {
  const int lane = threadIdx.x & 31;
  const int nLevelScale = LevelScale(level);
  const double invScale = 1.0 / nLevelScale;
  // ---- MakeSubPixTemplate (jni/PatchFinder.cc:242-267)
  const int Q = P - 2, QQ = Q * Q;
  for (int k = lane; k < QQ; k += 32) {
    const int x = k / Q + 1, y = k - (x - 1) * Q + 1;   // stored index (x-1)*Q + (y-1)
    sm.jx[k] = 0.5 * (tmpl[y * 12 + x + 1] - tmpl[y * 12 + x - 1]);
    sm.jy[k] = 0.5 * (tmpl[(y + 1) * 12 + x] - tmpl[(y - 1) * 12 + x]);
  }
  __syncwarp();
  // JtJ of (gx, gy, 1): sums of multiples of 0.25 below 2^53 are exact in any order, so a warp reduction is bit-exact
  double hxx = 0, hxy = 0, hyy = 0, hx = 0, hy = 0;
  for (int k = lane; k < QQ; k += 32) { const double gx = sm.jx[k], gy = sm.jy[k]; hxx += gx * gx; hxy += gx * gy; hyy += gy * gy; hx += gx; hy += gy; }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    hxx += __shfl_xor_sync(0xffffffffu, hxx, d); hxy += __shfl_xor_sync(0xffffffffu, hxy, d); hyy += __shfl_xor_sync(0xffffffffu, hyy, d);
    hx += __shfl_xor_sync(0xffffffffu, hx, d); hy += __shfl_xor_sync(0xffffffffu, hy, d);
  }
  const double H[9] = {hxx, hxy, hx, hxy, hyy, hy, hx, hy, (double)QQ};
  double hinv[9];
  {   // 3x3 inverse: adjugate * (1/det), evaluation order of the oracle (oracle/vslam_oracle.cc inverse3)
    const double c00 = H[4] * H[8] - H[5] * H[7], c10 = H[5] * H[6] - H[3] * H[8], c20 = H[3] * H[7] - H[4] * H[6];
    const double det = H[0] * c00 + H[1] * c10 + H[2] * c20, invdet = 1.0 / det;
    hinv[0] = c00 * invdet; hinv[3] = c10 * invdet; hinv[6] = c20 * invdet;
    hinv[1] = (H[2] * H[7] - H[1] * H[8]) * invdet; hinv[4] = (H[0] * H[8] - H[2] * H[6]) * invdet; hinv[7] = (H[1] * H[6] - H[0] * H[7]) * invdet;
    hinv[2] = (H[1] * H[5] - H[2] * H[4]) * invdet; hinv[5] = (H[2] * H[3] - H[0] * H[5]) * invdet; hinv[8] = (H[0] * H[4] - H[1] * H[3]) * invdet;
  }
  double sp0 = coarse0, sp1 = coarse1, meanDiff = 0.0;
  int ok = 0;
  // ---- IterateSubPixToConvergence / IterateSubPix (jni/PatchFinder.cc:272-350)
  for (int it = 0; it < subpix; it++) {
    const double c0 = (sp0 + 0.5) * invScale - 0.5, c1 = (sp1 + 0.5) * invScale - 0.5;   // LevelNPos
    const int xb = (c0 > 0.0 ? c0 + 0.5 : c0 - 0.5), yb = (c1 > 0.0 ? c1 + 0.5 : c1 - 0.5);
    const int bd = P / 2 + 1;
    if (!(xb >= bd && yb >= bd && xb < lw - bd && yb < lh - bd)) break;   // off the image: not converged
    const double b0 = c0 - (double)(P / 2), b1 = c1 - (double)(P / 2);
    const double dX = b0 - floor(b0), dY = b1 - floor(b1);
    const float fTL = (1.0 - dX) * (1.0 - dY), fTR = (dX) * (1.0 - dY), fBL = (1.0 - dX) * (dY), fBR = (dX) * (dY);
    for (int k = lane; k < QQ; k += 32) {   // k = (y-1)*Q + (x-1): the reference's loop order
      const int y = k / Q + 1, x = k - (y - 1) * Q + 1;
      const uint8_t* tl = img + (size_t)((int)b1 + y) * pitch + ((int)b0 + x);
      const float fPixel = fTL * tl[0] + fTR * tl[1] + fBL * tl[pitch] + fBR * tl[pitch + 1];
      const double dDiff = fPixel - tmpl[y * 12 + x] + meanDiff;
      const int j = (x - 1) * Q + (y - 1);
      sm.pos[k] = dDiff * sm.jx[j]; sm.pos[QQ + k] = dDiff * sm.jy[j]; sm.prod2[k] = dDiff;
    }
    __syncwarp();
    double acc = 0;   // lanes 0,1,2 add their accumulator's terms in pixel order, like the reference's serial loop
    if (lane < 3) {   // (nine terms are fetched before they are added: the chain is 81 additions, not 81 x (shared-memory load + addition))
      const double* p = lane == 0 ? sm.pos : (lane == 1 ? sm.pos + QQ : sm.prod2);
      int k = 0;
      for (; k + 9 <= QQ; k += 9) {
        double t[9];
#pragma unroll
        for (int u = 0; u < 9; u++) t[u] = p[k + u];
#pragma unroll
        for (int u = 0; u < 9; u++) acc += t[u];
      }
      for (; k < QQ; k++) acc += p[k];
    }
    const double a0 = __shfl_sync(0xffffffffu, acc, 0), a1 = __shfl_sync(0xffffffffu, acc, 1), a2 = __shfl_sync(0xffffffffu, acc, 2);
    __syncwarp();
    double upd[3];
#pragma unroll
    for (int r = 0; r < 3; r++) { double sacc = hinv[3 * r] * a0; sacc += hinv[3 * r + 1] * a1; sacc += hinv[3 * r + 2] * a2; upd[r] = sacc; }
    sp0 -= upd[0] * nLevelScale; sp1 -= upd[1] * nLevelScale;
    meanDiff -= upd[2];
    double d = 0; d += upd[0] * upd[0]; d += upd[1] * upd[1];
    const double lim = 0.03;
    if (d < lim * lim) { ok = 1; break; }
  }
  out0 = sp0; out1 = sp1;
  return ok;
}

}  // namespace
