// Tracker::SearchForPoints (jni/Tracker.cc:629-674) with EIGHT LANES PER MAP POINT, four points per warp.
//
// k_search (track.cu) spends one warp per list entry: about 1200 warp-instructions per point at the tracker's workload (a 21 x 21
// search window, about 130 FAST corners in its rows, about 6 of them inside the circle), most of them uniform bookkeeping executed by
// 32 lanes for one point, and the kernel is bound by instruction issue (71 % of the issue slots).  Here a point gets a quarter warp:
//   * per-point state, the re-use test of MakeTemplateCoarseCont (jni/PatchFinder.cc:79-125) and the search window (:170-209) cost one
//     instruction stream per FOUR points;
//   * the FAST corners of the window come straight from the level's corner bitmask (LevelDesc::cbits), in batches of 24 rows x 32 columns: a
//     lane fetches its three rows of a batch at once and cuts each down to a 32-bit mask; a set bit is a corner position (no corner list, no
//     row LUT); one corner per lane and step goes through the FP64 circle test of :216-219, survivors to a per-point queue in shared memory;
//   * ZMSSD (jni/PatchFinder.cc:352-380): ONE LANE PER CANDIDATE walks all P rows -- unaligned P-byte rows through aligned 32-bit loads
//     and a funnel shift, dp4a for sum I, sum I^2, sum I.T -- so the three sums stay in registers (no cross-lane reduction), and a round of up to
//     eight candidates per point costs one instruction stream for four points;
//   * argmin over the 64-bit key (ssd << 32 | y << 16 | x) = "first in raster order wins ties" (:223-226) by three shuffle steps.
// A template that has to be regenerated (rare: the 0.07 re-use test keeps > 99 % of them at tracking speed) is produced by the whole warp
// for that point, exactly like k_search does it (serial position accumulation of transform_image, jni/vision/ImageHandler.cpp:21-113).
// The sub-pixel refinement (jni/PatchFinder.cc:242-350) of the entries that ask for it runs in a second kernel, k_subpix, eight lanes per
// entry, on the coarse result this kernel leaves behind.
#include "search_common.cuh"
#include <algorithm>

namespace {

constexpr int kFW = 4;            // warps per CTA
constexpr int kGL = 8;            // lanes per list entry
constexpr int kGW = 32 / kGL;     // entries per warp
constexpr int kQ = 48;            // candidate queue per entry
#ifndef VS_SEARCH_MINB
#define VS_SEARCH_MINB 6
#endif
#ifndef VS_SEARCH_NB
#define VS_SEARCH_NB 4
#endif
constexpr int kNB = VS_SEARCH_NB; // template rows whose image words are requested together by the ZMSSD (registers: 4 per row)

struct FastWarp {
  union {
    struct { uint32_t cw[kGW][kQ]; } q;                        // candidates waiting for their ZMSSD: corner word (y << 16 | x)
    double pos[VS_MAXP * VS_MAXP * 2];                         // template regeneration: sample positions
  };
  uint32_t tmpl[kGW][VS_TMPL_BYTES / 4];                       // the entries' templates, rows of 3 zero-padded words (the dp4a operand layout)
};

struct Plan { int first, count, range, subpix_all, n_top, subpix_top; };
// which entries a launch covers and with what range / sub-pixel iterations (jni/Tracker.cc:464-532); false: nothing to do for this stream
__device__ __forceinline__ bool search_plan(const Dev& D, const StreamState* st, int mode, int range_arg, int subpix_arg, Plan& p) {
  if (mode != 0 && st->lost_frames >= 3 && !st->recovered) return false;
  if (mode != 0 && other_chain(D, st)) return false;
  p.n_top = 0; p.subpix_top = 0;
  if (mode == 0) { p.first = 0; p.count = st->nA; p.range = range_arg; p.subpix_all = subpix_arg; }
  else if (mode == 1) { if (!st->try_coarse) return false; p.first = 0; p.count = st->nA; p.range = st->coarse_range; p.subpix_all = D.prm.coarse_subpix_its; }
  else { p.first = st->nA; p.count = st->nB; p.range = st->did_coarse ? D.prm.fine_range_after_coarse : D.prm.fine_range; p.subpix_all = 0; p.n_top = st->nB_top; p.subpix_top = D.prm.fine_subpix_its_top_level; }
  return true;
}

// read-only load that keeps its place among its siblings (volatile asm): ptxas otherwise sinks the loads of a window between the dp4a
// of the previous rows to save registers, which turns eleven rows into eleven dependent round trips
__device__ __forceinline__ uint32_t ldg_ordered(const uint32_t* p) { uint32_t v; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }

constexpr int kSearchRefindF = 1;   // sflags: MapMaker::ReFind_Common's variant (jni/MapMaker.cc:967-1036), see k_search
constexpr int kSearchTemplateOnlyF = 2;   // sflags: stop after MakeTemplateCoarseCont (vslam_pf_make_template, the per-object PatchFinder path)

template <int PT>
__global__ void __launch_bounds__(kFW * 32, VS_SEARCH_MINB) k_search_fast(Dev D, int mode, int range_arg, int subpix_arg, int sflags) {
  cudaGridDependencySynchronize(); cudaTriggerProgrammaticLaunchCompletion();   // programmatic dependent launch (vs_launch_pdl); no-ops otherwise
  __shared__ FastWarp sm_all[kFW];
  const int s = blockIdx.y + D.s0, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane / kGL, j = lane % kGL;
  StreamState* st = D.ss + s;
  Plan pl;
  if (!search_plan(D, st, mode, range_arg, subpix_arg, pl)) return;
  const int e0 = (blockIdx.x * kFW + warp) * kGW;
  if (e0 >= pl.count) return;                                   // (warp-uniform)
  FastWarp& W = sm_all[warp];
  const int e = e0 + g;
  bool alive = e < pl.count;
  const int i = D.lists[(size_t)s * D.list_cap + pl.first + (alive ? e : e0)];
  const size_t SN = (size_t)D.S * D.N, gi = (size_t)s * D.N + i;
  const int P = PT ? PT : D.P, PP = P * P;
  const bool refind = (sflags & kSearchRefindF) != 0;
  int subpix = e < pl.n_top ? pl.subpix_top : pl.subpix_all;
  int flags = D.ps.flags[gi];
  const int level = refind ? D.ps.rlevel[gi] : D.ps.level[gi];
  if (refind) {
    flags &= ~(F_SEARCHED | F_FOUND | F_SUBPIX);
    if (alive && !(flags & F_INIMAGE)) { if (j == 0) D.ps.flags[gi] = flags; alive = false; }   // not in this keyframe's image: "never retry"
    if (level == 0) subpix = 0;
  }

  // ---- MakeTemplateCoarseCont (jni/PatchFinder.cc:79-125): the re-use test, per entry
  double m2[4];
  m2[0] = D.ps.m2[gi]; m2[1] = D.ps.m2[SN + gi]; m2[2] = D.ps.m2[2 * SN + gi]; m2[3] = D.ps.m2[3 * SN + gi];
  const double lw0 = D.ps.lastwarp[gi], lw1 = D.ps.lastwarp[SN + gi], lw2 = D.ps.lastwarp[2 * SN + gi], lw3 = D.ps.lastwarp[3 * SN + gi];
  const double v2i0 = D.ps.v2image[gi], v2i1 = D.ps.v2image[SN + gi];
  int tsum = D.ps.tsum[gi], tsumsq = D.ps.tsum[SN + gi];
  const uint32_t* tmpl_g = (const uint32_t*)(D.ps.tmpl + gi * VS_TMPL_BYTES);
  uint32_t told[5];
#pragma unroll
  for (int q = 0; q < 5; q++) told[q] = (j + kGL * q < VS_TMPL_BYTES / 4) ? tmpl_g[j + kGL * q] : 0u;
  bool refresh = refind || !(flags & F_HAVELAST);
  if (!refresh) { const double d0 = m2[0] - lw0, d1 = m2[2] - lw2; double dd = 0; dd += d0 * d0; dd += d1 * d1; if (dd > 0.07 * 0.07) refresh = true; }
  if (!refresh) { const double d0 = m2[1] - lw1, d1 = m2[3] - lw3; double dd = 0; dd += d0 * d0; dd += d1 * d1; if (dd > 0.07 * 0.07) refresh = true; }
  uint32_t* const tw = W.tmpl[g];
  if (!refresh) {
    flags &= ~F_NEWTMPL;
#pragma unroll
    for (int q = 0; q < 5; q++) if (j + kGL * q < VS_TMPL_BYTES / 4) tw[j + kGL * q] = told[q];
  }
  __syncwarp();
  // ---- templates that have to be regenerated: the whole warp, one entry after the other (transform_image, jni/vision/ImageHandler.cpp:21-113)
  unsigned need = __ballot_sync(0xffffffffu, alive && refresh && j == 0);
  while (need) {
    const int gl = __ffs(need) - 1, gq = gl / kGL; need &= need - 1;
    const double b0 = __shfl_sync(0xffffffffu, m2[0], gl), b1 = __shfl_sync(0xffffffffu, m2[1], gl), b2 = __shfl_sync(0xffffffffu, m2[2], gl), b3 = __shfl_sync(0xffffffffu, m2[3], gl);
    const int bi = __shfl_sync(0xffffffffu, i, gl);
    const size_t bgi = (size_t)s * D.N + bi;
    const int kf = D.map.srckf[bi], sl = D.map.srclevel[bi];
    const uint8_t* simg = D.src.img[sl] + (size_t)kf * D.src.h[sl] * D.src.pitch[sl];
    const int iw = D.src.w[sl], ih = D.src.h[sl], sp = D.src.pitch[sl];
    int inside = 0;
    if (lane == 0) {   // sample positions by sequential accumulation, like the reference
      const double across0 = b0, across1 = b2, down0 = b1, down1 = b3;
      const double o = (double)(P / 2);
      double a = b0 * o; a += b1 * o; double b = b2 * o; b += b3 * o;
      const double p00 = (double)D.map.ircenter[2 * bi] - a, p01 = (double)D.map.ircenter[2 * bi + 1] - b;
      double min_x = p00, min_y = p01, max_x = min_x, max_y = min_y;
      if (across0 < 0) min_x += P * across0; else max_x += P * across0;
      if (down0 < 0) min_x += P * down0; else max_x += P * down0;
      if (across1 < 0) min_y += P * across1; else max_y += P * across1;
      if (down1 < 0) min_y += P * down1; else max_y += P * down1;
      const double cr0 = down0 - P * across0, cr1 = down1 - P * across1;
      inside = (min_x >= 0 && min_y >= 0 && max_x < iw - 1 && max_y < ih - 1);
      double px = p00, py = p01;
      for (int r = 0; r < P; ++r, px += cr0, py += cr1)
        for (int c = 0; c < P; ++c, px += across0, py += across1) { W.pos[2 * (r * P + c)] = px; W.pos[2 * (r * P + c) + 1] = py; }
    }
    inside = __shfl_sync(0xffffffffu, inside, 0);
    uint32_t* const bw = W.tmpl[gq]; uint8_t* const bt = (uint8_t*)bw;
    bw[lane] = 0u; if (lane < VS_TMPL_BYTES / 4 - 32) bw[32 + lane] = 0u;   // row padding must be zero
    __syncwarp();
    const float x_bound = iw - 1, y_bound = ih - 1;
    int outside = 0;
    for (int k = lane; k < PP; k += 32) {
      double x = W.pos[2 * k], y = W.pos[2 * k + 1];
      uint8_t v = 0;
      if (inside || (0 <= x && 0 <= y && x < x_bound && y < y_bound)) {   // sample(u8) (jni/vision/ImageHandler.cpp:12-19)
        const int lx = (int)x, ly = (int)y;
        x -= lx; y -= ly;
        const uint8_t* r0 = simg + (size_t)ly * sp + lx; const uint8_t* r1 = r0 + sp;
        v = (uint8_t)((1 - y) * ((1 - x) * r0[0] + x * r0[1]) + y * ((1 - x) * r1[0] + x * r1[1]));
      } else outside++;
      const int r = k / P;
      bt[k + r * (12 - P)] = v;
    }
    outside = warp_sum(outside);
    __syncwarp();
    int ts = 0, tq = 0;
    for (int k = lane; k < 3 * P; k += 32) { const uint32_t w = bw[k]; ts += (int)__dp4a(w, 0x01010101u, 0u); tq += (int)__dp4a(w, w, 0u); }   // padding is zero
    { uint32_t* gt = (uint32_t*)(D.ps.tmpl + bgi * VS_TMPL_BYTES); gt[lane] = bw[lane]; if (lane < VS_TMPL_BYTES / 4 - 32) gt[32 + lane] = bw[32 + lane]; }
    ts = warp_sum(ts); tq = warp_sum(tq);
    if (g == gq) {
      tsum = ts; tsumsq = tq;
      flags = outside ? (flags | F_TBAD) : (flags & ~F_TBAD);
      flags |= F_HAVELAST | F_NEWTMPL;
    }
    if (lane == 0) {
      atomicAdd(D.evals + 2, 1ull);
      D.ps.tsum[bgi] = ts; D.ps.tsum[SN + bgi] = tq;
      D.ps.lastwarp[bgi] = b0; D.ps.lastwarp[SN + bgi] = b1; D.ps.lastwarp[2 * SN + bgi] = b2; D.ps.lastwarp[3 * SN + bgi] = b3;
    }
    __syncwarp();
  }
  if (sflags & kSearchTemplateOnlyF) { if (alive && j == 0) D.ps.flags[gi] = flags; return; }
  if (alive && (flags & F_TBAD)) {   // jni/Tracker.cc:637-640
    if (j == 0) D.ps.flags[gi] = flags & ~(F_INIMAGE | F_FOUND);
    alive = false;
  }
  if (alive && j == 0) atomicAdd(&st->attempted[level], 1);
  __syncwarp();

  // ---- FindPatchCoarse (jni/PatchFinder.cc:170-235)
  const int lv = alive ? level : 0;
  const LevelDesc& L = D.lev[lv];
  const uint8_t* img; int pitch;
  if (lv == 0) { img = D.l0_ptr[s]; pitch = D.l0_stride[s]; } else { img = L.img + (size_t)s * L.h * L.pitch; pitch = L.pitch; }
  const int lw = L.w, lh = L.h;
  const int maxSSD = PP * 500;
  const int nLevelScale = LevelScale(lv);
  const double invScale = 1.0 / nLevelScale;                    // 2^-level: x / 2^l == x * 2^-l exactly
  const double ix = v2i0 * invScale, iy = v2i1 * invScale;
  const unsigned nRange = ((unsigned)pl.range + nLevelScale - 1) / nLevelScale;
  int nTop = iy - nRange;
  const int nBottomPlusOne = iy + nRange + 1;
  const int nLeft = ix - nRange, nRight = ix + nRange;
  const double r2 = (double)(nRange * nRange);
  unsigned long long best = ((unsigned long long)(unsigned)(maxSSD + 1) << 32) | 0xffffffffull;
  flags |= F_SEARCHED;
  if (nTop < 0) nTop = 0;
  // The window's corners straight from the level's corner bitmask (LevelDesc::cbits, one word per 32 pixels, written by the FAST kernels):
  // rows [nTop, nBot), columns [nLeft, nRight] clipped to the image (the x range of jni/PatchFinder.cc:214-215, exact); a set bit IS a
  // corner position, so neither the corner list nor the row LUT is read.
  const int nBot = nBottomPlusOne < lh ? nBottomPlusOne : lh;
  const bool x_empty = nRight < 0 || nLeft > lw - 1;
  const int xl0 = nLeft > 0 ? nLeft : 0, xr0 = nRight < lw - 1 ? nRight : lw - 1;
  const int wr = xr0 >> 5, cwpr = (lw + 31) >> 5;
  const uint32_t* cb = L.cbits + (size_t)s * lh * cwpr;
  const bool more = alive && !x_empty && nTop < nBot;   // (nTop >= rows or nBottomPlusOne <= 0: nothing to search, jni/PatchFinder.cc:189-195)
  const int b = P / 2, nwords = (P + 3) >> 2;
  const uint32_t lastmask = (P & 3) ? ((1u << (8 * (P & 3))) - 1u) : 0xffffffffu;
  int nevals = 0, qn = 0;
  // The window is walked in batches of 24 rows x 32 columns: lane j fetches rows j, j + 8, j + 16 of the batch at once (six independent
  // loads, one round trip) and cuts each down to a 32-bit mask whose bit k stands for column x0b + k.  The tracker's windows (ranges up to
  // 11 pixels of the search level) and MapMaker's (4) are ONE batch; wide windows (the roofline sweep: range 40 = 81 x 81) take a few.
  const int ncb = more ? ((xr0 - xl0) >> 5) + 1 : 0;                                    // column blocks of 32
  const int nb = more ? ((nBot - nTop + 3 * kGL - 1) / (3 * kGL)) * ncb : 0;            // batches of this entry
  int bi = 0, x0b = 0, y0b = 0;
  uint32_t rm0 = 0u, rm1 = 0u, rm2 = 0u;
  while (true) {
    // -- scan: until the warp's entries have run out of corners or one of the queues could overflow in the next step
    while (!__any_sync(0xffffffffu, qn > kQ - kGL)) {
      const uint32_t left = rm0 | rm1 | rm2;
      if (!__any_sync(0xffffffffu, left != 0u)) {
        if (!__any_sync(0xffffffffu, bi < nb)) break;
        if (bi < nb) {
          const int rb = bi / ncb, cbk = bi - rb * ncb;
          y0b = nTop + 3 * kGL * rb; x0b = xl0 + 32 * cbk;
          const int w0 = x0b >> 5, width = xr0 - x0b + 1;
          uint32_t lo[3], hi[3];
#pragma unroll
          for (int r = 0; r < 3; r++) {
            const int y = y0b + j + kGL * r;
            lo[r] = 0u; hi[r] = 0u;
            if (y < nBot) { const uint32_t* row = cb + (size_t)y * cwpr; lo[r] = __ldg(row + w0); if (w0 < wr) hi[r] = __ldg(row + w0 + 1); }
          }
          const uint32_t wmask = width >= 32 ? 0xffffffffu : ((1u << width) - 1u);
          rm0 = __funnelshift_r(lo[0], hi[0], x0b & 31) & wmask; rm1 = __funnelshift_r(lo[1], hi[1], x0b & 31) & wmask; rm2 = __funnelshift_r(lo[2], hi[2], x0b & 31) & wmask;
          bi++;
        }
        continue;
      }
      // one corner per lane and step, from the first of the lane's rows that still has one: the circle test of jni/PatchFinder.cc:216-219,
      // survivors to the entry's queue
      bool pass = left != 0u;
      const int r = rm0 ? 0 : (rm1 ? 1 : 2);
      const uint32_t mm = rm0 ? rm0 : (rm1 ? rm1 : rm2);
      const int bit = pass ? __ffs(mm) - 1 : 0;
      if (rm0) rm0 &= rm0 - 1u; else if (rm1) rm1 &= rm1 - 1u; else rm2 &= rm2 - 1u;
      const int cx = x0b + bit, cy = y0b + j + kGL * r;
      if (pass) { const double dx = ix - (double)cx, dy = iy - (double)cy; double d2 = 0; d2 += dx * dx; d2 += dy * dy; pass = !(d2 > r2); }
      const unsigned gb = (__ballot_sync(0xffffffffu, pass) >> (kGL * g)) & ((1u << kGL) - 1u);
      if (pass) W.q.cw[g][qn + __popc(gb & ((1u << j) - 1u))] = ((uint32_t)cy << 16) | (uint32_t)cx;
      qn += __popc(gb);
    }
    if (!__any_sync(0xffffffffu, qn > 0)) break;
    __syncwarp();
    // -- ZMSSDAtPoint (jni/PatchFinder.cc:352-380) of the queued candidates: one lane per candidate, rounds of eight per entry
    for (int q0 = 0; __any_sync(0xffffffffu, q0 < qn); q0 += kGL) {
      const bool have = q0 + j < qn;
      const uint32_t cw = W.q.cw[g][have ? q0 + j : 0];
      const int cx = cw & 0xffff, cy = cw >> 16;
      const bool inb = have && (cx >= b && cy >= b && cx < lw - b && cy < lh - b);
      int ssd = maxSSD + 1;
      if ((pitch & 3) == 0) {
        // Rows a multiple of 4 bytes apart (every level image of the context; caller-owned level-0 frames with such a stride): one alignment
        // for the whole window.  The words of kNB rows are requested back to back, with a warp barrier between the requests and their first
        // use: left alone, ptxas sinks each row's loads between the dp4a of the previous row (fewer live registers), which turns a window
        // into P dependent round trips to L1 / L2.  Lanes without a (valid) candidate read the image origin and discard the result.
        const uint8_t* rp = inb ? img + (size_t)(cy - b) * pitch + (cx - b) : img;
        const unsigned a = (unsigned)((uintptr_t)rp & 3u), shf = a * 8;
        const uint32_t* wp = (const uint32_t*)(rp - a);
        const int pw = inb ? pitch >> 2 : 0;
        const bool need2 = a + P > 8, need3 = a + P > 12;
        constexpr int NR = PT ? PT : VS_MAXP;
        unsigned sum = 0, sumsq = 0, cross = 0;
#pragma unroll
        for (int r0 = 0; r0 < NR; r0 += kNB) {
          uint32_t w[kNB][4];
#pragma unroll
          for (int u = 0; u < kNB; u++) {
            const int r = r0 + u;
            w[u][0] = w[u][1] = w[u][2] = w[u][3] = 0u;
            if (r < NR && (PT != 0 || r < P)) {
              const uint32_t* q = wp + r * pw;
              w[u][0] = ldg_ordered(q); w[u][1] = ldg_ordered(q + 1);
              if (need2) w[u][2] = ldg_ordered(q + 2);
              if (need3) w[u][3] = ldg_ordered(q + 3);
            }
          }
          __syncwarp();
#pragma unroll
          for (int u = 0; u < kNB; u++) {
            const int r = r0 + u;
            if (r < NR && (PT != 0 || r < P)) {
              uint32_t n0 = __funnelshift_r(w[u][0], w[u][1], shf), n1 = __funnelshift_r(w[u][1], w[u][2], shf), n2 = __funnelshift_r(w[u][2], w[u][3], shf);
              if (nwords == 3) n2 &= lastmask; else if (nwords == 2) { n1 &= lastmask; n2 = 0; } else { n0 &= lastmask; n1 = 0; n2 = 0; }
              sum = __dp4a(n0, 0x01010101u, sum); sumsq = __dp4a(n0, n0, sumsq); cross = __dp4a(n0, tw[3 * r], cross);
              if (nwords > 1) { sum = __dp4a(n1, 0x01010101u, sum); sumsq = __dp4a(n1, n1, sumsq); cross = __dp4a(n1, tw[3 * r + 1], cross); }
              if (nwords > 2) { sum = __dp4a(n2, 0x01010101u, sum); sumsq = __dp4a(n2, n2, sumsq); cross = __dp4a(n2, tw[3 * r + 2], cross); }
            }
          }
        }
        if (inb) { const int SA = tsum, SB = (int)sum; ssd = ((2 * SA * SB - SA * SA - SB * SB) / PP + (int)sumsq + tsumsq - 2 * (int)cross); }
      } else if (inb) {
        unsigned sum = 0, sumsq = 0, cross = 0;
        const uint8_t* rp = img + (size_t)(cy - b) * pitch + (cx - b);
#pragma unroll 1
        for (int r = 0; r < P; r++) {
          const unsigned a = (unsigned)((uintptr_t)rp & 3u), shf = a * 8;
          const uint32_t* wp = (const uint32_t*)(rp - a);
          const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1);
          const uint32_t w2 = (a + P > 8) ? __ldg(wp + 2) : 0u, w3 = (a + P > 12) ? __ldg(wp + 3) : 0u;
          uint32_t n0 = __funnelshift_r(w0, w1, shf), n1 = __funnelshift_r(w1, w2, shf), n2 = __funnelshift_r(w2, w3, shf);
          if (nwords == 3) n2 &= lastmask; else if (nwords == 2) { n1 &= lastmask; n2 = 0; } else { n0 &= lastmask; n1 = 0; n2 = 0; }
          sum = __dp4a(n0, 0x01010101u, sum); sumsq = __dp4a(n0, n0, sumsq); cross = __dp4a(n0, tw[3 * r], cross);
          if (nwords > 1) { sum = __dp4a(n1, 0x01010101u, sum); sumsq = __dp4a(n1, n1, sumsq); cross = __dp4a(n1, tw[3 * r + 1], cross); }
          if (nwords > 2) { sum = __dp4a(n2, 0x01010101u, sum); sumsq = __dp4a(n2, n2, sumsq); cross = __dp4a(n2, tw[3 * r + 2], cross); }
          rp += pitch;
        }
        const int SA = tsum, SB = (int)sum;
        ssd = ((2 * SA * SB - SA * SA - SB * SB) / PP + (int)sumsq + tsumsq - 2 * (int)cross);
      }
      if (have) {
        const unsigned long long key = ((unsigned long long)(unsigned)ssd << 32) | cw;   // ssd >= 0; ties -> first corner in raster order (y << 16 | x)
        if (key < best) best = key;
      }
      __syncwarp();
    }
    nevals += qn; qn = 0;
    __syncwarp();
  }
#pragma unroll
  for (int d = kGL / 2; d; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
    if (o < best) best = o;
  }
  if (!alive) return;
  if (j == 0 && nevals) atomicAdd(D.evals, (unsigned long long)nevals);
  const int bestSSD = (int)(best >> 32);
  if (!(bestSSD < maxSSD)) {
    if (j == 0) D.ps.flags[gi] = flags & ~F_FOUND;
    return;
  }
  if (j == 0) {
    const uint32_t bc = (uint32_t)best;
    const double coarse0 = ((double)(bc & 0xffff) + 0.5) * nLevelScale - 0.5, coarse1 = ((double)(bc >> 16) + 0.5) * nLevelScale - 0.5;  // LevelZeroPos
    flags |= F_FOUND;
    D.ps.coarse[gi] = coarse0; D.ps.coarse[SN + gi] = coarse1;
    D.ps.sqrtinv[gi] = invScale;
    if (subpix > 0) flags |= F_SUBPIX;          // k_subpix refines it, counts it as found if the iteration converges and un-finds it otherwise
    else { flags &= ~F_SUBPIX; D.ps.v2found[gi] = coarse0; D.ps.v2found[SN + gi] = coarse1; atomicAdd(&st->found[level], 1); }
    D.ps.flags[gi] = flags;
  }
}

// MakeSubPixTemplate + IterateSubPixToConvergence (jni/PatchFinder.cc:242-350; jni/Tracker.cc:657-667) of the entries that k_search_fast
// found and marked for refinement: EIGHT LANES PER ENTRY, four entries per warp.  The iteration is a chain -- the three sums of an
// iteration are added in pixel order like the reference's serial loop, 81 dependent additions each, by three lanes -- so what a warp can
// do is run four such chains side by side; the bilinear samples and products of an iteration are spread over the entry's eight lanes.
// The template gradients are recomputed from the template bytes where they are needed (exact: multiples of 0.5).
constexpr int kSL = 8;            // k_subpix: lanes per entry
constexpr int kSW = 32 / kSL;     // k_subpix: entries per warp
constexpr int kWin = VS_MAXP + 5;   // image window staged per entry: the P - 1 pixels an iteration samples per axis + 3 pixels of travel either way
struct SubpixEntry { double prod[3][(VS_MAXP - 2) * (VS_MAXP - 2)]; uint32_t tmpl_w[VS_TMPL_BYTES / 4]; uint8_t win[kWin * kWin]; };   // dDiff*gx, dDiff*gy, dDiff per interior pixel; the template; the window

__global__ void __launch_bounds__(kFW * 32) k_subpix(Dev D, int mode, int range_arg, int subpix_arg, int sflags) {
  cudaGridDependencySynchronize(); cudaTriggerProgrammaticLaunchCompletion();
  __shared__ SubpixEntry sm_all[kFW][kSW];
  const int s = blockIdx.y + D.s0, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane / kSL, j = lane % kSL;
  StreamState* st = D.ss + s;
  Plan pl;
  if (!search_plan(D, st, mode, range_arg, subpix_arg, pl)) return;
  const int n_sub = pl.subpix_all > 0 ? pl.count : (pl.subpix_top > 0 ? min(pl.n_top, pl.count) : 0);
  const bool refind = (sflags & kSearchRefindF) != 0;
  const size_t SN = (size_t)D.S * D.N;
  const int P = D.P, Q = P - 2, QQ = Q * Q;
  const int rq = (65536 + Q - 1) / Q;                            // k / Q == (k * rq) >> 16 for k < Q * Q <= 81
  SubpixEntry& E = sm_all[warp][g];
  const uint8_t* const tmpl = (const uint8_t*)E.tmpl_w;
  for (int e0 = (blockIdx.x * kFW + warp) * kSW; e0 < n_sub; e0 += gridDim.x * kFW * kSW) {   // (the grid is a few CTAs per stream: the fine stage refines only its top-level entries)
    const int e = e0 + g;
    bool active = e < n_sub;
    const int i = D.lists[(size_t)s * D.list_cap + pl.first + (active ? e : e0)];
    const size_t gi = (size_t)s * D.N + i;
    int flags = D.ps.flags[gi];
    const int subpix = e < pl.n_top ? pl.subpix_top : pl.subpix_all;
    const int level = refind ? D.ps.rlevel[gi] : D.ps.level[gi];
    active = active && (flags & (F_FOUND | F_SUBPIX | F_SEARCHED)) == (F_FOUND | F_SUBPIX | F_SEARCHED) && subpix > 0 && !(refind && level == 0);
    const bool entry = active;                                   // this group has an entry to finish
    const int lv = active ? level : 0;
    const uint32_t* tmpl_g = (const uint32_t*)(D.ps.tmpl + gi * VS_TMPL_BYTES);
#pragma unroll
    for (int q = 0; q < 5; q++) if (j + kSL * q < VS_TMPL_BYTES / 4) E.tmpl_w[j + kSL * q] = tmpl_g[j + kSL * q];
    const double coarse0 = D.ps.coarse[gi], coarse1 = D.ps.coarse[SN + gi];
    const LevelDesc& L = D.lev[lv];
    const uint8_t* img; int pitch;
    if (lv == 0) { img = D.l0_ptr[s]; pitch = D.l0_stride[s]; } else { img = L.img + (size_t)s * L.h * L.pitch; pitch = L.pitch; }
    const int lw = L.w, lh = L.h;
    const int nLevelScale = LevelScale(lv);
    const double invScale = 1.0 / nLevelScale;
    __syncwarp();
    if (entry && j == 0) atomicAdd(D.evals + 3, 1ull);
    // ---- MakeSubPixTemplate (jni/PatchFinder.cc:242-267): JtJ of (gx, gy, 1); sums of multiples of 0.25 below 2^53 are exact in any order
    double hxx = 0, hxy = 0, hyy = 0, hx = 0, hy = 0;
    for (int k = j; k < QQ; k += kSL) {
      const int x = ((k * rq) >> 16) + 1, y = k - (x - 1) * Q + 1;
      const double gx = 0.5 * (tmpl[y * 12 + x + 1] - tmpl[y * 12 + x - 1]), gy = 0.5 * (tmpl[(y + 1) * 12 + x] - tmpl[(y - 1) * 12 + x]);
      hxx += gx * gx; hxy += gx * gy; hyy += gy * gy; hx += gx; hy += gy;
    }
    __syncwarp();
#pragma unroll
    for (int d = kSL / 2; d; d >>= 1) {
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
    // The iterations sample the image around a position that moves by fractions of a pixel: the kWin x kWin pixels around the coarse hit are
    // staged in shared memory once (32 independent loads per lane) instead of four dependent global loads per pixel and iteration.
    int wx0 = 0, wy0 = 0;
    if (active) {
      const double c0 = (coarse0 + 0.5) * invScale - 0.5, c1 = (coarse1 + 0.5) * invScale - 0.5;
      wx0 = (int)(c0 - (double)(P / 2)) - 2; wy0 = (int)(c1 - (double)(P / 2)) - 2;
      for (int k = j; k < kWin * kWin; k += kSL) {
        const int r = k / kWin, c = k - r * kWin;
        int yy = wy0 + r, xx = wx0 + c;
        yy = yy < 0 ? 0 : (yy >= lh ? lh - 1 : yy); xx = xx < 0 ? 0 : (xx >= lw ? lw - 1 : xx);   // (clamped pixels are never sampled: the border test below)
        E.win[k] = img[(size_t)yy * pitch + xx];
      }
    }
    __syncwarp();
    double sp0 = coarse0, sp1 = coarse1, meanDiff = 0.0;
    int ok = 0, it = 0;
    // ---- IterateSubPixToConvergence / IterateSubPix (jni/PatchFinder.cc:272-350); the four entries of the warp iterate in step
    while (__any_sync(0xffffffffu, active)) {
      if (active && it >= subpix) active = false;                  // iteration budget used up: not converged
      double b0 = 0, b1 = 0; float fTL = 0, fTR = 0, fBL = 0, fBR = 0;
      if (active) {
        const double c0 = (sp0 + 0.5) * invScale - 0.5, c1 = (sp1 + 0.5) * invScale - 0.5;   // LevelNPos
        const int xb = (c0 > 0.0 ? c0 + 0.5 : c0 - 0.5), yb = (c1 > 0.0 ? c1 + 0.5 : c1 - 0.5);
        const int bd = P / 2 + 1;
        if (!(xb >= bd && yb >= bd && xb < lw - bd && yb < lh - bd)) active = false;   // off the image: not converged
        else {
          b0 = c0 - (double)(P / 2); b1 = c1 - (double)(P / 2);
          const double dX = b0 - floor(b0), dY = b1 - floor(b1);
          fTL = (1.0 - dX) * (1.0 - dY); fTR = (dX) * (1.0 - dY); fBL = (1.0 - dX) * (dY); fBR = (dX) * (dY);
        }
      }
      if (active) {
        const int ox = (int)b0 - wx0, oy = (int)b1 - wy0;             // samples: columns ox + 1 .. ox + P - 1 of the window, rows likewise
        const bool inwin = ox >= -1 && oy >= -1 && ox + P - 1 < kWin && oy + P - 1 < kWin;
        const uint8_t* base = inwin ? E.win + oy * kWin + ox : img + (size_t)(int)b1 * pitch + (int)b0;
        const int bp = inwin ? kWin : pitch;
        for (int k = j; k < QQ; k += kSL) {   // k = (y-1)*Q + (x-1): the reference's loop order
          const int y = ((k * rq) >> 16) + 1, x = k - (y - 1) * Q + 1;
          const uint8_t* tl = base + y * bp + x;
          const float fPixel = fTL * tl[0] + fTR * tl[1] + fBL * tl[bp] + fBR * tl[bp + 1];
          const double dDiff = fPixel - tmpl[y * 12 + x] + meanDiff;
          const double gx = 0.5 * (tmpl[y * 12 + x + 1] - tmpl[y * 12 + x - 1]), gy = 0.5 * (tmpl[(y + 1) * 12 + x] - tmpl[(y - 1) * 12 + x]);
          E.prod[0][k] = dDiff * gx; E.prod[1][k] = dDiff * gy; E.prod[2][k] = dDiff;
        }
      }
      __syncwarp();
      double acc = 0;   // lanes 0,1,2 of the entry add their accumulator's terms in pixel order, like the reference's serial loop
      if (active && j < 3) {   // (nine terms are fetched before they are added: the chain is 81 additions, not 81 x (shared-memory load + addition))
        const double* p = E.prod[j];
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
      __syncwarp();
      const double a0 = __shfl_sync(0xffffffffu, acc, 0, kSL), a1 = __shfl_sync(0xffffffffu, acc, 1, kSL), a2 = __shfl_sync(0xffffffffu, acc, 2, kSL);
      if (active) {
        double upd[3];
#pragma unroll
        for (int r = 0; r < 3; r++) { double sacc = hinv[3 * r] * a0; sacc += hinv[3 * r + 1] * a1; sacc += hinv[3 * r + 2] * a2; upd[r] = sacc; }
        sp0 -= upd[0] * nLevelScale; sp1 -= upd[1] * nLevelScale;
        meanDiff -= upd[2];
        double d = 0; d += upd[0] * upd[0]; d += upd[1] * upd[1];
        const double lim = 0.03;
        if (d < lim * lim) { ok = 1; active = false; }
        it++;
      }
    }
    if (entry && j == 0) {
      if (ok || refind) { D.ps.v2found[gi] = sp0; D.ps.v2found[SN + gi] = sp1; atomicAdd(&st->found[level], 1); }
      else { flags &= ~F_FOUND; D.ps.flags[gi] = flags; }   // sub-pixel iteration did not converge (jni/Tracker.cc:660-666)
    }
    __syncwarp();
  }
}

}  // namespace

int vs_launch_search_fast(vslam_ctx* ctx, int which, int range, int subpix, int sflags) {
  const Dev D = make_dev(ctx);
  const int max_entries = which == 1 ? (int)(2 * ctx->params.coarse_max) : ctx->list_cap;
  if (max_entries <= 0) return VSLAM_OK;   // nCoarseMax == 0: the reference skips the coarse stage (jni/Tracker.cc:425)
  const int per_cta = kFW * kGW;
  dim3 grid((max_entries + per_cta - 1) / per_cta, ctx->cur_cnt);
  vs_time_begin(ctx, which == 2 ? VS_ST_SEARCH_FINE : VS_ST_SEARCH_COARSE);
  const bool pdl = ctx->pdl && !ctx->timing;
  if (ctx->P == 11) VS_CUDA(vs_launch_pdl(k_search_fast<11>, grid, dim3(kFW * 32), 0, ctx->stream, pdl, D, which, range, subpix, sflags));
  else if (ctx->P == 8) VS_CUDA(vs_launch_pdl(k_search_fast<8>, grid, dim3(kFW * 32), 0, ctx->stream, pdl, D, which, range, subpix, sflags));
  else VS_CUDA(vs_launch_pdl(k_search_fast<0>, grid, dim3(kFW * 32), 0, ctx->stream, pdl, D, which, range, subpix, sflags));
  ctx->launches++;
  // sub-pixel refinement: every entry of the coarse stage / of an explicit list with subpix > 0, the top-level entries of the fine stage
  const bool any_subpix = which == 0 ? subpix > 0 : (which == 1 ? ctx->params.coarse_subpix_its > 0 : ctx->params.fine_subpix_its_top_level > 0);
  if (any_subpix && !(sflags & 2)) {
    dim3 g2(std::min((max_entries + kFW * kSW - 1) / (kFW * kSW), 16), ctx->cur_cnt);
    VS_CUDA(vs_launch_pdl(k_subpix, g2, dim3(kFW * 32), 0, ctx->stream, pdl, D, which, range, subpix, sflags));
    ctx->launches++;
  }
  vs_time_end(ctx);
  VS_CUDA(cudaGetLastError());
  return VSLAM_OK;
}
