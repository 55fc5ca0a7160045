// KeyFrame::MakeKeyFrame_Lite on the GPU (reference: jni/KeyFrame.cc:5-51).
//
// Batched over all streams: k_pyramid_fast handles level 0 (and writes the images of levels 1..3), k_fast_levels handles levels 1..3,
// k_corner_count + k_corner_lists turn the corner bitmasks of all four levels into the corner lists and row tables.
// A CTA of the first two owns a strip of LevelDesc::strip_rows rows of one level of one stream.  It
//   1. stages the strip plus a 3-row halo into shared memory with ONE bulk async copy (cp.async.bulk, TMA engine);
//   2. (level 0) writes the matching rows of level 1:  (a+b+c+d+2)>>2  (cv::resize 2:1, jni/KeyFrame.cc:20-23; SURVEY.md
//      F2), and, from the same staged rows, the matching rows of levels 2 and 3;
//   3. runs FAST-10 (jni/vision/cvfast.cpp:6088-9241 == segment test), 16 pixels x 2 rows per lane:
//        a byte-SIMD rejection test (VABSDIFF4 + SWAR compares): a 10-arc contains at least one pixel of every opposite ring pair,
//        so both (0,8) and (4,12) must hold a pixel differing by more than t;
//        survivors are compacted into a per-warp queue and, 32 at a time, get the test on the eight even ring positions (five
//        consecutive of one polarity are necessary) and then the exact 16-pixel ring test, one lane per candidate;
//   4. sets the bit of every corner in the level's corner bitmask in global memory (one word per 32 pixels, cleared per frame)
//      and exits: no barrier after the staging, no inter-CTA dependency.
// k_corner_lists (one CTA per chunk of at most 4096 bitmask words): popcounts + one block scan give every corner its raster-order position; the running
// positions at the row starts ARE the row look-up table of jni/KeyFrame.cc:41-49.  (The tracker's patch search reads the bitmasks.)
// Integer / byte arithmetic only; results are bit-exact against the oracle.
#include "vslam_internal.cuh"
#include <cstdio>
#include <cstdlib>

#ifndef VS_PYR_REGS
#define VS_PYR_REGS 48   // registers per thread of the two FAST kernels (four 320-thread CTAs per SM; 40 / five CTAs measured 3 % slower, 32 / six CTAs 6 %)
#endif

namespace {

constexpr int kMaxThreads = 320;           // ten warps: a 16-row VGA strip is exactly ten work items (32 lanes x 16 pixels x 2 rows)
constexpr int kMaxWarps = kMaxThreads / 32;
constexpr int kQueue = 288;                // per-warp candidate queue: up to 31 carried over + up to 255 of one work item (denser items bypass the queue)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_%=;\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// per-byte (x > t) for 4 packed bytes, t < 128: flag in bit 7 of every byte (other bits are garbage)
__device__ __forceinline__ uint32_t bytes_gt(uint32_t x, uint32_t k /* (127 - t) * 0x01010101 */) { return ((x & 0x7f7f7f7fu) + k) | x; }

// Exact FAST-10 segment test of one candidate (queue entry = smem row << 16 | x); sets its bit in the level's corner bitmask.
// One IMAD per ring pixel yields both polarities: with kc = ((c - t - 1 + 2^15) << 16) + (2^15 - c - t - 1),
//   z = v * (1 - 2^16) + kc   has   bit 15 = (v - c - t - 1 >= 0) = brighter,   bit 31 = (c - t - 1 - v >= 0) = darker
// (both 16-bit fields stay inside [2^15 - 266, 2^15 + 254]: no borrow between them).
__device__ __forceinline__ void ring_test(uint32_t e, const uint8_t* img, int stride, int thr, uint32_t* gbits /* word of (staged row 0, x = 0) */, int words_per_row) {
  const int ry = e >> 16, x = e & 0xffff;
  const uint8_t* p0 = img + ry * stride + x;
  const uint8_t *pp1 = p0 + stride, *pp2 = pp1 + stride, *pp3 = pp2 + stride, *pm1 = p0 - stride, *pm2 = pm1 - stride, *pm3 = pm2 - stride;
  const uint32_t c = p0[0];
  const uint32_t kc = ((c + (0x8000u - 1u - (uint32_t)thr)) << 16) + ((0x8000u - 1u - (uint32_t)thr) - c);
  uint32_t a0 = 0, a1 = 0;   // ring positions 0..7 / 8..15: brighter flags end up in bits 8..15, darker flags in bits 24..31
#define RING(acc, ptr, off)                                             \
  {                                                                     \
    const uint32_t z = (uint32_t)(ptr)[(off)] * 0xFFFF0001u + kc;       \
    acc = (acc >> 1) | (z & 0x80008000u);                               \
  }
  RING(a0, pp3, 0) RING(a0, pp3, 1) RING(a0, pp2, 2) RING(a0, pp1, 3) RING(a0, p0, 3) RING(a0, pm1, 3) RING(a0, pm2, 2) RING(a0, pm3, 1)
  RING(a1, pm3, 0) RING(a1, pm3, -1) RING(a1, pm2, -2) RING(a1, pm1, -3) RING(a1, p0, -3) RING(a1, pp1, -3) RING(a1, pp2, -2) RING(a1, pp3, -1)
#undef RING
  uint32_t br = ((a0 >> 8) & 0xffu) | (a1 & 0xff00u), dk = (a0 >> 24) | ((a1 >> 16) & 0xff00u);
  br |= br << 16; dk |= dk << 16;
  uint32_t b1 = br & (br >> 1), d1 = dk & (dk >> 1);
  uint32_t b2 = b1 & (b1 >> 2), d2 = d1 & (d1 >> 2);
  b2 &= b2 >> 4; d2 &= d2 >> 4;
  b2 &= b1 >> 8; d2 &= d1 >> 8;
  if ((b2 | d2) & 0xffffu) atomicOr(&gbits[ry * words_per_row + (x >> 5)], 1u << (x & 31));   // result unused: a RED
}

// First stage of the exact test, on the eight even ring positions only: a 10-arc of the 16-ring covers five consecutive even positions,
// so a corner has five consecutive (circular, of eight) even-position pixels all brighter or all darker.  About one in five survivors
// of reject16 passes (level 0); only those get the full ring test.
__device__ __forceinline__ bool ring_even(uint32_t e, const uint8_t* img, int stride, int thr) {
  const int ry = e >> 16, x = e & 0xffff;
  const uint8_t* p0 = img + ry * stride + x;
  const uint8_t *pp2 = p0 + 2 * stride, *pp3 = pp2 + stride, *pm2 = p0 - 2 * stride, *pm3 = pm2 - stride;
  const uint32_t c = p0[0];
  const uint32_t kc = ((c + (0x8000u - 1u - (uint32_t)thr)) << 16) + ((0x8000u - 1u - (uint32_t)thr) - c);
  uint32_t a = 0;   // after eight steps: brighter flags of positions 0,2,..,14 in bits 8..15, darker flags in bits 24..31
#define RING(ptr, off)                                                  \
  {                                                                     \
    const uint32_t z = (uint32_t)(ptr)[(off)] * 0xFFFF0001u + kc;       \
    a = (a >> 1) | (z & 0x80008000u);                                   \
  }
  RING(pp3, 0) RING(pp2, 2) RING(p0, 3) RING(pm2, 2) RING(pm3, 0) RING(pm2, -2) RING(p0, -3) RING(pp2, -2)
#undef RING
  const uint32_t v = a | (a >> 8);            // each byte doubled: bits 0..15 brighter (circular), 16..31 darker; starts k = 0..7 read bits k..k+4 <= 11 only
  const uint32_t r2 = v & (v >> 1), r4 = r2 & (r2 >> 2);
  return (r4 & (v >> 4) & 0x00ff00ffu) != 0u;
}

// Rejection test of 16 pixels of one row (centre words c, staged row pointer r = &row[x0]): a 10-arc of the 16-ring contains at
// least one pixel of each opposite pair, so both (0,8) [rows -3 / +3] and (4,12) [columns -3 / +3] must hold a pixel differing
// from the centre by more than t.  Returns bit i = pixel x0 + i survives.
__device__ __forceinline__ uint32_t reject16(const uint8_t* r, int stride, const uint4 c, uint32_t kcmp) {
  const uint4 up = *(const uint4*)(r - 3 * stride), dn = *(const uint4*)(r + 3 * stride);
  const uint32_t cw[6] = {*(const uint32_t*)(r - 4), c.x, c.y, c.z, c.w, *(const uint32_t*)(r + 16)};
  const uint32_t uw[4] = {up.x, up.y, up.z, up.w}, dw[4] = {dn.x, dn.y, dn.z, dn.w};
  uint32_t bits = 0;
#pragma unroll
  for (int j = 3; j >= 0; j--) {
    const uint32_t cc = cw[j + 1];
    const uint32_t lf = __byte_perm(cw[j], cc, 0x4321), rt = __byte_perm(cc, cw[j + 2], 0x6543);
    const uint32_t m1 = bytes_gt(__vabsdiffu4(uw[j], cc), kcmp) | bytes_gt(__vabsdiffu4(dw[j], cc), kcmp);
    const uint32_t m2 = bytes_gt(__vabsdiffu4(lf, cc), kcmp) | bytes_gt(__vabsdiffu4(rt, cc), kcmp);
    // bit 7 of every byte -> the top nibble of the product (no two partial products meet), pushed into `bits` by a funnel shift
    bits = __funnelshift_l((m1 & m2 & 0x80808080u) * 0x00204081u, bits, 4);
  }
  return bits;
}


struct StripShared {
  uint64_t bar;
  int ticket, next_item;
};

// n / d for small n (n * d < 2^32) with the reciprocal m = 0xffffffff / d + 1 from the host (LevelDesc::mg_*; m == 0 stands for d == 1):
// one IMAD.HI instead of an integer division
__device__ __forceinline__ int div_small(int n, uint32_t m) { return m == 0u ? n : (int)__umulhi((uint32_t)n, m); }

// Levels 2 and 3 of a strip from its level-1 rows kept in shared memory (l1s: rows1 rows, w1 bytes apart, rows1 % 4 == 0 or
// rows1 == 4): cv::resize 2:1 once / twice more, every intermediate pixel rounded as the level-by-level computation
// rounds it, (a+b+c+d+2)>>2 (jni/KeyFrame.cc:20-23).  Threads t of nt.
__device__ __forceinline__ void emit_levels_2_3(const uint8_t* l1s, int w1, int rows1, const LevelDesc& L2, const LevelDesc& L3, int s, int y0, int t, int nt) {
  uint8_t* d2 = L2.img + ((size_t)s * L2.h + (y0 >> 2)) * L2.pitch;
  uint8_t* d3 = L3.img + ((size_t)s * L3.h + (y0 >> 3)) * L3.pitch;
  const int p2 = L2.w >> 1, n2 = (rows1 >> 1) * p2;       // level 2: two pixels per step
  for (int i = t; i < n2; i += nt) {
    const int r = div_small(i, L2.mg_hw), xp = i - r * p2;
    const uint32_t c0 = *(const uint32_t*)(l1s + (2 * r) * w1 + 4 * xp), c1 = *(const uint32_t*)(l1s + (2 * r + 1) * w1 + 4 * xp);
    const unsigned o0 = __dp4a(c0, 0x00000101u, __dp4a(c1, 0x00000101u, 2u)) >> 2;
    const unsigned o1 = __dp4a(c0, 0x01010000u, __dp4a(c1, 0x01010000u, 2u)) >> 2;
    *(uint16_t*)(d2 + (size_t)r * L2.pitch + 2 * xp) = (uint16_t)(o0 | (o1 << 8));
  }
  const int w3 = L3.w, n3 = (rows1 >> 2) * w3;            // level 3: one pixel (a 4x4 block of level 1) per step
  for (int i = t; i < n3; i += nt) {
    const int r = div_small(i, L3.mg_w), x = i - r * w3;
    const uint8_t* p = l1s + (4 * r) * w1 + 4 * x;
    const uint32_t q0 = *(const uint32_t*)p, q1 = *(const uint32_t*)(p + w1), q2 = *(const uint32_t*)(p + 2 * w1), q3 = *(const uint32_t*)(p + 3 * w1);
    const unsigned a = __dp4a(q0, 0x00000101u, __dp4a(q1, 0x00000101u, 2u)) >> 2, b = __dp4a(q0, 0x01010000u, __dp4a(q1, 0x01010000u, 2u)) >> 2;
    const unsigned c = __dp4a(q2, 0x00000101u, __dp4a(q3, 0x00000101u, 2u)) >> 2, d = __dp4a(q2, 0x01010000u, __dp4a(q3, 0x01010000u, 2u)) >> 2;
    d3[(size_t)r * L3.pitch + x] = (uint8_t)((a + b + c + d + 2u) >> 2);
  }
}

// four level-1 pixels from two words of row y and two words of row y + 1: (a+b+c+d+2)>>2 each
__device__ __forceinline__ uint32_t half4(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
  const uint32_t o0 = __dp4a(a0, 0x00000101u, __dp4a(b0, 0x00000101u, 2u)), o1 = __dp4a(a0, 0x01010000u, __dp4a(b0, 0x01010000u, 2u));
  const uint32_t o2 = __dp4a(a1, 0x00000101u, __dp4a(b1, 0x00000101u, 2u)), o3 = __dp4a(a1, 0x01010000u, __dp4a(b1, 0x01010000u, 2u));
  const uint32_t lo = ((o0 + (o1 << 16)) >> 2) & 0x00ff00ffu, hi = ((o2 + (o3 << 16)) >> 2) & 0x00ff00ffu;   // sums < 1024: the 16-bit halves do not meet
  return __byte_perm(lo, hi, 0x6420);
}

// One strip (rows [strip*R, +R) of level L of stream s): stage, [level 0: write levels 1..3], FAST-10 into the corner bitmask.
template <bool kLevel0>
__device__ __forceinline__ void fast_strip(const LevelDesc& L, const uint8_t* __restrict__ src, const int stride, const int s, const int strip, const int thr,
                                           const LevelDesc* Lchild /* [3] levels 1..3, level 0 only */, StripShared& sh, uint8_t* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nw = nthreads >> 5;
  const int W = L.w, H = L.h, R = L.strip_rows;
  const int y0 = strip * R, y1 = min(y0 + R, H);
  const int ya = max(y0 - 3, 0), yb = min(y1 + 3, H);
  const int words_per_row = (W + 31) >> 5;
  uint8_t* img = smem;                                                  // rows ya..yb-1, `stride` bytes apart
  uint32_t* qbase = (uint32_t*)(smem + (size_t)(R + 6) * stride);
  uint32_t* queue = qbase + warp * kQueue;
  uint32_t* gbits = L.cbits + ((size_t)s * H + ya) * words_per_row;     // corner bitmask word of (row ya, x = 0)

  if (tid == 0) {
    const uint32_t bytes = (uint32_t)(yb - ya) * (uint32_t)stride;
    mbar_expect_tx(&sh.bar, bytes);
    bulk_g2s(img, src + (size_t)ya * stride, bytes, &sh.bar);
  }
  // the strip's rows of the corner bitmask are set by this CTA alone: it clears them itself (no per-frame memset), behind the bulk copy
  {
    uint32_t* own = L.cbits + ((size_t)s * H + y0) * words_per_row;
    for (int i = tid; i < (y1 - y0) * words_per_row; i += nthreads) own[i] = 0u;
  }
  __syncthreads();     // ... before any warp of the CTA sets a bit
  mbar_wait(&sh.bar, 0);

  // ---- dense pass: a lane slot = 16 pixels x 2 rows (one 128-bit shared-memory load per row), slots of the strip in raster order,
  //      32 per work item
  const uint32_t kcmp = (uint32_t)(127 - thr) * 0x01010101u;
  const int n_pairs = (y1 - y0 + 1) >> 1;
  const int cpr = (W + 15) >> 4, n_slots = n_pairs * cpr;
  uint8_t* next_img = nullptr; int next_pitch = 0;
  const int l1p = ((W >> 1) + 7) & ~7;                                      // level 0 only: pitch of the strip's level-1 rows in shared memory
  uint8_t* l1s = (uint8_t*)(qbase + nw * (kQueue + 64));
  if (kLevel0) { next_pitch = Lchild[0].pitch; next_img = Lchild[0].img + ((size_t)s * Lchild[0].h + (y0 >> 1)) * next_pitch; }   // level-1 row of y0

  // Candidates go through two stages, each 32 at a time: queue -> ring_even -> queue2 -> ring_test.
  uint32_t* queue2 = qbase + nw * kQueue + warp * 64;
  const uint32_t lt_mask = (1u << lane) - 1u;
  int qn = 0, q2n = 0;   // entries waiting in this warp's two queues (warp-uniform)
  auto two_stage = [&](uint32_t e, bool have) {
    const bool pass = have && ring_even(e, img, stride, thr);
    const uint32_t bal = __ballot_sync(0xffffffffu, pass);
    if (pass) queue2[q2n + __popc(bal & lt_mask)] = e;
    q2n += __popc(bal);
    __syncwarp();
    if (q2n >= 32) {
      q2n -= 32;
      const uint32_t e2 = queue2[q2n + lane];
      __syncwarp();
      ring_test(e2, img, stride, thr, gbits, words_per_row);
    }
  };
  const int n_items = (n_slots + 31) >> 5;
  for (int item = warp; item < n_items;) {     // the first nw items are dealt out statically, the rest to whichever warp is free
    const int slot = item * 32 + lane;
    const int pr = div_small(slot, L.mg_cpr);
    const int y = y0 + 2 * pr;           // rows y and y+1
    const int x0 = (slot - pr * cpr) << 4;
    uint32_t mask = 0;                   // bit 16 k + i: pixel (x0 + i, y + k) survives the rejection test
    if (slot < n_slots) {
      const uint8_t* r0 = img + (y - ya) * stride + x0;
      const uint4 c0 = *(const uint4*)r0;
      const bool have1 = (y + 1) < y1;
      const uint4 c1 = have1 ? *(const uint4*)(r0 + stride) : make_uint4(0u, 0u, 0u, 0u);
      if (kLevel0 && have1) {   // half-sample: eight pixels of level 1 per lane (columns past W/2 land in the row padding)
        const uint2 o = make_uint2(half4(c0.x, c0.y, c1.x, c1.y), half4(c0.z, c0.w, c1.z, c1.w));
        *(uint2*)(next_img + (uint32_t)(pr * next_pitch + (x0 >> 1))) = o;
        *(uint2*)(l1s + pr * l1p + (x0 >> 1)) = o;
      }
      // x in [3, W-4] only (jni/vision/cvfast.cpp:6115-6119)
      const int n_valid = W - 3 - x0;
      uint32_t vm = n_valid >= 16 ? 0xffffu : ((1u << n_valid) - 1u);
      if (x0 == 0) vm &= ~7u;
      // rows < 3 and >= H-3 are never tested, so the words left of x0 == 0 / right of the last chunk lie inside the staged rows
      // and only feed pixels that vm clears
      if (y >= 3 && y < H - 3) mask = reject16(r0, stride, c0, kcmp) & vm;
      if (have1 && y + 1 >= 3 && y + 1 < H - 3) mask |= (reject16(r0 + stride, stride, c1, kcmp) & vm) << 16;
    }
    const int mine = __popc(mask);
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t ebase = ((uint32_t)(y - ya) << 16) | (uint32_t)x0;     // bit b -> ebase + (b & 15) + ((b >> 4) << 16)
    if (total > 255) {   // dense levels: every lane feeds its own survivors (lanes are about equally loaded), no first queue
      while (__any_sync(0xffffffffu, mask != 0u)) {
        const bool have = mask != 0u;
        const int b = have ? 31 - __clz(mask) : 0;
        mask &= ~(1u << b);
        two_stage(ebase + (uint32_t)b + (uint32_t)(b & 16) * 4095u, have);
      }
    } else if (total) {
      int pos = qn + incl - mine;
      while (mask) {
        const int b = 31 - __clz(mask); mask ^= 1u << b;
        queue[pos++] = ebase + (uint32_t)b + (uint32_t)(b & 16) * 4095u;
      }
      qn += total;
      __syncwarp();
      while (qn >= 32) {
        qn -= 32;
        const uint32_t e = queue[qn + lane];
        two_stage(e, true);
      }
      __syncwarp();
    }
    if (lane == 0) item = atomicAdd(&sh.next_item, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
  }
  { const uint32_t e = lane < qn ? queue[lane] : 0u; two_stage(e, lane < qn); }
  if (lane < q2n) ring_test(queue2[lane], img, stride, thr, gbits, words_per_row);

  if (kLevel0) {   // levels 2 and 3 of this strip from its level-1 rows
    __syncthreads();
    emit_levels_2_3(l1s, l1p, (y1 - y0) >> 1, Lchild[1], Lchild[2], s, y0, tid, nthreads);
  }
}

// Corner lists of the level images from their corner bitmasks (k_pyramid_fast / k_fast_levels set the bits).  An image is cut evenly into chunks of
// at most kListChunk bitmask words (a VGA level 0 is three chunks, a 4K level 0 sixty-four), one CTA per chunk: k_corner_count leaves every chunk's
// corner count, k_corner_lists adds up the counts of the chunks before its own (at most a few dozen) and then gives every word its
// raster-order position -- every warp owns a contiguous run of words, popcounts + one scan of the warp totals.  The position at a row's
// first word is the row's LUT entry (jni/KeyFrame.cc:41-49).
// Positions are clamped to the capacity: on overflow -- reported through status[0] -- the readers of the LUT must not index past the list.
constexpr int kListChunk = 4096;    // target chunk size; an image is split EVENLY into ceil(words / kListChunk) chunks
__host__ __device__ __forceinline__ int list_chunks(int n_words) { return (n_words + kListChunk - 1) / kListChunk; }
__host__ __device__ __forceinline__ int list_chunk_words(int n_words) { const int nc = list_chunks(n_words); return (((n_words + nc - 1) / nc) + 31) & ~31; }

struct ListJob { int level, s, chunk, n_chunks, w_begin, w_end; };
// block -> (level, stream, chunk): all chunks of the largest level first.  (The four descriptors stay kernel parameters -- constant bank --
// selected by compile-time index after unrolling; an array of them would be copied to local memory by every thread.)
__device__ __forceinline__ ListJob list_job(const LevelDesc& L0, const LevelDesc& L1, const LevelDesc& L2, const LevelDesc& L3, int first_stream, int count, int first_level) {
  int blk = blockIdx.x;
  ListJob J; J.level = VS_LEVELS - 1; J.s = first_stream; J.chunk = 0; J.n_chunks = 1; J.w_begin = 0; J.w_end = 0;
  bool done = false;
#pragma unroll
  for (int l = 0; l < VS_LEVELS; l++) {
    const LevelDesc& L = l == 0 ? L0 : l == 1 ? L1 : l == 2 ? L2 : L3;
    const int n_words = L.h * ((L.w + 31) >> 5), nc = list_chunks(n_words), cw = list_chunk_words(n_words);
    if (!done && l >= first_level) {
      if (blk < count * nc || l == VS_LEVELS - 1) {
        const int si = blk / nc;
        J.level = l; J.n_chunks = nc; J.s = first_stream + si; J.chunk = blk - si * nc;
        J.w_begin = min(J.chunk * cw, n_words); J.w_end = min(J.w_begin + cw, n_words);
        done = true;
      } else blk -= count * nc;
    }
  }
  return J;
}

__global__ void __launch_bounds__(kMaxThreads)
k_corner_count(LevelDesc L0, LevelDesc L1, LevelDesc L2, LevelDesc L3, int first_stream, int count, int first_level, int* __restrict__ chunk_counts) {
  __shared__ int wsum[kMaxWarps];
  const ListJob J = list_job(L0, L1, L2, L3, first_stream, count, first_level);
#define PICK(f) (J.level == 0 ? L0.f : J.level == 1 ? L1.f : J.level == 2 ? L2.f : L3.f)   /* a select among kernel parameters: no local copy of a descriptor */
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const uint32_t* bits = PICK(cbits) + (size_t)J.s * PICK(h) * ((PICK(w) + 31) >> 5);
  int c = 0;
  for (int w = J.w_begin + tid; w < J.w_end; w += blockDim.x) c += __popc(bits[w]);
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) wsum[warp] = c;
  __syncthreads();
  if (tid == 0) { int t = 0; for (int i = 0; i < nw; i++) t += wsum[i]; chunk_counts[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(kMaxThreads)
k_corner_lists(LevelDesc L0, LevelDesc L1, LevelDesc L2, LevelDesc L3, int first_stream, int count, int first_level, const int* __restrict__ chunk_counts,
               int* __restrict__ status) {
  __shared__ int wsum[kMaxWarps];
  __shared__ int s_base;
  const ListJob J = list_job(L0, L1, L2, L3, first_stream, count, first_level);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nw = nthreads >> 5;
  const int s = J.s, H = PICK(h), wpr = (PICK(w) + 31) >> 5, cap = PICK(cap);
  const uint32_t mg_wpr = PICK(mg_wpr);
  const uint32_t* bits = PICK(cbits) + (size_t)s * H * wpr;
  if (warp == 0) {   // corners of this image before this chunk, and (last chunk) the image's total
    int c = 0;
    const int* cc = chunk_counts + (blockIdx.x - J.chunk);
    for (int k = lane; k < J.chunk; k += 32) c += cc[k];
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) s_base = c;
  }
  // a warp owns a contiguous run of words (a multiple of 32, so that every pass reads whole 128-byte lines) and walks it 32 words at a time
  const int n_words = J.w_end - J.w_begin;
  const int per_warp = ((n_words + nw - 1) / nw + 31) & ~31;
  const int wb = J.w_begin + min(warp * per_warp, n_words), we = min(wb + per_warp, J.w_end);
  int c = 0;
  for (int w = wb + lane; w < we; w += 32) c += __popc(bits[w]);
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) wsum[warp] = c;
  __syncthreads();
  int ws = (lane < nw) ? wsum[lane] : 0, wi = ws;
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += v; }
  const int total = s_base + __shfl_sync(0xffffffffu, wi, 15);       // corners up to the end of this chunk
  int run = s_base + __shfl_sync(0xffffffffu, wi - ws, warp);        // corners before this warp's run
  int* lut = PICK(lut) + (size_t)s * (H + 1);
  uint32_t* out = PICK(corners) + (size_t)s * cap;
  uint32_t m_next = (wb + lane < we) ? bits[wb + lane] : 0u;
  for (int w0 = wb; w0 < we; w0 += 32) {
    const int w = w0 + lane;
    uint32_t m = m_next;
    m_next = (w + 32 < we) ? bits[w + 32] : 0u;                // the next 32 words are on their way while these are scanned
    const int mine = __popc(m);
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    int pos = run + incl - mine;
    run += __shfl_sync(0xffffffffu, incl, 31);
    if (w < we) {
      const int r = div_small(w, mg_wpr), col = w - r * wpr;
      if (col == 0) lut[r] = min(pos, cap);
      const uint32_t cw0 = ((uint32_t)r << 16) | ((uint32_t)col << 5);
      if (pos + mine <= cap) {               // (always, unless the list overflows) highest bit first, written from the back: one FLO per corner
        uint32_t* o = out + pos + mine;
        while (m) { const int b = 31 - __clz(m); m ^= 1u << b; *--o = cw0 + (uint32_t)b; }
      } else {
        while (m) { const int b = __ffs(m) - 1; m &= m - 1; if (pos < cap) out[pos] = cw0 + (uint32_t)b; pos++; }
      }
    }
  }
  if (tid == 0 && J.chunk == J.n_chunks - 1) {
    lut[H] = min(total, cap);
    if (total > cap) atomicExch(&status[0], 1);
  }
#undef PICK
}

// Level 0 of every stream: pyramid levels 1..3 + FAST-10 of level 0.
__global__ void __maxnreg__(VS_PYR_REGS)
k_pyramid_fast(LevelDesc L, LevelDesc L1, LevelDesc L2, LevelDesc L3, const uint8_t* const* __restrict__ l0_ptr, const int* __restrict__ l0_stride,
               int first_stream, int thr, unsigned* __restrict__ ticket) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ StripShared sh;
  __shared__ LevelDesc child[3];
  cudaGridDependencySynchronize(); cudaTriggerProgrammaticLaunchCompletion();   // programmatic dependent launch (vs_launch_pdl); no-ops otherwise
  if (threadIdx.x == 0) {
    const unsigned tk = atomicAdd(ticket, 1u);
    if (tk == gridDim.x - 1) *ticket = 0u;        // every ticket of this launch has been drawn: the counter is ready for the next frame (no memset)
    sh.ticket = (int)tk; sh.next_item = (int)(blockDim.x >> 5); mbar_init(&sh.bar, 1); child[0] = L1; child[1] = L2; child[2] = L3;
  }
  __syncthreads();
  const int t = sh.ticket;
  const int q = div_small(t, L.mg_strips), s = first_stream + q, strip = t - q * L.n_strips;
  fast_strip<true>(L, l0_ptr[s], l0_stride[s], s, strip, thr, child, sh, smem);
}

// FAST-10 of levels 1..3 of every stream in one launch: tickets run over level 1's strips, then level 2's, then level 3's.
__global__ void __maxnreg__(VS_PYR_REGS)
k_fast_levels(LevelDesc L1, LevelDesc L2, LevelDesc L3, int first_stream, int count, int thr1, int thr2, int thr3, unsigned* __restrict__ ticket) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ StripShared sh;
  cudaGridDependencySynchronize(); cudaTriggerProgrammaticLaunchCompletion();   // programmatic dependent launch (vs_launch_pdl); no-ops otherwise
  if (threadIdx.x == 0) {
    const unsigned tk = atomicAdd(ticket, 1u);
    if (tk == gridDim.x - 1) *ticket = 0u;
    sh.ticket = (int)tk; sh.next_item = (int)(blockDim.x >> 5); mbar_init(&sh.bar, 1);
  }
  __syncthreads();
  int t = sh.ticket, thr = thr1;
  const int n1 = count * L1.n_strips, n2 = count * L2.n_strips;
  LevelDesc L = L1;
  if (t >= n1) { t -= n1; L = L2; thr = thr2; if (t >= n2) { t -= n2; L = L3; thr = thr3; } }
  const int q = div_small(t, L.mg_strips), s = first_stream + q, strip = t - q * L.n_strips;
  fast_strip<false>(L, L.img + (size_t)s * L.h * L.pitch, L.pitch, s, strip, thr, nullptr, sh, smem);
}

// plain 2:1 half-sample for the (rare) source-keyframe uploads
__global__ void k_half_sample(const uint8_t* __restrict__ src, int sw, int sh, int spitch, uint8_t* __restrict__ dst, int dpitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= sw / 2 || y >= sh / 2) return;
  const uint8_t* a = src + (size_t)(2 * y) * spitch + 2 * x;
  dst[(size_t)y * dpitch + x] = (uint8_t)((a[0] + a[1] + a[spitch] + a[spitch + 1] + 2) >> 2);
}

size_t pyrfast_smem_bytes(int stride, int w, int rows, bool level0, int warps = kMaxWarps) {
  const int l1p = ((w >> 1) + 7) & ~7;
  return (size_t)(rows + 6) * stride + (size_t)warps * (kQueue + 64) * 4 + (level0 ? (size_t)(rows / 2) * l1p : 0);
}

// Threads per CTA for strips of `rows` rows of a w-pixel level: as many warps as the strip has work items (32 lanes x 16 pixels x 2
// rows), or an even share of them when there are more than ten; at least four (the write-out scan and the level 2-3 rows use every thread).
int pyrfast_threads(int w, int rows) {
  const int items = (((rows + 1) / 2) * ((w + 15) / 16) + 31) / 32;
  const int rounds = (items + kMaxWarps - 1) / kMaxWarps;
  int warps = (items + rounds - 1) / rounds;
  warps = warps < 4 ? 4 : warps;
  return 32 * warps;
}

}  // namespace

// Rows per CTA for level l.  Level 0: 32 where four to five CTAs still fit on an SM (the fixed cost per thread -- set-up, level 2-3 rows -- is
// spread over twice the pixels: 0.121 vs 0.140 ms for 256 VGA frames), else 16; the level-1..3 rows a CTA emits stay whole (32 -> 16, 8, 4).
// Levels 1-3: 16 (8 and 32 / 64 rows measured slower on B200).
int vs_strip_rows(int level, int w, int pitch) {
  // wide images (1080p, 4K): 16 rows of level 0 no longer leave room for several CTAs per SM (4K: over 100 KB per CTA, one CTA per SM);
  // 8-row strips keep 3-5 CTAs resident.  8 is the floor at level 0: a strip must cover whole rows of level 3.
  if (const char* e = getenv("VSLAM_STRIP_ROWS")) {     // tuning experiments only: "r0,r1,r2,r3" (r0 a multiple of 8, the others even, <= VS_MAX_STRIP_ROWS)
    int r[VS_LEVELS] = {0, 0, 0, 0};
    if (sscanf(e, "%d,%d,%d,%d", &r[0], &r[1], &r[2], &r[3]) == 4 && r[level] >= 2 && r[level] <= VS_MAX_STRIP_ROWS && r[level] % (level == 0 ? 8 : 2) == 0 &&
        pyrfast_smem_bytes(pitch, w, r[level], level == 0) <= 200 * 1024)
      return r[level];
  }
  if (level == 0 && pyrfast_smem_bytes(pitch, w, 32, true) <= 44 * 1024) return 32;   // (VGA: 43.5 KB; four CTAs of 48 registers per SM)
  if (pyrfast_smem_bytes(pitch, w, 16, level == 0) > 48 * 1024) return 8;
  return 16;
}

// FAST thresholds per level (jni/KeyFrame.cc:32-39)
static const int kFastThr[VS_LEVELS] = {10, 15, 15, 10};

// Level 0: pyramid levels 1..3 + FAST-10 of level 0 (one launch).
int vs_launch_pyramid_l0(vslam_ctx* ctx, int first_stream, int count) {
  // [0]: level-0 launch, [1]: levels 1-3 launch; one pair per stream group.  Zero at creation; the CTA that draws a launch's last ticket
  // zeroes the counter again, and every CTA clears its own rows of the corner bitmask: nothing to memset per frame.
  unsigned* tickets = ctx->tickets + 2 * ctx->cur_group;
  LevelDesc& L = ctx->lev[0];
  int stride = 0;
  for (int s = first_stream; s < first_stream + count; s++) stride = ctx->l0_stride_host[s] > stride ? ctx->l0_stride_host[s] : stride;
  const int threads = pyrfast_threads(L.w, L.strip_rows);
  const size_t smem = pyrfast_smem_bytes(stride, L.w, L.strip_rows, true, threads / 32);
  if (smem > 227 * 1024) { ctx->err = "pyramid_fast: strip does not fit in shared memory (row stride too large)"; return VSLAM_E_INVALID; }
  VS_CUDA(cudaFuncSetAttribute(k_pyramid_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  vs_time_begin(ctx, VS_ST_PYR0);
  VS_CUDA(vs_launch_pdl(k_pyramid_fast, dim3(count * L.n_strips), dim3(threads), smem, ctx->stream, ctx->pdl && !ctx->timing, L, ctx->lev[1], ctx->lev[2], ctx->lev[3],
                        (const uint8_t* const*)ctx->l0_ptr, (const int*)ctx->l0_stride, first_stream, kFastThr[0], tickets));
  vs_time_end(ctx);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}

// Levels 1..3: FAST-10 (one launch).  Needs the level images written by vs_launch_pyramid_l0.
int vs_launch_fast_levels(vslam_ctx* ctx, int first_stream, int count) {
  size_t smem = 0; int blocks = 0, threads = 0;
  for (int l = 1; l < VS_LEVELS; l++) {
    const LevelDesc& L = ctx->lev[l];
    blocks += count * L.n_strips;
    const int t = pyrfast_threads(L.w, L.strip_rows);
    threads = t > threads ? t : threads;
  }
  for (int l = 1; l < VS_LEVELS; l++) {
    const LevelDesc& L = ctx->lev[l];
    const size_t b = pyrfast_smem_bytes(L.pitch, L.w, L.strip_rows, false, threads / 32);
    smem = b > smem ? b : smem;
  }
  if (smem > 227 * 1024) { ctx->err = "fast_levels: strip does not fit in shared memory"; return VSLAM_E_INVALID; }
  VS_CUDA(cudaFuncSetAttribute(k_fast_levels, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  vs_time_begin(ctx, VS_ST_PYR1);
  VS_CUDA(vs_launch_pdl(k_fast_levels, dim3(blocks), dim3(threads), smem, ctx->stream, ctx->pdl && !ctx->timing, ctx->lev[1], ctx->lev[2], ctx->lev[3], first_stream, count,
                        kFastThr[1], kFastThr[2], kFastThr[3], ctx->tickets + 2 * ctx->cur_group + 1));
  vs_time_end(ctx);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}

// Corner lists and row LUTs of all four levels from the corner bitmasks (one launch).  The patch search of the tracker
// (search_fast.cu) reads the bitmasks themselves, so a tracked frame leaves this to whoever asks for the lists next (vs_ensure_lists).
int vs_launch_corner_lists(vslam_ctx* ctx, int first_stream, int count) {
  int blocks = 0;
  for (int l = 0; l < VS_LEVELS; l++) { const LevelDesc& L = ctx->lev[l]; blocks += count * list_chunks(L.h * ((L.w + 31) / 32)); }
  if ((size_t)blocks > ctx->list_counts_cap) {
    if (ctx->list_counts) cudaFree(ctx->list_counts);
    ctx->list_counts = nullptr; ctx->list_counts_cap = 0;
    VS_CUDA(cudaMalloc(&ctx->list_counts, sizeof(int) * (size_t)blocks));
    ctx->list_counts_cap = (size_t)blocks;
  }
  vs_time_begin(ctx, VS_ST_PYR2);
  k_corner_count<<<blocks, kMaxThreads, 0, ctx->stream>>>(ctx->lev[0], ctx->lev[1], ctx->lev[2], ctx->lev[3], first_stream, count, 0, ctx->list_counts);
  k_corner_lists<<<blocks, kMaxThreads, 0, ctx->stream>>>(ctx->lev[0], ctx->lev[1], ctx->lev[2], ctx->lev[3], first_stream, count, 0, ctx->list_counts, ctx->status);
  vs_time_end(ctx);
  VS_CUDA(cudaGetLastError());
  ctx->launches += 2;
  return VSLAM_OK;
}

int vs_ensure_lists(vslam_ctx* ctx) {
  if (!ctx->lists_stale) return VSLAM_OK;
  ctx->lists_stale = false;
  return vs_launch_corner_lists(ctx, 0, ctx->S);
}

int vs_launch_pyramid_fast(vslam_ctx* ctx, int first_stream, int count) {
  int rc = vs_launch_pyramid_l0(ctx, first_stream, count);
  if (!rc) rc = vs_launch_fast_levels(ctx, first_stream, count);
  return rc ? rc : vs_launch_corner_lists(ctx, first_stream, count);
}

int vs_launch_source_pyramid(vslam_ctx* ctx, int kf) {
  for (int l = 1; l < VS_LEVELS; l++) {
    const SourceKF& S = ctx->src;
    const uint8_t* src = S.img[l - 1] + (size_t)kf * S.h[l - 1] * S.pitch[l - 1];
    uint8_t* dst = S.img[l] + (size_t)kf * S.h[l] * S.pitch[l];
    dim3 b(32, 8), g((S.w[l] + 31) / 32, (S.h[l] + 7) / 8);
    k_half_sample<<<g, b, 0, ctx->stream>>>(src, S.w[l - 1], S.h[l - 1], S.pitch[l - 1], dst, S.pitch[l]);
    VS_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  return VSLAM_OK;
}
