// KeyFrame::MakeKeyFrame_Lite on the GPU (reference: jni/KeyFrame.cc:5-51).
//
// One kernel per pyramid level l, batched over all streams:  a CTA owns a strip of VS_STRIP_ROWS rows of level l
// of one stream.  It
//   1. stages the strip plus a 3-row halo into shared memory with ONE bulk async copy (cp.async.bulk, TMA engine);
//   2. writes the matching rows of level l+1:  (a+b+c+d+2)>>2  (cv::resize 2:1, jni/KeyFrame.cc:20-23);
//   3. runs FAST-10 (jni/vision/cvfast.cpp:6088-9241 == segment test):
//        a byte-SIMD rejection test on 4 pixels per lane (VABSDIFF4 + SWAR compares): a 10-arc contains at least one
//        pixel of every opposite ring pair, so both (0,8) and (4,12) must hold a pixel differing by more than t;
//        survivors are compacted into a per-warp queue and get the exact 16-pixel ring test, one lane per candidate;
//   4. appends its corners to the stream's list in raster order: per-row popcounts of a shared-memory corner bitmask,
//      a decoupled look-back across the strips of the image (tickets guarantee predecessors are resident), and the
//      running row offsets ARE the row look-up table of jni/KeyFrame.cc:41-49.
// Integer / byte arithmetic only; results are bit-exact against the oracle.
#include "vslam_internal.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kQueue = 288;  // per-warp candidate queue: up to 31 carried over + 256 of one work item (2 rows x 128 pixels)

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

// Exact FAST-10 segment test of one candidate (queue entry = smem row << 16 | x); sets its bit in the strip's corner bitmask.
__device__ __forceinline__ void ring_test(uint32_t e, const uint8_t* img, int stride, int thr, uint32_t* bitmask, int words_per_row, int row_bias) {
  const int ry = e >> 16, x = e & 0xffff;
  const uint8_t* p = img + (size_t)ry * stride + x;
  const int cb = (int)p[0] + thr, c_b = (int)p[0] - thr;
  const int s1 = stride, s2 = 2 * stride, s3 = 3 * stride;
  uint32_t br = 0, dk = 0;
  // sign bit of (cb - v) <=> v > cb ; sign bit of (v - c_b) <=> v < c_b : shifted into the masks with funnel shifts
#define RING(off)                                                       \
  {                                                                     \
    const int v = (int)p[(off)];                                        \
    br = __funnelshift_l((uint32_t)(cb - v), br, 1);                    \
    dk = __funnelshift_l((uint32_t)(v - c_b), dk, 1);                   \
  }
  RING(s3) RING(1 + s3) RING(2 + s2) RING(3 + s1) RING(3) RING(3 - s1) RING(2 - s2) RING(1 - s3)
  RING(-s3) RING(-1 - s3) RING(-2 - s2) RING(-3 - s1) RING(-3) RING(-3 + s1) RING(-2 + s2) RING(-1 + s3)
#undef RING
  br |= br << 16; dk |= dk << 16;
  uint32_t b1 = br & (br >> 1), d1 = dk & (dk >> 1);
  uint32_t b2 = b1 & (b1 >> 2), d2 = d1 & (d1 >> 2);
  b2 &= b2 >> 4; d2 &= d2 >> 4;
  b2 &= b1 >> 8; d2 &= d1 >> 8;
  if ((b2 | d2) & 0xffffu) atomicOr(&bitmask[(ry + row_bias) * words_per_row + (x >> 5)], 1u << (x & 31));
}

constexpr unsigned long long kFlagAgg = 1ull << 32, kFlagInc = 2ull << 32;

__global__ void __launch_bounds__(kThreads)
k_pyramid_fast(LevelDesc L, LevelDesc Ln, int has_next, int level, const uint8_t* const* __restrict__ l0_ptr, const int* __restrict__ l0_stride,
               int first_stream, int thr, unsigned* __restrict__ ticket, int* __restrict__ status) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ int s_ticket;
  __shared__ int s_rowcnt[VS_STRIP_ROWS], s_rowoff[VS_STRIP_ROWS + 1];
  __shared__ int s_base;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_ticket = (int)atomicAdd(ticket, 1u); mbar_init(&bar, 1); }
  __syncthreads();
  const int t = s_ticket;
  const int s = first_stream + t / L.n_strips, strip = t % L.n_strips;

  const uint8_t* src; int stride;
  if (level == 0) { src = l0_ptr[s]; stride = l0_stride[s]; }
  else { src = L.img + (size_t)s * L.h * L.pitch; stride = L.pitch; }

  const int W = L.w, H = L.h;
  const int y0 = strip * VS_STRIP_ROWS, y1 = min(y0 + VS_STRIP_ROWS, H);
  const int ya = max(y0 - 3, 0), yb = min(y1 + 3, H);
  const int words_per_row = (W + 31) >> 5;
  uint8_t* img = smem;                                                  // rows ya..yb-1, `stride` bytes apart
  uint32_t* bitmask = (uint32_t*)(smem + (size_t)(VS_STRIP_ROWS + 6) * stride);   // [VS_STRIP_ROWS][words_per_row]
  uint32_t* queue = bitmask + VS_STRIP_ROWS * words_per_row + warp * kQueue;

  if (tid == 0) {
    const uint32_t bytes = (uint32_t)(yb - ya) * (uint32_t)stride;
    mbar_expect_tx(&bar, bytes);
    bulk_g2s(img, src + (size_t)ya * stride, bytes, &bar);
  }
  for (int i = tid; i < VS_STRIP_ROWS * words_per_row; i += kThreads) bitmask[i] = 0;
  __syncthreads();
  mbar_wait(&bar, 0);

  const uint32_t kcmp = (uint32_t)(127 - thr) * 0x01010101u;
  const int n_chunks = (W + 127) >> 7;
  const int n_pairs = (y1 - y0 + 1) >> 1;
  uint8_t* next_img = has_next ? Ln.img + (size_t)s * Ln.h * Ln.pitch : nullptr;

  int qn = 0;   // candidates waiting in this warp's queue (warp-uniform)
  int pr = warp / n_chunks, ch = warp - pr * n_chunks;   // work item = (row pair, 128-pixel chunk); advanced incrementally below
  for (int item = warp; item < n_pairs * n_chunks; item += kWarps, ch += kWarps) {
    while (ch >= n_chunks) { ch -= n_chunks; pr++; }
    const int y = y0 + 2 * pr;           // rows y and y+1
    const int x0 = (ch << 7) + (lane << 2);
    const bool active = x0 < W;
    uint32_t cand[2] = {0u, 0u};
    if (active) {
      const uint8_t* r0 = img + (size_t)(y - ya) * stride + x0;
      const uint32_t c0 = *(const uint32_t*)r0;
      const bool have1 = (y + 1) < y1;
      const uint32_t c1 = have1 ? *(const uint32_t*)(r0 + stride) : 0u;
      if (has_next && have1) {   // half-sample: two output pixels per lane
        const unsigned o0 = __dp4a(c0, 0x00000101u, __dp4a(c1, 0x00000101u, 2u)) >> 2;
        const unsigned o1 = __dp4a(c0, 0x01010000u, __dp4a(c1, 0x01010000u, 2u)) >> 2;
        *(uint16_t*)(next_img + (size_t)(y >> 1) * Ln.pitch + (x0 >> 1)) = (uint16_t)(o0 | (o1 << 8));
      }
      // x in [3, W-4] only (jni/vision/cvfast.cpp:6115-6119)
      uint32_t vmask = 0x80808080u;
      if (x0 == 0) vmask = 0x80000000u;
      if (x0 == W - 4) vmask &= 0x00000080u;
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const int yy = y + k;
        if (yy < 3 || yy >= H - 3 || yy >= y1) continue;
        const uint8_t* r = r0 + k * stride;
        const uint32_t c = k ? c1 : c0;
        const uint32_t up = *(const uint32_t*)(r - 3 * stride), dn = *(const uint32_t*)(r + 3 * stride);
        const uint32_t m1 = bytes_gt(__vabsdiffu4(up, c), kcmp) | bytes_gt(__vabsdiffu4(dn, c), kcmp);
        const uint32_t lw = (x0 > 0) ? *(const uint32_t*)(r - 4) : 0u, rw = (x0 + 4 < W) ? *(const uint32_t*)(r + 4) : 0u;
        const uint32_t lf = __byte_perm(lw, c, 0x4321), rt = __byte_perm(c, rw, 0x6543);
        const uint32_t m2 = bytes_gt(__vabsdiffu4(lf, c), kcmp) | bytes_gt(__vabsdiffu4(rt, c), kcmp);
        cand[k] = m1 & m2 & vmask;
      }
    }
    // append the candidates of this work item to the warp's queue (order is irrelevant: corners land in a bitmask)
    const int mine = __popc(cand[0]) + __popc(cand[1]);
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    int pos = qn + incl - mine;
    qn += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
    for (int k = 0; k < 2; k++) {
      uint32_t m = cand[k];
      const uint32_t ebase = ((uint32_t)(y + k - ya) << 16) | (uint32_t)x0;
      while (m) {
        const int b = __ffs(m) - 1; m &= m - 1;
        queue[pos++] = ebase + (uint32_t)(b >> 3);
      }
    }
    __syncwarp();
    // exact ring test on full groups of 32 queued candidates; the remainder waits for the next work item
    while (qn >= 32) { qn -= 32; ring_test(queue[qn + lane], img, stride, thr, bitmask, words_per_row, ya - y0); }
    __syncwarp();
  }
  if (lane < qn) ring_test(queue[lane], img, stride, thr, bitmask, words_per_row, ya - y0);
  __syncthreads();

  // per-row corner counts
  const int rows = y1 - y0;
  for (int r = warp; r < rows; r += kWarps) {
    int c = 0;
    for (int w = lane; w < words_per_row; w += 32) c += __popc(bitmask[r * words_per_row + w]);
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane == 0) s_rowcnt[r] = c;
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int r = 0; r < rows; r++) { s_rowoff[r] = acc; acc += s_rowcnt[r]; }
    s_rowoff[rows] = acc;
    // decoupled look-back over the strips of this image (predecessors hold lower tickets, hence are resident or done)
    unsigned long long* st = L.strip_state + (size_t)s * L.n_strips;
    unsigned long long excl = 0;
    if (strip == 0) {
      atomicExch(&st[0], kFlagInc | (unsigned)acc);
    } else {
      atomicExch(&st[strip], kFlagAgg | (unsigned)acc);
      for (int j = strip - 1; j >= 0; j--) {
        unsigned long long v;
        do { v = *(volatile unsigned long long*)&st[j]; } while ((v >> 32) == 0);
        excl += (unsigned)v;
        if ((v >> 32) == 2) break;
      }
      atomicExch(&st[strip], kFlagInc | (unsigned)(excl + acc));
    }
    s_base = (int)excl;
    if ((long long)excl + acc > L.cap) atomicExch(&status[0], 1);
  }
  __syncthreads();
  const int base = s_base;
  int* lut = L.lut + (size_t)s * (H + 1);
  if (tid < rows) lut[y0 + tid] = base + s_rowoff[tid];
  if (tid == 0 && y1 == H) lut[H] = base + s_rowoff[rows];
  uint32_t* out = L.corners + (size_t)s * L.cap;
  for (int r = warp; r < rows; r += kWarps) {
    int run = base + s_rowoff[r];
    const uint32_t yy = (uint32_t)(y0 + r) << 16;
    for (int w0 = 0; w0 < words_per_row; w0 += 32) {
      const int w = w0 + lane;
      uint32_t m = (w < words_per_row) ? bitmask[r * words_per_row + w] : 0u;
      const int mine = __popc(m);
      int incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
      int pos = run + incl - mine;
      while (m) {
        const int b = __ffs(m) - 1; m &= m - 1;
        if (pos < L.cap) out[pos] = yy | (uint32_t)((w << 5) + b);
        pos++;
      }
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

// plain 2:1 half-sample for the (rare) source-keyframe uploads
__global__ void k_half_sample(const uint8_t* __restrict__ src, int sw, int sh, int spitch, uint8_t* __restrict__ dst, int dpitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= sw / 2 || y >= sh / 2) return;
  const uint8_t* a = src + (size_t)(2 * y) * spitch + 2 * x;
  dst[(size_t)y * dpitch + x] = (uint8_t)((a[0] + a[1] + a[spitch] + a[spitch + 1] + 2) >> 2);
}

size_t pyrfast_smem_bytes(int stride, int w) {
  const int words_per_row = (w + 31) >> 5;
  return (size_t)(VS_STRIP_ROWS + 6) * stride + (size_t)VS_STRIP_ROWS * words_per_row * 4 + (size_t)kWarps * kQueue * 4;
}

}  // namespace

// FAST thresholds per level (jni/KeyFrame.cc:32-39)
static const int kFastThr[VS_LEVELS] = {10, 15, 15, 10};

int vs_launch_pyramid_fast(vslam_ctx* ctx, int first_stream, int count) {
  VS_CUDA(cudaMemsetAsync(ctx->tickets, 0, sizeof(unsigned) * VS_LEVELS, ctx->stream));
  for (int l = 0; l < VS_LEVELS; l++) {
    LevelDesc& L = ctx->lev[l];
    VS_CUDA(cudaMemsetAsync(L.strip_state + (size_t)first_stream * L.n_strips, 0, sizeof(unsigned long long) * (size_t)count * L.n_strips, ctx->stream));
  }
  for (int l = 0; l < VS_LEVELS; l++) {
    LevelDesc& L = ctx->lev[l];
    int stride = L.pitch;
    if (l == 0) { stride = 0; for (int s = first_stream; s < first_stream + count; s++) stride = ctx->l0_stride_host[s] > stride ? ctx->l0_stride_host[s] : stride; }
    const size_t smem = pyrfast_smem_bytes(stride, L.w);
    if (smem > 227 * 1024) { ctx->err = "pyramid_fast: strip does not fit in shared memory (row stride too large)"; return VSLAM_E_INVALID; }
    VS_CUDA(cudaFuncSetAttribute(k_pyramid_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool has_next = l + 1 < VS_LEVELS;
    vs_time_begin(ctx, VS_ST_PYR0 + l);
    k_pyramid_fast<<<count * L.n_strips, kThreads, smem, ctx->stream>>>(L, has_next ? ctx->lev[l + 1] : L, has_next ? 1 : 0, l, ctx->l0_ptr, ctx->l0_stride,
                                                                       first_stream, kFastThr[l], ctx->tickets + l, ctx->status);
    vs_time_end(ctx);
    VS_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  return VSLAM_OK;
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
