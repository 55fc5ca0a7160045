// KeyFrame::MakeKeyFrame_Rest on the GPU (reference: jni/KeyFrame.cc:53-101, without its SmallBlurryImage tail):
//   k_fast_scores   old_style_corner_score for every FAST corner (jni/vision/cvfast.cpp:9337-9396), barrier 10 on all levels
//   k_nonmax_st     nonmax_suppression (jni/vision/cvfast.cpp:9243-9335) + FindShiTomasiScoreAtPoint for the maxima
//                   (jni/vision/ImageHandler.cpp:124-155) + raster-ordered compaction of vMaxCorners / vCandidates
// MiniPatch (jni/MiniPatch.cc:6-83): k_minipatch_sample (SampleFromImage) and k_minipatch_find (FindPatch, raw SSD).
// Both run only while a map is being bootstrapped or a keyframe is inserted: one stream per call, simple launches.
//
// The reference's non-max is a sequential cursor walk; per corner its outcome is a pure function of the 8 neighbours
// (left/right in the list, three above, three below), reproduced here with binary searches in the row LUT.  Quirk kept:
// the "right" test also requires the PREVIOUS list entry to share the row (jni/vision/cvfast.cpp:9284 reads corners[i-1]).
#include "vslam_internal.cuh"

namespace {

__device__ __forceinline__ int ring_score(const uint8_t* p, int stride, int barrier) {
  const int c = p[0], cb = c + barrier, c_b = c - barrier;
  const int s1 = stride, s2 = 2 * stride, s3 = 3 * stride;
  const int off[16] = {s3, 1 + s3, 2 + s2, 3 + s1, 3, 3 - s1, 2 - s2, 1 - s3, -s3, -1 - s3, -2 - s2, -3 - s1, -3, -3 + s1, -2 + s2, -1 + s3};
  int sp = 0, sn = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) { const int v = p[off[k]]; if (v > cb) sp += v - cb; else if (v < c_b) sn += c_b - v; }
  return sp > sn ? sp : sn;
}

__global__ void k_fast_scores(const uint8_t* img, int stride, const uint32_t* corners, int n, int barrier, int* scores) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = corners[i];
  scores[i] = ring_score(img + (size_t)(c >> 16) * stride + (c & 0xffff), stride, barrier);
}

// first index in [lo,hi) whose x >= x0 (corners of one row are sorted by x)
__device__ __forceinline__ int lower_x(const uint32_t* c, int lo, int hi, int x0) {
  while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)(c[mid] & 0xffff) < x0) lo = mid + 1; else hi = mid; }
  return lo;
}

// one CTA; flags[i] bit0 = maximal, bit1 = candidate; st[i] = Shi-Tomasi score of maxima inside the border
__global__ void k_nonmax_st(const uint8_t* img, int stride, int W, int H, const uint32_t* corners, const int* lut, int n, const int* scores,
                            uint32_t* max_out, uint32_t* cand_out, double* cand_score, int* counts) {
  __shared__ int s_base[2], s_warp[32][2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  if (tid < 2) s_base[tid] = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + tid;
    bool is_max = false, is_cand = false; double st = 0.0; uint32_t cw = 0;
    if (i < n) {
      cw = corners[i];
      const int px = cw & 0xffff, py = cw >> 16, score = scores[i];
      bool sup = false;
      if (i > 0) { const uint32_t l = corners[i - 1]; if ((int)(l & 0xffff) == px - 1 && (int)(l >> 16) == py && scores[i - 1] > score) sup = true; }
      if (!sup && i < n - 1 && i > 0) {
        const uint32_t r = corners[i + 1], l = corners[i - 1];
        if ((int)(r & 0xffff) == px + 1 && (int)(l >> 16) == py && scores[i + 1] > score) sup = true;
      }
      for (int dy = -1; dy <= 1 && !sup; dy += 2) {
        const int yy = py + dy;
        if (yy < 0 || yy >= H) continue;
        const int lo = lut[yy], hi = lut[yy + 1];
        for (int j = lower_x(corners, lo, hi, px - 1); j < hi && (int)(corners[j] & 0xffff) <= px + 1; j++) if (scores[j] > score) { sup = true; break; }
      }
      is_max = !sup;
      const int border = 10;
      if (is_max && px >= border && py >= border && px < W - border && py < H - border) {
        double dXX = 0, dYY = 0, dXY = 0;   // integer-valued sums: exact in any order
        for (int cy = py - 3; cy <= py + 3; cy++) {
          const uint8_t* r = img + (size_t)cy * stride;
          for (int cx = px - 3; cx <= px + 3; cx++) {
            const double dx = (double)((int)r[cx + 1] - (int)r[cx - 1]), dy = (double)((int)r[cx + stride] - (int)r[cx - stride]);
            dXX += dx * dx; dYY += dy * dy; dXY += dx * dy;
          }
        }
        const int nPixels = 49;
        dXX = dXX / (2.0 * nPixels); dYY = dYY / (2.0 * nPixels); dXY = dXY / (2.0 * nPixels);
        st = 0.5 * (dXX + dYY - sqrt((dXX + dYY) * (dXX + dYY) - 4 * (dXX * dYY - dXY * dXY)));
        is_cand = st > 70.0;   // gvdCandidateMinSTScore (jni/KeyFrame.cc:57)
      }
    }
    // raster-ordered compaction of both lists
    const unsigned bm = __ballot_sync(0xffffffffu, is_max), bc = __ballot_sync(0xffffffffu, is_cand);
    if (lane == 0) { s_warp[warp][0] = __popc(bm); s_warp[warp][1] = __popc(bc); }
    __syncthreads();
    int om = s_base[0], oc = s_base[1];
    for (int w = 0; w < warp; w++) { om += s_warp[w][0]; oc += s_warp[w][1]; }
    const unsigned lt = (1u << lane) - 1u;
    if (is_max) max_out[om + __popc(bm & lt)] = cw;
    if (is_cand) { const int k = oc + __popc(bc & lt); cand_out[k] = cw; cand_score[k] = st; }
    __syncthreads();
    if (tid == 0) { int a = 0, b = 0; for (int w = 0; w < nwarps; w++) { a += s_warp[w][0]; b += s_warp[w][1]; } s_base[0] += a; s_base[1] += b; }
    __syncthreads();
  }
  if (tid < 2) counts[tid] = s_base[tid];
}

// MiniPatch::SampleFromImage (jni/MiniPatch.cc:73-83): 9x9 bytes around (x,y) of level 0
__global__ void k_minipatch_sample(const uint8_t* img, int stride, const int* xy, int n, uint8_t* patches) {
  const int i = blockIdx.x, t = threadIdx.x;
  if (i >= n || t >= 81) return;
  const int r = t / 9, c = t - 9 * r;
  patches[(size_t)i * 81 + t] = img[(size_t)(xy[2 * i + 1] - 4 + r) * stride + (xy[2 * i] - 4 + c)];
}

// MiniPatch::FindPatch (jni/MiniPatch.cc:32-70) with SSDAtPoint (:6-27): one warp per trail.
// pos: n x 2 doubles in/out (integer valued); found: n ints out; best: n ints out (best SSD, max_ssd + 1 if nothing scored)
__global__ void k_minipatch_find(const uint8_t* img, int stride, int W, int H, const uint32_t* corners, const int* lut, int ncorners,
                                 const uint8_t* patches, int n, double* pos, int* found, int* best_out, int range, int max_ssd) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const double px = pos[2 * i], py = pos[2 * i + 1];
  const double tlx = px - range, tly = py - range, brx = px + range, bry = py + range;
  int top = tly; if (top < 0) top = 0; if (top >= H) top = H - 1;   // the pvRowLUT branch of the reference
  const int begin = lut[top];
  const uint8_t* patch = patches + (size_t)i * 81;
  unsigned long long best = ((unsigned long long)(unsigned)(max_ssd + 1) << 32) | 0xffffffffull;
  bool done = false;
  for (int c0 = begin; c0 < ncorners && !done; c0 += 32) {
    const int ci = c0 + lane;
    if (ci < ncorners) {
      const uint32_t cw = corners[ci];
      const int cx = cw & 0xffff, cy = cw >> 16;
      if ((double)cy > bry) done = true;                       // `break`: corners are in raster order, nothing further can match
      else if (!((double)cx < tlx || (double)cx > brx)) {
        int ssd;
        if (!(cx >= 4 && cy >= 4 && cx < W - 4 && cy < H - 4)) ssd = max_ssd + 1;
        else {
          ssd = 0;
          const uint8_t* ip = img + (size_t)(cy - 4) * stride + (cx - 4);
          for (int r = 0; r < 9; r++, ip += stride)
#pragma unroll
            for (int c = 0; c < 9; c++) { const int d = (int)ip[c] - (int)patch[r * 9 + c]; ssd += d * d; }
        }
        const unsigned long long key = ((unsigned long long)(unsigned)ssd << 32) | (unsigned)ci;
        best = key < best ? key : best;
      }
    }
    done = __any_sync(0xffffffffu, done);
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d); best = o < best ? o : best; }
  if (lane == 0) {
    const int ssd = (int)(best >> 32);
    best_out[i] = ssd;
    if (ssd < max_ssd) { const uint32_t cw = corners[(unsigned)best]; pos[2 * i] = (double)(cw & 0xffff); pos[2 * i + 1] = (double)(cw >> 16); found[i] = 1; }
    else found[i] = 0;
  }
}

}  // namespace

// Scratch for one stream's MakeKeyFrame_Rest results (allocated on first use).
int vs_keyframe_rest(vslam_ctx* ctx, int s) {
  { const int rc_ = vs_ensure_lists(ctx); if (rc_) return rc_; }
  if (!ctx->rest_scores) {
    size_t tot = 0; for (int l = 0; l < VS_LEVELS; l++) tot += ctx->lev[l].cap;
    VS_CUDA(cudaMalloc(&ctx->rest_scores, tot * sizeof(int)));
    VS_CUDA(cudaMalloc(&ctx->rest_max, tot * sizeof(uint32_t)));
    VS_CUDA(cudaMalloc(&ctx->rest_cand, tot * sizeof(uint32_t)));
    VS_CUDA(cudaMalloc(&ctx->rest_cand_score, tot * sizeof(double)));
    VS_CUDA(cudaMalloc(&ctx->rest_counts, 2 * VS_LEVELS * sizeof(int)));
  }
  size_t off = 0;
  for (int l = 0; l < VS_LEVELS; l++) {
    const LevelDesc& L = ctx->lev[l];
    int n = 0;
    VS_CUDA(cudaMemcpyAsync(&n, L.lut + (size_t)s * (L.h + 1) + L.h, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    VS_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint8_t* img = l == 0 ? ctx->l0_ptr_host[s] : L.img + (size_t)s * L.h * L.pitch;
    const int stride = l == 0 ? ctx->l0_stride_host[s] : L.pitch;
    const uint32_t* corners = L.corners + (size_t)s * L.cap;
    if (n > 0) { k_fast_scores<<<(n + 255) / 256, 256, 0, ctx->stream>>>(img, stride, corners, n, 10, ctx->rest_scores + off); ctx->launches++; }
    k_nonmax_st<<<1, 512, 0, ctx->stream>>>(img, stride, L.w, L.h, corners, L.lut + (size_t)s * (L.h + 1), n, ctx->rest_scores + off, ctx->rest_max + off,
                                             ctx->rest_cand + off, ctx->rest_cand_score + off, ctx->rest_counts + 2 * l);
    VS_CUDA(cudaGetLastError());
    ctx->launches++;
    ctx->rest_off[l] = off;
    off += L.cap;
  }
  ctx->rest_stream = s;
  return VSLAM_OK;
}

int vs_minipatch_sample(vslam_ctx* ctx, int s, int which, const int* xy_dev, int n, uint8_t* patches_dev) {
  const uint8_t* img = which ? ctx->snap_img + (size_t)s * ctx->lev[0].h * ctx->lev[0].pitch : ctx->l0_ptr_host[s];
  const int stride = which ? ctx->lev[0].pitch : ctx->l0_stride_host[s];
  k_minipatch_sample<<<n, 96, 0, ctx->stream>>>(img, stride, xy_dev, n, patches_dev);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}

int vs_minipatch_find(vslam_ctx* ctx, int s, int which, const uint8_t* patches_dev, int n, double* pos_dev, int* found_dev, int* best_dev, int range, int max_ssd) {
  { const int rc_ = vs_ensure_lists(ctx); if (rc_) return rc_; }
  const LevelDesc& L = ctx->lev[0];
  const uint8_t* img = which ? ctx->snap_img + (size_t)s * L.h * L.pitch : ctx->l0_ptr_host[s];
  const int stride = which ? L.pitch : ctx->l0_stride_host[s];
  const uint32_t* corners = which ? ctx->snap_corners + (size_t)s * L.cap : L.corners + (size_t)s * L.cap;
  const int* lut = which ? ctx->snap_lut + (size_t)s * (L.h + 1) : L.lut + (size_t)s * (L.h + 1);
  int nc = 0;
  VS_CUDA(cudaMemcpyAsync(&nc, lut + L.h, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  k_minipatch_find<<<(n + 3) / 4, 128, 0, ctx->stream>>>(img, stride, L.w, L.h, corners, lut, nc, patches_dev, n, pos_dev, found_dev, best_dev, range, max_ssd);
  VS_CUDA(cudaGetLastError());
  ctx->launches++;
  return VSLAM_OK;
}

// Integer-pipe micro-benchmark (the ZMSSD roofline denominator): back-to-back independent dp4a chains, 8 per thread.
namespace {
__global__ void __launch_bounds__(256) k_dp4a_peak(unsigned* out, int iters, unsigned seed) {
  unsigned a[8], b = seed + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = seed * (k + 1) + blockIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = __dp4a(a[k], b, a[k]);
    b += 0x01010101u;
  }
  unsigned r = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) r ^= a[k];
  if (r == 0x12345678u) out[0] = r;   // keeps the chains alive
}
}  // namespace

// Returns measured dp4a throughput in tera-MACs per second (4 MACs per dp4a lane-instruction) on the current device.
extern "C" int vslam_debug_dp4a_peak(double* tmacs_per_s) {
  if (!tmacs_per_s) return VSLAM_E_INVALID;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return VSLAM_E_CUDA;
  unsigned* out = nullptr;
  if (cudaMalloc(&out, 4) != cudaSuccess) return VSLAM_E_CUDA;
  const int blocks = sms * 8, iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dp4a_peak<<<blocks, 256>>>(out, 64, 1u);   // warm-up
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_dp4a_peak<<<blocks, 256>>>(out, iters, 7u + rep);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double macs = (double)blocks * 256 * iters * 8 * 4;
    const double t = macs / (ms * 1e-3) / 1e12;
    if (t > best) best = t;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  *tmacs_per_s = best;
  return cudaGetLastError() == cudaSuccess ? VSLAM_OK : VSLAM_E_CUDA;
}
