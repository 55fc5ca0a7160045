// Body of k_pose_fast, compiled once per CTA shape (pose_fast.cu includes this file with PF_NS / PF_FT / PF_PP / PF_MINB set).
namespace PF_NS {

constexpr int kFT = PF_FT;         // threads per CTA
constexpr int kPP = PF_PP;         // resident points per thread
constexpr int kMinB = PF_MINB;     // CTAs per SM
constexpr int kCap = kFT * kPP;    // resident points per stream
constexpr int kRadixBits = 11, kBins = 1 << kRadixBits;

struct FastSmem {
  double J[12][kCap + 1];          // si * m26Jacobian, [row * 6 + col][slot]; + 1: the serial mode reads 12 entries of ONE slot at a time (12 different banks)
  union {                          // 8 KB used in turn by
    int hist[kBins];               //   the radix-select histogram,
    int flist[kCap];               //   the found point indices in list order (set-up only: every thread then keeps its indices in registers),
    double wk[kCap];               //   serial mode: the weight of every resident point (0 = rejected)
  };
  double wsum[kFT / 32][28];
  double sums[28];
  double pose[12], mu[6], last[6];
  double sigma;
  double lu[36], inv[36]; int piv[6];
  unsigned long long sel[3];
  int warpcnt[kFT / 32];
  double red[kFT / 32][4];
#ifdef VS_POSE_TIMING
  long long tacc[12], tlast, tsub[16], tsl;
#endif
};
static_assert(kMinB * (sizeof(FastSmem) + 1024) <= 228 * 1024, "kMinB CTAs of k_pose_fast per SM");

#ifdef VS_POSE_TIMING   // instrumented build (scratch experiments): cycles per phase of one CTA, printed by stream 7
#define PF_MARK(k) do { __syncthreads(); if (threadIdx.x == 0) { const long long t_ = clock64(); sm.tacc[k] += t_ - sm.tlast; sm.tlast = t_; } } while (0)
#define PF_SUB(k) do { if (threadIdx.x == 0) { const long long t_ = clock64(); sm.tsub[k] += t_ - sm.tsl; sm.tsl = t_; } } while (0)
#define PF_SUB0() do { if (threadIdx.x == 0) sm.tsl = clock64(); } while (0)
#else
#define PF_MARK(k) do { } while (0)
#define PF_SUB(k) do { } while (0)
#define PF_SUB0() do { } while (0)
#endif

struct Pt { int idx; double im0, im1, f0, f1, si; };

// TrackerData::CalcJacobian (jni/TrackerData.h:107-123): unscaled entries to global memory (the state other entry points read), scaled rows to `Js`
__device__ __forceinline__ void jacobian_rows(const Dev& D, size_t gi, size_t SN, const double* c, const double* dv, double si, double* Js, int jstride) {
  const double invz = 1.0 / c[2];
  const double pos[4] = {c[0], c[1], c[2], 1.0};
#pragma unroll
  for (int m = 0; m < 6; m++) {
    double v4[3] = {0, 0, 0};
    if (m < 3) v4[m] = pos[3];
    else { v4[(m + 1) % 3] = -pos[(m + 2) % 3]; v4[(m + 2) % 3] = pos[(m + 1) % 3]; }
    const double c0 = (v4[0] - c[0] * v4[2] * invz) * invz, c1 = (v4[1] - c[1] * v4[2] * invz) * invz;
    double a0 = dv[0] * c0; a0 += dv[1] * c1;
    double a1 = dv[2] * c0; a1 += dv[3] * c1;
    D.ps.jac[(size_t)m * SN + gi] = a0; D.ps.jac[(size_t)(6 + m) * SN + gi] = a1;
    if (Js) { Js[(size_t)m * jstride] = si * a0; Js[(size_t)(6 + m) * jstride] = si * a1; }
  }
}

// TrackerData::ProjectAndDerivs (jni/TrackerData.h:91-103) of a FOUND point with the world position `w` (already loaded): Project
// (:69-86) with its early returns, then the derivatives `if(bFound)`.  c: v3Cam (always refreshed); P.im: refreshed when Cam.Project ran.
__device__ __forceinline__ void project_and_derivs(const Dev& D, const double* pose, const double* w, int flags, size_t gi, size_t SN, Pt& P, double* c, double* dv, int* quirk) {
  flags &= ~F_INIMAGE;
  se3_apply(pose, w, c);
  D.ps.v3cam[gi] = c[0]; D.ps.v3cam[SN + gi] = c[1]; D.ps.v3cam[2 * SN + gi] = c[2];
  bool projected = false; CamCache cc;
  if (!(c[2] < 0.001)) {
    const double px = c[0] / c[2], py = c[1] / c[2];
    double d = 0; d += px * px; d += py * py;
    if (!(d > D.cam.largestRadius * D.cam.largestRadius)) {
      double im[2]; cam_project(D.cam, px, py, im, cc);
      P.im0 = im[0]; P.im1 = im[1];
      D.ps.v2image[gi] = im[0]; D.ps.v2image[SN + gi] = im[1];
      projected = true;
      if (!cc.invalid && !(im[0] < 0 || im[1] < 0 || im[0] > D.cam.width || im[1] > D.cam.height)) flags |= F_INIMAGE;
    }
  }
  if (projected) { cam_derivs(D.cam, cc, dv); for (int q = 0; q < 4; q++) D.ps.derivs[q * SN + gi] = dv[q]; }
  else { atomicAdd(quirk, 1); for (int q = 0; q < 4; q++) dv[q] = D.ps.derivs[q * SN + gi]; }   // the reference reads another point's camera cache here; the old derivatives are kept
  D.ps.flags[gi] = flags;
}

// 27 per-lane values -> their 27 warp sums, sum q delivered to exactly one lane (returned with its index, -1 for lanes without one).
// Each step halves what a lane holds: 14 + 7 + 4 + 2 + 1 = 28 shuffles of doubles instead of 27 x 5.
__device__ __forceinline__ double warp_transpose_reduce27(const double (&a)[27], int lane, int& index) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
  double v14[14];
#pragma unroll
  for (int i = 0; i < 14; i++) {
    const double hi = (14 + i < 27) ? a[(14 + i < 27) ? 14 + i : 0] : 0.0;
    const double keep = b4 ? hi : a[i], give = b4 ? a[i] : hi;
    v14[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
  }
  double v7[7];
#pragma unroll
  for (int i = 0; i < 7; i++) { const double keep = b3 ? v14[7 + i] : v14[i], give = b3 ? v14[i] : v14[7 + i]; v7[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8); }
  double v4[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { const double hi = (4 + i < 7) ? v7[(4 + i < 7) ? 4 + i : 0] : 0.0; const double keep = b2 ? hi : v7[i], give = b2 ? v7[i] : hi; v4[i] = keep + __shfl_xor_sync(0xffffffffu, give, 4); }
  double v2[2];
#pragma unroll
  for (int i = 0; i < 2; i++) { const double keep = b1 ? v4[2 + i] : v4[i], give = b1 ? v4[i] : v4[2 + i]; v2[i] = keep + __shfl_xor_sync(0xffffffffu, give, 2); }
  const double keep = b0 ? v2[1] : v2[0], give = b0 ? v2[0] : v2[1];
  const double r = keep + __shfl_xor_sync(0xffffffffu, give, 1);
  const int r3 = (b1 ? 2 : 0) + (b0 ? 1 : 0), r2 = (b2 ? 4 : 0) + r3, r1 = (b3 ? 7 : 0) + r2;
  const bool valid = (!b2 || r3 <= 2) && (!b4 || r1 <= 12);
  index = valid ? (b4 ? 14 : 0) + r1 : -1;
  return r;
}

// One measurement row into the normal equations (jni/myWLS.h:39-50) with multiply-add contraction (the sums are a parallel reduction,
// whose order differs from the reference's serial loop anyway)
__device__ __forceinline__ void add_row(double (&acc)[27], const double* J, double m, double w) {
  int q = 0;
#pragma unroll
  for (int r = 0; r < 6; r++) {
    const double Jw = w * J[r];
    acc[21 + r] = __fma_rn(m, Jw, acc[21 + r]);
#pragma unroll
    for (int c = r; c < 6; c++) { acc[q] = __fma_rn(Jw, J[c], acc[q]); q++; }
  }
}

// k-th smallest (0-based) of the n squared errors: keys of the resident points in registers (key[p] of slot tid + p * kFT), the rest in
// spill[0 .. n - kCap).  MSB-first radix select on the IEEE bit patterns (monotonic for values >= 0; the sign bit is skipped), 11 bits per
// pass -- the first digit is the exponent -- stopping as soon as the selected bin holds a single element.  All threads call it.
// Per pass: equal digits inside a warp are combined with match.any before the shared-memory atomic (the keys of a pass share a handful of
// exponents: up to 60 lanes per address otherwise), and warp 0 finds the bin in two parallel steps (64 bins per lane, then 2 per lane).
__device__ __forceinline__ void radix_count(FastSmem& sm, bool part, int bin, int lane) {
  const unsigned peers = __match_any_sync(0xffffffffu, part ? bin : kBins + lane);     // non-participants: a group of their own
  if (part && lane == __ffs(peers) - 1) atomicAdd(&sm.hist[bin], __popc(peers));
  __syncwarp();   // reconverge before the next warp collective (a diverged warp takes the slow WARPSYNC.COLLECTIVE path, ~300 cycles per collective)
}
__device__ __forceinline__ double radix_select(FastSmem& sm, const double (&key)[kPP], int n, const double* spill, int k) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long prefix = 0;
  int rank = k, sh = 63;
  {
    // First level: 64 buckets by exponent (2^-63 .. 2^0, both ends clamped -- a monotone map, so bucket order is key order), counted in
    // WARP-PRIVATE histograms: the squared errors of a frame span a few dozen binades, i.e. a few dozen heavily shared counters, and
    // shared-memory atomics of different warps on one address serialise.  An interior bucket is one exponent = an 11-bit prefix, and the
    // select continues on the mantissa below; a clamped end bucket (median < 2^-63 or >= 1 px^2) restarts with the generic loop.
    constexpr int kEB = 64, kE0 = 1023 - 63;
    for (int b = tid; b < (kFT / 32) * kEB; b += kFT) sm.hist[b] = 0;
    __syncthreads();
    int* wh = sm.hist + warp * kEB;
#pragma unroll
    for (int p = 0; p < kPP; p++) {
      const int ex = (int)((unsigned long long)__double_as_longlong(key[p]) >> 52) - kE0;
      const int bkt = ex < 0 ? 0 : (ex > kEB - 1 ? kEB - 1 : ex);
      const bool part = tid + p * kFT < n;
      const unsigned peers = __match_any_sync(0xffffffffu, part ? bkt : kEB + lane);
      if (part && lane == __ffs(peers) - 1) wh[bkt] += __popc(peers);          // (one lane per distinct bucket of the warp: no atomic needed)
      __syncwarp();
    }
    for (int t0 = kCap; t0 < n; t0 += kFT) {
      const int t = t0 + tid;
      const int ex = t < n ? (int)((unsigned long long)__double_as_longlong(spill[t - kCap]) >> 52) - kE0 : 0;
      const int bkt = ex < 0 ? 0 : (ex > kEB - 1 ? kEB - 1 : ex);
      const unsigned peers = __match_any_sync(0xffffffffu, t < n ? bkt : kEB + lane);
      if (t < n && lane == __ffs(peers) - 1) wh[bkt] += __popc(peers);
      __syncwarp();
    }
    __syncthreads();
    if (tid < 32) {
      int v0 = 0, v1 = 0;
#pragma unroll
      for (int w = 0; w < kFT / 32; w++) { v0 += sm.hist[w * kEB + 2 * lane]; v1 += sm.hist[w * kEB + 2 * lane + 1]; }
      const int two = v0 + v1;
      int incl = two;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
      const int excl = incl - two;
      if (rank >= excl && rank < incl) {
        int r = rank - excl, bkt = 2 * lane, cnt = v0;
        if (r >= v0) { r -= v0; bkt++; cnt = v1; }
        sm.sel[0] = (unsigned long long)bkt; sm.sel[1] = (unsigned long long)r; sm.sel[2] = (unsigned long long)cnt;
      }
    }
    __syncthreads();
    const int bkt = (int)sm.sel[0];
    if (bkt > 0 && bkt < kEB - 1) {
      prefix = (unsigned long long)(bkt + kE0) << 52; rank = (int)sm.sel[1]; sh = 52;
      if (sm.sel[2] == 1ull) {   // a single key with this exponent: its owner publishes it
        __syncthreads();
#pragma unroll
        for (int p = 0; p < kPP; p++)
          if (tid + p * kFT < n) { const unsigned long long kk = (unsigned long long)__double_as_longlong(key[p]); if ((kk >> 52) == (prefix >> 52)) sm.sel[0] = kk; }
        for (int t = kCap + tid; t < n; t += kFT) { const unsigned long long kk = (unsigned long long)__double_as_longlong(spill[t - kCap]); if ((kk >> 52) == (prefix >> 52)) sm.sel[0] = kk; }
        __syncthreads();
        prefix = sm.sel[0]; sh = 0;
      }
    }
    __syncthreads();
  }
  while (sh > 0) {
    const int bits = sh >= kRadixBits ? kRadixBits : sh, nsh = sh - bits, nb = 1 << bits;
    PF_SUB0();
    for (int b = tid; b < nb; b += kFT) sm.hist[b] = 0;
    __syncthreads();
    PF_SUB(0);
#pragma unroll
    for (int p = 0; p < kPP; p++) {
      const unsigned long long kk = (unsigned long long)__double_as_longlong(key[p]);
      const bool part = (tid + p * kFT < n) && (sh == 63 || (kk >> sh) == (prefix >> sh));
      if (sh == 63) radix_count(sm, part, (int)((kk >> nsh) & (unsigned long long)(nb - 1)), lane);   // every key takes part, few distinct digits: combine equal ones first
      else if (part) atomicAdd(&sm.hist[(int)((kk >> nsh) & (unsigned long long)(nb - 1))], 1);          // one exponent's keys over 2048 bins: hardly any sharing
    }
    for (int t0 = kCap; t0 < n; t0 += kFT) {
      const int t = t0 + tid;
      const unsigned long long kk = t < n ? (unsigned long long)__double_as_longlong(spill[t - kCap]) : 0ull;
      const bool part = t < n && (sh == 63 || (kk >> sh) == (prefix >> sh));
      if (sh == 63) radix_count(sm, part, (int)((kk >> nsh) & (unsigned long long)(nb - 1)), lane);
      else if (part) atomicAdd(&sm.hist[(int)((kk >> nsh) & (unsigned long long)(nb - 1))], 1);
    }
    __syncthreads();
    PF_SUB(1);
    if (tid < 32) {
      __syncwarp();
      const int per = nb >> 5;            // bins per lane in the first step (nb >= 64 for every pass: 11, 11, 11, 11, 11, 8 bits)
      int mine = 0;
      if (per == 64) {   // the common case, unrolled: 64 independent loads in flight instead of a chain of load -> add
        int part4[4] = {0, 0, 0, 0};
#pragma unroll
        for (int q = 0; q < 64; q++) part4[q & 3] += sm.hist[lane * 64 + ((q + lane) & 63)];   // rotated start per lane: no bank conflicts
        mine = (part4[0] + part4[1]) + (part4[2] + part4[3]);
      } else for (int q = 0; q < per; q++) mine += sm.hist[lane * per + ((q + lane) & (per - 1))];
      int incl = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
      const int excl = incl - mine;
      const unsigned hit = __ballot_sync(0xffffffffu, rank >= excl && rank < incl);
      const int L = __ffs(hit) - 1;                                   // the lane whose bins hold the rank
      int r = rank - __shfl_sync(0xffffffffu, excl, L);               // rank inside that lane's bins
      // second step: the `per` bins of lane L, per / 32 of them to each lane (per = 64 -> 2, per = 8 -> the first 8 lanes take 1)
      const int sub = per >= 32 ? per >> 5 : 1;
      int v0 = 0, v1 = 0;
      if (lane * sub < per) { v0 = sm.hist[L * per + lane * sub]; if (sub == 2) v1 = sm.hist[L * per + lane * sub + 1]; }
      __syncwarp();
      const int two = v0 + v1;
      int inc2 = two;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc2, d); if (lane >= d) inc2 += v; }
      const int exc2 = inc2 - two;
      if (r >= exc2 && r < inc2) {
        r -= exc2;
        int bin = L * per + lane * sub, cnt = v0;
        if (r >= v0) { r -= v0; bin++; cnt = v1; }
        sm.sel[0] = prefix | ((unsigned long long)bin << nsh); sm.sel[1] = (unsigned long long)r; sm.sel[2] = (unsigned long long)cnt;
      }
    }
    __syncthreads();
    PF_SUB(2);
    prefix = sm.sel[0]; rank = (int)sm.sel[1];
    const bool single = sm.sel[2] == 1ull;
    sh = nsh;
#ifdef VS_POSE_TIMING
    if (threadIdx.x == 0) sm.tsub[15]++;
#endif
    if (single && sh > 0) {   // exactly one element carries this prefix: its owner publishes it
      __syncthreads();
#pragma unroll
      for (int p = 0; p < kPP; p++) {
        if (tid + p * kFT < n) { const unsigned long long kk = (unsigned long long)__double_as_longlong(key[p]); if ((kk >> sh) == (prefix >> sh)) sm.sel[0] = kk; }
      }
      for (int t = kCap + tid; t < n; t += kFT) { const unsigned long long kk = (unsigned long long)__double_as_longlong(spill[t - kCap]); if ((kk >> sh) == (prefix >> sh)) sm.sel[0] = kk; }
      __syncthreads();
      prefix = sm.sel[0];
      PF_SUB(3);
      break;
    }
  }
  __syncthreads();    // sel / hist may be rewritten by the caller
  return __longlong_as_double((long long)prefix);
}

// mode 1: coarse stage, 2: fine stage (+ scene depth; + motion model / quality if tail)
__global__ void __launch_bounds__(kFT, kMinB) k_pose_fast(Dev D, int mode, int tail, int serial, int sel) {
  cudaGridDependencySynchronize(); cudaTriggerProgrammaticLaunchCompletion();   // programmatic dependent launch (vs_launch_pdl); no-ops otherwise
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FastSmem& sm = *reinterpret_cast<FastSmem*>(smem_raw);
  const int s = blockIdx.x + D.s0, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  StreamState* st = D.ss + s;
  if (st->lost_frames >= 3 && !st->recovered) return;
  if (mode == 1 && !st->try_coarse) return;
  if (other_chain(D, st)) return;
  const size_t SN = (size_t)D.S * D.N, so = (size_t)s * D.N;
  const int* list = D.lists + (size_t)s * D.list_cap;
  const int nlist = (mode == 2) ? st->nA + st->nB : st->nA;
  if ((sel == 1 && nlist > 1024) || (sel == 2 && nlist <= 1024)) return;   // the other CTA shape's stream (vs_launch_pose_fast): lists that fit 1024 resident points or not
  int* gfl = D.pvs + (size_t)s * VS_LEVELS * D.N;                 // all found indices, list order (the first kCap also in shared memory during set-up)
  double* spill_key = D.sort_scratch + (size_t)s * D.sort_cap;    // squared errors of the found points beyond kCap
  if (tid < 12) sm.pose[tid] = st->pose[tid];
  if (tid < 6) sm.last[tid] = 0.0;

#ifdef VS_POSE_TIMING
  if (tid < 12) sm.tacc[tid] = 0;
  if (tid < 16) sm.tsub[tid] = 0;
  if (tid == 0) sm.tlast = clock64();
  __syncthreads();
#endif
  // ---- the found entries of the list (fixed for the whole stage) in ASCENDING POINT INDEX: a bitmap of the set, then an ordered expansion,
  // so that consecutive threads touch consecutive addresses of the per-point SoA arrays (the order of the set does not matter to the
  // parallel sums or the median).  The bitmap lives in the Jacobian block, which is not in use yet.
  unsigned* bitmap = reinterpret_cast<unsigned*>(&sm.J[0][0]);
  int* wordoff = reinterpret_cast<int*>(bitmap + 2048);            // serial mode: found points below each bitmap word
  const int words = (D.map.n + 31) >> 5;                          // <= 2048 (max_points <= 65536)
  for (int w = tid; w < words; w += kFT) bitmap[w] = 0u;
  __syncthreads();
  for (int k = tid; k < nlist; k += kFT) {
    const int idx = list[k];
    if (D.ps.flags[so + idx] & F_FOUND) atomicOr(&bitmap[idx >> 5], 1u << (idx & 31));
  }
  __syncthreads();
  int n = 0;
  for (int w0 = 0; w0 < words; w0 += kFT) {
    const int w = w0 + tid;
    unsigned m = w < words ? bitmap[w] : 0u;
    const int mine = __popc(m);
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    if (lane == 31) sm.warpcnt[warp] = incl;
    __syncthreads();
    int off = n, tot = 0;
#pragma unroll
    for (int q = 0; q < kFT / 32; q++) { const int c = sm.warpcnt[q]; if (q < warp) off += c; tot += c; }
    int pos = off + incl - mine;
    if (w < words) wordoff[w] = pos;
    while (m) { const int b = __ffs(m) - 1; m &= m - 1; const int idx = (w << 5) + b; if (pos < kCap) sm.flist[pos] = idx; gfl[pos] = idx; pos++; }
    n += tot;
    __syncthreads();
  }
  int* gorder = gfl + D.N;                                         // serial mode: slot of the k-th found entry in LIST order
  if (serial) {
    int cnt = 0;
    for (int base = 0; base < nlist; base += kFT) {
      const int k = base + tid;
      int idx = 0; bool fnd = false;
      if (k < nlist) { idx = list[k]; fnd = (bitmap[idx >> 5] >> (idx & 31)) & 1u; }
      const unsigned bal = __ballot_sync(0xffffffffu, fnd);
      if (lane == 0) sm.warpcnt[warp] = __popc(bal);
      __syncthreads();
      int off = cnt, tot = 0;
#pragma unroll
      for (int q = 0; q < kFT / 32; q++) { const int c = sm.warpcnt[q]; if (q < warp) off += c; tot += c; }
      if (fnd) gorder[off + __popc(bal & ((1u << lane) - 1u))] = wordoff[idx >> 5] + __popc(bitmap[idx >> 5] & ((1u << (idx & 31)) - 1u));
      cnt += tot;
      __syncthreads();
    }
  }
  if (mode == 1) {   // coarse stage (jni/Tracker.cc:464-489): needs nFound >= CoarseMin
    if ((unsigned)n < D.prm.coarse_min) return;
    if (tid == 0) st->did_coarse = 1;
  }
  const int nres = n < kCap ? n : kCap;
  PF_MARK(0);

  // ---- iteration 0 point pass: the projection of k_project_lists / k_reproject_fine stands; Jacobians from v3Cam and the derivatives
  Pt P[kPP]; double e2[kPP];
  {
    double c[kPP][3], dv[kPP][4];
#pragma unroll
    for (int p = 0; p < kPP; p++) {
      const int slot = tid + p * kFT;
      P[p].idx = slot < nres ? sm.flist[slot] : 0;
      const size_t gi = so + P[p].idx;
      P[p].im0 = D.ps.v2image[gi]; P[p].im1 = D.ps.v2image[SN + gi];
      P[p].f0 = D.ps.v2found[gi]; P[p].f1 = D.ps.v2found[SN + gi];
      P[p].si = D.ps.sqrtinv[gi];
      c[p][0] = D.ps.v3cam[gi]; c[p][1] = D.ps.v3cam[SN + gi]; c[p][2] = D.ps.v3cam[2 * SN + gi];
#pragma unroll
      for (int q = 0; q < 4; q++) dv[p][q] = D.ps.derivs[q * SN + gi];
    }
#pragma unroll
    for (int p = 0; p < kPP; p++) {
      const int slot = tid + p * kFT;
      e2[p] = 0.0;
      if (slot < nres) {
        const size_t gi = so + P[p].idx;
        jacobian_rows(D, gi, SN, c[p], dv[p], P[p].si, &sm.J[0][slot], kCap + 1);
        const double e0 = (P[p].f0 - P[p].im0) * P[p].si, e1 = (P[p].f1 - P[p].im1) * P[p].si;
        D.ps.err[gi] = e0; D.ps.err[SN + gi] = e1;
        double q2 = 0; q2 += e0 * e0; q2 += e1 * e1; e2[p] = q2;
      }
    }
    for (int k = kCap + tid; k < n; k += kFT) {   // spill path
      const size_t gi = so + gfl[k];
      const double cc[3] = {D.ps.v3cam[gi], D.ps.v3cam[SN + gi], D.ps.v3cam[2 * SN + gi]};
      const double dd[4] = {D.ps.derivs[gi], D.ps.derivs[SN + gi], D.ps.derivs[2 * SN + gi], D.ps.derivs[3 * SN + gi]};
      const double si = D.ps.sqrtinv[gi];
      jacobian_rows(D, gi, SN, cc, dd, si, nullptr, 0);
      const double e0 = (D.ps.v2found[gi] - D.ps.v2image[gi]) * si, e1 = (D.ps.v2found[SN + gi] - D.ps.v2image[SN + gi]) * si;
      D.ps.err[gi] = e0; D.ps.err[SN + gi] = e1;
      double q2 = 0; q2 += e0 * e0; q2 += e1 * e1; spill_key[k - kCap] = q2;
    }
  }
  __syncthreads();
  PF_MARK(1);

  for (int iter = 0; iter < 10; iter++) {
    const bool nonlinear = (mode == 1) || iter == 0 || iter == 4 || iter == 9;
    const double ov = (iter > 5) ? (mode == 1 ? 1.0 : 16.0) : 0.0;
    const bool mark = (mode == 2 && iter == 9);
    if (iter > 0) {
      if (nonlinear) {   // ProjectAndDerivs with the current pose + CalcJacobian
        double w[kPP][3]; int fl[kPP];
#pragma unroll
        for (int p = 0; p < kPP; p++) { const double* wp = D.map.world + 3 * (size_t)P[p].idx; w[p][0] = wp[0]; w[p][1] = wp[1]; w[p][2] = wp[2]; fl[p] = D.ps.flags[so + P[p].idx]; }
#pragma unroll
        for (int p = 0; p < kPP; p++) {
          const int slot = tid + p * kFT;
          if (slot < nres) {
            const size_t gi = so + P[p].idx;
            double c[3], dv[4];
            project_and_derivs(D, sm.pose, w[p], fl[p], gi, SN, P[p], c, dv, &st->quirk_stale_cache);
            jacobian_rows(D, gi, SN, c, dv, P[p].si, &sm.J[0][slot], kCap + 1);
            const double e0 = (P[p].f0 - P[p].im0) * P[p].si, e1 = (P[p].f1 - P[p].im1) * P[p].si;
            D.ps.err[gi] = e0; D.ps.err[SN + gi] = e1;
            double q2 = 0; q2 += e0 * e0; q2 += e1 * e1; e2[p] = q2;
          }
        }
        for (int k = kCap + tid; k < n; k += kFT) {
          const int i = gfl[k]; const size_t gi = so + i;
          Pt Q; Q.idx = i; Q.im0 = D.ps.v2image[gi]; Q.im1 = D.ps.v2image[SN + gi]; Q.f0 = D.ps.v2found[gi]; Q.f1 = D.ps.v2found[SN + gi]; Q.si = D.ps.sqrtinv[gi];
          const double* wp = D.map.world + 3 * (size_t)i; const double ww[3] = {wp[0], wp[1], wp[2]};
          double c[3], dv[4];
          project_and_derivs(D, sm.pose, ww, D.ps.flags[gi], gi, SN, Q, c, dv, &st->quirk_stale_cache);
          jacobian_rows(D, gi, SN, c, dv, Q.si, nullptr, 0);
          const double e0 = (Q.f0 - Q.im0) * Q.si, e1 = (Q.f1 - Q.im1) * Q.si;
          D.ps.err[gi] = e0; D.ps.err[SN + gi] = e1;
          double q2 = 0; q2 += e0 * e0; q2 += e1 * e1; spill_key[k - kCap] = q2;
        }
      } else {           // TrackerData::LinearUpdate (jni/TrackerData.h:126-132) with the last update
        double v6[6];
#pragma unroll
        for (int q = 0; q < 6; q++) v6[q] = sm.last[q];
#pragma unroll
        for (int p = 0; p < kPP; p++) {
          const int slot = tid + p * kFT;
          if (slot < nres) {
            const size_t gi = so + P[p].idx;
            // (si J).v = si (J.v) exactly (si is a power of two), so the unscaled product is recovered exactly by the division
            double a0 = sm.J[0][slot] * v6[0], a1 = sm.J[6][slot] * v6[0];
#pragma unroll
            for (int q = 1; q < 6; q++) { a0 += sm.J[q][slot] * v6[q]; a1 += sm.J[6 + q][slot] * v6[q]; }
            // 1 / si, si = 2^-level: the reciprocal of a power of two is the power of two with the mirrored exponent (exact, no division)
            const double sc = __longlong_as_double((long long)(2046ull - (((unsigned long long)__double_as_longlong(P[p].si) >> 52) & 0x7ffull)) << 52);
            P[p].im0 += a0 * sc; P[p].im1 += a1 * sc;
            D.ps.v2image[gi] = P[p].im0; D.ps.v2image[SN + gi] = P[p].im1;
            const double e0 = (P[p].f0 - P[p].im0) * P[p].si, e1 = (P[p].f1 - P[p].im1) * P[p].si;
            D.ps.err[gi] = e0; D.ps.err[SN + gi] = e1;
            double q2 = 0; q2 += e0 * e0; q2 += e1 * e1; e2[p] = q2;
          }
        }
        for (int k = kCap + tid; k < n; k += kFT) {
          const size_t gi = so + gfl[k];
          double a0 = D.ps.jac[gi] * v6[0], a1 = D.ps.jac[(size_t)6 * SN + gi] * v6[0];
#pragma unroll
          for (int q = 1; q < 6; q++) { a0 += D.ps.jac[(size_t)q * SN + gi] * v6[q]; a1 += D.ps.jac[(size_t)(6 + q) * SN + gi] * v6[q]; }
          const double im0 = D.ps.v2image[gi] + a0, im1 = D.ps.v2image[SN + gi] + a1, si = D.ps.sqrtinv[gi];
          D.ps.v2image[gi] = im0; D.ps.v2image[SN + gi] = im1;
          const double e0 = (D.ps.v2found[gi] - im0) * si, e1 = (D.ps.v2found[SN + gi] - im1) * si;
          D.ps.err[gi] = e0; D.ps.err[SN + gi] = e1;
          double q2 = 0; q2 += e0 * e0; q2 += e1 * e1; spill_key[k - kCap] = q2;
        }
      }
      if (n > kCap) __syncthreads();   // spill keys are read by other threads in the median
      PF_MARK(nonlinear ? 2 : 3);
    }

    // ---- CalcPoseUpdate (jni/Tracker.cc:683-774)
    if (n == 0) { if (tid < 6) sm.mu[tid] = 0.0; if (tid == 0) sm.sigma = 0.0; __syncthreads(); }
    else {
      if (ov > 0) { if (tid == 0) sm.sigma = ov; }
      else {   // Tukey::FindSigmaSquared (jni/MEstimator.h:67-77): the sort there only serves to pick v[n/2]
        const double med = radix_select(sm, e2, n, spill_key, n / 2);
        if (tid == 0) {
          const unsigned long long den = (unsigned long long)n * 2ull - 6ull;   // size_t arithmetic of the reference
          double sigma = 1.4826 * (1 + 5.0 / (double)den) * sqrt(med);
          sigma = 4.6851 * sigma;
          sm.sigma = sigma * sigma;
        }
      }
      __syncthreads();
      PF_MARK(4);
      const double sig2 = sm.sigma;
      PF_SUB0();
      if (!serial) {
        double acc[27];
#pragma unroll
        for (int k = 0; k < 27; k++) acc[k] = 0.0;
        // the Tukey weights of the thread's points first: kPP independent divisions in flight instead of one per point between its rows
        double wt[kPP];
#pragma unroll
        for (int p = 0; p < kPP; p++) { const double sq = (e2[p] > sig2) ? 0.0 : 1.0 - (e2[p] / sig2); wt[p] = sq * sq; }
#pragma unroll
        for (int p = 0; p < kPP; p++) {
          const int slot = tid + p * kFT;
          if (slot < nres) {
            const double w = wt[p];
            if (w == 0.0) { if (mark) D.ps.counts[so + P[p].idx]++; }
            else {
              if (mark) D.ps.counts[SN + so + P[p].idx]++;
              const double e0 = (P[p].f0 - P[p].im0) * P[p].si, e1 = (P[p].f1 - P[p].im1) * P[p].si;
              double Jr[6];
#pragma unroll
              for (int q = 0; q < 6; q++) Jr[q] = sm.J[q][slot];
              add_row(acc, Jr, D.truncate ? (double)(int)e0 : e0, w);      // (int) cast of jni/Tracker.cc:766-767
#pragma unroll
              for (int q = 0; q < 6; q++) Jr[q] = sm.J[6 + q][slot];
              add_row(acc, Jr, D.truncate ? (double)(int)e1 : e1, w);
            }
          }
        }
        for (int k = kCap + tid; k < n; k += kFT) {
          const size_t gi = so + gfl[k];
          const double e0 = D.ps.err[gi], e1 = D.ps.err[SN + gi], si = D.ps.sqrtinv[gi];
          double q2 = 0; q2 += e0 * e0; q2 += e1 * e1;
          const double sq = (q2 > sig2) ? 0.0 : 1.0 - (q2 / sig2);
          const double w = sq * sq;
          if (w == 0.0) { if (mark) D.ps.counts[gi]++; continue; }
          if (mark) D.ps.counts[SN + gi]++;
          double Jr[6];
#pragma unroll
          for (int q = 0; q < 6; q++) Jr[q] = si * D.ps.jac[(size_t)q * SN + gi];
          add_row(acc, Jr, D.truncate ? (double)(int)e0 : e0, w);
#pragma unroll
          for (int q = 0; q < 6; q++) Jr[q] = si * D.ps.jac[(size_t)(6 + q) * SN + gi];
          add_row(acc, Jr, D.truncate ? (double)(int)e1 : e1, w);
        }
        PF_SUB(4);
        __syncwarp();   // reconverge: shuffles issued by a diverged warp take the slow WARPSYNC.COLLECTIVE path
        int qi; const double r = warp_transpose_reduce27(acc, lane, qi);
        if (qi >= 0) sm.wsum[warp][qi] = r;
        PF_SUB(5);
        __syncthreads();
        PF_SUB(6);
        if (tid < 27) { double v = sm.wsum[0][tid]; for (int q = 1; q < kFT / 32; q++) v += sm.wsum[q][tid]; sm.sums[tid] = v; }
      } else {
        // the reference's order of operations: points in list order, row 0 then row 1, C(r,c) += (w J_r) J_c and b(r) += m (w J_r), no contraction
#pragma unroll
        for (int p = 0; p < kPP; p++) {
          const int slot = tid + p * kFT;
          if (slot < nres) {
            const double sq = (e2[p] > sig2) ? 0.0 : 1.0 - (e2[p] / sig2);
            const double w = sq * sq;
            sm.wk[slot] = w;
            if (w == 0.0) { if (mark) D.ps.counts[so + P[p].idx]++; } else if (mark) D.ps.counts[SN + so + P[p].idx]++;
          }
        }
        for (int k = kCap + tid; k < n; k += kFT) {   // inlier / outlier accounting of the spill points (their weights are recomputed below)
          if (!mark) break;
          const size_t gi = so + gfl[k];
          const double e0 = D.ps.err[gi], e1 = D.ps.err[SN + gi];
          double q2 = 0; q2 += e0 * e0; q2 += e1 * e1;
          const double sq = (q2 > sig2) ? 0.0 : 1.0 - (q2 / sig2);
          if (sq * sq == 0.0) D.ps.counts[gi]++; else D.ps.counts[SN + gi]++;
        }
        __syncthreads();
        if (warp == 0) {
          // lane q < 21: C(r,c) of the packed upper triangle; lane 21 + r: b(r).  Points arrive in chunks of 32 (weight and residuals of point
          // k0 + lane loaded by lane `lane`, the next chunk requested while the current one is summed) and are broadcast by shuffles.
          int r = 0, c = -1;
          if (lane < 21) { int q = lane; while (q >= 6 - r) { q -= 6 - r; r++; } c = r + q; } else if (lane < 27) r = lane - 21;   // (lanes 27..31 shadow b(0): no divergence inside the loop)
          __syncwarp();
          double a = 0.0;
          auto fetch = [&](int k, double& w, double& m0, double& m1, int& slot) {
            w = 0.0; m0 = 0.0; m1 = 0.0; slot = 0;
            if (k < n) {
              slot = gorder[k];
              const size_t gi = so + gfl[slot];
              const double e0 = D.ps.err[gi], e1 = D.ps.err[SN + gi];
              if (slot < kCap) w = sm.wk[slot];
              else { double q2 = 0; q2 += e0 * e0; q2 += e1 * e1; const double sq = (q2 > sig2) ? 0.0 : 1.0 - (q2 / sig2); w = sq * sq; }
              m0 = D.truncate ? (double)(int)e0 : e0; m1 = D.truncate ? (double)(int)e1 : e1;      // (int) cast of jni/Tracker.cc:766-767
            }
          };
          double nw, nm0, nm1; int nslot; fetch(lane, nw, nm0, nm1, nslot);
          for (int k0 = 0; k0 < n; k0 += 32) {
            const double w = nw, m0 = nm0, m1 = nm1; const int slot = nslot;
            fetch(k0 + 32 + lane, nw, nm0, nm1, nslot);
            __syncwarp();
            const int cnt = n - k0 < 32 ? n - k0 : 32;
            for (int j = 0; j < cnt; j++) {
              const double wj = __shfl_sync(0xffffffffu, w, j);
              if (wj == 0.0) continue;
              const double m0j = __shfl_sync(0xffffffffu, m0, j), m1j = __shfl_sync(0xffffffffu, m1, j);
              const int k = __shfl_sync(0xffffffffu, slot, j);
              {
                double J0r, J1r, J0c = 0, J1c = 0;
                if (k < kCap) { J0r = sm.J[r][k]; J1r = sm.J[6 + r][k]; if (c >= 0) { J0c = sm.J[c][k]; J1c = sm.J[6 + c][k]; } }
                else {
                  const size_t gi = so + gfl[k]; const double si = D.ps.sqrtinv[gi];
                  J0r = si * D.ps.jac[(size_t)r * SN + gi]; J1r = si * D.ps.jac[(size_t)(6 + r) * SN + gi];
                  if (c >= 0) { J0c = si * D.ps.jac[(size_t)c * SN + gi]; J1c = si * D.ps.jac[(size_t)(6 + c) * SN + gi]; }
                }
                const double Jw0 = wj * J0r, Jw1 = wj * J1r;
                if (c >= 0) { a += Jw0 * J0c; a += Jw1 * J1c; } else { a += m0j * Jw0; a += m1j * Jw1; }
              }
            }
          }
          if (lane < 27) sm.sums[lane] = a;
        }
      }
      __syncthreads();
      PF_MARK(5);
      if (warp == 0) 
#ifdef VS_POSE_TIMING
        wls_solve_warp_fast(sm.sums, sm.lu, sm.inv, sm.piv, sm.mu, &sm.tsub[7]);
#else
        wls_solve_warp_fast(sm.sums, sm.lu, sm.inv, sm.piv, sm.mu);
#endif

      __syncthreads();
      PF_MARK(6);
    }
    if (tid == 0) {
      double e[12], np[12]; se3_exp(sm.mu, e); se3_mul(e, sm.pose, np);
      for (int k = 0; k < 12; k++) sm.pose[k] = np[k];
      for (int k = 0; k < 6; k++) sm.last[k] = sm.mu[k];
      const int u = st->n_updates;
      if (u < VS_MAX_UPDATES) { for (int k = 0; k < 6; k++) st->updates[6 * u + k] = sm.mu[k]; st->sigmas[u] = sm.sigma; st->n_updates = u + 1; }
    }
    __syncthreads();
    PF_MARK(7);
  }
#ifdef VS_POSE_TIMING
  if (tid == 0 && s == 7 && mode == 2) printf("k_pose_fast cycles: found_list %lld setup %lld nonlinear %lld linear %lld median %lld accum %lld solve %lld exp %lld (n=%d)\n", sm.tacc[0], sm.tacc[1], sm.tacc[2], sm.tacc[3], sm.tacc[4], sm.tacc[5], sm.tacc[6], sm.tacc[7], n);
  if (tid == 0 && s == 7 && mode == 2) printf("   sub: radix clear %lld count %lld scan %lld fetch %lld passes %lld | acc rows %lld shuffle %lld sync %lld | solve elim %lld subst %lld\n", sm.tsub[0], sm.tsub[1], sm.tsub[2], sm.tsub[3], sm.tsub[15], sm.tsub[4], sm.tsub[5], sm.tsub[6], sm.tsub[7], sm.tsub[8]);
#endif
  if (tid < 12) st->pose[tid] = sm.pose[tid];
  if (mode != 2) return;

  // scene depth from the tracked features (jni/Tracker.cc:610-625); fixed-shape parallel sums
  {
    double a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
    for (int p = 0; p < kPP; p++) if (tid + p * kFT < nres) { const double z = D.ps.v3cam[2 * SN + so + P[p].idx]; a0 += z; a1 += z * z; a2 += 1.0; }
    for (int k = kCap + tid; k < n; k += kFT) { const double z = D.ps.v3cam[2 * SN + so + gfl[k]]; a0 += z; a1 += z * z; a2 += 1.0; }
    __syncwarp();
#pragma unroll
    for (int d = 16; d; d >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, d); a1 += __shfl_xor_sync(0xffffffffu, a1, d); a2 += __shfl_xor_sync(0xffffffffu, a2, d); }
    if (lane == 0) { sm.red[warp][0] = a0; sm.red[warp][1] = a1; sm.red[warp][2] = a2; }
  }
  __syncthreads();
  if (tid == 0) {
    double dSum = 0, dSumSq = 0, dNum = 0;
    for (int w = 0; w < kFT / 32; w++) { dSum += sm.red[w][0]; dSumSq += sm.red[w][1]; dNum += sm.red[w][2]; }
    pose_stage_tail(D, st, s, sm.pose, dSum, dSumSq, dNum, tail);
  }
}

}  // namespace PF_NS
