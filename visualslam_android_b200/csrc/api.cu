// C-ABI of the B200 tracking front-end (include/vslam_b200.h): context, memory layout in HBM, readers.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>
#include "vslam_internal.cuh"
#include <cstdlib>

static std::string g_create_error;

namespace {

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// cudaMemset / cudaMemset2D run on the legacy default stream and may return before the device has done them; the context's streams are
// non-blocking and do not wait for that stream, so every clear made from the host is followed by a wait for it.
cudaError_t memset_sync(void* p, int v, size_t n) { const cudaError_t e = cudaMemset(p, v, n); return e == cudaSuccess ? cudaStreamSynchronize(cudaStreamLegacy) : e; }
cudaError_t memset2d_sync(void* p, size_t pitch, int v, size_t w, size_t h) { const cudaError_t e = cudaMemset2D(p, pitch, v, w, h); return e == cudaSuccess ? cudaStreamSynchronize(cudaStreamLegacy) : e; }

template <class T> cudaError_t dalloc(T** p, size_t count) {
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T) + 256);   // +256: aligned word reads may touch the tail
  if (e == cudaSuccess) e = memset_sync(*p, 0, count * sizeof(T) + 256);
  return e;
}

// glibc srandom_r (TYPE_3) — state after srand(seed)
void glibc_seed(unsigned seed, int* ring, int* f, int* b) {
  if (seed == 0) seed = 1;
  int st[31];
  st[0] = (int)seed;
  for (int i = 1; i < 31; i++) {
    long hi = st[i - 1] / 127773, lo = st[i - 1] % 127773;
    long w = 16807 * lo - 2836 * hi;
    if (w < 0) w += 2147483647;
    st[i] = (int)w;
  }
  int ff = 3, bb = 0;
  for (int i = 0; i < 310; i++) {
    st[ff] = (int)((unsigned)st[ff] + (unsigned)st[bb]);
    if (++ff >= 31) ff = 0;
    if (++bb >= 31) bb = 0;
  }
  memcpy(ring, st, sizeof(st)); *f = ff; *b = bb;
}

int check_stream(vslam_ctx* ctx, int s) {
  if (!ctx) return VSLAM_E_INVALID;
  if (s < 0 || s >= ctx->S) { ctx->err = "stream index out of range"; return VSLAM_E_INVALID; }
  return VSLAM_OK;
}

int check_status(vslam_ctx* ctx) {   // after a sync: report corner-capacity overflow
  int st[4];
  VS_CUDA(cudaMemcpyAsync(st, ctx->status, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (st[0]) {
    VS_CUDA(cudaMemsetAsync(ctx->status, 0, sizeof(st), ctx->stream));
    ctx->err = "corner list capacity exceeded (raise vslam_config.max_corner_frac)";
    return VSLAM_E_CAPACITY;
  }
  return VSLAM_OK;
}

}  // namespace

extern "C" {

void vslam_default_config(vslam_config* c) {
  memset(c, 0, sizeof(*c));
  c->device = 0; c->width = 640; c->height = 480; c->n_streams = 1; c->max_points = 1000; c->patch_size = 11;
  c->max_source_keyframes = 1; c->max_corner_frac = 0.5f; c->cuda_stream = nullptr; c->truncate_error = 1; c->rand_seed = 1;
}

void vslam_default_params(vslam_params* p) {   // jni/Tracker.cc:405-410,495-497,518
  p->coarse_min = 20; p->coarse_max = 60; p->coarse_range = 30; p->coarse_subpix_its = 8; p->coarse_min_vel = 0.006;
  p->fine_range = 10; p->fine_range_after_coarse = 5; p->fine_subpix_its_top_level = 8; p->max_patches_per_frame = 1000; p->use_sbi = 1; p->stream_groups = 1; p->serial_normal_equations = 0; p->pose_kernel = 0; p->search_kernel = 0; p->frame_lookahead = -1; p->coarse_chain = -1;
}

const char* vslam_last_error(const vslam_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

unsigned long long vslam_kernel_launches(const vslam_ctx* ctx) { return ctx ? ctx->launches : 0; }

void vslam_camera_from_params(const double* p, int width, int height, int as_shipped_radius, double* c) {   // jni/ATANCamera.cc:37-86
  c[0] = width * p[0]; c[1] = height * p[1]; c[2] = width * p[2] - 0.5; c[3] = height * p[3] - 0.5;
  c[4] = p[4];
  if (p[4] != 0.0) { c[6] = 2.0 * tan(p[4] / 2.0); c[7] = 1.0 / c[6]; c[5] = 1.0 / p[4]; c[8] = 1.0; }
  else { c[5] = 0.0; c[6] = 0.0; c[7] = 0.0; c[8] = 0.0; }
  double v0, v1;
  if (as_shipped_radius) { const int a = (int)p[2], b = (int)(1.0 - p[2]); v0 = (a > b ? a : b) / p[0]; const int c2 = (int)p[3], d = (int)(1.0 - p[3]); v1 = (c2 > d ? c2 : d) / p[1]; }
  else { v0 = (p[2] > 1.0 - p[2] ? p[2] : 1.0 - p[2]) / p[0]; v1 = (p[3] > 1.0 - p[3] ? p[3] : 1.0 - p[3]) / p[1]; }
  double d2 = 0; d2 += v0 * v0; d2 += v1 * v1;
  const double r = sqrt(d2);
  c[9] = (p[4] == 0.0) ? r : tan(r * p[4]) * c[7];
  c[10] = 1.5 * c[9];
  c[11] = width; c[12] = height;
}

// ---------------------------------------------------------------------------------------------- frame sets (frame look-ahead)
static void use_set(vslam_ctx* ctx, int p) {
  const FrameSet& F = ctx->sets[p];
  for (int l = 0; l < VS_LEVELS; l++) { ctx->lev[l].img = F.img[l]; ctx->lev[l].cbits = F.cbits[l]; }
  ctx->l0_ptr = F.l0_ptr; ctx->l0_stride = F.l0_stride; ctx->l0_ptr_host = F.l0_ptr_host; ctx->l0_stride_host = F.l0_stride_host;
  ctx->cur_set = p;
}
// Second frame set + the front-end streams and events, on the first look-ahead frame.
static int alloc_set1_impl(vslam_ctx* ctx);
static int alloc_set1(vslam_ctx* ctx) {
  if (ctx->have_set1) return VSLAM_OK;
  const int rc = alloc_set1_impl(ctx);
  if (rc) {   // out of memory half-way: give back what was taken, the context carries on with one frame set
    FrameSet& F = ctx->sets[1];
    for (int l = 1; l < VS_LEVELS; l++) { cudaFree(F.img[l]); F.img[l] = nullptr; }
    cudaFree(F.cbits_block); F.cbits_block = nullptr; cudaFree(F.l0_ptr); F.l0_ptr = nullptr; cudaFree(F.l0_stride); F.l0_stride = nullptr;
    delete[] F.l0_ptr_host; F.l0_ptr_host = nullptr; delete[] F.l0_stride_host; F.l0_stride_host = nullptr;
    for (cudaStream_t* st : {&ctx->front_stream, &ctx->front_side, &ctx->back_stream}) if (*st) { cudaStreamDestroy(*st); *st = nullptr; }
    for (cudaEvent_t* e : {&ctx->ev_user, &ctx->ev_front_done, &ctx->ev_barrier, &ctx->ev_back_done[0], &ctx->ev_back_done[1], &ctx->ev_la_fork, &ctx->ev_la_join}) if (*e) { cudaEventDestroy(*e); *e = nullptr; }
    cudaGetLastError();
  }
  return rc;
}
static int alloc_set1_impl(vslam_ctx* ctx) {
  FrameSet& F = ctx->sets[1];
  const int S = ctx->S;
  size_t words = 0;
  for (int l = 0; l < VS_LEVELS; l++) words += ((size_t)S * ctx->lev[l].h * ((ctx->lev[l].w + 31) / 32) + 1) / 2;
  VS_CUDA(dalloc(&F.cbits_block, words));
  { size_t off = 0; for (int l = 0; l < VS_LEVELS; l++) { F.cbits[l] = (uint32_t*)(F.cbits_block + off); off += ((size_t)S * ctx->lev[l].h * ((ctx->lev[l].w + 31) / 32) + 1) / 2; } }
  F.img[0] = ctx->l0_alt;               // (null until a host-input path needs it: vs_ensure_own_l0)
  for (int l = 1; l < VS_LEVELS; l++) VS_CUDA(dalloc(&F.img[l], (size_t)S * ctx->lev[l].h * ctx->lev[l].pitch));
  VS_CUDA(dalloc(&F.l0_ptr, (size_t)S)); VS_CUDA(dalloc(&F.l0_stride, (size_t)S));
  F.l0_ptr_host = new const uint8_t*[S]; F.l0_stride_host = new int[S];
  for (int s = 0; s < S; s++) { F.l0_ptr_host[s] = nullptr; F.l0_stride_host[s] = 0; }   // no level 0 yet: the first frame of this set uploads its table
  // the back end is the latency chain a frame's result waits for: its kernels get the CTA slots first, the front end of the next frame fills in
  int prio_least = 0, prio_greatest = 0; VS_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  VS_CUDA(cudaStreamCreateWithPriority(&ctx->front_stream, cudaStreamNonBlocking, prio_least)); VS_CUDA(cudaStreamCreateWithPriority(&ctx->front_side, cudaStreamNonBlocking, prio_least));
  VS_CUDA(cudaStreamCreateWithPriority(&ctx->back_stream, cudaStreamNonBlocking, prio_greatest));
  for (cudaEvent_t* e : {&ctx->ev_user, &ctx->ev_front_done, &ctx->ev_barrier, &ctx->ev_back_done[0], &ctx->ev_back_done[1], &ctx->ev_la_fork, &ctx->ev_la_join}) VS_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  ctx->have_set1 = true;
  return VSLAM_OK;
}
static int vs_ensure_own_l0(vslam_ctx* ctx) {
  if (ctx->lev[0].img) return VSLAM_OK;     // (only set 1 starts without one)
  if (!ctx->l0_alt) { VS_CUDA(dalloc(&ctx->l0_alt, (size_t)ctx->S * ctx->lev[0].h * ctx->lev[0].pitch)); }
  ctx->sets[1].img[0] = ctx->l0_alt;
  if (ctx->cur_set == 1) ctx->lev[0].img = ctx->l0_alt;
  return VSLAM_OK;
}
// Start of every vslam_track_frame*.  With look-ahead the frame gets the OTHER frame set, and its input and front end go to front_stream, which
// waits only for the back end that last read that set (two frames ago) -- so they run beside the previous frame's back end, still in flight on
// ctx->stream.  That shortcut is taken only when nothing but frames was launched since the last look-ahead frame (ctx->launches unchanged);
// otherwise front_stream waits for everything enqueued on ctx->stream so far.
static bool lookahead_default(const vslam_ctx* ctx) {
  const char* e = getenv("VSLAM_LOOKAHEAD");
  if (e) return atoi(e) != 0;
  // Measured (B200, ms per step without / with): VGA 32 streams 0.315 / 0.253, 64: 0.390 / 0.319, 128: 0.495 / 0.412, 256: 0.726 / 0.745; 1080p x 148: 1.80 / 1.76;
  // 4K x 148: 7.15 / 6.89.  What decides is whether the one-CTA-per-stream kernels of the back end leave room on the SMs (up to about one stream per SM)
  return ctx->S <= 160;
}
static int lookahead_resolved(const vslam_ctx* ctx) {
  int la = ctx->params.frame_lookahead < 0 ? (lookahead_default(ctx) ? 1 : 0) : ctx->params.frame_lookahead;
  if (ctx->timing || ctx->params.stream_groups > 1 || ctx->la_unavailable) la = 0;      // per-stage timing serialises everything on ctx->stream; stream groups have their own streams
  return la;
}
int vslam_frame_lookahead_active(const vslam_ctx* ctx) { return ctx ? lookahead_resolved(ctx) : 0; }
static int vs_begin_frame(vslam_ctx* ctx) {
  const int la = lookahead_resolved(ctx);
  ctx->la_frame = la != 0; ctx->front = ctx->stream;
  if (!la) return VSLAM_OK;
  if (alloc_set1(ctx)) {   // no memory for a second frame set: this context runs without look-ahead from here on
    ctx->la_unavailable = true; ctx->la_frame = false; ctx->err.clear();
    return VSLAM_OK;
  }
  const bool chained = ctx->launches == ctx->launches_after_frame;
  use_set(ctx, ctx->cur_set ^ 1);
  if (chained) VS_CUDA(cudaStreamWaitEvent(ctx->front_stream, ctx->ev_back_done[ctx->cur_set], 0));
  else { VS_CUDA(cudaEventRecord(ctx->ev_barrier, ctx->stream)); VS_CUDA(cudaStreamWaitEvent(ctx->front_stream, ctx->ev_barrier, 0)); }
  ctx->front = ctx->front_stream;
  return VSLAM_OK;
}
static int end_frame(vslam_ctx* ctx, int rc) { ctx->front = nullptr; ctx->la_frame = false; return rc; }

int vslam_create(const vslam_config* cfg, vslam_ctx** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return VSLAM_E_INVALID; }
  *out = nullptr;
  if (cfg->width <= 0 || cfg->height <= 0 || cfg->width % 32 || cfg->height % 8 || cfg->width > 65535 || cfg->height > 65535) {
    g_create_error = "width must be a positive multiple of 32 and height of 8 (every pyramid level keeps even dimensions)"; return VSLAM_E_INVALID; }
  if (cfg->max_points > 65536) { g_create_error = "max_points must be <= 65536"; return VSLAM_E_INVALID; }
  if (cfg->n_streams < 1 || cfg->max_points < 1 || cfg->max_source_keyframes < 1) { g_create_error = "n_streams, max_points, max_source_keyframes must be >= 1"; return VSLAM_E_INVALID; }
  if (cfg->patch_size < 4 || cfg->patch_size > VSLAM_MAX_PATCH) { g_create_error = "patch_size must be in [4, 11]"; return VSLAM_E_INVALID; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device) { g_create_error = "no CUDA device (this library has no CPU fallback)"; cudaGetLastError(); return VSLAM_E_NO_DEVICE; }
  vslam_ctx* ctx = new (std::nothrow) vslam_ctx();
  if (!ctx) { g_create_error = "out of host memory"; return VSLAM_E_INVALID; }
  ctx->cfg = *cfg; vslam_default_params(&ctx->params);
  ctx->S = cfg->n_streams; ctx->N = cfg->max_points; ctx->P = cfg->patch_size; ctx->launches = 0; ctx->n_src = cfg->max_source_keyframes; ctx->src_have.assign(ctx->n_src, 0);
  ctx->reloc_n = 0; ctx->reloc_tmpl = nullptr; ctx->reloc_jac = nullptr; ctx->reloc_tmp = nullptr; ctx->reloc_small = nullptr; ctx->reloc_pose = nullptr; ctx->reloc_scores = nullptr;
  ctx->kf_req = nullptr; ctx->kf_policy = false; ctx->kf_wiggle = 0.1; ctx->kf_wiggle_dn = 0.1; ctx->kf_mult = 0.2; ctx->kf_min_frames = 20;
  ctx->unproj_lut = nullptr; ctx->unproj_ok = false; ctx->epi_buf = nullptr; ctx->epi_cap = 0; for (int g = 0; g < VS_MAX_GROUPS; g++) { ctx->side_stream[g] = nullptr; ctx->group_stream[g] = nullptr; ctx->ev_fork[g] = nullptr; ctx->ev_join[g] = nullptr; ctx->ev_end[g] = nullptr; }
  ctx->ev_begin = nullptr; ctx->cur_s0 = 0; ctx->cur_cnt = cfg->n_streams; ctx->cur_group = 0; ctx->scratch_host = nullptr; ctx->scratch_host_bytes = 0; ctx->timing = false; ctx->ev_used = 0; ctx->l0_alt = nullptr; ctx->copy_stream = nullptr; ctx->step = 0; ctx->pipe_ready = false; ctx->status_pin = nullptr;
  ctx->rest_scores = nullptr; ctx->rest_max = nullptr; ctx->rest_cand = nullptr; ctx->rest_cand_score = nullptr; ctx->rest_counts = nullptr; ctx->rest_stream = -1;
  ctx->snap_img = nullptr; ctx->snap_corners = nullptr; ctx->snap_lut = nullptr;
  ctx->sbi_on = false; ctx->sbi_tmpl = nullptr; ctx->sbi_scratch = nullptr; ctx->sbi_jac = nullptr; ctx->sbi_small = nullptr; ctx->sbi_have = nullptr;
  const float frac = cfg->max_corner_frac > 0 ? cfg->max_corner_frac : 0.5f;
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_create_error = std::string(#call) + ": " + cudaGetErrorString(e_); vslam_destroy(ctx); return VSLAM_E_CUDA; } } while (0)
  CK(cudaSetDevice(cfg->device));
  if (cfg->cuda_stream) { ctx->stream = (cudaStream_t)cfg->cuda_stream; ctx->own_stream = false; }
  else { CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
  for (int g = 0; g < VS_MAX_GROUPS; g++) {
    CK(cudaStreamCreateWithFlags(&ctx->side_stream[g], cudaStreamNonBlocking));
    if (g > 0) CK(cudaStreamCreateWithFlags(&ctx->group_stream[g], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_fork[g], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->ev_join[g], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_end[g], cudaEventDisableTiming));
  }
  CK(cudaEventCreateWithFlags(&ctx->ev_begin, cudaEventDisableTiming));
  const int S = ctx->S, N = ctx->N;
  ctx->user_events.assign(S, 0);
  // programmatic dependent launch of a frame's kernels: on by default up to 1080p frames (measured: VGA 0.7509 -> 0.7318 ms per step, 1080p neutral,
  // 4K 7.15 -> 7.51 ms: there every kernel runs for many waves and the parked CTAs of its successor only take registers and shared memory away);
  // VSLAM_PDL=0 / 1 overrides (A/B runs)
  { const char* e = getenv("VSLAM_PDL"); ctx->pdl = e ? atoi(e) != 0 : ((long long)cfg->width * cfg->height <= 1920ll * 1080ll); }
  int w = cfg->width, h = cfg->height;
  size_t strip_words = 0;
  for (int l = 0; l < VS_LEVELS; l++) {
    LevelDesc& L = ctx->lev[l];
    L.w = w; L.h = h; L.pitch = round_up(w, 128);
    L.cap = (int)((double)w * h * frac); if (L.cap < 64) L.cap = 64;
    L.strip_rows = vs_strip_rows(l, w, L.pitch);
    L.n_strips = (h + L.strip_rows - 1) / L.strip_rows;
    { auto mg = [](int d) { return d <= 1 ? 0u : 0xffffffffu / (uint32_t)d + 1u; };
      L.mg_strips = mg(L.n_strips); L.mg_cpr = mg((w + 15) / 16); L.mg_wpr = mg((w + 31) / 32); L.mg_hw = mg(w / 2); L.mg_w = mg(w); }
    CK(dalloc(&L.img, (size_t)S * h * L.pitch));
    CK(dalloc(&L.corners, (size_t)S * L.cap));
    CK(dalloc(&L.lut, (size_t)S * (h + 1)));
    strip_words += ((size_t)S * h * ((w + 31) / 32) + 1) / 2;      // corner bitmask words (32 bit) of this level, in 64-bit units
    ctx->src.w[l] = w; ctx->src.h[l] = h; ctx->src.pitch[l] = L.pitch;
    CK(dalloc(&ctx->src.img[l], (size_t)ctx->n_src * h * L.pitch));
    w /= 2; h /= 2;
  }
  // one allocation for the corner bitmasks of all levels and the tickets, so that a whole-context frame clears them with one memset
  CK(dalloc(&ctx->sync_words, strip_words + VS_MAX_GROUPS)); ctx->sync_words_n = strip_words + VS_MAX_GROUPS;
  { size_t off = 0;
    for (int l = 0; l < VS_LEVELS; l++) { LevelDesc& L = ctx->lev[l]; L.cbits = (uint32_t*)(ctx->sync_words + off); off += ((size_t)S * L.h * ((L.w + 31) / 32) + 1) / 2; }
    ctx->tickets = (unsigned*)(ctx->sync_words + off); }
  CK(dalloc(&ctx->l0_ptr, (size_t)S)); CK(dalloc(&ctx->l0_stride, (size_t)S));
  ctx->l0_ptr_host = new const uint8_t*[S]; ctx->l0_stride_host = new int[S];
  for (int s = 0; s < S; s++) { ctx->l0_ptr_host[s] = ctx->lev[0].img + (size_t)s * ctx->lev[0].h * ctx->lev[0].pitch; ctx->l0_stride_host[s] = ctx->lev[0].pitch; }
  CK(cudaMemcpy(ctx->l0_ptr, ctx->l0_ptr_host, sizeof(uint8_t*) * S, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->l0_stride, ctx->l0_stride_host, sizeof(int) * S, cudaMemcpyHostToDevice));
  { FrameSet& F = ctx->sets[0];   // the one frame set every context has (a second one comes with frame look-ahead: alloc_set1)
    for (int l = 0; l < VS_LEVELS; l++) { F.img[l] = ctx->lev[l].img; F.cbits[l] = ctx->lev[l].cbits; }
    F.cbits_block = nullptr; F.l0_ptr = ctx->l0_ptr; F.l0_stride = ctx->l0_stride; F.l0_ptr_host = ctx->l0_ptr_host; F.l0_stride_host = ctx->l0_stride_host; }
  CK(dalloc(&ctx->status, (size_t)4)); CK(dalloc(&ctx->evals, (size_t)4));   // [0] ZMSSD candidates scored, [1] unused, [2] templates generated, [3] sub-pixel refinements
  // map
  ctx->map.n = 0;
  CK(dalloc(&ctx->map.world, (size_t)3 * N)); CK(dalloc(&ctx->map.right, (size_t)3 * N)); CK(dalloc(&ctx->map.down, (size_t)3 * N));
  CK(dalloc(&ctx->map.ircenter, (size_t)2 * N)); CK(dalloc(&ctx->map.srclevel, (size_t)N)); CK(dalloc(&ctx->map.srckf, (size_t)N));
  // per (stream, point) state
  const size_t SN = (size_t)S * N;
  PointState& ps = ctx->ps;
  CK(dalloc(&ps.v3cam, 3 * SN)); CK(dalloc(&ps.v2image, 2 * SN)); CK(dalloc(&ps.derivs, 4 * SN)); CK(dalloc(&ps.warpinv, 4 * SN)); CK(dalloc(&ps.m2, 4 * SN));
  CK(dalloc(&ps.lastwarp, 4 * SN)); CK(dalloc(&ps.v2found, 2 * SN)); CK(dalloc(&ps.coarse, 2 * SN)); CK(dalloc(&ps.jac, 12 * SN));
  CK(dalloc(&ps.err, 2 * SN)); CK(dalloc(&ps.sqrtinv, SN)); CK(dalloc(&ps.flags, SN)); CK(dalloc(&ps.level, SN)); CK(dalloc(&ps.rlevel, SN));
  CK(dalloc(&ps.tmpl, SN * VS_TMPL_BYTES)); CK(dalloc(&ps.tsum, 2 * SN)); CK(dalloc(&ps.counts, 2 * SN));
  { std::vector<int> lv(SN, -1); CK(cudaMemcpy(ps.level, lv.data(), SN * sizeof(int), cudaMemcpyHostToDevice)); }
  // per-stream state
  CK(dalloc(&ctx->ss, (size_t)S)); CK(dalloc(&ctx->kf_req, (size_t)S));
  CK(cudaHostAlloc((void**)&ctx->coarse_hint_host, sizeof(int) * (size_t)S, cudaHostAllocMapped)); memset(ctx->coarse_hint_host, 0, sizeof(int) * (size_t)S);
  CK(cudaHostGetDevicePointer((void**)&ctx->coarse_hint_dev, ctx->coarse_hint_host, 0));
  {
    std::vector<StreamState> h(S);
    memset(h.data(), 0, sizeof(StreamState) * S);
    for (int s = 0; s < S; s++) {
      StreamState& st = h[s];
      st.pose[0] = st.pose[5] = st.pose[10] = 1.0; memcpy(st.start_pose, st.pose, sizeof(st.pose));
      st.depth_mean = 1.0; st.depth_sigma = 1.0; st.quality = 2;   // Tracker::Reset (jni/Tracker.cc:45-60)
      st.frame_no = 0; st.last_kf_dropped = -20; st.kf_closest = -1;
      glibc_seed(cfg->rand_seed, st.rng_ring, &st.rng_f, &st.rng_b);
    }
    CK(cudaMemcpy(ctx->ss, h.data(), sizeof(StreamState) * S, cudaMemcpyHostToDevice));
  }
  ctx->list_cap = N + 8;
  CK(dalloc(&ctx->lists, (size_t)S * ctx->list_cap));
  CK(dalloc(&ctx->pvs, (size_t)S * VS_LEVELS * N));
  { int c = 1; while (c < ctx->list_cap) c <<= 1; ctx->sort_cap = c; }
  CK(dalloc(&ctx->sort_scratch, (size_t)S * ctx->sort_cap));
  memset(&ctx->cam, 0, sizeof(ctx->cam));
#undef CK
  *out = ctx;
  return VSLAM_OK;
}

void vslam_destroy(vslam_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->front_stream) cudaStreamSynchronize(ctx->front_stream);
  use_set(ctx, 0);                      // the aliases freed below are set 0's; set 1 owns the rest (its level 0 is l0_alt)
  if (ctx->have_set1) {
    FrameSet& F = ctx->sets[1];
    for (int l = 1; l < VS_LEVELS; l++) cudaFree(F.img[l]);
    cudaFree(F.cbits_block); cudaFree(F.l0_ptr); cudaFree(F.l0_stride); delete[] F.l0_ptr_host; delete[] F.l0_stride_host;
  }
  if (ctx->front_stream) cudaStreamDestroy(ctx->front_stream);
  if (ctx->front_side) cudaStreamDestroy(ctx->front_side);
  if (ctx->back_stream) { cudaStreamSynchronize(ctx->back_stream); cudaStreamDestroy(ctx->back_stream); }
  for (cudaEvent_t e : {ctx->ev_user, ctx->ev_front_done, ctx->ev_barrier, ctx->ev_back_done[0], ctx->ev_back_done[1], ctx->ev_la_fork, ctx->ev_la_join}) if (e) cudaEventDestroy(e);
  cudaFree(ctx->sbi_rot_buf); cudaFree(ctx->reloc_frame_scratch); cudaFree(ctx->reloc_frame_small);
  for (int l = 0; l < VS_LEVELS; l++) { cudaFree(ctx->lev[l].img); cudaFree(ctx->lev[l].corners); cudaFree(ctx->lev[l].lut); cudaFree(ctx->src.img[l]); }
  cudaFree(ctx->epi_buf); cudaFree(ctx->pf_buf); cudaFree(ctx->list_counts); cudaFree(ctx->l0_ptr); cudaFree(ctx->l0_stride); cudaFree(ctx->sync_words); cudaFree(ctx->status); cudaFree(ctx->evals);
  cudaFree(ctx->map.world); cudaFree(ctx->map.right); cudaFree(ctx->map.down); cudaFree(ctx->map.ircenter); cudaFree(ctx->map.srclevel); cudaFree(ctx->map.srckf);
  PointState& ps = ctx->ps;
  cudaFree(ps.v3cam); cudaFree(ps.v2image); cudaFree(ps.derivs); cudaFree(ps.warpinv); cudaFree(ps.m2); cudaFree(ps.lastwarp); cudaFree(ps.v2found); cudaFree(ps.coarse);
  cudaFree(ps.jac); cudaFree(ps.err); cudaFree(ps.sqrtinv); cudaFree(ps.flags); cudaFree(ps.level); cudaFree(ps.rlevel); cudaFree(ps.tmpl); cudaFree(ps.tsum); cudaFree(ps.counts);
  cudaFree(ctx->ss); cudaFree(ctx->kf_req); cudaFree(ctx->lists); cudaFree(ctx->pvs); cudaFree(ctx->sort_scratch);
  if (ctx->scratch_host) cudaFreeHost(ctx->scratch_host);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->pipe_ready) { for (int k = 0; k < 2; k++) { cudaEventDestroy(ctx->ev_copied[k]); cudaEventDestroy(ctx->ev_computed[k]); cudaEventDestroy(ctx->ev_done[k]); } cudaStreamDestroy(ctx->copy_stream); }
  cudaFree(ctx->l0_alt);
  cudaFree(ctx->rest_scores); cudaFree(ctx->rest_max); cudaFree(ctx->rest_cand); cudaFree(ctx->rest_cand_score); cudaFree(ctx->rest_counts);
  cudaFree(ctx->snap_img); cudaFree(ctx->snap_corners); cudaFree(ctx->snap_lut);
  cudaFree(ctx->sbi_resize_tab); cudaFree(ctx->sbi_tmpl); cudaFree(ctx->sbi_scratch); cudaFree(ctx->sbi_jac); cudaFree(ctx->sbi_small); cudaFree(ctx->sbi_have);
  if (ctx->status_pin) cudaFreeHost(ctx->status_pin);
  if (ctx->coarse_hint_host) cudaFreeHost(ctx->coarse_hint_host);
  delete[] ctx->l0_ptr_host; delete[] ctx->l0_stride_host;
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaFree(ctx->unproj_lut); cudaFree(ctx->reloc_tmpl); cudaFree(ctx->reloc_jac); cudaFree(ctx->reloc_tmp); cudaFree(ctx->reloc_small); cudaFree(ctx->reloc_pose); cudaFree(ctx->reloc_scores);
  for (int g = 0; g < VS_MAX_GROUPS; g++) {
    if (ctx->side_stream[g]) cudaStreamDestroy(ctx->side_stream[g]);
    if (ctx->group_stream[g]) cudaStreamDestroy(ctx->group_stream[g]);
    if (ctx->ev_fork[g]) cudaEventDestroy(ctx->ev_fork[g]);
    if (ctx->ev_join[g]) cudaEventDestroy(ctx->ev_join[g]);
    if (ctx->ev_end[g]) cudaEventDestroy(ctx->ev_end[g]);
    if (ctx->chain_stream[g]) cudaStreamDestroy(ctx->chain_stream[g]);
    if (ctx->ev_chain_fork[g]) cudaEventDestroy(ctx->ev_chain_fork[g]);
    if (ctx->ev_chain_join[g]) cudaEventDestroy(ctx->ev_chain_join[g]);
  }
  if (ctx->chain_stream_hi) cudaStreamDestroy(ctx->chain_stream_hi);
  if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
  delete ctx;
}

int vslam_sync(vslam_ctx* ctx) {
  if (!ctx) return VSLAM_E_INVALID;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  return check_status(ctx);
}

int vslam_set_timing(vslam_ctx* ctx, int on) {
  if (!ctx) return VSLAM_E_INVALID;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->timing = on != 0; ctx->ev_used = 0; ctx->ev_stage.clear();
  return VSLAM_OK;
}

int vslam_get_stage_times(vslam_ctx* ctx, double* ms, int* launches) {
  if (!ctx || !ms || !launches) return VSLAM_E_INVALID;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < VSLAM_N_STAGES; k++) { ms[k] = 0.0; launches[k] = 0; }
  for (size_t k = 0; k < ctx->ev_stage.size(); k++) {
    float t = 0.f;
    VS_CUDA(cudaEventElapsedTime(&t, ctx->ev_pool[2 * k], ctx->ev_pool[2 * k + 1]));
    ms[ctx->ev_stage[k]] += t; launches[ctx->ev_stage[k]]++;
  }
  ctx->ev_used = 0; ctx->ev_stage.clear();
  return VSLAM_OK;
}

int vslam_set_params(vslam_ctx* ctx, const vslam_params* p) {
  if (!ctx || !p) return VSLAM_E_INVALID;
  if (2 * p->coarse_max > (unsigned)ctx->list_cap) { ctx->err = "coarse_max too large"; return VSLAM_E_INVALID; }
  if (p->stream_groups < 0 || p->stream_groups > VS_MAX_GROUPS) { ctx->err = "stream_groups must be 0 (default) .. 4"; return VSLAM_E_INVALID; }
  if (p->max_patches_per_frame < 0) { ctx->err = "max_patches_per_frame must be >= 0"; return VSLAM_E_INVALID; }
  if (p->coarse_chain < -1 || p->coarse_chain > 1) { ctx->err = "coarse_chain must be -1 (library default), 0 or 1"; return VSLAM_E_INVALID; }
  if (p->frame_lookahead < -1 || p->frame_lookahead > 1) { ctx->err = "frame_lookahead must be -1 (library default), 0 or 1"; return VSLAM_E_INVALID; }
  ctx->params = *p;
  return VSLAM_OK;
}

// SmallBlurryImage on the device: needs the camera at the SBI image size (level 3 halved = width/16 x height/16).
int vslam_enable_sbi(vslam_ctx* ctx, const double* c) {
  if (!ctx || !c) return VSLAM_E_INVALID;
  const LevelDesc& L3 = ctx->lev[3];
  {   // cv::resize(level 3, (cols / 2, rows / 2)) (jni/SmallBlurryImage.cc:22-30): exactly half unless a level-3 dimension is odd (1080p: 240 x 135 ->
      // 120 x 67), then OpenCV's fixed-point INTER_LINEAR.  Coefficient tables as resizeGeneric_ builds them (float sample positions, cvRound to 11 bits).
    const int sw = L3.w, sh = L3.h, dw = sw / 2, dh = sh / 2;
    if (dw < 1 || dh < 1) { ctx->err = "image too small for a SmallBlurryImage"; return VSLAM_E_INVALID; }
    const double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
    ctx->sbi_exact_half = (sw == 2 * dw && sh == 2 * dh) ? 1 : 0;
    std::vector<int> tab((size_t)3 * dw + 3 * dh);
    int* xofs = tab.data(); int* alpha = xofs + dw; int* yofs = alpha + 2 * dw; int* beta = yofs + dh;
    auto q11 = [](float v) { long r = lrintf(v); return (int)(r < -32768 ? -32768 : (r > 32767 ? 32767 : r)); };
    for (int dx = 0; dx < dw; dx++) {
      float fx = (float)((dx + 0.5) * scale_x - 0.5); int sx = (int)floorf(fx); fx -= sx;
      if (sx < 0) { fx = 0; sx = 0; }
      if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
      xofs[dx] = sx; alpha[2 * dx] = q11((1.f - fx) * 2048); alpha[2 * dx + 1] = q11(fx * 2048);
    }
    for (int dy = 0; dy < dh; dy++) {
      float fy = (float)((dy + 0.5) * scale_y - 0.5); const int sy = (int)floorf(fy); fy -= sy;
      yofs[dy] = sy; beta[2 * dy] = q11((1.f - fy) * 2048); beta[2 * dy + 1] = q11(fy * 2048);
    }
    if (!ctx->sbi_resize_tab) VS_CUDA(dalloc(&ctx->sbi_resize_tab, tab.size()));
    VS_CUDA(cudaMemcpy(ctx->sbi_resize_tab, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice));
  }
  CamDev& d = ctx->sbi_cam;
  d.fx = c[0]; d.fy = c[1]; d.cx = c[2]; d.cy = c[3]; d.W = c[4]; d.Winv = c[5]; d.twoTan = c[6]; d.oneOver2Tan = c[7]; d.distEnabled = c[8];
  d.largestRadius = c[9]; d.maxR = c[10]; d.width = c[11]; d.height = c[12];
  const int w = L3.w / 2, h = L3.h / 2; const size_t n = (size_t)w * h;
  {   // cv::getGaussianKernel(9, 0.75, CV_32F) as the stand-in computes it (host exp, float taps)
    const double sigma = 0.75, scale2x = -0.5 / (sigma * sigma); double sum = 0;
    for (int i = 0; i < 9; i++) { const double x = i - 4.0; ctx->sbi_taps[i] = (float)exp(scale2x * x * x); sum += ctx->sbi_taps[i]; }
    sum = 1. / sum; for (int i = 0; i < 9; i++) ctx->sbi_taps[i] = (float)(ctx->sbi_taps[i] * sum);
  }
  for (int k = 0; k < 2; k++) {   // ATANCamera::UnProject of (w/2 +- 5, h/2) on the host (jni/ATANCamera.cc:149-164)
    const double im[2] = {w / 2.0 + (k ? -5.0 : 5.0), h / 2.0};
    const double dx = (im[0] - d.cx) * (1.0 / d.fx), dy = (im[1] - d.cy) * (1.0 / d.fy);
    const double distR = sqrt(dx * dx + dy * dy);
    const double r = (d.W == 0.0) ? distR : tan(distR * d.W) * d.oneOver2Tan;
    const double factor = (distR > 0.01) ? r / distR : 1.0;
    ctx->sbi_orig[k][0] = dx * factor; ctx->sbi_orig[k][1] = dy * factor; ctx->sbi_orig[k][2] = 1.0;
  }
  if (!ctx->sbi_tmpl) {
    VS_CUDA(dalloc(&ctx->sbi_tmpl, (size_t)ctx->S * 2 * n)); VS_CUDA(dalloc(&ctx->sbi_scratch, (size_t)ctx->S * 3 * n));
    VS_CUDA(dalloc(&ctx->sbi_jac, (size_t)ctx->S * 2 * n)); VS_CUDA(dalloc(&ctx->sbi_small, (size_t)ctx->S * n)); VS_CUDA(dalloc(&ctx->sbi_have, (size_t)2 * ctx->S));
    VS_CUDA(dalloc(&ctx->sbi_rot_buf, (size_t)2 * ctx->S * 6)); VS_CUDA(dalloc(&ctx->reloc_frame_scratch, (size_t)ctx->S * 3 * n)); VS_CUDA(dalloc(&ctx->reloc_frame_small, (size_t)ctx->S * n));
  }
  ctx->sbi_on = true;
  return VSLAM_OK;
}

// Relocaliser keyframes (jni/Relocaliser.cc): n map keyframes given as source-keyframe ids (their pyramids are on the device) and poses.
int vslam_set_reloc_keyframes(vslam_ctx* ctx, int n, const int32_t* src_kf_ids, const double* poses12) {
  if (!ctx || n < 0 || (n > 0 && (!src_kf_ids || !poses12))) return VSLAM_E_INVALID;
  if (!ctx->sbi_on) { ctx->err = "vslam_set_reloc_keyframes needs vslam_enable_sbi first (SmallBlurryImage size and camera)"; return VSLAM_E_INVALID; }
  for (int k = 0; k < n; k++) if (src_kf_ids[k] < 0 || src_kf_ids[k] >= ctx->n_src) { ctx->err = "relocaliser keyframe is not a source keyframe id"; return VSLAM_E_INVALID; }
  for (int k = 0; k < n; k++) if (!ctx->src_have[src_kf_ids[k]]) { ctx->err = "relocaliser keyframe slot holds no image (vslam_upload_source_keyframe / vslam_add_keyframe_from_stream first)"; return VSLAM_E_INVALID; }
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(ctx->reloc_tmpl); cudaFree(ctx->reloc_jac); cudaFree(ctx->reloc_tmp); cudaFree(ctx->reloc_small); cudaFree(ctx->reloc_pose); cudaFree(ctx->reloc_scores);
  ctx->reloc_tmpl = nullptr; ctx->reloc_jac = nullptr; ctx->reloc_tmp = nullptr; ctx->reloc_small = nullptr; ctx->reloc_pose = nullptr; ctx->reloc_scores = nullptr;
  ctx->reloc_n = 0; ctx->reloc_ids.clear(); ctx->reloc_poses_host.clear();
  if (n == 0) {   // no relocaliser, no k_relocalise to reset the per-frame `recovered` mark
    VS_CUDA(memset2d_sync((char*)ctx->ss + offsetof(StreamState, recovered), sizeof(StreamState), 0, sizeof(int), ctx->S));
    return VSLAM_OK;
  }
  const size_t px = (size_t)(ctx->lev[3].w / 2) * (ctx->lev[3].h / 2);
  {   // cv::getGaussianKernel(17, 2.5, CV_32F) as the stand-in computes it (host exp, float taps)
    const double sigma = 2.5, scale2x = -0.5 / (sigma * sigma); double sum = 0;
    for (int i = 0; i < 17; i++) { const double x = i - 8.0; ctx->reloc_taps[i] = (float)exp(scale2x * x * x); sum += ctx->reloc_taps[i]; }
    sum = 1. / sum; for (int i = 0; i < 17; i++) ctx->reloc_taps[i] = (float)(ctx->reloc_taps[i] * sum);
  }
  VS_CUDA(dalloc(&ctx->reloc_tmpl, (size_t)n * px)); VS_CUDA(dalloc(&ctx->reloc_jac, (size_t)n * 2 * px)); VS_CUDA(dalloc(&ctx->reloc_tmp, (size_t)n * px));
  VS_CUDA(dalloc(&ctx->reloc_small, (size_t)n * px)); VS_CUDA(dalloc(&ctx->reloc_pose, (size_t)n * 12)); VS_CUDA(dalloc(&ctx->reloc_scores, (size_t)ctx->S * n));
  VS_CUDA(cudaMemcpy(ctx->reloc_pose, poses12, sizeof(double) * 12 * n, cudaMemcpyHostToDevice));
  int* ids = nullptr; VS_CUDA(cudaMalloc(&ids, sizeof(int) * n));
  cudaError_t e = cudaMemcpy(ids, src_kf_ids, sizeof(int) * n, cudaMemcpyHostToDevice);
  ctx->reloc_n = n;
  int rc = e == cudaSuccess ? vs_launch_reloc_make(ctx, ids) : VSLAM_E_CUDA;
  if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = VSLAM_E_CUDA;
  cudaFree(ids);
  if (rc) { ctx->reloc_n = 0; ctx->err = "vslam_set_reloc_keyframes: CUDA error"; }
  else { ctx->reloc_ids.assign(src_kf_ids, src_kf_ids + n); ctx->reloc_poses_host.assign(poses12, poses12 + 12 * (size_t)n); }
  return rc;
}
int vslam_set_camera(vslam_ctx* ctx, const double* c) {
  if (!ctx || !c) return VSLAM_E_INVALID;
  CamDev& d = ctx->cam;
  d.fx = c[0]; d.fy = c[1]; d.cx = c[2]; d.cy = c[3]; d.W = c[4]; d.Winv = c[5]; d.twoTan = c[6]; d.oneOver2Tan = c[7]; d.distEnabled = c[8];
  d.largestRadius = c[9]; d.maxR = c[10]; d.width = c[11]; d.height = c[12];
  ctx->unproj_ok = false;
  return VSLAM_OK;
}

int vslam_upload_source_keyframe(vslam_ctx* ctx, int kf, const uint8_t* gray, int stride) {
  if (!ctx || !gray) return VSLAM_E_INVALID;
  if (kf < 0 || kf >= ctx->n_src || stride < ctx->src.w[0]) { ctx->err = "bad source keyframe id or stride"; return VSLAM_E_INVALID; }
  VS_CUDA(cudaMemcpy2DAsync(ctx->src.img[0] + (size_t)kf * ctx->src.h[0] * ctx->src.pitch[0], ctx->src.pitch[0], gray, stride, ctx->src.w[0], ctx->src.h[0],
                            cudaMemcpyHostToDevice, ctx->stream));
  int rc = vs_launch_source_pyramid(ctx, kf);
  if (rc) return rc;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));   // the host buffer is only read during this call
  ctx->src_have[kf] = 1;
  return VSLAM_OK;
}

int vslam_set_map(vslam_ctx* ctx, int n, const double* world, const double* right, const double* down, const int32_t* irc, const int32_t* lvl, const int32_t* kf) {
  if (!ctx || !world || !right || !down || !irc || !lvl) return VSLAM_E_INVALID;
  if (n < 0 || n > ctx->N) { ctx->err = "map larger than max_points"; return VSLAM_E_INVALID; }
  for (int i = 0; i < n; i++) if (lvl[i] < 0 || lvl[i] >= VS_LEVELS || (kf && (kf[i] < 0 || kf[i] >= ctx->n_src))) { ctx->err = "map point with bad source level / keyframe"; return VSLAM_E_INVALID; }
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  VS_CUDA(cudaMemcpy(ctx->map.world, world, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
  VS_CUDA(cudaMemcpy(ctx->map.right, right, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
  VS_CUDA(cudaMemcpy(ctx->map.down, down, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
  VS_CUDA(cudaMemcpy(ctx->map.ircenter, irc, sizeof(int) * 2 * n, cudaMemcpyHostToDevice));
  VS_CUDA(cudaMemcpy(ctx->map.srclevel, lvl, sizeof(int) * n, cudaMemcpyHostToDevice));
  if (kf) VS_CUDA(cudaMemcpy(ctx->map.srckf, kf, sizeof(int) * n, cudaMemcpyHostToDevice));
  else VS_CUDA(memset_sync(ctx->map.srckf, 0, sizeof(int) * n));
  // a new map invalidates every per-point tracker state (TrackerData is created lazily per MapPoint, jni/Tracker.cc:372)
  const size_t SN = (size_t)ctx->S * ctx->N;
  VS_CUDA(memset_sync(ctx->ps.flags, 0, SN * sizeof(int)));
  VS_CUDA(memset_sync(ctx->ps.counts, 0, 2 * SN * sizeof(int)));
  ctx->map.n = n;
  return VSLAM_OK;
}

// Map::vpPoints.push_back while streams run (MapMaker::AddPointEpipolar from its thread, jni/MapMaker.cc:685): the new points go behind
// the existing ones; existing points keep their per-stream tracker state (TrackerData: template cache, found flags, M-estimator counters),
// the new ones start without (created on first use, jni/Tracker.cc:372).
int vslam_append_map_points(vslam_ctx* ctx, int n_new, const double* world, const double* right, const double* down, const int32_t* irc, const int32_t* lvl, const int32_t* kf) {
  if (!ctx || n_new < 0 || (n_new > 0 && (!world || !right || !down || !irc || !lvl))) return VSLAM_E_INVALID;
  const int n0 = ctx->map.n;
  if (n0 + n_new > ctx->N) { ctx->err = "map would exceed max_points"; return VSLAM_E_CAPACITY; }
  for (int i = 0; i < n_new; i++) if (lvl[i] < 0 || lvl[i] >= VS_LEVELS || (kf && (kf[i] < 0 || kf[i] >= ctx->n_src))) { ctx->err = "map point with bad source level / keyframe"; return VSLAM_E_INVALID; }
  if (n_new == 0) return VSLAM_OK;
  int rc = vslam_sync(ctx); if (rc) return rc;
  VS_CUDA(cudaMemcpy(ctx->map.world + 3 * (size_t)n0, world, sizeof(double) * 3 * n_new, cudaMemcpyHostToDevice));
  VS_CUDA(cudaMemcpy(ctx->map.right + 3 * (size_t)n0, right, sizeof(double) * 3 * n_new, cudaMemcpyHostToDevice));
  VS_CUDA(cudaMemcpy(ctx->map.down + 3 * (size_t)n0, down, sizeof(double) * 3 * n_new, cudaMemcpyHostToDevice));
  VS_CUDA(cudaMemcpy(ctx->map.ircenter + 2 * (size_t)n0, irc, sizeof(int) * 2 * n_new, cudaMemcpyHostToDevice));
  VS_CUDA(cudaMemcpy(ctx->map.srclevel + n0, lvl, sizeof(int) * n_new, cudaMemcpyHostToDevice));
  if (kf) VS_CUDA(cudaMemcpy(ctx->map.srckf + n0, kf, sizeof(int) * n_new, cudaMemcpyHostToDevice));
  else VS_CUDA(memset_sync(ctx->map.srckf + n0, 0, sizeof(int) * n_new));
  // per-stream state of the new points only: rows of [S][N] (flags) and [2][S][N] (counters)
  VS_CUDA(memset2d_sync(ctx->ps.flags + n0, sizeof(int) * ctx->N, 0, sizeof(int) * n_new, ctx->S));
  VS_CUDA(memset2d_sync(ctx->ps.counts + n0, sizeof(int) * ctx->N, 0, sizeof(int) * n_new, 2 * (size_t)ctx->S));
  ctx->map.n = n0 + n_new;
  return VSLAM_OK;
}

// ---------------------------------------------------------------------------------------------- MakeKeyFrame_Lite
static int adopt_l0(vslam_ctx* ctx, int first, int count, const uint8_t* base, int stride, size_t frame_stride, bool own) {
  if (own) { const int rc = vs_ensure_own_l0(ctx); if (rc) return rc; }
  for (int k = 0; k < count; k++) {
    const int s = first + k;
    ctx->l0_ptr_host[s] = own ? ctx->lev[0].img + (size_t)s * ctx->lev[0].h * ctx->lev[0].pitch : base + (size_t)k * frame_stride;
    ctx->l0_stride_host[s] = own ? ctx->lev[0].pitch : stride;
  }
  VS_CUDA(cudaMemcpyAsync(ctx->l0_ptr + first, ctx->l0_ptr_host + first, sizeof(uint8_t*) * count, cudaMemcpyHostToDevice, vs_in_stream(ctx)));
  VS_CUDA(cudaMemcpyAsync(ctx->l0_stride + first, ctx->l0_stride_host + first, sizeof(int) * count, cudaMemcpyHostToDevice, vs_in_stream(ctx)));
  return VSLAM_OK;
}

static int upload_frames_to(vslam_ctx* ctx, uint8_t* base, cudaStream_t st, int first, int count, const uint8_t* gray, int stride, size_t frame_stride) {
  const LevelDesc& L = ctx->lev[0];
  uint8_t* dst = base + (size_t)first * L.h * L.pitch;
  if (stride == L.pitch && frame_stride == (size_t)L.h * L.pitch) {
    VS_CUDA(cudaMemcpyAsync(dst, gray, (size_t)count * L.h * L.pitch, cudaMemcpyHostToDevice, st));
  } else {
    for (int k = 0; k < count; k++)
      VS_CUDA(cudaMemcpy2DAsync(dst + (size_t)k * L.h * L.pitch, L.pitch, gray + (size_t)k * frame_stride, stride, L.w, L.h, cudaMemcpyHostToDevice, st));
  }
  return VSLAM_OK;
}
static int upload_frames(vslam_ctx* ctx, int first, int count, const uint8_t* gray, int stride, size_t frame_stride) {
  vs_time_begin(ctx, VS_ST_H2D);
  int rc = upload_frames_to(ctx, ctx->lev[0].img, vs_in_stream(ctx), first, count, gray, stride, frame_stride);
  vs_time_end(ctx);
  return rc;
}

static int check_range(vslam_ctx* ctx, int first, int count, const void* p, int stride) {
  if (!ctx) return VSLAM_E_INVALID;
  if (!p || first < 0 || count < 1 || first + count > ctx->S || stride < ctx->lev[0].w) { ctx->err = "bad stream range, pointer or stride"; return VSLAM_E_INVALID; }
  return VSLAM_OK;
}

// Host frames: copy into the ctx-owned level-0 images (re-adopting them if a device buffer had been adopted).
static int stage_host_frames(vslam_ctx* ctx, int first, int count, const uint8_t* gray, int stride, size_t frame_stride) {
  int rc = check_range(ctx, first, count, gray, stride); if (rc) return rc;
  if ((rc = vs_ensure_own_l0(ctx))) return rc;
  bool was_adopted = false;
  for (int s = first; s < first + count; s++) was_adopted |= ctx->l0_stride_host[s] != ctx->lev[0].pitch || ctx->l0_ptr_host[s] != ctx->lev[0].img + (size_t)s * ctx->lev[0].h * ctx->lev[0].pitch;
  if (was_adopted && (rc = adopt_l0(ctx, first, count, nullptr, 0, 0, true))) return rc;
  return upload_frames(ctx, first, count, gray, stride, frame_stride);
}
// Device frames: adopt the caller's buffer as level 0 (zero copy).
static int adopt_device_frames(vslam_ctx* ctx, int first, int count, const uint8_t* gray, int stride, size_t frame_stride) {
  int rc = check_range(ctx, first, count, gray, stride); if (rc) return rc;
  if (((uintptr_t)gray & 15) || (stride & 15) || (frame_stride & 15)) { ctx->err = "device frames must be 16-byte aligned with stride % 16 == 0"; return VSLAM_E_INVALID; }
  bool same = true;
  for (int k = 0; k < count; k++) same &= ctx->l0_ptr_host[first + k] == gray + (size_t)k * frame_stride && ctx->l0_stride_host[first + k] == stride;
  if (!same && (rc = adopt_l0(ctx, first, count, gray, stride, frame_stride, false))) return rc;
  return VSLAM_OK;
}

int vslam_make_keyframe_lite(vslam_ctx* ctx, int first, int count, const uint8_t* gray, int stride, size_t frame_stride) {
  const int rc = stage_host_frames(ctx, first, count, gray, stride, frame_stride);
  return rc ? rc : vs_launch_pyramid_fast(ctx, first, count);
}

int vslam_make_keyframe_lite_dev(vslam_ctx* ctx, int first, int count, const uint8_t* gray, int stride, size_t frame_stride) {
  const int rc = adopt_device_frames(ctx, first, count, gray, stride, frame_stride);
  return rc ? rc : vs_launch_pyramid_fast(ctx, first, count);
}

// A map keyframe as a stream's current keyframe (corners + row LUT are needed of a search TARGET, and only streams carry them):
// zero-copy adoption of the source slot's level-0 image, then the usual pyramid + FAST pass.
int vslam_make_keyframe_from_source(vslam_ctx* ctx, int s, int kf_id) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  if (kf_id < 0 || kf_id >= ctx->n_src || !ctx->src_have[kf_id]) { ctx->err = "source keyframe was never uploaded"; return VSLAM_E_INVALID; }
  const size_t fs = (size_t)ctx->src.h[0] * ctx->src.pitch[0];
  return vslam_make_keyframe_lite_dev(ctx, s, 1, ctx->src.img[0] + (size_t)kf_id * fs, ctx->src.pitch[0], fs);
}

int vslam_level_dims(const vslam_ctx* ctx, int level, int* w, int* h) {
  if (!ctx || level < 0 || level >= VS_LEVELS) return VSLAM_E_INVALID;
  *w = ctx->lev[level].w; *h = ctx->lev[level].h; return VSLAM_OK;
}

int vslam_get_level(vslam_ctx* ctx, int s, int l, uint8_t* out, int out_stride) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  if (l < 0 || l >= VS_LEVELS || !out || out_stride < ctx->lev[l].w) { ctx->err = "bad level / stride"; return VSLAM_E_INVALID; }
  const LevelDesc& L = ctx->lev[l];
  if ((rc = vslam_sync(ctx))) return rc;
  const uint8_t* src = l == 0 ? ctx->l0_ptr_host[s] : L.img + (size_t)s * L.h * L.pitch;
  const int sp = l == 0 ? ctx->l0_stride_host[s] : L.pitch;
  VS_CUDA(cudaMemcpy2D(out, out_stride, src, sp, L.w, L.h, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

int vslam_get_num_corners(vslam_ctx* ctx, int s, int l, int* n) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  if (l < 0 || l >= VS_LEVELS || !n) return VSLAM_E_INVALID;
  if ((rc = vs_ensure_lists(ctx))) return rc;
  if ((rc = vslam_sync(ctx))) return rc;
  const LevelDesc& L = ctx->lev[l];
  VS_CUDA(cudaMemcpy(n, L.lut + (size_t)s * (L.h + 1) + L.h, sizeof(int), cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

int vslam_get_corners(vslam_ctx* ctx, int s, int l, int32_t* xy, int cap) {
  int n = 0; int rc = vslam_get_num_corners(ctx, s, l, &n); if (rc) return rc;
  if (n > cap) { ctx->err = "output buffer too small"; return VSLAM_E_INVALID; }
  std::vector<uint32_t> tmp(n);
  const LevelDesc& L = ctx->lev[l];
  if (n) VS_CUDA(cudaMemcpy(tmp.data(), L.corners + (size_t)s * L.cap, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) { xy[2 * i] = tmp[i] & 0xffff; xy[2 * i + 1] = tmp[i] >> 16; }
  return VSLAM_OK;
}

int vslam_get_row_lut(vslam_ctx* ctx, int s, int l, int32_t* lut) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  if (l < 0 || l >= VS_LEVELS || !lut) return VSLAM_E_INVALID;
  if ((rc = vs_ensure_lists(ctx))) return rc;
  if ((rc = vslam_sync(ctx))) return rc;
  const LevelDesc& L = ctx->lev[l];
  VS_CUDA(cudaMemcpy(lut, L.lut + (size_t)s * (L.h + 1), sizeof(int) * L.h, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

// ---------------------------------------------------------------------------------------------- MakeKeyFrame_Rest / MiniPatch
int vslam_make_keyframe_rest(vslam_ctx* ctx, int s) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  return vs_keyframe_rest(ctx, s);
}

static int rest_list(vslam_ctx* ctx, int s, int l, const uint32_t* src, int which, int32_t* xy, double* score, int cap, int* n) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  if (l < 0 || l >= VS_LEVELS || !n || ctx->rest_stream != s) { ctx->err = "call vslam_make_keyframe_rest for this stream first"; return VSLAM_E_INVALID; }
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  int counts[2];
  VS_CUDA(cudaMemcpy(counts, ctx->rest_counts + 2 * l, sizeof(counts), cudaMemcpyDeviceToHost));
  *n = counts[which];
  if (!xy) return VSLAM_OK;
  if (*n > cap) { ctx->err = "output buffer too small"; return VSLAM_E_INVALID; }
  std::vector<uint32_t> tmp(*n);
  if (*n) VS_CUDA(cudaMemcpy(tmp.data(), src + ctx->rest_off[l], sizeof(uint32_t) * *n, cudaMemcpyDeviceToHost));
  for (int i = 0; i < *n; i++) { xy[2 * i] = tmp[i] & 0xffff; xy[2 * i + 1] = tmp[i] >> 16; }
  if (score && *n) VS_CUDA(cudaMemcpy(score, ctx->rest_cand_score + ctx->rest_off[l], sizeof(double) * *n, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}
int vslam_get_max_corners(vslam_ctx* ctx, int s, int l, int32_t* xy, int cap, int* n) { return ctx ? rest_list(ctx, s, l, ctx->rest_max, 0, xy, nullptr, cap, n) : VSLAM_E_INVALID; }
int vslam_get_candidates(vslam_ctx* ctx, int s, int l, int32_t* xy, double* score, int cap, int* n) { return ctx ? rest_list(ctx, s, l, ctx->rest_cand, 1, xy, score, cap, n) : VSLAM_E_INVALID; }

int vslam_snapshot_keyframe(vslam_ctx* ctx, int s) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  if ((rc = vs_ensure_lists(ctx))) return rc;
  const LevelDesc& L = ctx->lev[0];
  if (!ctx->snap_img) {
    VS_CUDA(dalloc(&ctx->snap_img, (size_t)ctx->S * L.h * L.pitch)); VS_CUDA(dalloc(&ctx->snap_corners, (size_t)ctx->S * L.cap));
    VS_CUDA(dalloc(&ctx->snap_lut, (size_t)ctx->S * (L.h + 1)));
  }
  VS_CUDA(cudaMemcpy2DAsync(ctx->snap_img + (size_t)s * L.h * L.pitch, L.pitch, ctx->l0_ptr_host[s], ctx->l0_stride_host[s], L.w, L.h, cudaMemcpyDeviceToDevice, ctx->stream));
  VS_CUDA(cudaMemcpyAsync(ctx->snap_corners + (size_t)s * L.cap, L.corners + (size_t)s * L.cap, sizeof(uint32_t) * L.cap, cudaMemcpyDeviceToDevice, ctx->stream));
  VS_CUDA(cudaMemcpyAsync(ctx->snap_lut + (size_t)s * (L.h + 1), L.lut + (size_t)s * (L.h + 1), sizeof(int) * (L.h + 1), cudaMemcpyDeviceToDevice, ctx->stream));
  return VSLAM_OK;
}

static int minipatch_common(vslam_ctx* ctx, int s, int which, int n) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  if (n < 0 || (which != 0 && which != 1) || (which == 1 && !ctx->snap_img)) { ctx->err = "bad MiniPatch arguments (snapshot the keyframe first for which = 1)"; return VSLAM_E_INVALID; }
  return VSLAM_OK;
}

int vslam_minipatch_sample(vslam_ctx* ctx, int s, int which, int n, const int32_t* xy, uint8_t* patches81) {
  int rc = minipatch_common(ctx, s, which, n); if (rc) return rc;
  if (n == 0) return VSLAM_OK;
  const int W = ctx->lev[0].w, H = ctx->lev[0].h;
  for (int i = 0; i < n; i++) if (!(xy[2 * i] >= 4 && xy[2 * i + 1] >= 4 && xy[2 * i] < W - 4 && xy[2 * i + 1] < H - 4)) { ctx->err = "MiniPatch centre within 4 px of the border (the reference asserts, jni/MiniPatch.cc:75)"; return VSLAM_E_INVALID; }
  int* dxy = nullptr; uint8_t* dp = nullptr;
  VS_CUDA(cudaMalloc(&dxy, sizeof(int) * 2 * n)); VS_CUDA(cudaMalloc(&dp, (size_t)81 * n));
  VS_CUDA(cudaMemcpyAsync(dxy, xy, sizeof(int) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
  rc = vs_minipatch_sample(ctx, s, which, dxy, n, dp);
  if (!rc) { cudaError_t e = cudaMemcpyAsync(patches81, dp, (size_t)81 * n, cudaMemcpyDeviceToHost, ctx->stream); if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = VSLAM_E_CUDA; } }
  cudaFree(dxy); cudaFree(dp);
  return rc;
}

int vslam_minipatch_find(vslam_ctx* ctx, int s, int which, int n, const uint8_t* patches81, double* pos2, int32_t* found, int32_t* best_ssd, int range, int max_ssd) {
  int rc = minipatch_common(ctx, s, which, n); if (rc) return rc;
  if (n == 0) return VSLAM_OK;
  uint8_t* dp = nullptr; double* dpos = nullptr; int* df = nullptr; int* db = nullptr;
  VS_CUDA(cudaMalloc(&dp, (size_t)81 * n)); VS_CUDA(cudaMalloc(&dpos, sizeof(double) * 2 * n)); VS_CUDA(cudaMalloc(&df, sizeof(int) * n)); VS_CUDA(cudaMalloc(&db, sizeof(int) * n));
  VS_CUDA(cudaMemcpyAsync(dp, patches81, (size_t)81 * n, cudaMemcpyHostToDevice, ctx->stream));
  VS_CUDA(cudaMemcpyAsync(dpos, pos2, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
  rc = vs_minipatch_find(ctx, s, which, dp, n, dpos, df, db, range, max_ssd);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(pos2, dpos, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(found, df, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && best_ssd) e = cudaMemcpyAsync(best_ssd, db, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = VSLAM_E_CUDA; }
  }
  cudaFree(dp); cudaFree(dpos); cudaFree(df); cudaFree(db);
  return rc;
}

// ---------------------------------------------------------------------------------------------- per-stream state
static int read_ss(vslam_ctx* ctx, int s, StreamState* st) {
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  VS_CUDA(cudaMemcpy(st, ctx->ss + s, sizeof(StreamState), cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}
static int write_ss(vslam_ctx* ctx, int s, const StreamState* st) {
  VS_CUDA(cudaMemcpy(ctx->ss + s, st, sizeof(StreamState), cudaMemcpyHostToDevice));
  return VSLAM_OK;
}

int vslam_set_pose(vslam_ctx* ctx, int s, const double* p) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  VS_CUDA(cudaMemcpy((char*)(ctx->ss + s) + offsetof(StreamState, pose), p, sizeof(double) * 12, cudaMemcpyHostToDevice));
  return VSLAM_OK;
}
int vslam_get_pose(vslam_ctx* ctx, int s, double* p) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  VS_CUDA(cudaMemcpy(p, (char*)(ctx->ss + s) + offsetof(StreamState, pose), sizeof(double) * 12, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}
int vslam_get_poses(vslam_ctx* ctx, double* p) {
  if (!ctx || !p) return VSLAM_E_INVALID;
  VS_CUDA(cudaMemcpy2DAsync(p, sizeof(double) * 12, (char*)ctx->ss + offsetof(StreamState, pose), sizeof(StreamState), sizeof(double) * 12, ctx->S, cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  return VSLAM_OK;
}
int vslam_set_motion(vslam_ctx* ctx, int s, const double* v6, double msd, double dmean, double dsigma) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  StreamState st; if ((rc = read_ss(ctx, s, &st))) return rc;
  memcpy(st.velocity, v6, sizeof(st.velocity)); st.msd_scaled_vel = msd; st.depth_mean = dmean; st.depth_sigma = dsigma;
  return write_ss(ctx, s, &st);
}
int vslam_get_motion(vslam_ctx* ctx, int s, double* v6, double* msd, double* dmean, double* dsigma) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  StreamState st; if ((rc = read_ss(ctx, s, &st))) return rc;
  memcpy(v6, st.velocity, sizeof(st.velocity)); *msd = st.msd_scaled_vel; *dmean = st.depth_mean; *dsigma = st.depth_sigma;
  return VSLAM_OK;
}
int vslam_reset_stream(vslam_ctx* ctx, int s) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  StreamState st; if ((rc = read_ss(ctx, s, &st))) return rc;
  st.did_coarse = 0; st.quality = 2; st.lost_frames = 0; st.msd_scaled_vel = 0.0; st.vel_mag = 0.0; st.depth_mean = 1.0; st.depth_sigma = 1.0;
  st.just_recovered = 0; st.n_updates = 0; st.try_coarse = 0; st.recovered = 0;
  st.frame_no = 0; st.last_kf_dropped = -20; st.kf_request = 0; st.kf_closest = -1; st.kf_dist = 0.0;
  for (int k = 0; k < 6; k++) st.velocity[k] = 0.0;
  for (int l = 0; l < VS_LEVELS; l++) st.attempted[l] = st.found[l] = 0;
  return write_ss(ctx, s, &st);
}
// best keyframe and final alignment score of the stream's last relocalisation attempt, recoveries so far, and whether the last frame recovered it
int vslam_get_reloc_info(vslam_ctx* ctx, int s, int* best, double* score, int* n_recoveries, int* recovered_last_frame) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  StreamState st; if ((rc = read_ss(ctx, s, &st))) return rc;
  if (best) *best = st.reloc_best; if (score) *score = st.reloc_score; if (n_recoveries) *n_recoveries = st.n_recoveries; if (recovered_last_frame) *recovered_last_frame = st.recovered;
  return VSLAM_OK;
}
// ---- keyframe hand-off: the MapMaker heuristics Tracker::TrackFrame consults (jni/Tracker.cc:127-132, :869-871) ----------------------
int vslam_set_keyframe_policy(vslam_ctx* ctx, int enable, double wiggle_scale, double wiggle_scale_depth_normalized, double max_kf_dist_wiggle_mult, int min_frames_between) {
  if (!ctx) return VSLAM_E_INVALID;
  if (enable && !(wiggle_scale > 0.0 && wiggle_scale_depth_normalized > 0.0 && max_kf_dist_wiggle_mult > 0.0 && min_frames_between >= 0)) { ctx->err = "keyframe policy: scales must be positive"; return VSLAM_E_INVALID; }
  ctx->kf_policy = enable != 0;
  if (enable) { ctx->kf_wiggle = wiggle_scale; ctx->kf_wiggle_dn = wiggle_scale_depth_normalized; ctx->kf_mult = max_kf_dist_wiggle_mult; ctx->kf_min_frames = min_frames_between; }
  return VSLAM_OK;
}
int vslam_get_keyframe_requests(vslam_ctx* ctx, int32_t* request, int32_t* closest, double* dist) {
  if (!ctx || !request) return VSLAM_E_INVALID;
  int rc = vslam_sync(ctx); if (rc) return rc;
  if (!closest && !dist) {   // the per-frame poll: S ints
    VS_CUDA(cudaMemcpy(request, ctx->kf_req, sizeof(int) * ctx->S, cudaMemcpyDeviceToHost));
    return VSLAM_OK;
  }
  std::vector<StreamState> h(ctx->S);
  VS_CUDA(cudaMemcpy(h.data(), ctx->ss, sizeof(StreamState) * ctx->S, cudaMemcpyDeviceToHost));
  for (int s = 0; s < ctx->S; s++) { request[s] = h[s].kf_request; if (closest) closest[s] = h[s].kf_closest; if (dist) dist[s] = h[s].kf_dist; }
  return VSLAM_OK;
}
// Tracker::AddNewKeyFrame (jni/Tracker.cc:823-827) + MapMaker::AddKeyFrame (jni/MapMaker.cc:470-478: `*pK = k`): the stream's current
// keyframe (all four levels, device to device) becomes source keyframe `kf_id` at the stream's pose, joins the registered keyframes
// (relocaliser + keyframe policy) and the stream's mnLastKeyFrameDropped is set.
int vslam_add_keyframe_from_stream(vslam_ctx* ctx, int s, int kf_id) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  if (kf_id < 0 || kf_id >= ctx->n_src) { ctx->err = "keyframe id exceeds max_source_keyframes"; return VSLAM_E_CAPACITY; }
  if (!ctx->sbi_on) { ctx->err = "vslam_add_keyframe_from_stream needs vslam_enable_sbi first (the keyframe's SmallBlurryImage is built for the relocaliser)"; return VSLAM_E_INVALID; }
  StreamState st; if ((rc = read_ss(ctx, s, &st))) return rc;
  for (int l = 0; l < VS_LEVELS; l++) {
    const LevelDesc& L = ctx->lev[l];
    const uint8_t* from = l == 0 ? ctx->l0_ptr_host[s] : L.img + (size_t)s * L.h * L.pitch;
    const int fp = l == 0 ? ctx->l0_stride_host[s] : L.pitch;
    if (!from) { ctx->err = "stream has no current keyframe"; return VSLAM_E_INVALID; }
    uint8_t* to = ctx->src.img[l] + (size_t)kf_id * ctx->src.h[l] * ctx->src.pitch[l];
    if (to == from) continue;   // the stream's level 0 IS this slot (vslam_make_keyframe_from_source of the same keyframe): nothing to copy
    VS_CUDA(cudaMemcpy2DAsync(to, ctx->src.pitch[l], from, fp, L.w, L.h, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  ctx->src_have[kf_id] = 1;
  std::vector<int> ids = ctx->reloc_ids; std::vector<double> poses = ctx->reloc_poses_host;
  size_t at = 0; while (at < ids.size() && ids[at] != kf_id) at++;
  if (at == ids.size()) { ids.push_back(kf_id); poses.resize(12 * ids.size()); }
  memcpy(&poses[12 * at], st.pose, sizeof(double) * 12);
  if ((rc = vslam_set_reloc_keyframes(ctx, (int)ids.size(), ids.data(), poses.data()))) return rc;
  if ((rc = read_ss(ctx, s, &st))) return rc;
  st.last_kf_dropped = st.frame_no; st.kf_request = 0;
  VS_CUDA(memset_sync(ctx->kf_req + s, 0, sizeof(int)));
  return write_ss(ctx, s, &st);
}

// Test / hand-off hook: Tracker::mnLostFrames and mTrackingQuality of a stream
int vslam_set_lost(vslam_ctx* ctx, int s, int lost_frames, int quality) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  StreamState st; if ((rc = read_ss(ctx, s, &st))) return rc;
  st.lost_frames = lost_frames; st.quality = quality;
  return write_ss(ctx, s, &st);
}

int vslam_set_sbi_rotation(vslam_ctx* ctx, int s, const double* r6) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  VS_CUDA(cudaMemcpy((char*)(ctx->ss + s) + offsetof(StreamState, sbi_rot), r6, sizeof(double) * 6, cudaMemcpyHostToDevice));
  return VSLAM_OK;
}
int vslam_get_sbi_rotation(vslam_ctx* ctx, int s, double* r6) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  VS_CUDA(cudaMemcpy(r6, (char*)(ctx->ss + s) + offsetof(StreamState, sbi_rot), sizeof(double) * 6, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}
int vslam_get_counters(vslam_ctx* ctx, int s, int32_t* att, int32_t* fnd, int* quality, int* lost, int* did_coarse) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  StreamState st; if ((rc = read_ss(ctx, s, &st))) return rc;
  for (int l = 0; l < VS_LEVELS; l++) { att[l] = st.attempted[l]; fnd[l] = st.found[l]; }
  *quality = st.quality; *lost = st.lost_frames; *did_coarse = st.did_coarse;
  return VSLAM_OK;
}
int vslam_get_updates(vslam_ctx* ctx, int s, double* upd6, double* sig, int cap, int* n) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  StreamState st; if ((rc = read_ss(ctx, s, &st))) return rc;
  *n = st.n_updates;
  for (int k = 0; k < st.n_updates && k < cap; k++) { memcpy(upd6 + 6 * k, st.updates + 6 * k, sizeof(double) * 6); sig[k] = st.sigmas[k]; }
  return VSLAM_OK;
}
int vslam_get_zmssd_evals(vslam_ctx* ctx, unsigned long long* total) {
  if (!ctx || !total) return VSLAM_E_INVALID;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  VS_CUDA(cudaMemcpy(total, ctx->evals, sizeof(*total), cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

int vslam_get_search_stats(vslam_ctx* ctx, unsigned long long* stats4) {
  if (!ctx || !stats4) return VSLAM_E_INVALID;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  VS_CUDA(cudaMemcpy(stats4, ctx->evals, 4 * sizeof(*stats4), cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

int vslam_get_point_states(vslam_ctx* ctx, int s, int32_t* ints, double* dbl) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  const int n = ctx->map.n, N = ctx->N; const size_t SN = (size_t)ctx->S * N, o = (size_t)s * N;
  std::vector<double> buf((size_t)36 * n); std::vector<int> fl(n), lv(n);
  auto grab = [&](const double* src, int comps, double* dst) -> cudaError_t {
    for (int c = 0; c < comps; c++) { cudaError_t e = cudaMemcpy(dst + (size_t)c * n, src + c * SN + o, sizeof(double) * n, cudaMemcpyDeviceToHost); if (e != cudaSuccess) return e; }
    return cudaSuccess;
  };
  double* b = buf.data();
  const PointState& ps = ctx->ps;
  VS_CUDA(grab(ps.v2image, 2, b)); VS_CUDA(grab(ps.v2found, 2, b + 2 * n)); VS_CUDA(grab(ps.derivs, 4, b + 4 * n)); VS_CUDA(grab(ps.v3cam, 3, b + 8 * n));
  VS_CUDA(grab(ps.warpinv, 4, b + 11 * n)); VS_CUDA(grab(ps.sqrtinv, 1, b + 15 * n)); VS_CUDA(grab(ps.jac, 12, b + 16 * n)); VS_CUDA(grab(ps.err, 2, b + 28 * n));
  VS_CUDA(grab(ps.coarse, 2, b + 30 * n));
  VS_CUDA(cudaMemcpy(fl.data(), ps.flags + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
  VS_CUDA(cudaMemcpy(lv.data(), ps.level + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) {
    int32_t* I = ints + 8 * (size_t)i; double* Dd = dbl + 32 * (size_t)i;
    const int f = fl[i];
    I[0] = !!(f & F_INIMAGE); I[1] = lv[i]; I[2] = !!(f & F_SEARCHED); I[3] = !!(f & F_FOUND); I[4] = !!(f & F_SUBPIX); I[5] = !!(f & F_TBAD); I[6] = !!(f & F_HASTD); I[7] = !!(f & F_NEWTMPL);
    for (int c = 0; c < 32; c++) Dd[c] = b[(size_t)c * n + i];
  }
  return VSLAM_OK;
}

int vslam_get_point_template(vslam_ctx* ctx, int s, int i, uint8_t* tmpl, int* sum, int* sumsq) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  if (i < 0 || i >= ctx->map.n) return VSLAM_E_INVALID;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  const size_t gi = (size_t)s * ctx->N + i, SN = (size_t)ctx->S * ctx->N;
  uint8_t rows[VS_TMPL_BYTES];   // device layout: one row per 12 bytes
  VS_CUDA(cudaMemcpy(rows, ctx->ps.tmpl + gi * VS_TMPL_BYTES, VS_TMPL_BYTES, cudaMemcpyDeviceToHost));
  for (int r = 0; r < ctx->P; r++) memcpy(tmpl + r * ctx->P, rows + 12 * r, ctx->P);
  VS_CUDA(cudaMemcpy(sum, ctx->ps.tsum + gi, sizeof(int), cudaMemcpyDeviceToHost));
  VS_CUDA(cudaMemcpy(sumsq, ctx->ps.tsum + SN + gi, sizeof(int), cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

int vslam_get_point_counts(vslam_ctx* ctx, int s, int32_t* oi) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  const int n = ctx->map.n; const size_t SN = (size_t)ctx->S * ctx->N, o = (size_t)s * ctx->N;
  std::vector<int> a(n), b(n);
  VS_CUDA(cudaMemcpy(a.data(), ctx->ps.counts + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
  VS_CUDA(cudaMemcpy(b.data(), ctx->ps.counts + SN + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) { oi[2 * i] = a[i]; oi[2 * i + 1] = b[i]; }
  return VSLAM_OK;
}

// ---------------------------------------------------------------------------------------------- stages
int vslam_project_all(vslam_ctx* ctx) { if (!ctx) return VSLAM_E_INVALID; return vs_launch_project_all(ctx, 0); }

int vslam_set_point_projection(vslam_ctx* ctx, int s, const double* v2image, const double* warp, const int32_t* level) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  if (!v2image || !warp || !level) return VSLAM_E_INVALID;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  const int n = ctx->map.n; const size_t SN = (size_t)ctx->S * ctx->N, o = (size_t)s * ctx->N;
  std::vector<double> col(n);
  for (int c = 0; c < 2; c++) { for (int i = 0; i < n; i++) col[i] = v2image[2 * i + c]; VS_CUDA(cudaMemcpy(ctx->ps.v2image + c * SN + o, col.data(), sizeof(double) * n, cudaMemcpyHostToDevice)); }
  for (int c = 0; c < 4; c++) { for (int i = 0; i < n; i++) col[i] = warp[4 * i + c]; VS_CUDA(cudaMemcpy(ctx->ps.warpinv + c * SN + o, col.data(), sizeof(double) * n, cudaMemcpyHostToDevice)); }
  VS_CUDA(cudaMemcpy(ctx->ps.level + o, level, sizeof(int) * n, cudaMemcpyHostToDevice));
  // m2 = inverse(warp) * 2^level, the same IEEE operations as calc_level_warp (track.cu)
  for (int c = 0; c < 4; c++) {
    for (int i = 0; i < n; i++) {
      const double w00 = warp[4 * i], w01 = warp[4 * i + 1], w10 = warp[4 * i + 2], w11 = warp[4 * i + 3];
      const double invdet = 1.0 / (w00 * w11 - w01 * w10);
      const int sc = level[i] >= 0 ? 1 << level[i] : 1;
      const double a = c == 0 ? w11 : (c == 1 ? -w01 : (c == 2 ? -w10 : w00));
      col[i] = (a * invdet) * sc;
    }
    VS_CUDA(cudaMemcpy(ctx->ps.m2 + c * SN + o, col.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
  }
  return VSLAM_OK;
}

int vslam_set_lists(vslam_ctx* ctx, const int32_t* idx, const int32_t* n, int idx_stride) {
  if (!ctx || !idx || !n) return VSLAM_E_INVALID;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<StreamState> h(ctx->S);
  VS_CUDA(cudaMemcpy(h.data(), ctx->ss, sizeof(StreamState) * ctx->S, cudaMemcpyDeviceToHost));
  for (int s = 0; s < ctx->S; s++) {
    if (n[s] < 0 || n[s] > ctx->list_cap) { ctx->err = "list longer than max_points"; return VSLAM_E_INVALID; }
    for (int k = 0; k < n[s]; k++) if (idx[(size_t)s * idx_stride + k] < 0 || idx[(size_t)s * idx_stride + k] >= ctx->map.n) { ctx->err = "list entry is not a map point"; return VSLAM_E_INVALID; }
    if (n[s]) VS_CUDA(cudaMemcpy(ctx->lists + (size_t)s * ctx->list_cap, idx + (size_t)s * idx_stride, sizeof(int) * n[s], cudaMemcpyHostToDevice));
    h[s].nA = n[s]; h[s].nB = 0; h[s].nB_top = 0;
  }
  VS_CUDA(cudaMemcpy(ctx->ss, h.data(), sizeof(StreamState) * ctx->S, cudaMemcpyHostToDevice));
  return VSLAM_OK;
}

int vslam_clear_counters(vslam_ctx* ctx) {
  if (!ctx) return VSLAM_E_INVALID;
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<StreamState> h(ctx->S);
  VS_CUDA(cudaMemcpy(h.data(), ctx->ss, sizeof(StreamState) * ctx->S, cudaMemcpyDeviceToHost));
  for (int s = 0; s < ctx->S; s++) { for (int l = 0; l < VS_LEVELS; l++) h[s].attempted[l] = h[s].found[l] = 0; h[s].n_updates = 0; }
  VS_CUDA(cudaMemcpy(ctx->ss, h.data(), sizeof(StreamState) * ctx->S, cudaMemcpyHostToDevice));
  return VSLAM_OK;
}

int vslam_search_for_points(vslam_ctx* ctx, int range, int subpix) { if (!ctx) return VSLAM_E_INVALID; return vs_launch_search(ctx, 0, range, subpix, 0); }

// MapMaker::ReFind_Common (jni/MapMaker.cc:967-1036) for every (stream = keyframe, listed point) pair: projection with the stream's
// pose, cold-finder template, FindPatchCoarse with `range` (the reference uses 4), sub-pixel refinement (8 iterations) on levels > 0.
int vslam_refind(vslam_ctx* ctx, int range, int subpix_its) {
  if (!ctx) return VSLAM_E_INVALID;
  int rc = vs_launch_project_all(ctx, 0); if (rc) return rc;
  return vs_launch_search(ctx, 0, range, subpix_its, 1);
}
// Results of vslam_refind for the current list of `stream`, in list order: flags3[k] = {found, level, bSubPix}, pos2[k] = v2RootPos
int vslam_get_refind_results(vslam_ctx* ctx, int s, int32_t* flags3, double* pos2, int cap, int* n_out) {
  int rc = check_stream(ctx, s); if (rc) return rc;
  if (!flags3 || !pos2 || !n_out) return VSLAM_E_INVALID;
  StreamState st; if ((rc = read_ss(ctx, s, &st))) return rc;
  const int n = st.nA < cap ? st.nA : cap;
  *n_out = st.nA;
  const int N = ctx->N; const size_t SN = (size_t)ctx->S * N, o = (size_t)s * N;
  std::vector<int> list(n > 0 ? n : 1), fl(ctx->map.n), lv(ctx->map.n); std::vector<double> f0(ctx->map.n), f1(ctx->map.n);
  if (n > 0) VS_CUDA(cudaMemcpy(list.data(), ctx->lists + (size_t)s * ctx->list_cap, sizeof(int) * n, cudaMemcpyDeviceToHost));
  VS_CUDA(cudaMemcpy(fl.data(), ctx->ps.flags + o, sizeof(int) * ctx->map.n, cudaMemcpyDeviceToHost));
  VS_CUDA(cudaMemcpy(lv.data(), ctx->ps.rlevel + o, sizeof(int) * ctx->map.n, cudaMemcpyDeviceToHost));
  VS_CUDA(cudaMemcpy(f0.data(), ctx->ps.v2found + o, sizeof(double) * ctx->map.n, cudaMemcpyDeviceToHost));
  VS_CUDA(cudaMemcpy(f1.data(), ctx->ps.v2found + SN + o, sizeof(double) * ctx->map.n, cudaMemcpyDeviceToHost));
  for (int k = 0; k < n; k++) {
    const int i = list[k];
    flags3[3 * k] = !!(fl[i] & F_FOUND); flags3[3 * k + 1] = lv[i]; flags3[3 * k + 2] = !!(fl[i] & F_SUBPIX);
    pos2[2 * k] = f0[i]; pos2[2 * k + 1] = f1[i];
  }
  return VSLAM_OK;
}
// ---- MapMaker::AddPointEpipolar, the search (jni/MapMaker.cc:525-640) ------------------------------------------------------------
namespace {
// ATANCamera::UnProject (jni/ATANCamera.cc:149-164) with invrtrans (jni/ATANCamera.h:145-150); host libm, like the reference
void host_unproject(const CamDev& c, double x, double y, double* out) {
  const double dx = (x - c.cx) * (1.0 / c.fx), dy = (y - c.cy) * (1.0 / c.fy);
  const double distR = sqrt(dx * dx + dy * dy);
  const double R = (c.W == 0.0) ? distR : (tan(distR * c.W) * c.oneOver2Tan);
  const double factor = (distR > 0.01) ? R / distR : 1.0;
  out[0] = dx * factor; out[1] = dy * factor;
}
void rot_vec(const double* P, const double* v, double* o) { for (int i = 0; i < 3; i++) { double s = P[4 * i] * v[0]; s += P[4 * i + 1] * v[1]; s += P[4 * i + 2] * v[2]; o[i] = s; } }
void rot_t_vec(const double* P, const double* v, double* o) { for (int i = 0; i < 3; i++) { double s = P[i] * v[0]; s += P[4 + i] * v[1]; s += P[8 + i] * v[2]; o[i] = s; } }
}  // namespace

int vslam_epipolar_search(vslam_ctx* ctx, int stream, int src_kf, int level, int n, const int32_t* cand_xy, const double* src_pose, const double* tgt_pose,
                          double depth_mean, double depth_sigma, double wiggle_scale, int32_t* found, double* pos2, int32_t* best_corner, int32_t* best_zmssd) {
  int rc = check_stream(ctx, stream); if (rc) return rc;
  if (!cand_xy || !src_pose || !tgt_pose || !found || !pos2 || n < 0 || level < 0 || level >= VS_LEVELS || src_kf < 0 || src_kf >= ctx->n_src) { ctx->err = "bad argument"; return VSLAM_E_INVALID; }
  const CamDev& cam = ctx->cam;
  const int W = ctx->lev[0].w, H = ctx->lev[0].h;
  if (!ctx->unproj_ok) {   // imUnProj: UnProject of every integer pixel (jni/MapMaker.cc:530-540)
    std::vector<double> lut((size_t)2 * W * H);
    for (int j = 0; j < H; j++) for (int i = 0; i < W; i++) host_unproject(cam, (double)i, (double)j, &lut[2 * ((size_t)j * W + i)]);
    if (!ctx->unproj_lut) VS_CUDA(dalloc(&ctx->unproj_lut, lut.size()));
    VS_CUDA(cudaMemcpy(ctx->unproj_lut, lut.data(), sizeof(double) * lut.size(), cudaMemcpyHostToDevice));
    ctx->unproj_ok = true;
  }
  // mdOnePixelDist (jni/ATANCamera.cc:86-91)
  double uc[2], ur[2]; host_unproject(cam, cam.width / 2, cam.height / 2, uc); host_unproject(cam, cam.width / 2 + 1.0, cam.height / 2 + 1.0, ur);
  const double ddx = uc[0] - ur[0], ddy = uc[1] - ur[1];
  double dd = ddx * ddx; dd += ddy * ddy;
  const double onePixelDist = sqrt(dd) / sqrt(2.0);
  const int nLevelScale = 1 << level;
  const double dMaxDistDiff = onePixelDist * (4.0 + 1.0 * nLevelScale), dMaxDistSq = dMaxDistDiff * dMaxDistDiff;
  // The ray of every candidate through the camera model on the host (ATANCamera::UnProject calls libm's tan, like the imUnProj table above);
  // the line geometry of jni/MapMaker.cc:543-591 -- rotations, depth range, clipping, the line's normal form -- runs on the device
  // (k_epipolar_geometry, one thread per candidate, the reference's operations in the reference's order).
  if (n == 0) return VSLAM_OK;
  std::vector<double> rays(2 * (size_t)n);
  for (int k = 0; k < n; k++) {
    const double root0 = ((double)cand_xy[2 * k] + 0.5) * nLevelScale - 0.5, root1 = ((double)cand_xy[2 * k + 1] + 0.5) * nLevelScale - 0.5;   // LevelZeroPos
    host_unproject(cam, root0, root1, &rays[2 * (size_t)k]);
  }
  // one scratch allocation of the context, grown on demand: [EpiCand n][rays 2n f64][pos 2n f64][xy 2n i32][out 3n i32]
  const size_t need = sizeof(EpiCand) * n + sizeof(double) * 4 * n + sizeof(int) * 5 * n + 64;
  if (need > ctx->epi_cap) {
    VS_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->epi_buf); ctx->epi_buf = nullptr; ctx->epi_cap = 0;
    VS_CUDA(cudaMalloc(&ctx->epi_buf, need * 2));
    ctx->epi_cap = need * 2;
  }
  EpiCand* cd = (EpiCand*)ctx->epi_buf;
  double* rays_d = (double*)(cd + n); double* op = rays_d + 2 * (size_t)n;
  int* xy_d = (int*)(op + 2 * (size_t)n); int* oi = xy_d + 2 * (size_t)n;
  VS_CUDA(cudaMemcpyAsync(rays_d, rays.data(), sizeof(double) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
  VS_CUDA(cudaMemcpyAsync(xy_d, cand_xy, sizeof(int) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
  EpiGeom G;
  for (int q = 0; q < 12; q++) { G.src_pose[q] = src_pose[q]; G.tgt_pose[q] = tgt_pose[q]; }
  G.start_depth = std::max(wiggle_scale, depth_mean - depth_sigma); G.end_depth = std::min(40 * wiggle_scale, depth_mean + depth_sigma);
  G.max_dist_sq = dMaxDistSq; G.largest_radius = cam.largestRadius;
  if ((rc = vs_launch_epipolar_geometry(ctx, n, G, rays_d, xy_d, cd))) return rc;
  if ((rc = vs_launch_epipolar(ctx, stream, src_kf, level, n, cd, ctx->unproj_lut, 10, oi, op))) return rc;
  std::vector<int> hi(3 * (size_t)n);
  VS_CUDA(cudaMemcpyAsync(hi.data(), oi, sizeof(int) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(cudaMemcpyAsync(pos2, op, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < n; k++) { found[k] = hi[3 * k]; if (best_corner) best_corner[k] = hi[3 * k + 1]; if (best_zmssd) best_zmssd[k] = hi[3 * k + 2]; }
  return VSLAM_OK;
}

// ---- MapMaker::AddPointEpipolar, the new map point (jni/MapMaker.cc:646-690) -------------------------------------------------------
namespace {
// Right singular vector of the smallest singular value of a 4x4 matrix, by one-sided (Hestenes) Jacobi rotations of its columns.
// The reference asks Eigen's JacobiSVD for the same vector (jni/MapMaker.cc:191-192); Eigen is not part of the reference tree, the
// vector is unique up to sign, and the sign cancels in the dehomogenisation that follows.
void smallest_right_singular_vector4(const double A[4][4], double v[4]) {
  double U[4][4], V[4][4];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { U[i][j] = A[i][j]; V[i][j] = i == j ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0.0;
    for (int p = 0; p < 3; p++) for (int q = p + 1; q < 4; q++) {
      double alpha = 0, beta = 0, gamma = 0;
      for (int i = 0; i < 4; i++) { alpha += U[i][p] * U[i][p]; beta += U[i][q] * U[i][q]; gamma += U[i][p] * U[i][q]; }
      if (gamma == 0.0) continue;
      off = std::max(off, fabs(gamma) / sqrt(alpha * beta));
      const double zeta = (beta - alpha) / (2.0 * gamma);
      const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
      for (int i = 0; i < 4; i++) {
        const double up = U[i][p], uq = U[i][q]; U[i][p] = c * up - sn * uq; U[i][q] = sn * up + c * uq;
        const double vp = V[i][p], vq = V[i][q]; V[i][p] = c * vp - sn * vq; V[i][q] = sn * vp + c * vq;
      }
    }
    if (off < 1e-15) break;
  }
  int best = 0; double bn = -1;
  for (int j = 0; j < 4; j++) { double nn = 0; for (int i = 0; i < 4; i++) nn += U[i][j] * U[i][j]; if (bn < 0 || nn < bn) { bn = nn; best = j; } }
  for (int i = 0; i < 4; i++) v[i] = V[i][best];
}
void unit3(double* v) { double nn = v[0] * v[0]; nn += v[1] * v[1]; nn += v[2] * v[2]; const double nrm = sqrt(nn); v[0] /= nrm; v[1] /= nrm; v[2] /= nrm; }
}  // namespace

int vslam_epipolar_make_points(vslam_ctx* ctx, int level, int n, const int32_t* cand_xy, const double* found_pos2, const double* src_pose, const double* tgt_pose,
                               double* world3, double* pixel_right3, double* pixel_down3, int32_t* ir_center2, int32_t* src_level) {
  if (!ctx || n < 0 || level < 0 || level >= VS_LEVELS || !cand_xy || !found_pos2 || !src_pose || !tgt_pose || !world3 || !pixel_right3 || !pixel_down3 || !ir_center2 || !src_level) {
    if (ctx) ctx->err = "bad argument"; return VSLAM_E_INVALID; }
  const CamDev& cam = ctx->cam;
  const int scale = 1 << level;
  // se3AfromB = kSrc.se3CfromW * kTarget.se3CfromW.inverse(): R = Rs Rt^T, t = ts - R tt
  double P[12];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double a = src_pose[4 * i] * tgt_pose[4 * j]; a += src_pose[4 * i + 1] * tgt_pose[4 * j + 1]; a += src_pose[4 * i + 2] * tgt_pose[4 * j + 2]; P[4 * i + j] = a; }
  for (int i = 0; i < 3; i++) { double a = P[4 * i] * tgt_pose[3]; a += P[4 * i + 1] * tgt_pose[7]; a += P[4 * i + 2] * tgt_pose[11]; P[4 * i + 3] = src_pose[4 * i + 3] - a; }
  for (int k = 0; k < n; k++) {
    const double root[2] = {((double)cand_xy[2 * k] + 0.5) * scale - 0.5, ((double)cand_xy[2 * k + 1] + 0.5) * scale - 0.5};   // LevelZeroPos
    double a2[2], b2[2]; host_unproject(cam, root[0], root[1], a2); host_unproject(cam, found_pos2[2 * k], found_pos2[2 * k + 1], b2);
    // MapMaker::ReprojectPoint (jni/MapMaker.cc:176-200): the point, in the target camera's frame, that best satisfies both viewing rays
    double A[4][4] = {{-1.0, 0.0, b2[0], 0.0}, {0.0, -1.0, b2[1], 0.0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    for (int j = 0; j < 4; j++) { A[2][j] = a2[0] * P[8 + j] - P[j]; A[3][j] = a2[1] * P[8 + j] - P[4 + j]; }
    double v4[4]; smallest_right_singular_vector4(A, v4);
    if (v4[3] == 0.0) v4[3] = 0.00001;
    const double in_target[3] = {v4[0] / v4[3], v4[1] / v4[3], v4[2] / v4[3]};
    // v3New = kTarget.se3CfromW.inverse() * point = Rt^T (point - tt)
    const double d[3] = {in_target[0] - tgt_pose[3], in_target[1] - tgt_pose[7], in_target[2] - tgt_pose[11]};
    double* w = world3 + 3 * k; rot_t_vec(tgt_pose, d, w);
    // patch source fields (jni/MapMaker.cc:655-684) and MapPoint::RefreshPixelVectors (jni/MapPoint.cc:4-29) with the plane normal (0,0,-1)
    double uc[2], ur[2], ud[2];
    host_unproject(cam, root[0], root[1], uc); host_unproject(cam, root[0] + scale, root[1], ur); host_unproject(cam, root[0], root[1] + scale, ud);
    double c3[3] = {uc[0], uc[1], 1.0}, r3[3] = {ur[0], ur[1], 1.0}, d3[3] = {ud[0], ud[1], 1.0};
    unit3(c3); unit3(d3); unit3(r3);
    double pc[3]; rot_vec(src_pose, w, pc); pc[0] += src_pose[3]; pc[1] += src_pose[7]; pc[2] += src_pose[11];
    const double height = fabs(pc[2]);                       // |v3PlanePoint_C . (0,0,-1)|
    double con[3], ron[3], don[3];
    for (int q = 0; q < 3; q++) { con[q] = c3[q] * height / fabs(c3[2]); ron[q] = r3[q] * height / fabs(r3[2]); don[q] = d3[q] * height / fabs(d3[2]); }
    const double dr[3] = {ron[0] - con[0], ron[1] - con[1], ron[2] - con[2]}, dd[3] = {don[0] - con[0], don[1] - con[1], don[2] - con[2]};
    rot_t_vec(src_pose, dr, pixel_right3 + 3 * k); rot_t_vec(src_pose, dd, pixel_down3 + 3 * k);
    ir_center2[2 * k] = cand_xy[2 * k]; ir_center2[2 * k + 1] = cand_xy[2 * k + 1]; src_level[k] = level;
  }
  return VSLAM_OK;
}

int vslam_project_and_derivs(vslam_ctx* ctx, int only_found) { if (!ctx) return VSLAM_E_INVALID; return vs_launch_project_and_derivs(ctx, only_found); }
int vslam_calc_jacobians(vslam_ctx* ctx) { if (!ctx) return VSLAM_E_INVALID; return vs_launch_calc_jacobians(ctx); }

int vslam_calc_pose_update(vslam_ctx* ctx, double sigma, int mark, int apply, double* out) {
  if (!ctx) return VSLAM_E_INVALID;
  int rc = vs_launch_pose(ctx, 0, sigma, mark, apply); if (rc) return rc;
  if (out) {
    VS_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<StreamState> h(ctx->S);
    VS_CUDA(cudaMemcpy(h.data(), ctx->ss, sizeof(StreamState) * ctx->S, cudaMemcpyDeviceToHost));
    for (int s = 0; s < ctx->S; s++) { const int u = h[s].n_updates - 1; for (int k = 0; k < 6; k++) out[6 * s + k] = u >= 0 ? h[s].updates[6 * u + k] : 0.0; }
  }
  return VSLAM_OK;
}

int vslam_track_map(vslam_ctx* ctx) { if (!ctx) return VSLAM_E_INVALID; return vs_launch_track_map(ctx, 0); }

int vslam_track_frame(vslam_ctx* ctx, const uint8_t* gray, int stride, size_t frame_stride) {
  if (!ctx) return VSLAM_E_INVALID;
  int rc = vs_begin_frame(ctx);
  if (!rc) rc = stage_host_frames(ctx, 0, ctx->S, gray, stride, frame_stride);
  return end_frame(ctx, rc ? rc : vs_launch_frame(ctx));
}
// Pipelined host-input path: the copy of step k (copy stream, level-0 buffer k&1) overlaps the kernels of step k-1.
int vslam_track_frame_async(vslam_ctx* ctx, const uint8_t* gray, int stride, size_t frame_stride, double* poses_out) {
  int rc = check_range(ctx, 0, ctx ? ctx->S : 0, gray, stride); if (rc) return rc;
  const LevelDesc& L = ctx->lev[0];
  if (!ctx->pipe_ready) {
    VS_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    VS_CUDA(cudaHostAlloc((void**)&ctx->status_pin, sizeof(int) * 8, cudaHostAllocDefault));
    memset(ctx->status_pin, 0, sizeof(int) * 8);
    for (int k = 0; k < 2; k++) {
      VS_CUDA(cudaEventCreateWithFlags(&ctx->ev_copied[k], cudaEventDisableTiming)); VS_CUDA(cudaEventCreateWithFlags(&ctx->ev_computed[k], cudaEventDisableTiming));
      VS_CUDA(cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming));
      VS_CUDA(cudaEventRecord(ctx->ev_computed[k], ctx->stream)); VS_CUDA(cudaEventRecord(ctx->ev_done[k], ctx->stream));
    }
    ctx->pipe_ready = true;
  }
  const int slot = (int)(ctx->step & 1);
  VS_CUDA(cudaEventSynchronize(ctx->ev_done[slot]));                       // at most two steps in flight
  if ((rc = vs_begin_frame(ctx))) return end_frame(ctx, rc);
  // level-0 buffer of this step: with look-ahead the frame set's own one, else the two buffers in turn (b = which of the two: the events of the copy belong to the buffer)
  const int b = ctx->la_frame ? ctx->cur_set : slot;
  if (b == 1 && !ctx->l0_alt) { VS_CUDA(dalloc(&ctx->l0_alt, (size_t)ctx->S * L.h * L.pitch)); ctx->sets[1].img[0] = ctx->l0_alt; if (ctx->cur_set == 1) ctx->lev[0].img = ctx->l0_alt; }
  uint8_t* buf = b ? ctx->l0_alt : ctx->sets[0].img[0];
  VS_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_computed[b], 0));   // the kernels that last read this buffer are done
  if ((rc = upload_frames_to(ctx, buf, ctx->copy_stream, 0, ctx->S, gray, stride, frame_stride))) return end_frame(ctx, rc);
  VS_CUDA(cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
  VS_CUDA(cudaStreamWaitEvent(vs_in_stream(ctx), ctx->ev_copied[b], 0));
  bool same = true;
  for (int s = 0; s < ctx->S; s++) same &= ctx->l0_ptr_host[s] == buf + (size_t)s * L.h * L.pitch && ctx->l0_stride_host[s] == L.pitch;
  if (!same && (rc = adopt_l0(ctx, 0, ctx->S, buf, L.pitch, (size_t)L.h * L.pitch, false))) return end_frame(ctx, rc);
  if ((rc = end_frame(ctx, vs_launch_frame(ctx)))) return rc;
  VS_CUDA(cudaEventRecord(ctx->ev_computed[b], ctx->stream));
  if (poses_out)
    VS_CUDA(cudaMemcpy2DAsync(poses_out, sizeof(double) * 12, (char*)ctx->ss + offsetof(StreamState, pose), sizeof(StreamState), sizeof(double) * 12, ctx->S,
                              cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(cudaMemcpyAsync(ctx->status_pin + 4 * slot, ctx->status, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(cudaEventRecord(ctx->ev_done[slot], ctx->stream));
  const int id = (int)ctx->step;
  ctx->step = (ctx->step + 1) & 0x3fffffff;   // ids wrap at 2^30 (even, so the slot parity is preserved)
  return id;
}

int vslam_wait_step(vslam_ctx* ctx, int step) {
  if (!ctx || !ctx->pipe_ready || step < 0 || step > 0x3fffffff) { if (ctx) ctx->err = "unknown step id"; return VSLAM_E_INVALID; }
  // ids count modulo 2^30: `age` = how many steps were issued since `step` (1 = the latest); an id that was never handed out shows up as
  // a huge age and is treated like any long-finished step
  const long long age = (ctx->step - (long long)step) & 0x3fffffff;
  if (age == 0) { ctx->err = "unknown step id"; return VSLAM_E_INVALID; }     // = the id the NEXT step will get
  if (age > 2) return VSLAM_OK;   // older steps completed before their slot was reused
  VS_CUDA(cudaEventSynchronize(ctx->ev_done[step & 1]));
  if (ctx->status_pin[4 * (step & 1)]) {       // corner-capacity overflow seen by that step (the flag is sticky until vslam_sync)
    ctx->err = "corner list capacity exceeded (raise vslam_config.max_corner_frac)";
    return VSLAM_E_CAPACITY;
  }
  return VSLAM_OK;
}

int vslam_track_frame_dev(vslam_ctx* ctx, const uint8_t* gray, int stride, size_t frame_stride) {
  if (!ctx) return VSLAM_E_INVALID;
  int rc = vs_begin_frame(ctx);
  if (!rc) rc = adopt_device_frames(ctx, 0, ctx->S, gray, stride, frame_stride);
  return end_frame(ctx, rc ? rc : vs_launch_frame(ctx));
}

}  // extern "C"
