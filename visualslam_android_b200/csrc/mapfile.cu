// Map files: the on-disk form of what the tracking path needs from the map -- camera, source keyframes (level-0
// images; the library rebuilds their pyramids), map points (the vslam_set_map arrays) and the relocaliser registration -- so that a
// long-running multi-stream service can restart, or a second process / GPU rank can serve streams of the same map.
//
// The reference has no map format: its only persistence is the debug dump of MapMaker::GUICommandHandler("SaveMap")
// (jni/MapMaker.cc:1254-1297: map.dump = world position + source level per point, keyframes/<i>.info = pose; the image write is
// commented out).  vslam_export_map_text writes that layout for tools that read it; vslam_save_map_file / vslam_load_map_file are
// this library's own lossless binary format:
//
//   offset 0   MapFileHeader (little endian, 168 bytes)
//              keyframes     n_keyframes x { int32 id, int32 registered (index in the relocaliser list or -1), width*height bytes, padding to 8 }
//              points        world[n][3] right[n][3] down[n][3] (f64)  ircenter[n][2] srclevel[n] srckf[n] (i32), padding to 8
//              relocaliser   ids[n_reloc] (i32, padded to 8)  poses[n_reloc][12] (f64)
//   trailer    uint64 FNV-1a of every preceding byte
//
// Host code only (file I/O + the existing upload entry points); nothing here is on the per-frame path.
#include "vslam_internal.cuh"
#include <exception>
#include <cstdio>
#include <cerrno>
#include <sys/stat.h>

namespace {

struct MapFileHeader {
  char magic[8];          // "VSLMAP\0\1"
  uint32_t version;       // 1
  uint32_t header_bytes;  // sizeof(MapFileHeader)
  int32_t width, height;
  int32_t n_points, n_keyframes, n_reloc, reserved;
  double cam13[13];       // vslam_set_camera layout
  double reserved_f[3];
};
static_assert(sizeof(MapFileHeader) == 168, "map file header layout");
const char kMagic[8] = {'V', 'S', 'L', 'M', 'A', 'P', 0, 1};

struct Fnv {
  uint64_t h = 1469598103934665603ull;
  void add(const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } }
};

// Writer / reader that keep the running checksum and the 8-byte section alignment.
struct Out {
  FILE* f; Fnv sum; size_t pos = 0; bool ok = true;
  void put(const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) ok = false; sum.add(p, n); pos += n; }
  void pad8() { static const char z[8] = {0}; if (pos % 8) put(z, 8 - pos % 8); }
};
struct In {
  FILE* f; Fnv sum; size_t pos = 0; bool ok = true;
  void get(void* p, size_t n) { if (n && fread(p, 1, n, f) != n) { ok = false; memset(p, 0, n); } sum.add(p, n); pos += n; }
  void pad8() { char z[8]; if (pos % 8) get(z, 8 - pos % 8); }
};

void cam_to13(const CamDev& d, double* c) {
  c[0] = d.fx; c[1] = d.fy; c[2] = d.cx; c[3] = d.cy; c[4] = d.W; c[5] = d.Winv; c[6] = d.twoTan; c[7] = d.oneOver2Tan; c[8] = d.distEnabled;
  c[9] = d.largestRadius; c[10] = d.maxR; c[11] = d.width; c[12] = d.height;
}

int read_header(FILE* f, In& in, MapFileHeader& h, std::string& err) {
  in.f = f; in.get(&h, sizeof(h));
  if (!in.ok || memcmp(h.magic, kMagic, 8) != 0) { err = "not a vslam map file (bad magic)"; return VSLAM_E_INVALID; }
  if (h.version != 1 || h.header_bytes != sizeof(MapFileHeader)) { err = "unsupported map file version"; return VSLAM_E_INVALID; }
  if (h.width <= 0 || h.height <= 0 || h.n_points < 0 || h.n_keyframes < 0 || h.n_reloc < 0) { err = "corrupt map file header"; return VSLAM_E_INVALID; }
  return VSLAM_OK;
}

// Eigen's default stream format for a column vector (what `ofs << v3WorldPos` prints in the reference's dump): one coefficient per
// line, each right-aligned to the widest one, at the stream's precision (6 significant digits by default).  Eigen itself is not under
// /root/reference, so this layout is restated from Eigen's documented IOFormat defaults, unpinned.
void print_column(FILE* f, const double* v, int n) {
  char buf[8][64]; int width = 0;
  for (int i = 0; i < n; i++) { const int len = snprintf(buf[i], sizeof(buf[i]), "%g", v[i]); if (len > width) width = len; }
  for (int i = 0; i < n; i++) fprintf(f, "%*s%s", width, buf[i], i + 1 < n ? "\n" : "");
}

}  // namespace

extern "C" {

int vslam_map_file_info(const char* path, vslam_map_file_info_t* out) {
  if (!path || !out) return VSLAM_E_INVALID;
  FILE* f = fopen(path, "rb");
  if (!f) return VSLAM_E_IO;
  In in; MapFileHeader h; std::string err;
  const int rc = read_header(f, in, h, err);
  fclose(f);
  if (rc) return rc;
  out->width = h.width; out->height = h.height; out->n_points = h.n_points; out->n_keyframes = h.n_keyframes; out->n_reloc_keyframes = h.n_reloc;
  memcpy(out->cam13, h.cam13, sizeof(h.cam13));
  return VSLAM_OK;
}

int vslam_save_map_file(vslam_ctx* ctx, const char* path) {
  if (!ctx || !path) return VSLAM_E_INVALID;
  int nkf = 0;
  for (int k = 0; k < ctx->n_src; k++) if (ctx->src_have[k]) nkf++;
  const int n = ctx->map.n, W = ctx->src.w[0], H = ctx->src.h[0];
  {   // every point's source keyframe must be part of the file
    std::vector<int> kf(n);
    int rc = vslam_sync(ctx); if (rc) return rc;
    if (n) VS_CUDA(cudaMemcpy(kf.data(), ctx->map.srckf, sizeof(int) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) if (!ctx->src_have[kf[i]]) { ctx->err = "map point refers to a source keyframe that was never uploaded"; return VSLAM_E_INVALID; }
  }
  FILE* f = fopen(path, "wb");
  if (!f) { ctx->err = std::string("cannot open ") + path + " for writing: " + strerror(errno); return VSLAM_E_IO; }
  Out out; out.f = f;
  MapFileHeader h; memset(&h, 0, sizeof(h));
  memcpy(h.magic, kMagic, 8); h.version = 1; h.header_bytes = sizeof(h); h.width = W; h.height = H; h.n_points = n; h.n_keyframes = nkf; h.n_reloc = (int)ctx->reloc_ids.size();
  cam_to13(ctx->cam, h.cam13);
  out.put(&h, sizeof(h));
  std::vector<uint8_t> img((size_t)W * H);
  cudaError_t ce = cudaSuccess;
  for (int k = 0; k < ctx->n_src && ce == cudaSuccess; k++) {
    if (!ctx->src_have[k]) continue;
    int32_t rec[2] = {k, -1};
    for (size_t r = 0; r < ctx->reloc_ids.size(); r++) if (ctx->reloc_ids[r] == k) { rec[1] = (int)r; break; }
    ce = cudaMemcpy2D(img.data(), W, ctx->src.img[0] + (size_t)k * H * ctx->src.pitch[0], ctx->src.pitch[0], W, H, cudaMemcpyDeviceToHost);
    out.put(rec, sizeof(rec)); out.put(img.data(), img.size()); out.pad8();
  }
  if (ce == cudaSuccess && n) {
    std::vector<double> d(3 * (size_t)n); std::vector<int> iv(2 * (size_t)n);
    const double* dsrc[3] = {ctx->map.world, ctx->map.right, ctx->map.down};
    for (int a = 0; a < 3 && ce == cudaSuccess; a++) { ce = cudaMemcpy(d.data(), dsrc[a], sizeof(double) * 3 * n, cudaMemcpyDeviceToHost); out.put(d.data(), sizeof(double) * 3 * n); }
    if (ce == cudaSuccess) { ce = cudaMemcpy(iv.data(), ctx->map.ircenter, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost); out.put(iv.data(), sizeof(int) * 2 * n); }
    if (ce == cudaSuccess) { ce = cudaMemcpy(iv.data(), ctx->map.srclevel, sizeof(int) * n, cudaMemcpyDeviceToHost); out.put(iv.data(), sizeof(int) * n); }
    if (ce == cudaSuccess) { ce = cudaMemcpy(iv.data(), ctx->map.srckf, sizeof(int) * n, cudaMemcpyDeviceToHost); out.put(iv.data(), sizeof(int) * n); }
    out.pad8();
  }
  out.put(ctx->reloc_ids.data(), sizeof(int) * ctx->reloc_ids.size()); out.pad8();
  out.put(ctx->reloc_poses_host.data(), sizeof(double) * ctx->reloc_poses_host.size());
  const uint64_t sum = out.sum.h;
  if (fwrite(&sum, 1, 8, f) != 8) out.ok = false;
  if (fclose(f) != 0) out.ok = false;
  if (ce != cudaSuccess) { ctx->err = std::string("vslam_save_map_file: ") + cudaGetErrorString(ce); remove(path); return VSLAM_E_CUDA; }
  if (!out.ok) { ctx->err = std::string("short write to ") + path; remove(path); return VSLAM_E_IO; }
  return VSLAM_OK;
}

int vslam_load_map_file(vslam_ctx* ctx, const char* path, int flags) {
  if (!ctx || !path) return VSLAM_E_INVALID;
  FILE* f = fopen(path, "rb");
  if (!f) { ctx->err = std::string("cannot open ") + path + ": " + strerror(errno); return VSLAM_E_IO; }
  In in; MapFileHeader h;
  int rc = read_header(f, in, h, ctx->err);
  if (rc) { fclose(f); return rc; }
  const int W = ctx->src.w[0], H = ctx->src.h[0];
  if (h.width != W || h.height != H) { fclose(f); ctx->err = "map file was written for another image size"; return VSLAM_E_INVALID; }
  if (h.n_points < 0 || h.n_keyframes < 0 || h.n_reloc < 0) { fclose(f); ctx->err = "map file header holds negative counts"; return VSLAM_E_IO; }
  if (h.n_points > ctx->N || h.n_keyframes > ctx->n_src) { fclose(f); ctx->err = "map file exceeds this context's max_points / max_source_keyframes"; return VSLAM_E_CAPACITY; }
  // relocaliser keyframes are source keyframes: a count beyond the slots is a corrupt header (sizes below come from the header, before the checksum)
  if (h.n_reloc > ctx->n_src) { fclose(f); ctx->err = "map file registers more relocaliser keyframes than max_source_keyframes"; return VSLAM_E_CAPACITY; }
  try {
  if ((flags & VSLAM_MAP_LOAD_RELOC) && h.n_reloc > 0 && !ctx->sbi_on) { fclose(f); ctx->err = "VSLAM_MAP_LOAD_RELOC needs vslam_enable_sbi first"; return VSLAM_E_INVALID; }
  // read and verify the whole file before touching the context: a truncated or corrupt file must leave the loaded map as it was
  const size_t n = (size_t)h.n_points, px = (size_t)W * H;
  std::vector<int32_t> kf_rec(2 * (size_t)h.n_keyframes); std::vector<uint8_t> imgs(px * h.n_keyframes);
  for (int k = 0; k < h.n_keyframes; k++) { in.get(&kf_rec[2 * k], 8); in.get(imgs.data() + px * k, px); in.pad8(); }
  std::vector<double> world(3 * n), right(3 * n), down(3 * n); std::vector<int32_t> irc(2 * n), lvl(n), kf(n);
  if (n) {
    in.get(world.data(), 24 * n); in.get(right.data(), 24 * n); in.get(down.data(), 24 * n);
    in.get(irc.data(), 8 * n); in.get(lvl.data(), 4 * n); in.get(kf.data(), 4 * n); in.pad8();
  }
  std::vector<int32_t> rid(h.n_reloc); std::vector<double> rpose(12 * (size_t)h.n_reloc);
  in.get(rid.data(), 4 * (size_t)h.n_reloc); in.pad8(); in.get(rpose.data(), 96 * (size_t)h.n_reloc);
  uint64_t sum = 0; const bool have_sum = fread(&sum, 1, 8, f) == 8;
  const bool at_end = fgetc(f) == EOF;
  fclose(f); f = nullptr;
  if (!in.ok || !have_sum) { ctx->err = "map file is truncated"; return VSLAM_E_IO; }
  if (sum != in.sum.h || !at_end) { ctx->err = "map file checksum mismatch"; return VSLAM_E_IO; }
  for (int k = 0; k < h.n_keyframes; k++) if (kf_rec[2 * k] < 0 || kf_rec[2 * k] >= ctx->n_src) { ctx->err = "map file keyframe id exceeds max_source_keyframes"; return VSLAM_E_CAPACITY; }

  if (flags & VSLAM_MAP_LOAD_CAMERA) { if ((rc = vslam_set_camera(ctx, h.cam13))) return rc; }
  for (int k = 0; k < h.n_keyframes; k++) if ((rc = vslam_upload_source_keyframe(ctx, kf_rec[2 * k], imgs.data() + px * k, W))) return rc;
  static const double dz = 0; static const int32_t iz = 0;   // vslam_set_map wants non-NULL arrays even for an empty map
  if ((rc = vslam_set_map(ctx, (int)n, n ? world.data() : &dz, n ? right.data() : &dz, n ? down.data() : &dz, n ? irc.data() : &iz, n ? lvl.data() : &iz, n ? kf.data() : &iz))) return rc;
  if (flags & VSLAM_MAP_LOAD_RELOC) { if ((rc = vslam_set_reloc_keyframes(ctx, h.n_reloc, rid.data(), rpose.data()))) return rc; }
  return VSLAM_OK;
  } catch (const std::exception& e) {   // nothing may unwind across the C ABI (std::bad_alloc from the staging vectors)
    if (f) fclose(f);
    ctx->err = std::string("vslam_load_map_file: ") + e.what(); return VSLAM_E_IO;
  }
}

int vslam_export_map_text(vslam_ctx* ctx, const char* dir) {
  if (!ctx || !dir) return VSLAM_E_INVALID;
  int rc = vslam_sync(ctx); if (rc) return rc;
  const int n = ctx->map.n;
  std::vector<double> world(3 * (size_t)n); std::vector<int> lvl(n);
  if (n) { VS_CUDA(cudaMemcpy(world.data(), ctx->map.world, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost)); VS_CUDA(cudaMemcpy(lvl.data(), ctx->map.srclevel, sizeof(int) * n, cudaMemcpyDeviceToHost)); }
  const std::string base(dir);
  if (mkdir(base.c_str(), 0777) != 0 && errno != EEXIST) { ctx->err = "cannot create " + base; return VSLAM_E_IO; }
  if (mkdir((base + "/keyframes").c_str(), 0777) != 0 && errno != EEXIST) { ctx->err = "cannot create " + base + "/keyframes"; return VSLAM_E_IO; }
  FILE* f = fopen((base + "/map.dump").c_str(), "w");
  if (!f) { ctx->err = "cannot write " + base + "/map.dump"; return VSLAM_E_IO; }
  for (int i = 0; i < n; i++) { print_column(f, &world[3 * (size_t)i], 3); fprintf(f, "  %d\n", lvl[i]); }   // jni/MapMaker.cc:1260-1264
  fclose(f);
  // keyframes/<i>.info = `ofs << se3CfromW << endl` (jni/MapMaker.cc:1267-1282 with mySE3's operator<<, jni/RT.h:304-313): rows "r0 r1 r2 t\n".
  // The map's keyframes are the registered relocaliser keyframes, in registration order (the only keyframes whose poses the library holds).
  for (size_t k = 0; k < ctx->reloc_ids.size(); k++) {
    char name[64]; snprintf(name, sizeof(name), "/keyframes/%zu.info", k);
    f = fopen((base + name).c_str(), "w");
    if (!f) { ctx->err = "cannot write " + base + name; return VSLAM_E_IO; }
    const double* p = &ctx->reloc_poses_host[12 * k];
    for (int r = 0; r < 3; r++) fprintf(f, "%g %g %g %g\n", p[4 * r], p[4 * r + 1], p[4 * r + 2], p[4 * r + 3]);
    fprintf(f, "\n");
    fclose(f);
  }
  return VSLAM_OK;
}

}  // extern "C"
