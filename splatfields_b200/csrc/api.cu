// api.cu — the C ABI declared in include/splat_b200.h: host-side orchestration of the kernels.
// Mirrors the control flow of the external rasterizer's forward/backward entry points
// (SURVEY.md §3.2-3.3): preprocess -> [R to host] -> binning -> render ; render-bwd -> geometry-bwd.
#include "../../include/splat_b200.h"
#include "common.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace {

thread_local std::string g_err;
int g_launches = 0;   // process-wide: autograd runs backward on its own thread
thread_local uint32_t* g_pinned = nullptr;  // pinned (portable, mapped) word the preprocess kernel stores num_rendered into
constexpr int MAX_DEVICES = 64;
thread_local cudaEvent_t g_evt[MAX_DEVICES] = {};   // one per device: an event can only be recorded on its own device's streams

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  char buf[512];
  if (e != cudaSuccess) snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  else snprintf(buf, sizeof(buf), "%s", what);
  g_err = buf;
  return code;
}

#define CK(call)                                                        \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) return fail(SFB_ERR_CUDA, #call, e__);      \
  } while (0)

#define CK_LAUNCH(name, dbg, s)                                                     \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ == cudaSuccess && (dbg)) e__ = cudaStreamSynchronize(s);                \
    if (e__ != cudaSuccess) return fail(SFB_ERR_CUDA, "kernel " name, e__);         \
  } while (0)

// ---- optional per-kernel device timing (CUDA events on the launching stream) ----
// Every kernel launch is bracketed by an event pair when profiling is on; records of the last forward
// (which = 0) and backward (which = 1) are kept until the next profiled call.
constexpr int MAX_REC = 48;
const char* const kDepthSortNames[3] = {"depth_sort.hist", "depth_sort.scan", "depth_sort.scatter"};
const char* const kTileSortNames[3] = {"tile_sort.hist", "tile_sort.scan", "tile_sort.scatter"};
struct ProfRec { const char* name; cudaEvent_t a, b; };
// process-wide (not thread_local): PyTorch runs the backward on an autograd worker thread
bool g_prof = false;
int g_which = 0;
ProfRec g_rec[2][MAX_REC];
int g_nrec[2] = {0, 0};
bool g_rec_made = false;

// ---- red-zone verification (debug=True only) ----
int redzones_fill(const sfb::RedzoneList& rz, cudaStream_t s) {
  for (int i = 0; i < rz.n; i++)
    if (cudaMemsetAsync(rz.ptr[i], 0xA5, sfb::REDZONE_BYTES, s) != cudaSuccess) return -1;
  return 0;
}
// returns the index of the first clobbered red zone, -1 if all intact, -2 on a CUDA error
int redzones_check(const sfb::RedzoneList& rz, cudaStream_t s) {
  unsigned char host[sfb::REDZONE_BYTES];
  for (int i = 0; i < rz.n; i++) {
    if (cudaMemcpyAsync(host, rz.ptr[i], sfb::REDZONE_BYTES, cudaMemcpyDeviceToHost, s) != cudaSuccess) return -2;
    if (cudaStreamSynchronize(s) != cudaSuccess) return -2;
    for (size_t k = 0; k < sfb::REDZONE_BYTES; k++)
      if (host[k] != 0xA5) return i;
  }
  return -1;
}

int tile_sort_final(int T) {
  int bits = sfb::tile_bits(T);
  int npass = (bits + 7) / 8;
  return npass & 1;
}

}  // namespace

namespace sfb {
void set_error(const char* msg) { g_err = msg ? msg : ""; }
void prof_begin(const char* name, cudaStream_t s) {
  if (!g_prof) return;
  if (!g_rec_made) {
    for (int w = 0; w < 2; w++)
      for (int i = 0; i < MAX_REC; i++) { cudaEventCreate(&g_rec[w][i].a); cudaEventCreate(&g_rec[w][i].b); }
    g_rec_made = true;
  }
  int& n = g_nrec[g_which];
  if (n >= MAX_REC) return;
  g_rec[g_which][n].name = name;
  cudaEventRecord(g_rec[g_which][n].a, s);
}
void prof_end(cudaStream_t s) {
  if (!g_prof) return;
  int& n = g_nrec[g_which];
  if (n >= MAX_REC) return;
  cudaEventRecord(g_rec[g_which][n].b, s);
  n++;
}
}  // namespace sfb

extern "C" {

int sfb_abi_version(void) { return 6; }
const char* sfb_last_error(void) { return g_err.c_str(); }
int sfb_last_launch_count(void) { return g_launches; }

int sfb_rasterize_forward(int P, int sh_degree, int M, int W, int H, const float* bg, const float* means3D,
                          const float* shs, const float* colors_precomp, const float* opacities,
                          const float* scales, float scale_modifier, const float* rotations,
                          const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                          const float* campos, float tan_fovx, float tan_fovy, int prefiltered,
                          float* out_color, float* out_depth, float* out_alpha, int* radii, sfb_alloc_fn geom_alloc,
                          void* geom_user, sfb_alloc_fn binning_alloc, void* binning_user, sfb_alloc_fn img_alloc,
                          void* img_user, int* num_rendered, int debug, void* stream) {
  using namespace sfb;
  g_err.clear();
  g_launches = 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (num_rendered) *num_rendered = 0;
  if (P < 0 || W <= 0 || H <= 0) return fail(SFB_ERR_ARG, "bad sizes");
  if (!out_color || !out_depth || !bg || !geom_alloc || !binning_alloc || !img_alloc)
    return fail(SFB_ERR_ARG, "null output / allocator");
  const size_t HW = (size_t)H * W;
  if (P == 0) {  // like the reference: nothing is launched, outputs are zero-filled
    CK(cudaMemsetAsync(out_color, 0, 3 * HW * sizeof(float), s));
    CK(cudaMemsetAsync(out_depth, 0, HW * sizeof(float), s));
    if (out_alpha) CK(cudaMemsetAsync(out_alpha, 0, HW * sizeof(float), s));
    return SFB_OK;
  }
  if (!means3D || !opacities || !viewmatrix || !projmatrix || !campos || !radii)
    return fail(SFB_ERR_ARG, "null input");
  if ((shs == nullptr) == (colors_precomp == nullptr))
    return fail(SFB_ERR_ARG, "Please provide exactly one of either SHs or precomputed colors!");
  const bool has_sr = scales != nullptr && rotations != nullptr;
  if (has_sr == (cov3D_precomp != nullptr) || (scales == nullptr) != (rotations == nullptr))
    return fail(SFB_ERR_ARG, "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
  if (shs && (sh_degree < 0 || sh_degree > 3 || M < (sh_degree + 1) * (sh_degree + 1)))
    return fail(SFB_ERR_ARG, "sh_degree / M mismatch (degree 0..3, M >= (degree+1)^2)");
  const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y, T = gx * gy;
  if (gx > 65535 || gy > 65535) return fail(SFB_ERR_ARG, "image too large for packed tile rectangles");

  if (!g_pinned) CK(cudaHostAlloc((void**)&g_pinned, 64, cudaHostAllocMapped | cudaHostAllocPortable));
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAX_DEVICES) return fail(SFB_ERR_ARG, "device ordinal out of range");
  if (!g_evt[dev]) CK(cudaEventCreateWithFlags(&g_evt[dev], cudaEventDisableTiming));
  cudaEvent_t evt = g_evt[dev];

  RedzoneList rz;
  char* gchunk = (char*)geom_alloc(geom_user, GeomState::required((size_t)P));
  if (!gchunk) return fail(SFB_ERR_ALLOC, "geometry buffer allocation failed");
  if (debug) redzone_collector() = &rz;
  GeomState g = GeomState::from_chunk(gchunk, (size_t)P);
  redzone_collector() = nullptr;
  char* ichunk = (char*)img_alloc(img_user, ImgState::required(HW, (size_t)T));
  if (!ichunk) return fail(SFB_ERR_ALLOC, "image buffer allocation failed");
  if (debug) redzone_collector() = &rz;
  ImgState img = ImgState::from_chunk(ichunk, HW, (size_t)T);
  redzone_collector() = nullptr;
  if (debug && redzones_fill(rz, s)) return fail(SFB_ERR_CUDA, "red-zone fill");

  FwdParams fp;
  fp.P = P; fp.D = sh_degree; fp.M = M; fp.W = W; fp.H = H;
  fp.bg = bg; fp.means3D = means3D; fp.shs = shs; fp.colors_precomp = colors_precomp; fp.opacities = opacities;
  fp.scales = scales; fp.rotations = rotations; fp.cov3D_precomp = cov3D_precomp;
  fp.viewmatrix = viewmatrix; fp.projmatrix = projmatrix; fp.campos = campos;
  fp.scale_modifier = scale_modifier; fp.tan_fovx = tan_fovx; fp.tan_fovy = tan_fovy; fp.prefiltered = prefiltered;
  fp.wide256 = shs && ((M * 12) % 32 == 0) && ((reinterpret_cast<size_t>(shs) & 31) == 0);
  // the depth sort's scratch is cleared by the preprocess blocks on their way (one memset node less)
  const size_t dzero = radix_sort_zero_words(P, 32);
  fp.zero_ptr = dzero ? g.sort_hist : nullptr;
  fp.zero_words = (uint32_t)dzero;
  fp.nr_host = g_pinned;

  // K1 (+ num_rendered reduction) ; the 4-byte read-back is issued right behind it so that the host
  // wait overlaps the depth sort instead of draining the whole pipeline.
  g_which = 0; g_nrec[0] = 0;
  prof_begin("preprocess", s);
  CK(cudaMemsetAsync(g.counters, 0, 8 * sizeof(uint32_t), s));
  launch_preprocess(fp, g, radii, s);
  prof_end(s);
  g_launches++;
  CK_LAUNCH("preprocess", debug, s);
  CK(cudaEventRecord(evt, s));

  // stage 1: stable sort of the Gaussians by depth bits (culled ones carry 0xFFFFFFFF and sink)
  const uint32_t* top_total0 = nullptr;
  int dfinal = radix_sort_pairs(g.depth_key, g.depth_idx, g.sort_hist, P, 32, s, &g_launches, kDepthSortNames,
                                g.counters + 2, 0, dzero != 0, nullptr, 0, nullptr, 0, &top_total0);
  CK_LAUNCH("depth sort", debug, s);
  const SortedIdx sorted_idx{g.depth_idx[dfinal], g.depth_idx[dfinal ^ 1], top_total0, (uint32_t)P};
  launch_instance_block_sums(P, sorted_idx, g.tiles_touched, g.block_sums, s);
  g_launches += 2;
  CK_LAUNCH("instance scan", debug, s);

  CK(cudaEventSynchronize(evt));
  const uint32_t R = *reinterpret_cast<volatile uint32_t*>(g_pinned);
  if (R >= (1u << 30)) return fail(SFB_ERR_ARG, "num_rendered >= 2^30 is not supported");
  if (num_rendered) *num_rendered = (int)R;

  const InstPacking pk = inst_packing((size_t)P, (size_t)T);
  if (!pk.ok) return fail(SFB_ERR_ARG, "ceil(log2 P) + ceil(log2 tiles) > 40 bits is not supported");
  const bool split = pk.high_bits > 0;
  char* bchunk = (char*)binning_alloc(binning_user, BinState::required((size_t)R, (size_t)T, split));
  if (!bchunk) return fail(SFB_ERR_ALLOC, "binning buffer allocation failed");
  RedzoneList rzb;
  if (debug) redzone_collector() = &rzb;
  BinState b = BinState::from_chunk(bchunk, (size_t)R, (size_t)T, split);
  redzone_collector() = nullptr;
  if (debug && redzones_fill(rzb, s)) return fail(SFB_ERR_CUDA, "red-zone fill");

  int tfinal = 0;
  if (R > 0) {
    // stage 2: emit (tile, gaussian) instances in depth order (the blocks also clear the tile sort's scratch and set
    // every tile range to "empty"); stage 3: stable sort by tile id
    const size_t tzero = radix_sort_zero_words((int)R, tile_bits(T));
    prof_begin("duplicate", s);
    launch_duplicate(P, gx, sorted_idx, g.tiles_touched, g.rect, g.block_sums, b.tile_key[0], b.inst_hi[0],
                     pk.low_bits, tzero ? b.sort_hist : nullptr, tzero, b.ranges, T, img.tile_bcount, s);
    prof_end(s);
    g_launches++;
    CK_LAUNCH("duplicate", debug, s);
    // bare 32-bit words (+ one side byte per instance when the index does not fit), digits start above the index bits.
    // The last pass also writes the per-tile [start, end) ranges (K5): it sees every sorted key on its way out.
    tfinal = radix_sort_pairs(b.tile_key, nullptr, b.sort_hist, (int)R, tile_bits(T), s, &g_launches, kTileSortNames,
                              nullptr, pk.low_bits, tzero != 0, b.ranges, pk.low_bits, split ? b.inst_hi : nullptr,
                              split ? pk.low_bits : 0);
    CK_LAUNCH("tile sort", debug, s);
  } else {
    CK(cudaMemsetAsync(b.ranges, 0, sizeof(uint2) * (size_t)T, s));     // nothing to render: every tile is (0, 0)
    CK(cudaMemsetAsync(img.tile_bcount, 0, sizeof(uint32_t) * TILE_BUCKETS, s));
  }

  prof_begin("render_forward", s);
  if (launch_render_forward(W, H, b.ranges, b.point_list(tfinal), pk.idx_mask, g.rec, bg, out_color, out_depth,
                            out_alpha, img.final_T, img.n_contrib, b.hit, img.tile_bcount, img.tile_btile, g.grad,
                            (size_t)P, s) != 0)
    return fail(SFB_ERR_CUDA, g_err.c_str());
  prof_end(s);
  g_launches++;
  CK_LAUNCH("render forward", debug, s);
  if (debug) {
    int bad = redzones_check(rz, s);
    if (bad == -1) { bad = redzones_check(rzb, s); if (bad >= 0) bad += 100; }
    if (bad != -1) {
      char msg[128];
      snprintf(msg, sizeof(msg), "forward wrote past a scratch array (red zone %d clobbered)", bad);
      return fail(SFB_ERR_CUDA, msg);
    }
  }
  return SFB_OK;
}

// device-side view of the symmetric buffers for one step (records of parity `par`, this step's colour table)
static sfb::XchgDev make_xchg_dev(const sfb_xchg* x, const sfb::XchgLayout& xl, int par, unsigned epoch) {
  using namespace sfb;
  XchgDev d;
  d.rank = x->rank; d.world = x->world; d.P = x->P; d.ngeo = x->ngeo; d.nch = xl.nch;
  d.flags = reinterpret_cast<uint32_t*>(x->local);
  d.cflags = reinterpret_cast<uint32_t*>((char*)x->local + xl.cflag_off);
  d.geo = reinterpret_cast<float*>((char*)x->local + xl.geo_off[par]);
  d.geo_mc = x->mc ? reinterpret_cast<float*>((char*)x->mc + xl.geo_off[par]) : nullptr;
  for (int r = 0; r < XCHG_MAX_RANKS; r++) { d.peer_flags[r] = nullptr; d.peer_cflags[r] = nullptr; d.peer_geo[r] = nullptr; }
  for (int r = 0; r < x->world; r++) {
    d.peer_flags[r] = reinterpret_cast<uint32_t*>(x->peers[r]);
    d.peer_cflags[r] = reinterpret_cast<uint32_t*>((char*)x->peers[r] + xl.cflag_off);
    d.peer_geo[r] = reinterpret_cast<float*>((char*)x->peers[r] + xl.geo_off[par]);
  }
  d.gc = reinterpret_cast<const float*>((char*)x->local + xl.gc_off[epoch & 1u]);
  d.gc_slot_floats = xl.gc_slot_floats;
  return d;
}

int sfb_rasterize_backward(int P, int sh_degree, int M, int num_rendered, int W, int H, const float* bg,
                           const float* means3D, const float* shs, const float* colors_precomp,
                           const float* scales, float scale_modifier, const float* rotations,
                           const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                           const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                           void* geom_buffer, void* binning_buffer, void* img_buffer,
                           const float* dL_dout_color, const float* dL_dout_alpha, const float* dL_dout_depth,
                           float* dL_dmeans2D,
                           float* dL_dcolors, float* dL_dopacity,
                           float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscales,
                           float* dL_drotations, int debug, int flags, const sfb_xchg* xchg, unsigned xchg_epoch,
                           void* stream) {
  using namespace sfb;
  g_err.clear();
  g_launches = 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (P == 0) return SFB_OK;
  if (P < 0 || W <= 0 || H <= 0 || num_rendered < 0) return fail(SFB_ERR_ARG, "bad sizes");
  if (!geom_buffer || !binning_buffer || !img_buffer) return fail(SFB_ERR_ARG, "null scratch buffer");
  if (!dL_dout_color || !dL_dmeans2D) return fail(SFB_ERR_ARG, "null gradient buffer");
  const bool sh_factored = (flags & SFB_BWD_SH_FACTORED) != 0 || (xchg && shs);
  if (xchg) {
    // exchange mode: the parameter gradients leave through the symmetric buffer, the per-parameter outputs are unused
    if (cov3D_precomp || !scales || !rotations) return fail(SFB_ERR_ARG, "sfb_xchg needs scales / rotations (no cov3D_precomp)");
    if (xchg->P != P || xchg->world < 1 || xchg->world > XCHG_MAX_RANKS || xchg->rank < 0 || xchg->rank >= xchg->world ||
        !xchg->local || xchg->ngeo != (shs ? 12 : 16))
      return fail(SFB_ERR_ARG, "sfb_xchg does not match this call (P, world <= 16, rank, ngeo = 12 with shs / 16 without)");
    for (int r = 0; r < xchg->world; r++)
      if (!xchg->peers[r]) return fail(SFB_ERR_ARG, "sfb_xchg: null peer mapping");
  } else {
    if (!dL_dopacity || !dL_dmeans3D) return fail(SFB_ERR_ARG, "null gradient buffer");
    if (colors_precomp && !dL_dcolors) return fail(SFB_ERR_ARG, "dL_dcolors required with colors_precomp");
    if (cov3D_precomp && !dL_dcov3D) return fail(SFB_ERR_ARG, "dL_dcov3D required with cov3D_precomp");
    if (sh_factored && (!shs || !dL_dcolors)) return fail(SFB_ERR_ARG, "SFB_BWD_SH_FACTORED needs shs and dL_dcolors");
    if (shs && !dL_dsh && !sh_factored) return fail(SFB_ERR_ARG, "dL_dsh required with shs");
    if (!cov3D_precomp && (!dL_dscales || !dL_drotations || !scales || !rotations))
      return fail(SFB_ERR_ARG, "dL_dscales / dL_drotations required with scales / rotations");
  }
  const size_t HW = (size_t)H * W;
  const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y, T = gx * gy;
  RedzoneList rz;
  if (debug) redzone_collector() = &rz;
  char* gchunk = (char*)geom_buffer;
  GeomState g = GeomState::from_chunk(gchunk, (size_t)P);
  char* bchunk = (char*)binning_buffer;
  const InstPacking pk = inst_packing((size_t)P, (size_t)T);
  BinState b = BinState::from_chunk(bchunk, (size_t)num_rendered, (size_t)T, pk.high_bits > 0);
  char* ichunk = (char*)img_buffer;
  ImgState img = ImgState::from_chunk(ichunk, HW, (size_t)T);
  redzone_collector() = nullptr;
  const int tfinal = num_rendered > 0 ? tile_sort_final(T) : 0;

  g_which = 1; g_nrec[1] = 0;
  // the forward render cleared the accumulators as a prologue (render_fwd.cu) unless told otherwise
  const bool acc_fresh = (flags & SFB_BWD_ACC_FRESH) && (size_t)P < ((size_t)1 << 30);
  if (!acc_fresh) {
    prof_begin("zero_grad_acc", s);
    CK(cudaMemsetAsync(g.grad, 0, sizeof(GradRec) * (size_t)P, s));
    prof_end(s);
  }
  prof_begin("render_backward", s);
  if (launch_render_backward(W, H, b.ranges, b.point_list(tfinal), pk.idx_mask, g.rec, (size_t)P, bg, img.final_T,
                             img.n_contrib, dL_dout_color, dL_dout_alpha, dL_dout_depth, b.hit, img.tile_bcount,
                             img.tile_btile, g.grad,
                             s) != 0)
    return fail(SFB_ERR_CUDA, g_err.c_str());
  prof_end(s);
  g_launches++;
  CK_LAUNCH("render backward", debug, s);

  BwdParams bp;
  bp.P = P; bp.D = sh_degree; bp.M = M; bp.W = W; bp.H = H;
  bp.means3D = means3D; bp.shs = shs; bp.colors_precomp = colors_precomp; bp.scales = scales;
  bp.rotations = rotations; bp.cov3D_precomp = cov3D_precomp;
  bp.viewmatrix = viewmatrix; bp.projmatrix = projmatrix; bp.campos = campos;
  bp.scale_modifier = scale_modifier; bp.tan_fovx = tan_fovx; bp.tan_fovy = tan_fovy; bp.radii = radii;
  bp.wide256 = shs && ((M * 12) % 32 == 0) && ((reinterpret_cast<size_t>(shs) & 31) == 0) &&
               ((reinterpret_cast<size_t>(dL_dsh) & 31) == 0);
  bp.sh_factored = sh_factored ? 1 : 0;
  bp.dL_dmeans2D = dL_dmeans2D; bp.dL_dcolors = dL_dcolors; bp.dL_dopacity = dL_dopacity;
  bp.dL_dmeans3D = dL_dmeans3D; bp.dL_dcov3D = dL_dcov3D; bp.dL_dsh = dL_dsh; bp.dL_dscales = dL_dscales;
  bp.dL_drot = dL_drotations;
  bp.x_geo = nullptr; bp.x_ngeo = 0; bp.x_mc = 0; bp.x_ndst = 0; bp.x_nranks = 0;
  // Exchange mode with the summed outputs given: ONE fused kernel does the geometry backward AND the exchange
  // (geom_bwd.cu); without them the records stay in the symmetric buffer for sfb_xchg_finish.
  const bool want_fused = xchg && dL_dmeans3D && dL_dopacity && dL_dscales && dL_drotations &&
                          (shs ? (dL_dsh != nullptr && xchg->campos_views != nullptr) : dL_dcolors != nullptr);
  if (xchg) {
    const XchgLayout xl = XchgLayout::make((size_t)P, xchg->world, xchg->ngeo, shs != nullptr);
    const int par = (int)(xchg_epoch & 1u);      // records are double-buffered by step parity
    bp.x_geo = reinterpret_cast<float*>((char*)xchg->local + xl.geo_off[par]);
    bp.x_ngeo = xchg->ngeo;
    bp.x_nranks = xchg->world;
    // slot `rank` of this step's colour-gradient table, on every rank
    const size_t slot = xl.gc_off[xchg_epoch & 1u] + (size_t)xchg->rank * xl.gc_slot_floats * 4;
    for (int r = 0; r < xchg->world; r++) bp.x_gc_peer[r] = reinterpret_cast<float*>((char*)xchg->peers[r] + slot);
    if (xchg->mc) { bp.x_mc = 1; bp.x_ndst = 1; bp.x_gc_dst[0] = reinterpret_cast<float*>((char*)xchg->mc + slot); }
    else { bp.x_ndst = xchg->world; for (int r = 0; r < xchg->world; r++) bp.x_gc_dst[r] = bp.x_gc_peer[r]; }
  }
  bool fused_done = false;
  if (want_fused) {
    const XchgLayout xl = XchgLayout::make((size_t)P, xchg->world, xchg->ngeo, shs != nullptr);
    FusedXchg f;
    f.x = make_xchg_dev(xchg, xl, (int)(xchg_epoch & 1u), xchg_epoch);
    f.epoch = xchg_epoch; f.V = xchg->world; f.M = M;
    f.campos_views = reinterpret_cast<const float*>(xchg->campos_views);
    f.dL_dmeans3D = dL_dmeans3D; f.dL_dopacity = dL_dopacity; f.dL_dscales = dL_dscales; f.dL_drot = dL_drotations;
    f.dL_dcolors = dL_dcolors; f.dL_dsh = dL_dsh;
    BwdParams bf = bp;       // the geometry body writes records + pushed colour gradients, never the summed outputs
    bf.dL_dcolors = nullptr; bf.dL_dopacity = nullptr; bf.dL_dmeans3D = nullptr; bf.dL_dcov3D = nullptr; bf.dL_dsh = nullptr;
    bf.dL_dscales = nullptr; bf.dL_drot = nullptr;
    prof_begin("geom_backward_exchange", s);
    fused_done = launch_geom_exchange_fused(bf, g, f, xchg->max_ctas, s);
    prof_end(s);
    if (!fused_done) return fail(SFB_ERR_ARG, "fused exchange: unsupported SH layout (pass NULL gradient outputs and call sfb_xchg_finish)");
  } else {
    prof_begin("geom_backward", s);
    launch_geom_backward(bp, g, s);
    prof_end(s);
  }
  g_launches++;
  CK_LAUNCH("geometry backward", debug, s);
  if (debug) {   // the forward (run with debug) left the pattern in place; nothing may have touched it since
    int bad = redzones_check(rz, s);
    if (bad != -1) {
      char msg[128];
      snprintf(msg, sizeof(msg), "backward wrote past a scratch array (red zone %d clobbered)", bad);
      return fail(SFB_ERR_CUDA, msg);
    }
  }
  return SFB_OK;
}

size_t sfb_xchg_bytes(int P, int world, int ngeo, int with_colour_tables) {
  if (P < 0 || world < 1 || (ngeo != 12 && ngeo != 16)) return 0;
  return sfb::XchgLayout::make((size_t)P, world, ngeo, with_colour_tables != 0).bytes;
}

int sfb_xchg_finish(const sfb_xchg* x, unsigned epoch, int sh_degree, int M, const float* means3D,
                    const float* campos_views, float* dL_dmeans3D, float* dL_dopacity, float* dL_dscales,
                    float* dL_drotations, float* dL_dcolors, float* dL_dsh, void* stream) {
  using namespace sfb;
  g_err.clear();
  if (!x || x->world < 1 || x->world > XCHG_MAX_RANKS || x->rank < 0 || x->rank >= x->world || !x->local ||
      (x->ngeo != 12 && x->ngeo != 16) || x->P < 0)
    return fail(SFB_ERR_ARG, "sfb_xchg_finish: bad descriptor");
  if (x->P == 0) return SFB_OK;
  const bool has_sh = x->ngeo == 12;
  if (!dL_dmeans3D || !dL_dopacity || !dL_dscales || !dL_drotations || (has_sh ? (!dL_dsh || !means3D || !campos_views) : !dL_dcolors))
    return fail(SFB_ERR_ARG, "sfb_xchg_finish: null pointer");
  if (has_sh && (sh_degree < 0 || sh_degree > 3 || M < (sh_degree + 1) * (sh_degree + 1)))
    return fail(SFB_ERR_ARG, "sfb_xchg_finish: sh_degree / M mismatch");
  cudaStream_t s = (cudaStream_t)stream;
  const XchgLayout xl = XchgLayout::make((size_t)x->P, x->world, x->ngeo, has_sh);
  for (int r = 0; r < x->world; r++)
    if (!x->peers[r]) return fail(SFB_ERR_ARG, "sfb_xchg_finish: null peer mapping");
  XchgDev d = make_xchg_dev(x, xl, (int)(epoch & 1u), epoch);
  prof_begin("xchg_finish", s);
  launch_xchg_finish(d, x->max_ctas, epoch, sh_degree, M, means3D, campos_views, dL_dmeans3D, dL_dopacity, dL_dscales, dL_drotations,
                     dL_dcolors, has_sh ? dL_dsh : nullptr, s);
  prof_end(s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SFB_ERR_CUDA, "xchg finish", e);
  return SFB_OK;
}

int sfb_xchg_status(const sfb_xchg* x, unsigned* status, void* stream) {
  g_err.clear();
  if (!x || !x->local || !status) return fail(SFB_ERR_ARG, "sfb_xchg_status: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(status, reinterpret_cast<const uint32_t*>(x->local) + 34, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return SFB_OK;
}

int sfb_xchg_timeline(const sfb_xchg* x, unsigned long long* ns12, void* stream) {
  g_err.clear();
  if (!x || !x->local || !ns12) return fail(SFB_ERR_ARG, "sfb_xchg_timeline: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  uint32_t w[6];
  CK(cudaMemcpyAsync(ns12, reinterpret_cast<const uint32_t*>(x->local) + 36, 6 * sizeof(unsigned long long),
                     cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(w, reinterpret_cast<const uint32_t*>(x->local) + 56, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  ns12[0] = ~ns12[0];
  for (int k = 0; k < 6; k++) ns12[6 + k] = (unsigned long long)w[k] * 64ull;   // SM cycles summed over the CTAs
  return SFB_OK;
}

void sfb_xchg_tune(int nred_eighths, int depth) { sfb::xchg_tune(nred_eighths, depth); }

int sfb_sh_grad_combine(int P, int V, int sh_degree, int M, const float* means3D, const float* campos,
                        const float* dL_dcolor_views, float* dL_dsh, void* stream) {
  using namespace sfb;
  g_err.clear();
  if (P == 0) return SFB_OK;
  if (P < 0 || V < 1 || V > 64 || sh_degree < 0 || sh_degree > 3 || M < (sh_degree + 1) * (sh_degree + 1))
    return fail(SFB_ERR_ARG, "sfb_sh_grad_combine: bad sizes (1 <= V <= 64, 0 <= sh_degree <= 3, M >= (sh_degree+1)^2)");
  if (!means3D || !campos || !dL_dcolor_views || !dL_dsh) return fail(SFB_ERR_ARG, "sfb_sh_grad_combine: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  launch_sh_grad_combine(P, V, sh_degree, M, means3D, campos, dL_dcolor_views, dL_dsh, true, s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SFB_ERR_CUDA, "sh grad combine", e);
  return SFB_OK;
}

void sfb_profile_enable(int on) {
  g_prof = on != 0;
  if (g_prof) { g_nrec[0] = g_nrec[1] = 0; g_which = 0; }   // a fresh record list per profiling session (MAX_REC records each)
}

int sfb_profile_count(int which) { return (which == 0 || which == 1) ? g_nrec[which] : 0; }

int sfb_profile_read(int which, float* ms, int max_records) {
  g_err.clear();
  if (which < 0 || which > 1 || !ms) return fail(SFB_ERR_ARG, "bad arguments");
  int out = 0;
  for (int i = 0; i < g_nrec[which] && i < max_records; i++, out++) {
    ms[i] = 0.f;
    cudaError_t e = cudaEventSynchronize(g_rec[which][i].b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms[i], g_rec[which][i].a, g_rec[which][i].b);
    if (e != cudaSuccess) return fail(SFB_ERR_CUDA, "profile read", e);
  }
  return out;
}

const char* sfb_profile_name(int which, int i) {
  if ((which == 0 || which == 1) && i >= 0 && i < g_nrec[which]) return g_rec[which][i].name;
  return "";
}

int sfb_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream) {
  (void)projmatrix;
  g_err.clear();
  if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(SFB_ERR_ARG, "bad arguments");
  sfb::launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
  CK_LAUNCH("mark_visible", 0, (cudaStream_t)stream);
  return SFB_OK;
}

int sfb_export_geom(int P, const void* geom_buffer, const float* scales, float scale_modifier,
                    const float* rotations, const float* cov3D_precomp, float* means2D, float* depths,
                    float* cov3D, float* conic_opacity, float* rgb, uint8_t* clamped, uint32_t* tiles_touched,
                    void* stream) {
  using namespace sfb;
  g_err.clear();
  if (P <= 0 || !geom_buffer) return fail(SFB_ERR_ARG, "bad arguments");
  char* gchunk = (char*)geom_buffer;
  GeomState g = GeomState::from_chunk(gchunk, (size_t)P);
  launch_export_geom(P, g, scales, rotations, scale_modifier, cov3D_precomp, means2D, depths, cov3D, conic_opacity,
                     rgb, clamped, tiles_touched, (cudaStream_t)stream);
  CK_LAUNCH("export_geom", 0, (cudaStream_t)stream);
  return SFB_OK;
}

int sfb_export_binning(int P, int num_rendered, int W, int H, const void* geom_buffer,
                       const void* binning_buffer, uint64_t* point_list_keys, uint32_t* point_list,
                       uint32_t* ranges, void* stream) {
  using namespace sfb;
  g_err.clear();
  cudaStream_t s = (cudaStream_t)stream;
  if (P <= 0 || !geom_buffer || !binning_buffer) return fail(SFB_ERR_ARG, "bad arguments");
  const int gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y, T = gx * gy;
  char* gchunk = (char*)geom_buffer;
  GeomState g = GeomState::from_chunk(gchunk, (size_t)P);
  char* bchunk = (char*)binning_buffer;
  const InstPacking pk = inst_packing((size_t)P, (size_t)T);
  BinState b = BinState::from_chunk(bchunk, (size_t)num_rendered, (size_t)T, pk.high_bits > 0);
  const int tfinal = num_rendered > 0 ? tile_sort_final(T) : 0;
  if (point_list_keys || point_list)
    launch_export_keys(num_rendered, T, b.tile_key[tfinal], pk.high_bits > 0, pk.low_bits, b.ranges, g.rec,
                       point_list_keys, point_list, s);
  if (ranges) CK(cudaMemcpyAsync(ranges, b.ranges, sizeof(uint2) * (size_t)T, cudaMemcpyDeviceToDevice, s));
  CK_LAUNCH("export_binning", 0, s);
  return SFB_OK;
}

int sfb_debug_gather_rows(int P, const float* table, int n, const uint32_t* idx, float* out, void* stream) {
  g_err.clear();
  if (P <= 0 || n < 0 || !table || !idx || !out) return fail(SFB_ERR_ARG, "bad arguments");
  if ((reinterpret_cast<size_t>(table) & 15) != 0 || (reinterpret_cast<size_t>(out) & 15) != 0)
    return fail(SFB_ERR_ARG, "table / out must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (sfb::launch_gather_rows_probe(reinterpret_cast<const sfb::SplatRec*>(table), (size_t)P, idx, n, out, s) != 0)
    return fail(SFB_ERR_CUDA, g_err.c_str());
  CK_LAUNCH("gather_rows_probe", 0, s);
  return SFB_OK;
}

int sfb_export_img(int W, int H, const void* img_buffer, float* final_T, uint32_t* n_contrib, void* stream) {
  using namespace sfb;
  g_err.clear();
  cudaStream_t s = (cudaStream_t)stream;
  if (W <= 0 || H <= 0 || !img_buffer) return fail(SFB_ERR_ARG, "bad arguments");
  const size_t HW = (size_t)H * W;
  char* ichunk = (char*)img_buffer;
  ImgState img = ImgState::from_chunk(ichunk, HW, (size_t)(((W + 15) / 16) * ((H + 15) / 16)));
  if (final_T) CK(cudaMemcpyAsync(final_T, img.final_T, sizeof(float) * HW, cudaMemcpyDeviceToDevice, s));
  if (n_contrib) CK(cudaMemcpyAsync(n_contrib, img.n_contrib, sizeof(uint32_t) * HW, cudaMemcpyDeviceToDevice, s));
  return SFB_OK;
}

}  // extern "C"
