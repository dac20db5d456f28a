// loss.cu — SURVEY.md §8f-4: the photometric loss that follows the rasterizer in the training step,
//     loss = (1 - l) * mean|I - G| + l * (1 - SSIM(I, G))  [+ lm * mean|clamp(A, 0, 1) - M|]
// (reference train.py:183-184 and :189-193; utils/loss_utils.py:18-19 l1_loss, :33-76 gaussian / create_window /
// ssim / _ssim) and its gradient w.r.t. the rendered image, produced in the [C][H][W] layout the backward render reads.
//
// The reference evaluates SSIM with five depthwise 11x11 convolutions (zero padding 5) plus ~15 elementwise kernels, and
// autograd replays all of it backwards.  Here:
//   ssim_stats_kernel   one pass over (I, G): separable 11-tap Gaussian filtering of {x, y, x^2, y^2, xy} in shared
//                       memory, SSIM map, |x - y|, block partial sums, and the three per-pixel partial derivatives
//                       D1 = dmap/dmu1, D2 = dmap/dE[x^2], D3 = dmap/dE[xy] (everything the gradient needs);
//   ssim_grad_kernel    the adjoint of the (symmetric, zero-padded) filter applied to D1..D3:
//                       dL/dx = a sign(x - y) - b (w*D1 + 2 x w*D2 + y w*D3);
//   loss_finalize_kernel  fixed-order fp64 reduction of the block partials -> {l1, ssim, mask_l1, loss} (deterministic).
// HBM-bound elementwise/stencil work: 2 + 3 floats per pixel-channel in the first pass, 5 + 1 in the second = 44 B;
// no tensor cores.
//
// Conditioning: sigma^2 = E[x^2] - mu^2 cancels catastrophically in fp32 on flat regions (mu^2 ~ 1, C2 = 9e-4).  Every
// block therefore filters u = x - cx, v = y - cy with cx, cy = the values at the centre of its tile (the variance and
// covariance are shift-invariant; the means get the shift added back).  The window weights sum to 1 up to one fp32
// rounding, which is the size of the terms this drops.
#include "../../include/splat_b200.h"
#include "common.cuh"

#include <cmath>

namespace sfb {

constexpr int LT = 16;            // output tile
constexpr int LHALO = 5;          // window_size // 2
constexpr int LIN = LT + 2 * LHALO;   // 26
constexpr int LSTR = 48;          // row stride of the input tiles: rows r, r+1 land on disjoint bank halves
constexpr int LWIN = 11;

struct GaussWin { float g[LWIN]; };

// utils/loss_utils.py:33-35: gauss = Tensor([exp(-(x - 5)^2 / (2 sigma^2))]) (python doubles rounded to fp32),
// gauss / gauss.sum() in fp32; the 2-D window is the fp32 outer product (:38-41) — applied here as two 1-D passes.
// torch's sum of the 11 values is the correctly rounded exact sum (3.7592328f; a sequential fp32 sum is one ulp lower),
// hence the fp64 accumulation; tests/golden/next_rows.npz holds the reference's window for the bit-exact check.
static GaussWin make_window() {
  GaussWin w;
  double s = 0.0;
  for (int i = 0; i < LWIN; i++) {
    w.g[i] = (float)std::exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5));
    s += (double)w.g[i];
  }
  const float sf = (float)s;
  for (int i = 0; i < LWIN; i++) w.g[i] = w.g[i] / sf;
  return w;
}

__device__ __forceinline__ float block_sum_256(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; w++) t += s_red[w];
  }
  __syncthreads();
  return t;   // valid on thread 0
}

template <bool GRAD>
__global__ void __launch_bounds__(LT * LT)
ssim_stats_kernel(int H, int W, const float* __restrict__ img, const float* __restrict__ gt, GaussWin win,
                  float* __restrict__ dmaps /* [3][C][H][W] */, size_t plane_all /* C*H*W */,
                  float2* __restrict__ partials /* [blocks]: (sum map, sum |x-y|) */) {
  __shared__ float s_x[LIN][LSTR];
  __shared__ float s_y[LIN][LSTR];
  __shared__ float s_h[5][LIN][LT];
  __shared__ float s_red[8];
  const int tid = threadIdx.x;
  const int bx0 = blockIdx.x * LT, by0 = blockIdx.y * LT;
  const size_t plane = (size_t)blockIdx.z * H * W;
  const float* __restrict__ X = img + plane;
  const float* __restrict__ Y = gt + plane;

  // shift = values at the tile's centre pixel (clamped into the image): any constant works, a local one conditions best
  const int cyp = min(by0 + LT / 2, H - 1), cxp = min(bx0 + LT / 2, W - 1);
  const float cx = X[(size_t)cyp * W + cxp], cy = Y[(size_t)cyp * W + cxp];

  for (int i = tid; i < LIN * LIN; i += LT * LT) {
    const int r = i / LIN, c = i - r * LIN;
    const int gy = by0 + r - LHALO, gx = bx0 + c - LHALO;
    float x = 0.f, y = 0.f;                       // zero padding (F.conv2d padding = 5) ...
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) { x = X[(size_t)gy * W + gx]; y = Y[(size_t)gy * W + gx]; }
    s_x[r][c] = x - cx;                           // ... i.e. the padded value 0 is shifted like every other value
    s_y[r][c] = y - cy;
  }
  __syncthreads();
  for (int i = tid; i < LIN * LT; i += LT * LT) {   // horizontal pass: 26 rows x 16 columns
    const int r = i / LT, c = i - r * LT;
    float a = 0.f, b = 0.f, axx = 0.f, ayy = 0.f, axy = 0.f;
#pragma unroll
    for (int k = 0; k < LWIN; k++) {
      const float w = win.g[k], x = s_x[r][c + k], y = s_y[r][c + k];
      const float wx = w * x, wy = w * y;
      a += wx; b += wy;
      axx = fmaf(wx, x, axx); ayy = fmaf(wy, y, ayy); axy = fmaf(wx, y, axy);
    }
    s_h[0][r][c] = a; s_h[1][r][c] = b; s_h[2][r][c] = axx; s_h[3][r][c] = ayy; s_h[4][r][c] = axy;
  }
  __syncthreads();
  const int ty = tid / LT, tx = tid - ty * LT;
  const int py = by0 + ty, px = bx0 + tx;
  float map = 0.f, ad = 0.f;
  if (py < H && px < W) {
    float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < LWIN; k++) {
      const float w = win.g[k];
      m1 = fmaf(w, s_h[0][ty + k][tx], m1);
      m2 = fmaf(w, s_h[1][ty + k][tx], m2);
      e11 = fmaf(w, s_h[2][ty + k][tx], e11);
      e22 = fmaf(w, s_h[3][ty + k][tx], e22);
      e12 = fmaf(w, s_h[4][ty + k][tx], e12);
    }
    const float s11 = e11 - m1 * m1, s22 = e22 - m2 * m2, s12 = e12 - m1 * m2;   // shift-invariant
    const float mu1 = m1 + cx, mu2 = m2 + cy;
    constexpr float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float A1 = 2.f * mu1 * mu2 + C1, A2 = 2.f * s12 + C2;
    const float B1 = mu1 * mu1 + mu2 * mu2 + C1, B2 = s11 + s22 + C2;
    const float inv = 1.f / (B1 * B2);
    map = A1 * A2 * inv;
    ad = fabsf(X[(size_t)py * W + px] - Y[(size_t)py * W + px]);     // unshifted values (L1 hit): exact |x - y|
    if (GRAD) {
      // map as a function of (mu1, E[x^2], E[xy]) with sigma1^2 = E[x^2] - mu1^2, sigma12 = E[xy] - mu1 mu2
      const float d1 = 2.f * (mu2 * (A2 - A1) * inv - mu1 * map * (B2 - B1) * inv);
      const float d2 = -map / B2;
      const float d3 = 2.f * A1 * inv;
      const size_t o = plane + (size_t)py * W + px;
      dmaps[o] = d1;
      dmaps[plane_all + o] = d2;
      dmaps[2 * plane_all + o] = d3;
    }
  }
  const float sm = block_sum_256(map, s_red);
  const float sa = block_sum_256(ad, s_red);
  if (tid == 0) partials[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = make_float2(sm, sa);
}

__global__ void __launch_bounds__(LT * LT)
ssim_grad_kernel(int H, int W, const float* __restrict__ img, const float* __restrict__ gt, GaussWin win,
                 const float* __restrict__ dmaps, size_t plane_all, float wl1 /* g (1-l)/N */, float wss /* g l/N */,
                 float* __restrict__ dL_dimg) {
  __shared__ float s_d[3][LIN][LSTR];
  __shared__ float s_h[3][LIN][LT];
  const int tid = threadIdx.x;
  const int bx0 = blockIdx.x * LT, by0 = blockIdx.y * LT;
  const size_t plane = (size_t)blockIdx.z * H * W;
  for (int i = tid; i < LIN * LIN; i += LT * LT) {
    const int r = i / LIN, c = i - r * LIN;
    const int gy = by0 + r - LHALO, gx = bx0 + c - LHALO;
    float a = 0.f, b = 0.f, d = 0.f;           // map pixels outside the image do not exist
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const size_t o = plane + (size_t)gy * W + gx;
      a = dmaps[o]; b = dmaps[plane_all + o]; d = dmaps[2 * plane_all + o];
    }
    s_d[0][r][c] = a; s_d[1][r][c] = b; s_d[2][r][c] = d;
  }
  __syncthreads();
  for (int i = tid; i < LIN * LT; i += LT * LT) {
    const int r = i / LT, c = i - r * LT;
    float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
    for (int k = 0; k < LWIN; k++) {
      const float w = win.g[k];
      a = fmaf(w, s_d[0][r][c + k], a); b = fmaf(w, s_d[1][r][c + k], b); d = fmaf(w, s_d[2][r][c + k], d);
    }
    s_h[0][r][c] = a; s_h[1][r][c] = b; s_h[2][r][c] = d;
  }
  __syncthreads();
  const int ty = tid / LT, tx = tid - ty * LT;
  const int py = by0 + ty, px = bx0 + tx;
  if (py < H && px < W) {
    float c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
    for (int k = 0; k < LWIN; k++) {
      const float w = win.g[k];
      c1 = fmaf(w, s_h[0][ty + k][tx], c1); c2 = fmaf(w, s_h[1][ty + k][tx], c2); c3 = fmaf(w, s_h[2][ty + k][tx], c3);
    }
    const size_t o = plane + (size_t)py * W + px;
    const float x = img[o], y = gt[o];
    const float df = x - y;
    const float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);      // torch: d|t|/dt = sign(t), 0 at 0
    dL_dimg[o] = wl1 * sg - wss * (c1 + 2.f * x * c2 + y * c3);
  }
}

// mean |a - b| (CLAMP: a is clamped to [0, 1] first, train.py:190) with optional gradient; also the l1-only loss.
template <bool CLAMP>
__global__ void __launch_bounds__(256)
l1_kernel(size_t n, const float* __restrict__ a, const float* __restrict__ b, float wgt /* g w / n */,
          float* __restrict__ grad /* or nullptr */, float* __restrict__ partials /* [blocks] */) {
  __shared__ float s_red[8];
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float x = a[i];
    const float xc = CLAMP ? fminf(fmaxf(x, 0.f), 1.f) : x;
    const float df = xc - b[i];
    acc += fabsf(df);
    if (grad) {
      float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
      if (CLAMP && !(x >= 0.f && x <= 1.f)) sg = 0.f;      // clamp passes the gradient on [min, max] inclusive
      grad[i] = wgt * sg;
    }
  }
  const float t = block_sum_256(acc, s_red);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// One block; fixed summation order in fp64 -> the scalars are bit-reproducible run to run.
__global__ void __launch_bounds__(1024)
loss_finalize_kernel(const float2* __restrict__ p_ssim, int n_ssim, const float* __restrict__ p_l1, int n_l1,
                     const float* __restrict__ p_mask, int n_mask, double inv_n, double inv_nmask,
                     float lambda_dssim, float lambda_mask, float* __restrict__ out) {
  __shared__ double s_a[32], s_b[32], s_c[32];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < n_ssim; i += 1024) { const float2 v = p_ssim[i]; a += v.x; b += v.y; }
  for (int i = threadIdx.x; i < n_l1; i += 1024) b += p_l1[i];
  for (int i = threadIdx.x; i < n_mask; i += 1024) c += p_mask[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = b; s_c[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0, tc = 0.0;
    for (int w = 0; w < 32; w++) { ta += s_a[w]; tb += s_b[w]; tc += s_c[w]; }
    const double ssim = n_ssim > 0 ? ta * inv_n : 0.0;
    const double l1 = tb * inv_n;
    const double ml1 = n_mask > 0 ? tc * inv_nmask : 0.0;
    out[0] = (float)l1;
    out[1] = (float)ssim;
    out[2] = (float)ml1;
    out[3] = (float)((1.0 - (double)lambda_dssim) * l1 + (double)lambda_dssim * (1.0 - ssim) + (double)lambda_mask * ml1);
  }
}

constexpr int L1_BLOCKS = 4 * NUM_SMS_B200;

struct LossScratch {
  float* dmaps;        // [3][C*H*W]
  float2* p_ssim;      // [tiles * C]
  float* p_l1;         // [L1_BLOCKS]
  float* p_mask;       // [L1_BLOCKS]
  static LossScratch from_chunk(char*& chunk, size_t N, size_t nblk) {
    LossScratch s;
    s.dmaps = carve<float>(chunk, 3 * N);
    s.p_ssim = carve<float2>(chunk, nblk);
    s.p_l1 = carve<float>(chunk, L1_BLOCKS);
    s.p_mask = carve<float>(chunk, L1_BLOCKS);
    return s;
  }
};

}  // namespace sfb

extern "C" {

// the 11 window weights (test hook: compared bit-for-bit with the reference's gaussian(11, 1.5))
void sfb_loss_window(float* out11) {
  const sfb::GaussWin w = sfb::make_window();
  for (int i = 0; i < sfb::LWIN; i++) out11[i] = w.g[i];
}

size_t sfb_loss_scratch_bytes(int C, int H, int W) {
  using namespace sfb;
  if (C <= 0 || H <= 0 || W <= 0) return 0;
  const size_t N = (size_t)C * H * W;
  const size_t nblk = (size_t)((W + LT - 1) / LT) * ((H + LT - 1) / LT) * C;
  char* p = nullptr;
  LossScratch::from_chunk(p, N, nblk);
  return reinterpret_cast<size_t>(p) + 256;
}

int sfb_l1_ssim_loss(int C, int H, int W, const float* img, const float* gt, float lambda_dssim,
                     const float* opacity, const float* gt_mask, float lambda_mask, float grad_scale,
                     float* out_scalars, float* dL_dimg, float* dL_dopacity, void* scratch, void* stream) {
  using namespace sfb;
  cudaStream_t s = (cudaStream_t)stream;
  if (C <= 0 || H <= 0 || W <= 0 || !img || !gt || !out_scalars || !scratch)
    return sfb::set_error("sfb_l1_ssim_loss: bad sizes / null pointer"), SFB_ERR_ARG;
  if ((opacity == nullptr) != (gt_mask == nullptr))
    return sfb::set_error("sfb_l1_ssim_loss: opacity and gt_mask go together"), SFB_ERR_ARG;
  const int gx = (W + LT - 1) / LT, gy = (H + LT - 1) / LT;
  if (gy > 65535 || C > 65535) return sfb::set_error("sfb_l1_ssim_loss: image too large"), SFB_ERR_ARG;
  const size_t N = (size_t)C * H * W, HW = (size_t)H * W;
  const size_t nblk = (size_t)gx * gy * C;
  char* chunk = (char*)scratch;
  LossScratch ls = LossScratch::from_chunk(chunk, N, nblk);
  static const GaussWin win = make_window();
  const double inv_n = 1.0 / (double)N, inv_hw = 1.0 / (double)HW;
  const bool with_ssim = lambda_dssim != 0.f;
  int n_ssim = 0, n_l1 = 0, n_mask = 0;
  if (with_ssim) {
    const dim3 grid(gx, gy, C);
    prof_begin("loss.ssim_stats", s);
    if (dL_dimg) ssim_stats_kernel<true><<<grid, LT * LT, 0, s>>>(H, W, img, gt, win, ls.dmaps, N, ls.p_ssim);
    else ssim_stats_kernel<false><<<grid, LT * LT, 0, s>>>(H, W, img, gt, win, ls.dmaps, N, ls.p_ssim);
    prof_end(s);
    n_ssim = (int)nblk;
    if (dL_dimg) {
      prof_begin("loss.ssim_grad", s);
      ssim_grad_kernel<<<grid, LT * LT, 0, s>>>(H, W, img, gt, win, ls.dmaps, N,
                                                (float)((double)grad_scale * (1.0 - (double)lambda_dssim) * inv_n),
                                                (float)((double)grad_scale * (double)lambda_dssim * inv_n), dL_dimg);
      prof_end(s);
    }
  } else {
    n_l1 = (int)((N + 255) / 256 < (size_t)L1_BLOCKS ? (N + 255) / 256 : (size_t)L1_BLOCKS);
    prof_begin("loss.l1", s);
    l1_kernel<false><<<n_l1, 256, 0, s>>>(N, img, gt, (float)((double)grad_scale * inv_n), dL_dimg, ls.p_l1);
    prof_end(s);
  }
  if (opacity) {
    n_mask = (int)((HW + 255) / 256 < (size_t)L1_BLOCKS ? (HW + 255) / 256 : (size_t)L1_BLOCKS);
    prof_begin("loss.mask_l1", s);
    l1_kernel<true><<<n_mask, 256, 0, s>>>(HW, opacity, gt_mask, (float)((double)grad_scale * (double)lambda_mask * inv_hw),
                                           dL_dopacity, ls.p_mask);
    prof_end(s);
  }
  prof_begin("loss.finalize", s);
  loss_finalize_kernel<<<1, 1024, 0, s>>>(ls.p_ssim, n_ssim, ls.p_l1, n_l1, ls.p_mask, n_mask, inv_n, inv_hw, lambda_dssim,
                                          opacity ? lambda_mask : 0.f, out_scalars);
  prof_end(s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return sfb::set_error(cudaGetErrorString(e)), SFB_ERR_CUDA;
  return SFB_OK;
}

}  // extern "C"
