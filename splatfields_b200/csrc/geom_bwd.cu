// geom_bwd.cu — K8 + K9 fused: per-Gaussian backward of the EWA projection, the screen-space mean,
// the SH colour and the 3D covariance.
//
// One thread per Gaussian.  Inputs are the pixel sums the backward render left in GradRec; every output
// row is written exactly once (zeros for culled Gaussians), so the caller never has to pre-zero the
// ~256 B/Gaussian of gradient tensors — that removes a full memset pass over the largest buffers
// (dL_dsh alone is 192 B/Gaussian at degree 3).
#include "common.cuh"
#include "xchg.cuh"
#include <cstdlib>

namespace sfb {

// (SH_C0 / SH_C1 come from xchg.cuh)
__constant__ float SHB_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                -1.0925484305920792f, 0.5462742152960396f};
__constant__ float SHB_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                -0.5900435899266435f};

struct CamB {
  float view[16], proj[16], campos[3];
};

// VEC: shs / dL_dsh rows are 16-byte aligned and 3*M is a multiple of 4 -> 128-bit loads and stores.
// (Bulk-copy staging of the block's SH / dL_dsh slabs through shared memory was measured SLOWER here — 150 vs 134 us at
// 1M splats, round 1: the whole 48 KB slab has to land before any thread can start and only 2 CTAs fit per SM — and
// was removed; the rows move as L2-prefetched 256-bit loads and 256-bit stores.)
// FACT: factored SH gradient (SFB_BWD_SH_FACTORED) — dL_dcolors receives the clamp-masked colour gradient and the
// dL_dsh rows are not written (exchange.cu rebuilds their multi-view sum).  A template flag, so
// that the default instantiations carry no trace of it (a run-time branch cost 4 registers and 10 us at 1M splats).
// PUSH: view-parallel exchange over NVLink (exchange.cu): the 11 (SH colours) / 14 (precomputed colours)
// parameter gradients leave as ONE packed record per Gaussian in this rank's symmetric buffer (where the reduction
// kernel's multimem.ld_reduce finds them), and with FACT the clamp-masked colour gradient is written straight into
// slot `rank` of EVERY rank's buffer while this kernel is still computing — one multimem.st per 16 bytes through
// the NVSwitch multicast mapping (or one st.global per peer without multicast): the all-gather rides on the kernel.
// The body works on the 256 Gaussians of block `blk`: the plain kernel below runs it once per CTA, the fused
// backward + exchange kernel (further down) runs it from a work queue.
template <int D, bool VEC, bool W256, bool FACT, bool PUSH>
__device__ __forceinline__ void geom_backward_block(const BwdParams& p, const GeomState& g, const CamB& cam, const int blk) {
  const int idx = blk * 256 + (int)threadIdx.x;
  const bool in_range = idx < p.P;
  if (!PUSH && !in_range) return;          // (PUSH: the whole block meets again at the colour-gradient hand-off)
  const size_t i = (size_t)(in_range ? idx : 0);
  // Every input of the covariance chain is requested together with the radius, before the visibility decision that
  // depends on it: one memory round trip instead of three dependent ones (ncu, round 2: 21 % of the kernel's stall samples
  // sat on the first use of the radius and of the rotation / scale).  Four of five splats are visible and the rows of
  // neighbouring splats share their 128-byte lines, so the unconditional loads add next to no DRAM traffic.
  int rad = 0;
  float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0, g2 = g0;
  float mean[3] = {0.f, 0.f, 0.f};
  float hq[4] = {1.f, 0.f, 0.f, 0.f}, hs[3] = {0.f, 0.f, 0.f};
  if (in_range) {
    rad = p.radii[idx];
    const float4* gp = reinterpret_cast<const float4*>(g.grad + idx);
    g0 = gp[0]; g1 = gp[1]; g2 = gp[2];
    mean[0] = p.means3D[3 * i]; mean[1] = p.means3D[3 * i + 1]; mean[2] = p.means3D[3 * i + 2];
    if (!p.cov3D_precomp) {
      hq[0] = p.rotations[4 * i]; hq[1] = p.rotations[4 * i + 1]; hq[2] = p.rotations[4 * i + 2]; hq[3] = p.rotations[4 * i + 3];
      hs[0] = p.scales[3 * i]; hs[1] = p.scales[3 * i + 1]; hs[2] = p.scales[3 * i + 2];
    }
  }
  const bool visible = rad > 0;
  if (visible && p.shs) {   // pull the SH row towards L2 while the covariance math runs
    const char* row = reinterpret_cast<const char*>(p.shs + i * p.M * 3);
    prefetch_l2(row);
    prefetch_l2(row + 128);
  }

  float dmean[3] = {0.f, 0.f, 0.f};
  float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float dscale[3] = {0.f, 0.f, 0.f};
  float drot[4] = {0.f, 0.f, 0.f, 0.f};
  float gm2[2] = {0.f, 0.f};
  float gcol[3] = {0.f, 0.f, 0.f};
  float gop = 0.f;
  constexpr int NB = (D + 1) * (D + 1);

  if (visible) {
    gm2[0] = g0.x; gm2[1] = g0.y;
    const float gA = g0.z, gB = g0.w, gC = g1.x;
    gop = g1.y;
    gcol[0] = g1.z; gcol[1] = g1.w; gcol[2] = g2.x;

    // Sigma3D: re-derived (scale / rotation) or re-read (precomputed) instead of round-tripping 24 B through HBM
    float c6[6];
    if (p.cov3D_precomp) {
#pragma unroll
      for (int k = 0; k < 6; k++) c6[k] = p.cov3D_precomp[6 * i + k];
    } else {   // (R and s are re-derived again at the end instead of being kept live: registers, not ALU, are scarce)
      float Rm[9], sc3[3];
      const float qr = hq[0], qx = hq[1], qy = hq[2], qz = hq[3];
      Rm[0] = 1.f - 2.f * (qy * qy + qz * qz); Rm[1] = 2.f * (qx * qy - qr * qz); Rm[2] = 2.f * (qx * qz + qr * qy);
      Rm[3] = 2.f * (qx * qy + qr * qz); Rm[4] = 1.f - 2.f * (qx * qx + qz * qz); Rm[5] = 2.f * (qy * qz - qr * qx);
      Rm[6] = 2.f * (qx * qz - qr * qy); Rm[7] = 2.f * (qy * qz + qr * qx); Rm[8] = 1.f - 2.f * (qx * qx + qy * qy);
      sc3[0] = p.scale_modifier * hs[0]; sc3[1] = p.scale_modifier * hs[1]; sc3[2] = p.scale_modifier * hs[2];
      float L[9];
#pragma unroll
      for (int r = 0; r < 3; r++) { L[3 * r] = Rm[3 * r] * sc3[0]; L[3 * r + 1] = Rm[3 * r + 1] * sc3[1]; L[3 * r + 2] = Rm[3 * r + 2] * sc3[2]; }
      c6[0] = L[0] * L[0] + L[1] * L[1] + L[2] * L[2];
      c6[1] = L[0] * L[3] + L[1] * L[4] + L[2] * L[5];
      c6[2] = L[0] * L[6] + L[1] * L[7] + L[2] * L[8];
      c6[3] = L[3] * L[3] + L[4] * L[4] + L[5] * L[5];
      c6[4] = L[3] * L[6] + L[4] * L[7] + L[5] * L[8];
      c6[5] = L[6] * L[6] + L[7] * L[7] + L[8] * L[8];
    }
    const float fx = (float)p.W / (2.0f * p.tan_fovx), fy = (float)p.H / (2.0f * p.tan_fovy);

    // ---- conic -> cov2D -> (Sigma3D, view-space t) ----
    float t[3];
#pragma unroll
    for (int r = 0; r < 3; r++)
      t[r] = cam.view[r] * mean[0] + cam.view[4 + r] * mean[1] + cam.view[8 + r] * mean[2] + cam.view[12 + r];
    const float limx = 1.3f * p.tan_fovx, limy = 1.3f * p.tan_fovy;
    const float txtz = t[0] / t[2], tytz = t[1] / t[2];
    const float xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
    t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
    const float itz = 1.f / t[2], itz2 = itz * itz, itz3 = itz2 * itz;
    const float J00 = fx * itz, J02 = -(fx * t[0]) * itz2, J11 = fy * itz, J12 = -(fy * t[1]) * itz2;
    float m0[3], m1[3], v0[3], v1[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      m0[j] = J00 * cam.view[4 * j + 0] + J02 * cam.view[4 * j + 2];
      m1[j] = J11 * cam.view[4 * j + 1] + J12 * cam.view[4 * j + 2];
    }
    v0[0] = c6[0] * m0[0] + c6[1] * m0[1] + c6[2] * m0[2];
    v0[1] = c6[1] * m0[0] + c6[3] * m0[1] + c6[4] * m0[2];
    v0[2] = c6[2] * m0[0] + c6[4] * m0[1] + c6[5] * m0[2];
    v1[0] = c6[0] * m1[0] + c6[1] * m1[1] + c6[2] * m1[2];
    v1[1] = c6[1] * m1[0] + c6[3] * m1[1] + c6[4] * m1[2];
    v1[2] = c6[2] * m1[0] + c6[4] * m1[1] + c6[5] * m1[2];
    const float a = m0[0] * v0[0] + m0[1] * v0[1] + m0[2] * v0[2] + 0.3f;
    const float b = m1[0] * v0[0] + m1[1] * v0[1] + m1[2] * v0[2];
    const float c = m1[0] * v1[0] + m1[1] * v1[1] + m1[2] * v1[2] + 0.3f;
    const float denom = a * c - b * b;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    if (denom2inv != 0.f) {
      dL_da = denom2inv * (-c * c * gA + b * c * gB + (denom - a * c) * gC);
      dL_dc = denom2inv * (-a * a * gC + a * b * gB + (denom - a * c) * gA);
      dL_db = denom2inv * (2.f * b * c * gA - (denom + 2.f * b * b) * gB + 2.f * a * b * gC);
      dcov[0] = m0[0] * m0[0] * dL_da + m0[0] * m1[0] * dL_db + m1[0] * m1[0] * dL_dc;
      dcov[3] = m0[1] * m0[1] * dL_da + m0[1] * m1[1] * dL_db + m1[1] * m1[1] * dL_dc;
      dcov[5] = m0[2] * m0[2] * dL_da + m0[2] * m1[2] * dL_db + m1[2] * m1[2] * dL_dc;
      dcov[1] = 2.f * m0[0] * m0[1] * dL_da + (m0[0] * m1[1] + m0[1] * m1[0]) * dL_db + 2.f * m1[0] * m1[1] * dL_dc;
      dcov[2] = 2.f * m0[0] * m0[2] * dL_da + (m0[0] * m1[2] + m0[2] * m1[0]) * dL_db + 2.f * m1[0] * m1[2] * dL_dc;
      dcov[4] = 2.f * m0[2] * m0[1] * dL_da + (m0[1] * m1[2] + m0[2] * m1[1]) * dL_db + 2.f * m1[1] * m1[2] * dL_dc;
    }
    float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const float dm0 = 2.f * v0[j] * dL_da + v1[j] * dL_db;
      const float dm1 = 2.f * v1[j] * dL_dc + v0[j] * dL_db;
      dJ00 += cam.view[4 * j + 0] * dm0;
      dJ02 += cam.view[4 * j + 2] * dm0;
      dJ11 += cam.view[4 * j + 1] * dm1;
      dJ12 += cam.view[4 * j + 2] * dm1;
    }
    const float dtx = xmul * -fx * itz2 * dJ02;
    const float dty = ymul * -fy * itz2 * dJ12;
    const float dtz = -fx * itz2 * dJ00 - fy * itz2 * dJ11 + (2.f * fx * t[0]) * itz3 * dJ02 +
                      (2.f * fy * t[1]) * itz3 * dJ12;
    // (g2.y = GradRec::dz: gradient w.r.t. the view-space depth, non-zero only when the depth image had a cotangent;
    //  depth = view[2] x + view[6] y + view[10] z + view[14])
#pragma unroll
    for (int j = 0; j < 3; j++)
      dmean[j] = cam.view[4 * j + 0] * dtx + cam.view[4 * j + 1] * dty + cam.view[4 * j + 2] * (dtz + g2.y);

    // ---- NDC-scaled screen mean -> mean3D ----
    {
      float h[4];
#pragma unroll
      for (int r = 0; r < 4; r++)
        h[r] = cam.proj[r] * mean[0] + cam.proj[4 + r] * mean[1] + cam.proj[8 + r] * mean[2] + cam.proj[12 + r];
      const float m_w = 1.0f / (h[3] + 0.0000001f);
      const float mul1 = h[0] * m_w * m_w, mul2 = h[1] * m_w * m_w;
#pragma unroll
      for (int j = 0; j < 3; j++)
        dmean[j] += (cam.proj[4 * j + 0] * m_w - cam.proj[4 * j + 3] * mul1) * gm2[0] +
                    (cam.proj[4 * j + 1] * m_w - cam.proj[4 * j + 3] * mul2) * gm2[1];
    }

    // ---- colour -> SH coefficients and view direction ----
    if (p.shs) {
      constexpr int NF4 = (3 * NB + 3) / 4;   // float4s that hold the active coefficients
      float sh[NF4 * 4];
      float dshv[NF4 * 4];
      {
        const float* shp = p.shs + i * p.M * 3;
        if (VEC && W256 && (3 * NB) % 8 == 0) {
#pragma unroll
          for (int k = 0; k < (3 * NB) / 8; k++) ldg256(shp + 8 * k, sh + 8 * k);
        } else if (VEC) {
#pragma unroll
          for (int k = 0; k < NF4; k++) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(shp) + k);
            sh[4 * k] = q.x; sh[4 * k + 1] = q.y; sh[4 * k + 2] = q.z; sh[4 * k + 3] = q.w;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 3 * NB; k++) sh[k] = __ldg(shp + k);
        }
      }
      float* dsh = p.dL_dsh + i * p.M * 3;
      const float vx = mean[0] - cam.campos[0], vy = mean[1] - cam.campos[1], vz = mean[2] - cam.campos[2];
      const float sum2 = vx * vx + vy * vy + vz * vz;
      const float ilen = rsqrtf(sum2);
      const float x = vx * ilen, y = vy * ilen, z = vz * ilen;
      const uint8_t cm = g.clamped[idx];
      float gc[3];
#pragma unroll
      for (int ch = 0; ch < 3; ch++) gc[ch] = ((cm >> ch) & 1) ? 0.f : gcol[ch];
      if (FACT) { gcol[0] = gc[0]; gcol[1] = gc[1]; gcol[2] = gc[2]; }   // dL_dcolors <- clamp-masked gradient
      float ddx = 0.f, ddy = 0.f, ddz = 0.f;
      float basis[NB];
      basis[0] = SH_C0;
      float xx = 0, yy = 0, zz = 0, xy = 0, yz = 0, xz = 0;
      if (D > 0) { basis[1] = -SH_C1 * y; basis[2] = SH_C1 * z; basis[3] = -SH_C1 * x; }
      if (D > 1) {
        xx = x * x; yy = y * y; zz = z * z; xy = x * y; yz = y * z; xz = x * z;
        basis[4] = SHB_C2[0] * xy; basis[5] = SHB_C2[1] * yz; basis[6] = SHB_C2[2] * (2.f * zz - xx - yy);
        basis[7] = SHB_C2[3] * xz; basis[8] = SHB_C2[4] * (xx - yy);
      }
      if (D > 2) {
        basis[9] = SHB_C3[0] * y * (3.f * xx - yy); basis[10] = SHB_C3[1] * xy * z;
        basis[11] = SHB_C3[2] * y * (4.f * zz - xx - yy); basis[12] = SHB_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
        basis[13] = SHB_C3[4] * x * (4.f * zz - xx - yy); basis[14] = SHB_C3[5] * z * (xx - yy);
        basis[15] = SHB_C3[6] * x * (xx - 3.f * yy);
      }
#pragma unroll
      for (int k = 0; k < NF4 * 4; k++) dshv[k] = 0.f;
#pragma unroll
      for (int k = 0; k < NB; k++) {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) dshv[3 * k + ch] = basis[k] * gc[ch];
      }
      if (FACT) {
        // the 3*M-float row is rebuilt later from the masked colour gradients of all views
      } else if (VEC && W256 && (3 * NB) % 8 == 0) {   // launcher guarantees M == (D+1)^2 here
#pragma unroll
        for (int k = 0; k < (3 * NB) / 8; k++) stg256(dsh + 8 * k, dshv + 8 * k);
      } else if (VEC) {
        float4* d4 = reinterpret_cast<float4*>(dsh);
#pragma unroll
        for (int k = 0; k < NF4; k++) d4[k] = make_float4(dshv[4 * k], dshv[4 * k + 1], dshv[4 * k + 2], dshv[4 * k + 3]);
        for (int k = NF4; k < (3 * p.M) / 4; k++) d4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
#pragma unroll
        for (int k = 0; k < 3 * NB; k++) dsh[k] = dshv[k];
        for (int k = 3 * NB; k < 3 * p.M; k++) dsh[k] = 0.f;
      }
      if (D > 0) {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
          float rx = -SH_C1 * sh[3 * 3 + ch], ry = -SH_C1 * sh[1 * 3 + ch], rz = SH_C1 * sh[2 * 3 + ch];
          if (D > 1) {
            const float s4 = sh[4 * 3 + ch], s5 = sh[5 * 3 + ch], s6 = sh[6 * 3 + ch], s7 = sh[7 * 3 + ch],
                        s8 = sh[8 * 3 + ch];
            rx += SHB_C2[0] * y * s4 + SHB_C2[2] * 2.f * -x * s6 + SHB_C2[3] * z * s7 + SHB_C2[4] * 2.f * x * s8;
            ry += SHB_C2[0] * x * s4 + SHB_C2[1] * z * s5 + SHB_C2[2] * 2.f * -y * s6 + SHB_C2[4] * 2.f * -y * s8;
            rz += SHB_C2[1] * y * s5 + SHB_C2[2] * 4.f * z * s6 + SHB_C2[3] * x * s7;
            if (D > 2) {
              const float s9 = sh[9 * 3 + ch], s10 = sh[10 * 3 + ch], s11 = sh[11 * 3 + ch], s12 = sh[12 * 3 + ch],
                          s13 = sh[13 * 3 + ch], s14 = sh[14 * 3 + ch], s15 = sh[15 * 3 + ch];
              rx += SHB_C3[0] * s9 * 6.f * xy + SHB_C3[1] * s10 * yz + SHB_C3[2] * s11 * -2.f * xy +
                    SHB_C3[3] * s12 * -6.f * xz + SHB_C3[4] * s13 * (-3.f * xx + 4.f * zz - yy) +
                    SHB_C3[5] * s14 * 2.f * xz + SHB_C3[6] * s15 * 3.f * (xx - yy);
              ry += SHB_C3[0] * s9 * 3.f * (xx - yy) + SHB_C3[1] * s10 * xz +
                    SHB_C3[2] * s11 * (-3.f * yy + 4.f * zz - xx) + SHB_C3[3] * s12 * -6.f * yz +
                    SHB_C3[4] * s13 * -2.f * xy + SHB_C3[5] * s14 * -2.f * yz + SHB_C3[6] * s15 * -6.f * xy;
              rz += SHB_C3[1] * s10 * xy + SHB_C3[2] * s11 * 8.f * yz + SHB_C3[3] * s12 * 3.f * (2.f * zz - xx - yy) +
                    SHB_C3[4] * s13 * 8.f * xz + SHB_C3[5] * s14 * (xx - yy);
            }
          }
          ddx += rx * gc[ch]; ddy += ry * gc[ch]; ddz += rz * gc[ch];
        }
        const float invsum32 = ilen * ilen * ilen;
        dmean[0] += ((sum2 - vx * vx) * ddx - vy * vx * ddy - vz * vx * ddz) * invsum32;
        dmean[1] += (-vx * vy * ddx + (sum2 - vy * vy) * ddy - vz * vy * ddz) * invsum32;
        dmean[2] += (-vx * vz * ddx - vy * vz * ddy + (sum2 - vz * vz) * ddz) * invsum32;
      }
    }

    // ---- Sigma3D -> scale, quaternion ----
    if (!p.cov3D_precomp) {
      const float qr = p.rotations[4 * i], qx = p.rotations[4 * i + 1], qy = p.rotations[4 * i + 2],
                  qz = p.rotations[4 * i + 3];
      float R[9];
      R[0] = 1.f - 2.f * (qy * qy + qz * qz); R[1] = 2.f * (qx * qy - qr * qz); R[2] = 2.f * (qx * qz + qr * qy);
      R[3] = 2.f * (qx * qy + qr * qz); R[4] = 1.f - 2.f * (qx * qx + qz * qz); R[5] = 2.f * (qy * qz - qr * qx);
      R[6] = 2.f * (qx * qz - qr * qy); R[7] = 2.f * (qy * qz + qr * qx); R[8] = 1.f - 2.f * (qx * qx + qy * qy);
      const float s[3] = {p.scale_modifier * p.scales[3 * i], p.scale_modifier * p.scales[3 * i + 1],
                          p.scale_modifier * p.scales[3 * i + 2]};
      const float Gm[9] = {dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                           0.5f * dcov[2], 0.5f * dcov[4], dcov[5]};
      float Dr[9];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float ds = 0.f;
#pragma unroll
        for (int r = 0; r < 3; r++) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 3; k++) acc += Gm[3 * r + k] * (R[3 * k + j] * s[j]);
          const float dLm = 2.f * acc;
          ds += dLm * R[3 * r + j];
          Dr[3 * r + j] = dLm * s[j];
        }
        dscale[j] = p.scale_modifier * ds;
      }
      drot[0] = 2.f * (qz * (Dr[3] - Dr[1]) + qy * (Dr[2] - Dr[6]) + qx * (Dr[7] - Dr[5]));
      drot[1] = 2.f * (qy * (Dr[1] + Dr[3]) + qz * (Dr[2] + Dr[6]) + qr * (Dr[7] - Dr[5])) - 4.f * qx * (Dr[4] + Dr[8]);
      drot[2] = 2.f * (qx * (Dr[1] + Dr[3]) + qr * (Dr[2] - Dr[6]) + qz * (Dr[5] + Dr[7])) - 4.f * qy * (Dr[0] + Dr[8]);
      drot[3] = 2.f * (qr * (Dr[3] - Dr[1]) + qx * (Dr[2] + Dr[6]) + qy * (Dr[5] + Dr[7])) - 4.f * qz * (Dr[0] + Dr[4]);
    }
  } else if (in_range && p.shs && !FACT) {
    float* dsh = p.dL_dsh + i * p.M * 3;
    if (VEC && W256) {
      const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < (3 * p.M) / 8; k++) stg256(dsh + 8 * k, zero8);
    } else if (VEC) {
      float4* d4 = reinterpret_cast<float4*>(dsh);
      for (int k = 0; k < (3 * p.M) / 4; k++) d4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int k = 0; k < 3 * p.M; k++) dsh[k] = 0.f;
    }
  }

  if (PUSH) {
    // the block's colour gradients as 192 contiguous 16-byte chunks: one multicast (or per-peer) store each
    __shared__ __align__(16) float s_gc[256 * 3];
    if (FACT) {
      s_gc[3 * threadIdx.x] = gcol[0]; s_gc[3 * threadIdx.x + 1] = gcol[1]; s_gc[3 * threadIdx.x + 2] = gcol[2];
      __syncthreads();
      const size_t row0 = (size_t)blk * 256;
      const int nfl = 3 * (int)min((size_t)256, (size_t)p.P - row0);       // floats of this block's slab
      if ((int)threadIdx.x * 4 < nfl) {
        const float4 v = reinterpret_cast<const float4*>(s_gc)[threadIdx.x];
        const size_t off = row0 * 3 + (size_t)threadIdx.x * 4;                // float offset inside the [P][3] slot
        if ((int)threadIdx.x * 4 + 4 <= nfl) {
          if (p.x_mc) {
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p.x_gc_dst[0] + off), "f"(v.x),
                         "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
          } else {
            for (int r = 0; r < p.x_ndst; r++) *reinterpret_cast<float4*>(p.x_gc_dst[r] + off) = v;
          }
        } else {     // ragged tail of the last block (P * 3 not a multiple of 4): scalar stores to every rank
          const float vv[4] = {v.x, v.y, v.z, v.w};
          for (int e = 0; (int)threadIdx.x * 4 + e < nfl; e++)
            for (int r = 0; r < p.x_nranks; r++) p.x_gc_peer[r][off + e] = vv[e];
        }
      }
    }
    if (!in_range) return;
    float4* rec = reinterpret_cast<float4*>(p.x_geo + i * (size_t)p.x_ngeo);
    rec[0] = make_float4(dmean[0], dmean[1], dmean[2], gop);
    rec[1] = make_float4(dscale[0], dscale[1], dscale[2], drot[0]);
    rec[2] = make_float4(drot[1], drot[2], drot[3], 0.f);
    if (!FACT) rec[3] = make_float4(gcol[0], gcol[1], gcol[2], 0.f);          // precomputed colours: 14 floats (+2 pad)
    p.dL_dmeans2D[3 * i] = gm2[0]; p.dL_dmeans2D[3 * i + 1] = gm2[1]; p.dL_dmeans2D[3 * i + 2] = 0.f;
    return;
  }
  p.dL_dmeans3D[3 * i] = dmean[0]; p.dL_dmeans3D[3 * i + 1] = dmean[1]; p.dL_dmeans3D[3 * i + 2] = dmean[2];
  p.dL_dmeans2D[3 * i] = gm2[0]; p.dL_dmeans2D[3 * i + 1] = gm2[1]; p.dL_dmeans2D[3 * i + 2] = 0.f;
  if (p.dL_dcolors) { p.dL_dcolors[3 * i] = gcol[0]; p.dL_dcolors[3 * i + 1] = gcol[1]; p.dL_dcolors[3 * i + 2] = gcol[2]; }
  p.dL_dopacity[i] = gop;
  if (p.dL_dcov3D) {
#pragma unroll
    for (int k = 0; k < 6; k++) p.dL_dcov3D[6 * i + k] = dcov[k];
  }
  if (p.dL_dscales) { p.dL_dscales[3 * i] = dscale[0]; p.dL_dscales[3 * i + 1] = dscale[1]; p.dL_dscales[3 * i + 2] = dscale[2]; }
  if (p.dL_drot) { p.dL_drot[4 * i] = drot[0]; p.dL_drot[4 * i + 1] = drot[1]; p.dL_drot[4 * i + 2] = drot[2]; p.dL_drot[4 * i + 3] = drot[3]; }
}

__device__ __forceinline__ void load_cam(CamB& cam, const BwdParams& p) {
  if (threadIdx.x < 16) {
    cam.view[threadIdx.x] = p.viewmatrix[threadIdx.x];
    cam.proj[threadIdx.x] = p.projmatrix[threadIdx.x];
  }
  if (threadIdx.x < 3) cam.campos[threadIdx.x] = p.campos[threadIdx.x];
}

template <int D, bool VEC, int MINB = 1, bool W256 = false, bool FACT = false, bool PUSH = false>
__global__ void __launch_bounds__(256, MINB) geom_backward_kernel(BwdParams p, GeomState g) {
  __shared__ CamB cam;
  load_cam(cam, p);
  __syncthreads();
  geom_backward_block<D, VEC, W256, FACT, PUSH>(p, g, cam, (int)blockIdx.x);
}

// ---------------------------------------------------------------------------------------------------------------
// Fused geometry backward + view-parallel exchange: ONE persistent kernel per step and rank.
//
// The exchange of a step is bound by NVLink (60-150 MB per rank), the geometry backward, the SH row rebuild and the
// unpacking by HBM; run one after the other (geom_backward_kernel<PUSH>, then sfb_xchg_finish) the two resources idle
// in turn.  Here the splats are cut into chunks of XCHG_CHUNK and every CTA pulls work units from four queues:
//   G(c)  geometry backward of chunk c (4 blocks of 256 splats): packed records into this rank's buffer, colour gradients
//         pushed into every rank's table (multimem.st), then flag A[rank][c] raised in EVERY rank's buffer;
//   X(c)  NVLink unit of the chunk's owner (c % world == rank), ready when A[*][c] is up on this rank: sum the ranks'
//         records of chunk c inside the switch (multimem.ld_reduce) and broadcast the sums in place (multimem.st; peer
//         loads / stores without multicast), then raise flag B[c] everywhere;
//   S(c)  SH gradient rows of chunk c from the local colour table, ready when A[*][c] is up;
//   U(c)  unpack the broadcast sums of chunk c, ready when B[c] is up.
// Every fourth CTA prefers X over G (the NVLink round trips of a chunk start while later chunks are still being
// differentiated), the others run G, then S, then U; whatever is ready is taken when a CTA's preferred queue is empty.
// Queues are tickets in this rank's flag words, claims are CAS, waits are bounded (error word, sfb_xchg_status).
constexpr int FLAG_TK = 48;        // [+0..3] tickets of the queues G, X, S, U; [+4] abort (a wait timed out)
enum { FK_EXIT = 0, FK_G = 1, FK_X = 2, FK_S = 3, FK_U = 4 };

__device__ __forceinline__ uint32_t ld_relaxed_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <bool HAS_SH>
__device__ __forceinline__ void fused_store_sums(const FusedXchg& f, size_t i, const float4 a, const float4 b, const float4 c,
                                                 const float4 d) {
  f.dL_dmeans3D[3 * i] = a.x; f.dL_dmeans3D[3 * i + 1] = a.y; f.dL_dmeans3D[3 * i + 2] = a.z;
  f.dL_dopacity[i] = a.w;
  f.dL_dscales[3 * i] = b.x; f.dL_dscales[3 * i + 1] = b.y; f.dL_dscales[3 * i + 2] = b.z;
  f.dL_drot[4 * i] = b.w; f.dL_drot[4 * i + 1] = c.x; f.dL_drot[4 * i + 2] = c.y; f.dL_drot[4 * i + 3] = c.z;
  if (!HAS_SH) { f.dL_dcolors[3 * i] = d.x; f.dL_dcolors[3 * i + 1] = d.y; f.dL_dcolors[3 * i + 2] = d.z; }
}

template <int D, bool VEC, bool W256, bool HAS_SH>
__global__ void __launch_bounds__(256, 2) geom_exchange_fused_kernel(BwdParams p, GeomState g, FusedXchg f) {
  __shared__ CamB cam;
  __shared__ float s_camv[3 * XCHG_MAX_RANKS];
  __shared__ int s_kind, s_unit;
  const XchgDev& x = f.x;
  load_cam(cam, p);
  if (HAS_SH) for (int k = threadIdx.x; k < 3 * f.V; k += 256) s_camv[k] = f.campos_views[k];
  const int N = x.world, rank = x.rank, nch = x.nch;
  const int nblk = (x.P + 255) / 256;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t epoch = f.epoch;
  const uint32_t nG = (uint32_t)nch;
  const uint32_t nX = (uint32_t)(nch > rank ? (nch - rank + N - 1) / N : 0);
  const uint32_t nS = HAS_SH ? (uint32_t)nch : 0u;
  const uint32_t nU = (uint32_t)nch;
  const bool link_cta = (blockIdx.x & 3) == 0;
  uint32_t* const tk = x.flags + FLAG_TK;
  uint32_t tG = 0, tX = 0, tS = 0, tU = 0;      // this CTA's view of the tickets (they only grow)
  constexpr int NG4 = HAS_SH ? 3 : 4;           // 16-byte words per packed record
  if (threadIdx.x == 0) xchg_mark(x.flags, 0, true);
  __syncthreads();
  // where this CTA's time goes (thread 0; sfb_xchg_timeline): choosing work, G bodies, G flag releases, X, S, U units
  __shared__ long long s_cyc[7];      // (shared memory: registers are scarce here; [6] = start of the current lap)
  if (threadIdx.x < 6) s_cyc[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_cyc[6] = clock64();
  auto lap = [&](int slot) { const long long now = clock64(); s_cyc[slot] += now - s_cyc[6]; s_cyc[6] = now; };
  __syncthreads();

  for (;;) {
    if (warp == 0) {
      int kind = -1, unit = 0;
      long long t0 = 0;
      bool waiting = false;
      // flags of chunk c that gate a unit of queue q (1: X, 2: S -> A[*][c]; 3: U -> B[c]); all lanes get the answer
      auto chunk_ready = [&](int q, uint32_t c) -> bool {
        uint32_t fl = epoch;
        if (q == 3) { if (lane == 0) fl = ld_relaxed_sys_u32(x.cflags + (size_t)XCHG_MAX_RANKS * nch + c); }
        else if (lane < N) fl = ld_relaxed_sys_u32(x.cflags + (size_t)lane * nch + c);
        return __all_sync(0xffffffffu, (int32_t)(fl - epoch) >= 0);
      };
      // bounded wait (the other ranks are behind); false: gave up, the kernel retires with an error word
      auto keep_waiting = [&]() -> bool {
        if (!waiting) { waiting = true; t0 = clock64(); }
        bool give_up = ld_relaxed_sys_u32(tk + 4) != 0u;
        if (!give_up && clock64() - t0 > 4000000000LL) {
          give_up = true;
          if (lane == 0) { atomicMax(x.flags + FLAG_ERR, 3u | (epoch << 8)); atomicExch(tk + 4, 1u); }
        }
        give_up = __any_sync(0xffffffffu, give_up);
        if (!give_up) __nanosleep(200);
        return !give_up;
      };
      // A unit of queue q is taken with ONE atomicAdd once the queue's next unit (as this CTA last saw it) is ready:
      // every contender gets a ticket of its own (a compare-and-swap claim lets one CTA through per L2 round trip), and
      // the ticket it gets may lie a little further on than the one it looked at, in which case it waits for that chunk.
      auto take = [&](int q, uint32_t& t, uint32_t n, int k) {
        uint32_t got = 0;
        if (lane == 0) got = atomicAdd(tk + q, 1u);
        got = __shfl_sync(0xffffffffu, got, 0);
        t = got + 1u;
        if (got >= n) return;                                   // the queue ran out in the meantime
        const uint32_t c = q == 1 ? (uint32_t)rank + got * (uint32_t)N : got;
        while (!chunk_ready(q, c)) { if (!keep_waiting()) { kind = FK_EXIT; return; } }
        kind = k; unit = (int)c;
      };
      while (kind < 0) {
        const bool hasG = tG < nG, hasX = tX < nX, hasS = tS < nS, hasU = tU < nU;
        if (!(hasG | hasX | hasS | hasU)) { kind = FK_EXIT; break; }
        const uint32_t cX = (uint32_t)rank + tX * (uint32_t)N;
        if (hasX && (link_cta || !hasG) && chunk_ready(1, cX)) { take(1, tX, nX, FK_X); continue; }
        if (hasG) {
          uint32_t gt = 0;
          if (lane == 0) gt = atomicAdd(tk + 0, 1u);
          gt = __shfl_sync(0xffffffffu, gt, 0);
          tG = gt + 1u;
          if (gt < nG) { kind = FK_G; unit = (int)gt; }
          continue;
        }
        if (hasS && chunk_ready(2, tS)) { take(2, tS, nS, FK_S); continue; }
        // (U only once every X unit of this rank has been taken: a CTA that waits for another rank's broadcast must
        //  never be the one that rank is waiting for)
        if (hasU && !hasX && chunk_ready(3, tU)) { take(3, tU, nU, FK_U); continue; }
        if (!keep_waiting()) { kind = FK_EXIT; break; }
      }
      if (kind == FK_X || kind == FK_S || kind == FK_U) __threadfence_system();   // acquire side of the relaxed polls
      if (lane == 0) { s_kind = kind; s_unit = unit; }
    }
    __syncthreads();
    const int kind = s_kind, c = s_unit;
    if (threadIdx.x == 0) lap(0);
    if (kind == FK_EXIT) break;
    const size_t s0 = (size_t)c * XCHG_CHUNK;
    const size_t s1 = min((size_t)x.P, s0 + XCHG_CHUNK);

    if (kind == FK_G) {
#pragma unroll 1
      for (int b = 0; b < XCHG_CHUNK / 256; b++) {
        const int blk = c * (XCHG_CHUNK / 256) + b;
        if (blk < nblk) geom_backward_block<D, VEC, W256, HAS_SH, true>(p, g, cam, blk);
        __syncthreads();
      }
      // every store of the chunk has been issued by this CTA: release the flag in every rank's buffer (own included)
      if (threadIdx.x == 0) lap(1);
      if (warp == 0 && lane < N) st_release_sys(x.peer_cflags[lane] + (size_t)rank * nch + c, epoch);
      if (warp == 0) __syncwarp();
      if (threadIdx.x == 0) lap(2);
    } else if (kind == FK_X) {
      const size_t base16 = s0 * NG4;
      const int n16 = (int)(s1 - s0) * NG4;
      constexpr int RD = 12;
#pragma unroll 1
      for (int j = threadIdx.x; j < n16; j += 256 * RD) {
        float4 v[RD];
#pragma unroll
        for (int u = 0; u < RD; u++) {
          const int q = j + u * 256;
          if (q < n16) {
            if (x.geo_mc) v[u] = mm_ld_reduce_add(x.geo_mc + (base16 + q) * 4);
            else {
              v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              for (int r = 0; r < N; r++) {
                const float4 t = ld_relaxed_sys_v4(x.peer_geo[r] + (base16 + q) * 4);
                v[u].x += t.x; v[u].y += t.y; v[u].z += t.z; v[u].w += t.w;
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < RD; u++) {
          const int q = j + u * 256;
          if (q < n16) {
            if (x.geo_mc) mm_st(x.geo_mc + (base16 + q) * 4, v[u]);
            else for (int r = 0; r < N; r++) *reinterpret_cast<float4*>(x.peer_geo[r] + (base16 + q) * 4) = v[u];
          }
        }
      }
      __syncthreads();
      if (warp == 0 && lane < N) st_release_sys(x.peer_cflags[lane] + (size_t)XCHG_MAX_RANKS * nch + c, epoch);
    } else if (kind == FK_S) {
      if (HAS_SH) {
#pragma unroll 1
        for (int k = 0; k < XCHG_CHUNK / 256; k++) {
          const size_t i = s0 + threadIdx.x + 256 * k;
          if (i < s1) sh_row_rebuild<D, W256>(i, x.gc_slot_floats, f.V, f.M, p.means3D, s_camv, x.gc, f.dL_dsh);
        }
      }
    } else {   // FK_U
#pragma unroll
      for (int k = 0; k < XCHG_CHUNK / 256; k++) {
        const size_t i = s0 + threadIdx.x + 256 * k;
        if (i < s1) {
          const float4* rec = reinterpret_cast<const float4*>(x.geo + i * (size_t)x.ngeo);
          const float4 a = __ldcg(rec), b = __ldcg(rec + 1), cc = __ldcg(rec + 2);
          const float4 d = HAS_SH ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldcg(rec + 3);
          fused_store_sums<HAS_SH>(f, i, a, b, cc, d);
        }
      }
    }
    if (threadIdx.x == 0) {
      xchg_mark(x.flags, kind);      // timeline slots 1..4: last G / X / S / U unit to finish
      if (kind != FK_G) lap(kind + 1);
    }
    __syncthreads();     // s_kind / s_unit are rewritten by warp 0
  }
  if (threadIdx.x == 0) {
    xchg_mark(x.flags, 5);
#pragma unroll
    for (int k = 0; k < 6; k++) atomicAdd(x.flags + 56 + k, (uint32_t)(s_cyc[k] >> 6));
  }
}

bool launch_geom_exchange_fused(const BwdParams& p, const GeomState& g, const FusedXchg& f, int max_ctas, cudaStream_t s) {
  if (p.P <= 0) return true;
  const bool has_sh = p.shs != nullptr;
  // the SH rows are READ as 128-bit words when they are 16-byte aligned multiples of 16 bytes (any active degree inside
  // M coefficients), as 256-bit words (and the rebuilt gradient rows written as such) when rows are exactly the active
  // coefficients of degree 3 and everything is 32-byte aligned; scalar accesses otherwise
  const bool vec = has_sh && ((p.M * 3) % 4 == 0) && ((reinterpret_cast<size_t>(p.shs) & 15) == 0);
  const bool wide = vec && ((reinterpret_cast<size_t>(p.shs) & 31) == 0) && ((reinterpret_cast<size_t>(f.dL_dsh) & 31) == 0) &&
                    p.D == 3 && p.M == 16;
  if (p.cov3D_precomp || f.x.world > XCHG_MAX_RANKS) return false;
  cudaMemsetAsync(f.x.flags + FLAG_TK, 0, 16 * sizeof(uint32_t), s);      // tickets, abort word, CTA-time counters
  cudaMemsetAsync(f.x.flags + FLAG_TL, 0, 6 * sizeof(unsigned long long), s);
  const int grid = max_ctas > 0 ? min(max_ctas, 2 * NUM_SMS_B200) : 2 * NUM_SMS_B200;
#define SFB_FX(DD, VV, WW, SS) geom_exchange_fused_kernel<DD, VV, WW, SS><<<grid, 256, 0, s>>>(p, g, f)
#define SFB_FD(DD) do { if (vec) SFB_FX(DD, true, false, true); else SFB_FX(DD, false, false, true); } while (0)
  if (!has_sh) SFB_FX(0, false, false, false);
  else switch (p.D) {
    case 0: SFB_FD(0); break;
    case 1: SFB_FD(1); break;
    case 2: SFB_FD(2); break;
    default: if (wide) SFB_FX(3, true, true, true); else SFB_FD(3); break;
  }
#undef SFB_FD
#undef SFB_FX
  return true;
}

void launch_geom_backward(const BwdParams& p, const GeomState& g, cudaStream_t s) {
  if (p.P <= 0) return;
  const int blocks = (p.P + 255) / 256;
  const bool vec = p.shs && ((p.M * 3) % 4 == 0) && ((reinterpret_cast<size_t>(p.shs) & 15) == 0) &&
                   ((reinterpret_cast<size_t>(p.dL_dsh) & 15) == 0);
#define SFB_GB(DD)                                                                                         \
  if (p.x_geo) {     /* exchange over NVLink: packed records + pushed colour gradients */                 \
    if (p.shs && vec && p.wide256 && p.M == (DD + 1) * (DD + 1) && (3 * (DD + 1) * (DD + 1)) % 8 == 0)    \
      geom_backward_kernel<DD, true, 2, true, true, true><<<blocks, 256, 0, s>>>(p, g);                    \
    else if (p.shs && vec) geom_backward_kernel<DD, true, 1, false, true, true><<<blocks, 256, 0, s>>>(p, g); \
    else if (p.shs) geom_backward_kernel<DD, false, 1, false, true, true><<<blocks, 256, 0, s>>>(p, g);    \
    else geom_backward_kernel<DD, false, 1, false, false, true><<<blocks, 256, 0, s>>>(p, g);              \
  } else if (p.sh_factored) {                                                                                     \
    if (vec && p.wide256 && p.M == (DD + 1) * (DD + 1) && (3 * (DD + 1) * (DD + 1)) % 8 == 0)             \
      geom_backward_kernel<DD, true, 2, true, true><<<blocks, 256, 0, s>>>(p, g);                          \
    else if (vec) geom_backward_kernel<DD, true, 1, false, true><<<blocks, 256, 0, s>>>(p, g);             \
    else geom_backward_kernel<DD, false, 1, false, true><<<blocks, 256, 0, s>>>(p, g);                     \
  } else if (vec && p.wide256 && p.M == (DD + 1) * (DD + 1) && (3 * (DD + 1) * (DD + 1)) % 8 == 0) {       \
    geom_backward_kernel<DD, true, 2, true><<<blocks, 256, 0, s>>>(p, g);                                  \
  } else if (vec) geom_backward_kernel<DD, true><<<blocks, 256, 0, s>>>(p, g);                             \
  else            geom_backward_kernel<DD, false><<<blocks, 256, 0, s>>>(p, g);
  switch (p.shs ? p.D : 0) {
    case 0: SFB_GB(0) break;
    case 1: SFB_GB(1) break;
    case 2: SFB_GB(2) break;
    default: SFB_GB(3) break;
  }
#undef SFB_GB
}

}  // namespace sfb
