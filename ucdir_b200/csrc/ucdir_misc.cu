// Bandwidth-bound kernels around the UNet: guidance branch, timestep embedding, tile gather,
// tile scatter fused with the posterior step.  Plain coalesced / vectorised SIMT.
#include "common.cuh"

namespace ucdir {

// ------------------------------------------------------------------------------------------------
// Guidance map (step invariant): model/ucdir.py:133-135 without the attw factor.
//   g   = bilinear(guide, scale = W/GW, align_corners=False)     -> mean of the 2x2 centre taps (r >= 2)
//   u   = conv1x1(g) (3 -> 16) ; gate = u[0:8] * u[8:16]          (SimpleGate, ucdir.py:149-152)
//   out = conv3x3(gate) (8 -> 8, zero padding of the gate map)
// One thread per output pixel; gate values of the 3x3 neighbourhood are recomputed (cheap, runs once
// per image per block).  GUIDE is NHWC with 4 floats per pixel (3 used).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) guidance_kernel(const float4* __restrict__ guide, const float* __restrict__ w0,
                                                       const float* __restrict__ b0, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, float* __restrict__ dst, int B,
                                                       int GH, int GW, int H, int W) {
  __shared__ float sw0[16 * 3], sb0[16], sw2[8 * 8 * 9], sb2[8];
  for (int i = threadIdx.x; i < 48; i += blockDim.x) sw0[i] = w0[i];
  for (int i = threadIdx.x; i < 16; i += blockDim.x) sb0[i] = b0[i];
  for (int i = threadIdx.x; i < 576; i += blockDim.x) sw2[i] = w2[i];   // OIHW [o][i][ky][kx]
  for (int i = threadIdx.x; i < 8; i += blockDim.x) sb2[i] = b2[i];
  __syncthreads();
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)B * H * W;
  if (idx >= total) return;
  int x = idx % W; size_t t = idx / W;
  int y = t % H; int b = t / H;
  const int r = GW / W;
  const int off = (r >> 1) - 1;
  float out[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) out[o] = sb2[o];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      int yy = y + ky - 1, xx = x + kx - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      float g0, g1, g2;
      if (r == 1) {
        float4 v = __ldg(guide + ((size_t)b * GH + yy) * GW + xx);
        g0 = v.x; g1 = v.y; g2 = v.z;
      } else {
        const float4* base = guide + ((size_t)b * GH + yy * r + off) * GW + xx * r + off;
        float4 p00 = __ldg(base), p01 = __ldg(base + 1), p10 = __ldg(base + GW), p11 = __ldg(base + GW + 1);
        g0 = 0.5f * (0.5f * p00.x + 0.5f * p01.x) + 0.5f * (0.5f * p10.x + 0.5f * p11.x);
        g1 = 0.5f * (0.5f * p00.y + 0.5f * p01.y) + 0.5f * (0.5f * p10.y + 0.5f * p11.y);
        g2 = 0.5f * (0.5f * p00.z + 0.5f * p01.z) + 0.5f * (0.5f * p10.z + 0.5f * p11.z);
      }
      float gate[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float u1 = sb0[i] + sw0[i * 3] * g0 + sw0[i * 3 + 1] * g1 + sw0[i * 3 + 2] * g2;
        float u2 = sb0[i + 8] + sw0[(i + 8) * 3] * g0 + sw0[(i + 8) * 3 + 1] * g1 + sw0[(i + 8) * 3 + 2] * g2;
        gate[i] = u1 * u2;
      }
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        float a = out[o];
#pragma unroll
        for (int i = 0; i < 8; ++i) a = fmaf(sw2[(o * 8 + i) * 9 + ky * 3 + kx], gate[i], a);
        out[o] = a;
      }
    }
  }
  float4* d = reinterpret_cast<float4*>(dst + idx * 8);
  d[0] = make_float4(out[0], out[1], out[2], out[3]);
  d[1] = make_float4(out[4], out[5], out[6], out[7]);
}

int launch_guidance(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  int B = op.i[UCDIR_GUID_I_B], GH = op.i[UCDIR_GUID_I_GH], GW = op.i[UCDIR_GUID_I_GW], H = op.i[UCDIR_GUID_I_H], W = op.i[UCDIR_GUID_I_W];
  for (int k = 0; k <= UCDIR_GUID_P_DST; ++k) if (!op.p[k]) { set_error("guidance: null pointer %d", k); return -1; }
  if (B <= 0 || H <= 0 || W <= 0 || GW % W || GH % H || GW / W != GH / H) { set_error("guidance: bad dims"); return -1; }
  int r = GW / W;
  if (r != 1 && (r & 1)) { set_error("guidance: ratio %d must be 1 or even", r); return -2; }
  if (dry) return 0;
  size_t total = (size_t)B * H * W;
  guidance_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>((const float4*)op.p[UCDIR_GUID_P_GUIDE],
      (const float*)op.p[UCDIR_GUID_P_W0], (const float*)op.p[UCDIR_GUID_P_B0], (const float*)op.p[UCDIR_GUID_P_W2],
      (const float*)op.p[UCDIR_GUID_P_B2], (float*)op.p[UCDIR_GUID_P_DST], B, GH, GW, H, W);
  ++g_launches;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Timestep embedding -> per-block mixing weights attw[L][NBLK][8]   (model/ucdir.py:24-29,212-214,106,125)
// One CTA per level.  INNER <= 128.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) time_embed_kernel(const float* __restrict__ levels, float level_scalar,
                                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                                         const float* __restrict__ blk, float* __restrict__ dst,
                                                         float* __restrict__ temb_out, int nblk, int inner) {
  __shared__ float enc[128], h1[512], t[128], u[8 * 64];
  const int l = blockIdx.x, tid = threadIdx.x;
  const float level = levels ? levels[l] : level_scalar;
  const int count = inner / 2;
  if (tid < inner) {
    int k = tid < count ? tid : tid - count;
    float step = (float)k / (float)count;
    float e = level * expf(-9.210340371976184f * step);   // -math.log(1e4) rounded to fp32 by the tensor multiply
    enc[tid] = tid < count ? sinf(e) : cosf(e);
  }
  __syncthreads();
  const int hid = inner * 4;
  for (int j = tid; j < hid; j += 256) {
    float a = 0.f;
    for (int k = 0; k < inner; ++k) a = fmaf(w1[j * inner + k], enc[k], a);
    h1[j] = swish_f(a + b1[j]);
  }
  __syncthreads();
  for (int j = tid; j < inner; j += 256) {
    float a = 0.f;
    for (int k = 0; k < hid; ++k) a = fmaf(w2[j * hid + k], h1[k], a);
    t[j] = a + b2[j];
    if (temb_out) temb_out[(size_t)l * inner + j] = t[j];
  }
  __syncthreads();
  const int rec = 8 * inner + 8 + 64 + 8;
  for (int base = 0; base < nblk; base += 64) {
    int nb = min(64, nblk - base);
    for (int j = tid; j < nb * 8; j += 256) {
      const float* r = blk + (size_t)(base + j / 8) * rec;
      int o = j & 7;
      float a = 0.f;
      for (int k = 0; k < inner; ++k) a = fmaf(r[o * inner + k], t[k], a);
      u[j] = swish_f(a + r[8 * inner + o]);
    }
    __syncthreads();
    for (int j = tid; j < nb * 8; j += 256) {
      const float* r = blk + (size_t)(base + j / 8) * rec + 8 * inner + 8;
      int o = j & 7;
      float a = 0.f;
      for (int k = 0; k < 8; ++k) a = fmaf(r[o * 8 + k], u[(j & ~7) + k], a);
      dst[((size_t)l * nblk + base) * 8 + j] = a + r[64 + o];
    }
    __syncthreads();
  }
}

int launch_time_embed(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  int L = op.i[UCDIR_TEMB_I_L], nblk = op.i[UCDIR_TEMB_I_NBLK], inner = op.i[UCDIR_TEMB_I_INNER];
  for (int k = UCDIR_TEMB_P_W1; k <= UCDIR_TEMB_P_DST; ++k) if (!op.p[k]) { set_error("time_embed: null pointer %d", k); return -1; }
  if (L <= 0 || nblk <= 0 || inner <= 0 || inner > 128 || (inner & 1)) { set_error("time_embed: bad dims"); return -1; }
  if (dry) return 0;
  time_embed_kernel<<<L, 256, 0, st>>>((const float*)op.p[UCDIR_TEMB_P_LEVELS], op.f[UCDIR_TEMB_F_LEVEL],
      (const float*)op.p[UCDIR_TEMB_P_W1], (const float*)op.p[UCDIR_TEMB_P_B1], (const float*)op.p[UCDIR_TEMB_P_W2],
      (const float*)op.p[UCDIR_TEMB_P_B2], (const float*)op.p[UCDIR_TEMB_P_BLK], (float*)op.p[UCDIR_TEMB_P_DST],
      (float*)op.p[UCDIR_TEMB_P_TEMB_OUT], nblk, inner);
  ++g_launches;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Tile gather with on-the-fly reflect padding (utils/util.py:117-137, model/ucdir.py:303-306,
// model/diffusion.py:166).  One thread per destination pixel; NCHW fp32 sources, NHWC destination.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_idx(int q, int n) {
  if (q < 0) q = -q;
  if (q >= n) q = 2 * (n - 1) - q;
  return q;
}

template <int BF16>
__global__ void __launch_bounds__(256) gather_tiles_kernel(const float* __restrict__ srcA, const float* __restrict__ srcB,
                                                           const int* __restrict__ tab, void* __restrict__ dstv, int BT,
                                                           int TH, int TW, int IH, int IW, int PD, int CA, int CB, int CD) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)BT * TH * TW;
  if (idx >= total) return;
  int x = idx % TW; size_t t = idx / TW;
  int y = t % TH; int tile = t / TH;
  int img = tab[tile * 3], y0 = tab[tile * 3 + 1], x0 = tab[tile * 3 + 2];
  int sy = reflect_idx(y0 + y - PD, IH), sx = reflect_idx(x0 + x - PD, IW);
  size_t plane = (size_t)IH * IW;
  size_t sp = (size_t)sy * IW + sx;
  float v[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) v[c] = 0.f;
  for (int c = 0; c < CA; ++c) v[c] = __ldg(srcA + ((size_t)img * CA + c) * plane + sp);
  for (int c = 0; c < CB; ++c) v[CA + c] = __ldg(srcB + ((size_t)img * CB + c) * plane + sp);
  if (BF16 == 2) {               // (hi, lo) plane pairs for the fp32-tolerance tensor-core mode: [hi: CD | lo: CD] per pixel
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dstv) + idx * 2 * CD;
    if (CD == 16) {               // 64 bytes per pixel: four 16-byte stores
      __align__(16) __nv_bfloat162 oh[8];
      __align__(16) __nv_bfloat162 ol[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        oh[c] = __floats2bfloat162_rn(v[2 * c], v[2 * c + 1]);
        const float2 r = __bfloat1622float2(oh[c]);
        ol[c] = __floats2bfloat162_rn(v[2 * c] - r.x, v[2 * c + 1] - r.y);
      }
      reinterpret_cast<uint4*>(d)[0] = reinterpret_cast<const uint4*>(oh)[0];
      reinterpret_cast<uint4*>(d)[1] = reinterpret_cast<const uint4*>(oh)[1];
      reinterpret_cast<uint4*>(d)[2] = reinterpret_cast<const uint4*>(ol)[0];
      reinterpret_cast<uint4*>(d)[3] = reinterpret_cast<const uint4*>(ol)[1];
    } else {
      for (int c = 0; c < CD; ++c) {
        const __nv_bfloat16 hi = __float2bfloat16(v[c]);
        d[c] = hi;
        d[CD + c] = __float2bfloat16(v[c] - __bfloat162float(hi));
      }
    }
  } else if (BF16) {
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dstv) + idx * CD;
    if (CD == 16) {               // the tensor-core path's 16-channel pixel rows: two 16-byte stores instead of 16 two-byte ones
      __align__(16) __nv_bfloat162 o[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = __floats2bfloat162_rn(v[2 * c], v[2 * c + 1]);
      reinterpret_cast<uint4*>(d)[0] = reinterpret_cast<const uint4*>(o)[0];
      reinterpret_cast<uint4*>(d)[1] = reinterpret_cast<const uint4*>(o)[1];
    } else {
      for (int c = 0; c < CD; ++c) d[c] = __float2bfloat16(v[c]);
    }
  } else {
    float* d = reinterpret_cast<float*>(dstv) + idx * CD;
    if (CD % 4 == 0) {
      for (int c = 0; c < CD; c += 4) *reinterpret_cast<float4*>(d + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
    } else {
      for (int c = 0; c < CD; ++c) d[c] = v[c];
    }
  }
}

int launch_gather_tiles(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  int BT = op.i[UCDIR_GATHER_I_BT], TH = op.i[UCDIR_GATHER_I_TH], TW = op.i[UCDIR_GATHER_I_TW];
  int IH = op.i[UCDIR_GATHER_I_IMG_H], IW = op.i[UCDIR_GATHER_I_IMG_W], PD = op.i[UCDIR_GATHER_I_PD];
  int CA = op.i[UCDIR_GATHER_I_CA], CB = op.i[UCDIR_GATHER_I_CB], CD = op.i[UCDIR_GATHER_I_CD];
  if (!op.p[UCDIR_GATHER_P_SRC_A] || !op.p[UCDIR_GATHER_P_TAB] || !op.p[UCDIR_GATHER_P_DST] || (CB > 0 && !op.p[UCDIR_GATHER_P_SRC_B])) {
    set_error("gather_tiles: null pointer"); return -1; }
  if (BT <= 0 || TH <= 0 || TW <= 0 || IH <= 1 || IW <= 1 || CA <= 0 || CB < 0 || CA + CB > CD || CD > 16 || PD < 0) {
    set_error("gather_tiles: bad dims"); return -1; }
  if (PD >= IH || PD >= IW) { set_error("gather_tiles: reflect pad %d >= image dim (F.pad would raise)", PD); return -2; }
  if (dry) return 0;
  size_t total = (size_t)BT * TH * TW;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (op.i[UCDIR_GATHER_I_OUT_BF16] == 2)
    gather_tiles_kernel<2><<<grid, 256, 0, st>>>((const float*)op.p[UCDIR_GATHER_P_SRC_A], (const float*)op.p[UCDIR_GATHER_P_SRC_B],
        (const int*)op.p[UCDIR_GATHER_P_TAB], op.p[UCDIR_GATHER_P_DST], BT, TH, TW, IH, IW, PD, CA, CB, CD);
  else if (op.i[UCDIR_GATHER_I_OUT_BF16])
    gather_tiles_kernel<1><<<grid, 256, 0, st>>>((const float*)op.p[UCDIR_GATHER_P_SRC_A], (const float*)op.p[UCDIR_GATHER_P_SRC_B],
        (const int*)op.p[UCDIR_GATHER_P_TAB], op.p[UCDIR_GATHER_P_DST], BT, TH, TW, IH, IW, PD, CA, CB, CD);
  else
    gather_tiles_kernel<0><<<grid, 256, 0, st>>>((const float*)op.p[UCDIR_GATHER_P_SRC_A], (const float*)op.p[UCDIR_GATHER_P_SRC_B],
        (const int*)op.p[UCDIR_GATHER_P_TAB], op.p[UCDIR_GATHER_P_DST], BT, TH, TW, IH, IW, PD, CA, CB, CD);
  ++g_launches;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Tile scatter (+ posterior step).  One thread per image pixel; each pixel reads its owner tile.
//   utils/util.py:144-146 (interior write-back, later windows win, crop) and
//   model/diffusion.py:150-158,171-172,182-183 with separately rounded products as in eager PyTorch.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scatter_kernel(const float* __restrict__ eps, const int* __restrict__ ownY,
                                                      const int* __restrict__ ownX, const int* __restrict__ y0s,
                                                      const int* __restrict__ x0s, const float* __restrict__ xt,
                                                      const float* __restrict__ noise, float* __restrict__ out, int BI,
                                                      int IH, int IW, int NTY, int NTX, int TH, int TW, int PD, int CE, int C,
                                                      int mode, int clip, float ca, float cb, float c1, float c2, float sigma,
                                                      float c3, const float* __restrict__ params) {
  if (params) {   // per-step scalars resident on the device (CUDA-graph replay): {A, B, C1, C2, SIGMA, clip, use_noise, C3}
    ca = params[0]; cb = params[1]; c1 = params[2]; c2 = params[3]; sigma = params[4];
    clip = params[5] != 0.f;
    if (params[6] == 0.f) noise = nullptr;
    c3 = params[7];
  }
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t plane = (size_t)IH * IW;
  size_t total = (size_t)BI * plane;
  if (idx >= total) return;
  int x = idx % IW; size_t t = idx / IW;
  int y = t % IH; int img = t / IH;
  int ty = ownY[y], tx = ownX[x];
  float e[4] = {0.f, 0.f, 0.f, 0.f};
  if (ty >= 0 && tx >= 0) {
    size_t tile = ((size_t)img * NTY + ty) * NTX + tx;
    int py = y + PD - y0s[ty], px = x + PD - x0s[tx];
    const float* s = eps + ((tile * TH + py) * TW + px) * CE;
    for (int c = 0; c < C; ++c) e[c] = __ldg(s + c);
  }
  size_t o = (size_t)img * C * plane + (size_t)y * IW + x;
  for (int c = 0; c < C; ++c) {
    float r;
    if (mode == 0) {
      r = e[c];
    } else {
      float xv = xt[o + c * plane];
      float x0 = __fsub_rn(__fmul_rn(ca, xv), __fmul_rn(cb, e[c]));
      if (clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
      float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xv));
      if (c3 != 0.f) mean = __fadd_rn(mean, __fmul_rn(c3, e[c]));      // DDIM: + c * pred_noise (diffusion.py:287)
      float z = noise ? noise[o + c * plane] : 0.f;
      r = __fadd_rn(mean, __fmul_rn(z, sigma));
    }
    out[o + c * plane] = r;
  }
}

int launch_scatter(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  int BI = op.i[UCDIR_SCATTER_I_BIMG], IH = op.i[UCDIR_SCATTER_I_IMG_H], IW = op.i[UCDIR_SCATTER_I_IMG_W];
  int C = op.i[UCDIR_SCATTER_I_C], CE = op.i[UCDIR_SCATTER_I_CE], mode = op.i[UCDIR_SCATTER_I_MODE];
  for (int k = UCDIR_SCATTER_P_EPS; k <= UCDIR_SCATTER_P_X0; ++k) if (!op.p[k]) { set_error("scatter: null pointer %d", k); return -1; }
  if (!op.p[UCDIR_SCATTER_P_OUT] || (mode == 1 && !op.p[UCDIR_SCATTER_P_XT])) { set_error("scatter: null out/xt"); return -1; }
  if (BI <= 0 || IH <= 0 || IW <= 0 || C <= 0 || C > 4 || CE < C) { set_error("scatter: bad dims"); return -1; }
  if (dry) return 0;
  size_t total = (size_t)BI * IH * IW;
  scatter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float*)op.p[UCDIR_SCATTER_P_EPS],
      (const int*)op.p[UCDIR_SCATTER_P_OWNER_Y], (const int*)op.p[UCDIR_SCATTER_P_OWNER_X], (const int*)op.p[UCDIR_SCATTER_P_Y0],
      (const int*)op.p[UCDIR_SCATTER_P_X0], (const float*)op.p[UCDIR_SCATTER_P_XT], (const float*)op.p[UCDIR_SCATTER_P_NOISE],
      (float*)op.p[UCDIR_SCATTER_P_OUT], BI, IH, IW, op.i[UCDIR_SCATTER_I_NTY], op.i[UCDIR_SCATTER_I_NTX],
      op.i[UCDIR_SCATTER_I_TH], op.i[UCDIR_SCATTER_I_TW], op.i[UCDIR_SCATTER_I_PD], CE, C, mode, op.i[UCDIR_SCATTER_I_CLIP],
      op.f[UCDIR_SCATTER_F_A], op.f[UCDIR_SCATTER_F_B], op.f[UCDIR_SCATTER_F_C1], op.f[UCDIR_SCATTER_F_C2], op.f[UCDIR_SCATTER_F_SIGMA],
      op.f[UCDIR_SCATTER_F_C3], (const float*)op.p[UCDIR_SCATTER_P_PARAMS]);
  ++g_launches;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Tile interior crop: DST[BT, IH, IW, 4] = SRC[BT, TH, TW, 4][:, OY:OY+IH, OX:OX+IW]  (fp32, float4 per pixel).
// Only the interiors of the tiles are ever stitched (utils/util.py:144-145), so only they travel in the
// per-step all-gather of the tile-sharded mode.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) crop_tiles_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int BT, int TH,
                                                         int TW, int IH, int IW, int OY, int OX) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)BT * IH * IW;
  if (idx >= total) return;
  int x = idx % IW; size_t t = idx / IW;
  int y = t % IH; int b = t / IH;
  dst[idx] = __ldg(src + ((size_t)b * TH + y + OY) * TW + x + OX);
}

int launch_crop_tiles(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  int BT = op.i[UCDIR_CROP_I_BT], TH = op.i[UCDIR_CROP_I_TH], TW = op.i[UCDIR_CROP_I_TW], IH = op.i[UCDIR_CROP_I_IH],
      IW = op.i[UCDIR_CROP_I_IW], OY = op.i[UCDIR_CROP_I_OY], OX = op.i[UCDIR_CROP_I_OX];
  if (!op.p[UCDIR_CROP_P_SRC] || !op.p[UCDIR_CROP_P_DST]) { set_error("crop_tiles: null pointer"); return -1; }
  if (BT <= 0 || IH <= 0 || IW <= 0 || OY < 0 || OX < 0 || OY + IH > TH || OX + IW > TW) { set_error("crop_tiles: bad dims"); return -1; }
  if (dry) return 0;
  size_t total = (size_t)BT * IH * IW;
  crop_tiles_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float4*)op.p[UCDIR_CROP_P_SRC], (float4*)op.p[UCDIR_CROP_P_DST],
                                                                   BT, TH, TW, IH, IW, OY, OX);
  ++g_launches;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Result image: crop + clamp + rescale + HWC uint8 in one pass (model/model.py:137 crop, core/metrics.py:8-34 tensor2img).
// Arithmetic order of the reference: (clamp(x) - min) / (max - min) in fp32, * 255.0f in fp32, round half to even.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) to_image_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int B, int C, int H, int W,
                                                          int PD, float lo, float hi) {
  const int OH = H - 2 * PD, OW = W - 2 * PD;
  const size_t total = (size_t)B * OH * OW;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % OW); size_t t = idx / OW;
  const int y = (int)(t % OH); const int b = (int)(t / OH);
  const size_t plane = (size_t)H * W;
  const float* s = src + (size_t)b * C * plane + (size_t)(y + PD) * W + (x + PD);
  uint8_t* d = dst + idx * C;
  const float range = __fsub_rn(hi, lo);
  for (int c = 0; c < C; ++c) {
    float v = fminf(fmaxf(__ldg(s + c * plane), lo), hi);
    v = __fdiv_rn(__fsub_rn(v, lo), range);
    d[c] = (uint8_t)rintf(__fmul_rn(v, 255.0f));
  }
}

int launch_to_image_u8(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  const int B = op.i[UCDIR_IMG_I_B], C = op.i[UCDIR_IMG_I_C], H = op.i[UCDIR_IMG_I_H], W = op.i[UCDIR_IMG_I_W], PD = op.i[UCDIR_IMG_I_PD];
  if (!op.p[UCDIR_IMG_P_SRC] || !op.p[UCDIR_IMG_P_DST]) { set_error("to_image_u8: null pointer"); return -1; }
  if (B <= 0 || C <= 0 || C > 4 || H <= 0 || W <= 0 || PD < 0 || 2 * PD >= H || 2 * PD >= W) { set_error("to_image_u8: bad dims"); return -1; }
  if (!(op.f[UCDIR_IMG_F_MAX] > op.f[UCDIR_IMG_F_MIN])) { set_error("to_image_u8: MAX must exceed MIN"); return -1; }
  if (dry) return 0;
  const size_t total = (size_t)B * (H - 2 * PD) * (W - 2 * PD);
  to_image_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float*)op.p[UCDIR_IMG_P_SRC], (uint8_t*)op.p[UCDIR_IMG_P_DST], B, C, H, W,
                                                                    PD, op.f[UCDIR_IMG_F_MIN], op.f[UCDIR_IMG_F_MAX]);
  ++g_launches;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm with G > 1 groups (the SR3-style FiLM ResnetBlock, model/ucdir.py:75-100; the DY3h path uses G = 1 and
// folds its statistics into the convolution kernels instead).  fp32 NHWC.
//   gn_stats_f32:  STATS[b][g] = {sum, sum of squares} over H*W x C/G elements, one CTA per (sample, group)
//   gn_apply_f32:  DST = [Swish](GroupNorm(G, C)(SRC)),  eps as torch.nn.GroupNorm
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_stats_f32_kernel(const float* __restrict__ src, double* __restrict__ stats, int HW, int C, int G) {
  __shared__ double red[2][8];
  const int b = blockIdx.y, g = blockIdx.x, cpg = C / G;
  const float* base = src + (size_t)b * HW * C + g * cpg;
  double s1 = 0, s2 = 0;
  for (int i = threadIdx.x; i < HW * cpg; i += 256) {
    const int pix = i / cpg, c = i - pix * cpg;
    const double v = base[(size_t)pix * C + c];
    s1 += v; s2 += v * v;
  }
  s1 = warp_sum_d(s1); s2 = warp_sum_d(s2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t1 = 0, t2 = 0;
    for (int w = 0; w < 8; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
    stats[((size_t)b * G + g) * 2] = t1; stats[((size_t)b * G + g) * 2 + 1] = t2;
  }
}

__global__ void __launch_bounds__(256) gn_apply_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const double* __restrict__ stats, int HW, int C, int G,
                                                           float eps, int swish) {
  __shared__ float smean[64], srstd[64];
  const int b = blockIdx.y, cpg = C / G;
  if (threadIdx.x < G) {
    const double cnt = (double)HW * cpg;
    const double mean = stats[((size_t)b * G + threadIdx.x) * 2] / cnt;
    double var = stats[((size_t)b * G + threadIdx.x) * 2 + 1] / cnt - mean * mean;
    if (var < 0) var = 0;
    smean[threadIdx.x] = (float)mean; srstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const size_t per = (size_t)HW * C;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < per; i += (size_t)gridDim.x * 256) {
    const int c = (int)(i % C), g = c / cpg;
    float v = (src[(size_t)b * per + i] - smean[g]) * srstd[g] * gamma[c] + beta[c];
    if (swish) v = swish_f(v);
    dst[(size_t)b * per + i] = v;
  }
}

int launch_gn_stats_f32(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  const int B = op.i[UCDIR_GNS_I_B], HW = op.i[UCDIR_GNS_I_HW], C = op.i[UCDIR_GNS_I_C], G = op.i[UCDIR_GNS_I_G];
  if (!op.p[UCDIR_GNS_P_SRC] || !op.p[UCDIR_GNS_P_STATS]) { set_error("gn_stats: null pointer"); return -1; }
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G || B > 65535) { set_error("gn_stats: bad dims"); return -1; }
  if (dry) return 0;
  gn_stats_f32_kernel<<<dim3(G, B), 256, 0, st>>>((const float*)op.p[UCDIR_GNS_P_SRC], (double*)op.p[UCDIR_GNS_P_STATS], HW, C, G);
  ++g_launches;
  return 0;
}

int launch_gn_apply_f32(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  const int B = op.i[UCDIR_GNS_I_B], HW = op.i[UCDIR_GNS_I_HW], C = op.i[UCDIR_GNS_I_C], G = op.i[UCDIR_GNS_I_G];
  for (int k = 0; k <= UCDIR_GNF_P_STATS; ++k) if (!op.p[k]) { set_error("gn_apply_f32: null pointer %d", k); return -1; }
  if (B <= 0 || HW <= 0 || C <= 0 || G <= 0 || G > 64 || C % G || B > 65535) { set_error("gn_apply_f32: bad dims (G <= 64)"); return -1; }
  if (dry) return 0;
  const size_t per = (size_t)HW * C;
  unsigned gx = (unsigned)((per + 255) / 256); if (gx > 1184) gx = 1184;
  gn_apply_f32_kernel<<<dim3(gx, B), 256, 0, st>>>((const float*)op.p[UCDIR_GNF_P_SRC], (float*)op.p[UCDIR_GNF_P_DST], (const float*)op.p[UCDIR_GNF_P_GAMMA],
      (const float*)op.p[UCDIR_GNF_P_BETA], (const double*)op.p[UCDIR_GNF_P_STATS], HW, C, G, op.f[0], op.i[UCDIR_GNS_I_SWISH]);
  ++g_launches;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Layout change at a module boundary: fp32 NCHW <-> NHWC (DIR 0: NCHW -> NHWC, 1: NHWC -> NCHW).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layout_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int HW, int dir) {
  const size_t total = (size_t)B * C * HW;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    // i indexes the NHWC tensor
    const int c = (int)(i % C); const size_t t = i / C; const int pix = (int)(t % HW); const int b = (int)(t / HW);
    const size_t j = ((size_t)b * C + c) * HW + pix;      // NCHW index
    if (dir == 0) dst[i] = src[j]; else dst[j] = src[i];
  }
}

int launch_layout(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  const int B = op.i[0], C = op.i[1], HW = op.i[2], dir = op.i[3];
  if (!op.p[0] || !op.p[1] || B <= 0 || C <= 0 || HW <= 0) { set_error("layout: bad args"); return -1; }
  if (dry) return 0;
  const size_t total = (size_t)B * C * HW;
  unsigned g = (unsigned)((total + 255) / 256); if (g > 4736) g = 4736;
  layout_kernel<<<g, 256, 0, st>>>((const float*)op.p[0], (float*)op.p[1], B, C, HW, dir);
  ++g_launches;
  return 0;
}

}  // namespace ucdir
