// fp32 "parity mode" kernels: SIMT implicit-GEMM convolution with fused GroupNorm prologue and
// bias/activation/residual/integration-mix/statistics epilogue, batched SGEMM, row softmax, max-pool.
// NHWC fp32 activations.  These are the exact-semantics path (rtol 1e-3 / atol 1e-4 against the CPU
// oracle); the tcgen05 bf16 path lives in ucdir_tc.cu.
#include "common.cuh"

namespace ucdir {

struct ConvP {
  const float* src0; const float* src1; const float* w; const float* bias;
  const float* gamma; const float* beta; const double* stats0; const double* stats1;
  const float* res; const float* att; const float* attw; float* dst; double* dst_stats;
  const float* film_g; const float* film_b;
  int B, H, W, C0, C1, Cout, ks, stride, up, groups, pre, act, mode, srcH, srcW;
  int dstC, dstCoff, dstUp, dstPy, dstPx, resC, attwStride, ldw, filmStride;
  float eps;
};

// One CTA: 128 output pixels of sample blockIdx.z  x  BN output channels; K loop in steps of 8 input
// channels of one filter tap.  2*BN threads, each an 8x8 register tile.
template <int BN>
__global__ void __launch_bounds__(2 * BN) conv_f32_kernel(const ConvP p) {
  constexpr int BM = 128, BK = 8, LDA = BM + 4, THREADS = 2 * BN, NTX = BN / 8;
  constexpr int A_LOADS = (BM * 2) / THREADS;
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][BN];
  __shared__ double red[2][THREADS / 32];

  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int HW = p.H * p.W;
  const int Cin = p.C0 + p.C1;
  const int Cg = Cin / p.groups;
  const int Ng = p.Cout / p.groups;
  const int g = n0 / Ng;
  const int nl0 = n0 - g * Ng;
  const int K = p.ks * p.ks * Cg;
  const int nsteps = K / BK;
  const float* __restrict__ wg = p.w + (size_t)g * K * p.ldw;
  const int pad = p.ks >> 1;
  const int VH = p.srcH << p.up, VW = p.srcW << p.up;

  float mean = 0.f, rstd = 1.f;
  if (p.pre) {
    GnScalars s = gn_scalars(p.stats0, p.C1 > 0 ? p.stats1 : nullptr, b, (double)Cin * p.srcH * p.srcW, p.eps);
    mean = s.mean; rstd = s.rstd;
  }

  // per-thread A-load slots
  int a_oy[A_LOADS], a_ox[A_LOADS], a_pm[A_LOADS], a_half[A_LOADS];
  bool a_valid[A_LOADS];
#pragma unroll
  for (int j = 0; j < A_LOADS; ++j) {
    int li = tid + j * THREADS;
    a_pm[j] = li >> 1; a_half[j] = li & 1;
    int m = m0 + a_pm[j];
    a_valid[j] = m < HW;
    int mm = a_valid[j] ? m : 0;
    a_oy[j] = mm / p.W; a_ox[j] = mm - a_oy[j] * p.W;
  }
  // B-load slot
  const int b_kk = tid / (BN / 4), b_c4 = tid % (BN / 4);
  const int b_pos = ((b_c4 & 1) ? BN / 2 : 0) + (b_c4 >> 1) * 4;
  const bool b_ok = (nl0 + b_c4 * 4) < p.ldw;

  float4 ra[A_LOADS]; float4 rb;

  auto load_step = [&](int s) {
    const int kbase = s * BK;
    const int tap = kbase / Cg;
    const int c0 = kbase - tap * Cg;
    const int dy = tap / p.ks, dx = tap - dy * p.ks;
#pragma unroll
    for (int j = 0; j < A_LOADS; ++j) {
      int iy = a_oy[j] * p.stride + dy - pad, ix = a_ox[j] * p.stride + dx - pad;
      bool inb = a_valid[j] && iy >= 0 && iy < VH && ix >= 0 && ix < VW;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (inb) {
        int sy = iy >> p.up, sx = ix >> p.up;
        int ci = g * Cg + c0 + a_half[j] * 4;
        size_t pix = ((size_t)b * p.srcH + sy) * p.srcW + sx;
        const float* ptr = (ci < p.C0) ? p.src0 + pix * p.C0 + ci : p.src1 + pix * p.C1 + (ci - p.C0);
        v = __ldg(reinterpret_cast<const float4*>(ptr));
        if (p.pre) {
          float4 ga = __ldg(reinterpret_cast<const float4*>(p.gamma + ci));
          float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + ci));
          float sc;
          sc = rstd * ga.x; v.x = v.x * sc + (be.x - sc * mean);
          sc = rstd * ga.y; v.y = v.y * sc + (be.y - sc * mean);
          sc = rstd * ga.z; v.z = v.z * sc + (be.z - sc * mean);
          sc = rstd * ga.w; v.w = v.w * sc + (be.w - sc * mean);
          if (p.pre == 2) { v.x = swish_f(v.x); v.y = swish_f(v.y); v.z = swish_f(v.z); v.w = swish_f(v.w); }
        }
      }
      ra[j] = v;
    }
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b_ok) rb = __ldg(reinterpret_cast<const float4*>(wg + (size_t)(kbase + b_kk) * p.ldw + nl0 + b_c4 * 4));
  };
  auto store_step = [&](int buf) {
#pragma unroll
    for (int j = 0; j < A_LOADS; ++j) {
      int kk = a_half[j] * 4;
      As[buf][kk + 0][a_pm[j]] = ra[j].x; As[buf][kk + 1][a_pm[j]] = ra[j].y;
      As[buf][kk + 2][a_pm[j]] = ra[j].z; As[buf][kk + 3][a_pm[j]] = ra[j].w;
    }
    *reinterpret_cast<float4*>(&Bs[buf][b_kk][b_pos]) = rb;
  };

  const int tx = tid % NTX, ty = tid / NTX;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_step(0);
  store_step(0);
  __syncthreads();
  for (int s = 0; s < nsteps; ++s) {
    const int buf = s & 1;
    if (s + 1 < nsteps) load_step(s + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][BN / 2 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (s + 1 < nsteps) store_step(buf ^ 1);
    __syncthreads();
  }

  // ---------------- epilogue ----------------
  float s1 = 0.f, s2 = 0.f;
  const int nb = n0 + tx * 8;  // first of this thread's 8 output columns (global n)
  float bias8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bias8[j] = (p.bias && nb + j < p.Cout) ? __ldg(p.bias + nb + j) : 0.f;

  if (p.mode == 1) {
    float aw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) aw[j] = __ldg(p.attw + (size_t)b * p.attwStride + j);
    const int c = nb >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m = m0 + ty * 8 + i;
      if (m < HW && nb < p.Cout) {
        size_t pix = (size_t)b * HW + m;
        float4 t0 = __ldg(reinterpret_cast<const float4*>(p.att + pix * 8));
        float4 t1 = __ldg(reinterpret_cast<const float4*>(p.att + pix * 8 + 4));
        float at[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
        float h = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) h += (acc[i][j] + bias8[j]) * (at[j] * aw[j]);
        float v = swish_f(h) + __ldg(p.res + pix * p.resC + c);
        p.dst[pix * p.dstC + p.dstCoff + c] = v;
        s1 += v; s2 += v * v;
      }
    }
  } else {
    const bool vec = (p.Cout % 8 == 0) && (p.dstC % 4 == 0) && (p.dstCoff % 4 == 0);
    float fg[8], fb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      fg[j] = 0.f; fb[j] = 0.f;
      if (p.film_b && nb + j < p.Cout) {
        fb[j] = __ldg(p.film_b + (size_t)b * p.filmStride + nb + j);
        if (p.film_g) fg[j] = __ldg(p.film_g + (size_t)b * p.filmStride + nb + j);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m = m0 + ty * 8 + i;
      if (m >= HW) continue;
      int oy = m / p.W, ox = m - oy * p.W;
      size_t pix_in = (size_t)b * HW + m;
      size_t pix_out = p.dstUp ? ((size_t)b * 2 * p.H + 2 * oy + p.dstPy) * (2 * p.W) + 2 * ox + p.dstPx : pix_in;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = acc[i][j] + bias8[j];
        if (p.film_b) t = (1.0f + fg[j]) * t + fb[j];
        if (p.act == 1) t = swish_f(t); else if (p.act == 2) t = lrelu_f(t);
        if (p.res && nb + j < p.Cout) t += __ldg(p.res + pix_in * p.resC + nb + j);
        v[j] = t;
        if (nb + j < p.Cout) { s1 += t; s2 += t * t; }
      }
      float* d = p.dst + pix_out * p.dstC + p.dstCoff + nb;
      if (vec) {
        if (nb < p.Cout) {
          *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (nb + j < p.Cout) d[j] = v[j];
      }
    }
  }

  if (p.dst_stats) {
    double d1 = warp_sum_d((double)s1), d2 = warp_sum_d((double)s2);
    if ((tid & 31) == 0) { red[0][tid >> 5] = d1; red[1][tid >> 5] = d2; }
    __syncthreads();
    if (tid == 0) {
      double t1 = 0, t2 = 0;
#pragma unroll
      for (int w = 0; w < THREADS / 32; ++w) { t1 += red[0][w]; t2 += red[1][w]; }
      atomicAdd(p.dst_stats + 2 * b, t1);
      atomicAdd(p.dst_stats + 2 * b + 1, t2);
    }
  }
}

int launch_conv_f32(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  ConvP p;
  p.src0 = (const float*)op.p[UCDIR_CONV_P_SRC0]; p.src1 = (const float*)op.p[UCDIR_CONV_P_SRC1];
  p.w = (const float*)op.p[UCDIR_CONV_P_W]; p.bias = (const float*)op.p[UCDIR_CONV_P_BIAS];
  p.gamma = (const float*)op.p[UCDIR_CONV_P_GAMMA]; p.beta = (const float*)op.p[UCDIR_CONV_P_BETA];
  p.stats0 = (const double*)op.p[UCDIR_CONV_P_STATS0]; p.stats1 = (const double*)op.p[UCDIR_CONV_P_STATS1];
  p.res = (const float*)op.p[UCDIR_CONV_P_RES]; p.att = (const float*)op.p[UCDIR_CONV_P_ATT];
  p.attw = (const float*)op.p[UCDIR_CONV_P_ATTW]; p.dst = (float*)op.p[UCDIR_CONV_P_DST];
  p.dst_stats = (double*)op.p[UCDIR_CONV_P_DST_STATS];
  p.film_g = (const float*)op.p[UCDIR_CONV_P_FILM_G]; p.film_b = (const float*)op.p[UCDIR_CONV_P_FILM_B];
  p.B = op.i[UCDIR_CONV_I_B]; p.H = op.i[UCDIR_CONV_I_H]; p.W = op.i[UCDIR_CONV_I_W];
  p.C0 = op.i[UCDIR_CONV_I_C0]; p.C1 = op.i[UCDIR_CONV_I_C1]; p.Cout = op.i[UCDIR_CONV_I_COUT];
  p.ks = op.i[UCDIR_CONV_I_KSIZE]; p.stride = op.i[UCDIR_CONV_I_STRIDE]; p.up = op.i[UCDIR_CONV_I_UP];
  p.groups = op.i[UCDIR_CONV_I_GROUPS]; p.pre = op.i[UCDIR_CONV_I_PRE]; p.act = op.i[UCDIR_CONV_I_ACT];
  p.mode = op.i[UCDIR_CONV_I_MODE]; p.srcH = op.i[UCDIR_CONV_I_SRC_H]; p.srcW = op.i[UCDIR_CONV_I_SRC_W];
  p.dstC = op.i[UCDIR_CONV_I_DST_C]; p.dstCoff = op.i[UCDIR_CONV_I_DST_COFF]; p.dstUp = op.i[UCDIR_CONV_I_DST_UP];
  p.dstPy = op.i[UCDIR_CONV_I_DST_PY]; p.dstPx = op.i[UCDIR_CONV_I_DST_PX]; p.resC = op.i[UCDIR_CONV_I_RES_C];
  p.attwStride = op.i[UCDIR_CONV_I_ATTW_STRIDE];
  p.filmStride = op.i[UCDIR_CONV_I_FILM_STRIDE] ? op.i[UCDIR_CONV_I_FILM_STRIDE] : op.i[UCDIR_CONV_I_COUT];
  p.eps = op.f[UCDIR_CONV_F_EPS];
  if (!p.src0 || !p.w || !p.dst) { set_error("conv_f32: null src0/w/dst"); return -1; }
  if (p.B <= 0 || p.H <= 0 || p.W <= 0 || p.Cout <= 0 || p.C0 <= 0) { set_error("conv_f32: bad dims"); return -1; }
  if (p.groups < 1 || (p.ks != 1 && p.ks != 3) || (p.stride != 1 && p.stride != 2) || (p.up != 0 && p.up != 1)) {
    set_error("conv_f32: unsupported ksize/stride/up/groups"); return -2; }
  const int Cin = p.C0 + p.C1;
  if (Cin % p.groups || p.Cout % p.groups || (Cin / p.groups) % 8) { set_error("conv_f32: C/groups must be a multiple of 8 (got Cin=%d groups=%d)", Cin, p.groups); return -2; }
  if (p.C1 > 0 && (!p.src1 || p.C0 % 8 || p.groups != 1)) { set_error("conv_f32: bad dual-source config"); return -2; }
  if (p.pre && (!p.gamma || !p.beta || !p.stats0 || (p.C1 > 0 && !p.stats1))) { set_error("conv_f32: GN prologue needs gamma/beta/stats"); return -1; }
  if (op.i[UCDIR_CONV_I_GN_GROUPS] > 1) { set_error("conv_f32: GroupNorm groups > 1 not supported in the fused prologue"); return -2; }
  const int Ng = p.Cout / p.groups;
  p.ldw = (Ng + 3) & ~3;
  if (p.mode == 1) {
    if (!p.att || !p.attw || !p.res || Ng % 8 || p.dstUp) { set_error("conv_f32: mix epilogue needs att/attw/res and Cout/groups %% 8 == 0"); return -1; }
  }
  const int BN = (Ng >= 128 && Ng % 128 == 0) ? 128 : 64;
  if (p.groups > 1 && Ng % BN) { set_error("conv_f32: grouped conv needs Cout/groups %% %d == 0", BN); return -2; }
  if (dry) return 0;
  dim3 grid((p.H * p.W + 127) / 128, (p.Cout + BN - 1) / BN, p.B);
  if (BN == 128) conv_f32_kernel<128><<<grid, 256, 0, st>>>(p);
  else conv_f32_kernel<64><<<grid, 128, 0, st>>>(p);
  ++g_launches;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Batched SGEMM: C[b] = alpha * A[b] (MxK, lda) * B[b]   (TRANSB: B is NxK row-major; else KxN row-major)
// 64x64 tile, BK=16, 256 threads, 4x4 per thread.  Used by the fp32 attention path.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

template <bool TRANSB, typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256) sgemm_f32_kernel(const TA* __restrict__ A, const TB* __restrict__ Bm,
                                                        TC* __restrict__ C, int M, int N, int K, int lda, int ldb,
                                                        int ldc, long long sa, long long sb, long long sc, float alpha) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  const int tid = threadIdx.x;
  const int bz = blockIdx.z;
  A += (size_t)bz * sa; Bm += (size_t)bz * sb; C += (size_t)bz * sc;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    {
      int r = tid >> 2, c = (tid & 3) * 4;
      int m = m0 + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int k = k0 + c + j;
        As[c + j][r] = (m < M && k < K) ? ldf(A + (size_t)m * lda + k) : 0.f;
      }
    }
    if (TRANSB) {
      int r = tid >> 2, c = (tid & 3) * 4;
      int n = n0 + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int k = k0 + c + j;
        Bs[c + j][r] = (n < N && k < K) ? ldf(Bm + (size_t)n * ldb + k) : 0.f;
      }
    } else {
      int r = tid >> 4, c = (tid & 15) * 4;   // 16 rows x 64 cols
      int k = k0 + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int n = n0 + c + j;
        Bs[r][c + j] = (k < K && n < N) ? ldf(Bm + (size_t)k * ldb + n) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) stf(C + (size_t)m * ldc + n, acc[i][j] * alpha);
    }
  }
}

int launch_sgemm_f32(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  const void* A = op.p[UCDIR_SGEMM_P_A]; const void* B = op.p[UCDIR_SGEMM_P_B];
  void* C = op.p[UCDIR_SGEMM_P_C];
  int batch = op.i[UCDIR_SGEMM_I_BATCH], M = op.i[UCDIR_SGEMM_I_M], N = op.i[UCDIR_SGEMM_I_N], K = op.i[UCDIR_SGEMM_I_K];
  if (!A || !B || !C || batch <= 0 || M <= 0 || N <= 0 || K <= 0) { set_error("sgemm_f32: bad args"); return -1; }
  if (batch > 65535) { set_error("sgemm_f32: batch > 65535"); return -2; }
  const int tb = op.i[UCDIR_SGEMM_I_TRANSB] ? 1 : 0;
  const int ty = (op.i[UCDIR_SGEMM_I_A_BF16] ? 4 : 0) | (op.i[UCDIR_SGEMM_I_B_BF16] ? 2 : 0) | (op.i[UCDIR_SGEMM_I_C_BF16] ? 1 : 0);
  // supported operand type combinations: fp32 x fp32 -> fp32 (parity path); bf16 x bf16 -> fp32 (Q K^T);
  // fp32 x bf16 -> bf16 (P V)
  if (!((ty == 0) || (ty == 6 && tb) || (ty == 3 && !tb))) { set_error("sgemm_f32: unsupported operand types %d transb %d", ty, tb); return -2; }
  if (dry) return 0;
  dim3 grid((N + 63) / 64, (M + 63) / 64, batch);
  const int lda = op.i[UCDIR_SGEMM_I_LDA], ldb = op.i[UCDIR_SGEMM_I_LDB], ldc = op.i[UCDIR_SGEMM_I_LDC];
  const long long sa = op.i[UCDIR_SGEMM_I_SA], sb = op.i[UCDIR_SGEMM_I_SB], sc = op.i[UCDIR_SGEMM_I_SC];
  const float alpha = op.f[UCDIR_SGEMM_F_ALPHA];
  if (ty == 0 && tb) sgemm_f32_kernel<true, float, float, float><<<grid, 256, 0, st>>>((const float*)A, (const float*)B, (float*)C, M, N, K, lda, ldb, ldc, sa, sb, sc, alpha);
  else if (ty == 0) sgemm_f32_kernel<false, float, float, float><<<grid, 256, 0, st>>>((const float*)A, (const float*)B, (float*)C, M, N, K, lda, ldb, ldc, sa, sb, sc, alpha);
  else if (ty == 6) sgemm_f32_kernel<true, __nv_bfloat16, __nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, (float*)C, M, N, K, lda, ldb, ldc, sa, sb, sc, alpha);
  else sgemm_f32_kernel<false, float, __nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)A, (const __nv_bfloat16*)B, (__nv_bfloat16*)C, M, N, K, lda, ldb, ldc, sa, sb, sc, alpha);
  ++g_launches;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Row softmax in place (model/ucdir.py:176).  One CTA per row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ X, int cols, int in_ld, __nv_bfloat16* __restrict__ out, int out_ld) {
  __shared__ float red[8];
  __shared__ float bc;
  float* row = X + (size_t)blockIdx.x * in_ld;
  const int tid = threadIdx.x;
  float mx = -INFINITY;
  for (int c = tid; c < cols; c += 256) mx = fmaxf(mx, row[c]);
  mx = warp_max_f(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  if (tid == 0) { float m = red[0]; for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]); bc = m; }
  __syncthreads();
  mx = bc;
  float sum = 0.f;
  for (int c = tid; c < cols; c += 256) { float e = expf(row[c] - mx); row[c] = e; sum += e; }
  sum = warp_sum_f(sum);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) { float s = 0.f; for (int w = 0; w < 8; ++w) s += red[w]; bc = s; }
  __syncthreads();
  const float tot = bc;
  if (out) {
    __nv_bfloat16* o = out + (size_t)blockIdx.x * out_ld;
    for (int c = tid; c < out_ld; c += 256) o[c] = __float2bfloat16(c < cols ? row[c] / tot : 0.f);
  } else {
    for (int c = tid; c < cols; c += 256) row[c] = row[c] / tot;
  }
}

// One-pass variant for long rows (attention over 1024x1024 tiles: 16384 keys): the row is read from HBM once into
// shared memory, reduced there, and written once (bf16 probabilities or fp32 in place) -- 1.5 instead of 5 passes.
__global__ void __launch_bounds__(512) softmax_rows_smem_kernel(float* __restrict__ X, int cols, int in_ld, __nv_bfloat16* __restrict__ out, int out_ld) {
  extern __shared__ float srow[];
  __shared__ float red[16];
  __shared__ float bc;
  float* row = X + (size_t)blockIdx.x * in_ld;
  const int tid = threadIdx.x;
  float mx = -INFINITY;
  for (int c = tid * 4; c < cols; c += 512 * 4) {
    if (c + 3 < cols && (in_ld & 3) == 0) {
      const float4 v = *reinterpret_cast<const float4*>(row + c);
      *reinterpret_cast<float4*>(srow + c) = v;
      mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    } else {
      for (int k = c; k < cols && k < c + 4; ++k) { const float v = row[k]; srow[k] = v; mx = fmaxf(mx, v); }
    }
  }
  mx = warp_max_f(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  if (tid == 0) { float m = red[0]; for (int w = 1; w < 16; ++w) m = fmaxf(m, red[w]); bc = m; }
  __syncthreads();
  mx = bc;
  float sum = 0.f;
  for (int c = tid; c < cols; c += 512) { const float e = expf(srow[c] - mx); srow[c] = e; sum += e; }
  sum = warp_sum_f(sum);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int w = 0; w < 16; ++w) t += red[w]; bc = t; }
  __syncthreads();
  const float tot = bc;
  if (out) {
    __nv_bfloat16* o = out + (size_t)blockIdx.x * out_ld;
    for (int c = tid * 2; c < out_ld; c += 512 * 2) {
      const float a = c < cols ? srow[c] / tot : 0.f, b = c + 1 < cols ? srow[c + 1] / tot : 0.f;
      if (c + 1 < out_ld && (out_ld & 1) == 0) *reinterpret_cast<__nv_bfloat162*>(o + c) = __floats2bfloat162_rn(a, b);
      else { o[c] = __float2bfloat16(a); if (c + 1 < out_ld) o[c + 1] = __float2bfloat16(b); }
    }
  } else {
    for (int c = tid; c < cols; c += 512) row[c] = srow[c] / tot;
  }
}

// Short rows (attention over <= 32x32-pixel tiles: <= 1024 keys): one WARP per row, the row lives in registers (VPT float4
// per lane), reductions are warp shuffles, one read and one write of HBM.  The CTA-per-row kernel above spent its time in
// block barriers: 50 us per launch for 31k rows of 256 floats, against 8 us of HBM time.
template <int VPT>
__global__ void __launch_bounds__(256) softmax_rows_warp_kernel(float* __restrict__ X, int rows, int cols, int in_ld,
                                                                __nv_bfloat16* __restrict__ out, int out_ld) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    float* row = X + (size_t)r * in_ld;
    float v[4 * VPT];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * 32 + lane) * 4;
      float4 t = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (c + 3 < cols) t = *reinterpret_cast<const float4*>(row + c);
      else {
        if (c < cols) t.x = row[c];
        if (c + 1 < cols) t.y = row[c + 1];
        if (c + 2 < cols) t.z = row[c + 2];
      }
      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
      mx = fmaxf(fmaxf(mx, fmaxf(t.x, t.y)), fmaxf(t.z, t.w));
    }
    mx = warp_max_f(mx);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 4 * VPT; ++k) { v[k] = expf(v[k] - mx); sum += v[k]; }       // exp(-inf) = 0 for the padding
    const float tot = warp_sum_f(sum);
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
      const int c = (j * 32 + lane) * 4;
      const float a = v[4 * j] / tot, b = v[4 * j + 1] / tot, cc = v[4 * j + 2] / tot, d = v[4 * j + 3] / tot;
      if (out) {
        if (c < out_ld) {                                   // out_ld is a multiple of 4 (checked by the launcher); columns >= cols are 0
          __align__(8) __nv_bfloat162 o[2] = {__floats2bfloat162_rn(a, b), __floats2bfloat162_rn(cc, d)};
          *reinterpret_cast<uint2*>(out + (size_t)r * out_ld + c) = *reinterpret_cast<const uint2*>(o);
        }
      } else {
        if (c + 3 < cols) *reinterpret_cast<float4*>(row + c) = make_float4(a, b, cc, d);
        else {
          if (c < cols) row[c] = a;
          if (c + 1 < cols) row[c + 1] = b;
          if (c + 2 < cols) row[c + 2] = cc;
        }
      }
    }
  }
}

int launch_softmax_f32(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  float* X = (float*)op.p[UCDIR_SOFTMAX_P_X];
  int rows = op.i[UCDIR_SOFTMAX_I_ROWS], cols = op.i[UCDIR_SOFTMAX_I_COLS];
  if (!X || rows <= 0 || cols <= 0) { set_error("softmax: bad args"); return -1; }
  if (op.p[UCDIR_SOFTMAX_P_OUT_BF16] && op.i[UCDIR_SOFTMAX_I_OUT_LD] < cols) { set_error("softmax: OUT_LD < COLS"); return -1; }
  if (dry) return 0;
  const int in_ld = op.i[UCDIR_SOFTMAX_I_IN_LD] ? op.i[UCDIR_SOFTMAX_I_IN_LD] : cols;
  if (cols >= 2048 && cols <= 49152) {
    static int attr_dev[UCDIR_MAX_DEV] = {};
    int& attr = attr_dev[cur_dev()];
    const int bytes = cols * 4;
    if (attr < bytes) {
      if (cudaFuncSetAttribute(softmax_rows_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
        set_error("softmax: cannot opt in to %d bytes of shared memory", bytes); return -3; }
      attr = bytes;
    }
    softmax_rows_smem_kernel<<<rows, 512, bytes, st>>>(X, cols, in_ld, (__nv_bfloat16*)op.p[UCDIR_SOFTMAX_P_OUT_BF16], op.i[UCDIR_SOFTMAX_I_OUT_LD]);
  } else if (cols <= 1024 && (in_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
             (!op.p[UCDIR_SOFTMAX_P_OUT_BF16] || ((op.i[UCDIR_SOFTMAX_I_OUT_LD] & 3) == 0 && op.i[UCDIR_SOFTMAX_I_OUT_LD] <= ((cols + 127) / 128) * 128 &&
                                                  (reinterpret_cast<uintptr_t>(op.p[UCDIR_SOFTMAX_P_OUT_BF16]) & 7) == 0))) {
    __nv_bfloat16* o = (__nv_bfloat16*)op.p[UCDIR_SOFTMAX_P_OUT_BF16];
    const int out_ld = op.i[UCDIR_SOFTMAX_I_OUT_LD];
    const int vpt = (cols + 127) / 128;
    long long blocks = ((long long)rows + 7) / 8;                       // 8 warps per CTA, one row per warp and pass
    if (blocks > 148 * 8) blocks = 148 * 8;
    const unsigned g = (unsigned)blocks;
    if (vpt <= 1) softmax_rows_warp_kernel<1><<<g, 256, 0, st>>>(X, rows, cols, in_ld, o, out_ld);
    else if (vpt <= 2) softmax_rows_warp_kernel<2><<<g, 256, 0, st>>>(X, rows, cols, in_ld, o, out_ld);
    else if (vpt <= 4) softmax_rows_warp_kernel<4><<<g, 256, 0, st>>>(X, rows, cols, in_ld, o, out_ld);
    else softmax_rows_warp_kernel<8><<<g, 256, 0, st>>>(X, rows, cols, in_ld, o, out_ld);
  } else {
    softmax_rows_kernel<<<rows, 256, 0, st>>>(X, cols, in_ld, (__nv_bfloat16*)op.p[UCDIR_SOFTMAX_P_OUT_BF16], op.i[UCDIR_SOFTMAX_I_OUT_LD]);
  }
  ++g_launches;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// MaxPool 2x2 (predictor, model/ucdir.py:363-375), NHWC fp32, float4 over channels.
// ------------------------------------------------------------------------------------------------
__global__ void maxpool2_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int B, int H, int W, int C4) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)B * H * W * C4;
  if (idx >= total) return;
  int c = idx % C4; size_t t = idx / C4;
  int x = t % W; t /= W;
  int y = t % H; int b = t / H;
  const float4* s = src + (((size_t)b * 2 * H + 2 * y) * (2 * W) + 2 * x) * C4 + c;
  float4 v00 = s[0], v01 = s[C4], v10 = s[(size_t)2 * W * C4], v11 = s[(size_t)2 * W * C4 + C4];
  float4 o;
  o.x = fmaxf(fmaxf(v00.x, v01.x), fmaxf(v10.x, v11.x));
  o.y = fmaxf(fmaxf(v00.y, v01.y), fmaxf(v10.y, v11.y));
  o.z = fmaxf(fmaxf(v00.z, v01.z), fmaxf(v10.z, v11.z));
  o.w = fmaxf(fmaxf(v00.w, v01.w), fmaxf(v10.w, v11.w));
  dst[idx] = o;
}

// MaxPool 2x2 on (hi, lo) bf16 plane pairs [B][2H][2W][2C] -> [B][H][W][2C] (fp32_tc predictor): the maximum of the four hi + lo
// values, re-split.  One thread per output pixel and 8 channels.
__global__ void __launch_bounds__(256) maxpool2_split_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int H, int W, int C) {
  const int c8 = C / 8;
  const size_t total = (size_t)B * H * W * c8;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % c8) * 8; size_t t = idx / c8;
    const int x = (int)(t % W); t /= W;
    const int y = (int)(t % H); const int b = (int)(t / H);
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const __nv_bfloat16* s = src + ((((size_t)b * 2 * H + 2 * y + dy) * (2 * W)) + 2 * x + dx) * 2 * C + c;
        const uint4 uh = __ldg(reinterpret_cast<const uint4*>(s)), ul = __ldg(reinterpret_cast<const uint4*>(s + C));
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&uh);
        const __nv_bfloat162* l2 = reinterpret_cast<const __nv_bfloat162*>(&ul);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 fh = __bfloat1622float2(h2[k]), fl = __bfloat1622float2(l2[k]);
          m[2 * k] = fmaxf(m[2 * k], fh.x + fl.x); m[2 * k + 1] = fmaxf(m[2 * k + 1], fh.y + fl.y);
        }
      }
    __align__(16) __nv_bfloat162 oh[4];
    __align__(16) __nv_bfloat162 ol[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      oh[k] = __floats2bfloat162_rn(m[2 * k], m[2 * k + 1]);
      const float2 r = __bfloat1622float2(oh[k]);
      ol[k] = __floats2bfloat162_rn(m[2 * k] - r.x, m[2 * k + 1] - r.y);
    }
    __nv_bfloat16* d = dst + (((size_t)b * H + y) * W + x) * 2 * C + c;
    *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(oh);
    *reinterpret_cast<uint4*>(d + C) = *reinterpret_cast<const uint4*>(ol);
  }
}

int launch_maxpool2(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  int B = op.i[UCDIR_POOL_I_B], H = op.i[UCDIR_POOL_I_H], W = op.i[UCDIR_POOL_I_W], C = op.i[UCDIR_POOL_I_C];
  if (!op.p[0] || !op.p[1] || B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 4) { set_error("maxpool2: bad args"); return -1; }
  if (op.i[UCDIR_POOL_I_SPLIT]) {
    if (C % 8) { set_error("maxpool2: split planes need C %% 8 == 0"); return -1; }
    if (dry) return 0;
    const size_t total8 = (size_t)B * H * W * (C / 8);
    const unsigned g = (unsigned)((total8 + 255) / 256 > 148 * 16 ? 148 * 16 : (total8 + 255) / 256);
    maxpool2_split_kernel<<<g, 256, 0, st>>>((const __nv_bfloat16*)op.p[0], (__nv_bfloat16*)op.p[1], B, H, W, C);
    ++g_launches;
    return 0;
  }
  if (dry) return 0;
  size_t total = (size_t)B * H * W * (C / 4);
  maxpool2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float4*)op.p[0], (float4*)op.p[1], B, H, W, C / 4);
  ++g_launches;
  return 0;
}

}  // namespace ucdir
