// final_conv (model/ucdir.py:266-268): GroupNorm(1,C) -> Swish -> Conv3x3(C -> out_channel) as ONE tcgen05 kernel.
//
// The Swish between the norm and the conv keeps the norm from being folded into the weights, so the streamed path ran an
// elementwise pass (read + write of the full-resolution C-channel tensor) and then a 16-column conv that re-fetched the
// activation slab for each of its nine taps.  Here the activation is read from HBM once:
//
// * work item = super tile of 4 adjacent 8 x 16 pixel tiles; one TMA box brings its 34 x 18 pixel halo per 64-channel chunk
//   (ucdir_dhalo.cu / ucdir_mix.cu explain the halo operand views);
// * TRANSFORM WARPS (12 of the 20) map the landed box in place through Swish(GroupNorm(x)) -- bf16 in, bf16 out, exactly what
//   the elementwise pass stored -- skipping out-of-image pixels (they are the conv's zero padding and must stay zero), then
//   fence.proxy.async and hand the stage to the MMA warp.  A thread always meets the same 8 channels (its 16-byte chunk
//   position and row phase are fixed by its index), so its scale / shift live in registers;
// * the nine taps x 4 tiles are N = 16 MMAs against the resident 18 KB weight block; two items in TMEM (64 columns each);
// * 4 epilogue warps add the bias and store the `ncol_valid` (3) fp32 outputs per pixel.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace ucdir {

struct FhParams {
  const double* stats0;
  const float* gamma; const float* beta; const float* tb;
  float* dst;
  int B, H, W;
  int nchunk, tiles_x, tiles_y, n_items;
  int dstC, dstCoff, ncol_valid;
  int crop;                    // DST holds only the interior: [B][H - 2*crop][W - 2*crop][dstC] (the stitched part of a tile)
  double gn_count; float eps;
};

constexpr int FH_NT = 16, FH_MT = 4, FH_SW = 8 * FH_MT, FH_BW = FH_SW + 2, FH_BH = 18;
constexpr int FH_A_BYTES = FH_BW * FH_BH * 128;                  // 78336
constexpr int FH_A_STAGE = (FH_A_BYTES + 1023) & ~1023;          // 78848
constexpr int FH_ASTG = 2;
constexpr int FH_WTAP = FH_NT * 128;                             // one tap of one 64-channel chunk: 16 rows x 128 bytes
constexpr int FH_MAX_CHUNKS = 2;
constexpr int FH_OFF_W = FH_ASTG * FH_A_STAGE;
constexpr int FH_OFF_BARS = FH_OFF_W + FH_MAX_CHUNKS * 9 * FH_WTAP;
constexpr int FH_TOTAL = FH_OFF_BARS + 128 + 1024 /* align slack */;
constexpr int FH_EPI_WARPS = 4, FH_FIRST_EPI_WARP = 4, FH_XF_WARPS = 12, FH_FIRST_XF_WARP = 8;
constexpr int FH_THREADS = 32 * (FH_FIRST_XF_WARP + FH_XF_WARPS);
constexpr int FH_VECS = FH_BW * FH_BH * 8;                       // 16-byte vectors of one halo box
static_assert((32 * FH_XF_WARPS) % 64 == 0, "a transform thread must keep its chunk position and row phase");

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}

__global__ void __launch_bounds__(FH_THREADS, 1) final_halo_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                   const __grid_constant__ CUtensorMap mapB, const FhParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wsm = smem + FH_OFF_W;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + FH_OFF_BARS);
  uint64_t* a_ready = a_full + FH_ASTG;
  uint64_t* a_empty = a_ready + FH_ASTG;
  uint64_t* tmem_full = a_empty + FH_ASTG;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* w_full = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int it0 = (int)((long long)p.n_items * blockIdx.x / gridDim.x), it1 = (int)((long long)p.n_items * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    prefetch_tmap(&mapA); prefetch_tmap(&mapB);
    for (int s = 0; s < FH_ASTG; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_ready[s], FH_XF_WARPS); mbar_init(&a_empty[s], 1); }
    for (int j = 0; j < 2; ++j) { mbar_init(&tmem_full[j], 1); mbar_init(&tmem_empty[j], FH_EPI_WARPS); }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");     // everything below touches data of earlier kernels

  if (warp == 0) {
    // ===================== producer: the weight block once, then one halo box per (item, chunk) =====================
    if (elect_one()) {
      mbar_expect_tx(w_full, (uint32_t)(p.nchunk * 9 * FH_WTAP));
      for (int j = 0; j < p.nchunk; ++j)
        for (int tap = 0; tap < 9; ++tap) tma_load_2d(&mapB, w_full, wsm + (j * 9 + tap) * FH_WTAP, (tap * p.nchunk + j) * 64, 0);
    }
    __syncwarp();
    int tx = it0 % p.tiles_x, t = it0 / p.tiles_x;
    int ty = t % p.tiles_y, img = t / p.tiles_y;
    int stage = 0; uint32_t phase = 0;
    for (int it = it0; it < it1; ++it) {
      for (int j = 0; j < p.nchunk; ++j) {
        mbar_wait(&a_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&a_full[stage], (uint32_t)FH_A_BYTES);
          tma_load_4d(&mapA, &a_full[stage], smem + stage * FH_A_STAGE, j * 64, tx * FH_SW - 1, ty * 16 - 1, img);
        }
        __syncwarp();
        if (++stage == FH_ASTG) { stage = 0; phase ^= 1; }
      }
      if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++img; } }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FH_NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t a_hi = desc_hi(FH_BW * 128, 2u), b_hi = desc_hi(1024, 2u);
    mbar_wait(w_full, 0);
    int stage = 0; uint32_t phase = 0;
    int slot = 0; uint32_t sph = 0;
    for (int it = it0; it < it1; ++it) {
      mbar_wait(&tmem_empty[slot], sph ^ 1);             // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(slot * 64);
      for (int j = 0; j < p.nchunk; ++j) {
        mbar_wait(&a_ready[stage], phase);               // landed AND transformed
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo0 = desc_lo(smem_u32(smem + stage * FH_A_STAGE));
          const uint32_t b_lo0 = desc_lo(smem_u32(wsm + j * 9 * FH_WTAP));
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int tyy = tap / 3, txx = tap - tyy * 3;
            const uint32_t a_lo = a_lo0 + (uint32_t)(((tyy * FH_BW + txx) * 128) >> 4);
            const uint32_t b_lo = b_lo0 + (uint32_t)((tap * FH_WTAP) >> 4);
#pragma unroll
            for (int mt = 0; mt < FH_MT; ++mt) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_lohi(tacc + (uint32_t)(mt * FH_NT), a_lo + (uint32_t)((mt * 8 * 128) >> 4) + (uint32_t)(k * 2), a_hi,
                               b_lo + (uint32_t)(k * 2), b_hi, idesc, (j | tap | k) != 0);
            }
          }
          umma_commit(&a_empty[stage]);                  // halo box may be overwritten
          if (j == p.nchunk - 1) umma_commit(&tmem_full[slot]);
        }
        __syncwarp();
        if (++stage == FH_ASTG) { stage = 0; phase ^= 1; }
      }
      if (++slot == 2) { slot = 0; sph ^= 1; }
    }
  } else if (warp >= FH_FIRST_XF_WARP) {
    // ===================== transform: Swish(GroupNorm(x)) on the landed halo box, in place =====================
    const int tt = threadIdx.x - 32 * FH_FIRST_XF_WARP;
    // vector v = tt + 384*i: its 16-byte chunk position (v & 7) and its row phase ((v >> 3) & 7) do not depend on i, so the
    // logical channel octet (the 128-byte swizzle XORs the two) is a per-thread constant
    const int lc = (tt & 7) ^ ((tt >> 3) & 7);
    float sa[8], sb[8];
    int cur_img = -1, cur_j = -1;
    int tx = it0 % p.tiles_x, t = it0 / p.tiles_x;
    int ty = t % p.tiles_y, img = t / p.tiles_y;
    int stage = 0; uint32_t phase = 0;
    float mean = 0.f, rstd = 1.f;
    for (int it = it0; it < it1; ++it) {
      const int x0 = tx * FH_SW - 1, y0 = ty * 16 - 1;
      for (int j = 0; j < p.nchunk; ++j) {
        if (img != cur_img) { const GnScalars sc = gn_scalars(p.stats0, nullptr, img, p.gn_count, p.eps); mean = sc.mean; rstd = sc.rstd; }
        if (img != cur_img || j != cur_j) {
          cur_img = img; cur_j = j;
#pragma unroll
          for (int k = 0; k < 8; ++k) {                 // h = (x*a + b) / 2: the Swish below works on x / 2
            const float a = rstd * __ldg(p.gamma + j * 64 + lc * 8 + k);
            sa[k] = 0.5f * a;
            sb[k] = 0.5f * (__ldg(p.beta + j * 64 + lc * 8 + k) - a * mean);
          }
        }
        mbar_wait(&a_full[stage], phase);
        uint4* box = reinterpret_cast<uint4*>(smem + stage * FH_A_STAGE);
        for (int v = tt; v < FH_VECS; v += 32 * FH_XF_WARPS) {
          const int r = v >> 3;                          // pixel row of the box
          const int by = r / FH_BW, bx = r - by * FH_BW;
          const int y = y0 + by, x = x0 + bx;
          if (y < 0 || y >= p.H || x < 0 || x >= p.W) continue;      // zero padding stays zero
          uint4 u = box[v];
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float2 f = __bfloat1622float2(h2[k]);
            const float hx = fmaf(f.x, sa[2 * k], sb[2 * k]), hy = fmaf(f.y, sa[2 * k + 1], sb[2 * k + 1]);
            h2[k] = __floats2bfloat162_rn(swish_half(hx), swish_half(hy));
          }
          box[v] = u;
        }
        fence_proxy_async();                             // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[stage]);
        if (++stage == FH_ASTG) { stage = 0; phase ^= 1; }
      }
      if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++img; } }
    }
  } else if (warp >= FH_FIRST_EPI_WARP) {
    // ===================== epilogue: + bias, fp32 store of the valid columns =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int yy = r >> 3, xx = r & 7;
    float bias[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) bias[n] = n < p.ncol_valid ? __ldg(p.tb + n) : 0.f;
    int tx = it0 % p.tiles_x, t = it0 / p.tiles_x;
    int ty = t % p.tiles_y, img = t / p.tiles_y;
    int slot = 0; uint32_t sph = 0;
    for (int it = it0; it < it1; ++it) {
      mbar_wait(&tmem_full[slot], sph);
      tc_fence_after();
      uint32_t rv[FH_MT][4];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 64);
#pragma unroll
      for (int mt = 0; mt < FH_MT; ++mt) tmem_ld4(taddr + mt * FH_NT, rv[mt]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[slot]);
      if (++slot == 2) { slot = 0; sph ^= 1; }
      const int y = ty * 16 + yy;
#pragma unroll
      for (int mt = 0; mt < FH_MT; ++mt) {
        const int x = tx * FH_SW + mt * 8 + xx;
        if (y >= p.crop && y < p.H - p.crop && x >= p.crop && x < p.W - p.crop) {
          float* d = p.dst + (((size_t)img * (p.H - 2 * p.crop) + (y - p.crop)) * (p.W - 2 * p.crop) + (x - p.crop)) * p.dstC + p.dstCoff;
#pragma unroll
          for (int n = 0; n < 4; ++n)
            if (n < p.ncol_valid) d[n] = __uint_as_float(rv[mt][n]) + bias[n];
        }
      }
      if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++img; } }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static const bool g_fh_pdl = []() { const char* e = getenv("UCDIR_PDL"); return !(e && e[0] == '0'); }();

// true when the op asks for the fused GroupNorm + Swish source transform and fits this kernel (launch_tc_conv refuses a
// record that asks for the transform and does not fit: the streamed kernel has no such prologue)
bool tc_final_halo_applies(const ucdir_op_t& op) {
  const int C0 = op.i[UCDIR_TC_I_C0], H = op.i[UCDIR_TC_I_H], W = op.i[UCDIR_TC_I_W];
  const int KB = op.i[UCDIR_TC_I_KB] ? op.i[UCDIR_TC_I_KB] : op.i[UCDIR_TC_I_KC];
  const int ncv = op.i[UCDIR_TC_I_NCOL_VALID] ? op.i[UCDIR_TC_I_NCOL_VALID] : op.i[UCDIR_TC_I_NTOT];
  return op.i[UCDIR_TC_I_SRC_GN_SWISH] == 1 && op.i[UCDIR_TC_I_SPLIT] == 0 && op.i[UCDIR_TC_I_MODE] == 0 && op.i[UCDIR_TC_I_GROUPS] == 1 && op.i[UCDIR_TC_I_NT] == FH_NT &&
         op.i[UCDIR_TC_I_NTOT] == FH_NT && op.i[UCDIR_TC_I_KC] == 64 && KB == 64 && C0 % 64 == 0 && C0 / 64 <= FH_MAX_CHUNKS &&
         op.i[UCDIR_TC_I_C1] == 0 && op.i[UCDIR_TC_I_GN] == 0 && op.i[UCDIR_TC_I_ACT] == 0 && op.i[UCDIR_TC_I_NTY] == 3 &&
         op.i[UCDIR_TC_I_NTX] == 3 && op.i[UCDIR_TC_I_OY0] == -1 && op.i[UCDIR_TC_I_OX0] == -1 && op.i[UCDIR_TC_I_STRIDE] == 1 &&
         H >= 2 && W >= 2 && op.i[UCDIR_TC_I_SRC_H] == H && op.i[UCDIR_TC_I_SRC_W] == W && !op.p[UCDIR_TC_P_RES] &&
         op.i[UCDIR_TC_I_DST_F32] == 1 && ncv >= 1 && ncv <= 4 && !op.i[UCDIR_TC_I_DST_UP] && !op.i[UCDIR_TC_I_W_BATCHED] &&
         !op.p[UCDIR_TC_P_DST2] && !op.p[UCDIR_TC_P_DST_STATS] && (op.i[UCDIR_TC_I_SRC_CSTRIDE] == 0 || op.i[UCDIR_TC_I_SRC_CSTRIDE] == C0) &&
         op.p[UCDIR_TC_P_SRC_GAMMA] && op.p[UCDIR_TC_P_SRC_BETA] && op.p[UCDIR_TC_P_STATS0];
}

int launch_tc_final_halo(const ucdir_op_t& op, cudaStream_t st) {
  FhParams p;
  const int C0 = op.i[UCDIR_TC_I_C0];
  p.stats0 = (const double*)op.p[UCDIR_TC_P_STATS0];
  p.gamma = (const float*)op.p[UCDIR_TC_P_SRC_GAMMA]; p.beta = (const float*)op.p[UCDIR_TC_P_SRC_BETA];
  p.tb = (const float*)op.p[UCDIR_TC_P_TB];
  p.dst = (float*)op.p[UCDIR_TC_P_DST];
  p.B = op.i[UCDIR_TC_I_B]; p.H = op.i[UCDIR_TC_I_H]; p.W = op.i[UCDIR_TC_I_W];
  p.nchunk = C0 / 64;
  p.dstC = op.i[UCDIR_TC_I_DST_C]; p.dstCoff = op.i[UCDIR_TC_I_DST_COFF];
  p.crop = op.i[UCDIR_TC_I_DST_CROP];
  if (p.crop < 0 || 2 * p.crop >= op.i[UCDIR_TC_I_H] || 2 * p.crop >= op.i[UCDIR_TC_I_W]) { set_error("tc_final_halo: bad DST_CROP"); return -1; }
  p.ncol_valid = op.i[UCDIR_TC_I_NCOL_VALID] ? op.i[UCDIR_TC_I_NCOL_VALID] : FH_NT;
  p.eps = op.f[UCDIR_TC_F_EPS];
  p.gn_count = (double)C0 * p.H * p.W;
  p.tiles_x = (p.W + FH_SW - 1) / FH_SW; p.tiles_y = (p.H + 15) / 16;
  const long long items = (long long)p.tiles_x * p.tiles_y * p.B;
  if (items > 0x7fffffffLL) { set_error("tc_final_halo: too many items"); return -2; }
  p.n_items = (int)items;
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("tc_final_halo: cuTensorMapEncodeTiled unavailable"); return -3; }
  CUtensorMap ma, mb;
  {
    cuuint64_t dims[4] = {(cuuint64_t)C0, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)C0 * 2, (cuuint64_t)C0 * 2 * p.W, (cuuint64_t)C0 * 2 * p.W * p.H};
    cuuint32_t box[4] = {64, FH_BW, FH_BH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(op.p[UCDIR_TC_P_SRC0]), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tc_final_halo: cuTensorMapEncodeTiled(activation C=%d W=%d H=%d B=%d) failed: %d", C0, p.W, p.H, p.B, (int)r); return -3; }
  }
  {
    const int Ktot = 9 * C0;
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)FH_NT};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {64, FH_NT};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(op.p[UCDIR_TC_P_W]), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tc_final_halo: cuTensorMapEncodeTiled(weights K=%d) failed: %d", Ktot, (int)r); return -3; }
  }
  static bool attr_dev[UCDIR_MAX_DEV] = {};
  bool& attr = attr_dev[cur_dev()];
  if (!attr) {
    if (cudaFuncSetAttribute(final_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_TOTAL) != cudaSuccess) {
      set_error("tc_final_halo: cannot opt in to %d bytes of shared memory: %s", FH_TOTAL, cudaGetErrorString(cudaGetLastError())); return -3; }
    attr = true;
  }
  const int n_sm = sm_count();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(items < n_sm ? items : n_sm)); cfg.blockDim = dim3(FH_THREADS); cfg.dynamicSmemBytes = FH_TOTAL; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = g_fh_pdl ? 1 : 0;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, final_halo_kernel, ma, mb, p) != cudaSuccess) {
    set_error("tc_final_halo: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return -3; }
  ++g_launches;
  return 0;
}

}  // namespace ucdir
