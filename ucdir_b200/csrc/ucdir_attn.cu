// Fused self-attention core (sm_100a): O = softmax(Q K^T / sqrt(C)) V for ONE head of C = 512 channels -- the two einsums and
// the softmax of SelfAttention.forward (model/ucdir.py:174-179) in one kernel.  The N x N score matrix (1 GiB fp32 per 1024x1024
// tile, 16 384 tokens) and the probabilities never leave the SM: scores accumulate in TMEM, the online softmax runs out of
// TMEM into registers, probabilities go to shared memory as the A operand of the second GEMM.
//
//   per CTA:  128 queries x ONE half of the output channels (d-split: the 128 x 512 fp32 output alone would fill all 512 TMEM
//             columns; with a 128 x 256 half there is room for two 128 x 128 score buffers).  The two CTAs of a query block
//             compute the same scores (QK^T is done twice: 1.5x the algorithmic MMA work) but need no cross-CTA exchange.
//   TMEM      [0,128) S0 | [128,256) S1 | [256,512) O (fp32)
//   smem      Q 128 x 512 bf16 resident (8 swizzled 64-channel chunks, 128 KB) | P 128 x 128 bf16 (2 chunks, 32 KB)
//             | ring of 4 x 16 KB slots: K chunks (128 keys x 64 channels, 1 slot) and V^T chunks (256 channels x 64 keys, 2 slots)
//   warps     0: TMA producer   1: tcgen05.mma issuer (+ TMEM allocation)   2..5: softmax / rescale / epilogue, one thread per query
//   pipeline  key blocks of 128: QK(0); for j: { QK(j+1); PV(j) } -- the scores of block j+1 are produced while the softmax
//             warps work on block j and the tensor pipe runs PV(j-1 / j).
//   softmax   online with LAZY rescaling: the running maximum is only replaced when the block maximum exceeds it by more
//             than 8 (log2 units), so probabilities stay <= 2^8 and the 128 x 256 accumulator is rescaled (TMEM load,
//             multiply, TMEM store) only on the first blocks of a row; 1/l is applied once in the epilogue.
// Operands: Q / K = channel slices [0,C) / [C,2C) of the qkv convolution's output rows (bf16 [B][N][QK_LD]); V^T = bf16
// [B][C][VT_LD] written transposed by that convolution's epilogue (UCDIR_TC_P_DST2).  Keys beyond N are zero filled by TMA and
// masked to -inf; query rows beyond N are computed on zeros and not stored.
#include <cuda.h>
#include <cstdio>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace ucdir {

constexpr int AT_C = 512;                 // channels (= head dimension: n_head = 1, model/ucdir.py:157,192)
constexpr int AT_DH = 256;                // output channels per CTA
constexpr int AT_BM = 128;                // queries per CTA
constexpr int AT_BN = 128;                // keys per block
constexpr int AT_KC = 64;                 // channels per Q / K chunk, keys per V^T / P chunk (128-byte swizzled rows)
constexpr int AT_NQC = AT_C / AT_KC;      // 8 Q / K chunks per block
constexpr int AT_SLOT = 128 * 128;        // 16 KB ring slot
constexpr int AT_NSLOT = 4;
constexpr int AT_Q_BYTES = AT_NQC * AT_SLOT;             // 128 KB
constexpr int AT_P_BYTES = 2 * AT_SLOT;                  // 32 KB
constexpr int AT_RING_BYTES = AT_NSLOT * AT_SLOT;        // 64 KB
constexpr int AT_BARS = 32 * 8;
constexpr int AT_TOTAL = AT_Q_BYTES + AT_P_BYTES + AT_RING_BYTES + AT_BARS + 1024 /* align slack */;
constexpr int AT_THREADS = 192;
constexpr float AT_LAZY = 8.0f;           // log2 units

struct AttnParams {
  __nv_bfloat16* o;
  int N, o_ld;
  float scale_log2;                       // SCALE * log2(e)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(AT_THREADS, 1) flash_attn_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                                                                   const __grid_constant__ CUtensorMap mapV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + AT_Q_BYTES;
  uint8_t* ring = sP + AT_P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + AT_RING_BYTES);
  uint64_t* full = bars;                  // [4]  TMA -> MMA
  uint64_t* empty = full + AT_NSLOT;      // [4]  MMA -> TMA
  uint64_t* s_full = empty + AT_NSLOT;    // [2]  MMA -> softmax: scores of a block complete
  uint64_t* s_free = s_full + 2;          // [2]  softmax -> MMA: score buffer read
  uint64_t* q_full = s_free + 2;          //      TMA -> MMA: the query block landed
  uint64_t* p_full = q_full + 1;          //      softmax -> MMA: probabilities of a block in shared memory (accumulator rescaled)
  uint64_t* p_free = p_full + 1;          //      MMA -> softmax: PV of a block retired (P reusable, accumulator readable)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_free + 1);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = blockIdx.x & 1, qblk = blockIdx.x >> 1, img = blockIdx.y;
  const int q0 = qblk * AT_BM;
  const int nkb = (p.N + AT_BN - 1) / AT_BN;

  if (threadIdx.x == 0) {
    prefetch_tmap(&mapQ); prefetch_tmap(&mapK); prefetch_tmap(&mapV);
    for (int s = 0; s < AT_NSLOT; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_free[b], 4); }
    mbar_init(q_full, 1); mbar_init(p_full, 4); mbar_init(p_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(q_full, AT_Q_BYTES);
      for (int c = 0; c < AT_NQC; ++c) tma_load_3d(&mapQ, q_full, sQ + c * AT_SLOT, c * AT_KC, q0, img);
    }
    __syncwarp();
    int slot = 0; uint32_t phase = 0;
    auto load_k = [&](int j) {
      for (int c = 0; c < AT_NQC; ++c) {
        mbar_wait(&empty[slot], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[slot], AT_SLOT);
          tma_load_3d(&mapK, &full[slot], ring + slot * AT_SLOT, c * AT_KC, j * AT_BN, img);
        }
        __syncwarp();
        if (++slot == AT_NSLOT) { slot = 0; phase ^= 1; }
      }
    };
    auto load_v = [&](int j) {
      for (int c = 0; c < 2; ++c) {                     // 64 keys x 256 channels = two adjacent slots (slot is even here)
        mbar_wait(&empty[slot], phase ^ 1);
        mbar_wait(&empty[slot + 1], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[slot], 2 * AT_SLOT);
          tma_load_3d(&mapV, &full[slot], ring + slot * AT_SLOT, j * AT_BN + c * AT_KC, half * AT_DH, img);
          mbar_arrive(&full[slot + 1]);                 // keeps the second slot's phase in step (the MMA warp consumes it with the first)
        }
        __syncwarp();
        slot += 2;
        if (slot == AT_NSLOT) { slot = 0; phase ^= 1; }
      }
    };
    load_k(0);
    for (int j = 0; j < nkb; ++j) {
      if (j + 1 < nkb) load_k(j + 1);
      load_v(j);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_DH >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int slot = 0; uint32_t phase = 0;
    mbar_wait(q_full, 0);
    tc_fence_after();
    auto qk = [&](int j) {
      const int b = j & 1;
      mbar_wait(&s_free[b], ((j >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tS = tmem_base + (uint32_t)(b * AT_BN);
      for (int c = 0; c < AT_NQC; ++c) {
        mbar_wait(&full[slot], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = make_desc(smem_u32(sQ + c * AT_SLOT), 128), bd = make_desc(smem_u32(ring + slot * AT_SLOT), 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tS, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc_s, (c | k) != 0);
          umma_commit(&empty[slot]);
          if (c == AT_NQC - 1) umma_commit(&s_full[b]);
        }
        __syncwarp();
        if (++slot == AT_NSLOT) { slot = 0; phase ^= 1; }
      }
    };
    auto pv = [&](int j) {
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      const uint32_t tO = tmem_base + 256u;
      for (int c = 0; c < 2; ++c) {
        mbar_wait(&full[slot], phase);
        mbar_wait(&full[slot + 1], phase);              // the pair's second barrier (arrived by the producer at issue time): consumed
        tc_fence_after();                               // here so that every completed phase has a waiter
        if (elect_one()) {
          const uint64_t ad = make_desc(smem_u32(sP + c * AT_SLOT), 128), bd = make_desc(smem_u32(ring + slot * AT_SLOT), 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tO, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc_o, (j | c | k) != 0);
          umma_commit(&empty[slot]);
          umma_commit(&empty[slot + 1]);
          if (c == 1) umma_commit(p_free);
        }
        __syncwarp();
        slot += 2;
        if (slot == AT_NSLOT) { slot = 0; phase ^= 1; }
      }
    };
    qk(0);
    for (int j = 0; j < nkb; ++j) {
      if (j + 1 < nkb) qk(j + 1);
      pv(j);
    }
  } else {
    // ===================== softmax / rescale / epilogue: one thread per query row =====================
    const int q = warp & 3;                              // TMEM lane quadrant of this warp
    const int r = q * 32 + lane;                         // query row within the block = TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float m_used = -INFINITY, l = 0.f;
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;
    for (int j = 0; j < nkb; ++j) {
      const int b = j & 1;
      mbar_wait(&s_full[b], (j >> 1) & 1);
      tc_fence_after();
      float s[AT_BN];
#pragma unroll
      for (int c = 0; c < AT_BN; c += 32) tmem_ld32(lane_addr + (uint32_t)(b * AT_BN + c), reinterpret_cast<uint32_t*>(s + c));
      tmem_ld_wait();
      // the score buffer is in registers: hand it back so that QK(j+2) can start
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[b]);
      const int kvalid = p.N - j * AT_BN;               // keys of this block that exist (>= 128 except in the last block)
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < AT_BN; ++c) {
        s[c] = (c < kvalid) ? s[c] * p.scale_log2 : -INFINITY;
        mx = fmaxf(mx, s[c]);
      }
      float factor = 1.f;
      const bool need = mx > m_used + AT_LAZY;          // first block: m_used = -inf -> need
      if (need) { factor = ex2_approx(m_used - mx); m_used = mx; }      // exp2(-inf) = 0 on the first block
      float sum = 0.f;
      uint32_t pk[AT_BN / 2];
#pragma unroll
      for (int c = 0; c < AT_BN; c += 2) {
        const float a = ex2_approx(s[c] - m_used), bb = ex2_approx(s[c + 1] - m_used);
        sum += a + bb;
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, bb);
        pk[c >> 1] = *reinterpret_cast<const uint32_t*>(&h);
      }
      l = l * factor + sum;
      if (j > 0) {
        mbar_wait(p_free, (j - 1) & 1);                 // PV(j-1) retired: P may be overwritten, the accumulator may be touched
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {            // rescale this warp's 32 accumulator rows (rare after the first blocks)
#pragma unroll 1
          for (int c = 0; c < AT_DH; c += 32) {
            uint32_t o[32];
            tmem_ld32(lane_addr + (uint32_t)(256 + c), o);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * factor);
            tmem_st32(lane_addr + (uint32_t)(256 + c), o);
          }
          tmem_st_wait();
        }
      }
      // probabilities -> shared memory in the swizzled K-major layout the MMA reads: 16-byte chunk i of row r at (i ^ (r & 7))
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint8_t* dst = prow + (i >> 3) * AT_SLOT + (((i & 7) ^ sw) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: O / l -> bf16
    mbar_wait(p_free, (nkb - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    const int row = q0 + r;
    __nv_bfloat16* orow = p.o + ((size_t)img * p.N + row) * p.o_ld + half * AT_DH;
#pragma unroll 1
    for (int c = 0; c < AT_DH; c += 32) {
      uint32_t o[32];
      tmem_ld32(lane_addr + (uint32_t)(256 + c), o);
      tmem_ld_wait();
      if (row < p.N) {
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          __align__(16) __nv_bfloat162 h[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(__uint_as_float(o[k + 2 * e]) * inv, __uint_as_float(o[k + 2 * e + 1]) * inv);
          *reinterpret_cast<uint4*>(orow + c + k) = *reinterpret_cast<const uint4*>(h);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

static int make_map3(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes, uint64_t s2_bytes, uint32_t b0, uint32_t b1,
                     const char* what) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("tc_attn: cuTensorMapEncodeTiled unavailable"); return -3; }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1_bytes, s2_bytes};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("tc_attn: cuTensorMapEncodeTiled(%s) failed: %d", what, (int)r); return -3; }
  return 0;
}

int launch_tc_attn(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  const __nv_bfloat16* qk = (const __nv_bfloat16*)op.p[UCDIR_ATTN_P_QK];
  const __nv_bfloat16* vt = (const __nv_bfloat16*)op.p[UCDIR_ATTN_P_VT];
  AttnParams p;
  p.o = (__nv_bfloat16*)op.p[UCDIR_ATTN_P_O];
  const int B = op.i[UCDIR_ATTN_I_B], C = op.i[UCDIR_ATTN_I_C];
  p.N = op.i[UCDIR_ATTN_I_N];
  const int qk_ld = op.i[UCDIR_ATTN_I_QK_LD], vt_ld = op.i[UCDIR_ATTN_I_VT_LD];
  p.o_ld = op.i[UCDIR_ATTN_I_O_LD];
  if (!qk || !vt || !p.o) { set_error("tc_attn: null pointer"); return -1; }
  if (B <= 0 || p.N <= 0) { set_error("tc_attn: bad dims"); return -1; }
  if (C != AT_C) { set_error("tc_attn: the fused kernel is built for one head of %d channels (got %d)", AT_C, C); return -2; }
  if (qk_ld < 2 * C || qk_ld % 8 || vt_ld < p.N || vt_ld % 8 || p.o_ld < C || p.o_ld % 8) {
    set_error("tc_attn: row pitches must cover the operands and be multiples of 8 elements (QK_LD=%d VT_LD=%d O_LD=%d)", qk_ld, vt_ld, p.o_ld); return -2; }
  if (((uintptr_t)qk | (uintptr_t)vt | (uintptr_t)p.o) & 15) { set_error("tc_attn: operands must be 16-byte aligned"); return -2; }
  const float scale = op.f[UCDIR_ATTN_F_SCALE] != 0.f ? op.f[UCDIR_ATTN_F_SCALE] : 1.0f / sqrtf((float)C);
  p.scale_log2 = scale * 1.4426950408889634f;
  if (dry) return 0;
  CUtensorMap mq, mk, mv;
  int rc = make_map3(&mq, qk, (uint64_t)C, (uint64_t)p.N, (uint64_t)B, (uint64_t)qk_ld * 2, (uint64_t)qk_ld * 2 * p.N, AT_KC, AT_BM, "Q");
  if (rc) return rc;
  rc = make_map3(&mk, qk + C, (uint64_t)C, (uint64_t)p.N, (uint64_t)B, (uint64_t)qk_ld * 2, (uint64_t)qk_ld * 2 * p.N, AT_KC, AT_BN, "K");
  if (rc) return rc;
  rc = make_map3(&mv, vt, (uint64_t)p.N, (uint64_t)C, (uint64_t)B, (uint64_t)vt_ld * 2, (uint64_t)vt_ld * 2 * C, AT_KC, AT_DH, "V^T");
  if (rc) return rc;
  static bool attr_dev[UCDIR_MAX_DEV] = {};
  bool& attr = attr_dev[cur_dev()];
  if (!attr) {
    if (cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_TOTAL) != cudaSuccess) {
      set_error("tc_attn: cannot opt in to %d bytes of shared memory: %s", AT_TOTAL, cudaGetErrorString(cudaGetLastError())); return -3; }
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * ((p.N + AT_BM - 1) / AT_BM)), (unsigned)B, 1);
  cfg.blockDim = dim3(AT_THREADS); cfg.dynamicSmemBytes = AT_TOTAL; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, flash_attn_kernel, mq, mk, mv, p) != cudaSuccess) {
    set_error("tc_attn: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return -3; }
  ++g_launches;
  return 0;
}

}  // namespace ucdir
