// Dense 3x3 stride-1 convolution with few output channels (Cout = 64 / 128: conv1 of the ResnetBlockDY3h blocks at the two
// finest levels, model/ucdir.py:109-111,124-126) on tcgen05, "halo" schedule.
//
// The streamed form (tc_conv_kernel) is bound by L2 -> SM traffic on these layers: with only 64 / 128 accumulator columns
// per pixel tile, every filter tap re-fetches a 16 KB activation slab and an 8 / 16 KB weight slab for ~130 / 260 cycles
// of tensor work.  Here a work item is a SUPER TILE of MT = 256 / Cout adjacent 8 x 16 pixel tiles:
//
// * one TMA box per 64-channel chunk brings the (8*MT + 2) x 18 pixel halo of the super tile; the nine taps of the MT
//   tiles are views of that box (operand descriptor start = (ty*BW + tx + 8*mt) rows in, 8-row groups BW rows apart --
//   see ucdir_mix.cu for the addressing argument), so activations cross L2 -> SM ~1.2x instead of 9x (3x with ROW3);
// * every weight slab (one tap of one chunk) is used by the MT tiles of the item before it is released, so weights cross
//   MT times less often;  MT x Cout = 256 accumulator columns per item, two items in TMEM (MMA of item i+1 overlaps
//   the epilogue of item i);
// * separate producer warps for activations and weights (the activation box of chunk j+1 is in flight while the taps
//   of chunk j stream), 16 epilogue warps (warp j of a quadrant drains the 64-column stripe j of every item).
//
// SPLIT (fp32_tc): activations are (hi, lo) bf16 plane pairs, weights per tap [W_hi | W_hi | W_lo] per source.  The
// box loop then visits, per 64-channel chunk, the lo-plane box (one weight pass: W_hi) and the hi-plane box (two passes: W_hi,
// W_lo -- the hi box is fetched once for both); everything accumulates into the same fp32 accumulator; fp32 epilogue, (hi, lo) stores.
//
// torch.cat((x, skip)) inputs are two tensor maps visited by the same chunk loop; GroupNorm(1,C) of the input is folded
// exactly as in tc_conv_kernel (gamma in the weights, 9 border classes of additive terms); Swish epilogue; statistics
// of the stored tensor for the next GroupNorm.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace ucdir {

struct DhParams {
  const double* stats0; const double* stats1;
  const float* tb; const float* tg;
  __nv_bfloat16* dst; double* dst_stats;
  const float* tb2; __nv_bfloat16* dst2;                 // RES: bias and destination of the fused 1x1 res_conv
  int B, H, W;
  int nchunk, c0_chunks;                                 // boxes per item and boxes of source 0 (SPLIT: two boxes per channel chunk)
  int n0, n1;                                            // SPLIT: channel chunks of source 0 / 1
  int a0_lo, a1_lo;                                      // SPLIT: channel offset of the lo plane inside a pixel row of source 0 / 1
  int ktap;                                              // weight K elements per filter tap (SPLIT: 3 * Cin)
  int dst_lo, dst2_lo;                                   // SPLIT: element offset of the lo plane inside a destination row
  int tiles_x, tiles_y, n_items;
  int gn, act, dstC, dstCoff, st32, dst2C, st32_2;
  double gn_count; float eps;
};

constexpr int DH_EPI_WARPS = 16, DH_FIRST_EPI_WARP = 4;
constexpr int DH_THREADS = 32 * (DH_FIRST_EPI_WARP + DH_EPI_WARPS);
constexpr int DH_REGS_LOW = 40, DH_REGS_HIGH = 104;      // see MX_REGS_LOW / MX_REGS_HIGH in ucdir_mix.cu

// KC = channels per chunk = one pixel row of the halo box: 64 (128-byte rows, 128B swizzle) or 16 (the in-conv's zero-padded
// 6 -> 16 input channels: 32-byte rows, 32B swizzle, one K step per tap)
// RES = 1: the block's 1x1 res_conv (model/ucdir.py:120,140: res_conv(x) of the SAME un-normalised input) rides along: its
// operand is the centre-tap view of the halo box that is already in shared memory, so the input is not read from HBM a second
// time.  The item then is 128/NT tiles with NT conv1 + NT res_conv accumulator columns each.
template <int NT, int KC, int RES = 0>
struct DhCfg {
  static constexpr int ROWB = KC * 2;                     // bytes per pixel row of a chunk
  static constexpr uint32_t LAYOUT = KC == 64 ? 2u : 6u;  // operand descriptor swizzle mode: 128B / 32B
  static constexpr int KSTEPS = KC / 16;
  static constexpr int MT = (RES ? 128 : 256) / NT;       // 8 x 16 pixel tiles per item
  static constexpr int SLABS = 9 + RES;                   // weight slabs per chunk: nine taps (+ the res_conv)
  static constexpr int SW = 8 * MT;                       // super tile width
  static constexpr int BW = SW + 2, BH = 18;              // halo box
  static constexpr int A_BYTES = BW * BH * ROWB;
  static constexpr int A_STAGE = (A_BYTES + 1023) & ~1023;
  static constexpr int ASTG = RES ? (NT == 64 ? 3 : 4) : (NT == 64 ? 2 : 3);   // RES items are half as large: one more box in flight
  static constexpr int BSLAB = NT * ROWB;                 // one tap of one chunk
  static constexpr int BSTG = RES ? (NT == 64 ? 8 : 7) : (NT == 64 ? 6 : 5);
  static constexpr int OFF_B = ASTG * A_STAGE;
  static constexpr int OFF_CTAB = OFF_B + BSTG * BSLAB;
  static constexpr int OFF_BARS = OFF_CTAB + 9 * NT * 4 + NT * 4;   // + bias table of the res_conv
  static constexpr int TOTAL = OFF_BARS + 256 + 1024 /* align slack */;
  static_assert(TOTAL <= 227 * 1024, "shared memory");
};

struct SuperCursor {            // item = (image, super-tile row, super-tile column), column fastest
  int img, ty, tx;
  __device__ __forceinline__ void init(int it, int tiles_x, int tiles_y) {
    tx = it % tiles_x;
    const int t = it / tiles_x;
    ty = t % tiles_y;
    img = t / tiles_y;
  }
  __device__ __forceinline__ void next(int tiles_x, int tiles_y) {
    if (++tx == tiles_x) { tx = 0; if (++ty == tiles_y) { ty = 0; ++img; } }
  }
};

// Box j of an item: which tensor map / channel coordinate it loads and which weight slabs (K offset inside a filter tap) meet it.
struct BoxInfo { bool use1; int coff; int npass; int k[2]; };
template <int KC, bool SPLIT>
__device__ __forceinline__ BoxInfo box_info(const DhParams& p, int j) {
  BoxInfo b;
  if (!SPLIT) {
    b.use1 = j >= p.c0_chunks;
    b.coff = (b.use1 ? j - p.c0_chunks : j) * KC;
    b.npass = 1; b.k[0] = j * KC; b.k[1] = 0;
    return b;
  }
  // per tap [s0 W_hi | s0 W_hi | s1 W_hi | s1 W_hi | s0 W_lo | s1 W_lo] (engine.py:_split_k); boxes: (chunk 0 lo, chunk 0 hi, chunk 1 lo, ...)
  const int c = j >> 1;
  const bool lo = (j & 1) == 0;
  b.use1 = c >= p.n0;
  const int cc = b.use1 ? c - p.n0 : c;
  b.coff = cc * KC + (lo ? (b.use1 ? p.a1_lo : p.a0_lo) : 0);
  if (lo) { b.npass = 1; b.k[0] = (b.use1 ? 2 * p.n0 + p.n1 + cc : p.n0 + cc) * KC; b.k[1] = 0; }
  else { b.npass = 2; b.k[0] = (b.use1 ? 2 * p.n0 + cc : cc) * KC; b.k[1] = (b.use1 ? 3 * p.n0 + 2 * p.n1 + cc : 2 * p.n0 + 2 * p.n1 + cc) * KC; }
  return b;
}

template <int NT, int ACT, int KC, int RES, bool SPLIT>
__global__ void __launch_bounds__(DH_THREADS, 1) dense_halo_kernel(const __grid_constant__ CUtensorMap mapA0,
                                                                   const __grid_constant__ CUtensorMap mapA1,
                                                                   const __grid_constant__ CUtensorMap mapB,
                                                                   const __grid_constant__ CUtensorMap mapB2, const DhParams p) {
  using S = DhCfg<NT, KC, RES>;
  constexpr int MT = S::MT, ASTG = S::ASTG, BSTG = S::BSTG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* bring = smem + S::OFF_B;
  float* ctab = reinterpret_cast<float*>(smem + S::OFF_CTAB);          // [9][NT] additive terms of the current image
  float* btab2 = ctab + 9 * NT;                                         // [NT] bias of the fused res_conv
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + S::OFF_BARS);
  uint64_t* a_empty = a_full + ASTG;
  uint64_t* b_full = a_empty + ASTG;
  uint64_t* b_empty = b_full + BSTG;
  uint64_t* tmem_full = b_empty + BSTG;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int it0 = (int)((long long)p.n_items * blockIdx.x / gridDim.x), it1 = (int)((long long)p.n_items * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    prefetch_tmap(&mapA0); prefetch_tmap(&mapB);
    if (p.c0_chunks < p.nchunk) prefetch_tmap(&mapA1);
    if (RES) prefetch_tmap(&mapB2);
    for (int s = 0; s < ASTG; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < BSTG; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int j = 0; j < 2; ++j) { mbar_init(&tmem_full[j], 1); mbar_init(&tmem_empty[j], DH_EPI_WARPS); }
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
  asm volatile("griddepcontrol.wait;" ::: "memory");     // everything below touches data of earlier kernels

  if (warp < DH_FIRST_EPI_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DH_REGS_LOW));
    if (warp == 0) {
      // ===================== activation producer: one halo box per (item, 64-channel chunk) =====================
      SuperCursor cur; cur.init(it0, p.tiles_x, p.tiles_y);
      int stage = 0; uint32_t phase = 0;
      for (int it = it0; it < it1; ++it) {
        const int x0 = cur.tx * S::SW - 1, y0 = cur.ty * 16 - 1;
        for (int j = 0; j < p.nchunk; ++j) {
          mbar_wait(&a_empty[stage], phase ^ 1);
          if (elect_one()) {
            const BoxInfo bi = box_info<KC, SPLIT>(p, j);
            mbar_expect_tx(&a_full[stage], (uint32_t)S::A_BYTES);
            tma_load_4d(bi.use1 ? &mapA1 : &mapA0, &a_full[stage], smem + stage * S::A_STAGE, bi.coff, x0, y0, cur.img);
          }
          __syncwarp();
          if (++stage == ASTG) { stage = 0; phase ^= 1; }
        }
        cur.next(p.tiles_x, p.tiles_y);
      }
    } else if (warp == 2) {
      // ===================== weight producer: one slab per (item, chunk, tap) =====================
      int stage = 0; uint32_t phase = 0;
      for (int it = it0; it < it1; ++it) {
        for (int j = 0; j < p.nchunk; ++j) {
          const BoxInfo bi = box_info<KC, SPLIT>(p, j);
#pragma unroll 1
          for (int ps = 0; ps < bi.npass; ++ps) {
#pragma unroll 1
            for (int tap = 0; tap < S::SLABS; ++tap) {      // slab 9 = the res_conv's weights of this chunk
              mbar_wait(&b_empty[stage], phase ^ 1);
              if (elect_one()) {
                mbar_expect_tx(&b_full[stage], (uint32_t)S::BSLAB);
                if (tap < 9) tma_load_2d(&mapB, &b_full[stage], bring + stage * S::BSLAB, tap * p.ktap + bi.k[ps], 0);
                else tma_load_2d(&mapB2, &b_full[stage], bring + stage * S::BSLAB, bi.k[ps], 0);
              }
              __syncwarp();
              if (++stage == BSTG) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int as = 0; uint32_t aph = 0;
      int bs = 0; uint32_t bph = 0;
      int slot = 0; uint32_t sph = 0;
      for (int it = it0; it < it1; ++it) {
        mbar_wait(&tmem_empty[slot], sph ^ 1);             // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(slot * 256);
        for (int j = 0; j < p.nchunk; ++j) {
          mbar_wait(&a_full[as], aph);
          const uint32_t a_base = smem_u32(smem + as * S::A_STAGE);
          const int npass = SPLIT ? 2 - ((j & 1) == 0) : 1;   // SPLIT: lo box (even) one weight pass, hi box two (box_info)
#pragma unroll 1
          for (int ps = 0; ps < npass; ++ps) {
#pragma unroll 1
            for (int tap = 0; tap < S::SLABS; ++tap) {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              if (elect_one()) {
                const bool is_res = RES && tap == 9;        // the res_conv reads the centre-tap view
                const int ty = is_res ? 1 : tap / 3, tx = is_res ? 1 : tap - (tap / 3) * 3;
                constexpr uint32_t a_hi = desc_hi(S::BW * S::ROWB, S::LAYOUT), b_hi = desc_hi(8 * S::ROWB, S::LAYOUT);
                const uint32_t a_lo = desc_lo(a_base) + (uint32_t)(((ty * S::BW + tx) * S::ROWB) >> 4);
                const uint32_t b_lo = desc_lo(smem_u32(bring + bs * S::BSLAB));
                const uint32_t tdst = tacc + (is_res ? (uint32_t)(MT * NT) : 0u);
                const uint32_t first = is_res ? (uint32_t)(j | ps) : (uint32_t)(j | ps | tap);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                  for (int k = 0; k < S::KSTEPS; ++k)
                    umma_bf16_lohi(tdst + (uint32_t)(mt * NT), a_lo + (uint32_t)((mt * 8 * S::ROWB) >> 4) + (uint32_t)(k * 2), a_hi,
                                   b_lo + (uint32_t)(k * 2), b_hi, idesc, (first | (uint32_t)k) != 0);
                }
                umma_commit(&b_empty[bs]);                 // weight slab may be overwritten
                if (tap == S::SLABS - 1 && ps == npass - 1) {
                  umma_commit(&a_empty[as]);               // ... and so may the halo box
                  if (j == p.nchunk - 1) umma_commit(&tmem_full[slot]);
                }
              }
              __syncwarp();
              if (++bs == BSTG) { bs = 0; bph ^= 1; }
            }
          }
          if (++as == ASTG) { as = 0; aph ^= 1; }
        }
        if (++slot == 2) { slot = 0; sph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(DH_REGS_HIGH));
    const int q = warp & 3;                                // TMEM lane quadrant this warp may read
    const int stripe = (warp - DH_FIRST_EPI_WARP) >> 2;    // 64-column stripe of the item's 256 accumulator columns
    // RES: the first MT*NT columns are conv1 (tile mt, columns ncol0..), the next MT*NT the res_conv of the same tiles
    const bool is_res = RES && stripe * 64 >= MT * NT;
    const int local = stripe * 64 - (is_res ? MT * NT : 0);
    const int mt = local / NT, ncol0 = local % NT;
    const int r = q * 32 + lane;
    const int yy = r >> 3, xx = mt * 8 + (r & 7);
    const int et = threadIdx.x - 32 * DH_FIRST_EPI_WARP;
    SuperCursor cur; cur.init(it0, p.tiles_x, p.tiles_y);
    float s1 = 0.f, s2 = 0.f;
    int stat_img = -1;
    float rstd = 1.f;
    int slot = 0; uint32_t sph = 0;
    for (int it = it0; it < it1; ++it) {
      const int img = cur.img;
      if (img != stat_img) {
        if (p.dst_stats && stat_img >= 0) {
          const double d1 = warp_sum_d((double)s1), d2 = warp_sum_d((double)s2);
          if (lane == 0) { atomicAdd(p.dst_stats + 2 * stat_img, d1); atomicAdd(p.dst_stats + 2 * stat_img + 1, d2); }
        }
        stat_img = img; s1 = 0.f; s2 = 0.f;
        // all epilogue warps walk the same items, so they all rebuild the additive table at the same item
        asm volatile("bar.sync 1, %0;" ::"n"(32 * DH_EPI_WARPS) : "memory");
        const float half = (ACT && !SPLIT) ? 0.5f : 1.0f; // bf16: Swish works on x / 2 (tanh form)
        if (p.gn) {
          const GnScalars sc = gn_scalars(p.stats0, p.stats1, img, p.gn_count, p.eps);
          const float mri = sc.mean * sc.rstd;
          rstd = half * sc.rstd;
          for (int i = et; i < 9 * NT; i += 32 * DH_EPI_WARPS) ctab[i] = half * fmaf(-mri, __ldg(p.tg + i), __ldg(p.tb + i));
        } else {                                          // no GroupNorm in front: the additive term is the bias for every class
          rstd = half;
          for (int i = et; i < 9 * NT; i += 32 * DH_EPI_WARPS) ctab[i] = half * __ldg(p.tb + i % NT);
        }
        if (RES) for (int i = et; i < NT; i += 32 * DH_EPI_WARPS) btab2[i] = __ldg(p.tb2 + i);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * DH_EPI_WARPS) : "memory");
      }
      const int y = cur.ty * 16 + yy, x = cur.tx * S::SW + xx;
      const bool valid = y < p.H && x < p.W;
      const size_t pix = valid ? ((size_t)img * p.H + y) * p.W + x : 0;
      const int cls = (y == 0 ? 0 : (y == p.H - 1 ? 2 : 1)) * 3 + (x == 0 ? 0 : (x == p.W - 1 ? 2 : 1));
      // conv1 stripes: folded GroupNorm (+ Swish) into dst; res_conv stripes: + bias into dst2, no activation, no statistics
      const float4* ct = reinterpret_cast<const float4*>(is_res ? btab2 + ncol0 : ctab + cls * NT + ncol0);
      __nv_bfloat16* d = is_res ? p.dst2 + pix * p.dst2C + ncol0 : p.dst + pix * p.dstC + p.dstCoff + ncol0;
      const int d_lo = is_res ? p.dst2_lo : p.dst_lo;      // SPLIT: the lo plane of the destination row
      const bool st32 = is_res ? p.st32_2 != 0 : p.st32 != 0;
      const bool act_here = ACT && !is_res;
      const float rs = is_res ? 1.0f : rstd;               // (ACT: rstd / 2, and the table holds half the additive terms)
      const float2 rs2 = make_float2(rs, rs);
      float2 st1 = make_float2(0.f, 0.f), st2 = make_float2(0.f, 0.f);
      mbar_wait(&tmem_full[slot], sph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 256 + stripe * 64);
      // 16 columns per step; the TMEM load of step k+1 is in flight during the math of step k
      uint32_t rbuf[2][16];
      tmem_ld16(taddr, rbuf[0]);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < 3) tmem_ld16(taddr + 16 * (k + 1), rbuf[(k + 1) & 1]);
        if (valid) {
          const uint32_t* rv = rbuf[k & 1];
          uint32_t ow[8], lw[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 c4 = ct[4 * k + j];
            float2 va = __ffma2_rn(make_float2(__uint_as_float(rv[4 * j + 0]), __uint_as_float(rv[4 * j + 1])), rs2, make_float2(c4.x, c4.y));
            float2 vb = __ffma2_rn(make_float2(__uint_as_float(rv[4 * j + 2]), __uint_as_float(rv[4 * j + 3])), rs2, make_float2(c4.z, c4.w));
            if (act_here) {
              if (SPLIT) {                                // ex2.approx + rcp.approx: ~3e-7 relative against rtol 1e-3 / atol 1e-4
                va = make_float2(swish_fast(va.x), swish_fast(va.y));
                vb = make_float2(swish_fast(vb.x), swish_fast(vb.y));
              } else {
                va = __ffma2_rn(va, make_float2(tanh_approx(va.x), tanh_approx(va.y)), va);
                vb = __ffma2_rn(vb, make_float2(tanh_approx(vb.x), tanh_approx(vb.y)), vb);
              }
            }
            const __nv_bfloat162 ha = __floats2bfloat162_rn(va.x, va.y), hb = __floats2bfloat162_rn(vb.x, vb.y);
            ow[2 * j] = *reinterpret_cast<const uint32_t*>(&ha);
            ow[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&hb);
            if (SPLIT) {
              const float2 fa = __bfloat1622float2(ha), fb = __bfloat1622float2(hb);
              const __nv_bfloat162 la = __floats2bfloat162_rn(va.x - fa.x, va.y - fa.y), lb = __floats2bfloat162_rn(vb.x - fb.x, vb.y - fb.y);
              lw[2 * j] = *reinterpret_cast<const uint32_t*>(&la);
              lw[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&lb);
            }
            // statistics from the fp32 values (their bf16 rounding is zero-mean noise of relative size 2^-9)
            st1 = __fadd2_rn(st1, __fadd2_rn(va, vb));
            st2 = __ffma2_rn(va, va, __ffma2_rn(vb, vb, st2));
          }
          if (st32) {
            // one 256-bit store = one whole 32-byte sector per pixel (two 16-byte stores are two partial-sector writes at L2)
            st_global_v8(d + 16 * k, ow);
            if (SPLIT) st_global_v8(d + d_lo + 16 * k, lw);
          } else {
            *reinterpret_cast<uint4*>(d + 16 * k) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            *reinterpret_cast<uint4*>(d + 16 * k + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
            if (SPLIT) {
              *reinterpret_cast<uint4*>(d + d_lo + 16 * k) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              *reinterpret_cast<uint4*>(d + d_lo + 16 * k + 8) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
            }
          }
        }
        if (k < 3) tmem_ld_wait();
        if (k == 2) {
          // every accumulator value of this stripe is in registers: hand the slot back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[slot]);
        }
      }
      if (++slot == 2) { slot = 0; sph ^= 1; }
      if (!is_res) { s1 += st1.x + st1.y; s2 += st2.x + st2.y; }
      cur.next(p.tiles_x, p.tiles_y);
    }
    if (p.dst_stats && stat_img >= 0) {
      const double d1 = warp_sum_d((double)s1), d2 = warp_sum_d((double)s2);
      if (lane == 0) { atomicAdd(p.dst_stats + 2 * stat_img, d1); atomicAdd(p.dst_stats + 2 * stat_img + 1, d2); }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static const bool g_dh_pdl = []() { const char* e = getenv("UCDIR_PDL"); return !(e && e[0] == '0'); }();

template <int NT, int ACT, int KC, int RES, bool SPLIT = false>
static int launch_dh_inst(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& b2, const DhParams& p, int grid,
                          cudaStream_t st) {
  using S = DhCfg<NT, KC, RES>;
  static bool attr_dev[UCDIR_MAX_DEV] = {};
  bool& attr = attr_dev[cur_dev()];
  if (!attr) {
    if (int rc = check_reg_pool((const void*)dense_halo_kernel<NT, ACT, KC, RES, SPLIT>, "tc_dense_halo", 32 * DH_FIRST_EPI_WARP, DH_REGS_LOW, 32 * DH_EPI_WARPS, DH_REGS_HIGH)) return rc;
    if (cudaFuncSetAttribute(dense_halo_kernel<NT, ACT, KC, RES, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess) {
      set_error("tc_dense_halo: cannot opt in to %d bytes of shared memory: %s", S::TOTAL, cudaGetErrorString(cudaGetLastError())); return -3; }
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(DH_THREADS); cfg.dynamicSmemBytes = S::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = g_dh_pdl ? 1 : 0;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, dense_halo_kernel<NT, ACT, KC, RES, SPLIT>, a0, a1, b, b2, p) != cudaSuccess) {
    set_error("tc_dense_halo: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return -3; }
  return 0;
}

// true when the op (already validated by launch_tc_conv as a dense conv) fits this kernel
bool tc_dense_halo_applies(const ucdir_op_t& op) {
  const int C0 = op.i[UCDIR_TC_I_C0], C1 = op.i[UCDIR_TC_I_C1], H = op.i[UCDIR_TC_I_H], W = op.i[UCDIR_TC_I_W];
  const int NT = op.i[UCDIR_TC_I_NT];
  const int KB = op.i[UCDIR_TC_I_KB] ? op.i[UCDIR_TC_I_KB] : op.i[UCDIR_TC_I_KC];
  const int KC = op.i[UCDIR_TC_I_KC], gn = op.i[UCDIR_TC_I_GN];
  const bool chunks_ok = (KC == 64 && KB == 64 && C0 % 64 == 0 && C1 % 64 == 0) || (KC == 16 && KB == 16 && C0 == 16 && C1 == 0 && NT == 64);
  const bool gn_ok = gn == 1 ? (op.i[UCDIR_TC_I_NCLS] == 9 && op.p[UCDIR_TC_P_TG] && op.p[UCDIR_TC_P_STATS0] && (C1 == 0 || op.p[UCDIR_TC_P_STATS1]))
                             : (gn == 0);
  const bool split = op.i[UCDIR_TC_I_SPLIT] != 0;
  // SPLIT (fp32_tc): the layers the streamed split form serves badly -- the 16-channel in-conv (tiny slabs: the producer warp bounds it)
  // and the blocks whose 1x1 res_conv rides along (RES_FUSED); default plane layout [hi: C | lo: C]
  if (split && (!(KC == 16 || op.i[UCDIR_TC_I_RES_FUSED]) || (op.i[UCDIR_TC_I_SRC_LO_OFF] != 0 && op.i[UCDIR_TC_I_SRC_LO_OFF] != C0) || op.i[UCDIR_TC_I_W_LO_OFF] != 0)) return false;
  const int cs0 = split ? 2 * C0 : C0;
  return op.i[UCDIR_TC_I_HALO] == 1 && op.i[UCDIR_TC_I_MODE] == 0 && op.i[UCDIR_TC_I_GROUPS] == 1 && (NT == 64 || NT == 128) &&
         op.i[UCDIR_TC_I_NTOT] == NT && chunks_ok && gn_ok && op.i[UCDIR_TC_I_NTY] == 3 && op.i[UCDIR_TC_I_NTX] == 3 &&
         op.i[UCDIR_TC_I_OY0] == -1 && op.i[UCDIR_TC_I_OX0] == -1 && op.i[UCDIR_TC_I_STRIDE] == 1 && H >= 2 && W >= 2 &&
         op.i[UCDIR_TC_I_SRC_H] == H && op.i[UCDIR_TC_I_SRC_W] == W && !op.p[UCDIR_TC_P_RES] && !op.i[UCDIR_TC_I_DST_F32] &&
         !op.i[UCDIR_TC_I_DST_UP] && !op.i[UCDIR_TC_I_W_BATCHED] && !op.p[UCDIR_TC_P_DST2] && op.i[UCDIR_TC_I_ACT] <= 1 &&
         (op.i[UCDIR_TC_I_SRC_CSTRIDE] == 0 || op.i[UCDIR_TC_I_SRC_CSTRIDE] == cs0) && op.i[UCDIR_TC_I_DST_C] % 8 == 0 &&
         op.i[UCDIR_TC_I_DST_COFF] % 8 == 0 && (op.i[UCDIR_TC_I_NCOL_VALID] == 0 || op.i[UCDIR_TC_I_NCOL_VALID] == NT) &&
         (C1 == 0 || op.p[UCDIR_TC_P_SRC1]) &&
         (op.i[UCDIR_TC_I_RES_FUSED] == 0 || (KC == 64 && op.p[UCDIR_TC_P_W2] && op.p[UCDIR_TC_P_TB2] && op.p[UCDIR_TC_P_DST_RES] &&
                                              op.i[UCDIR_TC_I_DST_RES_C] >= NT && op.i[UCDIR_TC_I_DST_RES_C] % 8 == 0 && !op.i[UCDIR_TC_I_SRC_GN_SWISH]));
}

static int dh_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int B, int bw, int kc) {   // C: channels per pixel row (SPLIT: both planes)
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("tc_dense_halo: cuTensorMapEncodeTiled unavailable"); return -3; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)bw, 18, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("tc_dense_halo: cuTensorMapEncodeTiled(activation C=%d W=%d H=%d B=%d box %d) failed: %d", C, W, H, B, bw, (int)r); return -3; }
  return 0;
}

int launch_tc_dense_halo(const ucdir_op_t& op, cudaStream_t st) {
  DhParams p;
  const int C0 = op.i[UCDIR_TC_I_C0], C1 = op.i[UCDIR_TC_I_C1], NT = op.i[UCDIR_TC_I_NT], KC = op.i[UCDIR_TC_I_KC];
  p.gn = op.i[UCDIR_TC_I_GN];
  p.stats0 = (const double*)op.p[UCDIR_TC_P_STATS0]; p.stats1 = C1 ? (const double*)op.p[UCDIR_TC_P_STATS1] : nullptr;
  p.tb = (const float*)op.p[UCDIR_TC_P_TB]; p.tg = (const float*)op.p[UCDIR_TC_P_TG];
  p.dst = (__nv_bfloat16*)op.p[UCDIR_TC_P_DST]; p.dst_stats = (double*)op.p[UCDIR_TC_P_DST_STATS];
  p.B = op.i[UCDIR_TC_I_B]; p.H = op.i[UCDIR_TC_I_H]; p.W = op.i[UCDIR_TC_I_W];
  const bool split = op.i[UCDIR_TC_I_SPLIT] != 0;
  p.n0 = C0 / KC; p.n1 = C1 / KC;
  p.c0_chunks = (split ? 2 : 1) * p.n0; p.nchunk = (split ? 2 : 1) * (p.n0 + p.n1);
  p.a0_lo = C0; p.a1_lo = C1;
  p.ktap = (split ? 3 : 1) * (C0 + C1);
  p.act = op.i[UCDIR_TC_I_ACT]; p.dstC = op.i[UCDIR_TC_I_DST_C]; p.dstCoff = op.i[UCDIR_TC_I_DST_COFF];
  p.eps = op.f[UCDIR_TC_F_EPS];
  const int res = op.i[UCDIR_TC_I_RES_FUSED] ? 1 : 0;
  p.tb2 = (const float*)op.p[UCDIR_TC_P_TB2]; p.dst2 = (__nv_bfloat16*)op.p[UCDIR_TC_P_DST_RES]; p.dst2C = op.i[UCDIR_TC_I_DST_RES_C];
  p.dst_lo = p.dst2_lo = 0;
  if (split) { p.dst_lo = p.dstC; p.dstC *= 2; p.dst2_lo = p.dst2C; p.dst2C *= 2; }       // (hi, lo) plane pairs: rows twice as long
  p.st32 = (p.dstC % 16 == 0 && p.dstCoff % 16 == 0 && p.dst_lo % 16 == 0 && (reinterpret_cast<uintptr_t>(p.dst) & 31) == 0) ? 1 : 0;
  p.gn_count = (double)(C0 + C1) * p.H * p.W;
  p.st32_2 = (res && p.dst2C % 16 == 0 && p.dst2_lo % 16 == 0 && (reinterpret_cast<uintptr_t>(p.dst2) & 31) == 0) ? 1 : 0;
  const int sw = 8 * ((res ? 128 : 256) / NT);
  p.tiles_x = (p.W + sw - 1) / sw; p.tiles_y = (p.H + 15) / 16;
  const long long items = (long long)p.tiles_x * p.tiles_y * p.B;
  if (items > 0x7fffffffLL) { set_error("tc_dense_halo: too many items"); return -2; }
  p.n_items = (int)items;
  CUtensorMap a0, a1, mb, mb2;
  int rc = dh_act_map(&a0, op.p[UCDIR_TC_P_SRC0], split ? 2 * C0 : C0, p.W, p.H, p.B, sw + 2, KC);
  if (rc) return rc;
  if (C1 > 0) { rc = dh_act_map(&a1, op.p[UCDIR_TC_P_SRC1], split ? 2 * C1 : C1, p.W, p.H, p.B, sw + 2, KC); if (rc) return rc; }
  else a1 = a0;
  {
    EncodeTiledFn enc = get_encode();
    const int Ktot = 9 * p.ktap;
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)NT};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)NT};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(op.p[UCDIR_TC_P_W]), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tc_dense_halo: cuTensorMapEncodeTiled(weights K=%d N=%d) failed: %d", Ktot, NT, (int)r); return -3; }
  }
  mb2 = mb;
  if (res) {                                              // 1x1 res_conv weights [NT][C0 + C1], K-major (engine.py:pack_tc_dense)
    EncodeTiledFn enc = get_encode();
    const int Ktot = p.ktap;                                // SPLIT: [s0 W_hi | s0 W_hi | s1 W_hi | s1 W_hi | s0 W_lo | s1 W_lo]
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)NT};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)NT};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mb2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(op.p[UCDIR_TC_P_W2]), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tc_dense_halo: cuTensorMapEncodeTiled(res_conv weights K=%d N=%d) failed: %d", Ktot, NT, (int)r); return -3; }
  }
  const int n_sm = sm_count();
  const int grid = items < n_sm ? (int)items : n_sm;       // persistent: one CTA per SM
  const bool act = p.act == 1;
  if (split) {
    if (KC == 16) rc = act ? launch_dh_inst<64, 1, 16, 0, true>(a0, a1, mb, mb2, p, grid, st) : launch_dh_inst<64, 0, 16, 0, true>(a0, a1, mb, mb2, p, grid, st);
    else if (NT == 64) rc = act ? launch_dh_inst<64, 1, 64, 1, true>(a0, a1, mb, mb2, p, grid, st) : launch_dh_inst<64, 0, 64, 1, true>(a0, a1, mb, mb2, p, grid, st);
    else rc = act ? launch_dh_inst<128, 1, 64, 1, true>(a0, a1, mb, mb2, p, grid, st) : launch_dh_inst<128, 0, 64, 1, true>(a0, a1, mb, mb2, p, grid, st);
  } else if (KC == 16) rc = act ? launch_dh_inst<64, 1, 16, 0>(a0, a1, mb, mb2, p, grid, st) : launch_dh_inst<64, 0, 16, 0>(a0, a1, mb, mb2, p, grid, st);
  else if (NT == 64 && !res) rc = act ? launch_dh_inst<64, 1, 64, 0>(a0, a1, mb, mb2, p, grid, st) : launch_dh_inst<64, 0, 64, 0>(a0, a1, mb, mb2, p, grid, st);
  else if (NT == 64) rc = act ? launch_dh_inst<64, 1, 64, 1>(a0, a1, mb, mb2, p, grid, st) : launch_dh_inst<64, 0, 64, 1>(a0, a1, mb, mb2, p, grid, st);
  else if (!res) rc = act ? launch_dh_inst<128, 1, 64, 0>(a0, a1, mb, mb2, p, grid, st) : launch_dh_inst<128, 0, 64, 0>(a0, a1, mb, mb2, p, grid, st);
  else rc = act ? launch_dh_inst<128, 1, 64, 1>(a0, a1, mb, mb2, p, grid, st) : launch_dh_inst<128, 0, 64, 1>(a0, a1, mb, mb2, p, grid, st);
  if (rc) return rc;
  ++g_launches;
  return 0;
}

}  // namespace ucdir
