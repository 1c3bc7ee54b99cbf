// bf16 tensor-core path (sm_100a): implicit-GEMM convolution on tcgen05 with TMA-staged NHWC tiles -- the STREAMED form (one
// activation slab + one weight slab per filter tap and channel chunk).  launch_tc_conv routes three op families to dedicated
// halo-schedule kernels instead (ucdir_tc_schedule() tells which): the integration-module convs with C <= 256 (ucdir_mix.cu),
// the 3x3 convs with 64 / 128 output channels incl. the in-conv (ucdir_dhalo.cu) and final_conv with its GroupNorm + Swish
// (ucdir_fhalo.cu).  Everything else -- strided, 1x1, 2x2-phase upsample, >= 256-column and attention GEMMs -- runs here.
//
//   D[128 pixels x NT channels] (fp32, TMEM)  +=  A[128 pixels x KC] (bf16, smem)  *  B[NT x KC]^T (bf16, smem)
//
// * M tile = a bw x bh x bn box of output pixels (x, y, image); for filter tap (ty, tx) the A slab is ONE tiled
//   TMA box load of the NHWC activation at (x0*s + tx + ox0, y0*s + ty + oy0) -- out-of-bounds pixels are
//   zero-filled by the TMA unit, which is exactly the convolution's zero padding.  No im2col buffer.
// * K loop = taps x channel chunks of KC (64/32/16 channels = 128/64/32-byte swizzled rows, K-major).
//   torch.cat((x, skip)) is a second tensor map visited by the same loop; grouped convolution picks the
//   chunk(s) of its group; nearest-2x upsampling is four 2x2-tap phase convolutions (weights pre-summed).
// * GroupNorm(1,C) in front of a convolution is folded: gamma is multiplied into the packed weights and the
//   epilogue applies v = rstd*acc - mean*rstd*TG[cls][n] + TB[cls][n], where TG/TB are per-layer tables over
//   the 9 border classes (which taps fall outside the image) -- model/ucdir.py:109-112,161 never touch HBM.
// * persistent CTAs (one per SM) over work items (M tile, N sub-tile); warp roles: warp 0 = TMA producer, warp 1 = TMEM
//   allocator + tcgen05.mma issuer (whole warp runs the loop, one elect.sync lane issues), warps 4..15 = epilogue
//   (tcgen05.ld 32x32b, three warps per TMEM lane quadrant); setmaxnreg moves registers from warpgroup 0 to the
//   epilogue warpgroups.  smem ring (full/empty mbarriers) between producer and MMA; a ring of 512/NT TMEM
//   accumulators (tmem_full/tmem_empty) between MMA and epilogue, so the MMAs of item i+1 overlap the epilogue of i.
// * epilogues (compile-time EPI): plain + Swish + residual (optionally storing a column range transposed: attention
//   V^T), fp32 output (attention scores, eps), or the integration-module mix (model/ucdir.py:135-140): 8 adjacent
//   accumulator columns weighted by the per-pixel guidance map x per-step attw, Swish, + residual.
//   All emit the {sum, sum^2} of what they store for the next GroupNorm.
// * the same kernel is the batched GEMM of the attention core (model/ucdir.py:174,179): per-image "weights" through a
//   3-D tensor map (W_BATCHED), activation operand = a channel slice of a wider tensor (SRC_CSTRIDE).
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace ucdir {

// ------------------------------------------------------------------------------------------------
struct TcParams {
  const double* stats0; const double* stats1;
  const float* tb; const float* tg;
  const __nv_bfloat16* res; const float* att; const float* attw;
  void* dst; double* dst_stats;
  int B, H, W, srcH, srcW;
  int Ntot, ncol_valid;
  int bw, bh, bn, tiles_x, tiles_y, tiles_n, m_tiles, bres_bytes;
  int ctab;                    // mix epilogue: cache cadd[9][Ntot] of the current image in shared memory (bn == 1, small Ntot)
  int nty, ntx, oy0, ox0, stride;
  int nchunk, c0_chunks, groups, Ng, Cg, cg_eff;
  int gn, ncls, act, mode, dst_f32;
  int dstC, dstCoff, dstUp, dstPy, dstPx, resC, attwStride;
  double gn_count; float eps;
  float alpha;                 // scale on the accumulator when gn == 0 (attention: 1/sqrt(C))
  int w_batched;               // weights are per image: [B][Ntot][K] (attention K / V^T operands)
  __nv_bfloat16* dst2; int t_col0, t_ld;   // columns >= t_col0 are stored transposed: dst2[img][col - t_col0][pixel], row pitch t_ld
  // fp32-tolerance mode (UCDIR_TC_I_SPLIT): operands are (hi, lo) bf16 plane pairs, three K passes per tap
  int n0, n1;                  // chunks per plane of src0 / src1 (dense); grouped: n0 = chunks per pass
  int a0_lo, a1_lo;            // channel offset of the lo plane inside a pixel row of src0 / src1
  int b_lo;                    // batched weights: K offset of the lo plane inside a weight row
  int dst_lo, res_lo, t_lo;    // element offset of the lo plane inside a dst / residual / dst2 row
  // fused upsample phases (UCDIR_TC_I_PHASES = 4): the four output parities of nearest-2x + conv3x3 in one launch; the phase is an
  // extra, slowest work-item dimension folded into the image-tile index (tiles_n = phases * tiles_n_real)
  int phases, tiles_n_real;
  int io32;                    // bf16 destination (and residual) rows are 32-byte aligned per 16 columns: 256-bit stores / loads
};

// Which activation slab chunk j of a filter tap reads (map 0 / 1, channel coordinate) and, for batched weights, the K
// coordinate of its weight slab.  Plain mode: [src0 chunks | src1 chunks].  Split mode: [s0 hi | s0 lo | s1 hi | s1 lo | s0 hi | s1 hi]
// against weights [W_hi | W_hi | W_hi | W_hi | W_lo | W_lo]; grouped: [hi | lo | hi] of the group's chunk(s).
template <int KA, int KB, bool SPLIT>
__device__ __forceinline__ void chunk_source(const TcParams& p, int j, int cgrp0, int kb_seq, bool& use1, int& coff, int& kb) {
  kb = kb_seq;
  if (!SPLIT) {
    use1 = !(p.groups > 1 || j < p.c0_chunks);
    coff = use1 ? (j - p.c0_chunks) * KA : cgrp0 + j * KA;
    return;
  }
  use1 = false;
  if (p.groups > 1) {
    const int pass = j / p.n0, jj = j - pass * p.n0;
    coff = cgrp0 + jj * KA + (pass == 1 ? p.a0_lo : 0);
    return;
  }
  int q = j;
  if (q < 2 * p.n0) {
    const bool lo = q >= p.n0;
    if (lo) q -= p.n0;
    coff = q * KA + (lo ? p.a0_lo : 0);
    if (p.w_batched) kb = q * KB;
  } else if ((q -= 2 * p.n0) < 2 * p.n1) {
    use1 = true;
    const bool lo = q >= p.n1;
    if (lo) q -= p.n1;
    coff = q * KA + (lo ? p.a1_lo : 0);
  } else if ((q -= 2 * p.n1) < p.n0) {
    coff = q * KA;
    if (p.w_batched) kb = p.b_lo + q * KB;
  } else {
    use1 = true;
    coff = (q - p.n0) * KA;
  }
}

constexpr int TC_EPI_WARPS = 12;          // three per TMEM lane quadrant: the epilogue is latency bound, more warps hide it
constexpr int TC_FIRST_EPI_WARP = 4;      // warpgroup 0 = {TMA producer, MMA issuer, 2 idle warps}: shrinks to 72 registers
constexpr int TC_THREADS = 32 * (TC_FIRST_EPI_WARP + TC_EPI_WARPS);   // warpgroups 1..2 = epilogue: grow to 144 registers

// KA: channels per A slab row (64/32/16 -> 128/64/32-byte swizzled rows); KB: K elements per B slab row;
// NT: output columns per work item; NSPLIT: independent column groups of an item that read different K slices of
// the same A slab (grouped convolution with small groups: 4 groups of 64 columns share one 32-channel A slab).
// SPS: slabs (filter taps) per pipeline stage -- the small-N layers do ~130 cycles of MMA per slab but ~500 cycles of
// barrier round trip per stage, so they move three taps per stage.
template <int KA, int KB, int NT, int NSPLIT, int BSTAT = 0, int SPS = 1>
struct TcCfg {
  static constexpr int A_BYTES = 128 * KA * 2;
  static constexpr int B_BYTES = NT * KB * 2;
  static constexpr int B_PAD = (B_BYTES + 1023) & ~1023;
  static constexpr int SLAB = A_BYTES + (BSTAT ? 0 : B_PAD);    // weight-stationary: the ring holds activation slabs only
  // SPS == 4 ("ROW3"): one 130-pixel activation row serves the three horizontal taps of a filter row
  static constexpr bool ROW3 = (SPS == 4);
  static constexpr int A_ROW = (130 * KA * 2 + 1023) & ~1023;
  static constexpr int STAGE = ROW3 ? (A_ROW + 3 * B_PAD) : SPS * SLAB;
  static constexpr int STAGES_RAW = 196608 / STAGE;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : (STAGES_RAW < (SPS > 1 ? 2 : 4) ? (SPS > 1 ? 2 : 4) : STAGES_RAW);
  static_assert(STAGES * STAGE <= 200 * 1024, "smem ring too large");   // <= 8: the rest of the 228 KB stays L1 for the epilogue's loads
  static constexpr int SLOTW = NT < 32 ? 32 : NT;          // TMEM columns per accumulator slot
  static constexpr int NSLOT = 512 / SLOTW;                // accumulator ring: MMA of item i+1.. overlaps epilogue of item i
  static constexpr int BARS = (2 * STAGES + 2 * NSLOT + 2) * 8 + 64;
  static constexpr int TOTAL = STAGES * STAGE + 1024 /*align slack*/ + BARS;      // + the resident weight block when BSTAT
};

// Work-item cursor shared by the three roles: items are (M tile, N sub-tile) pairs, N fastest; advancing is
// increments and compares only (the producer issues one TMA pair per ~60 instructions, so divisions matter).
struct ItemCursor {
  int ns, tx_i, ty_i, tn_i;
  // MFAST = 0: item = m * n_sub + ns (N fastest: consecutive items share the activation tile).
  // MFAST = 1: item = ns * m_tiles + m (M fastest: consecutive items share the weight block -> weight-stationary).
  template <int MFAST>
  __device__ __forceinline__ void init(long long it, int n_sub, int m_tiles, int tiles_x, int tiles_y) {
    long long m;
    if (MFAST) { ns = (int)(it / m_tiles); m = it - (long long)ns * m_tiles; }
    else { m = it / n_sub; ns = (int)(it - m * n_sub); }
    tx_i = (int)(m % tiles_x);
    const long long t = m / tiles_x;
    ty_i = (int)(t % tiles_y);
    tn_i = (int)(t / tiles_y);
  }
  // returns true when the M tile changed
  template <int MFAST>
  __device__ __forceinline__ bool next(int n_sub, int tiles_x, int tiles_y, int tiles_n) {
    if (!MFAST) {
      if (++ns < n_sub) return false;
      ns = 0;
    }
    if (++tx_i == tiles_x) { tx_i = 0; if (++ty_i == tiles_y) { ty_i = 0; ++tn_i; } }
    if (MFAST && tn_i == tiles_n) { tn_i = 0; ++ns; }
    return true;
  }
};

// Persistent CTA (one per SM).  A CTA owns a contiguous range of work items.  Three pipelines:
//   smem ring   full[s]/empty[s]            TMA producer  <-> MMA issuer
//   TMEM ring   tmem_full[j]/tmem_empty[j]  MMA issuer    <-> epilogue warps   (NSLOT accumulators of NT columns)
// EPI selects the epilogue at compile time (the chunk loop is the hot code of the small-K layers):
enum { EPI_PLAIN = 0, EPI_MIX = 1, EPI_F32 = 2, EPI_PLAIN_T = 3, EPI_MIXC = 4 /* mix + additive table cached in smem */ };

template <int KA, int KB, int NT, int NSPLIT, int EPI, int BSTAT, int SPS, bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_conv_kernel(const __grid_constant__ CUtensorMap mapA0,
                                                                const __grid_constant__ CUtensorMap mapA1,
                                                                const __grid_constant__ CUtensorMap mapB, const TcParams p) {
  constexpr bool MIX = (EPI == EPI_MIX || EPI == EPI_MIXC);
  constexpr bool CTAB = (EPI == EPI_MIXC);
  using S = TcCfg<KA, KB, NT, NSPLIT, BSTAT, SPS>;
  constexpr int STAGES = S::STAGES, NSLOT = S::NSLOT;
  constexpr int BSLAB = NT * KB * 2;                  // one weight slab (all NT rows of one K slice)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bres = smem + STAGES * S::STAGE;           // BSTAT: resident weight block of the current N sub-tile
  uint64_t* full = reinterpret_cast<uint64_t*>(bres + (BSTAT ? p.bres_bytes : 0));
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* tmem_empty = tmem_full + NSLOT;
  uint64_t* bfull = tmem_empty + NSLOT;               // BSTAT: producer -> MMA, weight block landed
  uint64_t* bfree = bfull + 1;                        // BSTAT: MMA -> producer, all MMAs that read the old block retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfree + 1);
  float* ctab = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~uintptr_t(15));   // [9][Ntot] when p.ctab

  // Programmatic dependent launch: let the next kernel of the stream be scheduled while this one runs (its prologue
  // -- barrier init, TMEM allocation, tensor-map prefetch -- then overlaps our tail); it blocks in griddepcontrol.wait
  // until this grid has completed and flushed, so no data hazard is introduced.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_sub = p.Ntot / NT;
  const long long total = (long long)p.m_tiles * n_sub;
  const long long it0 = total * blockIdx.x / gridDim.x, it1 = total * (blockIdx.x + 1) / gridDim.x;
  const int n_items = (int)(it1 - it0);

  if (threadIdx.x == 0) {
    prefetch_tmap(&mapA0); prefetch_tmap(&mapB);
    if ((SPLIT ? p.n1 > 0 : p.c0_chunks < p.nchunk) && p.groups == 1) prefetch_tmap(&mapA1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int j = 0; j < NSLOT; ++j) { mbar_init(&tmem_full[j], 1); mbar_init(&tmem_empty[j], TC_EPI_WARPS); }
    mbar_init(bfull, 1); mbar_init(bfree, 1);
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
  // Register reallocation between warpgroups (setmaxnreg is warpgroup-wide, first statement of each role branch):
  // the epilogue keeps a chunk of accumulators, its folded-GroupNorm terms and the next chunk's table values in flight.
  if (warp < TC_FIRST_EPI_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop; one elected lane issues) =====================
    const uint32_t tx_bytes = (uint32_t)(p.bw * p.bh * p.bn * KA * 2 + (BSTAT ? 0 : NT * KB * 2));
    const int nslab_p = p.nty * p.ntx * p.nchunk;
    ItemCursor cur; cur.template init<BSTAT>(it0, n_sub, p.m_tiles, p.tiles_x, p.tiles_y);
    int stage = 0; uint32_t phase = 0;
    int cur_ns = -1; uint32_t bfree_phase = 0;
    for (int li = 0; li < n_items; ++li) {
      int tn = cur.tn_i, ph = 0;
      if (p.phases > 1) { ph = tn / p.tiles_n_real; tn -= ph * p.tiles_n_real; }
      const int x0 = cur.tx_i * p.bw * p.stride + p.ox0 + (ph & 1), y0 = cur.ty_i * p.bh * p.stride + p.oy0 + (ph >> 1), n0 = tn * p.bn;
      const int ncol0 = cur.ns * NT;
      const int brow0 = ncol0 + ph * p.Ntot;           // weight rows: the phases' blocks are stacked along N
      const int cgrp0 = p.groups > 1 ? ((ncol0 / p.Ng) * p.Cg) / p.cg_eff * p.cg_eff : 0;
      if (BSTAT && cur.ns != cur_ns) {
        // new N sub-tile: (re)load its whole weight block once; every following item streams activations only
        if (cur_ns >= 0) { mbar_wait(bfree, bfree_phase); bfree_phase ^= 1; }
        if (elect_one()) {
          mbar_expect_tx(bfull, (uint32_t)(nslab_p * BSLAB));
          for (int i = 0; i < nslab_p; ++i) tma_load_2d(&mapB, bfull, bres + i * BSLAB, i * KB, ncol0);
        }
        __syncwarp();
        cur_ns = cur.ns;
      }
      if (S::ROW3) {
        // one stage = one filter row (ty) x one channel chunk: a 130-pixel activation row + the three taps' weights
        const uint32_t row_bytes = (uint32_t)((p.bw + 2) * KA * 2 + 3 * NT * KB * 2);
        for (int ty = 0; ty < 3; ++ty) {
          for (int j = 0; j < p.nchunk; ++j) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
              uint8_t* sa = smem + stage * S::STAGE;
              mbar_expect_tx(&full[stage], row_bytes);
              bool use1; int coff, kbx;
              chunk_source<KA, KB, SPLIT>(p, j, 0, 0, use1, coff, kbx);
              tma_load_4d(use1 ? &mapA1 : &mapA0, &full[stage], sa, coff, x0, y0 + ty, n0);
#pragma unroll
              for (int tx = 0; tx < 3; ++tx)
                tma_load_2d(&mapB, &full[stage], sa + S::A_ROW + tx * S::B_PAD, ((ty * 3 + tx) * p.nchunk + j) * KB, ncol0);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        cur.template next<BSTAT>(n_sub, p.tiles_x, p.tiles_y, p.tiles_n);
        continue;
      }
      int kb = 0, ty = 0, tx = 0, j = 0;                // K coordinate of the B slab; tap and channel chunk of the slab
      for (int sidx = 0; sidx < nslab_p; ++sidx) {
        const int sub = sidx % SPS;
        if (sub == 0) mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * S::STAGE + sub * S::SLAB;
          if (sub == 0) mbar_expect_tx(&full[stage], tx_bytes * SPS);
          bool use1; int coff, kbx;
          chunk_source<KA, KB, SPLIT>(p, j, cgrp0, kb, use1, coff, kbx);
          tma_load_4d(use1 ? &mapA1 : &mapA0, &full[stage], sa, coff, x0 + tx, y0 + ty, n0);
          if (!BSTAT) {
            if (p.w_batched) tma_load_3d(&mapB, &full[stage], sa + S::A_BYTES, kbx, ncol0, n0);
            else tma_load_2d(&mapB, &full[stage], sa + S::A_BYTES, kbx, brow0);
          }
        }
        __syncwarp();
        kb += KB;
        if (++j == p.nchunk) { j = 0; if (++tx == p.ntx) { tx = 0; ++ty; } }
        if (sub == SPS - 1) { if (++stage == STAGES) { stage = 0; phase ^= 1; } }
      }
      cur.template next<BSTAT>(n_sub, p.tiles_x, p.tiles_y, p.tiles_n);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs the loop; one elected lane issues) =====================
    constexpr int NSUB = NT / NSPLIT;                   // columns per MMA
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NSUB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr int KSTEPS = KB / 16;
    constexpr int CGS = KA / NSPLIT;                    // channels of the A slab that belong to one split
    const int nslab = p.nty * p.ntx * p.nchunk;
    int stage = 0; uint32_t phase = 0;
    int slot = 0; uint32_t sph = 0;
    ItemCursor cur; cur.template init<BSTAT>(it0, n_sub, p.m_tiles, p.tiles_x, p.tiles_y);
    int cur_ns = -1; uint32_t bfull_phase = 0;
    for (int li = 0; li < n_items; ++li) {
      if (BSTAT && cur.ns != cur_ns) { mbar_wait(bfull, bfull_phase); bfull_phase ^= 1; cur_ns = cur.ns; }
      const int ns_now = cur.ns;
      cur.template next<BSTAT>(n_sub, p.tiles_x, p.tiles_y, p.tiles_n);
      const bool last_of_block = BSTAT && (cur.ns != ns_now) && (li + 1 < n_items);
      mbar_wait(&tmem_empty[slot], sph ^ 1);            // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(slot * S::SLOTW);
      const int nstage_item = S::ROW3 ? 3 * p.nchunk : nslab / SPS;
      for (int g = 0; g < nstage_item; ++g) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (S::ROW3) {
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * S::STAGE);
#pragma unroll
            for (int tx = 0; tx < 3; ++tx) {
              const uint64_t ad = make_desc_shifted(sa, (uint32_t)tx);
              const uint64_t bd = make_desc(sa + S::A_ROW + tx * S::B_PAD, KB * 2);
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                umma_bf16(tacc, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (g | tx | k) != 0);
            }
            umma_commit(&empty[stage]);
            if (g == nstage_item - 1) umma_commit(&tmem_full[slot]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          continue;
        }
        if (elect_one()) {
#pragma unroll
          for (int sub = 0; sub < (S::ROW3 ? 1 : SPS); ++sub) {
            const int i = g * SPS + sub;                 // slab index within the item
            const uint32_t sa = smem_u32(smem + stage * S::STAGE + sub * S::SLAB);
            const uint32_t sb = BSTAT ? smem_u32(bres + i * BSLAB) : sa + S::A_BYTES;
            const uint64_t ad = make_desc(sa, KA * 2), bd = make_desc(sb, KB * 2);
#pragma unroll
            for (int sp = 0; sp < NSPLIT; ++sp) {
              // split sp reads the 16-element K slice that holds its CGS channels, and its own rows of the B slab
              const uint32_t a_off = NSPLIT > 1 ? (uint32_t)((sp * CGS) / 16) * 32u : 0u;
              const uint32_t b_off = (uint32_t)(sp * NSUB * KB * 2);
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k)
                umma_bf16(tacc + (uint32_t)(sp * NSUB), ad + (uint64_t)((a_off + k * 32) >> 4), bd + (uint64_t)((b_off + k * 32) >> 4), idesc,
                          (i | k) != 0);
            }
          }
          umma_commit(&empty[stage]);          // frees the smem stage once these MMAs have read it
          if (g == nstage_item - 1) {
            umma_commit(&tmem_full[slot]);     // accumulator of this item complete
            if (last_of_block) umma_commit(bfree);   // ... and the weight block may be overwritten
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++slot == NSLOT) { slot = 0; sph ^= 1; }
    }
  }
  } else {
    // ===================== epilogue =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
    const int q = warp & 3;                          // TMEM lane quadrant this warp may read
    const int half = (warp - TC_FIRST_EPI_WARP) >> 2;  // the warps of a quadrant take alternate column chunks
    const int r = q * 32 + lane;                     // accumulator row = pixel slot of the tile
    const int box = p.bw * p.bh;
    const int nn = r / box, rr = r - nn * box;
    const int yy = rr / p.bw, xx = rr - yy * p.bw;
    // GroupNorm statistics of the stored tensor: when every lane of a warp works on the same image (one image per tile, or whole
    // 32-row groups per image) the per-thread sums live in registers across items and are reduced when the tile's images change
    // (the <= 8x8-pixel levels stack two images per tile: a warp reduction in double per item was 10 % of those launches)
    const bool wu_stats = p.bn == 1 || (box & 31) == 0;
    constexpr int CH = NT < 32 ? 16 : 32;
    ItemCursor cur; cur.template init<BSTAT>(it0, n_sub, p.m_tiles, p.tiles_x, p.tiles_y);
    bool new_m = true;
    bool valid = false;
    int img = 0, y = 0, x = 0, cls = 0;
    float rstd = 1.f, mr = 0.f;
    size_t pix_in = 0, pix_out = 0;
    const __nv_bfloat16* res_row = nullptr;          // residual row of this thread's pixel
    uint8_t* dst_row = nullptr;                      // output row of this thread's pixel (bf16 or fp32 elements)
    float aw[8];
    float s1 = 0.f, s2 = 0.f;                        // GroupNorm statistics of what this thread stored, current image
    int stat_img = -1;
    int gn_img = -1; float gn_rstd = 1.f, gn_mr = 0.f;   // GroupNorm scalars of the last image seen (FP64 math, cached)
    int ctab_img = -1;
    int slot = 0; uint32_t sph = 0;
    for (int li = 0; li < n_items; ++li) {
      const int ncol0 = cur.ns * NT;
      if (new_m) {                                    // per-pixel state changes only with the M tile
        int tn = cur.tn_i, ph = 0;
        if (p.phases > 1) { ph = tn / p.tiles_n_real; tn -= ph * p.tiles_n_real; }
        const int im0 = tn * p.bn;
        if (p.dst_stats && wu_stats && im0 != stat_img) {
          if (stat_img >= 0) {
            const double d1 = warp_sum_d((double)s1), d2 = warp_sum_d((double)s2);
            if (lane == 0 && stat_img + nn < p.B) { atomicAdd(p.dst_stats + 2 * (stat_img + nn), d1); atomicAdd(p.dst_stats + 2 * (stat_img + nn) + 1, d2); }
          }
          stat_img = im0; s1 = 0.f; s2 = 0.f;
        }
        img = im0 + nn; y = cur.ty_i * p.bh + yy; x = cur.tx_i * p.bw + xx;
        valid = (r < box * p.bn) && img < p.B && y < p.H && x < p.W;
        rstd = p.alpha; mr = 0.f; cls = 0;
        if (!valid) { img = 0; y = 0; x = 0; }
        if (p.gn && valid) {
          if (img != gn_img) {
            GnScalars sc = gn_scalars(p.stats0, p.stats1, img, p.gn_count, p.eps);
            gn_img = img; gn_rstd = sc.rstd; gn_mr = sc.mean * sc.rstd;
          }
          rstd = gn_rstd; mr = gn_mr;
          if (p.ncls == 9) cls = (y == 0 ? 0 : (y == p.H - 1 ? 2 : 1)) * 3 + (x == 0 ? 0 : (x == p.W - 1 ? 2 : 1));
        }
        pix_in = ((size_t)img * p.H + y) * p.W + x;
        pix_out = p.dstUp ? ((size_t)img * 2 * p.H + 2 * y + p.dstPy + (ph >> 1)) * (2 * p.W) + 2 * x + p.dstPx + (ph & 1) : pix_in;
        res_row = p.res ? p.res + pix_in * p.resC : nullptr;
        dst_row = reinterpret_cast<uint8_t*>(p.dst) + (pix_out * p.dstC + p.dstCoff) * (EPI == EPI_F32 ? 4 : 2);
        if (MIX) {
          const float4 t0 = __ldg(reinterpret_cast<const float4*>(p.att + pix_in * 8));
          const float4 t1 = __ldg(reinterpret_cast<const float4*>(p.att + pix_in * 8 + 4));
          const float* w8 = p.attw + (size_t)img * p.attwStride;
          aw[0] = t0.x * __ldg(w8 + 0); aw[1] = t0.y * __ldg(w8 + 1); aw[2] = t0.z * __ldg(w8 + 2); aw[3] = t0.w * __ldg(w8 + 3);
          aw[4] = t1.x * __ldg(w8 + 4); aw[5] = t1.y * __ldg(w8 + 5); aw[6] = t1.z * __ldg(w8 + 6); aw[7] = t1.w * __ldg(w8 + 7);
        }
      }
      if (CTAB) {
        // all epilogue warps walk the same items, so they all see the image change at the same item
        const int im0 = cur.tn_i * p.bn;
        if (im0 != ctab_img) {
          ctab_img = im0;
          asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");       // everyone is done with the old table
          const GnScalars sc = gn_scalars(p.stats0, p.stats1, im0 < p.B ? im0 : 0, p.gn_count, p.eps);
          const float mri = sc.mean * sc.rstd;
          const int et = threadIdx.x - 32 * TC_FIRST_EPI_WARP;
          for (int i = et * 4; i < 9 * p.Ntot; i += 32 * TC_EPI_WARPS * 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.tb + i));
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.tg + i));
            *reinterpret_cast<float4*>(ctab + i) = make_float4(fmaf(-mri, g.x, b.x), fmaf(-mri, g.y, b.y), fmaf(-mri, g.z, b.z), fmaf(-mri, g.w, b.w));
          }
          asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
        }
      }
      const float* tb = p.tb + (size_t)cls * p.Ntot + ncol0;
      const float* tg = p.tg ? p.tg + (size_t)cls * p.Ntot + ncol0 : nullptr;
      const float* ctab_row = ctab + (size_t)cls * p.Ntot + ncol0;
      float t1s = 0.f, t2s = 0.f;
      // Per-column additive term of the folded GroupNorm, cadd[j] = TB[cls][n] - mean*rstd*TG[cls][n]: the table loads
      // are issued one chunk ahead (and, for the first chunk, before waiting for the accumulator) so their L2 latency
      // overlaps the MMA wait / the previous chunk's math instead of stalling every chunk.
      constexpr int CSTEP = (TC_EPI_WARPS / 4) * CH;
      float cadd[CH];
      float4 tb4[CH / 4], tg4[CH / 4];
      auto issue_tables = [&](int c0) {
        if (CTAB) return;
#pragma unroll
        for (int j = 0; j < CH / 4; ++j) {
          tb4[j] = __ldg(reinterpret_cast<const float4*>(tb + c0) + j);
          tg4[j] = tg ? __ldg(reinterpret_cast<const float4*>(tg + c0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto finish_tables = [&](int c0) {
        if (CTAB) {
#pragma unroll
          for (int j = 0; j < CH / 4; ++j) {
            const float4 t = *reinterpret_cast<const float4*>(ctab_row + c0 + 4 * j);
            cadd[4 * j + 0] = t.x; cadd[4 * j + 1] = t.y; cadd[4 * j + 2] = t.z; cadd[4 * j + 3] = t.w;
          }
          return;
        }
        const float2 nm = make_float2(-mr, -mr);
#pragma unroll
        for (int j = 0; j < CH / 4; ++j) {
          const float2 lo = __ffma2_rn(nm, make_float2(tg4[j].x, tg4[j].y), make_float2(tb4[j].x, tb4[j].y));
          const float2 hi = __ffma2_rn(nm, make_float2(tg4[j].z, tg4[j].w), make_float2(tb4[j].z, tb4[j].w));
          cadd[4 * j + 0] = lo.x; cadd[4 * j + 1] = lo.y; cadd[4 * j + 2] = hi.x; cadd[4 * j + 3] = hi.y;
        }
      };
      uint2 res_next = make_uint2(0u, 0u);            // mix: residual of the next chunk, loaded one chunk ahead
      uint2 res_next_lo = make_uint2(0u, 0u);         // SPLIT: its lo plane
      auto issue_res = [&](int c0) {
        constexpr int NO_ = CH / 8;
        const __nv_bfloat16* rp = res_row + ((ncol0 + c0) >> 3);
        if (NO_ == 4) res_next = __ldg(reinterpret_cast<const uint2*>(rp));
        else res_next.x = __ldg(reinterpret_cast<const uint32_t*>(rp));
        if (SPLIT) {
          if (NO_ == 4) res_next_lo = __ldg(reinterpret_cast<const uint2*>(rp + p.res_lo));
          else res_next_lo.x = __ldg(reinterpret_cast<const uint32_t*>(rp + p.res_lo));
        }
      };
      if (valid && half * CH < NT) { issue_tables(half * CH); if (MIX) issue_res(half * CH); finish_tables(half * CH); }
      mbar_wait(&tmem_full[slot], sph);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = half * CH; c0 < NT; c0 += CSTEP) {
        // issue the residual load this chunk needs before waiting on the TMEM load
        constexpr int NO = CH / 8;
        uint2 res_mix = make_uint2(0u, 0u), res_mix_lo = make_uint2(0u, 0u);
        uint4 res_pl[CH / 8];
        uint4 res_pl_lo[SPLIT ? CH / 8 : 1];
        if (valid) {
          if (MIX) {
            res_mix = res_next; res_mix_lo = res_next_lo;
          } else if (EPI != EPI_F32 && p.res) {
            const uint4* rp = reinterpret_cast<const uint4*>(res_row + ncol0 + c0);
            if (CH % 16 == 0 && p.io32) {
#pragma unroll
              for (int j = 0; j < CH / 8; j += 2) ld_global_nc_v8(rp + j, reinterpret_cast<uint32_t*>(&res_pl[j]));
            } else {
#pragma unroll
              for (int j = 0; j < CH / 8; ++j) res_pl[j] = __ldg(rp + j);
            }
            if (SPLIT) {
              const uint4* rl = reinterpret_cast<const uint4*>(res_row + p.res_lo + ncol0 + c0);
              if (CH % 16 == 0 && p.io32) {
#pragma unroll
                for (int j = 0; j < CH / 8; j += 2) ld_global_nc_v8(rl + j, reinterpret_cast<uint32_t*>(&res_pl_lo[j]));
              } else {
#pragma unroll
                for (int j = 0; j < CH / 8; ++j) res_pl_lo[j] = __ldg(rl + j);
              }
            }
          }
        }
        uint32_t rv[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * S::SLOTW + c0);
#if defined(UCDIR_ABLATE) && UCDIR_ABLATE == 2      // ablation: no TMEM loads at all
#pragma unroll
        for (int j = 0; j < 32; ++j) rv[j] = 0u;
        (void)taddr;
#else
        if (CH == 32) tmem_ld32(taddr, rv); else tmem_ld16(taddr, rv);
        tmem_ld_wait();
#endif
#if defined(UCDIR_ABLATE) && (UCDIR_ABLATE == 1 || UCDIR_ABLATE == 2)   // ablation: drain TMEM, skip all epilogue math / IO
        if (MIX) { if (rv[0] == 0x7fc12345u) t1s += 1.f; continue; }
#endif
        if (!valid) continue;
        float v[CH];
        {
          const float2 rs2 = make_float2(rstd, rstd);
#pragma unroll
          for (int j = 0; j < CH; j += 2) {     // packed fp32 pairs (FFMA2)
            const float2 t = __ffma2_rn(make_float2(__uint_as_float(rv[j]), __uint_as_float(rv[j + 1])), rs2, make_float2(cadd[j], cadd[j + 1]));
            v[j] = t.x; v[j + 1] = t.y;
          }
        }
        const bool more = c0 + CSTEP < NT;
        if (more && MIX) issue_res(c0 + CSTEP);   // next chunk's residual in flight while this chunk's math runs
        if (MIX) {
          // integration-module mix: 8 adjacent columns (c*8+s) -> channel c   (model/ucdir.py:136-140)
          const int cbase = (ncol0 + c0) >> 3;
          __align__(8) __nv_bfloat16 o[NO];
          __align__(8) __nv_bfloat16 ol[NO];
          const __nv_bfloat16* rres = reinterpret_cast<const __nv_bfloat16*>(&res_mix);
          const __nv_bfloat16* rres_lo = reinterpret_cast<const __nv_bfloat16*>(&res_mix_lo);
#pragma unroll
          for (int c = 0; c < NO; ++c) {
            float2 h2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int s = 0; s < 8; s += 2) h2 = __ffma2_rn(make_float2(v[c * 8 + s], v[c * 8 + s + 1]), make_float2(aw[s], aw[s + 1]), h2);
            const float h = h2.x + h2.y;
            if (SPLIT) {
              const float t = swish_fast(h) + (__bfloat162float(rres[c]) + __bfloat162float(rres_lo[c]));   // ex2 / rcp.approx: ~3e-7 relative
              o[c] = __float2bfloat16(t);
              ol[c] = __float2bfloat16(t - __bfloat162float(o[c]));
              t1s += t; t2s += t * t;
            } else {
              const float t = swish_fast(h) + __bfloat162float(rres[c]);
              o[c] = __float2bfloat16(t);
              const float tr = __bfloat162float(o[c]);
              t1s += tr; t2s += tr * tr;
            }
          }
          __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst_row) + cbase;
          if (NO == 4) *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(o);
          else *reinterpret_cast<uint32_t*>(d) = *reinterpret_cast<const uint32_t*>(o);
          if (SPLIT) {
            if (NO == 4) *reinterpret_cast<uint2*>(d + p.dst_lo) = *reinterpret_cast<const uint2*>(ol);
            else *reinterpret_cast<uint32_t*>(d + p.dst_lo) = *reinterpret_cast<const uint32_t*>(ol);
          }
        } else {
          const int nb = ncol0 + c0;
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = swish_fast(v[j]);   // ex2.approx + rcp.approx (~3e-7 relative: ample for SPLIT's 1e-3 / 1e-4 too)
          } else if (p.act == 2) {              // LeakyReLU(0.2) = max(0.2 x, x), model/ucdir.py:414-416 (predictor)
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = fmaxf(0.2f * v[j], v[j]);
          }
          if (EPI != EPI_F32 && p.res) {
#pragma unroll
            for (int j = 0; j < CH; j += 8) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&res_pl[j / 8]);
              const __nv_bfloat162* l2 = reinterpret_cast<const __nv_bfloat162*>(&res_pl_lo[SPLIT ? j / 8 : 0]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 f = __bfloat1622float2(h2[e]);
                if (SPLIT) { const float2 g = __bfloat1622float2(l2[e]); f.x += g.x; f.y += g.y; }
                v[j + 2 * e] += f.x; v[j + 2 * e + 1] += f.y;
              }
            }
          }
          if (EPI == EPI_F32) {
            float* d = reinterpret_cast<float*>(dst_row) + nb;
            if (nb + CH <= p.ncol_valid && (p.dstC & 3) == 0 && (p.dstCoff & 3) == 0) {
              // full chunk: 16-byte stores (the attention scores: each thread writes 128 contiguous bytes of its row)
#pragma unroll
              for (int j = 0; j < CH; j += 4) {
                *reinterpret_cast<float4*>(d + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                t1s += (v[j] + v[j + 1]) + (v[j + 2] + v[j + 3]);
                t2s += (v[j] * v[j] + v[j + 1] * v[j + 1]) + (v[j + 2] * v[j + 2] + v[j + 3] * v[j + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (nb + j < p.ncol_valid) { d[j] = v[j]; t1s += v[j]; t2s += v[j] * v[j]; }
            }
          } else if (EPI == EPI_PLAIN_T && nb >= p.t_col0) {
            // transposed store (attention V^T): lanes hold consecutive pixels, so each store is one coalesced run
            __nv_bfloat16* d = p.dst2 + ((size_t)img * (p.Ntot - p.t_col0) + (nb - p.t_col0)) * p.t_ld + (y * p.W + x);
#pragma unroll
            for (int j = 0; j < CH; ++j) {
              const __nv_bfloat16 hi = __float2bfloat16(v[j]);
              d[(size_t)j * p.t_ld] = hi;
              if (SPLIT) d[(size_t)j * p.t_ld + p.t_lo] = __float2bfloat16(v[j] - __bfloat162float(hi));
            }
          } else {
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst_row) + nb;
            constexpr int SW = (CH % 16 == 0 && !SPLIT) ? 16 : 8;   // columns per store: 256-bit where the rows allow it (p.io32; SPLIT: register bound)
#pragma unroll
            for (int j = 0; j < CH; j += SW) {
              uint32_t o2[SW / 2], l2[SW / 2];
#pragma unroll
              for (int e = 0; e < SW / 2; ++e) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(v[j + 2 * e], v[j + 2 * e + 1]);
                o2[e] = *reinterpret_cast<const uint32_t*>(&h);
                const float2 f = __bfloat1622float2(h);
                if (SPLIT) {
                  const float a = v[j + 2 * e], b = v[j + 2 * e + 1];
                  const __nv_bfloat162 l = __floats2bfloat162_rn(a - f.x, b - f.y);
                  l2[e] = *reinterpret_cast<const uint32_t*>(&l);
                  t1s += a + b; t2s += a * a + b * b;
                } else {
                  t1s += f.x + f.y; t2s += f.x * f.x + f.y * f.y;
                }
              }
              if (SW == 16 && p.io32) {
                st_global_v8(d + j, o2);
                if (SPLIT) st_global_v8(d + p.dst_lo + j, l2);
              } else {
#pragma unroll
                for (int e = 0; e < SW / 8; ++e) {
                  *reinterpret_cast<uint4*>(d + j + 8 * e) = make_uint4(o2[4 * e], o2[4 * e + 1], o2[4 * e + 2], o2[4 * e + 3]);
                  if (SPLIT) *reinterpret_cast<uint4*>(d + p.dst_lo + j + 8 * e) = make_uint4(l2[4 * e], l2[4 * e + 1], l2[4 * e + 2], l2[4 * e + 3]);
                }
              }
            }
          }
        }
        if (more) { issue_tables(c0 + CSTEP); finish_tables(c0 + CSTEP); }   // next chunk's additive terms (short-lived registers)
      }
      // this warp is done reading the accumulator slot: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[slot]);
      if (++slot == NSLOT) { slot = 0; sph ^= 1; }
      if (p.dst_stats) {
        if (wu_stats) { s1 += t1s; s2 += t2s; }       // flushed when the tile's images change / at the end
        else if (valid) { atomicAdd(p.dst_stats + 2 * img, (double)t1s); atomicAdd(p.dst_stats + 2 * img + 1, (double)t2s); }
      }
      new_m = cur.template next<BSTAT>(n_sub, p.tiles_x, p.tiles_y, p.tiles_n);
    }
    if (p.dst_stats && wu_stats && stat_img >= 0) {
      const double d1 = warp_sum_d((double)s1), d2 = warp_sum_d((double)s2);
      if (lane == 0 && stat_img + nn < p.B) { atomicAdd(p.dst_stats + 2 * (stat_img + nn), d1); atomicAdd(p.dst_stats + 2 * (stat_img + nn) + 1, d2); }
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
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static CUtensorMapSwizzle swz(int kc) { return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B); }

static int make_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int B, int kc, int bw, int bh, int bn, int stride, int cstride) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("tc_conv: cuTensorMapEncodeTiled unavailable"); return -3; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)cstride * 2, (cuuint64_t)cstride * 2 * W, (cuuint64_t)cstride * 2 * W * H};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)bn};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("tc_conv: cuTensorMapEncodeTiled(activation C=%d W=%d H=%d B=%d box %dx%dx%dx%d stride %d) failed: %d",
                                     C, W, H, B, kc, bw, bh, bn, stride, (int)r); return -3; }
  return 0;
}
static int make_w_map(CUtensorMap* m, const void* base, int Ktot, int Ntot, int kc, int nt, int row_stride, int batch, long long batch_stride) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("tc_conv: cuTensorMapEncodeTiled unavailable"); return -3; }
  cuuint64_t dims[3] = {(cuuint64_t)Ktot, (cuuint64_t)Ntot, (cuuint64_t)(batch > 0 ? batch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride * 2, (cuuint64_t)batch_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)nt, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, batch > 0 ? 3 : 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz(kc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("tc_conv: cuTensorMapEncodeTiled(weights K=%d N=%d box %dx%d row stride %d batch %d) failed: %d", Ktot, Ntot, kc, nt, row_stride, batch, (int)r); return -3; }
  return 0;
}

// choose the M-tile rectangle: bw*bh*bn <= 128 maximising the useful fraction of the 128 accumulator rows
static void choose_tile(int W, int H, int B, int stride, int* bw, int* bh, int* bn) {
  double best = -1; int bbw = 1, bbh = 1, bbn = 1;
  const int lim = 256 / stride;                       // TMA box dimension limit
  for (int w = 1; w <= W && w <= 128 && w <= lim; ++w) {
    int hmax = 128 / w; if (hmax > H) hmax = H; if (hmax > lim) hmax = lim;
    for (int h = 1; h <= hmax; ++h) {
      int n = 128 / (w * h); if (n > B) n = B; if (n < 1) n = 1;
      if (h < H) n = 1;                              // only stack images when a tile covers a whole image
      const long tiles = (long)((W + w - 1) / w) * ((H + h - 1) / h) * ((B + n - 1) / n);
      const double eff = (double)W * H * B / (tiles * 128.0);
      if (eff > best + 1e-9 || (eff > best - 1e-9 && w > bbw)) { best = eff; bbw = w; bbh = h; bbn = n; }
    }
  }
  *bw = bbw; *bh = bbh; *bn = bbn;
}

int check_reg_pool(const void* kernel, const char* name, int low_threads, int low, int high_threads, int high) {
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) { set_error("%s: cudaFuncGetAttributes failed: %s", name, cudaGetErrorString(cudaGetLastError())); return -3; }
  const int r = fa.numRegs;
  if (r < low || low_threads * (r - low) < high_threads * (high - r)) {
    set_error("%s: compiled with %d registers per thread; %d threads releasing down to %d cannot cover %d threads growing to %d (setmaxnreg would spin forever)",
              name, r, low_threads, low, high_threads, high);
    return -3;
  }
  return 0;
}

static const bool g_pdl = []() { const char* e = getenv("UCDIR_PDL"); return !(e && e[0] == '0'); }();

template <int KA, int KB, int NT, int NSPLIT, int EPI, int BSTAT, int SPS, bool SPLIT = false>
static int launch_inst(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const TcParams& p, dim3 grid, cudaStream_t st) {
  using S = TcCfg<KA, KB, NT, NSPLIT, BSTAT, SPS>;
  const int smem_bytes = S::TOTAL + (BSTAT ? p.bres_bytes : 0) + (p.ctab ? 9 * p.Ntot * 4 + 16 : 0);
  static int attr_dev[UCDIR_MAX_DEV] = {};
  int& attr = attr_dev[cur_dev()];
  if (attr < smem_bytes) {
    if (cudaFuncSetAttribute(tc_conv_kernel<KA, KB, NT, NSPLIT, EPI, BSTAT, SPS, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) {
      set_error("tc_conv: cannot opt in to %d bytes of shared memory: %s", smem_bytes, cudaGetErrorString(cudaGetLastError())); return -3; }
    attr = smem_bytes;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, tc_conv_kernel<KA, KB, NT, NSPLIT, EPI, BSTAT, SPS, SPLIT>, a0, a1, b, p) != cudaSuccess) {
    set_error("tc_conv: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return -3; }
  return 0;
}

int launch_tc_conv(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  TcParams p;
  const void* src0 = op.p[UCDIR_TC_P_SRC0]; const void* src1 = op.p[UCDIR_TC_P_SRC1]; const void* w = op.p[UCDIR_TC_P_W];
  p.stats0 = (const double*)op.p[UCDIR_TC_P_STATS0]; p.stats1 = (const double*)op.p[UCDIR_TC_P_STATS1];
  p.tb = (const float*)op.p[UCDIR_TC_P_TB]; p.tg = (const float*)op.p[UCDIR_TC_P_TG];
  p.res = (const __nv_bfloat16*)op.p[UCDIR_TC_P_RES]; p.att = (const float*)op.p[UCDIR_TC_P_ATT];
  p.attw = (const float*)op.p[UCDIR_TC_P_ATTW]; p.dst = op.p[UCDIR_TC_P_DST]; p.dst_stats = (double*)op.p[UCDIR_TC_P_DST_STATS];
  p.B = op.i[UCDIR_TC_I_B]; p.H = op.i[UCDIR_TC_I_H]; p.W = op.i[UCDIR_TC_I_W];
  p.srcH = op.i[UCDIR_TC_I_SRC_H]; p.srcW = op.i[UCDIR_TC_I_SRC_W];
  const int C0 = op.i[UCDIR_TC_I_C0], C1 = op.i[UCDIR_TC_I_C1];
  p.Ntot = op.i[UCDIR_TC_I_NTOT]; p.ncol_valid = op.i[UCDIR_TC_I_NCOL_VALID] ? op.i[UCDIR_TC_I_NCOL_VALID] : p.Ntot;
  p.nty = op.i[UCDIR_TC_I_NTY]; p.ntx = op.i[UCDIR_TC_I_NTX]; p.oy0 = op.i[UCDIR_TC_I_OY0]; p.ox0 = op.i[UCDIR_TC_I_OX0];
  p.stride = op.i[UCDIR_TC_I_STRIDE]; p.groups = op.i[UCDIR_TC_I_GROUPS];
  const int KC = op.i[UCDIR_TC_I_KC];
  int NT = op.i[UCDIR_TC_I_NT];
  const int KB = op.i[UCDIR_TC_I_KB] ? op.i[UCDIR_TC_I_KB] : KC, NSPLIT = op.i[UCDIR_TC_I_NSPLIT] ? op.i[UCDIR_TC_I_NSPLIT] : 1;
  p.gn = op.i[UCDIR_TC_I_GN]; p.ncls = op.i[UCDIR_TC_I_NCLS]; p.act = op.i[UCDIR_TC_I_ACT]; p.mode = op.i[UCDIR_TC_I_MODE];
  p.dst_f32 = op.i[UCDIR_TC_I_DST_F32];
  p.dstC = op.i[UCDIR_TC_I_DST_C]; p.dstCoff = op.i[UCDIR_TC_I_DST_COFF]; p.dstUp = op.i[UCDIR_TC_I_DST_UP];
  p.dstPy = op.i[UCDIR_TC_I_DST_PY]; p.dstPx = op.i[UCDIR_TC_I_DST_PX]; p.resC = op.i[UCDIR_TC_I_RES_C];
  p.attwStride = op.i[UCDIR_TC_I_ATTW_STRIDE];
  p.eps = op.f[UCDIR_TC_F_EPS];
  p.alpha = op.f[UCDIR_TC_F_ALPHA] != 0.f ? op.f[UCDIR_TC_F_ALPHA] : 1.f;
  p.w_batched = op.i[UCDIR_TC_I_W_BATCHED];
  p.dst2 = (__nv_bfloat16*)op.p[UCDIR_TC_P_DST2]; p.t_col0 = op.i[UCDIR_TC_I_T_COL0]; p.t_ld = op.i[UCDIR_TC_I_T_LD];
  const int split = op.i[UCDIR_TC_I_SPLIT] ? 1 : 0;
  const int cstride0 = op.i[UCDIR_TC_I_SRC_CSTRIDE] ? op.i[UCDIR_TC_I_SRC_CSTRIDE] : (split ? 2 : 1) * op.i[UCDIR_TC_I_C0];
  const int w_rowstride = op.i[UCDIR_TC_I_W_ROWSTRIDE];
  const long long w_batchstride = (long long)op.i[UCDIR_TC_I_W_BATCHSTRIDE_LO] + ((long long)op.i[UCDIR_TC_I_W_BATCHSTRIDE_HI] << 31);
  if (!src0 || !w || !p.dst || !p.tb) { set_error("tc_conv: null src0/w/dst/tb"); return -1; }
  if (p.B <= 0 || p.H <= 0 || p.W <= 0 || C0 <= 0 || p.Ntot <= 0 || p.groups < 1) { set_error("tc_conv: bad dims"); return -1; }
  if (p.stride != 1 && p.stride != 2) { set_error("tc_conv: stride must be 1 or 2"); return -2; }
  if (p.nty < 1 || p.ntx < 1 || p.nty > 3 || p.ntx > 3) { set_error("tc_conv: tap grid must be 1..3 per axis"); return -2; }
  if (KC != 64 && KC != 32 && KC != 16) { set_error("tc_conv: KC must be 64, 32 or 16"); return -2; }
  const int Cin = C0 + C1;
  if (p.groups > 1) {
    if (C1 || Cin % p.groups || p.Ntot % p.groups) { set_error("tc_conv: bad grouped config"); return -2; }
    p.Cg = Cin / p.groups; p.Ng = p.Ntot / p.groups;
    // an item of NT columns covers max(1, NT/Ng) groups; its A slab holds their KC consecutive channels
    const int gpi = NT > p.Ng ? NT / p.Ng : 1;
    if ((NT > p.Ng ? NT % p.Ng : p.Ng % NT) || gpi != NSPLIT) { set_error("tc_conv: grouped conv needs NT/Ng = NSPLIT (NT=%d Ng=%d NSPLIT=%d)", NT, p.Ng, NSPLIT); return -2; }
    p.cg_eff = p.Cg * gpi > KC ? p.Cg * gpi : KC;
    if (p.cg_eff % KC || Cin % p.cg_eff) { set_error("tc_conv: grouped conv needs KC | max(Cg*NSPLIT, KC) | Cin"); return -2; }
    if (KB != (p.Cg > 16 ? p.Cg : 16) && !(NSPLIT == 1 && KB == KC)) { set_error("tc_conv: grouped conv needs KB = max(Cg, 16)"); return -2; }
    p.nchunk = p.cg_eff / KC; p.c0_chunks = p.nchunk;
    if (NSPLIT > 1 && p.nchunk != 1) { set_error("tc_conv: split items need one chunk per tap"); return -2; }
    p.n0 = p.nchunk; p.n1 = 0;
  } else {
    if ((C0 % KC && !(C1 == 0 && p.w_batched)) || C1 % KC) { set_error("tc_conv: C0=%d / C1=%d must be multiples of KC=%d", C0, C1, KC); return -2; }
    if (C1 > 0 && !src1) { set_error("tc_conv: null src1"); return -1; }
    if (NSPLIT != 1 || KB != KC) { set_error("tc_conv: dense conv needs NSPLIT = 1 and KB = KC"); return -2; }
    p.Cg = Cin; p.Ng = p.Ntot; p.cg_eff = Cin; p.nchunk = (Cin + KC - 1) / KC; p.c0_chunks = (C0 + KC - 1) / KC;   // a K tail is zero filled by TMA
    p.n0 = p.c0_chunks; p.n1 = C1 / KC;
  }
  // fp32-tolerance mode: (hi, lo) plane pairs, three K passes per tap
  p.a0_lo = p.a1_lo = p.b_lo = p.dst_lo = p.res_lo = p.t_lo = 0;
  if (split) {
    if (op.i[UCDIR_TC_I_SRC_GN_SWISH] || op.i[UCDIR_TC_I_BSTAT] || op.i[UCDIR_TC_I_SPS3]) {
      set_error("tc_conv: SPLIT has no SRC_GN_SWISH / BSTAT / SPS3 form"); return -2; }
    p.a0_lo = op.i[UCDIR_TC_I_SRC_LO_OFF] ? op.i[UCDIR_TC_I_SRC_LO_OFF] : (p.groups > 1 ? Cin : C0);
    p.a1_lo = C1;
    p.b_lo = op.i[UCDIR_TC_I_W_LO_OFF];
    if (p.w_batched && p.b_lo <= 0) { set_error("tc_conv: SPLIT with batched weights needs W_LO_OFF"); return -1; }
    if (p.a0_lo % 8 || (p.w_batched && p.b_lo % 8)) { set_error("tc_conv: SPLIT plane offsets must be multiples of 8 elements"); return -2; }
    p.nchunk *= 3;
    if (!p.dst_f32) { p.dst_lo = p.dstC; p.dstC *= 2; }
    if (p.res) { p.res_lo = p.resC; p.resC *= 2; }
    if (p.dst2) { p.t_lo = p.t_ld; p.t_ld *= 2; }
  }
  // 256-bit stores / residual loads: every 16-column step of a destination / residual row (and of its lo plane) starts on a 32-byte boundary
  p.io32 = (!p.dst_f32 && p.dstC % 16 == 0 && p.dstCoff % 16 == 0 && p.dst_lo % 16 == 0 && (reinterpret_cast<uintptr_t>(p.dst) & 31) == 0 &&
            (!p.res || (p.resC % 16 == 0 && p.res_lo % 16 == 0 && (reinterpret_cast<uintptr_t>(p.res) & 31) == 0))) ? 1 : 0;
  if (p.Ntot % NT) { set_error("tc_conv: Ntot=%d not a multiple of NT=%d", p.Ntot, NT); return -2; }
  if (p.gn) {
    if (!p.tg || !p.stats0 || (C1 > 0 && !p.stats1)) { set_error("tc_conv: GroupNorm fold needs tg and stats"); return -1; }
    if (p.ncls != 1 && p.ncls != 9) { set_error("tc_conv: ncls must be 1 or 9"); return -2; }
    if (p.ncls == 9 && (p.H < 2 || p.W < 2 || p.stride != 1 || p.nty != 3 || p.ntx != 3)) { set_error("tc_conv: 9-class fold needs a 3x3 stride-1 conv on >=2x2 pixels"); return -2; }
  } else { p.ncls = 1; }
  if (p.mode == 1 && (!p.att || !p.attw || !p.res || NT % 32 || p.dst_f32 || p.dstUp)) { set_error("tc_conv: mix epilogue needs att/attw/res, NT %% 32 == 0, bf16 dst"); return -2; }
  if (p.mode != 1 && !p.dst_f32 && (NT % 32 || (p.dstC % 8) || (p.dstCoff % 8))) { set_error("tc_conv: bf16 dst needs NT %% 32 == 0 and 16-byte aligned rows"); return -2; }
  if (p.res && p.mode != 1 && (p.resC % 8)) { set_error("tc_conv: residual rows must be 16-byte aligned"); return -2; }
  if (p.stats1 == nullptr && C1 > 0 && p.gn) { set_error("tc_conv: stats1 missing"); return -1; }
  if (C1 == 0) p.stats1 = nullptr;
  p.gn_count = (double)Cin * p.srcH * p.srcW;
  if (p.w_batched && (p.groups != 1 || C1 || w_rowstride <= 0)) { set_error("tc_conv: batched weights need a dense single-source op and a row stride"); return -2; }
  if (p.dst2 && (p.t_ld <= 0 || p.t_col0 % NT || p.mode == 1 || p.dstUp)) { set_error("tc_conv: bad transposed-store config"); return -2; }
  choose_tile(p.W, p.H, p.w_batched ? 1 : p.B, p.stride, &p.bw, &p.bh, &p.bn);
  p.tiles_x = (p.W + p.bw - 1) / p.bw; p.tiles_y = (p.H + p.bh - 1) / p.bh;
  const int tiles_n = (p.B + p.bn - 1) / p.bn;
  const long mt = (long)p.tiles_x * p.tiles_y * tiles_n;
  if (mt > 0x7fffffffL) { set_error("tc_conv: too many M tiles"); return -2; }
  p.phases = op.i[UCDIR_TC_I_PHASES] > 1 ? op.i[UCDIR_TC_I_PHASES] : 1;
  p.tiles_n_real = tiles_n;
  if (p.phases > 1) {
    if (p.phases != 4 || !p.dstUp || p.nty != 2 || p.ntx != 2 || p.stride != 1 || p.groups != 1 || p.w_batched || p.dst2 || p.mode == 1 ||
        op.i[UCDIR_TC_I_DST_PY] || op.i[UCDIR_TC_I_DST_PX] || mt * 4 > 0x7fffffffL) {
      set_error("tc_conv: PHASES = 4 needs a dense 2x2-tap stride-1 conv with DST_UP = 1 and DST_PY = DST_PX = 0"); return -2; }
  }
  p.m_tiles = (int)(mt * p.phases); p.tiles_n = tiles_n * p.phases;
  if (op.i[UCDIR_TC_I_SRC_GN_SWISH] && !tc_final_halo_applies(op)) {
    set_error("tc_conv: SRC_GN_SWISH needs a 3x3 stride-1 conv of <= 128 channels (multiple of 64) with GN = 0, NT = NTOT = 16, DST_F32 = 1, SRC_GAMMA / SRC_BETA / STATS0");
    return -2;
  }
  if (op.i[UCDIR_TC_I_DST_CROP] && !tc_final_halo_applies(op)) { set_error("tc_conv: DST_CROP needs the fused final conv (SRC_GN_SWISH, ucdir_fhalo.cu)"); return -2; }
  if (op.i[UCDIR_TC_I_RES_FUSED] && !tc_dense_halo_applies(op)) {
    set_error("tc_conv: RES_FUSED needs the halo schedule of a GroupNorm-folded 3x3 stride-1 conv with 64 / 128 output channels, KC = 64, W2 / TB2 / DST_RES");
    return -2;
  }
  if (dry) return 0;
  if (tc_final_halo_applies(op)) return launch_tc_final_halo(op, st);  // GroupNorm + Swish + conv of final_conv in one kernel (ucdir_fhalo.cu)
  if (tc_mix_halo_applies(op)) return launch_tc_mix_halo(op, st);      // halo / weight-stationary form of the mix convs (ucdir_mix.cu)
  if (tc_dense_halo_applies(op)) return launch_tc_dense_halo(op, st);  // halo / super-tile form of the Cout = 64 / 128 3x3 convs (ucdir_dhalo.cu)
  // Small grids (the <= 16x16-pixel levels of a 16-tile rank share: 8..32 pixel tiles x 2 column tiles for 148 SMs): a narrower
  // column tile multiplies the work items.  An M = 128, K = 16 MMA costs ~90 / 125 / 186 cycles at N = 64 / 128 / 256 (operand
  // fetch bound), so narrow tiles only pay when the grid is not full; pick the width with the lowest
  // ceil(items / CTAs) x cycles-per-item estimate among the instantiated ones.
  const int n_sm = sm_count();
  if (p.groups == 1 && p.mode != 1 && NSPLIT == 1 && KC == 64 && !op.i[UCDIR_TC_I_SPS3] && getenv("UCDIR_TC_NT_AUTO") == nullptr) {
    const bool plain_t = p.dst2 != nullptr;
    int best_nt = NT; double best = 1e30;
    const int cand[3] = {256, 128, 64};
    const double cyc[3] = {186.0, 125.0, 90.0};
    for (int c = 0; c < 3; ++c) {
      const int nt = cand[c];
      if (nt > NT || p.Ntot % nt || (plain_t && (nt < 128 || p.t_col0 % nt))) continue;
      const long long it = (long long)p.m_tiles * (p.Ntot / nt);
      const long long per_cta = (it + n_sm - 1) / n_sm;
      const double est = (double)per_cta * cyc[c] + 400.0 * (double)per_cta;      // + per-item epilogue / pipeline refill
      if (est < best * 0.97) { best = est; best_nt = nt; }
    }
    NT = best_nt;
  }
  // ROW3: row tiles (128 px x 1 row) of a dense 3x3 stride-1 conv load one 130-pixel activation row per filter row and
  // issue the three horizontal taps from shifted descriptors of that slab (3x less activation traffic from L2)
  const bool row3 = op.i[UCDIR_TC_I_ROW3] == 1 && p.nty == 3 && p.ntx == 3 && p.stride == 1 && p.groups == 1 && KC == 64 && KB == 64 &&
                    NSPLIT == 1 && p.bw == 128 && p.bh == 1 && p.bn == 1 && !p.w_batched && !p.dst2 && !p.dst_f32 && p.mode != 1 && NT <= 128;
  const int abw = row3 ? p.bw + 2 : p.bw;
  CUtensorMap a0, a1, bm;
  // split: the channel extent of a map covers both planes (the lo plane starts a*_lo channels into the pixel row)
  const int ext0 = split ? p.a0_lo + (p.groups > 1 ? Cin : p.n0 * KC) : C0;
  if (split && ext0 > cstride0) { set_error("tc_conv: SPLIT planes (%d + %d channels) exceed the source row pitch %d", p.a0_lo, ext0 - p.a0_lo, cstride0); return -2; }
  int rc = make_act_map(&a0, src0, ext0, p.srcW, p.srcH, p.B, KC, abw, p.bh, p.bn, p.stride, cstride0);
  if (rc) return rc;
  if (C1 > 0) { rc = make_act_map(&a1, src1, split ? 2 * C1 : C1, p.srcW, p.srcH, p.B, KC, abw, p.bh, p.bn, p.stride, split ? 2 * C1 : C1); if (rc) return rc; }
  else a1 = a0;
  const int Ktot = p.nty * p.ntx * p.nchunk * KB;
  if (p.w_batched) rc = make_w_map(&bm, w, split ? p.b_lo + p.n0 * KB : C0, op.i[UCDIR_TC_I_W_ROWS] ? op.i[UCDIR_TC_I_W_ROWS] : p.Ntot, KB, NT, w_rowstride, p.B, w_batchstride);
  else rc = make_w_map(&bm, w, Ktot, p.Ntot * p.phases, KB, NT, Ktot, 0, 0);
  if (rc) return rc;
  const long long items = (long long)p.m_tiles * (p.Ntot / NT);
  dim3 grid((unsigned)(items < n_sm ? items : n_sm), 1, 1);      // persistent: one CTA per SM
  p.ctab = (!split && p.mode == 1 && p.gn && p.ncls == 9 && p.bn == 1 && p.Ntot <= 1024 && KC == 32 && KB == 16 && op.i[UCDIR_TC_I_NO_CTAB] == 0 &&
            op.i[UCDIR_TC_I_BSTAT] == 0 && op.i[UCDIR_TC_I_SPS3] == 0) ? 1 : 0;
  const int epi = p.mode == 1 ? (p.ctab ? EPI_MIXC : EPI_MIX) : (p.dst_f32 ? EPI_F32 : (p.dst2 ? EPI_PLAIN_T : EPI_PLAIN));
  // weight-stationary schedule: the whole weight block of one N sub-tile stays in shared memory while the CTA walks
  // its M tiles (items ordered N-sub-tile major).  Used for the grouped integration-module convs whose weight slabs
  // are small, many and 32/64-byte rowed (the streamed form re-fetched them for every pixel tile).
  const int nslab = p.nty * p.ntx * p.nchunk;
  p.bres_bytes = nslab * NT * KB * 2;
  // (measured on B200, round 1: no gain over the streamed form -- the mix epilogue, not L2 traffic, bounds these ops --
  //  so it is opt-in)
  const int bstat = ((epi == EPI_MIX || epi == EPI_MIXC) && !p.w_batched && p.bres_bytes <= 150 * 1024 && op.i[UCDIR_TC_I_BSTAT] == 1) ? 1 : 0;
  // three filter taps per pipeline stage for the layers whose per-slab MMA work is small (N <= 128 columns per MMA)
  // (measured on B200, round 1: no gain -- those ops are bound by their epilogue, which already overlaps the main loop --
  //  so it is opt-in)
  int sps = (nslab % 3 == 0 && op.i[UCDIR_TC_I_SPS3] == 1 && (NT / NSPLIT <= 128 || ((epi == EPI_MIX || epi == EPI_MIXC) && KB < 64))) ? 3 : 1;
  if (row3) sps = 4;
#define INSTS(ka, kb, nt, ns, ep, sp) if (split && KC == ka && KB == kb && NT == nt && NSPLIT == ns && epi == ep && sps == sp) { rc = launch_inst<ka, kb, nt, ns, ep, 0, sp, true>(a0, a1, bm, p, grid, st); if (rc) return rc; ++g_launches; return 0; }
  INSTS(64, 64, 64, 1, EPI_PLAIN, 1) INSTS(64, 64, 128, 1, EPI_PLAIN, 1) INSTS(64, 64, 256, 1, EPI_PLAIN, 1) INSTS(16, 16, 64, 1, EPI_PLAIN, 1)
  INSTS(64, 64, 64, 1, EPI_PLAIN, 4) INSTS(64, 64, 128, 1, EPI_PLAIN, 4)
  INSTS(64, 64, 256, 1, EPI_PLAIN_T, 1) INSTS(64, 64, 128, 1, EPI_PLAIN_T, 1)
  INSTS(64, 64, 16, 1, EPI_F32, 1) INSTS(64, 64, 64, 1, EPI_F32, 1) INSTS(64, 64, 128, 1, EPI_F32, 1) INSTS(64, 64, 256, 1, EPI_F32, 1)
  INSTS(32, 16, 256, 4, EPI_MIX, 1) INSTS(32, 16, 256, 2, EPI_MIX, 1) INSTS(32, 32, 256, 1, EPI_MIX, 1) INSTS(64, 64, 256, 1, EPI_MIX, 1)
#undef INSTS
  if (split) { set_error("tc_conv: no SPLIT kernel instance for KC=%d KB=%d NT=%d NSPLIT=%d epilogue %d sps %d", KC, KB, NT, NSPLIT, epi, sps); return -2; }
#define INST(ka, kb, nt, ns, ep, bs, sp) if (KC == ka && KB == kb && NT == nt && NSPLIT == ns && epi == ep && bstat == bs && sps == sp) { rc = launch_inst<ka, kb, nt, ns, ep, bs, sp>(a0, a1, bm, p, grid, st); if (rc) return rc; ++g_launches; return 0; }
  INST(64, 64, 64, 1, EPI_PLAIN, 0, 1) INST(64, 64, 128, 1, EPI_PLAIN, 0, 1) INST(64, 64, 256, 1, EPI_PLAIN, 0, 1) INST(16, 16, 64, 1, EPI_PLAIN, 0, 1)
  INST(64, 64, 64, 1, EPI_PLAIN, 0, 3) INST(64, 64, 128, 1, EPI_PLAIN, 0, 3) INST(16, 16, 64, 1, EPI_PLAIN, 0, 3)
  INST(64, 64, 64, 1, EPI_PLAIN, 0, 4) INST(64, 64, 128, 1, EPI_PLAIN, 0, 4)
  INST(64, 64, 256, 1, EPI_PLAIN_T, 0, 1) INST(64, 64, 128, 1, EPI_PLAIN_T, 0, 1) INST(64, 64, 128, 1, EPI_PLAIN_T, 0, 3)
  INST(64, 64, 16, 1, EPI_F32, 0, 1) INST(64, 64, 64, 1, EPI_F32, 0, 1) INST(64, 64, 128, 1, EPI_F32, 0, 1) INST(64, 64, 256, 1, EPI_F32, 0, 1)
  INST(64, 64, 16, 1, EPI_F32, 0, 3) INST(64, 64, 64, 1, EPI_F32, 0, 3) INST(64, 64, 128, 1, EPI_F32, 0, 3)
  INST(32, 16, 256, 4, EPI_MIX, 1, 3) INST(32, 16, 256, 2, EPI_MIX, 1, 3) INST(32, 32, 256, 1, EPI_MIX, 1, 3)
  INST(32, 16, 256, 4, EPI_MIX, 0, 3) INST(32, 16, 256, 2, EPI_MIX, 0, 3) INST(32, 32, 256, 1, EPI_MIX, 0, 3) INST(64, 64, 256, 1, EPI_MIX, 0, 1)
  INST(32, 16, 256, 4, EPI_MIX, 0, 1) INST(32, 16, 256, 2, EPI_MIX, 0, 1) INST(32, 32, 256, 1, EPI_MIX, 0, 1)
  INST(32, 16, 256, 4, EPI_MIXC, 0, 1) INST(32, 16, 256, 2, EPI_MIXC, 0, 1)
#undef INST
  set_error("tc_conv: no kernel instance for KC=%d KB=%d NT=%d NSPLIT=%d epilogue %d bstat %d sps %d", KC, KB, NT, NSPLIT, epi, bstat, sps);
  return -2;
}

// ------------------------------------------------------------------------------------------------
// bf16 elementwise: GroupNorm(1,C) apply (+Swish) for the one place the fold does not reach
// (final_conv: GN -> Swish -> conv, model/ucdir.py:266-268), and fp32 <-> bf16 casts.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_apply_bf16_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const double* __restrict__ stats, int C, size_t per_sample,
                                                            double count, float eps, int swish) {
  const int b = blockIdx.y;
  const GnScalars sc = gn_scalars(stats, nullptr, b, count, eps);
  const size_t n8 = per_sample / 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // 8 channels per thread and iteration; when the grid stride is a multiple of the row length every iteration of a thread
  // sees the same 8 channels, so their scale / shift are computed once (16 scalar loads per 16 bytes of data otherwise)
  const bool invariant = (stride * 8) % (size_t)C == 0;
  float sa[8], sb[8];
  auto load_affine = [&](int c) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sa[k] = sc.rstd * __ldg(gamma + c + k);
      sb[k] = __ldg(beta + c + k) - sa[k] * sc.mean;
    }
  };
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (invariant && i0 < n8) load_affine((int)((i0 * 8) % C));
  for (size_t i = i0; i < n8; i += stride) {
    const size_t e = i * 8;
    if (!invariant) load_affine((int)(e % C));
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + (size_t)b * per_sample + e));
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
    __align__(16) __nv_bfloat162 o2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = __bfloat1622float2(h2[k]);
      f.x = f.x * sa[2 * k] + sb[2 * k];
      f.y = f.y * sa[2 * k + 1] + sb[2 * k + 1];
      if (swish) {      // x*sigmoid(x) = h + h*tanh(h), h = x/2: one MUFU op; the result is rounded to bf16 (tc_ptx.cuh)
        f.x = swish_half(0.5f * f.x); f.y = swish_half(0.5f * f.y);
      }
      o2[k] = __floats2bfloat162_rn(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(dst + (size_t)b * per_sample + e) = *reinterpret_cast<const uint4*>(o2);
  }
}

// fp32-tolerance mode: source and destination are (hi, lo) plane pairs [B][HW][2*C]; fp32 math, exact Swish.
__global__ void __launch_bounds__(256) gn_apply_split_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const double* __restrict__ stats, int C, size_t n_pix, double count, float eps, int swish) {
  const int b = blockIdx.y;
  const GnScalars sc = gn_scalars(stats, nullptr, b, count, eps);
  const int c8 = C / 8;
  const size_t total = n_pix * c8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i / c8; const int c = (int)(i - pix * c8) * 8;
    const size_t row = ((size_t)b * n_pix + pix) * 2 * C;
    const uint4 uh = __ldg(reinterpret_cast<const uint4*>(src + row + c));
    const uint4 ul = __ldg(reinterpret_cast<const uint4*>(src + row + C + c));
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&uh);
    const __nv_bfloat162* l2 = reinterpret_cast<const __nv_bfloat162*>(&ul);
    __align__(16) __nv_bfloat162 oh[4];
    __align__(16) __nv_bfloat162 ol[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fh = __bfloat1622float2(h2[k]), fl = __bfloat1622float2(l2[k]);
      float x0 = fh.x + fl.x, x1 = fh.y + fl.y;
      const float a0 = sc.rstd * __ldg(gamma + c + 2 * k), a1 = sc.rstd * __ldg(gamma + c + 2 * k + 1);
      x0 = (x0 - sc.mean) * a0 + __ldg(beta + c + 2 * k);
      x1 = (x1 - sc.mean) * a1 + __ldg(beta + c + 2 * k + 1);
      if (swish) { x0 = swish_f(x0); x1 = swish_f(x1); }
      oh[k] = __floats2bfloat162_rn(x0, x1);
      const float2 r = __bfloat1622float2(oh[k]);
      ol[k] = __floats2bfloat162_rn(x0 - r.x, x1 - r.y);
    }
    *reinterpret_cast<uint4*>(dst + row + c) = *reinterpret_cast<const uint4*>(oh);
    *reinterpret_cast<uint4*>(dst + row + C + c) = *reinterpret_cast<const uint4*>(ol);
  }
}

int launch_gn_apply_bf16(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  const int B = op.i[UCDIR_GNA_I_B], HW = op.i[UCDIR_GNA_I_HW], C = op.i[UCDIR_GNA_I_C];
  for (int k = 0; k <= UCDIR_GNA_P_STATS; ++k) if (!op.p[k]) { set_error("gn_apply: null pointer %d", k); return -1; }
  if (B <= 0 || HW <= 0 || C <= 0 || C % 8) { set_error("gn_apply: bad dims"); return -1; }
  if (dry) return 0;
  const size_t per = (size_t)HW * C;
  unsigned gx = (unsigned)((per / 8 + 255) / 256); if (gx > 2368) gx = 2368;       // 16 CTAs per SM x 148
  if (op.i[UCDIR_GNA_I_SPLIT]) {
    gn_apply_split_kernel<<<dim3(gx, B), 256, 0, st>>>((const __nv_bfloat16*)op.p[UCDIR_GNA_P_SRC], (__nv_bfloat16*)op.p[UCDIR_GNA_P_DST],
        (const float*)op.p[UCDIR_GNA_P_GAMMA], (const float*)op.p[UCDIR_GNA_P_BETA], (const double*)op.p[UCDIR_GNA_P_STATS], C, (size_t)HW,
        (double)per, op.f[0], op.i[UCDIR_GNA_I_SWISH]);
    ++g_launches;
    return 0;
  }
  gn_apply_bf16_kernel<<<dim3(gx, B), 256, 0, st>>>((const __nv_bfloat16*)op.p[UCDIR_GNA_P_SRC], (__nv_bfloat16*)op.p[UCDIR_GNA_P_DST],
      (const float*)op.p[UCDIR_GNA_P_GAMMA], (const float*)op.p[UCDIR_GNA_P_BETA], (const double*)op.p[UCDIR_GNA_P_STATS], C, per,
      (double)per, op.f[0], op.i[UCDIR_GNA_I_SWISH]);
  ++g_launches;
  return 0;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __float2bfloat16(s[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ s, float* __restrict__ d, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __bfloat162float(s[i]);
}
// fp32 rows -> (hi, lo) bf16 plane pairs (fp32_tc attention probabilities): one thread per output column pair
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, size_t rows, int cols, int in_ld, int out_ld) {
  const int half = out_ld / 2;                       // out_ld is even (checked by the launcher)
  const size_t total = rows * half;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / half; const int c = (int)(i - r * half) * 2;
    const float a = c < cols ? s[r * in_ld + c] : 0.f, b = c + 1 < cols ? s[r * in_ld + c + 1] : 0.f;
    const __nv_bfloat162 hi = __floats2bfloat162_rn(a, b);
    const float2 f = __bfloat1622float2(hi);
    __nv_bfloat16* o = d + r * 2 * out_ld + c;
    *reinterpret_cast<__nv_bfloat162*>(o) = hi;
    *reinterpret_cast<__nv_bfloat162*>(o + out_ld) = __floats2bfloat162_rn(a - f.x, b - f.y);
  }
}

int launch_cast(const ucdir_op_t& op, cudaStream_t st, bool dry) {
  const size_t n = (size_t)op.i[0] + ((size_t)op.i[1] << 31);
  if (!op.p[0] || !op.p[1] || n == 0) { set_error("cast: bad args"); return -1; }
  if (op.i[2] == 2) {
    const int cols = op.i[3], in_ld = op.i[4], out_ld = op.i[5];
    if (cols <= 0 || in_ld < cols || out_ld < cols || (out_ld & 1)) { set_error("cast: bad split-rows dims"); return -1; }
    if (dry) return 0;
    const size_t total = n * (size_t)(out_ld / 2);
    unsigned g2 = (unsigned)((total + 255) / 256 > 4736 ? 4736 : (total + 255) / 256);
    split_rows_kernel<<<g2, 256, 0, st>>>((const float*)op.p[0], (__nv_bfloat16*)op.p[1], n, cols, in_ld, out_ld);
    ++g_launches;
    return 0;
  }
  if (dry) return 0;
  unsigned g = (unsigned)((n + 255) / 256); if (g > 4736) g = 4736;
  if (op.i[2] == 0) cast_f32_bf16_kernel<<<g, 256, 0, st>>>((const float*)op.p[0], (__nv_bfloat16*)op.p[1], n);
  else cast_bf16_f32_kernel<<<g, 256, 0, st>>>((const __nv_bfloat16*)op.p[0], (float*)op.p[1], n);
  ++g_launches;
  return 0;
}

}  // namespace ucdir
