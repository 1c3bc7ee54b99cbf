// Integration module (model/ucdir.py:116,135-140) on tcgen05, "halo" schedule: grouped 3x3 conv C -> 8C whose 8 adjacent
// columns are mixed per pixel with the guidance map, for C = 64 / 128 / 256 (8 groups of CG = 8 / 16 / 32 channels).
//
// What bounds the streamed form of this op (tc_conv_kernel, EPI_MIX) is not the tensor pipe but L2 -> SM traffic (every
// filter tap re-fetches the activation tile and its weight slab: 288 KB per 128 pixels) and the latency of the mix
// epilogue.  This kernel removes both:
//
// * M tile = 8 x 16 output pixels.  ONE TMA box load brings the 10 x 18 pixel halo of the tile (64 channels = 128-byte
//   swizzled rows, out-of-image pixels zero filled = the convolution padding).  The nine filter taps are nine views of
//   that box: the operand descriptor of tap (ty, tx) starts (ty*10 + tx) rows into the box and strides 10 rows between
//   8-row groups (SBO = 1280 bytes), so 8-row group g = tile row g.  The 128-byte swizzle is a function of the shared
//   memory address only (measured on B200, see make_desc_shifted), so no base-offset correction is needed.
// * weights are stationary: the CTA keeps the 144 KB weight block of a "column set" (512 columns for C = 64 / 128, 256
//   for C = 256) resident in shared memory and walks the M tiles of its unit range; units are ordered set-major.
//   L2 -> SM traffic per 128 pixels drops from 288 KB to the 23 KB halo box.
// * 16 epilogue warps (four per TMEM lane quadrant); warp j of a quadrant owns the fixed 64-column stripe j of every
//   256-column item = 8 output channels.  TMEM is read 16 columns at a time, the load of step k+1 in flight during the math of
//   step k;  out[c] = Swish(sum_s aw[s] * (rstd*acc[c*8+s] + cadd[cls][c*8+s])) + res[c] evaluated as
//   Swish(sum_s (aw[s] rstd) acc[c*8+s] + D[c]), D[c] = sum_s aw[s] cadd[cls][c*8+s]: the D chain reads the folded-GroupNorm
//   additive table of the current (image, set) from shared memory while the TMEM load is in flight; the guidance row of the next
//   tile is prefetched into registers; Swish as h + h*tanh(h) (one MUFU op).
// * residual in / output out as TMA boxes (round 2).  One pixel per lane means a per-thread global access touches 32 different
//   128-byte lines per warp instruction; timing probes (profiles/r02_mix_probes.md) showed that these accesses -- not the
//   arithmetic, the TMEM reads, the table reads or the MUFU ops, which all hide behind the MMA side -- were what the C = 64 / 128
//   launches spent a third of their time on.  Now the residual tile of an item (32 channels x 8 x 16 pixels) lands in shared
//   memory by TMA two items ahead, every epilogue thread swaps its 16-byte chunk for its outputs, and the tile leaves by a TMA
//   store (warp 2 runs this: load -> [res_full] -> epilogue -> [out_done] -> store -> load of item n + 2); partial tiles are
//   clipped by the TMA unit.
// * measured (profiles/r02_mix_probes.md): the MMA side alone runs the C = 64 launch in 194 us (94 cycles per N = 128 MMA).
//
// Packed weights, TB / TG tables and every other operand are those of the streamed form (engine.py:pack_tc_grouped).
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "tc_ptx.cuh"

// Timing probes (wrong results, never part of the product build; UCDIR_NVCC_EXTRA=-DUCDIR_MIX_PROBE=n): 2 = MMAs with half the
// columns (half the B-operand bytes and tensor work: what cta_group::2 would leave per CTA), 3 = epilogue without its TMEM reads,
// 4 = no epilogue work at all (the MMA side alone).
// Round-2 result: probe 2 changes the C = 64 / 128 launches by -4 % / -5 % only -- these kernels are not MMA bound.
#ifndef UCDIR_MIX_PROBE
#define UCDIR_MIX_PROBE 1
#endif
#define UCDIR_MIX_PROBE_NDIV (UCDIR_MIX_PROBE == 2 ? 2 : 1)
#if UCDIR_MIX_PROBE == 3           // the epilogue without its TMEM reads (math, table reads, stores and barriers stay)
#define MX_TMEM_LD16(addr, dst) do { _Pragma("unroll") for (int z_ = 0; z_ < 16; ++z_) (dst)[z_] = __float_as_uint((float)(lane + z_)); } while (0)
#else
#define MX_TMEM_LD16(addr, dst) tmem_ld16(addr, dst)
#endif

namespace ucdir {

struct MixParams {
  const double* stats0;
  const float* tb; const float* tg;
  const __nv_bfloat16* res; const float* att; const float* attw;
  __nv_bfloat16* dst; double* dst_stats;
  int B, H, W, Ntot;
  int tiles_x, tiles_y, m_tiles, n_units;
  int resC, dstC, attwStride;
  int a_lo, res_lo, dst_lo;      // SPLIT: element offset of the lo plane inside a source / residual / destination pixel row
  double gn_count; float eps;
};

constexpr int MX_TW = 8, MX_TH = 16;                       // M tile: 8 x 16 output pixels = 128 accumulator rows
constexpr int MX_BW = MX_TW + 2, MX_BH = MX_TH + 2;        // halo box
constexpr int MX_A_BYTES = MX_BW * MX_BH * 128;            // 23040: 180 pixel rows of 64 bf16 channels
constexpr int MX_A_STAGE = (MX_A_BYTES + 1023) & ~1023;    // 23552
constexpr int MX_WRES = 147456;                            // resident weight block of one column set
constexpr int MX_EPI_WARPS = 16, MX_FIRST_EPI_WARP = 4;
constexpr int MX_THREADS = 32 * (MX_FIRST_EPI_WARP + MX_EPI_WARPS);
// 640 threads launch with 96 registers each; warpgroup 0 shrinks to 40 (releasing 56 x 128) so the four epilogue
// warpgroups can grow to 104 (8 x 512) -- the pool only holds what the CTA released (checked on the host before launch)
constexpr int MX_REGS_LOW = 40, MX_REGS_HIGH = 104;

// SPLIT (fp32_tc): activations are (hi, lo) bf16 plane pairs and the weights [W_hi | W_hi | W_lo] per tap; an item
// is three passes over the nine taps -- lo box x W_hi, hi box x W_hi, hi box x W_lo -- into one fp32 accumulator.  Both weight
// planes stay resident, so a column set is ONE 256-column item (2 x 72 KB), the hi and lo boxes of a tile travel as two stages of
// a three-deep ring (the lo box is consumed first: while the 18 hi MMAs run, the next tile's lo box is already resident and its
// hi box is in flight), and the epilogue works in fp32 (exact Swish, residual = hi + lo, (hi, lo) stores).
template <int CG, bool SPLIT = false>
struct MixCfg {
  static constexpr int C = 8 * CG;                         // channels; columns per group NG = 8C / 8 = C
  static constexpr int NG = C;
  // MMAs per tap and 256-column item.  C = 64: groups 2j and 2j+1 live in the same 32-byte K slice of the pixel row (their
  // weight rows carry zeros for the other group's 8 channels), so ONE N = 128 MMA serves both -- the A slice is read from
  // shared memory once instead of twice (at N = 64 the operand reads, 6 KB per 32 tensor cycles, exceed the 128 B/clk port).
  // columns of an item (one accumulator slot).  SPLIT, C = 256: both weight planes of a 256-column item exceed shared memory, so an
  // item is half a group (128 columns, N = 128 MMAs; the two upper epilogue stripes idle -- this mode is MMA bound)
  static constexpr int ICOLS = (SPLIT && CG == 32) ? 128 : 256;
  static constexpr int NSPLIT = CG == 8 ? 2 : (ICOLS > NG ? ICOLS / NG : 1);    // 2 / 2 / 1
  static constexpr int NSUB = ICOLS / NSPLIT;              // columns per MMA
  static constexpr int KB = CG < 16 ? 16 : CG;             // K elements per tap and group (C = 64: 8 real + 8 foreign, zero weights)
  static constexpr int KSTEPS = KB / 16;
  static constexpr int IPB = (CG == 32 || SPLIT) ? 1 : 2;  // items that share one halo box (= one 64-channel chunk)
  static constexpr int SETCOLS = ICOLS * IPB;
  static constexpr int BSLAB = ICOLS * KB * 2;             // one tap's weight slab of one item (and plane)
  static constexpr int EPI_ACTIVE = 4 * (ICOLS / 64);      // epilogue warps with a stripe inside the item
  static constexpr int NPL = SPLIT ? 2 : 1;                // resident weight planes: slab index = item * 9 + tap, SPLIT: tap * 2 + plane
  static constexpr int NBOX = SPLIT ? 2 : 1;               // halo boxes per unit (SPLIT: lo plane, then hi plane)
  static constexpr int ASTAGES = SPLIT ? 3 : 2;
  static_assert(IPB * 9 * NPL * BSLAB == MX_WRES, "resident weight block");
  static constexpr int CTAB_BYTES = 9 * SETCOLS * 4;
  static constexpr int OFF_W = ASTAGES * MX_A_STAGE;
  static constexpr int OFF_CTAB = OFF_W + MX_WRES;
  // bf16: two item tiles [128 pixels][32 channels] (64-byte rows, 64B swizzle): the residual of an item lands here by TMA, every
  // epilogue thread replaces its 16-byte chunk by its output, and the tile leaves by a TMA store.  SPLIT: none (no room).  (A third
  // tile -- paid for by a 5-class table -- was measured and buys nothing: the load of item n + 2 is not what the epilogue waits for.)
  static constexpr int OFF_RES = OFF_CTAB + CTAB_BYTES;
  static constexpr int IO_TILE = 128 * 64, IO_TILES = 2;
  static constexpr int OFF_BARS = OFF_RES + (SPLIT ? 0 : IO_TILES * IO_TILE);
  static_assert(SPLIT || OFF_RES % 1024 == 0, "I/O tile alignment (swizzle atom)");
  static constexpr int TOTAL = OFF_BARS + 256 + 1024 /* align slack */;
  static_assert(TOTAL <= 227 * 1024, "shared memory");
  static_assert(OFF_W % 1024 == 0, "weight block alignment");
};

// unit = (column set, M tile), set-major; M tile = (image, tile row, tile column), column fastest
struct UnitCursor {
  int set, img, ty, tx;
  __device__ __forceinline__ void init(int u, int m_tiles, int tiles_x, int tiles_y) {
    set = u / m_tiles;
    const int m = u - set * m_tiles;
    tx = m % tiles_x;
    const int t = m / tiles_x;
    ty = t % tiles_y;
    img = t / tiles_y;
  }
  __device__ __forceinline__ void next(int tiles_x, int tiles_y, int B) {
    if (++tx == tiles_x) { tx = 0; if (++ty == tiles_y) { ty = 0; if (++img == B) { img = 0; ++set; } } }
  }
};

template <int CG, bool SPLIT>
__global__ void __launch_bounds__(MX_THREADS, 1) mix_halo_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                 const __grid_constant__ CUtensorMap mapB,
                                                                 const __grid_constant__ CUtensorMap mapR,
                                                                 const __grid_constant__ CUtensorMap mapD, const MixParams p) {
  using S = MixCfg<CG, SPLIT>;
  constexpr int MX_ASTAGES = S::ASTAGES;
  extern __shared__ uint8_t smem_raw[];
  // keep the pointer arithmetic on the __shared__ array (no integer round trip) so table reads compile to LDS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wres = smem + S::OFF_W;
  float* ctab = reinterpret_cast<float*>(smem + S::OFF_CTAB);        // [9][SETCOLS] of the current (image, set)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + S::OFF_BARS);
  uint64_t* a_empty = a_full + MX_ASTAGES;
  uint64_t* tmem_full = a_empty + MX_ASTAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* bfull = tmem_empty + 2;
  uint64_t* bfree = bfull + 1;
  uint64_t* res_full = bfree + 1;                                     // [IO_TILES] residual tile of an item has landed (TMA)
  uint64_t* out_done = res_full + S::IO_TILES;                                  // [IO_TILES] every epilogue warp has written its outputs into the tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_done + S::IO_TILES);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = (int)((long long)p.n_units * blockIdx.x / gridDim.x), u1 = (int)((long long)p.n_units * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    prefetch_tmap(&mapA); prefetch_tmap(&mapB);
    for (int s = 0; s < MX_ASTAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int j = 0; j < 2; ++j) { mbar_init(&tmem_full[j], 1); mbar_init(&tmem_empty[j], S::EPI_ACTIVE); }
    mbar_init(bfull, 1); mbar_init(bfree, 1);
    for (int j = 0; j < S::IO_TILES; ++j) { mbar_init(&res_full[j], 1); mbar_init(&out_done[j], MX_EPI_WARPS); }
    if (!SPLIT) { prefetch_tmap(&mapR); prefetch_tmap(&mapD); }
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

  if (warp < MX_FIRST_EPI_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MX_REGS_LOW));
    if (warp == 0) {
      // ===================== TMA producer =====================
      UnitCursor cur; cur.init(u0, p.m_tiles, p.tiles_x, p.tiles_y);
      int stage = 0; uint32_t phase = 0;
      int cur_set = -1; uint32_t bfree_phase = 0;
      for (int u = u0; u < u1; ++u) {
        if (cur.set != cur_set) {
          // new column set: (re)load its whole weight block; every unit that follows streams one halo box only
          if (cur_set >= 0) { mbar_wait(bfree, bfree_phase); bfree_phase ^= 1; }
          if (elect_one()) {
            mbar_expect_tx(bfull, (uint32_t)MX_WRES);
#pragma unroll 1
            for (int i = 0; i < S::IPB * 9 * S::NPL; ++i) {
              if (SPLIT) {                                   // packed row per tap: [W_hi | W_hi | W_lo]; resident: [tap][W_hi, W_lo]
                const int tap = i >> 1, plane = i & 1;
                tma_load_2d(&mapB, bfull, wres + i * S::BSLAB, (tap * 3 + plane * 2) * S::KB, cur.set * S::SETCOLS);
              } else {
                const int item = i / 9, tap = i - item * 9;
                tma_load_2d(&mapB, bfull, wres + i * S::BSLAB, tap * S::KB, cur.set * S::SETCOLS + item * S::ICOLS);
              }
            }
          }
          __syncwarp();
          cur_set = cur.set;
        }
#pragma unroll
        for (int b = 0; b < S::NBOX; ++b) {                  // SPLIT: the lo box first (it is consumed first)
          mbar_wait(&a_empty[stage], phase ^ 1);
          if (elect_one()) {
            const int chunk = ((cur.set * S::SETCOLS) / S::NG * CG) >> 6;          // 64-channel chunk that holds the set's groups
            mbar_expect_tx(&a_full[stage], (uint32_t)MX_A_BYTES);
            tma_load_4d(&mapA, &a_full[stage], smem + stage * MX_A_STAGE, chunk * 64 + ((SPLIT && b == 0) ? p.a_lo : 0), cur.tx * MX_TW - 1,
                        cur.ty * MX_TH - 1, cur.img);
          }
          __syncwarp();
          if (++stage == MX_ASTAGES) { stage = 0; phase ^= 1; }
        }
        cur.next(p.tiles_x, p.tiles_y, p.B);
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((S::NSUB / UCDIR_MIX_PROBE_NDIV) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      UnitCursor cur; cur.init(u0, p.m_tiles, p.tiles_x, p.tiles_y);
      int stage = 0; uint32_t phase = 0;
      int slot = 0; uint32_t sph = 0;
      int cur_set = -1; uint32_t bfull_phase = 0;
      for (int u = u0; u < u1; ++u) {
        if (cur.set != cur_set) { mbar_wait(bfull, bfull_phase); bfull_phase ^= 1; cur_set = cur.set; }
        const int set = cur.set;
        cur.next(p.tiles_x, p.tiles_y, p.B);
        const bool last_of_set = (cur.set != set) && (u + 1 < u1);
        if constexpr (!SPLIT) {
          mbar_wait(&a_full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * MX_A_STAGE);
#pragma unroll
          for (int item = 0; item < S::IPB; ++item) {
            mbar_wait(&tmem_empty[slot], sph ^ 1);           // the epilogue has drained this accumulator
            tc_fence_after();
            if (elect_one()) {
              const uint32_t tacc = tmem_base + (uint32_t)(slot * 256);
              const int g0 = (set * S::SETCOLS + item * S::ICOLS) / S::NG;        // first group of the item
              constexpr int GPS = S::NSUB / S::NG;                                // groups per MMA
              constexpr uint32_t a_hi = desc_hi(MX_BW * 128, 2u), b_hi = desc_hi(8 * S::KB * 2, S::KB * 2 == 64 ? 4u : 6u);
              // the 32-byte K slice of the 128-byte pixel row that holds the first group of split sp, as a descriptor offset
              uint32_t a_lo0[S::NSPLIT];
#pragma unroll
              for (int sp = 0; sp < S::NSPLIT; ++sp) a_lo0[sp] = desc_lo(a_base) + (uint32_t)(((((g0 + sp * GPS) * CG * 2) & 127) & ~31) >> 4);
              uint32_t b_lo = desc_lo(smem_u32(wres + item * 9 * S::BSLAB));
#pragma unroll 1                                            // one 32-bit add per descriptor and tap: warpgroup 0 runs on 40 registers
              for (int tap = 0; tap < 9; ++tap) {
                const int ty = tap / 3, tx = tap - ty * 3;
                const uint32_t a_tap = (uint32_t)(((ty * MX_BW + tx) * 128) >> 4);
#pragma unroll
                for (int sp = 0; sp < S::NSPLIT; ++sp) {
#pragma unroll
                  for (int k = 0; k < S::KSTEPS; ++k)
                    umma_bf16_lohi(tacc + (uint32_t)(sp * S::NSUB), a_lo0[sp] + a_tap + (uint32_t)(k * 2), a_hi,
                                   b_lo + (uint32_t)((sp * S::NSUB * S::KB * 2) >> 4) + (uint32_t)(k * 2), b_hi, idesc, (tap | k) != 0);
                }
                b_lo += (uint32_t)(S::BSLAB >> 4);
              }
              umma_commit(&tmem_full[slot]);                 // accumulator of this item complete
              if (item == S::IPB - 1) {
                umma_commit(&a_empty[stage]);                // halo box may be overwritten
                if (last_of_set) umma_commit(bfree);         // ... and so may the weight block
              }
            }
            __syncwarp();
            if (++slot == 2) { slot = 0; sph ^= 1; }
          }
          if (++stage == MX_ASTAGES) { stage = 0; phase ^= 1; }
        } else {
          // three passes into one accumulator: lo box x W_hi (frees the lo stage early), hi box x W_hi, hi box x W_lo
          mbar_wait(&tmem_empty[slot], sph ^ 1);           // the epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t tacc = tmem_base + (uint32_t)(slot * 256);
          const int g0 = (set * S::SETCOLS) / S::NG;       // first group of the item
          constexpr int GPS = S::NSUB / S::NG;
          constexpr uint32_t a_hi = desc_hi(MX_BW * 128, 2u), b_hi = desc_hi(8 * S::KB * 2, S::KB * 2 == 64 ? 4u : 6u);
          const uint32_t w_lo0 = desc_lo(smem_u32(wres));
#pragma unroll 1
          for (int pass = 0; pass < 3; ++pass) {           // pass 0: lo box; passes 1, 2: hi box (the next stage)
            if (pass < 2) { mbar_wait(&a_full[stage], phase); tc_fence_after(); }
            if (elect_one()) {
              const uint32_t a_base = smem_u32(smem + stage * MX_A_STAGE);
              uint32_t a_lo0[S::NSPLIT];
#pragma unroll
              for (int sp = 0; sp < S::NSPLIT; ++sp) a_lo0[sp] = desc_lo(a_base) + (uint32_t)(((((g0 + sp * GPS) * CG * 2) & 127) & ~31) >> 4);
              uint32_t b_lo = w_lo0 + (pass == 2 ? (uint32_t)(S::BSLAB >> 4) : 0u);       // weight plane: W_hi, W_hi, W_lo
#pragma unroll 1
              for (int tap = 0; tap < 9; ++tap) {
                const int ty = tap / 3, tx = tap - ty * 3;
                const uint32_t a_tap = (uint32_t)(((ty * MX_BW + tx) * 128) >> 4);
#pragma unroll
                for (int sp = 0; sp < S::NSPLIT; ++sp) {
#pragma unroll
                  for (int k = 0; k < S::KSTEPS; ++k)
                    umma_bf16_lohi(tacc + (uint32_t)(sp * S::NSUB), a_lo0[sp] + a_tap + (uint32_t)(k * 2), a_hi,
                                   b_lo + (uint32_t)((sp * S::NSUB * S::KB * 2) >> 4) + (uint32_t)(k * 2), b_hi, idesc, (pass | tap | k) != 0);
                }
                b_lo += (uint32_t)((2 * S::BSLAB) >> 4);
              }
              if (pass == 0) umma_commit(&a_empty[stage]);   // lo box may be overwritten
              if (pass == 2) {
                umma_commit(&tmem_full[slot]);               // accumulator complete
                umma_commit(&a_empty[stage]);                // hi box may be overwritten
                if (last_of_set) umma_commit(bfree);         // ... and so may the weight block
              }
            }
            __syncwarp();
            if (pass != 1) { if (++stage == MX_ASTAGES) { stage = 0; phase ^= 1; } }
          }
          if (++slot == 2) { slot = 0; sph ^= 1; }
        }
      }
    }
    else if (warp == 2 && !SPLIT) {
      // ===================== tile I/O (bf16): residual tiles in, output tiles out, both as TMA boxes =====================
      // Item n of this CTA uses tile n % IO_TILES:  load residual(n) -> [res_full] -> the epilogue threads swap their 16-byte chunks for
      // outputs -> [out_done] -> store(n) -> (source read) -> load residual(n + IO_TILES).  One elected lane issues everything, so the
      // bulk async groups it waits on are its own.
      const int n_items = (u1 - u0) * S::IPB;
      if (n_items > 0 && elect_one()) {
        UnitCursor cl; cl.init(u0, p.m_tiles, p.tiles_x, p.tiles_y);      // cursor of the next item to load
        UnitCursor cs = cl;                                               // cursor of the next item to store
        int il = 0, is = 0;                                               // item of the unit (0 .. IPB-1) of either cursor
        int tl = 0;                                                       // tile of the next load
        auto chan0 = [&](const UnitCursor& c, int item) { return (c.set * S::SETCOLS + item * S::ICOLS) >> 3; };
        auto load_next = [&]() {
          mbar_expect_tx(&res_full[tl], (uint32_t)S::IO_TILE);
          tma_load_4d(&mapR, &res_full[tl], smem + S::OFF_RES + tl * S::IO_TILE, chan0(cl, il), cl.tx * MX_TW, cl.ty * MX_TH, cl.img);
          if (++tl == S::IO_TILES) tl = 0;
          if (++il == S::IPB) { il = 0; cl.next(p.tiles_x, p.tiles_y, p.B); }
        };
        for (int n = 0; n < S::IO_TILES && n < n_items; ++n) load_next();
        int ts = 0; uint32_t tph = 0;
        for (int n = 0; n < n_items; ++n) {
          mbar_wait(&out_done[ts], tph);
          tma_store_4d(&mapD, smem + S::OFF_RES + ts * S::IO_TILE, chan0(cs, is), cs.tx * MX_TW, cs.ty * MX_TH, cs.img);
          bulk_commit();
          if (++is == S::IPB) { is = 0; cs.next(p.tiles_x, p.tiles_y, p.B); }
          if (++ts == S::IO_TILES) { ts = 0; tph ^= 1; }
          if (n + S::IO_TILES < n_items) {
            bulk_wait_read0();                               // the store has read the tile: it may be overwritten
            load_next();
          }
        }
        bulk_wait0();                                        // every output tile is in global memory before the CTA retires
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MX_REGS_HIGH));
    const int q = warp & 3;                                // TMEM lane quadrant this warp may read
    const int stripe = (warp - MX_FIRST_EPI_WARP) >> 2;    // 64-column stripe of every item
    const int r = q * 32 + lane;                           // accumulator row = pixel of the tile
    const int yy = r >> 3, xx = r & 7;
    const int et = threadIdx.x - 32 * MX_FIRST_EPI_WARP;
    UnitCursor cur; cur.init(u0, p.m_tiles, p.tiles_x, p.tiles_y);
    float s1 = 0.f, s2 = 0.f;                              // GroupNorm statistics of what this thread stored, current image
    int stat_img = -1;
    int tab_img = -1, tab_set = -1;
    float rstd = 1.f;
    float w8[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) w8[s] = 0.f;
    int slot = 0; uint32_t sph = 0;
    // The guidance-map row of unit u+1 is requested while unit u is processed (requested at the point of use it stalled every
    // unit for a full DRAM round trip -- 30% of all warp samples in the first ncu capture of this kernel).
    bool n_valid = false; uint32_t n_pix = 0; int n_cls = 0;
    float4 n_att0 = make_float4(0.f, 0.f, 0.f, 0.f), n_att1 = n_att0;
    auto fetch_unit = [&]() {
      const int y = cur.ty * MX_TH + yy, x = cur.tx * MX_TW + xx;
      n_valid = y < p.H && x < p.W;
      n_pix = n_valid ? (uint32_t)((cur.img * p.H + y) * p.W + x) : 0u;
      n_cls = (y == 0 ? 0 : (y == p.H - 1 ? 2 : 1)) * 3 + (x == 0 ? 0 : (x == p.W - 1 ? 2 : 1));
      n_att0 = __ldg(reinterpret_cast<const float4*>(p.att + (size_t)n_pix * 8));
      n_att1 = __ldg(reinterpret_cast<const float4*>(p.att + (size_t)n_pix * 8 + 4));
    };
    // bf16: this thread's 16-byte chunk of an item tile [128 pixels][4 chunks], 64B swizzle (chunk ^= (row >> 1) & 3): the residual
    // of its 8 channels is read from it and their outputs are written back to it (tile I/O warp above)
    const uint32_t io_chunk = (uint32_t)(S::OFF_RES + r * 64 + ((stripe ^ ((r >> 1) & 3)) << 4));
    int io_t = 0; uint32_t io_ph = 0;                        // I/O tile of the current item and the phase of its barriers
    fetch_unit();
    for (int left = u1 - u0; left > 0; --left) {
      const int img = cur.img, set = cur.set;
      if (img != stat_img) {
        if (p.dst_stats && stat_img >= 0) {
          const double d1 = warp_sum_d((double)s1), d2 = warp_sum_d((double)s2);
          if (lane == 0) { atomicAdd(p.dst_stats + 2 * stat_img, d1); atomicAdd(p.dst_stats + 2 * stat_img + 1, d2); }
        }
        stat_img = img; s1 = 0.f; s2 = 0.f;
      }
      if (img != tab_img || set != tab_set) {
        // all epilogue warps walk the same units, so they all rebuild the additive table at the same unit
        asm volatile("bar.sync 1, %0;" ::"n"(32 * MX_EPI_WARPS) : "memory");       // everyone is done with the old table
        const GnScalars sc = gn_scalars(p.stats0, nullptr, img, p.gn_count, p.eps);
        const float mri = sc.mean * sc.rstd;
        rstd = sc.rstd;
        for (int i = et; i < 9 * S::SETCOLS / 4; i += 32 * MX_EPI_WARPS) {
          const int cls = i / (S::SETCOLS / 4), j4 = i - cls * (S::SETCOLS / 4);
          const size_t g = (size_t)cls * p.Ntot + (size_t)set * S::SETCOLS + (size_t)j4 * 4;
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.tb + g));
          const float4 t = __ldg(reinterpret_cast<const float4*>(p.tg + g));
          reinterpret_cast<float4*>(ctab)[i] = make_float4(fmaf(-mri, t.x, b.x), fmaf(-mri, t.y, b.y), fmaf(-mri, t.z, b.z), fmaf(-mri, t.w, b.w));
        }
        if (img != tab_img) {
          const float* wp = p.attw + (size_t)img * p.attwStride;
#pragma unroll
          for (int s = 0; s < 8; ++s) w8[s] = (SPLIT ? 1.0f : 0.5f) * __ldg(wp + s);     // 1/2: the bf16 Swish below works on x / 2
        }
        tab_img = img; tab_set = set;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * MX_EPI_WARPS) : "memory");
      }
      // operands of this unit were requested one unit ago
      const bool valid = n_valid;
      const uint32_t pix = n_pix;
      const int cls = n_cls;
      // out[c] = Swish(sum_s aw[s] (rstd acc[c*8+s] + cadd[cls][c*8+s])) + res[c]  (model/ucdir.py:136-140 on the folded GroupNorm)
      //        = Swish(sum_s (aw[s] rstd) acc[c*8+s] + D[c]),  D[c] = sum_s aw[s] cadd[cls][c*8+s]:
      // the D chain does not depend on the accumulator (it runs while the TMEM load is in flight, the table values never wait in
      // registers) and the accumulator chain continues from D[c]: 8 packed FMAs + 1 add per output
      float2 aw2[4], awr2[4];
      aw2[0] = make_float2(n_att0.x * w8[0], n_att0.y * w8[1]); aw2[1] = make_float2(n_att0.z * w8[2], n_att0.w * w8[3]);
      aw2[2] = make_float2(n_att1.x * w8[4], n_att1.y * w8[5]); aw2[3] = make_float2(n_att1.z * w8[6], n_att1.w * w8[7]);
      const float2 rs2 = make_float2(rstd, rstd);
#pragma unroll
      for (int j = 0; j < 4; ++j) awr2[j] = __fmul2_rn(aw2[j], rs2);
      cur.next(p.tiles_x, p.tiles_y, p.B);
      if (left > 1) fetch_unit();                            // ... and those of the next unit are requested now
#pragma unroll
      for (int item = 0; item < S::IPB; ++item) {
        if (stripe * 64 >= S::ICOLS) continue;               // (SPLIT, C = 256: 128-column items -- this warp's stripe does not exist)
        const int lcol = item * S::ICOLS + stripe * 64;      // first column of the stripe within the set
        const int ch0 = (set * S::SETCOLS + lcol) >> 3;      // its first output channel
        uint4 res_h = make_uint4(0u, 0u, 0u, 0u), res_l = res_h;
        if (SPLIT && valid) {                                // fp32_tc: plain loads of the two residual planes (the MMA side needs three
          const __nv_bfloat16* rp = p.res + (size_t)pix * p.resC + ch0;            // passes per item, this side has the slack)
          res_h = __ldg(reinterpret_cast<const uint4*>(rp));
          res_l = __ldg(reinterpret_cast<const uint4*>(rp + p.res_lo));
        }
        mbar_wait(&tmem_full[slot], sph);
        tc_fence_after();
#if UCDIR_MIX_PROBE == 4             // timing probe only: the MMA side alone (the epilogue hands every accumulator straight back)
        { tc_fence_before(); __syncwarp(); if (lane == 0) { mbar_arrive(&tmem_empty[slot]); if (!SPLIT) mbar_arrive(&out_done[io_t]); }
          if (++io_t == S::IO_TILES) { io_t = 0; io_ph ^= 1; } if (++slot == 2) { slot = 0; sph ^= 1; } continue; }
#endif
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 256 + stripe * 64);
        uint32_t rbuf[2][16];
        MX_TMEM_LD16(taddr, rbuf[0]);
        float2 d2[8];
        const float4* ct = reinterpret_cast<const float4*>(ctab + cls * S::SETCOLS + lcol);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 ca = ct[2 * c], cb = ct[2 * c + 1];
          float2 d = __fmul2_rn(make_float2(ca.x, ca.y), aw2[0]);
          d = __ffma2_rn(make_float2(ca.z, ca.w), aw2[1], d);
          d = __ffma2_rn(make_float2(cb.x, cb.y), aw2[2], d);
          d2[c] = __ffma2_rn(make_float2(cb.z, cb.w), aw2[3], d);
        }
        uint8_t* io_ptr = smem + io_chunk + io_t * S::IO_TILE;
        if (!SPLIT) {                                        // residual tile of this item (TMA, requested two items ago)
          mbar_wait(&res_full[io_t], io_ph);
          res_h = *reinterpret_cast<const uint4*>(io_ptr);
        }
        const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&res_h);
        const __nv_bfloat162* rl = reinterpret_cast<const __nv_bfloat162*>(&res_l);
        uint32_t o[4], ol[4];
        float2 st1 = make_float2(0.f, 0.f), st2 = make_float2(0.f, 0.f);
        tmem_ld_wait();
        // 16 columns (two output channels) per step; the TMEM load of step k+1 is in flight during the math of step k.  Rows of a
        // partial tile that lie outside the image compute on zero accumulators and are masked where results leave the thread.
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < 3) MX_TMEM_LD16(taddr + 16 * (k + 1), rbuf[(k + 1) & 1]);
          const uint32_t* rv = rbuf[k & 1];
          float hh[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float2 h2 = d2[2 * k + e];
            h2 = __ffma2_rn(make_float2(__uint_as_float(rv[e * 8 + 0]), __uint_as_float(rv[e * 8 + 1])), awr2[0], h2);
            h2 = __ffma2_rn(make_float2(__uint_as_float(rv[e * 8 + 2]), __uint_as_float(rv[e * 8 + 3])), awr2[1], h2);
            h2 = __ffma2_rn(make_float2(__uint_as_float(rv[e * 8 + 4]), __uint_as_float(rv[e * 8 + 5])), awr2[2], h2);
            h2 = __ffma2_rn(make_float2(__uint_as_float(rv[e * 8 + 6]), __uint_as_float(rv[e * 8 + 7])), awr2[3], h2);
            hh[e] = h2.x + h2.y;
          }
          float2 rf = __bfloat1622float2(rh[k]);
          float2 tv;
          if (SPLIT) {
            // fp32_tc: Swish from ex2.approx + rcp.approx (relative error ~3e-7 against a tolerance of rtol 1e-3 / atol 1e-4 and
            // operands of 16 mantissa bits), residual = hi + lo, the result stored as a (hi, lo) pair
            const float2 rg = __bfloat1622float2(rl[k]);
            tv = make_float2(swish_fast(hh[0]) + (rf.x + rg.x), swish_fast(hh[1]) + (rf.y + rg.y));
            const __nv_bfloat162 oh = __floats2bfloat162_rn(tv.x, tv.y);
            const float2 of = __bfloat1622float2(oh);
            const __nv_bfloat162 olo = __floats2bfloat162_rn(tv.x - of.x, tv.y - of.y);
            o[k] = *reinterpret_cast<const uint32_t*>(&oh);
            ol[k] = *reinterpret_cast<const uint32_t*>(&olo);
          } else {
            // Swish = h + h tanh(h), h = x / 2 (the 1/2 is folded into attw): one MUFU op
            tv = make_float2(swish_half(hh[0]) + rf.x, swish_half(hh[1]) + rf.y);
            const __nv_bfloat162 oh = __floats2bfloat162_rn(tv.x, tv.y);
            o[k] = *reinterpret_cast<const uint32_t*>(&oh);
          }
          // statistics from the fp32 values (their bf16 rounding is zero-mean noise of relative size 2^-9)
          st1 = __fadd2_rn(st1, tv);
          st2 = __ffma2_rn(tv, tv, st2);
          if (k < 3) tmem_ld_wait();
          if (k == 2) {
            // every accumulator value of this stripe is in registers: hand the slot back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[slot]);
          }
        }
        if (++slot == 2) { slot = 0; sph ^= 1; }
        if (SPLIT) {
          if (valid) {
            __nv_bfloat16* d = p.dst + (size_t)pix * p.dstC + ch0;
            *reinterpret_cast<uint4*>(d) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(d + p.dst_lo) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
          }
        } else {
          // outputs replace the residual chunk; once all 16 warps have arrived the I/O warp stores the tile (rows outside the image
          // are clipped by the TMA store)
          *reinterpret_cast<uint4*>(io_ptr) = make_uint4(o[0], o[1], o[2], o[3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&out_done[io_t]);
          if (++io_t == S::IO_TILES) { io_t = 0; io_ph ^= 1; }
        }
        if (valid) { s1 += st1.x + st1.y; s2 += st2.x + st2.y; }
      }
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
static const bool g_mix_pdl = []() { const char* e = getenv("UCDIR_PDL"); return !(e && e[0] == '0'); }();

template <int CG, bool SPLIT = false>
static int launch_mix_inst(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& r, const CUtensorMap& d, const MixParams& p, int grid,
                           cudaStream_t st) {
  using S = MixCfg<CG, SPLIT>;
  static bool attr_dev[UCDIR_MAX_DEV] = {};
  bool& attr = attr_dev[cur_dev()];
  if (!attr) {
    if (int rc = check_reg_pool((const void*)mix_halo_kernel<CG, SPLIT>, "tc_mix_halo", 32 * MX_FIRST_EPI_WARP, MX_REGS_LOW, 32 * MX_EPI_WARPS, MX_REGS_HIGH)) return rc;
    if (cudaFuncSetAttribute(mix_halo_kernel<CG, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess) {
      set_error("tc_mix_halo: cannot opt in to %d bytes of shared memory: %s", S::TOTAL, cudaGetErrorString(cudaGetLastError())); return -3; }
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(MX_THREADS); cfg.dynamicSmemBytes = S::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = g_mix_pdl ? 1 : 0;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, mix_halo_kernel<CG, SPLIT>, a, b, r, d, p) != cudaSuccess) {
    set_error("tc_mix_halo: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return -3; }
  return 0;
}

// true when the op (already validated by launch_tc_conv as a grouped mix conv) fits this kernel
bool tc_mix_halo_applies(const ucdir_op_t& op) {
  const int C = op.i[UCDIR_TC_I_C0], H = op.i[UCDIR_TC_I_H], W = op.i[UCDIR_TC_I_W];
  const int KB = op.i[UCDIR_TC_I_KB] ? op.i[UCDIR_TC_I_KB] : op.i[UCDIR_TC_I_KC];
  const bool split = op.i[UCDIR_TC_I_SPLIT] != 0;
  // SPLIT (fp32_tc): default plane layout [hi: C | lo: C]
  if (split && ((op.i[UCDIR_TC_I_SRC_LO_OFF] != 0 && op.i[UCDIR_TC_I_SRC_LO_OFF] != C) || op.i[UCDIR_TC_I_W_LO_OFF] != 0)) return false;
  const int cs = split ? 2 * C : C;
  return op.i[UCDIR_TC_I_HALO] == 1 && op.i[UCDIR_TC_I_MODE] == 1 && op.i[UCDIR_TC_I_GROUPS] == 8 && (C == 64 || C == 128 || C == 256) &&
         op.i[UCDIR_TC_I_C1] == 0 && op.i[UCDIR_TC_I_NTOT] == 8 * C && op.i[UCDIR_TC_I_GN] == 1 && op.i[UCDIR_TC_I_NCLS] == 9 &&
         op.i[UCDIR_TC_I_NTY] == 3 && op.i[UCDIR_TC_I_NTX] == 3 && op.i[UCDIR_TC_I_OY0] == -1 && op.i[UCDIR_TC_I_OX0] == -1 &&
         op.i[UCDIR_TC_I_STRIDE] == 1 && KB == (C / 8 < 16 ? 16 : C / 8) && H >= 2 && W >= 2 && op.i[UCDIR_TC_I_SRC_H] == H &&
         op.i[UCDIR_TC_I_SRC_W] == W && !op.i[UCDIR_TC_I_DST_F32] && !op.i[UCDIR_TC_I_DST_UP] && !op.i[UCDIR_TC_I_W_BATCHED] &&
         !op.p[UCDIR_TC_P_DST2] && op.i[UCDIR_TC_I_DST_COFF] == 0 && (op.i[UCDIR_TC_I_SRC_CSTRIDE] == 0 || op.i[UCDIR_TC_I_SRC_CSTRIDE] == cs) &&
         op.i[UCDIR_TC_I_DST_C] % 8 == 0 && op.i[UCDIR_TC_I_RES_C] % 8 == 0 && op.p[UCDIR_TC_P_TG] && op.p[UCDIR_TC_P_STATS0];
}

int launch_tc_mix_halo(const ucdir_op_t& op, cudaStream_t st) {
  MixParams p;
  p.stats0 = (const double*)op.p[UCDIR_TC_P_STATS0];
  p.tb = (const float*)op.p[UCDIR_TC_P_TB]; p.tg = (const float*)op.p[UCDIR_TC_P_TG];
  p.res = (const __nv_bfloat16*)op.p[UCDIR_TC_P_RES]; p.att = (const float*)op.p[UCDIR_TC_P_ATT]; p.attw = (const float*)op.p[UCDIR_TC_P_ATTW];
  p.dst = (__nv_bfloat16*)op.p[UCDIR_TC_P_DST]; p.dst_stats = (double*)op.p[UCDIR_TC_P_DST_STATS];
  p.B = op.i[UCDIR_TC_I_B]; p.H = op.i[UCDIR_TC_I_H]; p.W = op.i[UCDIR_TC_I_W]; p.Ntot = op.i[UCDIR_TC_I_NTOT];
  const int C = op.i[UCDIR_TC_I_C0], CG = C / 8, KB = CG < 16 ? 16 : CG;
  p.resC = op.i[UCDIR_TC_I_RES_C]; p.dstC = op.i[UCDIR_TC_I_DST_C]; p.attwStride = op.i[UCDIR_TC_I_ATTW_STRIDE];
  const bool split = op.i[UCDIR_TC_I_SPLIT] != 0;
  p.a_lo = p.res_lo = p.dst_lo = 0;
  if (split) {                                               // (hi, lo) plane pairs: rows are twice as long, lo plane after the hi plane
    p.a_lo = C; p.res_lo = p.resC; p.resC *= 2; p.dst_lo = p.dstC; p.dstC *= 2;
  }
  p.eps = op.f[UCDIR_TC_F_EPS];
  p.gn_count = (double)C * p.H * p.W;
  p.tiles_x = (p.W + MX_TW - 1) / MX_TW; p.tiles_y = (p.H + MX_TH - 1) / MX_TH;
  const long long mt = (long long)p.tiles_x * p.tiles_y * p.B;
  const int setcols = split ? (CG == 32 ? 128 : 256) : (CG == 32 ? 256 : 512);
  const long long units = mt * (p.Ntot / setcols);
  if (units > 0x7fffffffLL || (long long)p.B * p.H * p.W > 0x7fffffffLL) { set_error("tc_mix_halo: too many units / pixels"); return -2; }
  p.m_tiles = (int)mt; p.n_units = (int)units;
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("tc_mix_halo: cuTensorMapEncodeTiled unavailable"); return -3; }
  CUtensorMap ma, mb;
  {
    const cuuint64_t cs = split ? 2 * C : C;                 // channels per pixel row (both planes)
    cuuint64_t dims[4] = {cs, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {cs * 2, cs * 2 * p.W, cs * 2 * p.W * p.H};
    cuuint32_t box[4] = {64, MX_BW, MX_BH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(op.p[UCDIR_TC_P_SRC0]), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tc_mix_halo: cuTensorMapEncodeTiled(activation C=%d W=%d H=%d B=%d) failed: %d", C, p.W, p.H, p.B, (int)r); return -3; }
  }
  {
    const int Ktot = 9 * KB * (split ? 3 : 1);               // SPLIT: [W_hi | W_hi | W_lo] per tap
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)p.Ntot};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)((split && CG == 32) ? 128 : 256)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(op.p[UCDIR_TC_P_W]), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, KB == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tc_mix_halo: cuTensorMapEncodeTiled(weights K=%d N=%d) failed: %d", Ktot, p.Ntot, (int)r); return -3; }
  }
  // bf16: residual tiles in / output tiles out as TMA boxes of 32 channels x 8 x 16 pixels (64-byte rows, 64B swizzle; ucdir_mix.cu: tile I/O)
  CUtensorMap mr = ma, md = ma;
  if (!split) {
    for (int k = 0; k < 2; ++k) {
      const int pitch = k == 0 ? p.resC : p.dstC;
      void* base = k == 0 ? const_cast<__nv_bfloat16*>(p.res) : p.dst;
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
      cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)pitch * 2 * p.W, (cuuint64_t)pitch * 2 * p.W * p.H};
      cuuint32_t box[4] = {32, MX_TW, MX_TH, 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = enc(k == 0 ? &mr : &md, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("tc_mix_halo: cuTensorMapEncodeTiled(%s C=%d pitch=%d) failed: %d", k == 0 ? "residual" : "destination", C, pitch, (int)r); return -3; }
    }
  }
  const int n_sm = sm_count();
  const int grid = units < n_sm ? (int)units : n_sm;       // persistent: one CTA per SM
  int rc;
  if (split) rc = CG == 8 ? launch_mix_inst<8, true>(ma, mb, mr, md, p, grid, st)
                : (CG == 16 ? launch_mix_inst<16, true>(ma, mb, mr, md, p, grid, st) : launch_mix_inst<32, true>(ma, mb, mr, md, p, grid, st));
  else if (CG == 8) rc = launch_mix_inst<8>(ma, mb, mr, md, p, grid, st);
  else if (CG == 16) rc = launch_mix_inst<16>(ma, mb, mr, md, p, grid, st);
  else rc = launch_mix_inst<32>(ma, mb, mr, md, p, grid, st);
  if (rc) return rc;
  ++g_launches;
  return 0;
}

}  // namespace ucdir
