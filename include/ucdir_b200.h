/* ucdir_b200.h -- C ABI of libucdir_b200.so (hand-written sm_100a CUDA for UCDIR's denoising hot path).
 *
 * The reference (zhangyi-3/UCDIR) is pure Python/PyTorch and has no FFI of its own; every arithmetic
 * step of its hot path is a library call issued from model/ucdir.py, model/diffusion.py and
 * utils/util.py.  This ABI is what a binding for that path binds instead: plain
 * pointers and sizes, no torch types, no allocation, no implicit synchronisation, stream ordered,
 * int return codes (0 = ok, <0 = error, text via ucdir_last_error()).
 *
 * The unit of work is an "op": one kernel launch described by a POD record.  A whole UNet forward, a
 * whole denoising step, or a single op for a unit test are all arrays of ops handed to
 * ucdir_run_ops() in one call (host code builds the array once per shape -- see ucdir_b200/engine.py).
 *
 * Which reference code each op kind replaces (paths relative to the reference root):
 *   UCDIR_OP_CONV_F32      nn.Conv2d / GroupNorm(1,C) / Swish / grouped spdyconv + per-pixel mix + residual
 *                          model/ucdir.py:57,66,109-140,162-163,180-182,223,266-268 ; predictor convs :360-403
 *   UCDIR_OP_SGEMM_F32     the two attention einsums                    model/ucdir.py:174,179
 *   UCDIR_OP_SOFTMAX_F32   torch.softmax over keys                      model/ucdir.py:176
 *   UCDIR_OP_GUIDANCE      F.interpolate(bilinear) + conv2 branch       model/ucdir.py:133-135,113-114
 *   UCDIR_OP_TIME_EMBED    PositionalEncoding + noise_level_mlp + per-block noise_func
 *                                                                      model/ucdir.py:24-29,212-214,106,125
 *   UCDIR_OP_GATHER_TILES  F.pad(reflect) + window slicing + torch.cat([cond, x])
 *                          utils/util.py:117-137 ; model/ucdir.py:303-306 ; model/diffusion.py:166
 *   UCDIR_OP_SCATTER       interior write-back + crop (+ fused posterior step)
 *                          utils/util.py:144-146 ; model/diffusion.py:150-158,171-172,182-183
 *   UCDIR_OP_MAXPOOL2      nn.MaxPool2d(2)                              model/ucdir.py:363-375
 *   UCDIR_OP_TC_*          bf16 tcgen05/TMA versions of CONV / attention (same reference lines)
 *   UCDIR_OP_TO_IMAGE_U8   result crop + tensor2img                     model/model.py:137, core/metrics.py:8-34
 *
 * Activation layout inside the library: NHWC ("pixel-major, channel-innermost"), fp32 or bf16.
 * Image layout at the boundary: the reference's NCHW fp32.
 */
#ifndef UCDIR_B200_H
#define UCDIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UCDIR_ABI_VERSION 14

#define UCDIR_OP_NPTR 16
#define UCDIR_OP_NINT 56
#define UCDIR_OP_NFLT 8

typedef struct ucdir_op {
  int32_t kind;                 /* enum ucdir_op_kind */
  int32_t flags;                /* UCDIR_OP_FLAG_* (0 = none) */
  void*   p[UCDIR_OP_NPTR];     /* device pointers, meaning per kind (enums below) */
  int32_t i[UCDIR_OP_NINT];     /* integers, meaning per kind */
  float   f[UCDIR_OP_NFLT];     /* floats, meaning per kind */
} ucdir_op_t;

/* Op flags.  An op list is a sequential program; the flags only tell ucdir_graph_capture which ops need not wait for each other:
 * a BRANCH op is captured on a side stream forked from the main stream just before it (it depends on everything before it, nothing
 * after it depends on it) until the next op flagged JOIN, which waits for the side stream as well.  ucdir_run_ops ignores the flags and
 * runs the list in order.  The caller guarantees that no op between a BRANCH op and its JOIN reads what the BRANCH op writes or writes
 * what it reads (engine.py: a block's 1x1 res_conv beside its conv1, joined at the integration conv that reads both). */
#define UCDIR_OP_FLAG_BRANCH 1
#define UCDIR_OP_FLAG_JOIN 2

enum ucdir_op_kind {
  UCDIR_OP_CONV_F32 = 1,
  UCDIR_OP_SGEMM_F32 = 2,
  UCDIR_OP_SOFTMAX_F32 = 3,
  UCDIR_OP_GUIDANCE = 4,
  UCDIR_OP_TIME_EMBED = 5,
  UCDIR_OP_GATHER_TILES = 6,
  UCDIR_OP_SCATTER = 7,
  UCDIR_OP_MAXPOOL2 = 8,
  UCDIR_OP_MEMSET = 9,
  UCDIR_OP_TC_CONV = 10,
  UCDIR_OP_TC_ATTN = 11,
  UCDIR_OP_GN_APPLY_BF16 = 12,
  UCDIR_OP_CAST = 13,
  UCDIR_OP_CROP_TILES = 14,
  UCDIR_OP_GN_STATS_F32 = 15,
  UCDIR_OP_GN_APPLY_F32 = 16,
  UCDIR_OP_LAYOUT = 17,
  UCDIR_OP_TO_IMAGE_U8 = 18
};

/* ---- UCDIR_OP_CONV_F32: dst = epilogue( conv( prologue(concat(src0, src1)) ) ) -----------------
 * Implicit GEMM over NHWC fp32: M = output pixels of one sample, N = COUT, K = KSIZE^2 * (C0+C1)/GROUPS.
 * prologue PRE: 0 none | 1 GroupNorm(1,C) | 2 GroupNorm(1,C) then Swish.  Statistics come from STATS0/1
 *   (per sample {sum, sum of squares} in double, written by the producing op's epilogue); zero padding is
 *   applied after the prologue, as in the reference where the conv pads the normalised tensor.
 * UP=1: the source is read through a nearest x2 upsample (model/ucdir.py:56).
 * epilogue MODE 0 (plain): v = acc + bias[n]; ACT (0 none | 1 Swish | 2 LeakyReLU 0.2); + RES[pix, n] if RES;
 *                          FiLM: if FILM_B: v = (1 + FILM_G[b,n]) * v + FILM_B[b,n]  (FILM_G may be NULL => v + FILM_B)
 *                          applied before ACT/RES (model/ucdir.py:38-45).
 * epilogue MODE 1 (mix):   grouped conv C -> 8C whose 8 adjacent outputs n = c*8+s are mixed per pixel:
 *                          h = sum_s (acc[c*8+s] + bias) * ATT[pix, s] * ATTW[b, s]; v = Swish(h) + RES[pix, c]
 *                          (model/ucdir.py:135-140); the 8C tensor is never written.
 * Every epilogue also accumulates {sum v, sum v^2} per sample into DST_STATS when non-NULL (the next GN).
 * DST_UP=1 writes to pixel (2y+DST_PY, 2x+DST_PX) of a 2H x 2W tensor (ConvTranspose2d 2x2 s2 as 4 phases).
 */
enum ucdir_conv_ptr {
  UCDIR_CONV_P_SRC0 = 0, UCDIR_CONV_P_SRC1 = 1, UCDIR_CONV_P_W = 2, UCDIR_CONV_P_BIAS = 3,
  UCDIR_CONV_P_GAMMA = 4, UCDIR_CONV_P_BETA = 5, UCDIR_CONV_P_STATS0 = 6, UCDIR_CONV_P_STATS1 = 7,
  UCDIR_CONV_P_RES = 8, UCDIR_CONV_P_ATT = 9, UCDIR_CONV_P_ATTW = 10, UCDIR_CONV_P_DST = 11,
  UCDIR_CONV_P_DST_STATS = 12, UCDIR_CONV_P_FILM_G = 13, UCDIR_CONV_P_FILM_B = 14
};
enum ucdir_conv_int {
  UCDIR_CONV_I_B = 0, UCDIR_CONV_I_H = 1, UCDIR_CONV_I_W = 2,        /* output spatial size per sample */
  UCDIR_CONV_I_C0 = 3, UCDIR_CONV_I_C1 = 4, UCDIR_CONV_I_COUT = 5,
  UCDIR_CONV_I_KSIZE = 6, UCDIR_CONV_I_STRIDE = 7, UCDIR_CONV_I_UP = 8, UCDIR_CONV_I_GROUPS = 9,
  UCDIR_CONV_I_PRE = 10, UCDIR_CONV_I_ACT = 11, UCDIR_CONV_I_MODE = 12,
  UCDIR_CONV_I_SRC_H = 13, UCDIR_CONV_I_SRC_W = 14,                  /* stored source size (before UP) */
  UCDIR_CONV_I_DST_C = 15, UCDIR_CONV_I_DST_COFF = 16,               /* dst channel stride / offset */
  UCDIR_CONV_I_DST_UP = 17, UCDIR_CONV_I_DST_PY = 18, UCDIR_CONV_I_DST_PX = 19,
  UCDIR_CONV_I_RES_C = 20, UCDIR_CONV_I_ATTW_STRIDE = 21,            /* floats between samples in ATTW */
  UCDIR_CONV_I_GN_GROUPS = 22,                                       /* must be 1: G > 1 uses UCDIR_OP_GN_APPLY_F32 */
  UCDIR_CONV_I_FILM_STRIDE = 23                                      /* floats between samples in FILM_G / FILM_B (0 = COUT) */
};
enum ucdir_conv_flt { UCDIR_CONV_F_EPS = 0 };

/* ---- UCDIR_OP_SGEMM_F32: C[b] = alpha * A[b] * op(B[b]); TRANSB=1: B is N x K row-major ------------- */
enum ucdir_sgemm_ptr { UCDIR_SGEMM_P_A = 0, UCDIR_SGEMM_P_B = 1, UCDIR_SGEMM_P_C = 2 };
enum ucdir_sgemm_int {
  UCDIR_SGEMM_I_BATCH = 0, UCDIR_SGEMM_I_M = 1, UCDIR_SGEMM_I_N = 2, UCDIR_SGEMM_I_K = 3,
  UCDIR_SGEMM_I_LDA = 4, UCDIR_SGEMM_I_LDB = 5, UCDIR_SGEMM_I_LDC = 6,
  UCDIR_SGEMM_I_SA = 7, UCDIR_SGEMM_I_SB = 8, UCDIR_SGEMM_I_SC = 9,  /* batch strides in elements */
  UCDIR_SGEMM_I_TRANSB = 10,
  UCDIR_SGEMM_I_A_BF16 = 11, UCDIR_SGEMM_I_B_BF16 = 12, UCDIR_SGEMM_I_C_BF16 = 13   /* operand element types (0 = fp32) */
};
enum ucdir_sgemm_flt { UCDIR_SGEMM_F_ALPHA = 0 };

/* ---- UCDIR_OP_SOFTMAX_F32: in-place softmax over each row of X[ROWS][COLS] --------------------------- */
enum ucdir_softmax_ptr { UCDIR_SOFTMAX_P_X = 0, UCDIR_SOFTMAX_P_OUT_BF16 = 1 /* optional: write bf16 probabilities here, rows OUT_LD apart (tail zeroed), instead of in place */ };
enum ucdir_softmax_int { UCDIR_SOFTMAX_I_ROWS = 0, UCDIR_SOFTMAX_I_COLS = 1, UCDIR_SOFTMAX_I_OUT_LD = 2, UCDIR_SOFTMAX_I_IN_LD = 3 /* 0 = COLS */ };

/* ---- UCDIR_OP_GUIDANCE: DST[B,H,W,8] = conv3x3(SimpleGate(conv1x1(bilinear(GUIDE[B,GH,GW,4])))) --------- */
enum ucdir_guid_ptr {
  UCDIR_GUID_P_GUIDE = 0, UCDIR_GUID_P_W0 = 1, UCDIR_GUID_P_B0 = 2, UCDIR_GUID_P_W2 = 3, UCDIR_GUID_P_B2 = 4,
  UCDIR_GUID_P_DST = 5
};
enum ucdir_guid_int { UCDIR_GUID_I_B = 0, UCDIR_GUID_I_GH = 1, UCDIR_GUID_I_GW = 2, UCDIR_GUID_I_H = 3, UCDIR_GUID_I_W = 4 };

/* ---- UCDIR_OP_TIME_EMBED: DST[L][NBLK][8] from noise levels -------------------------------------------
 * LEVELS (float[L]) or, when NULL, the single value f[LEVEL] replicated L times.
 * BLK = NBLK records of { W_a[8][INNER], b_a[8], W_b[8][8], b_b[8] } (noise_func.0 / noise_func.2).     */
enum ucdir_temb_ptr {
  UCDIR_TEMB_P_LEVELS = 0, UCDIR_TEMB_P_W1 = 1, UCDIR_TEMB_P_B1 = 2, UCDIR_TEMB_P_W2 = 3, UCDIR_TEMB_P_B2 = 4,
  UCDIR_TEMB_P_BLK = 5, UCDIR_TEMB_P_DST = 6, UCDIR_TEMB_P_TEMB_OUT = 7 /* optional float[L][INNER] */
};
enum ucdir_temb_int { UCDIR_TEMB_I_L = 0, UCDIR_TEMB_I_NBLK = 1, UCDIR_TEMB_I_INNER = 2 };
enum ucdir_temb_flt { UCDIR_TEMB_F_LEVEL = 0 };

/* ---- UCDIR_OP_GATHER_TILES: DST[BT,TH,TW,CD] from NCHW fp32 images with on-the-fly reflect padding -------
 * TAB = int32[BT][3] {image index, y0, x0}: tile origin in padded coordinates; source pixel of padded
 * coordinate P is reflect(P - PD) (PD=0 with bottom/right overhang reproduces model/ucdir.py:303-306).
 * Channels: CA from SRC_A, then CB from SRC_B (may be NULL/0), zero-filled up to CD.  OUT_BF16=1 writes bf16;
 * OUT_BF16=2 writes (hi, lo) bf16 plane pairs, 2*CD elements per pixel (UCDIR_TC_I_SPLIT). */
enum ucdir_gather_ptr { UCDIR_GATHER_P_SRC_A = 0, UCDIR_GATHER_P_SRC_B = 1, UCDIR_GATHER_P_TAB = 2, UCDIR_GATHER_P_DST = 3 };
enum ucdir_gather_int {
  UCDIR_GATHER_I_BT = 0, UCDIR_GATHER_I_TH = 1, UCDIR_GATHER_I_TW = 2, UCDIR_GATHER_I_IMG_H = 3,
  UCDIR_GATHER_I_IMG_W = 4, UCDIR_GATHER_I_PD = 5, UCDIR_GATHER_I_CA = 6, UCDIR_GATHER_I_CB = 7,
  UCDIR_GATHER_I_CD = 8, UCDIR_GATHER_I_OUT_BF16 = 9
};

/* ---- UCDIR_OP_SCATTER: stitch tile outputs back to NCHW images; MODE 1 fuses the posterior step ---------
 * OWNER_Y[IMG_H] / OWNER_X[IMG_W]: index of the tile row / column whose interior owns that image row /
 * column (the LAST window in reference order, utils/util.py:124-145), -1 = never written (zeros).
 * Y0[NTY] / X0[NTX]: window origins in padded coordinates.  Tile index = (img*NTY + ty)*NTX + tx.
 * MODE 0: OUT = eps.   MODE 1: x0 = clamp(A*x - B*eps); OUT = C1*x0 + C2*x + C3*eps + SIGMA*noise
 * (model/diffusion.py:150-158,171-172,182-183; NOISE NULL => 0; CLIP=0 skips the clamp).                     */
enum ucdir_scatter_ptr {
  UCDIR_SCATTER_P_EPS = 0, UCDIR_SCATTER_P_OWNER_Y = 1, UCDIR_SCATTER_P_OWNER_X = 2, UCDIR_SCATTER_P_Y0 = 3,
  UCDIR_SCATTER_P_X0 = 4, UCDIR_SCATTER_P_XT = 5, UCDIR_SCATTER_P_NOISE = 6, UCDIR_SCATTER_P_OUT = 7,
  UCDIR_SCATTER_P_PARAMS = 8   /* optional device float[8] {A, B, C1, C2, SIGMA, clip, use_noise, C3} overriding f[] / CLIP / NOISE:
                                  lets a captured CUDA graph be replayed for every step of the schedule */
};
enum ucdir_scatter_int {
  UCDIR_SCATTER_I_BIMG = 0, UCDIR_SCATTER_I_IMG_H = 1, UCDIR_SCATTER_I_IMG_W = 2, UCDIR_SCATTER_I_NTY = 3,
  UCDIR_SCATTER_I_NTX = 4, UCDIR_SCATTER_I_TH = 5, UCDIR_SCATTER_I_TW = 6, UCDIR_SCATTER_I_PD = 7,
  UCDIR_SCATTER_I_CE = 8, UCDIR_SCATTER_I_MODE = 9, UCDIR_SCATTER_I_CLIP = 10, UCDIR_SCATTER_I_C = 11
};
enum ucdir_scatter_flt {
  UCDIR_SCATTER_F_A = 0, UCDIR_SCATTER_F_B = 1, UCDIR_SCATTER_F_C1 = 2, UCDIR_SCATTER_F_C2 = 3, UCDIR_SCATTER_F_SIGMA = 4,
  UCDIR_SCATTER_F_C3 = 5   /* DDIM (model/diffusion.py:287): OUT = C1*x0 + C2*x + C3*eps + SIGMA*noise */
};

/* ---- UCDIR_OP_MAXPOOL2: DST[B,H,W,C] = max 2x2 of SRC[B,2H,2W,C] (fp32 NHWC) -------------------------- */
enum ucdir_pool_ptr { UCDIR_POOL_P_SRC = 0, UCDIR_POOL_P_DST = 1 };
enum ucdir_pool_int { UCDIR_POOL_I_B = 0, UCDIR_POOL_I_H = 1, UCDIR_POOL_I_W = 2, UCDIR_POOL_I_C = 3,
                      UCDIR_POOL_I_SPLIT = 4 /* 1: (hi, lo) bf16 plane pairs [..][2*C] instead of fp32 (UCDIR_TC_I_SPLIT) */ };

/* ---- UCDIR_OP_TC_CONV: bf16 tcgen05 / TMA implicit-GEMM convolution (ucdir_tc.cu) ---------------------------
 * Same reference lines as UCDIR_OP_CONV_F32.  Activations bf16 NHWC; weights bf16 [NTOT][K] K-major with
 * K = (tap, channel chunk) slabs of KC channels, packed by ucdir_b200/engine.py:pack_tc_*.
 * Source pixel of output (y, x), tap (ty, tx): (y*STRIDE + ty + OY0, x*STRIDE + tx + OX0); outside the image = 0.
 * GN=1: GroupNorm(1,C) of the input is folded: weights carry gamma, the epilogue applies
 *   v = rstd*acc - mean*rstd*TG[cls][n] + TB[cls][n]  (cls = 3x3 border class when NCLS = 9, else 0);
 * GN=0: v = acc + TB[0][n] (TB = bias).  MODE / ACT (0 none | 1 Swish | 2 LeakyReLU 0.2) / RES / DST_UP as in UCDIR_OP_CONV_F32 (RES, DST bf16;
 * DST_F32=1 stores fp32 and only columns < NCOL_VALID).  DST_STATS as in UCDIR_OP_CONV_F32. */
enum ucdir_tc_ptr {
  UCDIR_TC_P_SRC0 = 0, UCDIR_TC_P_SRC1 = 1, UCDIR_TC_P_W = 2, UCDIR_TC_P_TB = 3, UCDIR_TC_P_TG = 4,
  UCDIR_TC_P_STATS0 = 5, UCDIR_TC_P_STATS1 = 6, UCDIR_TC_P_RES = 7, UCDIR_TC_P_ATT = 8, UCDIR_TC_P_ATTW = 9,
  UCDIR_TC_P_DST = 10, UCDIR_TC_P_DST_STATS = 11,
  UCDIR_TC_P_DST2 = 12,    /* bf16 [B][NTOT - T_COL0][T_LD]: columns >= T_COL0 are stored transposed here (attention V^T) */
  UCDIR_TC_P_SRC_GAMMA = 13, UCDIR_TC_P_SRC_BETA = 14,  /* fp32 [C0]: affine of the SRC_GN_SWISH source transform */
  /* RES_FUSED (exclusive with SRC_GN_SWISH; W2 / TB2 share its slots): the block's 1x1 res_conv of the same input */
  UCDIR_TC_P_W2 = 13,      /* bf16 [NTOT][C0 + C1], K-major */
  UCDIR_TC_P_TB2 = 14,     /* fp32 [NTOT] bias */
  UCDIR_TC_P_DST_RES = 15  /* bf16 [B][H][W][DST_RES_C] */
};
enum ucdir_tc_int {
  UCDIR_TC_I_B = 0, UCDIR_TC_I_H = 1, UCDIR_TC_I_W = 2, UCDIR_TC_I_SRC_H = 3, UCDIR_TC_I_SRC_W = 4,
  UCDIR_TC_I_C0 = 5, UCDIR_TC_I_C1 = 6, UCDIR_TC_I_NTOT = 7, UCDIR_TC_I_NCOL_VALID = 8,
  UCDIR_TC_I_NTY = 9, UCDIR_TC_I_NTX = 10, UCDIR_TC_I_OY0 = 11, UCDIR_TC_I_OX0 = 12, UCDIR_TC_I_STRIDE = 13,
  UCDIR_TC_I_GROUPS = 14, UCDIR_TC_I_KC = 15, UCDIR_TC_I_NT = 16, UCDIR_TC_I_GN = 17, UCDIR_TC_I_NCLS = 18,
  UCDIR_TC_I_ACT = 19, UCDIR_TC_I_MODE = 20, UCDIR_TC_I_DST_F32 = 21, UCDIR_TC_I_DST_C = 22,
  UCDIR_TC_I_DST_COFF = 23, UCDIR_TC_I_DST_UP = 24, UCDIR_TC_I_DST_PY = 25, UCDIR_TC_I_DST_PX = 26,
  UCDIR_TC_I_RES_C = 27, UCDIR_TC_I_ATTW_STRIDE = 28,
  UCDIR_TC_I_KB = 29,      /* K elements per weight slab row (0 = KC); grouped convs: max(Cin/groups, 16) */
  UCDIR_TC_I_NSPLIT = 30,  /* groups per NT-column work item that share one KC-channel activation slab (0 = 1) */
  /* batched GEMM use (the attention einsums, model/ucdir.py:174,179): SRC0 pixel rows are SRC_CSTRIDE elements apart
   * (0 = C0, lets a channel slice of a wider tensor be the A operand); W_BATCHED=1: W is per image, [B][NTOT][C0] with
   * rows W_ROWSTRIDE elements apart and images W_BATCHSTRIDE (LO + HI<<31) elements apart; f[ALPHA] scales the result */
  UCDIR_TC_I_SRC_CSTRIDE = 31, UCDIR_TC_I_W_BATCHED = 32, UCDIR_TC_I_W_ROWSTRIDE = 33,
  UCDIR_TC_I_W_BATCHSTRIDE_LO = 34, UCDIR_TC_I_W_BATCHSTRIDE_HI = 35,
  UCDIR_TC_I_T_COL0 = 36, UCDIR_TC_I_T_LD = 37,  /* transposed store of columns >= T_COL0 into DST2, row pitch T_LD */
  UCDIR_TC_I_W_ROWS = 38,                        /* batched weights: rows that exist per image (0 = NTOT); the rest read as 0 */
  UCDIR_TC_I_BSTAT = 39,                         /* 1: weight-stationary schedule for grouped mix convs (weights resident in smem, items N-tile major) */
  UCDIR_TC_I_SPS3 = 40,                          /* 1: three K slabs (filter taps) per pipeline stage for the small-N layers */
  UCDIR_TC_I_NO_CTAB = 42,                       /* 1: do not cache the folded-GroupNorm additive table of the current image in shared memory */
  UCDIR_TC_I_ROW3 = 41,                          /* 1: row tiles of dense 3x3 convs share one 130-pixel activation row among the three horizontal taps */
  UCDIR_TC_I_SRC_GN_SWISH = 44,                  /* 1: the source is first mapped through Swish(GroupNorm(1,C0)(SRC0)) (affine SRC_GAMMA / SRC_BETA, statistics
                                                  * STATS0, result rounded to bf16) -- final_conv, model/ucdir.py:266-268, whose Swish keeps the norm from being
                                                  * folded into the weights.  Applied to the landed halo box in shared memory (ucdir_fhalo.cu); needs a 3x3
                                                  * stride-1 conv of C0 <= 128 channels with GN = 0, NT = NTOT = 16 and DST_F32 = 1 */
  UCDIR_TC_I_RES_FUSED = 45,                     /* 1: also compute DST_RES = conv1x1(concat(SRC0, SRC1); W2) + TB2 (model/ucdir.py:120,140: res_conv(x) of the
                                                  * same un-normalised input as conv1) from the centre-tap view of the halo box already in shared memory;
                                                  * needs the 64-channel halo schedule (ucdir_dhalo.cu), refused otherwise */
  UCDIR_TC_I_DST_RES_C = 46,                     /* channels per pixel row of DST_RES */
  /* ---- fp32-tolerance mode ("fp32_tc", SURVEY 8d "Tolerances": split-operand MMA) --------------------------------------
   * SPLIT = 1: every bf16 tensor of the op is a PAIR of planes (hi, lo) with value = hi + lo (hi = bf16(v), lo = bf16(v - hi):
   * 16 mantissa bits), stored side by side in the pixel row: [hi: C channels | lo: C channels] (row pitch 2*C elements, so
   * C0 / C1 / DST_C / RES_C stay the LOGICAL channel counts).  The K loop runs three passes per filter tap,
   *     A_hi * W_hi  +  A_lo * W_hi  +  A_hi * W_lo      (fp32 accumulation in TMEM; the lo * lo term, 2^-18 relative, is dropped)
   * with weights packed per tap as [W_hi(src0) | W_hi(src0) | W_hi(src1) | W_hi(src1) | W_lo(src0) | W_lo(src1)]
   * (grouped: [W_hi | W_hi | W_lo] per group chunk).  The epilogue works in fp32 with the exact Swish, stores (hi, lo) pairs
   * and accumulates the GroupNorm statistics from the fp32 values.  DST_F32 outputs (attention scores, eps) are plain fp32.
   * DST2 (transposed V^T) rows are [hi: T_LD | lo: T_LD].  Only the streamed kernel implements it (no halo schedules). */
  UCDIR_TC_I_SPLIT = 47,
  UCDIR_TC_I_SRC_LO_OFF = 48,                    /* SPLIT: elements from the hi to the lo plane inside a SRC0 row (0 = C0); with SRC_CSTRIDE for channel slices */
  UCDIR_TC_I_W_LO_OFF = 49,                      /* SPLIT + W_BATCHED: elements from the hi to the lo plane inside a weight row (W_ROWSTRIDE = physical pitch) */
  UCDIR_TC_I_DST_CROP = 50,                      /* fused final conv only: DST is [B][H - 2*CROP][W - 2*CROP][DST_C] and receives only the interior of every
                                                  * sample -- the part of a tile that is stitched (utils/util.py:144-145) and, sharded, all-gathered; replaces a
                                                  * separate UCDIR_OP_CROP_TILES pass */
  UCDIR_TC_I_PHASES = 51,                        /* 4: nearest-2x upsample + conv3x3 (model/ucdir.py:53-60) as ONE launch: the four 2x2-tap phase convolutions
                                                  * are an extra work-item dimension; W = the phases' weight blocks stacked along N ([4*NTOT][K], phase = py*2 + px),
                                                  * OY0 = OX0 = -1, DST_UP = 1, DST_PY = DST_PX = 0 */
  UCDIR_TC_I_HALO = 43                           /* 1: halo schedule where it applies (grouped mix convs, C = 64 / 128 / 256): one 10 x 18 pixel TMA box per
                                                  * 8 x 16 pixel tile serves all nine taps, weights stay resident in shared memory (ucdir_mix.cu); 3x3 convs with 64 / 128 output
                                                  * channels: super tiles (ucdir_dhalo.cu).  With SPLIT = 1: the grouped mix convs with C = 64 / 128 / 256 (hi and lo boxes, both
                                                  * weight planes resident); every other SPLIT op runs the streamed schedule */
};
enum ucdir_tc_flt { UCDIR_TC_F_EPS = 0, UCDIR_TC_F_ALPHA = 1 /* 0 = 1.0 */ };

/* ---- UCDIR_OP_TC_ATTN: O = softmax(Q K^T * SCALE) V for one head of C = 512 channels, fused (ucdir_attn.cu) ---------------
 * SelfAttention.forward's two einsums and softmax (model/ucdir.py:174-179) in one kernel: scores in TMEM, online softmax in
 * registers, probabilities through shared memory; the N x N matrices never touch HBM.
 * QK = bf16 [B][N][QK_LD]: the qkv convolution's output rows, Q = channels [0, C), K = channels [C, 2C);
 * VT = bf16 [B][C][VT_LD]: V transposed (written by that convolution through UCDIR_TC_P_DST2); O = bf16 [B][N][O_LD].
 * f[SCALE] = 0 means 1 / sqrt(C) (model/ucdir.py:174). */
enum ucdir_attn_ptr { UCDIR_ATTN_P_QK = 0, UCDIR_ATTN_P_VT = 1, UCDIR_ATTN_P_O = 2 };
enum ucdir_attn_int { UCDIR_ATTN_I_B = 0, UCDIR_ATTN_I_N = 1, UCDIR_ATTN_I_C = 2, UCDIR_ATTN_I_QK_LD = 3, UCDIR_ATTN_I_VT_LD = 4, UCDIR_ATTN_I_O_LD = 5 };
enum ucdir_attn_flt { UCDIR_ATTN_F_SCALE = 0 };

/* ---- UCDIR_OP_GN_APPLY_BF16: DST = [Swish](GroupNorm(1,C)(SRC)) on bf16 NHWC [B][HW][C]; f[0] = eps ------------
 * SPLIT = 1: SRC and DST are (hi, lo) plane pairs [B][HW][2*C] (see UCDIR_TC_I_SPLIT), fp32 math, exact Swish. */
enum ucdir_gna_ptr { UCDIR_GNA_P_SRC = 0, UCDIR_GNA_P_DST = 1, UCDIR_GNA_P_GAMMA = 2, UCDIR_GNA_P_BETA = 3, UCDIR_GNA_P_STATS = 4 };
enum ucdir_gna_int { UCDIR_GNA_I_B = 0, UCDIR_GNA_I_HW = 1, UCDIR_GNA_I_C = 2, UCDIR_GNA_I_SWISH = 3, UCDIR_GNA_I_SPLIT = 4 };

/* ---- UCDIR_OP_CAST: p[1][k] = cast(p[0][k]) for k < i[0] + (i[1] << 31); i[2] = 0: fp32 -> bf16, 1: bf16 -> fp32 --
 * i[2] = 2 (fp32_tc attention probabilities): rows of fp32 -> rows of (hi, lo) bf16 plane pairs: for r < i[0] + (i[1] << 31),
 * c < i[3] (COLS): p[1][r][c] = hi, p[1][r][i[5] + c] = lo of p[0][r * i[4] + c]; i[4] = fp32 row pitch, i[5] = plane pitch
 * (output row = 2 * i[5] elements; columns COLS..i[5] of both planes are zeroed). */

/* ---- UCDIR_OP_GN_STATS_F32 / UCDIR_OP_GN_APPLY_F32: GroupNorm(G, C) with G >= 1 on fp32 NHWC [B][HW][C], for the SR3-style
 * FiLM ResnetBlock (model/ucdir.py:75-100).  STATS = double[B][G][2] {sum, sum of squares}; APPLY: f[0] = eps, SWISH as above. */
enum ucdir_gns_ptr { UCDIR_GNS_P_SRC = 0, UCDIR_GNS_P_STATS = 1 };
enum ucdir_gnf_ptr { UCDIR_GNF_P_SRC = 0, UCDIR_GNF_P_DST = 1, UCDIR_GNF_P_GAMMA = 2, UCDIR_GNF_P_BETA = 3, UCDIR_GNF_P_STATS = 4 };
enum ucdir_gns_int { UCDIR_GNS_I_B = 0, UCDIR_GNS_I_HW = 1, UCDIR_GNS_I_C = 2, UCDIR_GNS_I_G = 3, UCDIR_GNS_I_SWISH = 4 };

/* ---- UCDIR_OP_LAYOUT: fp32 layout change, p[0] -> p[1]; i = {B, C, HW, DIR}; DIR 0: NCHW -> NHWC, 1: NHWC -> NCHW ---------- */

/* ---- UCDIR_OP_TO_IMAGE_U8: DST[B][H-2*PD][W-2*PD][C] (uint8, HWC) = round(((clamp(SRC, MIN, MAX) - MIN) / (MAX - MIN)) * 255) of the
 * fp32 NCHW SRC[B][C][H][W] cropped by PD pixels per side: the caller-side tail of the path fused into one pass -- the
 * `[..., pd:-pd, pd:-pd]` crop of model/model.py:137 and core/metrics.py:8-34 (tensor2img: clamp, rescale, HWC, *255, round
 * half to even, uint8) -- so that one quarter of the bytes crosses PCIe.  f = {MIN, MAX}. */
enum ucdir_img_ptr { UCDIR_IMG_P_SRC = 0, UCDIR_IMG_P_DST = 1 };
enum ucdir_img_int { UCDIR_IMG_I_B = 0, UCDIR_IMG_I_C = 1, UCDIR_IMG_I_H = 2, UCDIR_IMG_I_W = 3, UCDIR_IMG_I_PD = 4 };
enum ucdir_img_flt { UCDIR_IMG_F_MIN = 0, UCDIR_IMG_F_MAX = 1 };

/* ---- UCDIR_OP_CROP_TILES: DST[BT,IH,IW,4] = SRC[BT,TH,TW,4][:, OY:OY+IH, OX:OX+IW] (fp32): the tile interiors that are
 * stitched (utils/util.py:144-145) and, in tile-sharded mode, all-gathered once per step ------------------------ */
enum ucdir_crop_ptr { UCDIR_CROP_P_SRC = 0, UCDIR_CROP_P_DST = 1 };
enum ucdir_crop_int { UCDIR_CROP_I_BT = 0, UCDIR_CROP_I_TH = 1, UCDIR_CROP_I_TW = 2, UCDIR_CROP_I_IH = 3, UCDIR_CROP_I_IW = 4,
                      UCDIR_CROP_I_OY = 5, UCDIR_CROP_I_OX = 6 };

/* ---- UCDIR_OP_MEMSET: cudaMemsetAsync(p[0], 0, i[0] + (i[1] << 31)) ------------------------------------ */

/* Run ops[0..n) in order on `stream` (a cudaStream_t).  No synchronisation.  Returns 0 or a negative error
 * (-1 bad argument, -2 unsupported shape, -3 CUDA launch error). */
int ucdir_run_ops(const ucdir_op_t* ops, int n_ops, void* stream);

/* Validate ops without launching (same return codes). */
int ucdir_check_ops(const ucdir_op_t* ops, int n_ops);

/* CUDA graphs: capture ops[0..n) once (tensor maps and launch parameters are baked into the nodes), replay with one
 * launch per step.  Everything that changes between steps must live in device memory the ops point at
 * (UCDIR_TEMB_P_LEVELS, UCDIR_SCATTER_P_PARAMS, fixed input / output buffers).  capture returns 0 and a handle. */
int ucdir_graph_capture(const ucdir_op_t* ops, int n_ops, void** graph_out);
int ucdir_graph_launch(void* graph, void* stream);
int ucdir_graph_destroy(void* graph);

/* Per-op device timing for benchmarks: between begin and end every op run by ucdir_run_ops is bracketed by CUDA
 * events on its stream.  ucdir_profile_end waits for the last event and fills up to `cap` records
 * {milliseconds, index of the op inside its ucdir_run_ops call, op kind}; returns the record count or <0. */
int ucdir_profile_begin(void);
int ucdir_profile_end(float* ms, int* op_index, int* kind, int cap);

int ucdir_abi_version(void);
int ucdir_op_sizeof(void);
const char* ucdir_last_error(void);
/* Number of kernel launches issued by this process through ucdir_run_ops since load. */
/* Which kernel a UCDIR_OP_TC_CONV record is routed to (no device needed): 0 = streamed (tc_conv_kernel), 1 = halo schedule of
 * the integration-module convs (ucdir_mix.cu), 2 = halo / super-tile schedule of the Cout = 64 / 128 3x3 convs (ucdir_dhalo.cu),
 * 3 = fused GroupNorm + Swish + conv of final_conv (ucdir_fhalo.cu); negative = not a TC_CONV op. */
int ucdir_tc_schedule(const ucdir_op_t* op);
long long ucdir_launch_count(void);
/* Device capability probe: returns 0 iff the current device is sm_100 (B200). */
int ucdir_device_ok(void);

/* ---- weight packing (host side, no device): OIHW fp32 checkpoint tensors -> kernel layouts ----------------------------------
 * For hosts that are not ucdir_b200/engine.py: everything UCDIR_OP_TC_CONV / UCDIR_OP_CONV_F32 expects in P_W / P_TB / P_TG can be
 * produced from a state_dict with these calls (done once after load_state_dict).  All pointers are HOST memory; outputs
 * are sized by the *_sizes / *_size queries (element counts); bf16 outputs are raw uint16.  Return >= 0 on success, < 0 on error.
 *   dense:   Conv2d weight [COUT][CIN][KS][KS] (KS 1 or 3) -> W bf16 [round_up(COUT, NT)][KS*KS*CIN (x3 when SPLIT)], TB fp32
 *            [NCLS][NTOT], TG fp32 [NCLS][NTOT] when GAMMA/BETA (the GroupNorm(1,C) in front of the conv) are given: NCLS = 9 border
 *            classes for KS = 3, 1 for KS = 1; without GAMMA: NCLS = 1, TB = bias, TG untouched.  C0 = channels of the first source
 *            of a concatenated input (SPLIT layout only; 0 = CIN).  Returns NCLS.
 *   grouped: spdyconv weight [COUT][CG][3][3] with GROUPS groups (model/ucdir.py:116), folded norm2; KC = MMA K chunk (engine:
 *            tc_mix_tiling(C)[1]); TB / TG fp32 [9][COUT].
 *   up_phase: Upsample conv [COUT][CIN][3][3] (model/ucdir.py:53-60) -> the 2x2-tap weights of output parity (PY, PX).
 *   conv_f32: the SIMT path's [GROUPS][KS*KS*CG][ld] layout (ld = round_up(COUT/GROUPS, 4)), CG zero padded to PAD_CIN_TO. */
int ucdir_pack_tc_dense_sizes(int cout, int cin, int ks, int nt, int split, int has_gn, int* w_elems, int* n_cls, int* n_tot);
int ucdir_pack_tc_dense(const float* w, const float* bias, const float* gamma, const float* beta, int cout, int cin, int ks, int nt,
                        int split, int c0, uint16_t* w_out, float* tb_out, float* tg_out);
int ucdir_pack_tc_grouped_sizes(int cout, int cg, int groups, int kc, int split, int* w_elems);
int ucdir_pack_tc_grouped(const float* w, const float* bias, const float* gamma, const float* beta, int cout, int cg, int groups, int kc,
                          int split, uint16_t* w_out, float* tb_out, float* tg_out);
int ucdir_pack_tc_up_phase_sizes(int cout, int cin, int nt, int split, int* w_elems, int* n_tot);
int ucdir_pack_tc_up_phase(const float* w, const float* bias, int cout, int cin, int py, int px, int nt, int split, uint16_t* w_out,
                           float* tb_out);
int ucdir_pack_conv_f32_size(int cout, int cg, int ks, int groups, int pad_cin_to);
int ucdir_pack_conv_f32(const float* w, int cout, int cg, int ks, int groups, int pad_cin_to, float* out);

#ifdef __cplusplus
}
#endif
#endif /* UCDIR_B200_H */
