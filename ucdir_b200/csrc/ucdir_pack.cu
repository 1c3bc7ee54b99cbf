// Host-side weight packing of the C ABI (include/ucdir_b200.h "ucdir_pack_*"): OIHW fp32 checkpoint tensors -> the layouts the
// kernels read, so that a non-Python host can drive libucdir_b200.so from a state_dict alone ("one
// ucdir_pack_weights_* per layer type ... done once after load_state_dict").  Pure host code: no device, no allocation; the caller
// sizes the outputs with the *_sizes queries.  ucdir_b200/engine.py keeps equivalent torch implementations (pack_tc_*), and
// tests/test_abi.py checks the two against each other (bf16 operands bit for bit, fp32 tables to summation order).
//
// What is folded here (see ucdir_tc.cu header):
//   * GroupNorm(1,C) in front of a convolution: gamma goes into the weights; the additive terms become per-border-class tables
//     TG[cls][n] = sum over the taps inside the image of sum_c W*gamma, TB[cls][n] = same with W*beta, + bias
//     (cls = cy*3 + cx, cy/cx in {first row/col, interior, last row/col}: model/ucdir.py:109-112,161 never touch HBM).
//   * grouped spdyconv (model/ucdir.py:116): K chunks of max(Cg, KC) channels, zero filled for the foreign groups.
//   * nearest-2x upsample + conv3x3 (model/ucdir.py:53-60) = four 2x2-tap phase convolutions with pre-summed weights.
//   * split-operand (fp32_tc) rows: per tap [W_hi(s0) | W_hi(s0) | W_hi(s1) | W_hi(s1) | W_lo(s0) | W_lo(s1)].
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "common.cuh"

namespace ucdir {

static inline uint16_t f2bf(float f) {                    // round to nearest even, NaN kept quiet
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline bool tap_inside(int cls, int ty, int tx) {  // is tap (ty,tx) of a 3x3 pad-1 conv inside the image for border class cls
  const int cy = cls / 3, cx = cls % 3;
  return !(cy == 0 && ty == 0) && !(cy == 2 && ty == 2) && !(cx == 0 && tx == 0) && !(cx == 2 && tx == 2);
}

// writes the K row of one output channel: plain [taps][ci] or split [taps][hi0 hi0 hi1 hi1 lo0 lo1]
static void put_row(uint16_t* dst, const float* wrow /* [taps][ci] */, int taps, int ci, int c0, int split) {
  if (!split) {
    for (int i = 0; i < taps * ci; ++i) dst[i] = f2bf(wrow[i]);
    return;
  }
  const int c1 = ci - c0;
  for (int t = 0; t < taps; ++t) {
    uint16_t* d = dst + (size_t)t * 3 * ci;
    const float* s = wrow + (size_t)t * ci;
    for (int c = 0; c < ci; ++c) {
      const uint16_t hi = f2bf(s[c]);
      const uint16_t lo = f2bf(s[c] - bf2f(hi));
      if (c < c0) { d[c] = hi; d[c0 + c] = hi; d[2 * c0 + 2 * c1 + c] = lo; }
      else { const int k = c - c0; d[2 * c0 + k] = hi; d[2 * c0 + c1 + k] = hi; d[3 * c0 + 2 * c1 + k] = lo; }
    }
  }
}

}  // namespace ucdir

using namespace ucdir;

extern "C" {

int ucdir_pack_conv_f32_size(int cout, int cg, int ks, int groups, int pad_cin_to) {
  if (cout <= 0 || cg <= 0 || ks <= 0 || groups <= 0 || cout % groups) return -1;
  const int cge = pad_cin_to > cg ? pad_cin_to : cg, ng = cout / groups, ldw = (ng + 3) & ~3;
  return groups * ks * ks * cge * ldw;
}
int ucdir_pack_conv_f32(const float* w, int cout, int cg, int ks, int groups, int pad_cin_to, float* out) {
  const int n = ucdir_pack_conv_f32_size(cout, cg, ks, groups, pad_cin_to);
  if (n < 0 || !w || !out) { set_error("pack_conv_f32: bad arguments"); return -1; }
  const int cge = pad_cin_to > cg ? pad_cin_to : cg, ng = cout / groups, ldw = (ng + 3) & ~3;
  memset(out, 0, (size_t)n * 4);
  for (int g = 0; g < groups; ++g)
    for (int o = 0; o < ng; ++o)
      for (int c = 0; c < cg; ++c)
        for (int t = 0; t < ks * ks; ++t)
          out[((size_t)g * ks * ks * cge + (size_t)t * cge + c) * ldw + o] = w[(((size_t)(g * ng + o) * cg + c) * ks * ks) + t];
  return 0;
}

int ucdir_pack_tc_dense_sizes(int cout, int cin, int ks, int nt, int split, int has_gn, int* w_elems, int* n_cls, int* n_tot) {
  if (cout <= 0 || cin <= 0 || (ks != 1 && ks != 3) || nt <= 0 || !w_elems || !n_cls || !n_tot) return -1;
  *n_tot = round_up(cout, nt);
  *w_elems = *n_tot * ks * ks * cin * (split ? 3 : 1);
  *n_cls = has_gn ? (ks == 3 ? 9 : 1) : 1;
  return 0;
}
int ucdir_pack_tc_dense(const float* w, const float* bias, const float* gamma, const float* beta, int cout, int cin, int ks, int nt, int split,
                        int c0, uint16_t* w_out, float* tb_out, float* tg_out) {
  int we, ncls, ntot;
  if (ucdir_pack_tc_dense_sizes(cout, cin, ks, nt, split, gamma != nullptr, &we, &ncls, &ntot) || !w || !w_out || !tb_out || (gamma && (!beta || !tg_out))) {
    set_error("pack_tc_dense: bad arguments"); return -1; }
  if (c0 <= 0 || c0 > cin) c0 = cin;
  const int taps = ks * ks, krow = taps * cin * (split ? 3 : 1);
  memset(w_out, 0, (size_t)we * 2);
  for (int i = 0; i < ncls * ntot; ++i) { tb_out[i] = 0.f; if (tg_out) tg_out[i] = 0.f; }
  std::vector<float> row((size_t)taps * cin);
  for (int n = 0; n < cout; ++n) {
    for (int t = 0; t < taps; ++t)
      for (int c = 0; c < cin; ++c) {
        const float v = w[((size_t)n * cin + c) * taps + t];
        row[(size_t)t * cin + c] = gamma ? v * gamma[c] : v;
      }
    put_row(w_out + (size_t)n * krow, row.data(), taps, cin, c0, split);
    const float b = bias ? bias[n] : 0.f;
    if (!gamma) { tb_out[n] = b; continue; }
    for (int cls = 0; cls < ncls; ++cls) {
      double sg = 0.0, sb = 0.0;
      for (int t = 0; t < taps; ++t) {
        if (ks == 3 && !tap_inside(cls, t / 3, t % 3)) continue;
        for (int c = 0; c < cin; ++c) {
          const float wg = row[(size_t)t * cin + c];
          sg += split ? (double)wg : (double)bf2f(f2bf(wg));          // bf16 mode: what the MMA really multiplies the mean with
          sb += (double)(w[((size_t)n * cin + c) * taps + t] * beta[c]);
        }
      }
      tg_out[(size_t)cls * ntot + n] = (float)sg;
      tb_out[(size_t)cls * ntot + n] = (float)sb + b;
    }
  }
  return ncls;
}

int ucdir_pack_tc_grouped_sizes(int cout, int cg, int groups, int kc, int split, int* w_elems) {
  if (cout <= 0 || cg <= 0 || groups <= 0 || cout % groups || kc <= 0 || !w_elems) return -1;
  const int cge = cg > kc ? cg : kc;
  *w_elems = cout * 9 * cge * (split ? 3 : 1);
  return 0;
}
int ucdir_pack_tc_grouped(const float* w, const float* bias, const float* gamma, const float* beta, int cout, int cg, int groups, int kc, int split,
                          uint16_t* w_out, float* tb_out /* [9][cout] */, float* tg_out /* [9][cout] */) {
  int we;
  if (ucdir_pack_tc_grouped_sizes(cout, cg, groups, kc, split, &we) || !w || !bias || !gamma || !beta || !w_out || !tb_out || !tg_out) {
    set_error("pack_tc_grouped: bad arguments"); return -1; }
  const int cge = cg > kc ? cg : kc, ng = cout / groups, krow = 9 * cge * (split ? 3 : 1);
  memset(w_out, 0, (size_t)we * 2);
  std::vector<float> row((size_t)9 * cge);
  for (int n = 0; n < cout; ++n) {
    const int g = n / ng;
    const int cbase = (g * cg) / cge * cge, off = g * cg - cbase;
    std::fill(row.begin(), row.end(), 0.f);
    for (int t = 0; t < 9; ++t)
      for (int c = 0; c < cg; ++c) row[(size_t)t * cge + off + c] = w[((size_t)n * cg + c) * 9 + t] * gamma[g * cg + c];
    put_row(w_out + (size_t)n * krow, row.data(), 9, cge, cge, split);
    for (int cls = 0; cls < 9; ++cls) {
      double sg = 0.0, sb = 0.0;
      for (int t = 0; t < 9; ++t) {
        if (!tap_inside(cls, t / 3, t % 3)) continue;
        for (int c = 0; c < cg; ++c) {
          const float wg = row[(size_t)t * cge + off + c];
          sg += split ? (double)wg : (double)bf2f(f2bf(wg));
          sb += (double)(w[((size_t)n * cg + c) * 9 + t] * beta[g * cg + c]);
        }
      }
      tg_out[(size_t)cls * cout + n] = (float)sg;
      tb_out[(size_t)cls * cout + n] = (float)sb + bias[n];
    }
  }
  return 9;
}

int ucdir_pack_tc_up_phase_sizes(int cout, int cin, int nt, int split, int* w_elems, int* n_tot) {
  if (cout <= 0 || cin <= 0 || nt <= 0 || !w_elems || !n_tot) return -1;
  *n_tot = round_up(cout, nt);
  *w_elems = *n_tot * 4 * cin * (split ? 3 : 1);
  return 0;
}
int ucdir_pack_tc_up_phase(const float* w, const float* bias, int cout, int cin, int py, int px, int nt, int split, uint16_t* w_out, float* tb_out) {
  int we, ntot;
  if (ucdir_pack_tc_up_phase_sizes(cout, cin, nt, split, &we, &ntot) || !w || !bias || !w_out || !tb_out || (py | px) & ~1) {
    set_error("pack_tc_up_phase: bad arguments"); return -1; }
  // output parity p, 2x2 tap t covers these rows of the 3x3 filter: p = 0 -> {0}, {1,2};  p = 1 -> {0,1}, {2}
  static const int lo[2][2] = {{0, 1}, {0, 2}}, hi[2][2] = {{0, 2}, {1, 2}};
  const int krow = 4 * cin * (split ? 3 : 1);
  memset(w_out, 0, (size_t)we * 2);
  for (int i = 0; i < ntot; ++i) tb_out[i] = i < cout ? bias[i] : 0.f;
  std::vector<float> row((size_t)4 * cin);
  for (int n = 0; n < cout; ++n) {
    for (int ty = 0; ty < 2; ++ty)
      for (int tx = 0; tx < 2; ++tx)
        for (int c = 0; c < cin; ++c) {
          float acc = 0.f;
          for (int dy = lo[py][ty]; dy <= hi[py][ty]; ++dy)
            for (int dx = lo[px][tx]; dx <= hi[px][tx]; ++dx) acc += w[((size_t)n * cin + c) * 9 + dy * 3 + dx];
          row[(size_t)(ty * 2 + tx) * cin + c] = acc;
        }
    put_row(w_out + (size_t)n * krow, row.data(), 4, cin, cin, split);
  }
  return 0;
}

}  // extern "C"
