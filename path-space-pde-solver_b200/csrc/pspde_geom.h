// pspde_geom.h -- on-chip layout of the network and the per-tile state (host + device, no CUDA types).
//
// Activation row (one per trajectory, in shared memory):
//   segment 0 = network input  [X | t | 1 | pad]   (t absent in TIME_NONE; X always starts at column 0 so that the
//                                                   SDE step can use aligned float4 accesses)
//   segment s = hidden layer s [h_s | (1) | pad]
// every segment is padded to a multiple of 4 floats so that all GEMM operands are float4-aligned.  The constant
// 1 column turns the bias into one more weight row (bias add and bias gradient come out of the GEMMs).
//   DenseNet (function_space.py:133-140): layer l reads the prefix [segment 0 .. segment l]
//   MySequential (:190-195):              layer l reads segment l only
// Weights live in shared memory in a k4-blocked layout: W_l[row r][col n] (row = activation column - in_start,
// i.e. the reference's input-major layout, function_space.py:122, with zero rows/cols at the pads) is stored at
//   w_off + ((r >> 2) * nng + (n >> 2)) * 16 + (r & 3) * 4 + (n & 3)
// so the four rows a thread needs for one k4 step of the forward GEMM are 64 contiguous bytes (one base register,
// immediate offsets) and a row segment W_l[r][n..n+3] is still one aligned float4 for the transposed GEMM.
#pragma once
#include <stdint.h>

#ifndef PSPDE_HD
#if defined(__CUDACC__)
#define PSPDE_HD __host__ __device__
#else
#define PSPDE_HD
#endif
#endif

#define PSPDE_MAXL 4

namespace pspde {

enum { NET_DENSENET = 0, NET_MLP_TANH = 1 };
enum { TIME_FIRST = 0, TIME_NONE = 1, TIME_LAST = 2 };

struct LayerGeom {
  int in_start;  // first activation column read by this layer
  int Kp;        // columns read (multiple of 4, pads included)
  int N;         // true output width
  int Np;        // output width padded to a multiple of 4
  int w_off;     // offset of W_l in the shared weight region (floats)
  int out_col;   // activation column of the first output (hidden layers); -1 for the output layer
  int th_w;      // offset of W_l in one parameter set of theta
  int th_b;      // offset of b_l
  int fan_in;    // true fan-in
  int nkg, nng;  // 4-row groups / 4-col groups (Kp / 4, Np / 4)
  int blk_begin; // first weight-gradient block of this layer in the global block list; a block owns the 4-row
  int kgh, ngh;  // groups {kg, kg + kgh} x the 4-col groups {ng, ng + ngh} (an 8x8 tile in two strided halves,
                 // so that neighbouring lanes read neighbouring float4s); kgh = ceil(nkg/2), ngh = ceil(nng/2)
};

struct NetGeom {
  int kind, L, time_mode;
  int dims[PSPDE_MAXL + 1];
  int seg_off[PSPDE_MAXL + 1];  // first column of segment s
  int seg_len[PSPDE_MAXL + 1];  // padded length
  int seg_one[PSPDE_MAXL + 1];  // column of the constant 1 (or -1)
  int x_col, t_col;             // column of X_0 / of t (-1: none)
  int d;                        // state dimension
  int lda;                      // activation row stride (floats), lda % 8 == 4
  int hid_off;                  // = seg_off[1]: delta tile column c <-> activation column hid_off + c
  int ldd;                      // delta tile row stride, ldd % 8 == 4
  int ldz;                      // row stride of the Z / xi tiles (d_out padded), ldz % 8 == 4
  int w_floats;                 // shared floats of all weights
  int n_params;                 // parameters in one set
  int n_blocks;                 // 8x8 weight-gradient blocks over all layers
  LayerGeom layer[PSPDE_MAXL];
};

PSPDE_HD inline int ceil4(int x) { return (x + 3) & ~3; }
PSPDE_HD inline int stride_cf(int x) {  // smallest s >= x with s % 8 == 4
  int s = ceil4(x);
  return (s & 7) == 4 ? s : s + 4;
}

// returns 0 on success
inline int build_geom(NetGeom& g, int kind, int L, const int* dims, int time_mode, int d) {
  if (L < 1 || L > PSPDE_MAXL) return -1;
  g.kind = kind; g.L = L; g.time_mode = time_mode; g.d = d;
  for (int i = 0; i <= L; ++i) { if (dims[i] < 1) return -2; g.dims[i] = dims[i]; }
  const int d_in_expected = d + (time_mode == TIME_NONE ? 0 : 1);
  if (dims[0] != d_in_expected) return -3;
  g.x_col = 0;
  g.t_col = (time_mode == TIME_NONE) ? -1 : d;
  int col = 0;
  for (int s = 0; s < L; ++s) {  // segments 0..L-1 are layer inputs
    const bool one = (kind == NET_MLP_TANH) || s == 0;
    g.seg_off[s] = col;
    g.seg_one[s] = one ? col + dims[s] : -1;
    g.seg_len[s] = ceil4(dims[s] + (one ? 1 : 0));
    col += g.seg_len[s];
  }
  g.seg_off[L] = col; g.seg_len[L] = 0; g.seg_one[L] = -1;
  g.lda = stride_cf(col);
  g.hid_off = (L > 1) ? g.seg_off[1] : col;
  g.ldd = stride_cf((L > 1) ? col - g.seg_off[1] : 4);
  g.ldz = stride_cf(dims[L]);
  int woff = 0, th = 0, blk = 0;
  for (int l = 0; l < L; ++l) {
    LayerGeom& y = g.layer[l];
    if (kind == NET_DENSENET) {
      y.in_start = 0; y.Kp = g.seg_off[l] + g.seg_len[l];
      y.fan_in = 0; for (int s = 0; s <= l; ++s) y.fan_in += dims[s];
    } else {
      y.in_start = g.seg_off[l]; y.Kp = g.seg_len[l]; y.fan_in = dims[l];
    }
    y.N = dims[l + 1]; y.Np = ceil4(y.N);
    y.w_off = woff; woff += y.Kp * y.Np;
    y.out_col = (l < L - 1) ? g.seg_off[l + 1] : -1;
    y.th_w = th; th += y.fan_in * y.N;
    y.th_b = th; th += y.N;
    y.nkg = y.Kp / 4; y.nng = y.Np / 4;
    y.kgh = (y.nkg + 1) / 2; y.ngh = (y.nng + 1) / 2;
    y.blk_begin = blk; blk += y.kgh * y.ngh;
  }
  g.w_floats = woff; g.n_params = th; g.n_blocks = blk;
  return 0;
}

// shared-memory offset (relative to w_off) of W_l[row r][col n] in the k4-blocked layout
PSPDE_HD inline int w_index(const LayerGeom& y, int r, int n) {
  return ((r >> 2) * y.nng + (n >> 2)) * 16 + (r & 3) * 4 + (n & 3);
}

// theta index (within one parameter set) of W_l[row r][col n] in the shared layout; -1 for pads.
PSPDE_HD inline int theta_index(const NetGeom& g, int l, int r, int n) {
  const LayerGeom& y = g.layer[l];
  if (n >= y.N || r >= y.Kp) return -1;
  const int col = y.in_start + r;
  int s = 0;
  while (s + 1 < g.L && col >= g.seg_off[s + 1]) ++s;   // segment containing this column
  const int i = col - g.seg_off[s];
  const bool bias_seg = (g.kind == NET_MLP_TANH) ? (s == l) : (s == 0);
  if (i == g.dims[s] && g.seg_one[s] == col) return bias_seg ? y.th_b + n : -1;
  if (i >= g.dims[s]) return -1;
  int fi = i;                                            // fan-in index inside segment s
  if (s == 0 && g.time_mode == TIME_FIRST) fi = (i == g.d) ? 0 : i + 1;   // columns [X | t]  <->  input [t, X]
  if (g.kind == NET_DENSENET) { for (int q = 0; q < s; ++q) fi += g.dims[q]; return y.th_w + fi * y.N + n; }
  return y.th_w + n * y.fan_in + fi;
}

// Weight-gradient block r (0 <= r < kgh * ngh) of a layer -> (kg, ng).  Blocks are ordered in column chunks of 8:
// the 8 lanes of a quarter-warp then read ONE 4-row group of the activations (a broadcast) and 8 consecutive
// float4s of the cotangent (one 128 B wavefront); with a plain row-major order a quarter-warp wraps around the
// end of a cotangent row and hits 2-way bank conflicts.
PSPDE_HD inline void bw_block_coords(const LayerGeom& y, int r, int& kg, int& ng) {
  const int full = (y.ngh >> 3) * 8 * y.kgh;     // blocks in the full-width chunks
  if (r < full) { const int c = r / (8 * y.kgh), q = r - c * 8 * y.kgh; kg = q >> 3; ng = 8 * c + (q & 7); }
  else { const int w = y.ngh & 7, q = r - full; kg = q / w; ng = (y.ngh & ~7) + q % w; }
}

}  // namespace pspde
