// Backward of the cross-attention Jacobian head (action-phase training).
//
// In the reference's action phase everything but the Jacobian head is frozen (models/model_wrapper.py:75-85,
// models/decoder/action_decoder_jacobian.py:251-258), and the loss is an MSE on the optical flow
// (model_wrapper.py:148-163), so the only gradient path is
//     optical_flow -> pw = p + Jbar^T u -> Jbar = sum_s w_s J_s -> J_s = head(q0_s) -> q0_s = W_q [enc | feat] + b_q .
// The train-mode forward is the fused render (field_kernel + xf_kernel); it leaves the fp16 query embedding q0 of
// every sample and the sample weights in the caller's workspace.  The kernels here take it from there:
//   xfb_layer_fwd  x3 : recompute the three attention / feed-forward layers in fp32, checkpoint x_0..x_3
//   xfb_head_bwd      : g x_3 = Wh^T (w_s g_Jbar[ray]);  d Wh, d bh
//   xfb_layer_bwd  x3 : recompute one layer from its checkpoint, back-propagate, accumulate d{M1,M2,W1,W2} (+ biases)
//   query_bwd_kernel  : d b_q, d W_q[:, :63] (positional-encoding columns) and the scatter of g q0 onto the context
//                       feature-map pixels (adjoint of the bilinear gather); d W_q[:, 63:] = (that map)^T . features
// Gradients are produced for the FOLDED matrices the forward uses (keys/values of the index embedding folded into
// M1 = scale W_q^T K, M2 = W_out V; LayerNorm affine folded into M1 / W1; field.cu), in fp32 with a natural-base
// softmax; the host chains them to the state-dict parameters through the (tiny, differentiable) fold.
//
// Layout of a thread block: 8 warps x 8 sample rows; a lane owns columns (lane, lane + 32) of every 64-wide row
// vector (registers).  y = W x reads W^T from shared memory ([k][n], row stride 65 floats: conflict-free both for
// y = W x and for g_x = W^T g) and broadcasts x from a per-warp staging row.  Weight gradients g (x) a are
// reduced over the block's 64 rows through two shared tiles; every warp keeps an 8 x 64 strip of each of the
// layer's four matrices in registers across all its tiles and flushes it once with atomicAdd.
#include <cuda_runtime.h>

#include "../../include/njf_b200.h"
#include "field.h"
#include "njf_internal.h"
#include "render.cuh"

namespace njf {

constexpr int kBR = 8;            // rows per warp
constexpr int kBWarps = 8;        // warps per block
constexpr int kBRows = kBR * kBWarps;  // 64 rows per block step = half a 128-row hand-over tile
constexpr int kWS = 65;           // row stride of the transposed weight images
constexpr int kMatFloats = 64 * kWS;

struct Vec {
  float lo[kBR], hi[kBR];
};

struct BwdSmem {
  float wt[4][kMatFloats];     // transposed weights of the current layer: wt[m][k * 65 + n] = W_m[n][k]
  float bias[4][64];
  float xs[kBWarps][kBR][64];  // per-warp broadcast staging of a row vector
  float gs[kBRows][64];        // weight-gradient staging: g rows
  float as[kBRows][64];        //                          a rows
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ float group8_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  return v;
}

__device__ __forceinline__ void stage(float (*xs)[64], const Vec& x, int lane) {
  __syncwarp();
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    xs[r][lane] = x.lo[r];
    xs[r][32 + lane] = x.hi[r];
  }
  __syncwarp();
}

// y[n] = sum_k W[n][k] x[k] + b[n]     (wt = W^T image, lane owns n = lane, lane + 32)
__device__ __forceinline__ void matvec(const float* __restrict__ wt, const float* __restrict__ bias, const Vec& x, Vec& y,
                                       float (*xs)[64], int lane) {
  stage(xs, x, lane);
  const float b0 = bias ? bias[lane] : 0.f, b1 = bias ? bias[32 + lane] : 0.f;
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    y.lo[r] = b0;
    y.hi[r] = b1;
  }
#pragma unroll 2
  for (int k4 = 0; k4 < 16; ++k4) {
    float4 xv[kBR];
#pragma unroll
    for (int r = 0; r < kBR; ++r) xv[r] = *reinterpret_cast<const float4*>(&xs[r][4 * k4]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float w0 = wt[(4 * k4 + j) * kWS + lane], w1 = wt[(4 * k4 + j) * kWS + 32 + lane];
#pragma unroll
      for (int r = 0; r < kBR; ++r) {
        const float xk = j == 0 ? xv[r].x : j == 1 ? xv[r].y : j == 2 ? xv[r].z : xv[r].w;
        y.lo[r] = fmaf(w0, xk, y.lo[r]);
        y.hi[r] = fmaf(w1, xk, y.hi[r]);
      }
    }
  }
}

// gx[k] = sum_n W[n][k] g[n]     (lane owns k = lane, lane + 32: reads rows lane / lane + 32 of the W^T image)
__device__ __forceinline__ void matvec_t(const float* __restrict__ wt, const Vec& g, Vec& gx, float (*xs)[64], int lane) {
  stage(xs, g, lane);
#pragma unroll
  for (int r = 0; r < kBR; ++r) gx.lo[r] = gx.hi[r] = 0.f;
#pragma unroll 2
  for (int n4 = 0; n4 < 16; ++n4) {
    float4 gv[kBR];
#pragma unroll
    for (int r = 0; r < kBR; ++r) gv[r] = *reinterpret_cast<const float4*>(&xs[r][4 * n4]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float w0 = wt[lane * kWS + 4 * n4 + j], w1 = wt[(32 + lane) * kWS + 4 * n4 + j];
#pragma unroll
      for (int r = 0; r < kBR; ++r) {
        const float gn = j == 0 ? gv[r].x : j == 1 ? gv[r].y : j == 2 ? gv[r].z : gv[r].w;
        gx.lo[r] = fmaf(w0, gn, gx.lo[r]);
        gx.hi[r] = fmaf(w1, gn, gx.hi[r]);
      }
    }
  }
}

// LayerNorm without affine (eps 1e-5, biased variance: nn.LayerNorm; the affine is folded into the next matrix)
__device__ __forceinline__ void ln_hat(const Vec& x, Vec& n, float (&rstd)[kBR]) {
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    const float mean = warp_sum(x.lo[r] + x.hi[r]) * (1.f / 64.f);
    const float d0 = x.lo[r] - mean, d1 = x.hi[r] - mean;
    const float var = warp_sum(d0 * d0 + d1 * d1) * (1.f / 64.f);
    rstd[r] = rsqrtf(var + 1e-5f);
    n.lo[r] = d0 * rstd[r];
    n.hi[r] = d1 * rstd[r];
  }
}
// gx += rstd * (g - mean(g) - n * mean(g . n))
__device__ __forceinline__ void ln_hat_bwd_add(const Vec& g, const Vec& n, const float (&rstd)[kBR], Vec& gx) {
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    const float mg = warp_sum(g.lo[r] + g.hi[r]) * (1.f / 64.f);
    const float mgn = warp_sum(g.lo[r] * n.lo[r] + g.hi[r] * n.hi[r]) * (1.f / 64.f);
    gx.lo[r] += rstd[r] * (g.lo[r] - mg - n.lo[r] * mgn);
    gx.hi[r] += rstd[r] * (g.hi[r] - mg - n.hi[r] * mgn);
  }
}
// softmax over the A real keys of each head (columns h*8 + a; lanes 8j..8j+7 of lo hold head j, of hi head 4 + j)
__device__ __forceinline__ void softmax_heads(const Vec& l, Vec& p, int A, int lane) {
  const bool real = (lane & 7) < A;
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    const float a = real ? l.lo[r] : -3.0e38f, b = real ? l.hi[r] : -3.0e38f;
    const float ma = group8_max(a), mb = group8_max(b);
    const float ea = real ? expf(a - ma) : 0.f, eb = real ? expf(b - mb) : 0.f;
    p.lo[r] = ea / group8_sum(ea);
    p.hi[r] = eb / group8_sum(eb);
  }
}
// gl = p * (gp - sum_head p gp)
__device__ __forceinline__ void softmax_heads_bwd(const Vec& p, const Vec& gp, Vec& gl) {
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    const float sa = group8_sum(p.lo[r] * gp.lo[r]), sb = group8_sum(p.hi[r] * gp.hi[r]);
    gl.lo[r] = p.lo[r] * (gp.lo[r] - sa);
    gl.hi[r] = p.hi[r] * (gp.hi[r] - sb);
  }
}
__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_d(float v) {
  return 0.5f * (1.f + erff(v * 0.70710678118654752f)) + v * 0.3989422804014327f * expf(-0.5f * v * v);
}

// acc[j][0/1] += sum over the block's 64 rows of g_row[8 warp + j] * a_row[lane / lane + 32]
__device__ __forceinline__ void outer_acc(BwdSmem* sm, const Vec& g, const Vec& a, float (&acc)[8][2], int warp, int lane) {
  __syncthreads();  // previous readers of the tiles are done
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    sm->gs[warp * kBR + r][lane] = g.lo[r];
    sm->gs[warp * kBR + r][32 + lane] = g.hi[r];
    sm->as[warp * kBR + r][lane] = a.lo[r];
    sm->as[warp * kBR + r][32 + lane] = a.hi[r];
  }
  __syncthreads();
#pragma unroll 4
  for (int r = 0; r < kBRows; ++r) {
    const float a0 = sm->as[r][lane], a1 = sm->as[r][32 + lane];
    const float4 g0 = *reinterpret_cast<const float4*>(&sm->gs[r][8 * warp]);
    const float4 g1 = *reinterpret_cast<const float4*>(&sm->gs[r][8 * warp + 4]);
    const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j][0] = fmaf(gv[j], a0, acc[j][0]);
      acc[j][1] = fmaf(gv[j], a1, acc[j][1]);
    }
  }
}
__device__ __forceinline__ void bias_acc(const Vec& g, float (&acc)[2]) {
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    acc[0] += g.lo[r];
    acc[1] += g.hi[r];
  }
}
// d W[n][k] (row-major [64][64]) += this warp's strip; d b += this warp's row sums
__device__ __forceinline__ void flush_mat(float* __restrict__ dW, const float (&acc)[8][2], int warp, int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(dW + (8 * warp + j) * 64 + lane, acc[j][0]);
    atomicAdd(dW + (8 * warp + j) * 64 + 32 + lane, acc[j][1]);
  }
}
__device__ __forceinline__ void flush_bias(float* __restrict__ db, const float (&acc)[2], int lane) {
  atomicAdd(db + lane, acc[0]);
  atomicAdd(db + 32 + lane, acc[1]);
}

// W (row-major [64][64] fp32 in global memory) -> transposed, padded shared image; biases
__device__ __forceinline__ void load_weights(BwdSmem* sm, const float* const (&W)[4], const float* const (&b)[4], int nmat) {
  for (int m = 0; m < nmat; ++m) {
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
      const int n = i >> 6, k = i & 63;
      sm->wt[m][k * kWS + n] = __ldg(W[m] + i);
    }
    if (threadIdx.x < 64) sm->bias[m][threadIdx.x] = b[m] ? __ldg(b[m] + threadIdx.x) : 0.f;
  }
  __syncthreads();
}

__device__ __forceinline__ void load_rows(const float* __restrict__ src, size_t row0, Vec& v, int lane) {
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    v.lo[r] = src[(row0 + r) * 64 + lane];
    v.hi[r] = src[(row0 + r) * 64 + 32 + lane];
  }
}
__device__ __forceinline__ void store_rows(float* __restrict__ dst, size_t row0, const Vec& v, int lane) {
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    dst[(row0 + r) * 64 + lane] = v.lo[r];
    dst[(row0 + r) * 64 + 32 + lane] = v.hi[r];
  }
}

// packed layout of the folded parameters (and of their gradients): per layer
// [M1 4096 | m1b 64 | M2 4096 | bo 64 | W1 4096 | w1b 64 | W2 4096 | b2 64], then [Wh 4096 (rows >= 3A zero) | bh 64]
constexpr int kLayerFloats = 4 * (4096 + 64);
constexpr int kFoldedFloats = 3 * kLayerFloats + 4096 + 64;
__host__ __device__ inline int off_mat(int layer, int m) { return layer * kLayerFloats + m * (4096 + 64); }

struct XfbParams {
  const float* folded;   // [kFoldedFloats]
  float* g_folded;       // [kFoldedFloats], accumulated
  int A;
  int n_units;           // 64-row block steps = 2 * hand-over tiles
  const uint4* qs;       // hand-over: [tile][8][128] x 8 fp16
  const float* wts;      // [tile][128]
  float* chk;            // [4][rows][64] checkpoints x_0..x_3
  float* gx;             // [rows][64] running gradient (in place); ends as g q0
  const float* g_jbar;   // [NR][3A]
  int NR, S, G, T;
  int layer;
};

__device__ __forceinline__ void layer_forward(BwdSmem* sm, const Vec& x, int A, int warp, int lane, Vec& n1, float (&rstd1)[kBR],
                                              Vec& p, Vec& xmid, Vec& n2, float (&rstd2)[kBR], Vec& f, Vec& xout) {
  float(*xs)[64] = sm->xs[warp];
  ln_hat(x, n1, rstd1);
  Vec l;
  matvec(sm->wt[0], sm->bias[0], n1, l, xs, lane);
  softmax_heads(l, p, A, lane);
  matvec(sm->wt[1], sm->bias[1], p, xmid, xs, lane);
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    xmid.lo[r] += x.lo[r];
    xmid.hi[r] += x.hi[r];
  }
  ln_hat(xmid, n2, rstd2);
  matvec(sm->wt[2], sm->bias[2], n2, f, xs, lane);
  Vec h;
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    h.lo[r] = gelu_f(f.lo[r]);
    h.hi[r] = gelu_f(f.hi[r]);
  }
  matvec(sm->wt[3], sm->bias[3], h, xout, xs, lane);
#pragma unroll
  for (int r = 0; r < kBR; ++r) {
    xout.lo[r] += xmid.lo[r];
    xout.hi[r] += xmid.hi[r];
  }
}

__global__ void __launch_bounds__(kBWarps * 32, 1) xfb_layer_fwd(const XfbParams p) {
  extern __shared__ uint8_t smem_raw[];
  BwdSmem* sm = reinterpret_cast<BwdSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = p.folded + off_mat(p.layer, 0);
  const float* const W[4] = {base, base + 4160, base + 2 * 4160, base + 3 * 4160};
  const float* const b[4] = {base + 4096, base + 4160 + 4096, base + 2 * 4160 + 4096, base + 3 * 4160 + 4096};
  load_weights(sm, W, b, 4);
  const size_t rows = static_cast<size_t>(p.n_units) * kBRows;
  for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
    const size_t row0 = static_cast<size_t>(u) * kBRows + warp * kBR;
    Vec x;
    if (p.layer == 0) {  // x_0 = q0 from the fp16 hand-over stream of the forward
      const size_t tile = row0 >> 7;
      const int trow = static_cast<int>(row0 & 127);
      const __half* q = reinterpret_cast<const __half*>(p.qs + tile * 8 * kRows);
#pragma unroll
      for (int r = 0; r < kBR; ++r) {
        x.lo[r] = __half2float(q[((lane >> 3) * kRows + trow + r) * 8 + (lane & 7)]);
        x.hi[r] = __half2float(q[((4 + (lane >> 3)) * kRows + trow + r) * 8 + (lane & 7)]);
      }
      store_rows(p.chk, row0, x, lane);
    } else {
      load_rows(p.chk + static_cast<size_t>(p.layer) * rows * 64, row0, x, lane);
    }
    Vec n1, pr, xmid, n2, f, xout;
    float r1[kBR], r2[kBR];
    layer_forward(sm, x, p.A, warp, lane, n1, r1, pr, xmid, n2, r2, f, xout);
    store_rows(p.chk + static_cast<size_t>(p.layer + 1) * rows * 64, row0, xout, lane);
  }
}

// row of a hand-over tile -> ray (the tiling of render.cuh row_setup / xf_kernel)
__device__ __forceinline__ int row_ray(const XfbParams& p, size_t grow) {
  const int tidx = static_cast<int>(grow >> 7), row = static_cast<int>(grow & 127);
  const int lgroup = tidx / p.T, tile = tidx - lgroup * p.T;
  if (p.T == 1) {
    const int lr = row / p.S;
    const int ray = lgroup * p.G + lr;
    return (lr < p.G && ray < p.NR) ? ray : -1;
  }
  return (tile * kRows + row < p.S && lgroup < p.NR) ? lgroup : -1;
}

__global__ void __launch_bounds__(kBWarps * 32, 1) xfb_head_bwd(const XfbParams p) {
  extern __shared__ uint8_t smem_raw[];
  BwdSmem* sm = reinterpret_cast<BwdSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* wh = p.folded + 3 * kLayerFloats;
  const float* const W[4] = {wh, nullptr, nullptr, nullptr};
  const float* const b[4] = {nullptr, nullptr, nullptr, nullptr};
  load_weights(sm, W, b, 1);
  const size_t rows = static_cast<size_t>(p.n_units) * kBRows;
  const int A3 = 3 * p.A;
  float acc[8][2], bacc[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.f;
  for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
    const size_t row0 = static_cast<size_t>(u) * kBRows + warp * kBR;
    Vec x3, gj, gx;
    load_rows(p.chk + 3 * rows * 64, row0, x3, lane);
#pragma unroll
    for (int r = 0; r < kBR; ++r) {  // g J_s = w_s * g Jbar[ray]   (Jbar = sum_s w_s J_s, models/model.py:281-286)
      const int ray = row_ray(p, row0 + r);
      const float w = p.wts[row0 + r];
      gj.lo[r] = (ray >= 0 && lane < A3) ? w * __ldg(p.g_jbar + static_cast<size_t>(ray) * A3 + lane) : 0.f;
      gj.hi[r] = 0.f;
    }
    matvec_t(sm->wt[0], gj, gx, sm->xs[warp], lane);
    store_rows(p.gx, row0, gx, lane);
    outer_acc(sm, gj, x3, acc, warp, lane);
    bias_acc(gj, bacc);
  }
  flush_mat(p.g_folded + 3 * kLayerFloats, acc, warp, lane);
  flush_bias(p.g_folded + 3 * kLayerFloats + 4096, bacc, lane);
}

__global__ void __launch_bounds__(kBWarps * 32, 1) xfb_layer_bwd(const XfbParams p) {
  extern __shared__ uint8_t smem_raw[];
  BwdSmem* sm = reinterpret_cast<BwdSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = p.folded + off_mat(p.layer, 0);
  const float* const W[4] = {base, base + 4160, base + 2 * 4160, base + 3 * 4160};
  const float* const b[4] = {base + 4096, base + 4160 + 4096, base + 2 * 4160 + 4096, base + 3 * 4160 + 4096};
  load_weights(sm, W, b, 4);
  const size_t rows = static_cast<size_t>(p.n_units) * kBRows;
  float aM1[8][2], aM2[8][2], aW1[8][2], aW2[8][2];
  float bM1[2] = {0.f, 0.f}, bM2[2] = {0.f, 0.f}, bW1[2] = {0.f, 0.f}, bW2[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) aM1[j][0] = aM1[j][1] = aM2[j][0] = aM2[j][1] = aW1[j][0] = aW1[j][1] = aW2[j][0] = aW2[j][1] = 0.f;
  float(*xs)[64] = sm->xs[warp];
  for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
    const size_t row0 = static_cast<size_t>(u) * kBRows + warp * kBR;
    Vec x, n1, pr, xmid, n2, f, xout;
    float r1[kBR], r2[kBR];
    load_rows(p.chk + static_cast<size_t>(p.layer) * rows * 64, row0, x, lane);
    layer_forward(sm, x, p.A, warp, lane, n1, r1, pr, xmid, n2, r2, f, xout);
    Vec g;
    load_rows(p.gx, row0, g, lane);
    // ---- feed-forward: x_out = x_mid + W2 gelu(W1 n2 + w1b) + b2
    {
      Vec h;
#pragma unroll
      for (int r = 0; r < kBR; ++r) {
        h.lo[r] = gelu_f(f.lo[r]);
        h.hi[r] = gelu_f(f.hi[r]);
      }
      outer_acc(sm, g, h, aW2, warp, lane);
      bias_acc(g, bW2);
    }
    Vec gf;
    matvec_t(sm->wt[3], g, gf, xs, lane);
#pragma unroll
    for (int r = 0; r < kBR; ++r) {
      gf.lo[r] *= gelu_d(f.lo[r]);
      gf.hi[r] *= gelu_d(f.hi[r]);
    }
    outer_acc(sm, gf, n2, aW1, warp, lane);
    bias_acc(gf, bW1);
    {
      Vec gn;
      matvec_t(sm->wt[2], gf, gn, xs, lane);
      ln_hat_bwd_add(gn, n2, r2, g);  // g = d loss / d x_mid
    }
    // ---- attention: x_mid = x + M2 softmax(M1 n1 + m1b) + bo
    outer_acc(sm, g, pr, aM2, warp, lane);
    bias_acc(g, bM2);
    Vec gl;
    {
      Vec gp;
      matvec_t(sm->wt[1], g, gp, xs, lane);
      softmax_heads_bwd(pr, gp, gl);
    }
    outer_acc(sm, gl, n1, aM1, warp, lane);
    bias_acc(gl, bM1);
    {
      Vec gn;
      matvec_t(sm->wt[0], gl, gn, xs, lane);
      ln_hat_bwd_add(gn, n1, r1, g);  // g = d loss / d x (layer input)
    }
    store_rows(p.gx, row0, g, lane);
  }
  float* gb = p.g_folded + off_mat(p.layer, 0);
  flush_mat(gb, aM1, warp, lane);
  flush_bias(gb + 4096, bM1, lane);
  flush_mat(gb + 4160, aM2, warp, lane);
  flush_bias(gb + 4160 + 4096, bM2, lane);
  flush_mat(gb + 2 * 4160, aW1, warp, lane);
  flush_bias(gb + 2 * 4160 + 4096, bW1, lane);
  flush_mat(gb + 3 * 4160, aW2, warp, lane);
  flush_bias(gb + 3 * 4160 + 4096, bW2, lane);
}

// ---- q0 = W_q [enc63 | feat512] + b_q (action_decoder_jacobian.py:423-430): gradients of the query MLP
// One warp per 128-row hand-over tile slice of 8 rows; geometry recomputed with the forward's own row_setup.
struct QueryBwdParams {
  PassGeom g;
  int n_tiles;
  const float* gq0;     // [tiles * 128][64]
  float* g_wq_enc;      // [64][64]: columns 0..62 = d W_q[:, :63] (nerfstudio column order), column 63 unused
  float* g_bq;          // [64]
  float* g_map;         // [B][Hf*Wf][64]: sum over samples of (bilinear tap weight) * g q0, zero-initialised by the caller
};

__global__ void __launch_bounds__(kBWarps * 32, 1) query_bwd_kernel(const QueryBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  BwdSmem* sm = reinterpret_cast<BwdSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const PassGeom& g = p.g;
  float acc[8][2], bacc[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.f;
  const int n_units = p.n_tiles * 2;
  for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
    const size_t row0 = static_cast<size_t>(u) * kBRows + warp * kBR;
    Vec gq, enc;
    load_rows(p.gq0, row0, gq, lane);
#pragma unroll
    for (int r = 0; r < kBR; ++r) {
      const size_t grow = row0 + r;
      const int tidx = static_cast<int>(grow >> 7), row = static_cast<int>(grow & 127);
      const int group = tidx / g.T, tile = tidx - group * g.T;
      RowState rs;
      row_setup(g, group, tile, row, rs);
      const bool valid = rs.ray >= 0;
      // positional encoding columns (nerfstudio order: sin block dim-major / freq-minor, cos block, xyz)
      float e0 = 0.f, e1 = 0.f;
      if (valid) {
        {
          const int c = lane;  // 0..31: sin block (30) + first 2 cos columns
          const int cc = c < 30 ? c : c - 30, i = cc / 10, k = cc - 10 * i;
          float t = __fmul_rn(6.2831855f, rs.cam[i]) * static_cast<float>(1 << k);
          if (c >= 30) t = __fadd_rn(t, 1.5707964f);
          e0 = sin_cw(t);
        }
        {
          const int c = 32 + lane;  // 32..62: cos block remainder (28) + xyz (3); 63: padding
          if (c < 60) {
            const int cc = c - 30, i = cc / 10, k = cc - 10 * i;
            e1 = sin_cw(__fadd_rn(__fmul_rn(6.2831855f, rs.cam[i]) * static_cast<float>(1 << k), 1.5707964f));
          } else if (c < 63) {
            e1 = rs.cam[c - 60];
          }
        }
      } else {
        gq.lo[r] = gq.hi[r] = 0.f;  // padding rows carry no gradient
      }
      enc.lo[r] = e0;
      enc.hi[r] = e1;
      // adjoint of the bilinear gather: g_map[pixel][n] += tap weight * g q0[n]
      if (valid) {
        const float x0 = floorf(rs.ix), y0 = floorf(rs.iy);
        const float x1 = x0 + 1.f, y1 = y0 + 1.f;
        const float wv[4] = {(x1 - rs.ix) * (y1 - rs.iy), (rs.ix - x0) * (y1 - rs.iy), (x1 - rs.ix) * (rs.iy - y0),
                             (rs.ix - x0) * (rs.iy - y0)};
        const int xi = static_cast<int>(x0), yi = static_cast<int>(y0);
        const int xj = min(xi + 1, g.Wf - 1), yj = min(yi + 1, g.Hf - 1);
        const int px[4] = {yi * g.Wf + xi, yi * g.Wf + xj, yj * g.Wf + xi, yj * g.Wf + xj};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (wv[t] != 0.f) {
            float* dst = p.g_map + (static_cast<size_t>(rs.pixbase) + px[t]) * 64;
            atomicAdd(dst + lane, wv[t] * gq.lo[r]);
            atomicAdd(dst + 32 + lane, wv[t] * gq.hi[r]);
          }
        }
      }
    }
    outer_acc(sm, gq, enc, acc, warp, lane);
    bias_acc(gq, bacc);
  }
  flush_mat(p.g_wq_enc, acc, warp, lane);
  flush_bias(p.g_bq, bacc, lane);
}

// ---- backward of finish_kernel's flow: flow = proj(pw) - proj(p), pw = p + Jbar^T u   (models/model.py:288-314)
struct FlowBwdParams {
  int NR, R, A;
  const float* g_flow;    // [NR][2]
  const float* g_pw_in;   // [NR][3] or null: gradient arriving directly at ray_positions_warped
  const float* jbar;      // [NR][3A]
  const float* p;         // [NR][3]
  const float* action;    // [B][A]
  const float* trgt_w2c;  // [B][16]
  const float* trgt_k;    // [B][9]
  float* g_jbar;          // [NR][3A] or null
  float* g_action;        // [B][A] or null (accumulated with atomicAdd; zeroed by the launcher)
};
__global__ void flow_bwd_kernel(const FlowBwdParams q) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = ray < q.NR;
  const int rr = live ? ray : q.NR - 1;
  const int b = rr / q.R;
  const int A3 = 3 * q.A;
  float px = q.p[rr * 3], py = q.p[rr * 3 + 1], pz = q.p[rr * 3 + 2];
  float f[3] = {0.f, 0.f, 0.f};
  for (int a = 0; a < q.A; ++a) {
    const float ua = __ldg(q.action + b * q.A + a);
#pragma unroll
    for (int d = 0; d < 3; ++d) f[d] = fmaf(q.jbar[static_cast<size_t>(rr) * A3 + a * 3 + d], ua, f[d]);
  }
  const float wx = px + f[0], wy = py + f[1], wz = pz + f[2];
  // uv = (K c)_{0,1} / ((K c)_2 + 1e-9), c = W[:3,:3] x + W[:3,3]
  const float* W = q.trgt_w2c + b * 16;
  const float* K = q.trgt_k + b * 9;
  float c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) c[i] = fmaf(W[4 * i + 2], wz, fmaf(W[4 * i + 1], wy, fmaf(W[4 * i], wx, W[4 * i + 3])));
  const float ka = fmaf(K[2], c[2], fmaf(K[1], c[1], K[0] * c[0]));
  const float kb = fmaf(K[5], c[2], fmaf(K[4], c[1], K[3] * c[0]));
  const float kw = fmaf(K[8], c[2], fmaf(K[7], c[1], K[6] * c[0])) + 1e-9f;
  const float gu = live ? q.g_flow[ray * 2] : 0.f, gv = live ? q.g_flow[ray * 2 + 1] : 0.f;
  // d loss / d (K c): (gu / kw, gv / kw, -(gu ka + gv kb) / kw^2)
  const float gk[3] = {gu / kw, gv / kw, -(gu * ka + gv * kb) / (kw * kw)};
  float gc[3], gx[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) gc[j] = K[j] * gk[0] + K[3 + j] * gk[1] + K[6 + j] * gk[2];
#pragma unroll
  for (int j = 0; j < 3; ++j) gx[j] = W[j] * gc[0] + W[4 + j] * gc[1] + W[8 + j] * gc[2];
  if (q.g_pw_in && live)
#pragma unroll
    for (int j = 0; j < 3; ++j) gx[j] += q.g_pw_in[ray * 3 + j];
  if (q.g_jbar && live)
    for (int a = 0; a < q.A; ++a) {
      const float ua = __ldg(q.action + b * q.A + a);
#pragma unroll
      for (int d = 0; d < 3; ++d) q.g_jbar[static_cast<size_t>(ray) * A3 + a * 3 + d] = ua * gx[d];
    }
  if (q.g_action) {
    // all 32 rays of a warp usually share the view: reduce before the atomic
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    const bool same = __all_sync(0xffffffffu, b == b0);
    for (int a = 0; a < q.A; ++a) {
      float ga = 0.f;
      if (live)
#pragma unroll
        for (int d = 0; d < 3; ++d) ga = fmaf(q.jbar[static_cast<size_t>(ray) * A3 + a * 3 + d], gx[d], ga);
      if (same) {
        ga = warp_sum(ga);
        if (lane == 0) atomicAdd(q.g_action + b0 * q.A + a, ga);
      } else if (live) {
        atomicAdd(q.g_action + b * q.A + a, ga);
      }
    }
  }
}

}  // namespace njf

using namespace njf;

namespace {
int bwd_grid() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}
template <class K>
int bwd_smem(K kernel) {
  NJF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(BwdSmem))));
  return 0;
}
}  // namespace

extern "C" int njf_xf_folded_floats(void) { return kFoldedFloats; }

extern "C" size_t njf_xf_backward_workspace_bytes(int n_tiles) {
  return static_cast<size_t>(n_tiles) * kRows * 64 * sizeof(float) * 4;  // checkpoints x_0..x_3
}

extern "C" int njf_xf_backward(const float* folded, int action_dim, const void* handover, int n_tiles, int n_rays,
                               int s_nerf, const float* g_jbar, float* g_folded, float* g_q0, void* workspace,
                               size_t workspace_bytes, void* stream_) {
  if (!folded || !handover || !g_jbar || !g_folded || !g_q0 || !workspace) NJF_FAIL("njf_xf_backward: null argument");
  if (action_dim < 1 || action_dim > 8) NJF_FAIL("njf_xf_backward: action_dim %d unsupported (1..8)", action_dim);
  if (n_tiles < 1 || n_rays < 1 || s_nerf < 1 || s_nerf > 512) NJF_FAIL("njf_xf_backward: bad sizes");
  if (workspace_bytes < njf_xf_backward_workspace_bytes(n_tiles)) NJF_FAIL("njf_xf_backward: workspace too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  XfbParams p{};
  p.folded = folded;
  p.g_folded = g_folded;
  p.A = action_dim;
  p.n_units = 2 * n_tiles;
  p.qs = static_cast<const uint4*>(handover);
  p.wts = reinterpret_cast<const float*>(static_cast<const uint8_t*>(handover) + static_cast<size_t>(n_tiles) * 8 * kRows * sizeof(uint4));
  p.chk = static_cast<float*>(workspace);
  p.gx = g_q0;
  p.g_jbar = g_jbar;
  p.NR = n_rays;
  p.S = s_nerf;
  p.G = s_nerf <= kRows ? kRows / s_nerf : 1;
  p.T = s_nerf <= kRows ? 1 : (s_nerf + kRows - 1) / kRows;
  const int NG = (n_rays + p.G - 1) / p.G;
  if (NG * p.T != n_tiles) NJF_FAIL("njf_xf_backward: %d hand-over tiles do not match %d rays x %d samples", n_tiles, n_rays, s_nerf);
  if (bwd_smem(xfb_layer_fwd) || bwd_smem(xfb_head_bwd) || bwd_smem(xfb_layer_bwd)) return 1;
  const int grid = p.n_units < bwd_grid() ? p.n_units : bwd_grid();
  const size_t smem = sizeof(BwdSmem);
  for (int l = 0; l < 3; ++l) {
    p.layer = l;
    xfb_layer_fwd<<<grid, kBWarps * 32, smem, stream>>>(p);
    njf::count_launch();
  }
  xfb_head_bwd<<<grid, kBWarps * 32, smem, stream>>>(p);
  njf::count_launch();
  for (int l = 2; l >= 0; --l) {
    p.layer = l;
    xfb_layer_bwd<<<grid, kBWarps * 32, smem, stream>>>(p);
    njf::count_launch();
  }
  NJF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int njf_query_backward(const NjfCameras* cams, const NjfRenderArgs* a, const float* final_bins, int bins_stride,
                                  const float* g_q0, int n_tiles, float* g_wq_enc, float* g_bq, float* g_map, void* stream_) {
  if (!cams || !a || !final_bins || !g_q0 || !g_wq_enc || !g_bq || !g_map) NJF_FAIL("njf_query_backward: null argument");
  if (!a->origins || !a->dirs || !a->z_near || !a->z_far || !cams->ctxt_w2c || !cams->ctxt_k)
    NJF_FAIL("njf_query_backward: rays / cameras required");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  QueryBwdParams q{};
  PassGeom& g = q.g;
  const int S = a->s_nerf;
  g.NR = a->B * a->R;
  g.R = a->R;
  g.S = S;
  g.G = S <= kRows ? kRows / S : 1;
  g.T = S <= kRows ? 1 : (S + kRows - 1) / kRows;
  g.NG = (g.NR + g.G - 1) / g.G;
  if (g.NG * g.T != n_tiles) NJF_FAIL("njf_query_backward: %d tiles do not match the render shape", n_tiles);
  g.origins = a->origins;
  g.dirs = a->dirs;
  g.z_near = a->z_near;
  g.z_far = a->z_far;
  g.bins = final_bins;
  g.bins_stride = bins_stride;
  g.ctxt_w2c = cams->ctxt_w2c;
  g.ctxt_k = cams->ctxt_k;
  g.Hf = a->Hf;
  g.Wf = a->Wf;
  q.n_tiles = n_tiles;
  q.gq0 = g_q0;
  q.g_wq_enc = g_wq_enc;
  q.g_bq = g_bq;
  q.g_map = g_map;
  if (bwd_smem(query_bwd_kernel)) return 1;
  const int units = 2 * n_tiles;
  const int grid = units < bwd_grid() ? units : bwd_grid();
  query_bwd_kernel<<<grid, kBWarps * 32, sizeof(BwdSmem), stream>>>(q);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int njf_flow_backward(const float* g_flow, const float* g_pw_in, const float* jbar, const float* p,
                                 const float* action, const float* trgt_w2c, const float* trgt_k_px, int n_rays,
                                 int rays_per_view, int action_dim, float* g_jbar, float* g_action, void* stream_) {
  if (!g_flow || !jbar || !p || !action || !trgt_w2c || !trgt_k_px) NJF_FAIL("njf_flow_backward: null argument");
  if (n_rays < 1 || rays_per_view < 1 || action_dim < 1) NJF_FAIL("njf_flow_backward: bad sizes");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int B = (n_rays + rays_per_view - 1) / rays_per_view;
  if (g_action) NJF_CUDA(cudaMemsetAsync(g_action, 0, static_cast<size_t>(B) * action_dim * sizeof(float), stream));
  FlowBwdParams q{n_rays, rays_per_view, action_dim, g_flow, g_pw_in, jbar, p, action, trgt_w2c, trgt_k_px, g_jbar, g_action};
  flow_bwd_kernel<<<(n_rays + 255) / 256, 256, 0, stream>>>(q);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}
