// xf_kernel: the cross-attention Jacobian head (action_decoder_jacobian.py:418-446, transformer.py:14-135)
// as its own persistent kernel.
//
// field_kernel leaves, per 128-row tile, the 64-wide query embedding  q0 = Linear(575->64)([enc63 | feat512])
// (fp32, `qs`) and the transmittance weight of every sample (`wts`).  Everything after that is row-local
// arithmetic on a 64-wide stream, so this kernel needs no geometry, no gathers and only 125 KB of weights:
//   * ALL layer images (3 x {M1, M2, W1, W2} + jacobian_head, key/value projections folded, LayerNorm affine
//     folded into the following linear map, softmax log2(e) folded into M1) stay RESIDENT in shared memory --
//     no weight ring, so tile slots never wait on each other;
//   * FOUR tile slots per SM (the trunk kernels fit two): slot = 128 rows, ONE thread per row (TMEM lane),
//     16 KB A tile, 128 TMEM columns (x: fp32 residual stream [0,64), acc [64,128)); residual adds are the
//     accumulate flag of tcgen05.mma onto x, biases ride in SW32 bias blocks (kStepBias);
//   * no issuer warps: after a layer's A tile is written the slot's 128 threads meet at a named barrier and its
//     first thread issues the MMAs (same ~650-cycle round trip as a dedicated issuer thread, tools/ubench_tmem.cu,
//     but 512 instead of 640 threads: 128 registers per thread and four polling warps fewer); the 13 dependent
//     round trips of a tile overlap across the four slots.
// Then J-bar = sum_s w_s J_s per ray (model.py:281-286), optionally the per-sample Jacobians.
#include "field.h"
#include "njf_internal.h"
#include <atomic>

namespace njf {

constexpr int kXfSlots = 4;
constexpr int kXfEpiThreads = kXfSlots * kRows;           // 512
constexpr int kXfThreads = kXfEpiThreads;  // no issuer warps: the slot's first thread issues after a named barrier
constexpr uint32_t kXfATile = kRows * 128;                 // 128 rows x 64 fp16
struct XfSmem {
  static constexpr uint32_t kW = 0;
  static constexpr uint32_t kA = kXfMaxBlob;                      // 1024-aligned
  static constexpr uint32_t kOnes = kA + kXfSlots * kXfATile;
  static constexpr uint32_t kBars = kOnes + kBiasBlkBytes;
  static constexpr uint32_t kPart = kBars + 256;
  static constexpr uint32_t kTotal = kPart + kXfSlots * 4 * 32 * 4 + 1024;
};
static_assert(kXfMaxBlob % 1024 == 0, "A tiles must stay 1024 B aligned");
static_assert(XfSmem::kTotal <= 232448, "shared memory budget");
struct XfBars {
  uint64_t acc_ready[kXfSlots];
  uint64_t w_full;
  uint32_t tmem_base;
};

// -DNJF_PROFILE: per-phase SM-cycle attribution (lane 0 of every row warp), read by njf_xf_prof_read
enum XfPhase { kXLoad = 0, kXWait, kXLn, kXSoftmax, kXGelu, kXTail, kXCount };
#ifdef NJF_PROFILE
__device__ unsigned long long g_xf_prof[8];
#define XPROF(e, ph)                                      \
  do {                                                    \
    const long long now_ = clock64();                     \
    (e).prof[ph] += static_cast<unsigned>(now_ - (e).pt); \
    (e).pt = now_;                                        \
  } while (0)
#else
#define XPROF(e, ph) do { } while (0)
#endif

struct XfEpi {
#ifdef NJF_PROFILE
  long long pt;
  unsigned prof[kXCount];
#endif
  uint8_t* a_row;   // this row of the slot's A tile
  uint32_t sw;      // row & 7 (chunk swizzle)
  uint64_t* acc_ready;
  uint32_t tm;      // TMEM address: this lane, column 0 of the slot
  uint32_t par;
  // MMA issue (used by the slot's first thread only)
  const MmaStep* steps;
  uint32_t a0, w0, d0;   // smem address of the slot's A tile / of the resident blob, TMEM column 0 of the slot
  uint64_t ones;
  uint32_t bar_id;
  bool issuer;
};
__device__ __forceinline__ void xf_store_a(const XfEpi& e, const uint32_t (&pk)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(e.a_row + ((c ^ e.sw) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}
// A tile of layer `s` written: all 128 threads of the slot meet, the first one issues the layer's MMAs
// (4 x K=16 + the bias block) and commits to the slot's accumulator barrier
__device__ __forceinline__ void xf_publish(const XfEpi& e, int s) {
  fence_proxy_async_smem();
  tc_fence_before();
  named_bar_sync(e.bar_id, kRows);
  if (e.issuer) {
    tc_fence_after();
    const MmaStep st = e.steps[s];
    const uint32_t idesc = make_idesc_f16(st.n);
    const uint32_t d = e.d0 + st.d_col;
    const uint32_t w = e.w0 + st.w_off;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_f16(d, make_sw128_desc(e.a0 + k * 32), make_sw128_desc(w + k * 32), idesc, (st.acc || k > 0) ? 1u : 0u);
    umma_f16(d, e.ones, make_sw32_desc(w + static_cast<uint32_t>(st.n) * 128u), idesc, 1u);
    umma_commit(e.acc_ready);
  }
}
__device__ __forceinline__ void xf_wait(XfEpi& e, int ph) {
  XPROF(e, ph);
  mbar_wait(e.acc_ready, e.par);
  e.par ^= 1u;
  tc_fence_after();
  XPROF(e, kXWait);
}
__device__ __forceinline__ void xf_ld64(const XfEpi& e, uint32_t col, float (&v)[64]) {
  uint32_t r0[32], r1[32];
  tmem_ld32(e.tm + col, r0);
  tmem_ld32(e.tm + col + 32, r1);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    v[j] = __uint_as_float(r0[j]);
    v[32 + j] = __uint_as_float(r1[j]);
  }
}
// LayerNorm statistics over the row's 64 values (two-pass, biased variance, eps 1e-5 like nn.LayerNorm);
// the affine part lives in the next layer's weights -> fp16 -> A tile
__device__ __forceinline__ void xf_ln_to_a(const XfEpi& e, float (&x)[64]) {
  float2 s[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) s[i] = make_float2(x[2 * i], x[2 * i + 1]);
#pragma unroll
  for (int j = 4; j < 32; ++j) s[j & 3] = fadd2(s[j & 3], make_float2(x[2 * j], x[2 * j + 1]));
  const float2 st = fadd2(fadd2(s[0], s[1]), fadd2(s[2], s[3]));
  const float mean = (st.x + st.y) * (1.f / 64.f);
  const float2 nm = make_float2(-mean, -mean);
  float2 q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float2 d = fadd2(make_float2(x[2 * j], x[2 * j + 1]), nm);
    x[2 * j] = d.x;
    x[2 * j + 1] = d.y;
    q[j & 3] = ffma2(d, d, q[j & 3]);
  }
  const float2 qt = fadd2(fadd2(q[0], q[1]), fadd2(q[2], q[3]));
  const float rstd = rsqrtf((qt.x + qt.y) * (1.f / 64.f) + 1e-5f);
  const float2 rs2 = make_float2(rstd, rstd), zero = make_float2(0.f, 0.f);
  uint32_t pk[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float2 y = ffma2(make_float2(x[2 * j], x[2 * j + 1]), rs2, zero);
    pk[j] = pack_f16x2(y.x, y.y);
  }
  xf_store_a(e, pk);
}
// softmax over the A real keys of each of the 8 heads (heads padded to 8 columns; the logits arrive
// pre-multiplied by log2(e), so exp is a bare ex2)
template <int A>
__device__ __forceinline__ void xf_softmax_to_a(const XfEpi& e, float (&lg)[64]) {
  uint32_t pk[32];
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    float m = lg[h * 8];
#pragma unroll
    for (int a = 1; a < A; ++a) m = fmaxf(m, lg[h * 8 + a]);
    const float2 nm = make_float2(-m, -m);
    float2 ev[4];
#pragma unroll
    for (int a2 = 0; a2 < 4; ++a2) {
      const float2 d = fadd2(make_float2(lg[h * 8 + 2 * a2], lg[h * 8 + 2 * a2 + 1]), nm);
      ev[a2].x = (2 * a2 < A) ? ex2_approx(d.x) : 0.f;
      ev[a2].y = (2 * a2 + 1 < A) ? ex2_approx(d.y) : 0.f;
    }
    const float2 s2 = fadd2(fadd2(ev[0], ev[1]), fadd2(ev[2], ev[3]));
    const float inv = __fdividef(1.f, s2.x + s2.y);
    const float2 inv2 = make_float2(inv, inv), zero = make_float2(0.f, 0.f);
#pragma unroll
    for (int a2 = 0; a2 < 4; ++a2) {
      const float2 pr = ffma2(ev[a2], inv2, zero);
      pk[h * 4 + a2] = pack_f16x2(pr.x, pr.y);
    }
  }
  xf_store_a(e, pk);
}
// exact-erf GELU (nn.GELU default) as  relu(v) - |v|/2 * erfc(|v|/sqrt2),  erfc(a/sqrt2) = 2^-Q(a) with a
// degree-6 fit of Q on [0,6] (|gelu error| < 3e-7 in fp32, tests/test_host_cpu.py::test_gelu_fit); one MUFU
__device__ __forceinline__ float2 xf_gelu2(float2 v) {
  // GELU_Q: Q(a) = -log2(erfc(a / sqrt(2))) ~ sum_k c_k a^k on [0, 6] (tools/fit_gelu.py)
  constexpr float c0 = -7.392460702e-06f, c1 = 1.151206732e+00f, c2 = 4.587528110e-01f, c3 = 5.344062671e-02f,
                  c4 = -8.102188818e-03f, c5 = 7.767766947e-04f, c6 = -3.408475459e-05f;
  const float2 a = make_float2(fminf(fabsf(v.x), 6.f), fminf(fabsf(v.y), 6.f));
  float2 q = ffma2(a, make_float2(c6, c6), make_float2(c5, c5));
  q = ffma2(q, a, make_float2(c4, c4));
  q = ffma2(q, a, make_float2(c3, c3));
  q = ffma2(q, a, make_float2(c2, c2));
  q = ffma2(q, a, make_float2(c1, c1));
  q = ffma2(q, a, make_float2(c0, c0));
  const float2 h = make_float2(fabsf(v.x) * ex2_approx(-q.x), fabsf(v.y) * ex2_approx(-q.y));
  return ffma2(h, make_float2(-0.5f, -0.5f), make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f)));
}

__global__ void __launch_bounds__(kXfThreads, 1) xf_kernel(const __grid_constant__ XfParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  XfBars* bars = reinterpret_cast<XfBars*>(smem + XfSmem::kBars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kXfSlots; ++s) mbar_init(&bars->acc_ready[s], 1);
    mbar_init(&bars->w_full, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&bars->tmem_base, 512);
  if (threadIdx.x < kRows) {  // ones block: element (row, 0) = 1.0
    uint4* rowp = reinterpret_cast<uint4*>(smem + XfSmem::kOnes + threadIdx.x * 32);
    rowp[0] = make_uint4(0, 0, 0, 0);
    rowp[1] = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint16_t*>(smem + XfSmem::kOnes + sw32_offset(threadIdx.x, 0)) = 0x3C00;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  if (threadIdx.x == 0) {  // the whole head stays resident: one burst of bulk copies
    mbar_arrive_expect_tx(&bars->w_full, p.blob_bytes);
    for (uint32_t off = 0; off < p.blob_bytes; off += 32768u)
      bulk_g2s(smem + XfSmem::kW + off, p.blob + off, min(32768u, p.blob_bytes - off), &bars->w_full);
  }
  const int nitems = (p.NG + kXfSlots - 1) / kXfSlots;

  {
    // ------------------------------------------------------------------ one thread per row
    const int slot = warp >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    XfEpi e;
    uint8_t* a_tile = smem + XfSmem::kA + slot * kXfATile;
    e.a_row = a_tile + row * 128;
    e.sw = row & 7;
    e.acc_ready = &bars->acc_ready[slot];
    e.steps = p.prog.steps;
    e.a0 = smem_u32(a_tile);
    e.w0 = smem_u32(smem + XfSmem::kW);
    e.d0 = tmem + slot * 128;
    e.ones = make_sw32_desc(smem_u32(smem + XfSmem::kOnes));
    e.bar_id = 1 + slot;
    e.issuer = (threadIdx.x & (kRows - 1)) == 0;
    mbar_wait(&bars->w_full, 0);  // the resident weights have landed (once per CTA)
    e.tm = tmem + (static_cast<uint32_t>(q * 32) << 16) + slot * 128;
    e.par = 0;
#ifdef NJF_PROFILE
    e.pt = clock64();
    for (int i = 0; i < kXCount; ++i) e.prof[i] = 0;
#endif
    float* part = reinterpret_cast<float*>(smem + XfSmem::kPart) + slot * 4 * 32;
    const int A3 = 3 * p.A;
    const uint32_t bar_id = e.bar_id;
    for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
      const int lgroup = kXfSlots * it + slot;
      if (lgroup >= p.NG) continue;
      const int group = p.group0 + lgroup;
      float cs = 0.f;  // running column sum (warp 0 of the slot) across the tiles of a long ray
      for (int tile = 0; tile < p.T; ++tile) {
        // row -> (ray, sample): the same tiling as field_kernel (render.cuh row_setup)
        int lr, s;
        if (p.T == 1) {
          lr = row / p.S;
          s = row - lr * p.S;
          if (lr >= p.G) lr = -1;
        } else {
          lr = 0;
          s = tile * kRows + row;
          if (s >= p.S) lr = -1;
        }
        const int ray = (lr < 0) ? -1 : group * p.G + lr;
        const bool valid = ray >= 0 && ray < p.NR;
        const size_t tidx = static_cast<size_t>(lgroup) * p.T + tile;
        const uint4* qp = p.qs + tidx * 8 * kRows + row;
        const float w = __ldcs(p.wts + tidx * kRows + row);
        float x[64];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 v = __ldcs(qp + j * kRows);
          const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = __half22float2(h[t]);
            x[8 * j + 2 * t] = f.x;
            x[8 * j + 2 * t + 1] = f.y;
          }
        }
        {  // the residual stream lives in TMEM; accumulating MMAs (M2, W2) add onto it
          uint32_t r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(x[j]);
          tmem_st32(e.tm, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(x[32 + j]);
          tmem_st32(e.tm + 32, r);
        }
        xf_ln_to_a(e, x);
        tmem_st_wait();
        xf_publish(e, 0);  // -> M1 (layer 0)
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
          xf_wait(e, l == 0 ? kXLoad : kXLn);
          xf_ld64(e, 64, x);
          switch (p.A) {
            case 1: xf_softmax_to_a<1>(e, x); break;
            case 2: xf_softmax_to_a<2>(e, x); break;
            case 3: xf_softmax_to_a<3>(e, x); break;
            case 4: xf_softmax_to_a<4>(e, x); break;
            case 5: xf_softmax_to_a<5>(e, x); break;
            case 6: xf_softmax_to_a<6>(e, x); break;
            case 7: xf_softmax_to_a<7>(e, x); break;
            default: xf_softmax_to_a<8>(e, x); break;
          }
          xf_publish(e, 4 * l + 1);  // -> M2: x += attention . (V W_out) + b_out
          xf_wait(e, kXSoftmax);
          xf_ld64(e, 0, x);
          xf_ln_to_a(e, x);
          xf_publish(e, 4 * l + 2);  // -> W1
          xf_wait(e, kXLn);
          xf_ld64(e, 64, x);
          {
            uint32_t pk[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float2 gl = xf_gelu2(make_float2(x[2 * j], x[2 * j + 1]));
              pk[j] = pack_f16x2(gl.x, gl.y);
            }
            xf_store_a(e, pk);
          }
          xf_publish(e, 4 * l + 3);  // -> W2: x += W2 . gelu + b2
          xf_wait(e, kXGelu);
          xf_ld64(e, 0, x);
          if (l < 2) {
            xf_ln_to_a(e, x);
          } else {
            uint32_t pk[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) pk[j] = pack_f16x2(x[2 * j], x[2 * j + 1]);
            xf_store_a(e, pk);
          }
          xf_publish(e, 4 * l + 4);  // -> M1 of the next layer / jacobian_head Linear(64, 3A)
        }
        xf_wait(e, kXLn);
        float J[32];
        {
          uint32_t r[32];
          tmem_ld32(e.tm + 64, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) J[j] = __uint_as_float(r[j]);
        }
        if (valid && p.jac_out) {
          float* jp = p.jac_out + (static_cast<size_t>(ray) * p.S + s) * A3;
#pragma unroll
          for (int j = 0; j < 24; ++j)
            if (j < A3) jp[j] = J[j];
        }
        if (p.jbar) {
          // ---- J-bar = sum_s w_s J_s: stage w*J in the (now idle) A tile, column sums by the slot's 4 warps
          const float ww = valid ? w : 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 o;
            o.x = valid ? ww * J[4 * c] : 0.f;
            o.y = valid ? ww * J[4 * c + 1] : 0.f;
            o.z = valid ? ww * J[4 * c + 2] : 0.f;
            o.w = valid ? ww * J[4 * c + 3] : 0.f;
            *reinterpret_cast<float4*>(e.a_row + ((c ^ e.sw) << 4)) = o;
          }
          named_bar_sync(bar_id, kRows);
          auto colsum = [&](int r0, int n) {
            float s0 = 0.f, s1 = 0.f;
            int r = r0;
            for (; r + 1 < r0 + n; r += 2) {
              s0 += *reinterpret_cast<const float*>(a_tile + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
              s1 += *reinterpret_cast<const float*>(a_tile + (r + 1) * 128 + (((lane >> 2) ^ ((r + 1) & 7)) << 4) + (lane & 3) * 4);
            }
            if (r < r0 + n)
              s0 += *reinterpret_cast<const float*>(a_tile + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
            return s0 + s1;
          };
          if (p.T > 1 || (p.S & 31) == 0) {
            part[q * 32 + lane] = colsum(q * 32, 32);
            named_bar_sync(bar_id, kRows);
            if (p.T > 1) {
              if (q == 0) {
                cs += (part[lane] + part[32 + lane]) + (part[64 + lane] + part[96 + lane]);
                if (tile == p.T - 1 && group < p.NR && lane < A3) p.jbar[static_cast<size_t>(group) * A3 + lane] = cs;
              }
            } else if (q < p.G) {
              const int wpr = p.S >> 5;  // warps per ray
              float sum = 0.f;
              for (int k = 0; k < wpr; ++k) sum += part[(q * wpr + k) * 32 + lane];
              const int rr = group * p.G + q;
              if (rr < p.NR && lane < A3) p.jbar[static_cast<size_t>(rr) * A3 + lane] = sum;
            }
          } else {
            for (int k = q; k < p.G; k += 4) {
              const int rr = group * p.G + k;
              if (rr >= p.NR) break;
              const float sum = colsum(k * p.S, p.S);
              if (lane < A3) p.jbar[static_cast<size_t>(rr) * A3 + lane] = sum;
            }
          }
          named_bar_sync(bar_id, kRows);  // staging reads done before the next tile's A writes
        }
        XPROF(e, kXTail);
      }
    }
#ifdef NJF_PROFILE
    if (lane == 0)
      for (int i = 0; i < kXCount; ++i) atomicAdd(&g_xf_prof[i], static_cast<unsigned long long>(e.prof[i]));
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace njf

using namespace njf;

#ifdef NJF_PROFILE
extern "C" int njf_xf_prof_read(unsigned long long* out8, int reset) {
  NJF_CUDA(cudaDeviceSynchronize());
  NJF_CUDA(cudaMemcpyFromSymbol(out8, g_xf_prof, 8 * sizeof(unsigned long long)));
  if (reset) {
    unsigned long long z[8] = {};
    NJF_CUDA(cudaMemcpyToSymbol(g_xf_prof, z, sizeof(z)));
  }
  return 0;
}
#endif

int njf_xf_launch(const NjfField* f, const XfParams& params, cudaStream_t stream) {
  int dev = 0;
  cudaGetDevice(&dev);
  static std::atomic<bool> attr[64];  // the opt-in applies to the current device only
  if (dev >= 0 && dev < 64 && !attr[dev].load(std::memory_order_acquire)) {
    NJF_CUDA(cudaFuncSetAttribute(xf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(XfSmem::kTotal)));
    attr[dev].store(true, std::memory_order_release);
  }
  XfParams p = params;
  p.prog = f->head_prog;
  p.blob = f->d_xf_blob;
  p.blob_bytes = f->xf_bytes;
  p.A = f->desc.action_dim;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int nitems = (p.NG + kXfSlots - 1) / kXfSlots;
  const int grid = nitems < sms ? nitems : sms;
  xf_kernel<<<grid, kXfThreads, XfSmem::kTotal, stream>>>(p);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}
