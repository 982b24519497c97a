// The render kernels of libnjf_b200.so (sm_100a):
//   proposal_kernel : spacing bins -> positions -> gather + posenc -> ResnetFC (tcgen05) -> delta * sigma
//   pdf_kernel      : transmittance weights + PDF resampling of one proposal level, one warp per ray
//   field_kernel    : final bins -> gather + posenc -> query embedding (cross-attention head: handed to
//                     xf_kernel, xf_head.cu) -> density trunk, colour head (and the MLP Jacobian trunk) ->
//                     transmittance weights -> alpha compositing
//   finish_kernel   : call-global depth clip, J.u flow, projection to the target camera
//   hoist_kernel    : fp32 SIMT variant of the hoist GEMM (A/B checks only; the product path is hoist_tc.cu)
// plus the standalone sampler / point-query kernels and the host launchers of the C ABI.
// See DESIGN.md for the data layout and rooflines.
#include "render.cuh"
#include "njf_internal.h"
#include <atomic>
#include <cstdlib>

namespace njf {

struct SlotScratch {
  float dd[kRows];       // delta * sigma of the current tile
  float cum[kRows];      // exclusive prefix sums (fp32) of dd
  float wrow[kRows];     // transmittance weight of each row (shared between the row's two threads)
  TapEntry taps[kRows];  // bilinear taps of the current tile
  float4 rgbp[kRows];    // colour-head partial dot products of the upper column half
  float part[8][64];     // per-warp partial column sums of the composite
};
constexpr uint32_t kScratchSlotBytes = (sizeof(SlotScratch) + 15) & ~15u;
constexpr uint32_t kSmemBytes = SmemMap::kScratch + kSlots * kScratchSlotBytes + 64 + 1024;
static_assert(kSmemBytes <= 232448, "shared memory budget");

__device__ __forceinline__ SlotScratch* slot_scratch(const CtaCtx& c, int slot) {
  return reinterpret_cast<SlotScratch*>(c.smem + SmemMap::kScratch + slot * kScratchSlotBytes);
}
__device__ __forceinline__ uint32_t* cta_minmax(const CtaCtx& c) {
  return reinterpret_cast<uint32_t*>(c.smem + SmemMap::kScratch + kSlots * kScratchSlotBytes);
}

// transmittance weights of this tile's rows (RaySamples.get_weights, ray_samplers.py:77-101): the h=0 thread
// of each row stores dd = delta * sigma; exclusive prefix sums are accumulated in double (like torch.cumsum on
// CPU) and rounded to fp32 per element; every thread of the row returns the row's weight.
//  * fast path (a warp's 32 rows belong to one ray: S a multiple of 32, or a long ray): ONE barrier; every warp
//    reduces the ray's rows before its own and scans its own 32 rows with shuffles (both warps of a row quarter
//    do this redundantly, so no second exchange is needed);
//  * general path (several ragged rays per tile): one warp per ray scans, results go through shared memory.
__device__ __forceinline__ float tile_weights(EpiCtx& e, SlotScratch* sc, const PassGeom& g, int tile,
                                              float dd, double& carry) {
  const int w8 = (threadIdx.x >> 5) & 7;
  const int lane = threadIdx.x & 31;
  PROF(e, kPOther);
  if (e.half == 0) sc->dd[e.row] = dd;
  slot_bar(e);
  float w;
  if (g.T > 1 || (g.S & 31) == 0) {
    const int row0 = 32 * e.q;                                    // first row of this warp
    const int ray0 = (g.T > 1) ? 0 : (row0 / g.S) * g.S;          // first row of the warp's ray inside the tile
    double before = 0.0;                                          // sum of the ray's rows before this warp's
    for (int r = ray0 + lane; r < row0; r += 32) before += static_cast<double>(sc->dd[r]);
    double tot = 0.0;                                             // long ray: whole-tile sum for the carry
    if (g.T > 1)
      for (int r = lane; r < kRows; r += 32) tot += static_cast<double>(sc->dd[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      before += __shfl_xor_sync(0xffffffffu, before, o);
      if (g.T > 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    }
    const double own = static_cast<double>(sc->dd[e.row]);
    double inc = own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const float cum = static_cast<float>(carry + before + (inc - own));
    carry += tot;
    w = (1.f - expf(-sc->dd[e.row])) * expf(-cum);
  } else {
    for (int lr = w8; lr < g.G; lr += 8) {
      double c0 = 0.0;
      excl_scan_warp(sc->dd + lr * g.S, g.S, c0, sc->cum + lr * g.S);
    }
    slot_bar(e);
    if (e.half == 0) {
      const float tr = expf(-sc->cum[e.row]);
      const float alpha = 1.f - expf(-dd);
      sc->wrow[e.row] = alpha * tr;
    }
    slot_bar(e);
    w = sc->wrow[e.row];
  }
  PROF(e, kPWeights);
  return w;
}

// ============================================================================= proposal pass
struct ProposalParams {
  Program prog;
  const uint8_t* blob;
  PassGeom g;
  int n_out;             // samples of the next level (PDF draws n_out+1 bin edges)
  const float* u;        // [n_out+1] shared or per ray
  int u_stride;
  float anneal;
  int sum_vec;
  float* bins_out;       // [NR][n_out+1]
  float* weights_out;    // [rays of this launch][S]: input of the PDF step, which runs as its own kernel afterwards
                         // (one warp per ray, fully parallel -- in here a single warp per tile would do it while
                         // seven wait); row 0 belongs to ray `w_ray0`
  int w_ray0;
  int32_t* inds_out;     // optional [NR][n_out+1]
  float* sigma_out;      // point-query mode (DensityDecoderMlp.get_density at explicit points): [NR] densities
};

__global__ void __launch_bounds__(kThreads, 1) proposal_kernel(const __grid_constant__ ProposalParams p) {
  extern __shared__ uint8_t smem_raw[];
  CtaCtx c = cta_setup(smem_raw);
  const PassGeom& g = p.g;
  const int warp = threadIdx.x >> 5;
  const bool one_slot = (g.debug & 4) != 0;  // NJF_DEBUG_SKIP bit 2 (measurement only): slot 1 idles
  const int nitems = one_slot ? g.NG : (g.NG + 1) / 2;
  int my_items = 0;
  for (int it = blockIdx.x; it < nitems; it += gridDim.x) ++my_items;

  if (warp == kLoaderWarp) {
    if ((threadIdx.x & 31) == 0) loader_role(c, p.prog, p.blob, my_items * g.T, g.debug);
  } else if (warp == kIssuerWarp) {
    if ((threadIdx.x & 31) == 0) {
      const int NG = g.NG, T = g.T, bx = blockIdx.x, gx = gridDim.x;
      issuer_role(c, p.prog, my_items * g.T, [=](int run) {
        const int it = bx + (run / T) * gx;
        return (!one_slot && 2 * it + 1 < NG) ? 2 : 1;
      });
    }
  } else {
    EpiCtx e = epi_ctx(c);
    SlotScratch* sc = slot_scratch(c, e.slot);
    const int lane = threadIdx.x & 31;
    // This slot's tiles form a software pipeline: while tile n runs blocks 2..4 (whose MMA wait windows carry no
    // gather of tile n), the warp sets up tile n+1 (bins -> position -> projection -> taps) and prefetches ITS part
    // of tile n+1's first hoisted segment into the staging buffer, which tile n stopped using after block 1.
    int n_my = 0;  // tiles of this slot in this CTA
    for (int it = blockIdx.x; it < nitems; it += gridDim.x)
      if ((one_slot ? it : 2 * it + e.slot) < g.NG && !(one_slot && e.slot)) n_my += g.T;
    auto tile_of = [&](int n, int& group, int& tile) {
      const int it = blockIdx.x + (n / g.T) * gridDim.x;
      group = g.group0 + (one_slot ? it : 2 * it + e.slot);
      tile = n - (n / g.T) * g.T;
    };
    RowState rs;
    if (n_my > 0) {
      int group, tile;
      tile_of(0, group, tile);
      PROF(e, kPOther);
      row_setup(g, group, tile, e.row, rs);
      PROF(e, kPPdf);
      if ((lane >> 4) == e.half) write_taps(sc->taps, e.row, rs, g.Hf, g.Wf, g.CH);
      PROF(e, kPHead);
      write_posenc(e, rs.cam, rs.ray >= 0, g.debug);
      PROF(e, kPColor);
      epi_publish(e);  // -> lin_in
      pair_bar(e);     // the row quarter's tap entries (16 written by each of its two warps) are complete
      PROF(e, kPSetup);
      gather_segment<128>(e, g, sc->taps, 0);
    }
    for (int n = 0; n < n_my; ++n) {
      const bool has_next = n + 1 < n_my;
      RowState nx;
      nx.ray = -1; nx.s = 0; nx.delta = 0.f; nx.cam[0] = nx.cam[1] = nx.cam[2] = 0.f;
      epi_wait_acc(e);  // lin_in
      trunk_blocks_epilogue(e, g, 0, sc->taps, [&](int k, int w) {
        if (!has_next) return;
        if (k == 2 && w == 1) {
          // every thread of the slot is past block 1 here (the fc_0(2) accumulator needed all 256 arrivals), so
          // nobody reads tile n's tap table any more
          int group, tile;
          tile_of(n + 1, group, tile);
          PROF(e, kPOther);
          row_setup(g, group, tile, e.row, nx);
          PROF(e, kPPdf);
          pair_bar(e);  // the partner warp has issued its last read of tile n's entries
          if ((lane >> 4) == e.half) write_taps(sc->taps, e.row, nx, g.Hf, g.Wf, g.CH);
          pair_bar(e);  // tile n+1's entries of this row quarter are complete
          PROF(e, kPHead);
        } else if (k == 3 && w == 0) {
          __syncwarp();
          gather_rows<128>(e, g, sc->taps, 0, 0, 8);
        } else if (k == 3 && w == 1) {
          gather_rows<128>(e, g, sc->taps, 0, 8, 16);
          __syncwarp();
        }
      });
      epi_wait_acc(e);  // lin_out accumulator ready
      float dd = 0.f, sigma_pt = 0.f;
      if (e.half == 0) {
        uint32_t r[16];
        tmem_ld16(e.tmem + 128, r);
        tmem_ld_wait();
        // density = trunc_exp(x - 1) (density_decoder.py:64-66, activations.py:32-35)
        const float sigma = expf(__fsub_rn(__uint_as_float(r[0]), 1.f));
        sigma_pt = sigma;
        dd = (rs.ray >= 0 && rs.delta > 0.f) ? __fmul_rn(rs.delta, sigma) : 0.f;
      }
      // delta*sigma goes out; the per-ray kernel that follows does the transmittance scan
      // (RaySamples.get_weights) and the PDF resampling with one warp per ray
      if (e.half == 0 && rs.ray >= 0) {
        if (p.sigma_out) p.sigma_out[rs.ray] = sigma_pt;
        else p.weights_out[static_cast<size_t>(rs.ray - p.w_ray0) * g.S + rs.s] = dd;
      }
      PROF(e, kPWeights);
      if (has_next) {
        rs = nx;
        tc_fence_before();  // this thread's TMEM reads of tile n are done before lin_in of tile n+1 overwrites x
        write_posenc(e, rs.cam, rs.ray >= 0, g.debug);
        PROF(e, kPColor);
        epi_publish(e);  // -> lin_in of tile n+1 (its first segment is already staged)
        PROF(e, kPSetup);
      }
    }
    prof_flush(e);
  }
  cta_teardown(c);
}

// ============================================================================= field pass
struct FieldParams {
  Program prog;
  const uint8_t* blob;
  ColorTab color;
  PassGeom g;
  int head_kind;   // NJF_HEAD_*
  int A;
  int sh_conv;     // NJF_SH_*
  // per-ray outputs
  float* rgb; float* depth; float* jbar; float* p;
  // per-sample outputs
  float* steps; float* weights; float* sigma; float* jac_out; float* positions; float* rgb_samples;
  float* geo_out;    // [NR*S][15] density-head geometry features (point queries)
  // transformer head: hand-over to xf_kernel (indexed by the launch-local tile number)
  uint4* qs;         // [tile][8 chunks][128 rows] x 8 fp16: query embedding
  float* wts;        // [tile][128] sample weights (0 for padding rows)
  uint32_t* minmax;  // [2] ordered-uint encoded min / max of steps
};

// ---- helpers on this thread's 32 of the 64 transformer / colour columns (columns 32h .. 32h+31)
__device__ __forceinline__ void ld_acc32(const EpiCtx& e, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld32(e.tmem + 128 + 32 * e.half, r);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
// Query embedding of the cross-attention head (action_decoder_jacobian.py:423-430):
//   q0 = W_q . [enc | xyz] + b_q (tensor core, q_enc step) + hoisted W_q[:, 63:] . feat
// Each of the row's two threads owns 32 of the 64 values and streams them, rounded to fp16 (they feed a
// LayerNorm and fp16 tensor-core operands next), to the `qs` hand-over that xf_kernel consumes
// ([tile][8 chunks of 8 values][128 rows] uint4: warp stores are 512 contiguous bytes; 128 B per sample).
__device__ __forceinline__ void store_query_stream(const EpiCtx& e, uint4* qs_tile) {
  float x[32];
  ld_acc32(e, x);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 q = *reinterpret_cast<const uint4*>(e.tz + tz_offset(e.row, 8 * e.half + j));  // gather_rows<64> layout
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = __half22float2(h[t]);
      x[8 * j + 2 * t] += f.x;
      x[8 * j + 2 * t + 1] += f.y;
    }
  }
  uint4* dst = qs_tile + (4 * e.half) * kRows + e.row;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    __stcs(dst + j * kRows, make_uint4(pack_f16x2(x[8 * j], x[8 * j + 1]), pack_f16x2(x[8 * j + 2], x[8 * j + 3]),
                                       pack_f16x2(x[8 * j + 4], x[8 * j + 5]), pack_f16x2(x[8 * j + 6], x[8 * j + 7])));
}

__global__ void __launch_bounds__(kThreads, 1) field_kernel(const __grid_constant__ FieldParams p) {
  extern __shared__ uint8_t smem_raw[];
  CtaCtx c = cta_setup(smem_raw);
  const PassGeom& g = p.g;
  const int warp = threadIdx.x >> 5;
  const int nitems = (g.NG + 1) / 2;
  int my_items = 0;
  for (int it = blockIdx.x; it < nitems; it += gridDim.x) ++my_items;
  uint32_t* mm = cta_minmax(c);
  if (threadIdx.x == 0) {
    mm[0] = 0xffffffffu;
    mm[1] = 0u;
  }
  __syncthreads();

  if (warp == kLoaderWarp) {
    if ((threadIdx.x & 31) == 0) loader_role(c, p.prog, p.blob, my_items * g.T, g.debug);
  } else if (warp == kIssuerWarp) {
    if ((threadIdx.x & 31) == 0) {
      const int NG = g.NG, T = g.T, bx = blockIdx.x, gx = gridDim.x;
      issuer_role(c, p.prog, my_items * g.T, [=](int run) {
        const int it = bx + (run / T) * gx;
        return (2 * it + 1 < NG) ? 2 : 1;
      });
    }
  } else {
    EpiCtx e = epi_ctx(c);
    SlotScratch* sc = slot_scratch(c, e.slot);
    const int w8 = warp & 7;
    const int lane = threadIdx.x & 31;
    const int A3 = 3 * p.A;
    const int nch = 8 + A3;  // composite channels: rgb3, t, 1, pos3, J(3A)
    // The slot's tiles form a software pipeline (like proposal_kernel): during blocks 2..4 of the LAST trunk of
    // tile n the warp sets up tile n+1 (row set-up, tap table) and prefetches its part of tile n+1's first hoisted
    // segment (the 64 query channels for the cross-attention head, segment 0 for the MLP head) into the staging
    // buffer.  The composite therefore stages its rows in the A tile (idle after the tile's last MMA), not in e.tz.
    const bool xf_head = p.head_kind == NJF_HEAD_TRANSFORMER;
    int n_my = 0;  // tiles of this slot in this CTA
    for (int it = blockIdx.x; it < nitems; it += gridDim.x)
      if (2 * it + e.slot < g.NG) n_my += g.T;
    auto tile_of = [&](int n, int& lgroup, int& tile) {
      const int it = blockIdx.x + (n / g.T) * gridDim.x;
      lgroup = 2 * it + e.slot;  // launch-local ray group
      tile = n - (n / g.T) * g.T;
    };
    RowState rs;
    if (n_my > 0) {
      int lgroup, tile;
      tile_of(0, lgroup, tile);
      PROF(e, kPOther);
      row_setup(g, g.group0 + lgroup, tile, e.row, rs);
      if ((lane >> 4) == e.half) write_taps(sc->taps, e.row, rs, g.Hf, g.Wf, g.CH);
      write_posenc(e, rs.cam, rs.ray >= 0, g.debug);
      epi_publish(e);  // -> lin_in (+ q_enc for the transformer head)
      pair_bar(e);     // the row quarter's tap entries (16 written by each of its two warps) are complete
      PROF(e, kPSetup);
      if (xf_head) gather_segment<64>(e, g, sc->taps, 384);
      else gather_segment<128>(e, g, sc->taps, 0);
    }
    double carry = 0.0;
    float cs0 = 0.f, cs1 = 0.f;  // column sums held by warp 0 of the slot across the tiles of a long ray
    for (int n = 0; n < n_my; ++n) {
      {
        int lgroup, tile;
        tile_of(n, lgroup, tile);
        const int group = g.group0 + lgroup;
        if (tile == 0) {
          carry = 0.0;
          cs0 = cs1 = 0.f;
        }
        const bool has_next = n + 1 < n_my;
        RowState nx;
        nx.ray = -1; nx.s = 0; nx.tmid = 0.f; nx.delta = 0.f;
        nx.pos[0] = nx.pos[1] = nx.pos[2] = 0.f;
        nx.cam[0] = nx.cam[1] = nx.cam[2] = 0.f;
        // set-up of tile n+1 in an MMA wait window of the last trunk's block 2: every thread of the slot is past
        // block 1 of that trunk then, so nobody reads tile n's tap table any more
        auto setup_next = [&]() {
          int lg2, t2;
          tile_of(n + 1, lg2, t2);
          PROF(e, kPOther);
          row_setup(g, g.group0 + lg2, t2, e.row, nx);
          pair_bar(e);  // the partner warp has issued its last read of tile n's entries
          if ((lane >> 4) == e.half) write_taps(sc->taps, e.row, nx, g.Hf, g.Wf, g.CH);
          pair_bar(e);  // tile n+1's entries of this row quarter are complete
          PROF(e, kPSetup);
        };
        const bool valid = rs.ray >= 0;
        float J[16];  // this thread's 16 of the 32 (padded) Jacobian columns: 16h .. 16h+15 (MLP head)
#pragma unroll
        for (int j = 0; j < 16; ++j) J[j] = 0.f;
        const size_t tidx = static_cast<size_t>(lgroup) * g.T + tile;
        if (xf_head) {
          // lin_in / q_enc window: the upper half of segment 0 goes into the staging chunks the query channels
          // (gathered by the prologue / the previous tile) do not occupy
          __syncwarp();
          gather_seg128_part(e, g, sc->taps, 0, 1);
          epi_wait_acc(e);  // lin_in + q_enc
          if (p.qs) store_query_stream(e, p.qs + tidx * 8 * kRows);
          PROF(e, kPHead);
          __syncwarp();     // the query channels are consumed: their chunks take the lower half of segment 0
          gather_seg128_part(e, g, sc->taps, 0, 0);
          __syncwarp();
          trunk_blocks_epilogue(e, g, 0, sc->taps, [&](int k, int w) {
            if (!has_next) return;
            if (k == 2 && w == 1) {
              setup_next();
            } else if (k == 3 && w == 0) {
              __syncwarp();
              gather_rows<64>(e, g, sc->taps, 384, 0, GatherShape<64>::kGroups);
              __syncwarp();
            }
          });
        } else {
          epi_wait_acc(e);  // lin_in; segment 0 was gathered by the prologue / the previous tile
          trunk_blocks_epilogue(e, g, 0, sc->taps);  // (the tap table is still needed by the Jacobian trunk)
        }
        // lin_out: 15 geometry features + density pre-activation (action_decoder_jacobian.py:106-112); colour head
        // input [geo15 | sh16 | 0...] (:208).  The h=0 thread handles the geometry half (columns 0..15: geo, sh0),
        // the h=1 thread the view direction (columns 16..31: sh1..15, 0) -- its loads are issued before the wait.
        PROF(e, kPOther);
        float sigma = 0.f;
        float dv[3] = {0.f, 0.f, 1.f};
        if (e.half == 1 && g.dirs) {
          const float* dp = g.dirs + static_cast<size_t>(valid ? rs.ray : 0) * 3;
          dv[0] = __ldg(dp); dv[1] = __ldg(dp + 1); dv[2] = __ldg(dp + 2);
        }
        epi_wait_acc(e);  // lin_out accumulator ready
        if (e.half == 0) {
          float geo[16];
          uint32_t r[16];
          tmem_ld16(e.tmem + 128, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) geo[j] = __uint_as_float(r[j]);
          sigma = expf(__fsub_rn(geo[15], 1.f));
          if (p.geo_out && valid) {
            float* gp = p.geo_out + (static_cast<size_t>(rs.ray) * g.S + rs.s) * 15;
#pragma unroll
            for (int j = 0; j < 15; ++j) gp[j] = geo[j];
          }
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 7; ++j) pk[j] = pack_f16x2(geo[2 * j], geo[2 * j + 1]);
          pk[7] = pack_f16x2(geo[14], 0.28209479177387814f);  // SH band 0 is a constant
          a_store16(e, 0, pk);
        } else {
          float sh[16];
          sh16(dv[0], dv[1], dv[2], p.sh_conv, sh);
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 7; ++j) pk[j] = pack_f16x2(sh[1 + 2 * j], sh[2 + 2 * j]);
          pk[7] = pack_f16x2(sh[15], 0.f);
          a_store16(e, 16, pk);
          a_zero32(e, 32);
        }
        epi_publish(e);  // -> color1
        epi_wait_acc(e);
        epi_relu_to_a(e, 128 + 32 * e.half, 32 * e.half);
        epi_publish(e);  // -> color2
        epi_wait_acc(e);
        float rgb[3] = {0.f, 0.f, 0.f};
        {
          float h2[32];
          ld_acc32(e, h2);
          float part[3];
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const float* w3 = p.color.w3 + ch * 64 + 32 * e.half;
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) a = fmaf(w3[j], fmaxf(h2[j], 0.f), a);
            part[ch] = a;
          }
          if (e.half == 1) sc->rgbp[e.row] = make_float4(part[0], part[1], part[2], 0.f);
          pair_bar(e);
          if (e.half == 0) {
            const float4 o = sc->rgbp[e.row];
            rgb[0] = 1.f / (1.f + expf(-(part[0] + o.x + p.color.b3[0])));
            rgb[1] = 1.f / (1.f + expf(-(part[1] + o.y + p.color.b3[1])));
            rgb[2] = 1.f / (1.f + expf(-(part[2] + o.z + p.color.b3[2])));
          }
        }
        PROF(e, kPColor);
        if (p.head_kind == NJF_HEAD_MLP) {
          // second ResnetFC on the same gathered point (action_decoder_jacobian.py:324-337)
          write_posenc(e, rs.cam, valid, g.debug);
          epi_publish(e);  // -> lin_in (jacobian head)
          gather_segment<128>(e, g, sc->taps, 384);
          epi_wait_acc(e);
          trunk_blocks_epilogue(e, g, 384, sc->taps, [&](int k, int w) {
            if (!has_next) return;
            if (k == 2 && w == 1) {
              setup_next();
            } else if (k == 3 && w == 0) {
              __syncwarp();
              gather_rows<128>(e, g, sc->taps, 0, 0, 8);
            } else if (k == 3 && w == 1) {
              gather_rows<128>(e, g, sc->taps, 0, 8, 16);
              __syncwarp();
            }
          });
          epi_wait_acc(e);
          uint32_t r[16];
          tmem_ld16(e.tmem + 128 + 16 * e.half, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) J[j] = __uint_as_float(r[j]);
        }
        // ---- weights + compositing (model.py:351-367, 384-394)
        const float dd = (valid && rs.delta > 0.f) ? __fmul_rn(rs.delta, sigma) : 0.f;
        const float w = tile_weights(e, sc, g, tile, dd, carry);
        const float ww = valid ? w : 0.f;
        // composite staging: [128 rows][64 fp32] in the slot's A tile (32 KB), 16 B chunks XOR-swizzled by row; the
        // tile's last MMA has completed, and the staging buffer e.tz may already hold tile n+1's prefetched segment
        uint8_t* rowp = e.a_tile + e.row * 256;
        const bool stage_j = p.jbar != nullptr;  // MLP head: J composited here; transformer: by xf_kernel
        if (e.half == 0) {
          if (p.wts) __stcs(p.wts + tidx * kRows + e.row, ww);
          if (valid) {
            const size_t si = static_cast<size_t>(rs.ray) * g.S + rs.s;
            if (p.steps) p.steps[si] = rs.tmid;
            if (p.weights) p.weights[si] = w;
            if (p.sigma) p.sigma[si] = sigma;
            if (p.positions) {
              p.positions[si * 3 + 0] = rs.pos[0];
              p.positions[si * 3 + 1] = rs.pos[1];
              p.positions[si * 3 + 2] = rs.pos[2];
            }
            if (p.rgb_samples) {
              p.rgb_samples[si * 3 + 0] = rgb[0];
              p.rgb_samples[si * 3 + 1] = rgb[1];
              p.rgb_samples[si * 3 + 2] = rgb[2];
            }
          }
          // stage w * [rgb, t, 1, pos, J0..15] : composite columns 0..23
          float v[24];
          v[0] = rgb[0]; v[1] = rgb[1]; v[2] = rgb[2]; v[3] = rs.tmid; v[4] = 1.f;
          v[5] = rs.pos[0]; v[6] = rs.pos[1]; v[7] = rs.pos[2];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[8 + j] = (j < A3) ? J[j] : 0.f;
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            if (q >= 2 && !stage_j) break;
            float4 o;
            o.x = valid ? ww * v[4 * q + 0] : 0.f;
            o.y = valid ? ww * v[4 * q + 1] : 0.f;
            o.z = valid ? ww * v[4 * q + 2] : 0.f;
            o.w = valid ? ww * v[4 * q + 3] : 0.f;
            *reinterpret_cast<float4*>(rowp + ((q ^ (e.row & 7)) << 4)) = o;
          }
          // min / max of steps over valid samples (render_depth's clip range, model.py:277)
          float tmn = valid ? rs.tmid : 3.0e38f, tmx = valid ? rs.tmid : -3.0e38f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            tmn = fminf(tmn, __shfl_xor_sync(0xffffffffu, tmn, o));
            tmx = fmaxf(tmx, __shfl_xor_sync(0xffffffffu, tmx, o));
          }
          if (lane == 0 && tmn <= tmx) {
            atomicMin(&mm[0], f2ord(tmn));
            atomicMax(&mm[1], f2ord(tmx));
          }
        } else if (stage_j) {
          // J16..31 : composite columns 24..39
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 o;
            o.x = (valid && 16 + 4 * q + 0 < A3) ? ww * J[4 * q + 0] : 0.f;
            o.y = (valid && 16 + 4 * q + 1 < A3) ? ww * J[4 * q + 1] : 0.f;
            o.z = (valid && 16 + 4 * q + 2 < A3) ? ww * J[4 * q + 2] : 0.f;
            o.w = (valid && 16 + 4 * q + 3 < A3) ? ww * J[4 * q + 3] : 0.f;
            *reinterpret_cast<float4*>(rowp + (((6 + q) ^ (e.row & 7)) << 4)) = o;
          }
        }
        if (valid && p.jac_out) {
          const size_t si = static_cast<size_t>(rs.ray) * g.S + rs.s;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (16 * e.half + j < A3) p.jac_out[si * A3 + 16 * e.half + j] = J[j];
        }
        slot_bar(e);
        auto colsum = [&](int r0, int n, float& s0, float& s1) {
          for (int r = r0; r < r0 + n; ++r) {
            const uint8_t* rp = e.a_tile + r * 256;
            if (stage_j || lane < 8)
              s0 += *reinterpret_cast<const float*>(rp + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
            if (stage_j && nch > 32)
              s1 += *reinterpret_cast<const float*>(rp + (((8 + (lane >> 2)) ^ (r & 7)) << 4) + (lane & 3) * 4);
          }
        };
        auto emit = [&](int ray, float s0, float s1) {
          // channel of this lane: s0 -> lane, s1 -> 32 + lane
          const float num = __shfl_sync(0xffffffffu, s0, 3), den = __shfl_sync(0xffffffffu, s0, 4);
          if (lane < 3 && p.rgb) p.rgb[static_cast<size_t>(ray) * 3 + lane] = s0;
          if (lane == 3 && p.depth) p.depth[ray] = num / (den + 1e-10f);
          if (lane >= 5 && lane < 8 && p.p) p.p[static_cast<size_t>(ray) * 3 + (lane - 5)] = s0;
          if (p.jbar) {
            if (lane >= 8 && lane - 8 < A3) p.jbar[static_cast<size_t>(ray) * A3 + (lane - 8)] = s0;
            if (24 + lane < A3) p.jbar[static_cast<size_t>(ray) * A3 + 24 + lane] = s1;
          }
        };
        if (g.T > 1 || (g.S & 31) == 0) {
          // fast path: every warp sums 16 rows (all of one ray), then one warp per ray adds the partials
          const int wi = 4 * e.half + e.q;
          float s0 = 0.f, s1 = 0.f;
          colsum(16 * wi, 16, s0, s1);
          sc->part[wi][lane] = s0;
          sc->part[wi][32 + lane] = s1;
          slot_bar(e);
          if (g.T > 1) {
            if (wi == 0) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                cs0 += sc->part[k][lane];
                cs1 += sc->part[k][32 + lane];
              }
              if (tile == g.T - 1 && group < g.NR) emit(group, cs0, cs1);
            }
          } else if (wi < g.G) {
            const int wpr = g.S >> 4;  // warps per ray
            float t0 = 0.f, t1 = 0.f;
            for (int k = 0; k < wpr; ++k) {
              t0 += sc->part[wi * wpr + k][lane];
              t1 += sc->part[wi * wpr + k][32 + lane];
            }
            const int ray = group * g.G + wi;
            if (ray < g.NR) emit(ray, t0, t1);
          }
        } else {
          for (int lr = w8; lr < g.G; lr += 8) {
            const int ray = group * g.G + lr;
            if (ray >= g.NR) break;
            float s0 = 0.f, s1 = 0.f;
            colsum(lr * g.S, g.S, s0, s1);
            emit(ray, s0, s1);
          }
          slot_bar(e);
        }
        PROF(e, kPComposite);
        if (has_next) {   // tile n+1: its rows are set up, its first segment is staged; the A tile is free again
          rs = nx;
          write_posenc(e, rs.cam, rs.ray >= 0, g.debug);
          epi_publish(e);  // -> lin_in (+ q_enc) of tile n+1
          PROF(e, kPSetup);
        }
      }
    }
    prof_flush(e);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && p.minmax) {
    if (mm[0] != 0xffffffffu) atomicMin(&p.minmax[0], mm[0]);
    if (mm[1] != 0u) atomicMax(&p.minmax[1], mm[1]);
  }
  cta_teardown(c);
}

// ============================================================================= small kernels
__global__ void init_minmax_kernel(uint32_t* mm) {
  mm[0] = 0xffffffffu;
  mm[1] = 0u;
}
// ordered-uint -> plain fp32 in place, so that ranks can all-reduce (min, max) between the passes
__global__ void decode_minmax_kernel(uint32_t* mm) {
  const float lo = ord2f(mm[0]), hi = ord2f(mm[1]);
  reinterpret_cast<float*>(mm)[0] = lo;
  reinterpret_cast<float*>(mm)[1] = hi;
}

// depth clip + optical flow (model.py:271-279, 288-314; geometry.py:206-215)
struct FinishParams {
  int NR, R, A;
  int ray0;
  const float* action;    // [B][A]
  const float* trgt_w2c;  // [B][16]
  const float* trgt_k;    // [B][9] pixel units
  const float* minmax;   // [2] call-global (min, max) of steps
  float* depth;
  const float* jbar;
  const float* p;
  float* pw;
  float* flow;
  const float* rgb;
  float* packed;         // optional [NR][12 + 3A]: rgb3 | depth1 | flow2 | jbar3A | p3 | pw3
};
__device__ __forceinline__ void project_uv(const float* W, const float* K, float x, float y, float z, float& u,
                                           float& v) {
  float c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) c[i] = fmaf(W[4 * i + 2], z, fmaf(W[4 * i + 1], y, fmaf(W[4 * i], x, W[4 * i + 3])));
  const float a = fmaf(K[2], c[2], fmaf(K[1], c[1], K[0] * c[0]));
  const float b = fmaf(K[5], c[2], fmaf(K[4], c[1], K[3] * c[0]));
  const float w = fmaf(K[8], c[2], fmaf(K[7], c[1], K[6] * c[0]));
  const float zd = __fadd_rn(w, 1e-9f);
  u = __fdiv_rn(a, zd);
  v = __fdiv_rn(b, zd);
}
__global__ void finish_kernel(const FinishParams q) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= q.NR) return;
  const int b = (ray + q.ray0) / q.R;
  const int A3 = 3 * q.A;
  float* pk = q.packed ? q.packed + static_cast<size_t>(ray) * (12 + A3) : nullptr;
  float dep = 0.f;
  if (q.depth) {
    dep = q.depth[ray];
    if (q.minmax) {
      dep = fminf(fmaxf(dep, q.minmax[0]), q.minmax[1]);
      q.depth[ray] = dep;
    }
  }
  if (pk) {
    if (q.rgb) { pk[0] = q.rgb[ray * 3]; pk[1] = q.rgb[ray * 3 + 1]; pk[2] = q.rgb[ray * 3 + 2]; }
    pk[3] = dep;
  }
  if (!q.p || !q.jbar) return;
  const float px = q.p[ray * 3], py = q.p[ray * 3 + 1], pz = q.p[ray * 3 + 2];
  float f[3] = {0.f, 0.f, 0.f};
  for (int a = 0; a < q.A; ++a) {
    const float ua = __ldg(q.action + b * q.A + a);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float jv = q.jbar[static_cast<size_t>(ray) * A3 + a * 3 + d];
      f[d] = fmaf(jv, ua, f[d]);
      if (pk) pk[6 + a * 3 + d] = jv;
    }
  }
  const float wx = px + f[0], wy = py + f[1], wz = pz + f[2];
  if (q.pw) {
    q.pw[ray * 3] = wx;
    q.pw[ray * 3 + 1] = wy;
    q.pw[ray * 3 + 2] = wz;
  }
  float fu = 0.f, fv = 0.f;
  if (q.flow || pk) {
    float u0, v0, u1, v1;
    project_uv(q.trgt_w2c + b * 16, q.trgt_k + b * 9, px, py, pz, u0, v0);
    project_uv(q.trgt_w2c + b * 16, q.trgt_k + b * 9, wx, wy, wz, u1, v1);
    fu = u1 - u0;
    fv = v1 - v0;
  }
  if (q.flow) {
    q.flow[ray * 2] = fu;
    q.flow[ray * 2 + 1] = fv;
  }
  if (pk) {
    pk[4] = fu; pk[5] = fv;
    pk[6 + A3] = px; pk[7 + A3] = py; pk[8 + A3] = pz;
    pk[9 + A3] = wx; pk[10 + A3] = wy; pk[11 + A3] = wz;
  }
}

// hoisted maps: out[b][px][n] = fp16( sum_c W[n][c] * feat[b][c][px] + bias[n] ), n in [n0, n0+CH)
// classic 128x128x16 SIMT tile, 256 threads, 8x8 outputs per thread.
__global__ void __launch_bounds__(256) hoist_kernel(const float* __restrict__ feat, const float* __restrict__ Wt,
                                                    const float* __restrict__ bias, __half* __restrict__ out,
                                                    int HW, int CH, int n0) {
  __shared__ float sF[16][128 + 4];
  __shared__ float sW[16][128 + 4];
  const int b = blockIdx.z;
  const int px0 = blockIdx.x * 128, nn0 = blockIdx.y * 128;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // tx: 8 px, ty: 8 n
  const float* F = feat + static_cast<size_t>(b) * 512 * HW;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < 512; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 128; i += 256) {
      const int kk = i >> 7, pp = i & 127;
      const int px = px0 + pp;
      sF[kk][pp] = (px < HW) ? F[static_cast<size_t>(k0 + kk) * HW + px] : 0.f;
    }
    for (int i = threadIdx.x; i < 16 * 128; i += 256) {
      const int nn = i >> 4, kk = i & 15;
      const int n = nn0 + nn;
      sW[kk][nn] = (n < CH) ? Wt[static_cast<size_t>(n0 + n) * 512 + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float fv[8], wv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) fv[i] = sF[kk][tx * 8 + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = sW[kk][ty * 8 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(fv[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int px = px0 + tx * 8 + i;
    if (px >= HW) continue;
    const int n = nn0 + ty * 8;
    if (n >= CH) continue;  // CH is a multiple of 8
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      ow[j] = pack_f16x2(acc[i][2 * j] + __ldg(bias + n0 + n + 2 * j), acc[i][2 * j + 1] + __ldg(bias + n0 + n + 2 * j + 1));
    *reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * HW + px) * CH + n) = o;
  }
}

// standalone PDFSampler: one warp per ray
// `from_dd` != 0: the input holds delta*sigma; the warp first turns it into transmittance weights
// (RaySamples.get_weights, ray_samplers.py:77-101) and, if `weights_io` is writable, stores them back.
__global__ void pdf_kernel(float* weights_io, int from_dd, int store_weights, const float* bins_in, int bins_stride,
                           const float* u, int u_stride, int n_rays, int S, int n_out, float anneal, int sum_vec,
                           float* bins_out, int32_t* inds_out) {
  extern __shared__ float sm[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * (blockDim.x >> 5) + wib;
  float* w = sm + wib * (2 * S + 8);
  float* cdf = w + S;
  if (ray >= n_rays) return;
  for (int j = lane; j < S; j += 32) w[j] = weights_io[static_cast<size_t>(ray) * S + j];
  __syncwarp();
  if (from_dd) {
    double carry = 0.0;
    excl_scan_warp(w, S, carry, cdf);
    __syncwarp();
    for (int j = lane; j < S; j += 32) {
      const float wt = (1.f - expf(-w[j])) * expf(-cdf[j]);
      w[j] = wt;
      if (store_weights) weights_io[static_cast<size_t>(ray) * S + j] = wt;
    }
    __syncwarp();
  }
  const int nb = n_out + 1;
  pdf_resample_warp(w, S, bins_in + static_cast<size_t>(ray) * bins_stride, u + static_cast<size_t>(ray) * u_stride,
                    nb, anneal, sum_vec, cdf, bins_out + static_cast<size_t>(ray) * nb,
                    inds_out ? inds_out + static_cast<size_t>(ray) * nb : nullptr);
}

// standalone RaySamples.get_weights: one warp per ray
__global__ void tw_kernel(const float* deltas, const float* sigma, int n_rays, int S, float* out) {
  extern __shared__ float sm[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * (blockDim.x >> 5) + wib;
  float* dd = sm + wib * 2 * S;
  float* cum = dd + S;
  if (ray >= n_rays) return;
  for (int j = lane; j < S; j += 32) {
    const float d = deltas[static_cast<size_t>(ray) * S + j];
    dd[j] = d > 0.f ? __fmul_rn(d, sigma[static_cast<size_t>(ray) * S + j]) : 0.f;
  }
  __syncwarp();
  double carry = 0.0;
  excl_scan_warp(dd, S, carry, cum);
  __syncwarp();
  for (int j = lane; j < S; j += 32) out[static_cast<size_t>(ray) * S + j] = (1.f - expf(-dd[j])) * expf(-cum[j]);
}

// ---- by-products of Model.compute_density (DensityHeadOutput.xyz_features / pixel_aligned_features)
// one warp per point: the 63-column positional encoding of the context-camera point and the
// bilinear gather of the RAW encoder features (NCHW fp32, like F.grid_sample in the reference)
__global__ void point_features_kernel(const float* __restrict__ points, const float* __restrict__ w2c,
                                      const float* __restrict__ kn, const float* __restrict__ feat, int B, int N,
                                      int C, int Hf, int Wf, float* __restrict__ xyz_feat,
                                      float* __restrict__ pix_feat) {
  const int lane = threadIdx.x & 31;
  const int pt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pt >= B * N) return;
  PassGeom g{};
  g.NR = B * N; g.R = N; g.S = 1; g.G = kRows; g.T = 1;
  g.points = points; g.ctxt_w2c = w2c; g.ctxt_k = kn; g.Hf = Hf; g.Wf = Wf;
  RowState rs;
  row_setup(g, pt / kRows, 0, pt % kRows, rs);
  if (xyz_feat) {
    for (int c = lane; c < 63; c += 32) {
      float v;
      if (c >= 60) {
        v = rs.cam[c - 60];
      } else {
        const int cc = c < 30 ? c : c - 30, i = cc / 10, k = cc - 10 * i;
        float t = __fmul_rn(6.2831855f, rs.cam[i]) * static_cast<float>(1 << k);
        if (c >= 30) t = __fadd_rn(t, 1.5707964f);
        v = sin_cw(t);
      }
      xyz_feat[static_cast<size_t>(pt) * 63 + c] = v;
    }
  }
  if (pix_feat) {
    const float x0 = floorf(rs.ix), y0 = floorf(rs.iy);
    const float x1 = x0 + 1.f, y1 = y0 + 1.f;
    const float wnw = (x1 - rs.ix) * (y1 - rs.iy), wne = (rs.ix - x0) * (y1 - rs.iy);
    const float wsw = (x1 - rs.ix) * (rs.iy - y0), wse = (rs.ix - x0) * (rs.iy - y0);
    const int xi = static_cast<int>(x0), yi = static_cast<int>(y0);
    const int xj = min(xi + 1, Wf - 1), yj = min(yi + 1, Hf - 1);
    const int b = pt / N;
    const float* F = feat + static_cast<size_t>(b) * C * Hf * Wf;
    for (int c = lane; c < C; c += 32) {
      const float* fc = F + static_cast<size_t>(c) * Hf * Wf;
      const float v = fc[yi * Wf + xi] * wnw + fc[yi * Wf + xj] * wne + fc[yj * Wf + xi] * wsw + fc[yj * Wf + xj] * wse;
      pix_feat[static_cast<size_t>(pt) * C + c] = v;
    }
  }
}

}  // namespace njf

// ============================================================================= host launchers
using namespace njf;

namespace {

// per-device state: SM count and the >48 KB dynamic shared memory opt-in (cudaFuncSetAttribute applies to the
// CURRENT device only, so a process that renders on several GPUs must opt in on each of them)
constexpr int kMaxDevices = 64;
int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
int num_sms() {
  static std::atomic<int> sms[kMaxDevices];
  const int dev = current_device();
  int n = sms[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    sms[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

int make_geom(const NjfField* f, const NjfCameras* cams, const NjfRenderArgs* a, int S, const float* bins,
              int bins_stride, const __half* map, int CH, PassGeom& g) {
  if (S < 1 || S > 512) NJF_FAIL("samples per ray %d unsupported (1..512)", S);
  g.NR = a->n_rays > 0 ? a->n_rays : a->B * a->R;
  g.ray0 = a->n_rays > 0 ? a->ray_offset : 0;
  g.R = a->R;
  g.S = S;
  g.G = S <= kRows ? kRows / S : 1;
  g.T = S <= kRows ? 1 : (S + kRows - 1) / kRows;
  g.NG = (g.NR + g.G - 1) / g.G;
  g.origins = a->origins;
  g.dirs = a->dirs;
  g.z_near = a->z_near;
  g.z_far = a->z_far;
  g.bins = bins;
  g.bins_stride = bins_stride;
  g.ctxt_w2c = cams->ctxt_w2c;
  g.ctxt_k = cams->ctxt_k;
  g.map = map;
  g.CH = CH;
  g.Hf = a->Hf;
  g.Wf = a->Wf;
  const char* dbg = getenv("NJF_DEBUG_SKIP");
  g.debug = dbg ? atoi(dbg) : 0;
  g.points = nullptr;
  g.n_const_views = 0;
  if (cams->h_ctxt_w2c && cams->h_ctxt_k && a->h_z_near && a->h_z_far && a->B <= kMaxConstViews) {
    for (int b = 0; b < a->B; ++b) {
      float* vc = g.view_const[b];
      for (int i = 0; i < 12; ++i) vc[i] = cams->h_ctxt_w2c[b * 16 + i];
      for (int i = 0; i < 9; ++i) vc[12 + i] = cams->h_ctxt_k[b * 9 + i];
      vc[21] = a->h_z_near[b];
      vc[22] = a->h_z_far[b];
    }
    g.n_const_views = a->B;
  }
  (void)f;
  return 0;
}

const __half* map_of(const NjfField* f, const NjfRenderArgs* a, int level /* -1 = main */) {
  const size_t px = static_cast<size_t>(a->B) * a->Hf * a->Wf;
  const __half* base = static_cast<const __half*>(a->maps);
  if (level < 0) return base + px * f->ch_prop * f->desc.n_proposal;
  return base + px * f->ch_prop * level;
}

int check_args(const NjfField* f, const NjfCameras* cams, const NjfRenderArgs* a) {
  if (!f || !cams || !a) NJF_FAIL("null argument");
  if (a->B < 1 || a->R < 1) NJF_FAIL("B=%d R=%d: nothing to render", a->B, a->R);
  if (a->n_rays < 0 || a->ray_offset < 0 ||
      (a->n_rays > 0 && static_cast<long long>(a->ray_offset) + a->n_rays > static_cast<long long>(a->B) * a->R))
    NJF_FAIL("ray range [%d, %d + %d) outside the %d x %d call", a->ray_offset, a->ray_offset, a->n_rays, a->B, a->R);
  if (a->n_levels != f->desc.n_proposal) NJF_FAIL("n_levels %d != field n_proposal %d", a->n_levels, f->desc.n_proposal);
  if (!a->origins || !a->dirs || !a->z_near || !a->z_far || !a->maps) NJF_FAIL("missing ray / map input");
  if (a->Hf < 1 || a->Wf < 1 || a->Hf > 16384 || a->Wf > 16384) NJF_FAIL("feature map %dx%d out of range", a->Hf, a->Wf);
  if (a->s_nerf < 1 || a->s_nerf > 512) NJF_FAIL("num_nerf_samples %d unsupported (1..512)", a->s_nerf);
  if (static_cast<long long>(a->B) * a->R > 0x7fffffffLL / 512) NJF_FAIL("B*R = %lld rays: render in groups", static_cast<long long>(a->B) * a->R);
  if (static_cast<size_t>(a->B) * a->Hf * a->Wf * 768 * 2 >= (1ull << 32))
    NJF_FAIL("hoisted maps of %d views of %dx%d exceed the 4 GiB tap-offset range; render the views in groups", a->B, a->Hf, a->Wf);
  return 0;
}

template <class K>
int set_smem(K kernel) {
  static std::atomic<bool> done[kMaxDevices];  // one flag per (kernel instantiation, device)
  const int dev = current_device();
  if (!done[dev].load(std::memory_order_acquire)) {
    NJF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBytes)));
    done[dev].store(true, std::memory_order_release);
  }
  return 0;
}

// field_kernel (+ xf_kernel for the cross-attention head) over all ray groups of the pass.  The transformer
// hand-over (16.5 KB per 128-row tile: fp16 query embedding + fp32 sample weights) lives in the CALLER's
// workspace; a pass with more tiles than the workspace holds runs as several launch pairs over ray-group
// ranges that re-use the same bytes (so a small workspace stays L2-resident and never reaches HBM).
constexpr size_t kXfTileBytes = 8 * kRows * sizeof(uint4) + kRows * sizeof(float);
constexpr size_t kXfDefaultCap = 256ull << 20;  // njf_workspace_bytes: hand-over capped at 256 MiB
// optional per-kernel timing of the field pass (njf_debug_field_timing): CUDA events on the launching stream
bool g_time_field = false;
std::vector<cudaEvent_t> g_field_events;  // triples (before field_kernel, between, after xf_kernel) per launch pair
size_t g_field_events_used = 0;

// ray groups per launch pair for a hand-over region of `bytes`: whole rounds of all SMs (4 xf slots x SMs groups,
// which is also a whole number of field_kernel slot pairs) so that no launch ends with a partly filled wave
int xf_groups_per_launch(size_t bytes, int T, int NGtot) {
  static const long env_tiles = [] {  // NJF_XF_MAX_TILES: smaller launch pairs (tests of the chunked path, L2 studies)
    const char* v = getenv("NJF_XF_MAX_TILES");
    return v ? atol(v) : 0L;
  }();
  size_t tiles = bytes / kXfTileBytes;
  if (env_tiles > 0 && static_cast<size_t>(env_tiles) < tiles)  // never below one ray group if the workspace holds one
    tiles = (static_cast<size_t>(env_tiles) < static_cast<size_t>(T) && tiles >= static_cast<size_t>(T))
                ? static_cast<size_t>(T) : static_cast<size_t>(env_tiles);
  long gpc = static_cast<long>(tiles / T);
  if (gpc >= NGtot) return NGtot;
  const long quantum = 4L * num_sms();
  if (gpc > quantum) gpc -= gpc % quantum;
  return static_cast<int>(gpc);
}

int launch_field(const NjfField* f, FieldParams& p, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (set_smem(field_kernel)) return 1;
  const int NGtot = p.g.NG, T = p.g.T;
  const bool xf = f->desc.head == NJF_HEAD_TRANSFORMER && (p.jbar || p.jac_out);
  float* jbar = p.jbar;
  float* jac = p.jac_out;
  int gpc = NGtot;  // groups per launch
  if (xf) {
    if (!ws) NJF_FAIL("cross-attention head: NjfRenderArgs.workspace required (njf_workspace_bytes)");
    if (reinterpret_cast<uintptr_t>(ws) & 15) NJF_FAIL("workspace must be 16-byte aligned");
    gpc = xf_groups_per_launch(ws_bytes, T, NGtot);
    if (gpc < 1)
      NJF_FAIL("workspace of %zu bytes cannot hold one ray group of the query hand-over (%zu bytes)", ws_bytes,
               static_cast<size_t>(T) * kXfTileBytes);
    const size_t tiles = static_cast<size_t>(gpc) * T;
    p.qs = static_cast<uint4*>(ws);
    p.wts = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + tiles * 8 * kRows * sizeof(uint4));
  }
  if (f->desc.head == NJF_HEAD_TRANSFORMER) {
    p.jbar = nullptr;
    p.jac_out = nullptr;
  }
  for (int g0 = 0; g0 < NGtot; g0 += gpc) {
    p.g.group0 = g0;
    p.g.NG = (NGtot - g0 < gpc) ? NGtot - g0 : gpc;
    const int nitems = (p.g.NG + 1) / 2;
    const int grid = nitems < num_sms() ? nitems : num_sms();
    cudaEvent_t* ev = nullptr;
    if (g_time_field) {
      while (g_field_events.size() < g_field_events_used + 3) {
        cudaEvent_t e;
        NJF_CUDA(cudaEventCreate(&e));
        g_field_events.push_back(e);
      }
      ev = &g_field_events[g_field_events_used];
      g_field_events_used += 3;
      NJF_CUDA(cudaEventRecord(ev[0], stream));
    }
    field_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
    njf::count_launch();
    NJF_CUDA(cudaGetLastError());
    if (ev) NJF_CUDA(cudaEventRecord(ev[1], stream));
    if (xf) {
      XfParams x{};
      x.NR = p.g.NR; x.S = p.g.S; x.G = p.g.G; x.T = T; x.NG = p.g.NG; x.group0 = g0;
      x.qs = p.qs;
      x.wts = p.wts;
      x.jbar = jbar;
      x.jac_out = jac;
      if (njf_xf_launch(f, x, stream)) return 1;
    }
    if (ev) NJF_CUDA(cudaEventRecord(ev[2], stream));
  }
  return 0;
}

// tile geometry of a pass with S samples per ray
void tile_geom(int S, int NR, int& G, int& T, int& NG) {
  G = S <= kRows ? kRows / S : 1;
  T = S <= kRows ? 1 : (S + kRows - 1) / kRows;
  NG = (NR + G - 1) / G;
}

}  // namespace

#ifdef NJF_PROFILE
extern "C" int njf_prof_read(unsigned long long* out16, int reset) {
  NJF_CUDA(cudaDeviceSynchronize());
  NJF_CUDA(cudaMemcpyFromSymbol(out16, g_prof, 16 * sizeof(unsigned long long)));
  if (reset) {
    unsigned long long z[16] = {};
    NJF_CUDA(cudaMemcpyToSymbol(g_prof, z, sizeof(z)));
  }
  return 0;
}
#endif

extern "C" int njf_debug_field_timing(int enable, float* field_kernel_ms, float* xf_kernel_ms) {
  if (field_kernel_ms || xf_kernel_ms) {
    float tf = 0.f, tx = 0.f;
    for (size_t i = 0; i + 3 <= g_field_events_used && i + 3 <= g_field_events.size(); i += 3) {
      float a = 0.f, b = 0.f;
      NJF_CUDA(cudaEventSynchronize(g_field_events[i + 2]));
      NJF_CUDA(cudaEventElapsedTime(&a, g_field_events[i], g_field_events[i + 1]));
      NJF_CUDA(cudaEventElapsedTime(&b, g_field_events[i + 1], g_field_events[i + 2]));
      tf += a;
      tx += b;
    }
    if (field_kernel_ms) *field_kernel_ms = tf;
    if (xf_kernel_ms) *xf_kernel_ms = tx;
  }
  g_field_events_used = 0;
  g_time_field = enable != 0;
  return 0;
}

extern "C" size_t njf_workspace_min_bytes(const NjfField* f, int n_levels, const int* s_prop, int s_nerf) {
  if (!f || !s_prop) return 0;
  size_t need = 16;
  int G, T, NG;
  for (int l = 0; l < n_levels && l < NJF_MAX_LEVELS; ++l) {
    tile_geom(s_prop[l], 1, G, T, NG);
    const size_t b = static_cast<size_t>(G) * s_prop[l] * sizeof(float);  // one ray group of delta*sigma
    if (b > need) need = b;
  }
  if (f->desc.head == NJF_HEAD_TRANSFORMER) {
    tile_geom(s_nerf, 1, G, T, NG);
    const size_t b = static_cast<size_t>(T) * kXfTileBytes;  // one ray group of the query hand-over
    if (b > need) need = b;
  }
  return (need + 255) & ~static_cast<size_t>(255);
}

extern "C" size_t njf_workspace_bytes(const NjfField* f, int B, int R, int n_levels, const int* s_prop, int s_nerf) {
  if (!f || !s_prop || B < 1 || R < 1) return 0;
  const int NR = B * R;
  size_t need = njf_workspace_min_bytes(f, n_levels, s_prop, s_nerf);
  int G, T, NG;
  for (int l = 0; l < n_levels && l < NJF_MAX_LEVELS; ++l) {
    size_t b = static_cast<size_t>(NR) * s_prop[l] * sizeof(float);
    if (b > kXfDefaultCap) b = kXfDefaultCap;
    if (b > need) need = b;
  }
  if (f->desc.head == NJF_HEAD_TRANSFORMER) {
    tile_geom(s_nerf, NR, G, T, NG);
    size_t b = static_cast<size_t>(NG) * T * kXfTileBytes;
    if (b > kXfDefaultCap) {
      // whole rounds of all SMs per launch pair (xf_groups_per_launch), at most the cap
      const size_t round = static_cast<size_t>(4) * num_sms() * T * kXfTileBytes;
      b = (kXfDefaultCap / round) * round;
      if (b == 0) b = round;
    }
    if (b > need) need = b;
  }
  return (need + 255) & ~static_cast<size_t>(255);
}

extern "C" size_t njf_query_workspace_bytes(const NjfField* f, int B, int N) {
  if (!f || f->desc.head != NJF_HEAD_TRANSFORMER || B < 1 || N < 1) return 256;
  size_t tiles = (static_cast<size_t>(B) * N + kRows - 1) / kRows;
  size_t b = tiles * kXfTileBytes;
  if (b > kXfDefaultCap) {
    const size_t round = static_cast<size_t>(4) * num_sms() * kXfTileBytes;
    b = (kXfDefaultCap / round) * round;
  }
  return (b + 255) & ~static_cast<size_t>(255);
}

extern "C" size_t njf_hoisted_bytes(const NjfField* f, int B, int Hf, int Wf) {
  return static_cast<size_t>(B) * Hf * Wf * f->ch_total * sizeof(__half);
}

extern "C" int njf_hoist_features(const NjfField* f, const float* feat_nchw, int B, int Hf, int Wf, void* maps_out,
                                  void* stream_) {
  if (!f || !feat_nchw || !maps_out) NJF_FAIL("njf_hoist_features: null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static const bool simt = getenv("NJF_HOIST_SIMT") != nullptr;  // fp32 SIMT variant, kept for A/B checks
  if (!simt) return njf_hoist_launch(f, feat_nchw, B, Hf, Wf, maps_out, stream);
  const int HW = Hf * Wf;
  __half* out = static_cast<__half*>(maps_out);
  int n0 = 0;
  for (int m = 0; m <= f->desc.n_proposal; ++m) {
    const int CH = (m < f->desc.n_proposal) ? f->ch_prop : f->ch_main;
    dim3 grid((HW + 127) / 128, (CH + 127) / 128, B);
    hoist_kernel<<<grid, 256, 0, stream>>>(feat_nchw, f->d_hoist_w, f->d_hoist_b, out, HW, CH, n0);
    njf::count_launch();
    NJF_CUDA(cudaGetLastError());
    out += static_cast<size_t>(B) * HW * CH;
    n0 += CH;
  }
  return 0;
}

extern "C" int njf_hoist_features_views(const NjfField* f, const float* feat_nchw, int B_local, int view0, int B_total,
                                        int Hf, int Wf, void* maps_out, void* stream_) {
  if (!f || !feat_nchw || !maps_out) NJF_FAIL("njf_hoist_features_views: null argument");
  if (B_local < 1 || view0 < 0 || view0 + B_local > B_total) NJF_FAIL("njf_hoist_features_views: views [%d, %d) outside %d", view0, view0 + B_local, B_total);
  return njf_hoist_launch(f, feat_nchw, B_local, Hf, Wf, maps_out, static_cast<cudaStream_t>(stream_), view0, B_total);
}

extern "C" int njf_hoist_features_nhwc16(const NjfField* f, const void* feat_nhwc_f16, int B_local, int view0, int B_total,
                                         int Hf, int Wf, void* maps_out, void* stream_) {
  if (!f || !feat_nhwc_f16 || !maps_out) NJF_FAIL("njf_hoist_features_nhwc16: null argument");
  if (B_local < 1 || view0 < 0 || view0 + B_local > B_total) NJF_FAIL("njf_hoist_features_nhwc16: views [%d, %d) outside %d", view0, view0 + B_local, B_total);
  if (reinterpret_cast<uintptr_t>(feat_nhwc_f16) & 15) NJF_FAIL("njf_hoist_features_nhwc16: the feature map must be 16-byte aligned");
  return njf_hoist_launch(f, nullptr, B_local, Hf, Wf, maps_out, static_cast<cudaStream_t>(stream_), view0, B_total, feat_nhwc_f16);
}

extern "C" int njf_proposal_pass(const NjfField* f, const NjfCameras* cams, const NjfRenderArgs* a, int level,
                                 const float* bins_in, int bins_in_stride, void* stream_) {
  if (check_args(f, cams, a)) return 1;
  if (level < 0 || level >= a->n_levels) NJF_FAIL("level %d out of range", level);
  if (!a->level_bins[level] || !a->u[level]) NJF_FAIL("level_bins[%d] / u[%d] required", level, level);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ProposalParams p{};
  p.prog = f->prop_prog[level];
  p.blob = f->prop_blob[level];
  if (make_geom(f, cams, a, a->s_prop[level], bins_in, bins_in_stride, map_of(f, a, level), f->ch_prop, p.g)) return 1;
  p.n_out = (level + 1 < a->n_levels) ? a->s_prop[level + 1] : a->s_nerf;
  if (p.n_out < 1 || p.n_out > 512) NJF_FAIL("n_out %d unsupported", p.n_out);
  p.u = a->u[level];
  p.u_stride = a->u_stride[level];
  p.anneal = a->anneal;
  p.sum_vec = a->sum_vec_width;
  p.bins_out = a->level_bins[level];
  p.weights_out = a->prop_weights[level];
  p.inds_out = a->level_inds[level];
  // delta*sigma travels from proposal_kernel to pdf_kernel through the caller's prop_weights output or, when that
  // is NULL, through the caller's workspace; a workspace smaller than [NR][S] floats splits the level into
  // launch pairs over ray ranges
  const int NGtot = p.g.NG, G = p.g.G, S = p.g.S, NR = p.g.NR;
  int gpc = NGtot;
  float* scratch = nullptr;
  if (!p.weights_out) {
    if (!a->workspace) NJF_FAIL("njf_proposal_pass: prop_weights[%d] or NjfRenderArgs.workspace required", level);
    if (reinterpret_cast<uintptr_t>(a->workspace) & 15) NJF_FAIL("workspace must be 16-byte aligned");
    scratch = static_cast<float*>(a->workspace);
    const size_t rays_fit = a->workspace_bytes / (static_cast<size_t>(S) * sizeof(float));
    long g = static_cast<long>(rays_fit / G);
    if (g < 1) NJF_FAIL("workspace of %zu bytes cannot hold one ray group of proposal level %d", a->workspace_bytes, level);
    if (g < NGtot) {
      const long quantum = 2L * num_sms();
      if (g > quantum) g -= g % quantum;
      gpc = static_cast<int>(g);
    }
  }
  if (set_smem(proposal_kernel)) return 1;
  // pdf_kernel: one warp per ray; 4 warps per block while their scan buffers fit the default 48 KB
  const int wpb = (static_cast<size_t>(4) * (2 * S + 8) * sizeof(float) <= 48 * 1024) ? 4 : 1;
  const size_t pdf_smem = static_cast<size_t>(wpb) * (2 * S + 8) * sizeof(float);
  const int nb = p.n_out + 1;
  for (int g0 = 0; g0 < NGtot; g0 += gpc) {
    p.g.group0 = g0;
    p.g.NG = (NGtot - g0 < gpc) ? NGtot - g0 : gpc;
    const int ray0 = g0 * G;
    const int nrays = (NR - ray0 < p.g.NG * G) ? NR - ray0 : p.g.NG * G;
    float* wbuf;
    if (scratch) {
      p.weights_out = scratch;
      p.w_ray0 = ray0;
      wbuf = scratch;
    } else {
      p.w_ray0 = 0;
      wbuf = a->prop_weights[level] + static_cast<size_t>(ray0) * S;
    }
    const int nitems = (p.g.NG + 1) / 2;
    const int grid = nitems < num_sms() ? nitems : num_sms();
    proposal_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
    njf::count_launch();
    NJF_CUDA(cudaGetLastError());
    pdf_kernel<<<(nrays + wpb - 1) / wpb, wpb * 32, pdf_smem, stream>>>(
        wbuf, /*from_dd=*/1, /*store_weights=*/a->prop_weights[level] != nullptr,
        bins_in + static_cast<size_t>(ray0) * bins_in_stride, bins_in_stride, p.u + static_cast<size_t>(ray0) * p.u_stride,
        p.u_stride, nrays, S, p.n_out, p.anneal, p.sum_vec, p.bins_out + static_cast<size_t>(ray0) * nb,
        p.inds_out ? p.inds_out + static_cast<size_t>(ray0) * nb : nullptr);
    njf::count_launch();
    NJF_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int njf_field_pass(const NjfField* f, const NjfCameras* cams, const NjfRenderArgs* a, const float* bins,
                              int bins_stride, void* stream_) {
  if (check_args(f, cams, a)) return 1;
  if (!a->minmax) NJF_FAIL("minmax workspace required");
  if (!a->action) NJF_FAIL("action required");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FieldParams p{};
  p.prog = f->field_prog;
  p.blob = f->field_blob;
  p.color = f->color;
  if (make_geom(f, cams, a, a->s_nerf, bins, bins_stride, map_of(f, a, -1), f->ch_main, p.g)) return 1;
  p.head_kind = f->desc.head;
  p.A = f->desc.action_dim;
  p.sh_conv = f->desc.sh_convention;
  p.rgb = a->rgb;
  p.depth = a->depth;
  p.jbar = a->jbar;
  p.p = a->p;
  p.steps = a->steps;
  p.weights = a->weights;
  p.sigma = a->sigma;
  p.jac_out = a->jac;
  p.positions = a->positions;
  p.rgb_samples = a->rgb_samples;
  p.minmax = reinterpret_cast<uint32_t*>(a->minmax);
  init_minmax_kernel<<<1, 1, 0, stream>>>(p.minmax);
  njf::count_launch();
  if (launch_field(f, p, a->workspace, a->workspace_bytes, stream)) return 1;
  decode_minmax_kernel<<<1, 1, 0, stream>>>(p.minmax);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int njf_query_points(const NjfField* f, const float* ctxt_w2c, const float* ctxt_k, const void* maps,
                                int Hf, int Wf, const float* points, const float* dirs, int B, int N, float* sigma,
                                float* geo, float* jac, float* rgb, void* workspace, size_t workspace_bytes,
                                void* stream_) {
  if (!f || !ctxt_w2c || !ctxt_k || !maps || !points) NJF_FAIL("njf_query_points: null argument");
  if (B < 1 || N < 1) NJF_FAIL("njf_query_points: B=%d N=%d", B, N);
  if (static_cast<size_t>(B) * Hf * Wf * 768 * 2 >= (1ull << 32)) NJF_FAIL("njf_query_points: maps too large");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FieldParams p{};
  p.prog = f->field_prog;
  p.blob = f->field_blob;
  p.color = f->color;
  PassGeom& g = p.g;
  g.NR = B * N; g.R = N; g.S = 1; g.G = kRows; g.T = 1;
  g.NG = (g.NR + g.G - 1) / g.G;
  g.ctxt_w2c = ctxt_w2c;
  g.ctxt_k = ctxt_k;
  const size_t px = static_cast<size_t>(B) * Hf * Wf;
  g.map = static_cast<const __half*>(maps) + px * f->ch_prop * f->desc.n_proposal;
  g.CH = f->ch_main; g.Hf = Hf; g.Wf = Wf;
  g.points = points;
  g.dirs = dirs;  // per point (point i plays the role of "ray" i); NULL: the colour head sees direction (0,0,1)
  if (rgb && !dirs) NJF_FAIL("njf_query_points: rgb output needs per-point view directions");
  p.head_kind = f->desc.head;
  p.A = f->desc.action_dim;
  p.sh_conv = f->desc.sh_convention;
  p.sigma = sigma;
  p.geo_out = geo;
  p.jac_out = jac;
  p.rgb_samples = rgb;
  return launch_field(f, p, workspace, workspace_bytes, stream);
}

extern "C" int njf_query_proposal_density(const NjfField* f, int level, const float* ctxt_w2c, const float* ctxt_k,
                                          const void* maps, int Hf, int Wf, const float* points, int B, int N,
                                          float* sigma, void* stream_) {
  if (!f || !ctxt_w2c || !ctxt_k || !maps || !points || !sigma) NJF_FAIL("njf_query_proposal_density: null argument");
  if (level < 0 || level >= f->desc.n_proposal) NJF_FAIL("njf_query_proposal_density: level %d out of range", level);
  if (B < 1 || N < 1) NJF_FAIL("njf_query_proposal_density: B=%d N=%d", B, N);
  if (static_cast<size_t>(B) * Hf * Wf * 768 * 2 >= (1ull << 32)) NJF_FAIL("njf_query_proposal_density: maps too large");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ProposalParams p{};
  p.prog = f->prop_prog[level];
  p.blob = f->prop_blob[level];
  PassGeom& g = p.g;
  g.NR = B * N; g.R = N; g.S = 1; g.G = kRows; g.T = 1;
  g.NG = (g.NR + g.G - 1) / g.G;
  g.ctxt_w2c = ctxt_w2c;
  g.ctxt_k = ctxt_k;
  const size_t px = static_cast<size_t>(B) * Hf * Wf;
  g.map = static_cast<const __half*>(maps) + px * f->ch_prop * level;
  g.CH = f->ch_prop; g.Hf = Hf; g.Wf = Wf;
  g.points = points;
  p.sigma_out = sigma;
  if (set_smem(proposal_kernel)) return 1;
  const int nitems = (g.NG + 1) / 2;
  const int grid = nitems < num_sms() ? nitems : num_sms();
  proposal_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int njf_point_features(const float* feat_nchw, const float* ctxt_w2c, const float* ctxt_k,
                                  const float* points, int B, int N, int C, int Hf, int Wf, float* xyz_features,
                                  float* pixel_aligned_features, void* stream_) {
  if (!ctxt_w2c || !ctxt_k || !points) NJF_FAIL("njf_point_features: null argument");
  if (pixel_aligned_features && !feat_nchw) NJF_FAIL("njf_point_features: feature map required");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int wpb = 8;
  point_features_kernel<<<(B * N + wpb - 1) / wpb, wpb * 32, 0, stream>>>(points, ctxt_w2c, ctxt_k, feat_nchw, B, N, C,
                                                                       Hf, Wf, xyz_features, pixel_aligned_features);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int njf_finish_pass(const NjfField* f, const NjfCameras* cams, const NjfRenderArgs* a, void* stream_) {
  if (check_args(f, cams, a)) return 1;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FinishParams q{};
  q.NR = a->n_rays > 0 ? a->n_rays : a->B * a->R;
  q.ray0 = a->n_rays > 0 ? a->ray_offset : 0;
  q.R = a->R;
  q.A = f->desc.action_dim;
  q.action = a->action;
  q.trgt_w2c = cams->trgt_w2c;
  q.trgt_k = cams->trgt_k_px;
  q.minmax = a->minmax;
  q.depth = a->depth;
  q.jbar = a->jbar;
  q.p = a->p;
  q.pw = a->pw;
  q.flow = a->flow;
  q.rgb = a->rgb;
  q.packed = a->packed;
  if (q.packed && (!q.jbar || !q.p || !q.rgb || !q.depth)) NJF_FAIL("packed output needs rgb, depth, jbar and p");
  if ((q.flow || q.pw || q.packed) && (!q.jbar || !q.p || !q.action || !q.trgt_w2c || !q.trgt_k))
    NJF_FAIL("flow / pw outputs need jbar, p, action and the target camera");
  finish_kernel<<<(q.NR + 255) / 256, 256, 0, stream>>>(q);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int njf_render_forward(const NjfField* f, const NjfCameras* cams, const NjfRenderArgs* a, void* stream) {
  if (check_args(f, cams, a)) return 1;
  if (!a->bins0) NJF_FAIL("bins0 required");
  const float* bins = a->bins0;
  int stride = a->bins0_stride;
  for (int l = 0; l < a->n_levels; ++l) {
    if (njf_proposal_pass(f, cams, a, l, bins, stride, stream)) return 1;
    bins = a->level_bins[l];
    stride = ((l + 1 < a->n_levels) ? a->s_prop[l + 1] : a->s_nerf) + 1;
  }
  if (njf_field_pass(f, cams, a, bins, stride, stream)) return 1;
  return njf_finish_pass(f, cams, a, stream);
}

extern "C" int njf_pdf_sample(const float* weights, const float* bins_in, int bins_in_stride, const float* u,
                              int u_stride, int n_rays, int s_in, int n_out, float anneal, int sum_vec_width,
                              float* bins_out, int32_t* inds_out, void* stream_) {
  if (!weights || !bins_in || !u || !bins_out) NJF_FAIL("njf_pdf_sample: null argument");
  if (s_in < 1 || s_in > 4096 || n_out < 1) NJF_FAIL("njf_pdf_sample: bad sizes (1 <= s_in <= 4096, n_out >= 1)");
  if (n_rays < 0) NJF_FAIL("njf_pdf_sample: negative ray count");
  if (n_rays == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int wpb = (static_cast<size_t>(4) * (2 * s_in + 8) * sizeof(float) <= 48 * 1024) ? 4 : 1;  // s_in <= 4096: <= 32.8 KB
  const size_t smem = static_cast<size_t>(wpb) * (2 * s_in + 8) * sizeof(float);
  pdf_kernel<<<(n_rays + wpb - 1) / wpb, wpb * 32, smem, stream>>>(const_cast<float*>(weights), 0, 0, bins_in,
                                                                 bins_in_stride, u, u_stride, n_rays, s_in, n_out,
                                                                 anneal, sum_vec_width, bins_out, inds_out);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int njf_transmittance_weights(const float* deltas, const float* sigma, int n_rays, int s,
                                         float* weights_out, void* stream_) {
  if (!deltas || !sigma || !weights_out) NJF_FAIL("njf_transmittance_weights: null argument");
  if (s < 1 || s > 4096) NJF_FAIL("njf_transmittance_weights: bad sizes");
  if (n_rays == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int wpb = (static_cast<size_t>(4) * 2 * s * sizeof(float) <= 48 * 1024) ? 4 : 1;  // s <= 4096: <= 32 KB
  tw_kernel<<<(n_rays + wpb - 1) / wpb, wpb * 32, static_cast<size_t>(wpb) * 2 * s * sizeof(float), stream>>>(
      deltas, sigma, n_rays, s, weights_out);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int njf_flow_from_encoding(const float* jbar, const float* p, const float* action, const float* trgt_w2c,
                                      const float* trgt_k_px, int n_rays, int rays_per_view, int action_dim,
                                      float* flow, float* pw, void* stream_) {
  if (!jbar || !p || !action || !trgt_w2c || !trgt_k_px) NJF_FAIL("njf_flow_from_encoding: null argument");
  if (n_rays == 0) return 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FinishParams q{};
  q.NR = n_rays;
  q.R = rays_per_view;
  q.A = action_dim;
  q.action = action;
  q.trgt_w2c = trgt_w2c;
  q.trgt_k = trgt_k_px;
  q.jbar = jbar;
  q.p = p;
  q.pw = pw;
  q.flow = flow;
  finish_kernel<<<(n_rays + 255) / 256, 256, 0, stream>>>(q);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}
