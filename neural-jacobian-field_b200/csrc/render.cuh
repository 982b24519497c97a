// Device-side building blocks of the render kernels: per-row ray/sample geometry, positional
// encoding into the tcgen05 A tile, the warp-cooperative bilinear gather of hoisted feature
// channels, the ResnetFC trunk driver, and the per-ray scan / PDF-resampling routines.
#pragma once
#include "field.h"

#ifndef NJF_GATHER_U
#define NJF_GATHER_U 8
#endif

namespace njf {

// ----------------------------------------------------------------------------- pass geometry
struct PassGeom {
  int NR;   // rays of this call (B*R, or the length of a ray-sharded range)
  int ray0; // index of the call's first ray in the flattened (view, ray) space (ray-sharded calls), else 0
  int R;    // rays per view
  int S;    // samples per ray in this pass
  int G;    // rays per 128-row tile (S <= 128) else 1
  int T;    // tiles per ray group (ceil(S/128) when S > 128) else 1
  int NG;   // ray groups handled by this launch (ceil(NR / G) unless the pass is chunked)
  int group0;  // first ray group of this launch
  const float* origins;   // [NR][3]
  const float* dirs;      // [NR][3]
  const float* z_near;    // [B]
  const float* z_far;     // [B]
  const float* bins;      // spacing-domain bin edges, [S+1] shared or per ray
  int bins_stride;        // 0 or S+1
  const float* ctxt_w2c;  // [B][16]
  const float* ctxt_k;    // [B][9]
  const __half* map;      // hoisted map of this pass [B][Hf*Wf][CH]
  int CH, Hf, Wf;
  int debug;              // NJF_DEBUG_SKIP bits (timing attribution only): 1 skip gather, 2 skip posenc
  const float* points;    // point-query mode (Model.compute_density): [NR][3] world points, S == 1
  // per-view constants of the first kMaxConstViews views, carried in the parameter constant bank
  // (row set-up then needs no global loads for cameras / near / far); n_const_views == 0 -> use pointers
  int n_const_views;
  float view_const[16][24];  // [view][ w2c rows 0-2 (12) | K (9) | near | far | pad ]
};
constexpr int kMaxConstViews = 16;

struct RowState {
  float pos[3];   // world-space sample position (RaySamples.get_positions, ray_samplers.py:48-55)
  float cam[3];   // context-camera coordinates (what the positional encoding sees)
  float tmid;     // (start+end)/2
  float delta;    // end-start
  float ix, iy;   // border-clamped, un-normalised feature-map coordinates
  int pixbase;    // view*Hf*Wf, or -1 for a padding row
  int ray;        // global ray index or -1
  int s;          // sample index along the ray
};

// All arithmetic that decides sample placement mirrors the reference op for op (explicit
// round-to-nearest intrinsics so that nvcc cannot contract mul+add into fma).
__device__ __forceinline__ void row_setup(const PassGeom& g, int group, int tile, int row, RowState& rs) {
  int lr, s;
  if (g.T == 1) {
    lr = row / g.S;
    s = row - lr * g.S;
    if (lr >= g.G) lr = -1;
  } else {
    lr = 0;
    s = tile * kRows + row;
    if (s >= g.S) lr = -1;
  }
  const int ray = (lr < 0) ? -1 : group * g.G + lr;
  rs.s = s;
  if (ray < 0 || ray >= g.NR) {
    rs.ray = -1;
    rs.pixbase = -1;
    rs.tmid = 0.f;
    rs.delta = 0.f;
    rs.ix = rs.iy = 0.f;
    rs.pos[0] = rs.pos[1] = rs.pos[2] = 0.f;
    rs.cam[0] = rs.cam[1] = rs.cam[2] = 0.f;
    return;
  }
  rs.ray = ray;
  const int b = (ray + g.ray0) / g.R;
  const bool cv = b < g.n_const_views;
  const float* vc = g.view_const[cv ? b : 0];
  if (g.points) {  // explicit world-space points instead of (ray, bin) samples
    rs.tmid = 0.f;
    rs.delta = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) rs.pos[i] = __ldg(g.points + static_cast<size_t>(ray) * 3 + i);
  } else {
  const float* bp = g.bins + static_cast<size_t>(ray) * g.bins_stride;
  const float b0 = __ldg(bp + s), b1 = __ldg(bp + s + 1);
  const float nr = cv ? vc[21] : __ldg(g.z_near + b), fr = cv ? vc[22] : __ldg(g.z_far + b);
  // spacing -> euclidean: x * s_far + (1 - x) * s_near  (ray_samplers.py:242-245)
  const float st = __fadd_rn(__fmul_rn(b0, fr), __fmul_rn(__fsub_rn(1.f, b0), nr));
  const float en = __fadd_rn(__fmul_rn(b1, fr), __fmul_rn(__fsub_rn(1.f, b1), nr));
  const float se = __fadd_rn(st, en);
  rs.tmid = __fmul_rn(se, 0.5f);
  rs.delta = __fsub_rn(en, st);
  const float* o = g.origins + static_cast<size_t>(ray) * 3;
  const float* d = g.dirs + static_cast<size_t>(ray) * 3;
#pragma unroll
  for (int i = 0; i < 3; ++i)  // origins + directions * (starts + ends) / 2
    rs.pos[i] = __fadd_rn(__ldg(o + i), __fmul_rn(__fmul_rn(__ldg(d + i), se), 0.5f));
  }
  // world -> context camera (pixel_aligned_features.py:18-20, geometry.py:59-65)
  float Wm[12], Km[9];
  if (cv) {
#pragma unroll
    for (int i = 0; i < 12; ++i) Wm[i] = vc[i];
#pragma unroll
    for (int i = 0; i < 9; ++i) Km[i] = vc[12 + i];
  } else {
#pragma unroll
    for (int i = 0; i < 12; ++i) Wm[i] = __ldg(g.ctxt_w2c + b * 16 + i);
#pragma unroll
    for (int i = 0; i < 9; ++i) Km[i] = __ldg(g.ctxt_k + b * 9 + i);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
    rs.cam[i] = fmaf(Wm[4 * i + 2], rs.pos[2], fmaf(Wm[4 * i + 1], rs.pos[1], fmaf(Wm[4 * i], rs.pos[0], Wm[4 * i + 3])));
  // project with normalised intrinsics, z-divide with +1e-9 (geometry.py:137-154)
  const float u = fmaf(Km[2], rs.cam[2], fmaf(Km[1], rs.cam[1], Km[0] * rs.cam[0]));
  const float v = fmaf(Km[5], rs.cam[2], fmaf(Km[4], rs.cam[1], Km[3] * rs.cam[0]));
  const float w = fmaf(Km[8], rs.cam[2], fmaf(Km[7], rs.cam[1], Km[6] * rs.cam[0]));
  const float zd = __fadd_rn(w, 1e-9f);
  const float un = __fdiv_rn(u, zd), vn = __fdiv_rn(v, zd);
  // grid = (uv - 0.5) * 2 ; align_corners=True: ((g + 1) / 2) * (size - 1) ; border clamp
  const float gx = __fmul_rn(__fsub_rn(un, 0.5f), 2.f), gy = __fmul_rn(__fsub_rn(vn, 0.5f), 2.f);
  float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), static_cast<float>(g.Wf - 1));
  float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), static_cast<float>(g.Hf - 1));
  rs.ix = fminf(fmaxf(ix, 0.f), static_cast<float>(g.Wf - 1));
  rs.iy = fminf(fmaxf(iy, 0.f), static_cast<float>(g.Hf - 1));
  if (!(rs.ix == rs.ix)) rs.ix = 0.f;  // NaN guard (degenerate projection)
  if (!(rs.iy == rs.iy)) rs.iy = 0.f;
  rs.pixbase = b * g.Hf * g.Wf;
}

// sin(t) for |t| up to ~1e5: 3-term Cody-Waite reduction by 2*pi (FMA, constants with short
// mantissas so n*C1 is exact), then the SFU sine on [-pi, pi] (abs error ~4e-7 incl. reduction;
// checked against float64 in DESIGN.md section 5).  The argument is the SAME fp32 number the
// reference feeds to torch.sin, so its large-argument rounding behaviour is reproduced.
__device__ __forceinline__ float sin_cw(float t) {
  // round-to-nearest-even of t / 2pi by the 1.5 * 2^23 trick (FMA pipe instead of FRND; |t / 2pi| < 2^22)
  const float n = __fsub_rn(__fmaf_rn(t, 0.15915494f, 12582912.f), 12582912.f);
  float r = fmaf(n, -6.28125f, t);
  r = fmaf(n, -1.9350052e-3f, r);
  r = fmaf(n, -3.019916e-7f, r);
  return __sinf(r);
}

// NeRFEncoding(63) of the camera-space point into the A tile.  nerfstudio's column order is
// [sin block | cos block | x], each block dim-major / freq-minor, cos(t) evaluated as sin(t + pi/2).  Here
// the row's thread of column half 0 writes the 30 sin columns (+2 zero columns) of K-block 0 and the thread
// of half 1 the 30 cos columns (+2 zeros), so (dim, freq) of every column is a compile-time constant; the
// raw xyz goes to the first 16 columns of K-block 1 as [x_hi | x_lo | x_hi | 0] (fp16 head + fp16 remainder,
// ~22 bits through an fp16 tensor-core product).  field.cu enc_cols packs lin_in / query-MLP weights to match.
__device__ __forceinline__ void write_posenc(const EpiCtx& e, const float (&cam)[3], bool valid,
                                             int debug = 0) {
  if (debug & 2) valid = false;
  const float off = e.half ? 1.5707964f : 0.f;  // t + 0 is exact
  const float base[3] = {__fmul_rn(6.2831855f, cam[0]), __fmul_rn(6.2831855f, cam[1]), __fmul_rn(6.2831855f, cam[2])};
  float v[32];
#pragma unroll
  for (int j = 0; j < 30; ++j) {
    const int i = j / 10, k = j - 10 * i;
    const float t = __fadd_rn(base[i] * static_cast<float>(1 << k), off);  // exact power-of-two scaling
    v[j] = sin_cw(t);
  }
  v[30] = v[31] = 0.f;
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) pk[j] = valid ? pack_f16x2(v[2 * j], v[2 * j + 1]) : 0u;
  a_store32(e, 32 * e.half, pk);
  if (e.half == 0) {
    float hi[3], lo[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const __half h = __ushort_as_half(static_cast<unsigned short>(pack_f16x2(cam[i], 0.f) & 0xffffu));  // satfinite
      hi[i] = __half2float(h);
      lo[i] = cam[i] - hi[i];
    }
    uint4 c0 = make_uint4(0u, 0u, 0u, 0u), c1 = make_uint4(0u, 0u, 0u, 0u);
    if (valid) {
      c0 = make_uint4(pack_f16x2(hi[0], hi[1]), pack_f16x2(hi[2], lo[0]), pack_f16x2(lo[1], lo[2]), pack_f16x2(hi[0], hi[1]));
      c1.x = pack_f16x2(hi[2], 0.f);
    }
    uint8_t* rowp = e.a_tile + kAKbStride + e.row * 128;
    *reinterpret_cast<uint4*>(rowp + ((0 ^ (e.row & 7)) << 4)) = c0;
    *reinterpret_cast<uint4*>(rowp + ((1 ^ (e.row & 7)) << 4)) = c1;
  }
}

// ----------------------------------------------------------------------------- gather
// Per-row bilinear taps (F.grid_sample align_corners=True, padding_mode="border"): computed once per row (by the
// thread of the row whose warp owns it: rows 32q+16h.. belong to warp (q,h)) and read by the gathering lanes of
// both warps of the row quarter for every channel segment; a named barrier between the two warps fences the table
// once per tile (after the writes, and before the next tile's writes).
struct TapEntry {
  uint32_t off[4];  // BYTE offset of the nw, ne, sw, se tap pixels inside the hoisted map (< 4 GiB)
  uint32_t w2[4];   // their weights as packed fp16 pairs {w, w} (all zero for padding rows)
};
__device__ __forceinline__ void write_taps(TapEntry* tab, int row, const RowState& rs, int Hf, int Wf, int CH) {
  uint4 px = make_uint4(0u, 0u, 0u, 0u);
  float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rs.pixbase >= 0) {
    const float x0 = floorf(rs.ix), y0 = floorf(rs.iy);
    const float x1 = x0 + 1.f, y1 = y0 + 1.f;
    w.x = (x1 - rs.ix) * (y1 - rs.iy);
    w.y = (rs.ix - x0) * (y1 - rs.iy);
    w.z = (x1 - rs.ix) * (rs.iy - y0);
    w.w = (rs.ix - x0) * (rs.iy - y0);
    const int xi = static_cast<int>(x0), yi = static_cast<int>(y0);
    const int xj = min(xi + 1, Wf - 1), yj = min(yi + 1, Hf - 1);  // out-of-range taps have weight 0
    const uint32_t pb = static_cast<uint32_t>(CH) * 2u;  // bytes per pixel
    px.x = static_cast<uint32_t>(rs.pixbase + yi * Wf + xi) * pb;
    px.y = static_cast<uint32_t>(rs.pixbase + yi * Wf + xj) * pb;
    px.z = static_cast<uint32_t>(rs.pixbase + yj * Wf + xi) * pb;
    px.w = static_cast<uint32_t>(rs.pixbase + yj * Wf + xj) * pb;
  }
  *reinterpret_cast<uint4*>(tab[row].off) = px;
  *reinterpret_cast<uint4*>(tab[row].w2) =
      make_uint4(pack_f16x2(w.x, w.x), pack_f16x2(w.y, w.y), pack_f16x2(w.z, w.z), pack_f16x2(w.w, w.w));
}

// Gather of NCH hoisted channels starting at channel ch0.  Warp (q, h) gathers, for ALL 32 rows of its row quarter,
// exactly the channel half [ch0 + h*NCH/2, ch0 + (h+1)*NCH/2) that its own threads consume afterwards (a thread
// owns columns 64h.. of its row): the staging buffer is written and read by the same warp, so the only fence
// is a __syncwarp -- no named barrier between the two warps of a row quarter, and no waiting for the slower one.
// A tap is one contiguous NCH-byte read spread over NCH/8 lanes (8 B per lane; 32*8/NCH rows per instruction).
// The four taps are blended in packed fp16 (HFMA2; the map, the weights and the staged result are fp16 anyway --
// 4 instead of 1 fp16 roundings).
// gather_rows: row groups [j_begin, j_end) of the warp's kGroups<NCH> groups; the trunk driver spreads a segment
// over BOTH MMA wait windows of a residual block (fc_0 and fc_1) instead of stacking it behind one.
template <int NCH>
struct GatherShape {
  static constexpr int kLanesPerRow = NCH / 8;              // 16 (128 channels) or 8 (64 channels)
  static constexpr int kRowsPerInstr = 32 / kLanesPerRow;   // 2 or 4
  static constexpr int kGroups = 32 / kRowsPerInstr;        // 16 or 8 row groups per warp and segment
};
template <int NCH>
__device__ __forceinline__ void gather_rows(EpiCtx& e, const PassGeom& g, const TapEntry* taps, int ch0, int j_begin,
                                            int j_end) {
  static_assert(NCH == 128 || NCH == 64, "segment width");
  using GS = GatherShape<NCH>;
  PROF(e, kPOther);
  if (g.debug & 1) return;
  const int lane = threadIdx.x & 31;
  const int sub = lane / GS::kLanesPerRow;        // which row of the group this lane works on
  const int piece = lane % GS::kLanesPerRow;      // which 8 B piece of the half-row
  const int row0 = e.q * 32 + sub;                // + kRowsPerInstr * group
  const uint8_t* mp = reinterpret_cast<const uint8_t*>(g.map + ch0) + e.half * NCH + piece * 8;
  // staging: 16 B chunk index inside the row's 256 B.  Half h always stays inside chunks [8h, 8h+8) -- also for a
  // 64-channel segment -- so that the XOR swizzle (row & 7) never moves one warp's data into the chunks the other
  // warp of the row quarter reads or writes (the two warps are not synchronised with each other)
  const int chunk = e.half * 8 + (piece >> 1);
  const int chunk_sub = (piece & 1) * 8;
  // U row groups x 4 taps of 8 B per lane are in flight per batch; the tap weights are fetched only when a row is
  // blended, which keeps a batch at 8 B x 4 x U registers
  constexpr int U = NJF_GATHER_U;
#pragma unroll 1
  for (int j0 = j_begin; j0 < j_end; j0 += U) {
    uint2 t[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (j0 + u < j_end) {
        const TapEntry* te = taps + row0 + GS::kRowsPerInstr * (j0 + u);
        const uint4 px = *reinterpret_cast<const uint4*>(te->off);
        t[u][0] = __ldg(reinterpret_cast<const uint2*>(mp + px.x));
        t[u][1] = __ldg(reinterpret_cast<const uint2*>(mp + px.y));
        t[u][2] = __ldg(reinterpret_cast<const uint2*>(mp + px.z));
        t[u][3] = __ldg(reinterpret_cast<const uint2*>(mp + px.w));
      } else {
        t[u][0] = t[u][1] = t[u][2] = t[u][3] = make_uint2(0u, 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (j0 + u >= j_end) break;
      const int row = row0 + GS::kRowsPerInstr * (j0 + u);
      const uint4 w = *reinterpret_cast<const uint4*>(taps[row].w2);
      const uint32_t wq[4] = {w.x, w.y, w.z, w.w};
      __half2 lo = __hmul2(*reinterpret_cast<const __half2*>(&t[u][0].x), *reinterpret_cast<const __half2*>(&wq[0]));
      __half2 hi = __hmul2(*reinterpret_cast<const __half2*>(&t[u][0].y), *reinterpret_cast<const __half2*>(&wq[0]));
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        lo = __hfma2(*reinterpret_cast<const __half2*>(&t[u][q].x), *reinterpret_cast<const __half2*>(&wq[q]), lo);
        hi = __hfma2(*reinterpret_cast<const __half2*>(&t[u][q].y), *reinterpret_cast<const __half2*>(&wq[q]), hi);
      }
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo);
      o.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(e.tz + tz_offset(row, chunk) + chunk_sub) = o;
    }
  }
  PROF(e, kPGather);
}

// the generic form: HB bytes per tap starting `byte_off` bytes into the pixel's channels, staged from 16 B chunk
// `chunk0` of the row on
template <int HB>
__device__ __forceinline__ void gather_bytes(EpiCtx& e, const PassGeom& g, const TapEntry* taps, int byte_off, int chunk0,
                                             int j_begin, int j_end) {
  static_assert(HB == 128 || HB == 64, "bytes per tap");
  using GS = GatherShape<HB>;  // HB bytes per warp and tap == the warp's half of an HB-channel segment
  PROF(e, kPOther);
  if (g.debug & 1) return;
  const int lane = threadIdx.x & 31;
  const int sub = lane / GS::kLanesPerRow;        // which row of the group this lane works on
  const int piece = lane % GS::kLanesPerRow;      // which 8 B piece of the row's HB bytes
  const int row0 = e.q * 32 + sub;                // + kRowsPerInstr * group
  const uint8_t* mp = reinterpret_cast<const uint8_t*>(g.map) + byte_off + piece * 8;
  const int chunk = chunk0 + (piece >> 1);
  const int chunk_sub = (piece & 1) * 8;
  // U row groups x 4 taps of 8 B per lane are in flight per batch; the tap weights are fetched only when a row is
  // blended, which keeps a batch at 8 B x 4 x U registers
  constexpr int U = NJF_GATHER_U;
#pragma unroll 1
  for (int j0 = j_begin; j0 < j_end; j0 += U) {
    uint2 t[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (j0 + u < j_end) {
        const TapEntry* te = taps + row0 + GS::kRowsPerInstr * (j0 + u);
        const uint4 px = *reinterpret_cast<const uint4*>(te->off);
        t[u][0] = __ldg(reinterpret_cast<const uint2*>(mp + px.x));
        t[u][1] = __ldg(reinterpret_cast<const uint2*>(mp + px.y));
        t[u][2] = __ldg(reinterpret_cast<const uint2*>(mp + px.z));
        t[u][3] = __ldg(reinterpret_cast<const uint2*>(mp + px.w));
      } else {
        t[u][0] = t[u][1] = t[u][2] = t[u][3] = make_uint2(0u, 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (j0 + u >= j_end) break;
      const int row = row0 + GS::kRowsPerInstr * (j0 + u);
      const uint4 w = *reinterpret_cast<const uint4*>(taps[row].w2);
      const uint32_t wq[4] = {w.x, w.y, w.z, w.w};
      __half2 lo = __hmul2(*reinterpret_cast<const __half2*>(&t[u][0].x), *reinterpret_cast<const __half2*>(&wq[0]));
      __half2 hi = __hmul2(*reinterpret_cast<const __half2*>(&t[u][0].y), *reinterpret_cast<const __half2*>(&wq[0]));
#pragma unroll
      for (int q = 1; q < 4; ++q) {
        lo = __hfma2(*reinterpret_cast<const __half2*>(&t[u][q].x), *reinterpret_cast<const __half2*>(&wq[q]), lo);
        hi = __hfma2(*reinterpret_cast<const __half2*>(&t[u][q].y), *reinterpret_cast<const __half2*>(&wq[q]), hi);
      }
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&lo);
      o.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(e.tz + tz_offset(row, chunk) + chunk_sub) = o;
    }
  }
  PROF(e, kPGather);
}
// a 128-channel segment in two 32-channel-per-warp parts: part 1 (the upper 32 channels of the warp's half) lands in
// chunks [8h+4, 8h+8) and can be gathered while a 64-channel segment still occupies [8h, 8h+4)
__device__ __forceinline__ void gather_seg128_part(EpiCtx& e, const PassGeom& g, const TapEntry* taps, int ch0, int part) {
  gather_bytes<64>(e, g, taps, (ch0 + e.half * 64 + part * 32) * 2, e.half * 8 + part * 4, 0, GatherShape<64>::kGroups);
}

// a whole segment at once; __syncwarp orders it against this warp's own reads of the staging buffer
template <int NCH>
__device__ __forceinline__ void gather_segment(EpiCtx& e, const PassGeom& g, const TapEntry* taps, int ch0) {
  __syncwarp();
  gather_rows<NCH>(e, g, taps, ch0, 0, GatherShape<NCH>::kGroups);
  __syncwarp();
}

// ----------------------------------------------------------------------------- trunk epilogues
// The 10 residual-block steps + lin_out of one ResnetFC trunk, epilogue side.  Pre-conditions:
// lin_in's accumulator wait has completed (x = W_in . [enc | xyz] + b_in in TMEM) and segment 0 of this
// trunk's hoisted channels is in the staging buffer.  Ends with lin_out issued: after the caller's
// epi_wait_acc its accumulator (bias included) is in TMEM columns [128, 128+n_out).
// `hook(k, w)` is called in the MMA wait windows of blocks 2..4 (w = 0: fc_0 is running, w = 1: fc_1 is running),
// which carry no gather of this tile: the kernels use them to set up the NEXT tile of the slot and to prefetch its
// first hoisted segment, so that neither sits on the next tile's critical path.
struct NoHook {
  __device__ __forceinline__ void operator()(int, int) const {}
};
template <class Hook = NoHook>
__device__ __forceinline__ void trunk_blocks_epilogue(EpiCtx& e, const PassGeom& g, int seg_ch0, const TapEntry* taps,
                                                      Hook hook = Hook()) {
  // E0: X_0 = lin_in(enc, xyz) + b_in + tz_0
  for (int c0 = e.col0; c0 < e.col0 + 64; c0 += 32) epi_x_update<true>(e, c0);
  epi_publish(e);  // -> fc_0 (block 0)
#pragma unroll 1
  for (int k = 0; k < 5; ++k) {
    // the next hoisted segment is gathered in two halves: one while the tensor pipe runs fc_0, one while it
    // runs fc_1 -- each half's load latency hides behind one MMA round trip
    if (k < 2) {
      __syncwarp();  // this warp has consumed segment k (the x update above)
      gather_rows<128>(e, g, taps, seg_ch0 + 128 * (k + 1), 0, 8);
    } else {
      hook(k, 0);
    }
    PROF(e, kPEpi);
    epi_wait_acc(e);
    epi_relu_to_a64(e, 128 + e.col0, e.col0);
    epi_publish(e);  // -> fc_1 (block k), accumulates onto x
    if (k < 2) {
      gather_rows<128>(e, g, taps, seg_ch0 + 128 * (k + 1), 8, 16);
      __syncwarp();  // segment k+1 complete: written and read by this warp only
    } else {
      hook(k, 1);
    }
    PROF(e, kPEpi);
    epi_wait_acc(e);
    if (k < 2) {
      for (int c0 = e.col0; c0 < e.col0 + 64; c0 += 32) epi_x_update<true>(e, c0);
    } else {
      epi_relu_to_a64(e, e.col0, e.col0);  // no hoisted segment after block 2: x -> ReLU -> A tile
    }
    epi_publish(e);  // -> fc_0 (block k+1) or lin_out
  }
  PROF(e, kPEpi);
  // the caller waits for the lin_out accumulator (it may issue independent loads first)
}

// ----------------------------------------------------------------------------- per-ray scans (one warp)
// exclusive prefix sums of dd[0..n) accumulated in double (like torch.cumsum on CPU), rounded to
// fp32 per element; `carry` holds the running total across the tiles of a long ray.
__device__ __forceinline__ void excl_scan_warp(const float* dd, int n, double& carry, float* cum) {
  const int lane = threadIdx.x & 31;
  const int C = (n + 31) >> 5;
  const int j0 = min(lane * C, n), j1 = min(j0 + C, n);
  double loc = 0.0;
  for (int j = j0; j < j1; ++j) loc += static_cast<double>(dd[j]);
  double inc = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  double run = carry + (inc - loc);
  for (int j = j0; j < j1; ++j) {
    cum[j] = static_cast<float>(run);
    run += static_cast<double>(dd[j]);
  }
  carry += __shfl_sync(0xffffffffu, inc, 31);
}

// PDFSampler.generate_ray_samples (ray_samplers.py:351-451), eval or train (u supplied), for ONE
// ray by ONE warp.  `w` (smem, S entries) holds the transmittance weights and is overwritten;
// `cdf` is smem scratch of S+1 entries.
__device__ __forceinline__ void pdf_resample_warp(float* w, int S, const float* __restrict__ bins_in,
                                                  const float* __restrict__ u, int nb, float anneal, int sum_vec,
                                                  float* cdf, float* __restrict__ bins_out,
                                                  int32_t* __restrict__ inds_out) {
  const int lane = threadIdx.x & 31;
  const int C = (S + 31) >> 5;
  const int j0 = min(lane * C, S), j1 = min(j0 + C, S);
  for (int j = j0; j < j1; ++j) {
    float x = w[j];
    if (anneal != 1.0f) x = powf(x, anneal);
    w[j] = __fadd_rn(x, 0.01f);  // histogram_padding
  }
  __syncwarp();
  float wsum;
  if (sum_vec == 8) {
    // ATen's vectorised fp32 row sum on CPU: 8 lanes x 4 interleaved accumulators, then the lanes
    // are added in order (SumKernel.cpp vectorized_inner_sum / row_sum); verified against
    // torch.sum in tests/test_oracle_golden.py.
    const int nvec = S >> 3, nilp = nvec >> 2;
    float r = 0.f;
    if (lane < 8) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int i = 0; i < nilp; ++i) {
        a0 = __fadd_rn(a0, w[(4 * i + 0) * 8 + lane]);
        a1 = __fadd_rn(a1, w[(4 * i + 1) * 8 + lane]);
        a2 = __fadd_rn(a2, w[(4 * i + 2) * 8 + lane]);
        a3 = __fadd_rn(a3, w[(4 * i + 3) * 8 + lane]);
      }
      for (int i = 4 * nilp; i < nvec; ++i) a0 = __fadd_rn(a0, w[i * 8 + lane]);
      r = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
    }
    float fin = 0.f;
    for (int k = nvec * 8; k < S; ++k) fin = __fadd_rn(fin, w[k]);
    for (int v = 0; v < 8; ++v) fin = __fadd_rn(fin, __shfl_sync(0xffffffffu, r, v));
    wsum = fin;
  } else {
    double loc = 0.0;
    for (int j = j0; j < j1; ++j) loc += static_cast<double>(w[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
    wsum = static_cast<float>(loc);
  }
  const float pad = fmaxf(__fsub_rn(1e-5f, wsum), 0.f);
  const float padj = __fdiv_rn(pad, static_cast<float>(S));
  wsum = __fadd_rn(wsum, pad);
  // cdf = [0, min(1, cumsum(pdf))]
  double loc = 0.0;
  for (int j = j0; j < j1; ++j) {
    const float pj = __fdiv_rn(__fadd_rn(w[j], padj), wsum);
    w[j] = pj;
    loc += static_cast<double>(pj);
  }
  double inc = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  double run = inc - loc;
  for (int j = j0; j < j1; ++j) {
    run += static_cast<double>(w[j]);
    cdf[j + 1] = fminf(1.0f, static_cast<float>(run));
  }
  if (lane == 0) cdf[0] = 0.f;
  __syncwarp();
  for (int i = lane; i < nb; i += 32) {
    const float uu = __ldg(u + i);
    int lo = 0, hi = S + 1;  // searchsorted(cdf, u, right=True)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= uu) lo = mid + 1; else hi = mid;
    }
    const int below = min(max(lo - 1, 0), S), above = min(max(lo, 0), S);
    const float c0 = cdf[below], c1 = cdf[above];
    const float b0 = __ldg(bins_in + below), b1 = __ldg(bins_in + above);
    float t = __fdiv_rn(__fsub_rn(uu, c0), __fsub_rn(c1, c0));
    if (t != t) t = 0.f;                       // nan_to_num(nan=0); +-inf are removed by the clip
    t = fminf(fmaxf(t, 0.f), 1.f);
    bins_out[i] = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
    if (inds_out) inds_out[i] = lo;
  }
  __syncwarp();
}

// ordered-uint encoding of a float for atomicMin/atomicMax
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// SH degree 4 of the unit direction (action_decoder_jacobian.py:194-199): the colour head's directional encoding.
//  conv = NJF_SH_TCNN            : tiny-cuda-nn's SphericalHarmonics (input d01 re-mapped x*2-1, tcnn's signs)
//  conv = NJF_SH_NERFSTUDIO_TORCH: nerfstudio's torch fallback (components_from_spherical_harmonics evaluated on
//                                  the [0,1] input as passed, all-positive leading signs)
__device__ __forceinline__ void sh16(float dx, float dy, float dz, int conv, float (&o)[16]) {
  // get_normalized_directions: (d + 1) / 2
  const float hx = __fmul_rn(__fadd_rn(dx, 1.f), 0.5f), hy = __fmul_rn(__fadd_rn(dy, 1.f), 0.5f),
              hz = __fmul_rn(__fadd_rn(dz, 1.f), 0.5f);
  const bool tc = conv == NJF_SH_TCNN;
  const float x = tc ? __fsub_rn(__fmul_rn(hx, 2.f), 1.f) : hx;
  const float y = tc ? __fsub_rn(__fmul_rn(hy, 2.f), 1.f) : hy;
  const float z = tc ? __fsub_rn(__fmul_rn(hz, 2.f), 1.f) : hz;
  const float sg = tc ? -1.f : 1.f;  // tcnn flips the sign of the odd-in-(x,y) terms below
  const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  o[0] = 0.28209479177387814f;
  o[1] = sg * 0.48860251190291987f * y;
  o[2] = 0.48860251190291987f * z;
  o[3] = sg * 0.48860251190291987f * x;
  o[4] = 1.0925484305920792f * xy;
  o[5] = sg * 1.0925484305920792f * yz;
  o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = sg * 1.0925484305920792f * xz;
  o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  o[9] = sg * 0.59004358992664352f * y * (3.0f * x2 - y2);
  o[10] = 2.8906114426405538f * xy * z;
  o[11] = sg * 0.45704579946446572f * y * (5.0f * z2 - 1.0f);
  o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  o[13] = sg * 0.45704579946446572f * x * (5.0f * z2 - 1.0f);
  o[14] = 1.4453057213202769f * z * (x2 - y2);
  o[15] = sg * 0.59004358992664352f * x * (x2 - 3.0f * y2);
}

}  // namespace njf
