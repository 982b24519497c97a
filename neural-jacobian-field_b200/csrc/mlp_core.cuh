// Warp-specialised tcgen05 "layer chain" machinery shared by the render kernels
// and the self-test.
//
// One CTA = 576 threads:
//   warps 0-7   : epilogue group of tile slot 0 -- TWO threads per sample point (TMEM lane):
//                 warp = 4*h + q handles rows 32q..32q+31, column half h (64 of every 128 columns)
//   warps 8-15  : epilogue group of tile slot 1
//   warp  16    : MMA issuer (one elected thread issues every tcgen05.mma / commit)
//   warp  17    : weight loader (one elected thread streams packed layer images with
//                 cp.async.bulk into a 2-stage shared-memory ring)
// Both slots run the SAME step program, so one weight stage feeds two 128-row tiles
// and the tensor pipe works on one slot while the other slot's warpgroup runs its
// epilogue (ping-pong).
//
// Per slot TMEM: 256 fp32 columns (x: [0,128) residual stream, net: [128,256)).
#pragma once
#include "ptx.cuh"

namespace njf {

constexpr int kRows = 128;
constexpr int kSlots = 2;
constexpr int kSlotThreads = 256;   // epilogue threads per slot (2 per row)
constexpr int kIssuerWarp = 16;
constexpr int kLoaderWarp = 17;
constexpr int kThreads = 576;
constexpr int kStages = 2;
constexpr uint32_t kBiasBlkBytes = 4096; // N<=128 rows x 16 cols fp16 (SW32): column 0 = bias
constexpr uint32_t kStageBytes = 32768 + kBiasBlkBytes;  // up to N=128 x K=128 fp16 + bias block
constexpr uint32_t kATileBytes = 32768;  // 128 rows x 128 cols fp16 (2 K-blocks of 16 KB)
constexpr uint32_t kAKbStride = kRows * 128;  // bytes between K-blocks of an A tile
constexpr uint32_t kTzBytes = 32768;     // 128 rows x 128 ch fp16 (256 B rows, chunk-swizzled)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSlotCols = 256;
constexpr int kMaxSteps = 48;

struct MmaStep {
  uint32_t w_off;    // byte offset of the layer image inside the packed blob (16 B aligned)
  uint32_t w_bytes;  // n * kblocks * 128
  uint16_t n;        // MMA N (multiple of 16, <= 128)
  uint8_t kblocks;   // K / 64 (1 or 2)
  uint8_t acc;       // 1: accumulate onto d_col
  uint16_t d_col;    // TMEM column offset inside the slot
  uint16_t flags;    // kStepReuseA: same A tile as the previous step (no a_ready wait);
                     // kStepNoCommit: the NEXT step's commit also covers this accumulator
};
constexpr uint16_t kStepReuseA = 1, kStepNoCommit = 2;
// kStepBias: the image is followed by an [N x 16] SW32 block whose column 0 is the layer bias; one
// extra MMA multiplies it with the CTA's constant "ones" A block (column 0 = 1), so the bias is
// accumulated by the tensor core and the epilogues carry no bias loads / adds.
constexpr uint16_t kStepBias = 4;
// kStepK16Tail: of the LAST K-block only the first 16 columns are multiplied (one MMA instead of four):
// lin_in / query-MLP steps carry the raw-xyz input there as hi/lo fp16 pairs (render.cuh write_posenc)
constexpr uint16_t kStepK16Tail = 8;
struct Program {
  int nsteps;
  MmaStep steps[kMaxSteps];
};

// dynamic shared memory map (base aligned to 1024 B)
struct SmemMap {
  static constexpr uint32_t kA = 0;                                 // 2 x 32 KB
  static constexpr uint32_t kW = kA + kSlots * kATileBytes;         // 2 x 32 KB
  static constexpr uint32_t kTz = kW + kStages * kStageBytes;       // 2 x 32 KB
  static constexpr uint32_t kOnes = kTz + kSlots * kTzBytes;        // 4 KB constant "ones" A block
  static constexpr uint32_t kMisc = kOnes + kBiasBlkBytes;          // barriers etc.
  static constexpr uint32_t kMiscBytes = 1024;
  static constexpr uint32_t kScratch = kMisc + kMiscBytes;          // kernel-specific
};
struct Barriers {
  uint64_t a_ready[kSlots];    // 256 arrivals: slot's A tile written (+ TMEM reads drained)
  uint64_t acc_ready[kSlots];  // tcgen05.commit: slot's accumulator complete
  uint64_t w_full[kStages];    // bulk-copy transaction bytes landed
  uint64_t w_empty[kStages];   // tcgen05.commit: both slots' MMAs done with the stage
  uint32_t tmem_base;
};

struct CtaCtx {
  uint8_t* smem;
  Barriers* bars;
  uint32_t tmem_base;
};

// all threads of the CTA
__device__ __forceinline__ CtaCtx cta_setup(uint8_t* smem_raw) {
  CtaCtx c;
  // align inside the shared window with pointer arithmetic only (keeps the shared address space
  // visible to the compiler: LDS/STS instead of generic LD/ST)
  c.smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  c.bars = reinterpret_cast<Barriers*>(c.smem + SmemMap::kMisc);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&c.bars->a_ready[s], kSlotThreads);
      mbar_init(&c.bars->acc_ready[s], 1);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&c.bars->w_full[s], 1);
      mbar_init(&c.bars->w_empty[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == kIssuerWarp) tmem_alloc(&c.bars->tmem_base, kTmemCols);
  if (threadIdx.x < kRows) {  // ones block: element (row, 0) = 1.0, everything else 0
    uint4* rowp = reinterpret_cast<uint4*>(c.smem + SmemMap::kOnes + threadIdx.x * 32);
    rowp[0] = make_uint4(0, 0, 0, 0);
    rowp[1] = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint16_t*>(c.smem + SmemMap::kOnes + sw32_offset(threadIdx.x, 0)) = 0x3C00;  // 1.0h
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  c.tmem_base = c.bars->tmem_base;
  return c;
}
__device__ __forceinline__ void cta_teardown(const CtaCtx& c) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == kIssuerWarp) tmem_dealloc(c.tmem_base, kTmemCols);
}

// ----------------------------------------------------------------------------- loader
// `nrun` = how many times the program is executed by this CTA (items x tiles-per-item).
__device__ __forceinline__ void loader_role(const CtaCtx& c, const Program& prog,
                                            const uint8_t* __restrict__ blob, int nrun, int debug = 0) {
  uint32_t cnt = 0;
  for (int run = 0; run < nrun; ++run) {
    for (int s = 0; s < prog.nsteps; ++s, ++cnt) {
      const uint32_t stage = cnt % kStages, par = (cnt / kStages) & 1u;
      mbar_wait_backoff(&c.bars->w_empty[stage], par ^ 1u);
      if (debug & 8) {  // NJF_DEBUG_SKIP bit 3 (timing attribution only): no weight traffic, stale operands
        mbar_arrive(&c.bars->w_full[stage]);
        continue;
      }
      const uint32_t bytes = prog.steps[s].w_bytes;
      mbar_arrive_expect_tx(&c.bars->w_full[stage], bytes);
      bulk_g2s(c.smem + SmemMap::kW + stage * kStageBytes, blob + prog.steps[s].w_off, bytes,
               &c.bars->w_full[stage]);
    }
  }
}

// ----------------------------------------------------------------------------- issuer
// `nact(run)` = number of active slots for that run (slot 1 idles on an odd tail).
template <class NActFn>
__device__ __forceinline__ void issuer_role(const CtaCtx& c, const Program& prog, int nrun,
                                            NActFn nact) {
  uint32_t cnt = 0;
  uint32_t apar[kSlots] = {0, 0};
  const uint32_t a_base = smem_u32(c.smem + SmemMap::kA);
  const uint32_t w_base = smem_u32(c.smem + SmemMap::kW);
  for (int run = 0; run < nrun; ++run) {
    const int na = nact(run);
    for (int s = 0; s < prog.nsteps; ++s, ++cnt) {
      const MmaStep st = prog.steps[s];
      const uint32_t stage = cnt % kStages, par = (cnt / kStages) & 1u;
      const uint32_t idesc = make_idesc_f16(st.n);
      const uint32_t w_kb_stride = static_cast<uint32_t>(st.n) * 128u;
      for (int slot = 0; slot < na; ++slot) {
        if (!(st.flags & kStepReuseA)) {
          mbar_wait(&c.bars->a_ready[slot], apar[slot]);
          apar[slot] ^= 1u;
        }
        if (slot == 0) mbar_wait(&c.bars->w_full[stage], par);
        tc_fence_after();
        const uint32_t d_tmem = c.tmem_base + slot * kSlotCols + st.d_col;
        const uint32_t a0 = a_base + slot * kATileBytes;
        const uint32_t w0 = w_base + stage * kStageBytes;
        uint32_t acc = st.acc;
        for (int kb = 0; kb < st.kblocks; ++kb) {
          const int nk = ((st.flags & kStepK16Tail) && kb == st.kblocks - 1) ? 1 : 4;
          for (int k = 0; k < nk; ++k) {
            umma_f16(d_tmem, make_sw128_desc(a0 + kb * kAKbStride + k * 32),
                     make_sw128_desc(w0 + kb * w_kb_stride + k * 32), idesc, acc);
            acc = 1;
          }
        }
        if (st.flags & kStepBias)
          umma_f16(d_tmem, make_sw32_desc(smem_u32(c.smem + SmemMap::kOnes)),
                   make_sw32_desc(w0 + st.kblocks * w_kb_stride), idesc, 1);
        if (!(st.flags & kStepNoCommit)) umma_commit(&c.bars->acc_ready[slot]);
      }
      umma_commit(&c.bars->w_empty[stage]);
    }
  }
}

// ----------------------------------------------------------------------------- optional phase profiler
// -DNJF_PROFILE: lane 0 of every epilogue warp attributes SM-clock cycles to phases; summed per
// phase into g_prof (read with njf_prof_read).  Zero cost in the normal build.
enum ProfPhase { kPSetup = 0, kPGather, kPWaitAcc, kPEpi, kPWeights, kPPdf, kPHead, kPColor, kPComposite, kPBar, kPOther,
                 kPLn, kPSoftmax, kPGelu, kPRes, kPCount };
#ifdef NJF_PROFILE
__device__ unsigned long long g_prof[16];
#define PROF(e, ph)                                   \
  do {                                                \
    const long long now_ = clock64();                 \
    (e).prof[ph] += static_cast<unsigned>(now_ - (e).pt); \
    (e).pt = now_;                                    \
  } while (0)
#else
#define PROF(e, ph) do { } while (0)
#endif

// ----------------------------------------------------------------------------- epilogue-side helpers
struct EpiCtx {
#ifdef NJF_PROFILE
  long long pt;
  unsigned prof[kPCount];
#endif
  uint8_t* a_tile;    // this slot's A tile (generic pointer)
  uint8_t* tz;        // this slot's 32 KB staging buffer
  uint64_t* a_ready;
  uint64_t* acc_ready;
  uint32_t tmem;      // TMEM address of this thread's lane, column 0 of the slot
  uint32_t acc_par;
  int row;            // 0..127 (== TMEM lane)
  int slot;
  int half;           // 0/1: which 64 of every 128 columns this thread handles
  int col0;           // 64 * half
  int q;              // row quarter (warp % 4)
};
// named barrier ids: 0 = __syncthreads, 1..2 = slot groups (256 threads),
// 3..10 = the two warps (64 threads) that share a row quarter of a slot
__device__ __forceinline__ void slot_bar(const EpiCtx& e) { named_bar_sync(1 + e.slot, kSlotThreads); }
__device__ __forceinline__ void pair_bar(const EpiCtx& e) { named_bar_sync(3 + e.slot * 4 + e.q, 64); }
__device__ __forceinline__ EpiCtx epi_ctx(const CtaCtx& c) {
  EpiCtx e;
  const int warp = threadIdx.x >> 5;
  e.slot = warp >> 3;
  e.half = (warp >> 2) & 1;
  e.col0 = e.half * 64;
  e.q = warp & 3;
  e.row = e.q * 32 + (threadIdx.x & 31);
  e.a_tile = c.smem + SmemMap::kA + e.slot * kATileBytes;
  e.tz = c.smem + SmemMap::kTz + e.slot * kTzBytes;
  e.a_ready = &c.bars->a_ready[e.slot];
  e.acc_ready = &c.bars->acc_ready[e.slot];
  e.tmem = c.tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + e.slot * kSlotCols;
  e.acc_par = 0;
#ifdef NJF_PROFILE
  e.pt = clock64();
  for (int i = 0; i < kPCount; ++i) e.prof[i] = 0;
#endif
  return e;
}
__device__ __forceinline__ void prof_flush(EpiCtx& e) {
#ifdef NJF_PROFILE
  if ((threadIdx.x & 31) == 0)
    for (int i = 0; i < kPCount; ++i) atomicAdd(&g_prof[i], static_cast<unsigned long long>(e.prof[i]));
#else
  (void)e;
#endif
}
// A tile fully written (generic proxy) and all of this thread's TMEM accesses retired
__device__ __forceinline__ void epi_publish(EpiCtx& e) {
  fence_proxy_async_smem();
  tc_fence_before();
  mbar_arrive(e.a_ready);
}
__device__ __forceinline__ void epi_wait_acc(EpiCtx& e) {
  PROF(e, kPOther);
  mbar_wait_backoff(e.acc_ready, e.acc_par);
  e.acc_par ^= 1u;
  tc_fence_after();
  PROF(e, kPWaitAcc);
}
// 32 packed fp16x2 words (64 columns? no: 16 words = 32 columns) -> A tile columns [c0, c0+32)
__device__ __forceinline__ void a_store32(const EpiCtx& e, int c0, const uint32_t (&p)[16]) {
  uint8_t* base = e.a_tile + (c0 >> 6) * kAKbStride + e.row * 128;
  const int ch0 = (c0 & 63) >> 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 v = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
    *reinterpret_cast<uint4*>(base + (((ch0 + j) ^ (e.row & 7)) << 4)) = v;
  }
}
// 8 packed words -> A tile columns [c0, c0+16) (c0 a multiple of 16)
__device__ __forceinline__ void a_store16(const EpiCtx& e, int c0, const uint32_t (&p)[8]) {
  uint8_t* base = e.a_tile + (c0 >> 6) * kAKbStride + e.row * 128;
  const int ch0 = (c0 & 63) >> 3;
#pragma unroll
  for (int j = 0; j < 2; ++j)
    *reinterpret_cast<uint4*>(base + (((ch0 + j) ^ (e.row & 7)) << 4)) = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
}
// zero-fill A tile columns [c0, c0+32)
__device__ __forceinline__ void a_zero32(const EpiCtx& e, int c0) {
  uint8_t* base = e.a_tile + (c0 >> 6) * kAKbStride + e.row * 128;
  const int ch0 = (c0 & 63) >> 3;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(base + (((ch0 + j) ^ (e.row & 7)) << 4)) = make_uint4(0, 0, 0, 0);
}

// staging buffer: row r = 256 B (128 fp16), 16 B chunk c stored at chunk (c ^ (r & 7))
__device__ __forceinline__ uint32_t tz_offset(int row, int chunk) {
  return row * 256 + ((chunk ^ (row & 7)) << 4);
}

// acc[c0..c0+32) of this slot (bias already accumulated by the tensor core) -> ReLU -> fp16 -> A tile.
// `tcol` = TMEM column of c0.
__device__ __forceinline__ void epi_relu_to_a(const EpiCtx& e, int tcol, int c0) {
  uint32_t r[32];
  tmem_ld32(e.tmem + tcol, r);
  tmem_ld_wait();
  uint32_t p[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) p[j] = pack_relu_f16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
  a_store32(e, c0, p);
}

// 64 accumulator columns [tcol, tcol+64) -> ReLU -> fp16 -> A tile columns [c0, c0+64): the second TMEM load is
// issued before the first half is converted and stored, so its latency hides behind that work
__device__ __forceinline__ void epi_relu_to_a64(const EpiCtx& e, int tcol, int c0) {
  uint32_t r0[32], r1[32];
  tmem_ld32(e.tmem + tcol, r0);
  tmem_ld_wait();
  tmem_ld32(e.tmem + tcol + 32, r1);
  {
    uint32_t p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] = pack_relu_f16x2(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]));
    a_store32(e, c0, p);
  }
  tmem_ld_wait();
  {
    uint32_t p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] = pack_relu_f16x2(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
    a_store32(e, c0 + 32, p);
  }
}

// x[c0..c0+32) (+= tz_staging[row][c0..], written back to TMEM so later accumulating MMAs see it)
// -> ReLU -> fp16 -> A tile.  Biases are accumulated by the tensor core (kStepBias).
template <bool kHasTz>
__device__ __forceinline__ void epi_x_update(const EpiCtx& e, int c0) {
  uint32_t r[32];
  tmem_ld32(e.tmem + c0, r);
  tmem_ld_wait();
  if (kHasTz) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 q = *reinterpret_cast<const uint4*>(e.tz + tz_offset(e.row, (c0 >> 3) + j));
      const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 s = fadd2(make_float2(__uint_as_float(r[8 * j + 2 * t]), __uint_as_float(r[8 * j + 2 * t + 1])),
                               __half22float2(h[t]));
        r[8 * j + 2 * t] = __float_as_uint(s.x);
        r[8 * j + 2 * t + 1] = __float_as_uint(s.y);
      }
    }
    tmem_st32(e.tmem + c0, r);
  }
  uint32_t p[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) p[j] = pack_relu_f16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
  a_store32(e, c0, p);
  if (kHasTz) tmem_st_wait();
}

}  // namespace njf
