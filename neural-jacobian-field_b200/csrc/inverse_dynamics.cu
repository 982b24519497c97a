// gn_terms_kernel: Gauss-Newton / Levenberg-Marquardt normal equations of the inverse-dynamics problem of
// notebooks/real_world/2_inverse_dynamics.ipynb (Adam loop over Model.infer_optical_flow, models/model.py:497-525)
// on the collapsed encoding (J-bar, p):
//     flow_i(u) = proj(W (p_i + Jbar_i^T u) + t) - proj(W p_i + t),   proj(c) = (K c)_{0,1} / ((K c)_2 + 1e-9)
//     r_i = flow_i(u) - target_i ;  G_i = d flow_i / d u  (2 x A, analytic)
//     H = sum_i w_i G_i^T G_i   (A x A) ;  g = sum_i w_i G_i^T r_i  (A) ;  loss = sum_i w_i |r_i|^2
// One thread per ray, fp64 accumulation, warp-shuffle + shared-memory block reduction, one atomicAdd per block
// and output element.  Per view (rays_per_view rays share one action row).
#include "njf_internal.h"
#include "../../include/njf_b200.h"

namespace njf {

constexpr int kGnMaxA = 10;
constexpr int kGnTerms = kGnMaxA * (kGnMaxA + 1) / 2 + kGnMaxA + 1;  // upper triangle + gradient + loss

struct GnParams {
  const float* jbar; const float* p; const float* action; const float* w2c; const float* kpx;
  const float* target; const float* weight;
  int N, R, A;
  double* out;  // [B][kGnTerms]
};

__global__ void __launch_bounds__(128) gn_terms_kernel(const GnParams q) {
  __shared__ double red[4][kGnTerms];
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // all rays of a block belong to one view (the launcher pads the grid per view)
  const int b = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = r < q.R;
  const int i = b * q.R + (live ? r : 0);
  (void)ray;
  const int A = q.A;
  float G[2][kGnMaxA];
  float res[2] = {0.f, 0.f};
  float wgt = 0.f;
#pragma unroll
  for (int a = 0; a < kGnMaxA; ++a) G[0][a] = G[1][a] = 0.f;
  if (live) {
    wgt = q.weight ? __ldg(q.weight + i) : 1.f;
    const float* W = q.w2c + b * 16;
    const float* K = q.kpx + b * 9;
    const float px = q.p[static_cast<size_t>(i) * 3], py = q.p[static_cast<size_t>(i) * 3 + 1], pz = q.p[static_cast<size_t>(i) * 3 + 2];
    const float* J = q.jbar + static_cast<size_t>(i) * 3 * A;
    float f[3] = {0.f, 0.f, 0.f};
    for (int a = 0; a < A; ++a) {
      const float ua = __ldg(q.action + b * A + a);
#pragma unroll
      for (int d = 0; d < 3; ++d) f[d] = fmaf(J[a * 3 + d], ua, f[d]);
    }
    auto cam = [&](float x, float y, float z, float (&c)[3]) {
#pragma unroll
      for (int k = 0; k < 3; ++k) c[k] = fmaf(W[4 * k + 2], z, fmaf(W[4 * k + 1], y, fmaf(W[4 * k], x, W[4 * k + 3])));
    };
    auto proj = [&](const float (&c)[3], float& u, float& v, float& zd) {
      const float a0 = fmaf(K[2], c[2], fmaf(K[1], c[1], K[0] * c[0]));
      const float a1 = fmaf(K[5], c[2], fmaf(K[4], c[1], K[3] * c[0]));
      const float a2 = fmaf(K[8], c[2], fmaf(K[7], c[1], K[6] * c[0]));
      zd = a2 + 1e-9f;
      u = a0 / zd;
      v = a1 / zd;
    };
    float c0[3], c1[3], u0, v0, z0, u1, v1, z1;
    cam(px, py, pz, c0);
    cam(px + f[0], py + f[1], pz + f[2], c1);
    proj(c0, u0, v0, z0);
    proj(c1, u1, v1, z1);
    res[0] = (u1 - u0) - q.target[static_cast<size_t>(i) * 2];
    res[1] = (v1 - v0) - q.target[static_cast<size_t>(i) * 2 + 1];
    // d(u,v)/dc at c1, then dc/dx = W[:3,:3], dx/du_a = Jbar[a]
    float du[3], dv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      du[k] = (K[k] - u1 * K[6 + k]) / z1;
      dv[k] = (K[3 + k] - v1 * K[6 + k]) / z1;
    }
    float dux[3], dvx[3];  // rows of d(u,v)/dx
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      dux[d] = du[0] * W[d] + du[1] * W[4 + d] + du[2] * W[8 + d];
      dvx[d] = dv[0] * W[d] + dv[1] * W[4 + d] + dv[2] * W[8 + d];
    }
    for (int a = 0; a < A; ++a) {
      G[0][a] = dux[0] * J[a * 3] + dux[1] * J[a * 3 + 1] + dux[2] * J[a * 3 + 2];
      G[1][a] = dvx[0] * J[a * 3] + dvx[1] * J[a * 3 + 1] + dvx[2] * J[a * 3 + 2];
    }
  }
  // per-thread terms -> warp reduction (fp64) -> block -> global
  int t = 0;
  const double w = static_cast<double>(wgt);
  auto reduce_store = [&](double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][t] = v;
    ++t;
  };
#pragma unroll
  for (int a = 0; a < kGnMaxA; ++a)
#pragma unroll
    for (int c = a; c < kGnMaxA; ++c)
      reduce_store(w * (static_cast<double>(G[0][a]) * G[0][c] + static_cast<double>(G[1][a]) * G[1][c]));
#pragma unroll
  for (int a = 0; a < kGnMaxA; ++a)
    reduce_store(w * (static_cast<double>(G[0][a]) * res[0] + static_cast<double>(G[1][a]) * res[1]));
  reduce_store(w * (static_cast<double>(res[0]) * res[0] + static_cast<double>(res[1]) * res[1]));
  __syncthreads();
  for (int k = threadIdx.x; k < kGnTerms; k += blockDim.x) {
    const double v = red[0][k] + red[1][k] + red[2][k] + red[3][k];
    if (v != 0.0) atomicAdd(q.out + static_cast<size_t>(b) * kGnTerms + k, v);
  }
}

// scratch terms [B][kGnTerms] -> H [B][A][A] (symmetric), g [B][A], loss [B]
__global__ void gn_unpack_kernel(const double* terms, int B, int A, double* H, double* g, double* loss) {
  const int b = blockIdx.x;
  if (b >= B) return;
  const double* t = terms + static_cast<size_t>(b) * kGnTerms;
  for (int k = threadIdx.x; k < kGnMaxA * kGnMaxA; k += blockDim.x) {
    const int a = k / kGnMaxA, c = k % kGnMaxA;
    if (a >= A || c >= A) continue;
    const int lo = a < c ? a : c, hi = a < c ? c : a;
    const int idx = lo * kGnMaxA - lo * (lo - 1) / 2 + (hi - lo);
    H[(static_cast<size_t>(b) * A + a) * A + c] = t[idx];
  }
  const int g0 = kGnMaxA * (kGnMaxA + 1) / 2;
  for (int a = threadIdx.x; a < A; a += blockDim.x) g[static_cast<size_t>(b) * A + a] = t[g0 + a];
  if (threadIdx.x == 0) loss[b] = t[g0 + kGnMaxA];
}

}  // namespace njf

extern "C" int njf_flow_gn_terms(const float* jbar, const float* p, const float* action, const float* trgt_w2c,
                                 const float* trgt_k_px, const float* target_flow, const float* ray_weight, int n_rays,
                                 int rays_per_view, int action_dim, double* workspace, double* H, double* g,
                                 double* loss, void* stream_) {
  using namespace njf;
  if (!jbar || !p || !action || !trgt_w2c || !trgt_k_px || !target_flow || !workspace || !H || !g || !loss)
    NJF_FAIL("njf_flow_gn_terms: null argument");
  if (action_dim < 1 || action_dim > kGnMaxA) NJF_FAIL("njf_flow_gn_terms: action_dim %d unsupported (1..%d)", action_dim, kGnMaxA);
  if (n_rays < 1 || rays_per_view < 1 || n_rays % rays_per_view) NJF_FAIL("njf_flow_gn_terms: n_rays %d / rays_per_view %d", n_rays, rays_per_view);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int B = n_rays / rays_per_view;
  NJF_CUDA(cudaMemsetAsync(workspace, 0, static_cast<size_t>(B) * kGnTerms * sizeof(double), stream));
  GnParams q{jbar, p, action, trgt_w2c, trgt_k_px, target_flow, ray_weight, n_rays, rays_per_view, action_dim, workspace};
  dim3 grid((rays_per_view + 127) / 128, B);
  gn_terms_kernel<<<grid, 128, 0, stream>>>(q);
  njf::count_launch();
  gn_unpack_kernel<<<B, 128, 0, stream>>>(workspace, B, action_dim, H, g, loss);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int njf_flow_gn_workspace_doubles(int n_views) { return n_views * njf::kGnTerms; }
