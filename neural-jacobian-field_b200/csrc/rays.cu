// rays_kernel: the ray bundle of a view (rendering/geometry.py:117-134 get_pixel_coordinates, :170-203
// get_world_rays_with_z; models/model.py:215-226 compute_ray_bundle): one thread per ray,
//   d_cam = K^-1 [x, y, 1];  d_cam /= |d_cam|;  z = d_cam.z;  d_world = R d_cam;  origin = t
// Pixel coordinates are either supplied (normalised xy per ray) or generated for an H x W grid in the
// reference's order (row-major, x fastest): x = (col + 0.5) / W, y = (row + 0.5) / H.
#include "njf_internal.h"
#include "../../include/njf_b200.h"

namespace njf {

struct RaysParams {
  const float* k_norm;   // [B][9]
  const float* c2w;      // [B][16]
  const float* coords;   // [B][R][2] or null (grid mode)
  int B, R, H, W;
  float* origins;        // [B][R][3]
  float* dirs;           // [B][R][3]
  float* z;              // [B][R] or null
};

__global__ void rays_kernel(const RaysParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.B * p.R) return;
  const int b = i / p.R, r = i - b * p.R;
  // inverse intrinsics by the adjugate in double (the reference calls torch.inverse on the fp32 matrix)
  double k[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) k[j] = static_cast<double>(__ldg(p.k_norm + b * 9 + j));
  const double c00 = k[4] * k[8] - k[5] * k[7], c01 = k[5] * k[6] - k[3] * k[8], c02 = k[3] * k[7] - k[4] * k[6];
  const double det = k[0] * c00 + k[1] * c01 + k[2] * c02;
  const double id = 1.0 / det;
  float inv[9];
  inv[0] = static_cast<float>(c00 * id);
  inv[1] = static_cast<float>((k[2] * k[7] - k[1] * k[8]) * id);
  inv[2] = static_cast<float>((k[1] * k[5] - k[2] * k[4]) * id);
  inv[3] = static_cast<float>(c01 * id);
  inv[4] = static_cast<float>((k[0] * k[8] - k[2] * k[6]) * id);
  inv[5] = static_cast<float>((k[2] * k[3] - k[0] * k[5]) * id);
  inv[6] = static_cast<float>(c02 * id);
  inv[7] = static_cast<float>((k[1] * k[6] - k[0] * k[7]) * id);
  inv[8] = static_cast<float>((k[0] * k[4] - k[1] * k[3]) * id);
  float x, y;
  if (p.coords) {
    x = __ldg(p.coords + static_cast<size_t>(i) * 2);
    y = __ldg(p.coords + static_cast<size_t>(i) * 2 + 1);
  } else {
    const int row = r / p.W, col = r - row * p.W;
    x = __fdiv_rn(__fadd_rn(static_cast<float>(col), 0.5f), static_cast<float>(p.W));
    y = __fdiv_rn(__fadd_rn(static_cast<float>(row), 0.5f), static_cast<float>(p.H));
  }
  float d[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) d[a] = fmaf(inv[3 * a + 1], y, fmaf(inv[3 * a], x, inv[3 * a + 2]));
  const float n = sqrtf(fmaf(d[2], d[2], fmaf(d[1], d[1], d[0] * d[0])));
#pragma unroll
  for (int a = 0; a < 3; ++a) d[a] = __fdiv_rn(d[a], n);
  const float* M = p.c2w + b * 16;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    p.dirs[static_cast<size_t>(i) * 3 + a] = fmaf(__ldg(M + 4 * a + 2), d[2], fmaf(__ldg(M + 4 * a + 1), d[1], __ldg(M + 4 * a) * d[0]));
    p.origins[static_cast<size_t>(i) * 3 + a] = __ldg(M + 4 * a + 3);
  }
  if (p.z) p.z[i] = d[2];
}

}  // namespace njf

extern "C" int njf_make_rays(const float* k_norm, const float* c2w, const float* coords_xy, int B, int R, int H, int W,
                             float* origins, float* dirs, float* z, void* stream_) {
  using namespace njf;
  if (!k_norm || !c2w || !origins || !dirs) NJF_FAIL("njf_make_rays: null argument");
  if (B < 1 || R < 1) NJF_FAIL("njf_make_rays: B=%d R=%d", B, R);
  if (!coords_xy && (H < 1 || W < 1 || static_cast<long long>(H) * W != R))
    NJF_FAIL("njf_make_rays: grid mode needs R == H*W (R=%d, H=%d, W=%d)", R, H, W);
  RaysParams p{k_norm, c2w, coords_xy, B, R, H, W, origins, dirs, z};
  const long long n = static_cast<long long>(B) * R;
  if (n > 0x7fffffffLL) NJF_FAIL("njf_make_rays: too many rays");
  rays_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(p);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}

// ---- inverse of 4x4 poses: Gauss-Jordan with partial pivoting in fp64, one thread per matrix, rounded to fp32
// (rendering/geometry.py:59-65 transform_world2cam and :206-215 call torch.inverse on the fp32 pose)
namespace njf {
__global__ void invert_poses_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      a[r][c] = static_cast<double>(in[i * 16 + r * 4 + c]);
      a[r][4 + c] = (r == c) ? 1.0 : 0.0;
    }
#pragma unroll
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    double best = fabs(a[col][col]);
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r > col && fabs(a[r][col]) > best) {
        best = fabs(a[r][col]);
        piv = r;
      }
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r == piv && piv != col) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const double t = a[col][c];
          a[col][c] = a[r][c];
          a[r][c] = t;
        }
      }
    const double inv = 1.0 / a[col][col];  // singular pose -> inf/nan, like torch.inverse raising / returning garbage
#pragma unroll
    for (int c = 0; c < 8; ++c) a[col][c] *= inv;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (r != col) {
        const double fct = a[r][col];
#pragma unroll
        for (int c = 0; c < 8; ++c) a[r][c] -= fct * a[col][c];
      }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) out[i * 16 + r * 4 + c] = static_cast<float>(a[r][4 + c]);
}
}  // namespace njf

extern "C" int njf_invert_poses(const float* c2w, float* w2c, int n, void* stream_) {
  using namespace njf;
  if (!c2w || !w2c) NJF_FAIL("njf_invert_poses: null argument");
  if (n < 1) return 0;
  invert_poses_kernel<<<(n + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream_)>>>(c2w, w2c, n);
  njf::count_launch();
  NJF_CUDA(cudaGetLastError());
  return 0;
}
