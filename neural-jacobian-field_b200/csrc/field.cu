// njf_field_create: pack the reference's state-dict tensors into the device-side layout the
// render kernels consume (see include/njf_b200.h).
//
// Layer order of the step programs MUST match the epilogue code in render.cu:
//   proposal : lin_in | (fc_0, fc_1) x5 | lin_out(N=16)
//   field/T  : lin_in | q_enc | (fc_0, fc_1) x5 | lin_out(16) | color1 | color2
//   head/T   : (M1, M2, W1, W2) x3 | jhead(N=32)            (xf_kernel, resident in shared memory)
//   field/M  : lin_in | (fc_0, fc_1) x5 | lin_out(16) | color1 | color2 | lin_in(jac) | (fc_0, fc_1) x5 | lin_out(N=32)
#include <cuda_fp16.h>

#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "field.h"
#include "njf_internal.h"

using namespace njf;

namespace {

struct Builder {
  std::map<std::string, std::pair<const float*, int64_t>> t;
  std::vector<uint8_t> blob;
  std::vector<float> hoist_w, hoist_b;  // rows of [512]
  std::string err;

  const float* get(const std::string& name, int64_t numel) {
    auto it = t.find(name);
    if (it == t.end()) {
      if (err.empty()) err = "missing tensor '" + name + "'";
      return nullptr;
    }
    if (it->second.second != numel) {
      if (err.empty())
        err = "tensor '" + name + "' has " + std::to_string(it->second.second) + " elements, expected " +
              std::to_string(numel);
      return nullptr;
    }
    return it->second.first;
  }
  // append an MMA step whose B image is W[n_real][k_real] (row stride ld)
  // bias != nullptr: append the [n_pad x 16] SW32 bias block (kStepBias)
  void step(Program& p, const float* w, int n_real, int k_real, int ld, int n_pad, int k_pad, int d_col,
            int acc, int flags = 0, const float* bias = nullptr, bool want_bias = false) {
    MmaStep st{};
    if (want_bias) flags |= kStepBias;
    st.flags = static_cast<uint16_t>(flags);
    st.w_off = static_cast<uint32_t>(blob.size());
    st.w_bytes = static_cast<uint32_t>(n_pad * k_pad * 2);
    st.n = static_cast<uint16_t>(n_pad);
    st.kblocks = static_cast<uint8_t>(k_pad / 64);
    st.acc = static_cast<uint8_t>(acc);
    st.d_col = static_cast<uint16_t>(d_col);
    blob.resize(blob.size() + st.w_bytes);
    if (w) pack_sw128_f16(w, n_real, k_real, ld, n_pad, k_pad, blob.data() + st.w_off);
    if (want_bias) {
      const size_t off = blob.size();
      blob.resize(off + static_cast<size_t>(n_pad) * 32);
      if (bias) pack_sw32_bias_f16(bias, n_real, n_pad, blob.data() + off);
      st.w_bytes += static_cast<uint32_t>(n_pad * 32);
    }
    p.steps[p.nsteps++] = st;
  }
};

// The kernels write the encoder input as two K-blocks (write_posenc):
//   K-block 0: [sin block (30) | 0 0 | cos block (30) | 0 0]  -- each of a row's two threads owns one block with
//              compile-time (dim, freq) per column; the 60 encoding columns of W[n][ld] are re-ordered to match;
//   K-block 1, first 16 columns: [x_hi(3) | x_lo(3) | x_hi(3) | 0...] -- the raw xyz input split into an fp16
//              head and an fp16 remainder; the matching weight columns are [W | W | W_lo] (W rounds to its fp16
//              head W_hi when packed, W_lo = W - W_hi), so x.W = x_hi W_hi + x_lo W_hi + x_hi W_lo keeps ~22 bits
//              although every operand is fp16 (kStepK16Tail: one extra K=16 MMA).
std::vector<float> enc_cols(const float* w, int n, int ld) {
  std::vector<float> o(static_cast<size_t>(n) * 128, 0.f);
  if (w)
    for (int r = 0; r < n; ++r) {
      float* orow = o.data() + static_cast<size_t>(r) * 128;
      const float* wr = w + static_cast<size_t>(r) * ld;
      for (int c = 0; c < 60; ++c) orow[c < 30 ? c : c + 2] = wr[c];
      for (int d = 0; d < 3; ++d) {
        const float wv = wr[60 + d];
        const uint16_t hb = f32_to_f16_bits(wv);
        __half hh;
        std::memcpy(&hh, &hb, 2);
        orow[64 + d] = wv;
        orow[67 + d] = wv;
        orow[70 + d] = wv - __half2float(hh);
      }
    }
  return o;
}

// lin_in step + hoisted lin_z rows; the block steps are appended separately (program order differs
// between heads).
void trunk_lin_in(Builder& b, Program& prog, const std::string& p, int flags = 0) {
  const float* w = b.get(p + ".lin_in.weight", 128 * 63);
  const float* bi = b.get(p + ".lin_in.bias", 128);
  b.step(prog, w ? enc_cols(w, 128, 63).data() : nullptr, 128, 128, 128, 128, 128, /*d_col=*/0, /*acc=*/0,
         flags | kStepK16Tail, bi, true);
  for (int k = 0; k < 3; ++k) {
    const float* wz = b.get(p + ".lin_z." + std::to_string(k) + ".weight", 128 * 512);
    const float* bz = b.get(p + ".lin_z." + std::to_string(k) + ".bias", 128);
    if (wz && bz) {
      b.hoist_w.insert(b.hoist_w.end(), wz, wz + 128 * 512);
      b.hoist_b.insert(b.hoist_b.end(), bz, bz + 128);
    }
  }
}
void trunk_blocks(Builder& b, Program& prog, const std::string& p, int d_out, int n_out_pad) {
  for (int k = 0; k < 5; ++k) {
    const std::string q = p + ".blocks." + std::to_string(k);
    const float* w0 = b.get(q + ".fc_0.weight", 128 * 128);
    const float* b0 = b.get(q + ".fc_0.bias", 128);
    const float* w1 = b.get(q + ".fc_1.weight", 128 * 128);
    const float* b1 = b.get(q + ".fc_1.bias", 128);
    b.step(prog, w0, 128, 128, 128, 128, 128, /*d_col=*/128, 0, 0, b0, true);
    b.step(prog, w1, 128, 128, 128, 128, 128, /*d_col=*/0, 1, 0, b1, true);  // x += fc_1(.) + b1
  }
  const float* wo = b.get(p + ".lin_out.weight", static_cast<int64_t>(d_out) * 128);
  const float* bo = b.get(p + ".lin_out.bias", d_out);
  b.step(prog, wo, d_out, 128, 128, n_out_pad, 128, /*d_col=*/128, 0, 0, bo, true);
}

// The attention / feed-forward layers and jacobian_head run in xf_kernel from their own blob.
// Exact folds done here in fp64: keys/values (they depend only on the learned index embedding,
// transformer.py:63-78, action_decoder_jacobian.py:431-435), the LayerNorm affine (PreNorm,
// transformer.py:14-21: W (g*n + b) = (W diag g) n + W b) and log2(e) into the logits.
void build_head(Builder& b, int A, Program& hp, std::vector<uint8_t>& xf_blob) {
  const float* emb = b.get("decoder.jacobian_index_embedding", static_cast<int64_t>(A) * 64);
  hp = Program{};
  {
    Builder hb;
    for (int l = 0; l < 3; ++l) {
      const std::string p = "decoder.jacobian_attn_decoder.layers." + std::to_string(l);
      const float* g1 = b.get(p + ".0.norm.weight", 64);
      const float* be1 = b.get(p + ".0.norm.bias", 64);
      const float* wq_ = b.get(p + ".0.fn.to_q.weight", 512 * 64);
      const float* wkv = b.get(p + ".0.fn.to_kv.weight", 1024 * 64);
      const float* wo = b.get(p + ".0.fn.to_out.0.weight", 64 * 512);
      const float* bo = b.get(p + ".0.fn.to_out.0.bias", 64);
      const float* g2 = b.get(p + ".1.norm.weight", 64);
      const float* be2 = b.get(p + ".1.norm.bias", 64);
      const float* w1 = b.get(p + ".1.fn.net.0.weight", 64 * 64);
      const float* bb1 = b.get(p + ".1.fn.net.0.bias", 64);
      const float* w2 = b.get(p + ".1.fn.net.3.weight", 64 * 64);
      const float* bb2 = b.get(p + ".1.fn.net.3.bias", 64);
      std::vector<float> m1(64 * 64, 0.f), m1b(64, 0.f), m2(64 * 64, 0.f), w1g(64 * 64, 0.f), w1b(64, 0.f);
      if (wq_ && wkv && wo && emb && g1 && be1) {
        std::vector<double> K(static_cast<size_t>(A) * 512), V(static_cast<size_t>(A) * 512);
        for (int a = 0; a < A; ++a)
          for (int j = 0; j < 512; ++j) {
            double sk = 0, sv = 0;
            for (int i = 0; i < 64; ++i) {
              sk += static_cast<double>(wkv[j * 64 + i]) * emb[a * 64 + i];
              sv += static_cast<double>(wkv[(512 + j) * 64 + i]) * emb[a * 64 + i];
            }
            K[a * 512 + j] = sk;
            V[a * 512 + j] = sv;
          }
        const double scale = 1.4426950408889634 / std::sqrt(64.0);  // log2(e) * dim_head^-0.5
        for (int h = 0; h < 8; ++h)
          for (int a = 0; a < A; ++a) {
            const int n = h * 8 + a;  // logit / attention column (heads padded to 8 keys)
            double sb = 0;
            for (int k = 0; k < 64; ++k) {
              double s = 0;
              for (int d = 0; d < 64; ++d) s += static_cast<double>(wq_[(h * 64 + d) * 64 + k]) * K[a * 512 + h * 64 + d];
              m1[n * 64 + k] = static_cast<float>(scale * s * g1[k]);
              sb += scale * s * be1[k];
            }
            m1b[n] = static_cast<float>(sb);
            for (int o = 0; o < 64; ++o) {
              double s = 0;
              for (int d = 0; d < 64; ++d) s += static_cast<double>(wo[o * 512 + h * 64 + d]) * V[a * 512 + h * 64 + d];
              m2[o * 64 + n] = static_cast<float>(s);
            }
          }
      }
      if (w1 && bb1 && g2 && be2)
        for (int n = 0; n < 64; ++n) {
          double sb = bb1[n];
          for (int k = 0; k < 64; ++k) {
            w1g[n * 64 + k] = w1[n * 64 + k] * g2[k];
            sb += static_cast<double>(w1[n * 64 + k]) * be2[k];
          }
          w1b[n] = static_cast<float>(sb);
        }
      hb.step(hp, m1.data(), 64, 64, 64, 64, 64, /*d_col=*/64, 0, 0, m1b.data(), true);
      hb.step(hp, m2.data(), 64, 64, 64, 64, 64, /*d_col=*/0, 1, 0, bo, true);      // x += attn out
      hb.step(hp, w1g.data(), 64, 64, 64, 64, 64, /*d_col=*/64, 0, 0, w1b.data(), true);
      hb.step(hp, w2, 64, 64, 64, 64, 64, /*d_col=*/0, 1, 0, bb2, true);             // x += feed-forward out
    }
    const float* wh = b.get("decoder.jacobian_head.weight", static_cast<int64_t>(3 * A) * 64);
    const float* bh = b.get("decoder.jacobian_head.bias", 3 * A);
    hb.step(hp, wh, 3 * A, 64, 64, 32, 64, /*d_col=*/64, 0, 0, bh, true);
    xf_blob.swap(hb.blob);
  }
}

// q_enc step image: jacobian_query_mlp's positional-encoding / xyz columns (+ bias block); `out` = n*k*2 + n*32 bytes
void pack_q_enc(const float* wq, const float* bq, uint8_t* out) {
  const std::vector<float> cols = enc_cols(wq, 64, 575);
  pack_sw128_f16(cols.data(), 64, 128, 128, 64, 128, out);
  pack_sw32_bias_f16(bq, 64, 64, out + 64 * 128 * 2);
}

}  // namespace

extern "C" int njf_field_create(const NjfFieldDesc* desc, const NjfTensor* tensors, int n_tensors,
                                NjfField** out) {
  if (!desc || !tensors || !out) NJF_FAIL("njf_field_create: null argument");
  if (desc->encoder_dim != 512) NJF_FAIL("encoder_dim %d unsupported (kernels are built for 512)", desc->encoder_dim);
  if (desc->n_proposal < 1 || desc->n_proposal > NJF_MAX_LEVELS) NJF_FAIL("n_proposal %d out of range", desc->n_proposal);
  const int A = desc->action_dim;
  if (desc->head == NJF_HEAD_TRANSFORMER) {
    if (A < 1 || A > 8) NJF_FAIL("jacobian_transformer: action_dim %d unsupported (1..8)", A);
  } else if (desc->head == NJF_HEAD_MLP) {
    if (A < 1 || A > 10) NJF_FAIL("jacobian_mlp: action_dim %d unsupported (1..10)", A);
  } else {
    NJF_FAIL("unknown head %d", desc->head);
  }
  Builder b;
  for (int i = 0; i < n_tensors; ++i) b.t[tensors[i].name] = {tensors[i].data, tensors[i].numel};

  // owned until success: every failure path below frees the device buffers already allocated
  struct Guard {
    NjfField* f;
    ~Guard() { if (f) njf_field_destroy(f); }
  } guard{new NjfField()};   // value-initialised: all tables start at zero
  NjfField* f = guard.f;
  f->desc = *desc;
  if (desc->sh_convention != NJF_SH_TCNN && desc->sh_convention != NJF_SH_NERFSTUDIO_TORCH)
    NJF_FAIL("unknown sh_convention %d", desc->sh_convention);
  // hoist channel order: proposal nets (384 each), then the main map (dens 384 + head part)
  for (int i = 0; i < desc->n_proposal; ++i) {
    const std::string p = "proposal_networks." + std::to_string(i) + ".density_head";
    trunk_lin_in(b, f->prop_prog[i], p);
    trunk_blocks(b, f->prop_prog[i], p, 1, 16);
  }
  Program& fp = f->field_prog;
  std::vector<uint8_t> xf_blob;
  // transformer head: lin_in and q_enc read the same A tile (the positional encoding) and are
  // covered by ONE accumulator commit (two arrivals on one mbarrier phase would be unsafe)
  trunk_lin_in(b, fp, "decoder.density_head", desc->head == NJF_HEAD_TRANSFORMER ? kStepNoCommit : 0);
  if (desc->head == NJF_HEAD_TRANSFORMER) {
    f->ch_main = 448;
    const float* wq = b.get("decoder.jacobian_query_mlp.weight", 64 * 575);
    const float* bq = b.get("decoder.jacobian_query_mlp.bias", 64);
    const float* emb = b.get("decoder.jacobian_index_embedding", static_cast<int64_t>(A) * 64);
    b.step(fp, wq ? enc_cols(wq, 64, 575).data() : nullptr, 64, 128, 128, 64, 128, /*d_col=*/128, 0,
           kStepReuseA | kStepK16Tail, bq, true);
    f->q_enc_step = fp.nsteps - 1;
    (void)emb;
    if (wq && bq) {
      // hoisted query channels: the 512 feature columns of jacobian_query_mlp (its bias rides in the q_enc step)
      for (int c = 0; c < 64; ++c) b.hoist_w.insert(b.hoist_w.end(), wq + c * 575 + 63, wq + c * 575 + 575);
      b.hoist_b.insert(b.hoist_b.end(), 64, 0.f);
    }
    build_head(b, A, f->head_prog, xf_blob);
  }
  trunk_blocks(b, fp, "decoder.density_head", 16, 16);
  {
    const float* w1 = b.get("decoder.color_head.0.weight", 64 * 31);
    const float* b1 = b.get("decoder.color_head.0.bias", 64);
    const float* w2 = b.get("decoder.color_head.2.weight", 64 * 64);
    const float* b2 = b.get("decoder.color_head.2.bias", 64);
    const float* w3 = b.get("decoder.color_head.4.weight", 3 * 64);
    const float* b3 = b.get("decoder.color_head.4.bias", 3);
    b.step(fp, w1, 64, 31, 31, 64, 64, 128, 0, 0, b1, true);
    b.step(fp, w2, 64, 64, 64, 64, 64, 128, 0, 0, b2, true);
    if (w3 && b3) {
      std::memcpy(f->color.w3, w3, 192 * 4);
      std::memcpy(f->color.b3, b3, 3 * 4);
    }
  }
  if (desc->head == NJF_HEAD_MLP) {
    f->ch_main = 768;
    trunk_lin_in(b, fp, "decoder.jacobian_head");
    trunk_blocks(b, fp, "decoder.jacobian_head", 3 * A, 32);
  }
  if (!b.err.empty()) NJF_FAIL("njf_field_create: %s", b.err.c_str());
  f->ch_total = desc->n_proposal * f->ch_prop + f->ch_main;
  if (static_cast<int>(b.hoist_b.size()) != f->ch_total)
    NJF_FAIL("internal: hoist rows %zu != %d", b.hoist_b.size(), f->ch_total);
  NJF_CUDA(cudaMalloc(&f->d_blob, b.blob.size()));
  NJF_CUDA(cudaMalloc(&f->d_hoist_w, b.hoist_w.size() * sizeof(float)));
  NJF_CUDA(cudaMalloc(&f->d_hoist_b, b.hoist_b.size() * sizeof(float)));
  NJF_CUDA(cudaMemcpy(f->d_blob, b.blob.data(), b.blob.size(), cudaMemcpyHostToDevice));
  NJF_CUDA(cudaMemcpy(f->d_hoist_w, b.hoist_w.data(), b.hoist_w.size() * sizeof(float), cudaMemcpyHostToDevice));
  NJF_CUDA(cudaMemcpy(f->d_hoist_b, b.hoist_b.data(), b.hoist_b.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (!xf_blob.empty()) {
    if (xf_blob.size() > kXfMaxBlob) NJF_FAIL("internal: head blob %zu B exceeds the resident budget", xf_blob.size());
    f->xf_bytes = static_cast<uint32_t>(xf_blob.size());
    NJF_CUDA(cudaMalloc(&f->d_xf_blob, xf_blob.size()));
    NJF_CUDA(cudaMemcpy(f->d_xf_blob, xf_blob.data(), xf_blob.size(), cudaMemcpyHostToDevice));
  }
  if (njf_hoist_build(f, b.hoist_w, b.hoist_b)) return 1;
  for (int i = 0; i < desc->n_proposal; ++i) f->prop_blob[i] = f->d_blob;
  f->field_blob = f->d_blob;
  guard.f = nullptr;
  *out = f;
  return 0;
}

extern "C" void njf_field_destroy(NjfField* f) {
  if (!f) return;
  cudaFree(f->d_blob);
  cudaFree(f->d_hoist_img);
  cudaFree(f->d_xf_blob);
  cudaFree(f->d_hoist_w);
  cudaFree(f->d_hoist_b);
  delete f;
}

// ---- partial re-pack: only the Jacobian-head tensors changed (an optimiser step of the action phase,
// models/model_wrapper.py:75-85).  Re-packs the cross-attention blob, the q_enc step image and the 64 hoisted query
// rows in place -- no allocation, ~0.6 MB of host->device copies instead of a full njf_field_create.
extern "C" int njf_field_update_head(NjfField* f, const NjfTensor* tensors, int n_tensors, void* stream_) {
  if (!f || !tensors) NJF_FAIL("njf_field_update_head: null argument");
  if (f->desc.head != NJF_HEAD_TRANSFORMER) NJF_FAIL("njf_field_update_head: cross-attention head only");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int A = f->desc.action_dim;
  Builder b;
  for (int i = 0; i < n_tensors; ++i) b.t[tensors[i].name] = {tensors[i].data, tensors[i].numel};
  Program hp{};
  std::vector<uint8_t> xf_blob;
  build_head(b, A, hp, xf_blob);
  const float* wq = b.get("decoder.jacobian_query_mlp.weight", 64 * 575);
  const float* bq = b.get("decoder.jacobian_query_mlp.bias", 64);
  if (!b.err.empty()) NJF_FAIL("njf_field_update_head: %s", b.err.c_str());
  if (xf_blob.size() != f->xf_bytes) NJF_FAIL("internal: head blob size changed");
  const MmaStep& st = f->field_prog.steps[f->q_enc_step];
  std::vector<uint8_t> qimg(st.w_bytes, 0);
  if (qimg.size() != 64 * 128 * 2 + 64 * 32) NJF_FAIL("internal: q_enc image size");
  pack_q_enc(wq, bq, qimg.data());
  // hoisted query rows: rows [384, 448) of the main map's slab, 8 K-block images of [N rows x 128 B]
  const NjfField::HoistJobHost* job = nullptr;
  for (const auto& j : f->hoist_jobs)
    if (j.map == f->desc.n_proposal && j.c0 <= 384 && 384 + 64 <= j.c0 + j.N) job = &j;
  if (!job) NJF_FAIL("internal: hoist slab of the query rows not found");
  std::vector<float> rows(64 * 512);
  for (int c = 0; c < 64; ++c) std::memcpy(rows.data() + c * 512, wq + c * 575 + 63, 512 * sizeof(float));
  std::vector<uint8_t> himg(8 * 64 * 128);
  for (int kb = 0; kb < 8; ++kb) pack_sw128_f16(rows.data() + kb * 64, 64, 64, 512, 64, 64, himg.data() + kb * 64 * 128);
  // the previous render on this stream may still read the old images
  NJF_CUDA(cudaStreamSynchronize(stream));
  NJF_CUDA(cudaMemcpy(f->d_xf_blob, xf_blob.data(), xf_blob.size(), cudaMemcpyHostToDevice));
  NJF_CUDA(cudaMemcpy(f->d_blob + st.w_off, qimg.data(), qimg.size(), cudaMemcpyHostToDevice));
  const int r0 = 384 - job->c0;
  for (int kb = 0; kb < 8; ++kb)
    NJF_CUDA(cudaMemcpy(f->d_hoist_img + job->w_off + static_cast<size_t>(kb) * job->N * 128 + static_cast<size_t>(r0) * 128,
                        himg.data() + kb * 64 * 128, 64 * 128, cudaMemcpyHostToDevice));
  const size_t grow = static_cast<size_t>(f->desc.n_proposal) * f->ch_prop + 384;   // row in the fp32 hoist matrix
  NJF_CUDA(cudaMemcpy(f->d_hoist_w + grow * 512, rows.data(), rows.size() * sizeof(float), cudaMemcpyHostToDevice));
  f->head_prog = hp;
  return 0;
}
