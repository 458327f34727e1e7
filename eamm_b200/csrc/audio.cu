// Kernels of the per-clip audio->motion-feature network AT_net2 (SURVEY.md section 8(f) rank 4;
// reference /root/reference/modules/util.py:514-613) that the conv kernels do not already cover:
//   eamm_linear      fp32 GEMM + bias + ReLU + scale, optional row-broadcast addend   nn.Linear  util.py:532-556
//   eamm_maxpool     k x k max pooling with independent strides, no padding           nn.MaxPool2d util.py:543,547
//   eamm_lstm_layer  one LSTM layer's recurrence over a whole clip                    nn.LSTM    util.py:557,597
// The LSTM recurrence is the only sequential part of the clip: one thread-block cluster of 8 CTAs per
// sequence keeps W_hh (1 MB fp32) entirely in registers (128 KB per CTA), and the 256-float hidden state
// is exchanged through distributed shared memory once per time step -- no global-memory round trip and
// no kernel launch per step.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace eamm {

// ---------------------------------------------------------------------------------------------
// Y[m][n] = scale * act( sum_k X[m][k] W[k][n] + bias[n] + add[m / add_period][n] )
// 64x64x16 tiles, 256 threads, 4x4 outputs per thread (same shape as conv_simt).
// ---------------------------------------------------------------------------------------------
struct LinearParams {
  const float* x; const float* w; const float* bias; const float* add; float* y;
  int M, K, N, ldx, ldy, add_period, relu, vec_a;
  float scale;
};

__global__ void __launch_bounds__(256)
linear_kernel(LinearParams p) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;       // A: 64 rows x 16 k, one float4 per thread
  const int b_row = tid >> 4, b_col = (tid & 15) * 4;    // B: 16 k x 64 n
  const bool a_ok = m0 + a_row < p.M;
  const float* xa = p.x + (long long)(m0 + a_row) * p.ldx;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += 16) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    if (a_ok) {
      if (p.vec_a && k0 + a_k + 3 < p.K) {
        float4 v = __ldg(reinterpret_cast<const float4*>(xa + k0 + a_k));
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (k0 + a_k + i < p.K) a[i] = __ldg(xa + k0 + a_k + i);
      }
    }
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k0 + b_row < p.K && n0 + b_col < p.N)
      b = __ldg(reinterpret_cast<const float4*>(p.w + (long long)(k0 + b_row) * p.N + n0 + b_col));
    __syncthreads();
    As[a_k + 0][a_row] = a[0]; As[a_k + 1][a_row] = a[1]; As[a_k + 2][a_row] = a[2]; As[a_k + 3][a_row] = a[3];
    *reinterpret_cast<float4*>(&Bs[b_row][b_col]) = b;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }

  const int n = n0 + tx * 4;
  if (n >= p.N) return;
  float4 bias = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    float4 v = make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w);
    if (p.add) {
      float4 r = __ldg(reinterpret_cast<const float4*>(p.add + (long long)(m / p.add_period) * p.N + n));
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    v.x *= p.scale; v.y *= p.scale; v.z *= p.scale; v.w *= p.scale;
    *reinterpret_cast<float4*>(p.y + (long long)m * p.ldy + n) = v;
  }
}

// ---------------------------------------------------------------------------------------------
// max pooling, window k x k, strides (sy, sx), no padding, floor mode (nn.MaxPool2d defaults)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool_kernel(ActView in, ActView out, int k, int sy, int sx, long long total) {
  const int c4 = out.c >> 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int g = (int)(idx % c4);
    long long pix = idx / c4;
    int ox = (int)(pix % out.w);
    int oy = (int)((pix / out.w) % out.h);
    int n = (int)(pix / ((long long)out.w * out.h));
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) {
        float4 v = act_load4(in, act_offset(in, n, oy * sy + dy, ox * sx + dx, 4 * g));
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    act_store4(out, act_offset(out, n, oy, ox, 4 * g), m);
  }
}

// ---------------------------------------------------------------------------------------------
// region copy between two views of different map sizes / storage formats: dst(n, y < h, x < w, :) = src(n, y, x, :);
// zero_rest also clears dst's other pixels (src == dst: only that).  The MFCC convs run on tensor cores over maps padded
// to powers of two (28x12 -> 32x16, 26x5 -> 32x8): this moves data into / out of the padded buffers and re-zeroes the
// padding a convolution has written, so that the next layer still sees the reference's zero padding.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
act_copy_kernel(ActView src, ActView dst, int h, int w, int zero_rest, int in_place, long long total) {
  const int c4 = dst.c >> 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int g = (int)(idx % c4);
    long long pix = idx / c4;
    int x = (int)(pix % dst.w);
    int y = (int)((pix / dst.w) % dst.h);
    int n = (int)(pix / ((long long)dst.w * dst.h));
    if (y < h && x < w) {
      if (!in_place) act_store4(dst, act_offset(dst, n, y, x, 4 * g), act_load4(src, act_offset(src, n, y, x, 4 * g)));
    } else if (zero_rest) {
      act_store4(dst, act_offset(dst, n, y, x, 4 * g), make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LSTM layer recurrence (gate order i, f, g, o as nn.LSTM; zero initial state, util.py:581-582).
// gates_x [B][T][4H] already holds W_ih x_t + b_ih + b_hh (eamm_linear).  Cluster of 8 CTAs per
// sequence; CTA r owns hidden units [32r, 32r+32) = 128 gate rows; thread (row, half) keeps half a row
// of W_hh (128 floats) in registers.  Per step: 128 FMAs per thread against the hidden state in shared
// memory, halves combined through shared memory, 32 threads finish the cell update and push the new h
// into every CTA's (double-buffered) copy through DSMEM, one cluster barrier.
// ---------------------------------------------------------------------------------------------
constexpr int LSTM_H = 256, LSTM_CL = 8, LSTM_U = LSTM_H / LSTM_CL;   // 32 units per CTA

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __cluster_dims__(LSTM_CL, 1, 1) __launch_bounds__(256, 1)
lstm_layer_kernel(const float* __restrict__ gx, const float* __restrict__ whh, float* __restrict__ hout, int T) {
  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();
  const int b = blockIdx.x / LSTM_CL;
  __shared__ __align__(16) float hbuf[2][LSTM_H];
  __shared__ float part[256];
  const int tid = threadIdx.x;
  const int row = tid & 127, gate = row >> 5, ul = row & 31, half = tid >> 7;

  float w[128];
  {
    const float4* wr = reinterpret_cast<const float4*>(whh + (size_t)(gate * LSTM_H + r * LSTM_U + ul) * LSTM_H + half * 128);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float4 v = __ldg(wr + i);
      w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
  }
  hbuf[0][tid] = 0.f;
  hbuf[1][tid] = 0.f;
  float c = 0.f;
  cluster.sync();                       // every CTA's state is initialised before any remote write

  const float* gxb = gx + (size_t)b * T * 4 * LSTM_H + r * LSTM_U + tid;
  float* hob = hout + (size_t)b * T * LSTM_H + r * LSTM_U + tid;
  for (int t = 0; t < T; ++t) {
    const int cur = t & 1;
    float gxv[4] = {0.f, 0.f, 0.f, 0.f};
    if (tid < LSTM_U) {
#pragma unroll
      for (int g = 0; g < 4; ++g) gxv[g] = __ldg(gxb + (size_t)t * 4 * LSTM_H + g * LSTM_H);
    }
    const float4* hv = reinterpret_cast<const float4*>(&hbuf[cur][half * 128]);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float4 h4 = hv[i];
      a0 = fmaf(w[4 * i], h4.x, a0);
      a1 = fmaf(w[4 * i + 1], h4.y, a1);
      a2 = fmaf(w[4 * i + 2], h4.z, a2);
      a3 = fmaf(w[4 * i + 3], h4.w, a3);
    }
    part[tid] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (tid < LSTM_U) {
      float pre[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) pre[g] = (part[g * 32 + tid] + part[128 + g * 32 + tid]) + gxv[g];
      const float ig = sigmoidf_acc(pre[0]), fg = sigmoidf_acc(pre[1]), gg = tanhf(pre[2]), og = sigmoidf_acc(pre[3]);
      c = fmaf(fg, c, ig * gg);
      const float h = og * tanhf(c);
      hob[(size_t)t * LSTM_H] = h;
      float* mine = &hbuf[cur ^ 1][r * LSTM_U + tid];
#pragma unroll
      for (int d = 0; d < LSTM_CL; ++d) *cluster.map_shared_rank(mine, d) = h;
    }
    cluster.sync();                     // new h visible everywhere; old buffer free for the next step
  }
}

}  // namespace eamm

using namespace eamm;

extern "C" int eamm_linear(const float* x, int ldx, const float* w, const float* bias, const float* add_rows,
                           int add_period, float* y, int ldy, int M, int K, int N, int relu, float scale,
                           void* stream) {
  if (!x || !w || !y || M <= 0 || K <= 0 || N <= 0) return EAMM_ERR_ARG;
  if (N % 4 || ldy % 4 || ldx < K || ldy < N) return EAMM_ERR_SHAPE;
  if (((uintptr_t)w | (uintptr_t)y | (uintptr_t)bias | (uintptr_t)add_rows) % 16) return EAMM_ERR_ALIGN;
  if (add_rows && add_period <= 0) return EAMM_ERR_ARG;
  LinearParams p;
  p.x = x; p.w = w; p.bias = bias; p.add = add_rows; p.y = y;
  p.M = M; p.K = K; p.N = N; p.ldx = ldx; p.ldy = ldy; p.add_period = add_rows ? add_period : 1;
  p.relu = relu; p.scale = scale;
  p.vec_a = (ldx % 4 == 0) && ((uintptr_t)x % 16 == 0);
  dim3 grid((M + 63) / 64, (N + 63) / 64);
  linear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  EAMM_LAUNCH_CHECK();
  return 0;
}

extern "C" int eamm_maxpool(const eamm_act* in, const eamm_act* out, int k, int stride_y, int stride_x, void* stream) {
  int rc = check_view(in); if (rc) return rc;
  rc = check_view(out); if (rc) return rc;
  if (k <= 0 || stride_y <= 0 || stride_x <= 0 || in->h < k || in->w < k) return EAMM_ERR_ARG;
  if (out->n != in->n || out->c != in->c || out->c % 4) return EAMM_ERR_SHAPE;
  if (out->h != (in->h - k) / stride_y + 1 || out->w != (in->w - k) / stride_x + 1) return EAMM_ERR_SHAPE;
  ActView vi = make_view(in), vo = make_view(out);
  long long total = (long long)out->n * out->h * out->w * (out->c / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  maxpool_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(vi, vo, k, stride_y, stride_x, total);
  EAMM_LAUNCH_CHECK();
  return 0;
}

extern "C" int eamm_act_copy(const eamm_act* src, const eamm_act* dst, int h, int w, int zero_rest, void* stream) {
  int rc = check_view(src); if (rc) return rc;
  rc = check_view(dst); if (rc) return rc;
  if (h <= 0 || w <= 0 || h > src->h || w > src->w || h > dst->h || w > dst->w) return EAMM_ERR_ARG;
  if (src->n != dst->n || src->c != dst->c || dst->c % 4) return EAMM_ERR_SHAPE;
  const int in_place = src->data == dst->data && src->c_off == dst->c_off;
  if (in_place && (src->h != dst->h || src->w != dst->w || src->dtype != dst->dtype || src->planes != dst->planes)) return EAMM_ERR_ARG;
  if (in_place && !zero_rest) return 0;
  ActView vs = make_view(src), vd = make_view(dst);
  long long total = (long long)dst->n * dst->h * dst->w * (dst->c / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  act_copy_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(vs, vd, h, w, zero_rest, in_place, total);
  EAMM_LAUNCH_CHECK();
  return 0;
}

extern "C" int eamm_lstm_layer(const float* gates_x, const float* w_hh, float* h_out, int B, int T, int hidden,
                               void* stream) {
  if (!gates_x || !w_hh || !h_out || B <= 0 || T <= 0) return EAMM_ERR_ARG;
  if (hidden != LSTM_H) return EAMM_ERR_UNSUPPORTED;
  if (((uintptr_t)w_hh) % 16) return EAMM_ERR_ALIGN;
  lstm_layer_kernel<<<B * LSTM_CL, 256, 0, (cudaStream_t)stream>>>(gates_x, w_hh, h_out, T);
  EAMM_LAUNCH_CHECK();
  return 0;
}
