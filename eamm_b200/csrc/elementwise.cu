// HBM-bound kernels of the EAMM generation path: anti-alias downsample, keypoint stage,
// flow combine, feature warp x occlusion, image warp, layout conversion.
// Reference semantics cited per kernel; arithmetic is fp32 throughout (north_star: "fp32 warp").
#include "common.cuh"

namespace eamm {

// =============================================================================================
// a3  AntiAliasInterpolation2d  (util.py:1044-1052)
//   out[c][i][j] = sum_{u,v} k2[u][v] * src[c][step*i+u-pad][step*j+v-pad],  k2 = outer(g1,g1) (sum 1), pad = taps/2
// Block = (batch item, TR output rows).  Phase 1: horizontal 13-tap filter at the subsampled
// columns for every needed input row (global reads, L1 absorbs the 13/step overlap); phase 2:
// vertical 13-tap filter from shared memory; one float4 (R,G,B,0) store per output pixel.
// =============================================================================================
constexpr int AA_MAX_TAPS = 13;   // the reference hard-codes sigma = 1.5 -> 13 taps (util.py:1011-1013); 1 tap = plain copy (scale 1)
constexpr int AA_TR = 4;  // output rows per block

__global__ void __launch_bounds__(256)
aa_downsample_kernel(const float* __restrict__ src, long long src_n_stride, float4* __restrict__ dst,
                     int H, int W, int Ho, int Wo, int step, int AA_TAPS, const float* __restrict__ g1, ActView act,
                     int to_act) {
  extern __shared__ float tmp[];  // [3][rows][Wo]
  __shared__ float g[AA_MAX_TAPS];
  const int n = blockIdx.y;
  const int r0 = blockIdx.x * AA_TR;
  const int AA_PAD = AA_TAPS >> 1;          // util.py:1044-1047: ka = ks // 2 on every side (ks is odd)
  const int rows = (AA_TR - 1) * step + AA_TAPS;
  if (threadIdx.x < AA_TAPS) g[threadIdx.x] = g1[threadIdx.x];
  __syncthreads();
  const float* img = src + (long long)n * src_n_stride;
  const int in_row0 = r0 * step - AA_PAD;
  for (int idx = threadIdx.x; idx < 3 * rows * Wo; idx += blockDim.x) {
    int j = idx % Wo;
    int rr = (idx / Wo) % rows;
    int c = idx / (Wo * rows);
    int y = in_row0 + rr;
    float acc = 0.f;
    if (y >= 0 && y < H) {
      const float* rowp = img + ((long long)c * H + y) * W;
      int x0 = j * step - AA_PAD;
      for (int v = 0; v < AA_TAPS; ++v) {
        int x = x0 + v;
        float s = (x >= 0 && x < W) ? __ldg(rowp + x) : 0.f;
        acc = fmaf(g[v], s, acc);
      }
    }
    tmp[idx] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < AA_TR * Wo; idx += blockDim.x) {
    int j = idx % Wo;
    int i = idx / Wo;
    if (r0 + i >= Ho) continue;
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float acc = 0.f;
      for (int u = 0; u < AA_TAPS; ++u) acc = fmaf(g[u], tmp[(c * rows + i * step + u) * Wo + j], acc);
      o[c] = acc;
    }
    if (to_act) act_store4(act, act_offset(act, n, r0 + i, j, 0), make_float4(o[0], o[1], o[2], 0.f));
    else dst[((long long)n * Ho + (r0 + i)) * Wo + j] = make_float4(o[0], o[1], o[2], 0.f);
  }
}

// =============================================================================================
// a4+a5+a6  keypoint stage  (dense_motion.py:32-79, util.py:815-855)
// For every pixel z of the h x w grid and every k in [0,K]:
//   H_k   = exp(-0.5|z-kp_d,k|^2/var) - exp(-0.5|z-kp_s,k|^2/var)          (H_0 = 0)
//   z'_k  = J_k (z - kp_d,k) + kp_s,k,  J_k = J_s,k inv(J_d,k)             (z'_0 = z)
//   S_k   = bilinear sample (zeros padding, align_corners=False) of the small source at z'_k
// Writes [H_k, S_k.R, S_k.G, S_k.B] into channels 4k..4k+3 of the hourglass input (NHWC) and
// S_k into sparse_deformed [n,K+1,3,h,w].  Block = one image row; sparse_deformed goes through
// shared memory so both outputs are written coalesced.
// =============================================================================================
constexpr int KP_MAX = 32;

struct KpDev {
  const float* value; const float* jac; long long vs, js;
};

__global__ void __launch_bounds__(256)
kp_stage_kernel(const float4* __restrict__ small_img, long long small_n_stride, KpDev kd, KpDev ks,
                int K, float kp_var, ActView hg, float* __restrict__ sdef, int* __restrict__ status) {
  extern __shared__ float s_sd[];  // [(K+1)*3][w]
  __shared__ float s_kd[KP_MAX * 2], s_ks[KP_MAX * 2], s_J[KP_MAX * 4];
  const int n = blockIdx.y, y = blockIdx.x;
  const int h = hg.h, w = hg.w, K1 = K + 1;
  if (threadIdx.x < K) {
    int k = threadIdx.x;
    const float* vd = kd.value + (long long)n * kd.vs + k * 2;
    const float* vs = ks.value + (long long)n * ks.vs + k * 2;
    s_kd[2 * k] = vd[0]; s_kd[2 * k + 1] = vd[1];
    s_ks[2 * k] = vs[0]; s_ks[2 * k + 1] = vs[1];
    float J[4] = {1.f, 0.f, 0.f, 1.f};
    if (kd.jac != nullptr) {
      bool ok = kp_affine(kd.jac + (long long)n * kd.js + k * 4, ks.jac + (long long)n * ks.js + k * 4, J);
      if (!ok && status != nullptr && y == 0) atomicOr(status, 1);
    }
    s_J[4 * k] = J[0]; s_J[4 * k + 1] = J[1]; s_J[4 * k + 2] = J[2]; s_J[4 * k + 3] = J[3];
  }
  __syncthreads();
  const float4* img = small_img + ((long long)n * small_n_stride) / 4;   // stride is in floats
  const float zy = grid_coord(y, h);
  for (int idx = threadIdx.x; idx < w * K1; idx += blockDim.x) {
    int k = idx % K1, x = idx / K1;
    float zx = grid_coord(x, w);
    float hm = 0.f, gx = zx, gy = zy;
    if (k > 0) {
      int kk = k - 1;
      float ddx = zx - s_kd[2 * kk], ddy = zy - s_kd[2 * kk + 1];
      float dsx = zx - s_ks[2 * kk], dsy = zy - s_ks[2 * kk + 1];
      hm = expf(-0.5f * (ddx * ddx + ddy * ddy) / kp_var) - expf(-0.5f * (dsx * dsx + dsy * dsy) / kp_var);
      float mx = ddx, my = ddy;
      if (kd.jac != nullptr) {
        mx = s_J[4 * kk] * ddx + s_J[4 * kk + 1] * ddy;
        my = s_J[4 * kk + 2] * ddx + s_J[4 * kk + 3] * ddy;
      }
      gx = mx + s_ks[2 * kk];
      gy = my + s_ks[2 * kk + 1];
    }
    Bilinear b = bilinear_setup(gx, gy, w, h);
    float wx0 = 1.f - b.wx1, wy0 = 1.f - b.wy1;
    float r = 0.f, g = 0.f, bl = 0.f;
    bool xin0 = b.x0 >= 0 && b.x0 < w, xin1 = b.x0 + 1 >= 0 && b.x0 + 1 < w;
    bool yin0 = b.y0 >= 0 && b.y0 < h, yin1 = b.y0 + 1 >= 0 && b.y0 + 1 < h;
    if (yin0 && xin0) { float4 t = __ldg(img + b.y0 * w + b.x0); float q = wy0 * wx0; r += t.x * q; g += t.y * q; bl += t.z * q; }
    if (yin0 && xin1) { float4 t = __ldg(img + b.y0 * w + b.x0 + 1); float q = wy0 * b.wx1; r += t.x * q; g += t.y * q; bl += t.z * q; }
    if (yin1 && xin0) { float4 t = __ldg(img + (b.y0 + 1) * w + b.x0); float q = b.wy1 * wx0; r += t.x * q; g += t.y * q; bl += t.z * q; }
    if (yin1 && xin1) { float4 t = __ldg(img + (b.y0 + 1) * w + b.x0 + 1); float q = b.wy1 * b.wx1; r += t.x * q; g += t.y * q; bl += t.z * q; }
    act_store4(hg, act_offset(hg, n, y, x, 4 * k), make_float4(hm, r, g, bl));
    s_sd[(k * 3 + 0) * w + x] = r;
    s_sd[(k * 3 + 1) * w + x] = g;
    s_sd[(k * 3 + 2) * w + x] = bl;
  }
  // zero the padding channels of the view (weights there are zero too, but keep NaNs out)
  const int pad4 = (hg.c - 4 * K1) / 4;
  for (int idx = threadIdx.x; idx < w * pad4; idx += blockDim.x) {
    int p = idx % pad4, x = idx / pad4;
    act_store4(hg, act_offset(hg, n, y, x, 4 * K1 + 4 * p), make_float4(0.f, 0.f, 0.f, 0.f));
  }
  __syncthreads();
  if (sdef != nullptr) {
    for (int idx = threadIdx.x; idx < K1 * 3 * w; idx += blockDim.x) {
      int x = idx % w, kc = idx / w;
      sdef[(((long long)n * K1 * 3 + kc) * h + y) * w + x] = s_sd[kc * w + x];
    }
  }
}

// =============================================================================================
// a8 epilogue  softmax(mask logits) -> mask; deformation = sum_k mask_k * z'_k; occlusion = sigmoid
// (dense_motion.py:98-111).  One thread per pixel; z'_k recomputed from the keypoints.
// =============================================================================================
__global__ void __launch_bounds__(128)
flow_combine_kernel(const float* __restrict__ logits, int ldl, KpDev kd, KpDev ks, int K, int has_occ,
                    int h, int w, float* __restrict__ mask, float2* __restrict__ deform,
                    float* __restrict__ occ) {
  __shared__ float s_kd[KP_MAX * 2], s_ks[KP_MAX * 2], s_J[KP_MAX * 4];
  const int n = blockIdx.y;
  const int K1 = K + 1;
  if (threadIdx.x < K) {
    int k = threadIdx.x;
    const float* vd = kd.value + (long long)n * kd.vs + k * 2;
    const float* vs = ks.value + (long long)n * ks.vs + k * 2;
    s_kd[2 * k] = vd[0]; s_kd[2 * k + 1] = vd[1];
    s_ks[2 * k] = vs[0]; s_ks[2 * k + 1] = vs[1];
    float J[4] = {1.f, 0.f, 0.f, 1.f};
    if (kd.jac != nullptr) kp_affine(kd.jac + (long long)n * kd.js + k * 4, ks.jac + (long long)n * ks.js + k * 4, J);
    s_J[4 * k] = J[0]; s_J[4 * k + 1] = J[1]; s_J[4 * k + 2] = J[2]; s_J[4 * k + 3] = J[3];
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= h * w) return;
  const int y = p / w, x = p % w;
  const float* lp = logits + ((long long)n * h * w + p) * ldl;
  float l[KP_MAX + 1];
  float m = -INFINITY;
  for (int k = 0; k < K1; ++k) { l[k] = __ldg(lp + k); m = fmaxf(m, l[k]); }
  float sum = 0.f;
  for (int k = 0; k < K1; ++k) { l[k] = expf(l[k] - m); sum += l[k]; }
  const float zx = grid_coord(x, w), zy = grid_coord(y, h);
  float dx = 0.f, dy = 0.f;
  for (int k = 0; k < K1; ++k) {
    float mk = l[k] / sum;
    mask[(((long long)n * K1 + k) * h + y) * w + x] = mk;
    float gx = zx, gy = zy;
    if (k > 0) {
      int kk = k - 1;
      float ddx = zx - s_kd[2 * kk], ddy = zy - s_kd[2 * kk + 1];
      float mx = ddx, my = ddy;
      if (kd.jac != nullptr) {
        mx = s_J[4 * kk] * ddx + s_J[4 * kk + 1] * ddy;
        my = s_J[4 * kk + 2] * ddx + s_J[4 * kk + 3] * ddy;
      }
      gx = mx + s_ks[2 * kk];
      gy = my + s_ks[2 * kk + 1];
    }
    dx += gx * mk;
    dy += gy * mk;
  }
  deform[(long long)n * h * w + p] = make_float2(dx, dy);
  if (has_occ) occ[(long long)n * h * w + p] = 1.f / (1.f + expf(-__ldg(lp + K1)));
}

// =============================================================================================
// a9-i  out = grid_sample(feat, deformation) * occlusion   (generator.py:57,79-84)
//       out2 = relu(out*scale2 + shift2)                    (next ResBlock2d norm1+relu, util.py:873-874)
// NHWC: each thread owns 4 channels of one pixel, so the 4 bilinear taps are four fully
// coalesced channel-vector reads.  Algorithmic traffic: one read + one (or two) writes of the map.
// =============================================================================================
// flow and occlusion of feature pixel (n, y, x): read directly when the motion grid equals the feature grid, else
// bilinearly resized on the fly (generator.py:53-56 and :82-83: F.interpolate(..., mode='bilinear'))
__device__ __forceinline__ float2 flow_at(const float2* __restrict__ deform, int fh, int fw, int n, int y, int x, int H, int W) {
  if (fh == H && fw == W) return __ldg(deform + ((long long)n * H + y) * W + x);
  return resize_flow(deform + (long long)n * fh * fw, fh, fw, y, x, H, W);
}
__device__ __forceinline__ float occ_at(const float* __restrict__ occ, int fh, int fw, int n, int y, int x, int H, int W) {
  if (fh == H && fw == W) return __ldg(occ + ((long long)n * H + y) * W + x);
  return resize_scalar(occ + (long long)n * fh * fw, fh, fw, y, x, H, W);
}

__global__ void __launch_bounds__(256)
warp_occlude_kernel(ActView feat, const float2* __restrict__ deform, const float* __restrict__ occ, int fh, int fw,
                    ActView out, ActView out2, int has_out2, const float* __restrict__ scale2,
                    const float* __restrict__ shift2, float* __restrict__ amax_out2, long long total) {
  const int c4 = feat.c >> 2;
  float amax = 0.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int cg = (int)(idx % c4);
    long long pix = idx / c4;
    int x = (int)(pix % feat.w);
    int y = (int)((pix / feat.w) % feat.h);
    int n = (int)(pix / ((long long)feat.w * feat.h));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (deform != nullptr) {
    float2 d = flow_at(deform, fh, fw, n, y, x, feat.h, feat.w);
    Bilinear b = bilinear_setup(d.x, d.y, feat.w, feat.h);
    float wx0 = 1.f - b.wx1, wy0 = 1.f - b.wy1;
    bool xin0 = b.x0 >= 0 && b.x0 < feat.w, xin1 = b.x0 + 1 >= 0 && b.x0 + 1 < feat.w;
    bool yin0 = b.y0 >= 0 && b.y0 < feat.h, yin1 = b.y0 + 1 >= 0 && b.y0 + 1 < feat.h;
    if (yin0 && xin0) { float4 t = act_load4(feat, act_offset(feat, n, b.y0, b.x0, 4 * cg)); float q = wy0 * wx0; acc.x += t.x * q; acc.y += t.y * q; acc.z += t.z * q; acc.w += t.w * q; }
    if (yin0 && xin1) { float4 t = act_load4(feat, act_offset(feat, n, b.y0, b.x0 + 1, 4 * cg)); float q = wy0 * b.wx1; acc.x += t.x * q; acc.y += t.y * q; acc.z += t.z * q; acc.w += t.w * q; }
    if (yin1 && xin0) { float4 t = act_load4(feat, act_offset(feat, n, b.y0 + 1, b.x0, 4 * cg)); float q = b.wy1 * wx0; acc.x += t.x * q; acc.y += t.y * q; acc.z += t.z * q; acc.w += t.w * q; }
    if (yin1 && xin1) { float4 t = act_load4(feat, act_offset(feat, n, b.y0 + 1, b.x0 + 1, 4 * cg)); float q = b.wy1 * b.wx1; acc.x += t.x * q; acc.y += t.y * q; acc.z += t.z * q; acc.w += t.w * q; }
    } else {
      acc = act_load4(feat, act_offset(feat, n, y, x, 4 * cg));     // no dense-motion network (generator.py:67): no warp
    }
    if (occ != nullptr) {
      float o = occ_at(occ, fh, fw, n, y, x, feat.h, feat.w);
      acc.x *= o; acc.y *= o; acc.z *= o; acc.w *= o;
    }
    act_store4(out, act_offset(out, n, y, x, 4 * cg), acc);
    if (has_out2) {
      float4 s = __ldg(reinterpret_cast<const float4*>(scale2) + cg);
      float4 t = __ldg(reinterpret_cast<const float4*>(shift2) + cg);
      float4 r;
      r.x = fmaxf(fmaf(acc.x, s.x, t.x), 0.f);
      r.y = fmaxf(fmaf(acc.y, s.y, t.y), 0.f);
      r.z = fmaxf(fmaf(acc.z, s.z, t.z), 0.f);
      r.w = fmaxf(fmaf(acc.w, s.w, t.w), 0.f);
      act_store4(out2, act_offset(out2, n, y, x, 4 * cg), r);
      amax = fmaxf(fmaxf(amax, fmaxf(r.x, r.y)), fmaxf(r.z, r.w));
    }
  }
  if (amax_out2 != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(amax_out2), __float_as_int(amax));   // non-negative floats
  }
}

// Vector fast path of a9-i: 8 channels per thread, so the per-pixel sampling setup is amortised over twice the
// data and every access is a 128-bit transaction.  Storage formats (template parameters): 0 = bf16, 1 = bf16 hi/lo
// planes, 2 = fp16, 3 = fp16 + e4m3 lo8 + e4m3 hi8 (mixed operand format, include/eamm_b200.h); feat and out
// share FIN, out2 (the next conv's A operand) has its own FOUT2.
constexpr int FMT_BF16 = 0, FMT_BF16X2 = 1, FMT_F16 = 2, FMT_MIX = 3;

__device__ __forceinline__ void bf16x8_fma(uint4 r, float q, float* acc) {
  float4 a = bf16x4_to_float4(make_uint2(r.x, r.y)), b = bf16x4_to_float4(make_uint2(r.z, r.w));
  acc[0] = fmaf(a.x, q, acc[0]); acc[1] = fmaf(a.y, q, acc[1]); acc[2] = fmaf(a.z, q, acc[2]); acc[3] = fmaf(a.w, q, acc[3]);
  acc[4] = fmaf(b.x, q, acc[4]); acc[5] = fmaf(b.y, q, acc[5]); acc[6] = fmaf(b.z, q, acc[6]); acc[7] = fmaf(b.w, q, acc[7]);
}
__device__ __forceinline__ void f16x8_fma(uint4 r, float q, float* acc) {
  const float2 a = f16x2_to_f32x2(r.x), b = f16x2_to_f32x2(r.y), c = f16x2_to_f32x2(r.z), d = f16x2_to_f32x2(r.w);
  acc[0] = fmaf(a.x, q, acc[0]); acc[1] = fmaf(a.y, q, acc[1]); acc[2] = fmaf(b.x, q, acc[2]); acc[3] = fmaf(b.y, q, acc[3]);
  acc[4] = fmaf(c.x, q, acc[4]); acc[5] = fmaf(c.y, q, acc[5]); acc[6] = fmaf(d.x, q, acc[6]); acc[7] = fmaf(d.y, q, acc[7]);
}
// acc += q * (8 channels starting at channel `ch` of the view at plane-0 element offset `off`)
template <int FMT>
__device__ __forceinline__ void vec8_tap(const ActView& v, long long off, int ch, float q, float* acc) {
  if (FMT == FMT_BF16 || FMT == FMT_BF16X2) {
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(v.data);
    bf16x8_fma(__ldg(reinterpret_cast<const uint4*>(base + off)), q, acc);
    if (FMT == FMT_BF16X2) bf16x8_fma(__ldg(reinterpret_cast<const uint4*>(base + off + v.c_buf)), q, acc);
  } else {
    const __half* base = static_cast<const __half*>(v.data);
    const float qs = q * v.inv_mul;
    f16x8_fma(__ldg(reinterpret_cast<const uint4*>(base + off)), qs, acc);
    if (FMT == FMT_MIX) {
      const uint2 r = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(base + off) + 2 * v.c_buf - (v.c_off + ch)));
      const float4 a = e4m3x4_to_f32x4(r.x), b = e4m3x4_to_f32x4(r.y);
      const float ql = qs * MIX_HI_GAIN;                       // lo8 / 64
      acc[0] = fmaf(a.x, ql, acc[0]); acc[1] = fmaf(a.y, ql, acc[1]); acc[2] = fmaf(a.z, ql, acc[2]); acc[3] = fmaf(a.w, ql, acc[3]);
      acc[4] = fmaf(b.x, ql, acc[4]); acc[5] = fmaf(b.y, ql, acc[5]); acc[6] = fmaf(b.z, ql, acc[6]); acc[7] = fmaf(b.w, ql, acc[7]);
    }
  }
}
template <int FMT>
__device__ __forceinline__ void vec8_store(const ActView& v, long long off, int ch, const float* f) {
  if (FMT == FMT_BF16 || FMT == FMT_BF16X2) {
    __nv_bfloat16* p = static_cast<__nv_bfloat16*>(v.data) + off;
    uint2 a = float4_to_bf16x4(make_float4(f[0], f[1], f[2], f[3]));
    uint2 b = float4_to_bf16x4(make_float4(f[4], f[5], f[6], f[7]));
    *reinterpret_cast<uint4*>(p) = make_uint4(a.x, a.y, b.x, b.y);
    if (FMT == FMT_BF16X2) {
      float4 ha = bf16x4_to_float4(a), hb = bf16x4_to_float4(b);
      uint2 la = float4_to_bf16x4(make_float4(f[0] - ha.x, f[1] - ha.y, f[2] - ha.z, f[3] - ha.w));
      uint2 lb = float4_to_bf16x4(make_float4(f[4] - hb.x, f[5] - hb.y, f[6] - hb.z, f[7] - hb.w));
      *reinterpret_cast<uint4*>(p + v.c_buf) = make_uint4(la.x, la.y, lb.x, lb.y);
    }
  } else {
    __half* p = static_cast<__half*>(v.data) + off;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = f[j] * v.mul;
    const uint4 h = make_uint4(f32x2_to_f16x2_sat(s[0], s[1]), f32x2_to_f16x2_sat(s[2], s[3]),
                               f32x2_to_f16x2_sat(s[4], s[5]), f32x2_to_f16x2_sat(s[6], s[7]));
    *reinterpret_cast<uint4*>(p) = h;
    if (FMT == FMT_MIX) {
      const float2 a = f16x2_to_f32x2(h.x), b = f16x2_to_f32x2(h.y), c = f16x2_to_f32x2(h.z), d = f16x2_to_f32x2(h.w);
      uint8_t* q = reinterpret_cast<uint8_t*>(p) + 2 * v.c_buf - (v.c_off + ch);
      *reinterpret_cast<uint2*>(q) = make_uint2(
          f32x4_to_e4m3x4_sat((s[0] - a.x) * MIX_LO_GAIN, (s[1] - a.y) * MIX_LO_GAIN, (s[2] - b.x) * MIX_LO_GAIN, (s[3] - b.y) * MIX_LO_GAIN),
          f32x4_to_e4m3x4_sat((s[4] - c.x) * MIX_LO_GAIN, (s[5] - c.y) * MIX_LO_GAIN, (s[6] - d.x) * MIX_LO_GAIN, (s[7] - d.y) * MIX_LO_GAIN));
      *reinterpret_cast<uint2*>(q + v.c_buf) = make_uint2(
          f32x4_to_e4m3x4_sat(a.x * MIX_HI_GAIN, a.y * MIX_HI_GAIN, b.x * MIX_HI_GAIN, b.y * MIX_HI_GAIN),
          f32x4_to_e4m3x4_sat(c.x * MIX_HI_GAIN, c.y * MIX_HI_GAIN, d.x * MIX_HI_GAIN, d.y * MIX_HI_GAIN));
    }
  }
}

template <int FIN, int FOUT2>
__global__ void __launch_bounds__(256)
warp_occlude_vec_kernel(ActView feat, const float2* __restrict__ deform, const float* __restrict__ occ, int fh, int fw,
                        ActView out, ActView out2, int has_out2, const float* __restrict__ scale2,
                        const float* __restrict__ shift2, float* __restrict__ amax_out2, long long total) {
  const int c8 = feat.c >> 3;
  float amax = 0.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % c8);
    const long long pix = idx / c8;
    const int x = (int)(pix % feat.w);
    const int y = (int)((pix / feat.w) % feat.h);
    const int n = (int)(pix / ((long long)feat.w * feat.h));
    const float2 d = flow_at(deform, fh, fw, n, y, x, feat.h, feat.w);
    const Bilinear b = bilinear_setup(d.x, d.y, feat.w, feat.h);
    const float wx0 = 1.f - b.wx1, wy0 = 1.f - b.wy1;
    const bool xin0 = b.x0 >= 0 && b.x0 < feat.w, xin1 = b.x0 + 1 >= 0 && b.x0 + 1 < feat.w;
    const bool yin0 = b.y0 >= 0 && b.y0 < feat.h, yin1 = b.y0 + 1 >= 0 && b.y0 + 1 < feat.h;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (yin0 && xin0) vec8_tap<FIN>(feat, act_offset(feat, n, b.y0, b.x0, 8 * cg), 8 * cg, wy0 * wx0, acc);
    if (yin0 && xin1) vec8_tap<FIN>(feat, act_offset(feat, n, b.y0, b.x0 + 1, 8 * cg), 8 * cg, wy0 * b.wx1, acc);
    if (yin1 && xin0) vec8_tap<FIN>(feat, act_offset(feat, n, b.y0 + 1, b.x0, 8 * cg), 8 * cg, b.wy1 * wx0, acc);
    if (yin1 && xin1) vec8_tap<FIN>(feat, act_offset(feat, n, b.y0 + 1, b.x0 + 1, 8 * cg), 8 * cg, b.wy1 * b.wx1, acc);
    if (occ != nullptr) {
      const float o = occ_at(occ, fh, fw, n, y, x, feat.h, feat.w);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= o;
    }
    vec8_store<FIN == FMT_MIX ? FMT_BF16X2 : FIN>(out, act_offset(out, n, y, x, 8 * cg), 8 * cg, acc);
    if (has_out2) {
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale2) + 2 * cg), s1 = __ldg(reinterpret_cast<const float4*>(scale2) + 2 * cg + 1);
      const float4 t0 = __ldg(reinterpret_cast<const float4*>(shift2) + 2 * cg), t1 = __ldg(reinterpret_cast<const float4*>(shift2) + 2 * cg + 1);
      float r[8];
      r[0] = fmaxf(fmaf(acc[0], s0.x, t0.x), 0.f); r[1] = fmaxf(fmaf(acc[1], s0.y, t0.y), 0.f);
      r[2] = fmaxf(fmaf(acc[2], s0.z, t0.z), 0.f); r[3] = fmaxf(fmaf(acc[3], s0.w, t0.w), 0.f);
      r[4] = fmaxf(fmaf(acc[4], s1.x, t1.x), 0.f); r[5] = fmaxf(fmaf(acc[5], s1.y, t1.y), 0.f);
      r[6] = fmaxf(fmaf(acc[6], s1.z, t1.z), 0.f); r[7] = fmaxf(fmaf(acc[7], s1.w, t1.w), 0.f);
      vec8_store<FOUT2>(out2, act_offset(out2, n, y, x, 8 * cg), 8 * cg, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) amax = fmaxf(amax, r[j]);
    }
  }
  if (amax_out2 != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(amax_out2), __float_as_int(amax));   // non-negative floats
  }
}

// =============================================================================================
// a9-ii  deformed = grid_sample(source, interpolate(deformation, (H,W), bilinear))  (generator.py:50-57,86)
// The flow upsample (align_corners=False: src = (dst+0.5)*h/H - 0.5, clamped at 0, upper tap
// clamped to h-1) is evaluated on the fly; source and output stay NCHW fp32 (API tensors).
// =============================================================================================
__global__ void __launch_bounds__(256)
warp_image_kernel(const float* __restrict__ src, long long src_n_stride, const float2* __restrict__ deform,
                  float* __restrict__ dst, int C, int H, int W, int h, int w) {
  const int n = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  float gx, gy;
  if (h == H && w == W) {
    float2 d = __ldg(deform + ((long long)n * h + y) * w + x);
    gx = d.x; gy = d.y;
  } else {
    const float2 g = resize_flow(deform + (long long)n * h * w, h, w, y, x, H, W);
    gx = g.x; gy = g.y;
  }
  Bilinear b = bilinear_setup(gx, gy, W, H);
  float wx0 = 1.f - b.wx1, wy0 = 1.f - b.wy1;
  bool xin0 = b.x0 >= 0 && b.x0 < W, xin1 = b.x0 + 1 >= 0 && b.x0 + 1 < W;
  bool yin0 = b.y0 >= 0 && b.y0 < H, yin1 = b.y0 + 1 >= 0 && b.y0 + 1 < H;
  const float* img = src + (long long)n * src_n_stride;
  for (int c = 0; c < C; ++c) {
    const float* pl = img + (long long)c * H * W;
    float acc = 0.f;
    if (yin0 && xin0) acc += __ldg(pl + b.y0 * W + b.x0) * (wy0 * wx0);
    if (yin0 && xin1) acc += __ldg(pl + b.y0 * W + b.x0 + 1) * (wy0 * b.wx1);
    if (yin1 && xin0) acc += __ldg(pl + (b.y0 + 1) * W + b.x0) * (b.wy1 * wx0);
    if (yin1 && xin1) acc += __ldg(pl + (b.y0 + 1) * W + b.x0 + 1) * (b.wy1 * b.wx1);
    dst[(((long long)n * C + c) * H + y) * W + x] = acc;
  }
}

// =============================================================================================
// source image NCHW fp32 -> NHWC activation view (zero-filled up to the view's channel count)
// =============================================================================================
__global__ void __launch_bounds__(256)
nchw_to_act_kernel(const float* __restrict__ src, int C, ActView dst, long long total) {
  const int c4 = dst.c >> 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long pix = idx;
    int x = (int)(pix % dst.w);
    int y = (int)((pix / dst.w) % dst.h);
    int n = (int)(pix / ((long long)dst.w * dst.h));
    for (int g = 0; g < c4; ++g) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int c = 4 * g + i;
        v[i] = c < C ? __ldg(src + (((long long)n * C + c) * dst.h + y) * dst.w + x) : 0.f;
      }
      act_store4(dst, act_offset(dst, n, y, x, 4 * g), make_float4(v[0], v[1], v[2], v[3]));
    }
  }
}

// =============================================================================================
// source image NCHW fp32 -> zero-bordered [n][H+6][W+8][8] bf16 (hi0,hi1,hi2,0,lo0,lo1,lo2,0):
// the operand of the packed 7x7 `first` convolution (one 16-byte pixel, 8 pixels = one K window).
// =============================================================================================
__global__ void __launch_bounds__(256)
pack_image_kernel(const float* __restrict__ src, int C, int H, int W, int split, uint4* __restrict__ dst,
                  long long total) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int x = (int)(idx % W);
    int y = (int)((idx / W) % H);
    int n = (int)(idx / ((long long)W * H));
    float v[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) v[c] = __ldg(src + (((long long)n * C + c) * H + y) * W + x);
    uint2 hi = float4_to_bf16x4(make_float4(v[0], v[1], v[2], 0.f));
    uint2 lo = make_uint2(0u, 0u);
    if (split == 2) hi = make_uint2(f32x2_to_f16x2_sat(v[0], v[1]), f32x2_to_f16x2_sat(v[2], 0.f));     // fp16, single plane
    else if (split) {
      float4 h = bf16x4_to_float4(hi);
      lo = float4_to_bf16x4(make_float4(v[0] - h.x, v[1] - h.y, v[2] - h.z, 0.f));
    }
    dst[((long long)n * (H + 6) + (y + 3)) * (W + 8) + (x + 3)] = make_uint4(hi.x, hi.y, lo.x, lo.y);
  }
}

// =============================================================================================
// keypoint heads  (keypoint_detector.py:40-50 gaussian2kp, :88-103 / :187-203)
//   heatmap_k = softmax(logit_k / T) over the (h-6+2pad) x (w-6+2pad) window of the same-padded 7x7 map,
//   value_k = sum heatmap_k * grid,  jacobian_k = sum heatmap_k * jacobian_map_k   (4 maps per keypoint)
// One block per image: each thread walks pixels of the window and keeps per-keypoint partial
// max / sums in registers; block reductions use warp shuffles + one shared-memory hop.
// =============================================================================================
constexpr int KH_MAX = 16;       // keypoints per image handled by the register arrays
constexpr int KH_THREADS = 256;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < KH_THREADS / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

__global__ void __launch_bounds__(KH_THREADS)
kp_head_kernel(const float* __restrict__ logits, int ldl, int h, int w, int K, int J, int off, int hh, int ww,
               float inv_t, float* __restrict__ heatmap, float* __restrict__ value, float* __restrict__ jac) {
  __shared__ float red[KH_THREADS / 32];
  __shared__ float s_max[KH_MAX], s_sum[KH_MAX];
  const int n = blockIdx.x;
  const float* img = logits + (long long)n * h * w * ldl;
  const int npix = hh * ww;
  // pass 1: per-keypoint maximum of logit / T
  float m[KH_MAX];
#pragma unroll
  for (int k = 0; k < KH_MAX; ++k) m[k] = -INFINITY;
  for (int p = threadIdx.x; p < npix; p += KH_THREADS) {
    const float* px = img + ((long long)(p / ww + off) * w + (p % ww + off)) * ldl;
#pragma unroll
    for (int k = 0; k < KH_MAX; ++k) if (k < K) m[k] = fmaxf(m[k], __ldg(px + k) * inv_t);
  }
#pragma unroll
  for (int k = 0; k < KH_MAX; ++k)
    if (k < K) { float r = block_reduce(m[k], red, true); if (threadIdx.x == 0) s_max[k] = r; }
  __syncthreads();
  // pass 2: sums of e, e*x, e*y and e*jacobian maps
  float se[KH_MAX], sx[KH_MAX], sy[KH_MAX], sj[KH_MAX][4];
#pragma unroll
  for (int k = 0; k < KH_MAX; ++k) { se[k] = sx[k] = sy[k] = 0.f; sj[k][0] = sj[k][1] = sj[k][2] = sj[k][3] = 0.f; }
  for (int p = threadIdx.x; p < npix; p += KH_THREADS) {
    const int y = p / ww, x = p % ww;
    const float* px = img + ((long long)(y + off) * w + (x + off)) * ldl;
    const float gx = grid_coord(x, ww), gy = grid_coord(y, hh);
#pragma unroll
    for (int k = 0; k < KH_MAX; ++k) {
      if (k < K) {
        const float e = expf(__ldg(px + k) * inv_t - s_max[k]);
        se[k] += e; sx[k] += e * gx; sy[k] += e * gy;
        if (J > 0) {
          const float* jm = px + K + 4 * (J == 1 ? 0 : k);
          sj[k][0] += e * __ldg(jm); sj[k][1] += e * __ldg(jm + 1); sj[k][2] += e * __ldg(jm + 2); sj[k][3] += e * __ldg(jm + 3);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KH_MAX; ++k) {
    if (k >= K) break;
    const float tot = block_reduce(se[k], red, false);
    const float vx = block_reduce(sx[k], red, false), vy = block_reduce(sy[k], red, false);
    if (threadIdx.x == 0) {
      s_sum[k] = tot;
      value[((long long)n * K + k) * 2] = vx / tot;
      value[((long long)n * K + k) * 2 + 1] = vy / tot;
    }
    if (J > 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float r = block_reduce(sj[k][c], red, false);
        if (threadIdx.x == 0) jac[((long long)n * K + k) * 4 + c] = r / tot;
      }
    }
  }
  __syncthreads();
  // pass 3: the normalised heatmap itself (an output of the module, keypoint_detector.py:90 / :189)
  if (heatmap != nullptr) {
    for (int idx = threadIdx.x; idx < K * npix; idx += KH_THREADS) {
      const int k = idx / npix, p = idx - k * npix;
      const float* px = img + ((long long)(p / ww + off) * w + (p % ww + off)) * ldl;
      heatmap[((long long)n * K + k) * npix + p] = expf(__ldg(px + k) * inv_t - s_max[k]) / s_sum[k];
    }
  }
}

// =============================================================================================
// per-clip keypoint glue  (filter1.py:14-47, demo.py:231-248, :263-271, :112-132)
// One block.  Phase A: one thread per scalar series (K*2 values + K*4 jacobian entries, same for the
// emotion keypoints) runs the One-Euro filter over the T frames -- the only sequential dependency of the
// clip.  Phase B: one thread per (frame, keypoint) adds the emotion rows and applies normalize_kp.
// Arithmetic mirrors the reference's float32 steps (separate multiplies/adds, IEEE division).
// =============================================================================================
struct OneEuroDev { float mincutoff, beta, alpha_d, one_minus_alpha_d, freq, scale, two_pi, te; };

__device__ __forceinline__ void one_euro_series(const float* __restrict__ in, float* __restrict__ out, int T, int stride,
                                                const OneEuroDev f) {
  float prev_x = 0.f, prev_dx = 0.f, prev_xf = 0.f;
  for (int t = 0; t < T; ++t) {
    const float x = __fmul_rn(in[(long long)t * stride], f.scale);
    float xf;
    if (t == 0) { xf = x; prev_dx = 0.f; }
    else {
      const float dx = __fmul_rn(__fsub_rn(x, prev_x), f.freq);
      const float edx = __fadd_rn(__fmul_rn(f.alpha_d, dx), __fmul_rn(f.one_minus_alpha_d, prev_dx));
      prev_dx = edx;
      const float cutoff = __fadd_rn(f.mincutoff, __fmul_rn(f.beta, fabsf(edx)));
      const float tau = __fdiv_rn(1.0f, __fmul_rn(f.two_pi, cutoff));
      const float a = __fdiv_rn(1.0f, __fadd_rn(1.0f, __fdiv_rn(tau, f.te)));
      xf = __fadd_rn(__fmul_rn(a, x), __fmul_rn(__fsub_rn(1.0f, a), prev_xf));
    }
    prev_x = x; prev_xf = xf;
    out[(long long)t * stride] = __fdiv_rn(xf, f.scale);
  }
}

__global__ void __launch_bounds__(256)
kp_clip_kernel(const float* __restrict__ drv_value, const float* __restrict__ drv_jac,
               const float* __restrict__ emo_value, const float* __restrict__ emo_jac, int T, int K, int Ke,
               OneEuroDev f_kp, OneEuroDev f_emo, const int* __restrict__ emo_rows, const float* __restrict__ emo_gain,
               int n_rows, const float* __restrict__ src_value, const float* __restrict__ src_jac,
               const float* __restrict__ init_value, const float* __restrict__ init_jac, float movement_scale,
               int relative, float* __restrict__ out_value, float* __restrict__ out_jac, float* __restrict__ emo_scratch) {
  // ---- phase A: filters (note filter1.py: the very first frame passes through, dx of frame 0 is 0 and is
  // fed to the dx low-pass as its initial state)
  const int nkp = K * 6, nemo = (emo_value != nullptr) ? Ke * 6 : 0;
  for (int e = threadIdx.x; e < nkp + nemo; e += blockDim.x) {
    if (e < nkp) {
      if (e < K * 2) one_euro_series(drv_value + e, out_value + e, T, K * 2, f_kp);
      else one_euro_series(drv_jac + (e - K * 2), out_jac + (e - K * 2), T, K * 4, f_kp);
    } else {
      const int q = e - nkp;
      if (q < Ke * 2) one_euro_series(emo_value + q, emo_scratch + q, T, Ke * 2, f_emo);
      else one_euro_series(emo_jac + (q - Ke * 2), emo_scratch + (long long)T * Ke * 2 + (q - Ke * 2), T, Ke * 4, f_emo);
    }
  }
  __syncthreads();
  // ---- phase B: emotion rows + normalize_kp
  for (int idx = threadIdx.x; idx < T * K; idx += blockDim.x) {
    const int t = idx / K, k = idx % K;
    float v[2] = {out_value[(long long)idx * 2], out_value[(long long)idx * 2 + 1]};
    float j[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) j[c] = out_jac[(long long)idx * 4 + c];
    if (nemo) {
      for (int r = 0; r < n_rows; ++r) {
        if (emo_rows[2 * r] == k) {
          const int s = emo_rows[2 * r + 1];
          const float g = emo_gain[r];
          const float* ev = emo_scratch + ((long long)t * Ke + s) * 2;
          const float* ej = emo_scratch + (long long)T * Ke * 2 + ((long long)t * Ke + s) * 4;
          v[0] = __fadd_rn(v[0], __fmul_rn(ev[0], g)); v[1] = __fadd_rn(v[1], __fmul_rn(ev[1], g));
#pragma unroll
          for (int c = 0; c < 4; ++c) j[c] = __fadd_rn(j[c], __fmul_rn(ej[c], g));
        }
      }
    }
    if (relative) {
      v[0] = __fadd_rn(__fmul_rn(__fsub_rn(v[0], init_value[k * 2]), movement_scale), src_value[k * 2]);
      v[1] = __fadd_rn(__fmul_rn(__fsub_rn(v[1], init_value[k * 2 + 1]), movement_scale), src_value[k * 2 + 1]);
      float M[4];
      kp_affine(init_jac + k * 4, j, M);                  // M = j * inv(init)
      const float* sj = src_jac + k * 4;
      j[0] = M[0] * sj[0] + M[1] * sj[2]; j[1] = M[0] * sj[1] + M[1] * sj[3];
      j[2] = M[2] * sj[0] + M[3] * sj[2]; j[3] = M[2] * sj[1] + M[3] * sj[3];
    }
    out_value[(long long)idx * 2] = v[0]; out_value[(long long)idx * 2 + 1] = v[1];
#pragma unroll
    for (int c = 0; c < 4; ++c) out_jac[(long long)idx * 4 + c] = j[c];
  }
}

static inline KpDev to_dev(const eamm_kp* k) {
  KpDev d; d.value = k->value; d.jac = k->jacobian; d.vs = k->value_stride; d.js = k->jacobian_stride;
  return d;
}

}  // namespace eamm

using namespace eamm;

extern "C" int eamm_abi_version(void) { return EAMM_ABI_VERSION; }

extern "C" int eamm_device_ok(int device) {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}

extern "C" int eamm_aa_downsample(const float* src, int64_t src_n_stride, float* dst, int n, int H, int W,
                                  int step, const float* g1, int taps, void* stream) {
  const int AA_TAPS = taps;
  if (!src || !dst || !g1 || n <= 0 || H <= 0 || W <= 0 || step <= 0) return EAMM_ERR_ARG;
  if (taps < 1 || taps > AA_MAX_TAPS || !(taps & 1)) return EAMM_ERR_UNSUPPORTED;
  if (H % step || W % step) return EAMM_ERR_SHAPE;
  if ((uintptr_t)dst % 16) return EAMM_ERR_ALIGN;
  int Ho = H / step, Wo = W / step;
  int rows = (AA_TR - 1) * step + AA_TAPS;
  size_t smem = (size_t)3 * rows * Wo * sizeof(float);
  if (smem > 200 * 1024) return EAMM_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(aa_downsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((Ho + AA_TR - 1) / AA_TR, n);
  ActView none = {};
  aa_downsample_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(src, src_n_stride, (float4*)dst, H, W, Ho, Wo, step, taps, g1, none, 0);
  EAMM_LAUNCH_CHECK();
  return 0;
}

extern "C" int eamm_aa_downsample_act(const float* src, int64_t src_n_stride, int n, int H, int W, int step,
                                      const float* g1, int taps, const eamm_act* dst, void* stream) {
  const int AA_TAPS = taps;
  if (!src || !g1 || n <= 0 || H <= 0 || W <= 0 || step <= 0) return EAMM_ERR_ARG;
  if (taps < 1 || taps > AA_MAX_TAPS || !(taps & 1)) return EAMM_ERR_UNSUPPORTED;
  if (H % step || W % step) return EAMM_ERR_SHAPE;
  int rc = check_view(dst); if (rc) return rc;
  int Ho = H / step, Wo = W / step;
  if (dst->n != n || dst->h != Ho || dst->w != Wo || dst->c < 4) return EAMM_ERR_SHAPE;
  int rows = (AA_TR - 1) * step + AA_TAPS;
  size_t smem = (size_t)3 * rows * Wo * sizeof(float);
  if (smem > 200 * 1024) return EAMM_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(aa_downsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((Ho + AA_TR - 1) / AA_TR, n);
  aa_downsample_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(src, src_n_stride, nullptr, H, W, Ho, Wo, step, taps, g1,
                                                                 make_view(dst), 1);
  EAMM_LAUNCH_CHECK();
  return 0;
}

extern "C" int eamm_kp_stage(const float* small_img, int64_t small_n_stride, const eamm_kp* kp_driving,
                             const eamm_kp* kp_source, int num_kp, float kp_variance, const eamm_act* hg_in,
                             float* sparse_deformed, int32_t* status, void* stream) {
  if (!small_img || !kp_driving || !kp_source || !kp_driving->value || !kp_source->value) return EAMM_ERR_ARG;
  int rc = check_view(hg_in);
  if (rc) return rc;
  if (num_kp <= 0 || num_kp > KP_MAX) return EAMM_ERR_UNSUPPORTED;
  if (hg_in->c < 4 * (num_kp + 1)) return EAMM_ERR_SHAPE;
  if ((kp_driving->jacobian == nullptr) != (kp_source->jacobian == nullptr)) return EAMM_ERR_ARG;
  if (hg_in->h < 2 || hg_in->w < 2) return EAMM_ERR_SHAPE;
  ActView hg = make_view(hg_in);
  size_t smem = (size_t)(num_kp + 1) * 3 * hg.w * sizeof(float);
  if (smem > 48 * 1024) return EAMM_ERR_UNSUPPORTED;
  dim3 grid(hg.h, hg.n);
  kp_stage_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const float4*)small_img, small_n_stride,
      to_dev(kp_driving), to_dev(kp_source), num_kp, kp_variance, hg, sparse_deformed, status);
  EAMM_LAUNCH_CHECK();
  return 0;
}

extern "C" int eamm_flow_combine(const float* logits, int ldl, const eamm_kp* kp_driving, const eamm_kp* kp_source,
                                 int num_kp, int has_occ, int n, int h, int w, float* mask, float* deformation,
                                 float* occlusion, void* stream) {
  if (!logits || !kp_driving || !kp_source || !mask || !deformation || n <= 0 || h < 2 || w < 2) return EAMM_ERR_ARG;
  if (has_occ && !occlusion) return EAMM_ERR_ARG;
  if (num_kp <= 0 || num_kp > KP_MAX) return EAMM_ERR_UNSUPPORTED;
  if (ldl < num_kp + 1 + (has_occ ? 1 : 0)) return EAMM_ERR_SHAPE;
  dim3 grid((h * w + 127) / 128, n);
  flow_combine_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(logits, ldl, to_dev(kp_driving), to_dev(kp_source),
      num_kp, has_occ, h, w, mask, (float2*)deformation, occlusion);
  EAMM_LAUNCH_CHECK();
  return 0;
}

static int view_fmt(const ActView& v) {
  if (v.dtype == EAMM_BF16) return v.planes == 2 ? FMT_BF16X2 : FMT_BF16;
  if (v.dtype == EAMM_F16) return v.planes == 2 ? FMT_MIX : FMT_F16;
  return -1;
}

extern "C" int eamm_warp_occlude(const eamm_act* feat, const float* deformation, const float* occlusion, int fh, int fw,
                                 const eamm_act* out, const eamm_act* out2, const float* scale2,
                                 const float* shift2, float* amax_out2, void* stream) {
  int rc = check_view(feat); if (rc) return rc;
  rc = check_view(out); if (rc) return rc;
  if (fh <= 0 || fw <= 0) { fh = feat->h; fw = feat->w; }
  if (!deformation && occlusion) return EAMM_ERR_ARG;         // an occlusion map only exists next to a flow
  if (out->n != feat->n || out->h != feat->h || out->w != feat->w || out->c != feat->c) return EAMM_ERR_SHAPE;
  ActView f = make_view(feat), o = make_view(out), o2 = o;
  int has2 = 0;
  if (out2) {
    rc = check_view(out2); if (rc) return rc;
    if (!scale2 || !shift2) return EAMM_ERR_ARG;
    if (out2->n != feat->n || out2->h != feat->h || out2->w != feat->w || out2->c != feat->c) return EAMM_ERR_SHAPE;
    o2 = make_view(out2); has2 = 1;
  }
  auto vec_ok = [](const ActView& v) {
    return v.dtype != EAMM_F32 && v.c % 8 == 0 && v.c_off % 8 == 0 && v.c_buf % 8 == 0 &&
           ((uintptr_t)v.data % 16) == 0 && v.n_stride % 8 == 0;
  };
  const int fin = view_fmt(f), fo = view_fmt(o), fo2 = has2 ? view_fmt(o2) : fin;
  // the vector kernel's format pairs: out has feat's format (a mixed feat writes bf16 hi/lo); out2 the same or mixed
  const bool pair_ok = fin >= 0 && fo == (fin == FMT_MIX ? FMT_BF16X2 : fin) && (fo2 == fo || fo2 == FMT_MIX) &&
                       !(fo2 == FMT_MIX && fin != FMT_BF16X2 && fin != FMT_MIX);
  if (deformation && pair_ok && vec_ok(f) && vec_ok(o) && (!has2 || vec_ok(o2))) {
    long long total = (long long)f.n * f.h * f.w * (f.c / 8);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    const float2* dp = (const float2*)deformation;
    cudaStream_t st = (cudaStream_t)stream;
#define EAMM_WO(FI, FO2) warp_occlude_vec_kernel<FI, FO2><<<blocks, 256, 0, st>>>(f, dp, occlusion, fh, fw, o, o2, has2, scale2, shift2, amax_out2, total)
    if (fin == FMT_BF16) EAMM_WO(FMT_BF16, FMT_BF16);
    else if (fin == FMT_F16) EAMM_WO(FMT_F16, FMT_F16);
    else if (fin == FMT_BF16X2 && fo2 == FMT_MIX) EAMM_WO(FMT_BF16X2, FMT_MIX);
    else if (fin == FMT_BF16X2) EAMM_WO(FMT_BF16X2, FMT_BF16X2);
    else if (fo2 == FMT_MIX) EAMM_WO(FMT_MIX, FMT_MIX);
    else EAMM_WO(FMT_MIX, FMT_BF16X2);
#undef EAMM_WO
    EAMM_LAUNCH_CHECK();
    return 0;
  }
  long long total = (long long)f.n * f.h * f.w * (f.c / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  warp_occlude_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(f, (const float2*)deformation, occlusion, fh, fw, o, o2, has2,
                                                               scale2, shift2, amax_out2, total);
  EAMM_LAUNCH_CHECK();
  return 0;
}

extern "C" int eamm_warp_image(const float* src, int64_t src_n_stride, const float* deformation, float* dst,
                               int n, int C, int H, int W, int h, int w, void* stream) {
  if (!src || !deformation || !dst || n <= 0 || C <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0) return EAMM_ERR_ARG;
  dim3 grid((W + 255) / 256, H, n);
  warp_image_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_n_stride, (const float2*)deformation, dst, C, H, W, h, w);
  EAMM_LAUNCH_CHECK();
  return 0;
}

extern "C" int eamm_nchw_to_act(const float* src, int n, int C, int H, int W, const eamm_act* dst, void* stream) {
  if (!src || n <= 0 || C <= 0) return EAMM_ERR_ARG;
  int rc = check_view(dst); if (rc) return rc;
  if (dst->n != n || dst->h != H || dst->w != W || dst->c < C) return EAMM_ERR_SHAPE;
  ActView d = make_view(dst);
  long long total = (long long)n * H * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  nchw_to_act_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, C, d, total);
  EAMM_LAUNCH_CHECK();
  return 0;
}

extern "C" int eamm_pack_image(const float* src, int n, int C, int H, int W, int split, void* dst, void* stream) {
  if (!src || !dst || n <= 0 || C <= 0 || C > 3 || H <= 0 || W <= 0 || split < 0 || split > 2) return EAMM_ERR_ARG;
  if ((uintptr_t)dst % 16) return EAMM_ERR_ALIGN;
  long long total = (long long)n * H * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_image_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, C, H, W, split, (uint4*)dst, total);
  EAMM_LAUNCH_CHECK();
  return 0;
}

extern "C" int eamm_kp_head(const float* logits, int ldl, int n, int h, int w, int num_kp, int num_jac_maps, int pad,
                            float temperature, float* heatmap, float* value, float* jacobian, void* stream) {
  if (!logits || !value || n <= 0 || h <= 0 || w <= 0 || !(temperature > 0.f)) return EAMM_ERR_ARG;
  if (num_kp <= 0 || num_kp > KH_MAX) return EAMM_ERR_UNSUPPORTED;
  if (num_jac_maps != 0 && num_jac_maps != 1 && num_jac_maps != num_kp) return EAMM_ERR_ARG;
  if (num_jac_maps && !jacobian) return EAMM_ERR_ARG;
  if (pad < 0 || pad > 3) return EAMM_ERR_UNSUPPORTED;
  if (ldl < num_kp + 4 * num_jac_maps) return EAMM_ERR_SHAPE;
  const int off = 3 - pad, hh = h - 2 * off, ww = w - 2 * off;
  if (hh < 2 || ww < 2) return EAMM_ERR_SHAPE;
  kp_head_kernel<<<n, KH_THREADS, 0, (cudaStream_t)stream>>>(logits, ldl, h, w, num_kp, num_jac_maps, off, hh, ww,
                                                              1.f / temperature, heatmap, value, jacobian);
  EAMM_LAUNCH_CHECK();
  return 0;
}

static OneEuroDev one_euro_dev(const eamm_one_euro* f) {
  // compute_alpha(dcutoff) is evaluated in double by the reference (python floats) and then multiplies fp32 tensors
  const double te = 1.0 / (double)f->freq;
  const double tau = 1.0 / (2.0 * 3.141592653589793 * (double)f->dcutoff);
  const double a_d = 1.0 / (1.0 + tau / te);
  OneEuroDev d;
  d.mincutoff = f->mincutoff; d.beta = f->beta; d.alpha_d = (float)a_d; d.one_minus_alpha_d = (float)(1.0 - a_d);
  d.freq = f->freq; d.scale = f->scale; d.two_pi = (float)(2.0 * 3.141592653589793); d.te = (float)te;
  return d;
}

extern "C" int eamm_kp_clip(const float* drv_value, const float* drv_jac, const float* emo_value, const float* emo_jac,
                            int T, int K, int Ke, const eamm_one_euro* f_kp, const eamm_one_euro* f_emo,
                            const int32_t* emo_rows, const float* emo_gain, int n_emo_rows, const float* src_value,
                            const float* src_jac, const float* init_value, const float* init_jac, float movement_scale,
                            int relative, float* out_value, float* out_jac, float* emo_scratch, void* stream) {
  if (!drv_value || !drv_jac || !f_kp || !out_value || !out_jac || T <= 0 || K <= 0) return EAMM_ERR_ARG;
  if (relative && (!src_value || !src_jac || !init_value || !init_jac)) return EAMM_ERR_ARG;
  if ((emo_value == nullptr) != (emo_jac == nullptr)) return EAMM_ERR_ARG;
  if (emo_value && (!f_emo || !emo_scratch || Ke <= 0 || n_emo_rows < 0 || (n_emo_rows && (!emo_rows || !emo_gain)))) return EAMM_ERR_ARG;
  OneEuroDev fe = emo_value ? one_euro_dev(f_emo) : one_euro_dev(f_kp);
  kp_clip_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(drv_value, drv_jac, emo_value, emo_jac, T, K, Ke, one_euro_dev(f_kp), fe,
                                                      emo_rows, emo_gain, emo_value ? n_emo_rows : 0, src_value, src_jac,
                                                      init_value, init_jac, movement_scale, relative, out_value, out_jac,
                                                      emo_scratch);
  EAMM_LAUNCH_CHECK();
  return 0;
}
