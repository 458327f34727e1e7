// Shared device/host helpers for the eamm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/eamm_b200.h"

#define EAMM_LAUNCH_CHECK()                                   \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    if (e__ != cudaSuccess) return (int)e__;                  \
  } while (0)

namespace eamm {

// Device-side copy of an eamm_act with the strides precomputed.
struct ActView {
  void* data;
  int dtype, n, h, w, c, c_off, c_buf, planes;
  long long n_stride;
  int pix_stride;  // planes * c_buf
};

inline ActView make_view(const eamm_act* a) {
  ActView v;
  v.data = a->data; v.dtype = a->dtype; v.n = a->n; v.h = a->h; v.w = a->w; v.c = a->c;
  v.c_off = a->c_off; v.c_buf = a->c_buf; v.planes = a->planes; v.n_stride = a->n_stride;
  v.pix_stride = a->planes * a->c_buf;
  return v;
}

inline int check_view(const eamm_act* a) {
  if (!a || !a->data) return EAMM_ERR_ARG;
  if (a->n <= 0 || a->h <= 0 || a->w <= 0 || a->c <= 0 || a->c_buf <= 0) return EAMM_ERR_ARG;
  if (a->c_off < 0 || a->c_off + a->c > a->c_buf) return EAMM_ERR_SHAPE;
  if (a->dtype == EAMM_F32) { if (a->planes != 1) return EAMM_ERR_DTYPE; }
  else if (a->dtype == EAMM_BF16) { if (a->planes != 1 && a->planes != 2) return EAMM_ERR_DTYPE; }
  else return EAMM_ERR_DTYPE;
  if (a->c_off % 4 || a->c_buf % 4) return EAMM_ERR_ALIGN;   // 4-channel vector access everywhere
  return 0;
}

// Offset (elements) of (n,y,x, plane 0, channel ch of the view).
__device__ __forceinline__ long long act_offset(const ActView& v, int n, int y, int x, int ch) {
  return (long long)n * v.n_stride + ((long long)y * v.w + x) * v.pix_stride + v.c_off + ch;
}

// ---- 4-channel vector load/store in any storage mode; values are fp32 in registers ----------
__device__ __forceinline__ float4 bf16x4_to_float4(uint2 r) {
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ uint2 float4_to_bf16x4(float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  return r;
}

__device__ __forceinline__ float4 act_load4(const ActView& v, long long off) {
  if (v.dtype == EAMM_F32) {
    return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(v.data) + off));
  }
  const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(v.data) + off;
  float4 r = bf16x4_to_float4(__ldg(reinterpret_cast<const uint2*>(p)));
  if (v.planes == 2) {
    float4 lo = bf16x4_to_float4(__ldg(reinterpret_cast<const uint2*>(p + v.c_buf)));
    r.x += lo.x; r.y += lo.y; r.z += lo.z; r.w += lo.w;
  }
  return r;
}

__device__ __forceinline__ void act_store4(const ActView& v, long long off, float4 val) {
  if (v.dtype == EAMM_F32) {
    *reinterpret_cast<float4*>(static_cast<float*>(v.data) + off) = val;
    return;
  }
  __nv_bfloat16* p = static_cast<__nv_bfloat16*>(v.data) + off;
  uint2 hi = float4_to_bf16x4(val);
  *reinterpret_cast<uint2*>(p) = hi;
  if (v.planes == 2) {
    float4 h = bf16x4_to_float4(hi);
    float4 rem = make_float4(val.x - h.x, val.y - h.y, val.z - h.z, val.w - h.w);
    *reinterpret_cast<uint2*>(p + v.c_buf) = float4_to_bf16x4(rem);
  }
}

// ---- sampling arithmetic shared by every warp kernel ----------------------------------------
// grid_sample(align_corners=False) un-normalisation: pixel = ((g + 1) * size - 1) / 2
// (ATen grid_sampler_unnormalize; call sites dense_motion.py:77, generator.py:57).
__device__ __forceinline__ float unnormalize(float g, int size) {
  return ((g + 1.f) * (float)size - 1.f) * 0.5f;
}

// make_coordinate_grid (util.py:847-848): 2*(i/(n-1)) - 1, evaluated in the same op order.
__device__ __forceinline__ float grid_coord(int i, int n) {
  return 2.f * ((float)i / (float)(n - 1)) - 1.f;
}

struct Bilinear {
  int x0, y0;          // floor taps (x0+1, y0+1 are the others)
  float wx1, wy1;      // weights of the +1 taps; 1-w for the floor taps
};
__device__ __forceinline__ Bilinear bilinear_setup(float gx, float gy, int W, int H) {
  float ix = unnormalize(gx, W), iy = unnormalize(gy, H);
  float fx = floorf(ix), fy = floorf(iy);
  Bilinear b;
  // clamp before the int conversion so huge/NaN coordinates stay defined (they sample zeros)
  b.x0 = (int)fminf(fmaxf(fx, -2.f), (float)W + 1.f);
  b.y0 = (int)fminf(fmaxf(fy, -2.f), (float)H + 1.f);
  b.wx1 = ix - fx;
  b.wy1 = iy - fy;
  if (!(ix == ix) || !(iy == iy)) { b.x0 = -2; b.y0 = -2; b.wx1 = 0.f; b.wy1 = 0.f; }
  return b;
}

// skimage.img_as_ubyte of a float32 image value in [0,1]: rint(v * 255) (round half to even), clipped to [0,255]
__device__ __forceinline__ unsigned char to_ubyte(float v) {
  return (unsigned char)fminf(fmaxf(rintf(__fmul_rn(v, 255.f)), 0.f), 255.f);
}

// J = J_s * inv(J_d) for one keypoint (dense_motion.py:56); closed-form 2x2 inverse.
// Returns false when J_d is singular (torch.inverse raises there).
__device__ __forceinline__ bool kp_affine(const float* jd, const float* js, float* J) {
  float a = jd[0], b = jd[1], c = jd[2], d = jd[3];
  float det = a * d - b * c;
  bool ok = (det != 0.f) && (det == det) && (fabsf(det) <= 3.0e38f);
  float r = 1.f / det;
  float i00 = d * r, i01 = -b * r, i10 = -c * r, i11 = a * r;
  J[0] = js[0] * i00 + js[1] * i10;
  J[1] = js[0] * i01 + js[1] * i11;
  J[2] = js[2] * i00 + js[3] * i10;
  J[3] = js[2] * i01 + js[3] * i11;
  return ok;
}

}  // namespace eamm
