// Shared device/host helpers for the eamm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>
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
  float mul, inv_mul;  // EAMM_F16: stored = value * mul (2^scale_exp), value = stored * inv_mul; 1 otherwise
};

inline ActView make_view(const eamm_act* a) {
  ActView v;
  v.data = a->data; v.dtype = a->dtype; v.n = a->n; v.h = a->h; v.w = a->w; v.c = a->c;
  v.c_off = a->c_off; v.c_buf = a->c_buf; v.planes = a->planes; v.n_stride = a->n_stride;
  v.pix_stride = a->planes * a->c_buf;
  const int e = a->dtype == EAMM_F16 ? a->scale_exp : 0;
  v.mul = ldexpf(1.f, e); v.inv_mul = ldexpf(1.f, -e);
  return v;
}

inline int check_view(const eamm_act* a) {
  if (!a || !a->data) return EAMM_ERR_ARG;
  if (a->n <= 0 || a->h <= 0 || a->w <= 0 || a->c <= 0 || a->c_buf <= 0) return EAMM_ERR_ARG;
  if (a->c_off < 0 || a->c_off + a->c > a->c_buf) return EAMM_ERR_SHAPE;
  if (a->dtype == EAMM_F32) { if (a->planes != 1) return EAMM_ERR_DTYPE; }
  else if (a->dtype == EAMM_BF16) { if (a->planes != 1 && a->planes != 2) return EAMM_ERR_DTYPE; }
  else if (a->dtype == EAMM_F16) {
    if (a->planes != 1 && a->planes != 2) return EAMM_ERR_DTYPE;
    if (a->scale_exp < -100 || a->scale_exp > 100) return EAMM_ERR_ARG;
  }
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

// ---- fp16 / e4m3 packs (saturating: an out-of-range value clamps instead of becoming inf / NaN) ----------
__device__ __forceinline__ uint32_t f32x2_to_f16x2_sat(float lo, float hi) {      // lo -> bits [0,16)
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 f16x2_to_f32x2(uint32_t r) {
  return __half22float2(*reinterpret_cast<const __half2*>(&r));
}
__device__ __forceinline__ uint32_t f32x4_to_e4m3x4_sat(float a, float b, float c, float d) {   // a -> byte 0
  uint16_t lo, hi;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
  return (uint32_t)lo | ((uint32_t)hi << 16);
}
__device__ __forceinline__ float4 e4m3x4_to_f32x4(uint32_t r) {
  uint32_t h0, h1;
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h0) : "h"((uint16_t)(r & 0xffffu)));
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h1) : "h"((uint16_t)(r >> 16)));
  const float2 a = f16x2_to_f32x2(h0), b = f16x2_to_f32x2(h1);
  return make_float4(a.x, a.y, b.x, b.y);
}
// byte address of the e4m3 planes of a mixed (EAMM_F16, planes == 2) view: lo8 of channel ch of the pixel whose
// plane-0 element offset is `off` (= act_offset(..., ch)); hi8 follows c_buf bytes later
__device__ __forceinline__ uint8_t* mix_lo8_ptr(const ActView& v, long long off, int ch_abs) {
  // off = pixel_base + c_off + ch (fp16 elements); pixel_base bytes = 2 * (off - ch_abs); plane 1 starts 2*c_buf bytes in
  return static_cast<uint8_t*>(v.data) + 2 * (off - ch_abs) + 2 * v.c_buf + ch_abs;
}
constexpr float MIX_LO_GAIN = 64.f;          // lo8 = e4m3((v' - hi) * 64), hi8 = e4m3(hi / 64): the products hi8*lo8 carry
constexpr float MIX_HI_GAIN = 1.f / 64.f;    // the scale of hi*hi, so all three terms share one accumulator

__device__ __forceinline__ float4 act_load4(const ActView& v, long long off) {
  if (v.dtype == EAMM_F32) {
    return __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(v.data) + off));
  }
  if (v.dtype == EAMM_F16) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(static_cast<const __half*>(v.data) + off));
    const float2 a = f16x2_to_f32x2(q.x), b = f16x2_to_f32x2(q.y);
    float4 r = make_float4(a.x, a.y, b.x, b.y);
    if (v.planes == 2) {
      // (the channel index inside the buffer is recovered from the offset: image and pixel bases are multiples of pix_stride)
      const int ch_abs = (int)(off % v.pix_stride);
      const float4 lo = e4m3x4_to_f32x4(__ldg(reinterpret_cast<const uint32_t*>(mix_lo8_ptr(v, off, ch_abs))));
      r.x += lo.x * MIX_HI_GAIN; r.y += lo.y * MIX_HI_GAIN; r.z += lo.z * MIX_HI_GAIN; r.w += lo.w * MIX_HI_GAIN;
    }
    r.x *= v.inv_mul; r.y *= v.inv_mul; r.z *= v.inv_mul; r.w *= v.inv_mul;
    return r;
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
  if (v.dtype == EAMM_F16) {
    const float4 s = make_float4(val.x * v.mul, val.y * v.mul, val.z * v.mul, val.w * v.mul);
    const uint2 hi = make_uint2(f32x2_to_f16x2_sat(s.x, s.y), f32x2_to_f16x2_sat(s.z, s.w));
    *reinterpret_cast<uint2*>(static_cast<__half*>(v.data) + off) = hi;
    if (v.planes == 2) {
      const float2 a = f16x2_to_f32x2(hi.x), b = f16x2_to_f32x2(hi.y);
      uint8_t* lo8 = mix_lo8_ptr(v, off, (int)(off % v.pix_stride));
      *reinterpret_cast<uint32_t*>(lo8) = f32x4_to_e4m3x4_sat((s.x - a.x) * MIX_LO_GAIN, (s.y - a.y) * MIX_LO_GAIN,
                                                              (s.z - b.x) * MIX_LO_GAIN, (s.w - b.y) * MIX_LO_GAIN);
      *reinterpret_cast<uint32_t*>(lo8 + v.c_buf) = f32x4_to_e4m3x4_sat(a.x * MIX_HI_GAIN, a.y * MIX_HI_GAIN,
                                                                        b.x * MIX_HI_GAIN, b.y * MIX_HI_GAIN);
    }
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
  // clamp before the int conversion so huge coordinates stay defined (they sample zeros, as in ATen)
  b.x0 = (int)fminf(fmaxf(fx, -2.f), (float)W + 1.f);
  b.y0 = (int)fminf(fmaxf(fy, -2.f), (float)H + 1.f);
  b.wx1 = ix - fx;
  b.wy1 = iy - fy;
  // NaN / +-inf coordinates: F.grid_sample's tap weights become NaN (inf - inf) and the output is NaN
  // (measured on the reference's CPU path); an in-range tap with NaN weights reproduces that.
  if (!(fabsf(ix) <= 3.0e38f) || !(fabsf(iy) <= 3.0e38f)) {
    b.x0 = 0; b.y0 = 0;
    b.wx1 = __int_as_float(0x7fc00000); b.wy1 = b.wx1;
  }
  return b;
}

// F.interpolate(mode='bilinear', align_corners=False) source taps for destination index d (ATen
// area_pixel_compute_source_index: src = scale * (d + 0.5) - 0.5 clamped at 0, upper tap clamped to in - 1);
// call sites generator.py:55 (flow -> feature / image grid) and generator.py:83 (occlusion map).
struct Resize1D { int i0, i1; float w1; };
__device__ __forceinline__ Resize1D resize_taps(int d, int in, int out) {
  const float scale = (float)in / (float)out;
  float f = scale * ((float)d + 0.5f) - 0.5f;
  if (f < 0.f) f = 0.f;
  Resize1D r;
  r.i0 = (int)f;
  r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
  r.w1 = f - (float)r.i0;
  return r;
}
// bilinear resize of a [h,w] map of float2 / float at destination pixel (y, x) of an [H,W] grid (same op order as ATen:
// w0 * (wx0 * a + wx1 * b) + w1 * (wx0 * c + wx1 * d))
__device__ __forceinline__ float2 resize_flow(const float2* __restrict__ dp, int h, int w, int y, int x, int H, int W) {
  const Resize1D ry = resize_taps(y, h, H), rx = resize_taps(x, w, W);
  const float ly0 = 1.f - ry.w1, lx0 = 1.f - rx.w1;
  const float2 d00 = __ldg(dp + ry.i0 * w + rx.i0), d01 = __ldg(dp + ry.i0 * w + rx.i1);
  const float2 d10 = __ldg(dp + ry.i1 * w + rx.i0), d11 = __ldg(dp + ry.i1 * w + rx.i1);
  float2 g;
  g.x = ly0 * (lx0 * d00.x + rx.w1 * d01.x) + ry.w1 * (lx0 * d10.x + rx.w1 * d11.x);
  g.y = ly0 * (lx0 * d00.y + rx.w1 * d01.y) + ry.w1 * (lx0 * d10.y + rx.w1 * d11.y);
  return g;
}
__device__ __forceinline__ float resize_scalar(const float* __restrict__ p, int h, int w, int y, int x, int H, int W) {
  const Resize1D ry = resize_taps(y, h, H), rx = resize_taps(x, w, W);
  const float ly0 = 1.f - ry.w1, lx0 = 1.f - rx.w1;
  return ly0 * (lx0 * __ldg(p + ry.i0 * w + rx.i0) + rx.w1 * __ldg(p + ry.i0 * w + rx.i1)) +
         ry.w1 * (lx0 * __ldg(p + ry.i1 * w + rx.i0) + rx.w1 * __ldg(p + ry.i1 * w + rx.i1));
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
