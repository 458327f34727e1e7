// fp32 CUDA-core implicit-GEMM convolution with the fused epilogues of the EAMM generator.
// This is the exact-fp32 parity path ("fp32_simt") and the on-device checker for the tcgen05
// kernel; it implements every conv flavour of the hot path:
//   3x3 pad 1 (+BN folded, ReLU, 2x2 avg-pool)        DownBlock2d      util.py:915-920
//   nearest-x2 + 3x3 as four parity-class 2x2 convs    UpBlock2d        util.py:895-900
//   3x3 + residual + fused next norm1/relu             ResBlock2d       util.py:872-880
//   7x7 pad 3 (+sigmoid, NCHW fp32 out)                first/final/mask generator.py:25,46; dense_motion.py:18,21
//
// GEMM view: M = pixels (tiles of 16 2x2 quads = 64 pixels), N = cout (tiles of 64),
// K = taps x cin (steps of 16 channels of one tap).  256 threads, 4 pixels (one quad) x 4 couts each.
#include "common.cuh"

namespace eamm {

constexpr int SM_BM = 64, SM_BN = 64, SM_BK = 16;

struct ConvSimtParams {
  ActView in, out, out2, res;
  const float* w;      // [classes][taps][cin][cout]
  const float* bias;
  const float* scale2;
  const float* shift2;
  float* out_nchw;
  float* out_nhwc;
  unsigned char* out_u8;
  int kind, flags, cin, cout, taps, ksize, out_nchw_c;
  int has_out, has_out2, has_res;
  long long quads;     // n * ceil(h/2) * ceil(w/2)
};

__global__ void __launch_bounds__(256)
conv_simt_kernel(ConvSimtParams p) {
  __shared__ __align__(16) float As[SM_BK][SM_BM + 4];
  __shared__ __align__(16) float Bs[SM_BK][SM_BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int H = p.in.h, W = p.in.w, hq = (H + 1) >> 1, wq = (W + 1) >> 1;   // odd maps: the last quads hang over the edge
  const int cls = blockIdx.z;                    // parity class for UP2 (a = cls>>1, b = cls&1)
  const int pa = cls >> 1, pb = cls & 1;
  const int n0 = blockIdx.y * SM_BN;

  // A-load role: pixel m_ld = tid/4 of the tile, channel group (tid%4)*4 of the 16-channel step.
  const int m_ld = tid >> 2, cg_ld = (tid & 3) * 4;
  long long q_ld = (long long)blockIdx.x * 16 + (m_ld >> 2);
  bool ld_valid = q_ld < p.quads;
  int ld_n = 0, ld_y = 0, ld_x = 0;
  if (ld_valid) {
    int qx = (int)(q_ld % wq);
    int qy = (int)((q_ld / wq) % hq);
    ld_n = (int)(q_ld / ((long long)wq * hq));
    ld_y = qy * 2 + ((m_ld & 3) >> 1);
    ld_x = qx * 2 + (m_ld & 1);
  }
  // B-load role: row tid/16 of the step, 4 couts at (tid%16)*4.
  const int b_row = tid >> 4, b_col = (tid & 15) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const float* wcls = p.w + (long long)cls * p.taps * p.cin * p.cout;
  for (int t = 0; t < p.taps; ++t) {
    int dy, dx;
    if (p.kind == EAMM_CONV_UP2_3X3) { dy = pa - 1 + (t >> 1); dx = pb - 1 + (t & 1); }
    else { dy = t / p.ksize - (p.ksize >> 1); dx = t % p.ksize - (p.ksize >> 1); }
    const int sy = ld_y + dy, sx = ld_x + dx;
    const bool in_img = ld_valid && sy >= 0 && sy < H && sx >= 0 && sx < W;
    for (int c0 = 0; c0 < p.cin; c0 += SM_BK) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in_img && c0 + cg_ld < p.cin) a = act_load4(p.in, act_offset(p.in, ld_n, sy, sx, c0 + cg_ld));
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + b_row < p.cin && n0 + b_col < p.cout)
        b = __ldg(reinterpret_cast<const float4*>(wcls + ((long long)t * p.cin + c0 + b_row) * p.cout + n0 + b_col));
      __syncthreads();
      As[cg_ld + 0][m_ld] = a.x; As[cg_ld + 1][m_ld] = a.y; As[cg_ld + 2][m_ld] = a.z; As[cg_ld + 3][m_ld] = a.w;
      *reinterpret_cast<float4*>(&Bs[b_row][b_col]) = b;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < SM_BK; ++kk) {
        float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
    }
  }

  // ---------------------------------------------------------------- epilogue
  const int co = n0 + tx * 4;
  long long q = (long long)blockIdx.x * 16 + ty;
  if (q >= p.quads || co >= p.cout) return;
  const int qx = (int)(q % wq), qy = (int)((q / wq) % hq), n = (int)(q / ((long long)wq * hq));
  const float4 bias = __ldg(reinterpret_cast<const float4*>(p.bias + co));
  const float bb[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[i][j] + bb[j];
      if (p.flags & EAMM_EPI_RELU) v = fmaxf(v, 0.f);
      acc[i][j] = v;
    }
  int npix = 4;
  if (p.flags & EAMM_EPI_POOL2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][j] = 0.25f * (acc[0][j] + acc[1][j] + acc[2][j] + acc[3][j]);
    npix = 1;
  }
  float4 s2 = make_float4(0, 0, 0, 0), t2 = s2;
  if (p.has_out2) {
    s2 = __ldg(reinterpret_cast<const float4*>(p.scale2 + co));
    t2 = __ldg(reinterpret_cast<const float4*>(p.shift2 + co));
  }
  for (int i = 0; i < npix; ++i) {
    int oy, ox, OH, OW;
    if (p.flags & EAMM_EPI_POOL2) { oy = qy; ox = qx; OH = hq; OW = wq; }
    else {
      int y = qy * 2 + (i >> 1), x = qx * 2 + (i & 1);
      if (y >= H || x >= W) continue;
      if (p.kind == EAMM_CONV_UP2_3X3) { oy = 2 * y + pa; ox = 2 * x + pb; OH = 2 * H; OW = 2 * W; }
      else { oy = y; ox = x; OH = H; OW = W; }
    }
    float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (p.has_res) {
      float4 r = act_load4(p.res, act_offset(p.res, n, oy, ox, co));
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (p.has_out) act_store4(p.out, act_offset(p.out, n, oy, ox, co), v);
    if (p.has_out2) {
      float4 r;
      r.x = fmaxf(fmaf(v.x, s2.x, t2.x), 0.f);
      r.y = fmaxf(fmaf(v.y, s2.y, t2.y), 0.f);
      r.z = fmaxf(fmaf(v.z, s2.z, t2.z), 0.f);
      r.w = fmaxf(fmaf(v.w, s2.w, t2.w), 0.f);
      act_store4(p.out2, act_offset(p.out2, n, oy, ox, co), r);
    }
    if (p.out_nhwc != nullptr)
      *reinterpret_cast<float4*>(p.out_nhwc + (((long long)n * OH + oy) * OW + ox) * p.cout + co) = v;
    if (p.out_nchw != nullptr) {
      float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (co + j < p.out_nchw_c) {
          float o = vv[j];
          if (p.flags & EAMM_EPI_SIGMOID) o = 1.f / (1.f + expf(-o));
          p.out_nchw[(((long long)n * p.out_nchw_c + co + j) * OH + oy) * OW + ox] = o;
          if (p.out_u8 != nullptr) p.out_u8[(((long long)n * OH + oy) * OW + ox) * p.out_nchw_c + co + j] = to_ubyte(o);
        }
      }
    }
  }
}

// Shared argument validation for every conv implementation.  Returns 0 or an EAMM_ERR_*.
int conv_check_args(const eamm_conv_args* a, int cout_align) {
  if (!a || !a->in || !a->weight || !a->bias) return EAMM_ERR_ARG;
  int rc = check_view(a->in); if (rc) return rc;
  if (a->kind < EAMM_CONV_3X3 || a->kind > EAMM_CONV_ROW7_PACKED) return EAMM_ERR_UNSUPPORTED;
  if (a->cin != a->in->c || a->cout <= 0 || a->cout % cout_align) return EAMM_ERR_SHAPE;
  const bool pool = a->flags & EAMM_EPI_POOL2;
  if (((a->in->h & 1) || (a->in->w & 1)) && pool) return EAMM_ERR_SHAPE;     // only pooling needs even maps
  if (pool && a->kind == EAMM_CONV_UP2_3X3) return EAMM_ERR_UNSUPPORTED;
  int OH = a->in->h, OW = a->in->w;
  if (pool) { OH >>= 1; OW >>= 1; }
  if (a->kind == EAMM_CONV_UP2_3X3) { OH *= 2; OW *= 2; }
  const eamm_act* outs[3] = {a->out, a->out2, a->residual};
  for (int i = 0; i < 3; ++i) {
    if (!outs[i]) continue;
    rc = check_view(outs[i]); if (rc) return rc;
    if (outs[i]->n != a->in->n || outs[i]->h != OH || outs[i]->w != OW) return EAMM_ERR_SHAPE;
    if (outs[i]->c != a->cout) return EAMM_ERR_SHAPE;
  }
  if (a->out2 && (!a->scale2 || !a->shift2)) return EAMM_ERR_ARG;
  if (a->out_nchw && (a->out_nchw_c <= 0 || a->out_nchw_c > a->cout)) return EAMM_ERR_SHAPE;
  if ((a->flags & EAMM_EPI_SIGMOID) && !a->out_nchw) return EAMM_ERR_UNSUPPORTED;
  if (a->out_u8_nhwc && !a->out_nchw) return EAMM_ERR_ARG;
  if (!a->out && !a->out2 && !a->out_nchw && !a->out_nhwc_f32) return EAMM_ERR_ARG;
  return 0;
}

}  // namespace eamm

using namespace eamm;

extern "C" int eamm_conv_simt(const eamm_conv_args* a, void* stream) {
  int rc = conv_check_args(a, 4);
  if (rc) return rc;
  if (a->kind == EAMM_CONV_ROW7_PACKED) return EAMM_ERR_UNSUPPORTED;
  ConvSimtParams p;
  p.in = make_view(a->in);
  p.has_out = a->out != nullptr; p.has_out2 = a->out2 != nullptr; p.has_res = a->residual != nullptr;
  p.out = p.has_out ? make_view(a->out) : p.in;
  p.out2 = p.has_out2 ? make_view(a->out2) : p.in;
  p.res = p.has_res ? make_view(a->residual) : p.in;
  p.w = static_cast<const float*>(a->weight);
  p.bias = a->bias; p.scale2 = a->scale2; p.shift2 = a->shift2;
  p.out_nchw = a->out_nchw; p.out_nhwc = a->out_nhwc_f32; p.out_nchw_c = a->out_nchw_c;
  p.out_u8 = a->out_u8_nhwc;
  p.kind = a->kind; p.flags = a->flags; p.cin = a->cin; p.cout = a->cout;
  p.ksize = a->kind == EAMM_CONV_7X7 ? 7 : 3;
  p.taps = a->kind == EAMM_CONV_UP2_3X3 ? 4 : p.ksize * p.ksize;
  p.quads = (long long)p.in.n * ((p.in.h + 1) >> 1) * ((p.in.w + 1) >> 1);
  long long tiles = (p.quads + 15) / 16;
  if (tiles > 0x7fffffffLL) return EAMM_ERR_UNSUPPORTED;
  dim3 grid((unsigned)tiles, (p.cout + SM_BN - 1) / SM_BN, a->kind == EAMM_CONV_UP2_3X3 ? 4 : 1);
  conv_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  EAMM_LAUNCH_CHECK();
  return 0;
}
