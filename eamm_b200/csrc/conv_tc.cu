// tcgen05 implicit-GEMM convolution for sm_100a: TMA-staged NHWC tiles, UMMA (bf16 x bf16 -> fp32
// accumulators in TMEM), warp-specialised persistent CTAs, fused epilogues.
//
//   GEMM view   D[128 pixels, BN couts] += A[128 pixels, 64 ch] * B[BN couts, 64 ch]^T
//   A tile      one 4-D TMA box {64 ch, bw, bh, bn} (bw*bh*bn = 128 pixels) of the NHWC input,
//               shifted by the filter tap (dy, dx); out-of-image pixels are zero-filled by TMA,
//               which is exactly the conv's zero padding.  SWIZZLE_128B, K-major.
//   B tile      2-D TMA box {64 k, BN rows} of the host-packed weight matrix [couts][K], K-major.
//   K loop      passes x taps x (cin/64).  passes = 1 (bf16) or 3 (split-bf16 "fp32" mode:
//               a_hi*b_lo + a_lo*b_hi + a_hi*b_hi with activations/weights stored as hi/lo planes;
//               small terms first, see the producer).
//   roles       warps 0-7 = epilogue (TMEM lane quadrant = warp % 4, two warps per quadrant splitting
//               the accumulator columns), warps 8-10 = TMA producers (round-robin over the ring
//               stages), warp 11 = MMA issuer (+TMEM alloc).
//   pipelines   smem ring (full/empty mbarriers, 1-2 K chunks per stage) and a 2-deep TMEM accumulator
//               ring so the epilogue of tile i overlaps the main loop of tile i+1.
//   variants    UP2        nearest-x2 + 3x3 as four parity classes of 2x2 taps, scattered by the epilogue
//               halo row   7x7: one 134-pixel row per ky, the 7 kx taps are row-shifted smem views
//               kx-in-N    7x7 -> <=4 NCHW channels: the 7 kx taps are GEMM columns, epilogue sums shifts
//               row7       7x7 over a packed <=3-channel image: K window = 8 pixels x (hi,lo) channels
//               fold       split mode, cout <= 128: weight planes stacked along N (2 A loads, not 3)
//               kx-in-N x4 `final` with four output rows per tile: N = (dr, kx, co) = 112, K walks 10 input rows
//               kx-in-N fw mask+occlusion logits (W <= 128, 16 couts): N = (kx, co) = 112, full-width tiles
//               split-K    layers with too few N = 256 tiles for the chip (hourglass 8x8 ... 2x2): partial tiles through
//                          an fp32 workspace, the last-arriving CTA reduces in split order and runs the epilogue
//               CTA pair   N tile 256 and folded layers: tcgen05.mma.cta_group::2, M = 256, half of B per CTA
//               halo tile  3x3 / UP2 with fp16 or mixed operands on large maps: 8 x 16-pixel tiles, ONE 10 x 18 halo tile per
//                          128-byte K chunk in shared memory, the filter taps are shifted descriptor views of it (every conv
//                          of the step was bound by the ~12 TB/s L2 -> SM rate of re-staging the A tile per tap); weights
//                          stream through their own ring; UP2 keeps the four parity classes of a tile in one work item
//   switches    EAMM_TC_HALO / _FOLD / _CTA2 = 0 disable a variant (_CTA2 is a bit mask: 1 pairs, 2 folded
//               pairs, 4 narrow unfolded pairs -- default 19; bit 4 = 16: pairs for mixed fp16+fp8 layers with N < 256),
//               EAMM_TC_KXW (bit 0: 7x7 scheme 3, bit 1: scheme 4, bit 2: compact scheme-3 epilogue buffer; default 7), EAMM_TC_SPLITK = 0 / EAMM_TC_ST256 = 0 switch
//               split-K / the 32-byte epilogue stores off, EAMM_TC_SPLITK_MAX caps the split factor (9), EAMM_TC_BNCOST = 0 restores the older N-tile rule,
//               EAMM_TC_NUM_SMS lets eamm_conv_tc_query plan on a host without a GPU, EAMM_TC_KSUB / _CTA2_KSUB force the
//               chunks per stage, EAMM_TC_PROF = 1 prints per-role cycle counters, EAMM_TC_DEBUG = 1..6
//               switches TMA / MMA / epilogue off (timing experiments; results are garbage).
// Epilogues: folded-BN bias, ReLU, 2x2 avg-pool (DownBlock2d), parity scatter (UpBlock2d as four
// 2x2 convs), residual add + fused next norm1/ReLU (ResBlock2d), fp32 NHWC logits, sigmoid NCHW.
// Reference call sites: util.py:872-880, 895-900, 915-920, 934-938; generator.py:92-93;
// dense_motion.py:98,110.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "common.cuh"

namespace eamm {

int conv_check_args(const eamm_conv_args* a, int cout_align);   // conv_simt.cu

constexpr int TC_EPI_WARPS = 8;                 // epilogue warps: TMEM lane quadrant = warp % 4, column half = warp / 4
constexpr int TC_PRODUCERS = 3;                 // TMA producer warps (warps TC_EPI_WARPS .. +TC_PRODUCERS-1)
constexpr int TC_MMA_WARP = TC_EPI_WARPS + TC_PRODUCERS;   // single MMA-issuing warp, highest warp id
constexpr int TC_THREADS = 32 * (TC_MMA_WARP + 1);
constexpr int TC_A_BYTES = 128 * 128;          // 128 pixels x 64 bf16
constexpr uint32_t TC_TMEM_COLS = 512;

struct ConvTcParams {
  int N, H, W;
  int bw, bh, bn, bw_log2, bh_log2;
  int tiles_x, tiles_y, tiles_n, n_tiles, classes;
  int BN, kind, flags, taps, ksize, cin_chunks, passes;
  int a_c_off, a_c_buf;
  int halo;            // 7x7 with a 134-pixel halo row per ky: the 7 kx taps are shifted smem views
  int kxn;             // 7x7 -> <=4 NCHW channels: the 7 kx taps live in the N dimension (N = 7*4 -> 32),
                       // one MMA group per (ky, pass, chunk); the epilogue sums the kx-shifted columns.
                       // kxn == 2: four output rows per tile as well -- N = (dr, kx, co) = 4*7*4 = 112, the K loop walks
                       // the 10 input rows y0-3 .. y0+6 (B row block j holds w[ky = j - dr], zero outside the filter):
                       // 80 wide MMAs per 4 rows instead of 224 narrow ones (these layers are bound by the
                       // per-instruction cost of fetching the 128-pixel A slab, not by the math).
                       // kxn == 3: full-width tiles of a <=128-wide map with <=16 fp32 NHWC couts (mask+occlusion):
                       // N = (kx, co) = 7*16 = 112, tile = (128/W) whole rows, no x halo (it is the zero padding)
  int x_stride;        // pixels between consecutive x tiles (122 in kxn mode 1/2, else bw)
  int y_stride;        // rows between consecutive y tiles (4 for kxn == 2, else bh)
  int st256;           // epilogue activation stores / residual loads as 32-byte pieces (every view 32-byte aligned)
  int s_nc;            // kxn == 2: channels kept per (dr, kx) group in the epilogue's smem buffer (4, or out_nchw_c: compact)
  int ntap;            // K-loop taps: 7 (halo, kxn 1/3), 10 (kxn 2), else taps
  int splitk;          // > 1: split-K.  `splitk` consecutive work items share one output tile, each runs 1/splitk of the
                       // K loop (contiguous in the (pass, tap, chunk) order), publishes its fp32 partial tile to sk_ws
                       // and bumps the tile's counter; the item that arrives last sums the partials in split order
                       // (deterministic) and runs the normal epilogue.  For the 8x8 ... 2x2 hourglass layers, which
                       // otherwise need N tiles of 32 columns to occupy the chip and then pay the per-MMA A-slab cost
                       // 8x more often than an N = 256 tile would.
  float* sk_ws;        // [out tile][split][128][BN] fp32 partials
  unsigned int* sk_cnt;// [out tile] arrival counters, zero before and after every launch
  int fold;            // split mode with 2*BN <= 256: the weight planes are stacked along N.  Chunk type 0 =
                       // a_hi x [b_hi; b_lo] (N = 2*BN), type 1 = a_lo x b_hi (N = BN); the epilogue adds
                       // accumulator columns [BN, 2BN) (the a_hi*b_lo cross term) to [0, BN).  2 A loads and
                       // 8 MMAs per (tap, 64 channels) instead of 3 and 12.  fold == 2: every chunk is type 0
                       // (packed `first` conv: both activation planes already sit in one K window).
  int b_rows_total;    // rows of one weight plane block (classes * cout): the lo block starts there
  int cta2;            // CTA pairs with cta_group::2 MMAs (tiles 2i, 2i+1 = adjacent M tiles of one class / N tile)
  int ksub;            // 64-channel K chunks per pipeline stage (1..4)
  int chunk_shift;     // log2(cin_chunks) (cin/64 is a power of two for every layer of the path)
  int debug;           // EAMM_TC_DEBUG: 1 = no TMA (MMA side alone), 2 = no MMA (TMA side alone); timing only
  unsigned long long* prof;  // EAMM_TC_PROF: per-CTA cycle counters [grid][8] (bring-up instrumentation)
  int a_slot_bytes;    // bytes reserved for the A operand in a stage
  int num_stages, cout;
  int has_out, has_out2, has_res;
  ActView out, out2, res;
  const float* bias; const float* scale2; const float* shift2;
  float* out_nchw; int out_nchw_c; float* out_nhwc; unsigned char* out_u8;
  long long total_tiles;
  int lean;            // the specialised MMA issue loop (mma_issuer_lean); 0 = the generic loop (EAMM_TC_LEAN=0, halo, INSTR)
  int f16in;           // the A/B operands are fp16 (EAMM_F16 input): tensor-map coordinates are BYTES (uint8 maps)
  int mix;             // EAMM_F16 two-plane input: K loop = [a_hi8 x w_lo8 | a_lo8 x w_hi8] as kind::f8f6f4 steps over 128-channel
                       // chunks (n8 = ntap * cin/128 chunks each), then a_hi x w_hi as kind::f16 steps over 64-channel chunks
  int nf8;             // mix: fp8 chunks at the head of the K loop (2 * ntap * cin/128, or ntap for mix64)
  int mix64;           // mix with cin == c_buf == 64: plane 1 of a pixel, [lo8 x 64 | hi8 x 64], is ONE 128-byte K chunk that
                       // meets the weight row [w_hi8 x 64 | w_lo8 x 64]: both cross terms in one fp8 chunk per tap
  int ah;              // halo-tile scheme (see the header): A ring of `ah_na` slots (one 10 x 18-pixel halo tile of a 128-byte K
                       // chunk each) + weight ring of `num_stages` stages (`ksub` taps each); tap (ty, tx) reads the view
                       // slot + (ty * 10 + tx) * 128 B with 8-row groups (= image rows of the 8 x 16 tile) 1280 B apart
  int ah_g;            // UP2: parity classes per work item (accumulator columns cls * BN); 1 otherwise
  int cls_groups;      // class groups per M tile (classes / ah_g)
  int b_res;           // the whole weight matrix of the CTA stays in shared memory (packed first conv: 7 chunks = 112 KB, loaded
                       // once; the ring then carries A only -- that layer ran at the L2 -> SM cap re-staging 16 KB of weights per
                       // 16 KB of pixels); barrier bars[44]
  int tma_st;          // fast epilogue stores through shared memory + TMA (cp.async.bulk.tensor store): a lane owns a pixel, so its
                       // 32-byte global stores touched 32 different lines per warp instruction and kept the LSU busy for ~50 % of a
                       // narrow layer's time, with the next chunk's bias loads queued behind them.  Each epilogue warp stages its
                       // 32 pixels x 32 channels (4 KB, swizzled: conflict-free 16-byte shared stores) and one lane issues the bulk
                       // stores.  st_base = offset of the staging area (8 warps x 2 x 4 KB) from the ring base.
  int st_base;
  int st_stride;       // bytes of staging per epilogue warp: 4 KB (out) or 8 KB (out + out2)
  int epi_fast;        // fast epilogue variant (epilogue_fast): 0 = generic, 1 = plain, 2 = pooled, 3 = residual, 4 = residual + out2, 5 / 6 = plain / pooled with folded weight planes
  int dec_shift;       // >= 0: decode_tile by shifts, log2 of (n_tiles, cls_groups, tiles_x, tiles_y) in 5-bit fields; -1: divisions
  int ah_na;           // A ring slots
  int ah_spc;          // weight stages per K chunk (= ah_g * ntap / ksub)
  int ah_nsec;         // K sections of a tile, each {first A byte column, 128-byte chunks, kind::f8f6f4?, first weight byte column}
  int ah_cbase[3], ah_nch[3], ah_f8[3], ah_bcol[3];
  const float* acc_scale;   // [cout] accumulator multiplier (undoes the operand pre-scales), or null
  float* amax_out; float* amax_out2;   // running max |value| of out / out2 (calibration statistic), or null
};

// ------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done;
}
// Slow path, out of line so that the hot loops stay a few hundred bytes of code.  Bounded: a
// protocol bug traps (launch error) instead of hanging the GPU.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 28); ++it)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// ---- cta_group::2 (CTA pair) variants -------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {     // own smem address -> peer CTA's
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Relaxed: the arrive only tells the leader's MMA warp that this warp's tcgen05.ld reads of the accumulator have completed
// (tcgen05.wait::ld + tcgen05.fence::before_thread_sync precede it); no generic-proxy data is handed over.  A release at
// cluster scope made the lane wait for all of its outstanding global stores first (ERRBAR + membar: 13-21 % of the warp
// samples of down0 / up1 in profiles/r2_ncu_summary.md).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster,
                                             int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc2_commit_mc(uint32_t bar) {       // arrive on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc2_mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// kind::f8f6f4 (e4m3 x e4m3, K = 32 per instruction): the cross terms of the fp16 + fp8 scheme
__device__ __forceinline__ void tc2_mma_f8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_mma_f8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// Output tensor maps of the TMA-store epilogue: 5-D byte maps {pixel bytes, column parity, x, row parity, image * H + y} of `out`
// and `out2` (parities = the UP2 scatter; 1 otherwise), one with 64-byte boxes (fp16 / bf16 planes, SWIZZLE_64B) and one with
// 32-byte boxes (the e4m3 planes of the mixed format, SWIZZLE_32B).
struct StoreMaps { CUtensorMap o64, o32, p64, p32; };
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// K-major SWIZZLE_128B smem matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(=1)<<16 |
// SBO(=1024B>>4)<<32 | version 1 <<46 | layout SWIZZLE_128B(2) <<61.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// The same descriptor from a pre-shifted address (smem byte address >> 4; the ring lives below 256 KB, so no masking):
// what the lean MMA issuer keeps as running 32-bit counters instead of re-deriving descriptors from byte addresses.
constexpr uint64_t SW128_DESC_BITS = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
__device__ __forceinline__ uint64_t sw128_desc16(uint32_t addr16) { return SW128_DESC_BITS | (uint64_t)addr16; }

// One operand chunk (128-byte rows) = four K slices 32 bytes apart: K = 16 per kind::f16 step, K = 32 per kind::f8f6f4 step.
// `acc` = 0 overwrites the accumulator with the first slice (first chunk of a tile).
template <bool CTA2, bool F8>
__device__ __forceinline__ void issue_chunk(uint32_t tmem_acc, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (CTA2) {
    if (F8) {
      tc2_mma_f8(tmem_acc, da, db, idesc, acc); tc2_mma_f8(tmem_acc, da + 2, db + 2, idesc, 1u);
      tc2_mma_f8(tmem_acc, da + 4, db + 4, idesc, 1u); tc2_mma_f8(tmem_acc, da + 6, db + 6, idesc, 1u);
    } else {
      tc2_mma_bf16(tmem_acc, da, db, idesc, acc); tc2_mma_bf16(tmem_acc, da + 2, db + 2, idesc, 1u);
      tc2_mma_bf16(tmem_acc, da + 4, db + 4, idesc, 1u); tc2_mma_bf16(tmem_acc, da + 6, db + 6, idesc, 1u);
    }
  } else {
    if (F8) {
      tc_mma_f8(tmem_acc, da, db, idesc, acc); tc_mma_f8(tmem_acc, da + 2, db + 2, idesc, 1u);
      tc_mma_f8(tmem_acc, da + 4, db + 4, idesc, 1u); tc_mma_f8(tmem_acc, da + 6, db + 6, idesc, 1u);
    } else {
      tc_mma_bf16(tmem_acc, da, db, idesc, acc); tc_mma_bf16(tmem_acc, da + 2, db + 2, idesc, 1u);
      tc_mma_bf16(tmem_acc, da + 4, db + 4, idesc, 1u); tc_mma_bf16(tmem_acc, da + 6, db + 6, idesc, 1u);
    }
  }
}

template <int CH> struct TmemLd;
template <> struct TmemLd<32> {
  __device__ static __forceinline__ void ld(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
  }
};
template <> struct TmemLd<16> {
  __device__ static __forceinline__ void ld(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
  }
};

// bf16 vector store/load of CH consecutive channels (hi plane, and lo plane when planes == 2)
// 256-bit global store / load (sm_100: STG.E.ENL2.256): one full 32-byte sector per lane
__device__ __forceinline__ void stg256(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}

// L2-coherent (.cg) 256-bit variants for the split-K partial tiles (written by one CTA, read by another)
__device__ __forceinline__ void stg256_cg(void* p, const uint32_t* v) {
  asm volatile("st.global.cg.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void ldg256_cg(const void* p, float* v) {
  asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p) : "memory");
}

// EAMM_F16 views (one fp16 plane, or fp16 + e4m3 lo8 + e4m3 hi8): stored = value * v.mul, saturating packs.
// The mixed format needs the 32-byte path with CH == 32 (one 32-byte store per e4m3 plane and chunk).
// e4m3(hi / 64) of four fp16 values straight from their f16x2 pairs (the product by 2^-6 is exact in fp16 wherever the
// e4m3 result is non-zero, so this equals rounding hi * 2^-6 from fp32)
__device__ __forceinline__ uint32_t f16x4_to_hi8x4(uint32_t h01, uint32_t h23) {
  uint32_t s01, s23;
  uint16_t lo, hi;
  asm("mul.f16x2 %0, %1, %2;" : "=r"(s01) : "r"(h01), "r"(0x24002400u));       // 0x2400 = 2^-6
  asm("mul.f16x2 %0, %1, %2;" : "=r"(s23) : "r"(h23), "r"(0x24002400u));
  asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(lo) : "r"(s01));
  asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(hi) : "r"(s23));
  return (uint32_t)lo | ((uint32_t)hi << 16);
}

// Eight consecutive channels of one pixel (the pooled epilogue: a lane owns an 8-channel slice).  `off` = element offset
// of channel `ch` of the view (act_offset), 8-element aligned.
__device__ __forceinline__ void store8(const ActView& v, long long off, int ch, const float* f) {
  if (v.dtype == EAMM_F16) {
    __half* p = static_cast<__half*>(v.data) + off;
    uint32_t h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = f32x2_to_f16x2_sat(f[2 * j] * v.mul, f[2 * j + 1] * v.mul);
    *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
    if (v.planes == 2) {
      uint32_t lo8[2], hi8[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float2 a = f16x2_to_f32x2(h[2 * j]), b = f16x2_to_f32x2(h[2 * j + 1]);
        lo8[j] = f32x4_to_e4m3x4_sat((f[4 * j] * v.mul - a.x) * MIX_LO_GAIN, (f[4 * j + 1] * v.mul - a.y) * MIX_LO_GAIN,
                                     (f[4 * j + 2] * v.mul - b.x) * MIX_LO_GAIN, (f[4 * j + 3] * v.mul - b.y) * MIX_LO_GAIN);
        hi8[j] = f16x4_to_hi8x4(h[2 * j], h[2 * j + 1]);
      }
      uint8_t* q = reinterpret_cast<uint8_t*>(p) + 2 * v.c_buf - (v.c_off + ch);
      *reinterpret_cast<uint2*>(q) = make_uint2(lo8[0], lo8[1]);
      *reinterpret_cast<uint2*>(q + v.c_buf) = make_uint2(hi8[0], hi8[1]);
    }
    return;
  }
  __nv_bfloat16* p = static_cast<__nv_bfloat16*>(v.data) + off;
  const uint2 a = float4_to_bf16x4(make_float4(f[0], f[1], f[2], f[3]));
  const uint2 b = float4_to_bf16x4(make_float4(f[4], f[5], f[6], f[7]));
  *reinterpret_cast<uint4*>(p) = make_uint4(a.x, a.y, b.x, b.y);
  if (v.planes == 2) {
    const float4 ha = bf16x4_to_float4(a), hb = bf16x4_to_float4(b);
    const uint2 la = float4_to_bf16x4(make_float4(f[0] - ha.x, f[1] - ha.y, f[2] - ha.z, f[3] - ha.w));
    const uint2 lb = float4_to_bf16x4(make_float4(f[4] - hb.x, f[5] - hb.y, f[6] - hb.z, f[7] - hb.w));
    *reinterpret_cast<uint4*>(p + v.c_buf) = make_uint4(la.x, la.y, lb.x, lb.y);
  }
}

template <int CH>
__device__ __forceinline__ void store_chunk_f16(const ActView& v, long long off, int ch, const float* f, bool wide) {
  __half* p = static_cast<__half*>(v.data) + off;
  uint32_t h[CH / 2];
#pragma unroll
  for (int j = 0; j < CH / 2; ++j) h[j] = f32x2_to_f16x2_sat(f[2 * j] * v.mul, f[2 * j + 1] * v.mul);
  if (wide) {
#pragma unroll
    for (int g = 0; g < CH / 16; ++g)
      stg256(p + 16 * g, make_uint4(h[8 * g], h[8 * g + 1], h[8 * g + 2], h[8 * g + 3]),
             make_uint4(h[8 * g + 4], h[8 * g + 5], h[8 * g + 6], h[8 * g + 7]));
  } else {
#pragma unroll
    for (int g = 0; g < CH / 8; ++g)
      *reinterpret_cast<uint4*>(p + 8 * g) = make_uint4(h[4 * g], h[4 * g + 1], h[4 * g + 2], h[4 * g + 3]);
  }
  if (v.planes == 2 && CH == 32) {
    uint32_t lo8[CH / 4], hi8[CH / 4];
#pragma unroll
    for (int j = 0; j < CH / 4; ++j) {
      const float2 a = f16x2_to_f32x2(h[2 * j]), b = f16x2_to_f32x2(h[2 * j + 1]);
      lo8[j] = f32x4_to_e4m3x4_sat((f[4 * j] * v.mul - a.x) * MIX_LO_GAIN, (f[4 * j + 1] * v.mul - a.y) * MIX_LO_GAIN,
                                   (f[4 * j + 2] * v.mul - b.x) * MIX_LO_GAIN, (f[4 * j + 3] * v.mul - b.y) * MIX_LO_GAIN);
      hi8[j] = f16x4_to_hi8x4(h[2 * j], h[2 * j + 1]);
    }
    // plane 1 of the pixel starts 2*c_buf bytes after plane 0 and holds [c_buf lo8 bytes | c_buf hi8 bytes]:
    // p points at fp16 element (c_off + ch) of plane 0, i.e. 2*(c_off + ch) bytes into the pixel
    uint8_t* q = reinterpret_cast<uint8_t*>(p) + 2 * v.c_buf - (v.c_off + ch);
    stg256(q, make_uint4(lo8[0], lo8[1], lo8[2], lo8[3]), make_uint4(lo8[4], lo8[5], lo8[6], lo8[7]));
    stg256(q + v.c_buf, make_uint4(hi8[0], hi8[1], hi8[2], hi8[3]), make_uint4(hi8[4], hi8[5], hi8[6], hi8[7]));
  }
}
template <int CH>
__device__ __forceinline__ void add_chunk_f16(const ActView& v, long long off, float* f, bool wide) {   // single plane only
  const __half* p = static_cast<const __half*>(v.data) + off;
  if (wide) {
#pragma unroll
    for (int g = 0; g < CH / 16; ++g) {
      uint4 a, b;
      ldg256(p + 16 * g, a, b);
      const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 x = f16x2_to_f32x2(w[j]);
        f[16 * g + 2 * j] += x.x * v.inv_mul; f[16 * g + 2 * j + 1] += x.y * v.inv_mul;
      }
    }
    return;
  }
#pragma unroll
  for (int g = 0; g < CH / 8; ++g) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p + 8 * g));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = f16x2_to_f32x2(w[j]);
      f[8 * g + 2 * j] += x.x * v.inv_mul; f[8 * g + 2 * j + 1] += x.y * v.inv_mul;
    }
  }
}

// WIDE: every lane writes 32-byte pieces (a thread owns one pixel, so a warp store spans 32 pixels: 16-byte
// pieces fill only half of each sector they touch)
template <int CH>
__device__ __forceinline__ void store_chunk(const ActView& v, long long off, int ch, const float* f, bool wide) {
  if (v.dtype == EAMM_F16) { store_chunk_f16<CH>(v, off, ch, f, wide); return; }
  __nv_bfloat16* p = static_cast<__nv_bfloat16*>(v.data) + off;
  if (wide) {
#pragma unroll
    for (int g = 0; g < CH / 16; ++g) {
      uint2 q[4];
      float4 h[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        q[i] = float4_to_bf16x4(make_float4(f[16 * g + 4 * i], f[16 * g + 4 * i + 1], f[16 * g + 4 * i + 2], f[16 * g + 4 * i + 3]));
        h[i] = bf16x4_to_float4(q[i]);
      }
      stg256(p + 16 * g, make_uint4(q[0].x, q[0].y, q[1].x, q[1].y), make_uint4(q[2].x, q[2].y, q[3].x, q[3].y));
      if (v.planes == 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          q[i] = float4_to_bf16x4(make_float4(f[16 * g + 4 * i] - h[i].x, f[16 * g + 4 * i + 1] - h[i].y,
                                              f[16 * g + 4 * i + 2] - h[i].z, f[16 * g + 4 * i + 3] - h[i].w));
        stg256(p + v.c_buf + 16 * g, make_uint4(q[0].x, q[0].y, q[1].x, q[1].y), make_uint4(q[2].x, q[2].y, q[3].x, q[3].y));
      }
    }
    return;
  }
#pragma unroll
  for (int g = 0; g < CH / 8; ++g) {
    uint2 a = float4_to_bf16x4(make_float4(f[8 * g], f[8 * g + 1], f[8 * g + 2], f[8 * g + 3]));
    uint2 b = float4_to_bf16x4(make_float4(f[8 * g + 4], f[8 * g + 5], f[8 * g + 6], f[8 * g + 7]));
    *reinterpret_cast<uint4*>(p + 8 * g) = make_uint4(a.x, a.y, b.x, b.y);
    if (v.planes == 2) {
      float4 ha = bf16x4_to_float4(a), hb = bf16x4_to_float4(b);
      uint2 la = float4_to_bf16x4(make_float4(f[8 * g] - ha.x, f[8 * g + 1] - ha.y, f[8 * g + 2] - ha.z, f[8 * g + 3] - ha.w));
      uint2 lb = float4_to_bf16x4(make_float4(f[8 * g + 4] - hb.x, f[8 * g + 5] - hb.y, f[8 * g + 6] - hb.z, f[8 * g + 7] - hb.w));
      *reinterpret_cast<uint4*>(p + v.c_buf + 8 * g) = make_uint4(la.x, la.y, lb.x, lb.y);
    }
  }
}
template <int CH>
__device__ __forceinline__ void add_chunk(const ActView& v, long long off, float* f, bool wide) {
  if (v.dtype == EAMM_F16) { add_chunk_f16<CH>(v, off, f, wide); return; }
  const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(v.data) + off;
  if (wide) {
#pragma unroll
    for (int g = 0; g < CH / 16; ++g) {
      for (int pl = 0; pl < v.planes; ++pl) {
        uint4 a, b;
        ldg256(p + pl * v.c_buf + 16 * g, a, b);
        const float4 x0 = bf16x4_to_float4(make_uint2(a.x, a.y)), x1 = bf16x4_to_float4(make_uint2(a.z, a.w));
        const float4 x2 = bf16x4_to_float4(make_uint2(b.x, b.y)), x3 = bf16x4_to_float4(make_uint2(b.z, b.w));
        float* o = f + 16 * g;
        o[0] += x0.x; o[1] += x0.y; o[2] += x0.z; o[3] += x0.w; o[4] += x1.x; o[5] += x1.y; o[6] += x1.z; o[7] += x1.w;
        o[8] += x2.x; o[9] += x2.y; o[10] += x2.z; o[11] += x2.w; o[12] += x3.x; o[13] += x3.y; o[14] += x3.z; o[15] += x3.w;
      }
    }
    return;
  }
#pragma unroll
  for (int g = 0; g < CH / 8; ++g) {
    uint4 r = __ldg(reinterpret_cast<const uint4*>(p + 8 * g));
    float4 a = bf16x4_to_float4(make_uint2(r.x, r.y)), b = bf16x4_to_float4(make_uint2(r.z, r.w));
    f[8 * g] += a.x; f[8 * g + 1] += a.y; f[8 * g + 2] += a.z; f[8 * g + 3] += a.w;
    f[8 * g + 4] += b.x; f[8 * g + 5] += b.y; f[8 * g + 6] += b.z; f[8 * g + 7] += b.w;
    if (v.planes == 2) {
      uint4 q = __ldg(reinterpret_cast<const uint4*>(p + v.c_buf + 8 * g));
      float4 c = bf16x4_to_float4(make_uint2(q.x, q.y)), d = bf16x4_to_float4(make_uint2(q.z, q.w));
      f[8 * g] += c.x; f[8 * g + 1] += c.y; f[8 * g + 2] += c.z; f[8 * g + 3] += c.w;
      f[8 * g + 4] += d.x; f[8 * g + 5] += d.y; f[8 * g + 6] += d.z; f[8 * g + 7] += d.w;
    }
  }
}

struct TileCoord { int x0, y0, n0, cls, nt, split; uint32_t out_tile; };

__device__ __forceinline__ TileCoord decode_tile(const ConvTcParams& p, uint32_t tile) {
  TileCoord t;
  t.split = 0;
  if (p.splitk > 1) { const uint32_t q0 = tile / (uint32_t)p.splitk; t.split = (int)(tile - q0 * p.splitk); tile = q0; }
  t.out_tile = tile;
  // CTA pairs: tiles 2i and 2i+1 are adjacent M tiles of the same (class, N tile) -- they share the weights
  const uint32_t rank = p.cta2 ? (tile & 1u) : 0u;
  if (p.cta2) tile >>= 1;
  uint32_t q;
  if (p.dec_shift >= 0) {
    // every divisor is a power of two (all power-of-two maps with plain tiles): shifts instead of four ~20-instruction
    // divisions per warp and tile (6 % of the instructions of the narrow layers)
    const uint32_t sh = (uint32_t)p.dec_shift;
    const uint32_t s_nt = sh & 31u, s_cg = (sh >> 5) & 31u, s_tx = (sh >> 10) & 31u, s_ty = (sh >> 15) & 31u;
    t.nt = (int)(tile & ((1u << s_nt) - 1u)); tile >>= s_nt;
    t.cls = (int)(tile & ((1u << s_cg) - 1u)) * p.ah_g; tile >>= s_cg;
    if (p.cta2) tile = 2u * tile + rank;
    t.x0 = (int)(tile & ((1u << s_tx) - 1u)) * p.x_stride; tile >>= s_tx;
    t.y0 = (int)(tile & ((1u << s_ty) - 1u)) * p.y_stride; tile >>= s_ty;
    t.n0 = (int)tile * p.bn;
    return t;
  }
  q = tile / (uint32_t)p.n_tiles; t.nt = (int)(tile - q * p.n_tiles); tile = q;
  q = tile / (uint32_t)p.cls_groups; t.cls = (int)(tile - q * p.cls_groups) * p.ah_g; tile = q;
  if (p.cta2) tile = 2u * tile + rank;
  q = tile / (uint32_t)p.tiles_x; t.x0 = (int)(tile - q * p.tiles_x) * p.x_stride; tile = q;
  q = tile / (uint32_t)p.tiles_y; t.y0 = (int)(tile - q * p.tiles_y) * p.y_stride; tile = q;
  t.n0 = (int)tile * p.bn;
  return t;
}

// The MMA-issuing warp of the production kernel.  What the tensor core needs per K step is short (measured with
// tools/experiments/mma_step_cost.cu on B200: 128 cycles at N = 256, 64 at N = 128, <= 48 at N <= 64, the same for
// kind::f16 K = 16 and kind::f8f6f4 K = 32, single CTA or pair, no penalty for alternating kinds), while one lane
// retires a dependent instruction only every ~8-10 cycles: the issue loop, not the MMA, set the ~110-cycle floor per
// step of the narrow layers and the 166-200 cycles of the N = 256 ones.  So this loop is specialised at compile time
// (MODE) instead of branching per chunk, keeps every parameter in registers (no constant-bank reloads) and walks the
// ring with running pre-shifted descriptor counters.
//   MODE 0  every chunk one step group with `idesc` (plain layers; fold 2 = packed first conv with N = 2*BN)
//   MODE 1  mixed fp16 + fp8 operands: chunks [0, nf8) of the K range are kind::f8f6f4, the rest kind::f16
//   MODE 2  fold 1, single CTA: type-0 chunks (kc == 0 or kc > T) are one N = 2*BN group, type-1 chunks one N = BN group
//   MODE 3  fold 1, CTA pair: type-0 chunks are two N = BN groups (b_hi halves -> columns [0, BN), b_lo halves -> [BN, 2BN))
template <bool CTA2, int MODE>
__device__ __forceinline__ void mma_issuer_lean(const ConvTcParams& p, uint32_t smem_base, uint32_t bar0, uint32_t tmem_base,
                                                uint32_t KC, uint32_t KS, uint32_t a_slot, uint32_t b_bytes, uint32_t b_half,
                                                uint32_t tile0, uint32_t tile_step, uint32_t total_tiles, uint32_t bres_base = 0u) {
  const uint32_t fmt = p.f16in ? 0u : ((1u << 7) | (1u << 10));
  const uint32_t mbits = ((CTA2 ? 256u : 128u) >> 4) << 24;
  uint32_t idesc1 = (1u << 4) | fmt | ((uint32_t)(p.BN >> 3) << 17) | mbits;
  uint32_t idesc2 = (1u << 4) | fmt | ((uint32_t)(p.BN >> 2) << 17) | mbits;          // N = 2*BN
  if (MODE == 0 && p.fold == 2) idesc1 = idesc2;
  uint32_t nstages = (uint32_t)p.num_stages, nf8 = (uint32_t)p.nf8, BN = (uint32_t)p.BN, splitk = (uint32_t)p.splitk;
  const uint32_t T = KC >> 1;
  uint32_t slotA16 = a_slot >> 4, slotB16 = b_bytes >> 4, bofs16 = (KS * a_slot) >> 4, bhalf16 = b_half >> 4;
  uint32_t stage16 = (KS * (a_slot + (bres_base ? 0u : b_bytes))) >> 4, base16 = smem_base >> 4;
  uint32_t bres16 = bres_base >> 4;                   // MODE 0 only: resident weights, chunk kc at bres16 + kc * slotB16
  if (bres_base) { mbar_wait(bar0 + 8u * 44u, 0u); tc_fence_after(); }
  // opaque copies: keep ptxas from re-reading the kernel parameters (constant bank) inside the loops
  asm volatile("mov.u32 %0, %0;" : "+r"(idesc1)); asm volatile("mov.u32 %0, %0;" : "+r"(idesc2));
  asm volatile("mov.u32 %0, %0;" : "+r"(nstages)); asm volatile("mov.u32 %0, %0;" : "+r"(nf8));
  asm volatile("mov.u32 %0, %0;" : "+r"(slotA16)); asm volatile("mov.u32 %0, %0;" : "+r"(slotB16));
  asm volatile("mov.u32 %0, %0;" : "+r"(bofs16)); asm volatile("mov.u32 %0, %0;" : "+r"(stage16));
  asm volatile("mov.u32 %0, %0;" : "+r"(KS)); asm volatile("mov.u32 %0, %0;" : "+r"(KC));
  uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
  uint32_t sa16 = base16, fb = bar0, eb = bar0 + 8u * 16u;
  for (uint32_t tile = tile0; tile < total_tiles; tile += tile_step) {
    mbar_wait(bar0 + 8u * (34u + as), aphase ^ 1u);                       // accumulator `as` drained by the epilogue
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base + as * 256u;
    const uint32_t tfull = bar0 + 8u * (32u + as);
    // mixed operands under split-K: this work item's K range starts at chunk kbase of the whole K loop
    uint32_t nf = 0;
    if (MODE == 1) {
      const uint32_t kbase = splitk > 1u ? (uint32_t)decode_tile(p, tile).split * KC : 0u;
      nf = nf8 > kbase ? nf8 - kbase : 0u;                                // chunks [0, nf) of this item are fp8
    }
    for (uint32_t kc = 0; kc < KC; kc += KS) {
      const uint32_t nsub = KC - kc < KS ? KC - kc : KS;
      mbar_wait(fb, phase);
      tc_fence_after();
      if (elect_one()) {
        uint32_t a16 = sa16, b16 = (MODE == 0 && bres16) ? bres16 + kc * slotB16 : sa16 + bofs16;
#pragma unroll 1
        for (uint32_t sub = 0; sub < nsub; ++sub, a16 += slotA16, b16 += slotB16) {
          const uint32_t kcs = kc + sub;
          const uint64_t da = sw128_desc16(a16), db = sw128_desc16(b16);
          const uint32_t acc = kcs ? 1u : 0u;
          if (MODE == 0) issue_chunk<CTA2, false>(tmem_acc, da, db, idesc1, acc);
          if (MODE == 1) {
            if (kcs < nf) issue_chunk<CTA2, true>(tmem_acc, da, db, idesc1, acc);
            else issue_chunk<CTA2, false>(tmem_acc, da, db, idesc1, acc);
          }
          if (MODE == 2) issue_chunk<CTA2, false>(tmem_acc, da, db, (kcs == 0u || kcs > T) ? idesc2 : idesc1, acc);
          if (MODE == 3) {
            issue_chunk<CTA2, false>(tmem_acc, da, db, idesc1, acc);
            if (kcs == 0u || kcs > T) issue_chunk<CTA2, false>(tmem_acc + BN, da, db + (uint64_t)bhalf16, idesc1, acc);
          }
        }
        if (CTA2) { tc2_commit_mc(eb); if (kc + KS >= KC) tc2_commit_mc(tfull); }
        else { tc_commit(eb); if (kc + KS >= KC) tc_commit(tfull); }
      }
      __syncwarp();
      ++stage; sa16 += stage16; fb += 8u; eb += 8u;
      if (stage == nstages) { stage = 0; phase ^= 1u; sa16 = base16; fb = bar0; eb = bar0 + 8u * 16u; }
    }
    as ^= 1u; if (as == 0u) aphase ^= 1u;
  }
}


// ------------------------------------------------------------------------------ halo-tile scheme (ConvTcParams::ah)
// Barriers: weight ring full/empty = bars[0..15] / [16..31] as in the plain scheme, A ring full = bars[36..39], empty = bars[40..43].
constexpr uint32_t AH_PITCH = 10, AH_ROWS = 18, AH_A_BYTES = AH_PITCH * AH_ROWS * 128;     // 180 pixel rows of 128 bytes
// K-major SWIZZLE_128B descriptor of a halo view: 8-row groups AH_PITCH * 128 bytes apart (tools/experiments/halo_desc.cu:
// the swizzle is a function of the absolute shared-memory address, so neither the start nor the group stride needs 1024-byte alignment)
constexpr uint64_t AH_DESC_BITS = ((uint64_t)1 << 16) | ((uint64_t)((AH_PITCH * 128) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);

// All TC_PRODUCERS warps walk every weight stage of every tile (cheap counters) and issue the ones whose running index is
// theirs; the producer that owns the first weight stage of a K chunk also stages that chunk's halo tile.
template <bool CTA2>
__device__ __forceinline__ void producer_ah(const ConvTcParams& p, const CUtensorMap* tmA, const CUtensorMap* tmB,
                                            uint32_t smem_base, uint32_t bar0, uint32_t w, uint32_t cta_rank,
                                            uint32_t tile0, uint32_t tile_step, uint32_t total_tiles) {
  const uint32_t KS = (uint32_t)p.ksub, NB = (uint32_t)p.num_stages, NA = (uint32_t)p.ah_na, SPC = (uint32_t)p.ah_spc;
  const uint32_t ntap = (uint32_t)p.ntap, a_slot = (uint32_t)p.a_slot_bytes;
  const uint32_t b_bytes = (uint32_t)p.BN * (CTA2 ? 64u : 128u);
  const uint32_t b_base = smem_base + NA * a_slot, stage_bytes = KS * b_bytes;
  const uint32_t a_tx = AH_A_BYTES * (CTA2 ? 2u : 1u), b_tx = stage_bytes * (CTA2 ? 2u : 1u);
  const int nsec = p.ah_nsec, cout = p.cout;
  uint32_t bslot = 0, bphase = 0, aslot = 0, aphase = 0, turn = 0;
  const bool prof = p.prof != nullptr;
  long long pw = 0, pstart = 0;
  if (prof) pstart = clock64();
  for (uint32_t tile = tile0; tile < total_tiles; tile += tile_step) {
    const TileCoord tc = decode_tile(p, tile);
    const int brow0 = tc.nt * p.BN + (CTA2 ? (int)cta_rank * (p.BN >> 1) : 0);
    for (int sec = 0; sec < nsec; ++sec) {
      const uint32_t nch = (uint32_t)p.ah_nch[sec];
      for (uint32_t cc = 0; cc < nch; ++cc) {
        uint32_t cls_l = 0, t0 = 0;
        for (uint32_t r = 0; r < SPC; ++r) {
          if (turn == w) {
            long long tw = 0;
            if (prof) tw = clock64();
            mbar_wait(bar0 + 8u * (16u + bslot), bphase ^ 1u);
            if (r == 0u) mbar_wait(bar0 + 8u * (40u + aslot), aphase ^ 1u);
            if (prof) pw += clock64() - tw;
            if (elect_one()) {
              const uint32_t fbl = bar0 + 8u * bslot, afl = bar0 + 8u * (36u + aslot);
              const uint32_t fb = CTA2 ? mapa_shared(fbl, 0u) : fbl, af = CTA2 ? mapa_shared(afl, 0u) : afl;
              if (r == 0u) {
                if (!CTA2 || cta_rank == 0u) mbar_expect_tx(afl, a_tx);
                const int cbase = p.ah_cbase[sec] + (int)cc * 128;
                if (CTA2) tma2_load_4d(smem_base + aslot * a_slot, tmA, af, cbase, tc.x0 - 1, tc.y0 - 1, tc.n0);
                else tma_load_4d(smem_base + aslot * a_slot, tmA, af, cbase, tc.x0 - 1, tc.y0 - 1, tc.n0);
              }
              if (!CTA2 || cta_rank == 0u) mbar_expect_tx(fbl, b_tx);
              const int brow = (tc.cls + (int)cls_l) * cout + brow0;
              const uint32_t sB = b_base + bslot * stage_bytes;
              for (uint32_t sub = 0; sub < KS; ++sub) {
                const int bcol = p.ah_bcol[sec] + (int)(((t0 + sub) * nch + cc) * 128u);
                if (CTA2) tma2_load_2d(sB + sub * b_bytes, tmB, fb, bcol, brow);
                else tma_load_2d(sB + sub * b_bytes, tmB, fb, bcol, brow);
              }
            }
            __syncwarp();
          }
          turn = turn + 1u == (uint32_t)TC_PRODUCERS ? 0u : turn + 1u;
          if (++bslot == NB) { bslot = 0; bphase ^= 1u; }
          t0 += KS;
          if (t0 == ntap) { t0 = 0; ++cls_l; }
        }
        if (++aslot == NA) { aslot = 0; aphase ^= 1u; }
      }
    }
  }
#ifndef EAMM_EPI_PHASES
  if (prof && w == 0 && (threadIdx.x & 31) == 0) {
    p.prof[blockIdx.x * 8 + 0] = pw;                             // producer 0: cycles waiting for free slots
    p.prof[blockIdx.x * 8 + 1] = clock64() - pstart;             // producer 0: total
  }
#endif
}

// One weight stage of the halo-tile issuer: KS taps of one parity class, fully unrolled.  KIND 0 = 3x3 (KS 3: stage = filter
// row, KS 1: stage = tap), 1 = UP2 (KS 4: stage = class, 2: class row, 1: tap).  Every descriptor is the stage's base plus a
// compile-time offset, so the elected lane's instructions are independent of each other (a lone lane retires a DEPENDENT
// instruction only every ~8-10 cycles: see mma_issuer_lean).  `a16` = halo view of the stage's first tap, in 16-byte units.
template <bool CTA2, bool F8, int KIND, int KS>
__device__ __forceinline__ void ah_issue_stage(uint32_t acc_col, uint32_t a16, uint32_t b16, uint32_t slotB16, uint32_t idesc, uint32_t acc0) {
#pragma unroll
  for (int sub = 0; sub < KS; ++sub) {
    // tap offsets inside a stage: x-adjacent taps are one pixel row (8 units) apart; UP2 KS 4 walks (0,0) (0,1) (1,0) (1,1)
    const uint32_t toff = (KIND == 1 && KS == 4) ? (uint32_t)((sub >> 1) * (int)AH_PITCH * 8 + (sub & 1) * 8) : (uint32_t)(sub * 8);
    const uint64_t da = AH_DESC_BITS | (uint64_t)(a16 + toff), db = sw128_desc16(b16 + (uint32_t)sub * slotB16);
    issue_chunk<CTA2, F8>(acc_col, da, db, idesc, sub == 0 ? acc0 : 1u);
  }
}

template <bool CTA2, int KIND, int KS>
__device__ __forceinline__ void mma_issuer_ah(const ConvTcParams& p, uint32_t smem_base, uint32_t bar0, uint32_t tmem_base,
                                              uint32_t tile0, uint32_t tile_step, uint32_t total_tiles) {
  uint32_t idesc = (1u << 4) | ((uint32_t)(p.BN >> 3) << 17) | (((CTA2 ? 256u : 128u) >> 4) << 24);   // fp16 / e4m3 operands: format 0
  uint32_t NB = (uint32_t)p.num_stages, NA = (uint32_t)p.ah_na, SPC = (uint32_t)p.ah_spc, BN = (uint32_t)p.BN;
  const uint32_t b_bytes = BN * (CTA2 ? 64u : 128u);
  uint32_t slotA16 = (uint32_t)p.a_slot_bytes >> 4, slotB16 = b_bytes >> 4, stage16 = ((uint32_t)KS * b_bytes) >> 4;
  uint32_t base16 = smem_base >> 4, bbase16 = (smem_base + NA * (uint32_t)p.a_slot_bytes) >> 4;
  uint32_t nsec = (uint32_t)p.ah_nsec;
  // per-section chunk counts and kinds packed into registers (no constant-bank reads inside the loops)
  uint32_t nch0 = (uint32_t)p.ah_nch[0], nch1 = nsec > 1u ? (uint32_t)p.ah_nch[1] : 0u, nch2 = nsec > 2u ? (uint32_t)p.ah_nch[2] : 0u;
  uint32_t f8mask = (p.ah_f8[0] ? 1u : 0u) | ((nsec > 1u && p.ah_f8[1]) ? 2u : 0u) | ((nsec > 2u && p.ah_f8[2]) ? 4u : 0u);
  const uint32_t grp = (uint32_t)p.ah_g;
  asm volatile("mov.u32 %0, %0;" : "+r"(idesc)); asm volatile("mov.u32 %0, %0;" : "+r"(NB));
  asm volatile("mov.u32 %0, %0;" : "+r"(NA)); asm volatile("mov.u32 %0, %0;" : "+r"(SPC));
  asm volatile("mov.u32 %0, %0;" : "+r"(slotA16)); asm volatile("mov.u32 %0, %0;" : "+r"(slotB16));
  asm volatile("mov.u32 %0, %0;" : "+r"(stage16)); asm volatile("mov.u32 %0, %0;" : "+r"(BN));
  asm volatile("mov.u32 %0, %0;" : "+r"(nch0)); asm volatile("mov.u32 %0, %0;" : "+r"(nch1));
  asm volatile("mov.u32 %0, %0;" : "+r"(nch2)); asm volatile("mov.u32 %0, %0;" : "+r"(f8mask));
  asm volatile("mov.u32 %0, %0;" : "+r"(nsec));
  uint32_t bslot = 0, bphase = 0, aslot = 0, aphase = 0, as = 0, aphase_t = 0;
  uint32_t a16 = base16, b16 = bbase16;
  uint32_t fb = bar0, eb = bar0 + 8u * 16u, afb = bar0 + 8u * 36u, aeb = bar0 + 8u * 40u;
  const bool prof = p.prof != nullptr;
  long long pm0 = 0, pm1 = 0, pstart = 0;
  if (prof) pstart = clock64();
  for (uint32_t tile = tile0; tile < total_tiles; tile += tile_step) {
    // UP2: halo offset of the item's first class (classes cls0 .. cls0 + grp - 1; grp 4 -> 0, grp 2 -> class row, grp 1 -> class)
    uint32_t cls0 = 0;
    if (KIND == 1 && grp < 4u) cls0 = (uint32_t)decode_tile(p, tile).cls;
    long long t0 = 0;
    if (prof) t0 = clock64();
    mbar_wait(bar0 + 8u * (34u + as), aphase_t ^ 1u);                       // accumulator `as` drained by the epilogue
    if (prof) pm0 += clock64() - t0;
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base + as * 256u;
    const uint32_t tfull = bar0 + 8u * (32u + as);
    uint32_t fresh = 0u;                                                    // 0 on the first K chunk: a class's first tap overwrites
    for (uint32_t sec = 0; sec < nsec; ++sec) {
      const uint32_t nch = sec == 0u ? nch0 : (sec == 1u ? nch1 : nch2);
      const bool f8 = (f8mask >> sec) & 1u;
      const bool last_sec = sec + 1u == nsec;
      for (uint32_t cc = 0; cc < nch; ++cc) {
        const bool last_chunk = last_sec && cc + 1u == nch;
        // stage walk state: 3x3 -> (row offset, x) ; UP2 -> (class, tap)
        uint32_t soff = 0, sx = 0, cls_l = 0;
        for (uint32_t r = 0; r < SPC; ++r) {
          // this stage's first-tap view and accumulate flag, computed before the waits
          uint32_t va16, acc0, acc_col;
          if (KIND == 0) { va16 = a16 + soff; acc0 = fresh | r; acc_col = tmem_acc; }
          else {
            const uint32_t c = cls0 + cls_l;
            va16 = a16 + ((c >> 1) * AH_PITCH + (c & 1u)) * 8u + soff;
            acc0 = fresh | sx; acc_col = tmem_acc + cls_l * BN;
          }
          if (prof) t0 = clock64();
          mbar_wait(fb, bphase);
          if (r == 0u) mbar_wait(afb, aphase);
          if (prof) pm1 += clock64() - t0;
          tc_fence_after();
          if (elect_one()) {
            if (f8) ah_issue_stage<CTA2, true, KIND, KS>(acc_col, va16, b16, slotB16, idesc, acc0);
            else ah_issue_stage<CTA2, false, KIND, KS>(acc_col, va16, b16, slotB16, idesc, acc0);
            const bool last_r = r + 1u == SPC;
            if (CTA2) { tc2_commit_mc(eb); if (last_r) { tc2_commit_mc(aeb); if (last_chunk) tc2_commit_mc(tfull); } }
            else { tc_commit(eb); if (last_r) { tc_commit(aeb); if (last_chunk) tc_commit(tfull); } }
          }
          __syncwarp();
          b16 += stage16; fb += 8u; eb += 8u;
          if (++bslot == NB) { bslot = 0; bphase ^= 1u; b16 = bbase16; fb = bar0; eb = bar0 + 8u * 16u; }
          if (KIND == 0) {
            if (KS == 3) soff += AH_PITCH * 8u;                                  // next filter row
            else { soff += 8u; if (++sx == 3u) { sx = 0; soff += (AH_PITCH - 3u) * 8u; } }
          } else {
            if (KS == 4) ++cls_l;                                                // next class
            else if (KS == 2) { if (++sx == 2u) { sx = 0; soff = 0; ++cls_l; } else soff = AH_PITCH * 8u; }
            else { ++sx; if (sx == 4u) { sx = 0; soff = 0; ++cls_l; } else soff = (sx >> 1) * AH_PITCH * 8u + (sx & 1u) * 8u; }
          }
        }
        fresh = 1u;
        a16 += slotA16; afb += 8u; aeb += 8u;
        if (++aslot == NA) { aslot = 0; aphase ^= 1u; a16 = base16; afb = bar0 + 8u * 36u; aeb = bar0 + 8u * 40u; }
      }
    }
    as ^= 1u; if (as == 0u) aphase_t ^= 1u;
  }
  if (prof && (threadIdx.x & 31) == 0) {
    p.prof[blockIdx.x * 8 + 2] = pm0;                            // MMA: waiting for a free accumulator
    p.prof[blockIdx.x * 8 + 3] = pm1;                            // MMA: waiting for operands
    p.prof[blockIdx.x * 8 + 4] = clock64() - pstart;             // MMA: total
  }
}

// ------------------------------------------------------------------------------ fast epilogue (ConvTcParams::epi_fast)
// The hot layers of the path (activation-view outputs, N tile a multiple of 32, no split-K) take this variant of
// epilogue_tile: same arithmetic in the same order, but (a) everything that depends only on the tile -- pixel addresses of
// the views, flags, vector bases -- is computed once per tile instead of per 32-column chunk (the generic code re-read ~40
// kernel parameters from the constant bank per chunk: LDCU results sit on the long scoreboard), (b) the residual / second
// output / pooling variants are compile-time, (c) the TMEM load of the next chunk is issued as soon as the current chunk
// has left its registers, so its latency hides behind the conversions and stores.
// View format codes: 0 = bf16, 1 = bf16 hi/lo planes, 2 = fp16, 3 = fp16 + e4m3 lo8 + e4m3 hi8 (mixed).
struct EpiView {
  char* p0;          // byte address of plane 0 at (pixel, first channel of the N tile)
  char* p1;          // second plane at the same channel: bf16 lo plane, or the e4m3 lo8 bytes of the mixed format
  int fmt, c_buf;
  int col0, col1;    // byte columns of p0 / p1 inside the pixel record (TMA-store coordinates)
  float mul, inv_mul;
};
__device__ __forceinline__ EpiView epi_view(const ActView& v, int n, int y, int x, int co0) {
  EpiView e;
  const long long off = act_offset(v, n, y, x, co0);                       // elements (2 bytes each in every non-f32 format)
  e.p0 = static_cast<char*>(v.data) + 2 * off;
  e.fmt = (v.dtype == EAMM_F16 ? 2 : 0) + (v.planes == 2 ? 1 : 0);
  e.c_buf = v.c_buf;
  // mixed: plane 1 of the pixel starts 2*c_buf bytes after plane 0 and holds [c_buf lo8 | c_buf hi8]
  e.p1 = e.fmt == 3 ? e.p0 - 2 * (v.c_off + co0) + 2 * v.c_buf + (v.c_off + co0) : e.p0 + 2 * v.c_buf;
  e.col0 = 2 * (v.c_off + co0);
  e.col1 = e.fmt == 3 ? 2 * v.c_buf + v.c_off + co0 : 2 * v.c_buf + 2 * (v.c_off + co0);
  e.mul = v.mul; e.inv_mul = v.inv_mul;
  return e;
}

// NCH consecutive channels (8 or 32) of one pixel, starting `c` channels into the N tile
template <int NCH>
__device__ __forceinline__ void epi_store(const EpiView& e, int c, const float* f) {
  char* q0 = e.p0 + 2 * c;
  if (e.fmt >= 2) {
    uint32_t h[NCH / 2];
#pragma unroll
    for (int j = 0; j < NCH / 2; ++j) h[j] = f32x2_to_f16x2_sat(f[2 * j] * e.mul, f[2 * j + 1] * e.mul);
    if (NCH == 32) {
#pragma unroll
      for (int g = 0; g < 2; ++g)
        stg256(q0 + 32 * g, make_uint4(h[8 * g], h[8 * g + 1], h[8 * g + 2], h[8 * g + 3]),
               make_uint4(h[8 * g + 4], h[8 * g + 5], h[8 * g + 6], h[8 * g + 7]));
    } else {
      *reinterpret_cast<uint4*>(q0) = make_uint4(h[0], h[1], h[2], h[3]);
    }
    if (e.fmt == 3) {
      uint32_t lo8[NCH / 4], hi8[NCH / 4];
#pragma unroll
      for (int j = 0; j < NCH / 4; ++j) {
        const float2 a = f16x2_to_f32x2(h[2 * j]), b = f16x2_to_f32x2(h[2 * j + 1]);
        lo8[j] = f32x4_to_e4m3x4_sat((f[4 * j] * e.mul - a.x) * MIX_LO_GAIN, (f[4 * j + 1] * e.mul - a.y) * MIX_LO_GAIN,
                                     (f[4 * j + 2] * e.mul - b.x) * MIX_LO_GAIN, (f[4 * j + 3] * e.mul - b.y) * MIX_LO_GAIN);
        hi8[j] = f16x4_to_hi8x4(h[2 * j], h[2 * j + 1]);
      }
      char* q1 = e.p1 + c;
      if (NCH == 32) {
        stg256(q1, make_uint4(lo8[0], lo8[1], lo8[2], lo8[3]), make_uint4(lo8[4], lo8[5], lo8[6], lo8[7]));
        stg256(q1 + e.c_buf, make_uint4(hi8[0], hi8[1], hi8[2], hi8[3]), make_uint4(hi8[4], hi8[5], hi8[6], hi8[7]));
      } else {
        *reinterpret_cast<uint2*>(q1) = make_uint2(lo8[0], lo8[1]);
        *reinterpret_cast<uint2*>(q1 + e.c_buf) = make_uint2(hi8[0], hi8[1]);
      }
    }
    return;
  }
  char* q1 = e.p1 + 2 * c;
#pragma unroll
  for (int g = 0; g < NCH / 16; ++g) {
    uint2 q[4];
    float4 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      q[i] = float4_to_bf16x4(make_float4(f[16 * g + 4 * i], f[16 * g + 4 * i + 1], f[16 * g + 4 * i + 2], f[16 * g + 4 * i + 3]));
      h[i] = bf16x4_to_float4(q[i]);
    }
    stg256(q0 + 32 * g, make_uint4(q[0].x, q[0].y, q[1].x, q[1].y), make_uint4(q[2].x, q[2].y, q[3].x, q[3].y));
    if (e.fmt == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        q[i] = float4_to_bf16x4(make_float4(f[16 * g + 4 * i] - h[i].x, f[16 * g + 4 * i + 1] - h[i].y,
                                            f[16 * g + 4 * i + 2] - h[i].z, f[16 * g + 4 * i + 3] - h[i].w));
      stg256(q1 + 32 * g, make_uint4(q[0].x, q[0].y, q[1].x, q[1].y), make_uint4(q[2].x, q[2].y, q[3].x, q[3].y));
    }
  }
  if (NCH == 8) {
    const uint2 a = float4_to_bf16x4(make_float4(f[0], f[1], f[2], f[3]));
    const uint2 b = float4_to_bf16x4(make_float4(f[4], f[5], f[6], f[7]));
    *reinterpret_cast<uint4*>(q0) = make_uint4(a.x, a.y, b.x, b.y);
    if (e.fmt == 1) {
      const float4 ha = bf16x4_to_float4(a), hb = bf16x4_to_float4(b);
      const uint2 la = float4_to_bf16x4(make_float4(f[0] - ha.x, f[1] - ha.y, f[2] - ha.z, f[3] - ha.w));
      const uint2 lb = float4_to_bf16x4(make_float4(f[4] - hb.x, f[5] - hb.y, f[6] - hb.z, f[7] - hb.w));
      *reinterpret_cast<uint4*>(q1) = make_uint4(la.x, la.y, lb.x, lb.y);
    }
  }
}

__device__ __forceinline__ void epi_add16_bf16(const uint4 a, const uint4 b, float* o) {
  const float4 x0 = bf16x4_to_float4(make_uint2(a.x, a.y)), x1 = bf16x4_to_float4(make_uint2(a.z, a.w));
  const float4 x2 = bf16x4_to_float4(make_uint2(b.x, b.y)), x3 = bf16x4_to_float4(make_uint2(b.z, b.w));
  o[0] += x0.x; o[1] += x0.y; o[2] += x0.z; o[3] += x0.w; o[4] += x1.x; o[5] += x1.y; o[6] += x1.z; o[7] += x1.w;
  o[8] += x2.x; o[9] += x2.y; o[10] += x2.z; o[11] += x2.w; o[12] += x3.x; o[13] += x3.y; o[14] += x3.z; o[15] += x3.w;
}
// TMA-store variant of epi_store<32> (ConvTcParams::tma_st): the lane's 32 channels go to the warp's staging area (row = lane,
// 16-byte pieces XOR-swizzled as the tensor maps expect: conflict-free st.shared.v4), lane 0 issues the bulk stores.
struct EpiTma { const CUtensorMap* m64; const CUtensorMap* m32; uint32_t area; int px, x, py, yr; };
__device__ __forceinline__ void epi_store_tma(const EpiView& e, const EpiTma& t, int lane, int c, const float* f) {
  if (lane == 0) bulk_wait_read0();               // earlier bulk stores of this warp have finished reading the staging areas
  __syncwarp();
  const uint32_t row64 = t.area + (uint32_t)lane * 64u, sw64 = ((uint32_t)lane >> 1) & 3u;
  const uint32_t row32 = t.area + 2048u + (uint32_t)lane * 32u, sw32 = ((uint32_t)lane >> 2) & 1u;
  if (e.fmt >= 2) {
    uint32_t h[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = f32x2_to_f16x2_sat(f[2 * j] * e.mul, f[2 * j + 1] * e.mul);
#pragma unroll
    for (int k = 0; k < 4; ++k) sts128(row64 + (((uint32_t)k ^ sw64) << 4), h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
    if (e.fmt == 3) {
      uint32_t lo8[8], hi8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 a = f16x2_to_f32x2(h[2 * j]), b = f16x2_to_f32x2(h[2 * j + 1]);
        lo8[j] = f32x4_to_e4m3x4_sat((f[4 * j] * e.mul - a.x) * MIX_LO_GAIN, (f[4 * j + 1] * e.mul - a.y) * MIX_LO_GAIN,
                                     (f[4 * j + 2] * e.mul - b.x) * MIX_LO_GAIN, (f[4 * j + 3] * e.mul - b.y) * MIX_LO_GAIN);
        hi8[j] = f16x4_to_hi8x4(h[2 * j], h[2 * j + 1]);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        sts128(row32 + (((uint32_t)k ^ sw32) << 4), lo8[4 * k], lo8[4 * k + 1], lo8[4 * k + 2], lo8[4 * k + 3]);
        sts128(row32 + 1024u + (((uint32_t)k ^ sw32) << 4), hi8[4 * k], hi8[4 * k + 1], hi8[4 * k + 2], hi8[4 * k + 3]);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint2 a = float4_to_bf16x4(make_float4(f[8 * k], f[8 * k + 1], f[8 * k + 2], f[8 * k + 3]));
      const uint2 b = float4_to_bf16x4(make_float4(f[8 * k + 4], f[8 * k + 5], f[8 * k + 6], f[8 * k + 7]));
      sts128(row64 + (((uint32_t)k ^ sw64) << 4), a.x, a.y, b.x, b.y);
      if (e.fmt == 1) {
        const float4 ha = bf16x4_to_float4(a), hb = bf16x4_to_float4(b);
        const uint2 la = float4_to_bf16x4(make_float4(f[8 * k] - ha.x, f[8 * k + 1] - ha.y, f[8 * k + 2] - ha.z, f[8 * k + 3] - ha.w));
        const uint2 lb = float4_to_bf16x4(make_float4(f[8 * k + 4] - hb.x, f[8 * k + 5] - hb.y, f[8 * k + 6] - hb.z, f[8 * k + 7] - hb.w));
        sts128(row64 + 2048u + (((uint32_t)k ^ sw64) << 4), la.x, la.y, lb.x, lb.y);
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the bulk copy
  __syncwarp();
  if (lane == 0) {
    tma_store_5d(t.m64, t.area, e.col0 + 2 * c, t.px, t.x, t.py, t.yr);
    if (e.fmt == 3) {
      tma_store_5d(t.m32, t.area + 2048u, e.col1 + c, t.px, t.x, t.py, t.yr);
      tma_store_5d(t.m32, t.area + 3072u, e.col1 + e.c_buf + c, t.px, t.x, t.py, t.yr);
    } else if (e.fmt == 1) {
      tma_store_5d(t.m64, t.area + 2048u, e.col1 + 2 * c, t.px, t.x, t.py, t.yr);
    }
    bulk_commit();
  }
}

// residual of one 32-channel chunk: the loads are issued ahead of use (all of them back to back -- a loop over the planes
// with a load -> use dependence per iteration serialised four DRAM round trips per chunk: 17 k cycles per conv2 tile),
// epi_res_add then adds plane by plane in add_chunk's order.  rr = [16-channel group][plane] x 32 bytes.
__device__ __forceinline__ void epi_res_load(const EpiView& e, int c, uint4* rr) {
  const char* q0 = e.p0 + 2 * c;
  const char* q1 = e.p1 + 2 * c;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    ldg256(q0 + 32 * g, rr[4 * g], rr[4 * g + 1]);
    if (e.fmt == 1) ldg256(q1 + 32 * g, rr[4 * g + 2], rr[4 * g + 3]);
  }
}
__device__ __forceinline__ void epi_res_add(const EpiView& e, const uint4* rr, float* f) {
  if (e.fmt == 2) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const uint32_t w[8] = {rr[4 * g].x, rr[4 * g].y, rr[4 * g].z, rr[4 * g].w, rr[4 * g + 1].x, rr[4 * g + 1].y, rr[4 * g + 1].z, rr[4 * g + 1].w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 x = f16x2_to_f32x2(w[j]);
        f[16 * g + 2 * j] += x.x * e.inv_mul; f[16 * g + 2 * j + 1] += x.y * e.inv_mul;
      }
    }
    return;
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    epi_add16_bf16(rr[4 * g], rr[4 * g + 1], f + 16 * g);
    if (e.fmt == 1) epi_add16_bf16(rr[4 * g + 2], rr[4 * g + 3], f + 16 * g);
  }
}

template <bool POOL, bool RES, bool OUT2, bool FOLD, bool TMA>
__device__ __forceinline__ void epilogue_fast(const ConvTcParams& p, const TileCoord& tc, uint32_t tmem_acc,
                                              int quadrant, int lane, int half, float& amax1, float& amax2,
                                              const StoreMaps* maps, uint32_t st_area) {
  const int r = quadrant * 32 + lane;
  const int xl = r & (p.bw - 1);
  const int yl = (r >> p.bw_log2) & (p.bh - 1);
  const int nl = r >> (p.bw_log2 + p.bh_log2);
  const int x = tc.x0 + xl, y = tc.y0 + yl, n = tc.n0 + nl;
  const bool valid = (y < p.H) && (n < p.N);
  int oy = y, ox = x;
  if (POOL) { oy = y >> 1; ox = x >> 1; }
  else if (p.kind == EAMM_CONV_UP2_3X3) { oy = 2 * y + (tc.cls >> 1); ox = 2 * x + (tc.cls & 1); }
  const int BN = p.BN, co0 = tc.nt * BN, bw = p.bw;
  const bool relu = (p.flags & EAMM_EPI_RELU) != 0;
  const bool track1 = p.amax_out != nullptr, track2 = OUT2 && p.amax_out2 != nullptr;
  const float* bias = p.bias + co0;
  const float* asc = p.acc_scale != nullptr ? p.acc_scale + co0 : nullptr;
  const float* sc2 = OUT2 ? p.scale2 + co0 : nullptr;
  const float* sh2 = OUT2 ? p.shift2 + co0 : nullptr;
  // out-of-range rows of a partial tile compute on garbage and store nothing: their addresses are never formed
  const EpiView vo = epi_view(p.out, valid ? n : 0, valid ? oy : 0, valid ? ox : 0, co0);
  EpiView vr = vo, v2 = vo;
  if (RES) vr = epi_view(p.res, valid ? n : 0, valid ? oy : 0, valid ? ox : 0, co0);
  if (OUT2) v2 = epi_view(p.out2, valid ? n : 0, valid ? oy : 0, valid ? ox : 0, co0);
  // TMA-store coordinates of the warp's 32 pixels (tiles lie inside one image: bn == 1, H a multiple of the tile height)
  constexpr bool tma = TMA;          // compile-time: the residual / pooled variants never carry the staging code (registers)
  EpiTma t1, t2;
  if (tma) {
    const int q32 = quadrant * 32;
    const bool up2 = p.kind == EAMM_CONV_UP2_3X3;
    t1.m64 = &maps->o64; t1.m32 = &maps->o32; t1.area = st_area;
    t1.px = up2 ? (tc.cls & 1) : 0; t1.py = up2 ? (tc.cls >> 1) : 0;
    t1.x = tc.x0 + (q32 & (p.bw - 1)); t1.yr = tc.n0 * p.H + tc.y0 + (q32 >> p.bw_log2);
    t2 = t1; t2.m64 = &maps->p64; t2.m32 = &maps->p32; t2.area = st_area + 4096u;
  }
  const uint32_t taddr = tmem_acc + ((uint32_t)(quadrant * 32) << 16);
  const int cstep = (TC_EPI_WARPS / 4) * 32;
  int c0 = half * 32;
  if (c0 >= BN) return;
  uint32_t raw[32], raw2[FOLD ? 32 : 1];
  uint4 rr[RES ? 8 : 1];
  TmemLd<32>::ld(taddr + c0, raw);
  if (FOLD) TmemLd<32>::ld(taddr + BN + c0, raw2);
  for (; c0 < BN; c0 += cstep) {
    if (RES && c0 != half * 32) TmemLd<32>::ld(taddr + c0, raw);
#ifdef EAMM_EPI_PHASES
    long long ph0 = clock64();
#endif
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#ifdef EAMM_EPI_PHASES
    long long ph1 = clock64();
#endif
    float f[32];
    if (FOLD) {
#pragma unroll
      for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(raw2[FOLD ? j : 0]));
    }
    if (asc != nullptr) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0) + g);
        const float4 sv = __ldg(reinterpret_cast<const float4*>(asc + c0) + g);
        f[4 * g] = fmaf(__uint_as_float(raw[4 * g]), sv.x, b.x);
        f[4 * g + 1] = fmaf(__uint_as_float(raw[4 * g + 1]), sv.y, b.y);
        f[4 * g + 2] = fmaf(__uint_as_float(raw[4 * g + 2]), sv.z, b.z);
        f[4 * g + 3] = fmaf(__uint_as_float(raw[4 * g + 3]), sv.w, b.w);
      }
    } else {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0) + g);
        f[4 * g] = __uint_as_float(raw[4 * g]) + b.x;
        f[4 * g + 1] = __uint_as_float(raw[4 * g + 1]) + b.y;
        f[4 * g + 2] = __uint_as_float(raw[4 * g + 2]) + b.z;
        f[4 * g + 3] = __uint_as_float(raw[4 * g + 3]) + b.w;
      }
    }
    // the accumulator columns of this chunk are in f: fetch the next chunk while this one is converted and stored
    // (the residual variants prefetch the residual instead: registers)
    if (!RES && c0 + cstep < BN) {
      TmemLd<32>::ld(taddr + c0 + cstep, raw);
      if (FOLD) TmemLd<32>::ld(taddr + BN + c0 + cstep, raw2);
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
    }
#ifdef EAMM_EPI_PHASES
    {
      float keep = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) keep += f[j];
      asm volatile("" :: "f"(keep));
    }
    long long ph2 = clock64();
    if (p.prof != nullptr && quadrant == 0 && half == 0 && lane == 0) {
      atomicAdd(p.prof + blockIdx.x * 8 + 0, (unsigned long long)(ph1 - ph0));
      atomicAdd(p.prof + blockIdx.x * 8 + 1, (unsigned long long)(ph2 - ph1));
    }
#endif
    if (POOL) {
      // 2x2 average as a reduce-scatter over the window's four lanes (see epilogue_tile)
      const bool hx = lane & 1, hy = lane & bw;
      float g[16], h8[8];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float send = hx ? f[j] : f[j + 16], keep = hx ? f[j + 16] : f[j];
        g[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float send = hy ? g[j] : g[j + 8], keep = hy ? g[j + 8] : g[j];
        h8[j] = 0.25f * (keep + __shfl_xor_sync(0xffffffffu, send, bw));
      }
      if (valid) {
        epi_store<8>(vo, c0 + (hx ? 16 : 0) + (hy ? 8 : 0), h8);
        if (track1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) amax1 = fmaxf(amax1, fabsf(h8[j]));
        }
      }
      continue;
    }
    if (valid) {
      if (RES) { epi_res_load(vr, c0, rr); epi_res_add(vr, rr, f); }      // all planes' loads first, then the adds
      if (tma) epi_store_tma(vo, t1, lane, c0, f);
      else epi_store<32>(vo, c0, f);
#ifdef EAMM_EPI_PHASES
      if (p.prof != nullptr && quadrant == 0 && half == 0 && lane == 0)
        atomicAdd(p.prof + blockIdx.x * 8 + 7, (unsigned long long)(clock64() - ph2));
#endif
      if (track1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) amax1 = fmaxf(amax1, fabsf(f[j]));
      }
      if (OUT2) {
        // second output in place (f is dead after the first store)
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 sv = __ldg(reinterpret_cast<const float4*>(sc2 + c0) + g);
          const float4 tv = __ldg(reinterpret_cast<const float4*>(sh2 + c0) + g);
          f[4 * g] = fmaxf(fmaf(f[4 * g], sv.x, tv.x), 0.f);
          f[4 * g + 1] = fmaxf(fmaf(f[4 * g + 1], sv.y, tv.y), 0.f);
          f[4 * g + 2] = fmaxf(fmaf(f[4 * g + 2], sv.z, tv.z), 0.f);
          f[4 * g + 3] = fmaxf(fmaf(f[4 * g + 3], sv.w, tv.w), 0.f);
        }
        if (tma) epi_store_tma(v2, t2, lane, c0, f);
        else epi_store<32>(v2, c0, f);
        if (track2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) amax2 = fmaxf(amax2, f[j]);          // post-ReLU: non-negative
        }
      }
    }
  }
}

// Epilogue for one accumulator tile, CH columns at a time.
template <int CH>
__device__ __forceinline__ void epilogue_tile(const ConvTcParams& p, const TileCoord& tc, uint32_t tmem_acc,
                                              int quadrant, int lane, int half, float& amax1, float& amax2) {
  const int r = quadrant * 32 + lane;
  const int xl = r & (p.bw - 1);
  const int yl = (r >> p.bw_log2) & (p.bh - 1);
  const int nl = r >> (p.bw_log2 + p.bh_log2);
  const int x = tc.x0 + xl, y = tc.y0 + yl, n = tc.n0 + nl;
  bool valid = (y < p.H) && (n < p.N);
  const bool pool = p.flags & EAMM_EPI_POOL2;
  // pooled layers that only write an activation view: the reduce-scatter variant below
  const bool pool_rs = pool && p.has_out && !p.has_res && !p.has_out2 && p.out_nhwc == nullptr && p.out_nchw == nullptr &&
                       p.out.dtype != EAMM_F32 && (p.out.c_off % 8) == 0 && (p.out.c_buf % 8) == 0;
  int oy = y, ox = x, OH = p.H, OW = p.W;
  if (pool) { oy = y >> 1; ox = x >> 1; OH = p.H >> 1; OW = p.W >> 1; valid = valid && !(xl & 1) && !(yl & 1); }
  else if (p.kind == EAMM_CONV_UP2_3X3) { oy = 2 * y + (tc.cls >> 1); ox = 2 * x + (tc.cls & 1); OH = 2 * p.H; OW = 2 * p.W; }
  const uint32_t taddr = tmem_acc + ((uint32_t)(quadrant * 32) << 16);
  // chunk walk: the two warps of a quadrant alternate over the CH-column chunks
  const int cfirst = half * CH;
  const int cstep = (TC_EPI_WARPS / 4) * CH;
  for (int c0 = cfirst; c0 < p.BN; c0 += cstep) {
    uint32_t raw[CH];
    if (p.splitk > 1) {                             // reducer of a split-K tile: partials summed in split order
      const float* src = p.sk_ws + (((size_t)tc.out_tile * p.splitk) * 128 + r) * p.BN + c0;
      float acc[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) acc[j] = 0.f;
      if ((y < p.H) && (n < p.N)) {                 // (dead rows were not published: splitk_publish)
#pragma unroll
        for (int g = 0; g < CH / 8; ++g) ldg256_cg(src + 8 * g, acc + 8 * g);
        for (int s2 = 1; s2 < p.splitk; ++s2) {
          src += 128 * p.BN;
          float v[CH];
#pragma unroll
          for (int g = 0; g < CH / 8; ++g) ldg256_cg(src + 8 * g, v + 8 * g);
#pragma unroll
          for (int j = 0; j < CH; ++j) acc[j] += v[j];
        }
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) raw[j] = __float_as_uint(acc[j]);
    } else {
      TmemLd<CH>::ld(taddr + c0, raw);
      if (p.fold) {                                   // add the a_hi*b_lo columns kept at [BN, 2BN)
        uint32_t raw2[CH];
        TmemLd<CH>::ld(taddr + p.BN + c0, raw2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < CH; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(raw2[j]));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    float f[CH];
    const int co = tc.nt * p.BN + c0;
    if (p.acc_scale != nullptr) {                     // pre-scaled operands: undo 2^(e_in + e_w[co]) before the bias
#pragma unroll
      for (int g = 0; g < CH / 4; ++g) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + co) + g);
        const float4 s = __ldg(reinterpret_cast<const float4*>(p.acc_scale + co) + g);
        f[4 * g] = fmaf(__uint_as_float(raw[4 * g]), s.x, b.x);
        f[4 * g + 1] = fmaf(__uint_as_float(raw[4 * g + 1]), s.y, b.y);
        f[4 * g + 2] = fmaf(__uint_as_float(raw[4 * g + 2]), s.z, b.z);
        f[4 * g + 3] = fmaf(__uint_as_float(raw[4 * g + 3]), s.w, b.w);
      }
    } else {
#pragma unroll
    for (int g = 0; g < CH / 4; ++g) {
      float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + co) + g);
      f[4 * g] = __uint_as_float(raw[4 * g]) + b.x;
      f[4 * g + 1] = __uint_as_float(raw[4 * g + 1]) + b.y;
      f[4 * g + 2] = __uint_as_float(raw[4 * g + 2]) + b.z;
      f[4 * g + 3] = __uint_as_float(raw[4 * g + 3]) + b.w;
    }
    }
    if (p.flags & EAMM_EPI_RELU) {
#pragma unroll
      for (int j = 0; j < CH; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (pool && CH == 32 && pool_rs) {
      // 2x2 average as a reduce-scatter over the window's four lanes (x partner = lane ^ 1, y partner = lane ^ bw): each
      // lane ends up with 8 of the chunk's 32 channels of the pooled pixel, so the conversions and stores that follow run
      // on a quarter of the values (the butterfly left all four lanes with identical copies of all 32).
      const bool hx = lane & 1, hy = lane & p.bw;
      float g[16], h8[8];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float send = hx ? f[j] : f[j + 16], keep = hx ? f[j + 16] : f[j];
        g[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float send = hy ? g[j] : g[j + 8], keep = hy ? g[j + 8] : g[j];
        h8[j] = 0.25f * (keep + __shfl_xor_sync(0xffffffffu, send, p.bw));
      }
      if ((y < p.H) && (n < p.N)) {
        const int c8 = co + (hx ? 16 : 0) + (hy ? 8 : 0);
        store8(p.out, act_offset(p.out, n, oy, ox, c8), c8, h8);
        if (p.amax_out != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) amax1 = fmaxf(amax1, fabsf(h8[j]));
        }
      }
      continue;
    }
    if (pool) {
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        float s = f[j] + __shfl_xor_sync(0xffffffffu, f[j], 1);
        s += __shfl_xor_sync(0xffffffffu, s, p.bw);
        f[j] = 0.25f * s;
      }
    }
    if (valid) {
      if (p.has_res) add_chunk<CH>(p.res, act_offset(p.res, n, oy, ox, co), f, p.st256);
      if (p.has_out) {
        store_chunk<CH>(p.out, act_offset(p.out, n, oy, ox, co), co, f, p.st256);
        if (p.amax_out != nullptr) {
#pragma unroll
          for (int j = 0; j < CH; ++j) amax1 = fmaxf(amax1, fabsf(f[j]));
        }
      }
      if (p.has_out2) {
        float g2[CH];
#pragma unroll
        for (int g = 0; g < CH / 4; ++g) {
          float4 s = __ldg(reinterpret_cast<const float4*>(p.scale2 + co) + g);
          float4 t = __ldg(reinterpret_cast<const float4*>(p.shift2 + co) + g);
          g2[4 * g] = fmaxf(fmaf(f[4 * g], s.x, t.x), 0.f);
          g2[4 * g + 1] = fmaxf(fmaf(f[4 * g + 1], s.y, t.y), 0.f);
          g2[4 * g + 2] = fmaxf(fmaf(f[4 * g + 2], s.z, t.z), 0.f);
          g2[4 * g + 3] = fmaxf(fmaf(f[4 * g + 3], s.w, t.w), 0.f);
        }
        store_chunk<CH>(p.out2, act_offset(p.out2, n, oy, ox, co), co, g2, p.st256);
        if (p.amax_out2 != nullptr) {
#pragma unroll
          for (int j = 0; j < CH; ++j) amax2 = fmaxf(amax2, g2[j]);          // post-ReLU: non-negative
        }
      }
      if (p.out_nhwc != nullptr) {
        float4* dst = reinterpret_cast<float4*>(p.out_nhwc + (((long long)n * OH + oy) * OW + ox) * p.cout + co);
#pragma unroll
        for (int g = 0; g < CH / 4; ++g) dst[g] = make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
      }
      if (p.out_nchw != nullptr) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (co + j < p.out_nchw_c) {
            float o = f[j];
            if (p.flags & EAMM_EPI_SIGMOID) o = 1.f / (1.f + expf(-o));
            p.out_nchw[(((long long)n * p.out_nchw_c + co + j) * OH + oy) * OW + ox] = o;
            if (p.out_u8 != nullptr) p.out_u8[(((long long)n * OH + oy) * OW + ox) * p.out_nchw_c + co + j] = to_ubyte(o);
          }
        }
      }
    }
  }
}

// Split-K: publish this work item's fp32 partial accumulator tile and elect the reducer.  Returns true in the
// CTA whose arrival completes the tile (all its epilogue threads): it then runs epilogue_tile, which sums the
// partials.  Classic last-block pattern: data stores, __threadfence, CTA barrier, one atomicAdd on the tile's
// counter; the reducer fences again and reads with ld.global.cg.  The counter is left at zero for the next launch.
__device__ __forceinline__ bool splitk_publish(const ConvTcParams& p, const TileCoord& tc, uint32_t tmem_acc,
                                               int quadrant, int lane, int half, volatile uint32_t* flag) {
  const int r = quadrant * 32 + lane;
  const uint32_t taddr = tmem_acc + ((uint32_t)(quadrant * 32) << 16);
  float* dst = p.sk_ws + (((size_t)tc.out_tile * p.splitk + tc.split) * 128 + r) * p.BN;
  // rows outside the batch / the map carry nothing (a batch-1 call on a 4x4 map has 16 live rows of 128): not published
  const bool live = (tc.y0 + ((r >> p.bw_log2) & (p.bh - 1)) < p.H) && (tc.n0 + (r >> (p.bw_log2 + p.bh_log2)) < p.N);
  for (int c0 = half * 32; c0 < p.BN; c0 += (TC_EPI_WARPS / 4) * 32) {
    uint32_t raw[32];
    TmemLd<32>::ld(taddr + (uint32_t)c0, raw);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (live) {
#pragma unroll
      for (int g = 0; g < 4; ++g) stg256_cg(dst + c0 + 8 * g, raw + 8 * g);
    }
  }
  __threadfence();
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (threadIdx.x == 0) {
    const unsigned int old = atomicAdd(p.sk_cnt + tc.out_tile, 1u);
    const bool last = old == (unsigned int)(p.splitk - 1);
    if (last) p.sk_cnt[tc.out_tile] = 0u;
    *flag = last ? 1u : 0u;
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const bool last = *flag != 0u;
  if (last) __threadfence();
  return last;
}

// kx-in-N epilogue (7x7 -> <=4 channels, sigmoid, NCHW fp32): accumulator row p holds, for the input
// column x0-3+p, the partial sums D[p][kx*4+co] over (ky, channels).  out[x0+j][co] = bias +
// sum_kx D[j+kx][kx*4+co]; the shifted rows are exchanged through shared memory (S, stride 29).
__device__ __forceinline__ void epilogue_kxn(const ConvTcParams& p, const TileCoord& tc, uint32_t tmem_acc,
                                             int quadrant, int lane, float* S) {
  const int r = quadrant * 32 + lane;
  uint32_t raw[32];
  TmemLd<32>::ld(tmem_acc + ((uint32_t)(quadrant * 32) << 16), raw);
  if (p.fold) {
    uint32_t raw2[32];
    TmemLd<32>::ld(tmem_acc + ((uint32_t)(quadrant * 32) << 16) + 32u, raw2);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < 28; ++c) raw[c] = __float_as_uint(__uint_as_float(raw[c]) + __uint_as_float(raw2[c]));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int c = 0; c < 28; ++c) S[r * 29 + c] = __uint_as_float(raw[c]);
  asm volatile("bar.sync 1, 128;" ::: "memory");
  const int x = tc.x0 + r;
  if (r < 122 && x < p.W && tc.y0 < p.H && tc.n0 < p.N) {
    for (int co = 0; co < p.out_nchw_c; ++co) {
      float acc = __ldg(p.bias + co);
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) acc += S[(r + kx) * 29 + kx * 4 + co];
      if (p.flags & EAMM_EPI_SIGMOID) acc = 1.f / (1.f + expf(-acc));
      p.out_nchw[(((long long)tc.n0 * p.out_nchw_c + co) * p.H + tc.y0) * p.W + x] = acc;
      if (p.out_u8 != nullptr) p.out_u8[(((long long)tc.n0 * p.H + tc.y0) * p.W + x) * p.out_nchw_c + co] = to_ubyte(acc);
    }
  }
}

// kx-in-N epilogue for the 112-column variants (kxn == 2: (dr, kx, co) = 4*7*4, NCHW + sigmoid; kxn == 3:
// (kx, co) = 7*16, fp32 NHWC logits).  All eight epilogue warps move the accumulator to shared memory (row
// stride 113 floats: conflict-free for lanes = rows; single buffer, the TMEM ring still decouples the MMA warp),
// then each thread sums the kx-shifted entries of a few outputs, in the same order as the 32-column variant.
constexpr int KXW_COLS = 112, KXW_LD = 113;
__device__ __forceinline__ void epilogue_kxn_wide(const ConvTcParams& p, const TileCoord& tc, uint32_t tmem_acc,
                                                  int quadrant, int lane, int half, float* S) {
  const int r = quadrant * 32 + lane;
  const uint32_t taddr = tmem_acc + ((uint32_t)(quadrant * 32) << 16);
  if (p.kxn == 2 && p.s_nc != 4) {
    // compact rows: only the nc valid channels of every (dr, kx) group are kept, row stride 28*nc + 1 floats
    // (odd): the smaller buffer buys a fourth pipeline stage.  A 32-column load per output row dr (28 used).
    const int nc = p.s_nc, ld = 28 * nc + 1;
    for (int dr = half; dr < 4; dr += 2) {
      uint32_t raw[32];
      TmemLd<32>::ld(taddr + (uint32_t)(dr * 28), raw);
      if (p.fold) {
        uint32_t raw2[32];
        TmemLd<32>::ld(taddr + (uint32_t)(KXW_COLS + dr * 28), raw2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 28; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(raw2[j]));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float* dst = S + r * ld + dr * 7 * nc;
#pragma unroll
      for (int j = 0; j < 28; ++j)
        if ((j & 3) < nc) dst[(j >> 2) * nc + (j & 3)] = __uint_as_float(raw[j]);
    }
  } else
  for (int c0 = half * 16; c0 < KXW_COLS; c0 += 32) {
    uint32_t raw[16];
    TmemLd<16>::ld(taddr + (uint32_t)c0, raw);
    if (p.fold) {
      uint32_t raw2[16];
      TmemLd<16>::ld(taddr + (uint32_t)(KXW_COLS + c0), raw2);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(raw2[j]));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) S[r * KXW_LD + c0 + j] = __uint_as_float(raw[j]);
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const int tid = (half * 4 + quadrant) * 32 + lane;
  if (p.kxn == 2) {
    const int nc = p.out_nchw_c, snc = p.s_nc, ld = 28 * snc + 1;
    for (int i = tid; i < 4 * nc * 128; i += 256) {
      const int rr = i & 127, q = i >> 7;
      const int dr = q / nc, co = q - dr * nc;
      const int x = tc.x0 + rr, y = tc.y0 + dr;
      if (rr < 122 && x < p.W && y < p.H && tc.n0 < p.N) {
        float acc = __ldg(p.bias + co);
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) acc += S[(rr + kx) * ld + (dr * 7 + kx) * snc + co];
        if (p.flags & EAMM_EPI_SIGMOID) acc = 1.f / (1.f + expf(-acc));
        p.out_nchw[(((long long)tc.n0 * nc + co) * p.H + y) * p.W + x] = acc;
        if (p.out_u8 != nullptr) p.out_u8[(((long long)tc.n0 * p.H + y) * p.W + x) * nc + co] = to_ubyte(acc);
      }
    }
  } else {
    for (int i = tid; i < 128 * 16; i += 256) {
      const int co = i & 15, px = i >> 4;
      const int xl = px & (p.bw - 1), yl = px >> p.bw_log2;
      const int y = tc.y0 + yl;
      if (y < p.H && tc.n0 < p.N) {
        float acc = __ldg(p.bias + co);
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int xs = xl + kx - 3;
          if (xs >= 0 && xs < p.W) acc += S[(px + kx - 3) * KXW_LD + kx * 16 + co];
        }
        p.out_nhwc[(((long long)tc.n0 * p.H + y) * p.W + xl) * p.cout + co] = acc;
      }
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");      // S is reused by the next tile
}

// INSTR = true compiles the bring-up instrumentation (EAMM_TC_PROF cycle counters, EAMM_TC_DEBUG
// role isolation); the production instantiation carries none of it.
// CTA2 = true: CTA pairs (cluster of 2) run `tcgen05.mma.cta_group::2`, M = 256 pixels x N = 256
// couts per instruction.  Each CTA stages its own 128-pixel A tile and HALF of the weight tile, so
// the shared-memory operand fetch per SM and the weight traffic per SM halve (the N=256 layers are
// bound by that fetch: ~190 cycles per K=16 step with cta_group::1 vs 128 of math).  The leader CTA
// (rank 0) owns the "full" barriers and issues every MMA; both CTAs run producers and epilogues.
template <bool INSTR, bool CTA2>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ StoreMaps tmS, const ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * 16 + 4 + 8 + 1];      // + halo-tile scheme: A ring full [36..39] / empty [40..43]
  __shared__ uint32_t tmem_base_smem;
  __shared__ uint32_t sk_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // dynamic smem is only guaranteed 16B-aligned by the ABI: align the ring to 1024 B by hand
  uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t bar0 = smem_u32(bars);
  // opaque copies: keep ptxas from re-deriving the shared-window base (S2R) inside the hot loops
  asm volatile("mov.u32 %0, %0;" : "+r"(smem_base));
  asm volatile("mov.u32 %0, %0;" : "+r"(bar0));
  const uint32_t a_slot = (uint32_t)p.a_slot_bytes;
  const uint32_t fold = (uint32_t)p.fold;
  // B sub-slot size (a CTA of a pair stages only its half of the N tile)
  const uint32_t b_bytes = CTA2 ? (uint32_t)p.BN * (fold ? 128u : 64u)
                                : (uint32_t)p.BN * 128u * (p.halo ? 7u : (fold ? 2u : 1u));
  const uint32_t KS = (uint32_t)p.ksub;                         // 64-channel sub-chunks per pipeline stage
  const uint32_t b_res = (uint32_t)p.b_res;
  const uint32_t stage_bytes = KS * (a_slot + (b_res ? 0u : b_bytes));
  const uint32_t sub_tx = (p.halo ? 134u * 128u : (uint32_t)TC_A_BYTES) + (b_res ? 0u : b_bytes);   // bytes of a type-0 chunk
  const uint32_t bres_base = smem_base + (uint32_t)p.num_stages * stage_bytes;      // resident weights behind the ring
  const uint32_t b_half = (uint32_t)p.BN * (CTA2 ? 64u : 128u);        // fold: offset of the b_lo rows inside a B slot
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (16 + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (32 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (34 + a); };
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), CTA2 ? 2 * TC_EPI_WARPS : TC_EPI_WARPS); }
    for (int a = 0; a < 8; ++a) mbar_init(bar0 + 8u * (36 + a), 1);
    mbar_init(bar0 + 8u * 44u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == TC_MMA_WARP) {
    if (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(&tmem_base_smem)), "r"(TC_TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(&tmem_base_smem)), "r"(TC_TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();          // the peer's barriers must be initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int KC = p.ntap * p.cin_chunks * (p.fold == 1 ? 2 : (p.fold == 2 ? 1 : p.passes)) / (p.splitk > 1 ? p.splitk : 1);
  const uint32_t total_tiles = (uint32_t)p.total_tiles;
  const int dbg = INSTR ? p.debug : 0;
  // tile walk: CTA i takes tiles i, i+grid, ...; a CTA pair takes adjacent M tiles (2c + rank), (2c + rank) + grid, ...
  const uint32_t tile0 = blockIdx.x, tile_step = gridDim.x;

  // Both single-issuer roles run as warp-uniform loops (all 32 lanes wait, one elected lane issues) so
  // ptxas keeps descriptors/coordinates in uniform registers.  One pipeline stage carries KS
  // consecutive 64-channel K chunks: the barrier round trip and loop bookkeeping of these
  // latency-bound single-warp loops are paid once per stage, not once per chunk.
  if (warp >= TC_EPI_WARPS && warp < TC_EPI_WARPS + TC_PRODUCERS) {
    // ================================================================ TMA producers
    // TC_PRODUCERS warps share the ring round-robin: producer w fills the stages whose running index
    // is congruent to w.  A lone warp retires a dependent scalar instruction every ~8-10 cycles, so
    // one producer could not feed short-K layers (measured 600-800 cycles per 64-channel chunk).
    const uint32_t w = (uint32_t)(warp - TC_EPI_WARPS);
    if (!INSTR && p.ah) producer_ah<CTA2>(p, &tmA, &tmB, smem_base, bar0, w, cta_rank, tile0, tile_step, total_tiles);
    else {
    const uint32_t ntap = (uint32_t)p.ntap;
    const uint32_t chunk_shift = (uint32_t)p.chunk_shift, chunk_mask = (1u << chunk_shift) - 1u;
    const int passes = p.passes, kind = p.kind;
    const uint32_t nstages = (uint32_t)p.num_stages;
    const uint32_t SPT = ((uint32_t)KC + KS - 1u) / KS;          // stages per tile
    const uint32_t foldT = (uint32_t)KC >> 1;                    // fold == 1: (tap, chunk) pairs per tile
    const bool haloish = p.halo || p.kxn;
    uint32_t slot = 0, phase = 0, si = w;
    for (uint32_t i = 0; i < w; ++i) { if (++slot == nstages) { slot = 0; phase ^= 1u; } }
    if (b_res && w == 0u && tile0 < total_tiles && elect_one()) {
      // resident weights (single CTA, one N tile): chunk kc = 64 K elements of every row ([hi rows | lo rows] when folded)
      mbar_expect_tx(bar0 + 8u * 44u, (uint32_t)KC * b_bytes);
      for (int kc = 0; kc < KC; ++kc) {
        const int bcol = kc * (p.f16in ? 128 : 64);                  // fp16 maps are addressed in bytes
        tma_load_2d(bres_base + (uint32_t)kc * b_bytes, &tmB, bar0 + 8u * 44u, bcol, 0);
        if (fold) tma_load_2d(bres_base + (uint32_t)kc * b_bytes + b_half, &tmB, bar0 + 8u * 44u, bcol, p.b_rows_total);
      }
    }
    __syncwarp();
    long long pw = 0, pstart = 0;
    if (INSTR) pstart = clock64();
    for (uint32_t tile = tile0; tile < total_tiles; tile += tile_step) {
      const TileCoord tc = decode_tile(p, tile);
      // CTA pair: this CTA stages rows [rank*BN/2, +BN/2) of the N tile; "full" lives in the leader CTA
      const int brow = tc.cls * p.cout + tc.nt * p.BN + (CTA2 ? (int)cta_rank * (p.BN >> 1) : 0);
      for (; si < SPT; si += TC_PRODUCERS) {
        const uint32_t kc0 = si * KS;
        const uint32_t nsub = (uint32_t)KC - kc0 < KS ? (uint32_t)KC - kc0 : KS;
        const uint32_t sa = smem_base + slot * stage_bytes;
        const uint32_t fb = CTA2 ? mapa_shared(full_bar((int)slot), 0u) : full_bar((int)slot);
        long long t0 = 0;
        if (INSTR) t0 = clock64();
        mbar_wait(empty_bar((int)slot), phase ^ 1u);
        if (INSTR) pw += clock64() - t0;
        if (elect_one()) {
          if (INSTR && (dbg == 1 || dbg == 4 || dbg == 5)) mbar_arrive(fb);
          else {
            // fold == 1 chunk order: kc 0 = type 0 (zero-initialises both accumulator halves), kc 1..T =
            // type 1 (a_lo x b_hi, small terms first), kc T+1..2T-1 = type 0; type 1 carries half the B bytes
            uint32_t tx = nsub * sub_tx;
            if (fold == 1) {
              const uint32_t lo1 = kc0 > 1u ? kc0 : 1u, hi1 = kc0 + nsub < foldT + 1u ? kc0 + nsub : foldT + 1u;
              if (hi1 > lo1) tx -= (hi1 - lo1) * b_half;
            }
            if (CTA2) { if (cta_rank == 0u) mbar_expect_tx(full_bar((int)slot), 2u * tx); }   // both CTAs' bytes
            else mbar_expect_tx(fb, tx);
            for (uint32_t sub = 0; sub < nsub; ++sub) {
              // K order = (pass, tap, chunk).  In split mode the two cross terms (a_hi*b_lo, a_lo*b_hi)
              // are accumulated first, while the TMEM accumulator is still small, and the dominant
              // a_hi*b_hi term last: the tensor core truncates on every accumulate, so the bias it
              // leaves scales with |accumulator| x (number of steps taken at that magnitude).
              // (fold mode keeps a_hi*b_lo in its own accumulator columns instead.)
              const uint32_t kc = kc0 + sub;
              const uint32_t type1 = (fold == 1 && kc >= 1u && kc <= foldT) ? 1u : 0u;
              const uint32_t kq = fold == 1 ? (kc == 0u ? 0u : (type1 ? kc - 1u : kc - foldT))
                                            : kc + (uint32_t)tc.split * (uint32_t)KC;     // split-K: this item's K range
              uint32_t cc, q;
              int cbase, bcol;
              if (p.f16in) {
                // byte coordinates (uint8 tensor maps).  mix: chunks [0, n8) = a_hi8 (plane 1, second half) x w_lo8,
                // [n8, 2 n8) = a_lo8 x w_hi8, 128 channels each (mix64: [0, ntap) = [a_lo8 | a_hi8] x [w_hi8 | w_lo8], 64
                // channels of both); then a_hi x w_hi, 64 fp16 channels each
                const uint32_t nf8 = (uint32_t)p.nf8;
                if (kq < nf8 && p.mix64) {
                  cc = 0u; q = kq;
                  cbase = 2 * p.a_c_buf;                       // (c_off == 0: the whole 64-channel pixel)
                } else if (kq < nf8) {
                  const uint32_t n8 = nf8 >> 1;
                  const uint32_t second = kq >= n8 ? 1u : 0u, i8 = kq - second * n8;
                  cc = i8 & (chunk_mask >> 1); q = i8 >> (chunk_shift - 1u);
                  cbase = 2 * p.a_c_buf + (second ? 0 : p.a_c_buf) + p.a_c_off + (int)cc * 128;
                } else {
                  const uint32_t i16 = kq - nf8;
                  cc = i16 & chunk_mask; q = i16 >> chunk_shift;
                  cbase = 2 * (p.a_c_off + (int)cc * 64);
                }
                bcol = (int)kq * 128;
              } else {
                cc = kq & chunk_mask; q = kq >> chunk_shift;                        // q = pass * ntap + tap
                const uint32_t psb = q >= 2u * ntap ? 2u : (q >= ntap ? 1u : 0u);
                cbase = p.a_c_off + (((passes == 3 && psb == 1u) || type1) ? p.a_c_buf : 0) + (int)cc * 64;
                bcol = (int)kq * 64;
              }
              const uint32_t ps = p.f16in ? 0u : (q >= 2u * ntap ? 2u : (q >= ntap ? 1u : 0u));
              const int t = (int)(q - ps * ntap);
              int dy, dx;
              if (haloish) { dy = t - 3; dx = p.kxn == 3 ? 0 : -3; }
              else if (kind == EAMM_CONV_ROW7_PACKED) { dy = t; dx = 0; cbase = 0; }   // both planes inside the K window
              else if (kind == EAMM_CONV_UP2_3X3) { dy = (tc.cls >> 1) - 1 + (t >> 1); dx = (tc.cls & 1) - 1 + (t & 1); }
              else if (kind == EAMM_CONV_3X3) { const int ty = (t * 11) >> 5; dy = ty - 1; dx = t - 3 * ty - 1; }
              else { const int ty = (t * 37) >> 8; dy = ty - 3; dx = t - 7 * ty - 3; }   // 7x7 per-tap
              const uint32_t sB = sa + KS * a_slot + sub * b_bytes;
              if (CTA2) {
                tma2_load_4d(sa + sub * a_slot, &tmA, fb, cbase, tc.x0 + dx, tc.y0 + dy, tc.n0);
                tma2_load_2d(sB, &tmB, fb, bcol, brow);
                if (fold && !type1) tma2_load_2d(sB + b_half, &tmB, fb, bcol, brow + p.b_rows_total);
              } else {
                tma_load_4d(sa + sub * a_slot, &tmA, fb, cbase, tc.x0 + dx, tc.y0 + dy, tc.n0);
                if (!b_res) {
                  tma_load_2d(sB, &tmB, fb, bcol, brow);
                  if (fold && !type1) tma_load_2d(sB + b_half, &tmB, fb, bcol, brow + p.b_rows_total);
                }
              }
            }
          }
        }
        slot += TC_PRODUCERS;
        while (slot >= nstages) { slot -= nstages; phase ^= 1u; }
      }
      si -= SPT;
    }
    if (INSTR && p.prof && w == 0 && lane == 0) {
      p.prof[blockIdx.x * 8 + 0] = pw;                             // producer 0: cycles waiting for a free slot
      p.prof[blockIdx.x * 8 + 1] = clock64() - pstart;             // producer 0: total
    }
    }
  } else if (warp == TC_MMA_WARP && (!CTA2 || cta_rank == 0u)) {
    // ================================================================ MMA issuer (leader CTA of a pair)
    if (!INSTR && p.ah) {
      const int ks = p.ksub;
      if (p.kind == EAMM_CONV_3X3) {
        if (ks == 3) mma_issuer_ah<CTA2, 0, 3>(p, smem_base, bar0, tmem_base, tile0, tile_step, total_tiles);
        else mma_issuer_ah<CTA2, 0, 1>(p, smem_base, bar0, tmem_base, tile0, tile_step, total_tiles);
      } else {
        if (ks == 4) mma_issuer_ah<CTA2, 1, 4>(p, smem_base, bar0, tmem_base, tile0, tile_step, total_tiles);
        else if (ks == 2) mma_issuer_ah<CTA2, 1, 2>(p, smem_base, bar0, tmem_base, tile0, tile_step, total_tiles);
        else mma_issuer_ah<CTA2, 1, 1>(p, smem_base, bar0, tmem_base, tile0, tile_step, total_tiles);
      }
    }
    else if (!INSTR && p.lean) {
      if (p.mix) mma_issuer_lean<CTA2, 1>(p, smem_base, bar0, tmem_base, (uint32_t)KC, KS, a_slot, b_bytes, b_half, tile0, tile_step, total_tiles);
      else if (fold == 1u && CTA2) mma_issuer_lean<CTA2, 3>(p, smem_base, bar0, tmem_base, (uint32_t)KC, KS, a_slot, b_bytes, b_half, tile0, tile_step, total_tiles);
      else if (fold == 1u) mma_issuer_lean<CTA2, 2>(p, smem_base, bar0, tmem_base, (uint32_t)KC, KS, a_slot, b_bytes, b_half, tile0, tile_step, total_tiles);
      else mma_issuer_lean<CTA2, 0>(p, smem_base, bar0, tmem_base, (uint32_t)KC, KS, a_slot, b_bytes, b_half, tile0, tile_step, total_tiles,
                                    b_res ? bres_base : 0u);
    } else {
    // instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major A/B, N>>3 at 17, M>>4 at 24
    // (a/b format code 0 is F16 under kind::f16 and E4M3 under kind::f8f6f4: fp16 and mixed inputs use one descriptor for both)
    const uint32_t fmt = p.f16in ? 0u : ((1u << 7) | (1u << 10));
    const uint32_t idesc1 = (1u << 4) | fmt | ((uint32_t)(p.BN >> 3) << 17) | (((CTA2 ? 256u : 128u) >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | fmt | ((uint32_t)(p.BN >> 2) << 17) | (((CTA2 ? 256u : 128u) >> 4) << 24);   // N = 2*BN
    const uint32_t n8x2 = p.mix ? (uint32_t)p.nf8 : 0u;          // mixed input: K chunks [0, nf8) are the fp8 cross terms
    const int nstages = p.num_stages, halo = p.halo, BN = p.BN;
    int stage = 0; uint32_t phase = 0; uint32_t as = 0, aphase = 0;
    uint32_t sa = smem_base, fb = full_bar(0), eb = empty_bar(0);
    long long pm0 = 0, pm1 = 0, pstart = 0;
    if (INSTR) pstart = clock64();
    for (uint32_t tile = tile0; tile < total_tiles; tile += tile_step) {
      long long t0 = 0;
      if (INSTR) t0 = clock64();
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      if (INSTR) pm0 += clock64() - t0;
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + as * 256u;
      const uint32_t kbase = (n8x2 && p.splitk > 1) ? (uint32_t)decode_tile(p, tile).split * (uint32_t)KC : 0u;
      for (int kc = 0; kc < KC; kc += (int)KS) {
        const uint32_t nsub = (uint32_t)(KC - kc) < KS ? (uint32_t)(KC - kc) : KS;
        if (INSTR) t0 = clock64();
        mbar_wait(fb, phase);
        if (INSTR) pm1 += clock64() - t0;
        tc_fence_after();
        if (elect_one()) {
          for (uint32_t sub = 0; sub < nsub; ++sub) {
            const uint32_t sA = sa + sub * a_slot, sB = sa + KS * a_slot + sub * b_bytes;
            const uint32_t first = (kc | (int)sub) ? 1u : 0u;
            const uint32_t kcs = (uint32_t)kc + sub;
            const uint32_t idesc = (fold == 2 || (fold == 1 && (kcs == 0u || kcs > ((uint32_t)KC >> 1)))) ? idesc2 : idesc1;
            if (INSTR && (dbg == 2 || dbg == 4 || dbg == 5)) {
            } else if (INSTR && dbg == 3) {
              tc_mma_bf16(tmem_acc, make_sw128_desc(sA), make_sw128_desc(sB), idesc, first);
            } else if (halo) {
              // one halo row of 134 pixels serves the 7 kx taps: tap kx reads rows [kx, kx+128)
              // (measured on B200: the 128B swizzle is a function of the absolute smem address, so a
              //  row-shifted start address needs no base_offset in the descriptor)
#pragma unroll 1
              for (int kx = 0; kx < 7; ++kx) {
                const uint64_t da = make_sw128_desc(sA + (uint32_t)kx * 128u);
                const uint64_t db = make_sw128_desc(sB + (uint32_t)(kx * BN) * 128u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  tc_mma_bf16(tmem_acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (first | (uint32_t)kx | (uint32_t)k) ? 1u : 0u);
              }
            } else if (kbase + kcs < n8x2) {
              // fp8 cross-term chunk (128 e4m3 channels = one 128-byte row): four K = 32 steps, 32 bytes apart
              const uint64_t da = make_sw128_desc(sA), db = make_sw128_desc(sB);
              if (CTA2) {
                tc2_mma_f8(tmem_acc, da, db, idesc1, first);
                tc2_mma_f8(tmem_acc, da + 2, db + 2, idesc1, 1u);
                tc2_mma_f8(tmem_acc, da + 4, db + 4, idesc1, 1u);
                tc2_mma_f8(tmem_acc, da + 6, db + 6, idesc1, 1u);
              } else {
                tc_mma_f8(tmem_acc, da, db, idesc1, first);
                tc_mma_f8(tmem_acc, da + 2, db + 2, idesc1, 1u);
                tc_mma_f8(tmem_acc, da + 4, db + 4, idesc1, 1u);
                tc_mma_f8(tmem_acc, da + 6, db + 6, idesc1, 1u);
              }
            } else {
              const uint64_t da = make_sw128_desc(sA), db = make_sw128_desc(sB);
              if (CTA2) {
                // pair + fold: a CTA's B slot is [its half of b_hi | its half of b_lo]; an N = 2*BN step would
                // interleave the columns as [hi|lo|hi|lo], so the type-0 chunk runs as two N = BN steps instead
                // (b_hi halves -> columns [0,BN), b_lo halves -> [BN,2BN)): same column layout and the same
                // per-column accumulation order as the single-CTA kernel, i.e. bit-identical results
                tc2_mma_bf16(tmem_acc, da, db, idesc1, first);
                tc2_mma_bf16(tmem_acc, da + 2, db + 2, idesc1, 1u);
                tc2_mma_bf16(tmem_acc, da + 4, db + 4, idesc1, 1u);
                tc2_mma_bf16(tmem_acc, da + 6, db + 6, idesc1, 1u);
                if (fold == 1 && idesc == idesc2) {
                  const uint64_t dl = db + (uint64_t)(b_half >> 4);
                  const uint32_t acc2 = tmem_acc + (uint32_t)BN;
                  tc2_mma_bf16(acc2, da, dl, idesc1, first);
                  tc2_mma_bf16(acc2, da + 2, dl + 2, idesc1, 1u);
                  tc2_mma_bf16(acc2, da + 4, dl + 4, idesc1, 1u);
                  tc2_mma_bf16(acc2, da + 6, dl + 6, idesc1, 1u);
                }
              } else {
                tc_mma_bf16(tmem_acc, da, db, idesc, first);
                tc_mma_bf16(tmem_acc, da + 2, db + 2, idesc, 1u);
                tc_mma_bf16(tmem_acc, da + 4, db + 4, idesc, 1u);
                tc_mma_bf16(tmem_acc, da + 6, db + 6, idesc, 1u);
              }
            }
          }
          if (CTA2) { tc2_commit_mc(eb); if (kc + (int)KS >= KC) tc2_commit_mc(tfull_bar(as)); }
          else { tc_commit(eb); if (kc + (int)KS >= KC) tc_commit(tfull_bar(as)); }
        }
        __syncwarp();
        ++stage; sa += stage_bytes; fb += 8; eb += 8;
        if (stage == nstages) { stage = 0; phase ^= 1u; sa = smem_base; fb = full_bar(0); eb = empty_bar(0); }
      }
      as ^= 1u; if (as == 0) aphase ^= 1u;
    }
    if (INSTR && p.prof && lane == 0) {
      p.prof[blockIdx.x * 8 + 2] = pm0;                            // MMA: waiting for a free accumulator
      p.prof[blockIdx.x * 8 + 3] = pm1;                            // MMA: waiting for operands
      p.prof[blockIdx.x * 8 + 4] = clock64() - pstart;             // MMA: total
    }
    }
  } else if (warp < TC_EPI_WARPS) {
    // ================================================================ epilogue warps (TMEM lanes by warp%4)
    const int quadrant = warp & 3, half = warp >> 2;
    int as = 0; uint32_t aphase = 0;
    long long pe = 0, pstart = 0;
    float amax1 = 0.f, amax2 = 0.f;              // running max |out|, |out2| of this thread (calibration statistic)
    const bool eprof = p.prof != nullptr;
    if (eprof) pstart = clock64();
    float* kxn_smem = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) +
                                               (size_t)p.num_stages * stage_bytes);
    const uint32_t st_area = smem_base + (uint32_t)p.st_base + (uint32_t)(warp * p.st_stride);     // TMA-store staging of this warp
    for (uint32_t tile = tile0; tile < total_tiles; tile += tile_step) {
      const TileCoord tc = decode_tile(p, tile);
      long long t0 = 0;
      if (eprof) t0 = clock64();
      mbar_wait(tfull_bar(as), aphase);
      if (eprof) pe += clock64() - t0;
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(as * 256);
      if (INSTR && dbg >= 5) {                     // 5: protocol only, 6: real main loop, no epilogue work
      } else if (p.kxn == 1) { if (half == 0) epilogue_kxn(p, tc, tmem_acc, quadrant, lane, kxn_smem + as * (128 * 29)); }
      else if (p.kxn) epilogue_kxn_wide(p, tc, tmem_acc, quadrant, lane, half, kxn_smem);
      else if (p.splitk > 1) {
        if (splitk_publish(p, tc, tmem_acc, quadrant, lane, half, &sk_flag)) epilogue_tile<32>(p, tc, tmem_acc, quadrant, lane, half, amax1, amax2);
      }
      else if (p.epi_fast) {
        for (int g = 0; g < p.ah_g; ++g) {          // halo-tile UP2: the item's parity classes sit side by side in the accumulator
          TileCoord tg = tc; tg.cls = tc.cls + g;
          const uint32_t ta = tmem_acc + (uint32_t)(g * p.BN);
          if (p.epi_fast == 1) {
            if (p.tma_st) epilogue_fast<false, false, false, false, true>(p, tg, ta, quadrant, lane, half, amax1, amax2, &tmS, st_area);
            else epilogue_fast<false, false, false, false, false>(p, tg, ta, quadrant, lane, half, amax1, amax2, &tmS, st_area);
          }
          else if (p.epi_fast == 2) epilogue_fast<true, false, false, false, false>(p, tg, ta, quadrant, lane, half, amax1, amax2, &tmS, st_area);
          else if (p.epi_fast == 3) epilogue_fast<false, true, false, false, false>(p, tg, ta, quadrant, lane, half, amax1, amax2, &tmS, st_area);
          else if (p.epi_fast == 4) epilogue_fast<false, true, true, false, false>(p, tg, ta, quadrant, lane, half, amax1, amax2, &tmS, st_area);
          else if (p.epi_fast == 5) {
            if (p.tma_st) epilogue_fast<false, false, false, true, true>(p, tg, ta, quadrant, lane, half, amax1, amax2, &tmS, st_area);
            else epilogue_fast<false, false, false, true, false>(p, tg, ta, quadrant, lane, half, amax1, amax2, &tmS, st_area);
          }
          else epilogue_fast<true, false, false, true, false>(p, tg, ta, quadrant, lane, half, amax1, amax2, &tmS, st_area);
        }
      }
      else if (p.ah_g > 1) {                       // halo-tile UP2: the item's parity classes sit side by side in the accumulator
        for (int g = 0; g < p.ah_g; ++g) {
          TileCoord tg = tc; tg.cls = tc.cls + g;
          if (p.BN % 32 == 0) epilogue_tile<32>(p, tg, tmem_acc + (uint32_t)(g * p.BN), quadrant, lane, half, amax1, amax2);
          else epilogue_tile<16>(p, tg, tmem_acc + (uint32_t)(g * p.BN), quadrant, lane, half, amax1, amax2);
        }
      }
      else if (p.BN % 32 == 0) epilogue_tile<32>(p, tc, tmem_acc, quadrant, lane, half, amax1, amax2);
      else epilogue_tile<16>(p, tc, tmem_acc, quadrant, lane, half, amax1, amax2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_cluster(mapa_shared(tempty_bar(as), 0u));     // the leader's MMA warp waits for both CTAs
        else mbar_arrive(tempty_bar(as));
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (p.tma_st && lane == 0) bulk_wait0();       // this warp's bulk stores have landed
    if (p.amax_out != nullptr || p.amax_out2 != nullptr) {
      // values are non-negative: the integer order of their bit patterns is the float order
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        amax1 = fmaxf(amax1, __shfl_xor_sync(0xffffffffu, amax1, o));
        amax2 = fmaxf(amax2, __shfl_xor_sync(0xffffffffu, amax2, o));
      }
      if (lane == 0) {
        if (p.amax_out != nullptr) atomicMax(reinterpret_cast<int*>(p.amax_out), __float_as_int(amax1));
        if (p.amax_out2 != nullptr) atomicMax(reinterpret_cast<int*>(p.amax_out2), __float_as_int(amax2));
      }
    }
    if (eprof && warp == 0 && lane == 0) {
      p.prof[blockIdx.x * 8 + 5] = pe;                             // epilogue warp 0: waiting for an accumulator
      p.prof[blockIdx.x * 8 + 6] = clock64() - pstart;             // epilogue: total
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();          // no CTA of the pair may exit while the other can still signal it
  if (warp == TC_MMA_WARP) {
    tc_fence_after();
    if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

static int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}

}  // namespace eamm

using namespace eamm;

/* Which 7x7 scheme eamm_conv_tc uses (decides the weight packing): 0 = one TMA tile per tap,
 * 1 = halo row with kx-shifted smem views, 2 = kx taps in the N dimension (<= 4 NCHW channels). */
extern "C" int eamm_conv_tc_uses_halo(int kind, int w, int cout, int out_nchw_c) {
  const char* e = getenv("EAMM_TC_HALO");
  const int halo_env = e ? atoi(e) : 1;
  if (kind != EAMM_CONV_7X7 || halo_env <= 0 || w % 128 != 0) return 0;
  if (out_nchw_c >= 1 && out_nchw_c <= 4 && halo_env != 3) return 2;
  return cout <= 32 ? 1 : 0;
}

/* Whether eamm_conv_tc folds the weight planes into the N axis for this layer (decides the packing):
 * 0 = K holds (pass, tap, channel); 1 = rows [hi block | lo block], K = (tap, channel), type-0/1 chunks;
 * 2 = packed first conv with both row blocks used by every chunk. */
extern "C" int eamm_conv_tc_fold(int kind, int split, int cout, int halo_scheme) {
  const char* e = getenv("EAMM_TC_FOLD");
  const int fold_env = e ? atoi(e) : 1;
  if (!fold_env || !split || halo_scheme == 1 || (halo_scheme != 2 && cout > 128)) return 0;
  return kind == EAMM_CONV_ROW7_PACKED ? 2 : 1;
}

static int conv_tc_run(const eamm_conv_args* a, void* stream, int* query);

extern "C" int eamm_conv_tc(const eamm_conv_args* a, void* stream) { return conv_tc_run(a, stream, nullptr); }

/* Dry run of eamm_conv_tc's planning: fills out[0..5] = {N tile, 7x7 scheme, fold, K chunks per stage,
 * CTA pairs (bit 0) / halo-tile scheme (bit 1) / split factor (bits 8+), pipeline stages} for these arguments (a->weight_fold is ignored) without launching anything. */
extern "C" int eamm_conv_tc_query(const eamm_conv_args* a, int* out) {
  if (!out) return EAMM_ERR_ARG;
  return conv_tc_run(a, nullptr, out);
}

static int conv_tc_run(const eamm_conv_args* a, void* stream, int* query) {
  int rc = conv_check_args(a, 16);
  if (rc) return rc;
  const eamm_act* in = a->in;
  const bool row7 = a->kind == EAMM_CONV_ROW7_PACKED;
  if (in->dtype != EAMM_BF16 && in->dtype != EAMM_F16) return EAMM_ERR_DTYPE;
  const bool f16in = in->dtype == EAMM_F16, mix = f16in && in->planes == 2;
  const bool mix64 = mix && a->cin == 64 && in->c_buf == 64 && in->c_off == 0;
  if (mix && (row7 || a->kind == EAMM_CONV_7X7 || (a->cin % 128 && !mix64))) return EAMM_ERR_UNSUPPORTED;
  if (row7) {
    if (in->c != 8 || in->c_buf != 8 || in->c_off != 0 || in->planes != 1 || a->cin != 8) return EAMM_ERR_SHAPE;
    if (a->pack_passes != 1 && a->pack_passes != 2) return EAMM_ERR_ARG;
  } else if (in->c_buf % 64 || in->c_off % 64 || a->cin % 64) {
    return EAMM_ERR_ALIGN;
  }
  const eamm_act* views[3] = {a->out, a->out2, a->residual};
  for (int i = 0; i < 3; ++i)
    if (views[i] && ((views[i]->dtype != EAMM_BF16 && views[i]->dtype != EAMM_F16) || views[i]->c_off % 8 || views[i]->c_buf % 8))
      return EAMM_ERR_DTYPE;
  if (a->residual && a->residual->dtype == EAMM_F16 && a->residual->planes != 1) return EAMM_ERR_UNSUPPORTED;
  if ((uintptr_t)in->data % 16 || (uintptr_t)a->weight % 16) return EAMM_ERR_ALIGN;

  ConvTcParams p;
  p.N = in->n; p.H = in->h; p.W = in->w;
  int wl = ilog2_exact(in->w), hl = ilog2_exact(in->h);
  if (wl < 0 || hl < 0) return EAMM_ERR_UNSUPPORTED;            // power-of-two maps only (1x1 included)
  p.kind = a->kind; p.flags = a->flags; p.cout = a->cout;
  p.ksize = a->kind == EAMM_CONV_7X7 ? 7 : 3;
  p.taps = a->kind == EAMM_CONV_UP2_3X3 ? 4 : (row7 ? 7 : p.ksize * p.ksize);
  p.classes = a->kind == EAMM_CONV_UP2_3X3 ? 4 : 1;
  p.cin_chunks = row7 ? 1 : a->cin / 64;
  p.chunk_shift = ilog2_exact(p.cin_chunks);
  if (p.chunk_shift < 0) return EAMM_ERR_UNSUPPORTED;          // cin/64 must be a power of two
  // (mixed input: the two fp8 passes over 128-channel chunks count as one pass of 64-channel chunks)
  p.passes = row7 ? a->pack_passes : (mix ? 2 : (in->planes == 2 ? 3 : 1));
  p.f16in = f16in ? 1 : 0; p.mix = mix ? 1 : 0;
  if (f16in && row7 && a->pack_passes != 1) return EAMM_ERR_UNSUPPORTED;
  p.a_c_off = in->c_off; p.a_c_buf = in->c_buf;
  static int halo_env = -1;
  if (halo_env < 0) { const char* e = getenv("EAMM_TC_HALO"); halo_env = e ? atoi(e) : 1; }
  int mode7 = halo_env > 0 ? eamm_conv_tc_uses_halo(a->kind, in->w, a->cout,
                                                    (a->out || a->out2 || a->out_nhwc_f32) ? 0 : a->out_nchw_c) : 0;
  // 112-column kx-in-N variants (EAMM_TC_KXW bit 0: scheme 3 = four output rows per tile for the <=4-channel
  // NCHW layer; bit 1: scheme 4 = full-width tiles for a <=128-wide map with 16 fp32 NHWC couts)
  static int kxw_env = -1;
  if (kxw_env < 0) { const char* e = getenv("EAMM_TC_KXW"); kxw_env = e ? atoi(e) : 7; }
  if (mode7 == 2 && (kxw_env & 1) && in->h % 4 == 0) mode7 = 3;
  if (halo_env > 0 && (kxw_env & 2) && a->kind == EAMM_CONV_7X7 && a->cout == 16 && a->flags == 0 && a->out_nhwc_f32 &&
      !(a->out || a->out2 || a->out_nchw || a->residual) && in->w <= 128 && in->w * in->h >= 128)
    mode7 = 4;
  p.kxn = mode7 >= 2 ? mode7 - 1 : 0; p.halo = mode7 == 1;
  static int debug_env = -1;
  if (debug_env < 0) { const char* e = getenv("EAMM_TC_DEBUG"); debug_env = e ? atoi(e) : 0; }
  p.debug = debug_env;
  // 128-pixel box: bw x bh x bn
  if (p.kxn == 3) { p.bw = in->w; p.bh = 128 / in->w; p.bn = 1; }
  else if (p.halo || p.kxn) { p.bw = 128; p.bh = 1; p.bn = 1; }
  else {
    p.bw = in->w >= 16 ? 16 : in->w;
    p.bh = 128 / p.bw; if (p.bh > in->h) p.bh = in->h;
    p.bn = 128 / (p.bw * p.bh);
  }
  p.bw_log2 = ilog2_exact(p.bw); p.bh_log2 = ilog2_exact(p.bh);
  p.x_stride = (p.kxn == 1 || p.kxn == 2) ? 122 : p.bw;
  p.y_stride = p.kxn == 2 ? 4 : p.bh;
  p.ntap = p.kxn == 2 ? 10 : ((p.halo || p.kxn) ? 7 : p.taps);
  p.tiles_x = (in->w + p.x_stride - 1) / p.x_stride;
  p.tiles_y = (in->h + p.y_stride - 1) / p.y_stride; p.tiles_n = (in->n + p.bn - 1) / p.bn;
  static int num_sms = 0;
  if (!num_sms) {
    const char* e = getenv("EAMM_TC_NUM_SMS");       // planning dry runs (eamm_conv_tc_query) on a host without a GPU
    if (e && atoi(e) > 0) num_sms = atoi(e);
    else {
      int dev = 0; cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
  }
  static int prof_env = -1;
  if (prof_env < 0) { const char* e = getenv("EAMM_TC_PROF"); prof_env = e ? atoi(e) : 0; }
  const bool instr = prof_env == 1 || p.debug;       // the instrumented instantiation is single-CTA, plain scheme only
                                                     // (EAMM_TC_PROF=2: role counters of the production kernels, halo-tile scheme)
  // mixed-format outputs are written 32 channels at a time (one 32-byte store per e4m3 plane): N tiles of 32+ columns
  bool need32 = false;
  for (int i = 0; i < 2; ++i)
    if (views[i] && views[i]->dtype == EAMM_F16 && views[i]->planes == 2) need32 = true;
  if (need32 && a->cout % 32) return EAMM_ERR_UNSUPPORTED;
  // Halo-tile scheme (ConvTcParams::ah): 3x3 / UP2 layers with fp16 or mixed operands whose 8 x 16-pixel tiles fill the chip.
  // EAMM_TC_AH=0 switches it off (A/B runs).
  static int ah_env = -1;
  if (ah_env < 0) { const char* e = getenv("EAMM_TC_AH"); ah_env = e ? atoi(e) : 1; }
  p.ah = 0; p.ah_g = 1; p.cls_groups = p.classes; p.ah_na = 0; p.ah_spc = 0; p.ah_nsec = 0;
  int ah_bn = 0;
  if (ah_env && f16in && !row7 && !instr && (a->kind == EAMM_CONV_3X3 || a->kind == EAMM_CONV_UP2_3X3) &&
      in->w % 8 == 0 && in->h % 16 == 0) {
    // N tile: the widest of {cout (<= 256) or 256, 128, 64} whose work items fill the chip; failing that the narrowest one if
    // it still occupies 60 % of the SMs (batch-1 calls: a 64x64 bottleneck conv is 32 tiles x 4 N tiles = 128 items, and the
    // per-tap scheme kept each of those SMs on its own L2 -> SM limit: 1.7 MB per item against 0.76 MB with a halo tile)
    const int first = a->cout <= 256 ? a->cout : (a->cout % 256 == 0 ? 256 : 0);
    const int cand[3] = {first, 128, 64};
    int last_ok = 0, last_g = 1;
    for (int i = 0; i < 3 && !p.ah; ++i) {
      const int bn = cand[i];
      if (bn <= 0 || bn > first || (i > 0 && bn == cand[i - 1]) || a->cout % bn || bn % 16 || (need32 && bn % 32) ||
          (ah_env == 2 && bn == 256))
        continue;
      const int g = a->kind == EAMM_CONV_UP2_3X3 ? (256 / bn >= 4 ? 4 : (256 / bn >= 2 ? 2 : 1)) : 1;
      const long long items = (long long)(in->w / 8) * (in->h / 16) * in->n * (p.classes / g) * (a->cout / bn);
      last_ok = bn; last_g = g;
      if (items >= num_sms) { p.ah = 1; p.ah_g = g; p.cls_groups = p.classes / g; ah_bn = bn; }
    }
    if (!p.ah && last_ok) {
      const long long items = (long long)(in->w / 8) * (in->h / 16) * in->n * (p.classes / last_g) * (a->cout / last_ok);
      if (items * 10 >= (long long)num_sms * 6) { p.ah = 1; p.ah_g = last_g; p.cls_groups = p.classes / last_g; ah_bn = last_ok; }
    }
  }
  if (p.ah) {
    p.bw = 8; p.bh = 16; p.bn = 1; p.bw_log2 = 3; p.bh_log2 = 4; p.x_stride = 8; p.y_stride = 16;
    p.tiles_x = in->w / 8; p.tiles_y = in->h / 16; p.tiles_n = in->n;
  }
  // N tile: the widest UMMA N (<= 256) dividing cout that still yields at least one tile per SM;
  // small maps (hourglass 8x8 ... 2x2) prefer narrow N tiles so that more SMs stream the weights.
  static int fold_env = -1;
  if (fold_env < 0) { const char* e = getenv("EAMM_TC_FOLD"); fold_env = e ? atoi(e) : 1; }
  p.fold = 0;
  // decided from the layer's cout, not from the batch-dependent N tile, so that a frame's result does
  // not depend on how many other frames share the launch
  if (fold_env && !p.halo && (p.kxn || a->cout <= 128)) {
    if (row7 && a->pack_passes == 2) p.fold = 2;
    else if (!row7 && p.passes == 3) p.fold = 1;
  }
  if (!query && a->weight_fold != p.fold) return EAMM_ERR_ARG; // the caller packed the weights for the other scheme
  if (p.kxn) p.BN = p.kxn == 1 ? 32 : KXW_COLS;
  else if (p.ah) p.BN = ah_bn;
  else {
    const long long m_tiles = (long long)p.tiles_x * p.tiles_y * p.tiles_n * p.classes;
    p.BN = 0;
    const int cand[5] = {256, 128, 64, 32, need32 ? 32 : 16};
    if (a->cout <= 256 && m_tiles >= num_sms) p.BN = a->cout;
    // Layers that cannot give every SM a full-width tile: the N tile with the lowest modelled time (waves x MMA steps
    // x per-step cycles, measured: ~110 up to N = 64, 124 at 128, 192 at 256, plus the epilogue) -- one wave of
    // N = 64 tiles beats two waves of N = 32 tiles even though it leaves SMs idle.  EAMM_TC_BNCOST=0: the older rule
    // (widest tile that still yields one tile per SM).  The choice never changes a result bit (same per-column order).
    static int bncost_env = -1;
    if (bncost_env < 0) { const char* e = getenv("EAMM_TC_BNCOST"); bncost_env = e ? atoi(e) : 1; }
    if (!p.BN && bncost_env) {
      auto step = [](int n) -> long long {
        return n >= 256 ? 192 : (n >= 128 ? 124 + (n - 128) * 68 / 128 : (n >= 64 ? 114 + (n - 64) * 10 / 64 : 110));
      };
      const long long pairs = (long long)p.taps * p.cin_chunks;          // (tap, 64-channel chunk) pairs per tile
      long long best = -1;
      const int cands[6] = {a->cout <= 256 ? a->cout : 0, 256, 128, 64, 32, 16};
      for (int i = 0; i < 6; ++i) {
        const int c = cands[i];
        if (c <= 0 || c > a->cout || a->cout % c != 0 || (p.fold && 2 * c > 256) || (need32 && c % 32)) continue;
        const long long waves = (m_tiles * (a->cout / c) + num_sms - 1) / num_sms;
        const long long per_pair = p.fold == 1 ? step(2 * c) + step(c) : (p.fold == 2 ? step(2 * c) : (long long)p.passes * step(c));
        const long long cost = waves * (4 * pairs * per_pair + 2000 + 14000ll * c / 256);
        if (best < 0 || cost < best) { best = cost; p.BN = c; }
      }
    }
    for (int i = 0; i < 5 && !p.BN; ++i)
      if (cand[i] <= a->cout && a->cout % cand[i] == 0 && m_tiles * (a->cout / cand[i]) >= num_sms) p.BN = cand[i];
    if (!p.BN) {                       // cannot fill the chip: narrowest tile of at least 64 columns
      if (a->cout % 64 == 0) p.BN = 64;
      else if (a->cout <= 256) p.BN = a->cout;
      else return EAMM_ERR_UNSUPPORTED;
    }
  }
  // split-K (see ConvTcParams::splitk): unfolded layers whose N = 256 tiling leaves more than half of the SMs idle.
  // S = the largest divisor of the K-chunk count that keeps tiles * S within the SM count (>= 8 chunks per item).
  // The partial sums are added in split order, so a result is reproducible for a given batch size, but its last
  // bits depend on S and hence on how many frames share the launch.
  p.splitk = 1; p.sk_ws = nullptr; p.sk_cnt = nullptr;
  static int splitk_env = -1;
  if (splitk_env < 0) { const char* e = getenv("EAMM_TC_SPLITK"); splitk_env = e ? atoi(e) : 1; }
  if (splitk_env && !p.kxn && !p.halo && !row7 && !p.fold && !p.ah && a->cout % 256 == 0 && (query || a->splitk_ws)) {
    const long long t256 = (long long)p.tiles_x * p.tiles_y * p.tiles_n * p.classes * (a->cout / 256);
    const int kc_all = p.taps * p.cin_chunks * p.passes;
    static int splitk_max = -1, sk_narrow_env = -1;
    if (splitk_max < 0) { const char* e = getenv("EAMM_TC_SPLITK_MAX"); splitk_max = e ? atoi(e) : 9; }
    if (sk_narrow_env < 0) { const char* e = getenv("EAMM_TC_SPLITK_NARROW"); sk_narrow_env = e ? atoi(e) : 1; }
    // N tile of the split layer: 256, or -- when 256-column tiles x splits still leave more than half of the SMs without
    // work (batch-1 calls: one M tile, a handful of N tiles) -- 128 / 64 columns, so that the weight matrix, which is what
    // these layers stream, is pulled through the L2 -> SM ports of (nearly) every SM instead of 36 of them
    int sk_bn = 256, S = 1;
    long long best_ctas = 0;
    for (int bn = 256; bn >= 64; bn >>= 1) {
      const long long t = t256 * (256 / bn);
      int s_ = 1;
      for (int d = 2; d <= splitk_max && t * d <= num_sms; ++d)
        if (kc_all % d == 0 && kc_all / d >= 8) s_ = d;
      if (t * s_ > best_ctas) { best_ctas = t * s_; sk_bn = bn; S = s_; }
      if (t * s_ * 2 > num_sms || !sk_narrow_env) break;
    }
    const long long tsk = t256 * (256 / sk_bn);
    const long long need = 4096 + tsk * S * 128ll * sk_bn * 4;
    // Worth it?  Measured cycle model (EAMM_TC_PROF role counters, B200): one K=16 MMA step costs ~110 cycles up to
    // N = 64, 124 at N = 128, 192 at N = 256 (the 128-row A slab is re-fetched per step whatever N is); publishing
    // a partial tile and electing costs ~20k cycles, the reducer's reload + epilogue ~10k + 3k per split (the
    // reducer walks the S partials one L2 round trip after the other, hence the cap: 27 splits of a batch-1 layer
    // were measured slower than no split); a plain 128 x BN epilogue ~2k + 14k * BN/256.
    // Split only when the model says >= 15 % faster: batch-1 bottleneck convs (one wave of N = 64 tiles) stay unsplit.
    const long long tiles_plain = (long long)p.tiles_x * p.tiles_y * p.tiles_n * p.classes * (a->cout / p.BN);
    const long long waves = (tiles_plain + num_sms - 1) / num_sms;
    const long long step_plain = p.BN >= 256 ? 192 : (p.BN >= 128 ? 124 : (p.BN >= 64 ? 114 : 110));
    const long long step_split = sk_bn >= 256 ? 192 : (sk_bn >= 128 ? 124 : 114);
    const long long cost_plain = waves * (4ll * kc_all * step_plain + 2000 + 14000ll * p.BN / 256);
    const long long cost_split = 4ll * (kc_all / S) * step_split + (30000ll * sk_bn) / 256 + 3000ll * S;
    if (t256 * 2 <= num_sms && tsk <= 1024 && S > 1 && cost_split * 100 < cost_plain * 85 &&
        (query || a->splitk_ws_bytes >= need) && (query || (uintptr_t)a->splitk_ws % 16 == 0)) {
      p.BN = sk_bn; p.splitk = S;
      if (!query) {
        p.sk_cnt = reinterpret_cast<unsigned int*>(a->splitk_ws);
        p.sk_ws = reinterpret_cast<float*>(static_cast<char*>(a->splitk_ws) + 4096);
      }
    }
  }
  p.n_tiles = p.kxn ? 1 : a->cout / p.BN;
  static int cta2_env = -1;
  if (cta2_env < 0) { const char* e = getenv("EAMM_TC_CTA2"); cta2_env = e ? atoi(e) : 19; }   // bit 0: pairs, bit 1: folded pairs, bit 2: unfolded pairs with N < 256 (measured slower in
                                                                                             // single-plane mode: the leader's one MMA warp issues for both CTAs; off by default)
  {
    const long long tiles_all = (long long)p.tiles_x * p.tiles_y * p.tiles_n * p.cls_groups * p.n_tiles;
    const bool common = cta2_env && !instr && !p.halo && !p.kxn && !row7 && p.splitk == 1 && (num_sms % 2) == 0;
    // (a) unfolded layers (3-pass cout > 128, every single-plane layer), (b) folded layers (split mode, cout <= 128); both
    // once they fill the chip and when a pair's two M tiles exist (even count).  The arithmetic (per-column
    // accumulation order) is the same as the single-CTA kernel's, so the choice may depend on the batch.
    const long long m_tiles_pc = (long long)p.tiles_x * p.tiles_y * p.tiles_n;     // M tiles per (class, N tile)
    const bool pair_a = !p.fold && (p.BN == 256 || ((cta2_env & 4) && p.BN % 32 == 0));
    const bool pair_b = ((cta2_env & 2) && p.fold == 1 && p.BN % 32 == 0) ||
                        ((cta2_env & 16) && p.mix && p.BN % 32 == 0) ||     // mixed operands: 8 steps per (tap, 64 ch) like a folded layer
                        ((cta2_env & 16) && p.ah && p.BN % 32 == 0);        // halo tiles, one-plane fp16: the weight stream is what is left of
                                                                            // the L2 traffic, a pair halves it per SM (down0 fp16 mode: L2-bound alone)
    const bool fills = m_tiles_pc % 2 == 0 && tiles_all >= num_sms;
    p.cta2 = (common && fills && (pair_a || pair_b)) ? 1 : 0;
  }
  p.b_rows_total = p.kxn ? p.BN : p.classes * a->cout;
  p.a_slot_bytes = p.halo ? 17 * 1024 : TC_A_BYTES;
  const uint32_t chunk_bytes = (uint32_t)p.a_slot_bytes +
      (p.cta2 ? (uint32_t)p.BN * (p.fold ? 128u : 64u) : (uint32_t)p.BN * 128u * (p.halo ? 7u : (p.fold ? 2u : 1u)));
  p.s_nc = (p.kxn == 2 && (kxw_env & 4) && a->out_nchw_c < 4) ? a->out_nchw_c : 4;
  const uint32_t extra_smem = p.kxn == 1 ? 2u * 128u * 29u * 4u
                            : p.kxn == 2 ? 128u * (28u * (uint32_t)p.s_nc + 1u) * 4u
                            : (p.kxn ? 128u * (uint32_t)KXW_LD * 4u : 0u);
  // (scheme 3 takes everything the SM has: with <= 3 NCHW channels that is a fourth 44 KB stage)
  // TMA-store epilogue (ConvTcParams::tma_st): plain (no residual / second output / pooling) activation-view outputs of layers
  // whose tiles lie inside one image and whose 32-pixel warp groups are whole rows / row segments; dense NHWC views.  By default
  // only where the epilogue is the critical path, N tiles of <= 64 columns (packed first conv, up1: measured -5 % / -8 % per tile);
  // wider layers are MMA- or L2-bound and need the shared memory for their weight ring.  EAMM_TC_TMAST=0: off, 2: every plain layer.
  static int tmast_env = -1;
  if (tmast_env < 0) { const char* e = getenv("EAMM_TC_TMAST"); tmast_env = e ? atoi(e) : 1; }
  auto tma_view_ok = [&](const eamm_act* v, int up) {
    return v && v->dtype != EAMM_F32 && v->n_stride == (int64_t)v->h * v->w * v->planes * v->c_buf && (uintptr_t)v->data % 128 == 0 &&
           (v->planes * v->c_buf) % 64 == 0 && v->c_off % 32 == 0 && v->c_buf % 32 == 0 && v->h == in->h * up && v->w == in->w * up &&
           v->n == in->n;
  };
  const int st_up = a->kind == EAMM_CONV_UP2_3X3 ? 2 : 1;
  p.tma_st = (tmast_env && !instr && !p.kxn && p.splitk == 1 && p.BN % 32 == 0 && !(a->flags & (EAMM_EPI_POOL2 | EAMM_EPI_SIGMOID)) &&
              !a->out_nhwc_f32 && !a->out_nchw && p.bn == 1 && p.bw >= 8 && in->h % p.bh == 0 && in->w % p.bw == 0 &&
              tma_view_ok(a->out, st_up) && !a->out2 && !a->residual && (p.BN <= 64 || tmast_env == 2))
                 ? 1 : 0;
  const uint32_t st_bytes = p.tma_st ? 32u * 1024u : 0u;      // 8 warps x 4 KB
  const uint32_t ring_bytes = (p.tma_st ? 226u * 1024u - st_bytes : (p.kxn == 2 ? 225u * 1024u : 200u * 1024u)) - extra_smem;
  // K chunks per stage: as many as keep >= 4 stages in the ring (>= 3 for the widest tiles); short
  // single-warp issue loops are latency-bound, so fewer, fatter stages win until smem runs out.
  static int ksub_env = -1;
  if (ksub_env < 0) { const char* e = getenv("EAMM_TC_KSUB"); ksub_env = e ? atoi(e) : 0; }
  const int kc_total = p.ntap * p.cin_chunks * (p.fold == 1 ? 2 : (p.fold == 2 ? 1 : p.passes)) / p.splitk;
  int ksub = 1;
  for (int k = 4; k >= 2; --k)
    if ((uint32_t)k * chunk_bytes * 4u <= ring_bytes && k <= kc_total) { ksub = k; break; }
  if (ksub == 1 && 2u * chunk_bytes * 3u <= ring_bytes && kc_total >= 2 && p.BN < 256) ksub = 2;
  static int cta2_ksub_env = -1;
  if (cta2_ksub_env < 0) { const char* e = getenv("EAMM_TC_CTA2_KSUB"); cta2_ksub_env = e ? atoi(e) : 1; }
  // pairs: N = 256 stages are long enough with one chunk; narrower ones (and folded pairs, whose N = BN step
  // alone is shorter than the per-stage overhead) take two
  if (p.cta2) ksub = p.BN == 256 && !p.fold ? (cta2_ksub_env > 0 ? cta2_ksub_env : 1) : (kc_total >= 2 ? 2 : 1);
  if (ksub_env > 0) ksub = ksub_env;
  while (ksub > 1 && (uint32_t)ksub * chunk_bytes * 2u > ring_bytes) --ksub;      // keep at least two stages
  uint32_t stage_bytes = (uint32_t)ksub * chunk_bytes;
  int stages = (int)(ring_bytes / stage_bytes);
  // resident weights (ConvTcParams::b_res): packed first conv in split mode -- one N tile, 7 chunks of [hi | lo] rows
  static int lean_env = -1, bres_env = -1;
  if (lean_env < 0) { const char* e = getenv("EAMM_TC_LEAN"); lean_env = e ? atoi(e) : 1; }
  if (bres_env < 0) { const char* e = getenv("EAMM_TC_BRES"); bres_env = e ? atoi(e) : 1; }
  p.b_res = 0;
  size_t bres_bytes = 0;
  if (bres_env && lean_env && !instr && row7 && (p.fold == 2 || p.passes == 1) && !p.cta2 && p.n_tiles == 1 && p.splitk == 1) {
    const uint32_t b_bytes = (uint32_t)p.BN * 128u * (p.fold ? 2u : 1u);
    if ((size_t)kc_total * b_bytes + 4u * (uint32_t)p.a_slot_bytes <= ring_bytes) {
      p.b_res = 1; bres_bytes = (size_t)kc_total * b_bytes;
      ksub = ksub_env > 0 ? ksub_env : 1;
      stage_bytes = (uint32_t)ksub * (uint32_t)p.a_slot_bytes;
      stages = (int)((ring_bytes - bres_bytes) / stage_bytes);
    }
  }
  if (p.ah) {
    // weight stage = `ksub` taps of one parity class (a whole filter row / a whole 2x2 class when the tile is narrow)
    const uint32_t b_bytes = (uint32_t)p.BN * (p.cta2 ? 64u : 128u);
    if (a->kind == EAMM_CONV_3X3) ksub = 3u * b_bytes <= 24u * 1024u ? 3 : 1;
    else ksub = 4u * b_bytes <= 32u * 1024u ? 4 : (2u * b_bytes <= 32u * 1024u ? 2 : 1);
    p.a_slot_bytes = 23 * 1024;                       // 180 pixel rows of 128 bytes, slots 1024-byte aligned
    p.ah_na = 3;
    p.ah_spc = p.ah_g * p.ntap / ksub;
    stage_bytes = (uint32_t)ksub * b_bytes;
    stages = (int)((ring_bytes - (uint32_t)p.ah_na * (uint32_t)p.a_slot_bytes) / stage_bytes);
    // K sections (byte columns of the uint8 tensor maps; see the plain producer for the operand layout)
    const int nc8 = a->cin / 128, nc16 = a->cin / 64;
    if (mix64) {
      p.ah_nsec = 2;
      p.ah_cbase[0] = 2 * in->c_buf; p.ah_nch[0] = 1; p.ah_f8[0] = 1; p.ah_bcol[0] = 0;
      p.ah_cbase[1] = 0; p.ah_nch[1] = 1; p.ah_f8[1] = 0; p.ah_bcol[1] = p.ntap * 128;
    } else if (mix) {
      p.ah_nsec = 3;
      p.ah_cbase[0] = 3 * in->c_buf + in->c_off; p.ah_nch[0] = nc8; p.ah_f8[0] = 1; p.ah_bcol[0] = 0;
      p.ah_cbase[1] = 2 * in->c_buf + in->c_off; p.ah_nch[1] = nc8; p.ah_f8[1] = 1; p.ah_bcol[1] = p.ntap * nc8 * 128;
      p.ah_cbase[2] = 2 * in->c_off; p.ah_nch[2] = nc16; p.ah_f8[2] = 0; p.ah_bcol[2] = 2 * p.ntap * nc8 * 128;
    } else {
      p.ah_nsec = 1;
      p.ah_cbase[0] = 2 * in->c_off; p.ah_nch[0] = nc16; p.ah_f8[0] = 0; p.ah_bcol[0] = 0;
    }
  }
  p.ksub = ksub;
  if (stages > 8) stages = 8;
  if (stages < 2) return EAMM_ERR_UNSUPPORTED;
  p.num_stages = stages;
  if (query) {
    query[0] = p.BN; query[1] = mode7; query[2] = p.fold; query[3] = p.ksub;
    query[4] = p.cta2 | (p.ah << 1) | (p.splitk << 8); query[5] = p.num_stages;
    return 0;
  }
  p.has_out = a->out != nullptr; p.has_out2 = a->out2 != nullptr; p.has_res = a->residual != nullptr;
  ActView dummy = make_view(in);
  p.out = p.has_out ? make_view(a->out) : dummy;
  p.out2 = p.has_out2 ? make_view(a->out2) : dummy;
  p.res = p.has_res ? make_view(a->residual) : dummy;
  static int st256_env = -1;
  if (st256_env < 0) { const char* e = getenv("EAMM_TC_ST256"); st256_env = e ? atoi(e) : 1; }
  p.st256 = st256_env ? 1 : 0;
  for (int i = 0; i < 3; ++i)
    if (views[i] && ((uintptr_t)views[i]->data % 32 || views[i]->c_off % 16 || views[i]->c_buf % 16 ||
                     views[i]->n_stride % 16 || (views[i]->planes * views[i]->c_buf) % 16))
      p.st256 = 0;
  // mixed-format outputs: one 32-byte store per e4m3 plane and 32-channel chunk
  for (int i = 0; i < 2; ++i)
    if (views[i] && views[i]->dtype == EAMM_F16 && views[i]->planes == 2 &&
        (!p.st256 || p.BN % 32 || views[i]->c_off % 32 || views[i]->c_buf % 32 || p.kxn))
      return EAMM_ERR_UNSUPPORTED;
  {
    // fast epilogue: activation-view outputs only, 32-column chunks, 32-byte accesses, no split-K / kx-in-N
    static int fast_env = -1;
    if (fast_env < 0) { const char* e = getenv("EAMM_TC_EPIFAST"); fast_env = e ? atoi(e) : 1; }
    const bool pool = (a->flags & EAMM_EPI_POOL2) != 0;
    bool ok = fast_env && !instr && !p.kxn && p.splitk == 1 && p.BN % 32 == 0 && p.st256 && a->out && !a->out_nhwc_f32 && !a->out_nchw &&
              !(a->flags & EAMM_EPI_SIGMOID);
    if (ok && pool && (a->residual || a->out2 || a->out->c_off % 8 || a->out->c_buf % 8)) ok = false;
    if (ok && a->out2 && !a->residual) ok = false;
    if (ok && p.fold && (a->residual || a->out2)) ok = false;
    p.epi_fast = !ok ? 0 : (p.fold ? (pool ? 6 : 5) : (pool ? 2 : (a->residual ? (a->out2 ? 4 : 3) : 1)));
    if (!p.epi_fast || pool) p.tma_st = 0;
  }
  p.acc_scale = a->acc_scale; p.amax_out = a->out ? a->amax_out : nullptr; p.amax_out2 = a->out2 ? a->amax_out2 : nullptr;
  p.lean = (lean_env && !p.halo) ? 1 : 0;
  p.mix64 = mix64 ? 1 : 0;
  p.nf8 = mix64 ? p.ntap : (mix ? p.ntap * p.cin_chunks : 0);
  p.bias = a->bias; p.scale2 = a->scale2; p.shift2 = a->shift2;
  p.out_nchw = a->out_nchw; p.out_nchw_c = a->out_nchw_c; p.out_nhwc = a->out_nhwc_f32; p.out_u8 = a->out_u8_nhwc;
  p.total_tiles = (long long)p.tiles_x * p.tiles_y * p.tiles_n * p.cls_groups * p.n_tiles * p.splitk;
  {
    const int l0 = ilog2_exact(p.n_tiles), l1 = ilog2_exact(p.cls_groups), l2 = ilog2_exact(p.tiles_x), l3 = ilog2_exact(p.tiles_y);
    p.dec_shift = (l0 >= 0 && l1 >= 0 && l2 >= 0 && l3 >= 0) ? (l0 | (l1 << 5) | (l2 << 10) | (l3 << 15)) : -1;
  }
  if (p.total_tiles > 0x7fffffffLL) return EAMM_ERR_UNSUPPORTED;

  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return EAMM_ERR_UNSUPPORTED;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4] = {64, (cuuint32_t)(p.halo ? 134 : p.bw), (cuuint32_t)p.bh, (cuuint32_t)p.bn};
    if (row7) {
      // Overlapping windows: the buffer is the source image packed as [n][H+6][W+8][8 ch] bf16 with a
      // zero border (3 rows top/bottom, 3 columns left, 5 right).  Row x of the map is the 64-element
      // window of 8 consecutive pixels starting at padded column x, i.e. dimension 1 has a 16-byte
      // stride although dimension 0 spans 128 bytes.
      const cuuint64_t wp = (cuuint64_t)in->w + 8, hp = (cuuint64_t)in->h + 6;
      dims[0] = 64; dims[1] = (cuuint64_t)in->w; dims[2] = hp; dims[3] = (cuuint64_t)in->n;
      strides[0] = 16; strides[1] = wp * 16; strides[2] = wp * hp * 16;
    } else {
      const cuuint64_t pix = (cuuint64_t)in->planes * in->c_buf;
      if (in->n_stride != (int64_t)in->h * in->w * (int64_t)pix) return EAMM_ERR_UNSUPPORTED;
      dims[0] = pix; dims[1] = (cuuint64_t)in->w; dims[2] = (cuuint64_t)in->h; dims[3] = (cuuint64_t)in->n;
      strides[0] = pix * 2; strides[1] = pix * 2 * in->w; strides[2] = pix * 2 * in->w * in->h;
    }
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (f16in) { dims[0] *= 2; box[0] = 128; }            // byte units: fp16 and e4m3 planes are addressed through one uint8 map
    if (p.ah) { box[1] = AH_PITCH; box[2] = AH_ROWS; box[3] = 1; }      // the 10 x 18-pixel halo of an 8 x 16 tile
    CUresult r = encode(&tmA, f16in ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in->data, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return EAMM_ERR_UNSUPPORTED - 100 - (int)r;      // distinguishable in bring-up logs
  }
  {
    // halo mode: rows = (kx, cout), K = (ky, pass, channel); otherwise rows = (class, cout), K = (tap, pass, channel)
    // kxn mode : rows = 32 (kx*4 + cout), K = (ky, pass, channel); kxn 2: rows = 112 (dr*28 + kx*4 + cout), K = (input
    //            row j of 10, pass, channel); kxn 3: rows = 112 (kx*16 + cout), K = (ky, pass, channel)
    // fold mode: K = (tap, channel) only, rows = [hi block | lo block]
    const cuuint64_t ktot = (cuuint64_t)p.ntap * (p.fold ? 1 : p.passes) * (row7 ? 64 : a->cin);
    const cuuint64_t rows = (p.kxn ? (cuuint64_t)p.BN : (cuuint64_t)(p.halo ? 7 : p.classes) * a->cout) * (p.fold ? 2 : 1);
    cuuint64_t dims[2] = {ktot, rows};
    cuuint64_t strides[1] = {ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(p.halo ? 7 * p.BN : (p.cta2 ? p.BN / 2 : p.BN))};
    cuuint32_t es[2] = {1, 1};
    if (f16in) {                                          // bytes; mixed: (tap, channel) x [lo8 | hi8 | fp16 hi] = 4 bytes each
      dims[0] = mix ? (cuuint64_t)p.ntap * a->cin * 4 : ktot * 2;
      strides[0] = dims[0]; box[0] = 128;
    }
    CUresult r = encode(&tmB, f16in ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a->weight), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return EAMM_ERR_UNSUPPORTED - 200 - (int)r;
  }
  size_t smem = (size_t)stages * stage_bytes + 1024 + extra_smem + (p.ah ? (size_t)p.ah_na * p.a_slot_bytes : 0) + bres_bytes;
  StoreMaps tmS;
  memset(&tmS, 0, sizeof(tmS));
  p.st_base = 0; p.st_stride = 0;
  if (p.tma_st) {
    p.st_base = (int)((smem - 1024 + 1023) / 1024 * 1024);      // staging areas: 1024-byte aligned, behind everything else
    smem = (size_t)p.st_base + st_bytes + 1024;
    p.st_stride = (int)(st_bytes / TC_EPI_WARPS);
    const int bx = p.bw < 32 ? p.bw : 32, by = 32 / bx;
    auto enc = [&](CUtensorMap* m, const eamm_act* v, int box_bytes) -> bool {
      // {pixel bytes, column parity, x, row parity, image * H_in + y}: output pixel (up*y + py, up*x + px)
      const cuuint64_t pixb = (cuuint64_t)v->planes * v->c_buf * 2, wout = (cuuint64_t)v->w;
      cuuint64_t dims[5] = {pixb, (cuuint64_t)st_up, (cuuint64_t)in->w, (cuuint64_t)st_up, (cuuint64_t)in->n * in->h};
      cuuint64_t strides[4] = {pixb, (cuuint64_t)st_up * pixb, wout * pixb, (cuuint64_t)st_up * wout * pixb};
      cuuint32_t box[5] = {(cuuint32_t)box_bytes, 1, (cuuint32_t)bx, 1, (cuuint32_t)by};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      return encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, v->data, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    box_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    bool okm = enc(&tmS.o64, a->out, 64) && enc(&tmS.o32, a->out, 32);
    if (okm && a->out2) okm = enc(&tmS.p64, a->out2, 64) && enc(&tmS.p32, a->out2, 32);
    if (!okm) return EAMM_ERR_UNSUPPORTED - 300;
  }
  // the attribute is per device: a process that drives several GPUs (nn.DataParallel replicas, train.py:53-60) must set it on each
  static size_t smem_set_dev[64] = {0};
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  size_t& smem_set = smem_set_dev[cur_dev & 63];
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    smem_set = smem;
  }
  long long grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  static unsigned long long* prof_buf = nullptr;
  p.prof = nullptr;
  if (prof_env) {
    if (!prof_buf) cudaMalloc(&prof_buf, 1024 * 8 * sizeof(unsigned long long));
    cudaMemsetAsync(prof_buf, 0, 1024 * 8 * sizeof(unsigned long long), (cudaStream_t)stream);
    p.prof = prof_buf;
  }
  if (p.cta2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, true>, tmA, tmB, tmS, p);
    if (e != cudaSuccess) return (int)e;
  } else if (instr) conv_tc_kernel<true, false><<<(unsigned)grid, TC_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, tmS, p);
  else conv_tc_kernel<false, false><<<(unsigned)grid, TC_THREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, tmS, p);
  EAMM_LAUNCH_CHECK();
  if (prof_env) {        // bring-up instrumentation only: synchronous read-back and print
    static unsigned long long host[1024 * 8];
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaMemcpy(host, prof_buf, grid * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long b = 0; b < grid; ++b) for (int j = 0; j < 8; ++j) acc[j] += (double)host[b * 8 + j];
    const double tiles_per_cta = (double)p.total_tiles / (double)grid;
    const int KCh = kc_total;
    const double mg = p.cta2 ? grid / 2.0 : (double)grid;          // the MMA counters exist in the leader CTA of a pair only
    fprintf(stderr, "[tc_prof] kind=%d %dx%dx%d cin=%d cout=%d BN=%d stages=%dx%d KC=%d ah=%d(g%d) cta2=%d epi=%d tiles/cta=%.1f | per tile (cycles): "
            "total=%.0f prod_wait_empty=%.0f mma_wait_acc=%.0f mma_wait_full=%.0f epi_wait_full=%.0f epi_busy=%.0f | per stage=%.0f\n",
            p.kind, p.N, p.H, p.W, a->cin, a->cout, p.BN, p.num_stages, p.ksub, KCh, p.ah, p.ah_g, p.cta2, p.epi_fast, tiles_per_cta,
            acc[4] / mg / tiles_per_cta, acc[0] / grid / tiles_per_cta, acc[2] / mg / tiles_per_cta,
            acc[3] / mg / tiles_per_cta, acc[5] / grid / tiles_per_cta, (acc[6] - acc[5]) / grid / tiles_per_cta,
            acc[4] / mg / tiles_per_cta / KCh);
#ifdef EAMM_EPI_PHASES
    fprintf(stderr, "[epi_phases] per chunk of warp 0 (cycles): tmem wait %.0f, bias/scale + relu %.0f, first store %.0f (chunks per tile %d, tma_st %d)\n",
            acc[0] / grid / tiles_per_cta / (p.BN / 64.0 * p.ah_g), acc[1] / grid / tiles_per_cta / (p.BN / 64.0 * p.ah_g),
            acc[7] / grid / tiles_per_cta / (p.BN / 64.0 * p.ah_g), p.BN / 64 * p.ah_g, p.tma_st);
#endif
  }
  return 0;
}

