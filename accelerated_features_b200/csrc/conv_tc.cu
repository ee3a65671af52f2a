// 64->64 convolutions (3x3 pad 1 / 1x1, stride 1) as implicit GEMM on the 5th-generation tensor cores:
// TMA-fed shared-memory operands, tcgen05.mma with fp32 accumulators in TMEM, bias+ReLU epilogue from TMEM.
// Used for block3.1/3.2, block4.1/4.2, block_fusion.0-2, heatmap_head.0-1, keypoint_head.0-2 (model.py:55-92) -- 48 % of
// the network's FLOPs.
//
// fp32-equivalent precision by operand splitting (single-pass TF32/BF16/FP16 moves keypoints and misses the
// 1e-3 descriptor tolerance; 3-term splits reproduce the fp32 result):  x = hi + lo in fp16, w*2^k = whi + wlo in fp16,
//      y = (hi.whi + hi.wlo + lo.whi) * 2^-k     (fp32 accumulation; the dropped lo.wlo term is ~2^-22 relative).
// Activations travel between tensor-core layers already split: NHWC with 128 halves per pixel, [hi(64) | lo(64)] -- the
// same 256 bytes per pixel as fp32, written by the producing epilogue.
//
// GEMM mapping: M = 128 output pixels (a TH x TW patch of one image), N = 64 output channels, K = taps x 64 x 3 terms.
//   * the A operand of tap (dy,dx) is the input patch shifted by (dy-1, dx-1): ONE 4-D TMA box {64 ch, TW, TH, 1} per
//     term, 128B-swizzled K-major, out-of-image coordinates zero-filled by the TMA unit (= the conv's zero padding);
//   * the folded, pre-split weights of all taps stay resident in shared memory for the life of the (persistent) CTA;
//   * warp 0 = TMA producer, warp 1 = single-thread MMA issuer, warps 2-5 = epilogue (one TMEM lane quarter each);
//     the accumulator is double buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda_fp16.h>

#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace xf {

constexpr int CT_THREADS = 192;
constexpr int CT_ABOX = 128 * 128;   // bytes: 128 pixels x 64 halves
constexpr int CT_WBOX = 64 * 128;    // bytes: 64 cout x 64 halves

template <int KS>
struct ConvTcCfg {
  static constexpr int TAPS = KS * KS;
  static constexpr int NS = (KS == 3) ? 2 : 4;                       // A stages (one tap = hi + lo box each)
  static constexpr size_t W_BYTES = (size_t)TAPS * 2 * CT_WBOX;
  static constexpr size_t A_BYTES = (size_t)NS * 2 * CT_ABOX;
  static constexpr size_t SMEM = 1024 + W_BYTES + A_BYTES + 512;
};

struct ConvTcParams {
  CUtensorMap amap;   // input activations, split fp16 (B,H,W,128)
  CUtensorMap wmap;   // weights [tap][term][cout][cin] fp16
  const float* bias;
  float inv_wscale;
  int B, H, W;
  int tw_log2;        // TW = 1 << tw_log2, TH = 128 / TW
  __half* out_split;  // (B,H,W,128) or null
  float* out_f32;     // (B,H,W,64) or null
  int relu;
};

template <int KS>
__global__ void __launch_bounds__(CT_THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams P) {
  using C = ConvTcCfg<KS>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sW = base;
  unsigned char* sA = base + C::W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + C::W_BYTES + C::A_BYTES);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;                 // [NS]
  uint64_t* a_empty = bars + 1 + C::NS;        // [NS]
  uint64_t* acc_full = bars + 1 + 2 * C::NS;   // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sBias = reinterpret_cast<float*>(tmem_slot + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int TW = 1 << P.tw_log2, TH = 128 >> P.tw_log2;
  const int tiles_x = (P.W + TW - 1) / TW, tiles_y = (P.H + TH - 1) / TH;
  const int tiles_img = tiles_x * tiles_y;
  const int n_tiles = tiles_img * P.B;

  if (threadIdx.x < 64) sBias[threadIdx.x] = __ldg(P.bias + threadIdx.x);
  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&P.amap);
    tc::tma_prefetch_desc(&P.wmap);
    tc::mbar_init(w_full, 1);
    for (int i = 0; i < C::NS; ++i) {
      tc::mbar_init(&a_full[i], 1);
      tc::mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 128);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      tc::mbar_expect_tx(w_full, (uint32_t)C::W_BYTES);
      for (int i = 0; i < C::TAPS * 2; ++i) tc::tma_load_2d(sW + i * CT_WBOX, &P.wmap, w_full, 0, i * 64);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / tiles_img, rem = tile - b * tiles_img;
        const int y0 = (rem / tiles_x) * TH, x0 = (rem % tiles_x) * TW;
        for (int tap = 0; tap < C::TAPS; ++tap, ++it) {
          const int s = it % C::NS;
          const uint32_t ph = (it / C::NS) & 1;
          const int dy = (KS == 3) ? tap / 3 - 1 : 0, dx = (KS == 3) ? tap % 3 - 1 : 0;
          tc::mbar_wait(&a_empty[s], ph ^ 1);
          tc::mbar_expect_tx(&a_full[s], 2 * CT_ABOX);
          unsigned char* dst = sA + (size_t)s * 2 * CT_ABOX;
          tc::tma_load_4d(dst, &P.amap, &a_full[s], 0, x0 + dx, y0 + dy, b);               // hi half of the channels
          tc::tma_load_4d(dst + CT_ABOX, &P.amap, &a_full[s], 64, x0 + dx, y0 + dy, b);    // lo half
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc = tc::make_idesc(/*F16*/ 0, 128, 64);
      tc::mbar_wait(w_full, 0);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
        const int a = tcount & 1;
        tc::mbar_wait(&acc_empty[a], ((tcount >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t d = tmem + a * 64;
        for (int tap = 0; tap < C::TAPS; ++tap, ++it) {
          const int s = it % C::NS;
          tc::mbar_wait(&a_full[s], (it / C::NS) & 1);
          tc::tc_fence_after();
          const uint32_t a_addr = tc::smem_u32(sA + (size_t)s * 2 * CT_ABOX);
          const uint32_t w_addr = tc::smem_u32(sW + (size_t)tap * 2 * CT_WBOX);
          const uint64_t ahi = tc::make_desc_sw128(a_addr, 1024), alo = tc::make_desc_sw128(a_addr + CT_ABOX, 1024);
          const uint64_t whi = tc::make_desc_sw128(w_addr, 1024), wlo = tc::make_desc_sw128(w_addr + CT_WBOX, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::umma_f16(d, ahi + 2 * k, whi + 2 * k, idesc, (tap | k) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::umma_f16(d, ahi + 2 * k, wlo + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::umma_f16(d, alo + 2 * k, whi + 2 * k, idesc, 1u);
          tc::umma_commit(&a_empty[s]);
        }
        tc::umma_commit(&acc_full[a]);
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue ----------------
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // pixel of the tile = TMEM lane
    const int ph_ = r >> P.tw_log2, pw_ = r & (TW - 1);
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const int a = tcount & 1;
      const int b = tile / tiles_img, rem = tile - b * tiles_img;
      const int y = (rem / tiles_x) * TH + ph_, x = (rem % tiles_x) * TW + pw_;
      tc::mbar_wait(&acc_full[a], (tcount >> 1) & 1);
      tc::tc_fence_after();
      uint32_t v0[32], v1[32];
      __syncwarp();
      tc::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + a * 64, v0);
      tc::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + a * 64 + 32, v1);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[a]);   // TMEM buffer is free again: the stores below overlap the next MMAs
      if (y < P.H && x < P.W) {
        const int64_t pix = ((int64_t)b * P.H + y) * P.W + x;
        float o[64];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float t0 = fmaf(__uint_as_float(v0[c]), P.inv_wscale, sBias[c]);
          float t1 = fmaf(__uint_as_float(v1[c]), P.inv_wscale, sBias[32 + c]);
          if (P.relu) { t0 = fmaxf(t0, 0.f); t1 = fmaxf(t1, 0.f); }
          o[c] = t0; o[32 + c] = t1;
        }
        if (P.out_f32) {
          float4* op = reinterpret_cast<float4*>(P.out_f32 + pix * 64);
#pragma unroll
          for (int c = 0; c < 16; ++c) op[c] = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
        }
        if (P.out_split) {
          uint4* hp = reinterpret_cast<uint4*>(P.out_split + pix * 128);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            __half2 h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float x0 = o[8 * c + 2 * j], x1 = o[8 * c + 2 * j + 1];
              h[j] = __floats2half2_rn(x0, x1);
              const float2 hf = __half22float2(h[j]);
              l[j] = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
            }
            hp[c] = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]),
                               *reinterpret_cast<uint32_t*>(&h[2]), *reinterpret_cast<uint32_t*>(&h[3]));
            hp[8 + c] = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]),
                                   *reinterpret_cast<uint32_t*>(&l[2]), *reinterpret_cast<uint32_t*>(&l[3]));
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 128);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// helpers: fp32 NHWC (64 ch) -> split fp16 NHWC (128 halves); unfold8 + split of the normalised gray image
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
    const float2 hf = __half22float2(h[j]);
    l[j] = __floats2half2_rn(x[2 * j] - hf.x, x[2 * j + 1] - hf.y);
  }
  hi = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]), *reinterpret_cast<uint32_t*>(&h[2]),
                  *reinterpret_cast<uint32_t*>(&h[3]));
  lo = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]), *reinterpret_cast<uint32_t*>(&l[2]),
                  *reinterpret_cast<uint32_t*>(&l[3]));
}

__global__ void __launch_bounds__(256) split_nhwc64_kernel(const float* __restrict__ in, __half* __restrict__ out,
                                                           int64_t npix) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (pixel, 8-channel group)
  if (gid >= npix * 8) return;
  const int64_t pix = gid >> 3;
  const int g = (int)(gid & 7);
  const float4 a = __ldg(reinterpret_cast<const float4*>(in + pix * 64) + 2 * g);
  const float4 b = __ldg(reinterpret_cast<const float4*>(in + pix * 64) + 2 * g + 1);
  const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint4 hi, lo;
  split8(x, hi, lo);
  reinterpret_cast<uint4*>(out + pix * 128)[g] = hi;
  reinterpret_cast<uint4*>(out + pix * 128)[8 + g] = lo;
}

// XFeatModel._unfold2d(x, 8) (model.py:113-120) fused with the split: channel 8i+j of cell (h,w) = xn[8h+i, 8w+j].
__global__ void __launch_bounds__(256) unfold8_split_kernel(const float* __restrict__ xn, __half* __restrict__ out, int Hc,
                                                            int Wc, int64_t ncell) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (cell, row i)
  if (gid >= ncell * 8) return;
  const int64_t cell = gid >> 3;
  const int i = (int)(gid & 7);
  const int64_t b = cell / ((int64_t)Hc * Wc);
  const int rem = (int)(cell - b * Hc * Wc);
  const int h = rem / Wc, w = rem - h * Wc;
  const float* p = xn + ((int64_t)b * Hc * 8 + h * 8 + i) * (Wc * 8) + w * 8;
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), c = __ldg(reinterpret_cast<const float4*>(p) + 1);
  const float x[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
  uint4 hi, lo;
  split8(x, hi, lo);
  reinterpret_cast<uint4*>(out + cell * 128)[i] = hi;
  reinterpret_cast<uint4*>(out + cell * 128)[8 + i] = lo;
}

int launch_split_nhwc64(const float* in, __half* out, int64_t npix, cudaStream_t st) {
  split_nhwc64_kernel<<<(unsigned)((npix * 8 + 255) / 256), 256, 0, st>>>(in, out, npix);
  XF_LAUNCH_CHECK();
  return XF_OK;
}
int launch_unfold8_split(const float* xn, __half* out, int B, int Hc, int Wc, cudaStream_t st) {
  const int64_t ncell = (int64_t)B * Hc * Wc;
  unfold8_split_kernel<<<(unsigned)((ncell * 8 + 255) / 256), 256, 0, st>>>(xn, out, Hc, Wc, ncell);
  XF_LAUNCH_CHECK();
  return XF_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// host: weight preparation at context creation, per-launch tensor maps
// ---------------------------------------------------------------------------------------------------------------------
bool conv_tc_eligible(int layer) {
  const LayerSpec& s = kLayers[layer];
  return layer < L_FM_0 && s.cin == 64 && s.cout == 64 && s.stride == 1;
}

int conv_tc_prepare(xfeat_ctx* ctx) {
  // split weights: W[tap][cin][cout] fp32 (BN folded) -> [tap][term][cout][cin] fp16 of W * 2^k
  size_t total = 0;
  for (int l = 0; l < L_COUNT; ++l) {
    ctx->tc_off[l] = (size_t)-1;
    if (conv_tc_eligible(l)) {
      ctx->tc_off[l] = total;
      total += (size_t)kLayers[l].ks * kLayers[l].ks * 2 * 64 * 64;
    }
  }
  std::vector<__half> h(total);
  for (int l = 0; l < L_COUNT; ++l) {
    if (ctx->tc_off[l] == (size_t)-1) continue;
    const int taps = kLayers[l].ks * kLayers[l].ks;
    const float* w = ctx->h_weights + ctx->table.w_off[l];
    float mx = 0.f;
    for (int i = 0; i < taps * 64 * 64; ++i) mx = fmaxf(mx, fabsf(w[i]));
    int e = 0;
    if (mx > 0.f) frexpf(mx, &e);
    const float s = (mx > 0.f) ? ldexpf(1.f, 13 - e) : 1.f;          // max|w| * s in [2^12, 2^13)
    ctx->tc_inv_wscale[l] = (mx > 0.f) ? ldexpf(1.f, e - 13) : 1.f;
    __half* dst = h.data() + ctx->tc_off[l];
    for (int t = 0; t < taps; ++t)
      for (int co = 0; co < 64; ++co)
        for (int ci = 0; ci < 64; ++ci) {
          const float v = w[((size_t)t * 64 + ci) * 64 + co] * s;
          const __half hi = __float2half_rn(v);
          const __half lo = __float2half_rn(v - __half2float(hi));
          dst[(((size_t)t * 2 + 0) * 64 + co) * 64 + ci] = hi;
          dst[(((size_t)t * 2 + 1) * 64 + co) * 64 + ci] = lo;
        }
  }
  XF_CUDA(cudaMalloc(&ctx->d_tcw, total * sizeof(__half)));
  XF_CUDA(cudaMemcpy(ctx->d_tcw, h.data(), total * sizeof(__half), cudaMemcpyHostToDevice));
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return XF_E_CUDA;
  }
  for (int l = 0; l < L_COUNT; ++l) {
    if (ctx->tc_off[l] == (size_t)-1) continue;
    const int taps = kLayers[l].ks * kLayers[l].ks;
    const cuuint64_t dims[2] = {64, (cuuint64_t)taps * 2 * 64};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, 64};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&ctx->tc_wmap[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)((__half*)ctx->d_tcw + ctx->tc_off[l]), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(weights, layer %d) failed: %d", l, (int)r);
      return XF_E_CUDA;
    }
  }
  return XF_OK;
}

static void pick_tile(int H, int W, int& tw_log2) {
  // TH x TW = 128: choose the shape that wastes the fewest out-of-image pixels
  int best = -1;
  long best_cost = 0;
  for (int l = 2; l <= 6; ++l) {
    const int TW = 1 << l, TH = 128 >> l;
    const long cost = (long)cdiv(H, TH) * TH * cdiv(W, TW) * TW;
    if (best < 0 || cost < best_cost) { best = l; best_cost = cost; }
  }
  tw_log2 = best;
}

// in_split: (B,H,W,128) halves. Any of out_split / out_f32 may be null (not both).
int launch_conv_tc(const xfeat_ctx* ctx, int layer, const __half* in_split, int B, int H, int W, __half* out_split,
                   float* out_f32, cudaStream_t st) {
  XF_REQUIRE(conv_tc_eligible(layer) && ctx->d_tcw, "conv_tc: layer %d not prepared for the tensor-core path", layer);
  XF_REQUIRE(out_split || out_f32, "conv_tc: no output");
  PFN_encodeTiled enc = get_encode_tiled();
  ConvTcParams P;
  int tw_log2;
  pick_tile(H, W, tw_log2);
  const int TW = 1 << tw_log2, TH = 128 >> tw_log2;
  const cuuint64_t dims[4] = {128, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {256, (cuuint64_t)W * 256, (cuuint64_t)H * W * 256};
  const cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&P.amap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)in_split, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(activations %dx%dx%d) failed: %d", B, H, W, (int)r);
    return XF_E_CUDA;
  }
  P.wmap = ctx->tc_wmap[layer];
  P.bias = ctx->d_weights + ctx->table.b_off[layer];
  P.inv_wscale = ctx->tc_inv_wscale[layer];
  P.B = B; P.H = H; P.W = W;
  P.tw_log2 = tw_log2;
  P.out_split = out_split;
  P.out_f32 = out_f32;
  P.relu = kLayers[layer].relu;
  const int n_tiles = cdiv(H, TH) * cdiv(W, TW) * B;
  const int grid = n_tiles < ctx->sm_count ? n_tiles : ctx->sm_count;
  if (kLayers[layer].ks == 3) {
    static bool attr3 = false;
    if (!attr3) {
      XF_CUDA(cudaFuncSetAttribute(conv_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ConvTcCfg<3>::SMEM));
      attr3 = true;
    }
    conv_tc_kernel<3><<<grid, CT_THREADS, ConvTcCfg<3>::SMEM, st>>>(P);
  } else {
    static bool attr1 = false;
    if (!attr1) {
      XF_CUDA(cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ConvTcCfg<1>::SMEM));
      attr1 = true;
    }
    conv_tc_kernel<1><<<grid, CT_THREADS, ConvTcCfg<1>::SMEM, st>>>(P);
  }
  XF_LAUNCH_CHECK();
  return XF_OK;
}

}  // namespace xf

// Test hook: one eligible layer through the tensor-core kernel with fp32 NHWC in/out (the split of the input runs first).
extern "C" int xfeat_debug_conv_layer_tc(xfeat_ctx* ctx, int layer, const float* d_in, int B, int H, int W, float* d_out,
                                         void* d_scratch, size_t scratch_bytes, void* stream) {
  XF_REQUIRE(ctx && d_in && d_out && d_scratch && layer >= 0 && layer < xf::L_COUNT, "debug_conv_layer_tc: bad arguments");
  const int64_t npix = (int64_t)B * H * W;
  XF_REQUIRE(scratch_bytes >= (size_t)npix * 256, "debug_conv_layer_tc: scratch must hold B*H*W*256 bytes");
  XF_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc = xf::launch_split_nhwc64(d_in, (__half*)d_scratch, npix, st);
  if (rc) return rc;
  return xf::launch_conv_tc(ctx, layer, (const __half*)d_scratch, B, H, W, nullptr, d_out, st);
}
