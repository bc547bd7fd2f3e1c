// Legacy-tensor-path GEMM (mma.sync m16n8k16, cp.async ring).  This is the recompiled-Ampere
// style kernel the tcgen05 path in gemm_tc.cu is measured against (gemm_impl = 1); it is also
// what llmlb_op_gemm(impl=1) runs so the two can be parity-checked on identical inputs.
//   out[t, n] = sum_k X[t, k] * W[n, k]
// CTA tile 128 tokens x 128 outputs x 32 k, 8 warps (2 x 4), warp tile 64 x 32, 3 stages.
#include <cuda.h>

#include <cstdlib>

#include "../../include/llmlb_b200.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace llmlb {

int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols,
                   uint32_t box_rows);
uint32_t tc_pick_bn(uint32_t n_tokens);
int gemm_tc_launch(const CUtensorMap& tw, const CUtensorMap& tx, void* out, uint32_t n_tokens,
                   uint32_t n_out, uint32_t k, uint32_t epi, uint32_t out_stride, cudaStream_t st,
                   const CUtensorMap* tx_half = nullptr, uint32_t* n_parts = nullptr,
                   const SkWorkspace* sk = nullptr);
int sk_workspace_create(SkWorkspace* sk, cudaStream_t st);

constexpr int kMT = 128, kMN = 128, kMK = 32, kMStages = 3, kMThreads = 256;

__device__ __forceinline__ void cpa16(void* smem, const void* g, bool valid) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(g), "r"(sz));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(s));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 64-byte rows (4 chunks of 16 B); chunk ^= (row>>1)&3 keeps the 8 rows of an ldmatrix apart
__device__ __forceinline__ uint32_t swz32(uint32_t row, uint32_t chunk) {
  return row * kMK + ((chunk ^ ((row >> 1) & 3)) << 3);
}

template <int EPI>
__global__ void __launch_bounds__(kMThreads)
gemm_mma_kernel(const __nv_bfloat16* __restrict__ W, const __nv_bfloat16* __restrict__ X,
                void* __restrict__ out, uint32_t n_tokens, uint32_t n_out, uint32_t K,
                uint32_t out_stride) {
  __shared__ __align__(128) __nv_bfloat16 sx[kMStages][kMT * kMK];
  __shared__ __align__(128) __nv_bfloat16 sw[kMStages][kMN * kMK];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t wm = warp >> 2, wn = warp & 3;  // 2 x 4
  const uint32_t t0 = blockIdx.y * kMT, n0 = blockIdx.x * kMN;
  const uint32_t k_tiles = (K + kMK - 1) / kMK;

  auto load_stage = [&](uint32_t kt, uint32_t st) {
#pragma unroll
    for (uint32_t i = 0; i < 2; ++i) {
      uint32_t c = tid + i * kMThreads;  // 512 chunks per operand
      uint32_t r = c >> 2, ch = c & 3;
      uint32_t kk = kt * kMK + ch * 8;
      bool vx = (t0 + r < n_tokens) && kk < K;
      bool vw = (n0 + r < n_out) && kk < K;
      cpa16(&sx[st][swz32(r, ch)], X + size_t(vx ? t0 + r : 0) * K + (vx ? kk : 0), vx);
      cpa16(&sw[st][swz32(r, ch)], W + size_t(vw ? n0 + r : 0) * K + (vw ? kk : 0), vw);
    }
  };

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

  for (uint32_t s = 0; s < kMStages - 1; ++s) {
    if (s < k_tiles) load_stage(s, s);
    asm volatile("cp.async.commit_group;");
  }
  for (uint32_t kt = 0; kt < k_tiles; ++kt) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kMStages - 2));
    __syncthreads();
    {
      uint32_t nk = kt + kMStages - 1;
      if (nk < k_tiles) load_stage(nk, nk % kMStages);
      asm volatile("cp.async.commit_group;");
    }
    const __nv_bfloat16* tx = sx[kt % kMStages];
    const __nv_bfloat16* tw = sw[kt % kMStages];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t af[4][4], bf[2][4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        uint32_t m = lane >> 3;
        uint32_t r = wm * 64 + mi * 16 + (lane & 7) + (m & 1) * 8;
        ldsm4(af[mi], tx + swz32(r, ks * 2 + (m >> 1)));
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {  // each x4 covers two 8-wide n blocks for this k step
        uint32_t m = lane >> 3;
        uint32_t r = wn * 32 + np * 16 + (m >> 1) * 8 + (lane & 7);
        ldsm4(bf[np], tw + swz32(r, ks * 2 + (m & 1)));
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
          mma16816(acc[mi][ni], af[mi], bf[ni >> 1][(ni & 1) * 2], bf[ni >> 1][(ni & 1) * 2 + 1]);
    }
  }
  asm volatile("cp.async.wait_group 0;");

  const uint32_t g = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t t = t0 + wm * 64 + mi * 16 + g + h * 8;
        uint32_t n = n0 + wn * 32 + ni * 8 + t4 * 2;
        if (t >= n_tokens || n >= n_out) continue;
        float v0 = acc[mi][ni][h * 2], v1 = acc[mi][ni][h * 2 + 1];
        if constexpr (EPI == LLMLB_EPI_SILU_MUL) {
          float s = v0 / (1.f + __expf(-v0));
          reinterpret_cast<__nv_bfloat16*>(out)[size_t(t) * out_stride + (n >> 1)] =
              __float2bfloat16_rn(s * v1);
        } else if constexpr (EPI == LLMLB_EPI_STORE_BF16) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + size_t(t) * out_stride + n;
          o[0] = __float2bfloat16_rn(v0);
          if (n + 1 < n_out) o[1] = __float2bfloat16_rn(v1);
        } else if constexpr (EPI == LLMLB_EPI_STORE_F32) {
          float* o = reinterpret_cast<float*>(out) + size_t(t) * out_stride + n;
          o[0] = v0;
          if (n + 1 < n_out) o[1] = v1;
        } else {
          float* o = reinterpret_cast<float*>(out) + size_t(t) * out_stride + n;
          o[0] += v0;
          if (n + 1 < n_out) o[1] += v1;
        }
      }
}

int gemm_mma_launch(const void* w, const void* x, void* out, uint32_t n_tokens, uint32_t n_out,
                    uint32_t k, uint32_t epi, uint32_t out_stride, cudaStream_t st) {
  dim3 grid((n_out + kMN - 1) / kMN, (n_tokens + kMT - 1) / kMT);
  const __nv_bfloat16* W = (const __nv_bfloat16*)w;
  const __nv_bfloat16* X = (const __nv_bfloat16*)x;
  switch (epi) {
    case LLMLB_EPI_STORE_BF16:
      gemm_mma_kernel<LLMLB_EPI_STORE_BF16><<<grid, kMThreads, 0, st>>>(W, X, out, n_tokens, n_out, k, out_stride);
      break;
    case LLMLB_EPI_RESID_F32:
      gemm_mma_kernel<LLMLB_EPI_RESID_F32><<<grid, kMThreads, 0, st>>>(W, X, out, n_tokens, n_out, k, out_stride);
      break;
    case LLMLB_EPI_SILU_MUL:
      gemm_mma_kernel<LLMLB_EPI_SILU_MUL><<<grid, kMThreads, 0, st>>>(W, X, out, n_tokens, n_out, k, out_stride);
      break;
    case LLMLB_EPI_STORE_F32:
      gemm_mma_kernel<LLMLB_EPI_STORE_F32><<<grid, kMThreads, 0, st>>>(W, X, out, n_tokens, n_out, k, out_stride);
      break;
    default:
      set_error("gemm_mma: unknown epilogue");
      return LLMLB_E_INVALID_ARG;
  }
  LLMLB_LAUNCH_CHECK();
  return LLMLB_OK;
}

}  // namespace llmlb

using namespace llmlb;

extern "C" int llmlb_op_gemm(const void* w, const void* x, void* out, uint32_t n_tokens,
                             uint32_t n_out, uint32_t k, uint32_t epilogue, uint32_t out_stride,
                             uint32_t impl, void* stream) {
  if (!w || !x || !out || n_out == 0 || k == 0 || k % 8 != 0) {
    set_error("llmlb_op_gemm: bad argument (k must be a multiple of 8)");
    return LLMLB_E_INVALID_ARG;
  }
  if (epilogue == LLMLB_EPI_SILU_MUL && (n_out & 1)) {
    set_error("llmlb_op_gemm: SILU_MUL needs interleaved gate/up rows (even n_out)");
    return LLMLB_E_INVALID_ARG;
  }
  if (n_tokens == 0) return LLMLB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == 1) return gemm_mma_launch(w, x, out, n_tokens, n_out, k, epilogue, out_stride, st);
  if (impl > 2 || (impl == 2 && n_tokens > 128)) {
    set_error("llmlb_op_gemm: impl must be 0 (tcgen05 tiles), 1 (mma.sync) or 2 (tcgen05 stream-K, n_tokens <= 128)");
    return LLMLB_E_INVALID_ARG;
  }
  // impl 2: stream-K work split (gemm_sk.cu, one token tile); scratch is per device and the
  // op-level entry point is synchronous test/bench surface, so one static workspace is enough
  const SkWorkspace* sk = nullptr;
  if (impl == 2) {
    static SkWorkspace sk_dev[64];
    int dev = 0;
    LLMLB_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("llmlb_op_gemm: device index"); return LLMLB_E_INVALID_ARG; }
    if (!sk_dev[dev].ws) {
      int rc = sk_workspace_create(&sk_dev[dev], st);
      if (rc) return rc;
      sk_dev[dev].force = true;
    }
    sk = &sk_dev[dev];
  }
  CUtensorMap tw, tx;
  int rc = make_tmap_bf16(&tw, w, n_out, k, 128);
  if (rc) return rc;
  rc = make_tmap_bf16(&tx, x, n_tokens, k, tc_pick_bn(n_tokens));
  if (rc) return rc;
  CUtensorMap txh;
  rc = make_tmap_bf16(&txh, x, n_tokens, k, 128);
  if (rc) return rc;
  return gemm_tc_launch(tw, tx, out, n_tokens, n_out, k, epilogue, out_stride, st, &txh, nullptr, sk);
}
