// Decode GEMV, bulk-copy variant: weights flow HBM -> shared memory through a ring of
// cp.async.bulk (TMA 1-D) stages, consumers read them with conflict-free LDS.128.
//
// Versus gemv_ks (register-staged loads) this keeps the loads-in-flight in SHARED MEMORY
// (3 x 32 KiB per CTA, ~100 registers/thread), so two CTAs fit on an SM: with programmatic
// dependent launch the NEXT projection's CTA becomes resident while the current one drains and its
// producer thread pulls 96 KiB of weights before `griddepcontrol.wait` — HBM stays busy across
// kernel boundaries.  Rows of a CTA are contiguous in memory, so one stage is one 32 KiB
// contiguous bulk copy (L2 evict-first: weights are read once per token).
//   warps 0..7  consumers: warp w owns K chunks [w*CPW, (w+1)*CPW) (256 elements each) of every
//               row; x slice in fp32 registers, RMSNorm fused; per-row shuffle reduce ->
//               partial[window][row][warp]; every 32 rows a consumer-only named barrier and the
//               first lanes finish the sums + epilogue (residual add / SiLU*up / store)
//   warp 8      producer: one thread, empty/full mbarrier ring
#include "../../include/llmlb_b200.h"
#include "common.cuh"

namespace llmlb {

constexpr int kBkConsumerWarps = 8;
constexpr int kBkThreads = (kBkConsumerWarps + 1) * 32;
constexpr int kBkStageBytes = 32768;
constexpr int kBkStages = 3;
constexpr int kBkWindow = 32;  // rows between cross-warp reductions

__device__ __forceinline__ uint32_t bk_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bk_mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bk_smem(b)), "r"(n));
}
__device__ __forceinline__ void bk_mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(bk_smem(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bk_mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bk_smem(b)) : "memory");
}
__device__ __forceinline__ void bk_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bk_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bk_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(bk_smem(dst)), "l"(src), "r"(bytes), "r"(bk_smem(bar)), "l"(policy) : "memory");
}

template <int B, int EPI, bool NORM, int CPW>
__global__ void __launch_bounds__(kBkThreads)
gemv_bulk_kernel(const __nv_bfloat16* __restrict__ W, const void* __restrict__ xin,
                 const __nv_bfloat16* __restrict__ gain, float eps, void* __restrict__ out,
                 uint32_t n_out, uint32_t K, uint32_t out_stride, uint32_t R /*rows per stage*/) {
  extern __shared__ __align__(128) uint8_t bk_dyn[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bk_dyn) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full_bar[kBkStages], empty_bar[kBkStages];
  __shared__ float partial[2][kBkWindow][kBkConsumerWarps][B];
  __shared__ float red[B][kBkConsumerWarps];

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n_pairs = (n_out + 1) / 2;
  const uint32_t row_begin = uint32_t((uint64_t(blockIdx.x) * n_pairs) / gridDim.x) * 2;
  const uint32_t row_end = min(n_out, uint32_t((uint64_t(blockIdx.x + 1) * n_pairs) / gridDim.x) * 2);
  const uint32_t n_rows = row_end - row_begin;
  const uint32_t n_stages = (n_rows + R - 1) / R;
  const uint32_t row_bytes = K * 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBkStages; ++i) {
      bk_mbar_init(&full_bar[i], 1);
      bk_mbar_init(&empty_bar[i], kBkConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // PDL: the next kernel may launch now; everything before griddepcontrol.wait touches weights only
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == kBkConsumerWarps) {
    // ---------------- producer ----------------
    if (lane == 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      for (uint32_t j = 0; j < n_stages; ++j) {
        const uint32_t s = j % kBkStages;
        bk_mbar_wait(&empty_bar[s], ((j / kBkStages) & 1) ^ 1);
        const uint32_t rows = min(R, n_rows - j * R);
        const uint32_t bytes = rows * row_bytes;
        bk_expect_tx(&full_bar[s], bytes);
        bk_bulk_load(ring + size_t(s) * kBkStageBytes, W + size_t(row_begin + j * R) * K, bytes, &full_bar[s], policy);
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float xr[B][CPW][8];
  if constexpr (NORM) {
    const float* xf = reinterpret_cast<const float*>(xin);
    float ss[B];
#pragma unroll
    for (int b = 0; b < B; ++b) ss[b] = 0.f;
    for (uint32_t i = threadIdx.x; i < K / 4; i += kBkConsumerWarps * 32) {
#pragma unroll
      for (int b = 0; b < B; ++b) {
        float4 v = reinterpret_cast<const float4*>(xf + size_t(b) * K)[i];
        ss[b] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
      float s = warp_sum(ss[b]);
      if (lane == 0) red[b][warp] = s;
    }
    asm volatile("bar.sync 2, %0;" ::"n"(kBkConsumerWarps * 32) : "memory");
#pragma unroll
    for (int b = 0; b < B; ++b) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kBkConsumerWarps; ++w) tot += red[b][w];
      const float rs = rsqrtf(tot / float(K) + eps);
#pragma unroll
      for (int c = 0; c < CPW; ++c) {
        const uint32_t k0 = (warp * CPW + c) * 256u + lane * 8u;
        const float4 v0 = *reinterpret_cast<const float4*>(xf + size_t(b) * K + k0);
        const float4 v1 = *reinterpret_cast<const float4*>(xf + size_t(b) * K + k0 + 4);
        const uint4 g = __ldg(reinterpret_cast<const uint4*>(gain + k0));
        xr[b][c][0] = __bfloat162float(__float2bfloat16_rn(v0.x * rs * bf16_lo(g.x)));
        xr[b][c][1] = __bfloat162float(__float2bfloat16_rn(v0.y * rs * bf16_hi(g.x)));
        xr[b][c][2] = __bfloat162float(__float2bfloat16_rn(v0.z * rs * bf16_lo(g.y)));
        xr[b][c][3] = __bfloat162float(__float2bfloat16_rn(v0.w * rs * bf16_hi(g.y)));
        xr[b][c][4] = __bfloat162float(__float2bfloat16_rn(v1.x * rs * bf16_lo(g.z)));
        xr[b][c][5] = __bfloat162float(__float2bfloat16_rn(v1.y * rs * bf16_hi(g.z)));
        xr[b][c][6] = __bfloat162float(__float2bfloat16_rn(v1.z * rs * bf16_lo(g.w)));
        xr[b][c][7] = __bfloat162float(__float2bfloat16_rn(v1.w * rs * bf16_hi(g.w)));
      }
    }
  } else {
    const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(xin);
#pragma unroll
    for (int b = 0; b < B; ++b)
#pragma unroll
      for (int c = 0; c < CPW; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(xb + size_t(b) * K + (warp * CPW + c) * 256u + lane * 8u);
        xr[b][c][0] = bf16_lo(v.x); xr[b][c][1] = bf16_hi(v.x);
        xr[b][c][2] = bf16_lo(v.y); xr[b][c][3] = bf16_hi(v.y);
        xr[b][c][4] = bf16_lo(v.z); xr[b][c][5] = bf16_hi(v.z);
        xr[b][c][6] = bf16_lo(v.w); xr[b][c][7] = bf16_hi(v.w);
      }
  }

  const uint32_t stages_per_window = kBkWindow / R;
  uint32_t wbuf = 0;
  for (uint32_t j = 0; j < n_stages; ++j) {
    const uint32_t s = j % kBkStages;
    bk_mbar_wait(&full_bar[s], (j / kBkStages) & 1);
    const uint8_t* st = ring + size_t(s) * kBkStageBytes;
    const uint32_t rows = min(R, n_rows - j * R);
    const uint32_t wrow0 = (j % stages_per_window) * R;  // row within the window
    for (uint32_t r = 0; r < rows; ++r) {
      float acc[B];
#pragma unroll
      for (int b = 0; b < B; ++b) acc[b] = 0.f;
#pragma unroll
      for (int c = 0; c < CPW; ++c) {
        const uint4 w = *reinterpret_cast<const uint4*>(st + size_t(r) * row_bytes + ((warp * CPW + c) * 256u + lane * 8u) * 2u);
        const float w0 = bf16_lo(w.x), w1 = bf16_hi(w.x), w2 = bf16_lo(w.y), w3 = bf16_hi(w.y);
        const float w4 = bf16_lo(w.z), w5 = bf16_hi(w.z), w6 = bf16_lo(w.w), w7 = bf16_hi(w.w);
#pragma unroll
        for (int b = 0; b < B; ++b) {
          float a = acc[b];
          a = fmaf(w0, xr[b][c][0], a); a = fmaf(w1, xr[b][c][1], a);
          a = fmaf(w2, xr[b][c][2], a); a = fmaf(w3, xr[b][c][3], a);
          a = fmaf(w4, xr[b][c][4], a); a = fmaf(w5, xr[b][c][5], a);
          a = fmaf(w6, xr[b][c][6], a); a = fmaf(w7, xr[b][c][7], a);
          acc[b] = a;
        }
      }
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const float sum = warp_sum(acc[b]);
        if (lane == 0) partial[wbuf][wrow0 + r][warp][b] = sum;
      }
    }
    __syncwarp();
    if (lane == 0) bk_mbar_arrive(&empty_bar[s]);  // this warp is done reading the stage

    const bool window_done = ((j + 1) % stages_per_window == 0) || (j + 1 == n_stages);
    if (window_done) {
      asm volatile("bar.sync 1, %0;" ::"n"(kBkConsumerWarps * 32) : "memory");
      const uint32_t win_first_stage = (j / stages_per_window) * stages_per_window;
      const uint32_t win_row0 = row_begin + win_first_stage * R;
      const uint32_t win_rows = min(uint32_t(kBkWindow), row_end - win_row0);
      // threads 0 .. win_rows*B-1 finish one (row, token) each
      const uint32_t t = threadIdx.x;
      const uint32_t r = t / B, b = t % B;
      float v = 0.f;
      if (t < win_rows * B) {
#pragma unroll
        for (int w = 0; w < kBkConsumerWarps; ++w) v += partial[wbuf][r][w][b];
      }
      const uint32_t row = win_row0 + r;
      if constexpr (EPI == LLMLB_EPI_SILU_MUL) {
        // rows r (gate) and r+1 (up) of the same token are B threads apart, within one warp
        // because kBkWindow*B <= 128 and pairs never straddle a 32-thread boundary for B in {1,2,4}
        const float up = __shfl_down_sync(0xffffffffu, v, B);
        if (t < win_rows * B && (r & 1) == 0 && row + 1 < row_end) {
          const float sg = v / (1.f + __expf(-v));
          reinterpret_cast<__nv_bfloat16*>(out)[size_t(b) * out_stride + (row >> 1)] = __float2bfloat16_rn(sg * up);
        }
      } else if (t < win_rows * B && row < row_end) {
        const size_t idx = size_t(b) * out_stride + row;
        if constexpr (EPI == LLMLB_EPI_STORE_BF16) reinterpret_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16_rn(v);
        else if constexpr (EPI == LLMLB_EPI_RESID_F32) reinterpret_cast<float*>(out)[idx] += v;
        else reinterpret_cast<float*>(out)[idx] = v;
      }
      wbuf ^= 1;
    }
  }
}

template <int B, int EPI, bool NORM, int CPW>
static int bulk_launch(const void* w, const void* x, const void* gain, float eps, void* out,
                       uint32_t n_out, uint32_t k, uint32_t out_stride, uint32_t rows_per_stage, cudaStream_t st) {
  auto kern = gemv_bulk_kernel<B, EPI, NORM, CPW>;
  constexpr int smem = kBkStages * kBkStageBytes + 128;
  static bool configured = false;
  if (!configured) {
    LLMLB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  uint32_t n_pairs = (n_out + 1) / 2;
  uint32_t grid = n_pairs < (uint32_t)kNumSMs ? n_pairs : (uint32_t)kNumSMs;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kBkThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LLMLB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, (const __nv_bfloat16*)w, x, (const __nv_bfloat16*)gain, eps, out,
                                      n_out, k, out_stride, rows_per_stage));
  LLMLB_LAUNCH_CHECK();
  return LLMLB_OK;
}

template <int B, int CPW>
static int bulk_dispatch(uint32_t epi, bool norm, const void* w, const void* x, const void* gain, float eps,
                         void* out, uint32_t n_out, uint32_t k, uint32_t out_stride, uint32_t rps, cudaStream_t st) {
#define BK_CASE(E)                                                                                 \
  case E:                                                                                          \
    return norm ? bulk_launch<B, E, true, CPW>(w, x, gain, eps, out, n_out, k, out_stride, rps, st) \
                : bulk_launch<B, E, false, CPW>(w, x, gain, eps, out, n_out, k, out_stride, rps, st);
  switch (epi) {
    BK_CASE(LLMLB_EPI_STORE_BF16)
    BK_CASE(LLMLB_EPI_RESID_F32)
    BK_CASE(LLMLB_EPI_SILU_MUL)
    BK_CASE(LLMLB_EPI_STORE_F32)
  }
#undef BK_CASE
  set_error("gemv_bulk: unknown epilogue");
  return LLMLB_E_INVALID_ARG;
}

// LLMLB_E_UNSUPPORTED when the shape does not fit (caller falls back to gemv_ks / gemv)
int gemv_bulk_try(const void* w, const void* x, const void* gain, float eps, void* out, uint32_t n_tokens,
                  uint32_t n_out, uint32_t k, uint32_t epi, uint32_t out_stride, cudaStream_t st) {
  if (k % (256 * kBkConsumerWarps) || n_tokens == 0 || n_tokens == 3 || n_tokens > 4) return LLMLB_E_UNSUPPORTED;
  const uint32_t cpw = k / (256 * kBkConsumerWarps);
  if (!(cpw == 1 || cpw == 2 || cpw == 4 || cpw == 7) || n_tokens * cpw > 8) return LLMLB_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(w) & 15) != 0) return LLMLB_E_UNSUPPORTED;
  uint32_t rps = kBkStageBytes / (k * 2);  // rows per stage, rounded down to a power of two <= 32
  uint32_t p2 = 1;
  while (p2 * 2 <= rps && p2 * 2 <= (uint32_t)kBkWindow) p2 *= 2;
  rps = p2;
  if (rps == 0 || k * 2 > (uint32_t)kBkStageBytes) return LLMLB_E_UNSUPPORTED;
  const bool norm = gain != nullptr;
#define BK_GO(BB, CC) return bulk_dispatch<BB, CC>(epi, norm, w, x, gain, eps, out, n_out, k, out_stride, rps, st)
  if (n_tokens == 1) { if (cpw == 1) BK_GO(1, 1); if (cpw == 2) BK_GO(1, 2); if (cpw == 4) BK_GO(1, 4); BK_GO(1, 7); }
  if (n_tokens == 2) { if (cpw == 1) BK_GO(2, 1); if (cpw == 2) BK_GO(2, 2); BK_GO(2, 4); }
  if (cpw == 1) BK_GO(4, 1);
  BK_GO(4, 2);
#undef BK_GO
}

}  // namespace llmlb
