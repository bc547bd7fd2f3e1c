// Stream-K form of the tcgen05 projection GEMM for steps whose tokens fit ONE N tile
// (batched decode, short prefill chunks).
//
//   out[t, n] = sum_k X[t, k] * W[n, k]        X bf16 [T <= 128, K], W bf16 [N, K]
//
// Such a step is a weight stream: every byte of W crosses HBM once and the token operand lives in
// L2, so the kernel is HBM-bound and what matters is that all 148 SMs pull for the whole launch.
// Tiling by 128-row slabs alone cannot do that (Llama-3-8B: QKV = 48 slabs, gate/up = 224 = 1.51
// waves).  Here the work is the flat list of (slab, 64-wide K block) units, cut into gridDim.x
// equal contiguous ranges; a CTA's range is [tail piece of a slab][whole slabs][head piece of a
// slab].  A CTA that starts in the middle of a slab computes that piece FIRST, parks the fp32
// accumulator tile in its workspace slot and raises its flag; the CTA that owns the slab's head
// (K block 0) computes it LAST, adds the parked pieces in CTA order and runs the fused epilogue.
// Pieces never wait, heads wait only on work that was started before theirs: no deadlock as long
// as the grid is co-resident (grid <= SM count, 1 CTA/SM), and the summation order is fixed by
// the shape, so results are bit-reproducible.
// Warp roles and pipeline are those of gemm_tc.cu (TMA producer, single-thread MMA issuer, TMEM
// double buffer, 8 epilogue warps).
#include <cuda.h>

#include <cstdlib>

#include "../../include/llmlb_b200.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace llmlb {

__device__ __forceinline__ uint32_t sk_begin(uint32_t cta, uint32_t units, uint32_t grid) {
  return uint32_t(uint64_t(cta) * units / grid);
}

static __device__ TraceBuf d_trace_sk;
void sk_set_trace(const TraceBuf& tb) { cudaMemcpyToSymbol(d_trace_sk, &tb, sizeof(tb)); }

template <int BN, int EPI>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_sk_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
               void* __restrict__ out, uint32_t n_tokens, uint32_t n_out, uint32_t K,
               uint32_t out_stride, uint32_t m_tiles, float* __restrict__ ws, uint32_t* __restrict__ flags) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t KB = (K + kBK - 1) / kBK;
  const uint32_t units = m_tiles * KB, G = gridDim.x;
  const uint32_t u0 = sk_begin(blockIdx.x, units, G), u1 = sk_begin(blockIdx.x + 1, units, G);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(empty_bar + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + i, 1);
      mbar_init(tempty_bar + i, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(uint32_t(Cfg::kTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // debug timeline (tools/batch_decode_profile.py): CTA lifetime on the global timer + cycles the
  // producer waited for free slots / the head waited for parked pieces / epilogue for accumulators
  const TraceBuf tb = d_trace_sk;
  __shared__ long long stall[3];
  const unsigned long long g_begin = tb.data ? gtime_ns() : 0ull;
  long long c_wait = 0, c_flag = 0;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      // L2 prefetch cursor runs ahead of the smem ring along this CTA's unit range
      constexpr uint32_t kPrefetchAhead = 12;
      uint32_t p_mt = u0 / KB, p_kb = u0 - p_mt * KB, p_u = u0;
      auto p_step = [&]() {
        if (p_u >= u1) return;
        tma_prefetch_l2_2d(&tmap_w, int32_t(p_kb * kBK), int32_t(p_mt * kBM));
        ++p_u;
        if (++p_kb == KB) { p_kb = 0; ++p_mt; }
      };
      for (uint32_t i = 0; i < kPrefetchAhead; ++i) p_step();
      uint32_t mt = u0 / KB, kb = u0 - mt * KB;
      for (uint32_t u = u0; u < u1; ++u) {
        p_step();
        { const long long c0 = clock64(); mbar_wait(empty_bar + stage, phase ^ 1); c_wait += clock64() - c0; }
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + kBM * kBK * 2;
        mbar_expect_tx(full_bar + stage, Cfg::kStageBytes);
        tma_load_2d(sa, &tmap_w, full_bar + stage, int32_t(kb * kBK), int32_t(mt * kBM));
        tma_load_2d(sb, &tmap_x, full_bar + stage, int32_t(kb * kBK), 0);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        if (++kb == KB) { kb = 0; ++mt; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (uint32_t u = u0; u < u1;) {
        const uint32_t mt = u / KB, kb0 = u - mt * KB;
        const uint32_t kb1 = min(KB, kb0 + (u1 - u));
        mbar_wait(tempty_bar + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (uint32_t kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + kBM * kBK * 2;
          const uint64_t adesc = make_sw128_desc(sa);
          const uint64_t bdesc = make_sw128_desc(sb);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            tc_mma(tmem_d, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), Cfg::kIdesc,
                   (kb > kb0 || k > 0) ? 1u : 0u);
          tc_commit(empty_bar + stage);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull_bar + acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        u += kb1 - kb0;
      }
    }
  } else if (warp >= 4) {
    const uint32_t q = warp & 3;  // TMEM lane quarter this warp may read
    const uint32_t row = q * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    constexpr uint32_t kColsPerWarp = (BN / 2 >= 16) ? BN / 2 : 16;
    const uint32_t c_begin = ((warp - 4) >> 2) * kColsPerWarp;
    for (uint32_t u = u0; u < u1;) {
      const uint32_t mt = u / KB, kb0 = u - mt * KB;
      const uint32_t kb1 = min(KB, kb0 + (u1 - u));
      const bool piece = kb0 > 0;                 // parked for the slab's head CTA
      const bool head = kb0 == 0 && kb1 < KB;     // finishes a slab other CTAs contributed to
      uint32_t f_lo = blockIdx.x + 1, f_hi = f_lo;
      if (head) {
        const uint32_t slab_end = (mt + 1) * KB;
        while (f_hi < G && sk_begin(f_hi, units, G) < slab_end) ++f_hi;
        const long long c0 = clock64();
        if (lane == 0)
          for (uint32_t f = f_lo; f < f_hi; ++f)
            while (ld_acquire_u32(flags + f) == 0) {}
        __syncwarp();
        c_flag += clock64() - c0;
      }
      const uint32_t n = mt * kBM + row;          // output feature of this thread
      constexpr uint32_t kGroup = kColsPerWarp < 64 ? kColsPerWarp : 64;  // columns gathered per round
      bool waited = false;
#pragma unroll 1
      for (uint32_t cg = c_begin; cg < c_begin + kColsPerWarp && cg < BN; cg += kGroup) {
        if (cg >= n_tokens) break;                // warp-uniform
        // head: the parked pieces for this row x kGroup columns, all loads of a piece in flight at
        // once and issued before the accumulator is awaited
        float p[kGroup];
        if (head) {
#pragma unroll
          for (uint32_t j = 0; j < kGroup; ++j) p[j] = 0.f;
          for (uint32_t f = f_lo; f < f_hi; ++f) {
            const float* slot = ws + (size_t(f) * BN + cg) * kBM + row;
#pragma unroll
            for (uint32_t j = 0; j < kGroup; ++j) p[j] += __ldcg(slot + size_t(j) * kBM);
          }
        }
        if (!waited) {
          const long long c0 = clock64();
          mbar_wait(tfull_bar + acc, acc_phase);
          c_wait += clock64() - c0;
          tc_fence_after();
          waited = true;
        }
#pragma unroll
        for (uint32_t cc = 0; cc < kGroup; cc += 16) {
          const uint32_t c = cg + cc;
          if (c < n_tokens) {
            uint32_t r[16];
            tc_ld16(tmem_base + ((q * 32) << 16) + acc * BN + c, r);
            tc_wait_ld();
            if (piece) {
              float* slot = ws + (size_t(blockIdx.x) * BN + c) * kBM + row;
#pragma unroll
              for (int j = 0; j < 16; ++j) slot[size_t(j) * kBM] = __uint_as_float(r[j]);
            } else {
              if (head) {
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + p[cc + j]);
              }
              if constexpr (EPI == LLMLB_EPI_SILU_MUL) {
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float v = __uint_as_float(r[j]);
                  const float other = __shfl_xor_sync(0xffffffffu, v, 1);
                  // lanes (2i, 2i+1) hold (gate_i, up_i); even lanes finish even columns, odd lanes odd ones
                  if (((j ^ lane) & 1) == 0 && (n | 1) < n_out && c + j < n_tokens) {
                    const float g = (lane & 1) ? other : v, up = (lane & 1) ? v : other;
                    const float sg = __fdividef(g, 1.f + __expf(-g));
                    o[size_t(c + j) * out_stride + (n >> 1)] = __float2bfloat16_rn(sg * up);
                  }
                }
              } else if (n < n_out) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const uint32_t t = c + j;
                  if (t < n_tokens) {
                    const float v = __uint_as_float(r[j]);
                    const size_t idx = size_t(t) * out_stride + n;
                    if constexpr (EPI == LLMLB_EPI_STORE_BF16)
                      reinterpret_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16_rn(v);
                    else if constexpr (EPI == LLMLB_EPI_STORE_F32 || EPI == kEpiPartialF32)
                      reinterpret_cast<float*>(out)[idx] = v;   // a single "part": the slab is complete
                    else
                      reinterpret_cast<float*>(out)[idx] += v;
                  }
                }
              }
            }
          }
        }
      }
      if (!waited) {  // no live columns for this warp: still observe the accumulator phase
        mbar_wait(tfull_bar + acc, acc_phase);
        tc_fence_after();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (piece) {
        __threadfence();
        epi_bar();
        if (warp == 4 && lane == 0) st_release_u32(flags + blockIdx.x, 1u);
      } else if (head) {
        epi_bar();                                 // every warp has consumed the parked pieces
        if (warp == 4 && lane == 0)
          for (uint32_t f = f_lo; f < f_hi; ++f) flags[f] = 0;   // clean for the next launch
      }
      u += kb1 - kb0;
    }
  }

  if (tb.data) {
    if (warp == 0 && lane == 0) stall[0] = c_wait;
    if (warp == 4 && lane == 0) { stall[1] = c_flag; stall[2] = c_wait; }
  }
  tc_fence_before();
  __syncthreads();
  if (tb.data && threadIdx.x == 0)
    trace_emit(tb, (3ull << 60) | ((unsigned long long)n_out << 32) | ((unsigned long long)EPI << 28) | K, g_begin,
               ((unsigned long long)stall[0] << 32) | (unsigned long long)(stall[1] & 0xFFFFFFFF),
               (unsigned long long)stall[2], gtime_ns());
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(uint32_t(Cfg::kTmemCols))
                 : "memory");
  }
}

// ----------------------------------------------------------------- host side ----------------
template <int BN, int EPI>
static int launch_sk(const CUtensorMap& tw, const CUtensorMap& tx, void* out, uint32_t n_tokens,
                     uint32_t n_out, uint32_t k, uint32_t out_stride, const SkWorkspace& sk,
                     cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  auto kern = gemm_sk_kernel<BN, EPI>;
  static bool configured = false;
  if (!configured) {
    LLMLB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::kSmemBytes));
    configured = true;
  }
  const uint32_t m_tiles = (n_out + kBM - 1) / kBM;
  const uint32_t units = m_tiles * ((k + kBK - 1) / kBK);
  // every CTA gets at least 4 K blocks, so no range is empty and heads never wait on idle CTAs
  uint32_t grid = units / 4;
  if (grid < 1) grid = 1;
  if (grid > (uint32_t)kNumSMs) grid = kNumSMs;
  kern<<<grid, kTcThreads, Cfg::kSmemBytes, st>>>(tw, tx, out, n_tokens, n_out, k, out_stride,
                                                  m_tiles, sk.ws, sk.flags);
  LLMLB_LAUNCH_CHECK();
  return LLMLB_OK;
}

template <int BN>
static int dispatch_sk_epi(uint32_t epi, const CUtensorMap& tw, const CUtensorMap& tx, void* out,
                           uint32_t n_tokens, uint32_t n_out, uint32_t k, uint32_t out_stride,
                           const SkWorkspace& sk, cudaStream_t st) {
  switch (epi) {
    case LLMLB_EPI_STORE_BF16:
      return launch_sk<BN, LLMLB_EPI_STORE_BF16>(tw, tx, out, n_tokens, n_out, k, out_stride, sk, st);
    case LLMLB_EPI_RESID_F32:
      return launch_sk<BN, LLMLB_EPI_RESID_F32>(tw, tx, out, n_tokens, n_out, k, out_stride, sk, st);
    case LLMLB_EPI_SILU_MUL:
      return launch_sk<BN, LLMLB_EPI_SILU_MUL>(tw, tx, out, n_tokens, n_out, k, out_stride, sk, st);
    case LLMLB_EPI_STORE_F32:
    case kEpiPartialF32:
      return launch_sk<BN, LLMLB_EPI_STORE_F32>(tw, tx, out, n_tokens, n_out, k, out_stride, sk, st);
  }
  set_error("gemm_sk: unknown epilogue");
  return LLMLB_E_INVALID_ARG;
}

uint32_t tc_pick_bn(uint32_t n_tokens);

// tx must be the activation map whose box holds tc_pick_bn(n_tokens) rows; n_tokens <= 128
int gemm_sk_launch(const CUtensorMap& tw, const CUtensorMap& tx, void* out, uint32_t n_tokens,
                   uint32_t n_out, uint32_t k, uint32_t epi, uint32_t out_stride,
                   const SkWorkspace& sk, cudaStream_t st) {
  if (n_tokens > 128 || !sk.ws || !sk.flags) {
    set_error("gemm_sk: needs n_tokens <= 128 and a workspace");
    return LLMLB_E_INVALID_ARG;
  }
  switch (tc_pick_bn(n_tokens)) {
    case 16: return dispatch_sk_epi<16>(epi, tw, tx, out, n_tokens, n_out, k, out_stride, sk, st);
    case 32: return dispatch_sk_epi<32>(epi, tw, tx, out, n_tokens, n_out, k, out_stride, sk, st);
    case 64: return dispatch_sk_epi<64>(epi, tw, tx, out, n_tokens, n_out, k, out_stride, sk, st);
    default: return dispatch_sk_epi<128>(epi, tw, tx, out, n_tokens, n_out, k, out_stride, sk, st);
  }
}

int sk_workspace_create(SkWorkspace* sk, cudaStream_t st) {
  LLMLB_CUDA_CHECK(cudaMalloc(&sk->ws, kSkWsBytes));
  LLMLB_CUDA_CHECK(cudaMalloc(&sk->flags, kNumSMs * sizeof(uint32_t)));
  LLMLB_CUDA_CHECK(cudaMemsetAsync(sk->flags, 0, kNumSMs * sizeof(uint32_t), st));
  return LLMLB_OK;
}
void sk_workspace_destroy(SkWorkspace* sk) {
  if (sk->ws) cudaFree(sk->ws);
  if (sk->flags) cudaFree(sk->flags);
  sk->ws = nullptr;
  sk->flags = nullptr;
}

}  // namespace llmlb
