// Batch-1 decode: a CHAIN of dependent GEMVs in one persistent kernel.
//
// Between two attention kernels a decoder layer runs four dependent projections
//   O-proj (+residual) -> [RMSNorm] gate/up (SiLU*up) -> down (+residual) -> [RMSNorm] next QKV
// As separate kernels each one restarts the HBM pipeline (~2-3 us of ramp per launch, x129 per
// token).  Here one CTA per SM walks the whole chain:
//   * warp 8 (one thread) is the weight PRODUCER: it streams this CTA's rows of op 0, then op 1,
//     ... as 32 KiB contiguous cp.async.bulk stages into a 6-deep shared-memory ring and never
//     waits for the activations — weights do not depend on them — so HBM stays busy across the
//     dependency points (192 KiB per SM in flight = 28 MB chip-wide = ~4 us of HBM time);
//   * warps 0..7 are CONSUMERS (K-split: warp w owns chunks [w*CPW,(w+1)*CPW) of every row; x
//     slice in registers; RMSNorm fused; epilogues fused as in gemv_ks);
//   * between ops a device-wide barrier (one atomic per CTA on a monotonically increasing
//     counter + acquire spin) orders the activations; x is re-read with ld.global.cg (L2).
// With PDL the producer starts streaming before the preceding attention kernel has finished.
#include "../../include/llmlb_b200.h"
#include "common.cuh"

namespace llmlb {

constexpr int kChConsumerWarps = 8;
constexpr int kChThreads = (kChConsumerWarps + 1) * 32;
constexpr int kChStageBytes = 32768;
constexpr int kChStages = 6;
constexpr int kChWindow = 32;
constexpr int kChMaxOps = 4;
constexpr int kChMaxCpw = 7;

struct ChainOp {
  const __nv_bfloat16* W;
  const void* x;               // fp32 residual (gain != null) or bf16 activation
  const __nv_bfloat16* gain;   // RMSNorm gain or null
  void* out;
  uint32_t n_out, K, out_stride, epi;
  uint32_t rows_per_stage, cpw;
};
struct ChainArgs {
  ChainOp op[kChMaxOps];
  uint32_t n_ops;
  float eps;
  uint32_t* state;  // [0] = epoch (launches so far), [1 + i] = arrivals at the barrier after op i
};

__device__ __forceinline__ uint32_t ch_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ch_mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ch_smem(b)), "r"(n));
}
__device__ __forceinline__ void ch_mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(ch_smem(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void ch_mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ch_smem(b)) : "memory");
}
__device__ __forceinline__ void ch_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ch_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ch_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(ch_smem(dst)), "l"(src), "r"(bytes), "r"(ch_smem(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void ch_rows(const ChainOp& o, uint32_t& row_begin, uint32_t& row_end) {
  const uint32_t n_pairs = (o.n_out + 1) / 2;
  row_begin = uint32_t((uint64_t(blockIdx.x) * n_pairs) / gridDim.x) * 2;
  row_end = min(o.n_out, uint32_t((uint64_t(blockIdx.x + 1) * n_pairs) / gridDim.x) * 2);
}

__global__ void __launch_bounds__(kChThreads, 1) gemv_chain_kernel(const __grid_constant__ ChainArgs A) {
  extern __shared__ __align__(128) uint8_t ch_dyn[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ch_dyn) + 127) & ~uintptr_t(127));
  __shared__ uint64_t full_bar[kChStages], empty_bar[kChStages];
  __shared__ float partial[2][kChWindow][kChConsumerWarps];
  __shared__ float red[kChConsumerWarps];

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kChStages; ++i) {
      ch_mbar_init(&full_bar[i], 1);
      ch_mbar_init(&empty_bar[i], kChConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == kChConsumerWarps) {
    // ---------------- producer: all ops back to back, gated only by ring slots ----------------
    if (lane == 0) {
      uint64_t policy;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      uint32_t j = 0;  // global stage counter
      for (uint32_t oi = 0; oi < A.n_ops; ++oi) {
        const ChainOp& o = A.op[oi];
        uint32_t rb, re;
        ch_rows(o, rb, re);
        const uint32_t n_rows = re - rb, R = o.rows_per_stage, row_bytes = o.K * 2;
        for (uint32_t r0 = 0; r0 < n_rows; r0 += R, ++j) {
          const uint32_t s = j % kChStages;
          ch_mbar_wait(&empty_bar[s], ((j / kChStages) & 1) ^ 1);
          const uint32_t bytes = min(R, n_rows - r0) * row_bytes;
          ch_expect_tx(&full_bar[s], bytes);
          ch_bulk_load(ring + size_t(s) * kChStageBytes, o.W + size_t(rb + r0) * o.K, bytes, &full_bar[s], policy);
        }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  const uint32_t epoch = A.state[0] + 1;  // same value in every CTA: state[0] is bumped at the very end
  asm volatile("griddepcontrol.wait;" ::: "memory");
  uint32_t j = 0;
  for (uint32_t oi = 0; oi < A.n_ops; ++oi) {
    const ChainOp& o = A.op[oi];
    const uint32_t K = o.K, cpw = o.cpw, R = o.rows_per_stage, row_bytes = K * 2;
    uint32_t row_begin, row_end;
    ch_rows(o, row_begin, row_end);
    const uint32_t n_rows = row_end - row_begin;

    // ---- x slice of this lane (fp32 registers) ----
    float xr[kChMaxCpw][8];
    if (o.gain) {
      const float* xf = reinterpret_cast<const float*>(o.x);
      float ss = 0.f;
      for (uint32_t i = threadIdx.x; i < K / 4; i += kChConsumerWarps * 32) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(xf) + i);
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
      ss = warp_sum(ss);
      if (lane == 0) red[warp] = ss;
      asm volatile("bar.sync 2, %0;" ::"n"(kChConsumerWarps * 32) : "memory");
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kChConsumerWarps; ++w) tot += red[w];
      const float rs = rsqrtf(tot / float(K) + A.eps);
#pragma unroll
      for (int c = 0; c < kChMaxCpw; ++c) {
        if (uint32_t(c) < cpw) {
          const uint32_t k0 = (warp * cpw + c) * 256u + lane * 8u;
          const float4 v0 = __ldcg(reinterpret_cast<const float4*>(xf + k0));
          const float4 v1 = __ldcg(reinterpret_cast<const float4*>(xf + k0 + 4));
          const uint4 g = __ldg(reinterpret_cast<const uint4*>(o.gain + k0));
          xr[c][0] = __bfloat162float(__float2bfloat16_rn(v0.x * rs * bf16_lo(g.x)));
          xr[c][1] = __bfloat162float(__float2bfloat16_rn(v0.y * rs * bf16_hi(g.x)));
          xr[c][2] = __bfloat162float(__float2bfloat16_rn(v0.z * rs * bf16_lo(g.y)));
          xr[c][3] = __bfloat162float(__float2bfloat16_rn(v0.w * rs * bf16_hi(g.y)));
          xr[c][4] = __bfloat162float(__float2bfloat16_rn(v1.x * rs * bf16_lo(g.z)));
          xr[c][5] = __bfloat162float(__float2bfloat16_rn(v1.y * rs * bf16_hi(g.z)));
          xr[c][6] = __bfloat162float(__float2bfloat16_rn(v1.z * rs * bf16_lo(g.w)));
          xr[c][7] = __bfloat162float(__float2bfloat16_rn(v1.w * rs * bf16_hi(g.w)));
        }
      }
    } else {
      const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(o.x);
#pragma unroll
      for (int c = 0; c < kChMaxCpw; ++c) {
        if (uint32_t(c) < cpw) {
          const uint4 v = __ldcg(reinterpret_cast<const uint4*>(xb + (warp * cpw + c) * 256u + lane * 8u));
          xr[c][0] = bf16_lo(v.x); xr[c][1] = bf16_hi(v.x); xr[c][2] = bf16_lo(v.y); xr[c][3] = bf16_hi(v.y);
          xr[c][4] = bf16_lo(v.z); xr[c][5] = bf16_hi(v.z); xr[c][6] = bf16_lo(v.w); xr[c][7] = bf16_hi(v.w);
        }
      }
    }

    // ---- stream this CTA's rows ----
    const uint32_t n_stages = (n_rows + R - 1) / R;
    const uint32_t stages_per_window = kChWindow / R;
    uint32_t wbuf = 0;
    for (uint32_t sj = 0; sj < n_stages; ++sj, ++j) {
      const uint32_t s = j % kChStages;
      ch_mbar_wait(&full_bar[s], (j / kChStages) & 1);
      const uint8_t* st = ring + size_t(s) * kChStageBytes;
      const uint32_t rows = min(R, n_rows - sj * R);
      const uint32_t wrow0 = (sj % stages_per_window) * R;
      for (uint32_t r = 0; r < rows; ++r) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < kChMaxCpw; ++c) {
          if (uint32_t(c) < cpw) {
            const uint4 w = *reinterpret_cast<const uint4*>(st + size_t(r) * row_bytes + ((warp * cpw + c) * 256u + lane * 8u) * 2u);
            acc = fmaf(bf16_lo(w.x), xr[c][0], acc); acc = fmaf(bf16_hi(w.x), xr[c][1], acc);
            acc = fmaf(bf16_lo(w.y), xr[c][2], acc); acc = fmaf(bf16_hi(w.y), xr[c][3], acc);
            acc = fmaf(bf16_lo(w.z), xr[c][4], acc); acc = fmaf(bf16_hi(w.z), xr[c][5], acc);
            acc = fmaf(bf16_lo(w.w), xr[c][6], acc); acc = fmaf(bf16_hi(w.w), xr[c][7], acc);
          }
        }
        acc = warp_sum(acc);
        if (lane == 0) partial[wbuf][wrow0 + r][warp] = acc;
      }
      __syncwarp();
      if (lane == 0) ch_mbar_arrive(&empty_bar[s]);

      const bool window_done = ((sj + 1) % stages_per_window == 0) || (sj + 1 == n_stages);
      if (window_done) {
        asm volatile("bar.sync 1, %0;" ::"n"(kChConsumerWarps * 32) : "memory");
        const uint32_t win_row0 = row_begin + (sj / stages_per_window) * stages_per_window * R;
        const uint32_t win_rows = min(uint32_t(kChWindow), row_end - win_row0);
        const uint32_t t = threadIdx.x;
        float v = 0.f;
        if (t < win_rows) {
#pragma unroll
          for (int w = 0; w < kChConsumerWarps; ++w) v += partial[wbuf][t][w];
        }
        const uint32_t row = win_row0 + t;
        if (o.epi == LLMLB_EPI_SILU_MUL) {
          const float up = __shfl_down_sync(0xffffffffu, v, 1);
          if (t < win_rows && (t & 1) == 0 && row + 1 < row_end) {
            const float sg = v / (1.f + __expf(-v));
            reinterpret_cast<__nv_bfloat16*>(o.out)[row >> 1] = __float2bfloat16_rn(sg * up);
          }
        } else if (t < win_rows && row < row_end) {
          if (o.epi == LLMLB_EPI_STORE_BF16) reinterpret_cast<__nv_bfloat16*>(o.out)[row] = __float2bfloat16_rn(v);
          else if (o.epi == LLMLB_EPI_RESID_F32) reinterpret_cast<float*>(o.out)[row] = __ldcg(reinterpret_cast<const float*>(o.out) + row) + v;
          else reinterpret_cast<float*>(o.out)[row] = v;
        }
        wbuf ^= 1;
      }
    }

    // ---- device-wide barrier before the next op reads what this one wrote ----
    if (oi + 1 < A.n_ops) {
      asm volatile("bar.sync 1, %0;" ::"n"(kChConsumerWarps * 32) : "memory");  // all epilogue stores issued
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(A.state + 1 + oi, 1u);
        const uint32_t target = epoch * gridDim.x;
        uint32_t spins = 0;
        while (ld_acquire_gpu(A.state + 1 + oi) < target) {
          if (++spins > (1u << 28)) { asm volatile("trap;"); }  // never hang the GPU on a lost CTA
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kChConsumerWarps * 32) : "memory");
    }
  }
  // last barrier counter of this launch done: bump the epoch (CTA 0 is itself past every barrier)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    __threadfence();
    A.state[0] = epoch;
  }
}

// rows per 32 KiB stage: largest power of two (<= 32) that fits
static uint32_t ch_rows_per_stage(uint32_t k) {
  uint32_t rps = kChStageBytes / (k * 2), p2 = 1;
  while (p2 * 2 <= rps && p2 * 2 <= (uint32_t)kChWindow) p2 *= 2;
  return p2;
}

bool gemv_chain_shape_ok(uint32_t n_out, uint32_t k) {
  if (k % (256 * kChConsumerWarps) || k * 2 > (uint32_t)kChStageBytes) return false;
  const uint32_t cpw = k / (256 * kChConsumerWarps);
  return cpw >= 1 && cpw <= (uint32_t)kChMaxCpw && n_out >= 2 * (uint32_t)kNumSMs;
}

struct ChainOpHost { const void* w; const void* x; const void* gain; void* out; uint32_t n_out, k, out_stride, epi; };

// Batch-1 chain of up to 4 GEMVs; state = device uint32[1 + kChMaxOps], zeroed once, private to
// the call site (its barrier counters count launches).
int gemv_chain_launch(const ChainOpHost* ops, uint32_t n_ops, float eps, uint32_t* state, cudaStream_t st) {
  if (n_ops == 0 || n_ops > (uint32_t)kChMaxOps || !state) { set_error("gemv_chain: bad op count"); return LLMLB_E_INVALID_ARG; }
  ChainArgs a{};
  a.n_ops = n_ops; a.eps = eps; a.state = state;
  for (uint32_t i = 0; i < n_ops; ++i) {
    if (!gemv_chain_shape_ok(ops[i].n_out, ops[i].k)) { set_error("gemv_chain: unsupported shape"); return LLMLB_E_UNSUPPORTED; }
    a.op[i] = ChainOp{(const __nv_bfloat16*)ops[i].w, ops[i].x, (const __nv_bfloat16*)ops[i].gain, ops[i].out,
                      ops[i].n_out, ops[i].k, ops[i].out_stride, ops[i].epi, ch_rows_per_stage(ops[i].k),
                      ops[i].k / (256 * kChConsumerWarps)};
  }
  constexpr int smem = kChStages * kChStageBytes + 128;
  static bool configured = false;
  if (!configured) {
    LLMLB_CUDA_CHECK(cudaFuncSetAttribute(gemv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    LLMLB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gemv_chain_kernel, kChThreads, smem));
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    if (per_sm < 1 || n_sm < kNumSMs) { set_error("gemv_chain: device cannot hold one CTA per SM x 148"); return LLMLB_E_UNSUPPORTED; }
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kNumSMs);
  cfg.blockDim = dim3(kChThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LLMLB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemv_chain_kernel, a));
  LLMLB_LAUNCH_CHECK();
  return LLMLB_OK;
}

}  // namespace llmlb
