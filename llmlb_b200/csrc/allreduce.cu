// Tensor-parallel exchange over NVLink peer memory.
//
// Every rank owns one CUDA-IPC-shared exchange region:
//   [ signal block | slot 0 | slot 1 ]      (slots alternate per collective: double buffering)
// A producer kernel (the O / down projection) writes its fp32 partial [T, hidden] into the local
// slot; allreduce_add_kernel then (1) raises this rank's flag in every peer's signal block,
// (2) waits until every peer raised theirs, (3) reads the T*hidden slice it owns from all peers
// in rank order (bit-identical sums on every rank keep the replicated residual stream identical)
// and adds it into the fp32 residual.  Start-barrier only: a slot is reused two collectives
// later, and no rank can get two collectives ahead of a peer.
// allgather_kernel uses the same handshake to assemble vocab-sharded logits.
#include "../../include/llmlb_b200.h"
#include "common.cuh"

namespace llmlb {

constexpr int kArMaxRanks = 8;
constexpr int kArMaxBlocks = 148;

struct ArSignals {  // lives at the start of every rank's exchange region
  // flag[block][src_rank]: epoch value written by src_rank's block `block`
  uint32_t flag[kArMaxBlocks][kArMaxRanks];
  uint32_t epoch[kArMaxBlocks];  // local only: last epoch this block used
};

constexpr size_t kArSignalBytes = (sizeof(ArSignals) + 255) & ~size_t(255);

struct ArPeers {
  uint8_t* base[kArMaxRanks];  // exchange region of every rank, mapped into this process
  uint32_t rank, size;
  uint64_t slot_bytes;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// cross-GPU barrier for block blockIdx.x; returns the epoch used
__device__ __forceinline__ void ar_barrier(const ArPeers& P) {
  ArSignals* mine = reinterpret_cast<ArSignals*>(P.base[P.rank]);
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) {
    s_epoch = mine->epoch[blockIdx.x] + 1;
    mine->epoch[blockIdx.x] = s_epoch;
  }
  __syncthreads();
  const uint32_t ep = s_epoch;
  if (threadIdx.x < P.size) {
    ArSignals* peer = reinterpret_cast<ArSignals*>(P.base[threadIdx.x]);
    st_release_sys(&peer->flag[blockIdx.x][P.rank], ep);
    while (ld_acquire_sys(&mine->flag[blockIdx.x][threadIdx.x]) < ep) {
    }
  }
  __syncthreads();
}

// x[i] += sum_r slot_r[i]   for i in [0, n)   (n multiple of 4)
__global__ void __launch_bounds__(512)
allreduce_add_kernel(ArPeers P, uint32_t slot, float* __restrict__ x, uint64_t n) {
  ar_barrier(P);
  const uint64_t n4 = n / 4;
  const size_t off = kArSignalBytes + size_t(slot) * P.slot_bytes;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += uint64_t(gridDim.x) * blockDim.x) {
    float4 acc = reinterpret_cast<float4*>(x)[i];
#pragma unroll
    for (int r = 0; r < kArMaxRanks; ++r) {
      if (r < int(P.size)) {
        float4 v = ld_peer_f4(reinterpret_cast<const float4*>(P.base[r] + off) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    reinterpret_cast<float4*>(x)[i] = acc;
  }
}

// out[row, r*cols_local + c] = slot_r[row, c]   (gathers vocab-sharded logits on every rank)
__global__ void __launch_bounds__(512)
allgather_cols_kernel(ArPeers P, uint32_t slot, float* __restrict__ out, uint32_t rows,
                      uint32_t cols_local) {
  ar_barrier(P);
  const size_t off = kArSignalBytes + size_t(slot) * P.slot_bytes;
  const uint32_t c4 = cols_local / 4;
  const uint64_t total = uint64_t(rows) * P.size * c4;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += uint64_t(gridDim.x) * blockDim.x) {
    uint32_t c = uint32_t(i % c4);
    uint64_t t = i / c4;
    uint32_t r = uint32_t(t % P.size);
    uint32_t row = uint32_t(t / P.size);
    float4 v = ld_peer_f4(reinterpret_cast<const float4*>(P.base[r] + off) + size_t(row) * c4 + c);
    reinterpret_cast<float4*>(out + (size_t(row) * P.size + r) * cols_local)[c] = v;
  }
}

size_t ar_signal_bytes() { return kArSignalBytes; }

int ar_allreduce_add(const ArPeers& P, uint32_t slot, float* x, uint64_t n, cudaStream_t st) {
  uint32_t blocks = (uint32_t)((n / 4 + 511) / 512);
  if (blocks > 64) blocks = 64;
  if (blocks < 1) blocks = 1;
  allreduce_add_kernel<<<blocks, 512, 0, st>>>(P, slot, x, n);
  LLMLB_LAUNCH_CHECK();
  return LLMLB_OK;
}
int ar_allgather_cols(const ArPeers& P, uint32_t slot, float* out, uint32_t rows,
                      uint32_t cols_local, cudaStream_t st) {
  uint64_t total = uint64_t(rows) * P.size * (cols_local / 4);
  uint32_t blocks = (uint32_t)((total + 511) / 512);
  if (blocks > 64) blocks = 64;
  if (blocks < 1) blocks = 1;
  allgather_cols_kernel<<<blocks, 512, 0, st>>>(P, slot, out, rows, cols_local);
  LLMLB_LAUNCH_CHECK();
  return LLMLB_OK;
}

}  // namespace llmlb
