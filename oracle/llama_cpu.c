/* Oracle / CPU baseline (test infrastructure, not product): the decoder's linear layers on the
 * host cores with bf16 weights and fp32 accumulation — the CPU stand-in for the reference path's
 * external llama.cpp endpoint, used by bench.py's cpu_baseline / --impl reference
 * legs through oracle/llama_ref.py.  Plain C + OpenMP; gcc vectorises the inner loops.
 *   gcc -O3 -march=native -fopenmp -shared -fPIC oracle/llama_cpu.c -o oracle/liboracle_cpu.so
 *
 *   y[t, n] = sum_k x[t, k] * W[n, k]      W bf16 [n_out, k] row-major, x/y fp32, T <= 64     */
#include <omp.h>
#include <stdint.h>
#include <string.h>

#define TB 8 /* tokens per register block */

static inline float bf16_to_f32(uint16_t v) {
  uint32_t u = (uint32_t)v << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

void oracle_linear_bf16(const uint16_t* W, const float* X, float* Y, int64_t n_out, int64_t k,
                        int64_t T) {
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < n_out; ++n) {
    const uint16_t* w = W + n * k;
    for (int64_t t0 = 0; t0 < T; t0 += TB) {
      const int64_t tb = (T - t0 < TB) ? (T - t0) : TB;
      float acc[TB] = {0};
      if (tb == 1) {
        float a = 0.f;
        const float* x = X + t0 * k;
#pragma omp simd reduction(+ : a)
        for (int64_t i = 0; i < k; ++i) a += bf16_to_f32(w[i]) * x[i];
        acc[0] = a;
      } else {
        for (int64_t t = 0; t < tb; ++t) {
          float a = 0.f;
          const float* x = X + (t0 + t) * k;
#pragma omp simd reduction(+ : a)
          for (int64_t i = 0; i < k; ++i) a += bf16_to_f32(w[i]) * x[i];
          acc[t] = a;
        }
      }
      for (int64_t t = 0; t < tb; ++t) Y[(t0 + t) * n_out + n] = acc[t];
    }
  }
}
