/* Oracle / CPU baseline (test infrastructure, not product): the decoder's linear layers on the
 * host cores with bf16 weights and fp32 accumulation — the CPU stand-in for the reference path's
 * external llama.cpp endpoint, used by bench.py's cpu_baseline / --impl reference
 * legs through oracle/llama_ref.py.  Plain C + OpenMP; gcc vectorises the inner loops.
 *   gcc -O3 -march=native -fopenmp -shared -fPIC oracle/llama_cpu.c -o oracle/liboracle_cpu.so
 *
 *   y[t, n] = sum_k x[t, k] * W[n, k]      W bf16 [n_out, k] row-major, x/y fp32, T <= 64     */
#include <immintrin.h>
#include <math.h>
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


/* ---- Q4_0: the weight format of the reference path's CPU endpoint in BASELINE.json configs[0] ("llama.cpp ... Llama-3-8B q4").
 * Restates ggml's published block format and reference quantiser (ggml-quants.c quantize_row_q4_0_ref /
 * dequantize_row_q4_0; llama.cpp is an external, un-vendored dependency of the reference: SURVEY.md §8c): a block is 32
 * weights = one fp16 scale d followed by 16 bytes of nibbles; w[j] = d * (q[j] - 8), low nibbles are elements 0..15, high
 * nibbles 16..31; d = (the value of largest magnitude, sign kept) / -8.  Pinned bit for bit to llama.cpp's own `gguf`
 * Python package in tests/test_oracle_q4.py.  The product engine computes in bf16; this exists only so that the CPU
 * baseline can also be quoted at the reference configuration's weight width (4.5 bits/weight streamed per token). */
static inline uint16_t f32_to_f16(float f) {          /* round to nearest even, like GGML_FP32_TO_FP16 (F16C) */
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  int32_t e = (int32_t)((x >> 23) & 0xFF) - 127 + 15;
  uint32_t m = x & 0x7FFFFFu;
  if (((x >> 23) & 0xFF) == 0xFF) return (uint16_t)(sign | 0x7C00u | (m ? 0x200u : 0));
  if (e >= 31) return (uint16_t)(sign | 0x7C00u);
  if (e <= 0) {
    if (e < -10) return (uint16_t)sign;
    m |= 0x800000u;
    const int shift = 14 - e;
    uint32_t h = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((uint32_t)e << 10) | (m >> 13);
  const uint32_t rem = m & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;   /* may carry into the exponent: still correct */
  return (uint16_t)(sign | h);
}
static inline float f16_to_f32(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 0x1F, m = h & 0x3FFu, x;
  if (e == 0) {
    if (m == 0) x = sign;
    else { e = 127 - 15 + 1; while (!(m & 0x400u)) { m <<= 1; --e; } x = sign | (e << 23) | ((m & 0x3FFu) << 13); }
  } else if (e == 31) x = sign | 0x7F800000u | (m << 13);
  else x = sign | ((e - 15 + 127) << 23) | (m << 13);
  float f;
  memcpy(&f, &x, 4);
  return f;
}

#define QK4_0 32
#define Q4_0_BLOCK_BYTES 18

/* W: bf16 bits [n_out, k] (k % 32 == 0) -> out: [n_out, k / 32] blocks of 18 bytes.
 * No FMA contraction here: x*id and +8.5 are two roundings in the reference quantiser (and in gguf-py, the pin). */
__attribute__((optimize("-ffp-contract=off")))
void oracle_quantize_q4_0(const uint16_t* W, uint8_t* out, int64_t n_out, int64_t k) {
  const int64_t nb = k / QK4_0;
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < n_out; ++n) {
    for (int64_t b = 0; b < nb; ++b) {
      float x[QK4_0], ax[QK4_0];
#pragma omp simd
      for (int j = 0; j < QK4_0; ++j) {
        x[j] = bf16_to_f32(W[n * k + b * QK4_0 + j]);
        ax[j] = fabsf(x[j]);
      }
      float amax = 0.f;
#pragma omp simd reduction(max : amax)
      for (int j = 0; j < QK4_0; ++j) amax = ax[j] > amax ? ax[j] : amax;
      float max = 0.f;                                  /* the FIRST element of largest magnitude, sign kept */
      for (int j = 0; j < QK4_0; ++j) if (ax[j] == amax) { max = x[j]; break; }
      const float d = max / -8.f;
      const float id = d ? 1.0f / d : 0.0f;
      uint8_t* blk = out + (n * nb + b) * Q4_0_BLOCK_BYTES;
      const uint16_t dh = f32_to_f16(d);
      memcpy(blk, &dh, 2);
      int32_t q[QK4_0];
#pragma omp simd
      for (int j = 0; j < QK4_0; ++j) {
        const float v = x[j] * id;                      /* two roundings (no FMA): as the reference quantiser */
        const float u = v + 8.5f;
        int32_t qi = (int32_t)u;                        /* u >= 0.5 by construction: truncation = the (int8_t) cast */
        q[j] = qi > 15 ? 15 : qi;
      }
      for (int j = 0; j < QK4_0 / 2; ++j) blk[2 + j] = (uint8_t)(q[j] | (q[QK4_0 / 2 + j] << 4));
    }
  }
}

/* y[t, n] = sum over blocks of d * sum_j (q_j - 8) * x[t, j]   (fp32 activations; llama.cpp itself quantises the
 * activations to Q8_0 and runs integer dot products — same bytes per token, different arithmetic: a port, not llama.cpp) */
void oracle_linear_q4_0(const uint8_t* Wq, const float* X, float* Y, int64_t n_out, int64_t k, int64_t T) {
  const int64_t nb = k / QK4_0;
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < n_out; ++n) {
    const uint8_t* row = Wq + n * nb * Q4_0_BLOCK_BYTES;
    for (int64_t t = 0; t < T; ++t) {
      const float* x = X + t * k;
      float acc = 0.f;
      for (int64_t b = 0; b < nb; ++b) {
        const uint8_t* blk = row + b * Q4_0_BLOCK_BYTES;
        uint16_t dh;
        memcpy(&dh, blk, 2);
        const float d = f16_to_f32(dh);
        const float* xb = x + b * QK4_0;
        float w[QK4_0];
#pragma omp simd
        for (int j = 0; j < QK4_0 / 2; ++j) {
          w[j] = (float)((int)(blk[2 + j] & 0x0F) - 8);
          w[QK4_0 / 2 + j] = (float)((int)(blk[2 + j] >> 4) - 8);
        }
        float s = 0.f;
#pragma omp simd reduction(+ : s)
        for (int j = 0; j < QK4_0; ++j) s += w[j] * xb[j];
        acc += d * s;
      }
      Y[t * n_out + n] = acc;
    }
  }
}


/* ---- Q4_0 x Q8_0: how llama.cpp's CPU backend actually multiplies a Q4_0 weight: the ACTIVATIONS are quantised to Q8_0
 * (block of 32 = fp16 scale + 32 int8, d = amax / 127, q = roundf(x / d); ggml-quants.c quantize_row_q8_0_ref) and the dot
 * product runs on integers, one (d4 * d8) scaling per block (ggml_vec_dot_q4_0_q8_0).  The AVX2 body below is that
 * published scheme (unsigned x signed bytes through abs/sign + maddubs + madd), restated; the quantiser is pinned bit for
 * bit to gguf-py, the product to a numpy evaluation of the same integers (tests/test_oracle_q4.py). */
#define Q8_0_BLOCK_BYTES 34

__attribute__((optimize("-ffp-contract=off")))
void oracle_quantize_q8_0(const float* X, uint8_t* out, int64_t T, int64_t k) {
  const int64_t nb = k / QK4_0;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < T * nb; ++i) {
    const float* x = X + i * QK4_0;
    float amax = 0.f;
    for (int j = 0; j < QK4_0; ++j) { const float a = fabsf(x[j]); if (a > amax) amax = a; }
    const float d = amax / 127.f;
    const float id = d ? 1.0f / d : 0.0f;
    uint8_t* blk = out + i * Q8_0_BLOCK_BYTES;
    const uint16_t dh = f32_to_f16(d);
    memcpy(blk, &dh, 2);
    for (int j = 0; j < QK4_0; ++j) ((int8_t*)(blk + 2))[j] = (int8_t)roundf(x[j] * id);
  }
}

#if defined(__AVX2__)
static inline float hsum8(__m256 v) {
  __m128 lo = _mm256_castps256_ps128(v), hi = _mm256_extractf128_ps(v, 1);
  lo = _mm_add_ps(lo, hi);
  lo = _mm_add_ps(lo, _mm_movehl_ps(lo, lo));
  lo = _mm_add_ss(lo, _mm_movehdup_ps(lo));
  return _mm_cvtss_f32(lo);
}
#endif

/* y[t, n] = sum_b d4[n,b] * d8[t,b] * sum_j (q4[n,b,j] - 8) * q8[t,b,j]      Xq: [T, k/32] Q8_0 blocks */
void oracle_linear_q4_0_q8_0(const uint8_t* Wq, const uint8_t* Xq, float* Y, int64_t n_out, int64_t k, int64_t T) {
  const int64_t nb = k / QK4_0;
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < n_out; ++n) {
    const uint8_t* row = Wq + n * nb * Q4_0_BLOCK_BYTES;
    for (int64_t t = 0; t < T; ++t) {
      const uint8_t* xr = Xq + t * nb * Q8_0_BLOCK_BYTES;
#if defined(__AVX2__)
      __m256 acc = _mm256_setzero_ps();
      const __m256i m4 = _mm256_set1_epi8(0x0F), off = _mm256_set1_epi8(8), ones = _mm256_set1_epi16(1);
      for (int64_t b = 0; b < nb; ++b) {
        const uint8_t* wb = row + b * Q4_0_BLOCK_BYTES;
        const uint8_t* xb = xr + b * Q8_0_BLOCK_BYTES;
        uint16_t dw, dx;
        memcpy(&dw, wb, 2);
        memcpy(&dx, xb, 2);
#if defined(__F16C__)
        const __m256 d = _mm256_set1_ps(_cvtsh_ss(dw) * _cvtsh_ss(dx));
#else
        const __m256 d = _mm256_set1_ps(f16_to_f32(dw) * f16_to_f32(dx));
#endif
        const __m128i nib = _mm_loadu_si128((const __m128i*)(wb + 2));
        __m256i q4 = _mm256_set_m128i(_mm_srli_epi16(nib, 4), nib);          /* low nibbles = elements 0..15, high = 16..31 */
        q4 = _mm256_sub_epi8(_mm256_and_si256(q4, m4), off);                   /* -8 .. 7 */
        const __m256i q8 = _mm256_loadu_si256((const __m256i*)(xb + 2));
        const __m256i ax = _mm256_sign_epi8(q4, q4), sy = _mm256_sign_epi8(q8, q4);
        const __m256i p16 = _mm256_maddubs_epi16(ax, sy);
        const __m256i p32 = _mm256_madd_epi16(p16, ones);
        acc = _mm256_fmadd_ps(d, _mm256_cvtepi32_ps(p32), acc);
      }
      Y[t * n_out + n] = hsum8(acc);
#else
      float acc = 0.f;
      for (int64_t b = 0; b < nb; ++b) {
        const uint8_t* wb = row + b * Q4_0_BLOCK_BYTES;
        const uint8_t* xb = xr + b * Q8_0_BLOCK_BYTES;
        uint16_t dw, dx;
        memcpy(&dw, wb, 2);
        memcpy(&dx, xb, 2);
        int32_t s = 0;
        for (int j = 0; j < QK4_0 / 2; ++j) {
          s += ((int)(wb[2 + j] & 0x0F) - 8) * (int)((const int8_t*)(xb + 2))[j];
          s += ((int)(wb[2 + j] >> 4) - 8) * (int)((const int8_t*)(xb + 2))[QK4_0 / 2 + j];
        }
        acc += f16_to_f32(dw) * f16_to_f32(dx) * (float)s;
      }
      Y[t * n_out + n] = acc;
#endif
    }
  }
}
