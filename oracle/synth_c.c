/* Oracle (test infrastructure, not product): plain-C restatement of the synthetic-weight
 * generator, used to materialise full-size (8B) weights on the host for the CPU baseline and the
 * --impl reference arm.  Must agree bit for bit with oracle/synth.py (checked in
 * tests/test_oracle_llama.py) and with llmlb_b200/csrc/common.cuh synth_value (checked on the GPU
 * in tests/test_ops_gpu.py::test_synth_bitexact).
 *   gcc -O3 -fopenmp -shared -fPIC oracle/synth_c.c -o oracle/liboracle_synth.so              */
#include <stdint.h>
#include <string.h>

static inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static inline uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

/* out[r*cols + c] = bf16(synth(seed, tensor_id, (row0+r)*ld + col0+c)) */
void oracle_synth_bf16(uint16_t* out, uint64_t rows, uint64_t cols, uint64_t row0, uint64_t col0,
                       uint64_t ld, uint64_t seed, uint32_t tensor_id, float std) {
  const uint64_t base = mix64(seed * 0xD1342543DE82EF95ull + tensor_id);
  const float scale = std / 37837.22f;
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < (int64_t)rows; ++r) {
    uint16_t* o = out + (uint64_t)r * cols;
    const uint64_t rb = (row0 + (uint64_t)r) * ld + col0;
    for (uint64_t c = 0; c < cols; ++c) {
      uint64_t h = mix64(base + rb + c);
      int s = (int)(h & 0xFFFF) + (int)((h >> 16) & 0xFFFF) + (int)((h >> 32) & 0xFFFF) +
              (int)((h >> 48) & 0xFFFF) - 131070;
      o[c] = f32_to_bf16_rne((float)s * scale);
    }
  }
}
