// TEST INFRASTRUCTURE: the product's fused VBEM transform (salmon_b200/csrc/em_math.h) compiled for the host, so that
// tests/test_em_math.py can compare it with the oracle's digamma without a GPU.
#include <stdint.h>
#include "../salmon_b200/csrc/em_math.h"
extern "C" void hem_exp_digamma(uint64_t n, const double* x, const double* logNorm, double* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = sb::exp_digamma_shifted(x[i], logNorm[i]);
}
