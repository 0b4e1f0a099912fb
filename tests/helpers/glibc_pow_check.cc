// Host check of csrc/glibc_pow.h (the restatement the reference-order CUDA kernels use for std::pow) against this
// machine's libm: prints the number of arguments whose result differs in any bit.  TEST INFRASTRUCTURE.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "glibc_pow.h"

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 2000000;
  srand48(20261018);
  long bad = 0, skipped = 0;
  const double ys[4] = {3.0, 0.5, 2.5, -1.75};
  for (long i = 0; i < n; ++i) {
    const double x = exp(log(1e-12) + drand48() * log(1e24));
    volatile double y = ys[i & 3];   // volatile: keep the compiler from folding pow() into something else
    double got;
    if (!pda::glibcpow::powPositive(x, y, &got)) { ++skipped; continue; }
    if (got != pow(x, y)) ++bad;
  }
  for (long i = 0; i < n / 8; ++i) {   // around 1 (tiny |y log x|)
    const double x = 1.0 + (drand48() - 0.5) * 1e-3;
    volatile double y = 3.0;
    double got;
    if (!pda::glibcpow::powPositive(x, y, &got)) { ++skipped; continue; }
    if (got != pow(x, y)) ++bad;
  }
  double one;
  volatile double y3 = 3.0, x1 = 1.0;
  if (!pda::glibcpow::powPositive(x1, y3, &one) || one != 1.0) ++bad;
  printf("%ld %ld\n", bad, skipped);
  return 0;
}
