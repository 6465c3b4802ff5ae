/* naive_dft.c — O(N^2) long-double DFT used to pin oracle/dft_oracle.py (test infrastructure only).
 * Definition: X[k] = sum_n x[n] exp(sign * 2 pi i k n / N), unnormalised — the transform cuFFT (the
 * reference's engine, /root/reference/include/cufft.hpp:20-66) documents for CUFFT_FORWARD (sign -1)
 * and CUFFT_INVERSE (sign +1).  Batched over `howmany` lines with element strides. */
#include <math.h>
#include <stddef.h>

void naive_dft_lines(const double* in, double* out, size_t n, size_t howmany, size_t stride, size_t dist, int sign) {
    const long double tau = 6.283185307179586476925286766559005768L;
    for (size_t b = 0; b < howmany; ++b) {
        const double* x = in + 2 * b * dist;
        double* y = out + 2 * b * dist;
        for (size_t k = 0; k < n; ++k) {
            long double sr = 0, si = 0;
            for (size_t m = 0; m < n; ++m) {
                long double a = sign * tau * (long double)((k * m) % n) / (long double)n;
                long double c = cosl(a), s = sinl(a);
                long double xr = x[2 * m * stride], xi = x[2 * m * stride + 1];
                sr += xr * c - xi * s;
                si += xr * s + xi * c;
            }
            y[2 * k * stride] = (double)sr;
            y[2 * k * stride + 1] = (double)si;
        }
    }
}
