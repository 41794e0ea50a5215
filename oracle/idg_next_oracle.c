/* TEST INFRASTRUCTURE - CPU statements of the "next" rows of SURVEY.md 8f: grid adder (8f-1),
 * subgrid FFT (8f-2) and splitter (8f-3).  Only tests/, __graft_entry__.smoke() and bench.py's
 * checks may call this file.
 *
 * Parity status: UNPINNED.  ska-sdp-idg-bench has none of these steps: only the unused idg::Grid
 * type (app/common/types.hpp:358-370, complex<float>[w][pol][y][x]) and the subgrid coordinate
 * (types.hpp:11-13; app/common/init.cpp:182-199 turns it into the offsets the gridder uses, so
 * coordinate.(x, y) is the grid position of the subgrid's first pixel).  They are the steps of
 * Image Domain Gridding (van der Tol, Veenboer, Offringa 2018, sec. 3.3) around the two kernels the
 * bench has: gridder -> subgrid FFT -> adder -> [grid] -> splitter -> inverse subgrid FFT ->
 * degridder.  There is nothing in the reference to check this file against: the GPU kernels are
 * tested against it, and it against numpy statements of the same sums (tests/test_next_rows.py).
 *
 * The bench's gridder stores unshifted pixels (gridder_reference.cpp:105-109).  IDGO_FFT_SHIFT
 * states the index shift the full pipeline applies between subgrid and grid: grid offset (y, x)
 * of a subgrid pairs with subgrid pixel ((y + N/2) mod N, (x + N/2) mod N).
 */
#include <math.h>
#include <stdlib.h>

#include "idg_oracle.h"

static long shifted(long i, long N, int flags) { return (flags & IDGO_FFT_SHIFT) ? (i + N / 2) % N : i; }

/* grid: complex64 [4][grid_size][grid_size] (interleaved re, im), accumulated into;
 * pixels that fall outside the grid are dropped */
void idgo_adder(int nr_subgrids, int grid_size, int subgrid_size, int flags, const idgo_metadata *metadata,
                const float *subgrids, float *grid) {
  const long G = grid_size, N = subgrid_size;
  for (int s = 0; s < nr_subgrids; s++) {   /* sequential: a fixed summation order */
    const long x0 = metadata[s].x, y0 = metadata[s].y;
    for (int pol = 0; pol < 4; pol++)
      for (long y = 0; y < N; y++) {
        const long Y = y0 + y;
        if (Y < 0 || Y >= G) continue;
        for (long x = 0; x < N; x++) {
          const long X = x0 + x;
          if (X < 0 || X >= G) continue;
          const float *src = subgrids + 2 * (((long)s * 4 + pol) * N * N + shifted(y, N, flags) * N +
                                             shifted(x, N, flags));
          float *dst = grid + 2 * ((pol * G + Y) * G + X);
          dst[0] += src[0];
          dst[1] += src[1];
        }
      }
  }
}

/* the adder's inverse: every subgrid pixel is read from the grid, 0 where the subgrid overhangs it */
void idgo_splitter(int nr_subgrids, int grid_size, int subgrid_size, int flags, const idgo_metadata *metadata,
                   float *subgrids, const float *grid) {
  const long G = grid_size, N = subgrid_size;
  for (int s = 0; s < nr_subgrids; s++) {
    const long x0 = metadata[s].x, y0 = metadata[s].y;
    for (int pol = 0; pol < 4; pol++)
      for (long y = 0; y < N; y++)
        for (long x = 0; x < N; x++) {
          const long Y = y0 + y, X = x0 + x;
          float *dst = subgrids + 2 * (((long)s * 4 + pol) * N * N + shifted(y, N, flags) * N +
                                       shifted(x, N, flags));
          if (Y < 0 || Y >= G || X < 0 || X >= G) {
            dst[0] = dst[1] = 0.f;
          } else {
            const float *src = grid + 2 * ((pol * G + Y) * G + X);
            dst[0] = src[0];
            dst[1] = src[1];
          }
        }
  }
}

/* In-place 2-D DFT of nr_planes planes of N x N complex64 (= nr_subgrids * 4 planes of a subgrid
 * array), evaluated as the two separable sums in double precision and rounded once:
 *   direction >= 0:  B[ky][kx] =        sum A[y][x] exp(-2 pi i (ky y + kx x) / N)
 *   direction <  0:  B[ky][kx] = 1/N^2  sum A[y][x] exp(+2 pi i (ky y + kx x) / N)       */
void idgo_subgrid_fft(long nr_planes, int subgrid_size, int direction, float *planes) {
  const long N = subgrid_size;
  const double sign = direction >= 0 ? -1.0 : 1.0;
  const double scale = direction >= 0 ? 1.0 : 1.0 / ((double)N * (double)N);
  double *wr = malloc(sizeof(double) * N), *wi = malloc(sizeof(double) * N);
  for (long k = 0; k < N; k++) {
    wr[k] = cos(2.0 * M_PI * (double)k / (double)N);
    wi[k] = sign * sin(2.0 * M_PI * (double)k / (double)N);
  }
#pragma omp parallel
  {
    double *t = malloc(sizeof(double) * 2 * N * N);
#pragma omp for
    for (long p = 0; p < nr_planes; p++) {
      float *a = planes + 2 * p * N * N;
      for (long ky = 0; ky < N; ky++)        /* over y */
        for (long x = 0; x < N; x++) {
          double sr = 0, si = 0;
          for (long y = 0; y < N; y++) {
            const long e = (ky * y) % N;
            const double ar = a[2 * (y * N + x)], ai = a[2 * (y * N + x) + 1];
            sr += ar * wr[e] - ai * wi[e];
            si += ar * wi[e] + ai * wr[e];
          }
          t[2 * (ky * N + x)] = sr;
          t[2 * (ky * N + x) + 1] = si;
        }
      for (long ky = 0; ky < N; ky++)        /* over x */
        for (long kx = 0; kx < N; kx++) {
          double sr = 0, si = 0;
          for (long x = 0; x < N; x++) {
            const long e = (kx * x) % N;
            const double ar = t[2 * (ky * N + x)], ai = t[2 * (ky * N + x) + 1];
            sr += ar * wr[e] - ai * wi[e];
            si += ar * wi[e] + ai * wr[e];
          }
          a[2 * (ky * N + kx)] = (float)(sr * scale);
          a[2 * (ky * N + kx) + 1] = (float)(si * scale);
        }
    }
    free(t);
  }
  free(wr);
  free(wi);
}
