/* TEST INFRASTRUCTURE - CPU statement of the grid adder (a "next" row).
 *
 * Parity status: UNPINNED.  ska-sdp-idg-bench has no adder: only the unused idg::Grid type
 * (app/common/types.hpp:358-370, complex<float>[w][pol][y][x]) and the subgrid coordinate
 * (types.hpp:11-13; app/common/init.cpp:182-199 turns it into the offsets the gridder uses, so
 * coordinate.(x, y) is the grid position of the subgrid's first pixel).  The adder is the step
 * of Image Domain Gridding (van der Tol, Veenboer, Offringa 2018, sec. 3.3) that follows the
 * gridder: every subgrid is accumulated into the grid at its coordinate.  The bench does no
 * subgrid FFT and no FFT shift (gridder_reference.cpp:105-109 stores unshifted pixels), so none is
 * applied here either; pixels that fall outside the grid are dropped.  There is nothing in the
 * reference to check this file against: the GPU adder is tested against it, and it against a
 * two-line numpy statement of the same sum (tests/test_adder.py).
 */
#include "idg_oracle.h"

/* grid: complex64 [4][grid_size][grid_size] (interleaved re, im), accumulated into */
void idgo_adder(int nr_subgrids, int grid_size, int subgrid_size, const idgo_metadata *metadata,
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
          const float *src = subgrids + 2 * (((long)s * 4 + pol) * N * N + y * N + x);
          float *dst = grid + 2 * ((pol * G + Y) * G + X);
          dst[0] += src[0];
          dst[1] += src[1];
        }
      }
  }
}
