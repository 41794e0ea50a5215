/* TEST INFRASTRUCTURE — CPU restatement of the reference's IDG gridder/degridder.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker or the reported
 * CPU baseline.  The product (ska_sdp_idg_bench_b200/) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py compares every
 * function below bit for bit with the reference's own code compiled from
 * /root/reference (oracle/_ref/libidgref.so), and tests/golden/ holds outputs of
 * that reference build for the reference's correctness shape.
 */
#ifndef IDG_ORACLE_H_
#define IDG_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* app/common/types.hpp:11-26 — 36-byte work descriptor of one subgrid */
typedef struct {
  int baseline_offset;
  int time_offset;
  int nr_timesteps;
  int aterm_index;
  unsigned int station1, station2; /* Baseline, types.hpp:15-17 */
  int x, y, z;                     /* Coordinate, types.hpp:11-13 */
} idgo_metadata;

/* app/common/types.hpp:46-50 */
typedef struct {
  float u, v, w;
} idgo_uvw;

/* app/common/types.hpp:15-17 */
typedef struct {
  unsigned int station1, station2;
} idgo_baseline;

int idgo_max_threads(void);
void idgo_set_threads(int n);

/* app/CPU/kernels/gridder_reference.cpp:6-114.  Complex arrays are passed as
 * interleaved (re, im) floats.
 *   visibilities [total_timesteps][nr_channels][4]      (types.hpp:28-44)
 *   aterms       [slots][nr_stations][N][N][4]
 *   subgrids     [nr_subgrids][4][N][N]                 (output) */
void idgo_gridder(int nr_subgrids, int grid_size, int subgrid_size,
                  float image_size, float w_step_in_lambda, int nr_channels,
                  int nr_stations, const idgo_uvw *uvw,
                  const float *wavenumbers, const float *visibilities,
                  const float *spheroidal, const float *aterms,
                  const idgo_metadata *metadata, float *subgrids);

/* app/CPU/kernels/degridder_reference.cpp:6-129 (visibilities are the output) */
void idgo_degridder(int nr_subgrids, int grid_size, int subgrid_size,
                    float image_size, float w_step_in_lambda, int nr_channels,
                    int nr_stations, const idgo_uvw *uvw,
                    const float *wavenumbers, float *visibilities,
                    const float *spheroidal, const float *aterms,
                    const idgo_metadata *metadata, const float *subgrids);

/* "Next" rows 8f-1..3, NOT part of the pinned path (idg_next_oracle.c: parity UNPINNED - the
 * reference has no adder, FFT or splitter).  grid is complex64 [4][grid_size][grid_size].
 *   adder:    grid[pol][y0 + y][x0 + x] += subgrids[s][pol][y'][x'], clipped at the grid edge
 *   splitter: subgrids[s][pol][y'][x']   = grid[pol][y0 + y][x0 + x], 0 outside the grid
 *   (y', x') = (y, x), or ((y + N/2) % N, (x + N/2) % N) with IDGO_FFT_SHIFT
 *   subgrid_fft: in-place 2-D DFT of nr_planes N x N planes, direction >= 0 forward (exp(-i),
 *   unscaled), < 0 backward (exp(+i), scaled by 1/N^2); double-precision sums rounded once. */
#define IDGO_FFT_SHIFT 1
void idgo_adder(int nr_subgrids, int grid_size, int subgrid_size, int flags, const idgo_metadata *metadata,
                const float *subgrids, float *grid);
void idgo_splitter(int nr_subgrids, int grid_size, int subgrid_size, int flags, const idgo_metadata *metadata,
                   float *subgrids, const float *grid);
void idgo_subgrid_fft(long nr_planes, int subgrid_size, int direction, float *planes);

/* Same formulas evaluated in float64 throughout (phase, sincos, sums) from the
 * same float32 inputs: the "truth" used to budget the float32 error.  Outputs
 * are interleaved doubles. */
void idgo_gridder_f64(int nr_subgrids, int grid_size, int subgrid_size,
                      float image_size, float w_step_in_lambda, int nr_channels,
                      int nr_stations, const idgo_uvw *uvw,
                      const float *wavenumbers, const float *visibilities,
                      const float *spheroidal, const float *aterms,
                      const idgo_metadata *metadata, double *subgrids);

void idgo_degridder_f64(int nr_subgrids, int grid_size, int subgrid_size,
                        float image_size, float w_step_in_lambda,
                        int nr_channels, int nr_stations, const idgo_uvw *uvw,
                        const float *wavenumbers, double *visibilities,
                        const float *spheroidal, const float *aterms,
                        const idgo_metadata *metadata, const float *subgrids);

/* Synthetic inputs, app/common/init.cpp.  rand() state is glibc's; call
 * idgo_srand(0) then uvw -> aterms -> metadata as the reference's test mains do
 * (tests/gridder_common.cpp:88-101). */
void idgo_srand(unsigned seed);
void idgo_init_uvw(unsigned grid_size, int nr_baselines, int nr_timesteps,
                   idgo_uvw *uvw);                                  /* init.cpp:4-25 */
void idgo_init_frequencies(int nr_channels, float *frequencies);    /* :27-36 */
void idgo_init_wavenumbers(int nr_channels, const float *frequencies,
                           float *wavenumbers);                     /* :38-46 */
void idgo_init_visibilities(unsigned grid_size, float image_size,
                            int nr_baselines, int nr_timesteps, int nr_channels,
                            const float *frequencies, const idgo_uvw *uvw,
                            float *visibilities);                   /* :48-79 */
void idgo_init_baselines(unsigned nr_stations, int nr_baselines,
                         idgo_baseline *baselines);                 /* :81-95 */
void idgo_init_spheroidal(int subgrid_size, float *spheroidal);     /* :97-107 */
void idgo_init_aterms(int nr_timeslots, int nr_stations, int subgrid_size,
                      const float *spheroidal, float *aterms);      /* :109-132 */
void idgo_init_metadata(unsigned grid_size, unsigned nr_timeslots,
                        unsigned nr_timesteps_subgrid, int nr_baselines,
                        const idgo_baseline *baselines,
                        idgo_metadata *metadata);                   /* :134-159 */
void idgo_init_subgrids(int nr_subgrids, int subgrid_size,
                        float *subgrids);                           /* :161-180 */

/* Metric model, app/common/common.cpp:100-159 */
uint64_t idgo_flops_gridder(uint64_t nr_channels, uint64_t nr_timesteps,
                            uint64_t nr_subgrids, uint64_t subgrid_size,
                            uint64_t nr_correlations);
uint64_t idgo_bytes_gridder(uint64_t nr_channels, uint64_t nr_timesteps,
                            uint64_t nr_subgrids, uint64_t subgrid_size,
                            uint64_t nr_correlations);

/* The reference's pass/fail number, tests/test_util.hpp:28-92 (A = candidate,
 * B = reference, n complex elements).  Returns mean_error. */
double idgo_check_error(int n, const float *A, const float *B);

#ifdef __cplusplus
}
#endif
#endif
