/* The IDG imaging cycle through the C ABI, from plain C: gridder -> subgrid FFT -> adder -> [grid]
 * -> splitter -> inverse FFT -> degridder on device-resident synthetic data (the library's own
 * generators, i.e. the distributions of the reference's app/common/init.cpp).
 *
 *   gcc -O2 -I include examples/imaging_cycle.c -L ska_sdp_idg_bench_b200 -lidgb200 \
 *       -Wl,-rpath,$PWD/ska_sdp_idg_bench_b200 -L/usr/local/cuda/lib64 -lcudart -o imaging_cycle
 *   ./imaging_cycle [nr_stations] [nr_timeslots]
 *
 * Checks (exit code 1 on failure): inverse FFT of the FFT returns the gridder's subgrids; the
 * visibilities degridded from the untouched subgrids are finite and not all zero.  tests/ builds
 * and runs it (test_c_example_*). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "idg_b200.h"

/* cudart, declared by hand so that the example needs no CUDA headers */
extern int cudaMalloc(void **p, size_t n);
extern int cudaFree(void *p);
extern int cudaMemset(void *p, int v, size_t n);
extern int cudaMemcpy(void *dst, const void *src, size_t n, int kind);
extern int cudaDeviceSynchronize(void);
enum { D2H = 2, D2D = 3 };

#define CHECK(x)                                                                            \
  do {                                                                                      \
    int rc_ = (x);                                                                          \
    if (rc_ != 0) {                                                                         \
      fprintf(stderr, "%s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #x, rc_, idgb200_error_string(rc_)); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

int main(int argc, char **argv) {
  const int nr_stations = argc > 1 ? atoi(argv[1]) : 8, nr_timeslots = argc > 2 ? atoi(argv[2]) : 2;
  const int G = 1024, N = 32, T = 128, C = 16;
  const float image_size = 0.01f;
  const int nr_baselines = nr_stations * (nr_stations - 1) / 2;
  const int S = nr_baselines * nr_timeslots;
  const long tt = (long)S * T;
  const size_t npix = (size_t)N * N, sg_elems = (size_t)S * 4 * npix;

  idgb200_uvw *uvw;
  float *wn, *sph;
  idgb200_cfloat *vis, *vis2, *at, *sg, *sg0, *grid;
  idgb200_metadata *meta;
  CHECK(cudaMalloc((void **)&uvw, tt * sizeof *uvw));
  CHECK(cudaMalloc((void **)&wn, C * sizeof *wn));
  CHECK(cudaMalloc((void **)&sph, npix * sizeof *sph));
  CHECK(cudaMalloc((void **)&vis, (size_t)tt * C * 4 * sizeof *vis));
  CHECK(cudaMalloc((void **)&vis2, (size_t)tt * C * 4 * sizeof *vis2));
  CHECK(cudaMalloc((void **)&at, (size_t)nr_timeslots * nr_stations * npix * 4 * sizeof *at));
  CHECK(cudaMalloc((void **)&sg, sg_elems * sizeof *sg));
  CHECK(cudaMalloc((void **)&sg0, sg_elems * sizeof *sg0));
  CHECK(cudaMalloc((void **)&grid, (size_t)4 * G * G * sizeof *grid));
  CHECK(cudaMalloc((void **)&meta, (size_t)S * sizeof *meta));

  CHECK(idgb200_init_uvw(G, S, T, 1, uvw, NULL));
  CHECK(idgb200_init_wavenumbers(C, wn, NULL));
  CHECK(idgb200_init_visibilities(G, image_size, tt, C, uvw, vis, NULL));
  CHECK(idgb200_init_spheroidal(N, sph, NULL));
  CHECK(idgb200_init_aterms(nr_timeslots, nr_stations, N, 2, at, NULL));
  CHECK(idgb200_init_metadata(G, nr_stations, nr_timeslots, T, 1, 3, meta, NULL));   /* per-slot A-terms, seed 3 */

  idgb200_params p;
  memset(&p, 0, sizeof p);
  p.nr_subgrids = S; p.grid_size = G; p.subgrid_size = N; p.image_size = image_size;
  p.w_step_in_lambda = 0.f; p.nr_channels = C; p.nr_stations = nr_stations;
  p.sincos_mode = IDGB200_SINCOS_FAST; p.variant = 0; p.flags = IDGB200_FLAG_FFT_SHIFT;

  idgb200_cfloat *parts[1] = {grid};
  CHECK(cudaMemset(grid, 0, (size_t)4 * G * G * sizeof *grid));
  CHECK(idgb200_gridder(&p, uvw, wn, vis, sph, at, meta, sg, NULL));
  CHECK(cudaMemcpy(sg0, sg, sg_elems * sizeof *sg, D2D));
  CHECK(idgb200_subgrid_fft(S, N, +1, sg, NULL));
  CHECK(idgb200_adder(&p, meta, sg, parts, 1, G, NULL));
  /* back: the subgrids overlap on the grid, so the splitter does not return the adder's input; the
   * inverse FFT is checked on the forward FFT's own output instead */
  CHECK(idgb200_subgrid_fft(S, N, -1, sg, NULL));
  CHECK(cudaDeviceSynchronize());

  idgb200_cfloat *h_a = malloc(sg_elems * sizeof *h_a), *h_b = malloc(sg_elems * sizeof *h_b);
  CHECK(cudaMemcpy(h_a, sg0, sg_elems * sizeof *h_a, D2H));
  CHECK(cudaMemcpy(h_b, sg, sg_elems * sizeof *h_b, D2H));
  double maxv = 0, maxd = 0;
  for (size_t i = 0; i < sg_elems; i++) {
    const double dr = h_a[i].re - h_b[i].re, di = h_a[i].im - h_b[i].im;
    maxd = fmax(maxd, fmax(fabs(dr), fabs(di)));
    maxv = fmax(maxv, fmax(fabs(h_a[i].re), fabs(h_a[i].im)));
  }
  printf("subgrids %d, |gridder output| max %.4g, inverse(forward) max difference %.3g (%.2g relative)\n", S, maxv, maxd,
         maxd / maxv);
  int bad = !(maxv > 0) || !(maxd <= 2e-5 * maxv);

  CHECK(idgb200_splitter(&p, meta, sg, (const idgb200_cfloat *const *)parts, 1, G, NULL));
  CHECK(idgb200_subgrid_fft(S, N, -1, sg, NULL));
  CHECK(idgb200_degridder(&p, uvw, wn, vis2, sph, at, meta, sg, NULL));
  CHECK(cudaDeviceSynchronize());
  const size_t nv = (size_t)tt * C * 4;
  idgb200_cfloat *h_v = malloc(nv * sizeof *h_v);
  CHECK(cudaMemcpy(h_v, vis2, nv * sizeof *h_v, D2H));
  double sum = 0;
  int finite = 1;
  for (size_t i = 0; i < nv; i++) {
    finite = finite && isfinite(h_v[i].re) && isfinite(h_v[i].im);
    sum += fabs(h_v[i].re) + fabs(h_v[i].im);
  }
  printf("degridded %zu visibilities from the grid: finite %d, mean |component| %.4g; kernel launches %llu\n", nv / 4, finite,
         sum / (2.0 * nv), (unsigned long long)idgb200_launch_count());
  bad = bad || !finite || !(sum > 0);
  free(h_a); free(h_b); free(h_v);
  cudaFree(uvw); cudaFree(wn); cudaFree(sph); cudaFree(vis); cudaFree(vis2); cudaFree(at); cudaFree(sg); cudaFree(sg0);
  cudaFree(grid); cudaFree(meta);
  printf(bad ? "imaging cycle FAILED\n" : "imaging cycle OK\n");
  return bad;
}
