/* TEST INFRASTRUCTURE — see idg_oracle.h.  Plain-C restatement of the reference's
 * CPU gridder/degridder, its synthetic-input generators and its metric model.
 *
 * Rounding contract.  The reference is built with GCC's default
 * -ffp-contract=fast, so which multiply-adds are fused is a property of its
 * binary, not of its source.  This file is compiled with -ffp-contract=off and
 * spells every fusion the reference binary performs as an explicit fmaf()
 * (read off `objdump -d` of oracle/_ref/libidgref.so), so the restatement is
 * bit-identical to the reference build and independent of compiler mood.
 *   complex a*b         re = fma(a.re, b.re, -(a.im*b.im))
 *                       im = fma(a.im, b.re,   a.re*b.im )
 *   gridder   phase_index = fma(w, n, fma(u, l, v*m))      (offset likewise)
 *             phase       = fma(-phase_index, k, phase_offset)
 *   degridder phase_index = fma(u, l, v*m) + w*n           (offset likewise)
 *             phase       = fma(phase_index, k, -phase_offset)
 */
#define _GNU_SOURCE /* sincosf */
#include "idg_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#if defined(_OPENMP)
#include <omp.h>
#endif

#define NR_POL 4 /* app/common/parameters.hpp:3 */

int idgo_max_threads(void) {
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void idgo_set_threads(int n) {
#if defined(_OPENMP)
  omp_set_num_threads(n > 0 ? n : 1);
#else
  (void)n;
#endif
}

typedef struct {
  float re, im;
} cf;

/* std::complex<float> operator* as GCC contracts it in the reference binary.
 * Two shapes occur (found by exhaustive search over the four possible
 * contractions of each product against the reference's output bits):
 *   cmul    the common one: the cross term a.re*b.im is rounded, a.im*b.re fused
 *   cmul_x  imaginary part fused the other way round; used for A1^H*P rows 0-1
 *           in the gridder and for every row of A1*P in the degridder        */
static inline cf cmul(cf a, cf b) {
  cf r;
  r.re = fmaf(a.re, b.re, -(a.im * b.im));
  r.im = fmaf(a.im, b.re, a.re * b.im);
  return r;
}

static inline cf cmul_x(cf a, cf b) {
  cf r;
  r.re = fmaf(a.re, b.re, -(a.im * b.im));
  r.im = fmaf(a.re, b.im, a.im * b.re);
  return r;
}

static inline cf cadd(cf a, cf b) {
  cf r = {a.re + b.re, a.im + b.im};
  return r;
}

/* app/common/math.hpp:9-17 — double intermediate, one rounding to float */
static inline float compute_l(int x, int subgrid_size, float image_size) {
  return (float)((x + 0.5 - (subgrid_size / 2)) * (double)image_size /
                 (double)subgrid_size);
}

/* app/common/math.hpp:19-24 */
static inline float compute_n(float l, float m) {
  const float tmp = fmaf(l, l, m * m);
  return tmp > 1.0f ? 1.0f : tmp / (1.0f + sqrtf(1.0f - tmp));
}

/* app/common/math.hpp:26-36.  x01 / x23: rows {0,1} / {2,3} of c use cmul_x. */
static inline void matmul(const cf *a, const cf *b, cf *c, int x01, int x23) {
  if (x01) {
    c[0] = cadd(cmul_x(a[0], b[0]), cmul_x(a[1], b[2]));
    c[1] = cadd(cmul_x(a[0], b[1]), cmul_x(a[1], b[3]));
  } else {
    c[0] = cadd(cmul(a[0], b[0]), cmul(a[1], b[2]));
    c[1] = cadd(cmul(a[0], b[1]), cmul(a[1], b[3]));
  }
  if (x23) {
    c[2] = cadd(cmul_x(a[2], b[0]), cmul_x(a[3], b[2]));
    c[3] = cadd(cmul_x(a[2], b[1]), cmul_x(a[3], b[3]));
  } else {
    c[2] = cadd(cmul(a[2], b[0]), cmul(a[3], b[2]));
    c[3] = cadd(cmul(a[2], b[1]), cmul(a[3], b[3]));
  }
}

/* app/common/math.hpp:38-62: conjugate (multiply by +-1) then transpose */
static inline void hermitian(const cf *a, cf *b) {
  cf t[4];
  for (int i = 0; i < 4; i++) {
    t[i].re = 1.0f * a[i].re;
    t[i].im = -1.0f * a[i].im;
  }
  b[0] = t[0];
  b[1] = t[2];
  b[2] = t[1];
  b[3] = t[3];
}

/* app/common/math.hpp:64-77: P <- A1^H * P * A2 */
static inline void apply_aterm_gridder(cf *pixels, const cf *aterm1,
                                       const cf *aterm2) {
  cf a1h[4], temp[4];
  hermitian(aterm1, a1h);
  matmul(a1h, pixels, temp, 1, 0);
  matmul(temp, aterm2, pixels, 0, 0);
}

/* app/common/math.hpp:79-92: P <- A1 * P * A2^H */
static inline void apply_aterm_degridder(cf *pixels, const cf *aterm1,
                                         const cf *aterm2) {
  cf temp[4], a2h[4];
  matmul(aterm1, pixels, temp, 1, 1);
  hermitian(aterm2, a2h);
  matmul(temp, a2h, pixels, 0, 0);
}

typedef struct {
  long time_offset;
  int nr_timesteps, aterm_index, station1, station2;
  float u_offset, v_offset, w_offset;
} subgrid_ctx;

/* gridder_reference.cpp:23-39 == degridder_reference.cpp:24-32,77-79 */
static inline subgrid_ctx load_ctx(const idgo_metadata *metadata, int s,
                                   int grid_size, int subgrid_size,
                                   float image_size, float w_step_in_lambda) {
  const idgo_metadata m = metadata[s];
  subgrid_ctx c;
  c.time_offset = (m.baseline_offset - metadata[0].baseline_offset) + m.time_offset;
  c.nr_timesteps = m.nr_timesteps;
  c.aterm_index = m.aterm_index;
  c.station1 = (int)m.station1;
  c.station2 = (int)m.station2;
  const float w_offset_in_lambda = (float)(w_step_in_lambda * (m.z + 0.5));
  c.u_offset = (float)((m.x + subgrid_size / 2 - grid_size / 2) *
                       (2 * M_PI / image_size));
  c.v_offset = (float)((m.y + subgrid_size / 2 - grid_size / 2) *
                       (2 * M_PI / image_size));
  c.w_offset = (float)(2 * M_PI * w_offset_in_lambda);
  return c;
}

void idgo_gridder(int nr_subgrids, int grid_size, int subgrid_size,
                  float image_size, float w_step_in_lambda, int nr_channels,
                  int nr_stations, const idgo_uvw *uvw,
                  const float *wavenumbers, const float *visibilities,
                  const float *spheroidal, const float *aterms,
                  const idgo_metadata *metadata, float *subgrids) {
  const int N = subgrid_size;
  const cf *vis = (const cf *)visibilities;
  const cf *at = (const cf *)aterms;
  cf *out = (cf *)subgrids;

#pragma omp parallel for schedule(dynamic, 1)
  for (int s = 0; s < nr_subgrids; s++) {
    const subgrid_ctx c =
        load_ctx(metadata, s, grid_size, N, image_size, w_step_in_lambda);

    for (int y = 0; y < N; y++) {
      for (int x = 0; x < N; x++) {
        cf pixels[NR_POL];
        memset(pixels, 0, sizeof pixels);

        const float l = compute_l(x, N, image_size);
        const float m = compute_l(y, N, image_size);
        const float n = compute_n(l, m);

        /* t-invariant; the reference recomputes it per timestep with the same
         * operands, hence the same bits (gridder_reference.cpp:64) */
        const float phase_offset =
            fmaf(c.w_offset, n, fmaf(c.u_offset, l, c.v_offset * m));

        for (int time = 0; time < c.nr_timesteps; time++) {
          const idgo_uvw p = uvw[c.time_offset + time];
          const float phase_index = fmaf(p.w, n, fmaf(p.u, l, p.v * m));

          for (int chan = 0; chan < nr_channels; chan++) {
            const float phase =
                fmaf(-phase_index, wavenumbers[chan], phase_offset);
            float sn, cs;
            sincosf(phase, &sn, &cs); /* GCC merges cosf+sinf (:72) into this */
            const cf phasor = {cs, sn};

            const size_t index =
                ((size_t)(c.time_offset + time) * nr_channels + chan) * NR_POL;
            for (int pol = 0; pol < NR_POL; pol++) {
              pixels[pol] = cadd(pixels[pol], cmul(vis[index + pol], phasor));
            }
          }
        }

        const size_t a1 =
            (((size_t)c.aterm_index * nr_stations + c.station1) * N * N +
             (size_t)y * N + x) * NR_POL;
        const size_t a2 =
            (((size_t)c.aterm_index * nr_stations + c.station2) * N * N +
             (size_t)y * N + x) * NR_POL;
        apply_aterm_gridder(pixels, &at[a1], &at[a2]);

        const float sph = spheroidal[y * N + x];
        for (int pol = 0; pol < NR_POL; pol++) {
          const size_t idx =
              (size_t)s * NR_POL * N * N + (size_t)pol * N * N + (size_t)y * N + x;
          out[idx].re = pixels[pol].re * sph;
          out[idx].im = pixels[pol].im * sph;
        }
      }
    }
  }
}

void idgo_degridder(int nr_subgrids, int grid_size, int subgrid_size,
                    float image_size, float w_step_in_lambda, int nr_channels,
                    int nr_stations, const idgo_uvw *uvw,
                    const float *wavenumbers, float *visibilities,
                    const float *spheroidal, const float *aterms,
                    const idgo_metadata *metadata, const float *subgrids) {
  const int N = subgrid_size;
  cf *vis = (cf *)visibilities;
  const cf *at = (const cf *)aterms;
  const cf *in = (const cf *)subgrids;

#pragma omp parallel
  {
    cf *pixels = (cf *)malloc((size_t)N * N * NR_POL * sizeof(cf));
    float *lmn = (float *)malloc((size_t)N * N * 4 * sizeof(float));

#pragma omp for schedule(dynamic, 1)
    for (int s = 0; s < nr_subgrids; s++) {
      const subgrid_ctx c =
          load_ctx(metadata, s, grid_size, N, image_size, w_step_in_lambda);

      /* degridder_reference.cpp:38-74: P' = A1 * (sph*S) * A2^H */
      for (int y = 0; y < N; y++) {
        for (int x = 0; x < N; x++) {
          const size_t a1 =
              (((size_t)c.aterm_index * nr_stations + c.station1) * N * N +
               (size_t)y * N + x) * NR_POL;
          const size_t a2 =
              (((size_t)c.aterm_index * nr_stations + c.station2) * N * N +
               (size_t)y * N + x) * NR_POL;
          const float sph = spheroidal[y * N + x];
          cf px[NR_POL];
          for (int pol = 0; pol < NR_POL; pol++) {
            const size_t idx = (size_t)s * NR_POL * N * N +
                               (size_t)pol * N * N + (size_t)y * N + x;
            px[pol].re = sph * in[idx].re;
            px[pol].im = sph * in[idx].im;
          }
          apply_aterm_degridder(px, &at[a1], &at[a2]);
          for (int pol = 0; pol < NR_POL; pol++)
            pixels[((size_t)y * N + x) * NR_POL + pol] = px[pol];

          /* l, m, n and phase_offset do not depend on (time, chan); the
           * reference recomputes them per visibility from identical operands
           * (:100-108), so caching them changes no bit. */
          const float l = compute_l(x, N, image_size);
          const float m = compute_l(y, N, image_size);
          const float n = compute_n(l, m);
          float *q = &lmn[((size_t)y * N + x) * 4];
          q[0] = l;
          q[1] = m;
          q[2] = n;
          /* NB: unlike the gridder, the reference binary leaves the w terms of
           * the degridder unfused (w*n and w_offset*n are rounded products
           * that are then added; objdump of kernel_degridder_reference). */
          q[3] = fmaf(c.u_offset, l, c.v_offset * m) + c.w_offset * n;
        }
      }

      /* :82-127 */
      for (int time = 0; time < c.nr_timesteps; time++) {
        const idgo_uvw p = uvw[c.time_offset + time];
        for (int chan = 0; chan < nr_channels; chan++) {
          cf sum[NR_POL];
          memset(sum, 0, sizeof sum);
          const float k = wavenumbers[chan];
          for (int i = 0; i < N * N; i++) {
            const float *q = &lmn[(size_t)i * 4];
            const float phase_index = fmaf(p.u, q[0], p.v * q[1]) + p.w * q[2];
            const float phase = fmaf(phase_index, k, -q[3]);
            float sn, cs;
            sincosf(phase, &sn, &cs);
            const cf phasor = {cs, sn};
            for (int pol = 0; pol < NR_POL; pol++)
              sum[pol] = cadd(sum[pol], cmul(pixels[(size_t)i * NR_POL + pol], phasor));
          }
          const size_t index =
              ((size_t)(c.time_offset + time) * nr_channels + chan) * NR_POL;
          for (int pol = 0; pol < NR_POL; pol++) vis[index + pol] = sum[pol];
        }
      }
    }
    free(pixels);
    free(lmn);
  }
}

/* ---------------------------------------------------------------- float64 -- */
typedef struct {
  double re, im;
} cd;

static inline cd dmul(cd a, cd b) {
  cd r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
  return r;
}
static inline cd dadd(cd a, cd b) {
  cd r = {a.re + b.re, a.im + b.im};
  return r;
}
static inline void dmatmul(const cd *a, const cd *b, cd *c) {
  c[0] = dadd(dmul(a[0], b[0]), dmul(a[1], b[2]));
  c[1] = dadd(dmul(a[0], b[1]), dmul(a[1], b[3]));
  c[2] = dadd(dmul(a[2], b[0]), dmul(a[3], b[2]));
  c[3] = dadd(dmul(a[2], b[1]), dmul(a[3], b[3]));
}
static inline void dherm(const cd *a, cd *b) {
  const cd t[4] = {{a[0].re, -a[0].im},
                   {a[2].re, -a[2].im},
                   {a[1].re, -a[1].im},
                   {a[3].re, -a[3].im}};
  memcpy(b, t, sizeof t);
}
static inline void dload(const cf *a, cd *b) {
  for (int i = 0; i < 4; i++) {
    b[i].re = a[i].re;
    b[i].im = a[i].im;
  }
}

typedef struct {
  double l, m, n, phase_offset;
} dlmn;

static inline dlmn d_lmn(int x, int y, int N, float image_size, double u_off,
                         double v_off, double w_off) {
  dlmn r;
  r.l = (x + 0.5 - (N / 2)) * (double)image_size / (double)N;
  r.m = (y + 0.5 - (N / 2)) * (double)image_size / (double)N;
  const double tmp = r.l * r.l + r.m * r.m;
  r.n = tmp > 1.0 ? 1.0 : tmp / (1.0 + sqrt(1.0 - tmp));
  r.phase_offset = u_off * r.l + v_off * r.m + w_off * r.n;
  return r;
}

void idgo_gridder_f64(int nr_subgrids, int grid_size, int subgrid_size,
                      float image_size, float w_step_in_lambda, int nr_channels,
                      int nr_stations, const idgo_uvw *uvw,
                      const float *wavenumbers, const float *visibilities,
                      const float *spheroidal, const float *aterms,
                      const idgo_metadata *metadata, double *subgrids) {
  const int N = subgrid_size;
  const cf *vis = (const cf *)visibilities;
  const cf *at = (const cf *)aterms;
  cd *out = (cd *)subgrids;

#pragma omp parallel for schedule(dynamic, 1)
  for (int s = 0; s < nr_subgrids; s++) {
    const idgo_metadata md = metadata[s];
    const long time_offset =
        (md.baseline_offset - metadata[0].baseline_offset) + md.time_offset;
    const double u_off =
        (md.x + N / 2 - grid_size / 2) * (2 * M_PI / (double)image_size);
    const double v_off =
        (md.y + N / 2 - grid_size / 2) * (2 * M_PI / (double)image_size);
    const double w_off = 2 * M_PI * ((double)w_step_in_lambda * (md.z + 0.5));

    for (int y = 0; y < N; y++) {
      for (int x = 0; x < N; x++) {
        const dlmn g = d_lmn(x, y, N, image_size, u_off, v_off, w_off);
        cd px[NR_POL] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        for (int time = 0; time < md.nr_timesteps; time++) {
          const idgo_uvw p = uvw[time_offset + time];
          const double phase_index =
              (double)p.u * g.l + (double)p.v * g.m + (double)p.w * g.n;
          for (int chan = 0; chan < nr_channels; chan++) {
            const double phase =
                g.phase_offset - phase_index * (double)wavenumbers[chan];
            const cd phasor = {cos(phase), sin(phase)};
            const size_t index =
                ((size_t)(time_offset + time) * nr_channels + chan) * NR_POL;
            for (int pol = 0; pol < NR_POL; pol++) {
              const cd v = {vis[index + pol].re, vis[index + pol].im};
              px[pol] = dadd(px[pol], dmul(v, phasor));
            }
          }
        }
        cd a1[4], a2[4], a1h[4], t[4], r[4];
        dload(&at[(((size_t)md.aterm_index * nr_stations + md.station1) * N * N +
                   (size_t)y * N + x) * NR_POL], a1);
        dload(&at[(((size_t)md.aterm_index * nr_stations + md.station2) * N * N +
                   (size_t)y * N + x) * NR_POL], a2);
        dherm(a1, a1h);
        dmatmul(a1h, px, t);
        dmatmul(t, a2, r);
        const double sph = spheroidal[y * N + x];
        for (int pol = 0; pol < NR_POL; pol++) {
          const size_t idx =
              (size_t)s * NR_POL * N * N + (size_t)pol * N * N + (size_t)y * N + x;
          out[idx].re = r[pol].re * sph;
          out[idx].im = r[pol].im * sph;
        }
      }
    }
  }
}

void idgo_degridder_f64(int nr_subgrids, int grid_size, int subgrid_size,
                        float image_size, float w_step_in_lambda,
                        int nr_channels, int nr_stations, const idgo_uvw *uvw,
                        const float *wavenumbers, double *visibilities,
                        const float *spheroidal, const float *aterms,
                        const idgo_metadata *metadata, const float *subgrids) {
  const int N = subgrid_size;
  cd *vis = (cd *)visibilities;
  const cf *at = (const cf *)aterms;
  const cf *in = (const cf *)subgrids;

#pragma omp parallel
  {
    cd *pixels = (cd *)malloc((size_t)N * N * NR_POL * sizeof(cd));
    dlmn *geo = (dlmn *)malloc((size_t)N * N * sizeof(dlmn));

#pragma omp for schedule(dynamic, 1)
    for (int s = 0; s < nr_subgrids; s++) {
      const idgo_metadata md = metadata[s];
      const long time_offset =
          (md.baseline_offset - metadata[0].baseline_offset) + md.time_offset;
      const double u_off =
          (md.x + N / 2 - grid_size / 2) * (2 * M_PI / (double)image_size);
      const double v_off =
          (md.y + N / 2 - grid_size / 2) * (2 * M_PI / (double)image_size);
      const double w_off = 2 * M_PI * ((double)w_step_in_lambda * (md.z + 0.5));

      for (int y = 0; y < N; y++) {
        for (int x = 0; x < N; x++) {
          cd a1[4], a2[4], a2h[4], t[4], px[4], r[4];
          dload(&at[(((size_t)md.aterm_index * nr_stations + md.station1) * N * N +
                     (size_t)y * N + x) * NR_POL], a1);
          dload(&at[(((size_t)md.aterm_index * nr_stations + md.station2) * N * N +
                     (size_t)y * N + x) * NR_POL], a2);
          const double sph = spheroidal[y * N + x];
          for (int pol = 0; pol < NR_POL; pol++) {
            const size_t idx = (size_t)s * NR_POL * N * N +
                               (size_t)pol * N * N + (size_t)y * N + x;
            px[pol].re = sph * in[idx].re;
            px[pol].im = sph * in[idx].im;
          }
          dmatmul(a1, px, t);
          dherm(a2, a2h);
          dmatmul(t, a2h, r);
          memcpy(&pixels[((size_t)y * N + x) * NR_POL], r, sizeof r);
          geo[(size_t)y * N + x] = d_lmn(x, y, N, image_size, u_off, v_off, w_off);
        }
      }

      for (int time = 0; time < md.nr_timesteps; time++) {
        const idgo_uvw p = uvw[time_offset + time];
        for (int chan = 0; chan < nr_channels; chan++) {
          cd sum[NR_POL] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
          const double k = wavenumbers[chan];
          for (int i = 0; i < N * N; i++) {
            const dlmn g = geo[i];
            const double phase_index =
                (double)p.u * g.l + (double)p.v * g.m + (double)p.w * g.n;
            const double phase = phase_index * k - g.phase_offset;
            const cd phasor = {cos(phase), sin(phase)};
            for (int pol = 0; pol < NR_POL; pol++)
              sum[pol] = dadd(sum[pol], dmul(pixels[(size_t)i * NR_POL + pol], phasor));
          }
          const size_t index =
              ((size_t)(time_offset + time) * nr_channels + chan) * NR_POL;
          for (int pol = 0; pol < NR_POL; pol++) vis[index + pol] = sum[pol];
        }
      }
    }
    free(pixels);
    free(geo);
  }
}

/* ------------------------------------------------------- synthetic inputs -- */
void idgo_srand(unsigned seed) { srand(seed); }

void idgo_init_uvw(unsigned grid_size, int nr_baselines, int nr_timesteps,
                   idgo_uvw *uvw) {
  for (int bl = 0; bl < nr_baselines; bl++) {
    const float radius_u = (float)((grid_size / 2) +
                                   (double)rand() / (double)(RAND_MAX) * (grid_size / 2));
    const float radius_v = (float)((grid_size / 2) +
                                   (double)rand() / (double)(RAND_MAX) * (grid_size / 2));
    for (int time = 0; time < nr_timesteps; time++) {
      const float angle =
          (float)(((unsigned)time + 0.5) / (double)(360.0f / (float)(unsigned)nr_timesteps));
      idgo_uvw p;
      p.u = (float)((double)radius_u * cos((double)angle * M_PI));
      p.v = (float)((double)radius_v * sin((double)angle * M_PI));
      p.w = 0;
      uvw[(size_t)bl * nr_timesteps + time] = p;
    }
  }
}

void idgo_init_frequencies(int nr_channels, float *frequencies) {
  const unsigned int start_frequency = 150e6;
  const float frequency_increment = 0.7e6;
  for (int i = 0; i < nr_channels; i++) {
    /* unsigned + float*unsigned is float arithmetic in the reference (:33) */
    const double frequency =
        (float)start_frequency + frequency_increment * (float)(unsigned)i;
    frequencies[i] = (float)frequency;
  }
}

void idgo_init_wavenumbers(int nr_channels, const float *frequencies,
                           float *wavenumbers) {
  const double speed_of_light = 299792458.0;
  for (int i = 0; i < nr_channels; i++)
    wavenumbers[i] = (float)(2 * M_PI * frequencies[i] / speed_of_light);
}

void idgo_init_visibilities(unsigned grid_size, float image_size,
                            int nr_baselines, int nr_timesteps, int nr_channels,
                            const float *frequencies, const idgo_uvw *uvw,
                            float *visibilities) {
  const float x_offset = (float)(0.6 * grid_size);
  const float y_offset = (float)(0.7 * grid_size);
  const float amplitude = 1.0f;
  const float l = x_offset * image_size / (float)grid_size;
  const float m = y_offset * image_size / (float)grid_size;
  cf *vis = (cf *)visibilities;
  const float scale[NR_POL] = {1.01f, 1.02f, 1.03f, 1.04f};

  for (int bl = 0; bl < nr_baselines; bl++) {
    for (int time = 0; time < nr_timesteps; time++) {
      const idgo_uvw p = uvw[(size_t)bl * nr_timesteps + time];
      for (int chan = 0; chan < nr_channels; chan++) {
        const double speed_of_light = 299792458.0;
        const float u = (float)((frequencies[chan] / speed_of_light) * p.u);
        const float v = (float)((frequencies[chan] / speed_of_light) * p.v);
        const float arg = (float)(-2 * M_PI * (double)fmaf(u, l, v * m));
        const float complex e = cexpf(CMPLXF(0.0f, arg));
        const cf value = {amplitude * crealf(e), amplitude * cimagf(e)};
        cf *dst = &vis[(((size_t)bl * nr_timesteps + time) * nr_channels + chan) * NR_POL];
        for (int pol = 0; pol < NR_POL; pol++) {
          dst[pol].re = value.re * scale[pol];
          dst[pol].im = value.im * scale[pol];
        }
      }
    }
  }
}

void idgo_init_baselines(unsigned nr_stations, int nr_baselines,
                         idgo_baseline *baselines) {
  int bl = 0;
  for (unsigned s1 = 0; s1 < nr_stations; s1++) {
    for (unsigned s2 = s1 + 1; s2 < nr_stations; s2++) {
      if (bl >= nr_baselines) break;
      baselines[bl].station1 = s1;
      baselines[bl].station2 = s2;
      bl++;
    }
  }
}

void idgo_init_spheroidal(int subgrid_size, float *spheroidal) {
  for (int y = 0; y < subgrid_size; y++) {
    const float tmp_y = fabsf(-1 + (unsigned)y * 2.0f / (float)subgrid_size);
    for (int x = 0; x < subgrid_size; x++) {
      const float tmp_x = fabsf(-1 + (unsigned)x * 2.0f / (float)subgrid_size);
      spheroidal[y * subgrid_size + x] = tmp_y * tmp_x;
    }
  }
}

void idgo_init_aterms(int nr_timeslots, int nr_stations, int subgrid_size,
                      const float *spheroidal, float *aterms) {
  cf *a = (cf *)aterms;
  const int N = subgrid_size;
  for (int ts = 0; ts < nr_timeslots; ts++)
    for (int st = 0; st < nr_stations; st++)
      for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++) {
          const float scale =
              (float)(0.8 + ((double)rand() / (double)(RAND_MAX) * 0.4));
          const float value = spheroidal[y * N + x] * scale;
          cf *dst = &a[((((size_t)ts * nr_stations + st) * N + y) * N + x) * NR_POL];
          dst[0].re = (float)(value + 0.1);
          dst[0].im = (float)-0.1;
          dst[1].re = (float)(value - 0.2);
          dst[1].im = (float)0.1;
          dst[2].re = (float)(value - 0.2);
          dst[2].im = (float)0.1;
          dst[3].re = (float)(value + 0.1);
          dst[3].im = (float)-0.1;
        }
}

void idgo_init_metadata(unsigned grid_size, unsigned nr_timeslots,
                        unsigned nr_timesteps_subgrid, int nr_baselines,
                        const idgo_baseline *baselines,
                        idgo_metadata *metadata) {
  for (unsigned bl = 0; bl < (unsigned)nr_baselines; bl++)
    for (unsigned ts = 0; ts < nr_timeslots; ts++) {
      idgo_metadata m;
      m.baseline_offset = 0;
      m.time_offset =
          (int)(bl * nr_timeslots * nr_timesteps_subgrid + ts * nr_timesteps_subgrid);
      m.nr_timesteps = (int)nr_timesteps_subgrid;
      m.aterm_index = 0;
      m.station1 = baselines[bl].station1;
      m.station2 = baselines[bl].station2;
      m.x = (int)((double)rand() / (double)(RAND_MAX) * grid_size);
      m.y = (int)((double)rand() / (double)(RAND_MAX) * grid_size);
      m.z = 0;
      metadata[bl * nr_timeslots + ts] = m;
    }
}

void idgo_init_subgrids(int nr_subgrids, int subgrid_size, float *subgrids) {
  cf *sg = (cf *)subgrids;
  const unsigned N = (unsigned)subgrid_size;
  for (unsigned s = 0; s < (unsigned)nr_subgrids; s++)
    for (unsigned c = 0; c < NR_POL; c++)
      for (unsigned y = 0; y < N; y++)
        for (unsigned x = 0; x < N; x++) {
          cf *dst = &sg[(((size_t)s * NR_POL + c) * N + y) * N + x];
          dst->re = (float)(y * N + x + 1) / ((float)100 * (float)N * (float)N);
          dst->im = (float)c / 10.0f;
        }
}

/* ------------------------------------------------------------ metric model -- */
uint64_t idgo_flops_gridder(uint64_t nr_channels, uint64_t nr_timesteps,
                            uint64_t nr_subgrids, uint64_t subgrid_size,
                            uint64_t nr_correlations) {
  /* common.cpp:100-120 */
  uint64_t per_vis = 0;
  per_vis += 5;                                /* phase index  */
  per_vis += 5;                                /* phase offset */
  per_vis += nr_channels * 2;                  /* phase        */
  per_vis += nr_channels * nr_correlations * 8; /* update       */
  const uint64_t per_subgrid = 6;              /* shift        */
  return nr_timesteps * subgrid_size * subgrid_size * per_vis +
         nr_subgrids * subgrid_size * subgrid_size * per_subgrid;
}

uint64_t idgo_bytes_gridder(uint64_t nr_channels, uint64_t nr_timesteps,
                            uint64_t nr_subgrids, uint64_t subgrid_size,
                            uint64_t nr_correlations) {
  /* common.cpp:122-159 */
  const uint64_t uvw = 3 * sizeof(float);
  const uint64_t vis = nr_channels * nr_correlations * 2 * sizeof(float);
  const uint64_t pix = 2 * nr_correlations * 2 * sizeof(float); /* read + write */
  const uint64_t aterm = 2 * nr_correlations * 2 * sizeof(float);
  const uint64_t sph = sizeof(float);
  const uint64_t npix = nr_subgrids * subgrid_size * subgrid_size;
  return nr_timesteps * uvw + nr_timesteps * vis + npix * pix + npix * aterm +
         npix * sph;
}

/* tests/test_util.hpp:28-92 */
double idgo_check_error(int n, const float *A_, const float *B_) {
  const cf *A = (const cf *)A_;
  const cf *B = (const cf *)B_;
  double r_error = 0.0, i_error = 0.0;
  int nnz = 0;
  float r_max = 1, i_max = 1;
  for (int i = 0; i < n; i++) {
    const float r = fabsf(A[i].re), im = fabsf(A[i].im);
    if (r > r_max) r_max = r;
    if (im > i_max) i_max = im;
  }
  for (int i = 0; i < n; i++) {
    const double r_diff = (double)(B[i].re - A[i].re);
    const double i_diff = (double)(B[i].im - A[i].im);
    if (hypotf(B[i].re, B[i].im) > 0.0f) {
      nnz++;
      r_error += (r_diff * r_diff) / r_max;
      i_error += (i_diff * i_diff) / i_max;
    }
  }
  r_error /= (nnz > 1 ? nnz : 1);
  i_error /= (nnz > 1 ? nnz : 1);
  return sqrt(r_error + i_error);
}
