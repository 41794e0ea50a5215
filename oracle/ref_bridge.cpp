// TEST INFRASTRUCTURE — not product code.
//
// extern "C" bridge over the *unmodified* reference CPU implementation
// (/root/reference/app/CPU/kernels/{gridder,degridder}_reference.cpp and
// app/common/{init,common}.cpp).  oracle/Makefile (target ref) compiles those sources
// where they lie and links them with this file into oracle/_ref/libidgref.so.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
// load that library, and only as the checker / reported baseline.
//
// The bridge exposes the same flat-pointer signatures as oracle/idg_oracle.c so
// a test can swap one for the other and compare them bit for bit.
//
// Reference interfaces wrapped here:
//   cpu::c_run_gridder_reference     app/lib-cpu.hpp:9-18
//   cpu::c_run_degridder_reference   app/lib-cpu.hpp:20-29
//   initialize_*                     app/common/init.hpp:8-41
//   flops_gridder / bytes_gridder    app/common/common.hpp:36-42
//   report / report_csv              app/common/common.hpp:28-34 (common.cpp:27-98)
#include <complex>
#include <cstdint>
#include <cstdlib>

#include "lib-cpu.hpp"

#if defined(_OPENMP)
#include <omp.h>
#endif

using cfloat = std::complex<float>;
using UVW = idg::UVWCoordinate<float>;
using Vis = idg::Visibility<cfloat>;
using Jones = idg::Matrix2x2<cfloat>;

static_assert(sizeof(idg::Metadata) == 36, "metadata ABI");
static_assert(sizeof(UVW) == 12, "uvw ABI");
static_assert(sizeof(Vis) == 32, "visibility ABI");

namespace {
// The reference kernels only use .data() of their Array arguments, so the
// non-owning views below just need the right element type; the dimensions are
// filled in faithfully anyway.
struct Views {
  idg::Array2D<UVW> uvw;
  idg::Array1D<float> wavenumbers;
  idg::Array3D<Vis> visibilities;
  idg::Array2D<float> spheroidal;
  idg::Array4D<Jones> aterms;
  idg::Array1D<idg::Metadata> metadata;
  idg::Array4D<cfloat> subgrids;
  Views(int nr_subgrids, int subgrid_size, int nr_channels, int nr_stations,
        long total_timesteps, int nr_aterm_slots, void *uvw_, float *wn,
        void *vis, float *sph, void *at, void *meta, void *sg)
      : uvw((UVW *)uvw_, 1, total_timesteps),
        wavenumbers(wn, nr_channels),
        visibilities((Vis *)vis, 1, total_timesteps, nr_channels),
        spheroidal(sph, subgrid_size, subgrid_size),
        aterms((Jones *)at, nr_aterm_slots, nr_stations, subgrid_size,
               subgrid_size),
        metadata((idg::Metadata *)meta, nr_subgrids),
        subgrids((cfloat *)sg, nr_subgrids, NR_CORRELATIONS, subgrid_size,
                 subgrid_size) {}
};
} // namespace

extern "C" {

int idgref_max_threads() {
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void idgref_set_threads(int n) {
#if defined(_OPENMP)
  omp_set_num_threads(n > 0 ? n : 1);
#else
  (void)n;
#endif
}

void idgref_gridder(int nr_subgrids, int grid_size, int subgrid_size,
                    float image_size, float w_step_in_lambda, int nr_channels,
                    int nr_stations, long total_timesteps, int nr_aterm_slots,
                    void *uvw, float *wavenumbers, void *visibilities,
                    float *spheroidal, void *aterms, void *metadata,
                    void *subgrids) {
  Views v(nr_subgrids, subgrid_size, nr_channels, nr_stations, total_timesteps,
          nr_aterm_slots, uvw, wavenumbers, visibilities, spheroidal, aterms,
          metadata, subgrids);
  cpu::c_run_gridder_reference(nr_subgrids, grid_size, subgrid_size, image_size,
                               w_step_in_lambda, nr_channels, nr_stations,
                               v.uvw, v.wavenumbers, v.visibilities,
                               v.spheroidal, v.aterms, v.metadata, v.subgrids);
}

void idgref_degridder(int nr_subgrids, int grid_size, int subgrid_size,
                      float image_size, float w_step_in_lambda, int nr_channels,
                      int nr_stations, long total_timesteps,
                      int nr_aterm_slots, void *uvw, float *wavenumbers,
                      void *visibilities, float *spheroidal, void *aterms,
                      void *metadata, void *subgrids) {
  Views v(nr_subgrids, subgrid_size, nr_channels, nr_stations, total_timesteps,
          nr_aterm_slots, uvw, wavenumbers, visibilities, spheroidal, aterms,
          metadata, subgrids);
  cpu::c_run_degridder_reference(
      nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda,
      nr_channels, nr_stations, v.uvw, v.wavenumbers, v.visibilities,
      v.spheroidal, v.aterms, v.metadata, v.subgrids);
}

// ---- synthetic inputs, in the reference's own call order -------------------
void idgref_srand(unsigned seed) { srand(seed); }

void idgref_init_uvw(unsigned grid_size, int nr_baselines, int nr_timesteps,
                     void *uvw) {
  idg::Array2D<UVW> a((UVW *)uvw, nr_baselines, nr_timesteps);
  initialize_uvw(grid_size, a);
}

void idgref_init_frequencies(int nr_channels, float *frequencies) {
  idg::Array1D<float> f(frequencies, nr_channels);
  initialize_frequencies(f);
}

void idgref_init_wavenumbers(int nr_channels, float *frequencies,
                             float *wavenumbers) {
  idg::Array1D<float> f(frequencies, nr_channels);
  idg::Array1D<float> w(wavenumbers, nr_channels);
  initialize_wavenumbers(f, w);
}

void idgref_init_visibilities(unsigned grid_size, float image_size,
                              int nr_baselines, int nr_timesteps,
                              int nr_channels, float *frequencies, void *uvw,
                              void *visibilities) {
  idg::Array1D<float> f(frequencies, nr_channels);
  idg::Array2D<UVW> u((UVW *)uvw, nr_baselines, nr_timesteps);
  idg::Array3D<Vis> v((Vis *)visibilities, nr_baselines, nr_timesteps,
                      nr_channels);
  initialize_visibilities(grid_size, image_size, f, u, v);
}

void idgref_init_baselines(unsigned nr_stations, int nr_baselines,
                           void *baselines) {
  idg::Array1D<idg::Baseline> b((idg::Baseline *)baselines, nr_baselines);
  initialize_baselines(nr_stations, b);
}

void idgref_init_spheroidal(int subgrid_size, float *spheroidal) {
  idg::Array2D<float> s(spheroidal, subgrid_size, subgrid_size);
  initialize_spheroidal(s);
}

void idgref_init_aterms(int nr_timeslots, int nr_stations, int subgrid_size,
                        float *spheroidal, void *aterms) {
  idg::Array2D<float> s(spheroidal, subgrid_size, subgrid_size);
  idg::Array4D<Jones> a((Jones *)aterms, nr_timeslots, nr_stations,
                        subgrid_size, subgrid_size);
  initialize_aterms(s, a);
}

void idgref_init_metadata(unsigned grid_size, unsigned nr_timeslots,
                          unsigned nr_timesteps_subgrid, int nr_baselines,
                          void *baselines, void *metadata) {
  idg::Array1D<idg::Baseline> b((idg::Baseline *)baselines, nr_baselines);
  idg::Array1D<idg::Metadata> m((idg::Metadata *)metadata,
                                (size_t)nr_baselines * nr_timeslots);
  initialize_metadata(grid_size, nr_timeslots, nr_timesteps_subgrid, b, m);
}

void idgref_init_subgrids(int nr_subgrids, int subgrid_size, void *subgrids) {
  idg::Array4D<cfloat> s((cfloat *)subgrids, nr_subgrids, NR_CORRELATIONS,
                         subgrid_size, subgrid_size);
  initialize_subgrids(s);
}

// ---- metric model -----------------------------------------------------------
uint64_t idgref_flops_gridder(uint64_t nr_channels, uint64_t nr_timesteps,
                              uint64_t nr_subgrids, uint64_t subgrid_size,
                              uint64_t nr_correlations) {
  return flops_gridder(nr_channels, nr_timesteps, nr_subgrids, subgrid_size,
                       nr_correlations);
}

uint64_t idgref_bytes_gridder(uint64_t nr_channels, uint64_t nr_timesteps,
                              uint64_t nr_subgrids, uint64_t subgrid_size,
                              uint64_t nr_correlations) {
  return bytes_gridder(nr_channels, nr_timesteps, nr_subgrids, subgrid_size,
                       nr_correlations);
}

// ---- report line and CSV (app/common/common.cpp:27-98), as the reference's runners call them
// (app/CUDA/util.cpp:157-160): the line goes to stdout, the CSV to $OUTPUT_PATH/<device>-<name><ext>
void idgref_report(const char *name, double seconds, double gflops, double gbytes, double mvis,
                   double joules) {
  report(name, seconds, gflops, gbytes, mvis, joules);
}

void idgref_report_csv(const char *name, const char *device_name, const char *file_extension,
                       double seconds, double gflops, double gbytes, double mvis, double joules) {
  report_csv(name, device_name, file_extension, seconds, gflops, gbytes, mvis, joules);
}

} // extern "C"
