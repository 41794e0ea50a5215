// Drop-in replacement for the reference's per-kernel translation unit + app/CUDA/util.cpp.
//
// Defines, with the reference's own C++ signatures, the symbols its test mains link
// against (tests/gridder_common.cpp:12-31, tests/degridder_common.cpp:12-31,
// app/lib-cuda.hpp:5-9):
//     cuda::p_run_gridder      cuda::c_run_gridder
//     cuda::p_run_degridder    cuda::c_run_degridder
//     cuda::print_device_info  cuda::print_benchmark
// and forwards them to the C ABI of libidgb200.so (include/idg_b200.h).  It is compiled
// against the reference's headers (-I<reference>/app ...), exactly as a maintainer would
// compile it inside the reference tree; nothing of the reference is copied here.
//
// Error convention: like cudaCheck (app/CUDA/util.cpp:5-15) a failing call prints the
// reason and exits the process with a non-zero code.
#include <cstdio>
#include <cstdlib>

#include "lib-common.hpp"  // reference: idg::ArrayND, idg::Metadata, parameters (via common.hpp)

#include "idg_b200.h"

namespace {

void check(int rc, const char *what) {
  if (rc != IDGB200_OK) {
    std::fprintf(stderr, "idgb200 assert: %s: %s (%d)\n", what, idgb200_error_string(rc), rc);
    std::exit(rc > 0 ? rc : 1);
  }
}

static_assert(sizeof(idg::Metadata) == sizeof(idgb200_metadata), "metadata layout");
static_assert(sizeof(idg::UVWCoordinate<float>) == sizeof(idgb200_uvw), "uvw layout");
static_assert(sizeof(idg::Visibility<std::complex<float>>) == 4 * sizeof(idgb200_cfloat), "visibility layout");
static_assert(sizeof(idg::Matrix2x2<std::complex<float>>) == 4 * sizeof(idgb200_cfloat), "aterm layout");

}  // namespace

namespace cuda {

void print_device_info() { check(idgb200_print_device_info(), "print_device_info"); }

void print_benchmark() { std::printf(">>> CUDA IDG BENCHMARK (idg-b200, sm_100a)\n"); }

void p_run_gridder() { check(idgb200_p_run_gridder(nullptr), "p_run_gridder"); }

void p_run_degridder() { check(idgb200_p_run_degridder(nullptr), "p_run_degridder"); }

void c_run_gridder(int nr_subgrids, int grid_size, int subgrid_size, float image_size,
                   float w_step_in_lambda, int nr_channels, int nr_stations,
                   idg::Array2D<idg::UVWCoordinate<float>> &uvw, idg::Array1D<float> &wavenumbers,
                   idg::Array3D<idg::Visibility<std::complex<float>>> &visibilities,
                   idg::Array2D<float> &spheroidal,
                   idg::Array4D<idg::Matrix2x2<std::complex<float>>> &aterms,
                   idg::Array1D<idg::Metadata> &metadata, idg::Array4D<std::complex<float>> &subgrids) {
  check(idgb200_c_run_gridder(
            nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda, nr_channels, nr_stations,
            (int64_t)uvw.size(), (int)aterms.get_w_dim(), reinterpret_cast<const idgb200_uvw *>(uvw.data()),
            wavenumbers.data(), reinterpret_cast<const idgb200_cfloat *>(visibilities.data()),
            spheroidal.data(), reinterpret_cast<const idgb200_cfloat *>(aterms.data()),
            reinterpret_cast<const idgb200_metadata *>(metadata.data()),
            reinterpret_cast<idgb200_cfloat *>(subgrids.data())),
        "c_run_gridder");
}

void c_run_degridder(int nr_subgrids, int grid_size, int subgrid_size, float image_size,
                     float w_step_in_lambda, int nr_channels, int nr_stations,
                     idg::Array2D<idg::UVWCoordinate<float>> &uvw, idg::Array1D<float> &wavenumbers,
                     idg::Array3D<idg::Visibility<std::complex<float>>> &visibilities,
                     idg::Array2D<float> &spheroidal,
                     idg::Array4D<idg::Matrix2x2<std::complex<float>>> &aterms,
                     idg::Array1D<idg::Metadata> &metadata, idg::Array4D<std::complex<float>> &subgrids) {
  check(idgb200_c_run_degridder(
            nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda, nr_channels, nr_stations,
            (int64_t)uvw.size(), (int)aterms.get_w_dim(), reinterpret_cast<const idgb200_uvw *>(uvw.data()),
            wavenumbers.data(), reinterpret_cast<idgb200_cfloat *>(visibilities.data()), spheroidal.data(),
            reinterpret_cast<const idgb200_cfloat *>(aterms.data()),
            reinterpret_cast<const idgb200_metadata *>(metadata.data()),
            reinterpret_cast<const idgb200_cfloat *>(subgrids.data())),
        "c_run_degridder");
}

}  // namespace cuda
