/* idg_b200.h — C ABI of the B200-native IDG gridder / degridder.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * Every entry point names the reference interface it replaces
 * (paths relative to the ska-sdp-idg-bench tree).
 *
 * Data layouts are the reference's, unchanged (app/common/types.hpp):
 *   uvw           idgb200_uvw      [total_timesteps]                      12 B
 *   wavenumbers   float            [nr_channels]
 *   visibilities  idgb200_cfloat   [total_timesteps][nr_channels][4]      32 B / vis
 *   spheroidal    float            [N][N]
 *   aterms        idgb200_cfloat   [nr_aterm_slots][nr_stations][N][N][4]
 *   metadata      idgb200_metadata [nr_subgrids]                          36 B
 *   subgrids      idgb200_cfloat   [nr_subgrids][4][N][N]
 * A subgrid s covers timesteps [time_offset(s), time_offset(s)+nr_timesteps(s)),
 * time_offset(s) = (metadata[s].baseline_offset - metadata[0].baseline_offset)
 *                  + metadata[s].time_offset     (gridder_reference.cpp:16-25).
 *
 * All functions return 0 on success or a negative IDGB200_E* / positive
 * cudaError_t value; idgb200_error_string() explains either.  Nothing here
 * falls back to a CPU path: without a CUDA device every compute call fails
 * with IDGB200_ENODEVICE.
 */
#ifndef IDG_B200_H_
#define IDG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDGB200_VERSION 100
#define IDGB200_NR_CORRELATIONS 4 /* app/common/parameters.hpp:3 */

/* app/common/types.hpp:19-26 (Metadata = 4 ints + Baseline + Coordinate) */
typedef struct {
  int32_t baseline_offset;
  int32_t time_offset;
  int32_t nr_timesteps;
  int32_t aterm_index;
  uint32_t station1, station2; /* types.hpp:15-17 */
  int32_t x, y, z;             /* types.hpp:11-13 */
} idgb200_metadata;

/* app/common/types.hpp:46-50 */
typedef struct {
  float u, v, w;
} idgb200_uvw;

/* std::complex<float> / float2 */
typedef struct {
  float re, im;
} idgb200_cfloat;

enum {
  IDGB200_OK = 0,
  IDGB200_EINVAL = -1,    /* bad argument (null pointer, size <= 0, odd subgrid size ...) */
  IDGB200_ENODEVICE = -2, /* no CUDA device / driver: there is no CPU fallback */
  IDGB200_EUNSUPPORTED = -3,
  IDGB200_ENOMEM = -4
};

/* How the per-(pixel, timestep, channel) phasor exp(i*phase) is evaluated.
 * |phase| reaches ~1.7e3 rad on the reference's shapes. */
enum {
  IDGB200_SINCOS_FAST = 0,    /* __sincosf: 1 FMUL + MUFU.SIN + MUFU.COS (what the reference's
                                 own best kernels use: gridder_v8.cu:143-150, degridder_v6.cu:110) */
  IDGB200_SINCOS_REDUCED = 1, /* 2-constant Cody-Waite reduction to [-pi,pi], then MUFU */
  IDGB200_SINCOS_ACCURATE = 2 /* sincosf (<= 2 ulp), FP32-pipe heavy */
};

/* Scalar arguments common to both kernels: the first seven parameters of the
 * reference's kernels (app/CUDA/kernels/gridder_v8.cu:286-291) plus tuning. */
typedef struct {
  int32_t nr_subgrids;
  int32_t grid_size;
  int32_t subgrid_size;
  float image_size;
  float w_step_in_lambda;
  int32_t nr_channels;
  int32_t nr_stations;
  int32_t sincos_mode; /* IDGB200_SINCOS_*  */
  int32_t variant;     /* 0 = default kernel; others are documented A/B variants */
  int32_t flags;       /* IDGB200_FLAG_*; 0 = the reference's behaviour */
  int32_t reserved[6];
} idgb200_params;

/* Subgrid FFT shift.  The bench's kernels store / read subgrid pixel (y, x) at [pol][y][x]
 * (gridder_reference.cpp:105-109, degridder_reference.cpp:38-46); the full IDG pipeline pairs
 * subgrid pixel ((y + N/2) mod N, (x + N/2) mod N) with offset (y, x).  With this flag the
 * gridder stores, the degridder reads, and the adder / splitter address the subgrid at the
 * shifted index; every value is unchanged.  Default off = the reference's layout. */
#define IDGB200_FLAG_FFT_SHIFT 1

/* ---- library / device ------------------------------------------------------ */
int idgb200_version(void);
const char *idgb200_error_string(int code);
/* replaces cuda::print_device_info (app/CUDA/util.cpp:17-71) */
int idgb200_print_device_info(void);
/* writes up to len-1 chars of the device name; replaces get_device_name (util.cpp:21-25) */
int idgb200_device_name(char *buf, size_t len);
int idgb200_sm_count(int *count);

/* ---- metric model: app/common/common.cpp:100-159 --------------------------- */
uint64_t idgb200_flops_gridder(uint64_t nr_channels, uint64_t nr_timesteps,
                               uint64_t nr_subgrids, uint64_t subgrid_size,
                               uint64_t nr_correlations);
uint64_t idgb200_bytes_gridder(uint64_t nr_channels, uint64_t nr_timesteps,
                               uint64_t nr_subgrids, uint64_t subgrid_size,
                               uint64_t nr_correlations);

/* ---- device-pointer API ----------------------------------------------------
 * One asynchronous launch on `stream` (a cudaStream_t, NULL = default stream)
 * of the current device.  Replaces the 13-parameter __global__ launched through
 * cudaLaunchKernel in app/CUDA/util.cpp:163-170 (c_run_kernel) / :81-128
 * (p_run_kernel).  The gridder overwrites every element of d_subgrids, the
 * degridder every visibility of every subgrid's time range. */
int idgb200_gridder(const idgb200_params *params, const idgb200_uvw *d_uvw,
                    const float *d_wavenumbers,
                    const idgb200_cfloat *d_visibilities,
                    const float *d_spheroidal, const idgb200_cfloat *d_aterms,
                    const idgb200_metadata *d_metadata,
                    idgb200_cfloat *d_subgrids, void *stream);

int idgb200_degridder(const idgb200_params *params, const idgb200_uvw *d_uvw,
                      const float *d_wavenumbers,
                      idgb200_cfloat *d_visibilities, const float *d_spheroidal,
                      const idgb200_cfloat *d_aterms,
                      const idgb200_metadata *d_metadata,
                      const idgb200_cfloat *d_subgrids, void *stream);

/* Number of KERNELS this library has put on a stream in this process, whatever the entry point
 * (gridder, degridder, adder, splitter, reduce_parts, subgrid FFT, the host-pointer and performance
 * runs): a call that enqueues two kernels counts two.  The idgb200_init_* generators are not counted.
 * bench.py reports the difference over its timed region as gpu_launches. */
uint64_t idgb200_launch_count(void);

/* ---- "next" rows: grid adder, subgrid FFT, splitter ------------
 * Not in the reference (it only declares idg::Grid, app/common/types.hpp:358-370,
 * and the subgrid coordinate, types.hpp:11-13); they surround the two kernels in IDG:
 *   gridder -> subgrid FFT -> adder -> [grid] -> splitter -> inverse FFT -> degridder.
 *
 * Adder:    grid[pol][y0 + y][x0 + x] += d_subgrids[s][pol][y'][x']  ((x0, y0) = metadata[s].x/.y,
 *           pixels outside the grid dropped), for the params->nr_subgrids subgrids.
 * Splitter: d_subgrids[s][pol][y'][x'] = grid[pol][y0 + y][x0 + x], 0 outside the grid.
 *           (y', x') = (y, x), or the shifted index with IDGB200_FLAG_FFT_SHIFT in params->flags.
 * The grid is cut into nr_parts (<= 16) blocks of rows_per_part rows; grid_parts is
 * a HOST array of nr_parts device pointers, part r laid out as complex64
 * [4][rows_per_part][grid_size].  One GPU: nr_parts = 1, rows_per_part = grid_size.
 * N GPUs: part r may be a peer (NVLink) address of the slice rank r owns - the
 * adder's atomics then are the reduce-scatter, the splitter's loads the all-gather.
 * The adder accumulates (the caller zeroes); the splitter overwrites every subgrid pixel.
 * Alignment: 8 bytes (one complex64) is enough for every pointer; 16-byte aligned subgrids / grid parts
 * get the 16-byte stores / reductions. */
int idgb200_adder(const idgb200_params *params, const idgb200_metadata *d_metadata,
                  const idgb200_cfloat *d_subgrids, idgb200_cfloat *const *grid_parts,
                  int nr_parts, int rows_per_part, void *stream);
int idgb200_splitter(const idgb200_params *params, const idgb200_metadata *d_metadata,
                     idgb200_cfloat *d_subgrids, const idgb200_cfloat *const *grid_parts,
                     int nr_parts, int rows_per_part, void *stream);

/* Multi-GPU adder, second half: out[i] = sources[0][i] + ... + sources[nr_sources-1][i], summed in
 * that order (bit-reproducible), count complex64 elements (even; every pointer 16-byte aligned).
 * sources is a HOST array of device pointers that may be peer (NVLink) addresses: rank r calls it
 * with every rank's local copy of grid part r and ends up owning the reduced part r - a
 * reduce-scatter by direct peer loads (tools/adder_reduce_scatter.py compares it with
 * ncclReduceScatter and with the adder's reductions into peer memory). nr_sources <= 16. */
int idgb200_reduce_parts(int nr_sources, const idgb200_cfloat *const *sources, int64_t count,
                         idgb200_cfloat *d_out, void *stream);

/* How N GPUs should add their shards of the subgrids into one row-scattered grid (measured on 2 and 8 B200s,
 * DESIGN.md 4.8): 1 = push - every rank's idgb200_adder reduces straight into the owners' slices through peer
 * addresses (grid_parts = the owners' slices) - when a rank has fewer subgrid pixels than the grid has cells
 * (nr_subgrids * N^2 < G^2); 0 = local grid + reduce-scatter (idgb200_reduce_parts over peer addresses, or
 * ncclReduceScatter).  ska_sdp_idg_bench_b200/grid_adder_rs.py implements both on symmetric memory. */
int idgb200_adder_rs_mode(int64_t nr_subgrids, int subgrid_size, int grid_size);

/* Subgrid FFT: in-place 2-D DFT of each of the nr_subgrids * 4 planes of N x N pixels.
 *   direction = +1 (forward, after the gridder):  B[ky][kx] = sum A[y][x] exp(-2 pi i (ky y + kx x) / N)
 *   direction = -1 (backward, before the degridder): exp(+...), scaled by 1 / N^2
 * so that backward(forward(A)) = A.  Any N >= 1; N = 8, 16, 24, 32, 48, 64 take the
 * register-resident kernel, other sizes a direct DFT. */
int idgb200_subgrid_fft(int64_t nr_subgrids, int subgrid_size, int direction,
                        idgb200_cfloat *d_subgrids, void *stream);

/* The kernel variant params->variant == 0 resolves to for this shape and sincos
 * mode (gridder != 0: the gridder, else the degridder); a non-zero variant is
 * returned unchanged.  FAST sincos: 30 = the row-column tcgen05 kernels (gridder:
 * subgrid sizes that are a multiple of 4; degridder: up to 64 x 64 pixels), which
 * leave subgrids whose phase does not separate - and, for the gridder, subgrids
 * whose sums cancel below the fp16 operand's error model - to the per-pixel / FP32
 * kernel launched behind them on the same stream; other FAST shapes: gridder
 * 24 / 21, degridder 24 / 22 = the per-pixel tcgen05 kernels where the shape fills
 * their tiles.  Other sincos modes and remaining shapes: gridder 10, degridder 4 =
 * the FP32 kernels.  IDGB200_EINVAL on bad params. */
int idgb200_resolve_variant(const idgb200_params *params, int gridder);

/* ---- host-pointer API --------------------------------------------------------
 * Replaces cuda::c_run_gridder_ / c_run_degridder_ (app/CUDA/util.cpp:251-307,
 * 388-444): device allocation, host->device copies, the kernel, device->host
 * copy of the result, free.  Unlike the reference the subgrid list is cut into
 * chunks that are copied and computed on three streams so that H2D, kernel and
 * D2H overlap; pinned host buffers (idgb200_host_alloc) make the copies
 * asynchronous.  total_timesteps / nr_aterm_slots are the array extents the
 * reference reads off its Array objects. */
int idgb200_c_run_gridder(int nr_subgrids, int grid_size, int subgrid_size,
                          float image_size, float w_step_in_lambda,
                          int nr_channels, int nr_stations,
                          int64_t total_timesteps, int nr_aterm_slots,
                          const idgb200_uvw *uvw, const float *wavenumbers,
                          const idgb200_cfloat *visibilities,
                          const float *spheroidal, const idgb200_cfloat *aterms,
                          const idgb200_metadata *metadata,
                          idgb200_cfloat *subgrids);

int idgb200_c_run_degridder(int nr_subgrids, int grid_size, int subgrid_size,
                            float image_size, float w_step_in_lambda,
                            int nr_channels, int nr_stations,
                            int64_t total_timesteps, int nr_aterm_slots,
                            const idgb200_uvw *uvw, const float *wavenumbers,
                            idgb200_cfloat *visibilities,
                            const float *spheroidal,
                            const idgb200_cfloat *aterms,
                            const idgb200_metadata *metadata,
                            const idgb200_cfloat *subgrids);

/* Same, with explicit sincos mode / kernel variant (the two calls above use
 * IDGB200_SINCOS_FAST, variant 0, or the IDGB200_SINCOS / IDGB200_VARIANT
 * environment variables when set). */
int idgb200_c_run_gridder_ex(const idgb200_params *params, int64_t total_timesteps,
                             int nr_aterm_slots, const idgb200_uvw *uvw,
                             const float *wavenumbers,
                             const idgb200_cfloat *visibilities,
                             const float *spheroidal,
                             const idgb200_cfloat *aterms,
                             const idgb200_metadata *metadata,
                             idgb200_cfloat *subgrids);

int idgb200_c_run_degridder_ex(const idgb200_params *params, int64_t total_timesteps,
                               int nr_aterm_slots, const idgb200_uvw *uvw,
                               const float *wavenumbers,
                               idgb200_cfloat *visibilities,
                               const float *spheroidal,
                               const idgb200_cfloat *aterms,
                               const idgb200_metadata *metadata,
                               const idgb200_cfloat *subgrids);

/* Pinned host memory for the host-pointer API (cudaHostAlloc / cudaFreeHost). */
int idgb200_host_alloc(void **ptr, size_t bytes);
int idgb200_host_free(void *ptr);

/* ---- performance runs --------------------------------------------------------
 * Replace cuda::p_run_gridder_ / p_run_degridder_ + p_run_kernel
 * (app/CUDA/util.cpp:172-249, 309-386, 81-161).  Shape from the same
 * environment variables (GRID_SIZE, SUBGRID_SIZE, NR_STATIONS, NR_TIMESLOTS,
 * NR_TIMESTEPS_SUBGRID, NR_CHANNELS, NR_WARM_UP_RUNS, NR_ITERATIONS), the same
 * report line on stdout.  Unlike the reference all device buffers are
 * initialised (idgb200_init_* below) before timing.  IDGB200_ENERGY_SECONDS (default 2,
 * 0 = off): how long the kernel is relaunched for the energy columns. */
typedef struct {
  double seconds;     /* mean kernel time over the timed iterations (CUDA events) */
  double gflops;      /* flops_gridder * 1e-9 (work per launch, not a rate)       */
  double gbytes;      /* bytes_gridder * 1e-9                                     */
  double mvis;        /* 1e-6 * total_timesteps * nr_channels                     */
  int32_t nr_subgrids;
  int32_t iterations;
  double joules;      /* energy per launch from the GPU's NVML counter (0 = not available); replaces the
                         PowerSensor measurement of util.cpp:131-155: W = joules / seconds,
                         GFLOP/s/W = gflops / joules, MVis/J = mvis / joules (common.cpp:47-54) */
} idgb200_perf;

/* The reference's report line and key,value CSV (app/common/common.cpp:27-56 and :58-98; called by its
 * runners at app/CUDA/util.cpp:157-160): same text, keys, order and number format, byte for byte
 * (tests/test_host_logic.py pins both against the reference's own functions).  seconds = time per launch,
 * gflops / gbytes / mvis = work per launch, joules = energy per launch (0 = columns left out).  The CSV goes
 * to $OUTPUT_PATH (default ".") as <device_name with '/' -> '-'>-<name><file_extension>. */
void idgb200_report(const char *name, double seconds, double gflops, double gbytes, double mvis, double joules);
void idgb200_report_csv(const char *name, const char *device_name, const char *file_extension, double seconds,
                        double gflops, double gbytes, double mvis, double joules);

int idgb200_p_run_gridder(idgb200_perf *result /* may be NULL */);
int idgb200_p_run_degridder(idgb200_perf *result /* may be NULL */);

/* ---- synthetic inputs on the device ------------------------------------------
 * The distributions of app/common/init.cpp evaluated by kernels, so that
 * benchmark-size inputs (GBs) never cross PCIe.  rand() is replaced by a
 * counter-based hash of (seed, index); everything else follows init.cpp.
 * All are asynchronous on `stream`. */
int idgb200_init_uvw(uint32_t grid_size, int64_t nr_baselines, int nr_timesteps,
                     uint32_t seed, idgb200_uvw *d_uvw, void *stream);          /* init.cpp:4-25   */
int idgb200_init_wavenumbers(int nr_channels, float *d_wavenumbers, void *stream); /* :27-46      */
int idgb200_init_visibilities(uint32_t grid_size, float image_size,
                              int64_t total_timesteps, int nr_channels,
                              const idgb200_uvw *d_uvw,
                              idgb200_cfloat *d_visibilities, void *stream);   /* :48-79          */
int idgb200_init_spheroidal(int subgrid_size, float *d_spheroidal, void *stream); /* :97-107     */
int idgb200_init_aterms(int nr_slots, int nr_stations, int subgrid_size,
                        uint32_t seed, idgb200_cfloat *d_aterms, void *stream); /* :109-132       */
int idgb200_init_metadata(uint32_t grid_size, int nr_stations, int nr_timeslots,
                          int nr_timesteps_subgrid, int per_slot_aterms,
                          uint32_t seed, idgb200_metadata *d_metadata,
                          void *stream);                                        /* :81-95,134-159 */
int idgb200_init_subgrids(int64_t nr_subgrids, int subgrid_size,
                          idgb200_cfloat *d_subgrids, void *stream);            /* :161-180        */

#ifdef __cplusplus
}
#endif
#endif /* IDG_B200_H_ */
