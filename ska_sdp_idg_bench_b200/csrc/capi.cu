// extern "C" layer of libidgb200.so (declared in include/idg_b200.h): argument
// validation, the device-pointer launches, the pipelined host-pointer runs, the
// env-driven performance runs and the device-side synthetic inputs.
//
// There is deliberately no CPU code path here: without a CUDA device every
// compute entry point returns IDGB200_ENODEVICE.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "idg_b200.h"
#include "kernels.h"

using namespace idgb200;

static_assert(sizeof(idgb200_metadata) == 36, "Metadata ABI (types.hpp:19-26)");
static_assert(sizeof(idgb200_uvw) == 12, "UVWCoordinate<float> ABI (types.hpp:46-50)");
static_assert(sizeof(idgb200_cfloat) == 8, "std::complex<float> ABI");

namespace {

std::atomic<uint64_t> g_launches{0};

#define CK(expr)                          \
  do {                                    \
    cudaError_t e__ = (expr);             \
    if (e__ != cudaSuccess) return (int)e__; \
  } while (0)

int have_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return IDGB200_ENODEVICE;
  }
  return IDGB200_OK;
}

long env_long(const char *name, long dflt) {
  // same semantics as get_env_var (app/common/common.cpp:10-17): atoi of the value
  const char *v = std::getenv(name);
  return v ? std::atol(v) : dflt;
}

int check_params(const idgb200_params *p) {
  if (!p) return IDGB200_EINVAL;
  if (p->nr_subgrids < 0 || p->subgrid_size <= 0 || p->grid_size <= 0 || p->nr_channels <= 0 ||
      p->nr_stations <= 0)
    return IDGB200_EINVAL;
  if (!(p->image_size > 0.0f)) return IDGB200_EINVAL;
  if (p->sincos_mode < 0 || p->sincos_mode > IDGB200_SINCOS_ACCURATE) return IDGB200_EINVAL;
  if (p->flags & ~IDGB200_FLAG_FFT_SHIFT) return IDGB200_EINVAL;
  // the shifted slot and the adder / splitter rows are found with a float reciprocal that is exact for
  // plane sizes below 2^22 pixels (common.cuh: subgrid_slot, adder.cu: div_small)
  if ((p->flags & IDGB200_FLAG_FFT_SHIFT) && p->subgrid_size > 1024) return IDGB200_EUNSUPPORTED;
  return IDGB200_OK;
}

KernelArgs make_args(const idgb200_params *p, const idgb200_uvw *uvw, const float *wn,
                     const idgb200_cfloat *vis, const float *sph, const idgb200_cfloat *at,
                     const idgb200_metadata *meta, const idgb200_cfloat *sg) {
  KernelArgs a;
  a.grid_size = p->grid_size;
  a.subgrid_size = p->subgrid_size;
  a.image_size = p->image_size;
  a.w_step_in_lambda = p->w_step_in_lambda;
  a.nr_channels = p->nr_channels;
  a.nr_stations = p->nr_stations;
  a.uvw = uvw;
  a.wavenumbers = wn;
  a.visibilities = reinterpret_cast<const float2 *>(vis);
  a.spheroidal = sph;
  a.aterms = reinterpret_cast<const float2 *>(at);
  a.metadata = meta;
  a.subgrids = reinterpret_cast<const float2 *>(sg);
  a.nr_subgrids = p->nr_subgrids;
  a.subgrid_offset = 0;
  a.flags = p->flags;
  a.list = nullptr;
  return a;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------- workspace
// Device buffers of the host-pointer API are kept between calls (grow-only) so a
// caller that degrids / grids repeatedly does not pay cudaMalloc/cudaFree each
// time as the reference does (util.cpp:270-276, 300-306).
struct Workspace {
  void *buf[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t cap[7] = {0, 0, 0, 0, 0, 0, 0};
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  int device = -1;
  std::mutex mu;

  int ensure_streams() {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (device != dev) {
      release();
      device = dev;
    }
    if (!s_in) {
      CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
      CK(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
      CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    }
    return 0;
  }
  int reserve(int i, size_t bytes) {
    if (bytes <= cap[i]) return 0;
    if (buf[i]) CK(cudaFree(buf[i]));
    buf[i] = nullptr;
    cap[i] = 0;
    cudaError_t e = cudaMalloc(&buf[i], bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return e == cudaErrorMemoryAllocation ? IDGB200_ENOMEM : (int)e;
    }
    cap[i] = bytes;
    return 0;
  }
  void release() {
    for (int i = 0; i < 7; i++) {
      if (buf[i]) cudaFree(buf[i]);
      buf[i] = nullptr;
      cap[i] = 0;
    }
    if (s_in) cudaStreamDestroy(s_in), s_in = nullptr;
    if (s_k) cudaStreamDestroy(s_k), s_k = nullptr;
    if (s_out) cudaStreamDestroy(s_out), s_out = nullptr;
  }
};
Workspace g_ws;

enum { B_UVW, B_WN, B_VIS, B_SPH, B_AT, B_META, B_SG };

// time range [t0, t1) touched by metadata[s0, s1)
void time_range(const idgb200_metadata *meta, int s0, int s1, int64_t *t0, int64_t *t1) {
  int64_t lo = INT64_MAX, hi = 0;
  const int64_t base0 = meta[0].baseline_offset;
  for (int s = s0; s < s1; s++) {
    const int64_t b = (int64_t)meta[s].baseline_offset - base0 + meta[s].time_offset;
    if (meta[s].nr_timesteps <= 0) continue;
    lo = b < lo ? b : lo;
    hi = b + meta[s].nr_timesteps > hi ? b + meta[s].nr_timesteps : hi;
  }
  if (lo == INT64_MAX) lo = hi = 0;
  *t0 = lo;
  *t1 = hi;
}

// The pipelined host-pointer run shared by gridder and degridder.
//   gridder:   per chunk  H2D(uvw, vis) -> kernel -> D2H(subgrids)
//   degridder: per chunk  H2D(uvw, subgrids) -> kernel -> D2H(vis)
int host_run(bool gridding, const idgb200_params *p, int64_t total_timesteps, int nr_aterm_slots,
             const idgb200_uvw *uvw, const float *wn, idgb200_cfloat *vis, const float *sph,
             const idgb200_cfloat *at, const idgb200_metadata *meta, idgb200_cfloat *sg) {
  int rc = check_params(p);
  if (rc) return rc;
  if (!uvw || !wn || !vis || !sph || !at || !meta || !sg) return IDGB200_EINVAL;
  if (total_timesteps < 0 || nr_aterm_slots <= 0) return IDGB200_EINVAL;
  if ((rc = have_device())) return rc;
  const int S = p->nr_subgrids;
  if (S == 0) return IDGB200_OK;

  const int N = p->subgrid_size, C = p->nr_channels;
  const size_t npix = (size_t)N * N;
  // validate the metadata against the array extents (the reference trusts it)
  for (int s = 0; s < S; s++) {
    const int64_t b = (int64_t)meta[s].baseline_offset - meta[0].baseline_offset + meta[s].time_offset;
    if (meta[s].nr_timesteps < 0 || b < 0 || b + meta[s].nr_timesteps > total_timesteps ||
        meta[s].aterm_index < 0 || meta[s].aterm_index >= nr_aterm_slots ||
        meta[s].station1 >= (uint32_t)p->nr_stations || meta[s].station2 >= (uint32_t)p->nr_stations)
      return IDGB200_EINVAL;
  }

  std::lock_guard<std::mutex> lock(g_ws.mu);
  if ((rc = g_ws.ensure_streams())) return rc;
  const size_t tt = (size_t)(total_timesteps > 0 ? total_timesteps : 1);
  if ((rc = g_ws.reserve(B_UVW, tt * sizeof(idgb200_uvw)))) return rc;
  if ((rc = g_ws.reserve(B_WN, (size_t)C * sizeof(float)))) return rc;
  if ((rc = g_ws.reserve(B_VIS, tt * C * NR_POL * sizeof(float2)))) return rc;
  if ((rc = g_ws.reserve(B_SPH, npix * sizeof(float)))) return rc;
  if ((rc = g_ws.reserve(B_AT, (size_t)nr_aterm_slots * p->nr_stations * npix * NR_POL * sizeof(float2))))
    return rc;
  if ((rc = g_ws.reserve(B_META, (size_t)S * sizeof(idgb200_metadata)))) return rc;
  if ((rc = g_ws.reserve(B_SG, (size_t)S * NR_POL * npix * sizeof(float2)))) return rc;

  auto *d_uvw = static_cast<idgb200_uvw *>(g_ws.buf[B_UVW]);
  auto *d_wn = static_cast<float *>(g_ws.buf[B_WN]);
  auto *d_vis = static_cast<idgb200_cfloat *>(g_ws.buf[B_VIS]);
  auto *d_sph = static_cast<float *>(g_ws.buf[B_SPH]);
  auto *d_at = static_cast<idgb200_cfloat *>(g_ws.buf[B_AT]);
  auto *d_meta = static_cast<idgb200_metadata *>(g_ws.buf[B_META]);
  auto *d_sg = static_cast<idgb200_cfloat *>(g_ws.buf[B_SG]);
  cudaStream_t s_in = g_ws.s_in, s_k = g_ws.s_k, s_out = g_ws.s_out;

  // shared read-only inputs
  CK(cudaMemcpyAsync(d_wn, wn, (size_t)C * sizeof(float), cudaMemcpyHostToDevice, s_in));
  CK(cudaMemcpyAsync(d_sph, sph, npix * sizeof(float), cudaMemcpyHostToDevice, s_in));
  CK(cudaMemcpyAsync(d_at, at, (size_t)nr_aterm_slots * p->nr_stations * npix * NR_POL * sizeof(float2),
                     cudaMemcpyHostToDevice, s_in));
  CK(cudaMemcpyAsync(d_meta, meta, (size_t)S * sizeof(idgb200_metadata), cudaMemcpyHostToDevice, s_in));

  // chunking: ~IDGB200_CHUNKS chunks (default 16) of at least 256 subgrids in the body of the run;
  // the first chunks ramp up from 256 subgrids by doubling and the last ones ramp down the same way,
  // so that the copy that nothing overlaps (first upload, last kernel + download) is short
  long want = env_long("IDGB200_CHUNKS", 16);
  if (want < 1) want = 1;
  int chunk = (int)((S + want - 1) / want);
  if (chunk < 256) chunk = 256;
  if (env_long("IDGB200_CHUNK_SUBGRIDS", 0) > 0) chunk = (int)env_long("IDGB200_CHUNK_SUBGRIDS", 0);
  std::vector<int> bounds;   // chunk i = subgrids [bounds[i], bounds[i + 1])
  {
    std::vector<int> head, tail;
    int lo = 0, hi = S;
    if (env_long("IDGB200_CHUNK_RAMP", 1) != 0)
      for (int c = 256; c < chunk && hi - lo > 2 * chunk; c *= 2) {
        head.push_back(lo + c);
        lo += c;
        tail.push_back(hi - c);
        hi -= c;
      }
    bounds.push_back(0);
    for (int b : head) bounds.push_back(b);
    for (int b = lo + chunk; b < hi; b += chunk) bounds.push_back(b);
    for (int i = (int)tail.size() - 1; i >= 0; i--)
      if (tail[i] > bounds.back()) bounds.push_back(tail[i]);
    if (bounds.back() != S) bounds.push_back(S);
  }
  const int nchunks = (int)bounds.size() - 1;
  std::vector<cudaEvent_t> ev_in(nchunks), ev_k(nchunks);
  for (int i = 0; i < nchunks; i++) {
    CK(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
  }

  KernelArgs a = make_args(p, d_uvw, d_wn, d_vis, d_sph, d_at, d_meta, d_sg);
  int status = 0;
  if (!gridding)  // rows no subgrid covers come back as zeros, not as stale device memory
    CK(cudaMemsetAsync(d_vis, 0, tt * C * NR_POL * sizeof(float2), s_k));
  // uvw / vis rows [up_lo, up_hi) are on the device (or on their way on s_in).
  // Metadata is normally time-ordered, so each chunk just extends the hull by its
  // own rows; for unordered metadata the hull also swallows the gap in between.
  int64_t up_lo = 0, up_hi = 0;
  auto upload = [&](int64_t r0, int64_t r1) -> cudaError_t {
    if (r1 <= r0) return cudaSuccess;
    cudaError_t e = cudaMemcpyAsync(d_uvw + r0, uvw + r0, (size_t)(r1 - r0) * sizeof(idgb200_uvw),
                                    cudaMemcpyHostToDevice, s_in);
    if (e == cudaSuccess && gridding)
      e = cudaMemcpyAsync(d_vis + (size_t)r0 * C * NR_POL, vis + (size_t)r0 * C * NR_POL,
                          (size_t)(r1 - r0) * C * NR_POL * sizeof(float2), cudaMemcpyHostToDevice, s_in);
    return e;
  };
  int64_t down_hi = 0;
  bool unordered = false;
  auto download = [&](int64_t r0, int64_t r1) -> cudaError_t {
    if (r1 <= r0) return cudaSuccess;
    return cudaMemcpyAsync(vis + (size_t)r0 * C * NR_POL, d_vis + (size_t)r0 * C * NR_POL,
                           (size_t)(r1 - r0) * C * NR_POL * sizeof(float2), cudaMemcpyDeviceToHost, s_out);
  };
  for (int i = 0; i < nchunks && !status; i++) {
    const int s0 = bounds[i], s1 = bounds[i + 1];
    int64_t t0, t1;
    time_range(meta, s0, s1, &t0, &t1);
    if (t1 > t0) {
      cudaError_t e = cudaSuccess;
      if (up_hi == up_lo) {
        e = upload(t0, t1);
        up_lo = t0;
        up_hi = t1;
      } else {
        if (t0 < up_lo) { e = upload(t0, up_lo); up_lo = t0; }
        if (e == cudaSuccess && t1 > up_hi) { e = upload(up_hi, t1); up_hi = t1; }
      }
      if (e != cudaSuccess) { status = (int)e; break; }
    }
    if (!gridding) {
      cudaError_t e = cudaMemcpyAsync(d_sg + (size_t)s0 * NR_POL * npix, sg + (size_t)s0 * NR_POL * npix,
                                      (size_t)(s1 - s0) * NR_POL * npix * sizeof(float2),
                                      cudaMemcpyHostToDevice, s_in);
      if (e != cudaSuccess) { status = (int)e; break; }
    }
    cudaEventRecord(ev_in[i], s_in);
    cudaStreamWaitEvent(s_k, ev_in[i], 0);

    a.nr_subgrids = s1 - s0;
    a.subgrid_offset = s0;
    int nk = 0;
    cudaError_t e = gridding ? launch_gridder(a, p->sincos_mode, p->variant, s_k, &nk)
                             : launch_degridder(a, p->sincos_mode, p->variant, s_k, &nk);
    if (e != cudaSuccess) { status = (int)e; break; }
    g_launches += nk;
    cudaEventRecord(ev_k[i], s_k);
    cudaStreamWaitEvent(s_out, ev_k[i], 0);

    if (gridding) {
      e = cudaMemcpyAsync(sg + (size_t)s0 * NR_POL * npix, d_sg + (size_t)s0 * NR_POL * npix,
                          (size_t)(s1 - s0) * NR_POL * npix * sizeof(float2), cudaMemcpyDeviceToHost, s_out);
    } else if (t1 > t0) {
      // time-ordered metadata: rows [down_hi, t1) are final once this chunk is done
      // (including never-covered rows, which the memset left at zero); otherwise
      // everything is fetched in one go after the last kernel
      if (!unordered && t0 >= down_hi) {
        e = download(down_hi, t1);
        down_hi = t1;
      } else {
        unordered = true;
      }
    }
    if (e != cudaSuccess) { status = (int)e; break; }
  }
  if (!gridding && !status) {  // the whole visibility array is (re)written exactly once
    cudaError_t e = unordered ? download(0, total_timesteps) : download(down_hi, total_timesteps);
    if (e != cudaSuccess) status = (int)e;
  }
  cudaError_t e1 = cudaStreamSynchronize(s_in);
  cudaError_t e2 = cudaStreamSynchronize(s_k);
  cudaError_t e3 = cudaStreamSynchronize(s_out);
  for (int i = 0; i < nchunks; i++) {
    cudaEventDestroy(ev_in[i]);
    cudaEventDestroy(ev_k[i]);
  }
  if (status) return status;
  if (e1 != cudaSuccess) return (int)e1;
  if (e2 != cudaSuccess) return (int)e2;
  if (e3 != cudaSuccess) return (int)e3;
  return IDGB200_OK;
}

idgb200_params env_params(int nr_subgrids, int grid_size, int subgrid_size, float image_size,
                          float w_step, int nr_channels, int nr_stations) {
  idgb200_params p;
  std::memset(&p, 0, sizeof p);
  p.nr_subgrids = nr_subgrids;
  p.grid_size = grid_size;
  p.subgrid_size = subgrid_size;
  p.image_size = image_size;
  p.w_step_in_lambda = w_step;
  p.nr_channels = nr_channels;
  p.nr_stations = nr_stations;
  p.sincos_mode = (int)env_long("IDGB200_SINCOS", IDGB200_SINCOS_FAST);
  p.variant = (int)env_long("IDGB200_VARIANT", 0);
  return p;
}

// ------------------------------------------------------ synthetic input kernels
__device__ __forceinline__ uint32_t mix32(uint32_t x) {  // lowbias32
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
// uniform in [0,1], replaces (double)rand() / RAND_MAX of init.cpp
__device__ __forceinline__ double u01(uint32_t seed, uint32_t stream, uint64_t i) {
  const uint32_t h = mix32((uint32_t)i ^ mix32((uint32_t)(i >> 32) + 0x9e3779b9U * (stream + 1) + seed));
  return (double)h / 4294967295.0;
}

__global__ void k_init_uvw(uint32_t grid_size, int64_t nr_baselines, int nr_timesteps, uint32_t seed,
                           idgb200_uvw *uvw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nr_baselines * nr_timesteps) return;
  const int64_t bl = i / nr_timesteps;
  const unsigned time = (unsigned)(i - bl * nr_timesteps);
  const float radius_u = (float)((grid_size / 2) + u01(seed, 0, (uint64_t)bl) * (grid_size / 2));
  const float radius_v = (float)((grid_size / 2) + u01(seed, 1, (uint64_t)bl) * (grid_size / 2));
  const float angle = (float)((time + 0.5) / (double)(360.0f / (float)(unsigned)nr_timesteps));
  const double pi = 3.14159265358979323846;
  idgb200_uvw p;
  p.u = (float)((double)radius_u * cos((double)angle * pi));
  p.v = (float)((double)radius_v * sin((double)angle * pi));
  p.w = 0.0f;
  uvw[i] = p;
}

__device__ __forceinline__ float frequency_of(int chan) {  // init.cpp:27-36 (float arithmetic)
  return (float)150000000u + 0.7e6f * (float)(unsigned)chan;
}

__global__ void k_init_wavenumbers(int nr_channels, float *wn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nr_channels) return;
  wn[i] = (float)(2.0 * 3.14159265358979323846 * (double)frequency_of(i) / 299792458.0);
}

__global__ void k_init_vis(uint32_t grid_size, float image_size, int64_t total_timesteps, int nr_channels,
                           const idgb200_uvw *uvw, float4 *vis) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_timesteps * nr_channels) return;
  const int64_t t = i / nr_channels;
  const int chan = (int)(i - t * nr_channels);
  const float x_offset = (float)(0.6 * grid_size), y_offset = (float)(0.7 * grid_size);
  const float l = x_offset * image_size / (float)grid_size;
  const float m = y_offset * image_size / (float)grid_size;
  const double f_over_c = (double)frequency_of(chan) / 299792458.0;
  const float u = (float)(f_over_c * uvw[t].u), v = (float)(f_over_c * uvw[t].v);
  const float arg = (float)(-2.0 * 3.14159265358979323846 * (double)fmaf(u, l, v * m));
  float sn, cs;
  sincosf(arg, &sn, &cs);
  vis[2 * i + 0] = make_float4(cs * 1.01f, sn * 1.01f, cs * 1.02f, sn * 1.02f);
  vis[2 * i + 1] = make_float4(cs * 1.03f, sn * 1.03f, cs * 1.04f, sn * 1.04f);
}

__device__ __forceinline__ float spheroidal_at(int y, int x, int N) {  // init.cpp:97-107
  const float ty = fabsf(-1 + (unsigned)y * 2.0f / (float)N);
  const float tx = fabsf(-1 + (unsigned)x * 2.0f / (float)N);
  return ty * tx;
}

__global__ void k_init_sph(int N, float *sph) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * N) return;
  sph[i] = spheroidal_at(i / N, i % N, N);
}

__global__ void k_init_aterms(int64_t count, int N, uint32_t seed, float4 *at) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (slot, station, y, x)
  if (i >= count) return;
  const int pix = (int)(i % ((int64_t)N * N));
  const float scale = (float)(0.8 + u01(seed, 2, (uint64_t)i) * 0.4);
  const float value = spheroidal_at(pix / N, pix % N, N) * scale;
  const float d = (float)(value + 0.1), o = (float)(value - 0.2);
  at[2 * i + 0] = make_float4(d, -0.1f, o, 0.1f);
  at[2 * i + 1] = make_float4(o, 0.1f, d, -0.1f);
}

__global__ void k_init_metadata(uint32_t grid_size, int nr_stations, int64_t nr_baselines, int nr_timeslots,
                                int nr_timesteps_subgrid, int per_slot_aterms, uint32_t seed,
                                idgb200_metadata *meta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nr_baselines * nr_timeslots) return;
  const int64_t bl = i / nr_timeslots;
  const int ts = (int)(i - bl * nr_timeslots);
  // baseline enumeration of initialize_baselines (init.cpp:81-95)
  int64_t rem = bl;
  int s1 = 0;
  while (rem >= nr_stations - 1 - s1) { rem -= nr_stations - 1 - s1; s1++; }
  idgb200_metadata m;
  m.baseline_offset = 0;
  m.time_offset = (int)(bl * nr_timeslots * nr_timesteps_subgrid + (int64_t)ts * nr_timesteps_subgrid);
  m.nr_timesteps = nr_timesteps_subgrid;
  m.aterm_index = per_slot_aterms ? ts : 0;
  m.station1 = (uint32_t)s1;
  m.station2 = (uint32_t)(s1 + 1 + rem);
  m.x = (int)(u01(seed, 3, (uint64_t)i) * grid_size);
  m.y = (int)(u01(seed, 4, (uint64_t)i) * grid_size);
  if (m.x >= (int)grid_size) m.x = grid_size - 1;
  if (m.y >= (int)grid_size) m.y = grid_size - 1;
  m.z = 0;
  meta[i] = m;
}

__global__ void k_init_subgrids(int64_t count, int N, float2 *sg) {  // init.cpp:161-180
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (s, c, y, x)
  if (i >= count) return;
  const unsigned npix = (unsigned)N * N;
  const unsigned pix = (unsigned)(i % npix);
  const unsigned c = (unsigned)((i / npix) % NR_POL);
  sg[i] = make_float2((float)(pix + 1) / ((float)100 * (float)N * (float)N), (float)c / 10.0f);
}

unsigned blocks_for(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

// report line of the reference (app/common/common.cpp:27-56)
void report(const char *name, double seconds, double gflops, double gbytes, double mvis, double joules) {
  std::printf("%20s: %7.2f ms", name, seconds * 1e3);
  if (gflops != 0) std::printf(", %7.2f GFLOP/s", gflops / seconds);
  if (gbytes != 0) std::printf(", %7.2f GB/s", gbytes / seconds);
  if (gflops != 0 && gbytes != 0) std::printf(", %7.2f FLOP/byte", (double)(float)(gflops / gbytes));
  if (mvis != 0) std::printf(", %7.2f MVis/s", mvis / seconds);
  if (joules != 0)   // common.cpp:47-54
    std::printf(", %7.2f W, %7.2f GFLOP/s/W, %7.2f MVis/J", joules / seconds, gflops / joules, mvis / joules);
  std::printf("\n");
  std::fflush(stdout);
}

// key,value CSV of the reference (app/common/common.cpp:58-98): same file name, keys, order and number format
void report_csv(const char *name, const char *device_name, const char *file_extension, double seconds, double gflops,
                double gbytes, double mvis, double joules) {
  if (!device_name || !*device_name || !file_extension || !*file_extension) {
    std::printf(">>> Device name or file extension not provided\n");
    return;
  }
  const char *dir = std::getenv("OUTPUT_PATH");
  const std::string file_path = dir ? dir : ".";
  std::printf("Saving output in %s\n", file_path.c_str());
  std::string dev = device_name;
  for (char &c : dev)
    if (c == '/') c = '-';
  const std::string path = file_path + "/" + dev + "-" + name + file_extension;
  std::printf("%s\n", path.c_str());
  std::fflush(stdout);
  if (FILE *f = std::fopen(path.c_str(), "w")) {
    std::fprintf(f, "ms,%.2f\n", seconds * 1e3);
    if (gflops != 0) std::fprintf(f, "GFLOP/s,%.2f\n", gflops / seconds);
    if (gbytes != 0) std::fprintf(f, "GB/s,%.2f\n", gbytes / seconds);
    if (gflops != 0 && gbytes != 0) std::fprintf(f, "FLOP/Byte,%.2f\n", (double)(float)(gflops / gbytes));
    if (mvis != 0) std::fprintf(f, "MVis/s,%.2f\n", mvis / seconds);
    if (joules != 0) {   // common.cpp:88-95
      std::fprintf(f, "W,%.2f\n", joules / seconds);
      std::fprintf(f, "GFLOP/s/W,%.2f\n", gflops / joules);
      std::fprintf(f, "MVis/J,%.2f\n", mvis / joules);
    }
    std::fclose(f);
  }
}

// Energy per launch.  The reference measures it with PowerSensor, an external library
// that is not vendored (app/CUDA/util.cpp:131-155: the kernel is relaunched for 10 s between two
// sensor reads).  Here the same protocol reads the GPU's own energy counter through NVML
// (nvmlDeviceGetTotalEnergyConsumption, millijoules since driver load), loaded at run time so that
// the library has no link-time dependency; no NVML or no counter -> 0 and the columns are left out,
// as in a reference build without PowerSensor.
class EnergyMeter {
 public:
  EnergyMeter() {
    lib_ = dlopen("libnvidia-ml.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!lib_) return;
    auto init = reinterpret_cast<int (*)()>(dlsym(lib_, "nvmlInit_v2"));
    by_bus_ = reinterpret_cast<int (*)(const char *, void **)>(dlsym(lib_, "nvmlDeviceGetHandleByPciBusId_v2"));
    energy_ = reinterpret_cast<int (*)(void *, unsigned long long *)>(dlsym(lib_, "nvmlDeviceGetTotalEnergyConsumption"));
    if (!init || !by_bus_ || !energy_ || init() != 0) return;
    int dev = 0;
    char bus[32] = "";
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) return;
    if (by_bus_(bus, &handle_) != 0) handle_ = nullptr;
    unsigned long long mj = 0;
    if (handle_ && energy_(handle_, &mj) != 0) handle_ = nullptr;
  }
  ~EnergyMeter() {
    if (lib_) {
      if (auto shutdown = reinterpret_cast<int (*)()>(dlsym(lib_, "nvmlShutdown"))) shutdown();
      dlclose(lib_);
    }
  }
  bool ok() const { return handle_ != nullptr; }
  double joules() const {
    unsigned long long mj = 0;
    return (handle_ && energy_(handle_, &mj) == 0) ? 1e-3 * (double)mj : 0.0;
  }

 private:
  void *lib_ = nullptr, *handle_ = nullptr;
  int (*by_bus_)(const char *, void **) = nullptr;
  int (*energy_)(void *, unsigned long long *) = nullptr;
};

int perf_run(bool gridding, idgb200_perf *result) {
  int rc = have_device();
  if (rc) return rc;
  // util.cpp:174-187
  const float image_size = 0.01f, w_step = 0.0f;  // parameters.hpp:4-5
  const int grid_size = (int)env_long("GRID_SIZE", 1024);
  const int N = (int)env_long("SUBGRID_SIZE", 32);
  const int nr_stations = (int)env_long("NR_STATIONS", 50);
  const int nr_timeslots = (int)env_long("NR_TIMESLOTS", 20);
  const int T = (int)env_long("NR_TIMESTEPS_SUBGRID", 128);
  const int C = (int)env_long("NR_CHANNELS", 16);
  const int warm = (int)env_long("NR_WARM_UP_RUNS", 2);
  const int iters = (int)env_long("NR_ITERATIONS", 5);
  if (grid_size <= 0 || N <= 0 || nr_stations < 2 || nr_timeslots <= 0 || T <= 0 || C <= 0 || iters <= 0)
    return IDGB200_EINVAL;
  const int64_t nr_baselines = (int64_t)nr_stations * (nr_stations - 1) / 2;
  const int64_t S64 = nr_baselines * nr_timeslots;
  if (S64 * T > INT32_MAX) return IDGB200_EUNSUPPORTED;  // metadata.time_offset is an int
  const int S = (int)S64;
  const int64_t tt = S64 * T;
  const size_t npix = (size_t)N * N;

  idgb200_params p = env_params(S, grid_size, N, image_size, w_step, C, nr_stations);
  std::printf(">>> idg-b200 %s: %d stations, %d timeslots, %d timesteps/subgrid, %d channels, "
              "subgrid %d, grid %d -> %d subgrids, %.3f MVis (sincos mode %d, variant %d)\n",
              gridding ? "gridder" : "degridder", nr_stations, nr_timeslots, T, C, N, grid_size, S,
              1e-6 * tt * C, p.sincos_mode, p.variant);

  idgb200_uvw *d_uvw = nullptr;
  float *d_wn = nullptr, *d_sph = nullptr;
  idgb200_cfloat *d_vis = nullptr, *d_at = nullptr, *d_sg = nullptr;
  idgb200_metadata *d_meta = nullptr;
  CK(cudaMalloc(&d_uvw, tt * sizeof(idgb200_uvw)));
  CK(cudaMalloc(&d_wn, C * sizeof(float)));
  CK(cudaMalloc(&d_sph, npix * sizeof(float)));
  CK(cudaMalloc(&d_vis, (size_t)tt * C * NR_POL * sizeof(float2)));
  CK(cudaMalloc(&d_at, (size_t)nr_timeslots * nr_stations * npix * NR_POL * sizeof(float2)));
  CK(cudaMalloc(&d_sg, (size_t)S * NR_POL * npix * sizeof(float2)));
  CK(cudaMalloc(&d_meta, (size_t)S * sizeof(idgb200_metadata)));

  // unlike util.cpp:216-231 every buffer is initialised
  rc = idgb200_init_uvw(grid_size, S, T, 0, d_uvw, nullptr);
  if (!rc) rc = idgb200_init_wavenumbers(C, d_wn, nullptr);
  if (!rc) rc = idgb200_init_visibilities(grid_size, image_size, tt, C, d_uvw, d_vis, nullptr);
  if (!rc) rc = idgb200_init_spheroidal(N, d_sph, nullptr);
  if (!rc) rc = idgb200_init_aterms(nr_timeslots, nr_stations, N, 0, d_at, nullptr);
  if (!rc) rc = idgb200_init_metadata(grid_size, nr_stations, nr_timeslots, T, 0, 0, d_meta, nullptr);
  if (!rc) rc = idgb200_init_subgrids(S, N, d_sg, nullptr);
  if (rc) return rc;
  CK(cudaDeviceSynchronize());

  auto run = [&]() {
    return gridding ? idgb200_gridder(&p, d_uvw, d_wn, d_vis, d_sph, d_at, d_meta, d_sg, nullptr)
                    : idgb200_degridder(&p, d_uvw, d_wn, d_vis, d_sph, d_at, d_meta, d_sg, nullptr);
  };
  // util.cpp:93-128
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < warm && !rc; i++) rc = run();
  if (rc) return rc;
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters && !rc; i++) rc = run();
  if (rc) return rc;
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double seconds = ms * 1e-3 / iters;

  const double gflops = 1e-9 * idgb200_flops_gridder(C, tt, S, N, NR_POL);
  const double gbytes = 1e-9 * idgb200_bytes_gridder(C, tt, S, N, NR_POL);
  const double mvis = 1e-6 * tt * C;
  // energy: relaunch for IDGB200_ENERGY_SECONDS (default 2; 0 = off) between two counter reads
  double joules = 0;
  const double energy_s = (double)env_long("IDGB200_ENERGY_SECONDS", 2);
  if (energy_s > 0) {
    EnergyMeter meter;
    if (meter.ok()) {
      const int launches = (int)std::max(1.0, energy_s / seconds);
      CK(cudaDeviceSynchronize());
      const double j0 = meter.joules();
      for (int i = 0; i < launches && !rc; i++) rc = run();
      if (rc) return rc;
      CK(cudaDeviceSynchronize());
      const double j1 = meter.joules();
      if (j1 > j0) joules = (j1 - j0) / launches;
    }
  }
  const char *name = gridding ? "gridder_b200" : "degridder_b200";
  report(name, seconds, gflops, gbytes, mvis, joules);
  if (std::getenv("OUTPUT_PATH")) {   // unlike the reference (which defaults to "."), only written when asked for
    char dev[256] = "";
    idgb200_device_name(dev, sizeof dev);
    report_csv(name, dev, "-cuda.csv", seconds, gflops, gbytes, mvis, joules);
  }
  if (result) {
    result->joules = joules;
    result->seconds = seconds;
    result->gflops = gflops;
    result->gbytes = gbytes;
    result->mvis = mvis;
    result->nr_subgrids = S;
    result->iterations = iters;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_uvw); cudaFree(d_wn); cudaFree(d_sph); cudaFree(d_vis);
  cudaFree(d_at); cudaFree(d_sg); cudaFree(d_meta);
  return IDGB200_OK;
}

}  // namespace

// ================================================================== extern "C"
extern "C" {

int idgb200_version(void) { return IDGB200_VERSION; }

const char *idgb200_error_string(int code) {
  switch (code) {
    case IDGB200_OK: return "ok";
    case IDGB200_EINVAL: return "idgb200: invalid argument";
    case IDGB200_ENODEVICE: return "idgb200: no CUDA device (there is no CPU fallback)";
    case IDGB200_EUNSUPPORTED: return "idgb200: unsupported configuration";
    case IDGB200_ENOMEM: return "idgb200: out of device memory";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "idgb200: unknown error";
}

int idgb200_device_name(char *buf, size_t len) {
  if (!buf || len == 0) return IDGB200_EINVAL;
  buf[0] = 0;
  int rc = have_device();
  if (rc) return rc;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  std::snprintf(buf, len, "%s", prop.name);
  return IDGB200_OK;
}

int idgb200_sm_count(int *count) {
  if (!count) return IDGB200_EINVAL;
  int rc = have_device();
  if (rc) return rc;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(count, cudaDevAttrMultiProcessorCount, dev));
  return IDGB200_OK;
}

int idgb200_print_device_info(void) {
  int rc = have_device();
  if (rc) return rc;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev);
  std::printf(">>> Device %d: %s (sm_%d%d), %d SMs, %.0f MHz, %.1f GB, %zu KB smem/SM, L2 %.0f MB\n", dev,
              prop.name, prop.major, prop.minor, prop.multiProcessorCount, clock_khz * 1e-3,
              prop.totalGlobalMem / 1073741824.0, prop.sharedMemPerMultiprocessor / 1024,
              prop.l2CacheSize / 1048576.0);
  std::fflush(stdout);
  return IDGB200_OK;
}

uint64_t idgb200_flops_gridder(uint64_t nr_channels, uint64_t nr_timesteps, uint64_t nr_subgrids,
                               uint64_t subgrid_size, uint64_t nr_correlations) {
  // common.cpp:100-120: per (timestep, pixel) 5 (phase index) + 5 (phase offset)
  // + per channel 2 (phase) + 8 per correlation (complex multiply-add); 6 per pixel (shift)
  const uint64_t per_vis = 5 + 5 + nr_channels * 2 + nr_channels * nr_correlations * 8;
  const uint64_t pixels = subgrid_size * subgrid_size;
  return nr_timesteps * pixels * per_vis + nr_subgrids * pixels * 6;
}

uint64_t idgb200_bytes_gridder(uint64_t nr_channels, uint64_t nr_timesteps, uint64_t nr_subgrids,
                               uint64_t subgrid_size, uint64_t nr_correlations) {
  // common.cpp:122-159: uvw + visibilities per timestep; pixel read+write, two
  // A-terms and the taper per subgrid pixel
  const uint64_t per_timestep = 3 * sizeof(float) + nr_channels * nr_correlations * 2 * sizeof(float);
  const uint64_t per_pixel = 2 * (nr_correlations * 2 * sizeof(float)) +
                             2 * nr_correlations * 2 * sizeof(float) + sizeof(float);
  return nr_timesteps * per_timestep + nr_subgrids * subgrid_size * subgrid_size * per_pixel;
}

int idgb200_gridder(const idgb200_params *p, const idgb200_uvw *d_uvw, const float *d_wn,
                    const idgb200_cfloat *d_vis, const float *d_sph, const idgb200_cfloat *d_at,
                    const idgb200_metadata *d_meta, idgb200_cfloat *d_sg, void *stream) {
  int rc = check_params(p);
  if (rc) return rc;
  if (!d_uvw || !d_wn || !d_vis || !d_sph || !d_at || !d_meta || !d_sg) return IDGB200_EINVAL;
  if (!aligned16(d_vis) || !aligned16(d_at) || !aligned16(d_sg)) return IDGB200_EINVAL;
  if ((rc = have_device())) return rc;
  KernelArgs a = make_args(p, d_uvw, d_wn, d_vis, d_sph, d_at, d_meta, d_sg);
  int nk = 0;
  cudaError_t e = launch_gridder(a, p->sincos_mode, p->variant, static_cast<cudaStream_t>(stream), &nk);
  if (e != cudaSuccess) return (int)e;
  g_launches += nk;
  return IDGB200_OK;
}

int idgb200_degridder(const idgb200_params *p, const idgb200_uvw *d_uvw, const float *d_wn,
                      idgb200_cfloat *d_vis, const float *d_sph, const idgb200_cfloat *d_at,
                      const idgb200_metadata *d_meta, const idgb200_cfloat *d_sg, void *stream) {
  int rc = check_params(p);
  if (rc) return rc;
  if (!d_uvw || !d_wn || !d_vis || !d_sph || !d_at || !d_meta || !d_sg) return IDGB200_EINVAL;
  if (!aligned16(d_vis) || !aligned16(d_at) || !aligned16(d_sg)) return IDGB200_EINVAL;
  if ((rc = have_device())) return rc;
  KernelArgs a = make_args(p, d_uvw, d_wn, d_vis, d_sph, d_at, d_meta, d_sg);
  int nk = 0;
  cudaError_t e = launch_degridder(a, p->sincos_mode, p->variant, static_cast<cudaStream_t>(stream), &nk);
  if (e != cudaSuccess) return (int)e;
  g_launches += nk;
  return IDGB200_OK;
}

uint64_t idgb200_launch_count(void) { return g_launches.load(); }

void idgb200_report(const char *name, double seconds, double gflops, double gbytes, double mvis, double joules) {
  report(name ? name : "", seconds, gflops, gbytes, mvis, joules);
}

void idgb200_report_csv(const char *name, const char *device_name, const char *file_extension, double seconds,
                        double gflops, double gbytes, double mvis, double joules) {
  report_csv(name ? name : "", device_name, file_extension, seconds, gflops, gbytes, mvis, joules);
}

int idgb200_adder(const idgb200_params *p, const idgb200_metadata *d_meta, const idgb200_cfloat *d_sg,
                  idgb200_cfloat *const *grid_parts, int nr_parts, int rows_per_part, void *stream) {
  int rc = check_params(p);
  if (rc) return rc;
  if (!d_meta || !d_sg || !grid_parts || nr_parts < 1 || nr_parts > 16 || rows_per_part < 1 ||
      (long long)nr_parts * rows_per_part < p->grid_size)
    return IDGB200_EINVAL;
  for (int i = 0; i < nr_parts; i++)
    if (!grid_parts[i]) return IDGB200_EINVAL;
  if (p->subgrid_size > 1024) return IDGB200_EUNSUPPORTED;   // adder.cu: div_small
  if ((rc = have_device())) return rc;
  cudaError_t e = launch_adder(p->nr_subgrids, 0, p->grid_size, p->subgrid_size, p->flags, d_meta,
                               reinterpret_cast<const float2 *>(d_sg),
                               reinterpret_cast<float2 *const *>(grid_parts), nr_parts, rows_per_part,
                               static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return (int)e;
  g_launches++;
  return IDGB200_OK;
}

int idgb200_splitter(const idgb200_params *p, const idgb200_metadata *d_meta, idgb200_cfloat *d_sg,
                     const idgb200_cfloat *const *grid_parts, int nr_parts, int rows_per_part, void *stream) {
  int rc = check_params(p);
  if (rc) return rc;
  if (!d_meta || !d_sg || !grid_parts || nr_parts < 1 || nr_parts > 16 || rows_per_part < 1 ||
      (long long)nr_parts * rows_per_part < p->grid_size)
    return IDGB200_EINVAL;
  for (int i = 0; i < nr_parts; i++)
    if (!grid_parts[i]) return IDGB200_EINVAL;
  if (p->subgrid_size > 1024) return IDGB200_EUNSUPPORTED;   // adder.cu: div_small
  if ((rc = have_device())) return rc;
  cudaError_t e = launch_splitter(p->nr_subgrids, 0, p->grid_size, p->subgrid_size, p->flags, d_meta,
                                  reinterpret_cast<float2 *>(d_sg),
                                  reinterpret_cast<const float2 *const *>(grid_parts), nr_parts, rows_per_part,
                                  static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return (int)e;
  g_launches++;
  return IDGB200_OK;
}

int idgb200_reduce_parts(int nr_sources, const idgb200_cfloat *const *sources, int64_t count,
                         idgb200_cfloat *d_out, void *stream) {
  if (nr_sources < 1 || nr_sources > 16 || !sources || count < 0 || (count & 1)) return IDGB200_EINVAL;
  if (count > 0 && (!d_out || ((uintptr_t)d_out & 15))) return IDGB200_EINVAL;
  for (int i = 0; i < nr_sources; i++)
    if (!sources[i] || ((uintptr_t)sources[i] & 15)) return IDGB200_EINVAL;
  int rc = have_device();
  if (rc) return rc;
  if (count == 0) return IDGB200_OK;
  int sms = 0;
  idgb200_sm_count(&sms);
  cudaError_t e = launch_reduce_parts(nr_sources, reinterpret_cast<const float2 *const *>(sources), count,
                                      reinterpret_cast<float2 *>(d_out), sms, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return (int)e;
  g_launches++;
  return IDGB200_OK;
}

int idgb200_adder_rs_mode(int64_t nr_subgrids, int subgrid_size, int grid_size) {
  if (nr_subgrids < 0 || subgrid_size <= 0 || grid_size <= 0) return IDGB200_EINVAL;
  return nr_subgrids * (int64_t)subgrid_size * subgrid_size < (int64_t)grid_size * grid_size ? 1 : 0;
}

int idgb200_subgrid_fft(int64_t nr_subgrids, int subgrid_size, int direction, idgb200_cfloat *d_sg,
                        void *stream) {
  if (nr_subgrids < 0 || subgrid_size < 1 || (direction != 1 && direction != -1) ||
      nr_subgrids * NR_POL > 0x7fffffffLL)
    return IDGB200_EINVAL;
  if (nr_subgrids > 0 && !d_sg) return IDGB200_EINVAL;
  // shared memory of the direct-DFT kernel that serves the unusual sizes
  if (((size_t)2 * subgrid_size * (subgrid_size + 1) + subgrid_size) * sizeof(float2) > 200 * 1024)
    return IDGB200_EUNSUPPORTED;
  int rc = have_device();
  if (rc) return rc;
  if (nr_subgrids == 0) return IDGB200_OK;
  cudaError_t e = launch_subgrid_fft(nr_subgrids * NR_POL, subgrid_size, direction,
                                     reinterpret_cast<float2 *>(d_sg), static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return (int)e;
  g_launches++;
  return IDGB200_OK;
}

int idgb200_resolve_variant(const idgb200_params *p, int gridder) {
  int rc = check_params(p);
  if (rc) return rc;
  return gridder ? resolve_gridder_variant(p->subgrid_size, p->nr_channels, p->sincos_mode, p->variant)
                 : resolve_degridder_variant(p->subgrid_size, p->nr_channels, p->sincos_mode, p->variant);
}

int idgb200_c_run_gridder_ex(const idgb200_params *p, int64_t total_timesteps, int nr_aterm_slots,
                             const idgb200_uvw *uvw, const float *wn, const idgb200_cfloat *vis,
                             const float *sph, const idgb200_cfloat *at, const idgb200_metadata *meta,
                             idgb200_cfloat *sg) {
  return host_run(true, p, total_timesteps, nr_aterm_slots, uvw, wn, const_cast<idgb200_cfloat *>(vis), sph,
                  at, meta, sg);
}

int idgb200_c_run_degridder_ex(const idgb200_params *p, int64_t total_timesteps, int nr_aterm_slots,
                               const idgb200_uvw *uvw, const float *wn, idgb200_cfloat *vis,
                               const float *sph, const idgb200_cfloat *at, const idgb200_metadata *meta,
                               const idgb200_cfloat *sg) {
  return host_run(false, p, total_timesteps, nr_aterm_slots, uvw, wn, vis, sph, at, meta,
                  const_cast<idgb200_cfloat *>(sg));
}

int idgb200_c_run_gridder(int nr_subgrids, int grid_size, int subgrid_size, float image_size,
                          float w_step_in_lambda, int nr_channels, int nr_stations, int64_t total_timesteps,
                          int nr_aterm_slots, const idgb200_uvw *uvw, const float *wn,
                          const idgb200_cfloat *vis, const float *sph, const idgb200_cfloat *at,
                          const idgb200_metadata *meta, idgb200_cfloat *sg) {
  idgb200_params p = env_params(nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda,
                                nr_channels, nr_stations);
  return idgb200_c_run_gridder_ex(&p, total_timesteps, nr_aterm_slots, uvw, wn, vis, sph, at, meta, sg);
}

int idgb200_c_run_degridder(int nr_subgrids, int grid_size, int subgrid_size, float image_size,
                            float w_step_in_lambda, int nr_channels, int nr_stations,
                            int64_t total_timesteps, int nr_aterm_slots, const idgb200_uvw *uvw,
                            const float *wn, idgb200_cfloat *vis, const float *sph, const idgb200_cfloat *at,
                            const idgb200_metadata *meta, const idgb200_cfloat *sg) {
  idgb200_params p = env_params(nr_subgrids, grid_size, subgrid_size, image_size, w_step_in_lambda,
                                nr_channels, nr_stations);
  return idgb200_c_run_degridder_ex(&p, total_timesteps, nr_aterm_slots, uvw, wn, vis, sph, at, meta, sg);
}

int idgb200_host_alloc(void **ptr, size_t bytes) {
  if (!ptr) return IDGB200_EINVAL;
  int rc = have_device();
  if (rc) return rc;
  CK(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return IDGB200_OK;
}

int idgb200_host_free(void *ptr) {
  if (!ptr) return IDGB200_OK;
  CK(cudaFreeHost(ptr));
  return IDGB200_OK;
}

int idgb200_p_run_gridder(idgb200_perf *result) { return perf_run(true, result); }
int idgb200_p_run_degridder(idgb200_perf *result) { return perf_run(false, result); }

// ------------------------------------------------------------ synthetic inputs
#define INIT_PROLOGUE(ptr)                 \
  if (!(ptr)) return IDGB200_EINVAL;       \
  {                                        \
    int rc__ = have_device();              \
    if (rc__) return rc__;                 \
  }                                        \
  cudaStream_t st = static_cast<cudaStream_t>(stream)

int idgb200_init_uvw(uint32_t grid_size, int64_t nr_baselines, int nr_timesteps, uint32_t seed,
                     idgb200_uvw *d_uvw, void *stream) {
  INIT_PROLOGUE(d_uvw);
  const int64_t n = nr_baselines * nr_timesteps;
  if (n <= 0) return n == 0 ? IDGB200_OK : IDGB200_EINVAL;
  k_init_uvw<<<blocks_for(n), 256, 0, st>>>(grid_size, nr_baselines, nr_timesteps, seed, d_uvw);
  return (int)cudaGetLastError();
}

int idgb200_init_wavenumbers(int nr_channels, float *d_wn, void *stream) {
  INIT_PROLOGUE(d_wn);
  if (nr_channels <= 0) return IDGB200_EINVAL;
  k_init_wavenumbers<<<blocks_for(nr_channels), 256, 0, st>>>(nr_channels, d_wn);
  return (int)cudaGetLastError();
}

int idgb200_init_visibilities(uint32_t grid_size, float image_size, int64_t total_timesteps, int nr_channels,
                              const idgb200_uvw *d_uvw, idgb200_cfloat *d_vis, void *stream) {
  INIT_PROLOGUE(d_vis);
  if (!d_uvw || nr_channels <= 0 || total_timesteps < 0) return IDGB200_EINVAL;
  const int64_t n = total_timesteps * nr_channels;
  if (n == 0) return IDGB200_OK;
  k_init_vis<<<blocks_for(n), 256, 0, st>>>(grid_size, image_size, total_timesteps, nr_channels, d_uvw,
                                            reinterpret_cast<float4 *>(d_vis));
  return (int)cudaGetLastError();
}

int idgb200_init_spheroidal(int subgrid_size, float *d_sph, void *stream) {
  INIT_PROLOGUE(d_sph);
  if (subgrid_size <= 0) return IDGB200_EINVAL;
  k_init_sph<<<blocks_for((int64_t)subgrid_size * subgrid_size), 256, 0, st>>>(subgrid_size, d_sph);
  return (int)cudaGetLastError();
}

int idgb200_init_aterms(int nr_slots, int nr_stations, int subgrid_size, uint32_t seed, idgb200_cfloat *d_at,
                        void *stream) {
  INIT_PROLOGUE(d_at);
  if (nr_slots <= 0 || nr_stations <= 0 || subgrid_size <= 0) return IDGB200_EINVAL;
  const int64_t n = (int64_t)nr_slots * nr_stations * subgrid_size * subgrid_size;
  k_init_aterms<<<blocks_for(n), 256, 0, st>>>(n, subgrid_size, seed, reinterpret_cast<float4 *>(d_at));
  return (int)cudaGetLastError();
}

int idgb200_init_metadata(uint32_t grid_size, int nr_stations, int nr_timeslots, int nr_timesteps_subgrid,
                          int per_slot_aterms, uint32_t seed, idgb200_metadata *d_meta, void *stream) {
  INIT_PROLOGUE(d_meta);
  if (nr_stations < 2 || nr_timeslots <= 0 || nr_timesteps_subgrid <= 0) return IDGB200_EINVAL;
  const int64_t nr_baselines = (int64_t)nr_stations * (nr_stations - 1) / 2;
  const int64_t n = nr_baselines * nr_timeslots;
  if (n * nr_timesteps_subgrid > INT32_MAX) return IDGB200_EUNSUPPORTED;
  k_init_metadata<<<blocks_for(n), 256, 0, st>>>(grid_size, nr_stations, nr_baselines, nr_timeslots,
                                                 nr_timesteps_subgrid, per_slot_aterms, seed, d_meta);
  return (int)cudaGetLastError();
}

int idgb200_init_subgrids(int64_t nr_subgrids, int subgrid_size, idgb200_cfloat *d_sg, void *stream) {
  INIT_PROLOGUE(d_sg);
  if (nr_subgrids < 0 || subgrid_size <= 0) return IDGB200_EINVAL;
  const int64_t n = nr_subgrids * NR_POL * subgrid_size * subgrid_size;
  if (n == 0) return IDGB200_OK;
  k_init_subgrids<<<blocks_for(n), 256, 0, st>>>(n, subgrid_size, reinterpret_cast<float2 *>(d_sg));
  return (int)cudaGetLastError();
}

}  // extern "C"
