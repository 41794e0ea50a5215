// Subgrid FFT (a "next" row): the step of Image Domain Gridding between the gridder
// and the adder (forward) and between the splitter and the degridder (backward): an in-place 2-D DFT
// of every [subgrid][pol] plane of N x N complex64 pixels,
//     forward   B[ky][kx] =        sum_{y,x} A[y][x] exp(-2 pi i (ky y + kx x) / N)
//     backward  B[ky][kx] = 1/N^2  sum_{y,x} A[y][x] exp(+2 pi i (ky y + kx x) / N)
// The reference has no FFT (its gridder stores image-domain pixels, gridder_reference.cpp:105-109,
// and nothing follows); oracle/idg_next_oracle.c states the sums the tests check (parity unpinned).
//
// The kernel is HBM-bound (one read and one write of every pixel, 16 B per pixel and launch):
//  * N threads own one plane, 96 or 128 threads (2 ... 16 planes) a CTA; a thread first owns a column:
//    its N loads are 8-byte accesses at consecutive addresses across the lanes (coalesced), the whole
//    column sits in registers and the 1-D DFT over y is a fully unrolled decimation-in-frequency
//    network (radix 2 down to a 3-point or 1-point base: N = 2^a or 3 * 2^a) whose twiddles are
//    compile-time constants (immediates in the FFMAs), output permutation folded into the indices;
//  * the columns go through padded shared memory (pitch N + 1: conflict-free both ways) and come
//    back as rows for the second 1-D DFT; the result takes the same way back and is stored coalesced.
//  * other even or odd N: a direct O(N^3) DFT per plane in shared memory (correct, not tuned).
// Bytes in flight per SM: regs/thread ~ 2.5 N -> at N = 32 about 20 planes x 8 KB.
#include <utility>

#include "common.cuh"
#include "kernels.h"

namespace idgb200 {

namespace {

// exp(2 pi i k / 192), k = 0..191: covers every N that divides 192 (8, 16, 24, 32, 48, 64, 96)
constexpr int TW_L = 192;
__device__ constexpr float TW_COS[192] = {
    1.0f, 0.9994645874763657f, 0.9978589232386035f, 0.9951847266721969f, 0.9914448613738104f, 0.986643332084879f,
    0.9807852804032304f, 0.9738769792773336f, 0.9659258262890683f, 0.9569403357322088f, 0.9469301294951057f, 0.9359059267573258f,
    0.9238795325112867f, 0.9108638249211758f, 0.8968727415326884f, 0.881921264348355f, 0.8660254037844387f, 0.8492021815265789f,
    0.8314696123025452f, 0.8128466845916152f, 0.7933533402912352f, 0.773010453362737f, 0.7518398074789774f, 0.7298640726978357f,
    0.7071067811865476f, 0.6835923020228712f, 0.6593458151000688f, 0.6343932841636455f, 0.6087614290087207f, 0.5824776968678022f,
    0.5555702330196024f, 0.5280678506503681f, 0.5000000000000001f, 0.4713967368259976f, 0.44228869021900125f, 0.4127070298043947f,
    0.38268343236508984f, 0.3522500479212336f, 0.3214394653031617f, 0.2902846772544625f, 0.25881904510252074f, 0.22707626303437345f,
    0.19509032201612833f, 0.16289547339458882f, 0.1305261922200517f, 0.09801714032956055f, 0.06540312923014327f, 0.032719082821776165f,
    0.0f, -0.03271908282177604f, -0.06540312923014314f, -0.09801714032956042f, -0.1305261922200516f, -0.1628954733945887f,
    -0.1950903220161282f, -0.2270762630343733f, -0.25881904510252063f, -0.29028467725446216f, -0.3214394653031616f, -0.3522500479212335f,
    -0.3826834323650895f, -0.4127070298043946f, -0.44228869021900113f, -0.4713967368259977f, -0.4999999999999998f, -0.528067850650368f,
    -0.5555702330196023f, -0.582477696867802f, -0.6087614290087207f, -0.6343932841636454f, -0.6593458151000688f, -0.6835923020228714f,
    -0.7071067811865475f, -0.7298640726978354f, -0.7518398074789773f, -0.773010453362737f, -0.793353340291235f, -0.8128466845916151f,
    -0.831469612302545f, -0.8492021815265788f, -0.8660254037844387f, -0.8819212643483549f, -0.8968727415326881f, -0.9108638249211759f,
    -0.9238795325112867f, -0.9359059267573258f, -0.9469301294951056f, -0.9569403357322087f, -0.9659258262890682f, -0.9738769792773336f,
    -0.9807852804032304f, -0.986643332084879f, -0.9914448613738104f, -0.9951847266721968f, -0.9978589232386035f, -0.9994645874763657f,
    -1.0f, -0.9994645874763657f, -0.9978589232386035f, -0.9951847266721969f, -0.9914448613738104f, -0.986643332084879f,
    -0.9807852804032305f, -0.9738769792773336f, -0.9659258262890683f, -0.9569403357322088f, -0.9469301294951057f, -0.9359059267573259f,
    -0.9238795325112868f, -0.910863824921176f, -0.8968727415326883f, -0.881921264348355f, -0.8660254037844388f, -0.8492021815265789f,
    -0.8314696123025455f, -0.812846684591615f, -0.7933533402912352f, -0.7730104533627371f, -0.7518398074789775f, -0.7298640726978359f,
    -0.7071067811865479f, -0.6835923020228712f, -0.6593458151000691f, -0.6343932841636459f, -0.6087614290087209f, -0.5824776968678023f,
    -0.5555702330196022f, -0.5280678506503678f, -0.5000000000000004f, -0.47139673682599786f, -0.44228869021900136f, -0.41270702980439467f,
    -0.3826834323650895f, -0.35225004792123393f, -0.3214394653031618f, -0.29028467725446244f, -0.25881904510252063f, -0.22707626303437292f,
    -0.19509032201612866f, -0.16289547339458896f, -0.13052619222005163f, -0.09801714032956134f, -0.06540312923014273f, -0.03271908282177651f,
    0.0f, 0.032719082821776144f, 0.06540312923014237f, 0.09801714032956096f, 0.13052619222005127f, 0.1628954733945886f,
    0.1950903220161283f, 0.22707626303437256f, 0.2588190451025203f, 0.29028467725446205f, 0.3214394653031615f, 0.35225004792123354f,
    0.38268343236508917f, 0.41270702980439433f, 0.442288690219001f, 0.4713967368259976f, 0.5000000000000001f, 0.5280678506503674f,
    0.5555702330196018f, 0.5824776968678019f, 0.6087614290087199f, 0.6343932841636449f, 0.6593458151000691f, 0.6835923020228717f,
    0.7071067811865474f, 0.7298640726978356f, 0.7518398074789775f, 0.7730104533627367f, 0.7933533402912349f, 0.8128466845916151f,
    0.8314696123025448f, 0.8492021815265786f, 0.8660254037844384f, 0.8819212643483553f, 0.8968727415326883f, 0.9108638249211758f,
    0.9238795325112868f, 0.9359059267573255f, 0.9469301294951056f, 0.9569403357322088f, 0.9659258262890681f, 0.9738769792773335f,
    0.9807852804032303f, 0.9866433320848791f, 0.9914448613738104f, 0.9951847266721969f, 0.9978589232386035f, 0.9994645874763657f,
};
__device__ constexpr float TW_SIN[192] = {
    0.0f, 0.03271908282177614f, 0.06540312923014306f, 0.0980171403295606f, 0.13052619222005157f, 0.16289547339458874f,
    0.19509032201612825f, 0.2270762630343732f, 0.25881904510252074f, 0.29028467725446233f, 0.3214394653031616f, 0.3522500479212335f,
    0.3826834323650898f, 0.4127070298043947f, 0.44228869021900125f, 0.4713967368259976f, 0.49999999999999994f, 0.528067850650368f,
    0.5555702330196022f, 0.5824776968678022f, 0.6087614290087207f, 0.6343932841636455f, 0.6593458151000688f, 0.6835923020228712f,
    0.7071067811865475f, 0.7298640726978357f, 0.7518398074789774f, 0.773010453362737f, 0.7933533402912352f, 0.8128466845916152f,
    0.8314696123025451f, 0.8492021815265789f, 0.8660254037844386f, 0.881921264348355f, 0.8968727415326884f, 0.9108638249211758f,
    0.9238795325112867f, 0.9359059267573256f, 0.9469301294951056f, 0.9569403357322088f, 0.9659258262890683f, 0.9738769792773336f,
    0.9807852804032304f, 0.986643332084879f, 0.9914448613738104f, 0.9951847266721969f, 0.9978589232386035f, 0.9994645874763657f,
    1.0f, 0.9994645874763657f, 0.9978589232386035f, 0.9951847266721969f, 0.9914448613738104f, 0.986643332084879f,
    0.9807852804032304f, 0.9738769792773336f, 0.9659258262890683f, 0.9569403357322089f, 0.9469301294951057f, 0.9359059267573258f,
    0.9238795325112868f, 0.9108638249211759f, 0.8968727415326884f, 0.881921264348355f, 0.8660254037844387f, 0.8492021815265789f,
    0.8314696123025451f, 0.8128466845916152f, 0.7933533402912352f, 0.7730104533627371f, 0.7518398074789774f, 0.7298640726978356f,
    0.7071067811865476f, 0.6835923020228716f, 0.659345815100069f, 0.6343932841636455f, 0.6087614290087209f, 0.5824776968678022f,
    0.5555702330196025f, 0.5280678506503681f, 0.49999999999999994f, 0.47139673682599786f, 0.4422886902190017f, 0.4127070298043946f,
    0.3826834323650899f, 0.35225004792123343f, 0.32143946530316175f, 0.2902846772544628f, 0.258819045102521f, 0.22707626303437328f,
    0.19509032201612816f, 0.1628954733945889f, 0.130526192220052f, 0.09801714032956083f, 0.06540312923014312f, 0.032719082821776005f,
    0.0f, -0.03271908282177576f, -0.06540312923014287f, -0.09801714032956059f, -0.13052619222005177f, -0.16289547339458865f,
    -0.19509032201612792f, -0.22707626303437303f, -0.2588190451025208f, -0.29028467725446255f, -0.32143946530316153f, -0.3522500479212332f,
    -0.38268343236508967f, -0.4127070298043944f, -0.44228869021900147f, -0.47139673682599764f, -0.4999999999999997f, -0.5280678506503679f,
    -0.555570233019602f, -0.5824776968678024f, -0.6087614290087207f, -0.6343932841636453f, -0.6593458151000688f, -0.683592302022871f,
    -0.7071067811865471f, -0.7298640726978357f, -0.7518398074789773f, -0.7730104533627367f, -0.7933533402912349f, -0.8128466845916151f,
    -0.8314696123025452f, -0.849202181526579f, -0.8660254037844384f, -0.8819212643483549f, -0.8968727415326883f, -0.9108638249211759f,
    -0.9238795325112868f, -0.9359059267573255f, -0.9469301294951056f, -0.9569403357322088f, -0.9659258262890683f, -0.9738769792773337f,
    -0.9807852804032303f, -0.986643332084879f, -0.9914448613738104f, -0.9951847266721968f, -0.9978589232386035f, -0.9994645874763657f,
    -1.0f, -0.9994645874763657f, -0.9978589232386036f, -0.9951847266721968f, -0.9914448613738105f, -0.986643332084879f,
    -0.9807852804032304f, -0.9738769792773339f, -0.9659258262890684f, -0.9569403357322089f, -0.9469301294951057f, -0.9359059267573256f,
    -0.923879532511287f, -0.910863824921176f, -0.8968727415326885f, -0.881921264348355f, -0.8660254037844386f, -0.8492021815265792f,
    -0.8314696123025455f, -0.8128466845916154f, -0.7933533402912357f, -0.7730104533627374f, -0.7518398074789772f, -0.7298640726978354f,
    -0.7071067811865477f, -0.6835923020228714f, -0.6593458151000687f, -0.6343932841636459f, -0.6087614290087209f, -0.5824776968678023f,
    -0.555570233019603f, -0.5280678506503685f, -0.5000000000000004f, -0.47139673682599714f, -0.4422886902190014f, -0.4127070298043947f,
    -0.38268343236508956f, -0.352250047921234f, -0.32143946530316186f, -0.2902846772544625f, -0.25881904510252157f, -0.22707626303437384f,
    -0.19509032201612872f, -0.16289547339458813f, -0.13052619222005168f, -0.0980171403295605f, -0.0654031292301428f, -0.032719082821776574f,
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// d * w_N^J, w_N = exp(-+ 2 pi i / N) (forward: -), trivial powers without multiplications
template <int N, bool INV, int J>
__device__ __forceinline__ float2 twiddle_mul(float2 d) {
  constexpr int K = (J % N) * (TW_L / N);
  if constexpr (K == 0) return d;
  else if constexpr (K == TW_L / 2) return make_float2(-d.x, -d.y);
  else if constexpr (K == TW_L / 4) return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
  else if constexpr (K == 3 * TW_L / 4) return INV ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x);
  else {
    constexpr float c = TW_COS[K];
    constexpr float s = INV ? TW_SIN[K] : -TW_SIN[K];
    return make_float2(fmaf(d.x, c, -d.y * s), fmaf(d.x, s, d.y * c));
  }
}

// where X[k] of the in-place network below ends up
constexpr int fft_pos(int n, int k) {
  return (n == 1 || n == 3) ? k : ((k & 1) ? n / 2 + fft_pos(n / 2, k / 2) : fft_pos(n / 2, k / 2));
}

template <int N, bool INV>
struct Dft;

template <int N, bool INV, int J>
__device__ __forceinline__ void butterfly(float2 *v) {
  const float2 a = v[J], b = v[J + N / 2];
  v[J] = cadd(a, b);
  v[J + N / 2] = twiddle_mul<N, INV, J>(csub(a, b));
}
template <int N, bool INV, int... J>
__device__ __forceinline__ void stage(float2 *v, std::integer_sequence<int, J...>) {
  (butterfly<N, INV, J>(v), ...);
}

// decimation in frequency: X[2k] = DFT_{N/2}(a_j + a_{j+N/2}), X[2k+1] = DFT_{N/2}((a_j - a_{j+N/2}) w^j)
template <int N, bool INV>
struct Dft {
  static __device__ __forceinline__ void run(float2 *v) {
    stage<N, INV>(v, std::make_integer_sequence<int, N / 2>{});
    Dft<N / 2, INV>::run(v);
    Dft<N / 2, INV>::run(v + N / 2);
  }
};
template <bool INV>
struct Dft<1, INV> {
  static __device__ __forceinline__ void run(float2 *) {}
};
template <bool INV>
struct Dft<3, INV> {
  static __device__ __forceinline__ void run(float2 *v) {
    const float2 a = v[0], t = cadd(v[1], v[2]), d = csub(v[1], v[2]);
    const float2 u = make_float2(fmaf(t.x, -0.5f, a.x), fmaf(t.y, -0.5f, a.y));
    const float h = 0.8660254037844386f;                         // sin(2 pi / 3)
    const float2 r = INV ? make_float2(-d.y * h, d.x * h)        // +i h d
                         : make_float2(d.y * h, -d.x * h);       // -i h d
    v[0] = cadd(a, t);
    v[1] = cadd(u, r);
    v[2] = csub(u, r);
  }
};

constexpr int fft_planes_per_cta(int n) { return 128 / n; }

template <int N, bool INV, int... K>
__device__ __forceinline__ void store_permuted(float2 *dst, int stride, const float2 *v, float scale,
                                               std::integer_sequence<int, K...>) {
  ((dst[K * stride] = make_float2(v[fft_pos(N, K)].x * scale, v[fft_pos(N, K)].y * scale)), ...);
}

template <int N, bool INV>
__global__ void __launch_bounds__(fft_planes_per_cta(N) * N)
subgrid_fft_kernel(float2 *__restrict__ planes, const long long nr_planes, const float scale) {
  constexpr int PPB = fft_planes_per_cta(N), PITCH = N + 1;
  extern __shared__ __align__(16) unsigned char fft_smem[];
  float2 *s_all = reinterpret_cast<float2 *>(fft_smem);
  const int p = threadIdx.x / N, i = threadIdx.x - p * N;
  const long long plane = (long long)blockIdx.x * PPB + p;
  const bool live = plane < nr_planes;
  float2 *g = planes + (size_t)(live ? plane : 0) * N * N;
  float2 *sp = s_all + p * N * PITCH;
  float2 v[N];
  // column i of the plane: lanes along x
  if (live) {
#pragma unroll
    for (int y = 0; y < N; y++) v[y] = __ldcs(&g[y * N + i]);
  } else {
#pragma unroll
    for (int y = 0; y < N; y++) v[y] = make_float2(0.f, 0.f);
  }
  Dft<N, INV>::run(v);
  store_permuted<N, INV>(sp + i, PITCH, v, 1.f, std::make_integer_sequence<int, N>{});   // sp[ky][x = i]
  __syncthreads();
  // row i (= ky): thread-private from here to the second barrier
#pragma unroll
  for (int x = 0; x < N; x++) v[x] = sp[i * PITCH + x];
  Dft<N, INV>::run(v);
  store_permuted<N, INV>(sp + i * PITCH, 1, v, scale, std::make_integer_sequence<int, N>{});   // sp[ky = i][kx]
  __syncthreads();
  if (live) {
#pragma unroll
    for (int y = 0; y < N; y++) __stcs(&g[y * N + i], sp[y * PITCH + i]);
  }
}

// any N: direct DFT, one CTA per plane
template <bool INV>
__global__ void __launch_bounds__(128)
subgrid_dft_generic_kernel(float2 *__restrict__ planes, const int N, const float scale) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  const int PITCH = N + 1;
  float2 *sa = reinterpret_cast<float2 *>(fft_smem);   // [N][PITCH]
  float2 *sb = sa + N * PITCH;                         // [N][PITCH]
  float2 *tw = sb + N * PITCH;                         // [N]
  float2 *g = planes + (size_t)blockIdx.x * N * N;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    float s, c;
    sincospif(2.f * (float)k / (float)N, &s, &c);
    tw[k] = make_float2(c, INV ? s : -s);
  }
  for (int q = threadIdx.x; q < N * N; q += blockDim.x) sa[(q / N) * PITCH + q % N] = g[q];
  __syncthreads();
  for (int q = threadIdx.x; q < N * N; q += blockDim.x) {   // over y: sb[ky][x]
    const int ky = q / N, x = q - ky * N;
    float2 acc = make_float2(0.f, 0.f);
    int e = 0;
    for (int y = 0; y < N; y++) {
      const float2 a = sa[y * PITCH + x], w = tw[e];
      acc.x = fmaf(a.x, w.x, fmaf(-a.y, w.y, acc.x));
      acc.y = fmaf(a.x, w.y, fmaf(a.y, w.x, acc.y));
      e += ky;
      if (e >= N) e -= N;
    }
    sb[ky * PITCH + x] = acc;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < N * N; q += blockDim.x) {   // over x
    const int ky = q / N, kx = q - ky * N;
    float2 acc = make_float2(0.f, 0.f);
    int e = 0;
    for (int x = 0; x < N; x++) {
      const float2 a = sb[ky * PITCH + x], w = tw[e];
      acc.x = fmaf(a.x, w.x, fmaf(-a.y, w.y, acc.x));
      acc.y = fmaf(a.x, w.y, fmaf(a.y, w.x, acc.y));
      e += kx;
      if (e >= N) e -= N;
    }
    g[q] = make_float2(acc.x * scale, acc.y * scale);
  }
}

template <int N, bool INV>
cudaError_t launch_fast(float2 *planes, long long nr_planes, float scale, cudaStream_t stream) {
  constexpr int PPB = fft_planes_per_cta(N);
  constexpr size_t smem = (size_t)PPB * N * (N + 1) * sizeof(float2);
  auto k = subgrid_fft_kernel<N, INV>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const long long ctas = (nr_planes + PPB - 1) / PPB;
  k<<<dim3((unsigned)ctas), dim3(PPB * N), smem, stream>>>(planes, nr_planes, scale);
  return cudaGetLastError();
}

template <bool INV>
cudaError_t launch_dir(int N, float2 *planes, long long nr_planes, float scale, cudaStream_t stream) {
  switch (N) {
    case 8: return launch_fast<8, INV>(planes, nr_planes, scale, stream);
    case 16: return launch_fast<16, INV>(planes, nr_planes, scale, stream);
    case 24: return launch_fast<24, INV>(planes, nr_planes, scale, stream);
    case 32: return launch_fast<32, INV>(planes, nr_planes, scale, stream);
    case 48: return launch_fast<48, INV>(planes, nr_planes, scale, stream);
    case 64: return launch_fast<64, INV>(planes, nr_planes, scale, stream);
    default: break;
  }
  const size_t smem = ((size_t)2 * N * (N + 1) + N) * sizeof(float2);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  auto k = subgrid_dft_generic_kernel<INV>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<dim3((unsigned)nr_planes), dim3(128), smem, stream>>>(planes, N, scale);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_subgrid_fft(long long nr_planes, int subgrid_size, int direction, float2 *planes,
                               cudaStream_t stream) {
  if (nr_planes == 0) return cudaSuccess;
  if (subgrid_size < 1 || nr_planes < 0 || nr_planes > 0x7fffffffLL) return cudaErrorInvalidValue;
  if (direction >= 0) return launch_dir<false>(subgrid_size, planes, nr_planes, 1.f, stream);
  const float scale = 1.f / ((float)subgrid_size * (float)subgrid_size);
  return launch_dir<true>(subgrid_size, planes, nr_planes, scale, stream);
}

}  // namespace idgb200
