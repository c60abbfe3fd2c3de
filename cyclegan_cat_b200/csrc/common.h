// Shared host/device helpers for libcyclegan_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/cyclegan_b200.h"

typedef __nv_bfloat16 bf16;

// ---- error plumbing (never throw across the C ABI) ---------------------------------
void cg_set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;      // kernels launched by this library

#define CG_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            cg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return CG_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)

#define CG_TRY(call)                       \
    do {                                   \
        int rc__ = (call);                 \
        if (rc__ != CG_OK) return rc__;    \
    } while (0)

#define CG_LAUNCH_CHECK()                                                                 \
    do {                                                                                  \
        g_launches.fetch_add(1, std::memory_order_relaxed);                               \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            cg_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return CG_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)

// Launch with programmatic stream serialization (PDL): the kernel may be scheduled while its predecessor in the stream
// is still draining; it must call griddep_wait() (ptx_async.h) before touching global memory.  CG_DISABLE_PDL=1 turns
// the attribute off (plain stream order) for A/B measurements.
bool cg_pdl_enabled();
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cg_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// cudaFuncSetAttribute is per device: `if (cg_first_on_device(mask)) { set attributes }` runs its body once per device
// (and per kernel / template instantiation: each call site owns its `static std::atomic<unsigned long long> mask{0}`)
static inline bool cg_first_on_device(std::atomic<unsigned long long>& mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return !(mask.load(std::memory_order_acquire) & bit) && !(mask.fetch_or(bit, std::memory_order_acq_rel) & bit);
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- geometry of one convolution (forward orientation) --------------------------------
// y[n,oh,ow,co] = sum_{kh,kw,ci} x[n, oh*s+kh-pt, ow*s+kw-pl, ci] * w[kh,kw,ci,co]   (w is HWIO)
struct ConvGeom {
    int N, Hi, Wi, Cin, Ho, Wo, Cout, k, s, pt, pl;
};

#ifdef __CUDACC__
// ---- scalar / vector access helpers ---------------------------------------------------
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16(v); }

template <typename T, int VEC>
struct Pack;   // VEC elements of T moved with one (<=16 byte) access
template <>
struct Pack<float, 1> { float v[1]; };
template <>
struct Pack<bf16, 1> { bf16 v[1]; };
template <>
struct __align__(16) Pack<float, 4> { float v[4]; };
template <>
struct __align__(16) Pack<bf16, 8> { bf16 v[8]; };
template <>
struct __align__(8) Pack<bf16, 4> { bf16 v[4]; };

template <typename T, int VEC>
__device__ __forceinline__ void load_vec(const T* p, float (&out)[VEC]) {
    Pack<T, VEC> pk = *reinterpret_cast<const Pack<T, VEC>*>(p);
#pragma unroll
    for (int i = 0; i < VEC; ++i) out[i] = ldf(&pk.v[i]);
}
template <typename T, int VEC>
__device__ __forceinline__ void store_vec(T* p, const float (&in)[VEC]) {
    Pack<T, VEC> pk;
#pragma unroll
    for (int i = 0; i < VEC; ++i) stf(&pk.v[i], in[i]);
    *reinterpret_cast<Pack<T, VEC>*>(p) = pk;
}

template <typename T>
struct VecWidth;     // elements per 16-byte access
template <>
struct VecWidth<float> { static constexpr int value = 4; };
template <>
struct VecWidth<bf16> { static constexpr int value = 8; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float act_fwd(float x, int act, float slope) {
    switch (act) {
        case CG_ACT_RELU: return x > 0.f ? x : 0.f;
        case CG_ACT_LEAKY: return x > 0.f ? x : slope * x;
        case CG_ACT_TANH: return tanhf(x);
        case CG_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
        default: return x;
    }
}
// Normalisation as every forward kernel evaluates it: pre = fma(v, sc, sh) with sc = rstd*gamma, sh = fma(-mean, sc, beta),
// all three roundings explicit.  Every backward kernel derives its activation mask from the SAME expression, so the
// gradient is the gradient of exactly the function the forward pass computed (no unit near zero can be "on" forward
// and "off" backward), and the parity tests can take the masks from the stored forward outputs.
__device__ __forceinline__ void in_scale_shift(float mean, float rstd, float ga, float be, float& sc, float& sh) {
    sc = __fmul_rn(rstd, ga);
    sh = __fmaf_rn(-mean, sc, be);
}
__device__ __forceinline__ float in_pre(float v, float sc, float sh) { return __fmaf_rn(v, sc, sh); }
// (sum x, sum x^2) over P elements -> (mean, rstd), biased variance (TFA InstanceNormalization / Keras moments).  ONE
// definition with explicit roundings: in_finalize_kernel and the streaming apply kernels that finalize on the fly must
// produce the same bits, because the backward kernels rebuild the forward's activation masks from the stored pair.
__device__ __forceinline__ void in_mean_rstd(float s0, float s1, float invP, float eps, float& mean, float& rstd) {
    mean = __fmul_rn(s0, invP);
    const float var = fmaxf(__fmaf_rn(-mean, mean, __fmul_rn(s1, invP)), 0.f);
    rstd = rsqrtf(__fadd_rn(var, eps));
}

// derivative expressed through the activation OUTPUT y (all four are invertible enough for that)
__device__ __forceinline__ float act_grad_from_out(float y, int act, float slope) {
    switch (act) {
        case CG_ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case CG_ACT_LEAKY: return y > 0.f ? 1.f : slope;
        case CG_ACT_TANH: return 1.f - y * y;
        case CG_ACT_SIGMOID: return y * (1.f - y);
        default: return 1.f;
    }
}
#endif
