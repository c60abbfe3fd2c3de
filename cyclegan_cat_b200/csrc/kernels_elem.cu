// Bandwidth-bound kernels of the CycleGAN step: dtype conversion, instance norm (statistics,
// apply+activation, backward), activations, reflect padding, concat/add, pooling, the loss
// reductions (value + gradient seed) and the fused Adam update.  NHWC everywhere; 16-byte
// vector accesses whenever the channel count allows; warp-shuffle reductions.
#include "kernels.h"
#include "ptx_async.h"

static const int EW_THREADS = 256;
static inline int ew_blocks(size_t work) {
    size_t b = (work + EW_THREADS - 1) / EW_THREADS;
    size_t cap = 148 * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------
// conversions
// ------------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void convert_kernel(const TS* __restrict__ s, TD* __restrict__ d, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        stf(d + i, ldf(s + i));
}
template <typename T> int k_convert_in(const float* src, T* dst, size_t n, cudaStream_t st) {
    if (n == 0) return CG_OK;
    convert_kernel<float, T><<<ew_blocks(n), EW_THREADS, 0, st>>>(src, dst, n);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
template <typename T> int k_convert_out(const T* src, float* dst, size_t n, cudaStream_t st) {
    if (n == 0) return CG_OK;
    convert_kernel<T, float><<<ew_blocks(n), EW_THREADS, 0, st>>>(src, dst, n);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// instance norm: statistics.  x[n] is a [P][C] matrix; lanes run over channels (coalesced),
// warps and grid.y over pixels; partial sums meet in shared memory, then one atomic per
// (block, channel).  MODE 0: (sum x, sum x^2).  MODE 1 (backward): (sum g, sum g*xhat) with
// g = dy * act'(xhat*gamma+beta).
// ------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(256) in_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                        const float* __restrict__ stats,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ sums,
                                                        int P, int C, int pchunk, int act, float slope) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane, n = blockIdx.z;
    const int p0 = blockIdx.y * pchunk;
    const int p1 = min(P, p0 + pchunk);
    float s = 0.f, ss = 0.f;
    if (c < C) {
        const size_t base = (size_t)n * P * C + c;
        float mean = 0.f, rstd = 1.f, ga = 1.f, be = 0.f;
        if (MODE == 1) {
            mean = stats[((size_t)n * C + c) * 2];
            rstd = stats[((size_t)n * C + c) * 2 + 1];
            if (gamma) { ga = gamma[c]; be = beta[c]; }
        }
        float sc, sh, xsc, xsh;
        in_scale_shift(mean, rstd, ga, be, sc, sh);
        in_scale_shift(mean, rstd, 1.f, 0.f, xsc, xsh);
        int p = p0 + warp;
        for (; p + 24 < p1; p += 32) {      // 4 independent pixels in flight per lane
            float v[4], gy[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = ldf(x + base + (size_t)(p + 8 * u) * C);
                if (MODE == 1) gy[u] = ldf(dy + base + (size_t)(p + 8 * u) * C);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (MODE == 0) { s += v[u]; ss += v[u] * v[u]; }
                else {
                    float xh = in_pre(v[u], xsc, xsh);
                    float g = gy[u] * act_grad_from_out(in_pre(v[u], sc, sh), act, slope);
                    s += g;
                    ss += g * xh;
                }
            }
        }
        for (; p < p1; p += 8) {
            float v = ldf(x + base + (size_t)p * C);
            if (MODE == 0) {
                s += v;
                ss += v * v;
            } else {
                float xh = in_pre(v, xsc, xsh);
                float g = ldf(dy + base + (size_t)p * C) * act_grad_from_out(in_pre(v, sc, sh), act, slope);
                s += g;
                ss += g * xh;
            }
        }
    }
    __shared__ float sh[2][8][33];
    sh[0][warp][lane] = s;
    sh[1][warp][lane] = ss;
    __syncthreads();
    if (warp == 0 && c < C) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a += sh[0][w][lane]; b += sh[1][w][lane]; }
        atomicAdd(&sums[((size_t)n * C + c) * 2], a);
        atomicAdd(&sums[((size_t)n * C + c) * 2 + 1], b);
    }
}

__global__ void in_finalize_kernel(const float* in, float* out, int NC, float invP, float eps) {
    griddep_launch();        // PDL: see ptx_async.h
    griddep_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NC) return;
    float mean, rstd;
    in_mean_rstd(in[2 * i], in[2 * i + 1], invP, eps, mean, rstd);
    out[2 * i] = mean;
    out[2 * i + 1] = rstd;
}

static inline void in_reduce_grid(int N, int P, int C, dim3& grid, int& pchunk) {
    int cb = cdiv(C, 32);
    int want = cdiv(148 * 4, (long long)cb * N);        // pixel splits so the grid fills the GPU
    int maxsplit = cdiv(P, 64);
    int ps = want < 1 ? 1 : (want > maxsplit ? maxsplit : want);
    pchunk = cdiv(P, ps);
    ps = cdiv(P, pchunk);
    grid = dim3(cb, ps, N);
}

int k_in_finalize(const float* raw, float* stats, int NC, int P, float eps, cudaStream_t st) {
    launch_pdl(in_finalize_kernel, dim3(cdiv(NC, 256)), dim3(256), 0, st, raw, stats, NC, 1.f / (float)P, eps);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// raw (sum x, sum x^2) per (sample, channel); k_in_stats turns them into (mean, rstd), k_bn_finalize pools them first
template <typename T> int k_in_stats_raw(const T* x, float* stats, int N, int P, int C, cudaStream_t st, bool zeroed) {
    if (!zeroed) CG_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * (size_t)N * C, st));
    if (k_in_stream_ok<T>(x, nullptr, nullptr, P, C)) return k_in_stats_stream<T>(x, stats, N, P, C, st);
    dim3 grid; int pchunk;
    in_reduce_grid(N, P, C, grid, pchunk);
    in_reduce_kernel<T, 0><<<grid, 256, 0, st>>>(x, nullptr, nullptr, nullptr, nullptr, stats, P, C, pchunk, 0, 0.f);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
template <typename T> int k_in_stats(const T* x, float* stats, int N, int P, int C, float eps, cudaStream_t st, bool zeroed) {
    CG_TRY(k_in_stats_raw<T>(x, stats, N, P, C, st, zeroed));
    return k_in_finalize(stats, stats, N * C, P, eps, st);
}

// y = act((x - mean) * rstd * gamma + beta); grid.y = sample, x-dimension strides over P*C/VEC
template <typename T, int VEC>
__global__ void in_apply_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                float slope, size_t PCv, int C) {
    const int n = blockIdx.y;
    const float* st = stats + (size_t)n * C * 2;
    const size_t base = (size_t)n * PCv * VEC;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < PCv; i += (size_t)gridDim.x * blockDim.x) {
        size_t e = i * VEC;
        int c = (int)(e % C);
        float v[VEC];
        load_vec<T, VEC>(x + base + e, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float mean = st[(c + j) * 2], rstd = st[(c + j) * 2 + 1];
            float ga = gamma ? gamma[c + j] : 1.f, be = beta ? beta[c + j] : 0.f;
            float sc, sh;
            in_scale_shift(mean, rstd, ga, be, sc, sh);
            v[j] = act_fwd(in_pre(v[j], sc, sh), act, slope);
        }
        store_vec<T, VEC>(y + base + e, v);
    }
}

// ------------------------------------------------------------------------------------------
// "fast" instance-norm elementwise kernels: when the number of 16-byte channel vectors per pixel (CV = C/VEC) divides
// the block size, every thread owns ONE channel vector for its whole life, so the per-channel constants (mean, rstd,
// gamma, beta, backward sums) sit in registers and the loop body is load -> FMA -> store.  (The generic kernels above
// re-read 2-6 per-channel scalars per element and are LSU-bound at ~1.3 TB/s.)
// grid = (pixel blocks, N); block = 256 threads = (256/CV) pixels x CV channel vectors.
// ------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) in_apply_fast_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                            const float* __restrict__ stats,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int act, float slope, int P,
                                                            int C) {
    const int CV = C / VEC, n = blockIdx.y;
    const int cv = threadIdx.x % CV, prow = threadIdx.x / CV, ppb = 256 / CV;
    float sc[VEC], sh[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int c = cv * VEC + j;
        const float mean = stats[((size_t)n * C + c) * 2], rstd = stats[((size_t)n * C + c) * 2 + 1];
        const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
        in_scale_shift(mean, rstd, ga, be, sc[j], sh[j]);
    }
    const size_t base = (size_t)n * P * C + (size_t)cv * VEC;
    const int stride = gridDim.x * ppb;
    int p = blockIdx.x * ppb + prow;
    for (; p + 3 * stride < P; p += 4 * stride) {       // 4 independent 16-byte loads in flight per thread
        float v[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; ++u) load_vec<T, VEC>(x + base + (size_t)(p + u * stride) * C, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[u][j] = act_fwd(in_pre(v[u][j], sc[j], sh[j]), act, slope);
            store_vec<T, VEC>(y + base + (size_t)(p + u * stride) * C, v[u]);
        }
    }
    for (; p < P; p += stride) {
        float v[VEC];
        load_vec<T, VEC>(x + base + (size_t)p * C, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] = act_fwd(in_pre(v[j], sc[j], sh[j]), act, slope);
        store_vec<T, VEC>(y + base + (size_t)p * C, v);
    }
}

// dx = rstd*gamma*(g - S1/P - xhat*S2/P), optionally into a zero-bordered [H+2h][W+2h] buffer (halo > 0).
// Per-channel constants are folded to keep the register count (and with it the occupancy) in check:
//   xhat = v*a + b,  r = k*gg - c1 - xhat*c2   with a = rstd, b = -mean*rstd, k = rstd*gamma, c1 = k*S1/P, c2 = k*S2/P
template <typename T, int VEC, bool AFFINE>
__global__ void __launch_bounds__(256, AFFINE ? 2 : 4) in_bwd_apply_fast_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                T* __restrict__ dx, const float* __restrict__ stats,
                                                                const float* __restrict__ sums,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, int act, float slope,
                                                                int H, int W, int C, int halo, float invP,
                                                                int accumulate) {
    const int CV = C / VEC, n = blockIdx.y;
    const int cv = threadIdx.x % CV, prow = threadIdx.x / CV, ppb = 256 / CV;
    float ca[VEC], cb[VEC], ck[VEC], c1[VEC], c2[VEC], cg[AFFINE ? VEC : 1], ce[AFFINE ? VEC : 1];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        const int c = cv * VEC + j;
        const float mean = stats[((size_t)n * C + c) * 2], rstd = stats[((size_t)n * C + c) * 2 + 1];
        const float ga = AFFINE ? gamma[c] : 1.f;
        in_scale_shift(mean, rstd, 1.f, 0.f, ca[j], cb[j]);
        ck[j] = rstd * ga;
        c1[j] = ck[j] * sums[((size_t)n * C + c) * 2] * invP;
        c2[j] = ck[j] * sums[((size_t)n * C + c) * 2 + 1] * invP;
        if (AFFINE) in_scale_shift(mean, rstd, ga, beta[c], cg[j], ce[j]);
    }
    const int Hp = H + 2 * halo, Wp = W + 2 * halo;
    const int PP = Hp * Wp;
    const size_t in_base = (size_t)n * H * W * C + (size_t)cv * VEC;
    const size_t out_base = (size_t)n * PP * C + (size_t)cv * VEC;
    for (int pp = blockIdx.x * ppb + prow; pp < PP; pp += gridDim.x * ppb) {
        int h = 0, w = pp;
        if (halo > 0) { const int hp = pp / Wp; w = pp - hp * Wp - halo; h = hp - halo; }
        float o[VEC];
        if (h >= 0 && h < H && w >= 0 && w < W) {
            const size_t e = in_base + ((size_t)h * W + w) * C;
            float v[VEC], g[VEC];
            load_vec<T, VEC>(x + e, v);
            load_vec<T, VEC>(dy + e, g);
            if (accumulate) load_vec<T, VEC>(dx + out_base + (size_t)pp * C, o);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float xh = in_pre(v[j], ca[j], cb[j]);
                const float pre = AFFINE ? in_pre(v[j], cg[j], ce[j]) : xh;
                const float gg = g[j] * act_grad_from_out(pre, act, slope);
                const float r = fmaf(ck[j], gg, -fmaf(xh, c2[j], c1[j]));
                o[j] = accumulate ? o[j] + r : r;
            }
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) o[j] = 0.f;
        }
        store_vec<T, VEC>(dx + out_base + (size_t)pp * C, o);
    }
}

static inline bool fast_cv_ok(int C, int VW) {
    if (C % VW) return false;
    const int cv = C / VW;
    return cv <= 256 && (256 % cv) == 0;
}
static inline int fast_blocks(int P, int C, int VW, int N) {
    const int ppb = 256 / (C / VW);
    long long want = (148LL * 8 + N - 1) / N;           // ~8 blocks per SM over all samples
    long long maxb = (P + ppb - 1) / ppb;
    long long b = want < maxb ? want : maxb;
    return (int)(b < 1 ? 1 : b);
}

template <typename T> int k_in_apply(const T* x, T* y, const float* stats, const float* gamma, const float* beta,
                                     int act, float slope, int N, int P, int C, cudaStream_t st) {
    constexpr int VW = VecWidth<T>::value;
    size_t PC = (size_t)P * C;
    if (k_in_stream_ok<T>(x, y, nullptr, P, C))
        return k_in_apply_stream<T>(x, nullptr, y, nullptr, const_cast<float*>(stats), gamma, beta, act, slope, N, P, C, 0, 0, st);
    if (fast_cv_ok(C, VW)) {
        dim3 grid(fast_blocks(P, C, VW, N), N);
        in_apply_fast_kernel<T, VW><<<grid, 256, 0, st>>>(x, y, stats, gamma, beta, act, slope, P, C);
    } else if (C % VW == 0) {
        dim3 grid(ew_blocks(PC / VW), N);
        in_apply_kernel<T, VW><<<grid, EW_THREADS, 0, st>>>(x, y, stats, gamma, beta, act, slope, PC / VW, C);
    } else {
        dim3 grid(ew_blocks(PC), N);
        in_apply_kernel<T, 1><<<grid, EW_THREADS, 0, st>>>(x, y, stats, gamma, beta, act, slope, PC, C);
    }
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// dx (+)= rstd*gamma*(g - S1/P - xhat*S2/P)
template <typename T, int VEC>
__global__ void in_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                    const float* __restrict__ stats, const float* __restrict__ sums,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                    float slope, size_t PCv, int C, float invP, int accumulate) {
    const int n = blockIdx.y;
    const float* st = stats + (size_t)n * C * 2;
    const float* sm = sums + (size_t)n * C * 2;
    const size_t base = (size_t)n * PCv * VEC;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < PCv; i += (size_t)gridDim.x * blockDim.x) {
        size_t e = i * VEC;
        int c = (int)(e % C);
        float v[VEC], g[VEC], o[VEC];
        load_vec<T, VEC>(x + base + e, v);
        load_vec<T, VEC>(dy + base + e, g);
        if (accumulate) load_vec<T, VEC>(dx + base + e, o);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float mean = st[(c + j) * 2], rstd = st[(c + j) * 2 + 1];
            float ga = gamma ? gamma[c + j] : 1.f, be = beta ? beta[c + j] : 0.f;
            float sc, sh, xsc, xsh;
            in_scale_shift(mean, rstd, ga, be, sc, sh);
            in_scale_shift(mean, rstd, 1.f, 0.f, xsc, xsh);
            float xh = in_pre(v[j], xsc, xsh);
            float gg = g[j] * act_grad_from_out(in_pre(v[j], sc, sh), act, slope);
            float r = rstd * ga * (gg - sm[(c + j) * 2] * invP - xh * sm[(c + j) * 2 + 1] * invP);
            o[j] = accumulate ? o[j] + r : r;
        }
        store_vec<T, VEC>(dx + base + e, o);
    }
}

// same, but dx is a zero-bordered buffer [N][H+2*halo][W+2*halo][C] (the layout the tensor-core data-gradient and
// weight-gradient kernels read with TMA); this kernel writes the zero border too.
template <typename T, int VEC>
__global__ void in_bwd_apply_halo_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                         const float* __restrict__ stats, const float* __restrict__ sums,
                                         const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                         float slope, int H, int W, int C, int halo, float invP) {
    const int n = blockIdx.y;
    const int Cv = C / VEC, Hp = H + 2 * halo, Wp = W + 2 * halo;
    const float* st = stats + (size_t)n * C * 2;
    const float* sm = sums + (size_t)n * C * 2;
    const size_t total = (size_t)Hp * Wp * Cv;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % Cv);
        size_t r = i / Cv;
        const int wp = (int)(r % Wp), hp = (int)(r / Wp);
        const int h = hp - halo, w = wp - halo, c = cv * VEC;
        float o[VEC];
        if (h >= 0 && h < H && w >= 0 && w < W) {
            const size_t e = (((size_t)n * H + h) * W + w) * C + c;
            float v[VEC], g[VEC];
            load_vec<T, VEC>(x + e, v);
            load_vec<T, VEC>(dy + e, g);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                float mean = st[(c + j) * 2], rstd = st[(c + j) * 2 + 1];
                float ga = gamma ? gamma[c + j] : 1.f, be = beta ? beta[c + j] : 0.f;
                float sc, sh, xsc, xsh;
                in_scale_shift(mean, rstd, ga, be, sc, sh);
                in_scale_shift(mean, rstd, 1.f, 0.f, xsc, xsh);
                float xh = in_pre(v[j], xsc, xsh);
                float gg = g[j] * act_grad_from_out(in_pre(v[j], sc, sh), act, slope);
                o[j] = rstd * ga * (gg - sm[(c + j) * 2] * invP - xh * sm[(c + j) * 2 + 1] * invP);
            }
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) o[j] = 0.f;
        }
        store_vec<T, VEC>(dx + (((size_t)n * Hp + hp) * Wp + wp) * C + c, o);
    }
}

__global__ void in_param_grad_kernel(const float* __restrict__ sums, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int N, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s1 = 0.f, s2 = 0.f;
    for (int n = 0; n < N; ++n) { s1 += sums[((size_t)n * C + c) * 2]; s2 += sums[((size_t)n * C + c) * 2 + 1]; }
    dbeta[c] += s1;
    dgamma[c] += s2;
}

template <typename T> int k_in_bwd(const T* x, const T* dy, T* dx, const float* stats, const float* gamma,
                                   const float* beta, float* dgamma, float* dbeta, float* scratch, int act,
                                   float slope, int N, int P, int C, int accumulate, cudaStream_t st, int halo, int W,
                                   bool zeroed, int bn_group) {
    if (!zeroed) CG_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * (size_t)N * C, st));
    // (a register-resident, 16-byte-vector variant of this reduction was measured slower -- 51 us vs 43 us per trunk
    // launch -- than the lanes-over-channels kernel with its 6 resident blocks per SM, so the simple kernel stays)
    const bool stream = k_in_stream_ok<T>(x, dy, dx, P, C);
    if (stream) {
        CG_TRY(k_in_bwd_reduce_stream<T>(x, dy, stats, gamma, beta, scratch, act, slope, N, P, C, st));
    } else {
        dim3 grid; int pchunk;
        in_reduce_grid(N, P, C, grid, pchunk);
        in_reduce_kernel<T, 1><<<grid, 256, 0, st>>>(x, dy, stats, gamma, beta, scratch, P, C, pchunk, act, slope);
        CG_LAUNCH_CHECK();
    }
    if (bn_group > 0) CG_TRY(k_bn_pool_sums(scratch, N, C, bn_group, st));
    const bool stream_apply = dx && stream && !accumulate && (halo == 0 || (W > 0 && P % W == 0));
    // the streaming apply adds d gamma / d beta itself; BatchNormalization (pooled sums spread over the samples of a group)
    // and the non-streaming fallbacks keep the separate kernel
    const bool fold_pg = dgamma && gamma && stream_apply && bn_group <= 0;
    if (dgamma && !fold_pg) {
        in_param_grad_kernel<<<cdiv(C, 128), 128, 0, st>>>(scratch, dgamma, dbeta, N, C);
        CG_LAUNCH_CHECK();
    }
    if (stream_apply) {
        CG_TRY(k_in_bwd_apply_stream<T>(x, dy, dx, stats, scratch, gamma, beta, act, slope, N, P, C, W, halo, st,
                                        fold_pg ? dgamma : nullptr, fold_pg ? dbeta : nullptr));
    } else if (dx && fast_cv_ok(C, VecWidth<T>::value) && (halo == 0 || (!accumulate && W > 0 && P % W == 0))) {
        constexpr int VW = VecWidth<T>::value;
        const int Wd = halo > 0 ? W : P, Hd = halo > 0 ? P / W : 1;     // halo == 0: treat the plane as one row
        dim3 g2(fast_blocks((Hd + 2 * halo) * (Wd + 2 * halo), C, VW, N), N);
        if (gamma)
            in_bwd_apply_fast_kernel<T, VW, true><<<g2, 256, 0, st>>>(x, dy, dx, stats, scratch, gamma, beta, act, slope, Hd,
                                                                     Wd, C, halo, 1.f / (float)P, accumulate);
        else
            in_bwd_apply_fast_kernel<T, VW, false><<<g2, 256, 0, st>>>(x, dy, dx, stats, scratch, gamma, beta, act, slope, Hd,
                                                                      Wd, C, halo, 1.f / (float)P, accumulate);
        CG_LAUNCH_CHECK();
    } else if (dx && halo > 0) {
        constexpr int VW = VecWidth<T>::value;
        if (accumulate || C % VW != 0 || W <= 0 || P % W != 0) {
            cg_set_error("halo IN backward: unsupported configuration");
            return CG_ERR_INVALID;
        }
        const int H = P / W;
        dim3 g2(ew_blocks((size_t)(H + 2 * halo) * (W + 2 * halo) * C / VW), N);
        in_bwd_apply_halo_kernel<T, VW><<<g2, EW_THREADS, 0, st>>>(x, dy, dx, stats, scratch, gamma, beta, act, slope, H,
                                                                   W, C, halo, 1.f / (float)P);
        CG_LAUNCH_CHECK();
    } else if (dx) {
        constexpr int VW = VecWidth<T>::value;
        size_t PC = (size_t)P * C;
        if (C % VW == 0) {
            dim3 g2(ew_blocks(PC / VW), N);
            in_bwd_apply_kernel<T, VW><<<g2, EW_THREADS, 0, st>>>(x, dy, dx, stats, scratch, gamma, beta, act, slope,
                                                                  PC / VW, C, 1.f / (float)P, accumulate);
        } else {
            dim3 g2(ew_blocks(PC), N);
            in_bwd_apply_kernel<T, 1><<<g2, EW_THREADS, 0, st>>>(x, dy, dx, stats, scratch, gamma, beta, act, slope, PC,
                                                                 C, 1.f / (float)P, accumulate);
        }
        CG_LAUNCH_CHECK();
    }
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// flat elementwise ops
// ------------------------------------------------------------------------------------------
enum { EW_ACT_FWD = 0, EW_ACT_BWD = 1, EW_ADD = 2, EW_COPY = 3 };

template <typename T, int VEC, int OP>
__global__ void ew_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, size_t nv, int act,
                          float slope, int accumulate) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x) {
        size_t e = i * VEC;
        float x[VEC], y[VEC], r[VEC];
        load_vec<T, VEC>(a + e, x);
        if (OP == EW_ACT_BWD || OP == EW_ADD) load_vec<T, VEC>(b + e, y);
        if (accumulate) load_vec<T, VEC>(o + e, r);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float v;
            if (OP == EW_ACT_FWD) v = act_fwd(x[j], act, slope);
            else if (OP == EW_ACT_BWD) v = y[j] * act_grad_from_out(x[j], act, slope);
            else if (OP == EW_ADD) v = x[j] + y[j];
            else v = x[j];
            r[j] = accumulate ? r[j] + v : v;
        }
        store_vec<T, VEC>(o + e, r);
    }
}

template <typename T, int OP>
static int ew_launch(const T* a, const T* b, T* o, size_t n, int act, float slope, int accumulate, cudaStream_t st) {
    if (n == 0) return CG_OK;
    constexpr int VW = VecWidth<T>::value;
    bool aligned = ((uintptr_t)a % 16 == 0) && ((uintptr_t)o % 16 == 0) && (b == nullptr || (uintptr_t)b % 16 == 0);
    if (n % VW == 0 && aligned)
        ew_kernel<T, VW, OP><<<ew_blocks(n / VW), EW_THREADS, 0, st>>>(a, b, o, n / VW, act, slope, accumulate);
    else
        ew_kernel<T, 1, OP><<<ew_blocks(n), EW_THREADS, 0, st>>>(a, b, o, n, act, slope, accumulate);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

template <typename T> int k_act_fwd(const T* x, T* y, size_t n, int act, float slope, cudaStream_t st) {
    return ew_launch<T, EW_ACT_FWD>(x, nullptr, y, n, act, slope, 0, st);
}
template <typename T> int k_act_bwd(const T* y, const T* dy, T* dx, size_t n, int act, float slope, int accumulate,
                                    cudaStream_t st) {
    return ew_launch<T, EW_ACT_BWD>(y, dy, dx, n, act, slope, accumulate, st);
}
template <typename T> int k_add(const T* a, const T* b, T* y, size_t n, cudaStream_t st) {
    return ew_launch<T, EW_ADD>(a, b, y, n, 0, 0.f, 0, st);
}
template <typename T> int k_copy_acc(const T* src, T* dst, size_t n, int accumulate, cudaStream_t st) {
    return ew_launch<T, EW_COPY>(src, nullptr, dst, n, 0, 0.f, accumulate, st);
}

// ------------------------------------------------------------------------------------------
// reflect padding (tf.pad REFLECT: mirror without repeating the edge) and its adjoint
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

template <typename T, int VEC>
__global__ void rpad_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int Cv, int p) {
    const int Ho = H + 2 * p, Wo = W + 2 * p;
    const size_t total = (size_t)N * Ho * Wo * Cv;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int cv = (int)(i % Cv);
        size_t r = i / Cv;
        int wo = (int)(r % Wo); r /= Wo;
        int ho = (int)(r % Ho);
        int n = (int)(r / Ho);
        int h = reflect_idx(ho - p, H), w = reflect_idx(wo - p, W);
        float v[VEC];
        load_vec<T, VEC>(x + (((size_t)n * H + h) * W + w) * Cv * VEC + (size_t)cv * VEC, v);
        store_vec<T, VEC>(y + i * VEC, v);
    }
}

// gather form of the adjoint: interior pixel h receives padded rows {h+p} U mirrors
__device__ __forceinline__ int reflect_sources(int i, int n, int p, int (&src)[3]) {
    int cnt = 0;
    src[cnt++] = i + p;
    if (i >= 1 && i <= p) src[cnt++] = p - i;
    if (i <= n - 2 && i >= n - 1 - p) src[cnt++] = 2 * (n - 1) - i + p;
    return cnt;
}

template <typename T, int VEC>
__global__ void rpad_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int Cv, int p,
                                int accumulate) {
    const int Ho = H + 2 * p, Wo = W + 2 * p;
    const size_t total = (size_t)N * H * W * Cv;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int cv = (int)(i % Cv);
        size_t r = i / Cv;
        int w = (int)(r % W); r /= W;
        int h = (int)(r % H);
        int n = (int)(r / H);
        int hs[3], ws[3];
        int nh = reflect_sources(h, H, p, hs), nw = reflect_sources(w, W, p, ws);
        float acc[VEC];
        if (accumulate) load_vec<T, VEC>(dx + i * VEC, acc);
        else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
        }
        for (int a = 0; a < nh; ++a)
            for (int b = 0; b < nw; ++b) {
                float v[VEC];
                load_vec<T, VEC>(dy + (((size_t)n * Ho + hs[a]) * Wo + ws[b]) * Cv * VEC + (size_t)cv * VEC, v);
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[j] += v[j];
            }
        store_vec<T, VEC>(dx + i * VEC, acc);
    }
}

// Border-only adjoint: dx already holds the interior part (padded pixel (h+p, w+p)) of every pixel; add the mirror terms.
// Only interior pixels within p of an edge (but not on it) have any: rows 1..p and H-1-p..H-2 completely, and of the other
// rows the columns 1..p and W-1-p..W-2.  One thread = one 16-byte vector of one affected pixel.
template <typename T, int VEC>
__global__ void rpad_bwd_border_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int Cv, int p) {
    griddep_launch();        // PDL: see ptx_async.h (this kernel sits between the data-gradient conv and the next norm backward)
    griddep_wait();
    const int Ho = H + 2 * p, Wo = W + 2 * p;
    const int full_rows = 2 * p, side_rows = H - 2 * p;       // affected pixels per image: full_rows*W + side_rows*2p
    const int per_img = full_rows * W + side_rows * 2 * p;
    const size_t total = (size_t)N * per_img * Cv;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % Cv);
        size_t r = i / Cv;
        int k = (int)(r % per_img);
        const int n = (int)(r / per_img);
        int h, w;
        if (k < full_rows * W) {
            const int fr = k / W;
            w = k - fr * W;
            h = fr < p ? 1 + fr : H - 1 - p + (fr - p);
        } else {
            k -= full_rows * W;
            const int sr = k / (2 * p), j = k - sr * 2 * p;
            // the side rows are the rows that are not "full" rows: 0, p+1 .. H-2-p, H-1
            h = sr == 0 ? 0 : (sr == side_rows - 1 ? H - 1 : p + sr);
            w = j < p ? 1 + j : W - 1 - p + (j - p);
        }
        int hs[3], ws[3];
        const int nh = reflect_sources(h, H, p, hs), nw = reflect_sources(w, W, p, ws);
        if (nh + nw == 2) continue;
        float acc[VEC];
        T* d = dx + ((((size_t)n * H + h) * W + w) * Cv + cv) * VEC;
        load_vec<T, VEC>(d, acc);
        for (int a = 0; a < nh; ++a)
            for (int b = 0; b < nw; ++b) {
                if (a + b == 0) continue;
                float v[VEC];
                load_vec<T, VEC>(dy + (((size_t)n * Ho + hs[a]) * Wo + ws[b]) * Cv * VEC + (size_t)cv * VEC, v);
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[j] += v[j];
            }
        store_vec<T, VEC>(d, acc);
    }
}
template <typename T> int k_rpad_bwd_border(const T* dy, T* dx, int N, int H, int W, int C, int p, cudaStream_t st) {
    constexpr int VW = VecWidth<T>::value;
    if (C % VW || H <= 2 * p + 1 || W <= 2 * p + 1) { cg_set_error("rpad_bwd_border: unsupported shape"); return CG_ERR_INVALID; }
    const size_t total = (size_t)N * (2 * p * W + (H - 2 * p) * 2 * p) * (C / VW);
    launch_pdl(rpad_bwd_border_kernel<T, VW>, dim3(ew_blocks(total)), dim3(EW_THREADS), 0, st, dy, dx, N, H, W, C / VW, p);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

template <typename T> int k_rpad_fwd(const T* x, T* y, int N, int H, int W, int C, int p, cudaStream_t st) {
    constexpr int VW = VecWidth<T>::value;
    size_t total = (size_t)N * (H + 2 * p) * (W + 2 * p) * C;
    if (C % VW == 0) rpad_fwd_kernel<T, VW><<<ew_blocks(total / VW), EW_THREADS, 0, st>>>(x, y, N, H, W, C / VW, p);
    else rpad_fwd_kernel<T, 1><<<ew_blocks(total), EW_THREADS, 0, st>>>(x, y, N, H, W, C, p);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
template <typename T> int k_rpad_bwd(const T* dy, T* dx, int N, int H, int W, int C, int p, int accumulate,
                                     cudaStream_t st) {
    constexpr int VW = VecWidth<T>::value;
    size_t total = (size_t)N * H * W * C;
    if (C % VW == 0)
        rpad_bwd_kernel<T, VW><<<ew_blocks(total / VW), EW_THREADS, 0, st>>>(dy, dx, N, H, W, C / VW, p, accumulate);
    else rpad_bwd_kernel<T, 1><<<ew_blocks(total), EW_THREADS, 0, st>>>(dy, dx, N, H, W, C, p, accumulate);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// channel-slice copy: dst[pix, doff:doff+Cc] (+)= src[pix, so:so+Cc]   (concat forward / backward)
// ------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void slice_copy_kernel(const T* __restrict__ src, int Cs, int so, T* __restrict__ dst, int Cd, int doff,
                                  int Ccv, size_t npix, int accumulate) {
    const size_t total = npix * Ccv;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int cv = (int)(i % Ccv);
        size_t pix = i / Ccv;
        float v[VEC], o[VEC];
        load_vec<T, VEC>(src + pix * Cs + so + (size_t)cv * VEC, v);
        T* d = dst + pix * Cd + doff + (size_t)cv * VEC;
        if (accumulate) {
            load_vec<T, VEC>(d, o);
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] += o[j];
        }
        store_vec<T, VEC>(d, v);
    }
}
template <typename T> int k_slice_copy(const T* src, int Cs, int so, T* dst, int Cd, int doff, int Cc, size_t npix,
                                       int accumulate, cudaStream_t st) {
    constexpr int VW = VecWidth<T>::value;
    if (Cs % VW == 0 && so % VW == 0 && Cd % VW == 0 && doff % VW == 0 && Cc % VW == 0)
        slice_copy_kernel<T, VW><<<ew_blocks(npix * Cc / VW), EW_THREADS, 0, st>>>(src, Cs, so, dst, Cd, doff, Cc / VW,
                                                                                    npix, accumulate);
    else
        slice_copy_kernel<T, 1><<<ew_blocks(npix * Cc), EW_THREADS, 0, st>>>(src, Cs, so, dst, Cd, doff, Cc, npix,
                                                                              accumulate);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// 2x2 average pooling / nearest x2 upsampling and their adjoints.  MODE: 0 pool fwd, 1 pool bwd,
// 2 upsample fwd, 3 upsample bwd.  (H, W) is always the LARGE grid; the small one is H/2 x W/2.
// ------------------------------------------------------------------------------------------
// Zero-copy concat (unet.py:109, `Concatenate()([skip, x])`): `cat` is the concat tensor (or its gradient) with `cat_cv`
// channel vectors per pixel, the slice of interest starting at vector `cat_ov`.
//   MODE 0 + cat : the pool also writes the copy of its input (the skip) into the concat tensor -- it reads all of it anyway
//   MODE 1 + cat : the pool's backward adds the skip slice of the concat gradient, so dSkip is written in one pass
//   MODE 2 + cat : the upsample writes straight into its slice of the concat tensor (`out` unused)
//   MODE 3 + cat : the upsample's backward gathers from its slice of the concat gradient (`in` unused)
template <typename T, int VEC, int MODE>
__global__ void resample_kernel(const T* __restrict__ in, T* __restrict__ out, int N, int H, int W, int Cv,
                                int accumulate, T* __restrict__ cat, int cat_cv, int cat_ov) {
    const int Hs = H / 2, Ws = W / 2;
    const bool out_small = (MODE == 0 || MODE == 3);
    const int Ho = out_small ? Hs : H, Wo = out_small ? Ws : W;
    const size_t total = (size_t)N * Ho * Wo * Cv;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int cv = (int)(i % Cv);
        size_t r = i / Cv;
        int w = (int)(r % Wo); r /= Wo;
        int h = (int)(r % Ho);
        int n = (int)(r / Ho);
        float acc[VEC];
        if (out_small) {   // gather 2x2 from the large grid
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b) {
                    float v[VEC];
                    const size_t pix = ((size_t)n * H + 2 * h + a) * W + 2 * w + b;
                    if (MODE == 3 && cat) load_vec<T, VEC>(cat + (pix * cat_cv + cat_ov + cv) * VEC, v);
                    else load_vec<T, VEC>(in + (pix * Cv + cv) * VEC, v);
                    if (MODE == 0 && cat) store_vec<T, VEC>(cat + (pix * cat_cv + cat_ov + cv) * VEC, v);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) acc[j] += v[j];
                }
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[j] *= 0.25f;
            }
        } else {           // broadcast from the small grid
            load_vec<T, VEC>(in + ((((size_t)n * Hs + h / 2) * Ws + w / 2) * Cv + cv) * VEC, acc);
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[j] *= 0.25f;
                if (cat) {
                    float v[VEC];
                    load_vec<T, VEC>(cat + ((r * Wo + w) * cat_cv + cat_ov + cv) * VEC, v);      // r = n * Ho + h
#pragma unroll
                    for (int j = 0; j < VEC; ++j) acc[j] += v[j];
                }
            }
        }
        if (MODE == 2 && cat) {
            store_vec<T, VEC>(cat + ((r * Wo + w) * cat_cv + cat_ov + cv) * VEC, acc);
            continue;
        }
        if (accumulate) {
            float o[VEC];
            load_vec<T, VEC>(out + i * VEC, o);
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] += o[j];
        }
        store_vec<T, VEC>(out + i * VEC, acc);
    }
}
template <typename T, int MODE>
static int resample_launch(const T* in, T* out, int N, int H, int W, int C, int accumulate, cudaStream_t st,
                           T* cat = nullptr, int cat_c = 0, int cat_off = 0) {
    constexpr int VW = VecWidth<T>::value;
    const bool out_small = (MODE == 0 || MODE == 3);
    size_t total = (size_t)N * (out_small ? (H / 2) * (W / 2) : H * W) * C;
    if (C % VW == 0 && cat_c % VW == 0 && cat_off % VW == 0)
        resample_kernel<T, VW, MODE><<<ew_blocks(total / VW), EW_THREADS, 0, st>>>(in, out, N, H, W, C / VW, accumulate, cat,
                                                                                  cat_c / VW, cat_off / VW);
    else resample_kernel<T, 1, MODE><<<ew_blocks(total), EW_THREADS, 0, st>>>(in, out, N, H, W, C, accumulate, cat, cat_c, cat_off);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
// H, W are the INPUT sizes of the forward op
template <typename T> int k_avgpool_fwd(const T* x, T* y, int N, int H, int W, int C, cudaStream_t st, T* cat, int cat_c,
                                        int cat_off) {
    return resample_launch<T, 0>(x, y, N, H, W, C, 0, st, cat, cat_c, cat_off);
}
template <typename T> int k_avgpool_bwd(const T* dy, T* dx, int N, int H, int W, int C, int accumulate, cudaStream_t st,
                                        const T* dcat, int cat_c, int cat_off) {
    return resample_launch<T, 1>(dy, dx, N, H, W, C, accumulate, st, const_cast<T*>(dcat), cat_c, cat_off);
}
template <typename T> int k_upsample_fwd(const T* x, T* y, int N, int H, int W, int C, cudaStream_t st, T* cat, int cat_c,
                                         int cat_off) {
    return resample_launch<T, 2>(x, y, N, 2 * H, 2 * W, C, 0, st, cat, cat_c, cat_off);
}
template <typename T> int k_upsample_bwd(const T* dy, T* dx, int N, int H, int W, int C, int accumulate, cudaStream_t st,
                                         const T* dcat, int cat_c, int cat_off) {
    return resample_launch<T, 3>(dy, dx, N, 2 * H, 2 * W, C, accumulate, st, const_cast<T*>(dcat), cat_c, cat_off);
}

// ------------------------------------------------------------------------------------------
// losses: block reduction with warp shuffles, one atomic per block
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_atomic_add2(float a, float b, float* pa, float* pb) {
    __shared__ float sh[2][32];
    a = warp_sum(a);
    b = warp_sum(b);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; }
    __syncthreads();
    if (warp == 0) {
        int nw = blockDim.x >> 5;
        a = lane < nw ? sh[0][lane] : 0.f;
        b = lane < nw ? sh[1][lane] : 0.f;
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) {
            if (pa) atomicAdd(pa, a);
            if (pb) atomicAdd(pb, b);
        }
    }
}

template <typename T>
__global__ void adv_loss_kernel(const T* __restrict__ d, size_t n, float target, int kind, float grad_scale,
                                T* __restrict__ grad, float* sum_out, float* correct_out) {
    float s = 0.f, c = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float x = ldf(d + i), l, g;
        if (kind == CG_LOSS_MSE) { float e = x - target; l = e * e; g = 2.f * e; }
        else if (kind == CG_LOSS_MAE) { float e = x - target; l = fabsf(e); g = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f); }
        else { l = fmaxf(x, 0.f) - x * target + log1pf(expf(-fabsf(x))); g = 1.f / (1.f + expf(-x)) - target; }
        s += l;
        c += ((x > 0.5f ? 1.f : 0.f) == target) ? 1.f : 0.f;
        if (grad) stf(grad + i, grad_scale * g);
    }
    block_atomic_add2(s, c, sum_out, correct_out);
}
template <typename T> int k_adv_loss(const T* d, size_t n, float target, int kind, float grad_scale, T* grad,
                                     float* sum_out, float* correct_out, cudaStream_t st) {
    adv_loss_kernel<T><<<ew_blocks(n), EW_THREADS, 0, st>>>(d, n, target, kind, grad_scale, grad, sum_out, correct_out);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

template <typename T>
__global__ void l1_loss_kernel(const T* __restrict__ real, const T* __restrict__ gen, size_t n, float grad_scale,
                               T* __restrict__ grad, int accumulate, float* sum_out) {
    float s = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float e = ldf(gen + i) - ldf(real + i);
        s += fabsf(e);
        if (grad) {
            float g = grad_scale * (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f));
            stf(grad + i, accumulate ? ldf(grad + i) + g : g);
        }
    }
    block_atomic_add2(s, 0.f, sum_out, nullptr);
}
template <typename T> int k_l1_loss(const T* real, const T* gen, size_t n, float grad_scale, T* grad, int accumulate,
                                    float* sum_out, cudaStream_t st) {
    l1_loss_kernel<T><<<ew_blocks(n), EW_THREADS, 0, st>>>(real, gen, n, grad_scale, grad, accumulate, sum_out);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// Keras/TF-form Adam over one flat buffer (all variables of a net in one launch):
//   m += (g-m)(1-b1); v += (g^2-v)(1-b2); p -= lr_t * m / (sqrt(v) + eps)      
// 28 B/param of HBM traffic: p,g,m,v read; p,m,v written.
// ------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr_t, float b1, float b2, float eps,
                            float gscale) {
    const size_t n4 = n / 4;
    const float o1 = 1.f - b1, o2 = 1.f - b2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float gr = ga[j] * gscale;
            ma[j] += (gr - ma[j]) * o1;
            va[j] += (gr * gr - va[j]) * o2;
            pa[j] -= lr_t * ma[j] / (sqrtf(va[j]) + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        size_t i = n4 * 4 + threadIdx.x;
        float gr = g[i] * gscale;
        float mi = m[i] + (gr - m[i]) * o1, vi = v[i] + (gr * gr - v[i]) * o2;
        m[i] = mi; v[i] = vi;
        p[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}
int k_adam(float* p, const float* g, float* m, float* v, size_t n, float lr_t, float b1, float b2, float eps,
           float grad_scale, cudaStream_t st) {
    if (n == 0) return CG_OK;
    adam_kernel<<<ew_blocks(n / 4 + 1), EW_THREADS, 0, st>>>(p, g, m, v, n, lr_t, b1, b2, eps, grad_scale);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// explicit instantiations
// ------------------------------------------------------------------------------------------
#define INSTANTIATE(T)                                                                                          \
    template int k_convert_in<T>(const float*, T*, size_t, cudaStream_t);                                       \
    template int k_convert_out<T>(const T*, float*, size_t, cudaStream_t);                                      \
    template int k_in_stats<T>(const T*, float*, int, int, int, float, cudaStream_t, bool);                     \
    template int k_in_stats_raw<T>(const T*, float*, int, int, int, cudaStream_t, bool);                        \
    template int k_in_apply<T>(const T*, T*, const float*, const float*, const float*, int, float, int, int, int, \
                               cudaStream_t);                                                                   \
    template int k_in_bwd<T>(const T*, const T*, T*, const float*, const float*, const float*, float*, float*,  \
                             float*, int, float, int, int, int, int, cudaStream_t, int, int, bool, int);        \
    template int k_act_fwd<T>(const T*, T*, size_t, int, float, cudaStream_t);                                  \
    template int k_act_bwd<T>(const T*, const T*, T*, size_t, int, float, int, cudaStream_t);                   \
    template int k_rpad_fwd<T>(const T*, T*, int, int, int, int, int, cudaStream_t);                            \
    template int k_rpad_bwd<T>(const T*, T*, int, int, int, int, int, int, cudaStream_t);                       \
    template int k_rpad_bwd_border<T>(const T*, T*, int, int, int, int, int, cudaStream_t);                     \
    template int k_add<T>(const T*, const T*, T*, size_t, cudaStream_t);                                        \
    template int k_copy_acc<T>(const T*, T*, size_t, int, cudaStream_t);                                        \
    template int k_slice_copy<T>(const T*, int, int, T*, int, int, int, size_t, int, cudaStream_t);             \
    template int k_avgpool_fwd<T>(const T*, T*, int, int, int, int, cudaStream_t, T*, int, int);                \
    template int k_avgpool_bwd<T>(const T*, T*, int, int, int, int, int, cudaStream_t, const T*, int, int);     \
    template int k_upsample_fwd<T>(const T*, T*, int, int, int, int, cudaStream_t, T*, int, int);               \
    template int k_upsample_bwd<T>(const T*, T*, int, int, int, int, int, cudaStream_t, const T*, int, int);    \
    template int k_adv_loss<T>(const T*, size_t, float, int, float, T*, float*, float*, cudaStream_t);          \
    template int k_l1_loss<T>(const T*, const T*, size_t, float, T*, int, float*, cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(bf16)
