// Kernels of the optional config paths and of the input pipeline, all HBM-bound streaming
// kernels: the non-Adam optimizers of optimizers.py:16-21, BatchNormalization on top of the instance-norm tables
// (unet.py:27-28,57-58,71-72; resnet.py:99-100), Dropout (unet.py:33-34) with a counter-based mask, and
// normalize / resize / random_jitter / postprocess (transform/data_load.py:20-34, predict.py:20-27).
#include "kernels.h"

static const int XT = 256;
static inline int x_blocks(size_t work) {
    size_t b = (work + XT - 1) / XT;
    size_t cap = 148 * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------
// optimizers over one flat buffer (all variables of a net in one launch), Keras / adabelief_tf update forms
//   SGD       : p -= lr*g                                                      12 B/param (p, g read; p written)
//   RMSprop   : v = rho*v + (1-rho) g^2 ; p -= lr*g/(sqrt(v)+eps)              20 B/param
//   AdaBelief : m = b1*m + (1-b1) g ; v = b2*v + (1-b2)(g-m)^2 + eps ;
//               p -= rect ? lr*r_t*(m*c_m)/(sqrt(v*c_v)+eps) : lr*(m*c_m)      28 B/param
// ------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ void opt_update(float& p, float g, float& m, float& v, const OptCoef& c) {
    g *= c.gscale;
    if (KIND == CG_OPT_SGD) {
        p -= c.lr * g;
    } else if (KIND == CG_OPT_RMSPROP) {
        v = c.b2 * v + (1.f - c.b2) * g * g;
        p -= c.lr * g / (sqrtf(v) + c.eps);
    } else {
        m = c.b1 * m + (1.f - c.b1) * g;
        const float d = g - m;
        v = c.b2 * v + (1.f - c.b2) * d * d + c.eps;
        const float mc = m * c.c_m;
        p -= c.rect ? c.lr * c.r_t * mc / (sqrtf(v * c.c_v) + c.eps) : c.lr * mc;
    }
}

template <int KIND>
__global__ void __launch_bounds__(XT) opt_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                 float* __restrict__ v, size_t n, OptCoef c) {
    const size_t n4 = n / 4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = make_float4(0.f, 0.f, 0.f, 0.f), vv = mm;
        if (KIND == CG_OPT_ADABELIEF) mm = reinterpret_cast<float4*>(m)[i];
        if (KIND != CG_OPT_SGD) vv = reinterpret_cast<float4*>(v)[i];
        float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) opt_update<KIND>(pa[j], ga[j], ma[j], va[j], c);
        reinterpret_cast<float4*>(p)[i] = pp;
        if (KIND == CG_OPT_ADABELIEF) reinterpret_cast<float4*>(m)[i] = mm;
        if (KIND != CG_OPT_SGD) reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = n4 * 4 + threadIdx.x;
        float pi = p[i], mi = KIND == CG_OPT_ADABELIEF ? m[i] : 0.f, vi = KIND != CG_OPT_SGD ? v[i] : 0.f;
        opt_update<KIND>(pi, g[i], mi, vi, c);
        p[i] = pi;
        if (KIND == CG_OPT_ADABELIEF) m[i] = mi;
        if (KIND != CG_OPT_SGD) v[i] = vi;
    }
}

int k_opt_step(int kind, float* p, const float* g, float* m, float* v, size_t n, const OptCoef& c, cudaStream_t st) {
    if (n == 0) return CG_OK;
    const int blocks = x_blocks(n / 4 + 1);
    switch (kind) {
        case CG_OPT_SGD: opt_kernel<CG_OPT_SGD><<<blocks, XT, 0, st>>>(p, g, m, v, n, c); break;
        case CG_OPT_RMSPROP: opt_kernel<CG_OPT_RMSPROP><<<blocks, XT, 0, st>>>(p, g, m, v, n, c); break;
        case CG_OPT_ADABELIEF: opt_kernel<CG_OPT_ADABELIEF><<<blocks, XT, 0, st>>>(p, g, m, v, n, c); break;
        default: cg_set_error("k_opt_step: unknown optimizer kind %d", kind); return CG_ERR_INVALID;
    }
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// BatchNormalization = the instance-norm machinery with its per-(sample, channel) tables pooled over the samples of
// one Keras call ("group": several calls are concatenated along the batch by the trainer).
//   forward, training : raw (sum x, sum x^2)[n][c]  ->  (mean, rstd) of the group in every sample's slot, plus the
//                       group's (mean, unbiased variance) in bstat[g][c] for the moving averages (Keras' fused batch
//                       norm feeds the Bessel-corrected variance to the moving average, the biased one to the output)
//   forward, inference: (moving_mean, rsqrt(moving_var + eps)) in every sample's slot
//   backward          : raw (sum g, sum g*xhat)[n][c]  ->  group sum / group size in every sample's slot, so that the
//                       instance-norm apply kernels (which divide by the pixels of ONE sample) need no change
// ------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(float* __restrict__ stats, float* __restrict__ bstat, int N, int C, int group, float P,
                                   float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int G = N / group;
    if (i >= G * C) return;
    const int g = i / C, c = i - g * C;
    float s = 0.f, ss = 0.f;
    for (int n = g * group; n < (g + 1) * group; ++n) {
        s += stats[((size_t)n * C + c) * 2];
        ss += stats[((size_t)n * C + c) * 2 + 1];
    }
    const float cnt = P * (float)group;
    const float mean = s / cnt;
    const float var = fmaxf(ss / cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    for (int n = g * group; n < (g + 1) * group; ++n) {
        stats[((size_t)n * C + c) * 2] = mean;
        stats[((size_t)n * C + c) * 2 + 1] = rstd;
    }
    bstat[((size_t)g * C + c) * 2] = mean;
    bstat[((size_t)g * C + c) * 2 + 1] = cnt > 1.f ? var * (cnt / (cnt - 1.f)) : var;
}
int k_bn_finalize(float* stats, float* bstat, int N, int C, int group, int P, float eps, cudaStream_t st) {
    bn_finalize_kernel<<<cdiv((long long)(N / group) * C, 128), 128, 0, st>>>(stats, bstat, N, C, group, (float)P, eps);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

__global__ void bn_fill_kernel(float* __restrict__ stats, const float* __restrict__ mm, const float* __restrict__ mv, int N,
                               int C, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * C) return;
    const int c = i % C;
    stats[(size_t)i * 2] = mm[c];
    stats[(size_t)i * 2 + 1] = rsqrtf(mv[c] + eps);
}
int k_bn_fill(float* stats, const float* moving_mean, const float* moving_var, int N, int C, float eps, cudaStream_t st) {
    bn_fill_kernel<<<cdiv((long long)N * C, 128), 128, 0, st>>>(stats, moving_mean, moving_var, N, C, eps);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// moving = moving*momentum + batch*(1-momentum), written the way Keras does (moving -= (moving - batch)*(1-momentum))
__global__ void bn_moving_kernel(float* __restrict__ mm, float* __restrict__ mv, const float* __restrict__ bstat, int C,
                                 float one_minus_mom) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mm[c] -= (mm[c] - bstat[2 * c]) * one_minus_mom;
    mv[c] -= (mv[c] - bstat[2 * c + 1]) * one_minus_mom;
}
int k_bn_update_moving(float* moving_mean, float* moving_var, const float* bstat, int C, float momentum, cudaStream_t st) {
    bn_moving_kernel<<<cdiv(C, 128), 128, 0, st>>>(moving_mean, moving_var, bstat, C, 1.f - momentum);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

__global__ void bn_pool_sums_kernel(float* __restrict__ sums, int N, int C, int group) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int G = N / group;
    if (i >= G * C) return;
    const int g = i / C, c = i - g * C;
    float s = 0.f, ss = 0.f;
    for (int n = g * group; n < (g + 1) * group; ++n) {
        s += sums[((size_t)n * C + c) * 2];
        ss += sums[((size_t)n * C + c) * 2 + 1];
    }
    const float inv = 1.f / (float)group;
    for (int n = g * group; n < (g + 1) * group; ++n) {
        sums[((size_t)n * C + c) * 2] = s * inv;
        sums[((size_t)n * C + c) * 2 + 1] = ss * inv;
    }
}
int k_bn_pool_sums(float* sums, int N, int C, int group, cudaStream_t st) {
    bn_pool_sums_kernel<<<cdiv((long long)(N / group) * C, 128), 128, 0, st>>>(sums, N, C, group);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// Dropout (inverted, training only): y = x * keep / (1 - rate), keep = [u >= rate], u = the top 24 bits of a splitmix64
// hash of (key, element index within the Keras call).  The key mixes seed, step counter, call id and layer id on the
// host; the oracle restates the same hash in numpy (oracle/tf_ops.py dropout_mask).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned long long dropout_key(const DropKey& k, int g) {
    const unsigned long long ctr = k.ctr_dev ? *k.ctr_dev : k.ctr_host;
    unsigned long long z = splitmix64(k.seed ^ (ctr * 0xD1342543DE82EF95ull));
    const int call = g == 0 ? k.call_id[0] : g == 1 ? k.call_id[1] : g == 2 ? k.call_id[2] : k.call_id[3];   // no local-memory indexing
    z = splitmix64(z ^ ((unsigned long long)(unsigned)call << 32 | (unsigned)k.layer));
    return z;
}

// MODE 0: forward.  MODE 1: backward, o (+)= a * mask / (1 - rate).  VEC elements (one 16-byte access for bf16 x8 /
// float x4) per thread and iteration; VEC == 1 is the fallback for group sizes that are not a multiple of the vector.
template <typename T, int MODE, int VEC>
__global__ void __launch_bounds__(XT) dropout_kernel(const T* __restrict__ a, T* __restrict__ o, size_t group_elems, int groups,
                                                     float rate, float scale, DropKey key, int training, int accumulate) {
    const size_t total_v = group_elems * (size_t)groups / VEC;
    int cur_g = -1;
    unsigned long long cur_key = 0;
    for (size_t iv = blockIdx.x * (size_t)blockDim.x + threadIdx.x; iv < total_v; iv += (size_t)gridDim.x * blockDim.x) {
        const size_t i = iv * VEC;
        float v[VEC];
        load_vec<T, VEC>(a + i, v);
        if (training) {
            const int g = (int)(i / group_elems);           // VEC divides group_elems: a vector never straddles two calls
            const size_t e = i - (size_t)g * group_elems;
            if (g != cur_g) { cur_key = dropout_key(key, g); cur_g = g; }
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const unsigned long long h = splitmix64(cur_key + (e + j) * 0x9E3779B97F4A7C15ull);
                const float u = (float)(h >> 40) * (1.f / 16777216.f);
                v[j] = u >= rate ? v[j] * scale : 0.f;
            }
        }
        if (MODE == 1 && accumulate) {
            float w[VEC];
            load_vec<T, VEC>(o + i, w);
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] += w[j];
        }
        store_vec<T, VEC>(o + i, v);
    }
}
template <typename T, int MODE>
static int dropout_launch(const T* a, T* o, size_t group_elems, int groups, float rate, const DropKey& key, int training,
                          int accumulate, cudaStream_t st) {
    if (!group_elems || !groups) return CG_OK;
    constexpr int VW = VecWidth<T>::value;
    const float scale = 1.f / (1.f - rate);
    const size_t total = group_elems * (size_t)groups;
    if (group_elems % VW == 0 && !(((uintptr_t)a | (uintptr_t)o) & 15))
        dropout_kernel<T, MODE, VW><<<x_blocks(total / VW), XT, 0, st>>>(a, o, group_elems, groups, rate, scale, key, training, accumulate);
    else
        dropout_kernel<T, MODE, 1><<<x_blocks(total), XT, 0, st>>>(a, o, group_elems, groups, rate, scale, key, training, accumulate);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
template <typename T> int k_dropout_fwd(const T* x, T* y, size_t group_elems, int groups, float rate, const DropKey& key,
                                        int training, cudaStream_t st) {
    return dropout_launch<T, 0>(x, y, group_elems, groups, rate, key, training, 0, st);
}
template <typename T> int k_dropout_bwd(const T* dy, T* dx, size_t group_elems, int groups, float rate, const DropKey& key,
                                        int training, int accumulate, cudaStream_t st) {
    return dropout_launch<T, 1>(dy, dx, group_elems, groups, rate, key, training, accumulate, st);
}
template int k_dropout_fwd<float>(const float*, float*, size_t, int, float, const DropKey&, int, cudaStream_t);
template int k_dropout_fwd<bf16>(const bf16*, bf16*, size_t, int, float, const DropKey&, int, cudaStream_t);
template int k_dropout_bwd<float>(const float*, float*, size_t, int, float, const DropKey&, int, int, cudaStream_t);
template int k_dropout_bwd<bf16>(const bf16*, bf16*, size_t, int, float, const DropKey&, int, int, cudaStream_t);

__global__ void set_counter_kernel(unsigned long long* p, unsigned long long v) { *p = v; }
int k_set_counter(unsigned long long* ctr_dev, unsigned long long value, cudaStream_t st) {
    set_counter_kernel<<<1, 1, 0, st>>>(ctr_dev, value);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// input pipeline
// ------------------------------------------------------------------------------------------
// normalize (data_load.py:31-34).  A thread converts 4-byte groups (one uchar4 in, one float4 out): a warp reads 128
// and writes 512 contiguous bytes per access, four independent groups in flight per thread.
__global__ void __launch_bounds__(XT) normalize_u8_kernel(const uint8_t* __restrict__ s, float* __restrict__ d, size_t n) {
    const size_t n4 = n / 4, stride = (size_t)gridDim.x * blockDim.x;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    auto cvt = [](unsigned w) {
        float4 o;
        o.x = (float)(w & 255u) / 127.5f - 1.f;
        o.y = (float)((w >> 8) & 255u) / 127.5f - 1.f;
        o.z = (float)((w >> 16) & 255u) / 127.5f - 1.f;
        o.w = (float)(w >> 24) / 127.5f - 1.f;
        return o;
    };
    for (; i + 3 * stride < n4; i += 4 * stride) {
        unsigned w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) w[u] = reinterpret_cast<const unsigned*>(s)[i + u * stride];
#pragma unroll
        for (int u = 0; u < 4; ++u) reinterpret_cast<float4*>(d)[i + u * stride] = cvt(w[u]);
    }
    for (; i < n4; i += stride) reinterpret_cast<float4*>(d)[i] = cvt(reinterpret_cast<const unsigned*>(s)[i]);
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t t = n4 * 4 + threadIdx.x;
        d[t] = (float)s[t] / 127.5f - 1.f;
    }
}
extern "C" int cg_normalize_u8(const uint8_t* src, float* dst, size_t n, void* stream) {
    if (!src || !dst) { cg_set_error("cg_normalize_u8: null argument"); return CG_ERR_INVALID; }
    if (n == 0) return CG_OK;
    if (((uintptr_t)src & 3) || ((uintptr_t)dst & 15)) { cg_set_error("cg_normalize_u8: src must be 4-byte, dst 16-byte aligned"); return CG_ERR_INVALID; }
    normalize_u8_kernel<<<x_blocks(n / 16 + 1), XT, 0, (cudaStream_t)stream>>>(src, dst, n);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// postprocess_prediction (predict.py:26-27): np.array((p + 1) * 127.5, np.uint8) truncates toward zero
__device__ __forceinline__ unsigned post_u8(float v) {
    const float t = (v + 1.f) * 127.5f;
    return (unsigned)fminf(fmaxf(truncf(t), 0.f), 255.f);
}
__global__ void __launch_bounds__(XT) postprocess_u8_kernel(const float* __restrict__ s, uint8_t* __restrict__ d, size_t n) {
    const size_t n4 = n / 4, stride = (size_t)gridDim.x * blockDim.x;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    auto cvt = [](float4 q) { return post_u8(q.x) | (post_u8(q.y) << 8) | (post_u8(q.z) << 16) | (post_u8(q.w) << 24); };
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = reinterpret_cast<const float4*>(s)[i + u * stride];
#pragma unroll
        for (int u = 0; u < 4; ++u) reinterpret_cast<unsigned*>(d)[i + u * stride] = cvt(q[u]);
    }
    for (; i < n4; i += stride) reinterpret_cast<unsigned*>(d)[i] = cvt(reinterpret_cast<const float4*>(s)[i]);
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t t = n4 * 4 + threadIdx.x;
        d[t] = (uint8_t)post_u8(s[t]);
    }
}
extern "C" int cg_postprocess_u8(const float* src, uint8_t* dst, size_t n, void* stream) {
    if (!src || !dst) { cg_set_error("cg_postprocess_u8: null argument"); return CG_ERR_INVALID; }
    if (n == 0) return CG_OK;
    if (((uintptr_t)src & 15) || ((uintptr_t)dst & 3)) { cg_set_error("cg_postprocess_u8: src must be 16-byte, dst 4-byte aligned"); return CG_ERR_INVALID; }
    postprocess_u8_kernel<<<x_blocks(n / 16 + 1), XT, 0, (cudaStream_t)stream>>>(src, dst, n);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// tf.image.resize bilinear with half-pixel centres (TF2 default, antialias=False):
//   in = (o + 0.5) * scale - 0.5 ; lo = max(floor(in), 0) ; hi = min(ceil(in), size - 1) ; lerp = in - floor(in)
__device__ __forceinline__ void resize_coord(int o, float scale, int size, int& lo, int& hi, float& lerp) {
    // explicit roundings (no fma contraction): the source coordinate decides which pixels are blended
    const float in = __fsub_rn(__fmul_rn(__fadd_rn((float)o, 0.5f), scale), 0.5f);
    const float f = floorf(in);
    lo = max((int)f, 0);
    hi = min((int)ceilf(in), size - 1);
    lerp = __fsub_rn(in, f);
}
// One thread per output PIXEL of one output row segment: grid = (row segments, Ho, N), so the row / image indices come
// from the block index (no 64-bit divisions) and the interpolation coordinates are computed once per pixel.  With CC = 3
// (images) the block's 256 x 3 results are staged in shared memory and stored as contiguous 4-byte lanes (a thread's own 12
// bytes would make every warp store touch three lines); CC = 0 is the run-time-C fallback with direct stores.
// Output pixel (y, x) of image n samples the virtual [Hr, Wr] resized image at (oy[n] + y, ox[n] + x') with x' mirrored
// when flip[n].
template <int CC>
__global__ void __launch_bounds__(XT) resize_kernel(const float* __restrict__ src, int H, int W, int Crt, int Hr, int Wr,
                                                    float* __restrict__ dst, int Ho, int Wo, const int* __restrict__ oy,
                                                    const int* __restrict__ ox, const int* __restrict__ flip) {
    __shared__ float stage[CC ? XT * CC : 1];
    const int C = CC ? CC : Crt;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    float* out_row = dst + ((size_t)n * Ho + y) * Wo * C;
    if (x < Wo) {
        const float sy = (float)H / (float)Hr, sx = (float)W / (float)Wr;
        const int yy = y + (oy ? oy[n] : 0);
        const int xr = (flip && flip[n]) ? Wo - 1 - x : x;
        const int xx = xr + (ox ? ox[n] : 0);
        int y0, y1, x0, x1;
        float ly, lx;
        resize_coord(yy, sy, H, y0, y1, ly);
        resize_coord(xx, sx, W, x0, x1, lx);
        const float* p00 = src + (((size_t)n * H + y0) * W + x0) * C;
        const float* p01 = src + (((size_t)n * H + y0) * W + x1) * C;
        const float* p10 = src + (((size_t)n * H + y1) * W + x0) * C;
        const float* p11 = src + (((size_t)n * H + y1) * W + x1) * C;
#pragma unroll
        for (int c = 0; c < (CC ? CC : C); ++c) {
            // TF's compute_lerp order, every operation rounded on its own (bit-identical to a float32 numpy restatement)
            const float top = __fadd_rn(p00[c], __fmul_rn(__fsub_rn(p01[c], p00[c]), lx));
            const float bot = __fadd_rn(p10[c], __fmul_rn(__fsub_rn(p11[c], p10[c]), lx));
            const float v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
            if (CC) stage[threadIdx.x * CC + c] = v;
            else out_row[(size_t)x * C + c] = v;
        }
    }
    if (CC) {
        __syncthreads();
        const int e0 = blockIdx.x * blockDim.x * CC;                 // first element of this segment in the output row
        const int ne = min(XT * CC, Wo * CC - e0);
        for (int e = threadIdx.x; e < ne; e += XT) out_row[e0 + e] = stage[e];
    }
}
extern "C" int cg_resize_crop_flip(const float* src, int N, int H, int W, int C, int Hr, int Wr, float* dst, int Ho, int Wo,
                                   const int32_t* oy, const int32_t* ox, const int32_t* flip, void* stream) {
    if (!src || !dst) { cg_set_error("cg_resize_crop_flip: null argument"); return CG_ERR_INVALID; }
    if (N < 0 || H <= 0 || W <= 0 || C <= 0 || Hr <= 0 || Wr <= 0 || Ho <= 0 || Wo <= 0 || Ho > Hr || Wo > Wr || Ho > 65535 ||
        N > 65535 || (long long)W * C > (1 << 30) || (long long)Wo * C > (1 << 30)) {
        cg_set_error("cg_resize_crop_flip: bad geometry %dx%dx%dx%d -> %dx%d -> crop %dx%d", N, H, W, C, Hr, Wr, Ho, Wo);
        return CG_ERR_INVALID;
    }
    if (N == 0) return CG_OK;
    const dim3 grid(cdiv(Wo, XT), Ho, N);
    if (C == 3) resize_kernel<3><<<grid, XT, 0, (cudaStream_t)stream>>>(src, H, W, C, Hr, Wr, dst, Ho, Wo, oy, ox, flip);
    else resize_kernel<0><<<grid, XT, 0, (cudaStream_t)stream>>>(src, H, W, C, Hr, Wr, dst, Ho, Wo, oy, ox, flip);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
extern "C" int cg_resize_bilinear(const float* src, int N, int H, int W, int C, float* dst, int Ho, int Wo, void* stream) {
    return cg_resize_crop_flip(src, N, H, W, C, Ho, Wo, dst, Ho, Wo, nullptr, nullptr, nullptr, stream);
}
