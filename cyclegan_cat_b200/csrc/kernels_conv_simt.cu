// CUDA-core implicit-GEMM convolutions: forward, data-gradient (= transposed-conv forward) and
// weight-gradient, for ANY geometry the four reference builders produce (k in 1..7, stride 1/2,
// TF 'same' asymmetric padding or 'valid', Cin/Cout from 1 to 1024+).  They are
//   * the whole conv path of the fp32 check mode (exact fp32 FMA accumulation), and
//   * the path of bf16-mode layers whose shape the tcgen05 kernel does not take
//     (Cin = 3 stems, Cout = 1/3 heads, ...).
// Tile 64x64x16, 256 threads, 4x4 register micro-tile, fp32 accumulation.  Weights are the
// float32 master copies in TensorFlow layout (HWIO).
#include "kernels.h"

#define BM 64
#define BN 64
#define BK 16

template <typename T>
__device__ __forceinline__ void load4_or_zero(const T* p, bool ok, float (&v)[4]) {
    if (ok) load_vec<T, 4>(p, v);
    else { v[0] = v[1] = v[2] = v[3] = 0.f; }
}

#define MICRO_FMA()                                                                   \
    _Pragma("unroll") for (int kk = 0; kk < BK; ++kk) {                               \
        float a[4], b[4];                                                             \
        *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]); \
        *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]); \
        _Pragma("unroll") for (int i = 0; i < 4; ++i)                                 \
            _Pragma("unroll") for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]); \
    }

// ------------------------------------------------------------------------------------------
// forward:  M = N*Ho*Wo pixels, Ncol = Cout, K = k*k*Cin
// ------------------------------------------------------------------------------------------
template <typename T, bool CIN4>
__global__ void __launch_bounds__(256) conv_fwd_simt(const T* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, T* __restrict__ y, ConvGeom g,
                                                     int accumulate) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int M = g.N * g.Ho * g.Wo, K = g.k * g.k * g.Cin;

    const int a_r = tid >> 2, a_k = (tid & 3) << 2;
    const int m = m0 + a_r;
    const bool mvalid = m < M;
    int n_img = 0, ih0 = 0, iw0 = 0;
    if (mvalid) {
        n_img = m / (g.Ho * g.Wo);
        int r = m - n_img * g.Ho * g.Wo;
        int oh = r / g.Wo, ow = r - oh * g.Wo;
        ih0 = oh * g.s - g.pt;
        iw0 = ow * g.s - g.pl;
    }
    const T* xn = x + (size_t)n_img * g.Hi * g.Wi * g.Cin;
    const int b_k = tid >> 4, b_n = (tid & 15) << 2;
    const bool cout4 = (g.Cout & 3) == 0;

    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
        {   // A tile: gathered input patch elements
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            int kk = k0 + a_k;
            if (CIN4) {
                if (mvalid && kk < K) {
                    int tap = kk / g.Cin, ci = kk - tap * g.Cin;
                    int kh = tap / g.k, kw = tap - kh * g.k;
                    int ih = ih0 + kh, iw = iw0 + kw;
                    bool ok = ih >= 0 && ih < g.Hi && iw >= 0 && iw < g.Wi;
                    load4_or_zero<T>(xn + ((size_t)ih * g.Wi + iw) * g.Cin + ci, ok, v);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int kj = kk + j;
                    if (mvalid && kj < K) {
                        int tap = kj / g.Cin, ci = kj - tap * g.Cin;
                        int kh = tap / g.k, kw = tap - kh * g.k;
                        int ih = ih0 + kh, iw = iw0 + kw;
                        if (ih >= 0 && ih < g.Hi && iw >= 0 && iw < g.Wi)
                            v[j] = ldf(xn + ((size_t)ih * g.Wi + iw) * g.Cin + ci);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) As[a_k + j][a_r] = v[j];
        }
        {   // B tile: w[kk][co], co contiguous
            int kk = k0 + b_k, co = n0 + b_n;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (kk < K) {
                if (cout4 && co + 3 < g.Cout) *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(w + (size_t)kk * g.Cout + co);
                else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (co + j < g.Cout) v[j] = w[(size_t)kk * g.Cout + co + j];
                }
            }
            *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = *reinterpret_cast<float4*>(v);
        }
        __syncthreads();
        MICRO_FMA();
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int mm = m0 + ty * 4 + i;
        if (mm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = n0 + tx * 4 + j;
            if (co < g.Cout) {
                T* o = y + (size_t)mm * g.Cout + co;
                float v = acc[i][j] + (bias ? bias[co] : 0.f);
                stf(o, accumulate ? ldf(o) + v : v);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------
// skinny convolutions (Cout <= 4: the tanh head c7s1-3 of resnet.py:82, 1x1 heads): a 64-wide GEMM tile would
// waste 95 % of its columns.  forward: one thread = two horizontally adjacent output pixels x all Cout channels,
// weights broadcast from shared memory as float4 rows; inputs read as 16-byte channel vectors.
// ------------------------------------------------------------------------------------------
template <typename T, bool VEC8>
__global__ void __launch_bounds__(128) conv_fwd_skinny(const T* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, T* __restrict__ y, ConvGeom g,
                                                       int accumulate) {
    extern __shared__ float4 wsm[];      // [K] rows of (w0, w1, w2, w3), zero padded
    const int K = g.k * g.k * g.Cin;
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        float* pv = &v.x;
        for (int c = 0; c < g.Cout; ++c) pv[c] = w[(size_t)i * g.Cout + c];
        wsm[i] = v;
    }
    __syncthreads();
    const int wpairs = (g.Wo + 1) / 2;
    const long long total = (long long)g.N * g.Ho * wpairs;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int wp, oh, n;          // 32-bit index arithmetic whenever it fits: a 64-bit divide is ~100 emulated instructions, more
        if (total <= 0x7fffffffLL) {       // than the 16 x 8 FMAs of a 1x1 head pixel pair
            const int i32 = (int)idx, r = i32 / wpairs;
            wp = i32 - r * wpairs; n = r / g.Ho; oh = r - n * g.Ho;
        } else {
            wp = (int)(idx % wpairs);
            const long long r = idx / wpairs;
            oh = (int)(r % g.Ho); n = (int)(r / g.Ho);
        }
        const int ow0 = wp * 2;
        const bool second = ow0 + 1 < g.Wo;
        float acc[2][4] = {};
        const T* xn = x + (size_t)n * g.Hi * g.Wi * g.Cin;
        for (int kh = 0; kh < g.k; ++kh) {
            const int ih = oh * g.s + kh - g.pt;
            if (ih < 0 || ih >= g.Hi) continue;
            for (int kw = 0; kw < g.k; ++kw) {
                const int iw0 = ow0 * g.s + kw - g.pl, iw1 = iw0 + g.s;
                const bool ok0 = iw0 >= 0 && iw0 < g.Wi, ok1 = second && iw1 >= 0 && iw1 < g.Wi;
                if (!ok0 && !ok1) continue;
                const T* p0 = xn + ((size_t)ih * g.Wi + iw0) * g.Cin;
                const T* p1 = xn + ((size_t)ih * g.Wi + iw1) * g.Cin;
                const float4* wr = wsm + (kh * g.k + kw) * g.Cin;
                if (VEC8) {
                    for (int ci = 0; ci < g.Cin; ci += 8) {
                        float a0[8], a1[8];
                        if (ok0) load_vec<T, VecWidth<T>::value == 8 ? 8 : 4>(p0 + ci, *reinterpret_cast<float(*)[VecWidth<T>::value == 8 ? 8 : 4]>(a0));
                        if (ok1) load_vec<T, VecWidth<T>::value == 8 ? 8 : 4>(p1 + ci, *reinterpret_cast<float(*)[VecWidth<T>::value == 8 ? 8 : 4]>(a1));
                        if (VecWidth<T>::value == 4) {
                            if (ok0) load_vec<T, 4>(p0 + ci + 4, *reinterpret_cast<float(*)[4]>(a0 + 4));
                            if (ok1) load_vec<T, 4>(p1 + ci + 4, *reinterpret_cast<float(*)[4]>(a1 + 4));
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 wv = wr[ci + j];
                            if (ok0) { acc[0][0] = fmaf(a0[j], wv.x, acc[0][0]); acc[0][1] = fmaf(a0[j], wv.y, acc[0][1]);
                                       acc[0][2] = fmaf(a0[j], wv.z, acc[0][2]); acc[0][3] = fmaf(a0[j], wv.w, acc[0][3]); }
                            if (ok1) { acc[1][0] = fmaf(a1[j], wv.x, acc[1][0]); acc[1][1] = fmaf(a1[j], wv.y, acc[1][1]);
                                       acc[1][2] = fmaf(a1[j], wv.z, acc[1][2]); acc[1][3] = fmaf(a1[j], wv.w, acc[1][3]); }
                        }
                    }
                } else {
                    for (int ci = 0; ci < g.Cin; ++ci) {
                        const float4 wv = wr[ci];
                        if (ok0) { const float a = ldf(p0 + ci);
                                   acc[0][0] = fmaf(a, wv.x, acc[0][0]); acc[0][1] = fmaf(a, wv.y, acc[0][1]);
                                   acc[0][2] = fmaf(a, wv.z, acc[0][2]); acc[0][3] = fmaf(a, wv.w, acc[0][3]); }
                        if (ok1) { const float a = ldf(p1 + ci);
                                   acc[1][0] = fmaf(a, wv.x, acc[1][0]); acc[1][1] = fmaf(a, wv.y, acc[1][1]);
                                   acc[1][2] = fmaf(a, wv.z, acc[1][2]); acc[1][3] = fmaf(a, wv.w, acc[1][3]); }
                    }
                }
            }
        }
        for (int px = 0; px < (second ? 2 : 1); ++px) {
            T* o = y + (((size_t)n * g.Ho + oh) * g.Wo + ow0 + px) * g.Cout;
            for (int c = 0; c < g.Cout; ++c) {
                float v = acc[px][c] + (bias ? bias[c] : 0.f);
                stf(o + c, accumulate ? ldf(o + c) + v : v);
            }
        }
    }
}

// stride-1 skinny forward with the input tile (output tile + halo) staged ONCE in shared memory (the 7x7 head reads
// each input pixel 49 times): block = 32 x (8*PX) output pixels, thread = PX vertically stacked pixels x Cout
// channels, 16-byte channel vectors XOR-swizzled by pixel so that a warp reading 32 neighbouring pixels is
// bank-conflict free, weights broadcast as float4 rows.
template <typename T, int PX>
__global__ void __launch_bounds__(256) conv_fwd_skinny_tiled(const T* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, T* __restrict__ y,
                                                             ConvGeom g, int accumulate, int tiles_w, int tiles_h) {
    constexpr int VEC = VecWidth<T>::value;
    constexpr int TW = 32, TH = 8 * PX;
    extern __shared__ float4 smem4[];
    const int K = g.k * g.k * g.Cin;
    float4* wsm = smem4;                                  // [K] weight rows
    uint4* xs = reinterpret_cast<uint4*>(smem4 + K);      // [IH*IW][nv] 16-byte vectors
    const int nv = g.Cin / VEC, IW = TW + g.k - 1, IH = TH + g.k - 1;
    for (int i = threadIdx.x; i < K; i += 256) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        float* pv = &v.x;
        for (int c = 0; c < g.Cout; ++c) pv[c] = w[(size_t)i * g.Cout + c];
        wsm[i] = v;
    }
    int t = blockIdx.x;
    const int tw = t % tiles_w; t /= tiles_w;
    const int th = t % tiles_h;
    const int n = t / tiles_h;
    const int ow0 = tw * TW, oh0 = th * TH;
    const int iw0 = ow0 - g.pl, ih0 = oh0 - g.pt;
    const T* xn = x + (size_t)n * g.Hi * g.Wi * g.Cin;
    for (int i = threadIdx.x; i < IH * IW * nv; i += 256) {
        const int cv = i % nv, p = i / nv;
        const int ph = p / IW, pw = p - ph * IW;
        const int ih = ih0 + ph, iw = iw0 + pw;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (ih >= 0 && ih < g.Hi && iw >= 0 && iw < g.Wi)
            v = *reinterpret_cast<const uint4*>(xn + ((size_t)ih * g.Wi + iw) * g.Cin + cv * VEC);
        xs[p * nv + (cv ^ (p & (nv - 1)))] = v;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float acc[PX][4] = {};
    for (int kh = 0; kh < g.k; ++kh)
        for (int kw = 0; kw < g.k; ++kw) {
            const float4* wr = wsm + (kh * g.k + kw) * g.Cin;
            int pbase[PX];
#pragma unroll
            for (int q = 0; q < PX; ++q) pbase[q] = (ty + 8 * q + kh) * IW + tx + kw;
            for (int cv = 0; cv < nv; ++cv) {
                float a[PX][VEC];
#pragma unroll
                for (int q = 0; q < PX; ++q) {
                    const uint4 raw = xs[pbase[q] * nv + (cv ^ (pbase[q] & (nv - 1)))];
                    load_vec<T, VEC>(reinterpret_cast<const T*>(&raw), a[q]);
                }
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const float4 wv = wr[cv * VEC + j];
#pragma unroll
                    for (int q = 0; q < PX; ++q) {
                        acc[q][0] = fmaf(a[q][j], wv.x, acc[q][0]); acc[q][1] = fmaf(a[q][j], wv.y, acc[q][1]);
                        acc[q][2] = fmaf(a[q][j], wv.z, acc[q][2]); acc[q][3] = fmaf(a[q][j], wv.w, acc[q][3]);
                    }
                }
            }
        }
#pragma unroll
    for (int q = 0; q < PX; ++q) {
        const int oh = oh0 + ty + 8 * q, ow = ow0 + tx;
        if (oh < g.Ho && ow < g.Wo) {
            T* o = y + (((size_t)n * g.Ho + oh) * g.Wo + ow) * g.Cout;
            for (int c = 0; c < g.Cout; ++c) {
                float v = acc[q][c] + (bias ? bias[c] : 0.f);
                stf(o + c, accumulate ? ldf(o + c) + v : v);
            }
        }
    }
}

// weight gradient for Cout <= 4: thread = (kw, pair of input channels), blockIdx.y = kh, blockIdx.x = (image, band of
// output rows); every thread walks its band with 4 independent pixels in flight, accumulating 2 x Cout sums, then one
// atomicAdd each (few blocks per address: the band is sized so that the grid is ~2 waves).
template <typename T>
__global__ void __launch_bounds__(1024) conv_wgrad_skinny(const T* __restrict__ x, const T* __restrict__ dy,
                                                          float* __restrict__ dw, ConvGeom g, int rows_per_block,
                                                          int bands) {
    const int half = g.Cin / 2;
    const int kw = threadIdx.x / half, ci = (threadIdx.x % half) * 2, kh = blockIdx.y;
    if (kw >= g.k) return;
    const int n = blockIdx.x / bands, band = blockIdx.x % bands;
    const int oh0 = band * rows_per_block;
    const int oh1 = min(g.Ho, oh0 + rows_per_block);
    float acc[2][4] = {};
    // valid output columns for this kw: 0 <= ow*s + kw - pl < Wi
    int ow_lo = 0, ow_hi = g.Wo;
    while (ow_lo < g.Wo && ow_lo * g.s + kw - g.pl < 0) ++ow_lo;
    while (ow_hi > ow_lo && (ow_hi - 1) * g.s + kw - g.pl >= g.Wi) --ow_hi;
    for (int oh = oh0; oh < oh1; ++oh) {
        const int ih = oh * g.s + kh - g.pt;
        if (ih < 0 || ih >= g.Hi) continue;
        const T* xrow = x + (((size_t)n * g.Hi + ih) * g.Wi + (kw - g.pl)) * g.Cin + ci;
        const T* drow = dy + (((size_t)n * g.Ho + oh) * g.Wo) * g.Cout;
        int ow = ow_lo;
        for (; ow + 4 <= ow_hi; ow += 4) {
            float a0[4], a1[4], d[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const T* xp = xrow + (size_t)(ow + u) * g.s * g.Cin;
                a0[u] = ldf(xp); a1[u] = ldf(xp + 1);
#pragma unroll
                for (int c = 0; c < 4; ++c) d[u][c] = c < g.Cout ? ldf(drow + (size_t)(ow + u) * g.Cout + c) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int c = 0; c < 4; ++c) { acc[0][c] = fmaf(a0[u], d[u][c], acc[0][c]); acc[1][c] = fmaf(a1[u], d[u][c], acc[1][c]); }
        }
        for (; ow < ow_hi; ++ow) {
            const T* xp = xrow + (size_t)ow * g.s * g.Cin;
            const float a0 = ldf(xp), a1 = ldf(xp + 1);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < g.Cout) {
                    const float dd = ldf(drow + (size_t)ow * g.Cout + c);
                    acc[0][c] = fmaf(a0, dd, acc[0][c]);
                    acc[1][c] = fmaf(a1, dd, acc[1][c]);
                }
        }
    }
    float* o = dw + ((size_t)(kh * g.k + kw) * g.Cin + ci) * g.Cout;
    for (int c = 0; c < g.Cout; ++c) {
        atomicAdd(o + c, acc[0][c]);
        atomicAdd(o + g.Cout + c, acc[1][c]);
    }
}

// vectorised variant: thread = (kw, 16-byte vector of input channels); 4 pixels in flight per thread
template <typename T>
__global__ void __launch_bounds__(1024) conv_wgrad_skinny_vec(const T* __restrict__ x, const T* __restrict__ dy,
                                                              float* __restrict__ dw, ConvGeom g, int rows_per_block,
                                                              int bands) {
    constexpr int VEC = VecWidth<T>::value;
    const int nv = g.Cin / VEC;
    const int kw = threadIdx.x / nv, ci = (threadIdx.x % nv) * VEC, kh = blockIdx.y;
    if (kw >= g.k) return;
    const int n = blockIdx.x / bands, band = blockIdx.x % bands;
    const int oh0 = band * rows_per_block;
    const int oh1 = min(g.Ho, oh0 + rows_per_block);
    float acc[VEC][4] = {};
    int ow_lo = 0, ow_hi = g.Wo;
    while (ow_lo < g.Wo && ow_lo * g.s + kw - g.pl < 0) ++ow_lo;
    while (ow_hi > ow_lo && (ow_hi - 1) * g.s + kw - g.pl >= g.Wi) --ow_hi;
    for (int oh = oh0; oh < oh1; ++oh) {
        const int ih = oh * g.s + kh - g.pt;
        if (ih < 0 || ih >= g.Hi) continue;
        const T* xrow = x + (((size_t)n * g.Hi + ih) * g.Wi + (kw - g.pl)) * g.Cin + ci;
        const T* drow = dy + (((size_t)n * g.Ho + oh) * g.Wo) * g.Cout;
        int ow = ow_lo;
        for (; ow + 4 <= ow_hi; ow += 4) {
            float a[4][VEC], d[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                load_vec<T, VEC>(xrow + (size_t)(ow + u) * g.s * g.Cin, a[u]);
#pragma unroll
                for (int c = 0; c < 4; ++c) d[u][c] = c < g.Cout ? ldf(drow + (size_t)(ow + u) * g.Cout + c) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < VEC; ++j)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[j][c] = fmaf(a[u][j], d[u][c], acc[j][c]);
        }
        for (; ow < ow_hi; ++ow) {
            float a[VEC];
            load_vec<T, VEC>(xrow + (size_t)ow * g.s * g.Cin, a);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < g.Cout) {
                    const float dd = ldf(drow + (size_t)ow * g.Cout + c);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) acc[j][c] = fmaf(a[j], dd, acc[j][c]);
                }
        }
    }
    float* o = dw + ((size_t)(kh * g.k + kw) * g.Cin + ci) * g.Cout;
#pragma unroll
    for (int j = 0; j < VEC; ++j)
        for (int c = 0; c < g.Cout; ++c) atomicAdd(o + (size_t)j * g.Cout + c, acc[j][c]);
}

// data gradient for Cin <= 4 (image-side layers: the 7x7 stem, the first discriminator conv): one thread = one input
// pixel x all Cin channels; weights [tap][co] -> float4 over ci in shared memory; dY read as 16-byte channel vectors.
template <typename T, bool VEC8>
__global__ void __launch_bounds__(128) conv_dgrad_skinny(const T* __restrict__ dy, const float* __restrict__ w,
                                                         const float* __restrict__ bias, T* __restrict__ dx, ConvGeom g,
                                                         int accumulate) {
    extern __shared__ float4 wsm[];      // [tap][co] -> (w[ci=0], w[1], w[2], w[3])
    const int KT = g.k * g.k * g.Cout;
    for (int i = threadIdx.x; i < KT; i += blockDim.x) {
        const int tap = i / g.Cout, co = i - tap * g.Cout;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        float* pv = &v.x;
        for (int c = 0; c < g.Cin; ++c) pv[c] = w[((size_t)tap * g.Cin + c) * g.Cout + co];
        wsm[i] = v;
    }
    __syncthreads();
    const long long total = (long long)g.N * g.Hi * g.Wi;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int iw, ih, n;          // 32-bit index arithmetic whenever it fits (see conv_fwd_skinny)
        if (total <= 0x7fffffffLL) {
            const int i32 = (int)idx, r = i32 / g.Wi;
            iw = i32 - r * g.Wi; n = r / g.Hi; ih = r - n * g.Hi;
        } else {
            iw = (int)(idx % g.Wi);
            const long long r = idx / g.Wi;
            ih = (int)(r % g.Hi); n = (int)(r / g.Hi);
        }
        const bool s1 = g.s == 1;          // stride 1: no divisibility test / divide per tap
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const T* dyn = dy + (size_t)n * g.Ho * g.Wo * g.Cout;
        for (int kh = 0; kh < g.k; ++kh) {
            const int th = ih + g.pt - kh;
            if (th < 0 || (!s1 && th % g.s)) continue;
            const int oh = s1 ? th : th / g.s;
            if (oh >= g.Ho) continue;
            for (int kw = 0; kw < g.k; ++kw) {
                const int tw = iw + g.pl - kw;
                if (tw < 0 || (!s1 && tw % g.s)) continue;
                const int ow = s1 ? tw : tw / g.s;
                if (ow >= g.Wo) continue;
                const T* p = dyn + ((size_t)oh * g.Wo + ow) * g.Cout;
                const float4* wr = wsm + (kh * g.k + kw) * g.Cout;
                if (VEC8) {
                    for (int co = 0; co < g.Cout; co += 8) {
                        float a[8];
                        load_vec<T, VecWidth<T>::value == 8 ? 8 : 4>(p + co, *reinterpret_cast<float(*)[VecWidth<T>::value == 8 ? 8 : 4]>(a));
                        if (VecWidth<T>::value == 4) load_vec<T, 4>(p + co + 4, *reinterpret_cast<float(*)[4]>(a + 4));
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 wv = wr[co + j];
                            acc[0] = fmaf(a[j], wv.x, acc[0]); acc[1] = fmaf(a[j], wv.y, acc[1]);
                            acc[2] = fmaf(a[j], wv.z, acc[2]); acc[3] = fmaf(a[j], wv.w, acc[3]);
                        }
                    }
                } else {
                    for (int co = 0; co < g.Cout; ++co) {
                        const float a = ldf(p + co);
                        const float4 wv = wr[co];
                        acc[0] = fmaf(a, wv.x, acc[0]); acc[1] = fmaf(a, wv.y, acc[1]);
                        acc[2] = fmaf(a, wv.z, acc[2]); acc[3] = fmaf(a, wv.w, acc[3]);
                    }
                }
            }
        }
        T* o = dx + (size_t)idx * g.Cin;
        for (int c = 0; c < g.Cin; ++c) {
            float v = acc[c] + (bias ? bias[c] : 0.f);
            stf(o + c, accumulate ? ldf(o + c) + v : v);
        }
    }
}

template <typename T> int k_conv_fwd(const T* x, const float* w, const float* bias, T* y, ConvGeom g, int accumulate,
                                     cudaStream_t st) {
    long long M = (long long)g.N * g.Ho * g.Wo;
    const size_t wbytes = (size_t)g.k * g.k * g.Cin * sizeof(float4);
    {   // stride-1 skinny conv with enough reuse: stage the input tile in shared memory
        constexpr int VW = VecWidth<T>::value;
        constexpr int PX = sizeof(T) == 2 ? 2 : 1;
        const int nv = g.Cin % VW == 0 ? g.Cin / VW : 0;
        const size_t tile_bytes = (size_t)(8 * PX + g.k - 1) * (32 + g.k - 1) * g.Cin * sizeof(T);
        if (g.Cout <= 4 && g.s == 1 && g.k >= 3 && nv >= 1 && (nv & (nv - 1)) == 0 && wbytes + tile_bytes <= 200 * 1024) {
            static std::atomic<unsigned long long> attr_t{0};
            if (cg_first_on_device(attr_t)) {
                CG_CUDA(cudaFuncSetAttribute(conv_fwd_skinny_tiled<T, PX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            }
            const int tiles_w = cdiv(g.Wo, 32), tiles_h = cdiv(g.Ho, 8 * PX);
            conv_fwd_skinny_tiled<T, PX><<<g.N * tiles_w * tiles_h, 256, wbytes + tile_bytes, st>>>(x, w, bias, y, g, accumulate,
                                                                                                 tiles_w, tiles_h);
            CG_LAUNCH_CHECK();
            return CG_OK;
        }
    }
    if (g.Cout <= 4 && wbytes <= 96 * 1024) {
        static std::atomic<unsigned long long> attr_done{0};
        if (cg_first_on_device(attr_done)) {
            CG_CUDA(cudaFuncSetAttribute(conv_fwd_skinny<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            CG_CUDA(cudaFuncSetAttribute(conv_fwd_skinny<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        }
        long long work = (long long)g.N * g.Ho * ((g.Wo + 1) / 2);
        int blocks = (int)((work + 127) / 128 < 148 * 16 ? (work + 127) / 128 : 148 * 16);
        if (g.Cin % 8 == 0) conv_fwd_skinny<T, true><<<blocks, 128, wbytes, st>>>(x, w, bias, y, g, accumulate);
        else conv_fwd_skinny<T, false><<<blocks, 128, wbytes, st>>>(x, w, bias, y, g, accumulate);
        CG_LAUNCH_CHECK();
        return CG_OK;
    }
    dim3 grid(cdiv(M, BM), cdiv(g.Cout, BN));
    if (g.Cin % 4 == 0) conv_fwd_simt<T, true><<<grid, 256, 0, st>>>(x, w, bias, y, g, accumulate);
    else conv_fwd_simt<T, false><<<grid, 256, 0, st>>>(x, w, bias, y, g, accumulate);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// data gradient / transposed-conv forward (gather form, one grid.z slice per stride-parity class
// so that only the taps that really hit a pixel are multiplied):
//   dx[n,ih,iw,ci] = sum_{kh,kw,co} dy[n,oh,ow,co] * w[kh,kw,ci,co],  oh*s + kh - pt = ih
//   M = pixels of the class, Ncol = Cin, K = (taps of the class) * Cout
// ------------------------------------------------------------------------------------------
template <typename T, bool COUT4>
__global__ void __launch_bounds__(256) conv_dgrad_simt(const T* __restrict__ dy, const float* __restrict__ w,
                                                       const float* __restrict__ bias, T* __restrict__ dx,
                                                       ConvGeom g, int accumulate) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ph = blockIdx.z / g.s, pw = blockIdx.z % g.s;
    const int Hc = (g.Hi - ph + g.s - 1) / g.s, Wc = (g.Wi - pw + g.s - 1) / g.s;
    const int M = g.N * Hc * Wc;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (m0 >= M) return;
    const int kh_first = (ph + g.pt) % g.s, kw_first = (pw + g.pl) % g.s;
    const int nkh = kh_first < g.k ? (g.k - kh_first + g.s - 1) / g.s : 0;
    const int nkw = kw_first < g.k ? (g.k - kw_first + g.s - 1) / g.s : 0;
    const int qh = (ph + g.pt - kh_first) / g.s, qw = (pw + g.pl - kw_first) / g.s;
    const int K = nkh * nkw * g.Cout;

    const int a_r = tid >> 2, a_k = (tid & 3) << 2;
    const int m = m0 + a_r;
    const bool mvalid = m < M;
    int n_img = 0, hc = 0, wc = 0;
    if (mvalid) {
        n_img = m / (Hc * Wc);
        int r = m - n_img * Hc * Wc;
        hc = r / Wc;
        wc = r - hc * Wc;
    }
    const T* dyn = dy + (size_t)n_img * g.Ho * g.Wo * g.Cout;
    const int b_c = tid >> 2, b_k = (tid & 3) << 2;     // B: column ci = n0 + b_c, 4 consecutive k

    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
        {
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            int kk = k0 + a_k;
            if (COUT4) {
                if (mvalid && kk < K) {
                    int ta = kk / g.Cout, co = kk - ta * g.Cout;
                    int a = ta / nkw, b = ta - a * nkw;
                    int oh = hc + qh - a, ow = wc + qw - b;
                    bool ok = oh >= 0 && oh < g.Ho && ow >= 0 && ow < g.Wo;
                    load4_or_zero<T>(dyn + ((size_t)oh * g.Wo + ow) * g.Cout + co, ok, v);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int kj = kk + j;
                    if (mvalid && kj < K) {
                        int ta = kj / g.Cout, co = kj - ta * g.Cout;
                        int a = ta / nkw, b = ta - a * nkw;
                        int oh = hc + qh - a, ow = wc + qw - b;
                        if (oh >= 0 && oh < g.Ho && ow >= 0 && ow < g.Wo)
                            v[j] = ldf(dyn + ((size_t)oh * g.Wo + ow) * g.Cout + co);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) As[a_k + j][a_r] = v[j];
        }
        {
            int ci = n0 + b_c;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (ci < g.Cin) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int kj = k0 + b_k + j;
                    if (kj < K) {
                        int ta = kj / g.Cout, co = kj - ta * g.Cout;
                        int a = ta / nkw, b = ta - a * nkw;
                        int kh = kh_first + a * g.s, kw = kw_first + b * g.s;
                        v[j] = w[(((size_t)kh * g.k + kw) * g.Cin + ci) * g.Cout + co];
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) Bs[b_k + j][b_c] = v[j];
        }
        __syncthreads();
        MICRO_FMA();
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int mm = m0 + ty * 4 + i;
        if (mm >= M) continue;
        int ni = mm / (Hc * Wc);
        int r = mm - ni * Hc * Wc;
        int h2 = r / Wc, w2 = r - h2 * Wc;
        size_t pix = ((size_t)ni * g.Hi + (h2 * g.s + ph)) * g.Wi + (w2 * g.s + pw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int ci = n0 + tx * 4 + j;
            if (ci < g.Cin) {
                T* o = dx + pix * g.Cin + ci;
                float v = acc[i][j] + (bias ? bias[ci] : 0.f);
                stf(o, accumulate ? ldf(o) + v : v);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// 1x1 convolutions with <= 4 output channels (the U-Net heads, unet.py:121: Conv2D(output_channels, 1)): both gradients are
// pure streaming passes over the input-side tensor (read x once / write dx once), so one thread owns one 16-byte channel
// vector of a pixel; the generic kernels above run these at ~0.2 TB/s (4 active lanes per warp / a K = 3 GEMM).
//   wgrad: dw[ci][co] += sum_p x[p][ci] * dy[p][co]        dgrad: dx[p][ci] (+)= sum_co dy[p][co] * w[ci][co]
// ------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) conv1x1_thin_wgrad(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw,
                                                          size_t npix, int Cin, int Cout) {
    const int nv = Cin / VEC, cv = threadIdx.x % nv, prow = threadIdx.x / nv, ppb = 256 / nv;
    float acc[VEC][4] = {};
    for (size_t p = blockIdx.x * (size_t)ppb + prow; p < npix; p += (size_t)gridDim.x * ppb) {
        float a[VEC], d[4];
        load_vec<T, VEC>(x + p * Cin + (size_t)cv * VEC, a);
#pragma unroll
        for (int c = 0; c < 4; ++c) d[c] = c < Cout ? ldf(dy + p * Cout + c) : 0.f;
#pragma unroll
        for (int j = 0; j < VEC; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[j][c] = fmaf(a[j], d[c], acc[j][c]);
    }
    // block reduction over the pixel rows of the same channel vector, then one atomic per (ci, co)
    __shared__ float red[256][4 * VEC + 1];
#pragma unroll
    for (int j = 0; j < VEC; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) red[threadIdx.x][j * 4 + c] = acc[j][c];
    __syncthreads();
    for (int o = threadIdx.x; o < Cin * Cout; o += 256) {
        const int ci = o / Cout, co = o - ci * Cout, v = ci / VEC, j = ci - v * VEC;
        float sum = 0.f;
        for (int r = 0; r < ppb; ++r) sum += red[r * nv + v][j * 4 + co];
        atomicAdd(dw + o, sum);
    }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) conv1x1_thin_dgrad(const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx,
                                                          size_t npix, int Cin, int Cout, int accumulate) {
    const int nv = Cin / VEC;
    const size_t total = npix * nv;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / nv;
        const int ci0 = (int)(i - p * nv) * VEC;
        float d[4], o[VEC];
#pragma unroll
        for (int c = 0; c < 4; ++c) d[c] = c < Cout ? ldf(dy + p * Cout + c) : 0.f;
        if (accumulate) load_vec<T, VEC>(dx + p * Cin + ci0, o);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float v = 0.f;
            for (int c = 0; c < Cout; ++c) v = fmaf(d[c], __ldg(w + (size_t)(ci0 + j) * Cout + c), v);
            o[j] = accumulate ? o[j] + v : v;
        }
        store_vec<T, VEC>(dx + p * Cin + ci0, o);
    }
}

template <typename T> int k_conv_dgrad(const T* dy, const float* w, const float* bias, T* dx, ConvGeom g,
                                       int accumulate, cudaStream_t st) {
    constexpr int VW1 = VecWidth<T>::value;
    if (g.k == 1 && g.s == 1 && g.Cout <= 4 && g.Cin % VW1 == 0 && !bias) {       // 1x1 thin-output head: streaming pass
        const size_t npix = (size_t)g.N * g.Hi * g.Wi, total = npix * (g.Cin / VW1);
        const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        conv1x1_thin_dgrad<T, VW1><<<blocks, 256, 0, st>>>(dy, w, dx, npix, g.Cin, g.Cout, accumulate);
        CG_LAUNCH_CHECK();
        return CG_OK;
    }
    const size_t wbytes = (size_t)g.k * g.k * g.Cout * sizeof(float4);
    if (g.Cin <= 4 && wbytes <= 96 * 1024) {
        static std::atomic<unsigned long long> attr_done{0};
        if (cg_first_on_device(attr_done)) {
            CG_CUDA(cudaFuncSetAttribute(conv_dgrad_skinny<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            CG_CUDA(cudaFuncSetAttribute(conv_dgrad_skinny<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        }
        long long work = (long long)g.N * g.Hi * g.Wi;
        int blocks = (int)((work + 127) / 128 < 148 * 16 ? (work + 127) / 128 : 148 * 16);
        if (g.Cout % 8 == 0) conv_dgrad_skinny<T, true><<<blocks, 128, wbytes, st>>>(dy, w, bias, dx, g, accumulate);
        else conv_dgrad_skinny<T, false><<<blocks, 128, wbytes, st>>>(dy, w, bias, dx, g, accumulate);
        CG_LAUNCH_CHECK();
        return CG_OK;
    }
    int Hc = cdiv(g.Hi, g.s), Wc = cdiv(g.Wi, g.s);
    long long M = (long long)g.N * Hc * Wc;
    dim3 grid(cdiv(M, BM), cdiv(g.Cin, BN), g.s * g.s);
    if (g.Cout % 4 == 0) conv_dgrad_simt<T, true><<<grid, 256, 0, st>>>(dy, w, bias, dx, g, accumulate);
    else conv_dgrad_simt<T, false><<<grid, 256, 0, st>>>(dy, w, bias, dx, g, accumulate);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// weight gradient: dw[(kh,kw,ci), co] += sum_pixels x[n, oh*s+kh-pt, ow*s+kw-pl, ci] * dy[n,oh,ow,co]
//   M = k*k*Cin, Ncol = Cout, K = N*Ho*Wo pixels, split over grid.z, fp32 atomics into dw
// ------------------------------------------------------------------------------------------
template <typename T, bool CIN4>
__global__ void __launch_bounds__(256) conv_wgrad_simt(const T* __restrict__ x, const T* __restrict__ dy,
                                                       float* __restrict__ dw, ConvGeom g, int pix_per_split) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int Mrows = g.k * g.k * g.Cin;
    const int P = g.N * g.Ho * g.Wo;
    const int p_begin = blockIdx.z * pix_per_split;
    const int p_end = min(P, p_begin + pix_per_split);

    // A: 4 consecutive rows (ci) for one pixel;  rows a_m..a_m+3, pixel slot a_p
    const int a_m = (tid & 15) << 2, a_p = tid >> 4;
    int a_tap[4], a_ci[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int r = m0 + a_m + j;
        if (r < Mrows) { a_tap[j] = r / g.Cin; a_ci[j] = r - a_tap[j] * g.Cin; }
        else { a_tap[j] = -1; a_ci[j] = 0; }
    }
    const int b_n = (tid & 15) << 2, b_p = tid >> 4;
    const bool cout4 = (g.Cout & 3) == 0;

    float acc[4][4] = {};
    for (int p0 = p_begin; p0 < p_end; p0 += BK) {
        {
            int p = p0 + a_p;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (p < p_end) {
                int n_img = p / (g.Ho * g.Wo);
                int r = p - n_img * g.Ho * g.Wo;
                int oh = r / g.Wo, ow = r - oh * g.Wo;
                const T* xn = x + (size_t)n_img * g.Hi * g.Wi * g.Cin;
                if (CIN4) {
                    if (a_tap[0] >= 0) {
                        int kh = a_tap[0] / g.k, kw = a_tap[0] - kh * g.k;
                        int ih = oh * g.s + kh - g.pt, iw = ow * g.s + kw - g.pl;
                        bool ok = ih >= 0 && ih < g.Hi && iw >= 0 && iw < g.Wi;
                        load4_or_zero<T>(xn + ((size_t)ih * g.Wi + iw) * g.Cin + a_ci[0], ok, v);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (a_tap[j] >= 0) {
                            int kh = a_tap[j] / g.k, kw = a_tap[j] - kh * g.k;
                            int ih = oh * g.s + kh - g.pt, iw = ow * g.s + kw - g.pl;
                            if (ih >= 0 && ih < g.Hi && iw >= 0 && iw < g.Wi)
                                v[j] = ldf(xn + ((size_t)ih * g.Wi + iw) * g.Cin + a_ci[j]);
                        }
                    }
                }
            }
            *reinterpret_cast<float4*>(&As[a_p][a_m]) = *reinterpret_cast<float4*>(v);
        }
        {
            int p = p0 + b_p, co = n0 + b_n;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (p < p_end) {
                if (cout4 && co + 3 < g.Cout) load_vec<T, 4>(dy + (size_t)p * g.Cout + co, v);
                else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (co + j < g.Cout) v[j] = ldf(dy + (size_t)p * g.Cout + co + j);
                }
            }
            *reinterpret_cast<float4*>(&Bs[b_p][b_n]) = *reinterpret_cast<float4*>(v);
        }
        __syncthreads();
        MICRO_FMA();
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int r = m0 + ty * 4 + i;
        if (r >= Mrows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = n0 + tx * 4 + j;
            if (co < g.Cout) atomicAdd(dw + (size_t)r * g.Cout + co, acc[i][j]);
        }
    }
}

template <typename T> int k_conv_wgrad(const T* x, const T* dy, float* dw, ConvGeom g, cudaStream_t st) {
    int Mrows = g.k * g.k * g.Cin;
    long long P = (long long)g.N * g.Ho * g.Wo;
    {
        constexpr int VW1 = VecWidth<T>::value;
        const int nv = g.Cin % VW1 == 0 ? g.Cin / VW1 : 0;
        if (g.k == 1 && g.s == 1 && g.Cout <= 4 && nv >= 1 && nv <= 256 && 256 % nv == 0) {      // 1x1 thin-output head
            const size_t npix = (size_t)P;
            const int ppb = 256 / nv;
            long long want = (long long)((npix + ppb - 1) / ppb);
            const int blocks = (int)(want < 148 * 4 ? want : 148 * 4);
            conv1x1_thin_wgrad<T, VW1><<<blocks, 256, 0, st>>>(x, dy, dw, npix, g.Cin, g.Cout);
            CG_LAUNCH_CHECK();
            return CG_OK;
        }
    }
    if (g.Cout <= 4 && g.Cin % VecWidth<T>::value == 0 && g.k * (g.Cin / VecWidth<T>::value) <= 1024) {
        const int threads = ((g.k * (g.Cin / VecWidth<T>::value) + 31) / 32) * 32;
        int per_sm = 2048 / threads;
        if (per_sm > 32) per_sm = 32;
        long long want_blocks = 148LL * per_sm * 2 / g.k + 1;
        int bands = (int)((want_blocks + g.N - 1) / g.N);
        if (bands < 1) bands = 1;
        if (bands > g.Ho) bands = g.Ho;
        const int rows = (g.Ho + bands - 1) / bands;
        bands = (g.Ho + rows - 1) / rows;
        dim3 grid((unsigned)(g.N * bands), g.k);
        conv_wgrad_skinny_vec<T><<<grid, threads, 0, st>>>(x, dy, dw, g, rows, bands);
        CG_LAUNCH_CHECK();
        return CG_OK;
    }
    if (g.Cout <= 4 && g.Cin % 2 == 0 && g.k * (g.Cin / 2) <= 1024) {
        const int threads = ((g.k * (g.Cin / 2) + 31) / 32) * 32;
        int per_sm = 2048 / threads;
        if (per_sm < 1) per_sm = 1;
        long long want_blocks = 148LL * per_sm * 2 / g.k + 1;          // ~2 waves over (image, row band) x kh
        int bands = (int)((want_blocks + g.N - 1) / g.N);
        if (bands < 1) bands = 1;
        if (bands > g.Ho) bands = g.Ho;
        const int rows = (g.Ho + bands - 1) / bands;
        bands = (g.Ho + rows - 1) / rows;
        dim3 grid((unsigned)(g.N * bands), g.k);
        conv_wgrad_skinny<T><<<grid, threads, 0, st>>>(x, dy, dw, g, rows, bands);
        CG_LAUNCH_CHECK();
        return CG_OK;
    }
    int tiles = cdiv(Mrows, BM) * cdiv(g.Cout, BN);
    int splits = cdiv(148 * 4, tiles);
    int maxsplits = cdiv(P, 256);
    if (splits > maxsplits) splits = maxsplits;
    if (splits < 1) splits = 1;
    int pps = cdiv(cdiv(P, splits), BK) * BK;
    splits = cdiv(P, pps);
    dim3 grid(cdiv(Mrows, BM), cdiv(g.Cout, BN), splits);
    if (g.Cin % 4 == 0) conv_wgrad_simt<T, true><<<grid, 256, 0, st>>>(x, dy, dw, g, pps);
    else conv_wgrad_simt<T, false><<<grid, 256, 0, st>>>(x, dy, dw, g, pps);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// bias gradient: db[c] += sum over rows of dy[rows][C]
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ dy, float* __restrict__ db, size_t rows,
                                                     int C, size_t rchunk) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const size_t r0 = blockIdx.y * rchunk;
    const size_t r1 = r0 + rchunk < rows ? r0 + rchunk : rows;
    float s = 0.f;
    if (c < C)
        for (size_t r = r0 + warp; r < r1; r += 8) s += ldf(dy + r * C + c);
    __shared__ float sh[8][33];
    sh[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && c < C) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += sh[w][lane];
        atomicAdd(db + c, a);
    }
}
template <typename T> int k_colsum(const T* dy, float* db, size_t rows, int C, cudaStream_t st) {
    int cb = cdiv(C, 32);
    int splits = cdiv(148 * 4, cb);
    size_t maxs = (rows + 63) / 64;
    if ((size_t)splits > maxs) splits = (int)maxs;
    if (splits < 1) splits = 1;
    size_t rchunk = (rows + splits - 1) / splits;
    dim3 grid(cb, (unsigned)((rows + rchunk - 1) / rchunk));
    colsum_kernel<T><<<grid, 256, 0, st>>>(dy, db, rows, C, rchunk);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

#define INSTANTIATE(T)                                                                                   \
    template int k_conv_fwd<T>(const T*, const float*, const float*, T*, ConvGeom, int, cudaStream_t);        \
    template int k_conv_dgrad<T>(const T*, const float*, const float*, T*, ConvGeom, int, cudaStream_t); \
    template int k_conv_wgrad<T>(const T*, const T*, float*, ConvGeom, cudaStream_t);                    \
    template int k_colsum<T>(const T*, float*, size_t, int, cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(bf16)
