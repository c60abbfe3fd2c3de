// Bulk-copy pipelined instance-norm kernels (forward apply, backward reduction, backward apply) for sm_100a.
//
// These three passes are pure HBM streaming (2-3 activation tensors per launch, tens to hundreds of MB) and the
// register-resident versions in kernels_elem.cu are limited by the number of loads a thread can keep in flight
// (in_bwd_apply_fast_kernel: 2 x 16 B per thread, ~32 KB per SM -> 2.4-3 TB/s).  Here the loads are issued by ONE
// producer thread per CTA as 8 KB cp.async.bulk copies into a shared-memory ring (mbarrier complete_tx), so
// 2 CTAs x 5-8 stages x 8-16 KB = 130-190 KB per SM is in flight independent of the consumers' registers; the 8 consumer
// warps read a tile from shared memory, free the slot, do the per-channel arithmetic with constants held in registers
// and store 16-byte vectors straight to global memory.
//
//   tile   = 8192 contiguous bytes of one image of one operand (= 512 16-byte channel vectors = 512/CV pixels)
//   grid   = (G, N) with G*N ~ 2 CTAs per SM; CTA (g, n) walks tiles g, g+G, ... of image n
//   block  = 288 threads: warps 0-7 consume, warp 8 lane 0 produces
// Semantics are those of in_apply_fast_kernel / in_reduce_kernel<MODE 1> / in_bwd_apply_fast_kernel (kernels_elem.cu),
// i.e. TFA InstanceNormalization forward and its gradient (reference: cyclegan/unet.py:28-33, resnet.py:26-34 use
// tfa.layers.InstanceNormalization followed by ReLU / LeakyReLU).
#include <stdint.h>
#include <stdlib.h>

#include "common.h"
#include "kernels.h"
#include "prof.h"
#include "ptx_async.h"

namespace {
constexpr int ST_TILE = 8192;
constexpr int ST_CONSUMERS = 256;
constexpr int ST_THREADS = ST_CONSUMERS + 32;

template <typename T>
struct StreamArgs {
    const T* x;
    const T* dy;
    T* out;
    T* out2;
    const float* stats;
    const float* sums_in;
    float* sums_out;
    const float* gamma;
    const float* beta;
    int act;
    float slope;
    int P, C, W, halo;
    int pad;                  // MODE 0: > 0 writes the reflection-padded [H+2p][W+2p] tensor (ReflectionPadding2D fused in)
    float invP;
    int stages, tiles_per_img;
    const float* raw;         // forward modes: raw (sum x, sum x^2) table to finalize on the fly (null: `stats` is final)
    float* stats_w;           //   ... and where one CTA per image stores the finalized (mean, rstd)
    float eps;
    float* dgamma;            // MODE 2, affine: d gamma / d beta += the per-image backward sums (one CTA per image adds them,
    float* dbeta;             // so the separate in_param_grad_kernel launch disappears)
};

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(ST_CONSUMERS) : "memory"); }

// MODE 0: out = act(x*sc + sh)                     (forward apply; one operand)
// MODE 1: sums_out += (sum g, sum g*xhat)          (backward reduction; x and dy)
// MODE 2: out = k*g - c1 - xhat*c2                 (backward apply; x and dy; optional zero-bordered output)
// MODE 3: out = act(x*sc + sh) + res               (forward apply fused with the residual add; res comes in as `dy`)
// MODE 4: sums_out += (sum x, sum x^2)             (forward statistics; one operand)
// MODES 0 and 3 write the plain tensor (`out`, nullable) and/or the reflection-padded one (`out2`, pad > 0).
template <typename T, int VEC, int MODE, bool AFFINE>
__global__ void __launch_bounds__(ST_THREADS, (MODE == 0 || MODE == 4) ? 3 : 2) in_stream_kernel(const StreamArgs<T> a) {
    extern __shared__ __align__(128) uint8_t st_smem[];
    constexpr int NOPS = (MODE == 0 || MODE == 4) ? 1 : 2;
    constexpr bool FWD = MODE == 0 || MODE == 3;
    constexpr bool RED = MODE == 1 || MODE == 4;      // reductions: per-thread partial sums, CTA-level combine, atomics
    const int S = a.stages;
    const uint32_t tiles_u32 = smem_u32(st_smem);
    const uint32_t bars_u32 = tiles_u32 + S * NOPS * ST_TILE;       // full[S], empty[S]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.y, G = gridDim.x, TI = a.tiles_per_img;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(bars_u32 + 8 * s, 1);
            mbar_init(bars_u32 + 8 * (S + s), ST_CONSUMERS / 32);
        }
        fence_barrier_init();
    }
    __syncthreads();
    griddep_launch();        // PDL: see ptx_async.h
    griddep_wait();

    const size_t img_elems = (size_t)a.P * a.C;
    if (warp == ST_CONSUMERS / 32) {                 // ---- producer ----
        if (lane == 0) {
            const uint8_t* xs = reinterpret_cast<const uint8_t*>(a.x + (size_t)n * img_elems);
            const uint8_t* gs = NOPS == 2 ? reinterpret_cast<const uint8_t*>(a.dy + (size_t)n * img_elems) : nullptr;
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < TI; t += G) {
                mbar_wait(bars_u32 + 8 * (S + s), ph ^ 1);
                const uint32_t full = bars_u32 + 8 * s;
                const uint32_t dst = tiles_u32 + s * NOPS * ST_TILE;
                mbar_expect_tx(full, NOPS * ST_TILE);
                bulk_load(dst, xs + (size_t)t * ST_TILE, ST_TILE, full);
                if (NOPS == 2) bulk_load(dst + ST_TILE, gs + (size_t)t * ST_TILE, ST_TILE, full);
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ---- consumers ----
    const int C = a.C, CV = C / VEC;
    if constexpr (MODE == 2 && AFFINE) {
        if (a.dgamma && blockIdx.x == 0)          // d beta = sum g, d gamma = sum g * xhat over the images: one CTA per image adds its sums
            for (int c = tid; c < C; c += ST_CONSUMERS) {
                atomicAdd(a.dbeta + c, a.sums_in[((size_t)n * C + c) * 2]);
                atomicAdd(a.dgamma + c, a.sums_in[((size_t)n * C + c) * 2 + 1]);
            }
    }
    const int cv = tid % CV, prow = tid / CV, rows = ST_CONSUMERS / CV;      // a tile holds 2*rows pixels
    float k0[VEC], k1[VEC], k2[MODE == 2 ? VEC : 1], k3[MODE == 2 ? VEC : 1], k4[MODE == 2 ? VEC : 1];
    float kg[(AFFINE && !FWD) ? VEC : 1], ke[(AFFINE && !FWD) ? VEC : 1];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        if constexpr (MODE == 4) { k0[j] = k1[j] = 0.f; continue; }
        const int c = cv * VEC + j;
        float mean, rstd;
        if (FWD && a.raw) {
            in_mean_rstd(a.raw[((size_t)n * C + c) * 2], a.raw[((size_t)n * C + c) * 2 + 1], a.invP, a.eps, mean, rstd);
            if (blockIdx.x == 0 && prow == 0) {
                a.stats_w[((size_t)n * C + c) * 2] = mean;
                a.stats_w[((size_t)n * C + c) * 2 + 1] = rstd;
            }
        } else {
            mean = a.stats[((size_t)n * C + c) * 2];
            rstd = a.stats[((size_t)n * C + c) * 2 + 1];
        }
        const float ga = AFFINE ? a.gamma[c] : 1.f, be = AFFINE ? a.beta[c] : 0.f;
        if constexpr (FWD) {
            in_scale_shift(mean, rstd, ga, be, k0[j], k1[j]);      // y = act(v*k0 + k1)
        } else {
            in_scale_shift(mean, rstd, 1.f, 0.f, k0[j], k1[j]);     // xhat = v*k0 + k1 (= the forward's pre-activation when not affine)
            // the activation mask is taken from the pre-activation computed EXACTLY as the forward pass computes it
            // (fmaf(v, rstd*gamma, beta - mean*rstd*gamma)): forward and backward then agree on every unit, bit for bit
            if constexpr (AFFINE) in_scale_shift(mean, rstd, ga, be, kg[j], ke[j]);
            if constexpr (MODE == 2) {
                k2[j] = rstd * ga;                 // r = k2*g - (xhat*k4 + k3)
                k3[j] = k2[j] * a.sums_in[((size_t)n * C + c) * 2] * a.invP;
                k4[j] = k2[j] * a.sums_in[((size_t)n * C + c) * 2 + 1] * a.invP;
            }
        }
    }
    float acc_s[RED ? VEC : 1], acc_ss[RED ? VEC : 1];
    if constexpr (RED) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc_s[j] = acc_ss[j] = 0.f;
    }
    const int Wp = a.W + 2 * a.halo;
    const size_t out_img = MODE == 2 && a.halo > 0 ? (size_t)(a.P / a.W + 2 * a.halo) * Wp * C : img_elems;
    T* outp = (RED || !a.out) ? nullptr : a.out + (size_t)n * out_img + (size_t)cv * VEC;
    T* out2p = (FWD && a.pad > 0) ? a.out2 + (size_t)n * (a.P / a.W + 2 * a.pad) * (a.W + 2 * a.pad) * C + (size_t)cv * VEC : nullptr;

    // zero-bordered output (MODE 2): row / first column of the current tile, advanced incrementally (valid when W is a
    // multiple of the tile) -- no division per pixel (backward apply 43.8 -> 39.2 us at the trunk shape)
    const int tpix = 2 * rows;
    const bool row_tiles = MODE == 2 && a.halo > 0 && (a.W % tpix) == 0;
    int th = 0, tw0 = 0, dth = 0, dtw = 0;
    if (row_tiles) {
        const int p0 = blockIdx.x * tpix, dp = G * tpix;
        th = p0 / a.W; tw0 = p0 - th * a.W;
        dth = dp / a.W; dtw = dp - dth * a.W;
    }
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < TI; t += G) {
        mbar_wait(bars_u32 + 8 * s, ph);
        const uint8_t* tile = st_smem + s * NOPS * ST_TILE;
        float v[2][VEC], g[NOPS == 2 ? 2 : 1][VEC];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            load_vec<T, VEC>(reinterpret_cast<const T*>(tile + (tid + u * ST_CONSUMERS) * 16), v[u]);
            if constexpr (NOPS == 2) load_vec<T, VEC>(reinterpret_cast<const T*>(tile + ST_TILE + (tid + u * ST_CONSUMERS) * 16), g[u]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars_u32 + 8 * (S + s));         // slot is free as soon as the tile sits in registers
        if (++s == S) { s = 0; ph ^= 1; }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int p = t * 2 * rows + u * rows + prow;              // pixel index inside the image
            if constexpr (MODE == 4) {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    acc_s[j] += v[u][j];
                    acc_ss[j] = fmaf(v[u][j], v[u][j], acc_ss[j]);
                }
            } else if constexpr (FWD) {
                // the activation switch is taken once per vector, not once per element (the forward kernels are
                // instruction-bound: ~170 warp instructions per 16-byte vector before this)
                if (a.act == CG_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) v[u][j] = fmaxf(in_pre(v[u][j], k0[j], k1[j]), 0.f);
                } else if (a.act == CG_ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) v[u][j] = in_pre(v[u][j], k0[j], k1[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) v[u][j] = act_fwd(in_pre(v[u][j], k0[j], k1[j]), a.act, a.slope);
                }
                if constexpr (MODE == 3) {
#pragma unroll
                    for (int j = 0; j < VEC; ++j) v[u][j] += g[u][j];
                }
                if (outp) store_vec<T, VEC>(outp + (size_t)p * C, v[u]);
                if (out2p) {
                    // reflection padding (cyclegan/resnet.py:5-23, tf.pad REFLECT): interior pixel (h, w) lands at
                    // (h+p, w+p) and, when it lies within p of an edge (but not on it), at its mirror image(s) too
                    const int pd = a.pad, H = a.P / a.W, Wq = a.W + 2 * pd;
                    const int h = p / a.W, w = p - h * a.W;
                    int hr[3], wr[3], nh = 1, nw = 1;
                    hr[0] = h + pd; wr[0] = w + pd;
                    if (h >= 1 && h <= pd) hr[nh++] = pd - h;
                    if (h >= H - 1 - pd && h <= H - 2) hr[nh++] = pd + 2 * (H - 1) - h;
                    if (w >= 1 && w <= pd) wr[nw++] = pd - w;
                    if (w >= a.W - 1 - pd && w <= a.W - 2) wr[nw++] = pd + 2 * (a.W - 1) - w;
                    for (int ia = 0; ia < nh; ++ia)
                        for (int ib = 0; ib < nw; ++ib) store_vec<T, VEC>(out2p + ((size_t)hr[ia] * Wq + wr[ib]) * C, v[u]);
                }
            } else {
                float o[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const float xh = in_pre(v[u][j], k0[j], k1[j]);
                    float pre = xh;
                    if constexpr (AFFINE) pre = in_pre(v[u][j], kg[j], ke[j]);
                    const float gg = g[u][j] * act_grad_from_out(pre, a.act, a.slope);
                    if constexpr (MODE == 1) {
                        acc_s[j] += gg;
                        acc_ss[j] = fmaf(gg, xh, acc_ss[j]);
                    } else {
                        o[j] = fmaf(k2[j], gg, -fmaf(xh, k4[j], k3[j]));
                    }
                }
                if constexpr (MODE == 2) {
                    size_t po = p;
                    if (a.halo > 0) {
                        int h, w;
                        if (row_tiles) { h = th; w = tw0 + u * rows + prow; }
                        else { h = p / a.W; w = p - h * a.W; }
                        po = (size_t)(h + a.halo) * Wp + (w + a.halo);
                    }
                    store_vec<T, VEC>(outp + po * C, o);
                }
            }
        }
        if (row_tiles) { th += dth; tw0 += dtw; if (tw0 >= a.W) { tw0 -= a.W; ++th; } }
    }

    if constexpr (MODE == 2) if (a.halo > 0) {      // zero border of the [H+2h][W+2h] output
        const int H = a.P / a.W, hl = a.halo;
        const int nb = 2 * hl * Wp + 2 * hl * H;
        float z[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) z[j] = 0.f;
        for (int b = blockIdx.x * rows + prow; b < nb; b += G * rows) {
            int hp, wp;
            if (b < 2 * hl * Wp) {
                hp = b / Wp;
                wp = b - hp * Wp;
                if (hp >= hl) hp += H;
            } else {
                const int r = b - 2 * hl * Wp;
                const int h = r / (2 * hl), j = r - h * 2 * hl;
                hp = h + hl;
                wp = j < hl ? j : a.W + j;
            }
            store_vec<T, VEC>(outp + ((size_t)hp * Wp + wp) * C, z);
        }
    }

    if constexpr (RED) {                            // CTA-level reduction over the pixel rows, then one atomic per channel
        consumer_sync();                            // every consumer is past its last tile read: reuse the ring
        float* red = reinterpret_cast<float*>(st_smem);              // [256][2*VEC+1]
        constexpr int RS = 2 * VEC + 1;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            red[tid * RS + j] = acc_s[j];
            red[tid * RS + VEC + j] = acc_ss[j];
        }
        consumer_sync();
        for (int o = tid; o < 2 * C; o += ST_CONSUMERS) {
            const int c = o >> 1, which = o & 1;
            const int ccv = c / VEC, j = c - ccv * VEC;
            float sum = 0.f;
            for (int r = 0; r < rows; ++r) sum += red[(r * CV + ccv) * RS + which * VEC + j];
            atomicAdd(&a.sums_out[((size_t)n * C + c) * 2 + which], sum);
        }
    }
}

inline bool cv_ok(int C, int VW) {
    if (C % VW) return false;
    const int cv = C / VW;
    return cv <= ST_CONSUMERS && (ST_CONSUMERS % cv) == 0;
}

template <typename T, int MODE, bool AFFINE>
int launch_stream(StreamArgs<T>& a, int N, cudaStream_t st) {
    constexpr int VW = VecWidth<T>::value;
    constexpr int NOPS = (MODE == 0 || MODE == 4) ? 1 : 2;
    a.tiles_per_img = (int)(((size_t)a.P * a.C * sizeof(T)) / ST_TILE);
    a.stages = NOPS == 1 ? 8 : 6;
    size_t smem = (size_t)a.stages * NOPS * ST_TILE + 16 * a.stages;
    if (MODE == 1 || MODE == 4) {
        const size_t red = (size_t)ST_CONSUMERS * (2 * VW + 1) * sizeof(float);
        if (smem < red) smem = red;
    }
    auto kern = in_stream_kernel<T, VW, MODE, AFFINE>;
    static std::atomic<unsigned long long> attr_done{0};                  // one flag per template instantiation
    if (cg_first_on_device(attr_done)) {
        CG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(6 * 2 * ST_TILE + 1024)));
    }
    const int per_sm = (MODE == 0 || MODE == 4) ? 3 : 2;      // one-operand modes: 64 KB of ring and <= 75 registers -> 3 CTAs per SM
    int G = (per_sm * 148) / N;                      // all CTAs resident at once, no second wave
    if (G > a.tiles_per_img) G = a.tiles_per_img;
    if (G < 1) G = 1;
    static const bool prof_stream = [] { const char* e = getenv("CG_PROF_STREAM"); return e && e[0] == '1'; }();
    const int pi = prof_stream ? prof_begin(st) : -1;      // diagnostic: in-step timing of the streaming kernels (CSV kinds 8+MODE)
    launch_pdl(kern, dim3(G, N), dim3(ST_THREADS), smem, st, a);
    if (pi >= 0) {
        const double bytes = (double)N * a.P * a.C * sizeof(T) * (MODE == 4 ? 1 : (MODE == 0 || MODE == 1) ? 2 : 3);
        prof_end(pi, st, bytes, prof_key(8 + MODE, MODE, 0, a.C, a.tiles_per_img, N));
    }
    CG_LAUNCH_CHECK();
    return CG_OK;
}
}   // namespace

template <typename T> bool k_in_stream_ok(const void* p0, const void* p1, const void* p2, int P, int C) {
    static const bool off = [] { const char* e = getenv("CG_DISABLE_STREAM"); return e && e[0] == '1'; }();   // test hook
    if (off) return false;
    if (!cv_ok(C, VecWidth<T>::value)) return false;
    if (((size_t)P * C * sizeof(T)) % ST_TILE) return false;
    return !(((uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2) & 15);
}

template <typename T> int k_in_apply_stream(const T* x, const T* res, T* y, T* ypad, float* stats, const float* gamma,
                                            const float* beta, int act, float slope, int N, int P, int C, int W, int pad,
                                            cudaStream_t st, const float* raw, float eps) {
    StreamArgs<T> a{};
    a.raw = raw; a.stats_w = stats; a.eps = eps;
    a.x = x; a.dy = res; a.out = y; a.out2 = ypad; a.stats = stats; a.gamma = gamma; a.beta = beta; a.act = act; a.slope = slope;
    a.P = P; a.C = C; a.W = (ypad && pad > 0) ? W : P; a.pad = ypad ? pad : 0; a.halo = 0; a.invP = 1.f / (float)P;
    if (res) return gamma ? launch_stream<T, 3, true>(a, N, st) : launch_stream<T, 3, false>(a, N, st);
    return gamma ? launch_stream<T, 0, true>(a, N, st) : launch_stream<T, 0, false>(a, N, st);
}

template <typename T> int k_in_stats_stream(const T* x, float* sums, int N, int P, int C, cudaStream_t st) {
    StreamArgs<T> a{};
    a.x = x; a.sums_out = sums; a.P = P; a.C = C; a.W = P; a.invP = 1.f / (float)P;
    return launch_stream<T, 4, false>(a, N, st);
}

template <typename T> int k_in_bwd_reduce_stream(const T* x, const T* dy, const float* stats, const float* gamma,
                                                 const float* beta, float* sums, int act, float slope, int N, int P, int C,
                                                 cudaStream_t st) {
    StreamArgs<T> a{};
    a.x = x; a.dy = dy; a.stats = stats; a.sums_out = sums; a.gamma = gamma; a.beta = beta; a.act = act; a.slope = slope;
    a.P = P; a.C = C; a.W = P; a.halo = 0; a.invP = 1.f / (float)P;
    return gamma ? launch_stream<T, 1, true>(a, N, st) : launch_stream<T, 1, false>(a, N, st);
}

template <typename T> int k_in_bwd_apply_stream(const T* x, const T* dy, T* dx, const float* stats, const float* sums,
                                                const float* gamma, const float* beta, int act, float slope, int N, int P,
                                                int C, int W, int halo, cudaStream_t st, float* dgamma, float* dbeta) {
    StreamArgs<T> a{};
    a.dgamma = gamma ? dgamma : nullptr; a.dbeta = dbeta;
    a.x = x; a.dy = dy; a.out = dx; a.stats = stats; a.sums_in = sums; a.gamma = gamma; a.beta = beta; a.act = act;
    a.slope = slope; a.P = P; a.C = C; a.W = halo > 0 ? W : P; a.halo = halo; a.invP = 1.f / (float)P;
    return gamma ? launch_stream<T, 2, true>(a, N, st) : launch_stream<T, 2, false>(a, N, st);
}

#define INSTANTIATE(T)                                                                                                     \
    template int k_in_stats_stream<T>(const T*, float*, int, int, int, cudaStream_t);                                      \
    template bool k_in_stream_ok<T>(const void*, const void*, const void*, int, int);                                      \
    template int k_in_apply_stream<T>(const T*, const T*, T*, T*, float*, const float*, const float*, int, float,          \
                                      int, int, int, int, int, cudaStream_t, const float*, float);                                              \
    template int k_in_bwd_reduce_stream<T>(const T*, const T*, const float*, const float*, const float*, float*, int,      \
                                           float, int, int, int, cudaStream_t);                                            \
    template int k_in_bwd_apply_stream<T>(const T*, const T*, T*, const float*, const float*, const float*, const float*,  \
                                          int, float, int, int, int, int, int, cudaStream_t, float*, float*);
INSTANTIATE(float)
INSTANTIATE(bf16)
