// Host launchers of the hand-written kernels.  T is the activation type: float (fp32 check
// mode) or bf16.  Parameters, parameter gradients and all statistics are always float32.
#pragma once
#include "common.h"

template <typename T> int k_convert_in(const float* src, T* dst, size_t n, cudaStream_t st);
template <typename T> int k_convert_out(const T* src, float* dst, size_t n, cudaStream_t st);

// ---- CUDA-core implicit-GEMM convolutions (any geometry) ----
template <typename T> int k_conv_fwd(const T* x, const float* w, const float* bias, T* y, ConvGeom g, int accumulate,
                                     cudaStream_t st);
template <typename T> int k_conv_dgrad(const T* dy, const float* w, const float* bias, T* dx, ConvGeom g,
                                       int accumulate, cudaStream_t st);
template <typename T> int k_conv_wgrad(const T* x, const T* dy, float* dw, ConvGeom g, cudaStream_t st);
template <typename T> int k_colsum(const T* dy, float* db, size_t rows, int C, cudaStream_t st);

// ---- instance norm ----
template <typename T> int k_in_stats(const T* x, float* stats, int N, int P, int C, float eps, cudaStream_t st,
                                     bool zeroed = false);      // zeroed: the caller has already cleared `stats`
int k_in_finalize(const float* raw, float* stats, int NC, int P, float eps, cudaStream_t st);      // raw may equal stats (in place)
template <typename T> int k_in_apply(const T* x, T* y, const float* stats, const float* gamma, const float* beta,
                                     int act, float slope, int N, int P, int C, cudaStream_t st);
// bulk-copy pipelined variants (kernels_stream.cu); k_in_stream_ok says whether a (P, C) plane qualifies
template <typename T> bool k_in_stream_ok(const void* p0, const void* p1, const void* p2, int P, int C);
// y (nullable) = act(norm(x)) [+ res]; ypad (nullable, pad > 0) = the same values as the reflection-padded
// [N][H+2p][W+2p][C] tensor (P = H*W)
// raw (nullable): `stats` has not been finalized yet -- the kernel computes (mean, rstd) from the raw (sum x, sum x^2) table
// itself and one CTA per image stores them into `stats` for the backward (no in_finalize_kernel launch)
template <typename T> int k_in_apply_stream(const T* x, const T* res, T* y, T* ypad, float* stats, const float* gamma,
                                            const float* beta, int act, float slope, int N, int P, int C, int W, int pad,
                                            cudaStream_t st, const float* raw = nullptr, float eps = 0.f);
template <typename T> int k_in_stats_stream(const T* x, float* sums, int N, int P, int C, cudaStream_t st);
template <typename T> int k_in_bwd_reduce_stream(const T* x, const T* dy, const float* stats, const float* gamma,
                                                 const float* beta, float* sums, int act, float slope, int N, int P, int C,
                                                 cudaStream_t st);
// dgamma / dbeta (nullable): the affine parameter gradients are added by the kernel itself (one CTA per image)
template <typename T> int k_in_bwd_apply_stream(const T* x, const T* dy, T* dx, const float* stats, const float* sums,
                                                const float* gamma, const float* beta, int act, float slope, int N, int P,
                                                int C, int W, int halo, cudaStream_t st, float* dgamma = nullptr,
                                                float* dbeta = nullptr);
template <typename T> int k_in_bwd(const T* x, const T* dy, T* dx, const float* stats, const float* gamma,
                                   const float* beta, float* dgamma, float* dbeta, float* scratch, int act,
                                   float slope, int N, int P, int C, int accumulate, cudaStream_t st,
                                   int halo = 0, int W = 0, bool zeroed = false,    // zeroed: `scratch` is already cleared
                                   int bn_group = 0);   // > 0: BatchNormalization, statistics pooled over bn_group samples

// ---- elementwise / data movement ----
template <typename T> int k_act_fwd(const T* x, T* y, size_t n, int act, float slope, cudaStream_t st);
template <typename T> int k_act_bwd(const T* y, const T* dy, T* dx, size_t n, int act, float slope, int accumulate,
                                    cudaStream_t st);
template <typename T> int k_rpad_fwd(const T* x, T* y, int N, int H, int W, int C, int p, cudaStream_t st);
template <typename T> int k_rpad_bwd(const T* dy, T* dx, int N, int H, int W, int C, int p, int accumulate,
                                     cudaStream_t st);
// dx already holds the interior term of every pixel (written by the tensor-core data gradient in fold mode): add the
// mirrored border terms of dy [N][H+2p][W+2p][C]
template <typename T> int k_rpad_bwd_border(const T* dy, T* dx, int N, int H, int W, int C, int p, cudaStream_t st);
template <typename T> int k_add(const T* a, const T* b, T* y, size_t n, cudaStream_t st);
template <typename T> int k_copy_acc(const T* src, T* dst, size_t n, int accumulate, cudaStream_t st);
template <typename T> int k_slice_copy(const T* src, int Cs, int so, T* dst, int Cd, int doff, int Cc, size_t npix,
                                       int accumulate, cudaStream_t st);
// `cat` (nullable): zero-copy concat -- the concat tensor / its gradient with cat_c channels, slice at channel cat_off
// (avgpool: the skip copy written / the skip slice added in the same pass; upsample: writes to / gathers from its slice)
template <typename T> int k_avgpool_fwd(const T* x, T* y, int N, int H, int W, int C, cudaStream_t st, T* cat = nullptr,
                                        int cat_c = 0, int cat_off = 0);
template <typename T> int k_avgpool_bwd(const T* dy, T* dx, int N, int H, int W, int C, int accumulate, cudaStream_t st,
                                        const T* dcat = nullptr, int cat_c = 0, int cat_off = 0);
template <typename T> int k_upsample_fwd(const T* x, T* y, int N, int H, int W, int C, cudaStream_t st, T* cat = nullptr,
                                         int cat_c = 0, int cat_off = 0);
template <typename T> int k_upsample_bwd(const T* dy, T* dx, int N, int H, int W, int C, int accumulate, cudaStream_t st,
                                         const T* dcat = nullptr, int cat_c = 0, int cat_off = 0);

// ---- losses (value + gradient seed in one pass), Adam ----
// sum_out += sum_i L(d_i, target); correct_out += #{(d_i > 0.5) == target}; grad_i = grad_scale * dL/dd_i
template <typename T> int k_adv_loss(const T* d, size_t n, float target, int kind, float grad_scale, T* grad,
                                     float* sum_out, float* correct_out, cudaStream_t st);
// sum_out += sum |real - gen|; grad (+)= grad_scale * sign(gen - real)
template <typename T> int k_l1_loss(const T* real, const T* gen, size_t n, float grad_scale, T* grad, int accumulate,
                                    float* sum_out, cudaStream_t st);
int k_adam(float* p, const float* g, float* m, float* v, size_t n, float lr_t, float b1, float b2, float eps,
           float grad_scale, cudaStream_t st);

// ---- optional config paths and input pipeline (kernels_extra.cu) ----
// coefficients of one optimizer step, computed on the host in double (trainer.cu)
struct OptCoef { float lr, b1, b2, eps, c_m, c_v, r_t, gscale; int rect; };
int k_opt_step(int kind, float* p, const float* g, float* m, float* v, size_t n, const OptCoef& c, cudaStream_t st);
// BatchNormalization on the instance-norm tables: `group` consecutive samples form one Keras call
template <typename T> int k_in_stats_raw(const T* x, float* stats, int N, int P, int C, cudaStream_t st, bool zeroed);
int k_bn_finalize(float* stats, float* bstat, int N, int C, int group, int P, float eps, cudaStream_t st);
int k_bn_fill(float* stats, const float* moving_mean, const float* moving_var, int N, int C, float eps, cudaStream_t st);
int k_bn_update_moving(float* moving_mean, float* moving_var, const float* bstat, int C, float momentum, cudaStream_t st);
int k_bn_pool_sums(float* sums, int N, int C, int group, cudaStream_t st);
// Dropout: the mask of element e of group g is a hash of (seed, counter, call_id[g], layer, e)
struct DropKey {
    unsigned long long seed, ctr_host;
    const unsigned long long* ctr_dev;      // when non-null the counter is read on the device (CUDA-graph replays)
    int call_id[4];
    int layer;
};
template <typename T> int k_dropout_fwd(const T* x, T* y, size_t group_elems, int groups, float rate, const DropKey& key,
                                        int training, cudaStream_t st);
template <typename T> int k_dropout_bwd(const T* dy, T* dx, size_t group_elems, int groups, float rate, const DropKey& key,
                                        int training, int accumulate, cudaStream_t st);
int k_set_counter(unsigned long long* ctr_dev, unsigned long long value, cudaStream_t st);
