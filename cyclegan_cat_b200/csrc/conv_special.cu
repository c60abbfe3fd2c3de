// Helper kernels that put the two image-side 7x7 convolutions of the ResNet generator on the tcgen05 kernels.
//
// A 7x7 conv with 3 channels on one side is a bad GEMM (K = 147 or N = 3).  Unfolding the HORIZONTAL taps into the
// channel dimension turns it into a 7x1 (vertical) conv with 21 -> 64/128 "channels":
//   stem (3 -> C):  U[r][ow][kw*3+ci] = xp[r][ow+kw][ci];      y[oh][ow][co]   = sum_kh U[oh+kh][ow][:] . Wv[kh][co][:]
//   head (C -> 3):  S[oh][q][kw*3+co] = sum_kh xp[oh+kh][q][:] . Wh[kh][kw*3+co][:];  y[oh][ow][co] = sum_kw S[oh][ow+kw][kw*3+co]
//   head dgrad:     T[r][q][kw*3+co]  = dy[r][q-kw][co];       dxp[ih][q][ci]  = sum_kh T[ih-kh][q][:] . Whd[kh][ci][:]
//   weight grads:   the same U / T tensors are the operands of wgrad_tc_kernel; small kernels fold the result back.
// The vertical convs run on conv_tc_kernel (box or flat mode), so the activation is re-read 7x instead of 49x.
#include "conv_special.h"

static inline int blocks_for(size_t n) {
    size_t b = (n + 255) / 256;
    return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

// dst[n][h][q][kw*Cs + c] = src[n][h][q + sign*kw][c]  (0 outside [0,Ws)), channels >= k*Cs are zero; dst has 128 channels.
// One thread = one destination pixel: it gathers the k*Cs (<= 32) live values once and writes 16 x 16-byte vectors.
__global__ void unfold_w_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int N, int H, int Ws, int Cs, int Wd,
                                int k, int sign) {
    const size_t total = (size_t)N * H * Wd;
    const int live = k * Cs;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % Wd);
        const size_t row = i / Wd;                          // n*H + h
        const bf16* srow = src + row * Ws * Cs;
        uint4* d = reinterpret_cast<uint4*>(dst + i * 128);
        const bf16 zero = __float2bfloat16(0.f);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            Pack<bf16, 8> pk;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int ch = v * 8 + j;
                bf16 val = zero;
                if (ch < live) {
                    const int kw = ch / Cs, c = ch - kw * Cs;
                    const int ws = q + sign * kw;
                    if (ws >= 0 && ws < Ws) val = srow[(size_t)ws * Cs + c];
                }
                pk.v[j] = val;
            }
            d[v] = *reinterpret_cast<uint4*>(&pk);
        }
#pragma unroll
        for (int v = 4; v < 16; ++v) d[v] = make_uint4(0u, 0u, 0u, 0u);
    }
}
int sp_unfold_w(const bf16* src, bf16* dst, int N, int H, int Ws, int Cs, int Wd, int k, int sign, cudaStream_t st) {
    unfold_w_kernel<<<blocks_for((size_t)N * H * Wd), 256, 0, st>>>(src, dst, N, H, Ws, Cs, Wd, k, sign);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// stem weights: w[kh][kw][ci][co] (fp32 HWIO) -> Wv[kh][co][kw*Cin+ci] bf16, 64 columns (zero padded)
__global__ void pack_stem_kernel(const float* __restrict__ w, bf16* __restrict__ wv, int k, int Cin, int Cout) {
    const int total = k * Cout * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int col = i % 64, co = (i / 64) % Cout, kh = i / (64 * Cout);
        float v = 0.f;
        if (col < k * Cin) { const int kw = col / Cin, ci = col - kw * Cin; v = w[(((size_t)kh * k + kw) * Cin + ci) * Cout + co]; }
        wv[i] = __float2bfloat16(v);
    }
}
// head weights: -> Wh[kh][kw*Cout+co (32 rows)][ci] and Whd[kh][ci][kw*Cout+co (64 columns)]
__global__ void pack_head_kernel(const float* __restrict__ w, bf16* __restrict__ wh, bf16* __restrict__ whd, int k, int Cin,
                                 int Cout) {
    const int n1 = k * 32 * Cin, n2 = k * Cin * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
        if (i < n1) {
            const int ci = i % Cin, row = (i / Cin) % 32, kh = i / (Cin * 32);
            float v = 0.f;
            if (row < k * Cout) { const int kw = row / Cout, co = row - kw * Cout; v = w[(((size_t)kh * k + kw) * Cin + ci) * Cout + co]; }
            wh[i] = __float2bfloat16(v);
        } else {
            const int j = i - n1;
            const int col = j % 64, ci = (j / 64) % Cin, kh = j / (64 * Cin);
            float v = 0.f;
            if (col < k * Cout) { const int kw = col / Cout, co = col - kw * Cout; v = w[(((size_t)kh * k + kw) * Cin + ci) * Cout + co]; }
            whd[j] = __float2bfloat16(v);
        }
    }
}
int sp_pack_stem(const float* w, bf16* wv, int k, int Cin, int Cout, cudaStream_t st) {
    pack_stem_kernel<<<blocks_for((size_t)k * Cout * 64), 256, 0, st>>>(w, wv, k, Cin, Cout);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
int sp_pack_head(const float* w, bf16* wh, bf16* whd, int k, int Cin, int Cout, cudaStream_t st) {
    pack_head_kernel<<<blocks_for((size_t)k * Cin * 96), 256, 0, st>>>(w, wh, whd, k, Cin, Cout);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// second half of the unfolded 7x1 convs: y[n][r][x][c] = bias[c] + sum_kw S[n][r][x + sign*kw][kw*C + c]  (S: 32 channels,
// width Ws; terms outside [0, Ws) are zero).  sign = +1: head forward (x < Wo <= Ws - k + 1);  sign = -1: stem data gradient.
__global__ void diag_sum_kernel(const bf16* __restrict__ S, const float* __restrict__ bias, bf16* __restrict__ y, int N,
                                int R, int Wy, int Ws, int k, int C, int sign) {
    const size_t total = (size_t)N * R * Wy;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wy);
        const size_t row = i / Wy;                          // n*R + r
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int kw = 0; kw < k; ++kw) {
            const int xs = x + sign * kw;
            if (xs < 0 || xs >= Ws) continue;
            const bf16* p = S + (row * Ws + xs) * 32 + kw * C;
            for (int c = 0; c < C; ++c) acc[c] += __bfloat162float(p[c]);
        }
        for (int c = 0; c < C; ++c) y[i * C + c] = __float2bfloat16(acc[c] + (bias ? bias[c] : 0.f));
    }
}
int sp_diag_sum(const bf16* S, const float* bias, bf16* y, int N, int R, int Wy, int Ws, int k, int C, int sign,
                cudaStream_t st) {
    diag_sum_kernel<<<blocks_for((size_t)N * R * Wy), 256, 0, st>>>(S, bias, y, N, R, Wy, Ws, k, C, sign);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// stem data-gradient weights: w[kh][kw][ci][co] -> Wsd[kh][kw*Cin+ci (32 rows)][co]
__global__ void pack_stem_d_kernel(const float* __restrict__ w, bf16* __restrict__ wsd, int k, int Cin, int Cout) {
    const int total = k * 32 * Cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % Cout, row = (i / Cout) % 32, kh = i / (Cout * 32);
        float v = 0.f;
        if (row < k * Cin) { const int kw = row / Cin, ci = row - kw * Cin; v = w[(((size_t)kh * k + kw) * Cin + ci) * Cout + co]; }
        wsd[i] = __float2bfloat16(v);
    }
}
int sp_pack_stem_d(const float* w, bf16* wsd, int k, int Cin, int Cout, cudaStream_t st) {
    pack_stem_d_kernel<<<blocks_for((size_t)k * 32 * Cout), 256, 0, st>>>(w, wsd, k, Cin, Cout);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// fold the tensor-core weight-gradient results back into TF's HWIO layout (+=)
//   stem: t[kh][kw*Cin+ci (128 rows)][co]   -> dw[kh][kw][ci][co]
//   head: t[kh][ci][kw*Cout+co (128 cols)]  -> dw[kh][kw][ci][co]
__global__ void unpack_dw_kernel(const float* __restrict__ t, float* __restrict__ dw, int k, int Cin, int Cout, int head) {
    const int total = k * k * Cin * Cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % Cout, ci = (i / Cout) % Cin, kw = (i / (Cout * Cin)) % k, kh = i / (Cout * Cin * k);
        const float v = head ? t[((size_t)kh * Cin + ci) * 128 + kw * Cout + co] : t[((size_t)kh * 128 + kw * Cin + ci) * Cout + co];
        dw[i] += v;
    }
}
int sp_unpack_dw(const float* t, float* dw, int k, int Cin, int Cout, int head, cudaStream_t st) {
    unpack_dw_kernel<<<blocks_for((size_t)k * k * Cin * Cout), 256, 0, st>>>(t, dw, k, Cin, Cout, head);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
