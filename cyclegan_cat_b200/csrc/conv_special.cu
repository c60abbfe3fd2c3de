// Helper kernels that put the two image-side 7x7 convolutions of the ResNet generator on the tcgen05 kernels.
//
// A 7x7 conv with 3 channels on one side is a bad GEMM (K = 147 or N = 3).
// Thin INPUT side (stem forward, head data gradient, both weight gradients): the 7 horizontal taps AND 3 vertical taps are
// unfolded into the channel dimension, 3*7*3 = 63 live channels of a dense 64-channel tensor (one SWIZZLE_128B K chunk),
// which leaves a 3-tap vertical conv with row offsets 0, 3, 6 (weights of the non-existing kernel rows 7, 8 are zero):
//   stem (3 -> C):  U[r][ow][(j*7+kw)*3+ci] = xp[r+j][ow+kw][ci], j<3;   y[oh][ow][co] = sum_t U[oh+3t][ow][:] . Wv[t][co][:]
//   head dgrad:     T[r][q][(j*7+kw)*3+co]  = dy[r-j][q-kw][co];         dxp[ih][q][ci] = sum_t T[ih-3t][q][:] . Whd[t][ci][:]
//   weight grads:   U / T are the X operand of wgrad_tc_kernel (two taps stacked into its 128 rows), the 64-channel
//                   activation (dy of the stem, padded input of the head) is the other; small kernels fold the result back.
// Thin OUTPUT side (head forward, stem data gradient): only the horizontal taps are unfolded, into the GEMM N dimension:
//   head (C -> 3):  S[oh][q][kw*3+co] = sum_kh xp[oh+kh][q][:] . Wh[kh][kw*3+co][:];  y[oh][ow][co] = sum_kw S[oh][ow+kw][kw*3+co]
// All of these run on conv_tc_kernel / wgrad_tc_kernel (box or flat mode).
#include "conv_special.h"

static inline int blocks_for(size_t n) {
    size_t b = (n + 255) / 256;
    return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

// dst[n][r][q][(j*k + kw)*Cs + c] = src[n][r + sign*j][q + sign*kw][c]  for j < 3 (0 outside the source), 64 channels,
// channels >= 3*k*Cs are zero.  grid = (N*Hd rows, row segments of UNF_SEG pixels).  A block first copies the three source
// row windows it needs ((UNF_SEG + k - 1) pixels each, zero outside the source) into shared memory with coalesced loads,
// then walks its segment 32 pixels at a time: one thread = one 16-byte vector of one destination pixel, gathering its 8
// values from the windows; the channel -> (j, kw, c) decode (two integer divisions) is done once per block.
constexpr int UNF_SEG = 256;                              // destination pixels per block
constexpr int UNF_MAXW = (UNF_SEG + 15) * 4;              // window elements per row: k <= 16, Cs <= 4
__global__ void __launch_bounds__(256) unfold_w_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int Hs, int Ws,
                                                       int Cs, int Hd, int Wd, int k, int sign) {
    __shared__ int lut[64];                              // offset into the window: j*winE + kw'*Cs + c, or -1 (zero channel)
    __shared__ bf16 win[3 * UNF_MAXW];
    const int row = blockIdx.x, n = row / Hd, r = row - n * Hd, q0 = blockIdx.y * UNF_SEG;
    const int seg = min(UNF_SEG, Wd - q0);
    const int winP = seg + k - 1, winE = winP * Cs;      // window: source columns [c0, c0 + winP)
    const int c0 = sign > 0 ? q0 : q0 - (k - 1);
    if (threadIdx.x < 64) {
        const int ch = threadIdx.x, row_live = k * Cs;
        int e = -1;
        if (ch < 3 * row_live) {
            const int j = ch / row_live, rem = ch - j * row_live, kw = rem / Cs, c = rem - kw * Cs;
            e = j * winE + (sign > 0 ? kw : (k - 1 - kw)) * Cs + c;      // pixel q reads window column (q - q0) + kw'
        }
        lut[ch] = e;
    }
    const bf16 zero = __float2bfloat16(0.f);
    for (int j = 0; j < 3; ++j) {                        // a window row is one contiguous run of the source row
        const int hs = r + sign * j;
        const bool row_ok = hs >= 0 && hs < Hs;
        const bf16* srow = src + ((size_t)n * Hs + (row_ok ? hs : 0)) * Ws * Cs;
        for (int i = threadIdx.x; i < winE; i += 256) {
            const int e = c0 * Cs + i;                   // element index inside the source row
            win[j * winE + i] = (row_ok && e >= 0 && e < Ws * Cs) ? srow[e] : zero;
        }
    }
    __syncthreads();
    const int v = threadIdx.x & 7;
    int d[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] = lut[v * 8 + e];
    uint4* drow = reinterpret_cast<uint4*>(dst) + (size_t)row * Wd * 8;
    for (int ql = threadIdx.x >> 3; ql < seg; ql += 32) {
        Pack<bf16, 8> pk;
#pragma unroll
        for (int e = 0; e < 8; ++e) pk.v[e] = d[e] >= 0 ? win[d[e] + ql * Cs] : zero;
        drow[(size_t)(q0 + ql) * 8 + v] = *reinterpret_cast<uint4*>(&pk);
    }
}
int sp_unfold_w(const bf16* src, bf16* dst, int N, int Hs, int Ws, int Cs, int Hd, int Wd, int k, int sign, cudaStream_t st) {
    unfold_w_kernel<<<dim3(N * Hd, (Wd + UNF_SEG - 1) / UNF_SEG), 256, 0, st>>>(src, dst, Hs, Ws, Cs, Hd, Wd, k, sign);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// stem weights: w[kh][kw][ci][co] (fp32 HWIO) -> Wv[t][co][(j*k+kw)*Cin+ci] bf16 with kh = 3t+j, 64 columns (zero padded)
__global__ void pack_stem_kernel(const float* __restrict__ w, bf16* __restrict__ wv, int k, int Cin, int Cout) {
    const int nt = (k + 2) / 3, total = nt * Cout * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int col = i % 64, co = (i / 64) % Cout, t = i / (64 * Cout);
        float v = 0.f;
        if (col < 3 * k * Cin) {
            const int j = col / (k * Cin), rem = col - j * k * Cin, kw = rem / Cin, ci = rem - kw * Cin, kh = 3 * t + j;
            if (kh < k) v = w[(((size_t)kh * k + kw) * Cin + ci) * Cout + co];
        }
        wv[i] = __float2bfloat16(v);
    }
}
// head weights: -> Wh[kh][kw*Cout+co (32 rows)][ci] and Whd[t][ci][(j*k+kw)*Cout+co (64 columns)] with kh = 3t+j
__global__ void pack_head_kernel(const float* __restrict__ w, bf16* __restrict__ wh, bf16* __restrict__ whd, int k, int Cin,
                                 int Cout) {
    const int n1 = k * 32 * Cin, n2 = ((k + 2) / 3) * Cin * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
        if (i < n1) {
            const int ci = i % Cin, row = (i / Cin) % 32, kh = i / (Cin * 32);
            float v = 0.f;
            if (row < k * Cout) { const int kw = row / Cout, co = row - kw * Cout; v = w[(((size_t)kh * k + kw) * Cin + ci) * Cout + co]; }
            wh[i] = __float2bfloat16(v);
        } else {
            const int j = i - n1;
            const int col = j % 64, ci = (j / 64) % Cin, t = j / (64 * Cin);
            float v = 0.f;
            if (col < 3 * k * Cout) {
                const int jj = col / (k * Cout), rem = col - jj * k * Cout, kw = rem / Cout, co = rem - kw * Cout, kh = 3 * t + jj;
                if (kh < k) v = w[(((size_t)kh * k + kw) * Cin + ci) * Cout + co];
            }
            whd[j] = __float2bfloat16(v);
        }
    }
}
int sp_pack_stem(const float* w, bf16* wv, int k, int Cin, int Cout, cudaStream_t st) {
    pack_stem_kernel<<<blocks_for((size_t)((k + 2) / 3) * Cout * 64), 256, 0, st>>>(w, wv, k, Cin, Cout);
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
//   stem: t[kh/3][((kh%3)*k+kw)*Cin+ci (64 rows)][co]    -> dw[kh][kw][ci][co]
//   head: t[kh/3][((kh%3)*k+kw)*Cout+co (64 rows)][ci]   -> dw[kh][kw][ci][co]
__global__ void unpack_dw_kernel(const float* __restrict__ t, float* __restrict__ dw, int k, int Cin, int Cout, int head) {
    const int total = k * k * Cin * Cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % Cout, ci = (i / Cout) % Cin, kw = (i / (Cout * Cin)) % k, kh = i / (Cout * Cin * k);
        const int tt = kh / 3, j = kh - 3 * tt;
        const float v = head ? t[((size_t)tt * 64 + (j * k + kw) * Cout + co) * Cin + ci] : t[((size_t)tt * 64 + (j * k + kw) * Cin + ci) * Cout + co];
        dw[i] += v;
    }
}
int sp_unpack_dw(const float* t, float* dw, int k, int Cin, int Cout, int head, cudaStream_t st) {
    unpack_dw_kernel<<<blocks_for((size_t)k * k * Cin * Cout), 256, 0, st>>>(t, dw, k, Cin, Cout, head);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// Thin-input strided conv (the discriminators' first layer, resnet.py:96: Conv k4 s2 'same' on a 3-channel image):
// k*k*Cin <= 64, so the whole receptive field of an output pixel is unfolded into ONE dense 64-channel chunk
//   U[n][oh][ow][(kh*k+kw)*Cin+ci] = x[n][oh*s + kh - pt][ow*s + kw - pl][ci]   (0 outside the image, 0 for the unused channels)
// and the conv is a 1x1 tensor-core GEMM over U (forward), a tap-stacked weight gradient with U as the X operand, and for
// the data gradient a 1x1 GEMM dU = dY . W^T followed by the adjoint of the unfolding (col2im).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_kernel(const bf16* __restrict__ x, bf16* __restrict__ U, int N, int H, int W, int Cin,
                                                     int Ho, int Wo, int k, int s, int pt, int pl) {
    __shared__ int lut[64];                              // (kh << 16) | (kw << 8) | ci, or -1 for an unused channel
    if (threadIdx.x < 64) {
        const int ch = threadIdx.x;
        int e = -1;
        if (ch < k * k * Cin) { const int tap = ch / Cin, ci = ch - tap * Cin, kh = tap / k, kw = tap - kh * k; e = (kh << 16) | (kw << 8) | ci; }
        lut[ch] = e;
    }
    __syncthreads();
    const size_t total = (size_t)N * Ho * Wo * 8;
    const bf16 zero = __float2bfloat16(0.f);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int v = (int)(i & 7);
        size_t px = i >> 3;
        const int ow = (int)(px % Wo);
        px /= Wo;
        const int oh = (int)(px % Ho), n = (int)(px / Ho);
        const bf16* img = x + (size_t)n * H * W * Cin;
        Pack<bf16, 8> pk;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int d = lut[v * 8 + e];
            bf16 val = zero;
            if (d >= 0) {
                const int ih = oh * s + (d >> 16) - pt, iw = ow * s + ((d >> 8) & 0xff) - pl;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W) val = img[((size_t)ih * W + iw) * Cin + (d & 0xff)];
            }
            pk.v[e] = val;
        }
        reinterpret_cast<uint4*>(U)[i] = *reinterpret_cast<uint4*>(&pk);
    }
}
int sp_im2col(const bf16* x, bf16* U, int N, int H, int W, int Cin, int Ho, int Wo, int k, int s, int pt, int pl, cudaStream_t st) {
    im2col_kernel<<<blocks_for((size_t)N * Ho * Wo * 8), 256, 0, st>>>(x, U, N, H, W, Cin, Ho, Wo, k, s, pt, pl);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// dx[n][ih][iw][ci] = sum over the taps (kh, kw) with (ih + pt - kh) and (iw + pl - kw) divisible by s and the quotient inside
// the output grid of dU[n][(ih+pt-kh)/s][(iw+pl-kw)/s][(kh*k+kw)*Cin+ci]
__global__ void col2im_kernel(const bf16* __restrict__ dU, bf16* __restrict__ dx, int N, int H, int W, int Cin, int Ho, int Wo,
                              int k, int s, int pt, int pl) {
    const size_t total = (size_t)N * H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int iw = (int)(i % W);
        size_t r = i / W;
        const int ih = (int)(r % H), n = (int)(r / H);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int kh = 0; kh < k; ++kh) {
            const int th = ih + pt - kh;
            if (th < 0 || th % s) continue;
            const int oh = th / s;
            if (oh >= Ho) continue;
            for (int kw = 0; kw < k; ++kw) {
                const int tw = iw + pl - kw;
                if (tw < 0 || tw % s) continue;
                const int ow = tw / s;
                if (ow >= Wo) continue;
                const bf16* p = dU + (((size_t)n * Ho + oh) * Wo + ow) * 64 + (kh * k + kw) * Cin;
                for (int c = 0; c < Cin; ++c) acc[c] += __bfloat162float(p[c]);
            }
        }
        for (int c = 0; c < Cin; ++c) dx[i * Cin + c] = __float2bfloat16(acc[c]);
    }
}
int sp_col2im(const bf16* dU, bf16* dx, int N, int H, int W, int Cin, int Ho, int Wo, int k, int s, int pt, int pl, cudaStream_t st) {
    col2im_kernel<<<blocks_for((size_t)N * H * W), 256, 0, st>>>(dU, dx, N, H, W, Cin, Ho, Wo, k, s, pt, pl);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// weights w[kh][kw][ci][co] (fp32 HWIO) -> wf[co][j] (64 columns, forward B operand) and wd[j][co] (64 rows, data-gradient
// B operand), j = (kh*k+kw)*Cin+ci, zero padded
__global__ void pack_im2col_kernel(const float* __restrict__ w, bf16* __restrict__ wf, bf16* __restrict__ wd, int k, int Cin, int Cout) {
    const int total = 64 * Cout, live = k * k * Cin;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % Cout, j = i / Cout;
        const bf16 v = __float2bfloat16(j < live ? w[(size_t)j * Cout + co] : 0.f);
        wd[(size_t)j * Cout + co] = v;
        wf[(size_t)co * 64 + j] = v;
    }
}
int sp_pack_im2col(const float* w, bf16* wf, bf16* wd, int k, int Cin, int Cout, cudaStream_t st) {
    pack_im2col_kernel<<<blocks_for((size_t)64 * Cout), 256, 0, st>>>(w, wf, wd, k, Cin, Cout);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
// dw[j][co] += t[j][co] for the live rows j < k*k*Cin (t: the first 64 rows of the tap-stacked weight-gradient result)
__global__ void unpack_im2col_kernel(const float* __restrict__ t, float* __restrict__ dw, int live, int Cout) {
    const int total = live * Cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) dw[i] += t[i];
}
int sp_unpack_im2col(const float* t, float* dw, int k, int Cin, int Cout, cudaStream_t st) {
    unpack_im2col_kernel<<<blocks_for((size_t)k * k * Cin * Cout), 256, 0, st>>>(t, dw, k * k * Cin, Cout);
    CG_LAUNCH_CHECK();
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// image-side layer of the U-Nets (unet.py:25 first double_conv conv, Cin = 3): the window conv (conv_tc.cu) needs a pixel
// stride of a multiple of 16 bytes, so the 3-channel input is re-laid as 8 channels (3 real + 5 zero) in the scratch
// ------------------------------------------------------------------------------------------
__global__ void pad_channels8_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, size_t npix, int cin) {
    const bf16 zero = __float2bfloat16(0.f);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        Pack<bf16, 8> pk;
#pragma unroll
        for (int c = 0; c < 8; ++c) pk.v[c] = c < cin ? x[i * cin + c] : zero;
        reinterpret_cast<uint4*>(y)[i] = *reinterpret_cast<uint4*>(&pk);
    }
}
int sp_pad_channels8(const bf16* x, bf16* y, size_t npix, int cin, cudaStream_t st) {
    pad_channels8_kernel<<<blocks_for(npix), 256, 0, st>>>(x, y, npix, cin);
    CG_LAUNCH_CHECK();
    return CG_OK;
}
