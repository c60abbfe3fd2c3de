#pragma once
#include "common.h"
int sp_unfold_w(const bf16* src, bf16* dst, int N, int Hs, int Ws, int Cs, int Hd, int Wd, int k, int sign, cudaStream_t st);
int sp_pack_stem(const float* w, bf16* wv, int k, int Cin, int Cout, cudaStream_t st);
int sp_pack_head(const float* w, bf16* wh, bf16* whd, int k, int Cin, int Cout, cudaStream_t st);
int sp_diag_sum(const bf16* S, const float* bias, bf16* y, int N, int R, int Wy, int Ws, int k, int C, int sign, cudaStream_t st);
int sp_pack_stem_d(const float* w, bf16* wsd, int k, int Cin, int Cout, cudaStream_t st);
int sp_unpack_dw(const float* t, float* dw, int k, int Cin, int Cout, int head, cudaStream_t st);
// thin-input strided conv (k*k*Cin <= 64) through a dense 64-channel unfolding (im2col) and its adjoint
int sp_im2col(const bf16* x, bf16* U, int N, int H, int W, int Cin, int Ho, int Wo, int k, int s, int pt, int pl, cudaStream_t st);
int sp_col2im(const bf16* dU, bf16* dx, int N, int H, int W, int Cin, int Ho, int Wo, int k, int s, int pt, int pl, cudaStream_t st);
int sp_pack_im2col(const float* w, bf16* wf, bf16* wd, int k, int Cin, int Cout, cudaStream_t st);
int sp_unpack_im2col(const float* t, float* dw, int k, int Cin, int Cout, cudaStream_t st);
int sp_pad_channels8(const bf16* x, bf16* y, size_t npix, int cin, cudaStream_t st);
