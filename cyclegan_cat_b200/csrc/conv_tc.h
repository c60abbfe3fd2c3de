// Host interface of the tcgen05 convolution kernels (conv_tc.cu).
#pragma once
#include <cuda.h>

#include "common.h"

struct TcConvArgs {
    int n_taps, cchunks, bn, n_blocks_n;        // K loop = taps x 64-channel chunks; N tile
    int tiles_per_img, tiles_w, Wb, Hb;         // M tiling: 128 output pixels = Wb x Hb box
    int n0, nb;                                 // images [n0, n0+nb)
    int out_P, out_wvalid, out_hvalid, out_H, out_W, Cout;   // epilogue: linear index -> (oh, ow), masks, dense output
    int b_rows_per_tap;                         // rows of the packed weight matrix per tap
    int stages;
    uint32_t idesc;
    short dw[49], dh[49];                       // TMA coordinate offsets per tap
};

struct TcWgradArgs {
    int n_taps, ci_blocks, co_blocks, bn, splits, stages;
    int n0, nb;
    int chunks_per_img, chunks_w, Wk, Hk;       // K chunk = 64 pixels = Wk x Hk box
    int dy_off;                                 // halo of the dY buffer
    int Cin, Cout;
    uint32_t idesc;
    short dw[49], dh[49];
};

int tc_make_map_4d(CUtensorMap* map, const void* base, int C, int W, int H, int N, int box_w, int box_h);
int tc_make_map_2d(CUtensorMap* map, const void* base, int cols, int rows, int box_rows);
int tc_pack_weights(const float* w, bf16* wf, bf16* wd, int taps, int Cin, int Cout, cudaStream_t st);
int tc_conv_launch(const CUtensorMap* mapA, const CUtensorMap* mapB, bf16* out, const float* bias, TcConvArgs a,
                   double flops, cudaStream_t st);
int tc_wgrad_launch(const CUtensorMap* mapX, const CUtensorMap* mapDY, float* dw, TcWgradArgs a, double flops,
                    cudaStream_t st);
