// Host interface of the tcgen05 convolution kernels (conv_tc.cu).
#pragma once
#include <cuda.h>

#include "common.h"

enum { TC_MAX_STEPS = 64 };
struct TcConvArgs {
    int n_taps, cchunks, bn, n_blocks_n;        // K loop = taps x 64-channel chunks; N tile
    int tiles_per_img, tiles_w, Wb, Hb;         // M tiling: 128 output pixels = Wb x Hb box
    int n0, nb;                                 // images [n0, n0+nb)
    int out_P, out_wvalid, out_hvalid, out_H, out_W, Cout;   // epilogue: linear index -> (oh, ow), masks, dense output
    int out_sy, out_oy, out_sx, out_ox;         // output pixel = (oh*sy + oy, ow*sx + ox): stride-2 scatter of a parity class
    int b_rows_per_tap;                         // rows of the packed weight matrix per tap
    int stages;
    uint32_t idesc;
    int bk16, groups, cin16;                    // 16-channel K groups (SWIZZLE_32B): `groups` per K step, cin16 = Cin/16 in total
    int dbg;                                    // measurement only (CG_TC_DBG): 1 = epilogue reads TMEM and stores nothing, 2 = no fused statistics
    int tps;                                    // bk16, Cin <= 64: filter taps per K step (each its own A box, ONE B box)
    float* stats;                               // nullable: [N][Cout][2] running (sum, sum of squares) of the outputs
    // fold mode (data gradient that feeds the adjoint of a reflection pad): output pixels inside the pad border go straight
    // to the UNPADDED gradient `out2` [nb][out_H-2p][out_W-2p][Cout] (added to its content when fold_acc), only the border
    // pixels go to `out`; rpad_bwd_border then folds those few pixels
    bf16* out2;
    int fold_pad, fold_acc;
    // window mode (conv_tc_kernel<.., WIN>): the A operand is the "window view" of an NHWC tensor with C % 8 == 0 -- row p of
    // a K step is the 128 contiguous bytes that start at element dc[step] of the k*C-element window [pixel p-pl .. p-pl+k-1]
    // of image row h0+dh[step] (tc_make_map_win).  Window pixels outside [0, W) are zeroed in shared memory before the MMAs.
    int win_C, win_k, win_pl, win_W;
    int win_pt;                                 // convw_tc_kernel: rows above the tile that the halo starts at (top padding)
    // per K-loop tap: TMA coordinate offsets into the 5-D activation view (c, w, p, h, n) and the weight row block
    short dc[TC_MAX_STEPS], dw[TC_MAX_STEPS], dp[TC_MAX_STEPS], dh[TC_MAX_STEPS], tb[TC_MAX_STEPS];
};

// weight gradient in window mode (wgradw_tc_kernel): M = two 64-element window chunks (K steps 2u, 2u+1 of the forward
// conv), N = Cout (16-channel groups of dY, SWIZZLE_32B), K = 64 pixels per stage, split over the pixel range
struct TcWgradWArgs {
    int steps, nch, k, C, Creal, Cout, units, splits, stages;
    int halves;                                 // window chunks per unit: 2 (one accumulator) or 4 (two, Cout <= 128)
    int n0, nb, y_n0;
    int chunks_per_img, chunks_w, Wk, Hk;
    int W, pl;
    int pt, acc_pitch, tmem_cols;               // wgradh_tc_kernel: top padding, TMEM column pitch of the accumulators, columns allocated
    uint32_t idesc;
    short dc[TC_MAX_STEPS], dh[TC_MAX_STEPS];
};
int tc_wgradw_launch(const CUtensorMap* mapX, const CUtensorMap* mapDY, float* dw, TcWgradWArgs a, double flops, cudaStream_t st);
// halo form (wgradh_tc_kernel): 16 x 4 pixel blocks, one unit per window chunk
int tc_wgradh_ok(int k, int Cout);
int tc_wgradh_launch(const CUtensorMap* mapX, const CUtensorMap* mapDY, float* dw, TcWgradWArgs a, double flops, cudaStream_t st);
// window view of an NHWC bf16 tensor (C % 8 == 0): dims (k*C, W, 1, H, N), pixel stride C (overlapping rows), base shifted
// left by `pl` pixels; box {64, box_w, 1, box_h, 1}, SWIZZLE_128B; elements beyond k*C are zero-filled by TMA
int tc_make_map_win(CUtensorMap* map, const void* x, int C, int k, int pl, int W, int H, int N, int box_w, int box_h);
// packed weights of the window form: wf[step][npad][64] with step = (kh, chunk j) and element e = 64j+i <-> (kw = e / C,
// ci = e % C); flip = data-gradient orientation (w[k-1-kh][k-1-kw], rows = input channels); rows >= n_rows and elements
// >= k*C (and channels >= Creal) are zero
// batched: tc_pack_win queues a job (flushing when the table is full), tc_pack_win_flush launches what is queued
struct TcPackWinJobs {
    const float* w[32]; bf16* wf[32];
    int k[32], C[32], Creal[32], n_rows[32], npad[32], cin_w[32], cout_w[32], flip[32], block0[33];
    int n;
};
int tc_pack_win(TcPackWinJobs& jobs, const float* w, bf16* wf, int k, int C, int Creal, int n_rows, int npad, int Cin_w,
                int Cout_w, int flip, cudaStream_t st);
int tc_pack_win_flush(TcPackWinJobs& jobs, cudaStream_t st);
// window conv with vertical halo reuse (convw_tc_kernel): stages that fit / launch
int tc_convw_stages(int bn, int k, int Wb, int Hb, bool* dual);
int tc_convw_launch(const CUtensorMap* mapA, const CUtensorMap* mapB, bf16* out, const float* bias, TcConvArgs a, double flops,
                    cudaStream_t st);

struct TcWgradArgs {
    int n_taps, a_blocks, b_blocks, bn, splits, stages;   // M = 128-row blocks of operand A, N = bn-column blocks of B
    int x_grouped, y_grouped;                   // operand map is the channel-grouped 5-D view: ONE TMA box loads all groups
    int transposed;                             // 0: A = X (rows = ci), B = dY (cols = co);  1: A = dY (rows = co), B = X
    int stack2;                                 // X has 64 channels: A rows 0-63 = tap 2u, rows 64-127 = tap 2u+1 (n_taps = units)
    int n0, nb, y_n0;                           // image offsets: X operand starts at n0, dY operand at y_n0
    int chunks_per_img, chunks_w, Wk, Hk;       // K chunk = 64 pixels = Wk x Hk box of the dY grid
    int dy_off;                                 // halo of the dY buffer
    int Cin, Cout;
    uint32_t idesc;
    short dc[TC_MAX_STEPS], dw[TC_MAX_STEPS], dp[TC_MAX_STEPS], dh[TC_MAX_STEPS];       // X coordinate offsets per tap (5-D view)
};

// 5-D activation view (c, w, p, h, n).  parity = 0: dense NHWC tensor, p is a dummy dim of size 1.
// parity = 1: stride-2 view of an NHWC tensor with even H, W: c' = pw*C + c (size 2C), w' = w/2, p = h%2, h' = h/2.
// weight gradient of stride-1 convs whose channel counts are multiples of 16 only (tap-stacked M, wgrad16_tc_kernel)
struct TcWgrad16Args {
    int n_taps, Cin, Cout, m_blocks, splits, stages;
    int n0, nb;                                 // X images start at n0, dY images at y_n0
    int y_n0;
    int chunks_per_img, chunks_w, Wk, Hk;       // chunks_w may round up: out-of-range pixels are zero in both operands
    uint32_t idesc;
    short dw[TC_MAX_STEPS], dh[TC_MAX_STEPS];
};
int tc_wgrad16_launch(const CUtensorMap* mapX, const CUtensorMap* mapDY, float* dw, TcWgrad16Args a, double flops,
                      cudaStream_t st);
int tc_make_map_act(CUtensorMap* map, const void* base, int C, int W, int H, int N, int parity, int box_w, int box_h);
// channel-grouped dense view (c%64, w, h, c/64, n): one box {64, box_w, groups, box_h, 1} lands as `groups` consecutive
// [box_h*box_w][64] tiles -- the MN-major operand layout of wgrad_tc_kernel -- with a single TMA instruction
int tc_make_map_act_grouped(CUtensorMap* map, const void* base, int C, int W, int H, int N, int box_w, int box_h, int groups);
int tc_make_map_2d(CUtensorMap* map, const void* base, int cols, int rows, int box_rows);
// 16-channel-group views (SWIZZLE_32B) for layers whose channel counts are multiples of 16 but not of 64 (the U-Nets):
//   activation (c%16, w, h, c/16, n), box {16, box_w, box_h, groups, 1}  ->  smem [group][box_h*box_w][16]
//   weights    (k%16, row, k/16),     box {16, box_rows, groups}         ->  smem [group][box_rows][16]
int tc_make_map_act16(CUtensorMap* map, const void* base, int C, int W, int H, int N, int box_w, int box_h, int groups);
int tc_make_map_w16(CUtensorMap* map, const void* base, int cols, int rows, int box_rows, int groups);
int tc_pack_weights(const float* w, bf16* wf, bf16* wd, int taps, int Cin, int Cout, cudaStream_t st);
// batched form: fill jobs (n <= TC_PACK_MAX), then one launch; tc_pack_weights_multi resets jobs.n
enum { TC_PACK_MAX = 32 };
struct TcPackJobs {
    const float* w[TC_PACK_MAX]; bf16* wf[TC_PACK_MAX]; bf16* wd[TC_PACK_MAX];
    int taps[TC_PACK_MAX], cin[TC_PACK_MAX], cout[TC_PACK_MAX], block0[TC_PACK_MAX + 1];
    int n;
};
int tc_pack_weights_multi(TcPackJobs& jobs, cudaStream_t st);
// mapB2 (nullable): the same weight matrix with a box of bn/2 rows, for the 2-CTA kernel
int tc_conv_launch(const CUtensorMap* mapA, const CUtensorMap* mapB, const CUtensorMap* mapB2, bf16* out, const float* bias,
                   TcConvArgs a, double flops, cudaStream_t st);
int tc_wgrad_launch(const CUtensorMap* mapX, const CUtensorMap* mapDY, float* dw, TcWgradArgs a, double flops,
                    cudaStream_t st);
