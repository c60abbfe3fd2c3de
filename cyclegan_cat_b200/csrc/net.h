// Layer-graph executor: plans buffers for one (N,H,W) call of a net and launches the kernels
// for forward and for the hand-scheduled backward.  No allocation: every buffer is a slice of
// caller-owned workspace.
#pragma once
#include <vector>

#include "common.h"

#include "conv_tc.h"

// tensor-core plan of one eligible conv layer for one planned call
struct TcConvLaunch { CUtensorMap mapA, mapB, mapB2; TcConvArgs a; double flop_share = 1.0; bool halo = false; /* convw_tc_kernel */ };
struct TcLayer {
    bool on = false;
    std::vector<TcConvLaunch> fwd;      // 1 launch (Conv2D) or 4 stride-parity classes (Conv2DTranspose)
    std::vector<TcConvLaunch> dgrad;    // 1 launch (stride 1: flat mode; Conv2DTranspose: a stride-2 conv of dY) or 4 classes
    CUtensorMap mapXw, mapDYw;
    TcWgradArgs wa;
    TcWgrad16Args wa16;
    TcWgradWArgs waw;
    bool wg16 = false;
    bool wgh = false;                   // ... in its halo form (wgradh_tc_kernel)
    bool wgw = false;                   // TC_S1_WIN: the window weight-gradient kernel applies (64-pixel chunks tile the grid)
    int wg_x_is_dy = 0;                 // Conv2DTranspose: the "X" operand of the weight gradient is dY
    size_t sc_tmp = 0;                  // TC_STEM / TC_HEAD: offset of the fp32 weight-gradient staging buffer in the scratch
};
enum { TC_NONE = 0, TC_S1_VALID = 1, TC_CONV_S2 = 2, TC_CONVT_S2 = 3, TC_STEM = 4, TC_HEAD = 5, TC_S1_16 = 6, TC_IM2COL = 7, TC_S1_WIN = 8 };

struct LayerInfo {
    cg_layer_desc d;
    long long w_off = -1, b_off = -1, g_off = -1, be_off = -1;   // offsets (floats) into the flat parameter buffer
    int out_t = 0;          // tensor id this layer writes (i+1, or i+2 when the following ACT is fused in)
    bool skipped = false;   // ACT folded into the preceding INORM
    int fused_act = CG_ACT_NONE;
    float fused_slope = 0.f;
    int fuse_rpad = -1;            // INORM whose only consumer is a reflection pad: index of that RPAD layer (the norm's
                                   // forward writes the padded tensor directly when the streaming kernel applies);
                                   // ADD: index of a reflection pad among its consumers (written as a second output)
    int fuse_add = -1;             // INORM whose only consumer is a residual ADD: index of that ADD layer
    bool feeds_in = false;         // the conv output feeds ONLY an instance norm (statistics fused into the conv epilogue)
    bool bias_grad_zero = false;   // the conv output feeds ONLY an instance norm: d(loss)/d(bias) == 0 exactly
    bool batch = false;            // INORM that is a BatchNormalization: statistics pooled over the samples of one call
    long long mm_off = -1, mv_off = -1;   // BatchNormalization: offsets (floats) of moving_mean / moving_variance in the state
    // zero-copy concat (unet.py:109 `Concatenate()([skip, x])`): AVGPOOL whose input is in0 of concat layer `cat_layer` (the
    // pool also writes the skip copy / its backward adds the skip slice), UPSAMPLE whose only consumer is that concat as in1
    // (it writes into / gathers from its slice); on the CONCAT layer: the pool / upsample layer serving each input, or -1
    int cat_layer = -1, cat_in0_pool = -1, cat_in1_up = -1;
    int drop_index = -1;           // DROPOUT: ordinal among the net's dropout layers (part of the mask key)
    int tc = 0;             // TC_* kind: which convs run on the tcgen05 kernels in bf16 mode
    long long pk_f = 0, pk_d = 0;   // byte offsets of the packed bf16 weights ([tap][Cout][Cin] / [tap][Cin][Cout])
};

struct cg_net_s {
    int mode = CG_MODE_BF16;
    std::vector<LayerInfo> layers;
    std::vector<int> chan;              // channels per tensor id
    std::vector<int> n_consumers;       // per tensor id
    std::vector<char> has_buffer;       // per tensor id
    std::vector<char> dep_params;       // tensor depends on some trainable variable
    std::vector<cg_var_info> vars;
    long long n_params = 0;
    long long n_state = 0;              // non-trainable floats (BatchNormalization moving statistics)
    float* state = nullptr;             // caller-owned device buffer of n_state floats (cg_net_bind_state)
    int training = 0;                   // Keras `training=` of the next single-net forward (cg_net_set_training)
    unsigned long long seed = 0;        // dropout stream (cg_net_set_seed)
    unsigned long long seed_epoch = 0;  // bumped by every cg_net_set_seed: a trainer whose captured graphs baked the old seed in re-captures
    unsigned long long calls = 0;       // training-mode single-net forwards so far (dropout counter)
    size_t packed_bytes = 0;            // bf16 weight copies for the tensor-core layers
    int out_tensor() const { return (int)layers.size(); }
    size_t elem_size() const { return mode == CG_MODE_BF16 ? 2 : 4; }
};

// One planned call: shapes and workspace offsets for a given (N,H,W)
struct CallCtx {
    const cg_net_s* net = nullptr;
    int N = 0, H = 0, W = 0;
    bool bwd = false;
    std::vector<int> th, tw;            // spatial size per tensor
    std::vector<size_t> act_off;        // byte offsets into `base` (activations)
    std::vector<size_t> stat_off;       // per layer: INORM statistics [N][C][2] float: (mean, rstd) once finalized
    std::vector<size_t> raw_off;        // per layer (plain instance norm): the raw (sum x, sum x^2) table a conv epilogue fills;
                                        // the streaming forward apply turns it into stat_off's (mean, rstd) on the fly
    size_t act_bytes = 0;
    size_t stat_begin = 0;              // the INORM statistics tables are one contiguous region [stat_begin, act_bytes)
    std::vector<size_t> grad_off;       // byte offsets into the shared gradient arena
    size_t grad_bytes = 0;              // includes the IN-backward scratch at scratch_off
    size_t scratch_off = 0;             // IN-backward sums: one [N][C][2] float table per INORM layer from here on,
    std::vector<size_t> sums_off;       // per layer (offset into the arena); the region is zeroed once per backward
    size_t sums_bytes = 0;
    char* base = nullptr;               // activations workspace
    char* arena = nullptr;              // gradient arena (may be shared between calls)
    void* ext_input = nullptr;          // if set, tensor 0 lives here instead of at act_off[0]
    int dx_nb = 0;                      // > 0: the caller reads dLoss/d(input) of the first dx_nb samples only (the layers that
                                        // have no parameter upstream may skip the data gradient of the rest)
    bool forwarded = false;
    char* packed = nullptr;             // packed bf16 weights of this net (net_pack)
    char* tcs = nullptr;                // scratch for the unfolded tensors of the 7x7 stem / head (conv_special.cu)
    size_t tcs_bytes = 0;
    std::vector<TcLayer> tc;            // per layer
    // BatchNormalization / Dropout: the batch of this call is `N / bn_group` Keras calls of bn_group samples each
    int bn_group = 0;                   // 0 = the whole batch is one call
    bool training = false;              // batch statistics + moving-average update, active dropout
    bool defer_moving = false;          // the caller applies the moving-average updates itself (net_update_moving), in the
                                        // reference's call order
    std::vector<size_t> bstat_off;      // per layer: (mean, unbiased variance) per group of a BatchNormalization [G][C][2]
    unsigned long long drop_ctr_host = 0;
    const unsigned long long* drop_ctr_dev = nullptr;
    int call_id[4] = {0, 1, 2, 3};      // dropout: id of each group's Keras call within the step
    std::vector<int> grad_halo;         // per tensor: zero border of the gradient buffer (tensor-core layers)
    std::vector<char> live;             // per tensor: the last forward materialised it (fused layers skip their own output)

    size_t sample_elems(int t) const { return (size_t)th[t] * tw[t] * net->chan[t]; }
    void* act(int t) const { return (t == 0 && ext_input) ? ext_input : (void*)(base + act_off[t]); }
};

int net_plan(const cg_net_s* net, int N, int H, int W, bool bwd, CallCtx* ctx);
int net_bind(CallCtx* ctx);           // after base/arena/packed are set: build the TMA descriptors
int net_pack(const cg_net_s* net, const float* params, void* packed, cudaStream_t st);
int net_out_hw(const cg_net_s* net, int H, int W, int* ho, int* wo);

// forward over the whole planned batch; output is ctx->act(out_tensor())
int net_forward(CallCtx* ctx, const float* params, cudaStream_t st);
// backward over samples [n0, n0+nb): dy is dLoss/d(output) for those samples (activation dtype);
// dx (nullable) receives dLoss/d(input); parameter gradients are ACCUMULATED into grads when non-null.
// `hook` (nullable) is called after the backward launches of each layer, last layer first: every parameter gradient of
// the layers >= that index is then complete on `st` (the data-parallel trainer all-reduces them bucket by bucket).
typedef int (*LayerHook)(void* user, int layer);
int net_backward(CallCtx* ctx, const float* params, const void* dy, void* dx, float* grads, int n0, int nb,
                 cudaStream_t st, LayerHook hook = nullptr, void* hook_user = nullptr);
// BatchNormalization: fold group g's batch statistics of the last training forward into the moving averages
int net_update_moving(CallCtx* ctx, int g, cudaStream_t st);
