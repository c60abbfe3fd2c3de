// C ABI of libcyclegan_b200.so: library init, model builder, single-net forward/backward.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "kernels.h"
#include "net.h"

std::atomic<long long> g_launches{0};
static thread_local char g_err[1024] = "";

void cg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool cg_pdl_enabled() {
    static const bool on = [] { const char* e = getenv("CG_DISABLE_PDL"); return !(e && e[0] == '1'); }();
    return on;
}

extern "C" const char* cg_last_error(void) { return g_err; }
extern "C" int cg_version(void) { return 110; }
extern "C" int cg_abi_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(cg_layer_desc);
        case 1: return (int)sizeof(cg_var_info);
        case 2: return (int)sizeof(cg_train_cfg);
        case 3: return (int)sizeof(cg_adam_cfg);
        default: return -1;
    }
}

extern "C" int cg_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cg_set_error("no CUDA device: %s (libcyclegan_b200 has no CPU fallback)", cudaGetErrorString(e));
        return CG_ERR_CUDA;
    }
    if (device < 0 || device >= n) { cg_set_error("device %d out of range (%d devices)", device, n); return CG_ERR_INVALID; }
    cudaDeviceProp p;
    CG_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        cg_set_error("device %d is sm_%d%d; this library contains sm_100a code only", device, p.major, p.minor);
        return CG_ERR_INVALID;
    }
    CG_CUDA(cudaSetDevice(device));
    return CG_OK;
}

extern "C" int cg_launch_count(int64_t* launches, int reset) {
    if (launches) *launches = g_launches.load();
    if (reset) g_launches.store(0);
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// model builder
// ------------------------------------------------------------------------------------------
extern "C" int cg_net_create(const cg_layer_desc* layers, int n_layers, int mode, cg_net_t* out) {
    if (!layers || n_layers <= 0 || !out) { cg_set_error("cg_net_create: null/empty layer list"); return CG_ERR_INVALID; }
    if (mode != CG_MODE_BF16 && mode != CG_MODE_FP32_CHECK) { cg_set_error("unknown mode %d", mode); return CG_ERR_INVALID; }
    cg_net_s* net = new cg_net_s();
    net->mode = mode;
    net->layers.resize(n_layers);
    net->chan.assign(n_layers + 1, 0);
    net->n_consumers.assign(n_layers + 1, 0);
    net->has_buffer.assign(n_layers + 1, 1);
    net->dep_params.assign(n_layers + 1, 0);
    auto fail = [&](const char* what, int i) {
        cg_set_error("cg_net_create: layer %d: %s", i, what);
        delete net;
        return CG_ERR_INVALID;
    };
    // input channels = cin of the first consumer of tensor 0
    for (int i = 0; i < n_layers; ++i)
        if (layers[i].in0 == 0) { net->chan[0] = layers[i].cin; break; }
    if (net->chan[0] <= 0) return fail("tensor 0 is never consumed", 0);
    long long off = 0, soff = 0;
    int n_drop = 0;
    for (int i = 0; i < n_layers; ++i) {
        LayerInfo& L = net->layers[i];
        L.d = layers[i];
        L.out_t = i + 1;
        const cg_layer_desc& d = L.d;
        if (d.in0 < 0 || d.in0 > i) return fail("in0 is not an earlier tensor", i);
        const bool two = d.op == CG_OP_ADD || d.op == CG_OP_CONCAT;
        if (two && (d.in1 < 0 || d.in1 > i)) return fail("in1 is not an earlier tensor", i);
        if (net->chan[d.in0] != d.cin) return fail("cin does not match the producer's channels", i);
        net->n_consumers[d.in0]++;
        if (two) net->n_consumers[d.in1]++;
        bool params_here = false;
        auto add_var = [&](int role, int ndim, int s0, int s1, int s2, int s3) {
            cg_var_info v;
            v.layer = i; v.role = role; v.ndim = ndim;
            v.shape[0] = s0; v.shape[1] = s1; v.shape[2] = s2; v.shape[3] = s3;
            v.offset = off;
            long long n = 1;
            for (int q = 0; q < ndim; ++q) n *= v.shape[q];
            off += n;
            net->vars.push_back(v);
            params_here = true;
            return v.offset;
        };
        switch (d.op) {
            case CG_OP_CONV:
            case CG_OP_CONVT:
                if (d.k < 1 || d.k > 15 || (d.stride != 1 && d.stride != 2) || d.cout < 1) return fail("bad conv geometry", i);
                if (d.op == CG_OP_CONVT && !d.same) return fail("Conv2DTranspose supports padding='same' only", i);
                net->chan[i + 1] = d.cout;
                L.w_off = d.op == CG_OP_CONV ? add_var(0, 4, d.k, d.k, d.cin, d.cout) : add_var(0, 4, d.k, d.k, d.cout, d.cin);
                if (d.has_bias) L.b_off = add_var(1, 1, d.cout, 0, 0, 0);
                break;
            case CG_OP_INORM:
                net->chan[i + 1] = d.cin;
                if (d.affine) { L.g_off = add_var(2, 1, d.cin, 0, 0, 0); L.be_off = add_var(3, 1, d.cin, 0, 0, 0); }
                break;
            case CG_OP_BNORM:
                // BatchNormalization runs on the instance-norm kernels with its statistics tables pooled over the samples
                // of a call; from here on the layer is an INORM with `batch` set.  Trainable: [gamma, beta] when affine
                // (Keras center/scale); non-trainable: [moving_mean, moving_variance].
                if (!(d.momentum >= 0.f && d.momentum < 1.f)) return fail("BatchNormalization momentum must be in [0, 1)", i);
                L.d.op = CG_OP_INORM;
                L.batch = true;
                net->chan[i + 1] = d.cin;
                if (d.affine) { L.g_off = add_var(2, 1, d.cin, 0, 0, 0); L.be_off = add_var(3, 1, d.cin, 0, 0, 0); }
                L.mm_off = soff; soff += d.cin;
                L.mv_off = soff; soff += d.cin;
                break;
            case CG_OP_DROPOUT:
                if (!(d.rate >= 0.f && d.rate < 1.f)) return fail("Dropout rate must be in [0, 1)", i);
                net->chan[i + 1] = d.cin;
                L.drop_index = n_drop++;
                break;
            case CG_OP_ACT:
                if (d.act < CG_ACT_RELU || d.act > CG_ACT_SIGMOID) return fail("unknown activation", i);
                net->chan[i + 1] = d.cin;
                break;
            case CG_OP_RPAD:
                if (d.pad < 1) return fail("reflect pad must be >= 1", i);
                net->chan[i + 1] = d.cin;
                break;
            case CG_OP_ADD:
                if (net->chan[d.in1] != d.cin) return fail("Add of different channel counts", i);
                net->chan[i + 1] = d.cin;
                break;
            case CG_OP_CONCAT:
                net->chan[i + 1] = d.cin + net->chan[d.in1];
                if (d.cout != net->chan[i + 1]) return fail("cout != sum of concat inputs", i);
                break;
            case CG_OP_AVGPOOL:
            case CG_OP_UPSAMPLE: net->chan[i + 1] = d.cin; break;
            default: return fail("unknown op", i);
        }
        net->dep_params[i + 1] = params_here || net->dep_params[d.in0] || (two && net->dep_params[d.in1]);
    }
    net->n_params = off;
    net->n_state = soff;
    // fold ReLU / LeakyReLU into the instance norm that feeds only them
    for (int i = 0; i + 1 < n_layers; ++i) {
        LayerInfo& L = net->layers[i];
        LayerInfo& A = net->layers[i + 1];
        if (L.d.op == CG_OP_INORM && A.d.op == CG_OP_ACT && A.d.in0 == i + 1 && net->n_consumers[i + 1] == 1 &&
            (A.d.act == CG_ACT_RELU || A.d.act == CG_ACT_LEAKY)) {
            L.out_t = i + 2;
            L.fused_act = A.d.act;
            L.fused_slope = A.d.slope;
            A.skipped = true;
            net->has_buffer[i + 1] = 0;
        }
    }
    // forward fusions of the streaming instance-norm kernel (kernels_stream.cu):
    //   INORM (+ folded activation) -> RPAD only          : the norm writes the reflection-padded tensor itself
    //   INORM -> ADD only (the residual sum, resnet.py:34) : the norm adds the skip tensor and writes the sum, and when a
    //                                                        reflection pad consumes the sum, its padded copy as well
    auto uses = [&](const LayerInfo& R, int t) {
        return R.d.in0 == t || ((R.d.op == CG_OP_ADD || R.d.op == CG_OP_CONCAT) && R.d.in1 == t);
    };
    for (int i = 0; i < n_layers; ++i) {
        LayerInfo& L = net->layers[i];
        if (L.skipped || L.out_t >= n_layers) continue;
        if (L.d.op == CG_OP_INORM && net->n_consumers[L.out_t] == 1) {
            for (int j = i + 1; j < n_layers; ++j) {
                const LayerInfo& R = net->layers[j];
                if (R.skipped || !uses(R, L.out_t)) continue;
                if (R.d.op == CG_OP_RPAD && R.d.pad > 0) L.fuse_rpad = j;
                if (R.d.op == CG_OP_ADD && R.d.in0 != R.d.in1) L.fuse_add = j;
                break;
            }
        } else if (L.d.op == CG_OP_ADD) {
            for (int j = i + 1; j < n_layers; ++j) {
                const LayerInfo& R = net->layers[j];
                if (!R.skipped && R.d.op == CG_OP_RPAD && R.d.in0 == L.out_t && R.d.pad > 0) { L.fuse_rpad = j; break; }
            }
        }
    }
    // zero-copy concat: see LayerInfo::cat_layer (CG_DISABLE_CATFUSE=1: plain slice copies, the A/B and test hook)
    {
        const char* off = getenv("CG_DISABLE_CATFUSE");
        if (!(off && off[0] == '1'))
            for (int j = 0; j < n_layers; ++j) {
                LayerInfo& Cc = net->layers[j];
                if (Cc.skipped || Cc.d.op != CG_OP_CONCAT || Cc.d.in0 == Cc.d.in1) continue;
                for (int i = 0; i < j; ++i) {
                    LayerInfo& L = net->layers[i];
                    if (L.skipped || L.cat_layer >= 0) continue;
                    if (L.d.op == CG_OP_AVGPOOL && L.d.in0 == Cc.d.in0 && Cc.cat_in0_pool < 0) { L.cat_layer = j; Cc.cat_in0_pool = i; }
                    if (L.d.op == CG_OP_UPSAMPLE && L.out_t == Cc.d.in1 && net->n_consumers[L.out_t] == 1 && Cc.cat_in1_up < 0) {
                        L.cat_layer = j; Cc.cat_in1_up = i;
                    }
                }
            }
    }
    // A conv / transposed-conv bias that feeds ONLY an instance norm has an identically zero gradient: the norm subtracts
    // the per-(sample, channel) mean, so sum_pixels dL/dy == 0 ('zero-by-construction gradients').  The
    // reference computes rounding noise there; this library writes the exact value 0 and skips the reduction.
    for (int i = 0; i + 1 < n_layers; ++i) {
        LayerInfo& L = net->layers[i];
        if ((L.d.op == CG_OP_CONV || L.d.op == CG_OP_CONVT) && net->n_consumers[i + 1] == 1 &&
            net->layers[i + 1].d.op == CG_OP_INORM && net->layers[i + 1].d.in0 == i + 1) {
            L.feeds_in = true;
            L.bias_grad_zero = L.d.has_bias != 0;
        }
        // the same holds through a channel concat (strided_unet, unet.py:66-70: Conv2DTranspose -> Concatenate -> norm over
        // the concatenated tensor): the norm is per channel, so the conv's slice still has a zero-sum gradient
        if ((L.d.op == CG_OP_CONV || L.d.op == CG_OP_CONVT) && L.d.has_bias && net->n_consumers[i + 1] == 1) {
            for (int j = i + 1; j + 1 < n_layers; ++j) {
                const cg_layer_desc& q = net->layers[j].d;
                if (q.op != CG_OP_CONCAT || (q.in0 != i + 1 && q.in1 != i + 1)) continue;
                if (q.in0 != q.in1 && net->n_consumers[j + 1] == 1 && net->layers[j + 1].d.op == CG_OP_INORM &&
                    net->layers[j + 1].d.in0 == j + 1)
                    L.bias_grad_zero = true;
                break;
            }
        }
    }
    // tensor-core layers (bf16 mode): 3x3 stride-1 'valid' convs with Cin % 128 == 0 and Cout % 64 == 0 whose output
    // feeds exactly one instance norm (its backward writes the zero-bordered dY the TMA loads expect)
    size_t pk = 0;
    const char* no_tc = getenv("CG_DISABLE_TC");     // test hook: force the CUDA-core convs in bf16 mode
    const char* no_win = getenv("CG_DISABLE_WIN");   // A/B hook: the 16-channel-group kernels instead of the window form
    const bool win_on = !(no_win && no_win[0] == '1');
    if (mode == CG_MODE_BF16 && !(no_tc && no_tc[0] == '1')) {
        auto chan_ok = [](int cin_f, int cout_f) {      // K chunks of 64, N tiles of <= 256, an M = 128 side for wgrad
            return cin_f % 64 == 0 && cout_f % 64 == 0 && (cin_f <= 256 || cin_f % 256 == 0) &&
                   (cout_f <= 256 || cout_f % 256 == 0) && (cin_f % 128 == 0 || cout_f % 128 == 0);
        };
        for (int i = 0; i < n_layers; ++i) {
            LayerInfo& L = net->layers[i];
            const cg_layer_desc& d = L.d;
            int kind = TC_NONE;
            if (d.op == CG_OP_CONV && d.k == 3 && d.stride == 1 && !d.same && d.cin % 128 == 0 && chan_ok(d.cin, d.cout) &&
                i + 1 < n_layers && net->n_consumers[i + 1] == 1 && net->layers[i + 1].d.op == CG_OP_INORM &&
                net->layers[i + 1].d.in0 == i + 1)
                kind = TC_S1_VALID;
            else if (d.op == CG_OP_CONV && d.stride == 2 && d.same && (d.k == 3 || d.k == 4) && chan_ok(d.cin, d.cout))
                kind = TC_CONV_S2;
            else if (d.op == CG_OP_CONVT && d.stride == 2 && (d.k == 3 || d.k == 4) && chan_ok(d.cout, d.cin))
                kind = TC_CONVT_S2;
            else if (d.op == CG_OP_CONV && d.stride == 2 && d.same && d.cin <= 4 && d.k * d.k * d.cin <= 64 && d.k <= 8 &&
                     d.cout % 64 == 0 && d.cout <= 256)
                kind = TC_IM2COL;       // discriminator input layer (resnet.py:96, Conv k4 s2 on the image): full im2col into 64 channels
            else if (d.op == CG_OP_CONV && d.stride == 1 && !d.same && d.k >= 3 && d.cin <= 4 && d.k * d.cin <= 21 && d.k <= 12 &&
                     d.cout % 64 == 0 && d.cout <= 256)
                kind = TC_STEM;         // c7s1-f stem (resnet.py:39-40): horizontal taps unfolded into channels
            else if (d.op == CG_OP_CONV && d.stride == 1 && !d.same && d.k >= 3 && d.cout <= 4 && d.k * d.cout <= 21 && d.k <= 12 &&
                     d.cin % 64 == 0 && d.cin <= 256)
                kind = TC_HEAD;         // c7s1-3 tanh head (resnet.py:82)
            else if (win_on && d.op == CG_OP_CONV && d.stride == 1 && (d.same || d.k == 1) && d.k <= 7 &&
                     ((d.cin % 8 == 0 && d.cin <= 2048) || d.cin <= 4) && d.cout % 16 == 0 && d.cout <= 256 &&
                     d.k * ((d.k * (d.cin <= 4 ? 8 : d.cin) + 63) / 64) <= TC_MAX_STEPS)
                kind = TC_S1_WIN;       // U-Net double_conv layers (unet.py:25) incl. the 3-channel image-side one: window form
            else if (d.op == CG_OP_CONV && d.stride == 1 && (d.same || d.k == 1) && d.k <= 7 && d.cin % 16 == 0 &&
                     d.cout % 16 == 0 && d.cout <= 256 && d.cin <= 256 * 8)
                kind = TC_S1_16;        // the same layers as 16-channel groups (CG_DISABLE_WIN=1: the round-1 path, kept for A/B)
            if (kind == TC_S1_WIN) {
                // packed window weights: forward [k*nch][Cout][64]; data gradient [k*nch_d][Cin][64] when Cin can be an N tile
                L.tc = kind;
                const int cp = d.cin <= 4 ? 8 : d.cin;
                const int nch = (d.k * cp + 63) / 64, nch_d = (d.k * d.cout + 63) / 64;
                L.pk_f = (long long)pk; pk += align_up((size_t)d.k * nch * d.cout * 64 * 2, 1024);
                L.pk_d = -1;
                if (d.cin % 16 == 0 && d.cin <= 256 && d.k * nch_d <= TC_MAX_STEPS) {
                    L.pk_d = (long long)pk; pk += align_up((size_t)d.k * nch_d * d.cin * 64 * 2, 1024);
                }
            } else if (kind == TC_IM2COL) {
                L.tc = kind;
                L.pk_f = (long long)pk; pk += align_up((size_t)d.cout * 64 * 2, 1024);
                L.pk_d = (long long)pk; pk += align_up((size_t)64 * d.cout * 2, 1024);
            } else if (kind == TC_STEM) {
                L.tc = kind;
                L.pk_f = (long long)pk; pk += align_up((size_t)d.k * d.cout * 64 * 2, 1024);
                L.pk_d = (long long)pk; pk += align_up((size_t)d.k * 32 * d.cout * 2, 1024);
            } else if (kind == TC_HEAD) {
                L.tc = kind;
                L.pk_f = (long long)pk; pk += align_up((size_t)d.k * 32 * d.cin * 2, 1024);
                L.pk_d = (long long)pk; pk += align_up((size_t)d.k * d.cin * 64 * 2, 1024);
            } else if (kind) {
                L.tc = kind;
                const size_t bytes = align_up((size_t)d.k * d.k * d.cin * d.cout * 2, 1024);
                L.pk_f = (long long)pk; pk += bytes;
                L.pk_d = (long long)pk; pk += bytes;
            }
        }
    }
    net->packed_bytes = pk;
    *out = net;
    return CG_OK;
}

static std::mutex g_single_mu;
struct SingleCall;
static std::vector<std::pair<cg_net_t, SingleCall*>> g_single;
static void single_forget(cg_net_t net);
extern "C" void cg_net_destroy(cg_net_t net) {
    if (!net) return;
    single_forget(net);
    delete net;
}

extern "C" int cg_net_param_floats(cg_net_t net, int64_t* n) {
    if (!net || !n) { cg_set_error("null argument"); return CG_ERR_INVALID; }
    *n = net->n_params;
    return CG_OK;
}
extern "C" int cg_net_state_floats(cg_net_t net, int64_t* n) {
    if (!net || !n) { cg_set_error("null argument"); return CG_ERR_INVALID; }
    *n = net->n_state;
    return CG_OK;
}
extern "C" int cg_net_bind_state(cg_net_t net, float* state_dev) {
    if (!net || (!state_dev && net->n_state)) { cg_set_error("cg_net_bind_state: null argument"); return CG_ERR_INVALID; }
    net->state = state_dev;
    return CG_OK;
}
extern "C" int cg_net_set_training(cg_net_t net, int training) {
    if (!net) { cg_set_error("null argument"); return CG_ERR_INVALID; }
    net->training = training != 0;
    return CG_OK;
}
extern "C" int cg_net_set_seed(cg_net_t net, uint64_t seed) {
    if (!net) { cg_set_error("null argument"); return CG_ERR_INVALID; }
    net->seed = seed;
    net->calls = 0;
    net->seed_epoch += 1;
    return CG_OK;
}
extern "C" int cg_net_var_count(cg_net_t net, int* n) {
    if (!net || !n) { cg_set_error("null argument"); return CG_ERR_INVALID; }
    *n = (int)net->vars.size();
    return CG_OK;
}
extern "C" int cg_net_var_info(cg_net_t net, int i, cg_var_info* out) {
    if (!net || !out || i < 0 || i >= (int)net->vars.size()) { cg_set_error("bad variable index %d", i); return CG_ERR_INVALID; }
    *out = net->vars[i];
    return CG_OK;
}
extern "C" int cg_net_out_shape(cg_net_t net, int N, int H, int W, int out[4]) {
    if (!net || !out) { cg_set_error("null argument"); return CG_ERR_INVALID; }
    int ho, wo;
    CG_TRY(net_out_hw(net, H, W, &ho, &wo));
    out[0] = N; out[1] = ho; out[2] = wo; out[3] = net->chan.back();
    return CG_OK;
}

// workspace of a single-net call: [activations | gradient arena | dy/dx staging in activation dtype]
struct SingleLayout { size_t act, arena, dy, dx, packed, tcs, total; };
static int single_layout(const cg_net_s* net, CallCtx* ctx, int N, int H, int W, bool bwd, SingleLayout* lay) {
    CG_TRY(net_plan(net, N, H, W, bwd, ctx));
    size_t es = net->elem_size();
    lay->act = 4096;                // slack: the window view of the first tensor starts a few pixels before it (tc_make_map_win)
    lay->arena = align_up(lay->act + ctx->act_bytes, 256);
    lay->dy = lay->arena + align_up(ctx->grad_bytes, 256);
    lay->dx = lay->dy + (bwd ? align_up((size_t)N * ctx->sample_elems(net->out_tensor()) * es, 256) : 0);
    lay->packed = align_up(lay->dx + (bwd ? align_up((size_t)N * ctx->sample_elems(0) * es, 256) : 0), 1024);
    lay->tcs = align_up(lay->packed + net->packed_bytes, 1024);
    lay->total = lay->tcs + ctx->tcs_bytes + 4096;
    return CG_OK;
}

extern "C" int cg_net_workspace_bytes(cg_net_t net, int N, int H, int W, int need_backward, size_t* bytes) {
    if (!net || !bytes) { cg_set_error("null argument"); return CG_ERR_INVALID; }
    CallCtx ctx; SingleLayout lay;
    CG_TRY(single_layout(net, &ctx, N, H, W, need_backward != 0, &lay));
    *bytes = lay.total;
    return CG_OK;
}

// the planned context of the last forward is kept per net so that cg_net_backward can follow it
struct SingleCall { CallCtx ctx; SingleLayout lay; void* ws = nullptr; };
static void single_forget(cg_net_t net) {
    std::lock_guard<std::mutex> lk(g_single_mu);
    for (size_t i = 0; i < g_single.size(); ++i)
        if (g_single[i].first == net) { delete g_single[i].second; g_single.erase(g_single.begin() + i); return; }
}
static SingleCall* single_of(cg_net_t net) {
    std::lock_guard<std::mutex> lk(g_single_mu);
    for (auto& p : g_single) if (p.first == net) return p.second;
    g_single.push_back({net, new SingleCall()});
    return g_single.back().second;
}

extern "C" int cg_net_forward(cg_net_t net, const float* params, const float* x, float* y, void* ws, size_t ws_bytes,
                              int N, int H, int W, int need_backward, void* stream) {
    if (!net || !params || !x || !y || !ws) { cg_set_error("cg_net_forward: null argument"); return CG_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    SingleCall* sc = single_of(net);
    CG_TRY(single_layout(net, &sc->ctx, N, H, W, need_backward != 0, &sc->lay));
    if (sc->lay.total > ws_bytes) { cg_set_error("workspace %zu < required %zu", ws_bytes, sc->lay.total); return CG_ERR_WORKSPACE; }
    sc->ws = ws;
    sc->ctx.base = (char*)ws + sc->lay.act;
    sc->ctx.arena = (char*)ws + sc->lay.arena;
    sc->ctx.ext_input = nullptr;
    sc->ctx.packed = (char*)ws + sc->lay.packed;
    sc->ctx.tcs = (char*)ws + sc->lay.tcs;
    sc->ctx.bn_group = 0;                       // one Keras call
    sc->ctx.training = net->training != 0;
    sc->ctx.defer_moving = false;
    sc->ctx.drop_ctr_dev = nullptr;
    sc->ctx.drop_ctr_host = net->calls;
    sc->ctx.call_id[0] = 0;
    if (net->training) net->calls += 1;
    CG_TRY(net_bind(&sc->ctx));
    CG_TRY(net_pack(net, params, sc->ctx.packed, st));
    const int tout = net->out_tensor();
    size_t nin = (size_t)N * sc->ctx.sample_elems(0), nout = (size_t)N * sc->ctx.sample_elems(tout);
    if (net->mode == CG_MODE_BF16) {
        CG_TRY(k_convert_in<bf16>(x, (bf16*)sc->ctx.act(0), nin, st));
        CG_TRY(net_forward(&sc->ctx, params, st));
        CG_TRY(k_convert_out<bf16>((const bf16*)sc->ctx.act(tout), y, nout, st));
    } else {
        CG_TRY(k_convert_in<float>(x, (float*)sc->ctx.act(0), nin, st));
        CG_TRY(net_forward(&sc->ctx, params, st));
        CG_TRY(k_convert_out<float>((const float*)sc->ctx.act(tout), y, nout, st));
    }
    return CG_OK;
}

// shared by cg_net_fetch_tensor / cg_trainer_fetch_tensor: copy tensor `t` of a planned, forwarded call out as float32 NHWC
int fetch_tensor(const CallCtx* c, int t, float* out, int* shape4, cudaStream_t st, int n0 = 0, int nb = -1) {
    if (nb < 0) nb = c->N - n0;        // samples [n0, n0 + nb) of the planned batch
    const cg_net_s* net = c->net;
    if (!net || t < 0 || t > net->out_tensor()) { cg_set_error("fetch_tensor: tensor id %d out of range", t); return CG_ERR_INVALID; }
    if (shape4) { shape4[0] = nb; shape4[1] = c->th.empty() ? 0 : c->th[t]; shape4[2] = c->tw.empty() ? 0 : c->tw[t]; shape4[3] = net->chan[t]; }
    const bool live = c->forwarded && t < (int)c->live.size() && c->live[t] && (t == 0 || net->has_buffer[t]);
    if (!out) return live ? CG_OK : 1;            // query form: 0 = materialised by the last forward, 1 = not (fused away)
    if (!live) { cg_set_error("fetch_tensor: tensor %d was not materialised by the last forward", t); return CG_ERR_STATE; }
    const size_t n = (size_t)nb * c->sample_elems(t), o = (size_t)n0 * c->sample_elems(t);
    if (net->mode == CG_MODE_BF16) return k_convert_out<bf16>((const bf16*)c->act(t) + o, out, n, st);
    return k_convert_out<float>((const float*)c->act(t) + o, out, n, st);
}

extern "C" int cg_net_fetch_tensor(cg_net_t net, int tensor, float* out, int shape4[4], void* stream) {
    if (!net) { cg_set_error("null argument"); return CG_ERR_INVALID; }
    SingleCall* sc = single_of(net);
    if (!sc->ctx.forwarded) { cg_set_error("cg_net_fetch_tensor: no cg_net_forward on this net yet"); return CG_ERR_STATE; }
    return fetch_tensor(&sc->ctx, tensor, out, shape4, (cudaStream_t)stream);
}

extern "C" int cg_net_backward(cg_net_t net, const float* params, const float* dy, float* dx, float* grads,
                               int accumulate, void* ws, size_t ws_bytes, void* stream) {
    if (!net || !params || !dy || !ws) { cg_set_error("cg_net_backward: null argument"); return CG_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    SingleCall* sc = single_of(net);
    if (sc->ws != ws || !sc->ctx.forwarded || !sc->ctx.bwd || sc->lay.total > ws_bytes) {
        cg_set_error("cg_net_backward: no matching cg_net_forward(need_backward=1) on this workspace");
        return CG_ERR_STATE;
    }
    CallCtx& c = sc->ctx;
    const int tout = net->out_tensor();
    size_t nin = (size_t)c.N * c.sample_elems(0), nout = (size_t)c.N * c.sample_elems(tout);
    if (grads && !accumulate) CG_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * (size_t)net->n_params, st));
    void* dy_t = (char*)ws + sc->lay.dy;
    void* dx_t = dx ? (void*)((char*)ws + sc->lay.dx) : nullptr;
    if (net->mode == CG_MODE_BF16) {
        CG_TRY(k_convert_in<bf16>(dy, (bf16*)dy_t, nout, st));
        CG_TRY(net_backward(&c, params, dy_t, dx_t, grads, 0, c.N, st));
        if (dx) CG_TRY(k_convert_out<bf16>((const bf16*)dx_t, dx, nin, st));
    } else {
        CG_TRY(k_convert_in<float>(dy, (float*)dy_t, nout, st));
        CG_TRY(net_backward(&c, params, dy_t, dx_t, grads, 0, c.N, st));
        if (dx) CG_TRY(k_convert_out<float>((const float*)dx_t, dx, nin, st));
    }
    return CG_OK;
}
