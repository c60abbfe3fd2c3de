// Graph planning and the forward / backward kernel schedules of one net call.
#include "net.h"

#include <stdarg.h>
#include <string.h>

#include "kernels.h"

static inline void same_pad(int in, int k, int s, int* before) {
    int out = (in + s - 1) / s;
    int total = (out - 1) * s + k - in;
    if (total < 0) total = 0;
    *before = total / 2;
}

// geometry of a CONV layer (forward orientation) or, for CONVT, of the forward conv F it is the
// data-gradient of (F maps the CONVT output grid back to the CONVT input grid).
static ConvGeom conv_geom(const cg_layer_desc& d, int N, int h_in, int w_in, int h_out, int w_out) {
    ConvGeom g;
    g.N = N; g.k = d.k; g.s = d.stride;
    if (d.op == CG_OP_CONV) {
        g.Hi = h_in; g.Wi = w_in; g.Cin = d.cin; g.Ho = h_out; g.Wo = w_out; g.Cout = d.cout;
    } else {   // CONVT: F: (h_out,w_out,cout) -> (h_in,w_in,cin)
        g.Hi = h_out; g.Wi = w_out; g.Cin = d.cout; g.Ho = h_in; g.Wo = w_in; g.Cout = d.cin;
    }
    g.pt = g.pl = 0;
    if (d.same) { same_pad(g.Hi, g.k, g.s, &g.pt); same_pad(g.Wi, g.k, g.s, &g.pl); }
    return g;
}

int net_out_hw(const cg_net_s* net, int H, int W, int* ho, int* wo) {
    std::vector<int> th(net->layers.size() + 1), tw(net->layers.size() + 1);
    th[0] = H; tw[0] = W;
    for (size_t i = 0; i < net->layers.size(); ++i) {
        const cg_layer_desc& d = net->layers[i].d;
        int h = th[d.in0], w = tw[d.in0], oh = h, ow = w;
        switch (d.op) {
            case CG_OP_CONV:
                if (d.same) { oh = (h + d.stride - 1) / d.stride; ow = (w + d.stride - 1) / d.stride; }
                else { oh = (h - d.k) / d.stride + 1; ow = (w - d.k) / d.stride + 1; }
                break;
            case CG_OP_CONVT: oh = h * d.stride; ow = w * d.stride; break;
            case CG_OP_RPAD: oh = h + 2 * d.pad; ow = w + 2 * d.pad; break;
            case CG_OP_AVGPOOL: oh = h / 2; ow = w / 2; break;
            case CG_OP_UPSAMPLE: oh = h * 2; ow = w * 2; break;
            default: break;
        }
        if (oh <= 0 || ow <= 0) { cg_set_error("layer %d: empty output for input %dx%d", (int)i, H, W); return CG_ERR_INVALID; }
        if ((d.op == CG_OP_ADD || d.op == CG_OP_CONCAT) && (th[d.in1] != h || tw[d.in1] != w)) {
            cg_set_error("layer %d: spatial mismatch %dx%d vs %dx%d (H, W must be multiples of the down factor)",
                         (int)i, h, w, th[d.in1], tw[d.in1]);
            return CG_ERR_INVALID;
        }
        if (d.op == CG_OP_AVGPOOL && ((h | w) & 1)) { cg_set_error("layer %d: odd size into 2x2 pooling", (int)i); return CG_ERR_INVALID; }
        if (d.op == CG_OP_RPAD && (d.pad >= h || d.pad >= w)) { cg_set_error("layer %d: reflect pad >= size", (int)i); return CG_ERR_INVALID; }
        th[i + 1] = oh; tw[i + 1] = ow;
    }
    *ho = th.back(); *wo = tw.back();
    return CG_OK;
}

int net_plan(const cg_net_s* net, int N, int H, int W, bool bwd, CallCtx* ctx) {
    if (N <= 0 || H <= 0 || W <= 0) { cg_set_error("bad call shape N=%d H=%d W=%d", N, H, W); return CG_ERR_INVALID; }
    const size_t nl = net->layers.size();
    ctx->net = net; ctx->N = N; ctx->H = H; ctx->W = W; ctx->bwd = bwd; ctx->forwarded = false;
    ctx->th.assign(nl + 1, 0); ctx->tw.assign(nl + 1, 0);
    int ho, wo;
    CG_TRY(net_out_hw(net, H, W, &ho, &wo));
    ctx->th[0] = H; ctx->tw[0] = W;
    for (size_t i = 0; i < nl; ++i) {
        const cg_layer_desc& d = net->layers[i].d;
        int h = ctx->th[d.in0], w = ctx->tw[d.in0], oh = h, ow = w;
        switch (d.op) {
            case CG_OP_CONV:
                if (d.same) { oh = (h + d.stride - 1) / d.stride; ow = (w + d.stride - 1) / d.stride; }
                else { oh = (h - d.k) / d.stride + 1; ow = (w - d.k) / d.stride + 1; }
                break;
            case CG_OP_CONVT: oh = h * d.stride; ow = w * d.stride; break;
            case CG_OP_RPAD: oh = h + 2 * d.pad; ow = w + 2 * d.pad; break;
            case CG_OP_AVGPOOL: oh = h / 2; ow = w / 2; break;
            case CG_OP_UPSAMPLE: oh = h * 2; ow = w * 2; break;
            default: break;
        }
        ctx->th[i + 1] = oh; ctx->tw[i + 1] = ow;
    }
    const size_t es = net->elem_size();
    size_t off = 0;
    ctx->act_off.assign(nl + 1, 0);
    for (size_t t = 0; t <= nl; ++t) {
        if (!net->has_buffer[t]) continue;
        ctx->act_off[t] = off;
        off += align_up((size_t)N * ctx->sample_elems((int)t) * es, 256);
    }
    ctx->stat_off.assign(nl, 0);
    size_t max_nc = 1;
    for (size_t i = 0; i < nl; ++i) {
        const cg_layer_desc& d = net->layers[i].d;
        if (d.op == CG_OP_INORM) {
            ctx->stat_off[i] = off;
            off += align_up((size_t)N * d.cin * 2 * sizeof(float), 256);
            if ((size_t)N * d.cin > max_nc) max_nc = (size_t)N * d.cin;
        }
    }
    ctx->act_bytes = off;
    // tensor-core eligibility for THIS shape: the 128-pixel M tile must be a Wb x Hb box of the output, and the
    // 64-pixel K chunk of the weight gradient a Wk x Hk box
    ctx->tc.assign(nl, TcLayer());
    ctx->grad_halo.assign(nl + 1, 0);
    for (size_t i = 0; i < nl; ++i) {
        if (!net->layers[i].tc) continue;
        const int wo = ctx->tw[i + 1], ho = ctx->th[i + 1];
        const bool box128 = (wo % 128 == 0) || (128 % wo == 0 && ho % (128 / wo) == 0);
        const bool box64 = (wo % 64 == 0) || (64 % wo == 0 && ho % (64 / wo) == 0);
        if (box128 && box64 && (size_t)(ho + 4) * (wo + 4) < (1u << 30)) {
            ctx->tc[i].on = true;
            ctx->grad_halo[i + 1] = 2;
        }
    }
    ctx->grad_off.assign(nl + 1, 0);
    size_t goff = 0;
    if (bwd) {
        for (size_t t = 1; t < nl; ++t) {      // tensor 0 -> caller's dx, last tensor -> caller's dy
            if (!net->has_buffer[t]) continue;
            ctx->grad_off[t] = goff;
            const int hl = ctx->grad_halo[t];
            goff += align_up((size_t)N * (ctx->th[t] + 2 * hl) * (ctx->tw[t] + 2 * hl) * net->chan[t] * es, 256);
        }
        ctx->scratch_off = goff;
        goff += align_up(max_nc * 2 * sizeof(float), 256);
    }
    ctx->grad_bytes = goff;
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
// tensor-core layers: packed weights and TMA descriptors
// ------------------------------------------------------------------------------------------
int net_pack(const cg_net_s* net, const float* params, void* packed, cudaStream_t st) {
    if (!net->packed_bytes) return CG_OK;
    if (!packed) { cg_set_error("net_pack: no packed-weight buffer"); return CG_ERR_STATE; }
    for (const LayerInfo& L : net->layers) {
        if (!L.tc) continue;
        CG_TRY(tc_pack_weights(params + L.w_off, (bf16*)((char*)packed + L.pk_f), (bf16*)((char*)packed + L.pk_d),
                               L.d.k * L.d.k, L.d.cin, L.d.cout, st));
    }
    return CG_OK;
}

int net_bind(CallCtx* c) {
    const cg_net_s* net = c->net;
    for (size_t i = 0; i < net->layers.size(); ++i) {
        TcLayer& t = c->tc[i];
        if (!t.on) continue;
        if (!c->packed) { cg_set_error("net_bind: tensor-core layer without packed weights"); return CG_ERR_STATE; }
        const LayerInfo& L = net->layers[i];
        const cg_layer_desc& d = L.d;
        const int tin = d.in0, tout = L.out_t;
        const int hi = c->th[tin], wi = c->tw[tin], ho = c->th[tout], wo = c->tw[tout];
        const int taps = d.k * d.k;
        const bf16* wf = (const bf16*)(c->packed + L.pk_f);
        const bf16* wd = (const bf16*)(c->packed + L.pk_d);
        // ---- forward: box mode on the (already reflect-padded) input
        {
            TcConvArgs& a = t.fa;
            memset(&a, 0, sizeof(a));
            a.Wb = wo % 128 == 0 ? 128 : wo;
            a.Hb = 128 / a.Wb;
            a.n_taps = taps; a.cchunks = d.cin / 64;
            a.bn = d.cout < 256 ? d.cout : 256; a.n_blocks_n = d.cout / a.bn;
            a.tiles_w = wo / a.Wb; a.tiles_per_img = a.tiles_w * (ho / a.Hb);
            a.n0 = 0; a.nb = c->N;
            a.out_P = wo; a.out_wvalid = wo; a.out_hvalid = ho; a.out_H = ho; a.out_W = wo; a.Cout = d.cout;
            a.b_rows_per_tap = d.cout;
            for (int kh = 0; kh < d.k; ++kh)
                for (int kw = 0; kw < d.k; ++kw) { a.dw[kh * d.k + kw] = (short)kw; a.dh[kh * d.k + kw] = (short)kh; }
            CG_TRY(tc_make_map_4d(&t.mapX, c->act(tin), d.cin, wi, hi, c->N, a.Wb, a.Hb));
            CG_TRY(tc_make_map_2d(&t.mapWf, wf, d.cin, taps * d.cout, a.bn));
        }
        if (!c->bwd) continue;
        const int hl = c->grad_halo[tout], P = wo + 2 * hl, HP = ho + 2 * hl;
        const bf16* dy = (const bf16*)(c->arena + c->grad_off[tout]);
        // ---- data gradient: flat mode over the zero-bordered dY, output = gradient of the padded input
        {
            TcConvArgs& a = t.da;
            memset(&a, 0, sizeof(a));
            a.Wb = 128; a.Hb = 1;
            a.n_taps = taps; a.cchunks = d.cout / 64;
            a.bn = d.cin < 256 ? d.cin : 256; a.n_blocks_n = d.cin / a.bn;
            a.tiles_per_img = (hi * P + 127) / 128; a.tiles_w = a.tiles_per_img;
            a.n0 = 0; a.nb = c->N;
            a.out_P = P; a.out_wvalid = wi; a.out_hvalid = hi; a.out_H = hi; a.out_W = wi; a.Cout = d.cin;
            a.b_rows_per_tap = d.cin;
            for (int kh = 0; kh < d.k; ++kh)
                for (int kw = 0; kw < d.k; ++kw) {
                    a.dw[kh * d.k + kw] = (short)((hl - kh) * P + (hl - kw));
                    a.dh[kh * d.k + kw] = 0;
                }
            CG_TRY(tc_make_map_4d(&t.mapDYflat, dy, d.cout, HP * P, 1, c->N, 128, 1));
            CG_TRY(tc_make_map_2d(&t.mapWd, wd, d.cout, taps * d.cin, a.bn));
        }
        // ---- weight gradient: 64-pixel K chunks of the input and of dY
        {
            TcWgradArgs& a = t.wa;
            memset(&a, 0, sizeof(a));
            a.Wk = wo % 64 == 0 ? 64 : wo;
            a.Hk = 64 / a.Wk;
            a.n_taps = taps; a.ci_blocks = d.cin / 128;
            a.bn = d.cout < 256 ? d.cout : 256; a.co_blocks = d.cout / a.bn;
            a.chunks_w = wo / a.Wk; a.chunks_per_img = a.chunks_w * (ho / a.Hk);
            a.dy_off = hl; a.Cin = d.cin; a.Cout = d.cout;
            for (int kh = 0; kh < d.k; ++kh)
                for (int kw = 0; kw < d.k; ++kw) { a.dw[kh * d.k + kw] = (short)kw; a.dh[kh * d.k + kw] = (short)kh; }
            CG_TRY(tc_make_map_4d(&t.mapXw, c->act(tin), d.cin, wi, hi, c->N, a.Wk, a.Hk));
            CG_TRY(tc_make_map_4d(&t.mapDYw, dy, d.cout, P, HP, c->N, a.Wk, a.Hk));
        }
    }
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
template <typename T>
static int forward_T(CallCtx* c, const float* params, cudaStream_t st) {
    const cg_net_s* net = c->net;
    const int N = c->N;
    for (size_t i = 0; i < net->layers.size(); ++i) {
        const LayerInfo& L = net->layers[i];
        if (L.skipped) continue;
        const cg_layer_desc& d = L.d;
        const int tin = d.in0, tout = L.out_t;
        const T* x = (const T*)c->act(tin);
        T* y = (T*)c->act(tout);
        const int h = c->th[tin], w = c->tw[tin], oh = c->th[tout], ow = c->tw[tout];
        switch (d.op) {
            case CG_OP_CONV: {
                ConvGeom g = conv_geom(d, N, h, w, oh, ow);
                if (c->tc[i].on) {
                    TcConvArgs a = c->tc[i].fa;
                    a.nb = N;
                    CG_TRY(tc_conv_launch(&c->tc[i].mapX, &c->tc[i].mapWf, (bf16*)y,
                                          L.b_off >= 0 ? params + L.b_off : nullptr, a,
                                          2.0 * N * oh * ow * (double)d.cout * d.k * d.k * d.cin, st));
                } else {
                    CG_TRY(k_conv_fwd<T>(x, params + L.w_off, L.b_off >= 0 ? params + L.b_off : nullptr, y, g, 0, st));
                }
                break;
            }
            case CG_OP_CONVT: {
                ConvGeom g = conv_geom(d, N, h, w, oh, ow);
                CG_TRY(k_conv_dgrad<T>(x, params + L.w_off, L.b_off >= 0 ? params + L.b_off : nullptr, y, g, 0, st));
                break;
            }
            case CG_OP_INORM: {
                float* stats = (float*)(c->base + c->stat_off[i]);
                CG_TRY(k_in_stats<T>(x, stats, N, h * w, d.cin, d.eps, st));
                CG_TRY(k_in_apply<T>(x, y, stats, L.g_off >= 0 ? params + L.g_off : nullptr,
                                     L.be_off >= 0 ? params + L.be_off : nullptr, L.fused_act, L.fused_slope, N,
                                     h * w, d.cin, st));
                break;
            }
            case CG_OP_ACT:
                CG_TRY(k_act_fwd<T>(x, y, (size_t)N * c->sample_elems(tin), d.act, d.slope, st));
                break;
            case CG_OP_RPAD:
                CG_TRY(k_rpad_fwd<T>(x, y, N, h, w, d.cin, d.pad, st));
                break;
            case CG_OP_ADD:
                CG_TRY(k_add<T>(x, (const T*)c->act(d.in1), y, (size_t)N * c->sample_elems(tin), st));
                break;
            case CG_OP_CONCAT: {
                int ca = net->chan[d.in0], cb = net->chan[d.in1];
                size_t npix = (size_t)N * h * w;
                CG_TRY(k_slice_copy<T>(x, ca, 0, y, ca + cb, 0, ca, npix, 0, st));
                CG_TRY(k_slice_copy<T>((const T*)c->act(d.in1), cb, 0, y, ca + cb, ca, cb, npix, 0, st));
                break;
            }
            case CG_OP_AVGPOOL: CG_TRY(k_avgpool_fwd<T>(x, y, N, h, w, d.cin, st)); break;
            case CG_OP_UPSAMPLE: CG_TRY(k_upsample_fwd<T>(x, y, N, h, w, d.cin, st)); break;
            default: cg_set_error("unknown op %d", d.op); return CG_ERR_INVALID;
        }
    }
    c->forwarded = true;
    return CG_OK;
}

int net_forward(CallCtx* ctx, const float* params, cudaStream_t st) {
    return ctx->net->mode == CG_MODE_BF16 ? forward_T<bf16>(ctx, params, st) : forward_T<float>(ctx, params, st);
}

// ------------------------------------------------------------------------------------------
template <typename T>
static int backward_T(CallCtx* c, const float* params, const T* dy_out, T* dx_in, float* grads, int n0, int nb,
                      cudaStream_t st) {
    const cg_net_s* net = c->net;
    const int nl = (int)net->layers.size();
    if (!c->bwd || !c->forwarded) { cg_set_error("backward without a planned forward"); return CG_ERR_STATE; }
    if (n0 < 0 || nb <= 0 || n0 + nb > c->N) { cg_set_error("bad sub-batch [%d,%d) of %d", n0, n0 + nb, c->N); return CG_ERR_INVALID; }
    // per-tensor pointers for this sub-batch
    auto A = [&](int t) -> const T* { return (const T*)c->act(t) + (size_t)n0 * c->sample_elems(t); };
    auto G = [&](int t) -> T* {
        if (t == nl) return const_cast<T*>(dy_out);
        if (t == 0) return dx_in;
        return (T*)(c->arena + c->grad_off[t]);
    };
    std::vector<char> written(nl + 1, 0);
    auto need = [&](int t) -> bool { return dx_in != nullptr || (t != 0 && net->dep_params[t]); };
    float* scratch = (float*)(c->arena + c->scratch_off);

    for (int i = nl - 1; i >= 0; --i) {
        const LayerInfo& L = net->layers[i];
        if (L.skipped) continue;
        const cg_layer_desc& d = L.d;
        const int tin = d.in0, tout = L.out_t;
        if (!need(tout) && tout != nl) continue;
        if (!written[tout] && tout != nl) continue;       // nothing downstream asked for it
        const T* dy = G(tout);
        const int h = c->th[tin], w = c->tw[tin], oh = c->th[tout], ow = c->tw[tout];
        const bool want_dx = need(tin);
        T* dx = want_dx ? G(tin) : nullptr;
        const int acc = want_dx ? (int)written[tin] : 0;
        switch (d.op) {
            case CG_OP_CONV: {
                ConvGeom g = conv_geom(d, nb, h, w, oh, ow);
                if (c->tc[i].on) {      // dy lives in a zero-bordered buffer (halo 2) written by the IN backward
                    const int hl = c->grad_halo[tout];
                    const double fl = 2.0 * nb * oh * ow * (double)d.cout * d.k * d.k * d.cin;
                    if (acc) { cg_set_error("tensor-core dgrad cannot accumulate"); return CG_ERR_STATE; }
                    if (grads) {
                        TcWgradArgs a = c->tc[i].wa;
                        a.n0 = n0; a.nb = nb;
                        CG_TRY(tc_wgrad_launch(&c->tc[i].mapXw, &c->tc[i].mapDYw, grads + L.w_off, a, fl, st));
                        if (L.b_off >= 0)
                            CG_TRY(k_colsum<T>(dy, grads + L.b_off, (size_t)nb * (oh + 2 * hl) * (ow + 2 * hl), d.cout, st));
                    }
                    if (want_dx) {
                        TcConvArgs a = c->tc[i].da;
                        a.nb = nb;
                        CG_TRY(tc_conv_launch(&c->tc[i].mapDYflat, &c->tc[i].mapWd, (bf16*)dx, nullptr, a, fl, st));
                    }
                    break;
                }
                if (grads) {
                    CG_TRY(k_conv_wgrad<T>(A(tin), dy, grads + L.w_off, g, st));
                    if (L.b_off >= 0) CG_TRY(k_colsum<T>(dy, grads + L.b_off, (size_t)nb * oh * ow, d.cout, st));
                }
                if (want_dx) CG_TRY(k_conv_dgrad<T>(dy, params + L.w_off, nullptr, dx, g, acc, st));
                break;
            }
            case CG_OP_CONVT: {
                ConvGeom g = conv_geom(d, nb, h, w, oh, ow);
                if (grads) {
                    CG_TRY(k_conv_wgrad<T>(dy, A(tin), grads + L.w_off, g, st));
                    if (L.b_off >= 0) CG_TRY(k_colsum<T>(dy, grads + L.b_off, (size_t)nb * oh * ow, d.cout, st));
                }
                if (want_dx) CG_TRY(k_conv_fwd<T>(dy, params + L.w_off, nullptr, dx, g, acc, st));
                break;
            }
            case CG_OP_INORM: {
                const float* stats = (const float*)(c->base + c->stat_off[i]) + (size_t)n0 * d.cin * 2;
                bool pg = grads && L.g_off >= 0;
                CG_TRY(k_in_bwd<T>(A(tin), dy, dx, stats, L.g_off >= 0 ? params + L.g_off : nullptr,
                                   L.be_off >= 0 ? params + L.be_off : nullptr, pg ? grads + L.g_off : nullptr,
                                   pg ? grads + L.be_off : nullptr, scratch, L.fused_act, L.fused_slope, nb, h * w,
                                   d.cin, acc, st, c->grad_halo[tin], w));
                break;
            }
            case CG_OP_ACT:
                if (want_dx)
                    CG_TRY(k_act_bwd<T>(A(tout), dy, dx, (size_t)nb * c->sample_elems(tin), d.act, d.slope, acc, st));
                break;
            case CG_OP_RPAD:
                if (want_dx) CG_TRY(k_rpad_bwd<T>(dy, dx, nb, h, w, d.cin, d.pad, acc, st));
                break;
            case CG_OP_ADD: {
                size_t n = (size_t)nb * c->sample_elems(tin);
                if (want_dx) CG_TRY(k_copy_acc<T>(dy, dx, n, acc, st));
                if (need(d.in1)) {
                    CG_TRY(k_copy_acc<T>(dy, G(d.in1), n, (int)written[d.in1] || (d.in1 == tin && want_dx), st));
                    written[d.in1] = 1;
                }
                break;
            }
            case CG_OP_CONCAT: {
                int ca = net->chan[d.in0], cb = net->chan[d.in1];
                size_t npix = (size_t)nb * h * w;
                if (want_dx) CG_TRY(k_slice_copy<T>(dy, ca + cb, 0, dx, ca, 0, ca, npix, acc, st));
                if (need(d.in1)) {
                    CG_TRY(k_slice_copy<T>(dy, ca + cb, ca, G(d.in1), cb, 0, cb, npix,
                                           (int)written[d.in1] || (d.in1 == tin && want_dx), st));
                    written[d.in1] = 1;
                }
                break;
            }
            case CG_OP_AVGPOOL:
                if (want_dx) CG_TRY(k_avgpool_bwd<T>(dy, dx, nb, h, w, d.cin, acc, st));
                break;
            case CG_OP_UPSAMPLE:
                if (want_dx) CG_TRY(k_upsample_bwd<T>(dy, dx, nb, h, w, d.cin, acc, st));
                break;
            default: cg_set_error("unknown op %d", d.op); return CG_ERR_INVALID;
        }
        if (want_dx) written[tin] = 1;
    }
    return CG_OK;
}

int net_backward(CallCtx* ctx, const float* params, const void* dy, void* dx, float* grads, int n0, int nb,
                 cudaStream_t st) {
    return ctx->net->mode == CG_MODE_BF16
               ? backward_T<bf16>(ctx, params, (const bf16*)dy, (bf16*)dx, grads, n0, nb, st)
               : backward_T<float>(ctx, params, (const float*)dy, (float*)dx, grads, n0, nb, st);
}
