// Graph planning and the forward / backward kernel schedules of one net call.
#include "net.h"

#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

#include "conv_special.h"
#include "kernels.h"

static int zero_region(void* p, size_t bytes, cudaStream_t st) {
    CG_CUDA(cudaMemsetAsync(p, 0, bytes, st));
    return CG_OK;
}

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
    ctx->raw_off.assign(nl, 0);
    ctx->bstat_off.assign(nl, 0);
    ctx->stat_begin = off;
    size_t max_nc = 1;
    for (size_t i = 0; i < nl; ++i) {
        const cg_layer_desc& d = net->layers[i].d;
        if (d.op == CG_OP_INORM) {
            ctx->stat_off[i] = off;
            off += align_up((size_t)N * d.cin * 2 * sizeof(float), 256);
            if ((size_t)N * d.cin > max_nc) max_nc = (size_t)N * d.cin;
            if (net->layers[i].batch) {      // per-group batch statistics (at most N groups)
                ctx->bstat_off[i] = off;
                off += align_up((size_t)N * d.cin * 2 * sizeof(float), 256);
            } else {                         // raw sums of a fused-statistics conv (see CallCtx::raw_off)
                ctx->raw_off[i] = off;
                off += align_up((size_t)N * d.cin * 2 * sizeof(float), 256);
            }
        }
    }
    ctx->act_bytes = off;
    // tensor-core eligibility for THIS shape: every 128-pixel M tile must be a Wb x Hb box of the (class) output grid
    // and every 64-pixel K chunk of the weight gradient a Wk x Hk box
    ctx->tc.assign(nl, TcLayer());
    ctx->grad_halo.assign(nl + 1, 0);
    ctx->tcs_bytes = 0;
    auto boxable = [](int w, int h, int px) { return (w % px == 0) || (px % w == 0 && h % (px / w) == 0); };
    for (size_t i = 0; i < nl; ++i) {
        const int kind = net->layers[i].tc;
        if (!kind) continue;
        const cg_layer_desc& d = net->layers[i].d;
        const int hi = ctx->th[d.in0], wi = ctx->tw[d.in0], ho = ctx->th[i + 1], wo = ctx->tw[i + 1];
        bool ok = false;
        if (kind == TC_S1_VALID) {
            ok = boxable(wo, ho, 128) && boxable(wo, ho, 64) && (size_t)(ho + 4) * (wo + 4) < (1u << 30);
            if (ok) ctx->grad_halo[i + 1] = 2;
        } else if (kind == TC_CONV_S2) {       // fwd tiles on (ho,wo); dgrad classes on (hi/2, wi/2)
            ok = !((hi | wi) & 1) && boxable(wo, ho, 128) && boxable(wo, ho, 64) && boxable(wi / 2, hi / 2, 128);
        } else if (kind == TC_CONVT_S2) {      // fwd classes on (hi,wi) = half the output grid; its dgrad tiles on (hi,wi)
            ok = boxable(wi, hi, 128) && boxable(wi, hi, 64);
        }
        else if (kind == TC_S1_16) {
            ok = boxable(wo, ho, 128) && d.cin <= 256;      // dgrad N tile = Cin
        } else if (kind == TC_S1_WIN) {
            ok = boxable(wo, ho, 128);
            if (ok && d.cin <= 4) {          // scratch: the input re-laid with 8 channels per pixel
                const size_t need = align_up((size_t)N * hi * wi * 8 * 2, 1024);
                if (need > ctx->tcs_bytes) ctx->tcs_bytes = need;
            }
        } else if (kind == TC_IM2COL) {
            ok = boxable(wo, ho, 128) && boxable(wo, ho, 64);
        } else if (kind == TC_STEM) {
            ok = boxable(wo, ho, 128) && boxable(wo, ho, 64);
        } else if (kind == TC_HEAD) {
            ok = (size_t)hi * wi < (1u << 30) && 6 * wi < 32767;
        }
        ctx->tc[i].on = ok;
        if (ok && kind == TC_IM2COL) {       // scratch: the unfolded tensor / its gradient [N][ho][wo][64], then the fp32 dW staging
            const size_t big = (size_t)N * ho * wo * 64 * 2, tmp = (size_t)2 * 64 * d.cout * 4;
            ctx->tc[i].sc_tmp = align_up(big, 1024);
            const size_t need = ctx->tc[i].sc_tmp + align_up(tmp, 1024);
            if (need > ctx->tcs_bytes) ctx->tcs_bytes = need;
        }
        if (ok && (kind == TC_STEM || kind == TC_HEAD)) {
            // scratch: the unfolded tensor (64 channels; S of the head forward is smaller) then the fp32 dW staging
            // ([vertical taps rounded up to pairs][64][C])
            const int ntp = 2 * (((d.k + 2) / 3 + 1) / 2);
            const size_t big = kind == TC_STEM ? (size_t)N * hi * wo * 64 * 2 : (size_t)N * (ho + 2) * wi * 64 * 2;
            const size_t tmp = (size_t)ntp * 64 * (kind == TC_STEM ? d.cout : d.cin) * 4;
            ctx->tc[i].sc_tmp = align_up(big, 1024);
            const size_t need = ctx->tc[i].sc_tmp + align_up(tmp, 1024);
            if (need > ctx->tcs_bytes) ctx->tcs_bytes = need;
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
        ctx->sums_off.assign(nl, 0);
        for (size_t i = 0; i < nl; ++i) {
            const cg_layer_desc& d = net->layers[i].d;
            if (d.op != CG_OP_INORM) continue;
            ctx->sums_off[i] = goff;
            goff += align_up((size_t)N * d.cin * 2 * sizeof(float), 256);
        }
        ctx->sums_bytes = goff - ctx->scratch_off;
        (void)max_nc;
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
    TcPackJobs jobs;
    jobs.n = 0;
    TcPackWinJobs wjobs;
    wjobs.n = 0;
    for (const LayerInfo& L : net->layers) {
        if (!L.tc) continue;
        if (L.tc == TC_STEM) {
            CG_TRY(sp_pack_stem(params + L.w_off, (bf16*)((char*)packed + L.pk_f), L.d.k, L.d.cin, L.d.cout, st));
            CG_TRY(sp_pack_stem_d(params + L.w_off, (bf16*)((char*)packed + L.pk_d), L.d.k, L.d.cin, L.d.cout, st));
            continue;
        }
        if (L.tc == TC_S1_WIN) {
            const int cp = L.d.cin <= 4 ? 8 : L.d.cin;
            CG_TRY(tc_pack_win(wjobs, params + L.w_off, (bf16*)((char*)packed + L.pk_f), L.d.k, cp, L.d.cin, L.d.cout, L.d.cout,
                               L.d.cin, L.d.cout, 0, st));
            if (L.pk_d >= 0)
                CG_TRY(tc_pack_win(wjobs, params + L.w_off, (bf16*)((char*)packed + L.pk_d), L.d.k, L.d.cout, L.d.cout, L.d.cin,
                                   L.d.cin, L.d.cin, L.d.cout, 1, st));
            continue;
        }
        if (L.tc == TC_IM2COL) {
            CG_TRY(sp_pack_im2col(params + L.w_off, (bf16*)((char*)packed + L.pk_f), (bf16*)((char*)packed + L.pk_d), L.d.k,
                                  L.d.cin, L.d.cout, st));
            continue;
        }
        if (L.tc == TC_HEAD) {
            CG_TRY(sp_pack_head(params + L.w_off, (bf16*)((char*)packed + L.pk_f), (bf16*)((char*)packed + L.pk_d), L.d.k,
                                L.d.cin, L.d.cout, st));
            continue;
        }
        // the TF kernel of a Conv2DTranspose (kh,kw,Cout,Cin) IS the HWIO kernel of the conv F it back-propagates
        const int cin_f = L.d.op == CG_OP_CONV ? L.d.cin : L.d.cout, cout_f = L.d.op == CG_OP_CONV ? L.d.cout : L.d.cin;
        const int j = jobs.n++;
        jobs.w[j] = params + L.w_off;
        jobs.wf[j] = (bf16*)((char*)packed + L.pk_f);
        jobs.wd[j] = (bf16*)((char*)packed + L.pk_d);
        jobs.taps[j] = L.d.k * L.d.k; jobs.cin[j] = cin_f; jobs.cout[j] = cout_f;
        if (jobs.n == TC_PACK_MAX) CG_TRY(tc_pack_weights_multi(jobs, st));
    }
    CG_TRY(tc_pack_win_flush(wjobs, st));
    return tc_pack_weights_multi(jobs, st);
}

static inline int floor_div2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

// tiles of 128 output pixels as a Wb x Hb box of a (w x h) grid
static void set_tiles(TcConvArgs& a, int w, int h) {
    a.Wb = w % 128 == 0 ? 128 : w;
    a.Hb = 128 / a.Wb;
    a.tiles_w = w / a.Wb;
    a.tiles_per_img = a.tiles_w * (h / a.Hb);
    a.out_P = w; a.out_wvalid = w; a.out_hvalid = h;
}

// 16-channel-group (SWIZZLE_32B) conv launch: out[n, p, :] = sum_taps in[n, p + (dw, dh)[tap], :] * W[tap]   with
// zero fill outside `in` (TMA out-of-bounds).  Box mode tiles the (out_w x out_h) grid; flat mode walks the output as
// one line of out_w*out_h pixels in 128-pixel tiles and `in` as one line of map_w pixels (map_h = 1, dw = flat offsets).
static int make_conv16(TcConvLaunch& Ln, const void* in, int cin_k, int map_w, int map_h, int N, const bf16* wmat, int n_out,
                       int n_taps, const short* dw, const short* dh, bool flat, int out_w, int out_h) {
    TcConvArgs& a = Ln.a;
    memset(&a, 0, sizeof(a));
    if (flat) {
        a.Wb = 128; a.Hb = 1;
        a.tiles_per_img = (out_w * out_h + 127) / 128; a.tiles_w = a.tiles_per_img;
        a.out_P = out_w; a.out_wvalid = out_w; a.out_hvalid = out_h;
    } else {
        set_tiles(a, out_w, out_h);
    }
    a.n_taps = n_taps;
    a.bk16 = 1; a.cin16 = cin_k / 16; a.groups = a.cin16 < 8 ? a.cin16 : 8;
    a.cchunks = (a.cin16 + a.groups - 1) / a.groups;
    a.bn = n_out; a.n_blocks_n = 1;
    a.nb = N; a.out_H = out_h; a.out_W = out_w; a.Cout = n_out; a.out_sy = a.out_sx = 1;
    a.b_rows_per_tap = n_out;
    for (int tp = 0; tp < n_taps; ++tp) { a.tb[tp] = (short)tp; a.dw[tp] = dw[tp]; a.dh[tp] = dh[tp]; }
    // few input channels: several taps share a K step (fewer barrier round trips and one weight box for all of them);
    // limits: 8 sixteen-channel groups per step, a weight box of <= 256 rows
    a.tps = 1;
    if (a.cchunks == 1) {
        int tps = 8 / a.cin16;
        if (tps * n_out > 256) tps = 256 / n_out;
        if (tps > n_taps) tps = n_taps;
        if (tps > 1) a.tps = tps;
    }
    CG_TRY(tc_make_map_act16(&Ln.mapA, in, cin_k, map_w, map_h, N, a.Wb, a.Hb, a.groups));
    CG_TRY(tc_make_map_w16(&Ln.mapB, wmat, cin_k, n_taps * n_out, a.tps * n_out, a.groups));
    Ln.mapB2 = Ln.mapB;
    return CG_OK;
}

// conv-type launch: out[n,oh,ow,:] = sum_taps in[n, oh*s + kh - pt, ow*s + kw - pl, :] * Wf[tap]  (+ bias)
//   `in` is [hi x wi x cin_k] (cin_k = GEMM K per tap), `Wf` is the [tap][n_out][cin_k] packing
static int make_conv_launch(TcConvLaunch& L, const void* in, int hi, int wi, int cin_k, int N, const bf16* wmat, int n_out,
                            int k, int s, int pt, int pl, int ho, int wo) {
    TcConvArgs& a = L.a;
    memset(&a, 0, sizeof(a));
    set_tiles(a, wo, ho);
    a.n_taps = k * k; a.cchunks = cin_k / 64;
    a.bn = n_out < 256 ? n_out : 256; a.n_blocks_n = n_out / a.bn;
    a.n0 = 0; a.nb = N;
    a.out_H = ho; a.out_W = wo; a.Cout = n_out;
    a.out_sy = a.out_sx = 1; a.out_oy = a.out_ox = 0;
    a.b_rows_per_tap = n_out;
    for (int kh = 0; kh < k; ++kh)
        for (int kw = 0; kw < k; ++kw) {
            const int t = kh * k + kw, oh = kh - pt, ow = kw - pl;
            a.tb[t] = (short)t;
            if (s == 1) { a.dc[t] = 0; a.dw[t] = (short)ow; a.dp[t] = 0; a.dh[t] = (short)oh; }
            else {
                const int ah = floor_div2(oh), aw = floor_div2(ow);
                a.dc[t] = (short)((ow - 2 * aw) * cin_k); a.dw[t] = (short)aw; a.dp[t] = (short)(oh - 2 * ah); a.dh[t] = (short)ah;
            }
        }
    CG_TRY(tc_make_map_act(&L.mapA, in, cin_k, wi, hi, N, s == 2, a.Wb, a.Hb));
    CG_TRY(tc_make_map_2d(&L.mapB, wmat, cin_k, k * k * n_out, a.bn));
    CG_TRY(tc_make_map_2d(&L.mapB2, wmat, cin_k, k * k * n_out, (a.bn) / 2 >= 8 ? (a.bn) / 2 : 8));
    return CG_OK;
}

// data-gradient of a stride-2 conv F, one launch per output parity class (ph,pw):
//   dx[n, 2hc+ph, 2wc+pw, ci] = sum_{taps of the class} dy[n, hc + qh - a, wc + qw - b, :] * Wd[tap][ci][:]  (+ bias)
//   `dy` is [ho x wo x cout_f] dense (zero padding = TMA out-of-bounds fill), `Wd` the native [tap][cin_f][cout_f] layout
static int make_class_launches(std::vector<TcConvLaunch>& out, const void* dy, int ho, int wo, int cout_f, int N,
                               const bf16* wd, int cin_f, int k, int pt, int pl, int hi, int wi) {
    out.assign(4, TcConvLaunch());
    int total_taps = 0;
    for (int cls = 0; cls < 4; ++cls) {
        const int ph = cls >> 1, pw = cls & 1;
        TcConvLaunch& L = out[cls];
        TcConvArgs& a = L.a;
        memset(&a, 0, sizeof(a));
        set_tiles(a, wi / 2, hi / 2);
        const int kh0 = (ph + pt) % 2, kw0 = (pw + pl) % 2;
        const int nkh = kh0 < k ? (k - kh0 + 1) / 2 : 0, nkw = kw0 < k ? (k - kw0 + 1) / 2 : 0;
        const int qh = (ph + pt - kh0) / 2, qw = (pw + pl - kw0) / 2;
        a.n_taps = nkh * nkw; a.cchunks = cout_f / 64;
        a.bn = cin_f < 256 ? cin_f : 256; a.n_blocks_n = cin_f / a.bn;
        a.n0 = 0; a.nb = N;
        a.out_H = hi; a.out_W = wi; a.Cout = cin_f;
        a.out_sy = a.out_sx = 2; a.out_oy = ph; a.out_ox = pw;
        a.b_rows_per_tap = cin_f;
        for (int ia = 0; ia < nkh; ++ia)
            for (int ib = 0; ib < nkw; ++ib) {
                const int t = ia * nkw + ib;
                a.tb[t] = (short)((kh0 + 2 * ia) * k + (kw0 + 2 * ib));
                a.dc[t] = 0; a.dp[t] = 0; a.dw[t] = (short)(qw - ib); a.dh[t] = (short)(qh - ia);
            }
        total_taps += a.n_taps;
        CG_TRY(tc_make_map_act(&L.mapA, dy, cout_f, wo, ho, N, 0, a.Wb, a.Hb));
        CG_TRY(tc_make_map_2d(&L.mapB, wd, cout_f, k * k * cin_f, a.bn));
        CG_TRY(tc_make_map_2d(&L.mapB2, wd, cout_f, k * k * cin_f, (a.bn) / 2 >= 8 ? (a.bn) / 2 : 8));
    }
    for (auto& L : out) L.flop_share = (double)L.a.n_taps / (double)(total_taps ? total_taps : 1);
    return CG_OK;
}

// weight gradient of conv F (stride s): X = F's input [hi x wi x cin_f] (parity view when s == 2), dY = F's output gradient
static int make_wgrad(TcLayer& t, const void* x, int hi, int wi, int cin_f, const void* dy, int ho, int wo, int cout_f, int N,
                      int k, int s, int pt, int pl, int dy_halo) {
    TcWgradArgs& a = t.wa;
    memset(&a, 0, sizeof(a));
    a.Wk = wo % 64 == 0 ? 64 : wo;
    a.Hk = 64 / a.Wk;
    a.n_taps = k * k;
    a.transposed = cin_f % 128 == 0 ? 0 : 1;
    if (!a.transposed) { a.a_blocks = cin_f / 128; a.bn = cout_f < 256 ? cout_f : 256; a.b_blocks = cout_f / a.bn; }
    else { a.a_blocks = cout_f / 128; a.bn = cin_f < 256 ? cin_f : 256; a.b_blocks = cin_f / a.bn; }
    a.chunks_w = wo / a.Wk; a.chunks_per_img = a.chunks_w * (ho / a.Hk);
    a.dy_off = dy_halo; a.Cin = cin_f; a.Cout = cout_f;
    for (int kh = 0; kh < k; ++kh)
        for (int kw = 0; kw < k; ++kw) {
            const int tp = kh * k + kw, oh = kh - pt, ow = kw - pl;
            if (s == 1) { a.dc[tp] = 0; a.dw[tp] = (short)ow; a.dp[tp] = 0; a.dh[tp] = (short)oh; }
            else {
                const int ah = floor_div2(oh), aw = floor_div2(ow);
                a.dc[tp] = (short)((ow - 2 * aw) * cin_f); a.dw[tp] = (short)aw; a.dp[tp] = (short)(oh - 2 * ah); a.dh[tp] = (short)ah;
            }
        }
    // operand roles: A (128 rows = 2 channel groups) and B (bn columns = bn/64 groups)
    const int x_groups = a.transposed ? a.bn / 64 : 2, y_groups = a.transposed ? 2 : a.bn / 64;
    if (s == 1) {
        a.x_grouped = 1;
        CG_TRY(tc_make_map_act_grouped(&t.mapXw, x, cin_f, wi, hi, N, a.Wk, a.Hk, x_groups));
    } else {
        CG_TRY(tc_make_map_act(&t.mapXw, x, cin_f, wi, hi, N, 1, a.Wk, a.Hk));
    }
    a.y_grouped = 1;
    CG_TRY(tc_make_map_act_grouped(&t.mapDYw, dy, cout_f, wo + 2 * dy_halo, ho + 2 * dy_halo, N, a.Wk, a.Hk, y_groups));
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
        const int tin = d.in0, tout = L.out_t, k = d.k;
        const int hi = c->th[tin], wi = c->tw[tin], ho = c->th[tout], wo = c->tw[tout];
        const bf16* wf = (const bf16*)(c->packed + L.pk_f);     // [tap][Cout_F][Cin_F]
        const bf16* wd = (const bf16*)(c->packed + L.pk_d);     // [tap][Cin_F][Cout_F]  (native TF layout)
        const void* x = c->act(tin);
        const void* dy = c->bwd ? (const void*)(c->arena + c->grad_off[tout]) : nullptr;
        if (L.tc == TC_S1_VALID) {
            t.fwd.assign(1, TcConvLaunch());
            CG_TRY(make_conv_launch(t.fwd[0], x, hi, wi, d.cin, c->N, wf, d.cout, k, 1, 0, 0, ho, wo));
            if (!c->bwd) continue;
            // data gradient in flat mode over the zero-bordered dY: output = gradient of the (padded) input, computed
            // in the row pitch P of the dY buffer so that every tap is a constant row shift
            const int hl = c->grad_halo[tout], P = wo + 2 * hl, HP = ho + 2 * hl;
            t.dgrad.assign(1, TcConvLaunch());
            TcConvArgs& a = t.dgrad[0].a;
            memset(&a, 0, sizeof(a));
            a.Wb = 128; a.Hb = 1;
            a.n_taps = k * k; a.cchunks = d.cout / 64;
            a.bn = d.cin < 256 ? d.cin : 256; a.n_blocks_n = d.cin / a.bn;
            a.tiles_per_img = (hi * P + 127) / 128; a.tiles_w = a.tiles_per_img;
            a.n0 = 0; a.nb = c->N;
            a.out_P = P; a.out_wvalid = wi; a.out_hvalid = hi; a.out_H = hi; a.out_W = wi; a.Cout = d.cin;
            a.out_sy = a.out_sx = 1;
            a.b_rows_per_tap = d.cin;
            for (int kh = 0; kh < k; ++kh)
                for (int kw = 0; kw < k; ++kw) {
                    const int tp = kh * k + kw;
                    a.tb[tp] = (short)tp;
                    a.dw[tp] = (short)((hl - kh) * P + (hl - kw));
                }
            CG_TRY(tc_make_map_act(&t.dgrad[0].mapA, dy, d.cout, HP * P, 1, c->N, 0, 128, 1));
            CG_TRY(tc_make_map_2d(&t.dgrad[0].mapB, wd, d.cout, k * k * d.cin, a.bn));
            CG_TRY(tc_make_map_2d(&t.dgrad[0].mapB2, wd, d.cout, k * k * d.cin, (a.bn) / 2 >= 8 ? (a.bn) / 2 : 8));
            CG_TRY(make_wgrad(t, x, hi, wi, d.cin, dy, ho, wo, d.cout, c->N, k, 1, 0, 0, hl));
        } else if (L.tc == TC_CONV_S2) {
            int pt, pl;
            same_pad(hi, k, 2, &pt);
            same_pad(wi, k, 2, &pl);
            t.fwd.assign(1, TcConvLaunch());
            CG_TRY(make_conv_launch(t.fwd[0], x, hi, wi, d.cin, c->N, wf, d.cout, k, 2, pt, pl, ho, wo));
            if (!c->bwd) continue;
            CG_TRY(make_class_launches(t.dgrad, dy, ho, wo, d.cout, c->N, wd, d.cin, k, pt, pl, hi, wi));
            CG_TRY(make_wgrad(t, x, hi, wi, d.cin, dy, ho, wo, d.cout, c->N, k, 2, pt, pl, 0));
        } else if (L.tc == TC_S1_16) {
            // stride-1 'same' conv with 16-multiple channels: forward and data gradient are both box-mode convs whose
            // zero padding is TMA out-of-bounds fill; K runs over 16-channel groups (SWIZZLE_32B tiles)
            int pt = 0, pl = 0;
            if (d.same) { same_pad(hi, k, 1, &pt); same_pad(wi, k, 1, &pl); }
            auto make16 = [&](TcConvLaunch& Ln, const void* in, int cin_k, const bf16* wmat, int n_out, bool flip) -> int {
                short tdw[49], tdh[49];
                for (int kh = 0; kh < k; ++kh)
                    for (int kw = 0; kw < k; ++kw) {
                        tdw[kh * k + kw] = (short)(flip ? pl - kw : kw - pl);
                        tdh[kh * k + kw] = (short)(flip ? pt - kh : kh - pt);
                    }
                return make_conv16(Ln, in, cin_k, wo, ho, c->N, wmat, n_out, k * k, tdw, tdh, false, wo, ho);
            };
            t.fwd.assign(1, TcConvLaunch());
            CG_TRY(make16(t.fwd[0], x, d.cin, wf, d.cout, false));
            if (!c->bwd) continue;
            t.dgrad.assign(1, TcConvLaunch());
            CG_TRY(make16(t.dgrad[0], dy, d.cout, wd, d.cin, true));
            const char* no16 = getenv("CG_DISABLE_WGRAD16");      // test hook: CUDA-core weight gradient for these layers
            if (!(no16 && no16[0] == '1') && ((wo % 64 == 0) || (64 % wo == 0 && ho % (64 / wo) == 0))) {      // weight gradient with tap-stacked M (wgrad16_tc_kernel)
                TcWgrad16Args& a = t.wa16;
                memset(&a, 0, sizeof(a));
                a.n_taps = k * k; a.Cin = d.cin; a.Cout = d.cout;
                a.Wk = wo % 64 == 0 ? 64 : wo; a.Hk = 64 / a.Wk;
                a.chunks_w = wo / a.Wk; a.chunks_per_img = a.chunks_w * (ho / a.Hk);
                for (int kh = 0; kh < k; ++kh)
                    for (int kw = 0; kw < k; ++kw) { a.dw[kh * k + kw] = (short)(kw - pl); a.dh[kh * k + kw] = (short)(kh - pt); }
                CG_TRY(tc_make_map_act16(&t.mapXw, x, d.cin, wi, hi, c->N, a.Wk, a.Hk, 1));
                CG_TRY(tc_make_map_act16(&t.mapDYw, dy, d.cout, wo, ho, c->N, a.Wk, a.Hk, d.cout / 16));
                t.wg16 = true;
            }
        } else if (L.tc == TC_S1_WIN) {
            // stride-1 'same' (or 1x1) conv in window form: a K step = 64 contiguous elements of the k*C-element window of one
            // kernel row; forward and data gradient are the same launch shape over x / dY
            int pt = 0, pl = 0;
            if (d.same) { same_pad(hi, k, 1, &pt); same_pad(wi, k, 1, &pl); }
            const int cp = d.cin <= 4 ? 8 : d.cin;
            const void* xs = d.cin <= 4 ? (const void*)c->tcs : x;
            if (d.cin <= 4 && !c->tcs) { cg_set_error("net_bind: no scratch for the 8-channel input"); return CG_ERR_STATE; }
            static const bool halo_on = [] { const char* e = getenv("CG_DISABLE_HALO"); return !(e && e[0] == '1'); }();   // A/B hook
            auto make_win = [&](TcConvLaunch& Ln, const void* in, int cin_k, const bf16* wmat, int n_out, int p_t, int p_l) -> int {
                TcConvArgs& a = Ln.a;
                memset(&a, 0, sizeof(a));
                const int nch = (k * cin_k + 63) / 64;
                a.bn = n_out; a.n_blocks_n = 1;
                a.nb = c->N; a.out_H = ho; a.out_W = wo; a.Cout = n_out; a.out_sy = a.out_sx = 1;
                a.b_rows_per_tap = n_out;
                a.win_C = cin_k; a.win_k = k; a.win_pl = p_l; a.win_W = wo; a.win_pt = p_t;
                bool dual_unused = false;
                Ln.halo = halo_on && wo % 16 == 0 && ho % 8 == 0 && tc_convw_stages(n_out, k, 16, 8, &dual_unused) >= 2;
                if (Ln.halo) {       // 8 x 16 pixel boxes; a stage = the (8 + k - 1)-row halo of one window chunk
                    a.Wb = 16; a.Hb = 8; a.tiles_w = wo / 16; a.tiles_per_img = a.tiles_w * (ho / 8);
                    a.out_P = wo; a.out_wvalid = wo; a.out_hvalid = ho;
                    a.n_taps = k * nch; a.cchunks = nch;
                    CG_TRY(tc_make_map_win(&Ln.mapA, in, cin_k, k, p_l, wo, ho, c->N, 16, 8 + k - 1));
                } else {
                    set_tiles(a, wo, ho);
                    a.n_taps = k * nch; a.cchunks = 1;
                    for (int kh = 0; kh < k; ++kh)
                        for (int j = 0; j < nch; ++j) {
                            const int st_ = kh * nch + j;
                            a.tb[st_] = (short)st_; a.dc[st_] = (short)(64 * j); a.dh[st_] = (short)(kh - p_t);
                        }
                    CG_TRY(tc_make_map_win(&Ln.mapA, in, cin_k, k, p_l, wo, ho, c->N, a.Wb, a.Hb));
                }
                CG_TRY(tc_make_map_2d(&Ln.mapB, wmat, 64, k * nch * n_out, n_out));
                Ln.mapB2 = Ln.mapB;
                return CG_OK;
            };
            t.fwd.assign(1, TcConvLaunch());
            CG_TRY(make_win(t.fwd[0], xs, cp, wf, d.cout, pt, pl));
            if (!c->bwd) continue;
            if (L.pk_d >= 0) {
                t.dgrad.assign(1, TcConvLaunch());
                CG_TRY(make_win(t.dgrad[0], dy, d.cout, wd, d.cin, k - 1 - pt, k - 1 - pl));
            }
            const char* now = getenv("CG_DISABLE_WGRADW");       // test hook: CUDA-core weight gradient for these layers
            if (!(now && now[0] == '1') && ((wo % 64 == 0) || (64 % wo == 0 && ho % (64 / wo) == 0))) {
                TcWgradWArgs& a = t.waw;
                memset(&a, 0, sizeof(a));
                a.k = k; a.C = cp; a.Creal = d.cin; a.Cout = d.cout;
                a.nch = (k * cp + 63) / 64; a.steps = k * a.nch;
                a.W = wi; a.pl = pl; a.pt = pt;
                t.wgh = halo_on && wo % 16 == 0 && ho % 4 == 0 && tc_wgradh_ok(k, d.cout) &&
                        (4 + k - 1) * 16 * 128 + (d.cout / 16) * 2048 <= 100 * 1024;
                if (t.wgh) { a.Wk = 16; a.Hk = 4; }
                else { a.Wk = wo % 64 == 0 ? 64 : wo; a.Hk = 64 / a.Wk; }
                a.chunks_w = wo / a.Wk; a.chunks_per_img = a.chunks_w * (ho / a.Hk);
                for (int kh = 0; kh < k; ++kh)
                    for (int j = 0; j < a.nch; ++j) { a.dc[kh * a.nch + j] = (short)(64 * j); a.dh[kh * a.nch + j] = (short)(kh - pt); }
                CG_TRY(tc_make_map_win(&t.mapXw, xs, cp, k, pl, wi, hi, c->N, a.Wk, t.wgh ? a.Hk + k - 1 : a.Hk));
                CG_TRY(tc_make_map_act16(&t.mapDYw, dy, d.cout, wo, ho, c->N, a.Wk, a.Hk, d.cout / 16));
                t.wgw = true;
            }
        } else if (L.tc == TC_IM2COL) {
            if (!c->tcs) { cg_set_error("net_bind: no scratch for the unfolded input"); return CG_ERR_STATE; }
            const void* U = c->tcs;                         // [N][ho][wo][64]: the receptive field of every output pixel
            t.fwd.assign(1, TcConvLaunch());
            {
                TcConvArgs& a = t.fwd[0].a;
                memset(&a, 0, sizeof(a));
                set_tiles(a, wo, ho);
                a.n_taps = 1; a.cchunks = 1; a.bn = d.cout; a.n_blocks_n = 1;
                a.nb = c->N; a.out_H = ho; a.out_W = wo; a.Cout = d.cout; a.out_sy = a.out_sx = 1;
                a.b_rows_per_tap = d.cout;
                CG_TRY(tc_make_map_act(&t.fwd[0].mapA, U, 64, wo, ho, c->N, 0, a.Wb, a.Hb));
                CG_TRY(tc_make_map_2d(&t.fwd[0].mapB, wf, 64, d.cout, a.bn));
                CG_TRY(tc_make_map_2d(&t.fwd[0].mapB2, wf, 64, d.cout, (a.bn) / 2 >= 8 ? (a.bn) / 2 : 8));
            }
            if (!c->bwd) continue;
            {   // dW[j][co] = sum_pixels U[p][j] * dy[p][co]: one unit, the second 64 MMA rows are out of bounds (zero)
                TcWgradArgs& a = t.wa;
                memset(&a, 0, sizeof(a));
                a.Wk = wo % 64 == 0 ? 64 : wo; a.Hk = 64 / a.Wk;
                a.n_taps = 1; a.stack2 = 1; a.transposed = 0; a.a_blocks = 1; a.bn = d.cout; a.b_blocks = 1;
                a.chunks_w = wo / a.Wk; a.chunks_per_img = a.chunks_w * (ho / a.Hk);
                a.Cin = 128; a.Cout = d.cout;
                a.dh[0] = 0; a.dh[1] = 20000;
                a.x_grouped = 0; a.y_grouped = 1;
                CG_TRY(tc_make_map_act(&t.mapXw, U, 64, wo, ho, c->N, 0, a.Wk, a.Hk));
                CG_TRY(tc_make_map_act_grouped(&t.mapDYw, dy, d.cout, wo, ho, c->N, a.Wk, a.Hk, d.cout / 64));
            }
            t.dgrad.assign(1, TcConvLaunch());
            {   // dU[p][j] = sum_co dy[p][co] * W[j][co]   (written to the scratch, then folded back by col2im)
                TcConvArgs& a = t.dgrad[0].a;
                memset(&a, 0, sizeof(a));
                set_tiles(a, wo, ho);
                a.n_taps = 1; a.cchunks = d.cout / 64; a.bn = 64; a.n_blocks_n = 1;
                a.nb = c->N; a.out_H = ho; a.out_W = wo; a.Cout = 64; a.out_sy = a.out_sx = 1;
                a.b_rows_per_tap = 64;
                CG_TRY(tc_make_map_act(&t.dgrad[0].mapA, dy, d.cout, wo, ho, c->N, 0, a.Wb, a.Hb));
                CG_TRY(tc_make_map_2d(&t.dgrad[0].mapB, wd, d.cout, 64, 64));
                CG_TRY(tc_make_map_2d(&t.dgrad[0].mapB2, wd, d.cout, 64, 32));
            }
        } else if (L.tc == TC_STEM) {
            if (!c->tcs) { cg_set_error("net_bind: no scratch for the unfolded stem input"); return CG_ERR_STATE; }
            const void* U = c->tcs;                         // [N][hi][wo][64]: U[r][ow][(j*k+kw)*cin+ci] = x[r+j][ow+kw][ci], j < 3
            const int nt = (k + 2) / 3;                     // vertical taps left after the unfolding: row offsets 0, 3, 6
            t.fwd.assign(1, TcConvLaunch());
            {
                TcConvArgs& a = t.fwd[0].a;
                memset(&a, 0, sizeof(a));
                set_tiles(a, wo, ho);
                a.n_taps = nt; a.cchunks = 1; a.bn = d.cout; a.n_blocks_n = 1;
                a.nb = c->N; a.out_H = ho; a.out_W = wo; a.Cout = d.cout; a.out_sy = a.out_sx = 1;
                a.b_rows_per_tap = d.cout;
                for (int tp = 0; tp < nt; ++tp) { a.tb[tp] = (short)tp; a.dh[tp] = (short)(3 * tp); }
                CG_TRY(tc_make_map_act(&t.fwd[0].mapA, U, 64, wo, hi, c->N, 0, a.Wb, a.Hb));
                CG_TRY(tc_make_map_2d(&t.fwd[0].mapB, wf, 64, nt * d.cout, a.bn));
                CG_TRY(tc_make_map_2d(&t.fwd[0].mapB2, wf, 64, nt * d.cout, (a.bn) / 2 >= 8 ? (a.bn) / 2 : 8));
            }
            if (!c->bwd) continue;
            {   // dWv[t][64 rows][co] = sum_pixels U[oh+3t][ow][:] (x) dy[oh][ow][co]; two taps share the 128 MMA rows
                TcWgradArgs& a = t.wa;
                memset(&a, 0, sizeof(a));
                a.Wk = wo % 64 == 0 ? 64 : wo; a.Hk = 64 / a.Wk;
                a.n_taps = (nt + 1) / 2; a.stack2 = 1; a.transposed = 0; a.a_blocks = 1; a.bn = d.cout; a.b_blocks = 1;
                a.chunks_w = wo / a.Wk; a.chunks_per_img = a.chunks_w * (ho / a.Hk);
                a.Cin = 128; a.Cout = d.cout;
                for (int tp = 0; tp < 2 * a.n_taps; ++tp) a.dh[tp] = (short)(tp < nt ? 3 * tp : 20000);   // odd tap count: the last half is out of bounds = 0
                a.x_grouped = 0; a.y_grouped = 1;
                CG_TRY(tc_make_map_act(&t.mapXw, U, 64, wo, hi, c->N, 0, a.Wk, a.Hk));
                CG_TRY(tc_make_map_act_grouped(&t.mapDYw, dy, d.cout, wo, ho, c->N, a.Wk, a.Hk, d.cout / 64));
            }
            // data gradient (only needed when the stem's input is itself a generated image: the cycle calls):
            //   S'[ih][q][kw*cin+ci] = sum_kh dy[ih-kh][q][:] . Wsd[kh][kw*cin+ci][:]  (flat mode, 7 vertical taps), then a
            //   horizontal diagonal sum; S' lives in the scratch
            t.dgrad.assign(1, TcConvLaunch());
            {
                TcConvArgs& a = t.dgrad[0].a;
                memset(&a, 0, sizeof(a));
                a.Wb = 128; a.Hb = 1;
                a.n_taps = k; a.cchunks = d.cout / 64; a.bn = 32; a.n_blocks_n = 1;
                a.tiles_per_img = (hi * wo + 127) / 128; a.tiles_w = a.tiles_per_img;
                a.nb = c->N;
                a.out_P = wo; a.out_wvalid = wo; a.out_hvalid = hi; a.out_H = hi; a.out_W = wo; a.Cout = 32;
                a.out_sy = a.out_sx = 1;
                a.b_rows_per_tap = 32;
                for (int kh = 0; kh < k; ++kh) { a.tb[kh] = (short)kh; a.dw[kh] = (short)(-kh * wo); }
                CG_TRY(tc_make_map_act(&t.dgrad[0].mapA, dy, d.cout, ho * wo, 1, c->N, 0, 128, 1));
                CG_TRY(tc_make_map_2d(&t.dgrad[0].mapB, wd, d.cout, k * 32, 32));
                CG_TRY(tc_make_map_2d(&t.dgrad[0].mapB2, wd, d.cout, k * 32, 16));
            }
        } else if (L.tc == TC_HEAD) {
            if (!c->tcs) { cg_set_error("net_bind: no scratch for the unfolded head tensors"); return CG_ERR_STATE; }
            const void* S = c->tcs;                         // forward:  [N][ho][wi][32]
            const void* T = c->tcs;                         // backward: [N][ho+2][wi][64], T[r][q][(j*k+kw)*cout+co] = dy[r-j][q-kw][co]
            t.fwd.assign(1, TcConvLaunch());
            {
                TcConvArgs& a = t.fwd[0].a;
                memset(&a, 0, sizeof(a));
                a.Wb = 128; a.Hb = 1;
                a.n_taps = k; a.cchunks = d.cin / 64; a.bn = 32; a.n_blocks_n = 1;
                a.tiles_per_img = (ho * wi + 127) / 128; a.tiles_w = a.tiles_per_img;
                a.nb = c->N;
                a.out_P = wi; a.out_wvalid = wi; a.out_hvalid = ho; a.out_H = ho; a.out_W = wi; a.Cout = 32;
                a.out_sy = a.out_sx = 1;
                a.b_rows_per_tap = 32;
                for (int kh = 0; kh < k; ++kh) { a.tb[kh] = (short)kh; a.dw[kh] = (short)(kh * wi); }
                CG_TRY(tc_make_map_act(&t.fwd[0].mapA, x, d.cin, hi * wi, 1, c->N, 0, 128, 1));
                CG_TRY(tc_make_map_2d(&t.fwd[0].mapB, wf, d.cin, k * 32, 32));
                CG_TRY(tc_make_map_2d(&t.fwd[0].mapB2, wf, d.cin, k * 32, (32) / 2 >= 8 ? (32) / 2 : 8));
                (void)S;
            }
            if (!c->bwd) continue;
            const int nt = (k + 2) / 3, Hd = ho + 2;
            t.dgrad.assign(1, TcConvLaunch());
            {   // dxp[ih][q][ci] = sum_t T[ih-3t][q][:] . Whd[t][ci][:]   (flat mode over the padded input grid)
                TcConvArgs& a = t.dgrad[0].a;
                memset(&a, 0, sizeof(a));
                a.Wb = 128; a.Hb = 1;
                a.n_taps = nt; a.cchunks = 1; a.bn = d.cin; a.n_blocks_n = 1;
                a.tiles_per_img = (hi * wi + 127) / 128; a.tiles_w = a.tiles_per_img;
                a.nb = c->N;
                a.out_P = wi; a.out_wvalid = wi; a.out_hvalid = hi; a.out_H = hi; a.out_W = wi; a.Cout = d.cin;
                a.out_sy = a.out_sx = 1;
                a.b_rows_per_tap = d.cin;
                for (int tp = 0; tp < nt; ++tp) { a.tb[tp] = (short)tp; a.dw[tp] = (short)(-3 * tp * wi); }
                CG_TRY(tc_make_map_act(&t.dgrad[0].mapA, T, 64, Hd * wi, 1, c->N, 0, 128, 1));
                CG_TRY(tc_make_map_2d(&t.dgrad[0].mapB, wd, 64, nt * d.cin, a.bn));
                CG_TRY(tc_make_map_2d(&t.dgrad[0].mapB2, wd, 64, nt * d.cin, (a.bn) / 2 >= 8 ? (a.bn) / 2 : 8));
            }
            {   // dWh[t][64 rows][ci] = sum_{r,q} T[r-3t][q][:] (x) xp[r][q][ci] over the padded input grid; the last 64-pixel
                // chunk of a row is partly out of bounds: zero-filled in BOTH operands
                TcWgradArgs& a = t.wa;
                memset(&a, 0, sizeof(a));
                a.Wk = 64; a.Hk = 1;
                a.n_taps = (nt + 1) / 2; a.stack2 = 1; a.transposed = 0; a.a_blocks = 1; a.bn = d.cin; a.b_blocks = 1;
                a.chunks_w = (wi + 63) / 64; a.chunks_per_img = a.chunks_w * hi;
                a.Cin = 128; a.Cout = d.cin;
                for (int tp = 0; tp < 2 * a.n_taps; ++tp) a.dh[tp] = (short)(tp < nt ? -3 * tp : 20000);
                a.x_grouped = 0; a.y_grouped = 1;
                CG_TRY(tc_make_map_act(&t.mapXw, T, 64, wi, Hd, c->N, 0, 64, 1));
                CG_TRY(tc_make_map_act_grouped(&t.mapDYw, x, d.cin, wi, hi, c->N, 64, 1, d.cin / 64));
            }
        } else if (L.tc == TC_CONVT_S2) {
            // F: (ho x wo x cout) -> (hi x wi x cin), stride 2, 'same' padding computed on the big grid
            int pt, pl;
            same_pad(ho, k, 2, &pt);
            same_pad(wo, k, 2, &pl);
            CG_TRY(make_class_launches(t.fwd, x, hi, wi, d.cin, c->N, wd, d.cout, k, pt, pl, ho, wo));
            if (!c->bwd) continue;
            t.dgrad.assign(1, TcConvLaunch());
            CG_TRY(make_conv_launch(t.dgrad[0], dy, ho, wo, d.cout, c->N, wf, d.cin, k, 2, pt, pl, hi, wi));
            CG_TRY(make_wgrad(t, dy, ho, wo, d.cout, x, hi, wi, d.cin, c->N, k, 2, pt, pl, 0));
            t.wg_x_is_dy = 1;
        }
    }
    return CG_OK;
}

// ------------------------------------------------------------------------------------------
static DropKey drop_key(const CallCtx* c, const LayerInfo& L) {
    DropKey k;
    k.seed = c->net->seed; k.ctr_host = c->drop_ctr_host; k.ctr_dev = c->drop_ctr_dev;
    for (int g = 0; g < 4; ++g) k.call_id[g] = c->call_id[g];
    k.layer = L.drop_index;
    return k;
}

int net_update_moving(CallCtx* c, int g, cudaStream_t st) {
    const cg_net_s* net = c->net;
    for (size_t i = 0; i < net->layers.size(); ++i) {
        const LayerInfo& L = net->layers[i];
        if (!L.batch) continue;
        if (!net->state) { cg_set_error("BatchNormalization without bound state"); return CG_ERR_STATE; }
        const float* bstat = (const float*)(c->base + c->bstat_off[i]) + (size_t)g * L.d.cin * 2;
        CG_TRY(k_bn_update_moving(net->state + L.mm_off, net->state + L.mv_off, bstat, L.d.cin, L.d.momentum, st));
    }
    return CG_OK;
}

template <typename T>
static int forward_T(CallCtx* c, const float* params, cudaStream_t st) {
    const cg_net_s* net = c->net;
    const int N = c->N;
    std::vector<char> stats_done(net->layers.size() + 1, 0), fused_done(net->layers.size() + 1, 0);
    std::vector<char> cat_done(net->layers.size() + 1, 0);      // CONCAT layer: bit 0 / 1 = in0 / in1 already in place
    bool upsample_fused = false;
    c->live.assign(net->layers.size() + 1, 0);
    c->live[0] = 1;
    // every statistics table of this call is zeroed by ONE memset (they are accumulated into by atomics)
    if (c->act_bytes > c->stat_begin) CG_TRY(zero_region(c->base + c->stat_begin, c->act_bytes - c->stat_begin, st));
    for (size_t i = 0; i < net->layers.size(); ++i) {
        const LayerInfo& L = net->layers[i];
        if (L.skipped) continue;
        const cg_layer_desc& d = L.d;
        const int tin = d.in0, tout = L.out_t;
        const T* x = (const T*)c->act(tin);
        // a tensor-core conv that feeds only an instance norm accumulates that norm's statistics in its epilogue
        float* fused_stats = nullptr;
        // ... when its main loop is long enough to hide the extra epilogue work (the column sums double the epilogue of a
        // 32-column chunk); the 3-K-step stem is faster with the separate streaming statistics pass
        const bool long_k = L.tc != TC_STEM && L.tc != TC_IM2COL;
        const bool bn_infer = i + 1 < net->layers.size() && net->layers[i + 1].batch && !c->training;
        if (c->tc[i].on && L.feeds_in && L.tc != TC_HEAD && long_k && i + 1 < net->layers.size() && !bn_infer) {
            // plain instance norm: the raw sums go to their own table and the streaming apply finalizes them on the fly
            fused_stats = (float*)(c->base + (net->layers[i + 1].batch ? c->stat_off[i + 1] : c->raw_off[i + 1]));
            stats_done[i + 1] = 1;
        }
        T* y = (T*)c->act(tout);
        const int h = c->th[tin], w = c->tw[tin], oh = c->th[tout], ow = c->tw[tout];
        switch (d.op) {
            case CG_OP_CONV: {
                ConvGeom g = conv_geom(d, N, h, w, oh, ow);
                const float* bias = L.b_off >= 0 ? params + L.b_off : nullptr;
                const double fl = 2.0 * N * oh * ow * (double)d.cout * d.k * d.k * d.cin;
                if (c->tc[i].on && L.tc == TC_S1_WIN) {
                    const TcConvLaunch& tl = c->tc[i].fwd[0];
                    if (d.cin <= 4) CG_TRY(sp_pad_channels8((const bf16*)x, (bf16*)c->tcs, (size_t)N * h * w, d.cin, st));
                    TcConvArgs a = tl.a;
                    a.nb = N;
                    a.stats = fused_stats;
                    if (tl.halo) CG_TRY(tc_convw_launch(&tl.mapA, &tl.mapB, (bf16*)y, bias, a, fl, st));
                    else CG_TRY(tc_conv_launch(&tl.mapA, &tl.mapB, &tl.mapB2, (bf16*)y, bias, a, fl, st));
                } else if (c->tc[i].on && L.tc == TC_IM2COL) {
                    const TcConvLaunch& tl = c->tc[i].fwd[0];
                    CG_TRY(sp_im2col((const bf16*)x, (bf16*)c->tcs, N, h, w, d.cin, oh, ow, d.k, d.stride, g.pt, g.pl, st));
                    TcConvArgs a = tl.a;
                    a.nb = N;
                    CG_TRY(tc_conv_launch(&tl.mapA, &tl.mapB, &tl.mapB2, (bf16*)y, bias, a, fl, st));
                } else if (c->tc[i].on && L.tc == TC_STEM) {
                    const TcConvLaunch& tl = c->tc[i].fwd[0];
                    CG_TRY(sp_unfold_w((const bf16*)x, (bf16*)c->tcs, N, h, w, d.cin, h, ow, d.k, +1, st));
                    TcConvArgs a = tl.a;
                    a.nb = N;
                    a.stats = fused_stats;
                    CG_TRY(tc_conv_launch(&tl.mapA, &tl.mapB, &tl.mapB2, (bf16*)y, bias, a, fl, st));
                } else if (c->tc[i].on && L.tc == TC_HEAD) {
                    const TcConvLaunch& tl = c->tc[i].fwd[0];
                    TcConvArgs a = tl.a;
                    a.nb = N;
                    CG_TRY(tc_conv_launch(&tl.mapA, &tl.mapB, &tl.mapB2, (bf16*)c->tcs, nullptr, a, fl, st));
                    CG_TRY(sp_diag_sum((const bf16*)c->tcs, bias, (bf16*)y, N, oh, ow, w, d.k, d.cout, +1, st));
                } else if (c->tc[i].on) {
                    for (const TcConvLaunch& tl : c->tc[i].fwd) {
                        TcConvArgs a = tl.a;
                        a.nb = N;
                        a.stats = fused_stats;
                        CG_TRY(tc_conv_launch(&tl.mapA, &tl.mapB, &tl.mapB2, (bf16*)y, bias, a, tl.flop_share * fl, st));
                    }
                } else {
                    CG_TRY(k_conv_fwd<T>(x, params + L.w_off, bias, y, g, 0, st));
                }
                break;
            }
            case CG_OP_CONVT: {
                ConvGeom g = conv_geom(d, N, h, w, oh, ow);
                if (c->tc[i].on) {
                    for (const TcConvLaunch& tl : c->tc[i].fwd) {
                        TcConvArgs a = tl.a;
                        a.nb = N;
                        a.stats = fused_stats;
                        CG_TRY(tc_conv_launch(&tl.mapA, &tl.mapB, &tl.mapB2, (bf16*)y, L.b_off >= 0 ? params + L.b_off : nullptr, a,
                                              tl.flop_share * 2.0 * N * h * w * (double)d.cout * d.k * d.k * d.cin, st));
                    }
                } else {
                    CG_TRY(k_conv_dgrad<T>(x, params + L.w_off, L.b_off >= 0 ? params + L.b_off : nullptr, y, g, 0, st));
                }
                break;
            }
            case CG_OP_INORM: {
                float* stats = (float*)(c->base + c->stat_off[i]);
                if (L.batch) {
                    if (!net->state) { cg_set_error("layer %d: BatchNormalization without bound state (cg_net_bind_state)", (int)i); return CG_ERR_STATE; }
                    float* mm = net->state + L.mm_off;
                    float* mv = net->state + L.mv_off;
                    if (c->training) {
                        const int grp = c->bn_group > 0 ? c->bn_group : N;
                        if (N % grp) { cg_set_error("batch %d is not a multiple of the call size %d", N, grp); return CG_ERR_INVALID; }
                        if (!stats_done[i]) CG_TRY(k_in_stats_raw<T>(x, stats, N, h * w, d.cin, st, /*zeroed=*/true));
                        float* bstat = (float*)(c->base + c->bstat_off[i]);
                        CG_TRY(k_bn_finalize(stats, bstat, N, d.cin, grp, h * w, d.eps, st));
                        if (!c->defer_moving)
                            for (int g = 0; g < N / grp; ++g)
                                CG_TRY(k_bn_update_moving(mm, mv, bstat + (size_t)g * d.cin * 2, d.cin, d.momentum, st));
                    } else {
                        CG_TRY(k_bn_fill(stats, mm, mv, N, d.cin, d.eps, st));
                    }
                } else if (!stats_done[i]) CG_TRY(k_in_stats<T>(x, stats, N, h * w, d.cin, d.eps, st, /*zeroed=*/true));
                // raw sums from the conv epilogue (plain instance norm): the streaming apply kernels finalize them themselves
                // (and store (mean, rstd) for the backward); every other path runs the finalize kernel first
                const float* raw = (!L.batch && stats_done[i]) ? (const float*)(c->base + c->raw_off[i]) : nullptr;
                auto finalize_now = [&]() -> int {
                    if (!raw) return CG_OK;
                    const int rc = k_in_finalize(raw, stats, N * d.cin, h * w, d.eps, st);
                    raw = nullptr;
                    return rc;
                };
                const float* gam = L.g_off >= 0 ? params + L.g_off : nullptr;
                const float* bet = L.be_off >= 0 ? params + L.be_off : nullptr;
                if (L.fuse_rpad >= 0) {
                    const LayerInfo& R = net->layers[L.fuse_rpad];
                    T* yp = (T*)c->act(R.out_t);
                    if (R.d.pad < h && R.d.pad < w && k_in_stream_ok<T>(x, yp, nullptr, h * w, d.cin)) {
                        CG_TRY(k_in_apply_stream<T>(x, nullptr, nullptr, yp, stats, gam, bet, L.fused_act, L.fused_slope, N, h * w,
                                                    d.cin, w, R.d.pad, st, raw, d.eps));
                        fused_done[L.fuse_rpad] = 1;
                        c->live[R.out_t] = 1;
                        break;
                    }
                }
                if (L.fuse_add >= 0) {
                    const LayerInfo& Ad = net->layers[L.fuse_add];
                    const int other = Ad.d.in0 == tout ? Ad.d.in1 : Ad.d.in0;
                    const T* res = (const T*)c->act(other);
                    T* ys = (T*)c->act(Ad.out_t);
                    T* yp = nullptr;
                    int pad = 0;
                    if (Ad.fuse_rpad >= 0 && net->layers[Ad.fuse_rpad].d.pad < h && net->layers[Ad.fuse_rpad].d.pad < w) {
                        yp = (T*)c->act(net->layers[Ad.fuse_rpad].out_t);
                        pad = net->layers[Ad.fuse_rpad].d.pad;
                    }
                    if (other <= (int)i && k_in_stream_ok<T>(x, res, ys, h * w, d.cin) && !((uintptr_t)yp & 15)) {
                        CG_TRY(k_in_apply_stream<T>(x, res, ys, yp, stats, gam, bet, L.fused_act, L.fused_slope, N, h * w, d.cin, w,
                                                    pad, st, raw, d.eps));
                        fused_done[L.fuse_add] = 1;
                        c->live[Ad.out_t] = 1;
                        if (yp) { fused_done[Ad.fuse_rpad] = 1; c->live[net->layers[Ad.fuse_rpad].out_t] = 1; }
                        break;
                    }
                }
                if (raw && k_in_stream_ok<T>(x, y, nullptr, h * w, d.cin)) {
                    CG_TRY(k_in_apply_stream<T>(x, nullptr, y, nullptr, stats, gam, bet, L.fused_act, L.fused_slope, N, h * w,
                                                d.cin, 0, 0, st, raw, d.eps));
                } else {
                    CG_TRY(finalize_now());
                    CG_TRY(k_in_apply<T>(x, y, stats, gam, bet, L.fused_act, L.fused_slope, N, h * w, d.cin, st));
                }
                c->live[tout] = 1;
                break;
            }
            case CG_OP_ACT:
                CG_TRY(k_act_fwd<T>(x, y, (size_t)N * c->sample_elems(tin), d.act, d.slope, st));
                break;
            case CG_OP_RPAD:
                if (!fused_done[i]) { CG_TRY(k_rpad_fwd<T>(x, y, N, h, w, d.cin, d.pad, st)); c->live[tout] = 1; }
                break;
            case CG_OP_ADD:
                if (!fused_done[i]) {
                    CG_TRY(k_add<T>(x, (const T*)c->act(d.in1), y, (size_t)N * c->sample_elems(tin), st));
                    c->live[tout] = 1;
                }
                break;
            case CG_OP_CONCAT: {
                int ca = net->chan[d.in0], cb = net->chan[d.in1];
                size_t npix = (size_t)N * h * w;
                // zero-copy halves: the pool that reads the skip / the upsample that produces x have already written them
                if (!(cat_done[i] & 1)) CG_TRY(k_slice_copy<T>(x, ca, 0, y, ca + cb, 0, ca, npix, 0, st));
                if (!(cat_done[i] & 2)) CG_TRY(k_slice_copy<T>((const T*)c->act(d.in1), cb, 0, y, ca + cb, ca, cb, npix, 0, st));
                break;
            }
            case CG_OP_AVGPOOL:
                if (L.cat_layer >= 0) {        // also writes the copy of its input into the concat tensor (channels [0, cin))
                    const LayerInfo& Cc = net->layers[L.cat_layer];
                    CG_TRY(k_avgpool_fwd<T>(x, y, N, h, w, d.cin, st, (T*)c->act(Cc.out_t), net->chan[Cc.out_t], 0));
                    cat_done[L.cat_layer] |= 1;
                } else CG_TRY(k_avgpool_fwd<T>(x, y, N, h, w, d.cin, st));
                break;
            case CG_OP_UPSAMPLE:
                if (L.cat_layer >= 0) {        // writes straight into its slice of the concat tensor; its own tensor is not materialised
                    const LayerInfo& Cc = net->layers[L.cat_layer];
                    CG_TRY(k_upsample_fwd<T>(x, (T*)nullptr, N, h, w, d.cin, st, (T*)c->act(Cc.out_t), net->chan[Cc.out_t],
                                             net->chan[Cc.d.in0]));
                    cat_done[L.cat_layer] |= 2;
                    upsample_fused = true;
                } else CG_TRY(k_upsample_fwd<T>(x, y, N, h, w, d.cin, st));
                break;
            case CG_OP_DROPOUT: {
                const int grp = c->bn_group > 0 ? c->bn_group : N;
                if (N % grp || N / grp > 4) { cg_set_error("dropout: batch %d / call size %d", N, grp); return CG_ERR_INVALID; }
                CG_TRY(k_dropout_fwd<T>(x, y, (size_t)grp * c->sample_elems(tin), N / grp, d.rate, drop_key(c, L), c->training ? 1 : 0, st));
                break;
            }
            default: cg_set_error("unknown op %d", d.op); return CG_ERR_INVALID;
        }
        if (d.op != CG_OP_INORM && d.op != CG_OP_RPAD && d.op != CG_OP_ADD) c->live[tout] = upsample_fused ? 0 : 1;
        upsample_fused = false;
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
                      cudaStream_t st, LayerHook hook, void* hook_user) {
    const cg_net_s* net = c->net;
    const int nl = (int)net->layers.size();
    if (!c->bwd || !c->forwarded) { cg_set_error("backward without a planned forward"); return CG_ERR_STATE; }
    if (n0 < 0 || nb <= 0 || n0 + nb > c->N) { cg_set_error("bad sub-batch [%d,%d) of %d", n0, n0 + nb, c->N); return CG_ERR_INVALID; }
    // per-tensor pointers for this sub-batch
    auto A = [&](int t) -> const T* { return (const T*)c->act(t) + (size_t)n0 * c->sample_elems(t); };
    // galias[t] = u: the gradient of tensor t IS the buffer of tensor u (residual add: d(sum)/d(input) = identity, so the
    // inputs' gradients start out as the sum's gradient; no copy is made when nothing else can touch the shared buffer first)
    std::vector<int> galias(nl + 1, -1);
    auto G = [&](int t) -> T* {
        while (t > 0 && t < nl && galias[t] >= 0) t = galias[t];
        if (t == nl) return const_cast<T*>(dy_out);
        if (t == 0) return dx_in;
        return (T*)(c->arena + c->grad_off[t]);
    };
    std::vector<char> written(nl + 1, 0), fold_done(nl + 1, 0);
    std::vector<int> pend_pool(nl, -1);       // AVGPOOL layer -> CONCAT layer whose skip slice its backward adds (zero-copy concat)
    std::vector<int> up_src(nl, -1);          // UPSAMPLE layer -> CONCAT layer whose gradient slice it gathers from
    bool defer_written = false;
    auto need = [&](int t) -> bool { return dx_in != nullptr || (t != 0 && net->dep_params[t]); };
    // the backward sums of every instance norm of this call are cleared by ONE memset (sub-batch relative tables)
    if (c->sums_bytes) CG_TRY(zero_region(c->arena + c->scratch_off, c->sums_bytes, st));
    // A tensor-core data gradient whose target is a reflection-padded tensor nobody else reads writes the interior of the
    // padded grid straight into the UNPADDED gradient (accumulating onto the skip path's contribution if that is already
    // there); the pad's backward then only folds the thin border.  Returns the RPAD layer index or -1.
    auto find_fold = [&](int i, int tin, int cin) -> int {
        if (tin < 2 || net->n_consumers[tin] != 1 || cin % 8) return -1;
        int prod = -1;
        for (int j = i - 1; j >= 0; --j)
            if (!net->layers[j].skipped && net->layers[j].out_t == tin) { prod = j; break; }
        if (prod < 0 || net->layers[prod].d.op != CG_OP_RPAD) return -1;
        const cg_layer_desc& rd = net->layers[prod].d;
        if (rd.in0 >= 1 && rd.pad > 0 && need(rd.in0) && c->grad_halo[rd.in0] == 0 && c->th[rd.in0] > 2 * rd.pad + 1 &&
            c->tw[rd.in0] > 2 * rd.pad + 1)
            return prod;
        return -1;
    };

    for (int i = nl - 1; i >= 0; --i) {
        const LayerInfo& L = net->layers[i];
        if (L.skipped) continue;
        const cg_layer_desc& d = L.d;
        const int tin = d.in0, tout = L.out_t;
        if ((!need(tout) || !written[tout]) && tout != nl) {      // nothing downstream asked for it
            if (d.op == CG_OP_AVGPOOL && pend_pool[i] >= 0 && need(d.in0)) {     // ... but a concat left its skip slice to this pool
                const LayerInfo& Cc = net->layers[pend_pool[i]];
                const int ca = net->chan[d.in0];
                CG_TRY(k_slice_copy<T>(G(Cc.out_t), net->chan[Cc.out_t], 0, G(d.in0), ca, 0, ca,
                                       (size_t)nb * c->th[d.in0] * c->tw[d.in0], (int)written[d.in0], st));
                written[d.in0] = 1;
            }
            continue;
        }
        const T* dy = G(tout);
        const int h = c->th[tin], w = c->tw[tin], oh = c->th[tout], ow = c->tw[tout];
        const bool want_dx = need(tin);
        T* dx = want_dx ? G(tin) : nullptr;
        const int acc = want_dx ? (int)written[tin] : 0;
        // samples whose gradient w.r.t. `tin` anyone reads: all of them when a parameter lies upstream of tin, else only the
        // first dx_nb the caller asked for (CallCtx::dx_nb; used by the 7x7 stem and the reflection pad in front of it)
        const int nbx = (c->dx_nb > 0 && n0 == 0 && c->dx_nb < nb && !net->dep_params[tin]) ? c->dx_nb : nb;
        switch (d.op) {
            case CG_OP_CONV: {
                ConvGeom g = conv_geom(d, nb, h, w, oh, ow);
                if (c->tc[i].on && L.tc == TC_S1_16) {
                    if (grads) {
                        if (c->tc[i].wg16) {
                            TcWgrad16Args a = c->tc[i].wa16;
                            a.n0 = n0; a.nb = nb;
                            CG_TRY(tc_wgrad16_launch(&c->tc[i].mapXw, &c->tc[i].mapDYw, grads + L.w_off, a,
                                                     2.0 * nb * oh * ow * (double)d.cout * d.k * d.k * d.cin, st));
                        } else {
                            CG_TRY(k_conv_wgrad<T>(A(tin), dy, grads + L.w_off, g, st));
                        }
                        if (L.b_off >= 0 && !L.bias_grad_zero)
                            CG_TRY(k_colsum<T>(dy, grads + L.b_off, (size_t)nb * oh * ow, d.cout, st));
                    }
                    if (want_dx) {
                        if (acc) {
                            CG_TRY(k_conv_dgrad<T>(dy, params + L.w_off, nullptr, dx, g, acc, st));
                        } else {
                            const TcConvLaunch& tl = c->tc[i].dgrad[0];
                            TcConvArgs a = tl.a;
                            a.nb = nb;
                            CG_TRY(tc_conv_launch(&tl.mapA, &tl.mapB, &tl.mapB2, (bf16*)dx, nullptr, a,
                                                  2.0 * nb * oh * ow * (double)d.cout * d.k * d.k * d.cin, st));
                        }
                    }
                    break;
                }
                if (c->tc[i].on && L.tc == TC_S1_WIN) {
                    const double fl = 2.0 * nb * oh * ow * (double)d.cout * d.k * d.k * d.cin;
                    if (grads) {
                        if (c->tc[i].wgw) {
                            if (d.cin <= 4) CG_TRY(sp_pad_channels8((const bf16*)A(tin), (bf16*)c->tcs, (size_t)nb * h * w, d.cin, st));
                            TcWgradWArgs a = c->tc[i].waw;
                            a.n0 = d.cin <= 4 ? 0 : n0; a.y_n0 = 0; a.nb = nb;      // dY lives in the (sub-batch relative) arena
                            if (c->tc[i].wgh) CG_TRY(tc_wgradh_launch(&c->tc[i].mapXw, &c->tc[i].mapDYw, grads + L.w_off, a, fl, st));
                            else CG_TRY(tc_wgradw_launch(&c->tc[i].mapXw, &c->tc[i].mapDYw, grads + L.w_off, a, fl, st));
                        } else {
                            CG_TRY(k_conv_wgrad<T>(A(tin), dy, grads + L.w_off, g, st));
                        }
                        if (L.b_off >= 0 && !L.bias_grad_zero)
                            CG_TRY(k_colsum<T>(dy, grads + L.b_off, (size_t)nb * oh * ow, d.cout, st));
                    }
                    if (want_dx) {
                        if (acc || c->tc[i].dgrad.empty()) {     // image-side layer: only the samples whose dx is read (dx_nb)
                            CG_TRY(k_conv_dgrad<T>(dy, params + L.w_off, nullptr, dx, conv_geom(d, nbx, h, w, oh, ow), acc, st));
                        } else {
                            const TcConvLaunch& tl = c->tc[i].dgrad[0];
                            TcConvArgs a = tl.a;
                            a.nb = nb;
                            if (tl.halo) CG_TRY(tc_convw_launch(&tl.mapA, &tl.mapB, (bf16*)dx, nullptr, a, fl, st));
                            else CG_TRY(tc_conv_launch(&tl.mapA, &tl.mapB, &tl.mapB2, (bf16*)dx, nullptr, a, fl, st));
                        }
                    }
                    break;
                }
                if (c->tc[i].on && !acc && L.tc == TC_IM2COL) {
                    const double fl = 2.0 * nb * oh * ow * (double)d.cout * d.k * d.k * d.cin;
                    float* tmp = (float*)(c->tcs + c->tc[i].sc_tmp);
                    bf16* big = (bf16*)c->tcs;
                    if (grads) {
                        CG_TRY(sp_im2col((const bf16*)A(tin), big, nb, h, w, d.cin, oh, ow, d.k, d.stride, g.pt, g.pl, st));
                        CG_CUDA(cudaMemsetAsync(tmp, 0, (size_t)2 * 64 * d.cout * sizeof(float), st));
                        TcWgradArgs a = c->tc[i].wa;
                        a.n0 = 0; a.nb = nb;
                        CG_TRY(tc_wgrad_launch(&c->tc[i].mapXw, &c->tc[i].mapDYw, tmp, a, fl, st));
                        CG_TRY(sp_unpack_im2col(tmp, grads + L.w_off, d.k, d.cin, d.cout, st));
                        if (L.b_off >= 0 && !L.bias_grad_zero) CG_TRY(k_colsum<T>(dy, grads + L.b_off, (size_t)nb * oh * ow, d.cout, st));
                    }
                    if (want_dx) {
                        const TcConvLaunch& tl = c->tc[i].dgrad[0];
                        TcConvArgs a = tl.a;
                        a.nb = nb;
                        CG_TRY(tc_conv_launch(&tl.mapA, &tl.mapB, &tl.mapB2, big, nullptr, a, fl, st));
                        CG_TRY(sp_col2im(big, (bf16*)dx, nb, h, w, d.cin, oh, ow, d.k, d.stride, g.pt, g.pl, st));
                    }
                    break;
                }
                if (c->tc[i].on && !acc && (L.tc == TC_STEM || L.tc == TC_HEAD)) {
                    const double fl = 2.0 * nb * oh * ow * (double)d.cout * d.k * d.k * d.cin;
                    float* tmp = (float*)(c->tcs + c->tc[i].sc_tmp);
                    bf16* big = (bf16*)c->tcs;
                    const int ntp = 2 * (((d.k + 2) / 3 + 1) / 2);
                    if (L.tc == TC_STEM) {
                        if (grads) {
                            CG_TRY(sp_unfold_w((const bf16*)A(tin), big, nb, h, w, d.cin, h, ow, d.k, +1, st));
                            CG_CUDA(cudaMemsetAsync(tmp, 0, (size_t)ntp * 64 * d.cout * sizeof(float), st));
                            TcWgradArgs a = c->tc[i].wa;
                            a.n0 = 0; a.nb = nb;
                            CG_TRY(tc_wgrad_launch(&c->tc[i].mapXw, &c->tc[i].mapDYw, tmp, a, fl, st));
                            CG_TRY(sp_unpack_dw(tmp, grads + L.w_off, d.k, d.cin, d.cout, 0, st));
                            if (L.b_off >= 0 && !L.bias_grad_zero) CG_TRY(k_colsum<T>(dy, grads + L.b_off, (size_t)nb * oh * ow, d.cout, st));
                        }
                        if (want_dx) {
                            const TcConvLaunch& tl = c->tc[i].dgrad[0];
                            TcConvArgs a = tl.a;
                            a.nb = nbx;
                            CG_TRY(tc_conv_launch(&tl.mapA, &tl.mapB, &tl.mapB2, big, nullptr, a, fl * nbx / nb, st));
                            CG_TRY(sp_diag_sum(big, nullptr, (bf16*)dx, nbx, h, w, ow, d.k, d.cin, -1, st));
                        }
                    } else {
                        CG_TRY(sp_unfold_w((const bf16*)dy, big, nb, oh, ow, d.cout, oh + 2, w, d.k, -1, st));
                        if (grads) {
                            CG_CUDA(cudaMemsetAsync(tmp, 0, (size_t)ntp * 64 * d.cin * sizeof(float), st));
                            TcWgradArgs a = c->tc[i].wa;
                            a.n0 = 0; a.y_n0 = n0; a.nb = nb;
                            CG_TRY(tc_wgrad_launch(&c->tc[i].mapXw, &c->tc[i].mapDYw, tmp, a, fl, st));
                            CG_TRY(sp_unpack_dw(tmp, grads + L.w_off, d.k, d.cin, d.cout, 1, st));
                            if (L.b_off >= 0 && !L.bias_grad_zero) CG_TRY(k_colsum<T>(dy, grads + L.b_off, (size_t)nb * oh * ow, d.cout, st));
                        }
                        if (want_dx) {
                            const TcConvLaunch& tl = c->tc[i].dgrad[0];
                            TcConvArgs a = tl.a;
                            a.nb = nb;
                            const int fold = find_fold(i, tin, d.cin);
                            if (fold >= 0) {
                                const cg_layer_desc& rd = net->layers[fold].d;
                                a.out2 = (bf16*)G(rd.in0);
                                a.fold_pad = rd.pad;
                                a.fold_acc = written[rd.in0] ? 1 : 0;
                            }
                            CG_TRY(tc_conv_launch(&tl.mapA, &tl.mapB, &tl.mapB2, (bf16*)dx, nullptr, a, fl, st));
                            if (fold >= 0) {
                                written[net->layers[fold].d.in0] = 1;
                                fold_done[fold] = 1;
                            }
                        }
                    }
                    break;
                }
                if (c->tc[i].on && !acc) {
                    // stride-1 layers: dy lives in a zero-bordered buffer (halo 2) written by the IN backward
                    const int hl = c->grad_halo[tout];
                    const double fl = 2.0 * nb * oh * ow * (double)d.cout * d.k * d.k * d.cin;
                    if (grads) {
                        TcWgradArgs a = c->tc[i].wa;
                        a.n0 = n0; a.nb = nb;
                        CG_TRY(tc_wgrad_launch(&c->tc[i].mapXw, &c->tc[i].mapDYw, grads + L.w_off, a, fl, st));
                        if (L.b_off >= 0 && !L.bias_grad_zero)
                            CG_TRY(k_colsum<T>(dy, grads + L.b_off, (size_t)nb * (oh + 2 * hl) * (ow + 2 * hl), d.cout, st));
                    }
                    if (want_dx) {
                        const int fold = L.tc == TC_S1_VALID ? find_fold(i, tin, d.cin) : -1;
                        for (const TcConvLaunch& tl : c->tc[i].dgrad) {
                            TcConvArgs a = tl.a;
                            a.nb = nb;
                            if (fold >= 0) {
                                const cg_layer_desc& rd = net->layers[fold].d;
                                a.out2 = (bf16*)G(rd.in0);
                                a.fold_pad = rd.pad;
                                a.fold_acc = written[rd.in0] ? 1 : 0;
                            }
                            CG_TRY(tc_conv_launch(&tl.mapA, &tl.mapB, &tl.mapB2, (bf16*)dx, nullptr, a, tl.flop_share * fl, st));
                        }
                        if (fold >= 0) {
                            written[net->layers[fold].d.in0] = 1;
                            fold_done[fold] = 1;
                        }
                    }
                    break;
                }
                if (c->grad_halo[tout]) { cg_set_error("layer %d: zero-bordered dY needs the tensor-core path", i); return CG_ERR_STATE; }
                if (grads) {
                    CG_TRY(k_conv_wgrad<T>(A(tin), dy, grads + L.w_off, g, st));
                    if (L.b_off >= 0 && !L.bias_grad_zero) CG_TRY(k_colsum<T>(dy, grads + L.b_off, (size_t)nb * oh * ow, d.cout, st));
                }
                if (want_dx) CG_TRY(k_conv_dgrad<T>(dy, params + L.w_off, nullptr, dx, conv_geom(d, nbx, h, w, oh, ow), acc, st));
                break;
            }
            case CG_OP_CONVT: {
                ConvGeom g = conv_geom(d, nb, h, w, oh, ow);
                if (c->tc[i].on && !acc) {
                    const double fl = 2.0 * nb * h * w * (double)d.cout * d.k * d.k * d.cin;
                    if (grads) {
                        TcWgradArgs a = c->tc[i].wa;      // X operand = dY (arena, sub-batch relative), dY operand = x (absolute)
                        a.n0 = 0; a.nb = nb; a.y_n0 = n0;
                        CG_TRY(tc_wgrad_launch(&c->tc[i].mapXw, &c->tc[i].mapDYw, grads + L.w_off, a, fl, st));
                        if (L.b_off >= 0 && !L.bias_grad_zero) CG_TRY(k_colsum<T>(dy, grads + L.b_off, (size_t)nb * oh * ow, d.cout, st));
                    }
                    if (want_dx)
                        for (const TcConvLaunch& tl : c->tc[i].dgrad) {
                            TcConvArgs a = tl.a;
                            a.nb = nb;
                            CG_TRY(tc_conv_launch(&tl.mapA, &tl.mapB, &tl.mapB2, (bf16*)dx, nullptr, a, tl.flop_share * fl, st));
                        }
                    break;
                }
                if (grads) {
                    CG_TRY(k_conv_wgrad<T>(dy, A(tin), grads + L.w_off, g, st));
                    if (L.b_off >= 0 && !L.bias_grad_zero) CG_TRY(k_colsum<T>(dy, grads + L.b_off, (size_t)nb * oh * ow, d.cout, st));
                }
                if (want_dx) CG_TRY(k_conv_fwd<T>(dy, params + L.w_off, nullptr, dx, g, acc, st));
                break;
            }
            case CG_OP_INORM: {
                const float* stats = (const float*)(c->base + c->stat_off[i]) + (size_t)n0 * d.cin * 2;
                bool pg = grads && L.g_off >= 0;
                int grp = 0;
                if (L.batch) {
                    if (!c->training) { cg_set_error("layer %d: backward through an inference-mode BatchNormalization", i); return CG_ERR_STATE; }
                    grp = c->bn_group > 0 ? c->bn_group : c->N;
                    if (n0 % grp || nb % grp) { cg_set_error("sub-batch [%d,%d) splits a BatchNormalization call of %d", n0, n0 + nb, grp); return CG_ERR_INVALID; }
                }
                CG_TRY(k_in_bwd<T>(A(tin), dy, dx, stats, L.g_off >= 0 ? params + L.g_off : nullptr,
                                   L.be_off >= 0 ? params + L.be_off : nullptr, pg ? grads + L.g_off : nullptr,
                                   pg ? grads + L.be_off : nullptr, (float*)(c->arena + c->sums_off[i]), L.fused_act,
                                   L.fused_slope, nb, h * w, d.cin, acc, st, c->grad_halo[tin], w, /*zeroed=*/true, grp));
                break;
            }
            case CG_OP_ACT:
                if (want_dx)
                    CG_TRY(k_act_bwd<T>(A(tout), dy, dx, (size_t)nb * c->sample_elems(tin), d.act, d.slope, acc, st));
                break;
            case CG_OP_RPAD:
                if (want_dx && fold_done[i]) CG_TRY(k_rpad_bwd_border<T>(dy, dx, nbx, h, w, d.cin, d.pad, st));
                else if (want_dx) CG_TRY(k_rpad_bwd<T>(dy, dx, nbx, h, w, d.cin, d.pad, acc, st));
                break;
            case CG_OP_ADD: {
                size_t n = (size_t)nb * c->sample_elems(tin);
                // an input whose gradient has no other contribution yet can share the sum's gradient buffer instead of
                // receiving a copy.  One input may always do so; both may when no consumer of either input sits between the
                // other input's producer and this layer (such a consumer would accumulate into the shared buffer before
                // that producer has read it) -- the case of the ResNet block, resnet.py:26-35
                auto producer = [&](int t) { for (int j = i - 1; j >= 0; --j) if (!net->layers[j].skipped && net->layers[j].out_t == t) return j; return -1; };
                auto can_alias = [&](int t) {
                    if (!(tout != nl && t > 0 && t < nl && !written[t] && need(t) && c->grad_halo[t] == 0 && c->grad_halo[tout] == 0 &&
                          galias[t] < 0 && d.in0 != d.in1))
                        return false;
                    const int pr = producer(t);       // tensor-core layers read their dY through TMA maps bound to its own buffer
                    return pr >= 0 && !c->tc[pr].on;
                };
                auto consumer_between = [&](int t, int lo) {      // a consumer of t with index in (lo, i)
                    for (int j = lo + 1; j < i; ++j) {
                        const cg_layer_desc& q = net->layers[j].d;
                        if (!net->layers[j].skipped && (q.in0 == t || ((q.op == CG_OP_ADD || q.op == CG_OP_CONCAT) && q.in1 == t))) return true;
                    }
                    return false;
                };
                bool a0 = want_dx && can_alias(tin), a1 = need(d.in1) && can_alias(d.in1);
                if (a0 && a1 && (consumer_between(tin, producer(d.in1)) || consumer_between(d.in1, producer(tin)))) a0 = false;
                if (a0) galias[tin] = tout;
                else if (want_dx) CG_TRY(k_copy_acc<T>(dy, dx, n, acc, st));
                if (need(d.in1)) {
                    if (a1) galias[d.in1] = tout;
                    else CG_TRY(k_copy_acc<T>(dy, G(d.in1), n, (int)written[d.in1] || (d.in1 == tin && want_dx), st));
                    written[d.in1] = 1;
                }
                break;
            }
            case CG_OP_CONCAT: {
                int ca = net->chan[d.in0], cb = net->chan[d.in1];
                size_t npix = (size_t)nb * h * w;
                const bool plain_grads = c->grad_halo[tout] == 0 && c->grad_halo[tin] == 0 && c->grad_halo[d.in1] == 0;
                if (want_dx) {
                    if (L.cat_in0_pool >= 0 && plain_grads) {      // the pool's backward adds this slice when it writes dSkip
                        pend_pool[L.cat_in0_pool] = i;
                        defer_written = true;
                    } else CG_TRY(k_slice_copy<T>(dy, ca + cb, 0, dx, ca, 0, ca, npix, acc, st));
                }
                if (need(d.in1)) {
                    if (L.cat_in1_up >= 0 && plain_grads && !written[d.in1])      // the upsample's backward gathers from the slice in place
                        up_src[L.cat_in1_up] = i;
                    else
                        CG_TRY(k_slice_copy<T>(dy, ca + cb, ca, G(d.in1), cb, 0, cb, npix,
                                               (int)written[d.in1] || (d.in1 == tin && want_dx), st));
                    written[d.in1] = 1;
                }
                break;
            }
            case CG_OP_AVGPOOL:
                if (want_dx) {
                    if (pend_pool[i] >= 0) {
                        const LayerInfo& Cc = net->layers[pend_pool[i]];
                        CG_TRY(k_avgpool_bwd<T>(dy, dx, nb, h, w, d.cin, acc, st, G(Cc.out_t), net->chan[Cc.out_t], 0));
                    } else CG_TRY(k_avgpool_bwd<T>(dy, dx, nb, h, w, d.cin, acc, st));
                }
                break;
            case CG_OP_UPSAMPLE:
                if (want_dx) {
                    if (up_src[i] >= 0) {
                        const LayerInfo& Cc = net->layers[up_src[i]];
                        CG_TRY(k_upsample_bwd<T>((const T*)nullptr, dx, nb, h, w, d.cin, acc, st, G(Cc.out_t), net->chan[Cc.out_t],
                                                 net->chan[Cc.d.in0]));
                    } else CG_TRY(k_upsample_bwd<T>(dy, dx, nb, h, w, d.cin, acc, st));
                }
                break;
            case CG_OP_DROPOUT:
                if (want_dx) {
                    const int grp = c->bn_group > 0 ? c->bn_group : c->N;
                    if (n0 % grp || nb % grp) { cg_set_error("sub-batch [%d,%d) splits a dropout call of %d", n0, n0 + nb, grp); return CG_ERR_INVALID; }
                    DropKey key = drop_key(c, L);
                    for (int g = 0; g < 4; ++g) key.call_id[g] = c->call_id[(n0 / grp + g) & 3];
                    CG_TRY(k_dropout_bwd<T>(dy, dx, (size_t)grp * c->sample_elems(tin), nb / grp, d.rate, key, c->training ? 1 : 0, acc, st));
                }
                break;
            default: cg_set_error("unknown op %d", d.op); return CG_ERR_INVALID;
        }
        if (want_dx && !defer_written) written[tin] = 1;
        defer_written = false;
        if (hook) CG_TRY(hook(hook_user, i));
    }
    return CG_OK;
}

int net_backward(CallCtx* ctx, const float* params, const void* dy, void* dx, float* grads, int n0, int nb,
                 cudaStream_t st, LayerHook hook, void* hook_user) {
    return ctx->net->mode == CG_MODE_BF16
               ? backward_T<bf16>(ctx, params, (const bf16*)dy, (bf16*)dx, grads, n0, nb, st, hook, hook_user)
               : backward_T<float>(ctx, params, (const float*)dy, (float*)dx, grads, n0, nb, st, hook, hook_user);
}
