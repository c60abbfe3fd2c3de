// CycleGan.train_step / validate_step as one native schedule (reference: cyclegan/model.py:91-154).
//
// Forward (instance norm is per sample, so concatenating calls along the batch is exact):
//   F1 = g_AB([real_a; real_b]) -> [fake_b ; same_b]        F2 = g_BA([real_b; real_a]) -> [fake_a ; same_a]
//   C1 = g_BA(fake_b) -> cycled_a                            C2 = g_AB(fake_a) -> cycled_b
//   DA = d_A([real_a; fake_a])                               DB = d_B([real_b; fake_b])
// Backward (one combined generator backward gives both reference generator gradients):
//   D loss   : DA, DB over both halves  -> d_A, d_B parameter gradients (no image gradient)
//   G adv    : DA, DB fake half, data-gradient only -> d fake_a, d fake_b
//   cycle    : C1 -> theta_BA += , d fake_b += ;   C2 -> theta_AB += , d fake_a +=
//   F1, F2   : [d fake ; d same(identity)] -> theta_AB += , theta_BA +=   (no input gradient)
// then four fused Adam updates that all see the pre-update weights.
#include <dlfcn.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "kernels.h"
#include "net.h"
#include "prof.h"

// ---- NCCL through dlopen (the torch wheel bundles libnccl.so.2; nothing to link at build time) ----
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
    if (g_nccl.h) return CG_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { cg_set_error("dlopen(libnccl.so.2) failed: %s", dlerror()); return CG_ERR_COMM; }
#define SYM(field, name)                                                         \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                   \
    if (!g_nccl.field) { cg_set_error("libnccl lacks %s", name); return CG_ERR_COMM; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(AllReduce, "ncclAllReduce")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GetErrorString, "ncclGetErrorString")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
#undef SYM
    g_nccl.h = h;
    return CG_OK;
}
#define CG_NCCL(call)                                                                              \
    do {                                                                                           \
        int r__ = (call);                                                                          \
        if (r__ != 0) { cg_set_error("%s -> %s", #call, g_nccl.GetErrorString(r__)); return CG_ERR_COMM; } \
    } while (0)

enum { S_ADV_AB = 0, S_ADV_BA, S_CYC_A, S_CYC_B, S_ID_A, S_ID_B, S_DA_REAL, S_DA_FAKE, S_DB_REAL, S_DB_FAKE,
       S_OK_A, S_OK_B, S_COUNT = 16 };

struct cg_trainer_s {
    cg_net_t net[4] = {nullptr, nullptr, nullptr, nullptr};     // g_AB, g_BA, d_A, d_B
    cg_train_cfg cfg;
    float* params[4] = {}; float* grads[4] = {}; float* m[4] = {}; float* v[4] = {};
    char* ws = nullptr; size_t ws_bytes = 0;
    long long iters[4] = {0, 0, 0, 0};
    unsigned long long train_calls = 0;         // training-mode steps so far: the dropout counter of the next one
    unsigned long long seed_epoch[4] = {0, 0, 0, 0};   // cg_net_set_seed generation each net had when the graphs were captured
    size_t o_ctr = 0;                           // device copy of that counter (read by the dropout kernels, also in graph replays)
    // planned shape
    int B = 0, H = 0, W = 0; bool planned = false;
    CallCtx F1, F2, C1, C2, DA, DB;
    // misc buffers (activation dtype unless noted)
    size_t o_seedF1 = 0, o_seedF2 = 0, o_seedC1 = 0, o_seedC2 = 0, o_dxC1 = 0, o_dxC2 = 0, o_seedDA = 0, o_seedDB = 0,
           o_advDA = 0, o_advDB = 0, o_sums = 0, o_arena = 0, o_packed[4] = {0, 0, 0, 0}, o_tcs = 0, total = 0;
    int d_h = 0, d_w = 0, d_c = 0;
    // data parallel
    // CUDA graph of the gradient step (everything between the input conversion and the metrics copy-out): captured the
    // second time a (train, stream) pair is seen for the planned shape, replayed afterwards
    size_t o_metrics = 0;
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};        // [validate, train]
    cudaStream_t cap_stream = nullptr;                         // capture happens here (the caller's stream may be the legacy
                                                               // default stream, which cannot be captured); replay on the caller's
    int graph_seen[2] = {0, 0};
    long long graph_launches[2] = {0, 0};                      // kernels per replay (for cg_launch_count)
    nccl_comm comm = nullptr; int world = 1, rank = 0;
    cudaStream_t comm_stream = nullptr; cudaEvent_t ev_d = nullptr, ev_g = nullptr, ev_done = nullptr;
    std::vector<cudaEvent_t> ev_bucket;                        // one fork event per gradient bucket of a step
    int ev_next = 0;
    // two-chain schedule (CG_DUAL_STREAM=1, EXPERIMENTAL, default off): the step is two independent chains until the losses and
    // again in the backward -- X = {g_AB([a;b]), g_BA(fake_b), d_B}, Y = {g_BA([b;a]), g_AB(fake_a), d_A} -- run on two streams
    // with their own gradient arena and unfold scratch, so one chain's kernels fill the SMs the other's last wave leaves idle.
    // Measured: C3 44.5 -> 44.0 ms, C2 24.9 -> 22.9 ms per step.  Off by default: with the two first-hop backward calls
    // running concurrently the layer-by-layer parity test (tests/test_gpu_layerwise.py::test_c3_full_size_gradients[fp32])
    // fails in ~3 of 4 runs with 1e-2 errors in one generator's gradients (0 of 7 with host-side stream synchronisation in
    // place of the events, 0 of 9 single-chain); the cause is not found yet.  CG_DUAL_PARTS selects the parts
    // that run on two streams (1 forward, 2 backward phase 1, 4 backward phase 2).
    bool dual = false;
    cudaStream_t st2 = nullptr;
    cudaEvent_t ev2[8] = {};
    size_t o_arena2 = 0, o_tcs2 = 0;
};

static bool dual_default() {      // read when a trainer is created (tests build single- and two-chain trainers side by side)
    const char* e = getenv("CG_DUAL_STREAM");
    return e && e[0] == '1';
}

// Bucketed all-reduce of a generator's gradients under its LAST backward call: the flat gradient buffer is
// in forward-layer order and the backward retires layers last to first, so after layer i every gradient at an offset >=
// (first variable of layer i) is final.  Whenever >= 1/4 of the net has accumulated, that tail range is all-reduced on
// the communication stream (fork by event) while the backward continues.
struct BucketCtx {
    cg_trainer_s* tr; int net; cudaStream_t st; long long done_from; long long min_floats; int rc_unused;
};
static int bucket_fire(BucketCtx* b, long long lo) {
    cg_trainer_s* tr = b->tr;
    if (lo >= b->done_from) return CG_OK;
    if (tr->ev_next >= (int)tr->ev_bucket.size()) { cg_set_error("gradient buckets: out of events"); return CG_ERR_STATE; }
    cudaEvent_t ev = tr->ev_bucket[tr->ev_next++];
    CG_CUDA(cudaEventRecord(ev, b->st));
    CG_CUDA(cudaStreamWaitEvent(tr->comm_stream, ev, 0));
    float* g = tr->grads[b->net] + lo;
    CG_NCCL(g_nccl.AllReduce(g, g, (size_t)(b->done_from - lo), 7, 0, tr->comm, tr->comm_stream));
    b->done_from = lo;
    return CG_OK;
}
static int bucket_hook(void* user, int layer) {
    BucketCtx* b = (BucketCtx*)user;
    const LayerInfo& L = b->tr->net[b->net]->layers[layer];
    const long long lo = L.w_off >= 0 ? L.w_off : L.g_off;      // first variable of the layer (kernel or gamma)
    if (lo < 0 || b->done_from - lo < b->min_floats) return CG_OK;
    return bucket_fire(b, lo);
}

static size_t es_of(cg_trainer_t tr) { return tr->net[0]->elem_size(); }

// compute (and optionally assign) the workspace layout for batch B
static int trainer_layout(cg_trainer_t tr, int B, int H, int W, bool assign) {
    CG_TRY(net_plan(tr->net[0], 2 * B, H, W, true, &tr->F1));
    CG_TRY(net_plan(tr->net[1], 2 * B, H, W, true, &tr->F2));
    CG_TRY(net_plan(tr->net[1], B, H, W, true, &tr->C1));
    CG_TRY(net_plan(tr->net[0], B, H, W, true, &tr->C2));
    CG_TRY(net_plan(tr->net[2], 2 * B, H, W, true, &tr->DA));
    CG_TRY(net_plan(tr->net[3], 2 * B, H, W, true, &tr->DB));
    for (int g = 0; g < 2; ++g) {
        int ho, wo;
        CG_TRY(net_out_hw(tr->net[g], H, W, &ho, &wo));
        if (ho != H || wo != W || tr->net[g]->chan.back() != 3 || tr->net[g]->chan[0] != 3) {
            cg_set_error("generator %d must map [H,W,3] -> [H,W,3] (got %dx%dx%d)", g, ho, wo, tr->net[g]->chan.back());
            return CG_ERR_INVALID;
        }
    }
    CG_TRY(net_out_hw(tr->net[2], H, W, &tr->d_h, &tr->d_w));
    tr->d_c = tr->net[2]->chan.back();
    int h2, w2;
    CG_TRY(net_out_hw(tr->net[3], H, W, &h2, &w2));
    if (h2 != tr->d_h || w2 != tr->d_w || tr->net[3]->chan.back() != tr->d_c) {
        cg_set_error("the two discriminators must have identical output shapes");
        return CG_ERR_INVALID;
    }
    const size_t es = es_of(tr);
    const size_t img = (size_t)H * W * 3 * es, dout = (size_t)tr->d_h * tr->d_w * tr->d_c * es;
    size_t off = 4096;          // slack before the first tensor: window views start a few pixels before their tensor
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    size_t oF1 = take(tr->F1.act_bytes), oF2 = take(tr->F2.act_bytes), oC1 = take(tr->C1.act_bytes),
           oC2 = take(tr->C2.act_bytes), oDA = take(tr->DA.act_bytes), oDB = take(tr->DB.act_bytes);
    size_t arena = 0;
    for (CallCtx* c : {&tr->F1, &tr->F2, &tr->C1, &tr->C2, &tr->DA, &tr->DB}) if (c->grad_bytes > arena) arena = c->grad_bytes;
    tr->o_arena = take(arena);
    tr->o_arena2 = tr->dual ? take(arena) : tr->o_arena;
    tr->o_seedF1 = take(2 * B * img); tr->o_seedF2 = take(2 * B * img);
    tr->o_seedC1 = take(B * img); tr->o_seedC2 = take(B * img);
    tr->o_dxC1 = take(B * img); tr->o_dxC2 = take(B * img);
    tr->o_seedDA = take(2 * B * dout); tr->o_seedDB = take(2 * B * dout);
    tr->o_advDA = take(B * dout); tr->o_advDB = take(B * dout);
    tr->o_sums = take(S_COUNT * sizeof(float));
    tr->o_metrics = take(8 * sizeof(float));
    tr->o_ctr = take(sizeof(unsigned long long));
    off = align_up(off, 1024);
    for (int i = 0; i < 4; ++i) tr->o_packed[i] = take(align_up(tr->net[i]->packed_bytes, 1024));
    size_t tcs = 0;
    for (CallCtx* c : {&tr->F1, &tr->F2, &tr->C1, &tr->C2, &tr->DA, &tr->DB}) if (c->tcs_bytes > tcs) tcs = c->tcs_bytes;
    tr->o_tcs = take(align_up(tcs, 1024));
    tr->o_tcs2 = tr->dual ? take(align_up(tcs, 1024)) : tr->o_tcs;
    off += 4096;                // ... and end a few pixels after it
    tr->total = off;
    if (assign) {
        if (off > tr->ws_bytes) { cg_set_error("trainer workspace %zu < required %zu", tr->ws_bytes, off); return CG_ERR_WORKSPACE; }
        size_t bases[6] = {oF1, oF2, oC1, oC2, oDA, oDB};
        CallCtx* cs[6] = {&tr->F1, &tr->F2, &tr->C1, &tr->C2, &tr->DA, &tr->DB};
        const int chain[6] = {0, 1, 0, 1, 1, 0};          // F1, C1, DB on chain X; F2, C2, DA on chain Y
        for (int i = 0; i < 6; ++i) {
            cs[i]->base = tr->ws + bases[i];
            cs[i]->arena = tr->ws + (chain[i] ? tr->o_arena2 : tr->o_arena);
            cs[i]->ext_input = nullptr;
        }
        // BatchNormalization / Dropout see the Keras calls of model.py:93-106, not the concatenated batches: B samples per
        // call; call ids = the order in which validate_step calls each model (g_AB: real_a, fake_a, real_b; g_BA: fake_b,
        // real_b, real_a; d_A: real_a, fake_a; d_B: real_b, fake_b)
        const int ids[6][2] = {{0, 2}, {1, 2}, {0, 0}, {1, 1}, {0, 1}, {0, 1}};
        for (int i = 0; i < 6; ++i) {
            cs[i]->bn_group = B;
            cs[i]->defer_moving = true;
            cs[i]->drop_ctr_dev = (const unsigned long long*)(tr->ws + tr->o_ctr);
            cs[i]->call_id[0] = ids[i][0]; cs[i]->call_id[1] = ids[i][1];
        }
        // cycle calls read the first half (the fakes) of F1 / F2's output in place
        tr->C1.ext_input = tr->F1.act(tr->net[0]->out_tensor());
        tr->C2.ext_input = tr->F2.act(tr->net[1]->out_tensor());
        const int owner[6] = {0, 1, 1, 0, 2, 3};
        for (int i = 0; i < 6; ++i) {
            cs[i]->packed = tr->ws + tr->o_packed[owner[i]];
            cs[i]->tcs = tr->ws + (chain[i] ? tr->o_tcs2 : tr->o_tcs);
            CG_TRY(net_bind(cs[i]));
        }
        for (int g = 0; g < 2; ++g) {
            if (tr->graph_exec[g]) { cudaGraphExecDestroy(tr->graph_exec[g]); tr->graph_exec[g] = nullptr; }
            tr->graph_seen[g] = 0;
        }
        tr->B = B; tr->H = H; tr->W = W; tr->planned = true;
    }
    return CG_OK;
}

extern "C" int cg_trainer_create(cg_net_t g_AB, cg_net_t g_BA, cg_net_t d_A, cg_net_t d_B, const cg_train_cfg* cfg,
                                 cg_trainer_t* out) {
    if (!g_AB || !g_BA || !d_A || !d_B || !cfg || !out) { cg_set_error("cg_trainer_create: null argument"); return CG_ERR_INVALID; }
    if (g_AB->mode != g_BA->mode || g_AB->mode != d_A->mode || g_AB->mode != d_B->mode) {
        cg_set_error("all four nets must use the same arithmetic mode");
        return CG_ERR_INVALID;
    }
    if (cfg->loss < CG_LOSS_MSE || cfg->loss > CG_LOSS_BCE) { cg_set_error("unknown loss %d", cfg->loss); return CG_ERR_INVALID; }
    for (int i = 0; i < 4; ++i)
        if (cfg->adam[i].kind < CG_OPT_ADAM || cfg->adam[i].kind > CG_OPT_ADABELIEF) {
            cg_set_error("unknown optimizer kind %d for net %d", cfg->adam[i].kind, i);
            return CG_ERR_INVALID;
        }
    cg_trainer_s* tr = new cg_trainer_s();
    tr->net[0] = g_AB; tr->net[1] = g_BA; tr->net[2] = d_A; tr->net[3] = d_B;
    tr->cfg = *cfg;
    tr->dual = dual_default();
    *out = tr;
    return CG_OK;
}

extern "C" void cg_trainer_destroy(cg_trainer_t tr) {
    if (!tr) return;
    for (int g = 0; g < 2; ++g)             // graphs first: they may hold NCCL kernels of the communicator
        if (tr->graph_exec[g]) cudaGraphExecDestroy(tr->graph_exec[g]);
    if (tr->cap_stream) cudaStreamDestroy(tr->cap_stream);
    if (tr->st2) cudaStreamDestroy(tr->st2);
    for (cudaEvent_t e : tr->ev2) if (e) cudaEventDestroy(e);
    if (tr->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(tr->comm);
    if (tr->comm_stream) cudaStreamDestroy(tr->comm_stream);
    if (tr->ev_d) cudaEventDestroy(tr->ev_d);
    if (tr->ev_g) cudaEventDestroy(tr->ev_g);
    if (tr->ev_done) cudaEventDestroy(tr->ev_done);
    for (cudaEvent_t e : tr->ev_bucket) cudaEventDestroy(e);
    delete tr;
}

extern "C" int cg_trainer_workspace_bytes(cg_trainer_t tr, int B, int H, int W, size_t* bytes) {
    if (!tr || !bytes) { cg_set_error("null argument"); return CG_ERR_INVALID; }
    CG_TRY(trainer_layout(tr, B, H, W, false));
    tr->planned = false;
    *bytes = tr->total;
    return CG_OK;
}

extern "C" int cg_trainer_bind(cg_trainer_t tr, float* const params[4], float* const grads[4], float* const m[4],
                               float* const v[4], void* ws, size_t ws_bytes) {
    if (!tr || !params || !grads || !m || !v || !ws) { cg_set_error("cg_trainer_bind: null argument"); return CG_ERR_INVALID; }
    for (int i = 0; i < 4; ++i) {
        if (!params[i] || !grads[i] || !m[i] || !v[i]) { cg_set_error("cg_trainer_bind: null buffer for net %d", i); return CG_ERR_INVALID; }
        tr->params[i] = params[i]; tr->grads[i] = grads[i]; tr->m[i] = m[i]; tr->v[i] = v[i];
    }
    tr->ws = (char*)ws; tr->ws_bytes = ws_bytes; tr->planned = false;
    return CG_OK;
}

__global__ void metrics_kernel(const float* __restrict__ s, float* __restrict__ out, float n_d, float n_img,
                               float w_cyc, float w_id, float w_gen, float w_disc) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float cyc = w_cyc * (s[S_CYC_A] + s[S_CYC_B]) / n_img;
    out[0] = w_gen * s[S_ADV_AB] / n_d + cyc + w_id * s[S_ID_B] / n_img;      // total_gAB_loss  model.py:115-116
    out[1] = w_gen * s[S_ADV_BA] / n_d + cyc + w_id * s[S_ID_A] / n_img;      // total_gBA_loss  model.py:117-118
    out[2] = w_disc * (s[S_DA_REAL] + s[S_DA_FAKE]) / n_d;                    // da_loss         model.py:120
    out[3] = w_disc * (s[S_DB_REAL] + s[S_DB_FAKE]) / n_d;                    // db_loss         model.py:121
    out[4] = s[S_OK_A] / (2.f * n_d);                                         // accuracy        model.py:123
    out[5] = s[S_OK_B] / (2.f * n_d);
}

// the part of a step that depends only on the trainer's own buffers (capturable into a CUDA graph)
template <typename T>
static int step_body(cg_trainer_t tr, int B, int H, int W, bool train, cudaStream_t st) {
    const cg_train_cfg& cfg = tr->cfg;
    const size_t img = (size_t)H * W * 3;               // elements per image
    const size_t dout = (size_t)tr->d_h * tr->d_w * tr->d_c;
    const int tG = tr->net[0]->out_tensor(), tGb = tr->net[1]->out_tensor();
    const int tDa = tr->net[2]->out_tensor(), tDb = tr->net[3]->out_tensor();
    char* ws = tr->ws;
    float* sums = (float*)(ws + tr->o_sums);
    float* metrics = (float*)(ws + tr->o_metrics);
    CG_CUDA(cudaMemsetAsync(sums, 0, S_COUNT * sizeof(float), st));

    for (int i = 0; i < 4; ++i) CG_TRY(net_pack(tr->net[i], tr->params[i], ws + tr->o_packed[i], st));
    T* Xab = (T*)tr->F1.act(0);
    const T* ra = Xab;
    const T* rb = Xab + B * img;

    // ---- two chains -------------------------------------------------------------------------
    // X = {F1, C1, DB} on `st`, Y = {F2, C2, DA} on `sy` (== st when the two-chain schedule is off).  fork(e): sy continues
    // after everything enqueued on st so far; join(e): st continues after everything enqueued on sy so far.  Inside a
    // stream capture the events become graph edges, so the captured step has two parallel branches.
    cudaStream_t sy = st;
    if (tr->dual) {
        if (!tr->st2) CG_CUDA(cudaStreamCreateWithFlags(&tr->st2, cudaStreamNonBlocking));
        for (cudaEvent_t& e : tr->ev2) if (!e) CG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        sy = tr->st2;
    }
    auto fork = [&](int e) -> int {
        if (sy == st) return CG_OK;
        CG_CUDA(cudaEventRecord(tr->ev2[e], st));
        CG_CUDA(cudaStreamWaitEvent(sy, tr->ev2[e], 0));
        return CG_OK;
    };
    auto join = [&](int e) -> int {
        if (sy == st) return CG_OK;
        CG_CUDA(cudaEventRecord(tr->ev2[e], sy));
        CG_CUDA(cudaStreamWaitEvent(st, tr->ev2[e], 0));
        return CG_OK;
    };

    // ---- forward --------------------------------------------------------------------------
    for (CallCtx* c : {&tr->F1, &tr->F2, &tr->C1, &tr->C2, &tr->DA, &tr->DB}) c->training = train;   // model.py:138-141 / :221
    const T* F1o = (const T*)tr->F1.act(tG);            // [fake_b ; same_b]
    const T* F2o = (const T*)tr->F2.act(tGb);           // [fake_a ; same_a]
    const T* cyc_a = (const T*)tr->C1.act(tGb);
    const T* cyc_b = (const T*)tr->C2.act(tG);
    T* DAin = (T*)tr->DA.act(0);
    T* DBin = (T*)tr->DB.act(0);
    static const int parts = [] { const char* e = getenv("CG_DUAL_PARTS"); return e ? atoi(e) : 7; }();   // diagnostic bisect
    cudaStream_t sy_all = sy;
    sy = (parts & 1) ? sy_all : st;
    CG_TRY(fork(0));
    CG_TRY(net_forward(&tr->F1, tr->params[0], st));
    CG_TRY(net_forward(&tr->F2, tr->params[1], sy));
    CG_TRY(net_forward(&tr->C1, tr->params[1], st));    // g_BA(fake_b)
    CG_TRY(net_forward(&tr->C2, tr->params[0], sy));    // g_AB(fake_a)
    CG_TRY(k_copy_acc<T>(rb, DBin, B * img, 0, st));
    CG_TRY(k_copy_acc<T>(F1o, DBin + B * img, B * img, 0, st));       // fake_b
    CG_TRY(k_copy_acc<T>(ra, DAin, B * img, 0, sy));
    CG_TRY(k_copy_acc<T>(F2o, DAin + B * img, B * img, 0, sy));       // fake_a
    CG_TRY(net_forward(&tr->DB, tr->params[3], st));
    CG_TRY(net_forward(&tr->DA, tr->params[2], sy));
    CG_TRY(join(1));
    const T* dA = (const T*)tr->DA.act(tDa);            // [disc_real_a ; disc_fake_a]
    const T* dB = (const T*)tr->DB.act(tDb);
    if (train) {    // BatchNormalization moving averages, one update per Keras call in the order of model.py:93-106
        if (tr->net[0]->n_state) { CG_TRY(net_update_moving(&tr->F1, 0, st)); CG_TRY(net_update_moving(&tr->C2, 0, st)); CG_TRY(net_update_moving(&tr->F1, 1, st)); }
        if (tr->net[1]->n_state) { CG_TRY(net_update_moving(&tr->C1, 0, st)); CG_TRY(net_update_moving(&tr->F2, 0, st)); CG_TRY(net_update_moving(&tr->F2, 1, st)); }
        if (tr->net[2]->n_state) { CG_TRY(net_update_moving(&tr->DA, 0, st)); CG_TRY(net_update_moving(&tr->DA, 1, st)); }
        if (tr->net[3]->n_state) { CG_TRY(net_update_moving(&tr->DB, 0, st)); CG_TRY(net_update_moving(&tr->DB, 1, st)); }
    }

    // ---- losses (+ gradient seeds when training) ---------------------------------------------
    const size_t n_d = (size_t)B * dout, n_img = (size_t)B * img;
    T* seedF1 = (T*)(ws + tr->o_seedF1); T* seedF2 = (T*)(ws + tr->o_seedF2);
    T* seedC1 = (T*)(ws + tr->o_seedC1); T* seedC2 = (T*)(ws + tr->o_seedC2);
    T* seedDA = (T*)(ws + tr->o_seedDA); T* seedDB = (T*)(ws + tr->o_seedDB);
    T* advDA = (T*)(ws + tr->o_advDA);   T* advDB = (T*)(ws + tr->o_advDB);
    const float gd = cfg.w_discriminator / (float)n_d, gg = cfg.w_generator / (float)n_d;
    const float gc = cfg.w_cycle / (float)n_img, gi = cfg.w_identity / (float)n_img;
    auto opt = [&](T* p) -> T* { return train ? p : nullptr; };
    CG_TRY(k_adv_loss<T>(dA, n_d, 1.f, cfg.loss, gd, opt(seedDA), sums + S_DA_REAL, sums + S_OK_A, st));
    CG_TRY(k_adv_loss<T>(dA + n_d, n_d, 0.f, cfg.loss, gd, opt(seedDA + n_d), sums + S_DA_FAKE, sums + S_OK_A, st));
    CG_TRY(k_adv_loss<T>(dB, n_d, 1.f, cfg.loss, gd, opt(seedDB), sums + S_DB_REAL, sums + S_OK_B, st));
    CG_TRY(k_adv_loss<T>(dB + n_d, n_d, 0.f, cfg.loss, gd, opt(seedDB + n_d), sums + S_DB_FAKE, sums + S_OK_B, st));
    CG_TRY(k_adv_loss<T>(dB + n_d, n_d, 1.f, cfg.loss, gg, opt(advDB), sums + S_ADV_AB, nullptr, st));   // gAB: D_B(fake_b)
    CG_TRY(k_adv_loss<T>(dA + n_d, n_d, 1.f, cfg.loss, gg, opt(advDA), sums + S_ADV_BA, nullptr, st));   // gBA: D_A(fake_a)
    CG_TRY(k_l1_loss<T>(ra, cyc_a, n_img, gc, opt(seedC1), 0, sums + S_CYC_A, st));
    CG_TRY(k_l1_loss<T>(rb, cyc_b, n_img, gc, opt(seedC2), 0, sums + S_CYC_B, st));
    CG_TRY(k_l1_loss<T>(rb, F1o + n_img, n_img, gi, opt(seedF1 + n_img), 0, sums + S_ID_B, st));          // same_b
    CG_TRY(k_l1_loss<T>(ra, F2o + n_img, n_img, gi, opt(seedF2 + n_img), 0, sums + S_ID_A, st));          // same_a
    metrics_kernel<<<1, 32, 0, st>>>(sums, metrics, (float)n_d, (float)n_img, cfg.w_cycle, cfg.w_identity,
                                     cfg.w_generator, cfg.w_discriminator);
    CG_LAUNCH_CHECK();
    if (!train) return CG_OK;

    // ---- backward ---------------------------------------------------------------------------
    // chain X: d_B loss -> d fake_b through the frozen d_B -> g_BA(fake_b) -> g_AB([a;b]);  chain Y: the mirror image.
    // Parameter gradients are accumulated with plain read-modify-writes in places, so the two chains never write the same
    // net at the same time: phase 1 = {X: d_B, g_BA | Y: d_A, g_AB}, phase 2 = {X: g_AB | Y: g_BA}, swapped by two events.
    for (int i = 0; i < 4; ++i)
        CG_CUDA(cudaMemsetAsync(tr->grads[i], 0, sizeof(float) * (size_t)tr->net[i]->n_params, st));
    sy = (parts & 2) ? sy_all : st;
    CG_TRY(fork(2));
    // discriminator losses: parameter gradients only
    CG_TRY(net_backward(&tr->DB, tr->params[3], seedDB, nullptr, tr->grads[3], 0, 2 * B, st));
    CG_TRY(net_backward(&tr->DA, tr->params[2], seedDA, nullptr, tr->grads[2], 0, 2 * B, sy));
    static const bool skip_ar = [] { const char* e = getenv("CG_DP_SKIP_AR"); return e && e[0] == '1'; }();   // measurement only
    if (tr->comm && !skip_ar) {     // d_A / d_B gradients are final: all-reduce them under the generator backward
        CG_CUDA(cudaEventRecord(tr->ev_d, st));
        CG_CUDA(cudaStreamWaitEvent(tr->comm_stream, tr->ev_d, 0));
        if (sy != st) {
            CG_CUDA(cudaEventRecord(tr->ev2[3], sy));
            CG_CUDA(cudaStreamWaitEvent(tr->comm_stream, tr->ev2[3], 0));
        }
        CG_NCCL(g_nccl.GroupStart());
        CG_NCCL(g_nccl.AllReduce(tr->grads[2], tr->grads[2], (size_t)tr->net[2]->n_params, 7, 0, tr->comm, tr->comm_stream));
        CG_NCCL(g_nccl.AllReduce(tr->grads[3], tr->grads[3], (size_t)tr->net[3]->n_params, 7, 0, tr->comm, tr->comm_stream));
        CG_NCCL(g_nccl.GroupEnd());
    }
    // adversarial generator terms: data gradient through the frozen discriminators (fake half)
    CG_TRY(net_backward(&tr->DB, tr->params[3], advDB, seedF1, nullptr, B, B, st));      // d fake_b
    CG_TRY(net_backward(&tr->DA, tr->params[2], advDA, seedF2, nullptr, B, B, sy));      // d fake_a
    // cycle terms
    T* dxC1 = (T*)(ws + tr->o_dxC1); T* dxC2 = (T*)(ws + tr->o_dxC2);
    CG_TRY(net_backward(&tr->C1, tr->params[1], seedC1, dxC1, tr->grads[1], 0, B, st));  // theta_BA, d fake_b
    CG_TRY(k_copy_acc<T>(dxC1, seedF1, n_img, 1, st));
    CG_TRY(net_backward(&tr->C2, tr->params[0], seedC2, dxC2, tr->grads[0], 0, B, sy));  // theta_AB, d fake_a
    CG_TRY(k_copy_acc<T>(dxC2, seedF2, n_img, 1, sy));
    // swap: X may touch theta_AB only after Y's cycle call has, Y theta_BA only after X's (a full join + fork)
    CG_TRY(join(4));
    sy = (parts & 4) ? sy_all : st;
    CG_TRY(fork(5));
    // first-hop generator calls: [d fake ; d same].  These are the last contributions to theta_AB / theta_BA, so with a
    // communicator their gradients can be all-reduced bucket by bucket while the backward is still running
    // (CG_DP_BUCKETS=1, single-chain schedule only).
    static const bool buckets_on = [] { const char* e = getenv("CG_DP_BUCKETS"); return e && e[0] == '1'; }();
    if (tr->comm && buckets_on && sy == st) {
        tr->ev_next = 0;
        BucketCtx b0{tr, 0, st, tr->net[0]->n_params, tr->net[0]->n_params / 4 + 1, 0};
        CG_TRY(net_backward(&tr->F1, tr->params[0], seedF1, nullptr, tr->grads[0], 0, 2 * B, st, bucket_hook, &b0));
        CG_TRY(bucket_fire(&b0, 0));
        BucketCtx b1{tr, 1, st, tr->net[1]->n_params, tr->net[1]->n_params / 4 + 1, 0};
        CG_TRY(net_backward(&tr->F2, tr->params[1], seedF2, nullptr, tr->grads[1], 0, 2 * B, st, bucket_hook, &b1));
        CG_TRY(bucket_fire(&b1, 0));
        CG_CUDA(cudaEventRecord(tr->ev_done, tr->comm_stream));
        CG_CUDA(cudaStreamWaitEvent(st, tr->ev_done, 0));
        return CG_OK;
    }
    CG_TRY(net_backward(&tr->F1, tr->params[0], seedF1, nullptr, tr->grads[0], 0, 2 * B, st));
    if (tr->comm && !skip_ar && sy == st) {     // theta_AB is final: its all-reduce runs under g_BA's last backward call
        if (tr->ev_bucket.empty()) { cg_set_error("communicator without events"); return CG_ERR_STATE; }
        CG_CUDA(cudaEventRecord(tr->ev_bucket[0], st));
        CG_CUDA(cudaStreamWaitEvent(tr->comm_stream, tr->ev_bucket[0], 0));
        CG_NCCL(g_nccl.AllReduce(tr->grads[0], tr->grads[0], (size_t)tr->net[0]->n_params, 7, 0, tr->comm, tr->comm_stream));
    }
    CG_TRY(net_backward(&tr->F2, tr->params[1], seedF2, nullptr, tr->grads[1], 0, 2 * B, sy));
    const bool ab_reduced = sy == st;
    CG_TRY(join(6));
    if (tr->comm && !skip_ar) {
        CG_CUDA(cudaEventRecord(tr->ev_g, st));
        CG_CUDA(cudaStreamWaitEvent(tr->comm_stream, tr->ev_g, 0));
        CG_NCCL(g_nccl.GroupStart());
        if (!ab_reduced)
            CG_NCCL(g_nccl.AllReduce(tr->grads[0], tr->grads[0], (size_t)tr->net[0]->n_params, 7, 0, tr->comm, tr->comm_stream));
        CG_NCCL(g_nccl.AllReduce(tr->grads[1], tr->grads[1], (size_t)tr->net[1]->n_params, 7, 0, tr->comm, tr->comm_stream));
        CG_NCCL(g_nccl.GroupEnd());
        CG_CUDA(cudaEventRecord(tr->ev_done, tr->comm_stream));
        CG_CUDA(cudaStreamWaitEvent(st, tr->ev_done, 0));
    }
    return CG_OK;
}

template <typename T>
static int step_T(cg_trainer_t tr, const float* real_a, const float* real_b, int B, int H, int W, float* metrics,
                  bool train, cudaStream_t st) {
    if (!tr->ws) { cg_set_error("trainer has no bound buffers (cg_trainer_bind)"); return CG_ERR_STATE; }
    if (!tr->planned || tr->B != B || tr->H != H || tr->W != W) CG_TRY(trainer_layout(tr, B, H, W, true));
    // a re-seeded net (cg_net_set_seed): the dropout key is a kernel argument baked into the captured graphs -> drop them and
    // restart the step counter, exactly like the single-net path restarts its call counter
    for (int i = 0; i < 4; ++i)
        if (tr->seed_epoch[i] != tr->net[i]->seed_epoch) {
            for (int g = 0; g < 2; ++g) {
                if (tr->graph_exec[g]) { cudaGraphExecDestroy(tr->graph_exec[g]); tr->graph_exec[g] = nullptr; }
                tr->graph_seen[g] = 0;
            }
            tr->train_calls = 0;
            for (int j = 0; j < 4; ++j) tr->seed_epoch[j] = tr->net[j]->seed_epoch;
            break;
        }
    // ---- inputs: Xab = [a; b], Xba = [b; a] (the only part that touches caller pointers) ------
    const size_t img = (size_t)H * W * 3;
    T* Xab = (T*)tr->F1.act(0);
    T* Xba = (T*)tr->F2.act(0);
    CG_TRY(k_convert_in<T>(real_a, Xab, B * img, st));
    CG_TRY(k_convert_in<T>(real_b, Xab + B * img, B * img, st));
    CG_TRY(k_convert_in<T>(real_b, Xba, B * img, st));
    CG_TRY(k_convert_in<T>(real_a, Xba + B * img, B * img, st));
    if (train) {
        bool any_drop = false;
        for (int i = 0; i < 4; ++i)
            for (const LayerInfo& L : tr->net[i]->layers) any_drop |= L.drop_index >= 0;
        if (any_drop) CG_TRY(k_set_counter((unsigned long long*)(tr->ws + tr->o_ctr), tr->train_calls, st));
        tr->train_calls += 1;
    }

    // ---- body: replay the captured graph, capture it, or run it eagerly ------------------------
    const char* goff = getenv("CG_DISABLE_GRAPH");      // read per call: tests switch it between trainers
    const bool graphs_off = goff && goff[0] == '1';
    const int gi = train ? 1 : 0;
    // with a communicator attached the all-reduces (side stream, fork/join by events) become graph nodes too; opt out
    // with CG_GRAPH_NCCL=0
    static const bool graphs_nccl = [] { const char* e = getenv("CG_GRAPH_NCCL"); return !(e && e[0] == '0'); }();
    const bool graph_ok = !graphs_off && !prof_enabled() && (!tr->comm || graphs_nccl);
    if (graph_ok && tr->graph_exec[gi]) {
        CG_CUDA(cudaGraphLaunch(tr->graph_exec[gi], st));
        g_launches.fetch_add(tr->graph_launches[gi], std::memory_order_relaxed);
    } else if (graph_ok && tr->graph_seen[gi] >= 1) {
        // second call for this shape: every lazy one-time initialisation (function attributes, NCCL channels) has happened
        // in the first, eager one, so the body can be captured
        if (!tr->cap_stream) CG_CUDA(cudaStreamCreateWithFlags(&tr->cap_stream, cudaStreamNonBlocking));
        const long long l0 = g_launches.load();
        cudaGraph_t graph = nullptr;
        CG_CUDA(cudaStreamBeginCapture(tr->cap_stream, cudaStreamCaptureModeThreadLocal));
        const int rc = step_body<T>(tr, B, H, W, train, tr->cap_stream);
        const cudaError_t ce = cudaStreamEndCapture(tr->cap_stream, &graph);
        bool replayed = false;
        if (rc == CG_OK && ce == cudaSuccess && graph) {
            tr->graph_launches[gi] = g_launches.load() - l0;
            if (cudaGraphInstantiate(&tr->graph_exec[gi], graph, 0) == cudaSuccess) {
                CG_CUDA(cudaGraphLaunch(tr->graph_exec[gi], st));
                replayed = true;
            } else {
                tr->graph_exec[gi] = nullptr;
            }
        }
        if (graph) cudaGraphDestroy(graph);
        if (!replayed) {                                // capture is not possible here: stay eager from now on
            cudaGetLastError();
            tr->graph_seen[gi] = -1000000;
            if (rc != CG_OK) return rc;
            CG_TRY(step_body<T>(tr, B, H, W, train, st));
        }
    } else {
        CG_TRY(step_body<T>(tr, B, H, W, train, st));
        if (graph_ok) tr->graph_seen[gi] += 1;
    }
    CG_CUDA(cudaMemcpyAsync(metrics, tr->ws + tr->o_metrics, 6 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return CG_OK;
}

static int step(cg_trainer_t tr, const float* a, const float* b, int B, int H, int W, float* metrics, bool train,
                void* stream) {
    if (!tr || !a || !b || !metrics) { cg_set_error("train/validate step: null argument"); return CG_ERR_INVALID; }
    if (B <= 0) { cg_set_error("empty batch"); return CG_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    return tr->net[0]->mode == CG_MODE_BF16 ? step_T<bf16>(tr, a, b, B, H, W, metrics, train, st)
                                            : step_T<float>(tr, a, b, B, H, W, metrics, train, st);
}

extern "C" int cg_validate_step(cg_trainer_t tr, const float* a, const float* b, int B, int H, int W, float* metrics,
                                void* stream) {
    return step(tr, a, b, B, H, W, metrics, false, stream);
}
extern "C" int cg_trainer_compute_gradients(cg_trainer_t tr, const float* a, const float* b, int B, int H, int W,
                                            float* metrics, void* stream) {
    return step(tr, a, b, B, H, W, metrics, true, stream);
}

// one optimizer.apply_gradients over a flat range: `iterations` is optimizer.iterations BEFORE the step
static int opt_apply(const cg_adam_cfg& a, float* p, const float* g, float* m, float* v, size_t n, long long iterations,
                     float gscale, cudaStream_t st) {
    const double t = (double)(iterations + 1);
    if (n == 0) return CG_OK;
    if (a.kind == CG_OPT_ADAM) {
        const double lr_t = (double)a.learning_rate * sqrt(1.0 - pow((double)a.beta_2, t)) / (1.0 - pow((double)a.beta_1, t));
        return k_adam(p, g, m, v, n, (float)lr_t, a.beta_1, a.beta_2, a.epsilon, gscale, st);
    }
    OptCoef c;
    memset(&c, 0, sizeof(c));
    c.lr = a.learning_rate; c.b1 = a.beta_1; c.b2 = a.beta_2; c.eps = a.epsilon; c.gscale = gscale;
    if (a.kind == CG_OPT_ADABELIEF) {       // adabelief_tf: bias corrections and the RAdam rectification term
        const double b1p = pow((double)a.beta_1, t), b2p = pow((double)a.beta_2, t);
        const double sma_inf = 2.0 / (1.0 - (double)a.beta_2) - 1.0;
        const double sma_t = sma_inf - 2.0 * t * b2p / (1.0 - b2p);
        c.c_m = (float)(1.0 / (1.0 - b1p));
        c.c_v = (float)(1.0 / (1.0 - b2p));
        c.rect = sma_t >= 5.0 ? 1 : 0;      // sma_threshold default
        c.r_t = c.rect ? (float)sqrt((sma_t - 4.0) / (sma_inf - 4.0) * (sma_t - 2.0) / (sma_inf - 2.0) * sma_inf / sma_t) : 0.f;
    }
    return k_opt_step(a.kind, p, g, m, v, n, c, st);
}

extern "C" int cg_optimizer_apply(const cg_adam_cfg* cfg, float* params, const float* grads, float* slot_m, float* slot_v,
                                  size_t n, int64_t iterations, void* stream) {
    if (!cfg || !params || !grads) { cg_set_error("cg_optimizer_apply: null argument"); return CG_ERR_INVALID; }
    if (cfg->kind < CG_OPT_ADAM || cfg->kind > CG_OPT_ADABELIEF) { cg_set_error("unknown optimizer kind %d", cfg->kind); return CG_ERR_INVALID; }
    const bool need_m = cfg->kind == CG_OPT_ADAM || cfg->kind == CG_OPT_ADABELIEF, need_v = cfg->kind != CG_OPT_SGD;
    if ((need_m && !slot_m) || (need_v && !slot_v)) { cg_set_error("cg_optimizer_apply: missing slot buffer for optimizer kind %d", cfg->kind); return CG_ERR_INVALID; }
    if (iterations < 0) { cg_set_error("cg_optimizer_apply: negative iteration count"); return CG_ERR_INVALID; }
    return opt_apply(*cfg, params, grads, slot_m, slot_v, n, iterations, 1.f, (cudaStream_t)stream);
}

extern "C" int cg_trainer_apply_gradients(cg_trainer_t tr, void* stream) {
    if (!tr || !tr->ws) { cg_set_error("trainer not bound"); return CG_ERR_STATE; }
    cudaStream_t st = (cudaStream_t)stream;
    const float gscale = 1.f / (float)tr->world;        // data parallel: mean of the per-rank mean gradients
    for (int i = 0; i < 4; ++i) {
        CG_TRY(opt_apply(tr->cfg.adam[i], tr->params[i], tr->grads[i], tr->m[i], tr->v[i], (size_t)tr->net[i]->n_params,
                         tr->iters[i], gscale, st));
        tr->iters[i] += 1;
    }
    return CG_OK;
}

extern "C" int cg_train_step(cg_trainer_t tr, const float* a, const float* b, int B, int H, int W, float* metrics,
                             void* stream) {
    CG_TRY(step(tr, a, b, B, H, W, metrics, true, stream));
    return cg_trainer_apply_gradients(tr, stream);
}

extern "C" int cg_trainer_get_iterations(cg_trainer_t tr, int64_t iters[4]) {
    if (!tr || !iters) { cg_set_error("null argument"); return CG_ERR_INVALID; }
    for (int i = 0; i < 4; ++i) iters[i] = tr->iters[i];
    return CG_OK;
}
extern "C" int cg_trainer_set_iterations(cg_trainer_t tr, const int64_t iters[4]) {
    if (!tr || !iters) { cg_set_error("null argument"); return CG_ERR_INVALID; }
    for (int i = 0; i < 4; ++i) tr->iters[i] = iters[i];
    return CG_OK;
}

extern "C" int cg_trainer_fetch_image(cg_trainer_t tr, int which, float* out, void* stream) {
    if (!tr || !out || !tr->planned || which < 0 || which > 5) { cg_set_error("fetch_image: bad argument / no step yet"); return CG_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)tr->B * tr->H * tr->W * 3;
    const int tG = tr->net[0]->out_tensor(), tGb = tr->net[1]->out_tensor();
    const char* src; size_t off = 0;
    switch (which) {
        case 0: src = (const char*)tr->F1.act(tG); break;
        case 1: src = (const char*)tr->F1.act(tG); off = n; break;
        case 2: src = (const char*)tr->F2.act(tGb); break;
        case 3: src = (const char*)tr->F2.act(tGb); off = n; break;
        case 4: src = (const char*)tr->C1.act(tGb); break;
        default: src = (const char*)tr->C2.act(tG); break;
    }
    if (tr->net[0]->mode == CG_MODE_BF16) return k_convert_out<bf16>((const bf16*)src + off, out, n, st);
    return k_convert_out<float>((const float*)src + off, out, n, st);
}

int fetch_tensor(const CallCtx* c, int t, float* out, int* shape4, cudaStream_t st);     // api.cu
extern "C" int cg_trainer_fetch_tensor(cg_trainer_t tr, int call, int tensor, float* out, int shape4[4], void* stream) {
    if (!tr || !tr->planned || call < 0 || call > 5) { cg_set_error("fetch_tensor: bad argument / no step yet"); return CG_ERR_INVALID; }
    const CallCtx* cs[6] = {&tr->F1, &tr->F2, &tr->C1, &tr->C2, &tr->DA, &tr->DB};
    return fetch_tensor(cs[call], tensor, out, shape4, (cudaStream_t)stream);
}

// ---- data parallel ---------------------------------------------------------------------------
extern "C" int cg_comm_unique_id(char id_out[128]) {
    CG_TRY(nccl_load());
    nccl_uid id;
    CG_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, id.internal, 128);
    return CG_OK;
}

extern "C" int cg_trainer_comm_init(cg_trainer_t tr, const char id_in[128], int rank, int world) {
    if (!tr || !id_in || world < 1 || rank < 0 || rank >= world) { cg_set_error("comm_init: bad argument"); return CG_ERR_INVALID; }
    CG_TRY(nccl_load());
    nccl_uid id;
    memcpy(id.internal, id_in, 128);
    CG_NCCL(g_nccl.CommInitRank(&tr->comm, world, id, rank));
    tr->world = world; tr->rank = rank;
    CG_CUDA(cudaStreamCreateWithFlags(&tr->comm_stream, cudaStreamNonBlocking));
    CG_CUDA(cudaEventCreateWithFlags(&tr->ev_d, cudaEventDisableTiming));
    CG_CUDA(cudaEventCreateWithFlags(&tr->ev_g, cudaEventDisableTiming));
    CG_CUDA(cudaEventCreateWithFlags(&tr->ev_done, cudaEventDisableTiming));
    tr->ev_bucket.resize(16);
    for (cudaEvent_t& e : tr->ev_bucket) CG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (int g = 0; g < 2; ++g) {          // the step body changes (all-reduces): captured graphs are stale
        if (tr->graph_exec[g]) { cudaGraphExecDestroy(tr->graph_exec[g]); tr->graph_exec[g] = nullptr; }
        tr->graph_seen[g] = 0;
    }
    return CG_OK;
}

