// CUDA-event timing of one kernel family inside a running step (bench.py's roofline numbers):
// every launch of the family is bracketed by two events on the launching stream.
#include <mutex>
#include <vector>

#include "common.h"
#include "prof.h"

namespace {
struct Pair { cudaEvent_t a, b; };
std::mutex mu;
bool on = false;
std::vector<Pair> pool;
size_t used = 0;
double flops = 0.0;
}   // namespace

bool prof_enabled() { return on; }

int prof_begin(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(mu);
    if (!on) return -1;
    if (used == pool.size()) {
        Pair p;
        if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return -1;
        pool.push_back(p);
    }
    cudaEventRecord(pool[used].a, st);
    return (int)used++;
}

void prof_end(int idx, cudaStream_t st, double fl) {
    if (idx < 0) return;
    std::lock_guard<std::mutex> lk(mu);
    cudaEventRecord(pool[idx].b, st);
    flops += fl;
}

extern "C" int cg_prof_enable(int enable) {
    std::lock_guard<std::mutex> lk(mu);
    on = enable != 0;
    used = 0;
    flops = 0.0;
    return CG_OK;
}

extern "C" int cg_prof_read(double* total_ms, int64_t* launches, double* total_flops) {
    CG_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(mu);
    double ms = 0.0;
    for (size_t i = 0; i < used; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, pool[i].a, pool[i].b) == cudaSuccess) ms += t;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = (int64_t)used;
    if (total_flops) *total_flops = flops;
    used = 0;
    flops = 0.0;
    return CG_OK;
}
