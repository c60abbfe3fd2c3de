// CUDA-event timing of one kernel family inside a running step (bench.py's roofline numbers):
// every launch of the family is bracketed by two events on the launching stream.
#include <stdio.h>
#include <stdlib.h>

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
std::vector<double> rec_flops;
std::vector<unsigned long long> rec_key;
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

void prof_end(int idx, cudaStream_t st, double fl, unsigned long long key) {
    if (idx < 0) return;
    std::lock_guard<std::mutex> lk(mu);
    cudaEventRecord(pool[idx].b, st);
    if ((key >> 60) < 8) flops += fl;
    if (rec_flops.size() <= (size_t)idx) { rec_flops.resize(idx + 1); rec_key.resize(idx + 1); }
    rec_flops[idx] = fl;
    rec_key[idx] = key;
}

extern "C" int cg_prof_enable(int enable) {
    std::lock_guard<std::mutex> lk(mu);
    if (enable < 0) {       // pause: stop recording, keep what has been recorded for cg_prof_read
        on = false;
        return CG_OK;
    }
    on = enable != 0;
    used = 0;
    flops = 0.0;
    while ((long long)pool.size() < (long long)enable) {      // enable > 1: pre-create that many event pairs (keeps creation out of a timed loop)
        Pair p;
        CG_CUDA(cudaEventCreate(&p.a));
        CG_CUDA(cudaEventCreate(&p.b));
        pool.push_back(p);
    }
    return CG_OK;
}

extern "C" int cg_prof_read(double* total_ms, int64_t* launches, double* total_flops) {
    CG_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(mu);
    double ms = 0.0;
    const char* dump = getenv("CG_PROF_DUMP");       // per-launch CSV: key, flops, ms
    FILE* f = dump ? fopen(dump, "w") : nullptr;
    if (f) fprintf(f, "kind,taps,cchunks,bn,tiles,nb,flops,ms\n");
    for (size_t i = 0; i < used; ++i) {
        float t = 0.f;
        const bool tc = i >= rec_key.size() || (rec_key[i] >> 60) < 8;     // kinds >= 8: streaming kernels (CG_PROF_STREAM=1),
        if (cudaEventElapsedTime(&t, pool[i].a, pool[i].b) == cudaSuccess && tc) ms += t;      // listed in the CSV only
        if (f && i < rec_key.size()) {
            const unsigned long long k = rec_key[i];
            fprintf(f, "%llu,%llu,%llu,%llu,%llu,%llu,%.0f,%.5f\n", k >> 60, (k >> 52) & 0xff, (k >> 44) & 0xff, (k >> 32) & 0xfff,
                    (k >> 12) & 0xfffff, k & 0xfff, rec_flops[i], t);
        }
    }
    if (f) fclose(f);
    if (total_ms) *total_ms = ms;
    if (launches) *launches = (int64_t)used;
    if (total_flops) *total_flops = flops;
    used = 0;
    flops = 0.0;
    return CG_OK;
}
