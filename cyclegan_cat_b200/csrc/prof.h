#pragma once
#include <cuda_runtime.h>
bool prof_enabled();
int prof_begin(cudaStream_t st);                       // -1 when profiling is off
void prof_end(int idx, cudaStream_t st, double flops, unsigned long long key = 0);
// key layout: kind(4) | taps(8) | cchunks(8) | bn(12) | tiles_per_img(20) | nb(12)
static inline unsigned long long prof_key(int kind, int taps, int cch, int bn, int tiles, int nb) {
    return ((unsigned long long)kind << 60) | ((unsigned long long)(taps & 0xff) << 52) | ((unsigned long long)(cch & 0xff) << 44) |
           ((unsigned long long)(bn & 0xfff) << 32) | ((unsigned long long)(tiles & 0xfffff) << 12) | (unsigned long long)(nb & 0xfff);
}
