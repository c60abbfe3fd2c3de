#pragma once
#include <cuda_runtime.h>
bool prof_enabled();
int prof_begin(cudaStream_t st);                       // -1 when profiling is off
void prof_end(int idx, cudaStream_t st, double flops);
