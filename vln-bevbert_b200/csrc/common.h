// Host-side helpers shared by the C-ABI translation units: thread-local error string, launch counter.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bb {
int set_error(const char* msg);          // stores msg, returns -1
int check_launch(const char* what);      // cudaGetLastError -> 0 / -1 (+message)
void count_launch(int n = 1);
}  // namespace bb
