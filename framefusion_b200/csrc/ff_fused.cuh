// Single-pass fused merge kernel (threshold branch): placeholder until the kernel lands.
#pragma once
#include "ff_common.cuh"
#include "ff_merge.cuh"

namespace ff {

constexpr int FUSED_MIN_ROWS = 4;

inline int launch_fused(int, int64_t*, int64_t*, int64_t*, const int*, int*, uint8_t*, float*, int*, int*,
                        unsigned long long*, const int*, const void*, void*, int, int64_t, int64_t, double, double,
                        const AuxPack&, cudaStream_t) {
    return FF_E_UNSUPPORTED;
}

}  // namespace ff
