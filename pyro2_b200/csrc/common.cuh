// common.cuh -- error plumbing shared by the .cu translation units of libpyro2b200.so
#pragma once
#ifdef P2B_EMU_HEADER
#include P2B_EMU_HEADER   // tests/emu only: the host emulation of the CUDA execution model
#else
#include <cuda_runtime.h>
#define P2B_LAUNCH(kernel, grid, block, smem, stream) kernel<<<grid, block, smem, stream>>>
#define P2B_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pyro2b200.h"

namespace pyro {

// thread-local last-error text (no global mutable state except this string)
char* last_error_buf();
void set_error(const char* fmt, ...);

#define P2B_CUDA_CHECK(expr)                                                                  \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            pyro::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,              \
                            cudaGetErrorString(e__));                                         \
            return P2B_ECUDA;                                                                 \
        }                                                                                     \
    } while (0)

#define P2B_REQUIRE(cond, msg)                                                                \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            pyro::set_error("invalid argument: %s (%s:%d)", msg, __FILE__, __LINE__);         \
            return P2B_EINVAL;                                                                \
        }                                                                                     \
    } while (0)

inline int num_sms()
{
    static int sms[64] = {0};                 // per device ordinal
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!sms[dev]) {
        cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (sms[dev] <= 0) sms[dev] = 148;
    }
    return sms[dev];
}

}  // namespace pyro
