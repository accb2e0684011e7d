// Shared helpers for the autolabel_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define AL_API extern "C" __attribute__((visibility("default")))

// Last-error slot (per host thread), filled by AL_CHECK / al_fail.
void al_set_error(const char* fmt, ...);

#define AL_CHECK(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            al_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,               \
                         cudaGetErrorString(_e));                                   \
            return (int)_e;                                                         \
        }                                                                           \
    } while (0)

// Every kernel launch of this library passes through here: counted (al_launch_count) and checked.
extern unsigned long long g_al_launches;
#define AL_LAUNCH_CHECK()                 \
    do {                                  \
        ++g_al_launches;                  \
        AL_CHECK(cudaGetLastError());     \
    } while (0)

#define AL_REQUIRE(cond, msg)                                                       \
    do {                                                                            \
        if (!(cond)) {                                                              \
            al_set_error("%s:%d invalid argument: %s (%s)", __FILE__, __LINE__,     \
                         msg, #cond);                                               \
            return (int)cudaErrorInvalidValue;                                      \
        }                                                                           \
    } while (0)

static inline unsigned int al_div_up(unsigned long long a, unsigned long long b) {
    return (unsigned int)((a + b - 1) / b);
}

// Number of SMs of the current device (cached; B200 = 148).
int al_num_sms();

__device__ __forceinline__ float al_clampf(float x, float lo, float hi) {
    return fminf(hi, fmaxf(lo, x));
}
