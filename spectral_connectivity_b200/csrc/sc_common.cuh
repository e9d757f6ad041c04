// Shared helpers for the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sc_b200.h"

#define SC_NUM_SMS_FALLBACK 148

void sc_set_error(const char* fmt, ...);
int sc_num_sms();
int sc_max_smem_optin();

#define SC_CHECK_ARG(cond, ...)               \
    do {                                      \
        if (!(cond)) {                        \
            sc_set_error(__VA_ARGS__);        \
            return SC_ERR_INVALID_ARGUMENT;   \
        }                                     \
    } while (0)

#define SC_CUDA_OK(expr)                                                                   \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            sc_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                        \
            return SC_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

#define SC_LAUNCH_OK()                                                                        \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            sc_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                           \
            return SC_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

struct ScMap {
    long long bw, bt, bk, rw, rt, rk;
};
