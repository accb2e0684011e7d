// Error slot + device queries shared by all translation units.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";
unsigned long long g_al_launches = 0;

void al_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int al_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

AL_API const char* al_last_error() { return g_err; }
AL_API int al_abi_version() { return 1; }
AL_API int al_sm_count() { return al_num_sms(); }
AL_API unsigned long long al_launch_count() { return g_al_launches; }
