/* mc_internal.h — what the translation units of libmicroclimate_b200.so share besides the public header (hidden symbols, not ABI). */
#pragma once
#ifndef MCW_HOST_EMULATION
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include "../../include/microclimate_b200.h"
#define MC_HIDDEN __attribute__((visibility("hidden")))
/* mc_api.cu */
MC_HIDDEN int mc_i_device(const mc_handle* h);
MC_HIDDEN cudaStream_t mc_i_stream(const mc_handle* h);               /* the handle's compute stream */
MC_HIDDEN int64_t mc_i_sites(const mc_handle* h);
MC_HIDDEN const mc_model_desc* mc_i_model(const mc_handle* h);
MC_HIDDEN int mc_i_fail(mc_handle* h, const char* msg, int rc);       /* sets mc_last_error(h), returns rc */
MC_HIDDEN int mc_i_multi_fail(mc_multi* m, int device_index, const char* msg, int rc);   /* sets mc_multi_last_error(m) to "device g: msg", returns rc */
/* mc_weather.cu */
MC_HIDDEN void mc_i_weather_release(mc_handle* h);                     /* frees the weather staging of a handle (called by mc_destroy) */
