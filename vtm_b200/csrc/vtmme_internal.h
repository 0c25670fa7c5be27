// Internal declarations shared by the .cu translation units of libvtmme.so.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/vtmme.h"

#define VTMME_CUDA_CHECK(ctx, call)                                                                   \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) return vtmme_set_error((ctx), VTMME_ERR_CUDA, #call, cudaGetErrorString(e__)); \
  } while (0)

int vtmme_set_error(vtmme_ctx* ctx, int code, const char* what, const char* detail);
