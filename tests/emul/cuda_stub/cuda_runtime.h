/*
 * cuda_stub/cuda_runtime.h — TEST INFRASTRUCTURE ONLY.
 * Lets the no-GPU suite compile the HOST logic of csrc/b2i_pipe.cpp (windows, ring,
 * worker threads, release / skip rules) against the oracle-backed shim
 * (tests/emul/b2i_shim.c): "pinned" memory is plain memory, nothing is ever on a device.
 */
#pragma once
#include <stdlib.h>
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum { cudaHostAllocPortable = 1 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { enum cudaMemoryType type; void *devicePointer; };
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned flags) { (void)flags; *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(struct cudaPointerAttributes *a, const void *p) { (void)p; a->type = cudaMemoryTypeUnregistered; a->devicePointer = 0; return cudaSuccess; }
static inline cudaError_t cudaGetLastError(void) { return cudaSuccess; }
