// fake_cudart.cpp - test infrastructure: the handful of CUDA runtime entry points libchannel_b200 calls, implemented
// on host memory, so that the library's own sources (chb_api.cu, the launchers, restart_io.cu, ...) can be built
// with g++ next to the CTA emulator and the single-GPU test-suite can exercise the whole C ABI path on the CPU.
// Everything is synchronous: streams and events are tokens, "device" memory is malloc'ed.  Never linked into the
// product library.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdlib>
#include <cstring>

namespace {
struct FakeEvent { double t_ms; };
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
thread_local cudaError_t g_last = cudaSuccess;
}  // namespace

extern "C" {

cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { cudaError_t e = g_last; g_last = cudaSuccess; return e; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime error"; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    *v = (a == cudaDevAttrMultiProcessorCount) ? 148 : (a == cudaDevAttrMaxSharedMemoryPerBlockOptin ? 232448 : 0);
    return cudaSuccess;
}
cudaError_t cudaMemGetInfo(size_t* free_b, size_t* total_b) { *free_b = (size_t)4 << 30; *total_b = (size_t)8 << 30; return cudaSuccess; }

cudaError_t cudaMalloc(void** p, size_t n) {
    *p = n ? malloc(n) : malloc(1);
    return *p ? cudaSuccess : (g_last = cudaErrorMemoryAllocation);
}
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaHostRegister(void*, size_t, unsigned int) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned int) { *s = reinterpret_cast<cudaStream_t>(malloc(8)); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }

cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = reinterpret_cast<cudaEvent_t>(new FakeEvent{0.0}); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned int) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete reinterpret_cast<FakeEvent*>(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { reinterpret_cast<FakeEvent*>(e)->t_ms = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = (float)(reinterpret_cast<FakeEvent*>(b)->t_ms - reinterpret_cast<FakeEvent*>(a)->t_ms);
    return cudaSuccess;
}

cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }

cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return g_last = cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned int) { return g_last = cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

// microbench.cu is not part of the emulated build
int chb_measure_device_peaks(double*) { return 1; }

}  // extern "C"
