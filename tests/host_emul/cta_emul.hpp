// cta_emul.hpp - test infrastructure: run a CUDA kernel's SOURCE on the CPU, one OS thread per CUDA thread
// of a thread block, `__syncthreads()` = a pthread barrier, shared memory = a process-wide buffer (one
// block runs at a time).  Enough of the CUDA surface for the FFT-pass kernels of channel_b200/csrc
// (blockIdx/threadIdx, dynamic and static shared memory, __ldg, warp shuffles, atomicMax on a 64-bit word);
// TMA / cp.async / mbarrier are replaced by plain copies where a kernel uses them (CHB_HOST_EMUL branches in
// fft_regs.cuh).  It checks kernel LOGIC (indexing, twiddles, layouts, barriers between phases) before a GPU
// minute is spent; it says nothing about performance and the product never loads it.
#pragma once
#include <cuda_runtime.h>
#include <pthread.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>
#include <math.h>
#include <cmath>
using std::max;
using std::min;

namespace cta_emul {
inline thread_local uint3 t_threadIdx, t_blockIdx;
inline thread_local int t_linear_tid;
inline dim3 g_blockDim, g_gridDim;
inline pthread_barrier_t g_block_barrier;
inline std::vector<pthread_barrier_t> g_warp_barrier;
inline double g_warp_buf[64][32];                 // [warp][lane] exchange buffer for shuffles (blocks of <= 2048 threads)
alignas(16) inline unsigned char g_dyn_smem[232448];   // 227 KB of dynamic shared memory
}  // namespace cta_emul

#define blockIdx cta_emul::t_blockIdx
#define threadIdx cta_emul::t_threadIdx
#define blockDim cta_emul::g_blockDim
#define gridDim cta_emul::g_gridDim
#undef __launch_bounds__
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define CHB_EMUL_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(cta_emul::g_dyn_smem)

template <class T>
static inline T __ldg(const T* p) { return *p; }
template <class T>
static inline T __ldcg(const T* p) { return *p; }
template <class T>
static inline T __ldcs(const T* p) { return *p; }
template <class T>
static inline void __stcg(T* p, T v) { *p = v; }
template <class T>
static inline void __stcs(T* p, T v) { *p = v; }
static inline void __syncthreads() { pthread_barrier_wait(&cta_emul::g_block_barrier); }
static inline double __shfl_xor_sync(unsigned, double v, int o) {
    const int tid = cta_emul::t_linear_tid, w = tid >> 5, l = tid & 31;
    cta_emul::g_warp_buf[w][l] = v;
    pthread_barrier_wait(&cta_emul::g_warp_barrier[w]);
    const double r = cta_emul::g_warp_buf[w][l ^ o];
    pthread_barrier_wait(&cta_emul::g_warp_barrier[w]);
    return r;
}
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline unsigned long long atomicMax(unsigned long long* a, unsigned long long v) {
    auto* p = reinterpret_cast<std::atomic<unsigned long long>*>(a);
    unsigned long long old = p->load();
    while (old < v && !p->compare_exchange_weak(old, v)) {}
    return old;
}

// cuda_runtime.h declares this convenience overload for nvcc only
template <class T>
static inline cudaError_t cudaFuncSetAttribute(T* /*kernel*/, cudaFuncAttribute, int) { return cudaSuccess; }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }

namespace cta_emul {
// run kern(args...) for every block of the grid; blocks of up to 2048 threads (a multiple of 32), 1-D to 3-D
template <class K, class... Args>
void launch(K kern, dim3 grid, dim3 block, Args... args) {
    const int threads = (int)(block.x * block.y * block.z);
    if (threads % 32 != 0 || threads > 2048) { fprintf(stderr, "cta_emul: block size %d\n", threads); abort(); }
    g_blockDim = block;
    g_gridDim = grid;
    pthread_barrier_init(&g_block_barrier, nullptr, threads);
    g_warp_barrier.resize(threads / 32);
    for (auto& b : g_warp_barrier) pthread_barrier_init(&b, nullptr, 32);
    // one OS thread per CUDA thread of a block, created once per launch; the blocks of the grid run one after the
    // other (a barrier between them: shared memory is reused)
    std::vector<std::thread> th;
    th.reserve(threads);
    for (int t = 0; t < threads; ++t)
        th.emplace_back([=]() {
            t_threadIdx = make_uint3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            t_linear_tid = t;
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        t_blockIdx = make_uint3(bx, by, bz);
                        kern(args...);
                        pthread_barrier_wait(&g_block_barrier);
                    }
        });
    for (auto& x : th) x.join();
    for (auto& b : g_warp_barrier) pthread_barrier_destroy(&b);
    pthread_barrier_destroy(&g_block_barrier);
}
template <class K, class... Args>
void launch(K kern, dim3 grid, int threads, Args... args) { launch(kern, grid, dim3(threads, 1, 1), args...); }

// CHB_LAUNCH(grid, block, smem, stream, kernel)(args...) of the product sources in the emulation build
template <class K>
struct Launcher {
    K kern;
    dim3 grid, block;
    template <class... Args>
    void operator()(Args... args) const { launch(kern, grid, block, args...); }
};
template <class K, class G, class B>
Launcher<K> launcher(K kern, G grid, B block) { return Launcher<K>{kern, dim3(grid), dim3(block)}; }
}  // namespace cta_emul
