// Microbenchmark: do FP64 math and shared-memory 16-byte loads overlap on sm_100a?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ub_pipes ub_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NL, int OP>   // NF fp64 ops + NL LDS.128 per inner iteration; OP 0=DFMA 1=DADD 2=DMUL
__global__ void __launch_bounds__(128) k(double* out, int iters, double a, double b) {
    __shared__ double2 sm[128 * 16];
    for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) sm[i] = make_double2(i, 1.0);
    __syncthreads();
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + i + threadIdx.x;
    double2 acc = make_double2(0, 0);
    int idx = threadIdx.x, iacc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < (NF > NL ? NF : NL); ++j) {
            if (j < NF) {
                if (OP == 0) x[j % 8] = fma(x[j % 8], b, a);
                else if (OP == 1) x[j % 8] = x[j % 8] + a;
                else x[j % 8] = x[j % 8] * b;
            }
            if (j < NL) {
                double2 v = sm[(threadIdx.x + ((j + it) & 15) * 128)];
                iacc ^= __double2loint(v.x) ^ __double2hiint(v.y);
            }
        }
    }
    double s = acc.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 1234.5678 || iacc == 0x12345) out[0] = s + idx + iacc;
}

template <int NF, int NL, int OP>
float run(int blocks, int iters) {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NF, NL, OP><<<blocks, 128>>>(d, 10, 1.0, 1.0000001);
    cudaEventRecord(e0);
    k<NF, NL, OP><<<blocks, 128>>>(d, iters, 1.0, 1.0000001);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(d);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, iters = 4000;   // 32 warps/SM
    const double clk = 1.965e9;
    auto rep = [&](const char* name, float ms, int nf, int nl) {
        double cyc = ms * 1e-3 * clk;             // cycles total
        double winst_f = (double)iters * nf * 32; // fp64 warp-instr per SM (32 warps)
        double winst_l = (double)iters * nl * 32;
        printf("%-28s %8.3f ms  fp64 warp-instr/clk/SM %.3f   LDS.128 warp-instr/clk/SM %.3f (B/clk/SM %.1f)\n", name, ms,
               winst_f / cyc, winst_l / cyc, winst_l * 512 / cyc);
    };
    rep("DFMA x16", run<16, 0, 0>(blocks, iters), 16, 0);
    rep("DADD x16", run<16, 0, 1>(blocks, iters), 16, 0);
    rep("DMUL x16", run<16, 0, 2>(blocks, iters), 16, 0);
    rep("LDS x16", run<0, 16, 0>(blocks, iters), 0, 16);
    rep("DFMA x16 + LDS x4", run<16, 4, 0>(blocks, iters), 16, 4);
    rep("DFMA x16 + LDS x8", run<16, 8, 0>(blocks, iters), 16, 8);
    rep("DFMA x16 + LDS x16", run<16, 16, 0>(blocks, iters), 16, 16);
    rep("DADD x16 + LDS x8", run<16, 8, 1>(blocks, iters), 16, 8);
    rep("DFMA x8 + LDS x16", run<8, 16, 0>(blocks, iters), 8, 16);
    return 0;
}
