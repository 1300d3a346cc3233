// Dependent-issue latency of the fp64 add the ordered sums are made of (phase B of the tally kernel, log_choose in the
// call kernel): cycles per step of one chain, alone on an SM.  nvcc -arch=sm_100a -O3 -o ub_fp64_latency ub_fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chains(double *out, long long *cyc, const double *lut, int n)
{
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = lut[i];
    __syncthreads();
    double x = out[0];
    long long t0, t1;
    // 0: dependent DADD with a register operand
    double a = x, b = 1.0 + x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a = __dadd_rn(a, b);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 1: dependent DADD, operand from shared memory (addresses independent of the chain)
    double c = x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) c = __dadd_rn(c, sm[i & 1023]);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // 2: dependent DADD, operand from global memory through __ldg (L1-resident LUT)
    double d = x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) d = __dadd_rn(d, __ldg(lut + (i & 1023)));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // 3: dependent DFMA
    double e = x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) e = __fma_rn(e, 1.0000001, b);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // 4: dependent FADD (fp32) for comparison
    float f = (float)x, g = 1.0f + f;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) f = __fadd_rn(f, g);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // 5: two independent DADD chains interleaved (throughput vs latency)
    double p = x, q = x + 2.0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) { p = __dadd_rn(p, b); q = __dadd_rn(q, b); }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    out[threadIdx.x] = a + c + d + e + f + p + q;
}

int main()
{
    const int n = 1 << 14;
    double *out, *lut; long long *cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&lut, 1024 * 8); cudaMalloc(&cyc, 64);
    cudaMemset(out, 0, 1024 * 8); cudaMemset(lut, 0, 1024 * 8);
    const char *names[6] = {"DADD reg", "DADD + LDS operand", "DADD + LDG operand", "DFMA", "FADD", "2 interleaved DADD chains (per pair)"};
    for (int threads = 1; threads <= 32; threads *= 32) {
        chains<<<1, threads>>>(out, cyc, lut, n);
        chains<<<1, threads>>>(out, cyc, lut, n);
        long long h[6];
        if (cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 1; }
        for (int k = 0; k < 6; ++k) printf("threads=%d %-40s %.2f cycles/step\n", threads, names[k], (double)h[k] / n);
    }
    return 0;
}
