// Microbenchmark: per-SM throughput of the instructions the stack kernel leans on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
// Each kernel runs ITER iterations of UNROLL independent dependent-chains per thread;
// reports lane-ops / clk / SM using clock64 deltas on one SM-resident block set.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
constexpr int NCHAIN = 8;

template <int OP>
__global__ void __launch_bounds__(1024) k(float *out, const float *in, long long *cyc) {
    float a[NCHAIN], b[NCHAIN];
    double d[NCHAIN];
    unsigned u[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) {
        a[i] = in[threadIdx.x + i * 32];
        b[i] = in[threadIdx.x + i * 32 + 7];
        d[i] = (double)a[i];
        u[i] = __float_as_uint(b[i]);
    }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) {
            if (OP == 0) {  // FMNMX pair (compare-exchange)
                float lo = fminf(a[i], b[i]);
                float hi = fmaxf(a[i], b[i]);
                a[i] = lo; b[i] = hi + 1.0f;  // FADD keeps the chain alive (fma pipe)
            } else if (OP == 1) {  // FADD only
                a[i] = a[i] + b[i];
            } else if (OP == 2) {  // F2F.F64.F32 + DADD
                d[i] += (double)a[i];
                a[i] = a[i] + 1.0f;
            } else if (OP == 3) {  // DADD only
                d[i] += 1.5;
            } else if (OP == 4) {  // DFMA
                d[i] = fma(d[i], 1.0000001, 0.5);
            } else if (OP == 5) {  // 3-input min
                float r;
                asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a[i]), "f"(b[i]), "f"(a[(i + 1) % NCHAIN]));
                a[i] = r + 1.0f;
            } else if (OP == 6) {  // integer CE: IMNMX + LOP3 xor
                unsigned x = __float_as_uint(a[i]);
                unsigned lo = min(x, u[i]);
                unsigned hi = x ^ u[i] ^ lo;
                a[i] = __uint_as_float(lo); u[i] = hi + 1u;
            } else if (OP == 7) {  // IMAD.WIDE based f32->f64 + DADD
                unsigned long long w = (unsigned long long)__float_as_uint(a[i]) * 536870912ull + 0x3800000000000000ull;
                d[i] += __longlong_as_double((long long)w);
                a[i] = a[i] + 1.0f;
            } else if (OP == 8) {  // FSETP + predicated select
                a[i] = (a[i] - b[i] >= 0.25f && a[i] - b[i] <= 8.0f) ? a[i] + 1.0f : b[i];
            } else if (OP == 9) {  // pure FMNMX x2 (no FADD)
                float lo = fminf(a[i], b[i]);
                float hi = fmaxf(a[i], b[i]);
                a[i] = lo; b[i] = hi;
            }
        }
    }
    long long t1 = clock64();
    float s = 0; double sd = 0; unsigned su = 0;
    for (int i = 0; i < NCHAIN; ++i) { s += a[i] + b[i]; sd += d[i]; su += u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)sd + (float)su;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int ops_per_iter, float *out, float *in, long long *cyc) {
    const int blocks = 148, threads = 1024;
    k<OP><<<blocks, threads>>>(out, in, cyc);
    cudaDeviceSynchronize();
    k<OP><<<blocks, threads>>>(out, in, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)h[i];
    avg /= blocks;
    double laneops = (double)threads * ITER * NCHAIN * ops_per_iter;
    printf("%-34s %8.1f lane-ops/clk/SM  (%.0f cyc, %s)\n", name, laneops / avg, avg, cudaGetErrorString(e));
}

int main() {
    float *out, *in;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&in, 4096 * 4);
    cudaMalloc(&cyc, 148 * 8);
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = 1.0f + (i % 97) * 0.37f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<9>("FMNMX min+max (2 ops)", 2, out, in, cyc);
    run<0>("FMNMX min+max + FADD (3 ops)", 3, out, in, cyc);
    run<1>("FADD (1 op)", 1, out, in, cyc);
    run<2>("F2F.F64.F32 + DADD + FADD (3 ops)", 3, out, in, cyc);
    run<3>("DADD (1 op)", 1, out, in, cyc);
    run<4>("DFMA (1 op)", 1, out, in, cyc);
    run<5>("min3.f32 + FADD (2 ops)", 2, out, in, cyc);
    run<6>("IMNMX + LOP3 + IADD (3 ops)", 3, out, in, cyc);
    run<7>("IMAD.WIDE cvt + DADD + FADD (3)", 3, out, in, cyc);
    run<8>("FADDx2 FSETPx2 SEL FADD (~6)", 6, out, in, cyc);
    return 0;
}
