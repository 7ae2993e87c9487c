// Microbenchmark: throughput of a register-resident 32-input sorting network (185 compare-exchanges)
// built from different compare-exchange (CE) forms, to find the form that spreads the work over the
// ALU and FMA pipes of an sm_100a SM.
//   form 0: lo = FMNMX.min, hi = FMNMX.max                        (2 ALU)
//   form 1: lo = FMNMX.min, hi = a + b - lo on integer bits, IMAD (1 ALU + 2 FMA)
//   form 2: lo = FMNMX.min, hi = a + b - lo, compiler's choice    (IADD3 expected: 2 ALU)
//   form 3: lo = FMNMX.min, hi = a ^ b ^ lo                       (LOP3: 2 ALU)
// MIX = n: every CE whose index % n == 0 uses form 0, the others form 1 (MIX 0 = all form F).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_ce ubench_ce.cu
#include <cstdio>
#include <cstdint>
#include <utility>
#include <cuda_runtime.h>

struct CePair { short a, b; };
struct Net32 {
    CePair pr[200];
    int count;
    constexpr Net32() : pr{}, count(0) {
        constexpr short g[60][2] = {
            {0, 13}, {1, 12}, {2, 15}, {3, 14}, {4, 8},   {5, 6},   {7, 11},  {9, 10},
            {0, 5},  {1, 7},  {2, 9},  {3, 4},  {6, 13},  {8, 14},  {10, 15}, {11, 12},
            {0, 1},  {2, 3},  {4, 5},  {6, 8},  {7, 9},   {10, 11}, {12, 13}, {14, 15},
            {0, 2},  {1, 3},  {4, 10}, {5, 11}, {6, 7},   {8, 9},   {12, 14}, {13, 15},
            {1, 2},  {3, 12}, {4, 6},  {5, 7},  {8, 10},  {9, 11},  {13, 14},
            {1, 4},  {2, 6},  {5, 8},  {7, 10}, {9, 13},  {11, 14},
            {2, 4},  {3, 6},  {9, 12}, {11, 13},
            {3, 5},  {6, 8},  {7, 9},  {10, 12},
            {3, 4},  {5, 6},  {7, 8},  {9, 10}, {11, 12},
            {6, 7},  {8, 9}};
        for (int blk = 0; blk < 2; ++blk)
            for (int i = 0; i < 60; ++i) {
                pr[count].a = (short)(blk * 16 + g[i][0]);
                pr[count].b = (short)(blk * 16 + g[i][1]);
                ++count;
            }
        const int P = 32, p = 16;
        for (int k = p; k >= 1; k >>= 1)
            for (int j = k % p; j + k < P; j += 2 * k)
                for (int i = 0; i < k; ++i)
                    if (i + j + k < P && (i + j) / (2 * p) == (i + j + k) / (2 * p)) {
                        pr[count].a = (short)(i + j);
                        pr[count].b = (short)(i + j + k);
                        ++count;
                    }
    }
};
struct NetHolder { static constexpr Net32 net{}; };

template <int FORM>
__device__ __forceinline__ void ce(float &a, float &b, int one, int mone) {
    const float lo = fminf(a, b);
    if constexpr (FORM == 0) {
        const float hi = fmaxf(a, b);
        a = lo; b = hi;
    } else if constexpr (FORM == 1) {
        // t = a + b (independent of lo), hi = t - lo; multipliers live in registers so ptxas must
        // keep real IMADs (fma pipe) instead of folding them into IADD3 (alu pipe)
        const int t = __float_as_int(a) * one + __float_as_int(b);
        const int hi = __float_as_int(lo) * mone + t;
        a = lo; b = __int_as_float(hi);
    } else if constexpr (FORM == 2) {
        const int hi = __float_as_int(a) + __float_as_int(b) - __float_as_int(lo);
        a = lo; b = __int_as_float(hi);
    } else {
        const int hi = __float_as_int(a) ^ __float_as_int(b) ^ __float_as_int(lo);
        a = lo; b = __int_as_float(hi);
    }
}

template <int FORM, int MIX, int... I>
__device__ __forceinline__ void sort_impl(float (&s)[32], int one, int mone, std::integer_sequence<int, I...>) {
    ((MIX > 0 && (I % MIX) == 0
          ? ce<0>(s[NetHolder::net.pr[I].a], s[NetHolder::net.pr[I].b], one, mone)
          : ce<FORM>(s[NetHolder::net.pr[I].a], s[NetHolder::net.pr[I].b], one, mone)),
     ...);
}

#define ITER 512
template <int FORM, int MIX>
__global__ void __launch_bounds__(256, 3) k(float *out, const float *in, long long *cyc, int one, int mone) {
    float s[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] = in[(threadIdx.x * 7 + i * 131) & 4095];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        sort_impl<FORM, MIX>(s, one, mone, std::make_integer_sequence<int, NetHolder::net.count>{});
        s[0] = fmaf(s[31], 1e-3f, s[0]);
    }
    const long long t1 = clock64();
    float acc = 0;
    bool sorted = true;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        acc += s[i] * (float)(i + 1);
        if (i && !(s[i - 1] <= s[i])) sorted = false;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = sorted ? acc : -1.0f;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int FORM, int MIX>
void run(const char *name, float *out, float *in, long long *cyc) {
    const int blocks = 148 * 3, threads = 256;
    k<FORM, MIX><<<blocks, threads>>>(out, in, cyc, 1, -1);
    cudaDeviceSynchronize();
    k<FORM, MIX><<<blocks, threads>>>(out, in, cyc, 1, -1);
    cudaError_t e = cudaDeviceSynchronize();
    static long long h[148 * 3];
    static float ho[148 * 3 * 256];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < blocks * threads; ++i) bad += ho[i] == -1.0f;
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)h[i];
    avg /= blocks;
    // 3 blocks x 256 threads resident per SM, each doing ITER sorts in `avg` cycles
    const double sorts_per_clk_sm = 3.0 * threads * ITER / avg;
    printf("%-44s %7.3f SM-cyc/sort  (%.0f cyc, unsorted=%d, %s)\n", name, 1.0 / sorts_per_clk_sm, avg,
           bad, cudaGetErrorString(e));
}

int main() {
    float *out, *in;
    long long *cyc;
    cudaMalloc(&out, 148 * 3 * 256 * 4);
    cudaMalloc(&in, 4096 * 4);
    cudaMalloc(&cyc, 148 * 3 * 8);
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = ((i * 2654435761u) % 10007) * 0.37f - 1500.0f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0, 0>("all FMNMX pairs", out, in, cyc);
    run<1, 0>("all FMNMX + 2 IMAD", out, in, cyc);
    run<2, 0>("all FMNMX + (a+b-lo) compiler", out, in, cyc);
    run<3, 0>("all FMNMX + LOP3 xor", out, in, cyc);
    run<1, 2>("1/2 FMNMX pairs, 1/2 IMAD form", out, in, cyc);
    run<1, 3>("1/3 FMNMX pairs, 2/3 IMAD form", out, in, cyc);
    run<1, 4>("1/4 FMNMX pairs, 3/4 IMAD form", out, in, cyc);
    run<1, 5>("1/5 FMNMX pairs, 4/5 IMAD form", out, in, cyc);
    return 0;
}
