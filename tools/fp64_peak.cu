// fp64_peak.cu -- measured FP64 DFMA issue peak of the GPU (SURVEY "facts" / hard part 6: the secondary ceiling
// of the update and utility kernels).  N_CHAIN independent FMA chains per thread, enough warps to fill every
// scheduler; reports TFLOP/s (2 flops per DFMA) as one JSON line.  Build: tools/build_tools.sh
#include <cstdio>
#include <cuda_runtime.h>

template <int NCH>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) x[i] = (double)(threadIdx.x + i) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < NCH; ++i) x[i] = fma(x[i], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += x[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chains alive
}

template <int NCH>
static double run(int sms, int blocks_per_sm, int iters) {
    double* out;
    cudaMalloc(&out, sizeof(double) * 256 * sms * blocks_per_sm);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_dfma<NCH><<<sms * blocks_per_sm, 256>>>(out, iters / 10, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_dfma<NCH><<<sms * blocks_per_sm, 256>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * NCH * 8.0 * (double)iters * 256.0 * sms * blocks_per_sm;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaFree(out);
    return best;
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    const int sms = p.multiProcessorCount;
    double best = 0.0; int best_ch = 0, best_b = 0;
    for (int b = 2; b <= 8; b *= 2) {
        double t4 = run<4>(sms, b, 20000), t8 = run<8>(sms, b, 10000);
        if (t4 > best) { best = t4; best_ch = 4; best_b = b; }
        if (t8 > best) { best = t8; best_ch = 8; best_b = b; }
    }
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"fp64_tflops\": %.3f, \"dfma_per_clk_per_sm_at_max_clock\": %.2f, \"gpu\": \"%s\", \"sms\": %d, "
           "\"max_clock_mhz\": %.0f, \"chains_per_thread\": %d, \"ctas_per_sm\": %d, "
           "\"how\": \"independent DFMA chains, 256 threads/CTA, best of 5 launches, CUDA events\"}\n",
           best, best * 1e12 / 2.0 / sms / (clk * 1e3), p.name, sms, clk / 1e3, best_ch, best_b);
    return 0;
}
