// int_peak.cu -- measures the integer issue peaks of the GPU (MEASURED_PEAKS.json has no
// INT32/DPX figure, the builder measures it).  Each kernel runs 8 independent dependency chains per
// thread of one SASS op class; reported as G thread-ops/s over all SMs.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#define ITER 4096
template <int OP>
__global__ void __launch_bounds__(1024) k(int *out, int a0, int b0) {
    int v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * (i + 1) + a0;
    int b = b0 + threadIdx.x, c = b0 * 3 + 1;
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) v[i] = v[i] * b + c;                                  // IMAD
            else if (OP == 1) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(b), "r"(c)); }  // LOP3
            else if (OP == 2) v[i] = max(v[i], b + i) ;                         // VIMNMX (+iadd folded?)
            else if (OP == 3) v[i] = __vimax3_s32(v[i], b, c + i);             // VIMNMX3
            else if (OP == 4) v[i] = __dp4a((unsigned)v[i], (unsigned)b, (unsigned)c);  // IDP.4A acc chain on c? no: chain on v
            else if (OP == 5) { asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(v[i]) : "r"(b), "r"(c)); } // IDP.2A
            else if (OP == 6) v[i] = (v[i] > b) ? c : v[i] + 1;                 // ISETP+SEL(+IADD)
            else if (OP == 7) v[i] = __byte_perm(v[i], b, 0x5140 + i);          // PRMT
            else if (OP == 8) v[i] = __funnelshift_r(v[i], b, 8);               // SHF
            else if (OP == 9) v[i] = v[i] + b + c;                              // IADD3
            else if (OP == 10) { v[i] = v[i] * b + c; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[(i+4)&7]) : "r"(b), "r"(c)); } // IMAD + LOP3 mix
            else if (OP == 11) { v[i] = v[i] * b + c; v[(i + 4) & 7] = __dp4a((unsigned)v[(i + 4) & 7], (unsigned)b, (unsigned)c); }   // IMAD + IDP mix
            else if (OP == 12) { v[i] = __viaddmax_s32(v[i], b, c); }           // VIADDMNMX
            else if (OP == 13) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(b), "r"(c)); v[(i + 4) & 7] = __dp4a((unsigned)v[(i + 4) & 7], (unsigned)b, (unsigned)c); }  // LOP3 + IDP mix
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= v[i];
    if (s == 0x12345678) out[0] = s;
}

template <int OP>
double run(const char *name, int opsPerInner, int sms) {
    int *d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 2;
    k<OP><<<blocks, 1024>>>(d, 1, 2);
    cudaDeviceSynchronize();
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<OP><<<blocks, 1024>>>(d, 1, 2);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double gops = (double)blocks * 1024 * ITER * 8 * opsPerInner / (ms * 1e-3) / 1e9;
        if (gops > best) best = gops;
    }
    printf("{\"op\": \"%s\", \"gops\": %.1f}\n", name, best);
    cudaFree(d);
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    run<0>("IMAD", 1, sms); run<1>("LOP3", 1, sms); run<2>("VIMNMX", 1, sms); run<3>("VIMNMX3", 1, sms);
    run<4>("IDP4A", 1, sms); run<5>("IDP2A", 1, sms); run<6>("ISETP+SEL", 1, sms); run<7>("PRMT", 1, sms);
    run<8>("SHF", 1, sms); run<9>("IADD3", 1, sms); run<10>("IMAD+LOP3", 2, sms); run<11>("IMAD+IDP4A", 2, sms);
    run<12>("VIADDMNMX", 1, sms); run<13>("LOP3+IDP4A", 2, sms);
    return 0;
}
