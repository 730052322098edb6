// Issue-rate probe for the FP32 instruction forms an in-register FFT is made of (B200, sm_100a):
// scalar FFMA / FADD / FMUL with register and immediate operands, the packed f32x2 forms (fma/add/mul.rn.f32x2),
// and 64-bit shared-memory loads issued next to them.  Prints warp-instructions per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_probe fp32_probe.cu && ./fp32_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITER 4096
#define CHAINS 8

__device__ __forceinline__ unsigned long long pack(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float lo(unsigned long long v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a + b;
}

template <int MODE>
__global__ void probe(float* out, float s, float t, const float2* tab) {
    __shared__ float2 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = tab[i];
    __syncthreads();
    float a[CHAINS], b[CHAINS];
    unsigned long long p[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        a[c] = s + c + threadIdx.x;
        b[c] = t - c;
        p[c] = pack(a[c], b[c]);
    }
    const unsigned long long ps = pack(s, t), pt = pack(t, s);
    int idx = threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (MODE == 0) a[c] = fmaf(a[c], s, t);                       // FFMA r,r,r
            if (MODE == 1) a[c] = fmaf(a[c], 0.999f, 0.001f);             // FFMA imm
            if (MODE == 2) a[c] = a[c] + b[c];                            // FADD r,r
            if (MODE == 3) a[c] = a[c] * 0.999f;                          // FMUL imm
            if (MODE == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[c]) : "l"(ps), "l"(pt));
            if (MODE == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(ps));
            if (MODE == 6) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(ps));
            if (MODE == 7) {  // butterfly-like mix: 2 FADD + 2 FFMA(imm)
                const float u = a[c] + b[c], v = a[c] - b[c];
                a[c] = fmaf(u, 0.7f, v);
                b[c] = fmaf(v, 0.7f, -u);
            }
            if (MODE == 8) {  // the same work in packed form: add2 + sub2(as fma2 with -1) ...
                unsigned long long u, v;
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(u) : "l"(p[c]), "l"(ps));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(v) : "l"(u), "l"(pt), "l"(p[c]));
                p[c] = v;
            }
            if (MODE == 9) {  // FFMA imm next to a conflict-free LDS.64 every 4 FFMAs
                a[c] = fmaf(a[c], 0.999f, 0.001f);
                if ((c & 3) == 0) {
                    const float2 w = sm[(idx + c) & 1023];
                    b[c] += w.x;
                    idx += 32;
                }
            }
            if (MODE == 10) {  // LDS.64 only (conflict-free), 1 FADD each
                const float2 w = sm[(idx + 32 * c) & 1023];
                a[c] += w.x;
                b[c] += w.y;
            }
        }
        if (MODE == 10) idx += 7;
    }
    float r = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) r += a[c] + b[c] + lo(p[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, double instr_per_iter, float* out, const float2* tab, int sms, double mhz) {
    const int blocks = sms * 4, threads = 256;
    probe<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, tab);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, tab);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = double(blocks) * (threads / 32) * ITER * CHAINS * instr_per_iter;
    const double clocks = ms * 1e-3 * mhz * 1e6;
    printf("%-44s %8.3f ms  %6.2f warp-instr/clk/SM (at %.0f MHz)\n", name, ms, warp_instr / clocks / sms, mhz);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    float* out;
    float2* tab;
    cudaMalloc(&out, sizeof(float) * sms * 4 * 256);
    cudaMalloc(&tab, sizeof(float2) * 1024);
    cudaMemset(tab, 0, sizeof(float2) * 1024);
    printf("%s, %d SMs, nominal %.0f MHz (rates assume the nominal clock)\n", p.name, sms, mhz);
    run<0>("FFMA r,r,r", 1, out, tab, sms, mhz);
    run<1>("FFMA r,imm,imm", 1, out, tab, sms, mhz);
    run<2>("FADD r,r", 1, out, tab, sms, mhz);
    run<3>("FMUL r,imm", 1, out, tab, sms, mhz);
    run<4>("FFMA2 (fma.rn.f32x2)", 1, out, tab, sms, mhz);
    run<5>("FADD2 (add.rn.f32x2)", 1, out, tab, sms, mhz);
    run<6>("FMUL2 (mul.rn.f32x2)", 1, out, tab, sms, mhz);
    run<7>("butterfly mix 2 FADD + 2 FFMA(imm)", 4, out, tab, sms, mhz);
    run<8>("packed mix FADD2 + FFMA2", 2, out, tab, sms, mhz);
    run<9>("FFMA imm + LDS.64 every 4th", 1.25 + 0.25, out, tab, sms, mhz);
    run<10>("LDS.64 + 2 FADD", 3, out, tab, sms, mhz);
    return 0;
}
