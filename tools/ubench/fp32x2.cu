// Issue-rate microbenchmark for scalar vs packed fp32 on sm_100a: warp instructions per cycle per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2 fp32x2.cu && ./fp32x2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 512
#define CH 8
enum { FADD_RR, FMUL_RR, FFMA_RRR, FFMA_RIR, FADD2_RR, FMUL2_RR, FFMA2_RRR, FADD2_RI, FMUL2_RI, FFMA2_RIR, IDP2A, IMAD_RRR, SHF_RI, MIX_FADD2_SHF, N_KINDS };
const char* NAMES[] = {"FADD r,r", "FMUL r,r", "FFMA r,r,r", "FFMA r,imm,r", "FADD2 r,r", "FMUL2 r,r", "FFMA2 r,r,r", "FADD2 r,imm", "FMUL2 r,imm",
                       "FFMA2 r,imm,r", "IDP.2A", "IMAD r,r,r", "SHF r,imm", "FADD2 + SHF interleaved"};

template <int KIND>
__global__ void k(float* out, float seed, long long* cycles) {
    float a[CH], b = seed + (float)(threadIdx.x & 1), c = seed * 0.5f + (float)(threadIdx.x & 2);
    u64 p[CH], pb, pc;
    unsigned ii[CH];
    for (int i = 0; i < CH; ++i) { a[i] = seed + i; ii[i] = (unsigned)(seed) + i; asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[i]), "f"(a[i] + 1.f)); }
    asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c), "f"(c));
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (KIND == FADD_RR) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
                if (KIND == FMUL_RR) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
                if (KIND == FFMA_RRR) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                if (KIND == FFMA_RIR) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+f"(a[i]) : "f"(c));
                if (KIND == FADD2_RR) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                if (KIND == FMUL2_RR) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                if (KIND == FFMA2_RRR) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
                if (KIND == FADD2_RI) asm volatile("{.reg .b64 t; mov.b64 t, {0f3F800001, 0f3F800001}; add.rn.f32x2 %0, %0, t;}" : "+l"(p[i]));
                if (KIND == FMUL2_RI) asm volatile("{.reg .b64 t; mov.b64 t, {0f3F800001, 0f3F800001}; mul.rn.f32x2 %0, %0, t;}" : "+l"(p[i]));
                if (KIND == FFMA2_RIR) asm volatile("{.reg .b64 t; mov.b64 t, {0f3F800001, 0f3F800001}; fma.rn.f32x2 %0, %0, t, %1;}" : "+l"(p[i]) : "l"(pc));
                if (KIND == IDP2A) asm volatile("dp2a.lo.u32.u32 %0, %1, %0, %0;" : "+r"(ii[i]) : "r"(0x12340567u));
                if (KIND == IMAD_RRR) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(ii[i]) : "r"(ii[(i + 1) % CH] | 1u), "r"(it));
                if (KIND == SHF_RI) asm volatile("shf.r.wrap.b32 %0, %0, %1, 14;" : "+r"(ii[i]) : "r"(0x12C0u));
                if (KIND == MIX_FADD2_SHF) {
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                    asm volatile("shf.r.wrap.b32 %0, %0, %1, 14;" : "+r"(ii[i]) : "r"(0x12C0u));
                }
            }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < CH; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += a[i] + lo + hi + (float)ii[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(int warps_per_smsp) {
    float* out; long long* cyc;
    const int threads = 128 * warps_per_smsp, blocks = 148;
    cudaMalloc(&out, sizeof(float) * threads * blocks);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    k<KIND><<<blocks, threads>>>(out, 1.0f, cyc);
    k<KIND><<<blocks, threads>>>(out, 1.0f, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += h[i];
    avg /= blocks;
    const double per_instr = (KIND == MIX_FADD2_SHF ? 2.0 : 1.0);
    const double instr = (double)ITERS * 4 * CH * per_instr * warps_per_smsp;       // warp instructions per SMSP
    printf("%-26s warps/SMSP %d : %.3f warp-instr/cycle/SMSP  (%.2f cycles per instr)\n", NAMES[KIND], warps_per_smsp, instr / avg, avg / instr);
    cudaFree(out); cudaFree(cyc);
}
template <int K> void all() { run<K>(1); run<K>(4); run<K>(8); }
int main() {
    all<FADD_RR>(); all<FMUL_RR>(); all<FFMA_RRR>(); all<FFMA_RIR>(); all<FADD2_RR>(); all<FMUL2_RR>(); all<FFMA2_RRR>(); all<FADD2_RI>();
    all<FMUL2_RI>(); all<FFMA2_RIR>(); all<IDP2A>(); all<IMAD_RRR>(); all<SHF_RI>(); all<MIX_FADD2_SHF>();
    return 0;
}
