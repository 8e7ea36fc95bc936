// Throughput microbenchmark of the integer instructions the field arithmetic is made of (sm_100a).
// Each kernel runs ITERS x 16 independent-chain instructions per thread; prints warp-instructions / clk / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
#define K16(OP) OP(0) OP(1) OP(2) OP(3) OP(4) OP(5) OP(6) OP(7) OP(8) OP(9) OP(10) OP(11) OP(12) OP(13) OP(14) OP(15)
template <int MODE>
__global__ void bench(uint32_t* out, uint32_t seed) {
    uint32_t a[16]; uint64_t w[16];
    for (int i = 0; i < 16; i++) { a[i] = seed * (i + 1) + threadIdx.x; w[i] = (uint64_t)a[i] * 0x9E3779B97F4A7C15ull; }
    uint32_t m = seed | 1, c = seed + 7;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {  // IMAD 32-bit lo
#define OP(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c));
            K16(OP)
#undef OP
        } else if (MODE == 1) {  // IMAD.WIDE.U32 with 64-bit accumulate
#define OP(i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(m));
            K16(OP)
#undef OP
        } else if (MODE == 2) {  // IADD3 (3-input add)
#define OP(i) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a[i]) : "r"(m), "r"(c));
            K16(OP)
#undef OP
        } else if (MODE == 3) {  // LOP3
#define OP(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m), "r"(c));
            K16(OP)
#undef OP
        } else if (MODE == 4) {  // SHF (funnel shift)
#define OP(i) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(m));
            K16(OP)
#undef OP
        } else if (MODE == 5) {  // LEA-like: (a << 3) + c
#define OP(i) asm volatile("{.reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1;}" : "+r"(a[i]) : "r"(c));
            K16(OP)
#undef OP
        } else if (MODE == 6) {  // 64-bit add (IADD3 + IADD3.X)
#define OP(i) asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) & 15]));
            K16(OP)
#undef OP
        } else if (MODE == 7) {  // mixed: 1 IMAD.WIDE + 1 IADD3 per pair (dual-pipe co-issue)
#define OP(i) asm volatile("mad.wide.u32 %0, %2, %3, %0;\n\tlop3.b32 %1, %1, %3, %4, 0x96;" : "+l"(w[i]), "+r"(a[i]) : "r"(a[(i+1)&15]), "r"(m), "r"(c));
            K16(OP)
#undef OP
        } else if (MODE == 8) {  // mul.hi.u32
#define OP(i) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
            K16(OP)
#undef OP
        } else if (MODE == 9) {  // FFMA (fp32) for reference
#define OP(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&a[i]) : "f"(*(float*)&m), "f"(*(float*)&c));
            K16(OP)
#undef OP
        } else if (MODE == 10) {  // PRMT
#define OP(i) asm volatile("prmt.b32 %0, %0, %1, 0x3201;" : "+r"(a[i]) : "r"(m));
            K16(OP)
#undef OP
        } else if (MODE == 11) {  // mixed 1 IMAD.WIDE : 2 LOP3
#define OP(i) asm volatile("mad.wide.u32 %0, %2, %3, %0;\n\tlop3.b32 %1, %1, %3, %4, 0x96;\n\tlop3.b32 %1, %1, %4, %3, 0x96;" : "+l"(w[i]), "+r"(a[i]) : "r"(a[(i+1)&15]), "r"(m), "r"(c));
            K16(OP)
#undef OP
        }
    }
    uint32_t r = 0;
    for (int i = 0; i < 16; i++) r ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name, int per_iter, int sms, float mhz) {
    uint32_t* d; cudaMalloc(&d, sms * 8 * 256 * 4);
    bench<MODE><<<sms * 8, 256>>>(d, 12345);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    bench<MODE><<<sms * 8, 256>>>(d, 12345);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double winst = (double)sms * 8 * 8 * ITERS * per_iter;  // warp instructions
    double clk = ms * 1e-3 * mhz * 1e6;
    printf("%-34s %8.3f ms  %6.3f warp-inst/clk/SM  (%5.1f lanes/clk/SM)\n", name, ms, winst / clk / sms, winst / clk / sms * 32);
    cudaFree(d);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; float mhz = p.clockRate / 1000.0f;
    printf("%s, %d SMs, nominal %0.f MHz (rates assume that clock)\n", p.name, sms, mhz);
    run<0>("IMAD (mad.lo.u32)", 16, sms, mhz);
    run<1>("IMAD.WIDE.U32 (+64-bit acc)", 16, sms, mhz);
    run<8>("mul.hi.u32", 16, sms, mhz);
    run<2>("IADD3 (two adds fused)", 16, sms, mhz);
    run<3>("LOP3", 16, sms, mhz);
    run<4>("SHF", 16, sms, mhz);
    run<5>("LEA (shl+add)", 16, sms, mhz);
    run<10>("PRMT", 16, sms, mhz);
    run<6>("add.u64 (2 instr)", 32, sms, mhz);
    run<9>("FFMA", 16, sms, mhz);
    run<7>("IMAD.WIDE + LOP3 pairs", 32, sms, mhz);
    run<11>("IMAD.WIDE + 2 LOP3", 48, sms, mhz);
    return 0;
}
