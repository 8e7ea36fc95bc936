#include <cstdint>
typedef uint64_t gl;
#define EPS 0xFFFFFFFFULL
__device__ __forceinline__ gl red(gl lo, gl hi) {
    gl t0, t2, m;
    const gl hh = hi >> 32, hl = hi & EPS;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u64 %1, 0, 0;" : "=l"(t0), "=l"(m) : "l"(lo), "l"(hh));
    t0 -= (m & EPS);
    const gl t1 = (hl << 32) - hl;
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u64 %1, 0, 0;" : "=l"(t2), "=l"(m) : "l"(t0), "l"(t1));
    return t2 + ((0 - m) & EPS);
}
// reduce with 32-bit word arithmetic: x = lo + c2*2^32 - (c2 + c3) mod p ; 
__device__ __forceinline__ gl red2(gl lo, gl hi) {
    const uint32_t c2 = (uint32_t)hi, c3 = (uint32_t)(hi >> 32);
    // t = lo - c3 (borrow -> -eps), then + c2*eps
    gl t0, m, t2;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u64 %1, 0, 0;" : "=l"(t0), "=l"(m) : "l"(lo), "l"((gl)c3));
    t0 -= (m & EPS);
    // c2 * eps + t0 as one wide multiply-add with carry
    const gl t1 = (gl)c2 * 0xFFFFFFFFu;
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u64 %1, 0, 0;" : "=l"(t2), "=l"(m) : "l"(t0), "l"(t1));
    return t2 + ((0 - m) & EPS);
}
__device__ __forceinline__ gl mul_i128(gl a, gl b) { unsigned __int128 m = (unsigned __int128)a * b; return red((gl)m, (gl)(m >> 64)); }
__device__ __forceinline__ gl mul_i128r2(gl a, gl b) { unsigned __int128 m = (unsigned __int128)a * b; return red2((gl)m, (gl)(m >> 64)); }
__device__ __forceinline__ gl mul_ptx(gl a, gl b) { gl lo, hi; asm("mul.lo.u64 %0, %2, %3;\n\tmul.hi.u64 %1, %2, %3;" : "=l"(lo), "=l"(hi) : "l"(a), "l"(b)); return red(lo, hi); }
// 32-bit carry-chain schoolbook
__device__ __forceinline__ gl mul_cc(gl a, gl b) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32), c0, c1, c2, c3;
    asm("{\n\t"
        "mul.lo.u32 %0, %4, %6;\n\t"
        "mul.hi.u32 %1, %4, %6;\n\t"
        "mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
        "madc.hi.u32 %2, %4, %7, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
        "addc.u32 %3, 0, 0;\n\t"
        "mad.lo.cc.u32 %2, %5, %7, %2;\n\t"
        "madc.hi.u32 %3, %5, %7, %3;\n\t"
        "}" : "=&r"(c0), "=&r"(c1), "=&r"(c2), "=&r"(c3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return red(((gl)c1 << 32) | c0, ((gl)c3 << 32) | c2);
}
extern "C" __global__ void k_i128(gl* p) { gl x = p[threadIdx.x], y = p[threadIdx.x + 32]; p[threadIdx.x] = mul_i128(x, y); }
extern "C" __global__ void k_i128r2(gl* p) { gl x = p[threadIdx.x], y = p[threadIdx.x + 32]; p[threadIdx.x] = mul_i128r2(x, y); }
extern "C" __global__ void k_ptx(gl* p) { gl x = p[threadIdx.x], y = p[threadIdx.x + 32]; p[threadIdx.x] = mul_ptx(x, y); }
extern "C" __global__ void k_cc(gl* p) { gl x = p[threadIdx.x], y = p[threadIdx.x + 32]; p[threadIdx.x] = mul_cc(x, y); }
extern "C" __global__ void k_sq128(gl* p) { gl x = p[threadIdx.x]; p[threadIdx.x] = mul_i128(x, x); }
