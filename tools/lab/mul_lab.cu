#include <cstdint>
typedef uint64_t gl;
#define EPS 0xFFFFFFFFULL
__device__ __forceinline__ gl red_a(gl lo, gl hi) {
    gl t0, t2, m;
    const gl hh = hi >> 32, hl = hi & EPS;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u64 %1, 0, 0;" : "=l"(t0), "=l"(m) : "l"(lo), "l"(hh));
    t0 -= (m & EPS);
    const gl t1 = (hl << 32) - hl;
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u64 %1, 0, 0;" : "=l"(t2), "=l"(m) : "l"(t0), "l"(t1));
    return t2 + ((0 - m) & EPS);
}
__device__ __forceinline__ gl mul_a(gl a, gl b) { return red_a(a * b, __umul64hi(a, b)); }

// variant b: explicit 32-bit schoolbook + word-level reduce
__device__ __forceinline__ gl red_b(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    // x = lo + c2*2^32 - (c2 + c3)  (mod p), lo = c0 + c1*2^32
    uint32_t r0, r1, br, cy;
    // A = lo - c3 ; with borrow -> subtract eps (i.e. add 2^64 - eps ... ) 
    asm("{\n\t.reg .u32 z;\n\t"
        "sub.cc.u32 %0, %4, %6;\n\t"
        "subc.cc.u32 %1, %5, 0;\n\t"
        "subc.u32 %2, 0, 0;\n\t"        // br = 0 or 0xffffffff
        "}" : "=r"(r0), "=r"(r1), "=r"(br), "=r"(cy) : "r"(c0), "r"(c1), "r"(c3));
    // borrow: value + 2^64 was computed, need -eps: subtract (br & eps) -> r0 -= br&1? eps = 2^32-1: subtracting eps = subtract 2^32 then add 1
    // t = A - (br ? eps : 0)
    asm("sub.cc.u32 %0, %0, %2;\n\tsubc.u32 %1, %1, 0;" : "+r"(r0), "+r"(r1) : "r"(br));   // br = 0xffffffff = eps when borrow
    // B = t + c2*eps = t + (c2<<32) - c2
    uint32_t s0, s1, m;
    asm("sub.cc.u32 %0, %3, %5;\n\t"
        "subc.cc.u32 %1, %4, 0;\n\t"
        "subc.u32 %2, 0, 0;" : "=r"(s0), "=r"(s1), "=r"(m) : "r"(r0), "r"(r1), "r"(c2));
    // if borrow (m): we wrapped +2^64; but we then add c2<<32. combine: s1 += c2 with carry; net carry = carry - borrow
    uint32_t cc;
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, 0, 0;" : "+r"(s1), "=r"(cc) : "r"(c2));
    // net = cc - (m&1): in {-1,0,1}; -1 impossible? (t - c2 + c2<<32 >= 0 always since c2<<32 >= c2). so net in {0,1}: add eps if net==1
    uint32_t net = cc + m;  // m is 0 or -1
    uint32_t e = 0 - net;   // 0 or 0xffffffff
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(s0), "+r"(s1) : "r"(e));
    return ((gl)s1 << 32) | s0;
}
__device__ __forceinline__ gl mul_b(gl a, gl b) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    uint64_t p0 = (uint64_t)a0 * b0;
    uint64_t t = (uint64_t)a0 * b1 + (p0 >> 32);
    uint64_t u = (uint64_t)a1 * b0 + (uint32_t)t;
    uint64_t hi = (uint64_t)a1 * b1 + (t >> 32) + (u >> 32);
    return red_b((uint32_t)p0, (uint32_t)u, (uint32_t)hi, (uint32_t)(hi >> 32));
}
// variant c: schoolbook + 64-bit reduce a
__device__ __forceinline__ gl mul_c(gl a, gl b) {
    uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    uint64_t p0 = (uint64_t)a0 * b0;
    uint64_t t = (uint64_t)a0 * b1 + (p0 >> 32);
    uint64_t u = (uint64_t)a1 * b0 + (uint32_t)t;
    uint64_t hi = (uint64_t)a1 * b1 + (t >> 32) + (u >> 32);
    return red_a((u << 32) | (uint32_t)p0, hi);
}
#ifndef V
#define V a
#endif
#define CAT(x,y) x##y
#define XCAT(x,y) CAT(x,y)
extern "C" __global__ void k(gl* p) {
    gl x = p[threadIdx.x], y = p[threadIdx.x + 32];
    p[threadIdx.x] = XCAT(mul_, V)(x, y);
}
extern "C" __global__ void ksq(gl* p) {
    gl x = p[threadIdx.x];
    p[threadIdx.x] = XCAT(mul_, V)(x, x);
}
