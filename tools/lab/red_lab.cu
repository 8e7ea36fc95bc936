#include <cstdint>
typedef uint64_t gl;
#define EPS 0xFFFFFFFFULL
__device__ __forceinline__ gl red_cur(gl lo, gl hi) {
    gl t0, t2, m, c;
    const gl hh = hi >> 32, hl = hi & EPS;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u64 %1, 0, 0;" : "=l"(t0), "=l"(m) : "l"(lo), "l"(hh));
    t0 -= (m & EPS);
    const gl t1 = (hl << 32) - hl;
    asm("add.cc.u64 %0, %2, %3;\n\taddc.u64 %1, 0, 0;" : "=l"(t2), "=l"(c) : "l"(t0), "l"(t1));
    return t2 + ((0 - c) & EPS);
}
// hl * eps + t0 through one wide multiply-add with carry out
__device__ __forceinline__ gl red_mad(gl lo, gl hi) {
    gl t0, t2, m, c;
    const gl hh = hi >> 32;
    const uint32_t hl = (uint32_t)hi;
    asm("sub.cc.u64 %0, %2, %3;\n\tsubc.u64 %1, 0, 0;" : "=l"(t0), "=l"(m) : "l"(lo), "l"(hh));
    t0 -= (m & EPS);
    asm("{\n\t.reg .u64 a;\n\tcvt.u64.u32 a, %3;\n\tmad.lo.cc.u64 %0, a, 0xFFFFFFFF, %2;\n\taddc.u64 %1, 0, 0;\n\t}" : "=l"(t2), "=l"(c) : "l"(t0), "r"(hl));
    return t2 + ((0 - c) & EPS);
}
// all three terms in one go: lo - hh + hl*eps, corrections folded: result = lo + hl*eps - hh, with carry c1 and borrow b: + (c1 - b) * eps
__device__ __forceinline__ gl red_c(gl lo, gl hi) {
    const uint32_t hl = (uint32_t)hi, hh = (uint32_t)(hi >> 32);
    const gl t1 = (gl)hl * 0xFFFFFFFFu + lo;           // may wrap
    const gl c1 = t1 < lo;
    gl t2 = t1 + (c1 ? EPS : 0);                        // no second wrap
    const gl b = t2 < hh;
    t2 -= hh;
    return t2 - (b ? EPS : 0);
}
__device__ __forceinline__ gl canon(gl a) { gl t, c; asm("add.cc.u64 %0, %2, %3;\n\taddc.u64 %1, 0, 0;" : "=l"(t), "=l"(c) : "l"(a), "l"((gl)EPS)); return c ? t : a; }
#define K(name, RED) extern "C" __global__ void name(gl* p) { gl x = p[threadIdx.x], y = p[threadIdx.x + 32], lo, hi; \
    asm("mul.lo.u64 %0, %2, %3;\n\tmul.hi.u64 %1, %2, %3;" : "=l"(lo), "=l"(hi) : "l"(x), "l"(y)); p[threadIdx.x] = canon(RED(lo, hi)); }
K(k_cur, red_cur) K(k_mad, red_mad) K(k_c, red_c)
