#include "../../tendermintx_b200/csrc/gl.cuh"
using namespace tmx;
// variant A: current ops
__device__ __forceinline__ void bfly_a(gl& a, gl& b, gl w) { gl s = gl_add(a, b); gl d = gl_sub(a, b); a = s; b = gl_mul(d, w); }
// variant B: 32-bit carry words, add via (a - (p - b))
__device__ __forceinline__ gl sub_b(gl a, gl b) {
    gl d; uint32_t m;
    asm("{\n\t.reg .u32 al, ah, bl, bh, dl, dh;\n\tmov.b64 {al, ah}, %2;\n\tmov.b64 {bl, bh}, %3;\n\t"
        "sub.cc.u32 dl, al, bl;\n\tsubc.cc.u32 dh, ah, bh;\n\tsubc.u32 %1, 0, 0;\n\t"   // m = 0 or 0xffffffff (= eps when borrow)
        "sub.cc.u32 dl, dl, %1;\n\tsubc.u32 dh, dh, 0;\n\tmov.b64 %0, {dl, dh};\n\t}" : "=l"(d), "=r"(m) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ gl add_b(gl a, gl b) { return sub_b(a, GL_P - b); }   // b = 0 -> a - p + p = a
__device__ __forceinline__ gl canon_b(gl a) {   // a in [0, 2^64)
    gl t; uint32_t c;
    asm("{\n\t.reg .u32 al, ah, tl, th;\n\t.reg .pred q;\n\tmov.b64 {al, ah}, %2;\n\t"
        "add.cc.u32 tl, al, 0xffffffff;\n\taddc.cc.u32 th, ah, 0;\n\taddc.u32 %1, 0, 0;\n\t"
        "setp.ne.u32 q, %1, 0;\n\tselp.u32 tl, tl, al, q;\n\tselp.u32 th, th, ah, q;\n\tmov.b64 %0, {tl, th};\n\t}" : "=l"(t), "=r"(c) : "l"(a));
    return t;
}
__device__ __forceinline__ void bfly_b(gl& a, gl& b, gl w) { gl s = add_b(a, b); gl d = sub_b(a, b); a = s; b = canon_b(gl_mul_nc(d, w)); }
// variant C: lazy: keep b-lane (product) non-canonical, canonicalise when it is next used as a subtrahend / both-operands
__device__ __forceinline__ void bfly_c(gl& a, gl& b, gl w) { gl s = add_b(a, b); gl d = sub_b(a, b); a = s; b = gl_mul_nc(d, w); }
#define K(name, F) extern "C" __global__ void name(gl* p) { gl a = p[threadIdx.x], b = p[threadIdx.x + 32], w = p[threadIdx.x + 64]; F(a, b, w); p[threadIdx.x] = a; p[threadIdx.x + 32] = b; }
K(k_a, bfly_a) K(k_b, bfly_b) K(k_c, bfly_c)
