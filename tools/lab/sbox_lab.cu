#include "../../tendermintx_b200/csrc/poseidon.cuh"
using namespace tmx;
extern "C" __global__ void k_mul(gl* p) { gl x = p[threadIdx.x], y = p[threadIdx.x + 32]; p[threadIdx.x] = gl_mul_nc(x, y); }
extern "C" __global__ void k_sqr(gl* p) { gl x = p[threadIdx.x]; p[threadIdx.x] = gl_sqr_nc(x); }
extern "C" __global__ void k_sbox(gl* p) { gl x = p[threadIdx.x]; p[threadIdx.x] = poseidon_sbox_nc(x); }
extern "C" __global__ void k_mds(gl* p) { gl s[12]; for (int i = 0; i < 12; i++) s[i] = p[threadIdx.x + 32 * i]; poseidon_mds_rc_alu(s, 12); for (int i = 0; i < 12; i++) p[threadIdx.x + 32 * i] = s[i]; }
extern "C" __global__ void k_conv(uint32_t* p) { uint32_t s[12], o[12]; for (int i = 0; i < 12; i++) s[i] = p[threadIdx.x + 32 * i]; mds_conv12_pieces(s, o); for (int i = 0; i < 12; i++) p[threadIdx.x + 32 * i] = o[i]; }
