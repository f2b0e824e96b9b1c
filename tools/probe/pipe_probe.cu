// pipe_probe.cu -- issue / pipe throughput of the instruction mixes of the LANE solver on sm_100a (tuning aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/pipe_probe tools/probe/pipe_probe.cu
// Every test is a loop of independent dependency chains in inline PTX; reported: cycles per warp-instruction per
// SM sub-partition at 1..4 warps per sub-partition (one CTA per SM).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2000

template <int MIX>
__global__ void probe(float *out, float seed, unsigned long long *cycles) {
    float a0 = seed + threadIdx.x, a1 = a0 * 1.1f, a2 = a0 * 1.2f, a3 = a0 * 1.3f, a4 = a0 * 1.4f, a5 = a0 * 1.5f, a6 = a0 * 1.6f, a7 = a0 * 1.7f;
    unsigned long long p0, p1, p2, p3, p4, p5, p6, p7, w;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
    p4 = p0; p5 = p1; p6 = p2; p7 = p3;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(w) : "f"(0.999f));
    double d0 = a0, d1 = a1, d2 = a2, d3 = a3, d4 = a4, d5 = a5, e0 = a6, e1 = a7, e2 = a3;
    unsigned u0 = threadIdx.x, u1 = u0 * 3, u2 = u0 * 5, u3 = u0 * 7;
    __syncthreads();
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MIX == 0) {  // 8 x mul.f32x2
            asm volatile("mul.rn.f32x2 %0, %0, %8; mul.rn.f32x2 %1, %1, %8; mul.rn.f32x2 %2, %2, %8; mul.rn.f32x2 %3, %3, %8;"
                         "mul.rn.f32x2 %4, %4, %8; mul.rn.f32x2 %5, %5, %8; mul.rn.f32x2 %6, %6, %8; mul.rn.f32x2 %7, %7, %8;"
                         : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+l"(p4), "+l"(p5), "+l"(p6), "+l"(p7) : "l"(w));
        } else if (MIX == 1) {  // 8 x fma.f32x2
            asm volatile("fma.rn.f32x2 %0, %0, %8, %1; fma.rn.f32x2 %1, %1, %8, %2; fma.rn.f32x2 %2, %2, %8, %3; fma.rn.f32x2 %3, %3, %8, %4;"
                         "fma.rn.f32x2 %4, %4, %8, %5; fma.rn.f32x2 %5, %5, %8, %6; fma.rn.f32x2 %6, %6, %8, %7; fma.rn.f32x2 %7, %7, %8, %0;"
                         : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+l"(p4), "+l"(p5), "+l"(p6), "+l"(p7) : "l"(w));
        } else if (MIX == 2) {  // 8 x add.f32 (scalar)
            asm volatile("add.rn.f32 %0, %0, %8; add.rn.f32 %1, %1, %8; add.rn.f32 %2, %2, %8; add.rn.f32 %3, %3, %8;"
                         "add.rn.f32 %4, %4, %8; add.rn.f32 %5, %5, %8; add.rn.f32 %6, %6, %8; add.rn.f32 %7, %7, %8;"
                         : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(seed));
        } else if (MIX == 3) {  // 6 x fma.f64, independent
            asm volatile("fma.rn.f64 %0, %0, %6, %0; fma.rn.f64 %1, %1, %6, %1; fma.rn.f64 %2, %2, %6, %2;"
                         "fma.rn.f64 %3, %3, %6, %3; fma.rn.f64 %4, %4, %6, %4; fma.rn.f64 %5, %5, %6, %5;"
                         : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5) : "d"(0.999));
        } else if (MIX == 4) {  // 6 x cvt.f64.f32 (results folded with xor so that they are not dead)
            asm volatile("cvt.f64.f32 %0, %6; cvt.f64.f32 %1, %7; cvt.f64.f32 %2, %8; cvt.f64.f32 %3, %9; cvt.f64.f32 %4, %10; cvt.f64.f32 %5, %11;"
                         : "=d"(d0), "=d"(d1), "=d"(d2), "=d"(d3), "=d"(d4), "=d"(d5) : "f"(a0), "f"(a1), "f"(a2), "f"(a3), "f"(a4), "f"(a5));
            a0 += (float)it;  // 1 I2F + 1 FADD per 6: keeps the inputs changing
        } else if (MIX == 5) {  // 8 x prmt
            asm volatile("prmt.b32 %0, %0, %4, 0x7440; prmt.b32 %1, %1, %4, 0x7441; prmt.b32 %2, %2, %4, 0x7442; prmt.b32 %3, %3, %4, 0x7443;"
                         "prmt.b32 %0, %0, %4, 0x7441; prmt.b32 %1, %1, %4, 0x7442; prmt.b32 %2, %2, %4, 0x7443; prmt.b32 %3, %3, %4, 0x7440;"
                         : "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3) : "r"(0x4B000000u));
        } else if (MIX == 6) {  // the per-pixel block: 3 add.f32, 3 cvt.f64.f32, 6 fma.f64
            asm volatile("add.rn.f32 %6, %6, %9; add.rn.f32 %7, %7, %9; add.rn.f32 %8, %8, %9;"
                         "cvt.f64.f32 %10, %6; cvt.f64.f32 %11, %7; cvt.f64.f32 %12, %8;"
                         "fma.rn.f64 %0, %10, %11, %0; fma.rn.f64 %1, %10, %12, %1; fma.rn.f64 %2, %10, %10, %2;"
                         "fma.rn.f64 %3, %11, %11, %3; fma.rn.f64 %4, %11, %12, %4; fma.rn.f64 %5, %12, %12, %5;"
                         : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5), "+f"(a0), "+f"(a1), "+f"(a2), "+f"(seed), "=d"(e0), "=d"(e1), "=d"(e2));
        } else if (MIX == 7) {  // the sample block: 8 mul.f32x2 + 3 fma.f32x2
            asm volatile("mul.rn.f32x2 %0, %0, %8; mul.rn.f32x2 %1, %1, %8; mul.rn.f32x2 %2, %2, %8; mul.rn.f32x2 %3, %3, %8;"
                         "mul.rn.f32x2 %0, %0, %4; mul.rn.f32x2 %1, %1, %5; mul.rn.f32x2 %2, %2, %6; mul.rn.f32x2 %3, %3, %7;"
                         "fma.rn.f32x2 %0, %1, %8, %0; fma.rn.f32x2 %0, %2, %8, %0; fma.rn.f32x2 %0, %3, %8, %0;"
                         : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+l"(p4), "+l"(p5), "+l"(p6), "+l"(p7) : "l"(w));
        } else if (MIX == 8) {  // per-pixel block + sample block (the row loop's ratio is about 49 : 45)
            asm volatile("add.rn.f32 %6, %6, %9; add.rn.f32 %7, %7, %9; add.rn.f32 %8, %8, %9;"
                         "cvt.f64.f32 %10, %6; cvt.f64.f32 %11, %7; cvt.f64.f32 %12, %8;"
                         "fma.rn.f64 %0, %10, %11, %0; fma.rn.f64 %1, %10, %12, %1; fma.rn.f64 %2, %10, %10, %2;"
                         "fma.rn.f64 %3, %11, %11, %3; fma.rn.f64 %4, %11, %12, %4; fma.rn.f64 %5, %12, %12, %5;"
                         : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5), "+f"(a0), "+f"(a1), "+f"(a2), "+f"(seed), "=d"(e0), "=d"(e1), "=d"(e2));
            asm volatile("mul.rn.f32x2 %0, %0, %8; mul.rn.f32x2 %1, %1, %8; mul.rn.f32x2 %2, %2, %8; mul.rn.f32x2 %3, %3, %8;"
                         "mul.rn.f32x2 %0, %0, %4; mul.rn.f32x2 %1, %1, %5; mul.rn.f32x2 %2, %2, %6; mul.rn.f32x2 %3, %3, %7;"
                         "fma.rn.f32x2 %0, %1, %8, %0; fma.rn.f32x2 %0, %2, %8, %0; fma.rn.f32x2 %0, %3, %8, %0;"
                         : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+l"(p4), "+l"(p5), "+l"(p6), "+l"(p7) : "l"(w));
        } else if (MIX == 9) {  // 8 x mov (register rotation)
            asm volatile("mov.b32 %0, %1; mov.b32 %1, %2; mov.b32 %2, %3; mov.b32 %3, %4; mov.b32 %4, %5; mov.b32 %5, %6; mov.b32 %6, %7; mov.b32 %7, %0;"
                         : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7));
        } else if (MIX == 10) {  // 6 x cvt.f64.f32 + 12 independent fma.f64 (do conversions hide behind the fp64 pipe?)
            asm volatile("cvt.f64.f32 %6, %9; cvt.f64.f32 %7, %10; cvt.f64.f32 %8, %11;"
                         "fma.rn.f64 %0, %6, %7, %0; fma.rn.f64 %1, %6, %8, %1; fma.rn.f64 %2, %6, %6, %2;"
                         "fma.rn.f64 %3, %7, %7, %3; fma.rn.f64 %4, %7, %8, %4; fma.rn.f64 %5, %8, %8, %5;"
                         : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5), "=d"(e0), "=d"(e1), "=d"(e2) : "f"(a0), "f"(a1), "f"(a2));
        } else if (MIX == 12) {  // 4 x mul.f32 + 4 x add.f32, independent (un-fused fp32 peak, scalar)
            asm volatile("mul.rn.f32 %0, %0, %8; add.rn.f32 %1, %1, %8; mul.rn.f32 %2, %2, %8; add.rn.f32 %3, %3, %8;"
                         "mul.rn.f32 %4, %4, %8; add.rn.f32 %5, %5, %8; mul.rn.f32 %6, %6, %8; add.rn.f32 %7, %7, %8;"
                         : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(seed));
        } else if (MIX == 13) {  // 4 x mul.f32x2 + 4 x add.f32x2, independent (un-fused fp32 peak, packed)
            asm volatile("mul.rn.f32x2 %0, %0, %8; add.rn.f32x2 %1, %1, %8; mul.rn.f32x2 %2, %2, %8; add.rn.f32x2 %3, %3, %8;"
                         "mul.rn.f32x2 %4, %4, %8; add.rn.f32x2 %5, %5, %8; mul.rn.f32x2 %6, %6, %8; add.rn.f32x2 %7, %7, %8;"
                         : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+l"(p4), "+l"(p5), "+l"(p6), "+l"(p7) : "l"(w));
        } else if (MIX == 14) {  // the per-pixel block WITHOUT its conversions: 3 add.f32 + 6 fma.f64 (what does an F2F cost?)
            asm volatile("add.rn.f32 %6, %6, %9; add.rn.f32 %7, %7, %9; add.rn.f32 %8, %8, %9;"
                         "fma.rn.f64 %0, %10, %11, %0; fma.rn.f64 %1, %10, %12, %1; fma.rn.f64 %2, %10, %10, %2;"
                         "fma.rn.f64 %3, %11, %11, %3; fma.rn.f64 %4, %11, %12, %4; fma.rn.f64 %5, %12, %12, %5;"
                         : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5), "+f"(a0), "+f"(a1), "+f"(a2) : "f"(seed), "d"(e0), "d"(e1), "d"(e2));
        } else if (MIX == 15) {  // 3 add.f32 + 3 cvt.f64.f32 of their results (inputs change every trip), results xor-folded
            asm volatile("add.rn.f32 %3, %3, %6; add.rn.f32 %4, %4, %6; add.rn.f32 %5, %5, %6;"
                         "cvt.f64.f32 %0, %3; cvt.f64.f32 %1, %4; cvt.f64.f32 %2, %5;"
                         : "=d"(e0), "=d"(e1), "=d"(e2), "+f"(a0), "+f"(a1), "+f"(a2) : "f"(seed));
            asm volatile("xor.b64 %0, %0, %1; xor.b64 %0, %0, %2; xor.b64 %0, %0, %3;" : "+l"(p7) : "l"(__double_as_longlong(e0)), "l"(__double_as_longlong(e1)), "l"(__double_as_longlong(e2)));
        } else if (MIX == 16) {  // two pipes: 4 mul.f32x2 (FMA pipe) + 4 prmt (ALU pipe), independent: do they overlap?
            asm volatile("mul.rn.f32x2 %0, %0, %8; prmt.b32 %4, %4, %9, 0x7440; mul.rn.f32x2 %1, %1, %8; prmt.b32 %5, %5, %9, 0x7441;"
                         "mul.rn.f32x2 %2, %2, %8; prmt.b32 %6, %6, %9, 0x7442; mul.rn.f32x2 %3, %3, %8; prmt.b32 %7, %7, %9, 0x7443;"
                         : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+r"(u0), "+r"(u1), "+r"(u2), "+r"(u3) : "l"(w), "r"(0x4B000000u));
        } else if (MIX == 17) {  // two pipes: 4 fma.f64 (fp64 pipe) + 4 mul.f32x2 (FMA pipe), independent
            asm volatile("fma.rn.f64 %0, %0, %8, %0; mul.rn.f32x2 %4, %4, %9; fma.rn.f64 %1, %1, %8, %1; mul.rn.f32x2 %5, %5, %9;"
                         "fma.rn.f64 %2, %2, %8, %2; mul.rn.f32x2 %6, %6, %9; fma.rn.f64 %3, %3, %8, %3; mul.rn.f32x2 %7, %7, %9;"
                         : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3) : "d"(0.999), "l"(w));
        } else if (MIX == 18) {  // three pipes: 3 fma.f64 + 3 mul.f32x2 + 3 prmt, independent
            asm volatile("fma.rn.f64 %0, %0, %9, %0; mul.rn.f32x2 %3, %3, %10; prmt.b32 %6, %6, %11, 0x7440;"
                         "fma.rn.f64 %1, %1, %9, %1; mul.rn.f32x2 %4, %4, %10; prmt.b32 %7, %7, %11, 0x7441;"
                         "fma.rn.f64 %2, %2, %9, %2; mul.rn.f32x2 %5, %5, %10; prmt.b32 %8, %8, %11, 0x7442;"
                         : "+d"(d0), "+d"(d1), "+d"(d2), "+l"(p0), "+l"(p1), "+l"(p2), "+r"(u0), "+r"(u1), "+r"(u2) : "d"(0.999), "l"(w), "r"(0x4B000000u));
        } else if (MIX == 19) {  // 8 x cvt.rn.f32.u8 of a byte of a changing word (I2F.U8 Rx, Ry.Bk on the XU pipe), xor-folded
            float c0, c1, c2, c3, c4, c5, c6, c7;
            c0 = __uint2float_rn(u0 & 0xffu); c1 = __uint2float_rn((u0 >> 8) & 0xffu); c2 = __uint2float_rn((u0 >> 16) & 0xffu); c3 = __uint2float_rn(u0 >> 24);
            c4 = __uint2float_rn(u1 & 0xffu); c5 = __uint2float_rn((u1 >> 8) & 0xffu); c6 = __uint2float_rn((u1 >> 16) & 0xffu); c7 = __uint2float_rn(u1 >> 24);
            u2 ^= __float_as_uint(c0) ^ __float_as_uint(c1) ^ __float_as_uint(c2) ^ __float_as_uint(c3);
            u3 ^= __float_as_uint(c4) ^ __float_as_uint(c5) ^ __float_as_uint(c6) ^ __float_as_uint(c7);
            u0 += 0x01010101u; u1 += 0x03050709u;
        } else if (MIX == 20) {  // the same 8 conversions + 8 mul.f32x2: do the conversions hide behind the FMA pipe?
            float c0, c1, c2, c3, c4, c5, c6, c7;
            c0 = __uint2float_rn(u0 & 0xffu); c1 = __uint2float_rn((u0 >> 8) & 0xffu); c2 = __uint2float_rn((u0 >> 16) & 0xffu); c3 = __uint2float_rn(u0 >> 24);
            c4 = __uint2float_rn(u1 & 0xffu); c5 = __uint2float_rn((u1 >> 8) & 0xffu); c6 = __uint2float_rn((u1 >> 16) & 0xffu); c7 = __uint2float_rn(u1 >> 24);
            u2 ^= __float_as_uint(c0) ^ __float_as_uint(c1) ^ __float_as_uint(c2) ^ __float_as_uint(c3);
            u3 ^= __float_as_uint(c4) ^ __float_as_uint(c5) ^ __float_as_uint(c6) ^ __float_as_uint(c7);
            u0 += 0x01010101u; u1 += 0x03050709u;
            asm volatile("mul.rn.f32x2 %0, %0, %8; mul.rn.f32x2 %1, %1, %8; mul.rn.f32x2 %2, %2, %8; mul.rn.f32x2 %3, %3, %8;"
                         "mul.rn.f32x2 %4, %4, %8; mul.rn.f32x2 %5, %5, %8; mul.rn.f32x2 %6, %6, %8; mul.rn.f32x2 %7, %7, %8;"
                         : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3), "+l"(p4), "+l"(p5), "+l"(p6), "+l"(p7) : "l"(w));
        } else if (MIX == 21) {  // what the solver does today for 8 pixels: 8 prmt + 4 add.f32x2 (+ the same 8 mul.f32x2)
            unsigned q0, q1, q2, q3, q4, q5, q6, q7;
            asm volatile("prmt.b32 %0, %8, %10, 0x7440; prmt.b32 %1, %8, %10, 0x7441; prmt.b32 %2, %8, %10, 0x7442; prmt.b32 %3, %8, %10, 0x7443;"
                         "prmt.b32 %4, %9, %10, 0x7440; prmt.b32 %5, %9, %10, 0x7441; prmt.b32 %6, %9, %10, 0x7442; prmt.b32 %7, %9, %10, 0x7443;"
                         : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3), "=r"(q4), "=r"(q5), "=r"(q6), "=r"(q7) : "r"(u0), "r"(u1), "r"(0x4B000000u));
            unsigned long long g0, g1, g2, g3, m;
            asm volatile("mov.b64 %0, {%1, %2};" : "=l"(g0) : "r"(q0), "r"(q1));
            asm volatile("mov.b64 %0, {%1, %2};" : "=l"(g1) : "r"(q2), "r"(q3));
            asm volatile("mov.b64 %0, {%1, %2};" : "=l"(g2) : "r"(q4), "r"(q5));
            asm volatile("mov.b64 %0, {%1, %2};" : "=l"(g3) : "r"(q6), "r"(q7));
            asm volatile("mov.b64 %0, {%1, %1};" : "=l"(m) : "f"(-8388608.0f));
            asm volatile("add.rn.f32x2 %0, %0, %4; add.rn.f32x2 %1, %1, %4; add.rn.f32x2 %2, %2, %4; add.rn.f32x2 %3, %3, %4;"
                         : "+l"(g0), "+l"(g1), "+l"(g2), "+l"(g3) : "l"(m));
            p4 ^= g0 ^ g1; p5 ^= g2 ^ g3;
            u0 += 0x01010101u; u1 += 0x03050709u;
            asm volatile("mul.rn.f32x2 %0, %0, %4; mul.rn.f32x2 %1, %1, %4; mul.rn.f32x2 %2, %2, %4; mul.rn.f32x2 %3, %3, %4;"
                         : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3) : "l"(w));
            asm volatile("mul.rn.f32x2 %0, %0, %4; mul.rn.f32x2 %1, %1, %4; mul.rn.f32x2 %2, %2, %4; mul.rn.f32x2 %3, %3, %4;"
                         : "+l"(p6), "+l"(p7), "+l"(p0), "+l"(p1) : "l"(w));
        } else if (MIX == 11) {  // 8 x scalar mul.f32 with three distinct registers
            asm volatile("mul.rn.f32 %0, %1, %8; mul.rn.f32 %1, %2, %8; mul.rn.f32 %2, %3, %8; mul.rn.f32 %3, %4, %8;"
                         "mul.rn.f32 %4, %5, %8; mul.rn.f32 %5, %6, %8; mul.rn.f32 %6, %7, %8; mul.rn.f32 %7, %0, %8;"
                         : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(seed));
        }
    }
    unsigned long long t1 = clock64();
    float2 q;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(q.x), "=f"(q.y) : "l"(p0 ^ p1 ^ p2 ^ p3 ^ p4 ^ p5 ^ p6 ^ p7));
    out[blockIdx.x * blockDim.x + threadIdx.x] = q.x + q.y + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (float)(d0 + d1 + d2 + d3 + d4 + d5) + (float)(u0 ^ u1 ^ u2 ^ u3);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

static int g_sms = 148;
static double g_best[24];  // best (lowest) cycles per warp-instruction per SMSP of each mix

template <int MIX>
void run(const char *name, int insts_per_iter) {
    float *out;
    unsigned long long *cyc, h;
    cudaMalloc(&out, (size_t)g_sms * 2 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 8);
    printf("%-58s", name);
    g_best[MIX] = 1e9;
    const int wps_list[5] = {1, 2, 4, 6, 8};   // warps per sub-partition, one CTA per SM (its own clock64 brackets all of them)
    for (int k = 0; k < 5; ++k) {
        const int wps = wps_list[k];
        const int ctas = 1, threads = 128 * wps;
        probe<MIX><<<g_sms * ctas, threads>>>(out, 1.0f, cyc);
        probe<MIX><<<g_sms * ctas, threads>>>(out, 1.0f, cyc);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        // block 0's cycles for its own warps; with two CTAs per SM the sub-partition holds twice as many warps
        const double c = (double)h / ((double)ITERS * insts_per_iter * wps);
        if (c < g_best[MIX]) g_best[MIX] = c;
        printf("  w%d: %5.2f", wps, c);
    }
    printf("   (cycles per warp-instruction per SMSP)\n");
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    g_sms = prop.multiProcessorCount;
    run<0>("8 x mul.f32x2 (FMUL2)", 8);
    run<1>("8 x fma.f32x2 (FFMA2)", 8);
    run<2>("8 x add.f32 (FADD)", 8);
    run<11>("8 x mul.f32 three distinct registers (FMUL)", 8);
    run<3>("6 x fma.f64 (DFMA)", 6);
    run<4>("6 x cvt.f64.f32 (F2F) [+1 I2F +1 FADD]", 6);
    run<5>("8 x prmt (PRMT)", 8);
    run<9>("8 x mov (MOV / IMAD.MOV)", 8);
    run<10>("3 cvt.f64.f32 + 6 fma.f64", 9);
    run<6>("pixel block: 3 FADD + 3 F2F + 6 DFMA", 12);
    run<7>("sample block: 8 FMUL2 + 3 FFMA2", 11);
    run<8>("pixel block + sample block", 23);
    run<14>("pixel block without conversions: 3 FADD + 6 DFMA", 9);
    run<15>("3 FADD + 3 F2F (changing inputs) [+3 xor.b64]", 6);
    run<16>("two pipes: 4 FMUL2 + 4 PRMT", 8);
    run<17>("two pipes: 4 DFMA + 4 FMUL2", 8);
    run<18>("three pipes: 3 DFMA + 3 FMUL2 + 3 PRMT", 9);
    run<19>("8 x cvt.f32.u8 of a byte (I2F.U8) [+8 xor +2 add]", 8);
    run<20>("8 I2F.U8 [+10 int] + 8 FMUL2", 16);
    run<21>("today: 8 PRMT + 4 FADD2 [+6 xor/add] + 8 FMUL2", 20);
    run<12>("4 x mul.f32 + 4 x add.f32 (scalar, un-fused)", 8);
    run<13>("4 x mul.f32x2 + 4 x add.f32x2 (packed, un-fused)", 8);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    // Un-fused fp32 peak = the best sustained rate of individually rounded mul / add lane-operations, in flop per
    // cycle per SM: 4 sub-partitions x 32 lanes x (1 or 2 operations per instruction) / cycles per instruction.
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double scalar = 4 * 32 * 1.0 / g_best[12], packed = 4 * 32 * 2.0 / g_best[13];
    const double per_sm = scalar > packed ? scalar : packed;
    printf("{\"fp32_unfused_flop_per_cycle_per_sm\": %.2f, \"scalar\": %.2f, \"packed\": %.2f, \"sms\": %d, "
           "\"clock_mhz\": %.0f, \"fp32_unfused_tflops\": %.3f, \"how\": \"tools/probe/pipe_probe.cu: independent "
           "mul.rn/add.rn chains (scalar and f32x2), best over 1..8 warps per sub-partition, clock64 cycles of one "
           "CTA\"}\n",
           per_sm, scalar, packed, g_sms, clk_khz / 1e3, per_sm * g_sms * clk_khz * 1e3 / 1e12);
    return e != cudaSuccess;
}
