// MUFU throughput on one SM sub-partition (B200): cycles per warp-instruction for ex2.f32 / ex2.f16x2 / tanh.f32 / tanh.f16x2
// with 1, 2 and 4 warps per sub-partition, 8 independent chains per thread.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/mufu tools/mufu_bench.cu && /tmp/mufu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t v) {
  uint32_t r;
  if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(r) : "r"(v));
  else if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(v));
  else if (OP == 2) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(r) : "r"(v));
  else if (OP == 3) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(v));
  else if (OP == 4) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=r"(r) : "r"(v));
  else if (OP == 6) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(v)), "f"(__uint_as_float(v ^ 0x1234u)));
  else if (OP == 7) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(v)), "f"(__uint_as_float(v ^ 0x1234u)));
  else if (OP == 8) asm volatile("fma.rn.f16x2 %0, %1, %1, %2;" : "=r"(r) : "r"(v), "r"(0x3C003C00u));
  else if (OP == 9) asm volatile("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(0x54005400u ^ v));
  else if (OP == 10) asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(r) : "r"(v), "r"(v >> 3));
  else if (OP == 11) asm volatile("fma.rn.f32 %0, %1, %1, %2;" : "=r"(r) : "r"(v), "r"(0x3F800000u));
  else if (OP == 12) asm volatile("add.f16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(0x3C003C00u));
  else if (OP == 13) asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(v >> 1));
  else if (OP == 14) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=r"(r) : "r"(v));
  else if (OP == 16) asm volatile("max.f32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(v >> 1));
  else if (OP == 17) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(v), "r"(v >> 1), "r"(0x55u));
  else if (OP == 18) asm volatile("shf.l.wrap.b32 %0, %1, %1, 7;" : "=r"(r) : "r"(v));
  else if (OP == 19) asm volatile("mad.lo.u32 %0, %1, %1, %2;" : "=r"(r) : "r"(v), "r"(77u));
  else if (OP == 20) {   // 1 MUFU + 6 dependent-free FFMA
    uint32_t e, a = v, b = v ^ 1u, c = v ^ 2u;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(e) : "r"(v));
    asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+r"(a) : "r"(0x3F800000u));
    asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+r"(b) : "r"(0x3F800000u));
    asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+r"(c) : "r"(0x3F800000u));
    asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+r"(a) : "r"(0x3F800000u));
    asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+r"(b) : "r"(0x3F800000u));
    asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+r"(c) : "r"(0x3F800000u));
    r = e ^ a ^ b ^ c;   // + 3 LOP3
  } else if (OP == 21) { // 1 MUFU + 3 HFMA2 (6 cycles of the half pipe)
    uint32_t e, a = v, b = v ^ 1u, c = v ^ 2u;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(e) : "r"(v));
    asm volatile("fma.rn.f16x2 %0, %0, %0, %1;" : "+r"(a) : "r"(0x3C003C00u));
    asm volatile("fma.rn.f16x2 %0, %0, %0, %1;" : "+r"(b) : "r"(0x3C003C00u));
    asm volatile("fma.rn.f16x2 %0, %0, %0, %1;" : "+r"(c) : "r"(0x3C003C00u));
    r = e ^ a ^ b ^ c;
  } else if (OP == 22) { // 1 MUFU + 2 PRMT (8 cycles of that pipe)
    uint32_t e, a, b;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(e) : "r"(v));
    asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(a) : "r"(v), "r"(v >> 3));
    asm volatile("prmt.b32 %0, %1, %2, 0x1054;" : "=r"(b) : "r"(a), "r"(v >> 5));
    r = e ^ b;
  } else if (OP == 23) { // 2 PRMT + 4 HFMA2 (8 + 8): do the ALU-ish and the half pipe overlap?
    uint32_t a, b, c = v, d = v ^ 3u;
    asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(a) : "r"(v), "r"(v >> 3));
    asm volatile("prmt.b32 %0, %1, %2, 0x1054;" : "=r"(b) : "r"(a), "r"(v >> 5));
    asm volatile("fma.rn.f16x2 %0, %0, %0, %1;" : "+r"(c) : "r"(0x3C003C00u));
    asm volatile("fma.rn.f16x2 %0, %0, %0, %1;" : "+r"(d) : "r"(0x3C003C00u));
    asm volatile("fma.rn.f16x2 %0, %0, %0, %1;" : "+r"(c) : "r"(0x3C003C00u));
    asm volatile("fma.rn.f16x2 %0, %0, %0, %1;" : "+r"(d) : "r"(0x3C003C00u));
    r = b ^ c ^ d;
  } else if (OP == 24) { // 4 FFMA + 4 HFMA2 : fp32 and half pipes
    uint32_t a = v, b = v ^ 1u, c = v ^ 2u, d = v ^ 3u;
    asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+r"(a) : "r"(0x3F800000u));
    asm volatile("fma.rn.f16x2 %0, %0, %0, %1;" : "+r"(c) : "r"(0x3C003C00u));
    asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+r"(b) : "r"(0x3F800000u));
    asm volatile("fma.rn.f16x2 %0, %0, %0, %1;" : "+r"(d) : "r"(0x3C003C00u));
    asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+r"(a) : "r"(0x3F800000u));
    asm volatile("fma.rn.f16x2 %0, %0, %0, %1;" : "+r"(c) : "r"(0x3C003C00u));
    asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+r"(b) : "r"(0x3F800000u));
    asm volatile("fma.rn.f16x2 %0, %0, %0, %1;" : "+r"(d) : "r"(0x3C003C00u));
    r = a ^ b ^ c ^ d;
  }
  else if (OP == 30) {   // the MLP kernel's GELU stage on one pair: F2FP, x*x, fma, mul, 2 MUFU + PRMT, fma
    uint32_t x, u, p, in, t;
    asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(x) : "f"(__uint_as_float(v)), "f"(__uint_as_float(v ^ 0x1234u)));
    asm volatile("mul.f16x2 %0, %1, %1;" : "=r"(u) : "r"(x));
    asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(u), "r"(0x28712871u), "r"(0x3A673A67u));
    asm volatile("mul.f16x2 %0, %1, %2;" : "=r"(in) : "r"(x), "r"(p));
    asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(in));
    asm volatile("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(r) : "r"(x), "r"(t));
  }
  else if (OP == 15) asm volatile("cvt.rn.f16.f32 %0, %1;" : "=h"(*reinterpret_cast<unsigned short*>(&r)) : "f"(__uint_as_float(v)));
  else {  // OP 5: half2 polynomial exp2 (x <= 0): clamp, round via magic add, cubic, scale by exponent bits
    uint32_t x, t, n, f, p, sc;
    asm volatile("max.f16x2 %0, %1, %2;" : "=r"(x) : "r"(v), "r"(0xCB80CB80u));          // >= -15
    asm volatile("add.f16x2 %0, %1, %2;" : "=r"(t) : "r"(x), "r"(0x660F660Fu));          // + 1551
    asm volatile("add.f16x2 %0, %1, %2;" : "=r"(n) : "r"(t), "r"(0xE60FE60Fu));          // - 1551
    asm volatile("sub.f16x2 %0, %1, %2;" : "=r"(f) : "r"(x), "r"(n));
    asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(f), "r"(0x2B1B2B1Bu), "r"(0x33B033B0u));
    asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(p), "r"(f), "r"(0x398C398Cu));
    asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(p), "r"(f), "r"(0x3C003C00u));
    sc = (t << 10) & 0x7C007C00u;
    asm volatile("mul.f16x2 %0, %1, %2;" : "=r"(r) : "r"(p), "r"(sc));
  }
  return r;
}
template <int OP>
__global__ void bench(uint32_t* out, long long* cyc, int iters) {
  uint32_t a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0x3c003c00u + threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = op<OP>(a[i]);
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name, int mufu_per_op) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int wps = 1; wps <= 4; wps *= 2) {
    const int iters = 2000;
    bench<OP><<<148, 128 * wps>>>(out, cyc, iters);
    bench<OP><<<148, 128 * wps>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 8.0 * wps);
    printf("%-14s warps/SMSP=%d: %.2f cycles per warp-op on its sub-partition (%.2f per MUFU instruction; %.1f results/clk/SM)\n", name, wps, per,
           per / mufu_per_op, 4 * 32.0 * (OP == 1 || OP == 3 || OP == 5 ? 2 : 1) / per);
  }
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("ex2.f32", 1); run<1>("ex2.f16x2", 2); run<2>("tanh.f32", 1); run<3>("tanh.f16x2", 2); run<5>("poly ex2 h2", 1);
  run<6>("cvt f16x2.f32", 1); run<7>("cvt bf16x2.f32", 1); run<8>("fma.f16x2", 1); run<9>("min.f16x2", 1); run<10>("prmt", 1);
  run<11>("fma.f32", 1); run<12>("add.f16x2", 1); run<13>("max.f16x2", 1); run<14>("cvt f32.f16", 1); run<15>("cvt f16.f32", 1);
  run<16>("max.f32", 1); run<17>("lop3", 1); run<18>("shf", 1); run<19>("imad", 1);
  run<30>("gelu pair", 2);
  run<20>("mufu+6ffma+3lop", 1); run<21>("mufu+3hfma2+3lop", 1); run<22>("mufu+2prmt+lop", 1); run<23>("2prmt+4hfma2+2lop", 1); run<24>("4ffma+4hfma2+3lop", 1);
  return 0;
}
