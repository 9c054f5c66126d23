// Issue-rate probe for the instruction mix of the attention softmax on sm_100a (development tool):
// how many clocks does ONE SM sub-partition need per warp-instruction of FFMA / FFMA2 / FADD2 /
// FMNMX3 / F2FP / MUFU.EX2 / IMAD, alone and mixed, with 1, 2 or 4 warps resident per sub-partition?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_probe tools/pipe_probe.cu && /tmp/pipe_probe
//
// Each kernel runs ITER iterations of 32 independent operations per thread (8 accumulators x 4), one
// CTA of 128 * W threads on one SM; clocks per warp-instruction per sub-partition = elapsed /
// (ITER * 32 * W).  Results feed DESIGN.md's account of what bounds the softmax warps.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2048;

__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm volatile("{ .reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %4}; mov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd; }" : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1) {
  asm volatile("{ .reg .b64 ra, rd;\n\t"
      "mov.b64 ra, {%2, %3}; mov.b64 rd, {%0, %1};\n\t"
      "add.rn.f32x2 rd, rd, ra;\n\t"
      "mov.b64 {%0, %1}, rd; }" : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1));
}

template <int KIND>
__global__ void probe(float* out, long long* clocks, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  unsigned int u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (KIND == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[8 + i]), "f"(seed));                 // FFMA
        if (KIND == 1) fma2(a[2 * i], a[2 * i + 1], a[2 * i], a[2 * i + 1], seed, 0.5f);                                     // FFMA2
        if (KIND == 2) add2(a[2 * i], a[2 * i + 1], seed, 0.25f);                                                            // FADD2
        if (KIND == 3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[8 + i]), "f"(seed));                     // FMNMX3
        if (KIND == 4) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[8 + i]));               // F2FP
        if (KIND == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));                                             // MUFU.EX2
        if (KIND == 6) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(u[(i + 2) & 7])); // IMAD
        if (KIND == 7) {                                                   // the MUFU-pair mix: FFMA2, EX2, EX2, FADD2, F2FP
          fma2(a[2 * i], a[2 * i + 1], a[2 * i], a[2 * i + 1], seed, 0.5f);
          if ((i & 1) == 0) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[8 + i]));
            asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[8 + i]));
          } else {
            add2(a[2 * i], a[2 * i + 1], seed, 0.25f);
          }
        }
        if (KIND == 10) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));                                           // MUFU.EX2 on a packed fp16 pair
        if (KIND == 11) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[8 + i]));           // F2FP to fp16x2
        if (KIND == 12) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));                     // HADD2
        if (KIND == 13) {                                                  // fp16 softmax mix per pair: FFMA2, F2FP.F16, EX2.F16x2, HADD2
          fma2(a[2 * i], a[2 * i + 1], a[2 * i], a[2 * i + 1], seed, 0.5f);
          asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
          asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
          asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(u[(i + 4) & 7]) : "r"(u[i]));
        }
        if (KIND == 14) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));                                      // MUFU.EX2 on a packed bf16 pair
        if (KIND == 8) asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[8 + i]), "f"(seed));             // FFMA.SAT
        if (KIND == 9) asm volatile("fma.rm.ftz.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[8 + i]), "f"(seed));             // FFMA.RM
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += __uint_as_float(u[i]);
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) clocks[0] = t1 - t0;
}

template <int KIND>
void run(const char* name, int ops_per_slot) {
  float* out;
  long long* clk;
  cudaMalloc(&out, 4096 * 4);
  cudaMalloc(&clk, 8);
  for (int W : {1, 2, 4}) {
    probe<KIND><<<1, 128 * W>>>(out, clk, 0.999f);
    probe<KIND><<<1, 128 * W>>>(out, clk, 0.999f);
    long long c = 0;
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    const double per = double(c) / (double(ITER) * 32 * ops_per_slot * W);
    printf("%-34s warps/SMSP=%d  %.2f clk per warp-instruction per sub-partition\n", name, W, per);
  }
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  run<0>("FFMA (3-reg)", 1);
  run<1>("FFMA2 (fma.rn.f32x2)", 1);
  run<2>("FADD2 (add.rn.f32x2)", 1);
  run<3>("FMNMX3 (max.f32 a,b,c)", 1);
  run<4>("F2FP (cvt.rn.bf16x2.f32)", 1);
  run<5>("MUFU.EX2", 1);
  run<6>("IMAD (mad.lo.u32)", 1);
  run<8>("FFMA.SAT", 1);
  run<9>("FFMA.RM", 1);
  run<7>("mix / slot (avg of FFMA2,EX2,EX2,F2FP | FFMA2,FADD2)", 3);
  run<10>("MUFU.EX2 f16x2 (ex2.approx.f16x2)", 1);
  run<14>("MUFU.EX2 bf16x2", 1);
  run<11>("F2FP f16x2 (cvt.rn.f16x2.f32)", 1);
  run<12>("HADD2 (add.rn.f16x2)", 1);
  run<13>("fp16 softmax mix / instr (FFMA2,F2FP,EX2.F16x2,HADD2)", 4);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
