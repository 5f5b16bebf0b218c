// Operand-read cost of DFMA on sm_100a (dev tool, evidence for DESIGN.md): how many cycles does a
// warp-wide DFMA hold the FP64 pipe when it reads 1, 2 or 3 distinct 64-bit register operands,
// a constant-bank operand, or an immediate?  16 independent chains per thread, 4 warps / SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_fp64_operands.bin scripts/ubench_fp64_operands.cu
#include <cuda_runtime.h>
#include <cstdio>

__constant__ double CC[4] = {0.999999, 1e-7, 0.5, 0.25};

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double* out, int iters, double a, double b) {
  double x[16], y[16], z[16];
#pragma unroll
  for (int u = 0; u < 16; u++) {
    x[u] = threadIdx.x * 1e-9 + u;
    y[u] = 0.999999 + 1e-9 * (threadIdx.x + u);
    z[u] = 1e-7 * (1 + threadIdx.x + u);
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (MODE == 0) x[u] = fma(x[u], a, b);              // 1 varying reg + 2 shared regs (reuse cache)
      if (MODE == 1) x[u] = fma(x[u], y[u], z[u]);        // 3 distinct regs
      if (MODE == 2) x[u] = fma(x[u], y[u], b);           // 2 distinct + 1 shared reg
      if (MODE == 3) x[u] = fma(x[u], y[u], CC[1]);       // 2 distinct + constant bank
      if (MODE == 4) x[u] = fma(x[u], CC[0], CC[1] * 0 + 6755399441055744.0);  // 1 reg + const + imm
      if (MODE == 5) x[u] = fma(x[u], x[u], x[u]);        // 1 distinct reg in 3 slots
      if (MODE == 6) x[u] = fma(x[u], y[u], x[u]);        // 2 distinct regs, 3 slots
      if (MODE == 7) x[u] = fma(x[u], y[(u + 1) & 15], z[(u + 2) & 15]);  // 3 distinct, shuffled banks
      if (MODE == 8) x[u] = x[u] * y[u];                  // DMUL 2 regs
      if (MODE == 9) x[u] = x[u] + z[u];                  // DADD 2 regs
      if (MODE == 10) { x[u] = fma(x[u], y[u], b); z[u] = fma(z[u], CC[0], CC[1]); }   // shared-reg DFMA interleaved with unrelated DFMA
      if (MODE == 11) { x[u] = fma(x[u], y[u], b); z[u] = fma(z[u], z[u], a); }        // two shared regs alternating in slot C
      if (MODE == 12) { x[u] = fma(b, x[u], y[u]); z[u] = fma(z[u], CC[0], CC[1]); }   // shared reg in slot A, interleaved
    }
  }
  double s = 0;
#pragma unroll
  for (int u = 0; u < 16; u++) s += x[u] + y[u] + z[u];
  if (s == 12345.678) out[0] = s;
}

template <int MODE>
void run(const char* label, double* d, int sms) {
  const int threads = 512, iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<sms, threads>>>(d, 1000, 0.999999, 1e-7);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k<MODE><<<sms, threads>>>(d, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double cyc = best * 1e-3 * clk_khz * 1e3 / (iters * 4.0 * 16.0 * (MODE >= 10 ? 2.0 : 1.0));
  printf("%-44s %8.3f ms  %5.2f cycles per warp-instruction (max clock)\n", label, best, cyc);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double* d;
  cudaMalloc(&d, 64);
  int n = p.multiProcessorCount;
  run<0>("dfma x,a,b (1 varying + 2 shared regs)", d, n);
  run<1>("dfma x,y,z (3 distinct regs)", d, n);
  run<2>("dfma x,y,b (2 distinct + 1 shared reg)", d, n);
  run<3>("dfma x,y,c[] (2 distinct + const bank)", d, n);
  run<4>("dfma x,c[],imm", d, n);
  run<5>("dfma x,x,x", d, n);
  run<6>("dfma x,y,x", d, n);
  run<7>("dfma x,y',z' (3 distinct, shuffled)", d, n);
  run<8>("dmul x,y", d, n);
  run<9>("dadd x,z", d, n);
  run<10>("[dfma x,y,b ; dfma z,c,c] per instr", d, n);
  run<11>("[dfma x,y,b ; dfma z,z,a] per instr", d, n);
  run<12>("[dfma b,x,y ; dfma z,c,c] per instr", d, n);
  return 0;
}
