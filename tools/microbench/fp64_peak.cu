// Microbenchmark: what the FP64 pipe of one B200 SM sub-partition sustains, alone and mixed with ALU work.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, int ALU_PER_FP64, int KIND>
__global__ void __launch_bounds__(256) k(double * out, int iters, double seed)
{
  double x[CHAINS];
  int    y[CHAINS];
  for (int c = 0; c < CHAINS; ++c)
  {
    x[c] = seed + threadIdx.x * 1e-3 + c;
    y[c] = threadIdx.x + c;
  }
  const double a = 1.0000001, b = 1e-9;
  for (int it = 0; it < iters; ++it)
  {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
    {
      if (KIND == 0)
        x[c] = __fma_rn(x[c], a, b);
      else if (KIND == 1)
        x[c] = __dmul_rn(x[c], a);
      else if (KIND == 2)
        x[c] = __dadd_rn(x[c], b);
      else
        x[c] = (x[c] > a) ? __dadd_rn(x[c], b) : __dmul_rn(x[c], a); // DSETP + select + ops
#pragma unroll
      for (int k = 0; k < ALU_PER_FP64; ++k)
        y[c] = (y[c] ^ (y[c] << 1)) + k; // LOP3/SHF/IADD-class work
      // ALU_PER_FP64 < 0: exactly -ALU_PER_FP64 single ALU instructions (LOP3 with two live operands) per FP64 one
#pragma unroll
      for (int k = 0; k < -ALU_PER_FP64; ++k)
        y[c] = (y[c] ^ it) & (y[(c + 1) % CHAINS] | k);
    }
  }
  double s = 0;
  int    t = 0;
  for (int c = 0; c < CHAINS; ++c)
  {
    s += x[c];
    t += y[c];
  }
  if (s == 12345.678 || t == 123456789)
    out[0] = s + t;
}

template <int CHAINS, int ALU, int KIND>
void
run(const char * name, int warps_per_smsp)
{
  double * d;
  cudaMalloc(&d, 8);
  const int   iters = 32768;
  const int   threads = 128; // one warp per SM sub-partition per block
  const int   blocks = 148 * warps_per_smsp;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<CHAINS, ALU, KIND><<<blocks, threads>>>(d, 16, 1.0);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<CHAINS, ALU, KIND><<<blocks, threads>>>(d, iters, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  if (cudaGetLastError() != cudaSuccess)
    printf("LAUNCH FAILED\n");
  const double fp64_warp_instr = (double)blocks * (threads / 32) * iters * CHAINS;
  const double per_smsp_per_s = fp64_warp_instr / (148.0 * 4) / (ms * 1e-3);
  printf("%-34s warps/SMSP %2d chains %2d alu/fp64 %d : %.3f ms, %.3f G FP64 warp-instr/s/SMSP (= %.3f per clk @1.965 GHz), "
         "%.1f TFLOP/s-equiv(FMA)\n",
         name, warps_per_smsp, CHAINS, ALU, ms, per_smsp_per_s * 1e-9, per_smsp_per_s / 1.965e9,
         fp64_warp_instr * 32 * 2 / (ms * 1e-3) * 1e-12);
  cudaFree(d);
}

int
main()
{
  run<8, 0, 0>("DFMA only", 8);
  run<8, 0, 0>("DFMA only", 4);
  run<8, 0, 0>("DFMA only", 2);
  run<8, 0, 0>("DFMA only", 1);
  run<2, 0, 0>("DFMA only", 4);
  run<1, 0, 0>("DFMA dependent chain", 1);
  run<1, 0, 0>("DFMA dependent chain", 4);
  run<8, 0, 1>("DMUL only", 4);
  run<8, 0, 2>("DADD only", 4);
  run<8, 1, 0>("DFMA + 3 ALU each", 4);
  run<8, 2, 0>("DFMA + 6 ALU each", 4);
  run<8, 0, 3>("DSETP+select+DADD+DMUL", 4);
  run<4, 1, 0>("DFMA + 3 ALU each", 3);
  // does an FP64 instruction leave its second dispatch cycle to another pipe?  N DFMA + N (2N) single ALU instructions
  run<8, -1, 0>("DFMA + 1 LOP3 each", 4);
  run<8, -2, 0>("DFMA + 2 LOP3 each", 4);
  run<8, -3, 0>("DFMA + 3 LOP3 each", 4);
  run<8, -1, 0>("DFMA + 1 LOP3 each", 8);
  return 0;
}
