// Development aid: what the PCIe link gives the host-resident march (e2d_march_host) at 8192^2 — 2.15 GB each way per
// step — for several ways of issuing the copies, with no kernel in between:
//   mono       one cudaMemcpyAsync per direction and step (bench.py's pcie ceiling)
//   2d         per chunk of rows one cudaMemcpy2DAsync over the four variable planes, directions independent
//   1d         per chunk and plane one cudaMemcpyAsync, directions independent
//   2d-dep     as 2d, with the march's event dependencies (H2D of step s+1 chunk k after D2H of step s chunk k;
//              D2H of chunk k after H2D of chunk k+1)
// build: nvcc -O2 -o tools/microbench/pcie_duplex tools/microbench/pcie_duplex.cu ; run: pcie_duplex [chunk_rows] [steps]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                   \
  do                                                                            \
  {                                                                             \
    cudaError_t e_ = (x);                                                       \
    if (e_ != cudaSuccess)                                                      \
    {                                                                           \
      std::printf("%s: %s\n", #x, cudaGetErrorString(e_));                      \
      return 1;                                                                 \
    }                                                                           \
  } while (0)

int
main(int argc, char ** argv)
{
  const int    n = 8196, chunk = argc > 1 ? std::atoi(argv[1]) : 128, steps = argc > 2 ? std::atoi(argv[2]) : 6;
  const size_t plane = (size_t)n * n, bytes = 4 * plane * sizeof(double);
  double *     h[2], *d[2];
  for (int k = 0; k < 2; ++k)
  {
    CK(cudaMallocHost(&h[k], bytes));
    CK(cudaMalloc(&d[k], bytes));
    CK(cudaMemset(d[k], 0, bytes));
    for (size_t i = 0; i < 4 * plane; i += 512)
      h[k][i] = 1.0;
  }
  cudaStream_t si, so;
  CK(cudaStreamCreateWithFlags(&si, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&so, cudaStreamNonBlocking));
  std::vector<int> jb;
  for (int j = 2; j < n - 2; j += chunk)
    jb.push_back(j);
  jb.push_back(n - 2);
  const int                nchunk = (int)jb.size() - 1;
  std::vector<cudaEvent_t> ev(4 * nchunk);
  for (auto & e : ev)
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  auto copy2d = [&](double * dst, const double * src, int k, cudaMemcpyKind kind, cudaStream_t s) {
    const size_t o = (size_t)jb[k] * n;
    return cudaMemcpy2DAsync(dst + o, plane * 8, src + o, plane * 8, (size_t)(jb[k + 1] - jb[k]) * n * 8, 4, kind, s);
  };
  auto copy1d = [&](double * dst, const double * src, int k, cudaMemcpyKind kind, cudaStream_t s) {
    const size_t o = (size_t)jb[k] * n;
    cudaError_t  e = cudaSuccess;
    for (int v = 0; v < 4 && e == cudaSuccess; ++v)
      e = cudaMemcpyAsync(dst + o + v * plane, src + o + v * plane, (size_t)(jb[k + 1] - jb[k]) * n * 8, kind, s);
    return e;
  };
  for (int mode = 0; mode < 4; ++mode)
  {
    for (int rep = 0; rep < 2; ++rep) // rep 0 warms up
    {
      CK(cudaDeviceSynchronize());
      const auto t0 = std::chrono::steady_clock::now();
      for (int s = 0; s < steps; ++s)
      {
        const int par = s & 1;
        if (mode == 0)
        {
          CK(cudaMemcpyAsync(d[0], h[par], bytes, cudaMemcpyHostToDevice, si));
          CK(cudaMemcpyAsync(h[1 - par], d[1], bytes, cudaMemcpyDeviceToHost, so));
          continue;
        }
        for (int k = 0; k < nchunk; ++k)
        {
          if (mode == 3 && s > 0)
            CK(cudaStreamWaitEvent(si, ev[(size_t)((1 - par) * 2 + 1) * nchunk + k], 0));
          CK(mode == 2 ? copy1d(d[0], h[par], k, cudaMemcpyHostToDevice, si)
                       : copy2d(d[0], h[par], k, cudaMemcpyHostToDevice, si));
          if (mode == 3)
            CK(cudaEventRecord(ev[(size_t)(par * 2 + 0) * nchunk + k], si));
          const int m = mode == 3 ? k - 1 : k; // the march sends chunk k-1 back once chunk k has arrived
          for (int mm = m; mm >= 0 && mm <= (k == nchunk - 1 ? k : m); ++mm)
          {
            if (mode == 3)
              CK(cudaStreamWaitEvent(so, ev[(size_t)(par * 2 + 0) * nchunk + k], 0));
            CK(mode == 2 ? copy1d(h[1 - par], d[1], mm, cudaMemcpyDeviceToHost, so)
                         : copy2d(h[1 - par], d[1], mm, cudaMemcpyDeviceToHost, so));
            if (mode == 3)
              CK(cudaEventRecord(ev[(size_t)(par * 2 + 1) * nchunk + mm], so));
          }
        }
      }
      CK(cudaDeviceSynchronize());
      const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      static const char * names[4] = { "mono", "2d", "1d", "2d-dep" };
      if (rep)
        std::printf("%-7s chunk_rows %4d  %2d steps: %7.2f ms/step  %5.1f GB/s per direction\n", names[mode], chunk, steps,
                    t / steps * 1e3, bytes / (t / steps) * 1e-9);
    }
  }
  return 0;
}
