// Which stream operations can release a kernel that spins on a device flag while it occupies every SM?
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o flag_probe flag_probe.cu -lcuda
// The spin is bounded (2 s), so the probe can never hang the GPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void spin_kernel(const unsigned* flag, const double* data, size_t n, int* result, double* sink) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned v = 0;
  while (true) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 2000000000ull) break;
    __nanosleep(256);
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    result[0] = v ? 1 : 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    result[1] = (int)((t - t0) / 1000);
    sink[0] = data[n - 1];
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

int main() {
  const size_t n = 12 << 20;  // 96 MB of doubles
  double *h, *d, *sink;
  unsigned *flag, *h_one;
  int *res, h_res[2];
  CK(cudaMallocHost(&h, n * 8));
  for (size_t i = 0; i < n; ++i) h[i] = (double)i;
  CK(cudaMallocHost(&h_one, 4));
  *h_one = 1;
  CK(cudaMalloc(&d, n * 8));
  CK(cudaMalloc(&flag, 4));
  CK(cudaMalloc(&res, 8));
  CK(cudaMalloc(&sink, 8));
  cudaStream_t copy, compute;
  CK(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking));
  int per_sm = 0, sms = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spin_kernel, 128, 0));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (int mode = 0; mode < 5; ++mode) {
    // 0: kernel first, then big copy + 4-byte flag copy      1: copies first, then kernel
    // 2: copies first, flag via cuStreamWriteValue32          3: like 1 but the grid leaves half of every SM free
    // 4: like 1, flag by a 4-byte copy from a SECOND stream after an event (flag copy alone in its stream)
    CK(cudaMemset(flag, 0, 4));
    CK(cudaMemset(res, 0xff, 8));
    CK(cudaDeviceSynchronize());
    const int grid = (mode == 3) ? sms * per_sm / 2 : sms * per_sm;
    if (mode == 0) spin_kernel<<<grid, 128, 0, compute>>>(flag, d, n, res, sink);
    CK(cudaMemcpyAsync(d, h, n * 8, cudaMemcpyHostToDevice, copy));
    if (mode == 2) {
      CUresult r = cuStreamWriteValue32((CUstream)copy, (CUdeviceptr)flag, 1u, 0);
      if (r != CUDA_SUCCESS) printf("cuStreamWriteValue32 failed %d\n", (int)r);
    } else {
      CK(cudaMemcpyAsync(flag, h_one, 4, cudaMemcpyHostToDevice, copy));
    }
    if (mode != 0) spin_kernel<<<grid, 128, 0, compute>>>(flag, d, n, res, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_res, res, 8, cudaMemcpyDeviceToHost));
    printf("mode %d grid %d (%d/SM): flag seen %d after %d us\n", mode, grid, per_sm, h_res[0], h_res[1]);
  }
  return 0;
}
