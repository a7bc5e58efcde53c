// timing of the CUDA start-up costs a short-lived host program pays (context, first launch, first pool growth)
#include <cstdio>
#include <chrono>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__global__ void k(int *p) { if (p) *p = 1; }
int main()
{
  double t0 = now();
  cudaFree(0);
  double t1 = now();
  int *d; cudaMalloc(&d, 4); k<<<1, 1>>>(d); cudaDeviceSynchronize();
  double t2 = now();
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  void *p[8];
  for (int i = 0; i < 8; i++) cudaMallocAsync(&p[i], 512ull << 20, s);
  cudaStreamSynchronize(s);
  double t3 = now();
  for (int i = 0; i < 8; i++) cudaMemsetAsync(p[i], 0, 512ull << 20, s);
  cudaStreamSynchronize(s);
  double t4 = now();
  void *h; cudaHostAlloc(&h, 400ull << 20, cudaHostAllocDefault);
  double t5 = now();
  printf("cuda_init=%.3f first_launch=%.3f mallocasync_4GB=%.3f memset_4GB=%.3f hostalloc_400MB=%.3f\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4);
  return 0;
}
