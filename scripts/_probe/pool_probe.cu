// cost of growing the stream-ordered memory pool: one big block, or many smaller ones (first use in a fresh process)
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <vector>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv)
{
  const size_t total = (size_t)atol(argv[1]) << 20, piece = (size_t)atol(argv[2]) << 20;
  const int pre = argc > 3 ? atoi(argv[3]) : 0;
  cudaFree(0);
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaMemPool_t pool; cudaDeviceGetDefaultMemPool(&pool, 0);
  unsigned long long keep = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  double t0 = now();
  if (pre) { void *p; cudaMallocAsync(&p, total, s); cudaFreeAsync(p, s); cudaStreamSynchronize(s); }
  double t1 = now();
  std::vector<void *> v;
  for (size_t b = 0; b < total; b += piece) { void *p; cudaMallocAsync(&p, piece, s); v.push_back(p); }
  cudaStreamSynchronize(s);
  double t2 = now();
  for (void *p : v) cudaMemsetAsync(p, 0, piece, s);
  cudaStreamSynchronize(s);
  double t3 = now();
  void *q; cudaMalloc(&q, total); double t4 = now();
  printf("total=%zu MB piece=%zu MB pre=%d: reserve=%.3f allocs=%.3f memset=%.3f plain_cudaMalloc_same_total=%.3f\n", total >> 20, piece >> 20, pre, t1 - t0, t2 - t1, t3 - t2, t4 - t3);
  return 0;
}
