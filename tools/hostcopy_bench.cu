// host-side copy ceilings on the GPU box (what bounds the D2H leg of meshify()):
//   nvcc -O2 -o /tmp/hostcopy_bench tools/hostcopy_bench.cu -lpthread && /tmp/hostcopy_bench
#include <cuda_runtime.h>
#include <pthread.h>
#include <sys/mman.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct job { char *dst; const char *src; size_t n, slice; std::atomic<long long> next; int nt_stream; };
static void *work(void *a) {
  job *j = (job *)a;
  for (;;) {
    long long i = j->next.fetch_add(1);
    size_t o = (size_t)i * j->slice;
    if (o >= j->n) break;
    size_t len = j->n - o < j->slice ? j->n - o : j->slice;
    if (j->src) memcpy(j->dst + o, j->src + o, len);
    else for (size_t q = 0; q < len; q += 4096) ((volatile char *)j->dst)[o + q] = 0;
  }
  return nullptr;
}
static double par(char *dst, const char *src, size_t n, int nt, size_t slice) {
  job j; j.dst = dst; j.src = src; j.n = n; j.slice = slice; j.next = 0;
  std::vector<pthread_t> th(nt);
  double t0 = now();
  for (int i = 0; i < nt; i++) pthread_create(&th[i], nullptr, work, &j);
  for (int i = 0; i < nt; i++) pthread_join(th[i], nullptr);
  return now() - t0;
}
int main() {
  const size_t N = (size_t)2 << 30;
  printf("cores %ld\n", sysconf(_SC_NPROCESSORS_ONLN));
  char *pin; cudaMallocHost(&pin, (size_t)256 << 20);
  char *dev; cudaMalloc(&dev, N); cudaMemset(dev, 1, N);
  for (int huge = 0; huge < 4; huge++) {
    char *m = (char *)malloc(N + (2 << 20));
    char *a = (char *)(((uintptr_t)m + (2 << 20) - 1) & ~(uintptr_t)((2 << 20) - 1));
    if (huge & 1) madvise(a, N, MADV_HUGEPAGE);
    if (huge < 2) printf("huge=%d touch(16 thr) %.1f ms\n", huge, par(a, nullptr, N, 16, 1 << 20));
    else {
#ifdef MADV_POPULATE_WRITE
      double tp0 = now(); int rcp = madvise(a, N, MADV_POPULATE_WRITE); double tp1 = now();
      printf("huge=%d MADV_POPULATE_WRITE (1 thr) rc %d: %.1f ms\n", huge, rcp, tp1 - tp0);
#endif
    }
    for (int nt : {4, 8, 16, 32}) {
      double t = 0;
      for (size_t o = 0; o < N; o += (size_t)256 << 20) t += par(a + o, pin, (size_t)256 << 20, nt, 1 << 20);
      printf("  memcpy pinned->malloc %2d thr: %.1f GB/s\n", nt, N / t / 1e6);
    }
    double t0 = now();
    cudaError_t e = cudaHostRegister(a, N, cudaHostRegisterDefault);
    double t1 = now();
    printf("  cudaHostRegister 2 GiB: %.1f ms (%s)\n", t1 - t0, cudaGetErrorString(e));
    if (e == cudaSuccess) {
      cudaMemcpy(a, dev, N, cudaMemcpyDeviceToHost);
      t0 = now(); cudaMemcpy(a, dev, N, cudaMemcpyDeviceToHost); t1 = now();
      printf("  direct D2H into registered: %.1f ms = %.1f GB/s\n", t1 - t0, N / (t1 - t0) / 1e6);
      t0 = now(); cudaHostUnregister(a); t1 = now();
      printf("  cudaHostUnregister: %.1f ms\n", t1 - t0);
      // chunked register / copy / unregister pipeline cost per 64 MiB
      t0 = now();
      for (size_t o = 0; o < N; o += (size_t)64 << 20) { cudaHostRegister(a + o, (size_t)64 << 20, 0); cudaHostUnregister(a + o); }
      t1 = now();
      printf("  register+unregister in 64 MiB pieces: %.1f ms total\n", t1 - t0);
    }
    // pageable cudaMemcpy as the driver does it
    t0 = now(); cudaMemcpy(a, dev, N, cudaMemcpyDeviceToHost); t1 = now();
    printf("  plain cudaMemcpy D2H into pageable: %.1f ms = %.1f GB/s\n", t1 - t0, N / (t1 - t0) / 1e6);
    free(m);
  }
  // pinned D2H ceiling
  double t0 = now();
  for (int i = 0; i < 8; i++) cudaMemcpy(pin, dev, (size_t)256 << 20, cudaMemcpyDeviceToHost);
  double t1 = now();
  printf("D2H into pinned: %.1f GB/s\n", 8 * 256.0 * 1048576 / (t1 - t0) / 1e6);
  return 0;
}
