// alloc_probe.cu — what a cold c2b_visibility_graph call pays for memory: pinned host allocation by
// cudaHostAlloc against mmap + transparent huge pages + cudaHostRegister, and cudaMalloc.
//   nvcc -O2 -o /tmp/alloc_probe profiles/probes/alloc_probe.cu && /tmp/alloc_probe
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <chrono>
#include <cstdio>
#include <cstring>

static double now() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main() {
  cudaFree(0);
  const size_t GB = 1ull << 30;
  void *d = nullptr;
  for (size_t sz : {GB / 4, GB, 2 * GB}) {
    double t0 = now();
    cudaMalloc(&d, sz);
    double t1 = now();
    cudaFree(d);
    double t2 = now();
    printf("cudaMalloc %5.2f GB: %8.3f ms   cudaFree %8.3f ms\n", sz / 1e9, t1 - t0, t2 - t1);
  }
  for (unsigned flags : {cudaHostAllocDefault, cudaHostAllocPortable})
    for (size_t sz : {(size_t)4096, (size_t)1 << 20, (size_t)64 << 20}) {
      void *h = nullptr;
      double t0 = now();
      cudaHostAlloc(&h, sz, flags);
      double t1 = now();
      cudaFreeHost(h);
      printf("cudaHostAlloc %9zu B flags %u: %8.3f ms   cudaFreeHost %8.3f ms\n", sz, flags, t1 - t0, now() - t1);
    }
  for (int rep = 0; rep < 2; ++rep) {
    void *h = nullptr;
    double t0 = now();
    cudaHostAlloc(&h, GB, cudaHostAllocPortable);
    double t1 = now();
    cudaFreeHost(h);
    double t2 = now();
    printf("cudaHostAlloc 1 GB: %8.3f ms   cudaFreeHost %8.3f ms\n", t1 - t0, t2 - t1);
  }
  for (int huge = 0; huge < 2; ++huge)
    for (int rep = 0; rep < 2; ++rep) {
      double t0 = now();
      void *p = mmap(nullptr, GB + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
      char *a = (char *)(((uintptr_t)p + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1));
      int adv = huge ? madvise(a, GB, MADV_HUGEPAGE) : 0;
      double t1 = now();
      for (size_t o = 0; o < GB; o += 4096) a[o] = 0;  // fault the pages in
      double t2 = now();
      cudaError_t e = cudaHostRegister(a, GB, cudaHostRegisterPortable);
      double t3 = now();
      // a copy into it, to see that it behaves like pinned memory
      cudaMalloc(&d, GB);
      cudaMemset(d, 1, GB);
      cudaDeviceSynchronize();
      double t4 = now();
      cudaMemcpy(a, d, GB, cudaMemcpyDeviceToHost);
      double t5 = now();
      cudaFree(d);
      cudaHostUnregister(a);
      munmap(p, GB + (2u << 20));
      printf("mmap%s (madvise rc %d) %7.3f ms  touch %8.3f ms  cudaHostRegister %8.3f ms (%s)  D2H 1 GB %7.3f ms = %.1f GB/s\n",
             huge ? "+THP" : "    ", adv, t1 - t0, t2 - t1, t3 - t2, cudaGetErrorString(e), t5 - t4, GB / 1e6 / (t5 - t4));
    }
  FILE *f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
  if (f) {
    char buf[128] = {0};
    if (fgets(buf, sizeof buf, f)) printf("THP enabled: %s", buf);
    fclose(f);
  }
  return 0;
}
