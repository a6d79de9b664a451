// d2h_ceiling.cu -- what this box's host can absorb: plain pinned cudaMemcpyAsync device-to-host of the bytes one
// bench step returns per GPU (6.468 GB = 10^6 MAPK trajectories x 101 rows x 8 species), on 1, 2, 4, 8 GPUs at once.
// One host thread per GPU, all copies start behind a barrier, time = max over GPUs (wall clock around the
// synchronised region and CUDA events per GPU).  Two shapes: the whole buffer in one call, and 2^14-set groups
// (106 MB) like the streamed host pipeline of ODESBatch_runHost issues them.
// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o d2h_ceiling d2h_ceiling.cu -lpthread
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define OK(c) do { cudaError_t e_ = (c); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #c, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Barrier {
  std::atomic<int> count{0}, gen{0}; int n;
  explicit Barrier(int n_) : n(n_) {}
  void wait() { int g = gen.load(); if (count.fetch_add(1) + 1 == n) { count.store(0); gen.fetch_add(1); } else while (gen.load() == g) std::this_thread::yield(); }
};

int main(int argc, char **argv)
{
  const size_t bytes = argc > 1 ? (size_t)atof(argv[1]) : (size_t)6468000000ull;
  const size_t group = (size_t)16384 * 101 * 8 * 8;                  // one 2^14-set group of the pipeline
  int ndev = 0; OK(cudaGetDeviceCount(&ndev));
  std::vector<void *> dbuf((size_t)ndev), hbuf((size_t)ndev);
  std::vector<cudaStream_t> st((size_t)ndev);
  std::vector<cudaEvent_t> e0((size_t)ndev), e1((size_t)ndev);
  for (int g = 0; g < ndev; g++) {
    OK(cudaSetDevice(g)); OK(cudaMalloc(&dbuf[g], bytes)); OK(cudaMemset(dbuf[g], 1, bytes));
    OK(cudaHostAlloc(&hbuf[g], bytes, cudaHostAllocDefault));
    OK(cudaStreamCreateWithFlags(&st[g], cudaStreamNonBlocking)); OK(cudaEventCreate(&e0[g])); OK(cudaEventCreate(&e1[g]));
    OK(cudaDeviceSynchronize());
  }
  printf("{\"bytes_per_gpu\": %zu, \"devices\": %d, \"runs\": [\n", bytes, ndev);
  bool first = true;
  for (int G = 1; G <= ndev; G *= 2) {
    for (int dir = 0; dir < 2; dir++) {                              // 0: device to host, 1: host to device
      for (int shape = 0; shape < 2; shape++) {
        double best = 1e30; float bestEv = 0;
        for (int rep = 0; rep < 3; rep++) {
          Barrier bar(G + 1);
          std::vector<float> ms((size_t)G, 0.f);
          std::vector<std::thread> th;
          for (int g = 0; g < G; g++) th.emplace_back([&, g]() {
            OK(cudaSetDevice(g));
            bar.wait();
            OK(cudaEventRecord(e0[g], st[g]));
            const size_t step = shape ? group : bytes;
            for (size_t off = 0; off < bytes; off += step) {
              const size_t n = bytes - off < step ? bytes - off : step;
              if (dir == 0) OK(cudaMemcpyAsync((char *)hbuf[g] + off, (char *)dbuf[g] + off, n, cudaMemcpyDeviceToHost, st[g]));
              else OK(cudaMemcpyAsync((char *)dbuf[g] + off, (char *)hbuf[g] + off, n, cudaMemcpyHostToDevice, st[g]));
            }
            OK(cudaEventRecord(e1[g], st[g]));
            OK(cudaStreamSynchronize(st[g]));
            OK(cudaEventElapsedTime(&ms[g], e0[g], e1[g]));
            bar.wait();
          });
          bar.wait();
          const auto t0 = std::chrono::steady_clock::now();
          bar.wait();
          const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
          for (auto &t : th) t.join();
          float mx = 0; for (float m : ms) mx = m > mx ? m : mx;
          if (dt < best) { best = dt; bestEv = mx; }
        }
        printf("%s  {\"gpus\": %d, \"direction\": \"%s\", \"shape\": \"%s\", \"wall_s\": %.4f, \"max_event_ms\": %.1f, \"aggregate_gbs\": %.1f, \"per_gpu_gbs\": %.1f}",
               first ? "" : ",\n", G, dir ? "h2d" : "d2h", shape ? "106MB groups" : "one call", best, bestEv, G * bytes / best / 1e9, bytes / best / 1e9);
        first = false;
        fflush(stdout);
      }
    }
  }
  printf("\n],\n");
  // the reference-shaped entry receives the caller's PAGEABLE arrays: plain copy into malloc'd memory, and the price of
  // pinning that memory in place (cudaHostRegister) before copying at full speed
  {
    OK(cudaSetDevice(0));
    char *pg = (char *)malloc(bytes);
    for (size_t i = 0; i < bytes; i += 4096) pg[i] = 1;                // touched: the pages exist
    auto t0 = std::chrono::steady_clock::now();
    OK(cudaMemcpy(pg, dbuf[0], bytes, cudaMemcpyDeviceToHost));
    const double tp = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    t0 = std::chrono::steady_clock::now();
    OK(cudaHostRegister(pg, bytes, cudaHostRegisterDefault));
    const double tr = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    t0 = std::chrono::steady_clock::now();
    OK(cudaMemcpyAsync(pg, dbuf[0], bytes, cudaMemcpyDeviceToHost, st[0])); OK(cudaStreamSynchronize(st[0]));
    const double tc = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    t0 = std::chrono::steady_clock::now();
    OK(cudaHostUnregister(pg));
    const double tu = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // staging: pinned ring of 106 MB blocks, host threads memcpy the blocks out into the pageable destination
    t0 = std::chrono::steady_clock::now();
    {
      const int nthreads = 8;
      std::vector<std::thread> th;
      std::atomic<size_t> nextBlock{0};
      const size_t nblocks = (bytes + group - 1) / group;
      std::vector<cudaEvent_t> ev(nblocks);
      for (size_t k = 0; k < nblocks; k++) { OK(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming)); }
      for (size_t k = 0; k < nblocks; k++) {                           // the pinned buffer of GPU 0 serves as the (oversized) ring
        const size_t off = k * group, n = bytes - off < group ? bytes - off : group;
        OK(cudaMemcpyAsync((char *)hbuf[0] + off, (char *)dbuf[0] + off, n, cudaMemcpyDeviceToHost, st[0]));
        OK(cudaEventRecord(ev[k], st[0]));
      }
      for (int t = 0; t < nthreads; t++) th.emplace_back([&]() {
        for (;;) {
          const size_t k = nextBlock.fetch_add(1); if (k >= nblocks) break;
          while (cudaEventQuery(ev[k]) == cudaErrorNotReady) std::this_thread::yield();
          const size_t off = k * group, n = bytes - off < group ? bytes - off : group;
          memcpy(pg + off, (char *)hbuf[0] + off, n);
        }
      });
      for (auto &t : th) t.join();
      for (size_t k = 0; k < nblocks; k++) cudaEventDestroy(ev[k]);
    }
    const double ts = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("\"pageable_one_gpu\": {\"plain_memcpy_gbs\": %.2f, \"host_register_s\": %.3f, \"host_register_gbs\": %.2f, \"copy_after_register_gbs\": %.2f, \"host_unregister_s\": %.3f, "
           "\"register_copy_unregister_gbs\": %.2f, \"pinned_staging_8_threads_gbs\": %.2f}}\n",
           bytes / tp / 1e9, tr, bytes / tr / 1e9, bytes / tc / 1e9, tu, bytes / (tr + tc + tu) / 1e9, bytes / ts / 1e9);
    free(pg);
  }
  return 0;
}
