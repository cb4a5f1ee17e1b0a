// Probe: how does a short H2D (or D2H) transfer fare while another stream keeps a long one going in the same direction?
// Variants: depth of the long transfer's pieces in flight, priority of the short transfer's stream, short transfer done by
// a kernel reading page-locked host memory directly.  Prints a table.  nvcc -O2 -arch=sm_100a -o link_interleave ...
#include <cuda_runtime.h>
#include <stdio.h>
#include <chrono>
#include <thread>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__global__ void k_pull(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main() {
  const size_t BIG = (size_t)1 << 30, SMALL = (size_t)128 << 20;
  char *h_big, *h_small, *d_big, *d_small;
  CK(cudaMallocHost(&h_big, BIG)); CK(cudaMallocHost(&h_small, SMALL));
  CK(cudaMalloc(&d_big, BIG)); CK(cudaMalloc(&d_small, SMALL));
  memset(h_big, 1, BIG); memset(h_small, 2, SMALL);
  int lo, hi; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  cudaStream_t sa, sb, sb_hi; CK(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking)); CK(cudaStreamCreateWithPriority(&sb_hi, cudaStreamNonBlocking, hi));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  printf("asyncEngineCount %d, priorities %d..%d\n", pr.asyncEngineCount, lo, hi);
  cudaEvent_t ev[8]; for (auto &evi : ev) CK(cudaEventCreateWithFlags(&evi, cudaEventDisableTiming));
  for (int dir = 0; dir < 2; dir++) {       // 0 = H2D, 1 = D2H
    for (int variant = 0; variant < 6; variant++) {
      // variant: 0 big whole; 1 pieces 8 MB depth 3; 2 pieces 8 MB depth 1; 3 = 1 + small on the high-priority stream; 4 = 1 + small by kernel; 5 pieces 2 MB depth 2
      const size_t piece = variant == 0 ? BIG : variant == 5 ? (size_t)2 << 20 : (size_t)8 << 20;
      const int depth = variant == 2 ? 1 : variant == 5 ? 2 : 3;
      CK(cudaDeviceSynchronize());
      double t_small_start = 0, t_small_end = 0, t_big_end = 0;
      const double t0 = now_ms();
      std::thread small([&] {
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
        t_small_start = now_ms();
        cudaStream_t s = variant == 3 ? sb_hi : sb;
        if (variant == 4) {
          if (dir == 0) k_pull<<<296, 512, 0, s>>>((const uint4 *)h_small, (uint4 *)d_small, SMALL / 16);
          else k_pull<<<296, 512, 0, s>>>((const uint4 *)d_small, (uint4 *)h_small, SMALL / 16);
        } else if (dir == 0) cudaMemcpyAsync(d_small, h_small, SMALL, cudaMemcpyHostToDevice, s);
        else cudaMemcpyAsync(h_small, d_small, SMALL, cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        t_small_end = now_ms();
      });
      size_t q = 0;
      for (size_t off = 0; off < BIG; off += piece, q++) {
        const size_t len = BIG - off < piece ? BIG - off : piece;
        const int s = (int)(q % depth);
        if (q >= (size_t)depth) CK(cudaEventSynchronize(ev[s]));
        if (dir == 0) CK(cudaMemcpyAsync(d_big + off, h_big + off, len, cudaMemcpyHostToDevice, sa));
        else CK(cudaMemcpyAsync(h_big + off, d_big + off, len, cudaMemcpyDeviceToHost, sa));
        CK(cudaEventRecord(ev[s], sa));
      }
      CK(cudaStreamSynchronize(sa));
      t_big_end = now_ms();
      small.join();
      printf("%s variant %d: big done at %.1f ms; small (128 MiB) submitted at %.1f, done at %.1f (took %.1f ms)\n", dir ? "D2H" : "H2D", variant,
             t_big_end - t0, t_small_start - t0, t_small_end - t0, t_small_end - t_small_start);
    }
  }
  return 0;
}
