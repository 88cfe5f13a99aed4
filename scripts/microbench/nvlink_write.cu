// nvlink_write.cu -- how fast can SM-issued stores fill a peer GPU's memory over NVLink 5?  (r02 experiment:
// the push exchange is bound by this.)  One process, two devices with peer access, both directions at once.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o nvlink_write nvlink_write.cu && ./nvlink_write
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// variant 0: 16 B per lane, consecutive lanes consecutive (512 B per warp instruction), .cs
// variant 1: same, default caching
// variant 2: 2 x 16 B per lane (lane writes 32 B contiguous: 1 KiB per warp in two instructions)
// variant 3: 128-byte runs scattered with a 4 KiB stride between the four runs of a warp (what a tile store
//            looks like when tile bits 3, 4 are not index bits 3, 4)
template <int V>
__global__ void __launch_bounds__(256) k_write(const double2 *__restrict__ src, double2 *__restrict__ dst, uint64_t n) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
    double2 v[4];
    uint64_t idx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint64_t e = i + u * stride;
      if (V == 2) e = ((e >> 1) << 1) | (e & 1);   // same mapping; the pairing is done by the unroll below
      if (V == 3) {  // swap index bits (3,4) with bits (8,9): runs of 8 amplitudes, the warp's four runs 4 KiB apart
        const uint64_t lo = (e >> 3) & 3, hi = (e >> 8) & 3;
        e = (e & ~((uint64_t(3) << 3) | (uint64_t(3) << 8))) | (hi << 3) | (lo << 8);
      }
      idx[u] = e;
      if (e < n) v[u] = __ldcs(src + e);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (idx[u] < n) {
        if (V == 1) dst[idx[u]] = v[u];
        else __stcs(dst + idx[u], v[u]);
      }
  }
}

// variant 4: TMA bulk copies shared -> peer global, 8 KiB per CTA iteration (loaded with plain loads first)
__global__ void __launch_bounds__(256) k_bulk(const double2 *__restrict__ src, double2 *__restrict__ dst, uint64_t n) {
  __shared__ __align__(128) double2 buf[2][512];
  const uint64_t chunks = n / 512;
  int slot = 0;
  for (uint64_t c = blockIdx.x; c < chunks; c += gridDim.x, slot ^= 1) {
    // the bulk store issued two iterations ago from this slot must have finished reading it
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();
    buf[slot][threadIdx.x] = __ldcs(src + c * 512 + threadIdx.x);
    buf[slot][threadIdx.x + 256] = __ldcs(src + c * 512 + 256 + threadIdx.x);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t sa = uint32_t(__cvta_generic_to_shared(&buf[slot][0]));
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * 512), "r"(sa), "n"(8192) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  const uint64_t n = uint64_t(1) << 28;   // 4 GiB per buffer
  double2 *src[2], *dst[2];
  cudaStream_t st[2];
  cudaEvent_t e0[2], e1[2];
  for (int d = 0; d < 2; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaDeviceEnablePeerAccess(1 - d, 0));
    CK(cudaMalloc(&src[d], n * 16));
    CK(cudaMalloc(&dst[d], n * 16));
    CK(cudaMemset(src[d], 1, n * 16));
    CK(cudaStreamCreate(&st[d]));
    CK(cudaEventCreate(&e0[d]));
    CK(cudaEventCreate(&e1[d]));
  }
  const char *names[] = {"16B/lane .cs", "16B/lane default", "16B/lane .cs (pairs)", "128B runs 4KiB apart .cs", "TMA bulk store 8KiB", "cudaMemcpyPeerAsync"};
  for (int both = 0; both < 2; ++both)
    for (int v = 0; v < 6; ++v) {
      for (int blocksPerSm : {4, 8, 16}) {
        if ((v == 5) && blocksPerSm != 4) continue;
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          for (int d = 0; d < (both ? 2 : 1); ++d) {
            CK(cudaSetDevice(d));
            CK(cudaEventRecord(e0[d], st[d]));
            const int blocks = 148 * blocksPerSm;
            double2 *to = dst[1 - d];
            if (v == 0) k_write<0><<<blocks, 256, 0, st[d]>>>(src[d], to, n);
            else if (v == 1) k_write<1><<<blocks, 256, 0, st[d]>>>(src[d], to, n);
            else if (v == 2) k_write<2><<<blocks, 256, 0, st[d]>>>(src[d], to, n);
            else if (v == 3) k_write<3><<<blocks, 256, 0, st[d]>>>(src[d], to, n);
            else if (v == 4) k_bulk<<<blocks, 256, 0, st[d]>>>(src[d], to, n);
            else CK(cudaMemcpyPeerAsync(to, 1 - d, src[d], d, n * 16, st[d]));
            CK(cudaEventRecord(e1[d], st[d]));
          }
          float worst = 0;
          for (int d = 0; d < (both ? 2 : 1); ++d) {
            CK(cudaSetDevice(d));
            CK(cudaEventSynchronize(e1[d]));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0[d], e1[d]));
            if (ms > worst) worst = ms;
          }
          if (worst < best) best = worst;
        }
        printf("%s  %-28s blocks/SM %2d  %.2f ms  %.0f GB/s per direction\n", both ? "both directions" : "one direction ", names[v],
               blocksPerSm, best, double(n) * 16 / (best * 1e-3) / 1e9);
      }
    }
  return 0;
}
