// Second microbenchmark behind the backward design: is the fp32 reduction rate (tools/gather_microbench.cu: ~6 SM cycles per 128-byte
// row at 148 SMs) set by the SM side or by L2?  (a) the same RED.v4 kernel on fewer SMs; (b) the same rows reduced by the TMA
// engine (cp.reduce.async.bulk.global.shared::cta.add.f32, 128 B per operation, issued by every thread from a shared-memory row).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

__global__ void __launch_bounds__(256) red_v4_kernel(float* __restrict__ table, uint32_t row_mask, int iters) {
  const int lane = threadIdx.x & 31;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t seed = wid * 2654435761u + 12345u;
  for (int it = 0; it < iters; ++it) {
    const uint32_t r = lcg(seed);
    const uint32_t row = (r * 4u + (lane >> 3) * 2654435761u) >> 4 & row_mask;
    float* p = table + static_cast<size_t>(row) * 32 + (lane & 7) * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(1.0f) : "memory");
  }
}

// every thread owns one 128-byte row in shared memory and bulk-reduces it into random table rows; ROWS_PER_OP rows per operation
template <int ROWS_PER_OP>
__global__ void __launch_bounds__(256) red_bulk_kernel(float* __restrict__ table, uint32_t row_mask, int iters) {
  __shared__ __align__(128) float rows[256][32];
  for (int i = threadIdx.x; i < 256 * 32; i += 256) (&rows[0][0])[i] = 1.0f;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  const uint32_t src = static_cast<uint32_t>(__cvta_generic_to_shared(&rows[threadIdx.x & ~(ROWS_PER_OP - 1)][0]));
  for (int it = 0; it < iters; ++it) {
    const uint32_t row = (lcg(seed) & row_mask) & ~static_cast<uint32_t>(ROWS_PER_OP - 1);
    float* dst = table + static_cast<size_t>(row) * 32;
    if ((threadIdx.x & (ROWS_PER_OP - 1)) == 0)
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(128 * ROWS_PER_OP) : "memory");
    if ((it & 7) == 7) {
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float time_us(F launch) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < 5; ++i) {
    CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best * 1e3f;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clock_khz = 0; CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  const double ghz = clock_khz * 1e-6;
  const uint32_t rows = 1u << 17;                              // 16 MB table (L2 resident), like grad_value of an encoder call (21 MB)
  float* table; CK(cudaMalloc(&table, static_cast<size_t>(rows) * 128)); CK(cudaMemset(table, 0, static_cast<size_t>(rows) * 128));
  for (int nsm : {sms, sms / 2, sms / 4}) {
    for (int occ : {1, 2, 5}) {
      const int grid = nsm * occ, iters = 512;
      const float us = time_us([&] { red_v4_kernel<<<grid, 256>>>(table, rows - 1, iters); });
      const double n_rows = static_cast<double>(grid) * 8 * iters * 4;
      printf("RED.v4   ctas=%4d (%3d SMs x %d)  %8.1f us  %6.2f Grows/s  %5.2f TB/s  %5.2f cyc/row per busy SM\n", grid, nsm, occ, us,
             n_rows / us * 1e-3, n_rows * 128 / us * 1e-6, us * 1e-6 * ghz * 1e9 * (grid < sms ? grid : sms) / n_rows);
    }
  }
  for (int occ : {1, 2, 4}) {
    const int grid = sms * occ, iters = 256;
    float us = time_us([&] { red_bulk_kernel<1><<<grid, 256>>>(table, rows - 1, iters); });
    double n_rows = static_cast<double>(grid) * 256 * iters;
    printf("bulk 128B ctas=%4d  %8.1f us  %6.2f Grows/s  %5.2f TB/s  %5.2f cyc/row/SM\n", grid, us, n_rows / us * 1e-3, n_rows * 128 / us * 1e-6,
           us * 1e-6 * ghz * 1e9 * sms / n_rows);
    us = time_us([&] { red_bulk_kernel<4><<<grid, 256>>>(table, rows - 1, iters); });
    printf("bulk 512B ctas=%4d  %8.1f us  %6.2f Grows/s  %5.2f TB/s  %5.2f cyc/row/SM\n", grid, us, n_rows / us * 1e-3, n_rows * 128 / us * 1e-6,
           us * 1e-6 * ghz * 1e9 * sms / n_rows);
  }
  return 0;
}
