// Microbenchmark behind the MSDA kernel design: what does one gathered 128-byte row cost an SM, as a function of how the
// warp instruction is shaped (rows per instruction, bytes per lane) and of where the rows live (L1 / L2 / HBM)?  Same for
// the fp32 vector reductions of the backward.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_microbench
// gather_microbench.cu ; run on the GPU box.  Output: one line per (mode, table size): ns, rows/s, SM cycles per row.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// MODE 0: LDG.128, 8 lanes per row, 4 random rows per instruction
// MODE 1: LDG.32, 32 lanes per row, 1 random row per instruction
// MODE 2: LDG.64, 16 lanes per row, 2 random rows per instruction
// MODE 3: LDG.128, 4 CONTIGUOUS rows per instruction (512 B)
// MODE 4: LDG.128, 2 random pairs of adjacent rows per instruction
// MODE 5: LDG.128, 4 random rows, but only 64 B of each (4 lanes per row, 8 rows per instruction; the 2-byte value case)
template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ table, uint32_t row_mask, int iters, float* sink) {
  const int lane = threadIdx.x & 31;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t seed = wid * 2654435761u + 12345u;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint32_t r = lcg(seed);               // warp-uniform random number
      if (MODE == 0) {
        const uint32_t row = (r * 4u + (lane >> 3) * 2654435761u) >> 4 & row_mask;      // 4 different rows
        v[u] = __ldg(reinterpret_cast<const float4*>(table + static_cast<size_t>(row) * 32) + (lane & 7));
      } else if (MODE == 1) {
        const uint32_t row = r & row_mask;
        v[u].x = __ldg(table + static_cast<size_t>(row) * 32 + lane); v[u].y = v[u].z = v[u].w = 0.f;
      } else if (MODE == 2) {
        const uint32_t row = (r + (lane >> 4) * 2654435761u) >> 4 & row_mask;
        const float2 t = __ldg(reinterpret_cast<const float2*>(table + static_cast<size_t>(row) * 32) + (lane & 15));
        v[u].x = t.x; v[u].y = t.y; v[u].z = v[u].w = 0.f;
      } else if (MODE == 3) {
        const uint32_t row = ((r & row_mask) & ~3u) + (lane >> 3);
        v[u] = __ldg(reinterpret_cast<const float4*>(table + static_cast<size_t>(row) * 32) + (lane & 7));
      } else if (MODE == 4) {
        const uint32_t row = ((((r + (lane >> 4) * 2654435761u) >> 4) & row_mask) & ~1u) + ((lane >> 3) & 1);
        v[u] = __ldg(reinterpret_cast<const float4*>(table + static_cast<size_t>(row) * 32) + (lane & 7));
      } else {
        const uint32_t row = (r * 4u + (lane >> 2) * 2654435761u) >> 4 & row_mask;      // 8 different rows, first 64 B of each
        v[u] = __ldg(reinterpret_cast<const float4*>(table + static_cast<size_t>(row) * 32) + (lane & 3));
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) sink[0] = acc;
}

// MODE 0: red.v4.f32, 8 lanes per row, 4 random rows per instruction;  1: red.f32, 1 row;  2: red.v2.f32, 2 rows
template <int MODE>
__global__ void __launch_bounds__(256) red_kernel(float* __restrict__ table, uint32_t row_mask, int iters) {
  const int lane = threadIdx.x & 31;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t seed = wid * 2654435761u + 12345u;
  for (int it = 0; it < iters; ++it) {
    const uint32_t r = lcg(seed);
    if (MODE == 0) {
      const uint32_t row = (r * 4u + (lane >> 3) * 2654435761u) >> 4 & row_mask;
      float* p = table + static_cast<size_t>(row) * 32 + (lane & 7) * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(1.0f) : "memory");
    } else if (MODE == 1) {
      const uint32_t row = r & row_mask;
      float* p = table + static_cast<size_t>(row) * 32 + lane;
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
    } else {
      const uint32_t row = (r + (lane >> 4) * 2654435761u) >> 4 & row_mask;
      float* p = table + static_cast<size_t>(row) * 32 + (lane & 15) * 2;
      asm volatile("red.global.add.v2.f32 [%0], {%1, %1};" ::"l"(p), "f"(1.0f) : "memory");
    }
  }
}

template <typename F>
float time_us(F launch) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < 5; ++i) {
    CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best * 1e3f;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clock_khz = 0; CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  const double ghz = clock_khz * 1e-6;
  printf("device %s, %d SMs, %.3f GHz (nominal)\n", prop.name, sms, ghz);
  const size_t max_rows = 1u << 22;                         // 4 M rows x 128 B = 512 MB
  float* table; CK(cudaMalloc(&table, max_rows * 128));
  CK(cudaMemset(table, 0, max_rows * 128));
  float* sink; CK(cudaMalloc(&sink, 4));
  const int ctas_per_sm[] = {3, 6};
  const uint32_t sizes[] = {1u << 6, 1u << 12, 1u << 17, 1u << 22};     // 8 KB (L1), 512 KB, 16 MB (L2), 512 MB (HBM)
  const char* gnames[] = {"LDG.128 4 rows/instr", "LDG.32 1 row/instr", "LDG.64 2 rows/instr", "LDG.128 4 contiguous rows",
                          "LDG.128 2 adjacent pairs", "LDG.128 8 half rows (64 B)"};
  for (int occ : ctas_per_sm) {
    const int grid = sms * occ;
    for (uint32_t rows : sizes) {
      for (int mode = 0; mode < 6; ++mode) {
        const int rows_per_instr[] = {4, 1, 2, 4, 4, 8};
        const int iters = (rows >= (1u << 22)) ? 64 : 256;
        constexpr int U = 8;
        float us = 0;
        switch (mode) {
          case 0: us = time_us([&] { gather_kernel<0, U><<<grid, 256>>>(table, rows - 1, iters, sink); }); break;
          case 1: us = time_us([&] { gather_kernel<1, U><<<grid, 256>>>(table, rows - 1, iters, sink); }); break;
          case 2: us = time_us([&] { gather_kernel<2, U><<<grid, 256>>>(table, rows - 1, iters, sink); }); break;
          case 3: us = time_us([&] { gather_kernel<3, U><<<grid, 256>>>(table, rows - 1, iters, sink); }); break;
          case 4: us = time_us([&] { gather_kernel<4, U><<<grid, 256>>>(table, rows - 1, iters, sink); }); break;
          case 5: us = time_us([&] { gather_kernel<5, U><<<grid, 256>>>(table, rows - 1, iters, sink); }); break;
        }
        const double n_instr = static_cast<double>(grid) * 8 * iters * U;          // warp instructions
        const double n_rows = n_instr * rows_per_instr[mode];
        printf("gather occ=%d rows=%8u %-28s %9.1f us  %7.2f Grows/s  %6.2f cyc/row/SM  %6.2f cyc/instr/SM\n", occ, rows, gnames[mode], us,
               n_rows / us * 1e-3, us * 1e-6 * ghz * 1e9 * sms / n_rows, us * 1e-6 * ghz * 1e9 * sms / n_instr);
      }
    }
  }
  const char* rnames[] = {"RED.v4 4 rows/instr", "RED.32 1 row/instr", "RED.v2 2 rows/instr"};
  for (uint32_t rows : {1u << 11, 1u << 14, 1u << 17}) {       // 2 K rows (the coarse levels: contention) .. 16 MB
    for (int mode = 0; mode < 3; ++mode) {
      const int rows_per_instr[] = {4, 1, 2};
      const int grid = sms * 5, iters = 512;
      float us = 0;
      switch (mode) {
        case 0: us = time_us([&] { red_kernel<0><<<grid, 256>>>(table, rows - 1, iters); }); break;
        case 1: us = time_us([&] { red_kernel<1><<<grid, 256>>>(table, rows - 1, iters); }); break;
        case 2: us = time_us([&] { red_kernel<2><<<grid, 256>>>(table, rows - 1, iters); }); break;
      }
      const double n_instr = static_cast<double>(grid) * 8 * iters;
      const double n_rows = n_instr * rows_per_instr[mode];
      printf("red    rows=%8u %-28s %9.1f us  %7.2f Grows/s  %6.2f cyc/row/SM  (%.2f TB/s of 128-byte rows)\n", rows, rnames[mode], us,
             n_rows / us * 1e-3, us * 1e-6 * ghz * 1e9 * sms / n_rows, n_rows * 128 / us * 1e-6);
    }
  }
  return 0;
}
