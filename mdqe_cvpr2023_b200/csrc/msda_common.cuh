// Shared device helpers for the multi-scale deformable attention kernels (sm_100a).
//
// Semantics follow the reference operator (paths relative to /root/reference/mdqe/models/ops):
//   pixel mapping  x = loc_x * W - 0.5, y = loc_y * H - 0.5      src/cuda/ms_deform_im2col_cuda.cuh:285-286
//   a corner contributes only when it lies inside the level        src/cuda/ms_deform_im2col_cuda.cuh:60-82
//   which equals grid_sample(bilinear, zeros, align_corners=False) functions/ms_deform_attn_func.py:58-59
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace msda {

constexpr int kMaxLevels = 32;       // "levels" seen by the op (pyramid levels, or T frames in temporal mode)
constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;

struct LevelInfo {                    // staged once per CTA in shared memory
  int H, W, start, pad;
};

// One bilinear corner of one sample: where to read, and with which (attention-scaled) weight.
// Out-of-range corners carry w == 0 and are never dereferenced.
struct __align__(8) Slot {
  uint32_t off;                       // element offset of the corner's channel row in `value`
  float w;
};

__device__ __forceinline__ float ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ double ld_as_float(const double* p) { return __ldg(p); }

__device__ __forceinline__ void st_from_float(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_from_float(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void st_from_float(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void st_from_float(double* p, double v) { *p = v; }

// 16-byte channel vector of the value tensor -> fp32 registers.
template <typename VT> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int kN = 4;
  __device__ __forceinline__ static void load(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ static void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int kN = 8;
  __device__ __forceinline__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {       // bf16 -> fp32 is a 16-bit shift
      v[2 * i] = __uint_as_float(u[i] << 16);
      v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      u[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(u[0], u[1], u[2], u[3]);
  }
};

// fp32 vector reduction into global memory (one 16-byte L2 atomic): SASS REDG.E.ADD.F32x4.
__device__ __forceinline__ void red_add_f32x4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// the same, predicated inside the instruction: no branch around it, so a batch of gathers / reductions stays one basic block
__device__ __forceinline__ void red_add_f32x4_if(bool on, float* p, float a, float b, float c, float d) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
               ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "r"(static_cast<uint32_t>(on)) : "memory");
}

// A level table entry is usable only if its window [start, start + H*W) lies inside the S rows of a batch element.  The
// reference asserts sum(H_l*W_l) == S on the host (ms_deform_attn.py:134, one device sync per call); here a table that does not
// fit disables that level (H = W = 0: every sample of it is out of range, contributes 0 and gets zero gradients) instead of
// letting the gathers and, worse, the grad_value reductions run past the tensor.
__device__ __forceinline__ bool level_fits(int64_t H, int64_t W, int64_t start, int S) {
  return H >= 0 && W >= 0 && start >= 0 && H <= 0x7fffffff && W <= 0x7fffffff && H * W <= static_cast<int64_t>(S) &&
         start <= static_cast<int64_t>(S) - H * W;
}

// Programmatic dependent launch (see launch_kernel in msda_launch.cuh).  pdl_wait(): blocks until the kernel this one was
// serialised behind has completed and its writes are visible (a no-op for a normal launch) -- must precede every global
// memory access.  pdl_trigger(): lets the NEXT kernel's CTAs be scheduled once every CTA of this grid has passed this point.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Stage the per-level geometry (int64 on device in the reference API) into shared memory.
__device__ __forceinline__ void stage_levels(LevelInfo* s_lvl, const int64_t* __restrict__ shapes,
                                             const int64_t* __restrict__ level_start, int L, int S) {
  if (threadIdx.x < L) {
    const int64_t H = shapes[2 * threadIdx.x], W = shapes[2 * threadIdx.x + 1], start = level_start[threadIdx.x];
    const bool ok = level_fits(H, W, start, S);
    LevelInfo li;
    li.H = ok ? static_cast<int>(H) : 0;
    li.W = ok ? static_cast<int>(W) : 0;
    li.start = ok ? static_cast<int>(start) : 0;
    li.pad = 0;
    s_lvl[threadIdx.x] = li;
  }
}

// Paired-corner bf16 value layout (csrc/msda_packed.cu; also written directly by the value_proj GEMM epilogue, gemm3x.cuh):
//   packed[n][prow][m] = { value[n, cell(y, xp - 1), m, 0:32], value[n, cell(y, xp), m, 0:32] } as 2 x 64 bytes of bf16,
//   prow = pstart_l + y * (W_l + 1) + xp, xp in [0, W_l]; batch stride 2 S M lines.
struct PackedLevel { int H, W, start, pstart; };

// level table + packed prefix (thread 0 adds the <= 32 terms); returns nothing: read s_lvl / s_total after a barrier
__device__ __forceinline__ void stage_packed_levels(PackedLevel* s_lvl, int* s_total, const int64_t* __restrict__ shapes,
                                                    const int64_t* __restrict__ level_start, int L, int S) {
  if (threadIdx.x == 0) {
    int p = 0;
    for (int l = 0; l < L; ++l) {
      const int64_t H = shapes[2 * l], W = shapes[2 * l + 1], st = level_start ? level_start[l] : 0;
      const bool ok = level_fits(H, W, st, S) && static_cast<int64_t>(p) + H * (W + 1) <= 2 * static_cast<int64_t>(S);
      PackedLevel pl;
      pl.H = ok ? static_cast<int>(H) : 0;
      pl.W = ok ? static_cast<int>(W) : 0;
      pl.start = ok ? static_cast<int>(st) : 0;
      pl.pstart = p;
      s_lvl[l] = pl;
      p += pl.H * (pl.W + 1);
    }
    *s_total = p;
  }
}


// Geometry of one sample: integer corner, fractional parts and per-corner validity.
struct SampleGeom {
  int x0, y0;
  float lx, ly;                        // fractional position inside the cell
  bool okx0, okx1, oky0, oky1;
};

__device__ __forceinline__ SampleGeom sample_geom(float locx, float locy, int H, int W) {
  SampleGeom g;
  const float x = locx * static_cast<float>(W) - 0.5f;
  const float y = locy * static_cast<float>(H) - 0.5f;
  const float fx = floorf(x), fy = floorf(y);
  g.lx = x - fx;
  g.ly = y - fy;
  // The reference takes a sample only if -1 < x < W and -1 < y < H, strictly (ms_deform_im2col_cuda.cuh:288); NaN /
  // huge coordinates fail the comparisons -> all corners invalid (sample skipped, zero gradient).
  const bool sane = (x > -1.f) && (x < static_cast<float>(W)) && (y > -1.f) && (y < static_cast<float>(H));
  g.x0 = sane ? static_cast<int>(fx) : -8;
  g.y0 = sane ? static_cast<int>(fy) : -8;
  g.okx0 = g.x0 >= 0 && g.x0 < W;
  g.okx1 = g.x0 + 1 >= 0 && g.x0 + 1 < W;
  g.oky0 = g.y0 >= 0 && g.y0 < H;
  g.oky1 = g.y0 + 1 >= 0 && g.y0 + 1 < H;
  return g;
}

}  // namespace msda
