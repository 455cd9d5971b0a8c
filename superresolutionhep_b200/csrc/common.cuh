// Shared device helpers and parameter structs for the srhep sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

namespace srhep {

// Bounds-asserting debug build (nvcc -DSRHEP_BOUNDS -> libsrhep_bounds.so, exercised by tests/test_gpu_bounds.py; compute-sanitizer is not
// available on the GPU pool): every global-memory row / event index of the hot kernels is checked against the extent of its buffer, a
// violation prints the site and traps, which the host sees as a launch failure.  Compiles to nothing in the production library.
#ifdef SRHEP_BOUNDS
#define SRHEP_CHECK(cond) do { if (!(cond)) { printf("srhep bounds violation %s:%d: %s (block %d thread %d)\n", __FILE__, __LINE__, #cond, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define SRHEP_CHECK(cond) do { } while (0)
#endif
struct Extents { int rows_cap = 0; int n_events = 0; };      // rows of the pass workspace buffers; events of the binding (per-event arrays)

constexpr float kLnEps = 1e-5f;      // nn.LayerNorm default eps (models/dense.py:62)
constexpr float kLeaky = 0.01f;      // nn.LeakyReLU default slope

// max(x, 0.01 x): the same value as `x > 0 ? x : 0.01 x` for every finite x, in two instructions (FMUL + FMNMX) instead of three (FSETP + FMUL + FSEL)
__device__ __forceinline__ float leaky_relu(float x) { return fmaxf(x, kLeaky * x); }
__device__ __forceinline__ float silu(float x) { return x / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Blocked layout of the fp32 residual stream (h_dim = 256) used by the tcgen05 chain path: the epilogue threads
// of a 128-row tile each own one row (TMEM lane = row), so a warp's 32-byte accesses would touch 32 different
// 128-byte lines of a row-major matrix.  Here the 8 floats (32 B) of column group cg of the 128 rows of a tile
// are contiguous: [tile][cg = col / 8 (32 groups)][row in tile (128)][8 floats] -> a warp access = 1 KB contiguous.
__host__ __device__ __forceinline__ size_t xblk_index(int row, int col) {
    return ((((size_t)(row >> 7) * 32 + (size_t)(col >> 3)) << 7) + (size_t)(row & 127)) * 8 + (size_t)(col & 7);
}

// One ODE stage = one network evaluation followed by `out = base + coef * v`.
// Lives in device memory so that a captured CUDA graph of one evaluation can be replayed
// for every stage: kernels read sp[*stage_idx].
struct StageParams {
    float        t;       // time of this evaluation (all events of a sampling pass share it)
    float        coef;    // dt (euler / 2nd midpoint stage) or dt/2 (1st midpoint stage)
    const float* x_in;    // network input state (packed rows of this pass)
    const float* base;    // y0 of the update (may be null when out is null)
    float*       out;     // base + coef*v (may be null)
    float*       vout;    // raw velocity (may be null)
};

// How a kernel finds the stage it is working for.  Graph replay: sp[*idx] (the captured
// launch arguments never change, the device-side counter does).  Direct launch: `fixed`.
// t_event (optional, indexed by global event id) overrides the shared stage time -- the
// FlowModel.forward(batch, x, time_step) entry point takes one time per event.
struct StageRef {
    const StageParams* sp      = nullptr;
    const int*         idx     = nullptr;
    StageParams        fixed   = {};
    const float*       t_event = nullptr;
};
__device__ __forceinline__ StageParams load_stage(const StageRef& r) { return r.sp ? r.sp[*r.idx] : r.fixed; }

// Epilogue description shared by the fp32 and the bf16 GEMM kernels:
//   val = act(acc + bias[n] + row_bias[event(row)][n])
//   C   = resid ? resid[row][n] + gate[event(row)][n] * val : val
struct GemmEpilogue {
    const float* bias        = nullptr;
    const float* row_bias    = nullptr;  int ld_row_bias = 0;
    const int*   row_event   = nullptr;  // null: event(row) = row
    int          act         = 0;        // 0 none, 1 LeakyReLU(0.01)
    const float* gate        = nullptr;  int ld_gate = 0;
    const float* resid       = nullptr;  int ld_resid = 0;
    // Optional fused LayerNorm of the produced row (tcgen05 GEMM with BN == row width only):
    //   y = (LN(C_row) * ln_w + ln_b) * (1 + ln_scale[event]) + ln_shift[event];  if ln_second: y = LN(y)
    // written as a 16-bit A operand for the next GEMM (diffusion_transformer.py:8-9,38-52; dense.py:62).
    void*        ln_out      = nullptr;  int ld_ln = 0;
    const float* ln_w        = nullptr;  const float* ln_b = nullptr;
    const float* ln_shift    = nullptr;  const float* ln_scale = nullptr;  int ld_lnmod = 0;
    int          ln_second   = 0;
};

}  // namespace srhep
