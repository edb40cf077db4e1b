// kernels_vec.cu -- BLAS-1 kernels of the Krylov solvers and Newton updates (HBM-bound, FP64).
//
// Replaces the Trilinos vector calls the reference makes (l2_norm, add, sadd, scale, *=, -=, =;
// NSSolverStationary.cpp:698, 720-721; NSSolverStationary.hpp:208, 293, 301, 306-307) and the
// deal.II Vector::add_and_dot chains inside SolverGMRES / SolverFGMRES / SolverCG.
// Reductions are two-stage and deterministic: fixed grid, per-CTA partials in a fixed order, the
// last CTA to finish sums them.  Results stay in device "slots" so that dependent kernels
// (the modified Gram-Schmidt chain) read their coefficient without a host round trip.
#include "device.cuh"

namespace nsx {

namespace {

constexpr int VT = 256;
constexpr int RED_MAX_BLOCKS = 1184;  // 148 SMs x 8

__global__ void __launch_bounds__(VT) k_copy(double *__restrict__ y, const double *__restrict__ x, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) y[i] = x[i];
}
__global__ void __launch_bounds__(VT) k_set(double *__restrict__ y, double a, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) y[i] = a;
}
__global__ void __launch_bounds__(VT) k_scale(double *__restrict__ y, double a, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) y[i] *= a;
}
__global__ void __launch_bounds__(VT) k_mul(double *__restrict__ y, const double *__restrict__ d, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) y[i] *= d[i];
}
__global__ void __launch_bounds__(VT) k_sadd(double *__restrict__ y, double s, double a, const double *__restrict__ x, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) y[i] = s * y[i] + a * x[i];
}
__global__ void __launch_bounds__(VT) k_equ(double *__restrict__ y, double a, const double *__restrict__ x, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) y[i] = a * x[i];
}
__global__ void __launch_bounds__(VT) k_axpy(double *__restrict__ y, double a, const double *dev_coef, const double *__restrict__ x, int64_t n) {
  const double c = dev_coef ? a * (*dev_coef) : a;
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) y[i] += c * x[i];
}

__device__ __forceinline__ double block_sum(double v, double *sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0;
  if (w == 0) {
    r = lane < VT / 32 ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

// FUSE = 0: result = a.b ; FUSE = 1: w += c x, result = w.v
template <int FUSE>
__global__ void __launch_bounds__(VT) k_dot(const double *__restrict__ a, const double *__restrict__ b, double *w, double coef,
                                            const double *dev_coef, const double *x, const double *v, int64_t n,
                                            double *partial, unsigned int *counter, double *result) {
  __shared__ double sh[VT / 32];
  __shared__ bool last;
  double acc = 0;
  if (FUSE == 0) {
    for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) acc += a[i] * b[i];
  } else {
    const double c = dev_coef ? coef * (*dev_coef) : coef;
    const bool alias = (v == w);
    for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) {
      const double wi = w[i] + c * x[i];
      w[i] = wi;
      acc += wi * (alias ? wi : v[i]);
    }
  }
  const double s = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = s;
    __threadfence();
    const unsigned int t = atomicInc(counter, gridDim.x - 1);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double r = 0;
    for (int i = threadIdx.x; i < gridDim.x; i += VT) r += __ldcg(&partial[i]);
    r = block_sum(r, sh);
    if (threadIdx.x == 0) *result = r;
  }
}

inline int vgrid(Ctx &c, int64_t n) { return grid_for(n, VT * 4, c.num_sms * 8); }

}  // namespace

#define LAUNCHED(c) ((c).stat_launches++)

void vec_copy(Ctx &c, double *y, const double *x, int64_t n) {
  if (!n || y == x) return;
  NSX_CUDA(cudaMemcpyAsync(y, x, n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
}
void vec_set(Ctx &c, double *y, double a, int64_t n) {
  if (!n) return;
  k_set<<<vgrid(c, n), VT, 0, c.stream>>>(y, a, n); LAUNCHED(c);
}
void vec_scale(Ctx &c, double *y, double a, int64_t n) {
  if (!n) return;
  k_scale<<<vgrid(c, n), VT, 0, c.stream>>>(y, a, n); LAUNCHED(c);
}
void vec_mul(Ctx &c, double *y, const double *d, int64_t n) {
  if (!n) return;
  k_mul<<<vgrid(c, n), VT, 0, c.stream>>>(y, d, n); LAUNCHED(c);
}
void vec_axpy(Ctx &c, double *y, double a, const double *x, int64_t n) {
  if (!n) return;
  k_axpy<<<vgrid(c, n), VT, 0, c.stream>>>(y, a, nullptr, x, n); LAUNCHED(c);
}
void vec_axpy_dev(Ctx &c, double *y, double sign, const double *dev_coef, const double *x, int64_t n) {
  if (!n) return;
  k_axpy<<<vgrid(c, n), VT, 0, c.stream>>>(y, sign, dev_coef, x, n); LAUNCHED(c);
}
void vec_sadd(Ctx &c, double *y, double s, double a, const double *x, int64_t n) {
  if (!n) return;
  k_sadd<<<vgrid(c, n), VT, 0, c.stream>>>(y, s, a, x, n); LAUNCHED(c);
}
void vec_equ(Ctx &c, double *y, double a, const double *x, int64_t n) {
  if (!n) return;
  k_equ<<<vgrid(c, n), VT, 0, c.stream>>>(y, a, x, n); LAUNCHED(c);
}

double *slot_ptr(Ctx &c, int slot) { return c.red_result.p + slot; }

static void ensure_red(Ctx &c) {
  if (!c.red_partial.p) {
    c.red_partial.alloc(RED_MAX_BLOCKS);
    c.red_result.alloc(RED_SLOTS);
    c.red_result.zero(c.stream);
    c.red_counter.alloc(1);
    c.red_counter.zero(c.stream);
    NSX_CUDA(cudaMallocHost(&c.h_scalars, RED_SLOTS * sizeof(double)));
  }
}

void vec_dot_dev(Ctx &c, int slot, const double *a, const double *b, int64_t n) {
  ensure_red(c);
  if (!n) { NSX_CUDA(cudaMemsetAsync(slot_ptr(c, slot), 0, sizeof(double), c.stream)); return; }
  k_dot<0><<<vgrid(c, n), VT, 0, c.stream>>>(a, b, nullptr, 0.0, nullptr, nullptr, nullptr, n, c.red_partial.p, c.red_counter.p, slot_ptr(c, slot));
  LAUNCHED(c);
}
void vec_add_and_dot_dev(Ctx &c, int slot, double *w, double sign, const double *dev_coef, const double *x, const double *v, int64_t n) {
  ensure_red(c);
  if (!n) { NSX_CUDA(cudaMemsetAsync(slot_ptr(c, slot), 0, sizeof(double), c.stream)); return; }
  k_dot<1><<<vgrid(c, n), VT, 0, c.stream>>>(nullptr, nullptr, w, sign, dev_coef, x, v, n, c.red_partial.p, c.red_counter.p, slot_ptr(c, slot));
  LAUNCHED(c);
}
double read_slot(Ctx &c, int slot) {
  ensure_red(c);
  NSX_CUDA(cudaMemcpyAsync(c.h_scalars + slot, slot_ptr(c, slot), sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  NSX_CUDA(cudaStreamSynchronize(c.stream));
  return c.h_scalars[slot];
}
void read_slots(Ctx &c, int first, int count, double *out) {
  ensure_red(c);
  NSX_CUDA(cudaMemcpyAsync(c.h_scalars + first, slot_ptr(c, first), count * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  NSX_CUDA(cudaStreamSynchronize(c.stream));
  for (int i = 0; i < count; ++i) out[i] = c.h_scalars[first + i];
}
double vec_dot(Ctx &c, const double *a, const double *b, int64_t n) {
  vec_dot_dev(c, RED_SLOTS - 1, a, b, n);
  return read_slot(c, RED_SLOTS - 1);
}
double vec_norm(Ctx &c, const double *a, int64_t n) { return sqrt(vec_dot(c, a, a, n)); }
double vec_add_and_dot(Ctx &c, double *w, double a, const double *x, const double *v, int64_t n) {
  vec_add_and_dot_dev(c, RED_SLOTS - 1, w, a, nullptr, x, v, n);
  return read_slot(c, RED_SLOTS - 1);
}

}  // namespace nsx
