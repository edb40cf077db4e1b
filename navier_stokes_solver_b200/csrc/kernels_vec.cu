// kernels_vec.cu -- BLAS-1 kernels of the Krylov solvers and Newton updates (HBM-bound, FP64).
//
// Replaces the Trilinos vector calls the reference makes (l2_norm, add, sadd, scale, *=, -=, =;
// NSSolverStationary.cpp:698, 720-721; NSSolverStationary.hpp:208, 293, 301, 306-307) and the
// deal.II Vector::add_and_dot chains inside SolverGMRES / SolverFGMRES / SolverCG.
// Reductions are two-stage and deterministic: fixed grid, per-CTA partials in a fixed order, the
// last CTA to finish sums them.  Results stay in device "slots" so that dependent kernels
// (the modified Gram-Schmidt chain) read their coefficient without a host round trip.
#include <cstring>

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

// ---- batched classical Gram-Schmidt: all projections of one pass in two kernels -------------
constexpr int MD = 8;  // vectors per register chunk of the multi-dot

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// result[m] = v_m . w for m < k  (k <= 31)
__global__ void __launch_bounds__(VT) k_multi_dot(VecList V, int k, const double *__restrict__ w, int64_t n, double *partial,
                                                  unsigned int *counter, double *result) {
  __shared__ const double *sv[32];
  __shared__ double sh[VT / 32][MD];
  __shared__ bool last;
  if (threadIdx.x < 32) sv[threadIdx.x] = V.v[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < k; c0 += MD) {
    double acc[MD];
#pragma unroll
    for (int m = 0; m < MD; ++m) acc[m] = 0.0;
    const int kc = min(MD, k - c0);
    for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) {
      const double wi = w[i];
#pragma unroll
      for (int m = 0; m < MD; ++m)
        if (m < kc) acc[m] += wi * sv[c0 + m][i];
    }
#pragma unroll
    for (int m = 0; m < MD; ++m) {
      const double r = warp_sum(acc[m]);
      if (lane == 0) sh[wp][m] = r;
    }
    __syncthreads();
    if (threadIdx.x < kc) {
      double r = 0;
#pragma unroll
      for (int q = 0; q < VT / 32; ++q) r += sh[q][threadIdx.x];
      partial[(size_t)blockIdx.x * 32 + c0 + threadIdx.x] = r;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicInc(counter, gridDim.x - 1);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    for (int m = wp; m < k; m += VT / 32) {
      double r = 0;
      for (int b = lane; b < gridDim.x; b += 32) r += __ldcg(&partial[(size_t)b * 32 + m]);
      r = warp_sum(r);
      if (lane == 0) result[m] = r;
    }
  }
}

// w -= sum_m coef[m] v_m ; *norm2 = w . w
__global__ void __launch_bounds__(VT) k_multi_axpy_norm(VecList V, int k, const double *coef, double *__restrict__ w, int64_t n, double *partial,
                                                        unsigned int *counter, double *norm2) {
  __shared__ const double *sv[32];
  __shared__ double sc[32];
  __shared__ double sh[VT / 32];
  __shared__ bool last;
  if (threadIdx.x < 32) { sv[threadIdx.x] = V.v[threadIdx.x]; sc[threadIdx.x] = threadIdx.x < k ? coef[threadIdx.x] : 0.0; }
  __syncthreads();
  double acc = 0;
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) {
    double wi = w[i];
    for (int m = 0; m < k; ++m) wi -= sc[m] * sv[m][i];
    w[i] = wi;
    acc += wi * wi;
  }
  const double s = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    partial[(size_t)blockIdx.x * 32] = s;
    __threadfence();
    const unsigned int t = atomicInc(counter, gridDim.x - 1);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double r = 0;
    for (int i = threadIdx.x; i < gridDim.x; i += VT) r += __ldcg(&partial[(size_t)i * 32]);
    r = block_sum(r, sh);
    if (threadIdx.x == 0) *norm2 = r;
  }
}

__device__ __forceinline__ int fg_check(const FgDev *st, int step, double value) {  // SolverControl::check -> gate code
  if (value <= st->tol) return 2;
  if (step >= st->max_it || isnan(value)) return 3;
  return 0;
}
__device__ __forceinline__ void fg_publish(const FgDev *st, FgRec *rec, long long seq) {
  rec->gate = st->gate; rec->it = st->it; rec->ny = st->ny; rec->res = st->res;
  __threadfence_system();
  *(volatile long long *)&rec->seq = seq;
}

__global__ void k_fg_begin(FgDev *st, const double *beta2, double tol, int max_it, int it0, FgRec *rec, long long seq) {
  const double beta = sqrt(*beta2);
  st->tol = tol; st->max_it = max_it; st->it = it0; st->ny = 0;
  st->res = beta; st->a = beta; st->g[0] = beta;
  const int v = fg_check(st, it0, beta);
  st->gate = v == 2 ? 2 : 0;   // deal.II leaves the cycle at its start only on success
  fg_publish(st, rec, seq);
}

// One warp: the inputs are fetched in parallel into shared memory, lane 0 runs the short sequential recurrences there.
__device__ void fg_step_body(FgDev *st, const double *slots, int j, int mode, FgRec *rec, long long seq) {
  __shared__ double col[34], cs[32], sn[32], g[34], h2s[32], ys[32], Rs[30 * 32];
  const int lane = threadIdx.x & 31;
  if (st->gate >= 2) { if (lane == 0) fg_publish(st, rec, seq); return; }
  // gather: coefficients of this column, the rotations so far, the rotated right-hand side
  {
    double v = 0.0;
    if (lane <= j) {
      if (mode == 0 || mode == 1 || mode == 4) v = slots[lane];
      else v = (mode == 2 ? st->R[j][lane] : slots[lane]) + slots[32 + lane];
    }
    col[lane] = v;
    h2s[lane] = lane <= j ? v * v : 0.0;
    cs[lane] = lane < j ? st->cs[lane] : 0.0;
    sn[lane] = lane < j ? st->sn[lane] : 0.0;
    g[lane] = lane <= j ? st->g[lane] : 0.0;
  }
  __syncwarp();
  double nrm2 = mode == 0 ? slots[j + 1] : (mode == 1 ? slots[64] : slots[65]);
  if (mode == 1 || mode == 4) {
    double h2 = 0;
    for (int i = 0; i <= j; ++i) h2 += h2s[i];   // same order on every lane
    // mode 4 (partitioned runs): |w'|^2 = |w|^2 - sum h^2 from the ONE reduction that brought the coefficients and |w|^2 = slots[j+1];
    // exact to ~100 eps relative whenever it is accepted (the same cancellation test sends the rest to a second pass with true norms)
    if (mode == 4) nrm2 = slots[j + 1] - h2;
    if (nrm2 < 0.01 * (nrm2 + h2)) {            // heavy cancellation: keep the first-pass coefficients, ask for a second pass
      if (lane <= j) st->R[j][lane] = col[lane];
      __syncwarp();
      if (lane == 0) { st->gate = 1; fg_publish(st, rec, seq); }
      return;
    }
  }
  const double a = sqrt(nrm2);
  int verdict = 0;
  double res = st->res;
  int it = st->it;
  if (j > 0) {
    res = fabs(g[j]);
    verdict = fg_check(st, ++it, res);
  }
  if (verdict != 0 || j == 29) {   // y of the (j+1) x j least-squares problem: back substitution in the rotated columns
    // the rotated columns come into shared memory in parallel; lane 0 substitutes back in the textbook order
    for (int e = lane; e < j * 32; e += 32) Rs[e] = st->R[e >> 5][e & 31];
    __syncwarp();
    if (lane == 0) {
      for (int i = j - 1; i >= 0; --i) {
        double s = g[i];
        for (int cc = i + 1; cc < j; ++cc) s -= Rs[cc * 32 + i] * ys[cc];
        ys[i] = s / Rs[i * 32 + i];
      }
    }
    __syncwarp();
    const double ycc = lane < j ? ys[lane] : 0.0;
    if (lane < j) st->y[lane] = ycc;
    __syncwarp();
    if (lane == 0) {
      st->a = a; st->res = res; st->it = it; st->ny = j; st->gate = verdict;
      fg_publish(st, rec, seq);
    }
    return;
  }
  if (lane == 0) {
    col[j + 1] = a;
    for (int i = 0; i < j; ++i) {
      const double t = cs[i] * col[i] + sn[i] * col[i + 1];
      col[i + 1] = -sn[i] * col[i] + cs[i] * col[i + 1];
      col[i] = t;
    }
    const double r = 1.0 / sqrt(col[j] * col[j] + col[j + 1] * col[j + 1]);
    const double snj = col[j + 1] * r, csj = col[j] * r;
    col[j] = csj * col[j] + snj * col[j + 1];
    st->sn[j] = snj; st->cs[j] = csj;
    st->g[j + 1] = -snj * g[j];
    st->g[j] = csj * g[j];
    st->a = a; st->res = res; st->it = it; st->gate = 0;
  }
  __syncwarp();
  if (lane <= j) st->R[j][lane] = col[lane];
  __syncwarp();
  if (lane == 0) fg_publish(st, rec, seq);
}

__global__ void __launch_bounds__(32) k_fg_step(FgDev *st, const double *slots, int j, int mode, FgRec *rec, long long seq) {
  fg_step_body(st, slots, j, mode, rec, seq);
}

// what the last CTA of the norm kernel does for the inner FGMRES on one GPU: the step kernel's work without its launch
struct FgFuse { FgDev *st; const double *slots; int j, mode; FgRec *rec; long long seq; };

// ---- wide variants for 16-byte aligned vectors: two elements per load, every vector of a chunk of 8 in flight at once --------
// The Krylov vectors of the README size are 4.3 MB each: a pass over 1 + k of them is latency-bound unless every thread has all its
// loads of a step in flight together.  One or two steps per thread, one reduction per chunk, 2 x #SMs CTAs of 512 threads.
constexpr int DT = 512, KC = 8;

// The Krylov basis is re-read by every Gram-Schmidt pass of a restart cycle (up to 30 x 4.3 MB at the README size, about the size of
// the 126 MB L2): its loads ask the L2 to keep the lines (evict-last), while the matrix streams of the SpMV and the sweeps pass
// through evict-first (tma.cuh).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double2 ld_keep(const double2 *p, uint64_t policy) {
  double2 v;
  asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy));
  return v;
}

__global__ void __launch_bounds__(DT, 2) k_multi_dot2(VecList V, int k, const double *__restrict__ w, int64_t n, double *partial, unsigned int *counter, double *result,
                                                      int keep_basis) {
  __shared__ const double *sv[32];
  __shared__ double sh[DT / 32][KC];
  __shared__ bool last;
  if (threadIdx.x < 32) sv[threadIdx.x] = V.v[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int64_t n2 = n >> 1;
  const double2 *w2 = reinterpret_cast<const double2 *>(w);
  const uint64_t keep = l2_policy_evict_last();
  for (int c0 = 0; c0 < k; c0 += KC) {
    const int kc = min(KC, k - c0);
    double acc[KC];
#pragma unroll
    for (int m = 0; m < KC; ++m) acc[m] = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)DT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * DT) {
      const double2 wv = w2[i];
      double2 vv[KC];
#pragma unroll
      for (int m = 0; m < KC; ++m) vv[m] = m < kc ? (keep_basis ? ld_keep(reinterpret_cast<const double2 *>(sv[c0 + m]) + i, keep) : reinterpret_cast<const double2 *>(sv[c0 + m])[i]) : make_double2(0.0, 0.0);
#pragma unroll
      for (int m = 0; m < KC; ++m) acc[m] = fma(wv.y, vv[m].y, fma(wv.x, vv[m].x, acc[m]));
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      const double wl = w[n - 1];
#pragma unroll
      for (int m = 0; m < KC; ++m) if (m < kc) acc[m] = fma(wl, sv[c0 + m][n - 1], acc[m]);
    }
#pragma unroll
    for (int m = 0; m < KC; ++m) {
      const double r = warp_sum(acc[m]);
      if (lane == 0) sh[wp][m] = r;
    }
    __syncthreads();
    if (threadIdx.x < kc) {
      double r = 0;
#pragma unroll
      for (int q = 0; q < DT / 32; ++q) r += sh[q][threadIdx.x];
      partial[(size_t)blockIdx.x * 32 + c0 + threadIdx.x] = r;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicInc(counter, gridDim.x - 1);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    for (int m = wp; m < k; m += DT / 32) {
      double r = 0;
      for (int b = lane; b < gridDim.x; b += 32) r += __ldcg(&partial[(size_t)b * 32 + m]);
      r = warp_sum(r);
      if (lane == 0) result[m] = r;
    }
  }
}

__global__ void __launch_bounds__(DT, 2) k_multi_axpy_norm2(VecList V, int k, const double *coef, double *__restrict__ w, int64_t n, double *partial,
                                                         unsigned int *counter, double *norm2, FgFuse fuse, int keep_basis) {
  __shared__ const double *sv[32];
  __shared__ double sc[32];
  __shared__ double sh[DT / 32];
  __shared__ bool last;
  if (threadIdx.x < 32) { sv[threadIdx.x] = V.v[threadIdx.x]; sc[threadIdx.x] = threadIdx.x < k ? coef[threadIdx.x] : 0.0; }
  __syncthreads();
  const int64_t n2 = n >> 1;
  double2 *w2 = reinterpret_cast<double2 *>(w);
  const uint64_t keep = l2_policy_evict_last();
  double acc = 0;
  for (int64_t i = blockIdx.x * (int64_t)DT + threadIdx.x; i < n2; i += (int64_t)gridDim.x * DT) {
    double2 wv = w2[i];
    for (int m0 = 0; m0 < k; m0 += KC) {
      double2 vv[KC];
#pragma unroll
      for (int m = 0; m < KC; ++m) vv[m] = m0 + m < k ? (keep_basis ? ld_keep(reinterpret_cast<const double2 *>(sv[m0 + m]) + i, keep) : reinterpret_cast<const double2 *>(sv[m0 + m])[i]) : make_double2(0.0, 0.0);
#pragma unroll
      for (int m = 0; m < KC; ++m) {   // coefficients beyond k are zero: same order of subtractions as the narrow kernel
        const double cm = sc[(m0 + m) & 31];
        if (m0 + m < k) { wv.x -= cm * vv[m].x; wv.y -= cm * vv[m].y; }
      }
    }
    w2[i] = wv;
    acc += wv.x * wv.x + wv.y * wv.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    double wl = w[n - 1];
    for (int m = 0; m < k; ++m) wl -= sc[m] * sv[m][n - 1];
    w[n - 1] = wl;
    acc += wl * wl;
  }
  // block sum over DT threads
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
#pragma unroll
    for (int q = 0; q < DT / 32; ++q) s += sh[q];
    partial[(size_t)blockIdx.x * 32] = s;
    __threadfence();
    const unsigned int t = atomicInc(counter, gridDim.x - 1);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double r = 0;
    for (int i = threadIdx.x; i < gridDim.x; i += DT) r += __ldcg(&partial[(size_t)i * 32]);
    r = warp_sum(r);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0;
#pragma unroll
      for (int q = 0; q < DT / 32; ++q) s += sh[q];
      *norm2 = s;
      __threadfence();
    }
    if (fuse.st) {   // uniform over the grid
      __syncthreads();
      if (threadIdx.x < 32) fg_step_body(fuse.st, fuse.slots, fuse.j, fuse.mode, fuse.rec, fuse.seq);
    }
  }
}

__global__ void __launch_bounds__(VT) k_scale_to(double *__restrict__ v, const double *__restrict__ x, const double *a, const int *gate, int64_t n) {
  if (gate && *gate != 0) return;
  const double av = *a;
  const bool ok = isfinite(av);
  const double inv = 1.0 / av;
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) v[i] = ok ? inv * x[i] : 0.0;
}

__global__ void __launch_bounds__(VT) k_multi_add(double *__restrict__ x, VecList V, const double *coef, const int *count, int64_t n) {
  __shared__ const double *sv[32];
  __shared__ double sc[32];
  const int k = *count;
  if (threadIdx.x < 32) { sv[threadIdx.x] = V.v[threadIdx.x]; sc[threadIdx.x] = threadIdx.x < k ? coef[threadIdx.x] : 0.0; }
  __syncthreads();
  if (k <= 0) return;
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) {
    double xi = x[i];
    for (int m = 0; m < k; ++m) xi += sc[m] * sv[m][i];
    x[i] = xi;
  }
}

}  // namespace

#define LAUNCHED(c) ((c).stat_launches++)

static void ensure_red(Ctx &c);
// the wide kernels need 16-byte aligned vectors and pay off once a vector spans a few CTAs
static bool wide_ok(const VecList &V, int k, const double *w, int64_t n) {
  if (n < 8192 || ((uintptr_t)w & 15)) return false;
  for (int m = 0; m < k; ++m) if ((uintptr_t)V.v[m] & 15) return false;
  return true;
}
static int wide_grid(Ctx &c, int64_t n) { return grid_for(n >> 1, DT, c.num_sms * 2); }

FgDev *fg_state(Ctx &c) {
  if (!c.fg_dev.p) {
    c.fg_dev.alloc(sizeof(FgDev));
    c.fg_dev.zero(c.stream);
    NSX_CUDA(cudaHostAlloc(&c.fg_rec, sizeof(FgRec), cudaHostAllocMapped));
    memset(c.fg_rec, 0, sizeof(FgRec));
    c.fg_seq = 0;
  }
  return reinterpret_cast<FgDev *>(c.fg_dev.p);
}
void fg_begin(Ctx &c, const double *beta2, double tol, int max_it, int it0) {
  FgDev *st = fg_state(c);
  k_fg_begin<<<1, 1, 0, c.stream>>>(st, beta2, tol, max_it, it0, (FgRec *)c.fg_rec, ++c.fg_seq); LAUNCHED(c);
}
void fg_step(Ctx &c, const double *slots, int j, int mode) {
  FgDev *st = fg_state(c);
  k_fg_step<<<1, 32, 0, c.stream>>>(st, slots, j, mode, (FgRec *)c.fg_rec, ++c.fg_seq); LAUNCHED(c);
}
// the classical pass's update + norm with the FGMRES step folded into the kernel's last CTA (one GPU, wide kernel); false if the
// caller has to launch fg_step itself
bool vec_multi_axpy_norm_fg(Ctx &c, int slot_norm, const VecList &V, int k, int slot_coef, double *w, int64_t n, const double *slots, int j, int mode) {
  ensure_red(c);
  if (c.comm || !wide_ok(V, k, w, n)) { vec_multi_axpy_norm_dev(c, slot_norm, V, k, slot_coef, w, n); return false; }
  FgDev *st = fg_state(c);
  k_multi_axpy_norm2<<<wide_grid(c, n), DT, 0, c.stream>>>(V, k, slot_ptr(c, slot_coef), w, n, c.red_partial.p, c.red_counter.p, slot_ptr(c, slot_norm),
                                                          FgFuse{st, slots, j, mode, (FgRec *)c.fg_rec, ++c.fg_seq}, c.l2_hints ? 1 : 0);
  LAUNCHED(c);
  return true;
}

FgRec fg_wait(Ctx &c) {
  const FgRec *rec = (const FgRec *)c.fg_rec;
  for (unsigned long long spins = 0;; ++spins) {
    if (__atomic_load_n(&rec->seq, __ATOMIC_ACQUIRE) == c.fg_seq) break;
    if ((spins & 0xfffff) == 0xfffff) {   // every ~1M polls: has the stream died?
      const cudaError_t e = cudaStreamQuery(c.stream);
      if (e != cudaSuccess && e != cudaErrorNotReady) throw CudaError(std::string("inner FGMRES: ") + cudaGetErrorString(e));
      if (e == cudaSuccess && __atomic_load_n(&rec->seq, __ATOMIC_ACQUIRE) != c.fg_seq) throw CudaError("inner FGMRES: the device record never arrived");
    }
  }
  return *rec;
}
void vec_scale_to_dev(Ctx &c, double *v, const double *x, const double *a, const int *gate, int64_t n) {
  if (!n) return;
  k_scale_to<<<vgrid(c, n), VT, 0, c.stream>>>(v, x, a, gate, n); LAUNCHED(c);
}
void vec_multi_add_dev(Ctx &c, double *x, const VecList &V, const double *coef, const int *count, int64_t n) {
  if (!n) return;
  k_multi_add<<<vgrid(c, n), VT, 0, c.stream>>>(x, V, coef, count, n); LAUNCHED(c);
}

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

static void ensure_red(Ctx &c);

double *slot_ptr(Ctx &c, int slot) { ensure_red(c); return c.red_result.p + slot; }


void vec_multi_dot_dev(Ctx &c, int slot0, const VecList &V, int k, const double *w, int64_t n) {
  ensure_red(c);
  if (k < 1 || k > 31) throw std::invalid_argument("multi-dot handles 1..31 vectors");
  if (wide_ok(V, k, w, n)) k_multi_dot2<<<wide_grid(c, n), DT, 0, c.stream>>>(V, k, w, n, c.red_partial.p, c.red_counter.p, slot_ptr(c, slot0), c.l2_hints ? 1 : 0);
  else k_multi_dot<<<vgrid(c, n), VT, 0, c.stream>>>(V, k, w, n, c.red_partial.p, c.red_counter.p, slot_ptr(c, slot0));
  LAUNCHED(c);
  allreduce_slots(c, slot0, k);
}
void vec_multi_axpy_norm_dev(Ctx &c, int slot_norm, const VecList &V, int k, int slot_coef, double *w, int64_t n, bool reduce) {
  ensure_red(c);
  if (wide_ok(V, k, w, n)) k_multi_axpy_norm2<<<wide_grid(c, n), DT, 0, c.stream>>>(V, k, slot_ptr(c, slot_coef), w, n, c.red_partial.p, c.red_counter.p, slot_ptr(c, slot_norm), FgFuse{nullptr, nullptr, 0, 0, nullptr, 0}, c.l2_hints ? 1 : 0);
  else k_multi_axpy_norm<<<vgrid(c, n), VT, 0, c.stream>>>(V, k, slot_ptr(c, slot_coef), w, n, c.red_partial.p, c.red_counter.p, slot_ptr(c, slot_norm));
  LAUNCHED(c);
  if (reduce) allreduce_slots(c, slot_norm, 1);
}

static void ensure_red(Ctx &c) {
  if (!c.red_partial.p) {
    c.red_partial.alloc((size_t)std::max(RED_MAX_BLOCKS, c.num_sms * 8) * 32);
    c.red_result.alloc(RED_SLOTS);
    c.red_result.zero(c.stream);
    c.red_counter.alloc(1);
    c.red_counter.zero(c.stream);
    NSX_CUDA(cudaMallocHost(&c.h_scalars, RED_SLOTS * sizeof(double)));
  }
}

void vec_dot_dev(Ctx &c, int slot, const double *a, const double *b, int64_t n) {
  ensure_red(c);
  if (!n) NSX_CUDA(cudaMemsetAsync(slot_ptr(c, slot), 0, sizeof(double), c.stream));
  else { k_dot<0><<<vgrid(c, n), VT, 0, c.stream>>>(a, b, nullptr, 0.0, nullptr, nullptr, nullptr, n, c.red_partial.p, c.red_counter.p, slot_ptr(c, slot)); LAUNCHED(c); }
  allreduce_slots(c, slot, 1);
}
void vec_add_and_dot_dev(Ctx &c, int slot, double *w, double sign, const double *dev_coef, const double *x, const double *v, int64_t n) {
  ensure_red(c);
  if (!n) NSX_CUDA(cudaMemsetAsync(slot_ptr(c, slot), 0, sizeof(double), c.stream));
  else { k_dot<1><<<vgrid(c, n), VT, 0, c.stream>>>(nullptr, nullptr, w, sign, dev_coef, x, v, n, c.red_partial.p, c.red_counter.p, slot_ptr(c, slot)); LAUNCHED(c); }
  allreduce_slots(c, slot, 1);
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
// L2 flush between timed launches: write 256 MiB (evicts everything, leaves dirty lines), then read
// another 256 MiB so that the lines the timed kernel evicts are clean -- otherwise the kernel under
// test pays the write-back of the flush itself.
__global__ void k_flush_read(const int4 *__restrict__ p, int64_t n, int *sink) {
  int acc = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc ^= __ldcg(p + i).x;
  if (acc == 0x5a5a5a5a) *sink = acc;
}
void flush_l2_cache(Ctx &c, int round) {
  const size_t half = c.flush.n / 2;
  NSX_CUDA(cudaMemsetAsync(c.flush.p, round & 0xff, half, c.stream));
  k_flush_read<<<c.num_sms * 8, 256, 0, c.stream>>>((const int4 *)(c.flush.p + half), (int64_t)(half / sizeof(int4)), (int *)c.flush.p);
}

// FP64 FMA peak of the device (measurement only): 8 independent dependent chains per thread, FP64_PEAK_ITERS trips
__global__ void __launch_bounds__(256) k_fp64_peak(double *sink, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  const double m = 1.0000001, b = 1e-9;
#pragma unroll 8
  for (int i = 0; i < FP64_PEAK_ITERS; ++i) {
    a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
    a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 1.2345e-300) *sink = s;
}
double fp64_peak_launch(Ctx &c, double *sink) {
  const int grid = c.num_sms * 8;
  k_fp64_peak<<<grid, 256, 0, c.stream>>>(sink, 0.5);
  c.stat_launches++;
  return (double)grid * 256.0 * FP64_PEAK_ITERS * 8.0 * 2.0;  // flops of one launch
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
