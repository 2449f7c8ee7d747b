// EST - the learned "Event Spike Tensor" quantisation layer of the reference, forward pass
// (ev-YOLOv6/yolov6/models/learned_repr.py:143-179, QuantizationLayer.forward):
//
//   t <- t / max(t) per sample;  for every temporal bin i < C:  vox[b, p, i, y, x] += t * f(t - i / (C - 1))
//
// where f is the ValueLayer MLP (1 -> 100 -> 100 -> 1, LeakyReLU 0.1) of ONE scalar.  The reference evaluates that MLP
// C times per event (2 x 10^4 MACs per event and bin).  A LeakyReLU network of a scalar input is a piecewise-linear
// function, so the host compiles the weights once into sorted breakpoints with a slope and an intercept per segment
// (event_representation_study_b200/est.py, float64) and the kernel evaluates f with a binary search and one FMA: the
// layer becomes the same HBM / L2-atomic bound scatter as the voxel grids.  The backward pass with respect to the weights
// (k_est_backward below) reduces the output gradient to two sums per linear segment.
//
// Output layout: (B, H, W, 2C) float32, channel = p * C + i (the reference's torch.cat([vox[:, 0], vox[:, 1]], 1) in HWC),
// which is what k_image_pipeline reads for the letterbox that follows (learned_repr.py:94-141).
#include <math.h>

#include <algorithm>

#include "evrep_common.cuh"

namespace evrep {

__global__ void __launch_bounds__(256) k_est_tmax(const float* __restrict__ t, const int64_t* __restrict__ offsets, float* __restrict__ tmax) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  const int64_t s = offsets[b], e = offsets[b + 1];
  float m = -INFINITY;
  for (int64_t i = s + threadIdx.x; i < e; i += 256) m = fmaxf(m, t[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    tmax[b] = m;
  }
}

__global__ void __launch_bounds__(256) k_est_scatter(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y, const float* __restrict__ t,
                                                     const int8_t* __restrict__ p, const int64_t* __restrict__ offsets,
                                                     const float* __restrict__ tmax, int H, int W, int C, const double* __restrict__ breaks,
                                                     const double* __restrict__ slope, const double* __restrict__ icpt, int K,
                                                     float* __restrict__ out, uint32_t* __restrict__ flags) {
  const int b = blockIdx.y;
  const int64_t s = offsets[b], n = offsets[b + 1] - s;
  const float tm = tmax[b];
  float* grid = out + (size_t)b * H * W * 2 * C;
  uint32_t bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t xv = x[s + i], yv = y[s + i];
    if (xv >= (uint32_t)W || yv >= (uint32_t)H) { bad |= EVREP_WF_OUT_OF_RANGE; continue; }
    const int pv = p[s + i] > 0 ? 1 : 0;
    const float tn = t[s + i] / tm;  // float32 division like `t[...] /= t[...].max()`
    float* px = grid + ((size_t)yv * W + xv) * (2 * C) + pv * C;
    for (int ib = 0; ib < C; ++ib) {
      const float u = tn - (float)((double)ib / (double)(C > 1 ? C - 1 : 1));  // float32 tensor minus a Python float
      // segment of u: number of breakpoints <= u
      int lo = 0, hi = K;
      const double ud = (double)u;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(breaks + mid) <= ud) lo = mid + 1; else hi = mid;
      }
      const float f = (float)fma(__ldg(slope + lo), ud, __ldg(icpt + lo));
      atomicAdd(px + ib, tn * f);
    }
  }
  bad = __reduce_or_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicOr(flags + b, bad);
}

// Backward pass with respect to the ValueLayer weights.  Inside segment j the layer is f(u) = a_j u + c_j with a_j, c_j
// functions of the weights, so dL/dtheta = sum_j (G1_j da_j/dtheta + G0_j dc_j/dtheta) with
//     G0_j = sum g,   G1_j = sum g u   over the (event, bin) samples whose u falls into segment j,
//     g = dL/dout[b, y, x, p C + i] * tn   (out += tn * f(u)).
// This kernel produces G0 / G1 (2 (K + 1) doubles); est.py turns them into parameter gradients with autograd on two
// points per segment.  Shared-memory accumulation per CTA (double atomics on shared memory are CAS loops, but the
// contention is spread over the segments), one global double atomic per touched segment and CTA.
__global__ void __launch_bounds__(256) k_est_backward(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y, const float* __restrict__ t,
                                                      const int8_t* __restrict__ p, const int64_t* __restrict__ offsets,
                                                      const float* __restrict__ tmax, int H, int W, int C, const double* __restrict__ breaks, int K,
                                                      const float* __restrict__ grad_out, double* __restrict__ seg) {
  extern __shared__ double sh_seg[];  // [2][K + 1]
  for (int i = threadIdx.x; i < 2 * (K + 1); i += blockDim.x) sh_seg[i] = 0.0;
  __syncthreads();
  const int b = blockIdx.y;
  const int64_t s = offsets[b], n = offsets[b + 1] - s;
  const float tm = tmax[b];
  const float* grid = grad_out + (size_t)b * H * W * 2 * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t xv = x[s + i], yv = y[s + i];
    if (xv >= (uint32_t)W || yv >= (uint32_t)H) continue;
    const int pv = p[s + i] > 0 ? 1 : 0;
    const float tn = t[s + i] / tm;
    const float* px = grid + ((size_t)yv * W + xv) * (2 * C) + pv * C;
    for (int ib = 0; ib < C; ++ib) {
      const float u = tn - (float)((double)ib / (double)(C > 1 ? C - 1 : 1));
      int lo = 0, hi = K;
      const double ud = (double)u;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(breaks + mid) <= ud) lo = mid + 1; else hi = mid;
      }
      const double gv = (double)__ldg(px + ib) * (double)tn;
      if (gv != 0.0) {
        atomicAdd(sh_seg + lo, gv);
        atomicAdd(sh_seg + (K + 1) + lo, gv * ud);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * (K + 1); i += blockDim.x)
    if (sh_seg[i] != 0.0) atomicAdd(seg + i, sh_seg[i]);
}

int launch_est_backward(const uint16_t* x, const uint16_t* y, const float* t, const int8_t* p, const int64_t* win_offsets_host, int B, int H, int W,
                        int C, const double* breaks, int K, const float* grad_out, double* seg_sums, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream);

size_t est_workspace_bytes(int B) { return align_up(sizeof(int64_t) * (size_t)(B + 1), 256) + align_up(sizeof(float) * (size_t)B, 256) + align_up(sizeof(uint32_t) * (size_t)B, 256); }

int launch_est(const uint16_t* x, const uint16_t* y, const float* t, const int8_t* p, const int64_t* win_offsets_host, int B, int H, int W, int C,
               const double* breaks, const double* slope, const double* icpt, int K, float* out, void* workspace, size_t workspace_bytes,
               cudaStream_t stream) {
  if (workspace_bytes < est_workspace_bytes(B) || !workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) {
    set_error("EST: workspace must be 256-byte aligned and hold %zu bytes", est_workspace_bytes(B));
    return EVREP_EWORKSPACE;
  }
  char* wsp = (char*)workspace;
  int64_t* offsets = (int64_t*)wsp;
  float* tmax = (float*)(wsp + align_up(sizeof(int64_t) * (size_t)(B + 1), 256));
  uint32_t* flags = (uint32_t*)((char*)tmax + align_up(sizeof(float) * (size_t)B, 256));
  int64_t n_max = 0;
  for (int b = 0; b < B; ++b) {
    if (win_offsets_host[b + 1] < win_offsets_host[b] || win_offsets_host[b] < 0) {
      set_error("win_offsets must be non-decreasing and non-negative");
      return EVREP_EINVAL;
    }
    n_max = std::max<int64_t>(n_max, win_offsets_host[b + 1] - win_offsets_host[b]);
  }
  EVREP_CUDA_OK(cudaMemcpyAsync(offsets, win_offsets_host, sizeof(int64_t) * (size_t)(B + 1), cudaMemcpyHostToDevice, stream));
  EVREP_CUDA_OK(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (size_t)B, stream));
  EVREP_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * H * W * 2 * C, stream));
  k_est_tmax<<<B, 256, 0, stream>>>(t, offsets, tmax);
  if (n_max > 0) {
    dim3 grid((unsigned)std::min<int64_t>((n_max + 255) / 256, 148 * 8), (unsigned)B);
    k_est_scatter<<<grid, 256, 0, stream>>>(x, y, t, p, offsets, tmax, H, W, C, breaks, slope, icpt, K, out, flags);
  }
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

int launch_est_backward(const uint16_t* x, const uint16_t* y, const float* t, const int8_t* p, const int64_t* win_offsets_host, int B, int H, int W,
                        int C, const double* breaks, int K, const float* grad_out, double* seg_sums, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream) {
  if (workspace_bytes < est_workspace_bytes(B) || !workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) {
    set_error("EST: workspace must be 256-byte aligned and hold %zu bytes", est_workspace_bytes(B));
    return EVREP_EWORKSPACE;
  }
  const size_t smem = sizeof(double) * 2 * (size_t)(K + 1);
  if (smem > 200 * 1024) {
    set_error("EST backward: %d segments do not fit shared memory", K + 1);
    return EVREP_EUNSUPPORTED;
  }
  char* wsp = (char*)workspace;
  int64_t* offsets = (int64_t*)wsp;
  float* tmax = (float*)(wsp + align_up(sizeof(int64_t) * (size_t)(B + 1), 256));
  int64_t n_max = 0;
  for (int b = 0; b < B; ++b) {
    if (win_offsets_host[b + 1] < win_offsets_host[b] || win_offsets_host[b] < 0) {
      set_error("win_offsets must be non-decreasing and non-negative");
      return EVREP_EINVAL;
    }
    n_max = std::max<int64_t>(n_max, win_offsets_host[b + 1] - win_offsets_host[b]);
  }
  EVREP_CUDA_OK(cudaMemcpyAsync(offsets, win_offsets_host, sizeof(int64_t) * (size_t)(B + 1), cudaMemcpyHostToDevice, stream));
  EVREP_CUDA_OK(cudaMemsetAsync(seg_sums, 0, smem, stream));
  k_est_tmax<<<B, 256, 0, stream>>>(t, offsets, tmax);
  if (n_max > 0) {
    EVREP_CUDA_OK(cudaFuncSetAttribute(k_est_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)std::min<int64_t>((n_max + 255) / 256, 148 * 2), (unsigned)B);
    k_est_backward<<<grid, 256, smem, stream>>>(x, y, t, p, offsets, tmax, H, W, C, breaks, K, grad_out, seg_sums);
  }
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

}  // namespace evrep
