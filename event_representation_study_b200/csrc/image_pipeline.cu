// The step that follows every representation in the reference's detector pipelines, fused into one pass on the GPU
// (SURVEY.md 8f, rank 1):
//
//   rep (H, W, C)  --x255-->  per-channel cv2.resize  -->  letterbox (pad 114)  -->  HWC -> CHW, channel order reversed
//                  -->  / 255 (Trainer.prepro_data)
//
// Reference: ev-YOLOv6/yolov6/data/gen1_2yolo.py:230-265 (resize_image: keep the aspect ratio, INTER_AREA when shrinking
// without augmentation, else INTER_LINEAR), :321-341 + data_augment.py:31-83 (letterbox, pad value 114), :397
// (transpose + [::-1]); gen4/precompute_reps.py:216-251 (resize_image_process: squash to img_size x img_size);
// yolov6/core/engine.py:629-635 (/ 255).
//
// One thread per output pixel, all C channels: the taps of a pixel are C contiguous floats in the HWC input (read as
// float4) and its C results go to C planes, each written coalesced along x.  HBM bound: reads H W C 4 bytes, writes
// C S S 4 bytes per window.  The OpenCV arithmetic is followed tap for tap (float32 weights computed from the double
// scale as cv::resize does: half-pixel centres and border clamping for INTER_LINEAR, the cell-overlap table of
// computeResizeAreaTab for INTER_AREA).
#include <string.h>

#include "evrep_common.cuh"

namespace evrep {

struct ImgArgs {
  int B, H, W, C;            // input windows (B, H, W, C)
  int rw, rh;                // resized width / height
  int left, top;             // where the resized image sits in the canvas
  int out_w, out_h;          // canvas
  int interp;                // 1 linear (cv2), 2 area (cv2), 3 linear with float32 coordinates (torch interpolate)
  int reverse;               // 1: output channel c holds input channel C - 1 - c
  float scale_in, scale_out, pad;
  double sx, sy;             // source pixels per destination pixel (cv::resize: 1 / (dsize / ssize))
};

constexpr int IMG_MAX_TAPS = 6;  // INTER_AREA: floor(scale) + 2 taps per axis -> scale < 5 (1280 -> 320 is scale 4)

// cv::resize INTER_LINEAR coordinate: f = (d + 0.5) * scale - 0.5, s = floor(f), weight = f - s, clamped at the borders.
// OpenCV 4.13 (the reference's cv2 here) carries the coordinate in double for CV_64F images - which is what the
// reference resizes (float64 representations x 255) - and its CV_32F path (IPP) agrees to 1e-7; casting the coordinate
// to float first, as older generic code did, costs 1.5e-5 in the weight at x ~ 600 (checked against cv2 with numpy).
__device__ __forceinline__ void linear_taps(int d, double scale, int ssize, int* idx, float* wgt, int* n) {
  double fx = ((double)d + 0.5) * scale - 0.5;
  int s = (int)floor(fx);
  fx -= (double)s;
  if (s < 0) { s = 0; fx = 0.0; }
  if (s >= ssize - 1) { s = ssize - 1; fx = 0.0; }
  idx[0] = s;
  idx[1] = min(s + 1, ssize - 1);
  wgt[0] = (float)(1.0 - fx);
  wgt[1] = (float)fx;
  *n = 2;
}

// torch.nn.functional.interpolate(mode="bilinear", align_corners=False) on float32 tensors (what the reference's learned
// representation uses, learned_repr.py:113-115): the same taps, but scale and source coordinate are float32
// (area_pixel_compute_source_index: scale * (d + 0.5f) - 0.5f, clamped at 0)
__device__ __forceinline__ void linear_taps_f32(int d, float scale, int ssize, int* idx, float* wgt, int* n) {
  float fx = scale * ((float)d + 0.5f) - 0.5f;
  if (fx < 0.f) fx = 0.f;
  int s = (int)fx;
  if (s > ssize - 1) s = ssize - 1;
  const float l1 = fx - (float)s;
  idx[0] = s;
  idx[1] = s + (s < ssize - 1 ? 1 : 0);
  wgt[0] = 1.f - l1;
  wgt[1] = l1;
  *n = 2;
}

// cv::resize INTER_AREA (general path), computeResizeAreaTab for one destination index
__device__ __forceinline__ void area_taps(int d, double scale, int ssize, int* idx, float* wgt, int* n) {
  const double fsx1 = d * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, (double)ssize - fsx1);
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  int k = 0;
  if (sx1 - fsx1 > 1e-3 && k < IMG_MAX_TAPS) {
    idx[k] = sx1 - 1;
    wgt[k++] = (float)((sx1 - fsx1) / cell);
  }
  for (int s = sx1; s < sx2 && k < IMG_MAX_TAPS; ++s) {
    idx[k] = s;
    wgt[k++] = (float)(1.0 / cell);
  }
  if (fsx2 - sx2 > 1e-3 && k < IMG_MAX_TAPS) {
    idx[k] = sx2;
    wgt[k++] = (float)(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell);
  }
  *n = k;
}

// INTER_AREA at shrink factors beyond the tap tables (scale > IMG_MAX_TAPS - 2): the taps of a destination index are a run of
// consecutive source cells - an optional partial cell in front, whole cells, an optional partial cell behind - so they are
// generated on the fly from (first index, count, three weights) instead of being tabulated per thread.  Slower per tap than the
// tables (measured: +25 % on the 2x / 4x shrinks), so only the large factors take this path.
struct AreaRun {
  int i0, n;
  float wf, wm, wl;
  bool hf, hl;
  __device__ __forceinline__ float w(int k) const { return (k == 0 && hf) ? wf : ((k == n - 1 && hl) ? wl : wm); }
};
__device__ __forceinline__ AreaRun area_run(int d, double scale, int ssize) {
  const double fsx1 = d * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, (double)ssize - fsx1);
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  AreaRun t;
  t.hf = sx1 - fsx1 > 1e-3;
  t.hl = fsx2 - sx2 > 1e-3;
  t.i0 = t.hf ? sx1 - 1 : sx1;
  t.n = (t.hf ? 1 : 0) + (sx2 - sx1) + (t.hl ? 1 : 0);
  t.wf = (float)((sx1 - fsx1) / cell);
  t.wm = (float)(1.0 / cell);
  t.wl = (float)(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell);
  return t;
}

// one thread per output pixel, every channel in turn: the large-factor INTER_AREA path (any channel count)
__global__ void __launch_bounds__(256) k_image_pipeline_area_run(const float* __restrict__ rep, const ImgArgs a, float* __restrict__ out) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  const int b = blockIdx.z;
  if (ox >= a.out_w) return;
  const int C = a.C;
  const size_t plane = (size_t)a.out_w * a.out_h;
  float* dst = out + (size_t)b * C * plane + (size_t)oy * a.out_w + ox;
  const int rx = ox - a.left, ry = oy - a.top;
  if (rx < 0 || rx >= a.rw || ry < 0 || ry >= a.rh) {  // letterbox border
    const float v = a.pad * a.scale_out;
    for (int c = 0; c < C; ++c) __stcs(dst + (size_t)c * plane, v);
    return;
  }
  const AreaRun tx = area_run(rx, a.sx, a.W), ty = area_run(ry, a.sy, a.H);
  const float* src = rep + (size_t)b * a.H * a.W * C;
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
    for (int j = 0; j < ty.n; ++j) {
      float row = 0.f;
      const float* srow = src + ((size_t)(ty.i0 + j) * a.W + tx.i0) * C + c;
      for (int i = 0; i < tx.n; ++i) row += (__ldg(srow + (size_t)i * C) * a.scale_in) * tx.w(i);  // horizontal pass first, like cv::resize
      acc += row * ty.w(j);
    }
    __stcs(dst + (size_t)(a.reverse ? C - 1 - c : c) * plane, acc * a.scale_out);
  }
}

template <int CT>  // CT = compile-time channel count (multiple of 4), or 0 for the generic path
__global__ void __launch_bounds__(256) k_image_pipeline(const float* __restrict__ rep, const ImgArgs a, float* __restrict__ out) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  const int b = blockIdx.z;
  if (ox >= a.out_w) return;
  const int C = CT ? CT : a.C;
  const size_t plane = (size_t)a.out_w * a.out_h;
  float* dst = out + (size_t)b * C * plane + (size_t)oy * a.out_w + ox;
  const int rx = ox - a.left, ry = oy - a.top;
  if (rx < 0 || rx >= a.rw || ry < 0 || ry >= a.rh) {  // letterbox border
    const float v = a.pad * a.scale_out;
    for (int c = 0; c < C; ++c) __stcs(dst + (size_t)c * plane, v);
    return;
  }
  int xi[IMG_MAX_TAPS], yi[IMG_MAX_TAPS], nx, ny;
  float xw[IMG_MAX_TAPS], yw[IMG_MAX_TAPS];
  if (a.interp == 1) {
    linear_taps(rx, a.sx, a.W, xi, xw, &nx);
    linear_taps(ry, a.sy, a.H, yi, yw, &ny);
  } else if (a.interp == 3) {
    linear_taps_f32(rx, (float)a.W / (float)a.rw, a.W, xi, xw, &nx);
    linear_taps_f32(ry, (float)a.H / (float)a.rh, a.H, yi, yw, &ny);
  } else {
    area_taps(rx, a.sx, a.W, xi, xw, &nx);
    area_taps(ry, a.sy, a.H, yi, yw, &ny);
  }
  const float* src = rep + (size_t)b * a.H * a.W * C;
  if (CT) {
    float acc[CT ? CT : 4];
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[c] = 0.f;
    for (int j = 0; j < ny; ++j) {
      float row[CT ? CT : 4];
#pragma unroll
      for (int c = 0; c < CT; ++c) row[c] = 0.f;
      const float* srow = src + (size_t)yi[j] * a.W * CT;
      for (int i = 0; i < nx; ++i) {  // horizontal pass first, like cv::resize
        const float4* px = reinterpret_cast<const float4*>(srow + (size_t)xi[i] * CT);
        const float w = xw[i];
#pragma unroll
        for (int q = 0; q < CT / 4; ++q) {
          const float4 v = __ldg(px + q);
          row[4 * q + 0] += (v.x * a.scale_in) * w;
          row[4 * q + 1] += (v.y * a.scale_in) * w;
          row[4 * q + 2] += (v.z * a.scale_in) * w;
          row[4 * q + 3] += (v.w * a.scale_in) * w;
        }
      }
      const float wy = yw[j];
#pragma unroll
      for (int c = 0; c < CT; ++c) acc[c] += row[c] * wy;
    }
#pragma unroll
    for (int c = 0; c < CT; ++c) __stcs(dst + (size_t)(a.reverse ? CT - 1 - c : c) * plane, acc[c] * a.scale_out);
  } else {
    for (int c = 0; c < C; ++c) {
      float acc = 0.f;
      for (int j = 0; j < ny; ++j) {
        float row = 0.f;
        const float* srow = src + ((size_t)yi[j] * a.W) * C + c;
        for (int i = 0; i < nx; ++i) row += (__ldg(srow + (size_t)xi[i] * C) * a.scale_in) * xw[i];
        acc += row * yw[j];
      }
      __stcs(dst + (size_t)(a.reverse ? C - 1 - c : c) * plane, acc * a.scale_out);
    }
  }
}

// mode 0: gen1_2yolo.py (keep aspect ratio, then letterbox to img_size x img_size); mode 1: precompute_reps.py (squash)
int launch_image_pipeline(const float* rep, int B, int H, int W, int C, int img_size, int mode, int interp, float scale_in, float scale_out,
                          float pad, int reverse, float* out, cudaStream_t stream) {
  ImgArgs a;
  a.B = B; a.H = H; a.W = W; a.C = C;
  a.out_w = a.out_h = img_size;
  a.scale_in = scale_in; a.scale_out = scale_out; a.pad = pad; a.reverse = reverse;
  const double r = (double)img_size / (double)(H > W ? H : W);  // resize_image: r = img_size / max(h0, w0)
  if (mode == 0) {
    a.rw = (int)((double)W * r);  // int(w0 * r)
    a.rh = (int)((double)H * r);
    if (r == 1.0) { a.rw = W; a.rh = H; }
    // letterbox(auto=False, scaleup=False) on the resized image: it already fits, so only the border is added
    const double dw = (double)(img_size - a.rw) / 2.0, dh = (double)(img_size - a.rh) / 2.0;
    a.left = (int)lround(dw - 0.1);
    a.top = (int)lround(dh - 0.1);
    if (interp == 3) {  // letterbox_image_batch (learned_repr.py:130-131): floor division
      a.left = (img_size - a.rw) / 2;
      a.top = (img_size - a.rh) / 2;
    }
  } else {
    a.rw = a.rh = img_size;
    a.left = a.top = 0;
  }
  if (a.rw < 1 || a.rh < 1) {
    set_error("image pipeline: resized image would be empty");
    return EVREP_EINVAL;
  }
  if (interp == 0) interp = (r < 1.0) ? 2 : 1;  // INTER_AREA if r < 1 and not augment else INTER_LINEAR
  a.interp = interp;
  a.sx = 1.0 / ((double)a.rw / (double)W);  // cv::resize: inv_scale = dsize / ssize, scale = 1 / inv_scale
  a.sy = 1.0 / ((double)a.rh / (double)H);
  if (interp == 2) {
    if (a.sx < 1.0 || a.sy < 1.0) {  // cv::resize turns INTER_AREA on an enlarging axis into a linear variant: not mirrored here
      set_error("image pipeline: INTER_AREA needs both axes to shrink (got scale %.4f x %.4f)", a.sx, a.sy);
      return EVREP_EUNSUPPORTED;
    }
  }
  dim3 grid((unsigned)((img_size + 255) / 256), (unsigned)img_size, (unsigned)B);
  if (interp == 2 && (a.sx > IMG_MAX_TAPS - 2 || a.sy > IMG_MAX_TAPS - 2)) {  // beyond the tap tables
    k_image_pipeline_area_run<<<grid, 256, 0, stream>>>(rep, a, out);
    EVREP_CUDA_OK(cudaGetLastError());
    return EVREP_OK;
  }
  const bool vec = (reinterpret_cast<uintptr_t>(rep) & 15u) == 0;
  if (C == 12 && vec)
    k_image_pipeline<12><<<grid, 256, 0, stream>>>(rep, a, out);
  else
    k_image_pipeline<0><<<grid, 256, 0, stream>>>(rep, a, out);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// The training-time augmentation that follows the letterbox (gen1_2yolo.py:365-391): random_affine = cv2.warpAffine(img,
// M[:2], dsize, borderValue=(114, 114, 114)) (data_augment.py:110-123), then the up-down / left-right flips of general_augment
// (:210-228).  The random draws (get_transform_matrix, the flip coins) stay with the caller, like the label bookkeeping; this
// is the image side.  cv::warpAffine with INTER_LINEAR on a float image: the matrix is inverted in double, destination
// coordinates are mapped in FIXED POINT (10 fractional bits, rounded to 1/32 pixel), the four taps are blended with float
// weights from a 32 x 32 table (accumulated in double there, in a float FMA chain here); taps outside the image take the border value of their channel, and
// the 3-entry borderValue tuple becomes a 4-entry scalar that channel k indexes with k & 3 - so every fourth channel of a
// 12-channel representation is padded with 0, not 114.  A numpy restatement of exactly this agrees with cv2 4.13 to the last
// bit on float64 images (oracle/image_pipeline.py::warp_affine_restated, checked there against cv2 itself).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int WARP_MAX_WINDOWS = 64;  // per launch: the per-window matrices travel as kernel arguments
struct WarpParams {
  double m[WARP_MAX_WINDOWS][6];  // inverse maps (destination -> source), as cv::warpAffine computes them
  int flip[WARP_MAX_WINDOWS];     // bit 0: up-down, bit 1: left-right (applied after the warp)
  float border[4];
  int C, in_h, in_w, out_h, out_w, reverse;
  float scale_out;
};

template <int CT>  // compile-time channel count (unrolled), or 0 for the generic loop
__global__ void __launch_bounds__(256) k_warp_affine(const float* __restrict__ in, const __grid_constant__ WarpParams p, float* __restrict__ out) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y, b = blockIdx.z;
  if (ox >= p.out_w) return;
  // the flips act on the warped image: output pixel (oy, ox) is warped pixel (y, x)
  const int y = (p.flip[b] & 1) ? p.out_h - 1 - oy : oy;
  const int x = (p.flip[b] & 2) ? p.out_w - 1 - ox : ox;
  const double* M = p.m[b];
  // imgwarp.cpp WarpAffineInvoker: AB_BITS = 10, round_delta = AB_SCALE / INTER_TAB_SIZE / 2 = 16; saturate_cast<int>(double)
  // rounds half to even; the products must not be contracted into FMAs (the CPU code is not)
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(M[0], (double)x), 1024.0));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(M[3], (double)x), 1024.0));
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[1], (double)y), M[2]), 1024.0)) + 16;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[4], (double)y), M[5]), 1024.0)) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));  // saturate_cast<short>
  const float fx = (float)(X & 31) * (1.f / 32.f), fy = (float)(Y & 31) * (1.f / 32.f);
  const float w00 = __fmul_rn(1.f - fy, 1.f - fx), w01 = __fmul_rn(1.f - fy, fx), w10 = __fmul_rn(fy, 1.f - fx), w11 = __fmul_rn(fy, fx);
  const bool x0 = sx >= 0 && sx < p.in_w, x1 = sx + 1 >= 0 && sx + 1 < p.in_w;
  const bool y0 = sy >= 0 && sy < p.in_h, y1 = sy + 1 >= 0 && sy + 1 < p.in_h;
  const size_t in_plane = (size_t)p.in_h * p.in_w, out_plane = (size_t)p.out_h * p.out_w;
  const float* src = in + (size_t)b * p.C * in_plane;
  float* dst = out + (size_t)b * p.C * out_plane + (size_t)oy * p.out_w + ox;
  const size_t o00 = (size_t)(y0 ? sy : 0) * p.in_w + (x0 ? sx : 0), o01 = (size_t)(y0 ? sy : 0) * p.in_w + (x1 ? sx + 1 : 0);
  const size_t o10 = (size_t)(y1 ? sy + 1 : 0) * p.in_w + (x0 ? sx : 0), o11 = (size_t)(y1 ? sy + 1 : 0) * p.in_w + (x1 ? sx + 1 : 0);
  const int C = CT ? CT : p.C;
  if (y0 && y1 && x0 && x1) {  // all four taps inside the image (nearly every pixel): no border selects
#pragma unroll
    for (int c = 0; c < (CT ? CT : 1); ++c) {
      for (int cc = c; cc < C; cc += (CT ? C : 1)) {
        const float* pl = src + (size_t)cc * in_plane;
        const float sum = fmaf(__ldg(pl + o11), w11, fmaf(__ldg(pl + o10), w10, fmaf(__ldg(pl + o01), w01, __ldg(pl + o00) * w00)));
        __stcs(dst + (size_t)(p.reverse ? C - 1 - cc : cc) * out_plane, sum * p.scale_out);
      }
    }
    return;
  }
  for (int c = 0; c < C; ++c) {
    const float* pl = src + (size_t)c * in_plane;
    const float bv = p.border[c & 3];
    const float v00 = (y0 && x0) ? __ldg(pl + o00) : bv, v01 = (y0 && x1) ? __ldg(pl + o01) : bv;
    const float v10 = (y1 && x0) ? __ldg(pl + o10) : bv, v11 = (y1 && x1) ? __ldg(pl + o11) : bv;
    // remapBilinear accumulates S[0] w[0] + S[1] w[1] + S[step] w[2] + S[step + 1] w[3] in double on the reference's float64 image;
    // here the image is float32 already, and a float FMA chain stays within 2 ulp of that sum (1e-7 relative, against a bar of
    // 1e-5) at a quarter of the instructions (the double version was issue bound at 2.1 TB/s)
    const float sum = fmaf(v11, w11, fmaf(v10, w10, fmaf(v01, w01, v00 * w00)));
    __stcs(dst + (size_t)(p.reverse ? C - 1 - c : c) * out_plane, sum * p.scale_out);
  }
}

// in: DEVICE float32 (B, C, in_h, in_w); M_host: B x 6 doubles, the FORWARD matrices handed to cv2.warpAffine (rows of M[:2]);
// flips_host: B ints (bit 0 up-down, bit 1 left-right) or NULL
int launch_warp_affine(const float* in, int B, int C, int in_h, int in_w, const double* M_host, const int* flips_host, int out_h, int out_w,
                       const float* border4, int reverse, float scale_out, float* out, cudaStream_t stream) {
  for (int b0 = 0; b0 < B; b0 += WARP_MAX_WINDOWS) {
    const int nb = B - b0 < WARP_MAX_WINDOWS ? B - b0 : WARP_MAX_WINDOWS;
    WarpParams p;
    memset(&p, 0, sizeof(p));
    for (int b = 0; b < nb; ++b) {
      // cv::warpAffine without WARP_INVERSE_MAP (imgwarp.cpp): the inverse of the 2 x 3 map, in double, in this operation order
      double M[6];
      for (int k = 0; k < 6; ++k) M[k] = M_host[(size_t)(b0 + b) * 6 + k];
      double D = M[0] * M[4] - M[1] * M[3];
      D = D != 0 ? 1. / D : 0;
      const double A11 = M[4] * D, A22 = M[0] * D;
      M[0] = A11; M[1] *= -D;
      M[3] *= -D; M[4] = A22;
      const double b1 = -M[0] * M[2] - M[1] * M[5];
      const double b2 = -M[3] * M[2] - M[4] * M[5];
      M[2] = b1; M[5] = b2;
      for (int k = 0; k < 6; ++k) p.m[b][k] = M[k];
      p.flip[b] = flips_host ? flips_host[b0 + b] : 0;
    }
    for (int k = 0; k < 4; ++k) p.border[k] = border4[k];
    p.C = C; p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w; p.reverse = reverse; p.scale_out = scale_out;
    dim3 grid((unsigned)((out_w + 255) / 256), (unsigned)out_h, (unsigned)nb);
    if (C == 12)
      k_warp_affine<12><<<grid, 256, 0, stream>>>(in + (size_t)b0 * C * in_h * in_w, p, out + (size_t)b0 * C * out_h * out_w);
    else
      k_warp_affine<0><<<grid, 256, 0, stream>>>(in + (size_t)b0 * C * in_h * in_w, p, out + (size_t)b0 * C * out_h * out_w);
    EVREP_CUDA_OK(cudaGetLastError());
  }
  return EVREP_OK;
}

}  // namespace evrep
