// Packed host wire format -> the SoA event arrays of the engine.
//
// End to end the engine is bound by the host link: 9 B/event (x u16, y u16, t i32, p i8) at ~54 GB/s is 6 Gevents/s into a
// pipeline that consumes 56 Gevents/s from HBM.  The data loader can pack an event into ONE 32-bit word
//     x | y << xb | pc << (xb + yb) | dt << (xb + yb + 2)        pc = p & 3 (0, +1, -1 -> 0, 1, 3)
// where dt = t - (base timestamp of the event's block of 2^blk_shift consecutive events of its window) as long as every dt
// fits the 30 - xb - yb bits that are left (format 4: 4 B/event + 4 B/block; at 1280x720, 9 bits: blocks of 64 events may span
// 511 us), or into that word plus a 16-bit dt (format 6: 6 B/event, any sensor up to 16384^2, blocks may span 65 ms).
// Block bases are relative to the window's first timestamp: every representation only uses timestamp differences inside
// a window, so the decoded int32 timestamps equal the originals up to one constant per window.
//
// One decode kernel writes x, y, t, p (9 B/event) for the regular entry points; its 13 B/event of HBM traffic are
// ~0.1 ms per 32 M events, against 2.4 ms that the 4-byte format saves on the link.
#include <algorithm>
#include <vector>

#include "evrep_common.cuh"

namespace evrep {

// grid-stride over chunks of UNPACK_CHUNK consecutive events of ONE window; chunk_prefix[b] = number of chunks before window b
// (the window of a chunk is found by binary search: the three per-window tables are a few hundred bytes, small enough for
// the driver to embed the upload in the command stream - a larger pageable upload would make every call wait for the stream)
constexpr int UNPACK_CHUNK = 1024;

template <bool SEP16>
__global__ void __launch_bounds__(256) k_unpack(const uint32_t* __restrict__ word, const uint16_t* __restrict__ dt16, const int32_t* __restrict__ tbase,
                                                const int64_t* __restrict__ offsets, const int64_t* __restrict__ blk_prefix,
                                                const int64_t* __restrict__ chunk_prefix, int B, int xb, int yb, int blk_shift,
                                                uint16_t* __restrict__ x, uint16_t* __restrict__ y, int32_t* __restrict__ t, int8_t* __restrict__ p) {
  const uint32_t xm = (1u << xb) - 1u, ym = (1u << yb) - 1u;
  const int ps = xb + yb, ds = xb + yb + 2;
  const int64_t n_chunks = __ldg(chunk_prefix + B);
  for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    int lo = 0, hi = B;  // last window with chunk_prefix[w] <= c (windows without events own no chunk and are skipped)
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(chunk_prefix + mid) <= c) lo = mid; else hi = mid;
    }
    const int w = lo;
    const int64_t w0 = __ldg(offsets + w), n = __ldg(offsets + w + 1) - w0;
    const int64_t l0 = (c - __ldg(chunk_prefix + w)) * UNPACK_CHUNK;
    const int32_t* base = tbase + __ldg(blk_prefix + w);
#pragma unroll
    for (int k = 0; k < UNPACK_CHUNK / 256; ++k) {
      const int64_t l = l0 + k * 256 + threadIdx.x;
      if (l >= n) break;
      const int64_t e = w0 + l;
      const uint32_t v = __ldg(word + e);
      const uint32_t pc = (v >> ps) & 3u;
      const int32_t d = SEP16 ? (int32_t)__ldg(dt16 + e) : (int32_t)(v >> ds);
      x[e] = (uint16_t)(v & xm);
      y[e] = (uint16_t)((v >> xb) & ym);
      t[e] = __ldg(base + (l >> blk_shift)) + d;
      p[e] = (int8_t)(pc == 3u ? -1 : (int)pc);  // pc == 2 never leaves the packer
    }
  }
}

size_t unpack_workspace_bytes(int B, int64_t total) {
  (void)total;
  return align_up(3 * sizeof(int64_t) * (size_t)(B + 1), 256);
}

// win_offsets: HOST, B + 1.  word / dt16 / tbase: DEVICE (the packed payload, already uploaded).  workspace: DEVICE.
int launch_unpack(const uint32_t* word, const uint16_t* dt16, const int32_t* tbase, const int64_t* win_offsets_host, int B, int fmt, int xb, int yb,
                  int blk_shift, uint16_t* x, uint16_t* y, int32_t* t, int8_t* p, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if ((fmt != 4 && fmt != 6) || xb < 1 || yb < 1 || blk_shift < 0 || blk_shift > 16 || (fmt == 4 && xb + yb > 29) || (fmt == 6 && xb + yb > 30)) {
    set_error("unpack: bad format (fmt %d, xb %d, yb %d, blk_shift %d)", fmt, xb, yb, blk_shift);
    return EVREP_EINVAL;
  }
  const int64_t total = win_offsets_host[B];
  if (total == 0) return EVREP_OK;
  if (workspace_bytes < unpack_workspace_bytes(B, total) || !workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) {
    set_error("unpack: workspace must be 256-byte aligned and hold %zu bytes", unpack_workspace_bytes(B, total));
    return EVREP_EWORKSPACE;
  }
  std::vector<int64_t> host(3 * (size_t)(B + 1));
  int64_t* offs = host.data();
  int64_t* bpre = offs + (B + 1);
  int64_t* cpre = bpre + (B + 1);
  int64_t nb = 0, nc = 0;
  const int64_t bs = (int64_t)1 << blk_shift;
  for (int b = 0; b <= B; ++b) offs[b] = win_offsets_host[b];
  for (int b = 0; b < B; ++b) {
    const int64_t n = win_offsets_host[b + 1] - win_offsets_host[b];
    if (n < 0) {
      set_error("win_offsets must be non-decreasing");
      return EVREP_EINVAL;
    }
    bpre[b] = nb;
    cpre[b] = nc;
    nb += (n + bs - 1) / bs;
    nc += (n + UNPACK_CHUNK - 1) / UNPACK_CHUNK;
  }
  bpre[B] = nb;
  cpre[B] = nc;
  EVREP_CUDA_OK(cudaMemcpyAsync(workspace, host.data(), sizeof(int64_t) * host.size(), cudaMemcpyHostToDevice, stream));
  const int64_t* d = (const int64_t*)workspace;
  const int grid = (int)std::min<int64_t>(nc, 148 * 16);
  if (fmt == 6)
    k_unpack<true><<<grid, 256, 0, stream>>>(word, dt16, tbase, d, d + (B + 1), d + 2 * (B + 1), B, xb, yb, blk_shift, x, y, t, p);
  else
    k_unpack<false><<<grid, 256, 0, stream>>>(word, nullptr, tbase, d, d + (B + 1), d + 2 * (B + 1), B, xb, yb, blk_shift, x, y, t, p);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// Format 3: 3 bytes per event.  A time-sorted stream rarely moves more than a microsecond or two between neighbouring events
// (1 M events in 300 ms: 0.3 us on average), so the time of an event is coded as the DIFFERENCE to its predecessor in 2 bits
// (0, 1, 2; 3 = escape: the difference is the next entry of a side table), next to x, y and a polarity bit:
//     x | y << xb | (p > 0) << (xb + yb) | code << (xb + yb + 1)          xb + yb <= 21, p in {-1, +1}
// Events are grouped in blocks of 64 of one window; a block occupies exactly 192 bytes (the last block of a window is padded),
// starts with its first event's timestamp (tbase, relative to the window's first) and knows how many escapes precede it
// (esc_prefix).  One warp decodes one block: two events per lane, the escapes located with a ballot, the differences summed
// with a warp scan.  3.13 B/event on the link against 4.06 for format 4 and 9 for the SoA arrays.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_unpack_delta(const uint8_t* __restrict__ rec3, const int32_t* __restrict__ tbase, const uint32_t* __restrict__ esc_prefix,
                                                      const uint32_t* __restrict__ esc_dt, const int64_t* __restrict__ offsets,
                                                      const int64_t* __restrict__ blk_prefix, int B, int xb, int yb, uint16_t* __restrict__ x,
                                                      uint16_t* __restrict__ y, int32_t* __restrict__ t, int8_t* __restrict__ p) {
  const uint32_t xm = (1u << xb) - 1u, ym = (1u << yb) - 1u;
  const int ps = xb + yb, cs = xb + yb + 1;
  const int lane = threadIdx.x & 31;
  const int64_t n_blocks = __ldg(blk_prefix + B);
  const uint32_t esc0 = __ldg(esc_prefix);  // a group's tables are slices of the batch's: prefixes are relative to the first
  for (int64_t blk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); blk < n_blocks; blk += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    int lo = 0, hi = B;  // last window with blk_prefix[w] <= blk
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(blk_prefix + mid) <= blk) lo = mid; else hi = mid;
    }
    const int w = lo;
    const int64_t w0 = __ldg(offsets + w), n = __ldg(offsets + w + 1) - w0;
    const int64_t l0 = (blk - __ldg(blk_prefix + w)) * 64;  // first event of the block inside its window
    const int cnt = (int)min((int64_t)64, n - l0);
    // the lane's two records: bytes [6 lane, 6 lane + 6) of the block, read as three aligned 16-bit words
    const uint16_t* h = reinterpret_cast<const uint16_t*>(rec3 + (size_t)blk * 192) + 3 * lane;
    const uint32_t h0 = __ldg(h), h1 = __ldg(h + 1), h2 = __ldg(h + 2);
    const uint32_t r0 = h0 | ((h1 & 0xffu) << 16), r1 = (h1 >> 8) | (h2 << 8);
    const int e0 = 2 * lane, e1 = e0 + 1;
    const uint32_t c0 = e0 < cnt ? (r0 >> cs) & 3u : 0u, c1 = e1 < cnt ? (r1 >> cs) & 3u : 0u;
    const bool x0 = c0 == 3u, x1 = c1 == 3u;
    const uint32_t b0 = __ballot_sync(0xffffffffu, x0), b1 = __ballot_sync(0xffffffffu, x1);
    const uint32_t below = (1u << lane) - 1u;
    const uint32_t k0 = __popc(b0 & below) + __popc(b1 & below);  // escapes of the block before event e0
    const uint32_t ebase = __ldg(esc_prefix + blk) - esc0;
    uint32_t d0 = x0 ? __ldg(esc_dt + ebase + k0) : c0;
    uint32_t d1 = x1 ? __ldg(esc_dt + ebase + k0 + (x0 ? 1u : 0u)) : c1;
    uint32_t pair = d0 + d1;  // inclusive scan of the lane sums
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, pair, o);
      if (lane >= o) pair += up;
    }
    const int32_t tb = __ldg(tbase + blk);
    const int64_t g = w0 + l0;
    if (e0 < cnt) {
      x[g + e0] = (uint16_t)(r0 & xm);
      y[g + e0] = (uint16_t)((r0 >> xb) & ym);
      t[g + e0] = tb + (int32_t)(pair - d1);
      p[g + e0] = ((r0 >> ps) & 1u) ? (int8_t)1 : (int8_t)-1;
    }
    if (e1 < cnt) {
      x[g + e1] = (uint16_t)(r1 & xm);
      y[g + e1] = (uint16_t)((r1 >> xb) & ym);
      t[g + e1] = tb + (int32_t)pair;
      p[g + e1] = ((r1 >> ps) & 1u) ? (int8_t)1 : (int8_t)-1;
    }
  }
}

size_t unpack_delta_workspace_bytes(int B) { return align_up(2 * sizeof(int64_t) * (size_t)(B + 1), 256); }

// rec3: DEVICE, 192 bytes per block; tbase: DEVICE int32 per block; esc_prefix: DEVICE uint32 per block + 1 (may start at any
// value: a slice of a larger table); esc_dt: DEVICE uint32 per escape of these blocks; win_offsets: HOST, B + 1
int launch_unpack_delta(const uint8_t* rec3, const int32_t* tbase, const uint32_t* esc_prefix, const uint32_t* esc_dt, const int64_t* win_offsets_host,
                        int B, int xb, int yb, uint16_t* x, uint16_t* y, int32_t* t, int8_t* p, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream) {
  if (xb < 1 || yb < 1 || xb + yb > 21) {
    set_error("unpack (format 3): x_bits + y_bits must be at most 21 (got %d + %d)", xb, yb);
    return EVREP_EINVAL;
  }
  const int64_t total = win_offsets_host[B];
  if (total == 0) return EVREP_OK;
  if (workspace_bytes < unpack_delta_workspace_bytes(B) || !workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) {
    set_error("unpack (format 3): workspace must be 256-byte aligned and hold %zu bytes", unpack_delta_workspace_bytes(B));
    return EVREP_EWORKSPACE;
  }
  if ((reinterpret_cast<uintptr_t>(rec3) & 1u)) { set_error("unpack (format 3): the record array must be 2-byte aligned"); return EVREP_EINVAL; }
  std::vector<int64_t> host(2 * (size_t)(B + 1));
  int64_t* offs = host.data();
  int64_t* bpre = offs + (B + 1);
  int64_t nb = 0;
  for (int b = 0; b <= B; ++b) offs[b] = win_offsets_host[b];
  for (int b = 0; b < B; ++b) {
    const int64_t n = win_offsets_host[b + 1] - win_offsets_host[b];
    if (n < 0) {
      set_error("win_offsets must be non-decreasing");
      return EVREP_EINVAL;
    }
    bpre[b] = nb;
    nb += (n + 63) / 64;
  }
  bpre[B] = nb;
  EVREP_CUDA_OK(cudaMemcpyAsync(workspace, host.data(), sizeof(int64_t) * host.size(), cudaMemcpyHostToDevice, stream));
  const int64_t* d = (const int64_t*)workspace;
  const int grid = (int)std::min<int64_t>((nb + 7) / 8, 148 * 16);
  k_unpack_delta<<<grid, 256, 0, stream>>>(rec3, tbase, esc_prefix, esc_dt, d, d + (B + 1), B, xb, yb, x, y, t, p);
  EVREP_CUDA_OK(cudaGetLastError());
  return EVREP_OK;
}

}  // namespace evrep
