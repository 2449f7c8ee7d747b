// Design-time microbenchmarks (round 1): which primitives can carry the
// event -> pixel scatter on a B200?  Numbers decide between
//   (A) global L2 atomics straight into per-pixel accumulators and
//   (B) spatial pre-bucketing + shared-memory per-tile reduction.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics_bench atomics_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// ---------- shared-memory atomics ----------
template <int MODE>  // 0 add+return, 1 exch, 2 add no return, 3 plain store (LSU floor)
__global__ void __launch_bounds__(1024) smem_atom(uint32_t* sink, int iters, int table_words) {
  extern __shared__ uint32_t tab[];
  for (int i = threadIdx.x; i < table_words; i += blockDim.x) tab[i] = 0;
  __syncthreads();
  uint32_t s = mix(blockIdx.x * 1024u + threadIdx.x + 1u), acc = 0;
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    uint32_t a = (s >> 8) % (uint32_t)table_words;
    if (MODE == 0) acc += atomicAdd(&tab[a], 1u);
    else if (MODE == 1) acc += atomicExch(&tab[a], s);
    else if (MODE == 2) atomicAdd(&tab[a], 1u);
    else tab[a] = s;
  }
  __syncthreads();
  if (acc == 0xdeadbeef || tab[threadIdx.x % table_words] == 0xdeadbeef) sink[0] = acc;
}

// ---------- global atomics ----------
template <int MODE>  // 0 red.add.u32, 1 red.add.u64, 2 red.max.u32, 3 red.v4.f32, 4 atom.add.u32 w/ return, 5 red.v2.f32
__global__ void __launch_bounds__(256) gmem_atom(void* table, uint64_t n_slots, int iters, uint32_t* sink) {
  uint32_t s = mix(blockIdx.x * 256u + threadIdx.x + 1u), acc = 0;
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    uint64_t a = (uint64_t)(mix(s)) % n_slots;
    if (MODE == 0) atomicAdd((uint32_t*)table + a, 1u);
    else if (MODE == 1) atomicAdd((unsigned long long*)table + a, 1ull);
    else if (MODE == 2) atomicMax((uint32_t*)table + a, s);
    else if (MODE == 3) {
      float* p = (float*)table + a * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(1.f), "f"(0.f), "f"(1.f) : "memory");
    } else if (MODE == 4) acc += atomicAdd((uint32_t*)table + a, 1u);
    else if (MODE == 5) {
      float* p = (float*)table + a * 2;
      asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(1.f), "f"(1.f) : "memory");
    }
  }
  if (acc == 0xdeadbeef) sink[0] = acc;
}

// same pixel gets k consecutive-word atomics (models "several accumulators of one pixel")
__global__ void __launch_bounds__(256) gmem_multi(uint32_t* table, uint64_t n_pix, int words_per_pix, int k, int iters) {
  uint32_t s = mix(blockIdx.x * 256u + threadIdx.x + 1u);
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    uint64_t a = (uint64_t)(mix(s)) % n_pix;
    uint32_t* p = table + a * words_per_pix;
    for (int j = 0; j < k; ++j) atomicAdd(p + j, 1u);
  }
}

// ---------- streaming ----------
__global__ void __launch_bounds__(256) stream_copy(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) out[i] = in[i];
}
__global__ void __launch_bounds__(256) stream_write(uint4* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  uint4 v = make_uint4(1, 2, 3, 4);
  for (; i < n; i += st) __stcs(out + i, v);
}
__global__ void __launch_bounds__(256) stream_read(const uint4* __restrict__ in, size_t n, uint32_t* sink) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (; i < n; i += st) { uint4 v = __ldcs(in + i); acc += v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0xdeadbeef) sink[0] = acc;
}
// scattered 8-byte stores: each warp lane writes to a different run (models the bucket scatter)
__global__ void __launch_bounds__(256) scatter8(uint2* __restrict__ out, uint32_t n_runs, uint32_t run_len, int iters) {
  // thread owns nothing; position = run r (random) + cursor (iteration) -> 8B store
  uint32_t s = mix(blockIdx.x * 256u + threadIdx.x + 1u);
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    uint32_t r = mix(s) % n_runs;
    uint32_t c = (blockIdx.x * iters + i) % run_len;
    out[(size_t)r * run_len + c] = make_uint2(s, i);
  }
}

template <typename F>
float time_it(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s SMs %d smem/blk optin %zu L2 %d MB clock %d kHz\n", prop.name, prop.multiProcessorCount,
         prop.sharedMemPerBlockOptin, prop.l2CacheSize >> 20, prop.clockRate);
  uint32_t* sink; CK(cudaMalloc(&sink, 4));
  const int SM = prop.multiProcessorCount;

  // shared atomics
  {
    int iters = 4096;
    for (int words : {1024, 8192, 24576}) {
      size_t sh = words * 4;
      for (int mode = 0; mode < 4; ++mode) {
        for (int cps : {1, 2}) {
          auto launch = [&]() {
            int grid = SM * cps;
            if (mode == 0) smem_atom<0><<<grid, 1024, sh>>>(sink, iters, words);
            if (mode == 1) smem_atom<1><<<grid, 1024, sh>>>(sink, iters, words);
            if (mode == 2) smem_atom<2><<<grid, 1024, sh>>>(sink, iters, words);
            if (mode == 3) smem_atom<3><<<grid, 1024, sh>>>(sink, iters, words);
          };
          cudaFuncSetAttribute(smem_atom<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
          cudaFuncSetAttribute(smem_atom<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
          cudaFuncSetAttribute(smem_atom<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
          cudaFuncSetAttribute(smem_atom<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
          float ms = time_it(launch);
          double ops = (double)SM * cps * 1024 * iters;
          printf("SMEM mode=%d(0 add+ret,1 exch,2 add,3 store) words=%d ctas/sm=%d : %.1f Gops/s (%.3f ms)\n", mode, words, cps,
                 ops / ms * 1e-6, ms);
        }
      }
    }
  }
  CK(cudaGetLastError());

  // global atomics
  {
    size_t maxbytes = (size_t)1 << 30;
    void* table; CK(cudaMalloc(&table, maxbytes)); CK(cudaMemset(table, 0, maxbytes));
    int iters = 256; int grid = SM * 32;
    for (size_t mb : {4, 32, 96, 512}) {
      for (int mode = 0; mode < 6; ++mode) {
        size_t slot = (mode == 1) ? 8 : (mode == 3) ? 16 : (mode == 5) ? 8 : 4;
        uint64_t n_slots = (mb << 20) / slot;
        auto launch = [&]() {
          if (mode == 0) gmem_atom<0><<<grid, 256>>>(table, n_slots, iters, sink);
          if (mode == 1) gmem_atom<1><<<grid, 256>>>(table, n_slots, iters, sink);
          if (mode == 2) gmem_atom<2><<<grid, 256>>>(table, n_slots, iters, sink);
          if (mode == 3) gmem_atom<3><<<grid, 256>>>(table, n_slots, iters, sink);
          if (mode == 4) gmem_atom<4><<<grid, 256>>>(table, n_slots, iters, sink);
          if (mode == 5) gmem_atom<5><<<grid, 256>>>(table, n_slots, iters, sink);
        };
        float ms = time_it(launch, 3);
        double ops = (double)grid * 256 * iters;
        printf("GMEM mode=%d(0 red.u32,1 red.u64,2 max.u32,3 red.v4f32,4 atom.u32+ret,5 red.v2f32) table=%zuMB : %.1f Gops/s (%.3f ms)\n",
               mode, mb, ops / ms * 1e-6, ms);
      }
    }
    CK(cudaGetLastError());
    // k atomics to consecutive words of one pixel record (16 words = 64B per pixel), table 59 MB (1 Mpx * 64 B)
    for (int k : {1, 2, 4, 8}) {
      uint64_t n_pix = 921600;
      auto launch = [&]() { gmem_multi<<<grid, 256>>>((uint32_t*)table, n_pix, 16, k, iters); };
      float ms = time_it(launch, 3);
      double ev = (double)grid * 256 * iters;
      printf("GMEM multi k=%d atomics/event on one 64B pixel record (1 Mpx): %.1f Gevents/s, %.1f Gatomics/s\n", k, ev / ms * 1e-6,
             ev * k / ms * 1e-6);
    }
    CK(cudaGetLastError());

    // streaming
    size_t n16 = maxbytes / 32;  // 512 MB in, 512 MB out
    uint4* in = (uint4*)table; uint4* out = in + n16;
    for (int g : {SM * 8, SM * 16, SM * 32}) {
      float ms = time_it([&]() { stream_copy<<<g, 256>>>(in, out, n16); });
      printf("STREAM copy grid=%d: %.1f GB/s (r+w)\n", g, 2.0 * n16 * 16 / ms * 1e-6);
      ms = time_it([&]() { stream_write<<<g, 256>>>(out, n16); });
      printf("STREAM write grid=%d: %.1f GB/s\n", g, 1.0 * n16 * 16 / ms * 1e-6);
      ms = time_it([&]() { stream_read<<<g, 256>>>(in, n16, sink); });
      printf("STREAM read grid=%d: %.1f GB/s\n", g, 1.0 * n16 * 16 / ms * 1e-6);
    }
    CK(cudaGetLastError());
    // scattered 8B stores into n_runs runs (256 MB region)
    for (uint32_t n_runs : {128u, 1024u, 8192u}) {
      uint32_t run_len = (256u << 20) / 8 / n_runs;
      int it = 128;
      float ms = time_it([&]() { scatter8<<<grid, 256>>>((uint2*)table, n_runs, run_len, it); });
      double n = (double)grid * 256 * it;
      printf("SCATTER8 runs=%u: %.1f Gstores/s = %.1f GB/s payload\n", n_runs, n / ms * 1e-6, n * 8 / ms * 1e-6);
    }
    CK(cudaGetLastError());
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
