// Micro-benchmark: how many bytes does a random 32-byte read cost on B200, per load flavour?
// Each thread reads one random 32-byte sector of a 4 GiB buffer.  Run under
//   ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum
// usage: sector_fetch [granularity_limit]   (limit set with cudaDeviceSetLimit before anything else when given)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

template <int V>
__device__ __forceinline__ uint64_t load(const unsigned char* p) {
  uint64_t a, b, c, d;
  if constexpr (V == 0) asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  if constexpr (V == 1) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  if constexpr (V == 2) asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  if constexpr (V == 3) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  if constexpr (V == 4) asm volatile("ld.global.nc.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  if constexpr (V == 5) asm volatile("ld.global.cs.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  if constexpr (V == 6) asm volatile("ld.global.nc.L1::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  if constexpr (V == 7) {  // two 16-byte loads of the same sector
    uint64_t e, f;
    asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(e), "=l"(f) : "l"(p + 16));
    c = e; d = f;
  }
  if constexpr (V == 8) {  // a single 8-byte load
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(a) : "l"(p));
    b = c = d = 0;
  }
  if constexpr (V == 9) {  // 8-byte load, L1 bypass
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(a) : "l"(p));
    b = c = d = 0;
  }
  return a ^ b ^ c ^ d;
}

template <int V>
__global__ void probe(const unsigned char* buf, uint64_t sectors, uint64_t n, uint64_t* out) {
  uint64_t acc = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t s = mix(i * 0x9e3779b97f4a7c15ull + V) % sectors;
    acc ^= load<V>(buf + s * 32);
  }
  if (acc == 0x1234567) out[0] = acc;
}

template <int V>
void run(const char* name, const unsigned char* buf, uint64_t sectors, uint64_t n, uint64_t* out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<V><<<148 * 16, 256>>>(buf, sectors, n, out);
  cudaEventRecord(e0);
  probe<V><<<148 * 16, 256>>>(buf, sectors, n, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%-28s %8.3f ms  %7.1f M reads/ms  (%s)\n", name, ms, n / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

// footprint / memory-level-parallelism sweep with the default load flavour
template <int ILP>
__global__ void probe_ilp(const unsigned char* buf, uint64_t sectors, uint64_t n, uint64_t* out) {
  uint64_t acc = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * ILP) {
    uint64_t v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      const uint64_t s = mix((i + k * stride) * 0x9e3779b97f4a7c15ull) % sectors;
      v[k] = load<0>(buf + s * 32);
    }
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc ^= v[k];
  }
  if (acc == 0x1234567) out[0] = acc;
}

template <int ILP>
void sweep(const unsigned char* buf, uint64_t bytes, uint64_t n, uint64_t* out, int blocks_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe_ilp<ILP><<<148 * blocks_per_sm, 256>>>(buf, bytes / 32, n, out);
  cudaEventRecord(e0);
  probe_ilp<ILP><<<148 * blocks_per_sm, 256>>>(buf, bytes / 32, n, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("footprint %6llu MiB  ilp %d  blocks/SM %2d : %8.3f ms  %6.1f G reads/s\n", (unsigned long long)(bytes >> 20), ILP,
         blocks_per_sm, ms, n / ms * 1e-6);
}

// random WRITES: one aligned 32-byte store (a full sector), or three 8-byte stores at a random 8-byte offset (a 24-byte
// particle record: partial sectors, read-modify-write in L2 / DRAM)
template <int MODE>
__global__ void probe_write(unsigned char* buf, uint64_t units, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t s = mix(i * 0x9e3779b97f4a7c15ull + 77 + MODE) % units;
    if (MODE == 0) {
      uint64_t* p = reinterpret_cast<uint64_t*>(buf + s * 32);
      asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(i), "l"(i + 1), "l"(i + 2), "l"(i + 3) : "memory");
    } else {
      uint64_t* p = reinterpret_cast<uint64_t*>(buf + s * 24);
      p[0] = i; p[1] = i + 1; p[2] = i + 2;
    }
  }
}

template <int MODE>
void run_write(const char* name, unsigned char* buf, uint64_t bytes, uint64_t n) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const uint64_t units = MODE == 0 ? bytes / 32 : bytes / 24 - 1;
  probe_write<MODE><<<148 * 8, 256>>>(buf, units, n);
  cudaEventRecord(e0);
  probe_write<MODE><<<148 * 8, 256>>>(buf, units, n);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%-40s footprint %6llu MiB : %8.3f ms  %6.1f G writes/s\n", name, (unsigned long long)(bytes >> 20), ms, n / ms * 1e-6);
}

int main(int argc, char** argv) {
  if (argc > 1) {
    cudaError_t r = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
    size_t v = 0;
    cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity);
    printf("limit -> %zu (%s)\n", v, cudaGetErrorString(r));
  }
  const uint64_t bytes = 4ull << 30, sectors = bytes / 32, n = 1ull << 26;
  unsigned char* buf; uint64_t* out;
  cudaMalloc(&buf, bytes); cudaMalloc(&out, 8);
  cudaMemset(buf, 1, bytes);
  run<0>("nc.v4.u64", buf, sectors, n, out);
  run<1>("v4.u64", buf, sectors, n, out);
  run<2>("cg.v4.u64", buf, sectors, n, out);
  run<3>("nc.L1::no_allocate.v4.u64", buf, sectors, n, out);
  run<4>("nc.L2::64B.v4.u64", buf, sectors, n, out);
  run<5>("cs.v4.u64", buf, sectors, n, out);
  run<6>("nc.L1::evict_first.v4.u64", buf, sectors, n, out);
  run<7>("2 x nc.v2.u64", buf, sectors, n, out);
  run<8>("nc.u64", buf, sectors, n, out);
  run<9>("cg.u64", buf, sectors, n, out);
  cudaFree(buf);
  const uint64_t big = 32ull << 30;
  cudaMalloc(&buf, big);
  cudaMemset(buf, 1, big);
  for (uint64_t fp = 64ull << 20; fp <= big; fp *= 4) {
    sweep<1>(buf, fp, n, out, 8);
    sweep<4>(buf, fp, n, out, 8);
    sweep<4>(buf, fp, n, out, 2);
  }
  sweep<1>(buf, big, n, out, 8);
  sweep<4>(buf, big, n, out, 8);
  for (uint64_t fp = 64ull << 20; fp <= (16ull << 30); fp *= 16) {
    run_write<0>("random 32-byte sector stores", buf, fp, n);
    run_write<1>("random 24-byte records (3 x 8-byte stores)", buf, fp, n);
  }
  return 0;
}
