// Probe (GPU box): does compute-data compression (cuMemCreate, CU_MEM_ALLOCATION_COMP_GENERIC) cut the DRAM cost of
// writing mostly-zero egocentric crops?  Times zero fills, sparse fills (1 byte in 64 set to 254, like a thin wall in a
// crop) and incompressible fills on a plain and a compressible allocation, with 16-byte stores and with the bulk-copy
// engine (cp.async.bulk.global.shared::cta from a shared page, the way ego_sparse_kernel zeroes its crops).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cprobe profiles/probes/compressible_probe.cu -lcuda && /tmp/cprobe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); \
  printf("%s failed: %s\n", #x, s_); exit(1); } } while (0)
#define CR(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(r_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// mode 0: zeros, 1: zeros with one 254 byte per 64, 2: pseudo-random
__global__ void fill16(uint4* dst, size_t n16, int mode) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (mode == 1 && (i & 3) == 1) v.y = 254u << 8;
    if (mode == 2) { uint32_t h = hash32((uint32_t)i); v = make_uint4(h, hash32(h), hash32(h + 1), hash32(h + 2)); }
    dst[i] = v;
  }
}

// 4 KB zero page in shared memory, bulk stores of 4 KB each (ego_sparse_kernel's zeroing), then byte hits on top
__global__ void fill_bulk(uint8_t* dst, size_t bytes, int hits) {
  __shared__ __align__(128) uint8_t page[4096];
  for (int i = threadIdx.x * 16; i < 4096; i += blockDim.x * 16) *reinterpret_cast<uint4*>(page + i) = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t src = (uint32_t)__cvta_generic_to_shared(page);
  const size_t crop = 15616;                       // 16-byte multiple near 133 x 117
  const size_t ncrops = bytes / crop;
  for (size_t c = blockIdx.x; c < ncrops; c += gridDim.x) {
    uint8_t* d = dst + c * crop;
    if (threadIdx.x == 0) {
      for (size_t o = 0; o < crop; o += 4096) {
        const uint32_t nb = (uint32_t)(crop - o < 4096 ? crop - o : 4096);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + o), "r"(src), "r"(nb) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    for (int k = threadIdx.x; k < hits; k += blockDim.x) d[(hash32((uint32_t)(c * 977 + k)) % (uint32_t)crop)] = 254;
    __syncthreads();
  }
}

__global__ void read16(const uint4* src, size_t n16, uint32_t* out) {
  uint32_t acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = src[i];
    acc += v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *out = acc;
}

static CUdeviceptr vmm_alloc(size_t bytes, bool compress, int dev) {
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = dev;
  if (compress) prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
  size_t gran = 0;
  CK(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  bytes = (bytes + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h;
  CK(cuMemCreate(&h, bytes, &prop, 0));
  CUmemAllocationProp got = {};
  CK(cuMemGetAllocationPropertiesFromHandle(&got, h));
  printf("  allocation: requested compression %d, got %d, granularity %zu\n", compress ? 1 : 0, (int)got.allocFlags.compressionType, gran);
  CUdeviceptr p;
  CK(cuMemAddressReserve(&p, bytes, 0, 0, 0));
  CK(cuMemMap(p, bytes, 0, h, 0));
  CUmemAccessDesc acc = {};
  acc.location = prop.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CK(cuMemSetAccess(p, bytes, &acc, 1));
  return p;
}

template <class F>
static float best_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  CR(cudaEventCreate(&a)); CR(cudaEventCreate(&b));
  float best = 1e9f;
  for (int r = 0; r < reps; ++r) {
    CR(cudaEventRecord(a)); f(); CR(cudaEventRecord(b)); CR(cudaEventSynchronize(b));
    float ms; CR(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  CR(cudaSetDevice(0));
  CR(cudaFree(0));
  CUdevice dev; CK(cuDeviceGet(&dev, 0));
  int sup = 0;
  CK(cuDeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev));
  printf("GENERIC_COMPRESSION_SUPPORTED = %d\n", sup);
  const size_t bytes = (size_t)1 << 31;                 // 2 GiB >> L2
  uint32_t* sink; CR(cudaMalloc(&sink, 4));
  for (int compress = 0; compress <= (sup ? 1 : 0); ++compress) {
    printf("%s allocation\n", compress ? "compressible" : "plain");
    CUdeviceptr p = vmm_alloc(bytes, compress != 0, 0);
    const int grid = 148 * 8;
    for (int mode = 0; mode < 3; ++mode) {
      const float ms = best_ms([&] { fill16<<<grid, 512>>>((uint4*)p, bytes / 16, mode); });
      printf("  fill16 mode %d: %.3f ms  %.0f GB/s\n", mode, ms, bytes / ms / 1e6);
      const float rms = best_ms([&] { read16<<<grid, 512>>>((const uint4*)p, bytes / 16, sink); });
      printf("  read16 after mode %d: %.3f ms  %.0f GB/s\n", mode, rms, bytes / rms / 1e6);
    }
    for (int hits = 0; hits <= 512; hits += 256) {
      const float ms = best_ms([&] { fill_bulk<<<148 * 16, 64>>>((uint8_t*)p, bytes, hits); });
      printf("  bulk zero + %d byte hits per crop: %.3f ms  %.0f GB/s\n", hits, ms, bytes / ms / 1e6);
    }
    CR(cudaDeviceSynchronize());
  }
  return 0;
}
