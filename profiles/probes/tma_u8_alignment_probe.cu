#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const void* tmap_g, const __grid_constant__ CUtensorMap tmap_p, int use_param, int x, int y, int bw, int bh, uint8_t* out) {
  extern __shared__ __align__(128) uint8_t raw[];
  uint8_t* tile = raw + ((128u - (smem_u32(raw) & 127u)) & 127u);
  __shared__ __align__(8) uint64_t mbar_s;
  uint32_t mbar = smem_u32(&mbar_s);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bw * bh) : "memory");
    const void* tm = use_param ? (const void*)&tmap_p : tmap_g;
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(tile)), "l"(tm), "r"(x), "r"(y), "r"(mbar) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(mbar), "r"(0) : "memory");
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
  int bw = argc > 1 ? atoi(argv[1]) : 208, bh = argc > 2 ? atoi(argv[2]) : 16; int W = argc > 3 ? atoi(argv[3]) : 428, H = 416, pitch = argc > 4 ? atoi(argv[4]) : 448; int sx = argc > 5 ? atoi(argv[5]) : 100;
  std::vector<uint8_t> img(pitch * H);
  for (int y = 0; y < H; ++y) for (int x = 0; x < pitch; ++x) img[y * pitch + x] = x < W ? (uint8_t)((x * 7 + y * 13) & 0xff) : 0;
  uint8_t *d_img, *d_out; void* d_tm;
  cudaMalloc(&d_img, img.size()); cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, bw * bh); cudaMalloc(&d_tm, 128);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q);
  auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  alignas(64) CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t strides[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}; cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode result %d\n", (int)r);
  cudaMemcpy(d_tm, &tm, 128, cudaMemcpyHostToDevice);
  for (int use_param = 1; use_param >= 0; --use_param) {
    for (int t = 0; t < 2; ++t) {
      int x = t ? -5 : sx, y = t ? -3 : 50;
      cudaMemset(d_out, 0xAA, bw * bh);
      probe<<<1, 128, bw * bh + 128>>>(d_tm, tm, use_param, x, y, bw, bh, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<uint8_t> out(bw * bh);
      cudaMemcpy(out.data(), d_out, out.size(), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int j = 0; j < bh; ++j) for (int i = 0; i < bw; ++i) {
        int X = x + i, Y = y + j; uint8_t want = (X >= 0 && X < W && Y >= 0 && Y < H) ? img[Y * pitch + X] : 0;
        bad += out[j * bw + i] != want;
      }
      printf("use_param=%d start=(%d,%d) bw=%d: %s, mismatches %d\n", use_param, x, y, bw, cudaGetErrorString(e), bad);
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
