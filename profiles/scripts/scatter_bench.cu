// Micro-benchmark behind the design of ec_bwd_apply_kernel's scatter (profiles/README.md): how fast can E = P*k rows
// of F fp32 gradients be scatter-added into a [P, 2F] table on one B200?
//   A  REDG.128 from registers (16-byte vector atomics, half a warp per 256-byte row)
//   B  TMA bulk reduction: rows staged in shared memory, cp.reduce.async.bulk ... .add.f32 of 256 bytes per row
//   C  A with 8-byte (v2) atomics, D scalar atomics -- for the per-request cost
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scatter_bench scatter_bench.cu ; run: ./scatter_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int F = 64, K = 20, N = 2048, B = 24, P = B * N;
constexpr unsigned FULL = 0xffffffffu;

template <int MODE>
__global__ void __launch_bounds__(256) scatter_regs(const int* __restrict__ idx, const float* __restrict__ src,
                                                    float* __restrict__ dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, l16 = lane & 15;
  for (int p = blockIdx.x * 8 + warp; p < P; p += gridDim.x * 8) {
    const int base = (p / N) * N;
    const int r = lane < K ? base + idx[p * K + lane] : 0;
    const float4 g = *reinterpret_cast<const float4*>(src + (size_t)p * F + 4 * l16);
#pragma unroll
    for (int t = 0; t < K / 2; ++t) {
      const int row = __shfl_sync(FULL, r, 2 * t + half);
      float* d = dst + (size_t)row * (2 * F) + F + 4 * l16;
      const float s = 1.0f + 0.001f * t;
      if (MODE == 0) atomicAdd(reinterpret_cast<float4*>(d), make_float4(g.x * s, g.y * s, g.z * s, g.w * s));
      if (MODE == 1) {
        atomicAdd(reinterpret_cast<float2*>(d), make_float2(g.x * s, g.y * s));
        atomicAdd(reinterpret_cast<float2*>(d + 2), make_float2(g.z * s, g.w * s));
      }
      if (MODE == 2) {
        atomicAdd(d, g.x * s); atomicAdd(d + 1, g.y * s); atomicAdd(d + 2, g.z * s); atomicAdd(d + 3, g.w * s);
      }
    }
  }
}

// E: pure gather -- the read-side twin of A: every point sums its k neighbour rows (256 B each) with 16-byte lanes,
// half a warp per row, all loads of a point in flight at once; one 256-byte store per point.  The L2 gather roof of
// ec_fwd_stats / ec_fwd_apply.
__global__ void __launch_bounds__(256) gather_rows(const int* __restrict__ idx, const float* __restrict__ tab,
                                                   float* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, l16 = lane & 15;
  for (int p = blockIdx.x * 8 + warp; p < P; p += gridDim.x * 8) {
    const int base = (p / N) * N;
    const int r = lane < K ? base + idx[p * K + lane] : 0;
    float4 v[K / 2];
#pragma unroll
    for (int t = 0; t < K / 2; ++t) {
      const int row = __shfl_sync(FULL, r, 2 * t + half);
      v[t] = __ldg(reinterpret_cast<const float4*>(tab + (size_t)row * (2 * F) + F + 4 * l16));
    }
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < K / 2; ++t) { s.x += v[t].x; s.y += v[t].y; s.z += v[t].z; s.w += v[t].w; }
    s.x += __shfl_xor_sync(FULL, s.x, 16); s.y += __shfl_xor_sync(FULL, s.y, 16);
    s.z += __shfl_xor_sync(FULL, s.z, 16); s.w += __shfl_xor_sync(FULL, s.w, 16);
    if (half == 0) *reinterpret_cast<float4*>(out + (size_t)p * F + 4 * l16) = s;
  }
}

// TMA bulk reduce: each warp stages its k rows (k * 256 B) in shared memory, one lane issues k bulk reductions
__global__ void __launch_bounds__(256) scatter_bulk(const int* __restrict__ idx, const float* __restrict__ src,
                                                    float* __restrict__ dst) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, l16 = lane & 15;
  float* stage = reinterpret_cast<float*>(smem_raw) + warp * (K * F);
  for (int p = blockIdx.x * 8 + warp; p < P; p += gridDim.x * 8) {
    const int base = (p / N) * N;
    const int r = lane < K ? base + idx[p * K + lane] : 0;
    const float4 g = *reinterpret_cast<const float4*>(src + (size_t)p * F + 4 * l16);
    // the previous point's bulk reads of this staging area must be complete
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int t = 0; t < K / 2; ++t) {
      const float s = 1.0f + 0.001f * t;
      *reinterpret_cast<float4*>(stage + (2 * t + half) * F + 4 * l16) = make_float4(g.x * s, g.y * s, g.z * s, g.w * s);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane < K) {
      float* d = dst + (size_t)r * (2 * F) + F;
      const unsigned s_addr = (unsigned)__cvta_generic_to_shared(stage + lane * F);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(d), "r"(s_addr),
                   "n"(F * 4)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  std::vector<int> hidx((size_t)P * K);
  srand(1);
  for (auto& v : hidx) v = rand() % N;
  int* idx;
  float *src, *dst;
  cudaMalloc(&idx, hidx.size() * 4);
  cudaMalloc(&src, (size_t)P * F * 4);
  cudaMalloc(&dst, (size_t)P * 2 * F * 4);
  cudaMemcpy(idx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(src, 0, (size_t)P * F * 4);
  cudaMemset(dst, 0, (size_t)P * 2 * F * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaFuncSetAttribute(scatter_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * K * F * 4);
  for (int grid : {148 * 2, 148 * 4, 148 * 8}) {
    for (int mode = 0; mode < 5; ++mode) {
      float best = 1e9f;
      for (int it = 0; it < 6; ++it) {
        cudaEventRecord(e0);
        if (mode == 0) scatter_regs<0><<<grid, 256>>>(idx, src, dst);
        if (mode == 1) scatter_regs<1><<<grid, 256>>>(idx, src, dst);
        if (mode == 2) scatter_regs<2><<<grid, 256>>>(idx, src, dst);
        if (mode == 3) scatter_bulk<<<grid, 256, 8 * K * F * 4>>>(idx, src, dst);
        if (mode == 4) gather_rows<<<grid, 256>>>(idx, dst, src);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
      }
      const char* names[] = {"REDG.128 regs", "REDG.64 regs", "REDG.32 regs", "TMA bulk reduce 256B", "pure gather LDG.128"};
      printf("grid %4d  %-22s %8.1f us  (%.2f TB/s of row bytes)  err=%s\n", grid, names[mode], best * 1e3,
             (double)P * K * F * 4 / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
