// Probe: tcgen05.mma with the A operand in TENSOR MEMORY (the "TS" form) - the layout question the attention kernel's
// P-in-TMEM design rests on.  O[128,64] = P[128,K] V[K,64], P written by each thread (lane = row) with tcgen05.st as
// packed bf16 pairs (column j of the lane = elements k = 2j, 2j+1), V in shared memory as the MN-major B operand
// ([key][dim] rows of 128 B, 128-byte swizzle).  Prints the max error against a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../../opensearch-neural-pre-train_b200/csrc -o tmem_a_operand tmem_a_operand.cu
#include "common.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace splade;

__device__ __forceinline__ uint64_t desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1024 >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

constexpr int KEYS = 64;

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* P, const __nv_bfloat16* V, float* O) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  // V[key][dim] -> swizzled rows
  for (int i = tid; i < KEYS * 64; i += 128) {
    const int key = i >> 6, dim = i & 63;
    const uint32_t off = key * 128 + (((dim >> 3) ^ (key & 7)) << 4) + (dim & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(smem + off) = V[key * 64 + dim];
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = tmem + ((warp * 32u) << 16);
  // P row `tid`: 64 keys -> 32 packed columns at TMEM columns 64..95
  for (int half = 0; half < 2; ++half) {
    uint32_t v[16];
    for (int j = 0; j < 16; ++j) {
      const uint32_t lo = __bfloat16_as_ushort(P[tid * KEYS + half * 32 + 2 * j]);
      const uint32_t hi = __bfloat16_as_ushort(P[tid * KEYS + half * 32 + 2 * j + 1]);
      v[j] = lo | (hi << 16);
    }
    tmem_st_x16(lane_base + 64u + 16u * half, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) {
    tcgen05_fence_after();
    const uint32_t idesc = make_idesc_bf16_f32(128, 64) | (1u << 16);  // B is MN-major
    for (int ks = 0; ks < KEYS / 16; ++ks)
      umma_ts(tmem, tmem + 64u + 8u * ks, desc_sw128_mn(smem_u32(smem) + ks * 2048), idesc, ks != 0 ? 1u : 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(lane_base + 32u * c, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) O[tid * 64 + 32 * c + j] = __uint_as_float(r[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

int main() {
  std::vector<__nv_bfloat16> P(128 * KEYS), V(KEYS * 64);
  std::vector<float> Pf(128 * KEYS), Vf(KEYS * 64), ref(128 * 64, 0.f), out(128 * 64);
  srand(1);
  for (size_t i = 0; i < P.size(); ++i) { P[i] = __float2bfloat16(rand() / float(RAND_MAX)); Pf[i] = __bfloat162float(P[i]); }
  for (size_t i = 0; i < V.size(); ++i) { V[i] = __float2bfloat16(rand() / float(RAND_MAX) - 0.5f); Vf[i] = __bfloat162float(V[i]); }
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < KEYS; ++k)
      for (int d = 0; d < 64; ++d) ref[m * 64 + d] += Pf[m * KEYS + k] * Vf[k * 64 + d];
  __nv_bfloat16 *dP, *dV;
  float* dO;
  cudaMalloc(&dP, P.size() * 2); cudaMalloc(&dV, V.size() * 2); cudaMalloc(&dO, out.size() * 4);
  cudaMemcpy(dP, P.data(), P.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dV, V.data(), V.size() * 2, cudaMemcpyHostToDevice);
  probe<<<1, 128, 16384 + 1024>>>(dP, dV, dO);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
  double err = 0;
  for (size_t i = 0; i < out.size(); ++i) err = fmax(err, fabs(out[i] - ref[i]));
  printf("cuda: %s  max |O - ref| = %g  (O[0]=%g ref[0]=%g, O[64*5+3]=%g ref=%g)\n", cudaGetErrorString(e), err, out[0], ref[0],
         out[64 * 5 + 3], ref[64 * 5 + 3]);
  return err < 1e-2 ? 0 : 1;
}
