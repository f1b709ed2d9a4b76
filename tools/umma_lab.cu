// umma_lab: stand-alone hardware checks of the tcgen05 building blocks used by pg_tc.cu.
// Each test runs in its own process (a faulting kernel poisons the CUDA context):
//   umma_lab 1cta <swap>      D[128xN] = A[128xK] * B[NxK]^T, cta_group::1, no-swizzle K-major operands
//   umma_lab 2cta <swap>      D[256xN], cta_group::2 (A split by rows, B split by N across the CTA pair)
//   umma_lab redux            redux.sync.max.f32 with full and partial member masks
// <swap> = 0: descriptor LBO = K-adjacent core stride, SBO = row-group stride (CUTLASS reading);
//          1: the two swapped.  Exactly one of them must reproduce the CPU result.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../point-gnn_b200/csrc/pg_umma.cuh"

using namespace pg::umma;

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e = (x);                                                          \
    if (e != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                    \
    }                                                                             \
  } while (0)

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return uint16_t(u >> 16);  // exact for the small integers used here
}

// K-major no-swizzle image of a [rows x k] matrix: core (rg, kc) at rg*rgs + kc*kcs
static void pack(const std::vector<float>& m, int rows, int k, int kcs, int rgs, std::vector<uint16_t>& img) {
  img.assign(size_t(rows / 8) * rgs / 2, 0);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < k; ++c) {
      size_t off = size_t(r / 8) * rgs + size_t(c / 8) * kcs + (r % 8) * 16 + (c % 8) * 2;
      img[off / 2] = f2bf(m[size_t(r) * k + c]);
    }
}

template <int kCtaGroup>
__global__ void __launch_bounds__(128) gemm_kernel(const uint16_t* __restrict__ a_img, const uint16_t* __restrict__ b_img,
                                                   int a_bytes, int b_bytes, int n, int k, int kcs, int lbo, int sbo,
                                                   float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const uint32_t rank = kCtaGroup == 2 ? cluster_ctarank() : 0;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_bytes + 1023) / 1024) * 1024;
  const uint16_t* ga = a_img + size_t(rank) * a_bytes / 2;
  const uint16_t* gb = b_img + size_t(rank) * b_bytes / 2;
  for (int i = threadIdx.x; i < a_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sa)[i] = reinterpret_cast<const uint32_t*>(ga)[i];
  for (int i = threadIdx.x; i < b_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sb)[i] = reinterpret_cast<const uint32_t*>(gb)[i];
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc<kCtaGroup>(&tmem_base, 256);
    tmem_relinquish<kCtaGroup>();
  }
  tc_fence_before();
  __syncthreads();
  if (kCtaGroup == 2) cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(128 * kCtaGroup, n);
    for (int s = 0; s < k / 16; ++s) {
      const uint64_t da = make_smem_desc(smem_u32(sa) + s * 2 * kcs, lbo, sbo);
      const uint64_t db = make_smem_desc(smem_u32(sb) + s * 2 * kcs, lbo, sbo);
      mma_bf16<kCtaGroup>(tmem, da, db, idesc, s > 0);
    }
    if (kCtaGroup == 1)
      mma_commit_1cta(&bar);
    else
      mma_commit_2cta(&bar, 0x3);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < n; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(size_t(rank) * 128 + row) * n + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (kCtaGroup == 2) cluster_sync();
  if (threadIdx.x < 32) tmem_dealloc<kCtaGroup>(tmem, 256);
}

static int run_gemm(int cta_group, int swap) {
  const int k = 64, n = 160, m = 128 * cta_group;
  const int kcs = 128, rgs = (k / 8) * 128;
  std::vector<float> a(size_t(m) * k), b(size_t(n) * k);
  srand(1);
  for (auto& x : a) x = float(rand() % 7 - 3);
  for (auto& x : b) x = float(rand() % 7 - 3);
  // per-CTA images: A rows [128*rank, +128); B rows: cta_group 1 -> all n; 2 -> [n/2*rank, +n/2)
  const int a_rows = 128, b_rows = n / cta_group;
  std::vector<uint16_t> a_img, b_img, tmp;
  for (int r = 0; r < cta_group; ++r) {
    std::vector<float> sub(a.begin() + size_t(r) * a_rows * k, a.begin() + size_t(r + 1) * a_rows * k);
    pack(sub, a_rows, k, kcs, rgs, tmp);
    a_img.insert(a_img.end(), tmp.begin(), tmp.end());
    std::vector<float> subb(b.begin() + size_t(r) * b_rows * k, b.begin() + size_t(r + 1) * b_rows * k);
    pack(subb, b_rows, k, kcs, rgs, tmp);
    b_img.insert(b_img.end(), tmp.begin(), tmp.end());
  }
  const int a_bytes = a_rows / 8 * rgs, b_bytes = b_rows / 8 * rgs;
  uint16_t *da, *db;
  float* dout;
  CK(cudaMalloc(&da, a_img.size() * 2));
  CK(cudaMalloc(&db, b_img.size() * 2));
  CK(cudaMalloc(&dout, size_t(m) * n * 4));
  CK(cudaMemcpy(da, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, size_t(m) * n * 4));
  const int lbo = swap ? rgs : kcs, sbo = swap ? kcs : rgs;
  const size_t smem = ((a_bytes + 1023) / 1024) * 1024 + b_bytes + 1024;
  if (cta_group == 1) {
    CK(cudaFuncSetAttribute(gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    gemm_kernel<1><<<1, 128, smem>>>(da, db, a_bytes, b_bytes, n, k, kcs, lbo, sbo, dout);
  } else {
    CK(cudaFuncSetAttribute(gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, gemm_kernel<2>, (const uint16_t*)da, (const uint16_t*)db, a_bytes, b_bytes, n, k, kcs,
                          lbo, sbo, dout));
  }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> out(size_t(m) * n);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  int bad = 0;
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      float ref = 0;
      for (int c = 0; c < k; ++c) ref += a[size_t(i) * k + c] * b[size_t(j) * k + c];
      const double e = fabs(double(ref) - out[size_t(i) * n + j]);
      if (!(e <= 1e-3)) ++bad;
      if (e > maxerr || e != e) maxerr = e;
    }
  printf("gemm cta_group=%d swap=%d: maxerr=%g bad=%d/%d -> %s\n", cta_group, swap, maxerr, bad, m * n,
         bad == 0 ? "MATCH" : "mismatch");
  return bad == 0 ? 0 : 1;
}

__global__ void redux_kernel(float* out, long long* cycles) {
  const int lane = threadIdx.x & 31;
  float v = float((lane * 37) % 32) - 7.5f;
  float r_full, r_part = -1.f;
  asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(r_full) : "f"(v), "r"(0xffffffffu));
  const unsigned mask = lane < 11 ? 0x7ffu : 0xfffff800u;  // two segments
  asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(r_part) : "f"(v), "r"(mask));
  out[threadIdx.x] = r_full;
  out[32 + threadIdx.x] = r_part;
  // throughput: 256 dependent-free redux per warp
  float acc = 0.f;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 256; ++i) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(r) : "f"(v + float(i)), "r"(0xffffffffu));
    acc += r;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  out[64 + threadIdx.x] = acc;
}

static int run_redux() {
  float* d;
  long long* c;
  CK(cudaMalloc(&d, 96 * 4));
  CK(cudaMalloc(&c, 8));
  redux_kernel<<<1, 32>>>(d, c);
  CK(cudaDeviceSynchronize());
  float h[96];
  long long hc;
  CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost));
  float v[32], mfull = -1e9f, m0 = -1e9f, m1 = -1e9f;
  for (int l = 0; l < 32; ++l) {
    v[l] = float((l * 37) % 32) - 7.5f;
    mfull = fmaxf(mfull, v[l]);
    if (l < 11) m0 = fmaxf(m0, v[l]); else m1 = fmaxf(m1, v[l]);
  }
  int bad = 0;
  for (int l = 0; l < 32; ++l) {
    if (h[l] != mfull) ++bad;
    if (h[32 + l] != (l < 11 ? m0 : m1)) ++bad;
  }
  printf("redux.sync.max.f32: bad=%d  (256 redux in %lld cycles = %.2f cyc/op, one warp)\n", bad, hc, hc / 256.0);
  return bad;
}

// ---- raw tensor-core rate of the production instruction mix -------------------------------------
// cta_group::2, M = 256, no-swizzle K-major operands resident in shared memory, kp/16 k-steps per
// "tile", three products per k-step (the BF16x3 pattern) for each of the instructions N1 (+ N2).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) mma_rate_kernel(int n1, int n2, int kp, int tiles,
                                                                                  int commit_every_stage,
                                                                                  long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[3];
  __shared__ uint32_t tmem_base;
  const uint32_t rank = cluster_ctarank();
  const int np = n1 + n2;
  const uint32_t part = uint32_t(np / 16) * uint32_t(kp / 8) * 128u;
  uint8_t* s_b = smem;
  uint8_t* s_a = smem + 2 * part;
  for (int i = threadIdx.x; i < int(2 * part + 3 * 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 3; ++i) mbar_init(&bar2[i], 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc<2>(&tmem_base, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc1 = make_idesc_bf16(256, n1), idesc2 = make_idesc_bf16(256, n2 > 0 ? n2 : 16);
    const uint32_t sbo_b = uint32_t(kp / 8) * 128u;
    const uint32_t b_hi = smem_u32(s_b), b_lo = b_hi + part, b2_off = uint32_t(n1 / 16) * sbo_b;
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      for (int s = 0; s < kp / 16; ++s) {
        const uint32_t a_hi = smem_u32(s_a + (s % 3) * 8192), a_lo = a_hi + 4096;
        const uint64_t da_hi = make_smem_desc(a_hi, 128, 256), da_lo = make_smem_desc(a_lo, 128, 256);
        const uint32_t koff = uint32_t(s) * 256u;
        const uint64_t db_hi = make_smem_desc(b_hi + koff, 128, sbo_b), db_lo = make_smem_desc(b_lo + koff, 128, sbo_b);
        if (commit_every_stage == 2 && n2 > 0) {   // interleave the two accumulators
          const uint64_t eb_hi = make_smem_desc(b_hi + b2_off + koff, 128, sbo_b), eb_lo = make_smem_desc(b_lo + b2_off + koff, 128, sbo_b);
          mma_bf16<2>(tmem, da_hi, db_hi, idesc1, s > 0);
          mma_bf16<2>(tmem + n1, da_hi, eb_hi, idesc2, s > 0);
          mma_bf16<2>(tmem, da_lo, db_hi, idesc1, true);
          mma_bf16<2>(tmem + n1, da_lo, eb_hi, idesc2, true);
          mma_bf16<2>(tmem, da_hi, db_lo, idesc1, true);
          mma_bf16<2>(tmem + n1, da_hi, eb_lo, idesc2, true);
        } else {
        mma_bf16<2>(tmem, da_hi, db_hi, idesc1, s > 0);
        mma_bf16<2>(tmem, da_lo, db_hi, idesc1, true);
        mma_bf16<2>(tmem, da_hi, db_lo, idesc1, true);
        if (n2 > 0) {
          const uint64_t eb_hi = make_smem_desc(b_hi + b2_off + koff, 128, sbo_b), eb_lo = make_smem_desc(b_lo + b2_off + koff, 128, sbo_b);
          mma_bf16<2>(tmem + n1, da_hi, eb_hi, idesc2, s > 0);
          mma_bf16<2>(tmem + n1, da_lo, eb_hi, idesc2, true);
          mma_bf16<2>(tmem + n1, da_hi, eb_lo, idesc2, true);
        }
        }
        if (commit_every_stage == 1) {
          mma_commit_2cta(&bar, 0x1);
          mbar_wait(&bar, phase);
          phase ^= 1;
        }
        if (commit_every_stage == 3) mma_commit_2cta(&bar2[s % 3], 0x3);   // commit, nobody waits
      }
      if (commit_every_stage != 1) {
        mma_commit_2cta(&bar, 0x1);
        mbar_wait(&bar, phase);
        phase ^= 1;
      }
    }
    cycles[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) tmem_dealloc<2>(tmem, 512);
}

static int run_rate(int n1, int n2, int kp, int per_stage) {
  long long* c;
  CK(cudaMalloc(&c, 8));
  const int np = n1 + n2;
  const size_t smem = size_t(2) * (np / 16) * (kp / 8) * 128 + 3 * 8192;
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int tiles = 200;
  mma_rate_kernel<<<2, 128, smem>>>(n1, n2, kp, tiles, per_stage, c);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  long long h;
  CK(cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost));
  const double per_tile = double(h) / tiles;
  const double ideal = double(kp / 16) * 3.0 * (256.0 * np / 512.0);
  printf("mma rate n1=%d n2=%d kp=%d commit/wait per %s: %.0f cycles per 256-row tile (ideal %.0f, ratio %.2f)\n", n1, n2, kp,
         per_stage ? "stage" : "tile", per_tile, ideal, per_tile / ideal);
  return 0;
}


// ---- TMEM drain rate: how fast can warps read a 128-lane x 256-column fp32 accumulator? -----------------
//   umma_lab drain <warps 4|8> <variant>
// variant 0: x32 loads, one at a time (ld, wait, 32 max)   1: x32, next load issued before the maxes (as pg_tc.cu)
//         2: x64 double buffered   3: x128 one at a time   4: x32, three loads in flight   5: x16 one at a time
__device__ __forceinline__ void ld64(uint32_t a, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
        "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
        "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
        "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
        "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(a)
      : "memory");
}
template <int N>
__device__ __forceinline__ float vmax(const uint32_t (&v)[N], float m) {
#pragma unroll
  for (int j = 0; j < N; ++j) m = fmaxf(m, __uint_as_float(v[j]));
  return m;
}

__global__ void __launch_bounds__(256) drain_kernel(int variant, int reps, float* out, long long* cycles) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  if (warp == 0) {
    tmem_alloc<1>(&tmem_base, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  const int quarter = warp & 3, part = warp >> 2, parts = nwarps / 4;
  const int cols = 256 / parts;                               // columns this warp drains per repetition
  const uint32_t t0 = tmem + (uint32_t(quarter * 32) << 16) + uint32_t(part * cols);
  float m = -1e30f;
  __syncthreads();
  const long long c0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (variant == 0) {
      for (int c = 0; c < cols; c += 32) {
        uint32_t v[32];
        tmem_ld32(t0 + c, v);
        tmem_ld_wait();
        m = vmax(v, m);
      }
    } else if (variant == 1) {
      uint32_t va[32], vb[32];
      tmem_ld32(t0, va);
      for (int c = 0; c < cols; c += 64) {
        tmem_ld_wait();
        tmem_ld32(t0 + c + 32, vb);
        m = vmax(va, m);
        tmem_ld_wait();
        if (c + 64 < cols) tmem_ld32(t0 + c + 64, va);
        m = vmax(vb, m);
      }
    } else if (variant == 2) {
      uint32_t va[64], vb[64];
      ld64(t0, va);
      for (int c = 0; c < cols; c += 128) {
        tmem_ld_wait();
        if (c + 64 < cols) ld64(t0 + c + 64, vb);
        m = vmax(va, m);
        tmem_ld_wait();
        if (c + 128 < cols) ld64(t0 + c + 128, va);
        if (c + 64 < cols) m = vmax(vb, m);
      }
    } else if (variant == 3) {
      for (int c = 0; c < cols; c += 128) {
        uint32_t va[64], vb[64];
        ld64(t0 + c, va);
        if (c + 64 < cols) ld64(t0 + c + 64, vb);
        tmem_ld_wait();
        m = vmax(va, m);
        if (c + 64 < cols) m = vmax(vb, m);
      }
    } else if (variant == 4) {
      uint32_t va[32], vb[32], vc[32], vd[32];
      for (int c = 0; c < cols; c += 128) {
        tmem_ld32(t0 + c, va);
        if (c + 32 < cols) tmem_ld32(t0 + c + 32, vb);
        if (c + 64 < cols) tmem_ld32(t0 + c + 64, vc);
        if (c + 96 < cols) tmem_ld32(t0 + c + 96, vd);
        tmem_ld_wait();
        m = vmax(va, m);
        if (c + 32 < cols) m = vmax(vb, m);
        if (c + 64 < cols) m = vmax(vc, m);
        if (c + 96 < cols) m = vmax(vd, m);
      }
    } else {
      for (int c = 0; c < cols; c += 16) {
        uint32_t v[16];
        tmem_ld16(t0 + c, v);
        tmem_ld_wait();
        m = vmax(v, m);
      }
    }
  }
  const long long c1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = m;
  if (threadIdx.x == 0) cycles[blockIdx.x] = c1 - c0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem, 512);
}

static int run_drain(int warps, int variant) {
  float* out;
  long long* c;
  const int blocks = 148, reps = 200;
  CK(cudaMalloc(&out, sizeof(float) * blocks * 256));
  CK(cudaMalloc(&c, 8 * blocks));
  drain_kernel<<<blocks, warps * 32>>>(variant, reps, out, c);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  drain_kernel<<<blocks, warps * 32>>>(variant, reps, out, c);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(blocks);
  CK(cudaMemcpy(h.data(), c, 8 * blocks, cudaMemcpyDeviceToHost));
  double avg = 0;
  for (long long v : h) avg += double(v) / blocks;
  const double per = avg / reps;
  printf("drain warps=%d variant=%d: %.0f cycles per 128 KB (128 lanes x 256 columns) = %.1f B/clk per SM\n", warps, variant,
         per, 131072.0 / per);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  if (!strcmp(argv[1], "1cta")) return run_gemm(1, argc > 2 ? atoi(argv[2]) : 0);
  if (!strcmp(argv[1], "2cta")) return run_gemm(2, argc > 2 ? atoi(argv[2]) : 0);
  if (!strcmp(argv[1], "redux")) return run_redux();
  if (!strcmp(argv[1], "drain")) return run_drain(atoi(argv[2]), atoi(argv[3]));
  if (!strcmp(argv[1], "rate")) return run_rate(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]));
  return 1;
}
