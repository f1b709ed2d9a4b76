// tcgen05 (5th-generation tensor core) kernels: the fast path (precision = 1, "BF16x3").
//
// edge_gnn_tc_kernel - one GNN iteration's gather -> edge MLP -> segment max
// (/root/reference/models/gnn.py:338-365) as ONE persistent kernel:
//
//   algebra   e0 @ W1 + b1 = (F @ W1[:C] + b1)[src] + (x_src - x_dst') @ W1[C:]          (hoisting)
//             so the first edge layer is a per-VERTEX table P plus a 3-term per-edge correction,
//             and only the second layer h1 @ W2 (E x D x D) is per-edge tensor work.
//             relu / bias commute with max:  max_e relu(y_e + b2) = relu(max_e y_e + b2),
//             so the epilogue reduces raw accumulators and applies bias + relu once per segment.
//   precision every fp32 operand is split x = hi + lo (two BF16), and hi*hi' + lo*hi' + hi*lo' is
//             accumulated in fp32 in tensor memory: ~2^-16 relative per product, which keeps the
//             whole network within 1e-4 of the fp32 CPU path (precision study).
//   mapping   CTA pairs (cluster 2x1x1) issue cta_group::2 MMAs, M = 256 edges per pair-tile
//             (128 per CTA), N = D output features, split N1 + N2 <= 256 each.  W2 (hi and lo) is
//             RESIDENT in shared memory for the whole kernel, each CTA holding its N/2 half - the
//             pair owns one copy, so no weight traffic after the prologue.
//   pipeline  warps 0-3  epilogue : TMEM -> registers, per-destination max with redux.sync.max.f32
//                                   over match_any segments, bias + relu, atomicMax to HBM
//             warp  4    MMA      : one thread issues tcgen05.mma, commits to mbarriers
//             warps 5-12 producers: gather P[src] (L2 resident), add the coordinate term, relu,
//                                   BF16 split, write the A operand straight into the UMMA
//                                   K-major core-matrix layout (one k-step = one stage)
//
// Operand layout (no swizzle, K-major): 8-row x 16-byte core matrices, 128 contiguous bytes each;
// A stage: core(rg, kc) at rg*256 + kc*128; resident B: core(g, kc) at g*(KP/8)*128 + kc*128.
#include "pg_common.cuh"
#include "pg_umma.cuh"

namespace pg {

int fill_async(float* p, int64_t n, float v, cudaStream_t s);
int fc_fp32_launch(const float* x, int64_t m, int k, const float* w, const float* bias, int n, int act,
                   const float* residual, float* out, int ldo, cudaStream_t s);
int edge_mlp_max_fp32(int mode, const float* features, int c_in, const float* xyz_src, const float* xyz_dst,
                      const int32_t* dst_index, const int32_t* src, const int32_t* dst, int64_t num_edges,
                      int64_t num_src, int64_t num_dst, const float* const* weights, const float* const* biases,
                      const int32_t* dims, int num_layers, float* out, cudaStream_t s);

namespace {
using namespace umma;

constexpr int kStages = 4;
constexpr int kEpiWarps = 4;
constexpr int kProdWarps = 8;
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;  // 416
constexpr int kStageBytes = 8192;                            // A hi (4096) + A lo (4096): 128 rows x 16 k
constexpr int kTileRows = 128;                               // edges per CTA per pair-tile

struct TcEdgeParams {
  const float* P;         // [num_src, ldp] = F @ W1[:C] + b1, zero padded to ldp = KP
  int ldp;
  const float* xyz_src;   // [num_src, 3]
  const float* xyz_dst;   // [num_dst', 3] (already offset)
  const int32_t* dst_index;  // optional indirection dst -> row of xyz_dst
  const int32_t* src;
  const int32_t* dst;
  int64_t num_edges, num_src, num_dst;
  const float* w1x;       // [3, kp] zero padded
  const float* b2;        // [np] zero padded
  int kp, ks;             // padded K, k-steps (kp / 16)
  int n, np, n1, n2;      // real N, padded N, instruction split
  const uint8_t* wimg;    // per rank: [hi part | lo part], each part_bytes
  uint32_t part_bytes;
  uint32_t tmem_cols;
  float* out;             // [num_dst, n], pre-filled with -FLT_MAX
  int* err;
  int64_t num_pair_tiles;
};

// ---- W2 -> resident B image -------------------------------------------------------------------
// w2 is [K, N] row-major fp32.  B operand rows are OUTPUT features (N), K-major.  Rank r of the
// pair holds rows [r*N1/2, (r+1)*N1/2) of instruction 1 followed by [N1 + r*N2/2, ...) of
// instruction 2, as 8-row groups g: core(g, kc) at g*sbo + kc*128, element (row%8)*16 + (k%8)*2.
__global__ void pack_w2_kernel(const float* __restrict__ w2, int k, int n, int kp, int n1, int n2,
                               uint8_t* __restrict__ img, uint32_t part_bytes) {
  const int rows_per_rank = (n1 + n2) / 2;
  const int total = 2 * rows_per_rank * kp;
  const uint32_t sbo = uint32_t(kp / 8) * 128u;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int rank = i / (rows_per_rank * kp);
    const int rem = i - rank * rows_per_rank * kp;
    const int lr = rem / kp;       // local B row
    const int kk = rem - lr * kp;  // k
    int feature;
    if (lr < n1 / 2) feature = rank * (n1 / 2) + lr;
    else feature = n1 + rank * (n2 / 2) + (lr - n1 / 2);
    const float v = (feature < n && kk < k) ? w2[int64_t(kk) * n + feature] : 0.0f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const uint32_t off = uint32_t(lr / 8) * sbo + uint32_t(kk / 8) * 128u + uint32_t(lr % 8) * 16u + uint32_t(kk % 8) * 2u;
    uint8_t* base = img + size_t(rank) * 2 * part_bytes;
    *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(base + part_bytes + off) = lo;
  }
}

__global__ void pad_rows_kernel(const float* __restrict__ in, int rows, int cols, int ld_out, float* __restrict__ out) {
  const int total = rows * ld_out;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / ld_out, c = i - r * ld_out;
    out[i] = c < cols ? in[int64_t(r) * cols + c] : 0.0f;
  }
}

__device__ __forceinline__ float redux_max_f32(float v, uint32_t mask) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "r"(mask));
  return r;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) edge_gnn_tc_kernel(TcEdgeParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve-up (all offsets identical in both CTAs of the pair: cta_group::2 requires it)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_bres = smem;                                   // 2 * part_bytes
  uint8_t* s_a = s_bres + 2 * size_t(p.part_bytes);         // kStages * kStageBytes
  float* s_w1x = reinterpret_cast<float*>(s_a + kStages * kStageBytes);   // 3 * kp
  float* s_b2 = s_w1x + 3 * p.kp;                                         // np
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b2 + p.np + (p.np & 1));
  uint64_t* bar_full = bars;                 // [kStages]  (used in the leader CTA)
  uint64_t* bar_empty = bars + kStages;      // [kStages]
  uint64_t* bar_tmem_full = bars + 2 * kStages;
  uint64_t* bar_tmem_empty = bars + 2 * kStages + 1;   // (leader)
  uint64_t* bar_wres = bars + 2 * kStages + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t cluster_id = blockIdx.x >> 1;
  const int64_t num_clusters = gridDim.x >> 1;

  // ---- prologue ------------------------------------------------------------------------------
  for (int i = threadIdx.x; i < 3 * p.kp; i += kThreads) s_w1x[i] = p.w1x[i];
  for (int i = threadIdx.x; i < p.np; i += kThreads) s_b2[i] = p.b2[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bar_full[i], 2 * kProdWarps);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(bar_tmem_full, 1);
    mbar_init(bar_tmem_empty, 2 * kEpiWarps);
    mbar_init(bar_wres, 1);
    fence_barrier_init();
  }
  if (warp == kEpiWarps) {
    tmem_alloc<2>(s_tmem, p.tmem_cols);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == kEpiWarps) {
    // =================================== MMA warp =============================================
    if (lane == 0) {
      // resident weights: one bulk copy per part into this CTA's shared memory
      mbar_arrive_expect_tx(bar_wres, 2 * p.part_bytes);
      const uint8_t* g = p.wimg + size_t(rank) * 2 * p.part_bytes;
      bulk_g2s(s_bres, g, p.part_bytes, bar_wres);
      bulk_g2s(s_bres + p.part_bytes, g + p.part_bytes, p.part_bytes, bar_wres);
      mbar_wait(bar_wres, 0);
    }
    __syncwarp();
    cluster_sync();   // both CTAs' weights are resident before the leader issues any MMA   [sync A]
    if (rank == 0 && lane == 0) {
      const uint32_t idesc1 = make_idesc_bf16(256, p.n1);
      const uint32_t idesc2 = make_idesc_bf16(256, p.n2 > 0 ? p.n2 : 16);
      const uint32_t sbo_b = uint32_t(p.kp / 8) * 128u;
      const uint32_t b_hi = smem_u32(s_bres), b_lo = b_hi + p.part_bytes;
      const uint32_t b2_off = uint32_t(p.n1 / 16) * sbo_b;   // first row group of instruction 2
      uint32_t it = 0;
      uint32_t tile_iter = 0;
      for (int64_t tile = cluster_id; tile < p.num_pair_tiles; tile += num_clusters, ++tile_iter) {
        mbar_wait_cluster(bar_tmem_empty, (tile_iter & 1) ^ 1);
        tc_fence_after();
        for (int s = 0; s < p.ks; ++s, ++it) {
          const uint32_t stage = it % kStages;
          mbar_wait_cluster(&bar_full[stage], (it / kStages) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(s_a + stage * kStageBytes), a_lo = a_hi + kStageBytes / 2;
          const uint64_t da_hi = make_smem_desc(a_hi, 128, 256);
          const uint64_t da_lo = make_smem_desc(a_lo, 128, 256);
          const uint32_t koff = uint32_t(s) * 256u;   // two K-adjacent cores per k-step
          {
            const uint64_t db_hi = make_smem_desc(b_hi + koff, 128, sbo_b);
            const uint64_t db_lo = make_smem_desc(b_lo + koff, 128, sbo_b);
            mma_bf16<2>(tmem, da_hi, db_hi, idesc1, s > 0);
            mma_bf16<2>(tmem, da_lo, db_hi, idesc1, true);
            mma_bf16<2>(tmem, da_hi, db_lo, idesc1, true);
          }
          if (p.n2 > 0) {
            const uint64_t db_hi = make_smem_desc(b_hi + b2_off + koff, 128, sbo_b);
            const uint64_t db_lo = make_smem_desc(b_lo + b2_off + koff, 128, sbo_b);
            mma_bf16<2>(tmem + p.n1, da_hi, db_hi, idesc2, s > 0);
            mma_bf16<2>(tmem + p.n1, da_lo, db_hi, idesc2, true);
            mma_bf16<2>(tmem + p.n1, da_hi, db_lo, idesc2, true);
          }
          mma_commit_2cta(&bar_empty[stage], 0x3);    // frees this A stage in both CTAs
        }
        mma_commit_2cta(bar_tmem_full, 0x3);          // accumulators of this tile are complete
      }
    }
    __syncwarp();
  } else if (warp < kEpiWarps) {
    // =================================== epilogue warps =======================================
    cluster_sync();   // [sync A]
    uint32_t tile_iter = 0;
    for (int64_t tile = cluster_id; tile < p.num_pair_tiles; tile += num_clusters, ++tile_iter) {
      const int64_t row = tile * 256 + int64_t(rank) * kTileRows + warp * 32 + lane;
      int d = -1;
      if (row < p.num_edges) {
        d = p.dst[row];
        if (d < 0 || d >= p.num_dst) { *p.err = 1; d = -1; }
      }
      const uint32_t my_mask = __match_any_sync(0xffffffffu, d);
      const int my_rank = __popc(my_mask & ((1u << lane) - 1u));
      const int my_count = __popc(my_mask);
      mbar_wait(bar_tmem_full, tile_iter & 1);
      tc_fence_after();
      const uint32_t taddr = tmem + (uint32_t(warp * 32) << 16);
      for (int c0 = 0; c0 < p.np; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        // segments of this warp: iterate distinct match masks (warp-uniform loop)
        uint32_t remaining = 0xffffffffu;
        while (remaining) {
          const int leader = __ffs(remaining) - 1;
          const uint32_t m = __shfl_sync(0xffffffffu, my_mask, leader);
          const int seg_dst = __shfl_sync(0xffffffffu, d, leader);
          remaining &= ~m;
          if (seg_dst < 0) continue;
          if (my_mask == m) {
            float* orow = p.out + int64_t(seg_dst) * p.n;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float r = redux_max_f32(__uint_as_float(v[j]), m);
              const int c = c0 + j;
              if ((j % my_count) == my_rank && c < p.n) {
                const float val = fmaxf(r + s_b2[c], 0.0f);
                atomicMax(reinterpret_cast<int*>(orow + c), __float_as_int(val));
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar_tmem_empty, 0);
    }
  } else {
    // =================================== producer warps =======================================
    cluster_sync();   // [sync A]
    const int pt = threadIdx.x - (kEpiWarps + 1) * 32;   // 0..255
    const int r = pt & 127;                              // tile row
    const int kc = pt >> 7;                              // which 8-wide K chunk of the k-step
    const uint32_t a_off = uint32_t(r >> 3) * 256u + uint32_t(kc) * 128u + uint32_t(r & 7) * 16u;
    uint32_t it = 0;
    for (int64_t tile = cluster_id; tile < p.num_pair_tiles; tile += num_clusters) {
      const int64_t row = tile * 256 + int64_t(rank) * kTileRows + r;
      bool valid = row < p.num_edges;
      int sidx = 0, didx = 0;
      if (valid) {
        sidx = p.src[row];
        didx = p.dst[row];
        if (sidx < 0 || sidx >= p.num_src || didx < 0 || didx >= p.num_dst) { *p.err = 1; valid = false; sidx = 0; didx = 0; }
      }
      const int64_t drow = p.dst_index ? int64_t(p.dst_index[didx]) : int64_t(didx);
      const float rx = p.xyz_src[int64_t(sidx) * 3 + 0] - p.xyz_dst[drow * 3 + 0];
      const float ry = p.xyz_src[int64_t(sidx) * 3 + 1] - p.xyz_dst[drow * 3 + 1];
      const float rz = p.xyz_src[int64_t(sidx) * 3 + 2] - p.xyz_dst[drow * 3 + 2];
      const float* prow = p.P + int64_t(sidx) * p.ldp + kc * 8;
      float4 n0 = *reinterpret_cast<const float4*>(prow);
      float4 n1 = *reinterpret_cast<const float4*>(prow + 4);
      for (int s = 0; s < p.ks; ++s, ++it) {
        const float4 c0 = n0, c1 = n1;
        if (s + 1 < p.ks) {   // prefetch the next k-step's slice of P[src]
          n0 = *reinterpret_cast<const float4*>(prow + (s + 1) * 16);
          n1 = *reinterpret_cast<const float4*>(prow + (s + 1) * 16 + 4);
        }
        const int k0 = s * 16 + kc * 8;
        const float* wx = s_w1x + k0;
        const float* wy = s_w1x + p.kp + k0;
        const float* wz = s_w1x + 2 * p.kp + k0;
        float h[8];
        const float pv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = fmaf(rx, wx[j], pv[j]);
          t = fmaf(ry, wy[j], t);
          t = fmaf(rz, wz[j], t);
          h[j] = valid ? fmaxf(t, 0.0f) : 0.0f;
        }
        uint4 hi, lo;
        split_bf16x2(h[0], h[1], &hi.x, &lo.x);
        split_bf16x2(h[2], h[3], &hi.y, &lo.y);
        split_bf16x2(h[4], h[5], &hi.z, &lo.z);
        split_bf16x2(h[6], h[7], &hi.w, &lo.w);
        const uint32_t stage = it % kStages;
        mbar_wait(&bar_empty[stage], ((it / kStages) & 1) ^ 1);
        uint8_t* st = s_a + stage * kStageBytes + a_off;
        *reinterpret_cast<uint4*>(st) = hi;
        *reinterpret_cast<uint4*>(st + kStageBytes / 2) = lo;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&bar_full[stage], 0);
      }
    }
  }

  // ---- teardown ------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == kEpiWarps) tmem_dealloc<2>(tmem, p.tmem_cols);
}

size_t tc_edge_smem_bytes(int kp, int np) {
  const size_t part = size_t(np / 16) * size_t(kp / 8) * 128;
  return 1024 + 2 * part + kStages * kStageBytes + (3 * kp + np + (np & 1)) * sizeof(float) + (2 * kStages + 3) * 8 + 16;
}

}  // namespace

extern "C" int pg_tc_available(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

int fc_tc_bf16x3(const float* x, int64_t m, int k, const float* w, const float* bias, int n, int act,
                 const float* residual, float* out, cudaStream_t s) {
  // The per-vertex layers are <1% of the frame's FLOPs; they run on the fp32 FFMA kernel (exact
  // fp32, no split needed).  The tensor-core budget goes to the per-edge GEMMs.
  return fc_fp32_launch(x, m, k, w, bias, n, act, residual, out, n, s);
}

int edge_mlp_max_tc(int mode, const float* features, int c_in, const float* xyz_src, const float* xyz_dst,
                    const int32_t* dst_index, const int32_t* src, const int32_t* dst, int64_t num_edges,
                    int64_t num_src, int64_t num_dst, const float* const* weights, const float* const* biases,
                    const int32_t* dims, int num_layers, float* out, cudaStream_t s) {
  bool fits = (mode == PG_EDGE_GNN) && num_layers == 2 && pg_tc_available();
  int kp = 0, np = 0, n1 = 0, n2 = 0;
  if (fits) {
    kp = (dims[1] + 15) / 16 * 16;
    np = (dims[2] + 15) / 16 * 16;
    if (np <= 256) { n1 = np; n2 = 0; }
    else { n1 = ((np / 2) + 15) / 16 * 16; n2 = np - n1; }
    fits = np <= 512 && kp / 16 > kStages && tc_edge_smem_bytes(kp, np) <= 227 * 1024 && dims[1] >= 8;
  }
  if (!fits || num_edges == 0)
    return edge_mlp_max_fp32(mode, features, c_in, xyz_src, xyz_dst, dst_index, src, dst, num_edges, num_src,
                             num_dst, weights, biases, dims, num_layers, out, s);
  PG_REQUIRE(dims[0] == c_in + 3, "dims[0]=%d must equal feature channels + 3 = %d", dims[0], c_in + 3);
  const int d1 = dims[1], n = dims[2];
  TcEdgeParams p{};
  Temp t_p, t_w1x, t_b2, t_img, t_err;
  // P = F @ W1[:C] + b1, row stride kp, pad columns zero
  PG_CUDA_OK(t_p.alloc(sizeof(float) * num_src * kp, s));
  if (int rc = fc_fp32_launch(features, num_src, c_in, weights[0], biases[0], d1, 0, nullptr, t_p.as<float>(), kp, s))
    return rc;
  PG_CUDA_OK(t_w1x.alloc(sizeof(float) * 3 * kp, s));
  pad_rows_kernel<<<4, 256, 0, s>>>(weights[0] + int64_t(c_in) * d1, 3, d1, kp, t_w1x.as<float>());
  PG_LAUNCH_CHECK();
  PG_CUDA_OK(t_b2.alloc(sizeof(float) * np, s));
  pad_rows_kernel<<<2, 256, 0, s>>>(biases[1], 1, n, np, t_b2.as<float>());
  PG_LAUNCH_CHECK();
  const uint32_t part = uint32_t(np / 16) * uint32_t(kp / 8) * 128u;
  PG_CUDA_OK(t_img.alloc(size_t(4) * part, s));
  pack_w2_kernel<<<std::min(num_sms(), 64), 256, 0, s>>>(weights[1], d1, n, kp, n1, n2, t_img.as<uint8_t>(), part);
  PG_LAUNCH_CHECK();
  PG_CUDA_OK(t_err.alloc(sizeof(int), s));
  PG_CUDA_OK(cudaMemsetAsync(t_err.ptr, 0, sizeof(int), s));
  if (int rc = fill_async(out, num_dst * n, -FLT_MAX, s)) return rc;

  p.P = t_p.as<float>();
  p.ldp = kp;
  p.xyz_src = xyz_src;
  p.xyz_dst = xyz_dst;
  p.dst_index = dst_index;
  p.src = src;
  p.dst = dst;
  p.num_edges = num_edges;
  p.num_src = num_src;
  p.num_dst = num_dst;
  p.w1x = t_w1x.as<float>();
  p.b2 = t_b2.as<float>();
  p.kp = kp;
  p.ks = kp / 16;
  p.n = n;
  p.np = np;
  p.n1 = n1;
  p.n2 = n2;
  p.wimg = t_img.as<uint8_t>();
  p.part_bytes = part;
  uint32_t cols = 32;
  while (cols < uint32_t(np)) cols <<= 1;
  p.tmem_cols = cols;
  p.out = out;
  p.err = t_err.as<int>();
  p.num_pair_tiles = ceil_div(num_edges, 2 * kTileRows);
  const size_t smem = tc_edge_smem_bytes(kp, np);
  PG_CUDA_OK(cudaFuncSetAttribute(edge_gnn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int clusters = int(std::min<int64_t>(p.num_pair_tiles, num_sms() / 2));
  edge_gnn_tc_kernel<<<2 * clusters, kThreads, smem, s>>>(p);
  PG_LAUNCH_CHECK();
  int h = 0;
  PG_CUDA_OK(cudaMemcpyAsync(&h, t_err.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
  PG_CUDA_OK(cudaStreamSynchronize(s));
  PG_REQUIRE(h == 0, "edge index out of range (src in [0,%lld), dst in [0,%lld))", (long long)num_src,
             (long long)num_dst);
  return PG_OK;
}

}  // namespace pg
