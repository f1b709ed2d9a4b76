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
#include <atomic>
#include <cstdlib>
#include <vector>

#include "pg_common.cuh"
#include "pg_umma.cuh"

extern "C" int pg_tc_available(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

namespace pg {
static std::atomic<long long> g_tc_launches[2];   // [0] fused edge kernel, [1] dense-layer kernel
}

// tcgen05 kernel launches so far: which = 0 fused edge (segment-max) kernel, 1 dense-layer kernel
extern "C" int64_t pg_tc_launch_count(int32_t which) {
  return (which == 0 || which == 1) ? pg::g_tc_launches[which].load(std::memory_order_relaxed) : -1;
}

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

constexpr int kStages = 3;
constexpr int kEpiWarps = 8;     // warps 0-7: warp w drains TMEM lane quarter (w & 3), column chunks of parity (w >> 2)
constexpr int kMmaWarp = 8;
constexpr int kProdWarps = 8;    // warps 9-16
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;  // 544
constexpr int kStageBytes = 8192;                            // A hi (4096) + A lo (4096): 128 rows x 16 k
constexpr int kTileRows = 128;                               // rows (edges) per CTA per pair-tile
constexpr int kScratchStride = 36;                           // floats per scratch column: 32 rows + pad, 16 B aligned
constexpr int kScratchFloats = 16 * kScratchStride;          // per epilogue warp: 16 columns x 32 rows

enum { PROD_GNN = 0, PROD_ROWS = 1, PROD_POOL = 2 };
constexpr int kPoolC1 = 32, kPoolC2 = 64, kPoolC3 = 128;   // point MLP widths the pooling producer is built for
constexpr int kPoolWFloats = 4 * kPoolC1 + kPoolC1 + kPoolC1 * kPoolC2 + kPoolC2 + kPoolC2 * kPoolC3 + kPoolC3;
enum { EPI_SEGMAX = 0, EPI_STORE = 1 };

struct TcParams {
  // A-operand producer
  const float* P;         // GNN: [num_src, ldp] = F @ W1[:C] + b1 (zero padded to ldp = kp);  ROWS: x [num_rows, ldp]
  int ldp;
  int k_real;             // ROWS: true K (multiple of 4)
  const float* xyz_src;   // [num_src, 3]
  const float* xyz_dst;   // [*, 3] (already offset)
  const int32_t* dst_index;  // optional indirection dst -> row of xyz_dst
  const int32_t* src;
  const int32_t* dst;
  int64_t num_rows;       // edges (GNN) or matrix rows (ROWS)
  int64_t num_src, num_dst;
  const float* w1x;       // [3, kp] zero padded (GNN)
  const float* pool_w;    // POOL: packed [W1 4x32 | b1 | W2 32x64 | b2 | W3 64x128 | b3] (fp32)
  const float* pool_feat; // POOL: point features [num_src, 1]
  // GEMM shape
  const float* bias;      // [np] zero padded
  int kp, ks;             // padded K, k-steps (kp / 16)
  int n, np, n1, n2;      // real N, padded N, instruction split
  const uint8_t* wimg;    // per rank: [hi part | lo part], each part_bytes
  uint32_t part_bytes;
  uint32_t tmem_cols;
  // epilogue
  float* out;             // SEGMAX: [num_dst, n] pre-filled with -FLT_MAX;  STORE: [num_rows, ldo]
  int ldo;
  int act;                // STORE: 0 linear, 1 relu
  const float* residual;  // STORE: optional [num_rows, n]
  int* err;
  int64_t num_pair_tiles;
  unsigned long long* trace;   // optional (PG_TC_TRACE): [role 0..7][slot 0..127][3] globaltimer ns, cluster 0 only
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PG_TRACE(role, slot, k)                                                                   \
  do {                                                                                            \
    if (p.trace != nullptr && cluster_id == 0 && (slot) < 128) p.trace[((role) * 128 + (slot)) * 3 + (k)] = gtime(); \
  } while (0)

// ---- W [K, N] -> resident B image ---------------------------------------------------------------
// B operand rows are OUTPUT features (N), K-major.  Rank r of the pair holds rows
// [r*N1/2, (r+1)*N1/2) of instruction 1 followed by [N1 + r*N2/2, ...) of instruction 2, as 8-row
// groups g: core(g, kc) at g*sbo + kc*128, element (row%8)*16 + (k%8)*2.  hi / lo = BF16 split.
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

struct SmemMap {
  uint8_t* bres;
  uint8_t* a;
  float* w1x;
  float* scratch;
  uint64_t* bar_full;      // [kStages]   (leader)
  uint64_t* bar_empty;     // [kStages]
  uint64_t* bar_tmem_full;
  uint64_t* bar_i1_empty;  // [2]         (leader)
  uint64_t* bar_i2_empty;  //             (leader)
  uint64_t* bar_wres;
  uint32_t* tmem;
};

__host__ __device__ inline size_t smem_layout(uint8_t* base, int kp, int np, uint32_t part_bytes, SmemMap* m,
                                              int prod = PROD_GNN) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return o; };
  const size_t o_bres = take(2 * size_t(part_bytes));
  const size_t o_a = take(size_t(kStages) * kStageBytes);
  const size_t o_w1x = take((prod == PROD_POOL ? size_t(kPoolWFloats) : size_t(3) * kp) * sizeof(float));
  const size_t o_scr = take(size_t(kEpiWarps) * kScratchFloats * sizeof(float));
  const size_t o_bar = take((2 * kStages + 5) * sizeof(uint64_t));
  const size_t o_tmem = take(16);
  if (m != nullptr) {
    m->bres = base + o_bres;
    m->a = base + o_a;
    m->w1x = reinterpret_cast<float*>(base + o_w1x);
    m->scratch = reinterpret_cast<float*>(base + o_scr);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + o_bar);
    m->bar_full = bars;
    m->bar_empty = bars + kStages;
    m->bar_tmem_full = bars + 2 * kStages;
    m->bar_i1_empty = bars + 2 * kStages + 1;
    m->bar_i2_empty = bars + 2 * kStages + 3;
    m->bar_wres = bars + 2 * kStages + 4;
    m->tmem = reinterpret_cast<uint32_t*>(base + o_tmem);
  }
  return off;
}

// One 16-column chunk of the accumulator, segment-max flavour.  Thread = TMEM lane = tile row.
// The chunk is transposed through a warp-private shared-memory scratch so that each lane then owns
// one COLUMN of one 16-row half (h = lane / 16): four 128-bit loads + 15 max.  `d` is the lane's own
// destination; when no destination boundary falls inside either half (warp-uniform `slow` == false,
// the common case: segments are ~150-250 edges long) the half's max is flushed with one atomic,
// otherwise the rows are walked one by one and a flush happens at every boundary.
struct SegState {
  int d;          // destination of this lane's row (-1 = row beyond the edge list)
  int cur0;       // destination of the first row of this lane's 16-row half
  int nb;         // number of destination boundaries strictly inside this lane's half
  int b;          // position (1..15) of the first such boundary
  int d_b;        // destination that starts at that boundary
  bool pair;      // both halves of the warp lie in ONE destination (no boundary anywhere in the warp)
};

__device__ __forceinline__ void seg_flush(const TcParams& p, int cur, int c, bool col_ok, float m, float bias) {
  if (cur >= 0 && col_ok && m > -FLT_MAX && p.act != 99)   // act == 99: PG_TC_NOFLUSH experiment (results invalid)
    atomicMax(reinterpret_cast<int*>(p.out + int64_t(cur) * p.n + c), __float_as_int(fmaxf(m + bias, 0.0f)));
}

// Second half of a chunk (the first half = tcgen05.ld + 16 stores into the scratch, see the caller).
// Lane = (column jj, 16-row half h).  Destinations are non-decreasing along the rows, so a half
// contains 0 boundaries (plain max), 1 boundary (two masked maxima, branch free) or - only for
// destinations with fewer than 16 edges - several (sequential walk).  Every partial max is flushed
// with one atomicMax; when the whole warp lies in one destination the two halves are combined first.
__device__ __forceinline__ void epi_reduce_segmax(const TcParams& p, int c_out, const float* scratch,
                                                   const SegState& st, int lane, float bias, int64_t warp_row0) {
  const int jj = lane & 15, h = lane >> 4;
  const float* col = scratch + jj * kScratchStride + 16 * h;
  const float4 q0 = reinterpret_cast<const float4*>(col)[0];
  const float4 q1 = reinterpret_cast<const float4*>(col)[1];
  const float4 q2 = reinterpret_cast<const float4*>(col)[2];
  const float4 q3 = reinterpret_cast<const float4*>(col)[3];
  const int c = c_out + jj;
  const bool col_ok = c < p.n;
  if (st.pair) {   // warp uniform
    float m = fmaxf(fmaxf(fmaxf(fmaxf(q0.x, q0.y), fmaxf(q0.z, q0.w)), fmaxf(fmaxf(q1.x, q1.y), fmaxf(q1.z, q1.w))),
                    fmaxf(fmaxf(fmaxf(q2.x, q2.y), fmaxf(q2.z, q2.w)), fmaxf(fmaxf(q3.x, q3.y), fmaxf(q3.z, q3.w))));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
    if (h == 0) seg_flush(p, st.cur0, c, col_ok, m, bias);
    return;
  }
  const float val[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
  if (st.nb <= 1) {
    float lo = -FLT_MAX, hi = -FLT_MAX;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const bool first = st.nb == 0 || i < st.b;
      lo = fmaxf(lo, first ? val[i] : -FLT_MAX);
      hi = fmaxf(hi, first ? -FLT_MAX : val[i]);
    }
    seg_flush(p, st.cur0, c, col_ok, lo, bias);
    seg_flush(p, st.d_b, c, col_ok, hi, bias);   // hi stays -FLT_MAX (no-op) when nb == 0
    return;
  }
  // several short segments inside 16 rows: walk them (destinations via shared memory is not
  // available here, so they are re-read from global memory - rare path)
  int cur = st.cur0;
  float m = -FLT_MAX;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int64_t r = warp_row0 + 16 * h + i;
    int di = r < p.num_rows ? p.dst[r] : -1;
    if (di >= p.num_dst) di = -1;
    if (di != cur) {
      seg_flush(p, cur, c, col_ok, m, bias);
      cur = di;
      m = -FLT_MAX;
    }
    m = fmaxf(m, val[i]);
  }
  seg_flush(p, cur, c, col_ok, m, bias);
}

// Plain GEMM epilogue: out[row, c] = act(acc + bias[c]) (+ residual[row, c]).
__device__ __forceinline__ void epi_chunk_store(const TcParams& p, uint32_t taddr, int c_out, int64_t row,
                                                 bool row_ok) {
  uint32_t v[16];
  tmem_ld16(taddr, v);
  tmem_ld_wait();
  if (!row_ok) return;
  float* o = p.out + row * p.ldo + c_out;
  const float* res = p.residual ? p.residual + row * p.n + c_out : nullptr;
  const bool vec = ((p.n & 3) == 0) && ((p.ldo & 3) == 0);
#pragma unroll
  for (int j4 = 0; j4 < 16; j4 += 4) {
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c_out + j4 + j;
      float t = __uint_as_float(v[j4 + j]) + __ldg(p.bias + c);   // bias is padded to np; warp-uniform address
      if (p.act == 1) t = fmaxf(t, 0.0f);
      r[j] = t;
    }
    if (vec) {
      if (c_out + j4 + 3 < p.n) {
        if (res) {
          const float4 q = *reinterpret_cast<const float4*>(res + j4);
          r[0] += q.x; r[1] += q.y; r[2] += q.z; r[3] += q.w;
        }
        *reinterpret_cast<float4*>(o + j4) = make_float4(r[0], r[1], r[2], r[3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c_out + j4 + j < p.n) o[j4 + j] = r[j] + (res ? res[j4 + j] : 0.0f);
    }
  }
}

// One accumulator section (instruction 1 or 2) for this warp: chunks par, par+2, ... of 16 columns.
// The tcgen05.ld of chunk k+1 is in flight while chunk k is reduced out of the scratch.
template <int kEpi, int kMaxChunks>
__device__ __forceinline__ void epi_section(const TcParams& p, uint32_t tbase, int col0, int ncols, int par,
                                             float* scratch, const SegState& st, int lane, int64_t row, bool row_ok,
                                             int64_t warp_row0, const float (&bias)[kMaxChunks], int trace_slot = -1) {
  const int64_t cluster_id = (trace_slot >= 0) ? 0 : 1;   // PG_TRACE only fires for cluster 0
  if (kEpi == EPI_STORE) {
#pragma unroll
    for (int k = 0; k < kMaxChunks; ++k) {
      const int ci = par + 2 * k;
      if (ci * 16 < ncols) epi_chunk_store(p, tbase + ci * 16, col0 + ci * 16, row, row_ok);
    }
    return;
  }
  uint32_t v[16];
  if (par * 16 < ncols) tmem_ld16(tbase + par * 16, v);
#pragma unroll
  for (int k = 0; k < kMaxChunks; ++k) {
    const int ci = par + 2 * k;
    if (ci * 16 < ncols) {
      if (k < 2) PG_TRACE(6 + k, trace_slot, 0);
      tmem_ld_wait();
      __syncwarp();   // the previous chunk's scratch reads are done
#pragma unroll
      for (int j = 0; j < 16; ++j) scratch[j * kScratchStride + lane] = __uint_as_float(v[j]);
      if ((ci + 2) * 16 < ncols) tmem_ld16(tbase + (ci + 2) * 16, v);
      __syncwarp();
      if (k < 2) PG_TRACE(6 + k, trace_slot, 1);
      epi_reduce_segmax(p, col0 + ci * 16, scratch, st, lane, bias[k], warp_row0);
      if (k < 2) PG_TRACE(6 + k, trace_slot, 2);
    }
  }
}

template <int kProd, int kEpi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) row_gemm_tc_kernel(TcParams p) {
  // No-swizzle operands, bulk copies and mbarriers only need 16-byte alignment; the carve-up is
  // identical in both CTAs of the pair (cta_group::2 addresses the peer's operands by offset).
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemMap sm;
  smem_layout(smem_raw, p.kp, p.np, p.part_bytes, &sm, kProd);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t cluster_id = blockIdx.x >> 1;
  const int64_t num_clusters = gridDim.x >> 1;

  // ---- prologue ------------------------------------------------------------------------------
  if (kProd == PROD_GNN)
    for (int i = threadIdx.x; i < 3 * p.kp; i += kThreads) sm.w1x[i] = p.w1x[i];
  if (kProd == PROD_POOL)
    for (int i = threadIdx.x; i < kPoolWFloats; i += kThreads) sm.w1x[i] = p.pool_w[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&sm.bar_full[i], 2 * kProdWarps);
      mbar_init(&sm.bar_empty[i], 1);
    }
    mbar_init(sm.bar_tmem_full, 1);
    mbar_init(&sm.bar_i1_empty[0], 2 * kEpiWarps);
    mbar_init(&sm.bar_i1_empty[1], 2 * kEpiWarps);
    mbar_init(sm.bar_i2_empty, 2 * kEpiWarps);
    mbar_init(sm.bar_wres, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc<2>(sm.tmem, p.tmem_cols);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *sm.tmem;
  // accumulator columns: instruction 1 is double buffered (a / b), instruction 2 single
  const uint32_t col_i1[2] = {0u, uint32_t(p.n1 + p.n2)};
  const uint32_t col_i2 = uint32_t(p.n1);

  if (warp == kMmaWarp) {
    // =================================== MMA warp =============================================
    if (lane == 0) {
      // resident weights: one bulk copy per part into this CTA's shared memory
      mbar_arrive_expect_tx(sm.bar_wres, 2 * p.part_bytes);
      const uint8_t* g = p.wimg + size_t(rank) * 2 * p.part_bytes;
      bulk_g2s(sm.bres, g, p.part_bytes, sm.bar_wres);
      bulk_g2s(sm.bres + p.part_bytes, g + p.part_bytes, p.part_bytes, sm.bar_wres);
      mbar_wait(sm.bar_wres, 0);
    }
    __syncwarp();
    cluster_sync();   // both CTAs' weights are resident before the leader issues any MMA   [sync A]
    if (rank == 0 && lane == 0) {
      const uint32_t idesc1 = make_idesc_bf16(256, p.n1);
      const uint32_t idesc2 = make_idesc_bf16(256, p.n2 > 0 ? p.n2 : 16);
      const uint32_t sbo_b = uint32_t(p.kp / 8) * 128u;
      // descriptors differ only in the 14-bit start-address field: build them once, then add offsets
      const uint64_t a_hi0 = make_smem_desc(smem_u32(sm.a), 128, 256);
      const uint64_t b_hi0 = make_smem_desc(smem_u32(sm.bres), 128, sbo_b);
      const uint64_t b_lo0 = make_smem_desc(smem_u32(sm.bres) + p.part_bytes, 128, sbo_b);
      const uint64_t b2_off = uint64_t((uint32_t(p.n1 / 16) * sbo_b) >> 4);   // first row group of instruction 2
      uint32_t it = 0, stage = 0, phase = 0;
      uint32_t tile_iter = 0;
      for (int64_t tile = cluster_id; tile < p.num_pair_tiles; tile += num_clusters, ++tile_iter) {
        const uint32_t buf = tile_iter & 1u;
        mbar_wait(&sm.bar_i1_empty[buf], ((tile_iter >> 1) & 1u) ^ 1u);
        if (p.n2 > 0) mbar_wait(sm.bar_i2_empty, (tile_iter & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d1 = tmem + col_i1[buf], d2 = tmem + col_i2;
        uint64_t kb = 0;   // (k-step * 256 bytes) >> 4: two K-adjacent cores per k-step
        bool ready = false;   // full[stage] already observed complete by the probe of the previous k-step
        for (int s = 0; s < p.ks; ++s, ++it, kb += 16) {
          PG_TRACE(0, it, 0);
          if (!ready) mbar_wait(&sm.bar_full[stage], phase);
          PG_TRACE(0, it, 1);
          tc_fence_after();
          // probe the NEXT stage's barrier now: its ~90-cycle latency hides behind the MMA issue below
          const uint32_t nstage = (stage + 1 == kStages) ? 0u : stage + 1;
          const uint32_t nphase = (stage + 1 == kStages) ? (phase ^ 1u) : phase;
          ready = mbar_try_wait(&sm.bar_full[nstage], nphase);
          const uint64_t da_hi = a_hi0 + uint64_t(stage * (kStageBytes >> 4));
          const uint64_t da_lo = da_hi + uint64_t((kStageBytes / 2) >> 4);
          const uint64_t db_hi = b_hi0 + kb, db_lo = b_lo0 + kb;
          mma_bf16<2>(d1, da_hi, db_hi, idesc1, s > 0);
          mma_bf16<2>(d1, da_lo, db_hi, idesc1, true);
          mma_bf16<2>(d1, da_hi, db_lo, idesc1, true);
          if (p.n2 > 0) {
            mma_bf16<2>(d2, da_hi, db_hi + b2_off, idesc2, s > 0);
            mma_bf16<2>(d2, da_lo, db_hi + b2_off, idesc2, true);
            mma_bf16<2>(d2, da_hi, db_lo + b2_off, idesc2, true);
          }
          mma_commit_2cta(&sm.bar_empty[stage], 0x3);    // frees this A stage in both CTAs
          PG_TRACE(0, it, 2);
          stage = nstage;
          phase = nphase;
        }
        mma_commit_2cta(sm.bar_tmem_full, 0x3);          // accumulators of this tile are complete
      }
    }
    __syncwarp();
  } else if (warp < kEpiWarps) {
    // =================================== epilogue warps =======================================
    cluster_sync();   // [sync A]
    const int quarter = warp & 3, par = warp >> 2;
    float* scratch = sm.scratch + warp * kScratchFloats;
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    constexpr int kMaxChunks = 8;   // chunks of one section handled by one warp: <= 256 / 16 / 2
    // this lane's column in each of its chunks never changes: keep the biases in registers
    float bias1[kMaxChunks], bias2[kMaxChunks];
    if (kEpi == EPI_SEGMAX) {
#pragma unroll
      for (int k = 0; k < kMaxChunks; ++k) {
        const int c1 = (par + 2 * k) * 16 + (lane & 15);
        const int c2 = p.n1 + c1;
        bias1[k] = ((par + 2 * k) * 16 < p.n1 && c1 < p.n) ? __ldg(p.bias + c1) : 0.0f;
        bias2[k] = ((par + 2 * k) * 16 < p.n2 && c2 < p.n) ? __ldg(p.bias + c2) : 0.0f;
      }
    }
    uint32_t tile_iter = 0;
    for (int64_t tile = cluster_id; tile < p.num_pair_tiles; tile += num_clusters, ++tile_iter) {
      const uint32_t buf = tile_iter & 1u;
      const int64_t row = tile * 256 + int64_t(rank) * kTileRows + quarter * 32 + lane;
      const bool row_ok = row < p.num_rows;
      const int64_t warp_row0 = tile * 256 + int64_t(rank) * kTileRows + quarter * 32;
      SegState st{-1, -1, 0, 0, -1, false};
      if (kEpi == EPI_SEGMAX) {
        if (row_ok) {
          st.d = p.dst[row];
          if (st.d < 0 || st.d >= p.num_dst) { *p.err = 1; st.d = -1; }
        }
        const int prev = __shfl_up_sync(0xffffffffu, st.d, 1);
        const uint32_t bits = __ballot_sync(0xffffffffu, (lane & 15) != 0 && prev != st.d);
        const uint32_t mine = (bits >> (lane & 16)) & 0xffffu;    // boundaries inside this lane's half
        st.nb = __popc(mine);
        st.b = mine ? __ffs(mine) - 1 : 16;
        st.cur0 = __shfl_sync(0xffffffffu, st.d, lane & 16);
        st.d_b = __shfl_sync(0xffffffffu, st.d, (lane & 16) + (st.b & 15));
        const int d16 = __shfl_sync(0xffffffffu, st.d, 16);
        st.pair = bits == 0 && __shfl_sync(0xffffffffu, st.d, 0) == d16;
      }
      if (warp == 0 && lane == 0) PG_TRACE(3 + rank, tile_iter, 0);
      mbar_wait(sm.bar_tmem_full, tile_iter & 1u);
      if (warp == 0 && lane == 0) PG_TRACE(3 + rank, tile_iter, 1);
      tc_fence_after();
      // ---- instruction-2 columns first: they are single buffered, free them as early as possible.
      // TMEM is handed back with a relaxed arrive (ordered by the tcgen05 fences): a release arrive
      // would wait for the reductions still in flight (~2 us, measured with PG_TC_TRACE).
      if (p.n2 > 0) {
        epi_section<kEpi, kMaxChunks>(p, tmem + lane_base + col_i2, p.n1, p.n2, par, scratch, st, lane, row, row_ok,
                                       warp_row0, bias2,
                                       (warp == 0 && lane == 0 && rank == 0 && cluster_id == 0) ? int(tile_iter) : -1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(sm.bar_i2_empty, 0);
        if (warp == 0 && lane == 0 && rank == 0) PG_TRACE(5, tile_iter, 0);
      }
      if (warp == 0 && lane == 0 && rank == 0) PG_TRACE(5, tile_iter, 1);
      epi_section<kEpi, kMaxChunks>(p, tmem + lane_base + col_i1[buf], 0, p.n1, par, scratch, st, lane, row, row_ok,
                                     warp_row0, bias1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(&sm.bar_i1_empty[buf], 0);
      if (warp == 0 && lane == 0 && rank == 0) PG_TRACE(5, tile_iter, 2);
      if (warp == 0 && lane == 0) PG_TRACE(3 + rank, tile_iter, 2);
    }
  } else {
    // =================================== producer warps =======================================
    cluster_sync();   // [sync A]
    const int pt = threadIdx.x - (kEpiWarps + 1) * 32;   // 0..255
    const int r = pt & 127;                              // tile row
    const int kc = pt >> 7;                              // which 8-wide K chunk of the k-step
    const uint32_t a_off = uint32_t(r >> 3) * 256u + uint32_t(kc) * 128u + uint32_t(r & 7) * 16u;
    uint32_t it = 0, stage = 0, phase = 0;
    for (int64_t tile = cluster_id; tile < p.num_pair_tiles; tile += num_clusters) {
      const int64_t row = tile * 256 + int64_t(rank) * kTileRows + r;
      bool valid = row < p.num_rows;
      float rx = 0.f, ry = 0.f, rz = 0.f;
      const float* prow = p.P;
      int sidx = 0, didx = 0;
      if (kProd != PROD_ROWS) {
        if (valid) {
          sidx = p.src[row];
          didx = p.dst[row];
          if (sidx < 0 || sidx >= p.num_src || didx < 0 || didx >= p.num_dst) { *p.err = 1; valid = false; sidx = 0; didx = 0; }
        }
        const int64_t drow = p.dst_index ? int64_t(p.dst_index[didx]) : int64_t(didx);
        rx = p.xyz_src[int64_t(sidx) * 3 + 0] - p.xyz_dst[drow * 3 + 0];
        ry = p.xyz_src[int64_t(sidx) * 3 + 1] - p.xyz_dst[drow * 3 + 1];
        rz = p.xyz_src[int64_t(sidx) * 3 + 2] - p.xyz_dst[drow * 3 + 2];
        if (kProd == PROD_GNN) prow = p.P + int64_t(sidx) * p.ldp + kc * 8;
      } else {
        prow = p.P + (valid ? row : 0) * int64_t(p.ldp) + kc * 8;
      }
      // hand one k-step (8 values of this thread's row) to the tensor core
      auto publish = [&](const float (&h)[8]) {
        uint4 hi, lo;
        split_bf16x2(h[0], h[1], &hi.x, &lo.x);
        split_bf16x2(h[2], h[3], &hi.y, &lo.y);
        split_bf16x2(h[4], h[5], &hi.z, &lo.z);
        split_bf16x2(h[6], h[7], &hi.w, &lo.w);
        if (pt == 0) PG_TRACE(1 + rank, it, 0);
        mbar_wait(&sm.bar_empty[stage], phase ^ 1u);
        if (pt == 0) PG_TRACE(1 + rank, it, 1);
        uint8_t* st = sm.a + stage * kStageBytes + a_off;
        *reinterpret_cast<uint4*>(st) = hi;
        *reinterpret_cast<uint4*>(st + kStageBytes / 2) = lo;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&sm.bar_full[stage], 0);
        if (pt == 0) PG_TRACE(1 + rank, it, 2);
        ++it;
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      };
      if (kProd == PROD_POOL) {
        // per-edge point MLP 4 -> 32 -> 64 (registers), then 64 -> 128 eight outputs per k-step;
        // weights are warp-uniform shared-memory broadcasts.  gnn.py:264-274 (layers 1-3 of 4).
        const float* w1 = sm.w1x;
        const float* b1 = w1 + 4 * kPoolC1;
        const float* w2 = b1 + kPoolC1;
        const float* b2 = w2 + kPoolC1 * kPoolC2;
        const float* w3 = b2 + kPoolC2;
        const float* b3 = w3 + kPoolC2 * kPoolC3;
        const float f0 = p.pool_feat[sidx];
        float h2[kPoolC2];
#pragma unroll
        for (int j = 0; j < kPoolC2; ++j) h2[j] = b2[j];
#pragma unroll 4
        for (int i = 0; i < kPoolC1; ++i) {
          float t = b1[i];
          t = fmaf(f0, w1[i], t);
          t = fmaf(rx, w1[kPoolC1 + i], t);
          t = fmaf(ry, w1[2 * kPoolC1 + i], t);
          t = fmaf(rz, w1[3 * kPoolC1 + i], t);
          t = fmaxf(t, 0.0f);
          const float4* wr = reinterpret_cast<const float4*>(w2 + i * kPoolC2);
#pragma unroll
          for (int j4 = 0; j4 < kPoolC2 / 4; ++j4) {
            const float4 q = wr[j4];
            h2[4 * j4 + 0] = fmaf(t, q.x, h2[4 * j4 + 0]);
            h2[4 * j4 + 1] = fmaf(t, q.y, h2[4 * j4 + 1]);
            h2[4 * j4 + 2] = fmaf(t, q.z, h2[4 * j4 + 2]);
            h2[4 * j4 + 3] = fmaf(t, q.w, h2[4 * j4 + 3]);
          }
        }
#pragma unroll
        for (int j = 0; j < kPoolC2; ++j) h2[j] = fmaxf(h2[j], 0.0f);
        for (int s = 0; s < p.ks; ++s) {
          const int k0 = s * 16 + kc * 8;
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = b3[k0 + j];
#pragma unroll
          for (int i = 0; i < kPoolC2; ++i) {
            const float4 q0 = *reinterpret_cast<const float4*>(w3 + i * kPoolC3 + k0);
            const float4 q1 = *reinterpret_cast<const float4*>(w3 + i * kPoolC3 + k0 + 4);
            acc[0] = fmaf(h2[i], q0.x, acc[0]);
            acc[1] = fmaf(h2[i], q0.y, acc[1]);
            acc[2] = fmaf(h2[i], q0.z, acc[2]);
            acc[3] = fmaf(h2[i], q0.w, acc[3]);
            acc[4] = fmaf(h2[i], q1.x, acc[4]);
            acc[5] = fmaf(h2[i], q1.y, acc[5]);
            acc[6] = fmaf(h2[i], q1.z, acc[6]);
            acc[7] = fmaf(h2[i], q1.w, acc[7]);
          }
          float h[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = fmaxf(acc[j], 0.0f);   // rows past the end are dropped by the epilogue
          publish(h);
        }
      } else {
        auto load_chunk = [&](int s, float4& q0, float4& q1) {
          if (kProd == PROD_GNN) {
            q0 = *reinterpret_cast<const float4*>(prow + s * 16);
            q1 = *reinterpret_cast<const float4*>(prow + s * 16 + 4);
          } else {
            const int k0 = s * 16 + kc * 8;
            q0 = (k0 + 4 <= p.k_real) ? *reinterpret_cast<const float4*>(prow + s * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
            q1 = (k0 + 8 <= p.k_real) ? *reinterpret_cast<const float4*>(prow + s * 16 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        float4 n0, n1;
        load_chunk(0, n0, n1);
        for (int s = 0; s < p.ks; ++s) {
          const float4 c0 = n0, c1 = n1;
          if (s + 1 < p.ks) load_chunk(s + 1, n0, n1);   // prefetch the next k-step's slice
          float h[8];
          const float pv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
          if (kProd == PROD_GNN) {
            const int k0 = s * 16 + kc * 8;
            const float* wx = sm.w1x + k0;
            const float* wy = sm.w1x + p.kp + k0;
            const float* wz = sm.w1x + 2 * p.kp + k0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float t = fmaf(rx, wx[j], pv[j]);
              t = fmaf(ry, wy[j], t);
              t = fmaf(rz, wz[j], t);
              h[j] = fmaxf(t, 0.0f);   // rows past the end are dropped by the epilogue
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = pv[j];
          }
          publish(h);
        }
      }
    }
  }

  // ---- teardown ------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == kMmaWarp) tmem_dealloc<2>(tmem, p.tmem_cols);
}

size_t tc_smem_bytes(int kp, int np, int prod = PROD_GNN) {
  const uint32_t part = uint32_t(np / 16) * uint32_t(kp / 8) * 128u;
  return smem_layout(nullptr, kp, np, part, nullptr, prod);
}

struct TcShape {
  int kp, np, n1, n2;
  uint32_t part, tmem_cols;
  bool ok;
};

TcShape tc_shape(int k, int n) {
  TcShape t{};
  t.kp = (k + 15) / 16 * 16;
  t.np = (n + 15) / 16 * 16;
  // instruction 1 is double buffered in TMEM, instruction 2 single: 2*n1 + n2 <= 512 columns.
  // Make n1 as large as that allows so that the single-buffered part (whose drain the MMA of the
  // next tile has to wait for) is as small as possible: np = 304 -> 208 + 96.
  if (t.np <= 256) { t.n1 = t.np; t.n2 = 0; }
  else { t.n1 = std::min(256, (512 - t.np) / 16 * 16); t.n2 = t.np - t.n1; }
  t.part = uint32_t(t.np / 16) * uint32_t(t.kp / 8) * 128u;
  const int cols_needed = 2 * t.n1 + t.n2;
  uint32_t cols = 32;
  while (cols < uint32_t(cols_needed)) cols <<= 1;
  t.tmem_cols = cols;
  t.ok = pg_tc_available() && cols_needed <= 512 && t.n1 <= 256 && t.n2 <= 256 && t.kp / 16 > kStages && n >= 8 &&
         tc_smem_bytes(t.kp, t.np) <= 227 * 1024;
  return t;
}

template <int kProd, int kEpi>
int launch_row_gemm(TcParams& p, const TcShape& t, const float* w, int k, int n, const float* bias, Temp& t_img,
                    Temp& t_bias, cudaStream_t s) {
  PG_CUDA_OK(t_bias.alloc(sizeof(float) * t.np, s));
  pad_rows_kernel<<<2, 256, 0, s>>>(bias, 1, n, t.np, t_bias.as<float>());
  PG_LAUNCH_CHECK();
  PG_CUDA_OK(t_img.alloc(size_t(4) * t.part, s));
  pack_w2_kernel<<<std::min(num_sms(), 64), 256, 0, s>>>(w, k, n, t.kp, t.n1, t.n2, t_img.as<uint8_t>(), t.part);
  PG_LAUNCH_CHECK();
  p.bias = t_bias.as<float>();
  p.kp = t.kp;
  p.ks = t.kp / 16;
  p.n = n;
  p.np = t.np;
  p.n1 = t.n1;
  p.n2 = t.n2;
  p.wimg = t_img.as<uint8_t>();
  p.part_bytes = t.part;
  p.tmem_cols = t.tmem_cols;
  p.num_pair_tiles = ceil_div(p.num_rows, 2 * kTileRows);
  const size_t smem = tc_smem_bytes(t.kp, t.np, kProd);
  PG_REQUIRE(smem <= 227 * 1024, "tcgen05 kernel needs %zu B of shared memory", smem);
  PG_CUDA_OK(cudaFuncSetAttribute(row_gemm_tc_kernel<kProd, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int clusters = int(std::min<int64_t>(p.num_pair_tiles, num_sms() / 2));
  const char* trace_path = (kProd == PROD_GNN) ? getenv("PG_TC_TRACE") : nullptr;   // debugging aid
  if (kEpi == EPI_SEGMAX && getenv("PG_TC_NOFLUSH") != nullptr) p.act = 99;       // timing experiment only
  Temp t_trace;
  const size_t trace_words = size_t(8) * 128 * 3;
  if (trace_path != nullptr) {
    PG_CUDA_OK(t_trace.alloc(trace_words * 8, s));
    PG_CUDA_OK(cudaMemsetAsync(t_trace.ptr, 0, trace_words * 8, s));
    p.trace = t_trace.as<unsigned long long>();
  }
  row_gemm_tc_kernel<kProd, kEpi><<<2 * clusters, kThreads, smem, s>>>(p);
  PG_LAUNCH_CHECK();
  if (trace_path != nullptr) {
    std::vector<unsigned long long> h(trace_words);
    PG_CUDA_OK(cudaMemcpyAsync(h.data(), t_trace.ptr, trace_words * 8, cudaMemcpyDeviceToHost, s));
    PG_CUDA_OK(cudaStreamSynchronize(s));
    if (FILE* f = fopen(trace_path, "w")) {
      for (size_t i = 0; i < trace_words; i += 3)
        fprintf(f, "%zu %zu %llu %llu %llu\n", i / 3 / 128, (i / 3) % 128, h[i], h[i + 1], h[i + 2]);
      fclose(f);
    }
  }
  g_tc_launches[kEpi == EPI_SEGMAX ? 0 : 1].fetch_add(1, std::memory_order_relaxed);
  return PG_OK;
}

}  // namespace

int fc_tc_bf16x3(const float* x, int64_t m, int k, const float* w, const float* bias, int n, int act,
                 const float* residual, float* out, cudaStream_t s) {
  const TcShape t = tc_shape(k, n);
  // narrow / shallow layers (N < 8, K < 64: the 64->3, 64->4, 64->7 heads) stay on the fp32 FFMA kernel
  if (!t.ok || (k & 3) != 0 || m < 1) return fc_fp32_launch(x, m, k, w, bias, n, act, residual, out, n, s);
  TcParams p{};
  p.P = x;
  p.ldp = k;
  p.k_real = k;
  p.num_rows = m;
  p.out = out;
  p.ldo = n;
  p.act = act;
  p.residual = residual;
  Temp t_img, t_bias;
  return launch_row_gemm<PROD_ROWS, EPI_STORE>(p, t, w, k, n, bias, t_img, t_bias, s);
}

int edge_mlp_max_tc(int mode, const float* features, int c_in, const float* xyz_src, const float* xyz_dst,
                    const int32_t* dst_index, const int32_t* src, const int32_t* dst, int64_t num_edges,
                    int64_t num_src, int64_t num_dst, const float* const* weights, const float* const* biases,
                    const int32_t* dims, int num_layers, float* out, cudaStream_t s) {
  TcShape t{};
  const bool pool_tc = mode == PG_EDGE_POOL && num_layers == 4 && c_in == 1 && dims[1] == kPoolC1 &&
                       dims[2] == kPoolC2 && dims[3] == kPoolC3;
  if (mode == PG_EDGE_GNN && num_layers == 2) t = tc_shape(dims[1], dims[2]);
  if (pool_tc) {
    t = tc_shape(dims[3], dims[4]);
    t.ok = t.ok && tc_smem_bytes(t.kp, t.np, PROD_POOL) <= 227 * 1024;
  }
  if (t.ok && pool_tc && num_edges > 0) {
    // PointSetPooling (gnn.py:256-277): layers 1-3 of the point MLP run in the producer warps (FFMA),
    // the 128 -> 300 layer (79% of the FLOPs) on the tensor cores, max / bias / relu in the epilogue.
    Temp t_w, t_img, t_bias, t_err;
    PG_CUDA_OK(t_w.alloc(sizeof(float) * kPoolWFloats, s));
    float* pw = t_w.as<float>();
    const int sizes[6] = {4 * kPoolC1, kPoolC1, kPoolC1 * kPoolC2, kPoolC2, kPoolC2 * kPoolC3, kPoolC3};
    const float* srcs[6] = {weights[0], biases[0], weights[1], biases[1], weights[2], biases[2]};
    size_t off = 0;
    for (int i = 0; i < 6; ++i) {
      PG_CUDA_OK(cudaMemcpyAsync(pw + off, srcs[i], sizeof(float) * sizes[i], cudaMemcpyDeviceToDevice, s));
      off += sizes[i];
    }
    PG_CUDA_OK(t_err.alloc(sizeof(int), s));
    PG_CUDA_OK(cudaMemsetAsync(t_err.ptr, 0, sizeof(int), s));
    const int n = dims[4];
    if (int rc = fill_async(out, num_dst * n, -FLT_MAX, s)) return rc;
    TcParams p{};
    p.pool_w = pw;
    p.pool_feat = features;
    p.xyz_src = xyz_src;
    p.xyz_dst = xyz_dst;
    p.dst_index = dst_index;
    p.src = src;
    p.dst = dst;
    p.num_rows = num_edges;
    p.num_src = num_src;
    p.num_dst = num_dst;
    p.out = out;
    p.err = t_err.as<int>();
    if (int rc = launch_row_gemm<PROD_POOL, EPI_SEGMAX>(p, t, weights[3], dims[3], n, biases[3], t_img, t_bias, s)) return rc;
    int h = 0;
    PG_CUDA_OK(cudaMemcpyAsync(&h, t_err.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
    PG_CUDA_OK(cudaStreamSynchronize(s));
    PG_REQUIRE(h == 0, "set index out of range (point in [0,%lld), keypoint in [0,%lld))", (long long)num_src,
               (long long)num_dst);
    return PG_OK;
  }
  if (!t.ok || pool_tc || num_edges == 0)
    return edge_mlp_max_fp32(mode, features, c_in, xyz_src, xyz_dst, dst_index, src, dst, num_edges, num_src,
                             num_dst, weights, biases, dims, num_layers, out, s);
  PG_REQUIRE(dims[0] == c_in + 3, "dims[0]=%d must equal feature channels + 3 = %d", dims[0], c_in + 3);
  const int d1 = dims[1], n = dims[2];
  Temp t_p, t_w1x, t_img, t_bias, t_err;
  // hoisted first layer: P = F @ W1[:C] + b1 (row stride kp, pad columns zero) on the tensor cores too
  PG_CUDA_OK(t_p.alloc(sizeof(float) * num_src * t.kp, s));
  {
    const TcShape tp = tc_shape(c_in, d1);
    if (tp.ok && (c_in & 3) == 0 && t.kp == tp.np) {
      TcParams q{};
      q.P = features;
      q.ldp = c_in;
      q.k_real = c_in;
      q.num_rows = num_src;
      q.out = t_p.as<float>();
      q.ldo = t.kp;
      q.act = 0;
      q.residual = nullptr;
      Temp q_img, q_bias;
      // n = kp here: the pad columns [d1, kp) get bias 0 and zero weights -> written as exact zeros
      Temp w_pad, b_pad;
      PG_CUDA_OK(w_pad.alloc(sizeof(float) * size_t(c_in) * t.kp, s));
      pad_rows_kernel<<<64, 256, 0, s>>>(weights[0], c_in, d1, t.kp, w_pad.as<float>());
      PG_LAUNCH_CHECK();
      PG_CUDA_OK(b_pad.alloc(sizeof(float) * t.kp, s));
      pad_rows_kernel<<<2, 256, 0, s>>>(biases[0], 1, d1, t.kp, b_pad.as<float>());
      PG_LAUNCH_CHECK();
      const TcShape tq = tc_shape(c_in, t.kp);
      if (int rc = launch_row_gemm<PROD_ROWS, EPI_STORE>(q, tq, w_pad.as<float>(), c_in, t.kp, b_pad.as<float>(), q_img,
                                                         q_bias, s))
        return rc;
    } else if (int rc = fc_fp32_launch(features, num_src, c_in, weights[0], biases[0], d1, 0, nullptr, t_p.as<float>(),
                                       t.kp, s)) {
      return rc;
    }
  }
  PG_CUDA_OK(t_w1x.alloc(sizeof(float) * 3 * t.kp, s));
  pad_rows_kernel<<<4, 256, 0, s>>>(weights[0] + int64_t(c_in) * d1, 3, d1, t.kp, t_w1x.as<float>());
  PG_LAUNCH_CHECK();
  PG_CUDA_OK(t_err.alloc(sizeof(int), s));
  PG_CUDA_OK(cudaMemsetAsync(t_err.ptr, 0, sizeof(int), s));
  if (int rc = fill_async(out, num_dst * n, -FLT_MAX, s)) return rc;
  TcParams p{};
  p.P = t_p.as<float>();
  p.ldp = t.kp;
  p.xyz_src = xyz_src;
  p.xyz_dst = xyz_dst;
  p.dst_index = dst_index;
  p.src = src;
  p.dst = dst;
  p.num_rows = num_edges;
  p.num_src = num_src;
  p.num_dst = num_dst;
  p.w1x = t_w1x.as<float>();
  p.out = out;
  p.err = t_err.as<int>();
  if (int rc = launch_row_gemm<PROD_GNN, EPI_SEGMAX>(p, t, weights[1], d1, n, biases[1], t_img, t_bias, s)) return rc;
  int h = 0;
  PG_CUDA_OK(cudaMemcpyAsync(&h, t_err.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
  PG_CUDA_OK(cudaStreamSynchronize(s));
  PG_REQUIRE(h == 0, "edge index out of range (src in [0,%lld), dst in [0,%lld))", (long long)num_src,
             (long long)num_dst);
  return PG_OK;
}

}  // namespace pg
