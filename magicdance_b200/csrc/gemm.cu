// tcgen05 GEMM / 3x3 implicit-GEMM convolution for sm_100a.
//
//   D[M,N] = epilogue(A[M,K] * B[N,K]^T), fp16 operands, fp32 accumulation in TMEM.
//
// One CTA computes one 128 x BN output tile (optionally one K-split of it).  Warp roles:
//   warp 0     TMA producer   — streams 128x64 A tiles and BNx64 B tiles (128B-swizzled) through a
//                               kStages-deep smem ring; completion on `full` mbarriers.  In conv
//                               mode the A tile of tap (kh,kw) is a 4-D TMA box over the NHWC
//                               activation shifted by (kw-1, kh-1): out-of-bounds rows/cols are
//                               zero-filled by TMA, which IS the pad-1 halo — no im2col buffer.
//   warp 1     MMA issuer     — one elected lane issues tcgen05.mma (M=128, N=BN, K=16) x4 per
//                               stage, tcgen05.commit releases the stage (`empty`) and finally
//                               signals `acc_full`.  Also owns the TMEM allocation.
//   warps 2-5  epilogue       — tcgen05.ld the accumulator (thread == output row), apply bias /
//                               per-batch bias (timestep embedding) / residual / GEGLU, store fp16.
// Two CTAs fit per SM (3 stages, <=108 KB smem, <=256 TMEM columns each) so one CTA's epilogue
// overlaps the other's main loop.
#include <stdlib.h>

#include "common.cuh"

namespace mdb {

int get_gemm_tuning(int key);
void set_gemm_tuning(int key, int value);

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 halves = 128 B = one swizzle row
constexpr int kGemmThreads = 192;

struct GemmKParams {
  CUtensorMap tmA;
  CUtensorMap tmA2;
  CUtensorMap tmB;
  __half* d;
  long long ldd;
  const float* bias;
  long long bias_batch_stride;
  const __half* residual;
  long long ldr;
  float* ws;
  int rows_per_batch;
  int m, n;
  int k_chunks;        // total K / 64
  int k1_chunks;       // chunks taken from tmA (plain mode); rest from tmA2
  int chunks_per_split;
  int splits;
  int conv;            // conv mode
  int chunks_per_tap;  // c / 64
  int w, hw;           // conv geometry of the OUTPUT (== input for stride 1)
  int cs;              // conv stride (1 | 2): input pixel = cs * output pixel + tap - 1
  int cluster_reduce;  // split-K partners form a cluster (1,1,splits) and reduce through distributed smem
  // LayerNorm folded into this GEMM (gemm_tc_kernel only): D = rstd_r (A W'^T - mean_r u) + bias with W' = W diag(gamma),
  // u[n] = sum_k W'[n][k]; the epilogue warps compute (mean_r, rstd_r) of their A rows from the staged tiles
  const float* ln_u;
  float ln_eps;
};

// STAGES = 3: <=108 KB, two CTAs per SM (large grids: the co-resident CTA hides the TMA round trip).
// STAGES = 6 (8 for 80-wide tiles): one CTA per SM with a ring deep enough to cover the TMA latency on
// its own — used when the grid has at most one CTA per SM anyway and the K loop is long.
// PAIR (gemm_pair_kernel below): two CTAs run ONE cta_group::2 UMMA of shape 256 x BN: each keeps its own
// 128 A rows and only BN/2 of the B rows, so a 256 x 160 pair tile pulls 26 KB per CTA and K chunk through
// the L2 -> SM fabric where two independent 128 x 160 tiles pull 36 KB — the fabric (~6.3 KB/clk chip-wide)
// is what bounds the large GEMMs of this path, not the tensor pipe.
template <int BN, int STAGES, bool PAIR = false>
struct GemmSmem {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBRows = PAIR ? BN / 2 : BN;
  static constexpr int kBBytes = kBRows * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTotal = STAGES * kStageBytes + 1024;  // + alignment slack
};

__device__ __forceinline__ void epi_store_chunk(const GemmKParams& p, long long row, int col0, int ncols,
                                                float (&v)[32], const float* bias_chunk, const uint4* res_pref) {
  // v holds columns col0 .. col0+31 of `row` (fp32 accumulators); bias + residual, then fp16 store.
  // Whole groups of 8 columns go through 16-byte accesses, a ragged tail (N = 77) is scalar.
  // bias_chunk: this chunk's 32 bias values (shared or global memory) or nullptr;
  // res_pref:   this chunk's residual, prefetched into registers during the main loop, or nullptr.
  const __half* rp = (p.residual != nullptr) ? p.residual + row * p.ldr + col0 : nullptr;
  __half* dp = p.d + row * p.ldd + col0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q * 8 + 8 <= ncols) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = v[q * 8 + e];
      if (bias_chunk != nullptr) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias_chunk + q * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias_chunk + q * 8 + 4);
        o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w;
        o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
      }
      if (rp != nullptr) {
        const uint4 r4 = (res_pref != nullptr) ? res_pref[q] : *reinterpret_cast<const uint4*>(rp + q * 8);
        const __half2* h2 = reinterpret_cast<const __half2*>(&r4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          o[2 * e] += f.x;
          o[2 * e + 1] += f.y;
        }
      }
      uint4 o4;
      o4.x = pack_half2(o[0], o[1]);
      o4.y = pack_half2(o[2], o[3]);
      o4.z = pack_half2(o[4], o[5]);
      o4.w = pack_half2(o[6], o[7]);
      *reinterpret_cast<uint4*>(dp + q * 8) = o4;
    } else if (q * 8 < ncols) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int j = q * 8 + e;
        if (j < ncols) {
          float x = v[j];
          if (bias_chunk != nullptr) x += bias_chunk[j];
          if (rp != nullptr) x += __half2float(rp[j]);
          dp[j] = __float2half_rn(x);
        }
      }
    }
  }
}

template <int BN, bool GEGLU, int kStages>
__global__ void __launch_bounds__(kGemmThreads, (kStages <= 3) ? 2 : 1)
    gemm_tc_kernel(const __grid_constant__ GemmKParams p) {
  using S = GemmSmem<BN, kStages, false>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_bias[BN];  // this N tile's bias row (when one row serves all batches)
  __shared__ __align__(16) float s_lnu[BN];   // this N tile's u (LayerNorm folded into the GEMM)

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBM;
  const int n0 = blockIdx.y * BN;
  const int split = blockIdx.z;
  const int kc_begin = split * p.chunks_per_split;
  const int kc_end = min(p.k_chunks, kc_begin + p.chunks_per_split);
  const int n_iter = kc_end - kc_begin;
  constexpr uint32_t kTmemCols = (BN <= 64) ? 64 : (BN <= 128 ? 128 : 256);
  constexpr int kRedLd = BN + 4;  // fp32 row pitch of the split-K partial tile parked in shared memory
  static_assert(kBM * kRedLd * 4 <= kStages * S::kStageBytes, "partial tile must fit in the operand ring");

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], p.ln_u != nullptr ? 5 : 1);  // LayerNorm fusion: + one arrival per epilogue warp
    }
    mbar_init(&acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, kTmemCols);
  // a bias row that serves all batches is a constant weight: stage it before the PDL wait (overlaps the
  // previous kernel's tail).  Per-batch biases (timestep embedding) are produced upstream and are read later.
  const bool bias_in_smem = (p.bias != nullptr) && (p.bias_batch_stride == 0) && !p.cluster_reduce;
  if (bias_in_smem && warp >= 2) {
    for (int j = threadIdx.x - 64; j < BN; j += 128) s_bias[j] = (n0 + j < p.n) ? p.bias[n0 + j] : 0.f;
  }
  if (p.ln_u != nullptr && warp >= 2) {  // a constant of the weights, like the bias
    for (int j = threadIdx.x - 64; j < BN; j += 128) s_lnu[j] = (n0 + j < p.n) ? p.ln_u[n0 + j] : 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory from here on

  if (warp == 0) {
    if (lane == 0 && n_iter > 0) {
      // conv geometry of this M tile
      int b0 = 0, y0 = 0, x0 = 0;
      if (p.conv) {
        b0 = m0 / p.hw;
        y0 = (p.hw >= kBM) ? (m0 % p.hw) / p.w : 0;
        x0 = (p.hw >= kBM) ? (m0 % p.hw) % p.w : 0;  // != 0 only for rows wider than the 128-pixel tile
      }
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * S::kStageBytes;
        uint8_t* sb = sa + S::kABytes;
        const int kc = kc_begin + it;
        mbar_expect_tx(&full_bar[s], S::kStageBytes);
        if (p.conv) {
          const int tap = kc / p.chunks_per_tap;
          const int cc = kc - tap * p.chunks_per_tap;
          const int kh = tap / 3, kw = tap - kh * 3;
          tma_load_4d(sa, &p.tmA, &full_bar[s], cc * kBK, p.cs * x0 + kw - 1, p.cs * y0 + kh - 1, b0);
        } else if (kc < p.k1_chunks) {
          tma_load_2d(sa, &p.tmA, &full_bar[s], kc * kBK, m0);
        } else {
          tma_load_2d(sa, &p.tmA2, &full_bar[s], (kc - p.k1_chunks) * kBK, m0);
        }
        tma_load_2d(sb, &p.tmB, &full_bar[s], kc * kBK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_iter > 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kBM, BN);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(smem + s * S::kStageBytes);
        const uint32_t b_addr = a_addr + S::kABytes;
        const uint64_t da = umma_desc_k_sw128(a_addr);
        const uint64_t db = umma_desc_k_sw128(b_addr);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // advance 16 halves (32 B) along K inside the 128B swizzle row: +2 in the >>4 address field
          umma_f16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&acc_bar);
    }
  } else if (n_iter > 0) {
    // ---------------- epilogue warps 2..5 ----------------
    const int g = warp & 3;  // TMEM lane group this warp may access
    const long long row = static_cast<long long>(m0) + g * 32 + lane;
    const bool row_ok = row < p.m;
    // While the main loop runs these warps are idle: fetch what the epilogue will need — the bias row into
    // shared memory and this thread's residual row into registers — so that the tail of the kernel does not
    // pay two dependent global-memory round trips per 32-column chunk.
    constexpr int kResVecs = (GEGLU ? 0 : ((BN + 31) / 32) * 4);
    uint4 res_pref[kResVecs > 0 ? kResVecs : 1];
    const bool res_in_regs = !GEGLU && (p.residual != nullptr) && !p.cluster_reduce && (p.splits == 1);
    if constexpr (!GEGLU) {
      if (res_in_regs && row_ok) {
        const __half* rrow = p.residual + row * p.ldr + n0;
#pragma unroll
        for (int q = 0; q < kResVecs; ++q)
          if (q * 8 + 8 <= BN && n0 + q * 8 + 8 <= p.n) res_pref[q] = *reinterpret_cast<const uint4*>(rrow + q * 8);
      }
    }
    // LayerNorm folded into the GEMM: statistics of this thread's A row (pivot-shifted sums, biased variance as
    // nn.LayerNorm), taken from the A tiles AS THEY PASS THROUGH SHARED MEMORY on their way to the tensor core — the
    // epilogue warps are idle during the main loop, wait on the same `full` barriers as the MMA thread and add one
    // arrival per warp to `empty` (armed with 1 + 4 arrivals in this mode); no extra global or L2 traffic.
    float ln_mean = 0.f, ln_rstd = 1.f;
    const bool ln = p.ln_u != nullptr;
    if (ln) {
      const int rt = g * 32 + lane;  // row inside the tile == TMEM lane
      float pivot = 0.f, s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
      for (int it = 0; it < n_iter; ++it) {
        const int s_ = it % kStages;
        mbar_wait(&full_bar[s_], (it / kStages) & 1);
        const uint32_t arow = smem_u32(smem + s_ * S::kStageBytes) + (rt >> 3) * 1024 + (rt & 7) * 128;
#pragma unroll
        for (int l = 0; l < 8; ++l) {  // logical 16-byte chunk l lives at physical chunk l ^ (row & 7): conflict-free
          uint4 u4;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(u4.x), "=r"(u4.y), "=r"(u4.z), "=r"(u4.w)
                       : "r"(arow + ((l ^ (rt & 7)) << 4)));
          const __half2* h2 = reinterpret_cast<const __half2*>(&u4);
          if (it == 0 && l == 0) pivot = __low2float(h2[0]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h2[e]);
            const float d0 = f.x - pivot, d1 = f.y - pivot;
            s0 += d0; q0 = fmaf(d0, d0, q0);
            s1 += d1; q1 = fmaf(d1, d1, q1);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s_]);  // this warp is done reading stage s_
      }
      const float inv_k = 1.0f / static_cast<float>(p.k_chunks * kBK);
      const float ms = (s0 + s1) * inv_k;
      ln_mean = pivot + ms;
      ln_rstd = rsqrtf(fmaxf(fmaf(-ms, ms, (q0 + q1) * inv_k), 0.f) + p.ln_eps);
    }
    mbar_wait(&acc_bar, 0);
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(g * 32) << 16);
    if constexpr (!GEGLU) {
#pragma unroll
      for (int ch = 0; ch < (BN + 31) / 32; ++ch) {
        const int col0 = n0 + ch * 32;
        if (col0 >= p.n) break;  // warp-uniform
        uint32_t r[32];
        if (BN % 32 != 0 && ch == BN / 32) {  // 16-column tail of an 80-wide tile
          uint32_t r16[16];
          tmem_ld_x16(taddr + ch * 32, r16);
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = r16[j];
#pragma unroll
          for (int j = 16; j < 32; ++j) r[j] = 0u;
        } else {
          tmem_ld_x32(taddr + ch * 32, r);
        }
        tmem_wait_ld();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (ln) {  // rstd_r (acc - mean_r u[n]); the bias row carries W beta (+ the layer's own bias)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ch * 32 + j < BN) v[j] = ln_rstd * fmaf(-ln_mean, s_lnu[ch * 32 + j], v[j]);
        }
        const int ncols = min(min(32, BN - ch * 32), p.n - col0);
        if (p.cluster_reduce) {
          // split-K inside a cluster: park this CTA's fp32 partial tile in its own shared memory (the
          // operand ring is idle by now); the cluster reduces it through DSMEM below.
          float* rp = reinterpret_cast<float*>(smem) + (g * 32 + lane) * kRedLd + ch * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (j < ncols) *reinterpret_cast<float4*>(rp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else if (p.splits > 1) {
          // split-K through global memory: this split's fp32 partial goes to its own workspace slab
          // (plain vector stores); splitk_finalize_kernel sums the slabs in a fixed order.
          if (row_ok) {
            float* wp = p.ws + (static_cast<long long>(split) * p.m + row) * p.n + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j + 4 <= ncols && (p.n & 3) == 0) {
                *reinterpret_cast<float4*>(wp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (j + e < ncols) wp[j + e] = v[j + e];
              }
            }
          }
        } else if (row_ok) {
          const float* bias_chunk = nullptr;
          if (bias_in_smem) {
            bias_chunk = s_bias + ch * 32;
          } else if (p.bias != nullptr) {
            const long long brow = (p.bias_batch_stride != 0) ? (row / p.rows_per_batch) : 0;
            bias_chunk = p.bias + brow * p.bias_batch_stride + col0;
          }
          epi_store_chunk(p, row, col0, ncols, v, bias_chunk, (res_in_regs && ncols == 32) ? &res_pref[ch * 4] : nullptr);
        }
      }
    } else {
      // GEGLU: chunk pairs (value, gate); output column = n0/2 + pair*32 + j
      const long long brow = (p.bias_batch_stride != 0) ? (row / p.rows_per_batch) : 0;
#pragma unroll 1
      for (int pr = 0; pr < BN / 64; ++pr) {
        const int col0 = n0 + pr * 64;
        if (col0 >= p.n) break;
        uint32_t rv[32], rg[32];
        tmem_ld_x32(taddr + pr * 64, rv);
        tmem_ld_x32(taddr + pr * 64 + 32, rg);
        tmem_wait_ld();
        if (row_ok) {
          const float* bp = bias_in_smem ? (s_bias + pr * 64)
                                         : ((p.bias != nullptr) ? p.bias + brow * p.bias_batch_stride + col0 : nullptr);
          uint4* d4 = reinterpret_cast<uint4*>(p.d + row * p.ldd + (col0 >> 1));
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int j = q * 8 + e;
              float a = __uint_as_float(rv[j]);
              float gt = __uint_as_float(rg[j]);
              if (bp != nullptr) {
                a += bp[j];
                gt += bp[32 + j];
              }
              o[e] = a * gelu_erf_f(gt);
            }
            uint4 o4;
            o4.x = pack_half2(o[0], o[1]);
            o4.y = pack_half2(o[2], o[3]);
            o4.z = pack_half2(o[4], o[5]);
            o4.w = pack_half2(o[6], o[7]);
            d4[q] = o4;
          }
        }
      }
    }
  }

  if constexpr (!GEGLU) {
    if (p.cluster_reduce) {
      // ---- split-K reduction across the cluster through distributed shared memory ----
      // cluster = the `splits` CTAs of this output tile.  CTA r owns rows [r*R, (r+1)*R) of the tile: it
      // sums that slice over all partners' parked partials (ld.shared::cluster), applies the epilogue
      // and stores fp16.  No global workspace, no second kernel.
      const int S_ = p.splits;
      const int R = kBM / S_;
      const int groups = BN / 8;
      const uint32_t crank = cluster_ctarank();
      // this thread's share of the residual is known up front: fetch it before the cluster barrier
      constexpr int kMaxItems = 4;
      uint4 rpre[kMaxItems];
      if (p.residual != nullptr) {
#pragma unroll
        for (int q = 0; q < kMaxItems; ++q) {
          const int item = threadIdx.x + q * kGemmThreads;
          if (item < R * groups) {
            const int rl = item / groups, cgp = item - rl * groups;
            const long long row = static_cast<long long>(m0) + static_cast<int>(crank) * R + rl;
            const int col0 = n0 + cgp * 8;
            if (row < p.m && col0 + 8 <= p.n) rpre[q] = *reinterpret_cast<const uint4*>(p.residual + row * p.ldr + col0);
          }
        }
      }
      cluster_sync_all();
      const uint32_t red_base = smem_u32(smem);
#pragma unroll 1
      for (int it_ = 0; it_ * kGemmThreads < R * groups; ++it_) {
        const int item = threadIdx.x + it_ * kGemmThreads;
        if (item >= R * groups) break;
        const int rl = item / groups, cgp = item - rl * groups;
        const int rt = static_cast<int>(crank) * R + rl;  // row inside the tile
        const long long row = static_cast<long long>(m0) + rt;
        const int col0 = n0 + cgp * 8;
        if (row >= p.m || col0 >= p.n) continue;
        const uint32_t off = red_base + static_cast<uint32_t>((rt * kRedLd + cgp * 8) * 4);
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = 0.f;
        for (int pr = 0; pr < S_; ++pr) {
          const uint32_t ra = dsmem_map(off, static_cast<uint32_t>(pr));
          const float4 a = dsmem_ld_f4(ra), b = dsmem_ld_f4(ra + 16);
          o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
          o[4] += b.x; o[5] += b.y; o[6] += b.z; o[7] += b.w;
        }
        const int ncols = min(8, p.n - col0);
        const long long brow = (p.bias_batch_stride != 0) ? (row / p.rows_per_batch) : 0;
        if (p.bias != nullptr) {
          const float* bp = p.bias + brow * p.bias_batch_stride + col0;
          for (int e = 0; e < ncols; ++e) o[e] += bp[e];
        }
        if (ncols == 8) {
          if (p.residual != nullptr) {
            uint4 r4 = make_uint4(0, 0, 0, 0);
            if (it_ < kMaxItems) {
#pragma unroll
              for (int q = 0; q < kMaxItems; ++q)
                if (q == it_) r4 = rpre[q];
            } else {
              r4 = *reinterpret_cast<const uint4*>(p.residual + row * p.ldr + col0);
            }
            const __half2* h2 = reinterpret_cast<const __half2*>(&r4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(h2[e]);
              o[2 * e] += f.x;
              o[2 * e + 1] += f.y;
            }
          }
          uint4 o4;
          o4.x = pack_half2(o[0], o[1]);
          o4.y = pack_half2(o[2], o[3]);
          o4.z = pack_half2(o[4], o[5]);
          o4.w = pack_half2(o[6], o[7]);
          *reinterpret_cast<uint4*>(p.d + row * p.ldd + col0) = o4;
        } else {
          for (int e = 0; e < ncols; ++e) {
            float x = o[e];
            if (p.residual != nullptr) x += __half2float(p.residual[row * p.ldr + col0 + e]);
            p.d[row * p.ldd + col0 + e] = __float2half_rn(x);
          }
        }
      }
      cluster_sync_all();  // nobody leaves while a partner may still read its shared memory
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent CTA-pair GEMM (the large grids: eight frames per GPU, the batched appearance passes): one cluster of
// two CTAs per TPC, ONE CTA per SM (the ~200 KB operand ring keeps every other tensor-memory kernel off the SM),
// all 512 TMEM columns owned by the pair and used as two accumulator buffers.  Each pair walks the 256 x BN output
// tiles t = cluster, cluster + #clusters, ... (M fastest, so the pairs running at one time share a B tile); the
// leader's MMA thread fills buffer (i & 1) for the i-th tile while the epilogue warps of both CTAs drain the other
// one, and the producers run ahead into the next tile's operands.  Barriers:
//   full/empty[stage]  both CTAs' TMA bytes are credited to the LEADER's `full`; tcgen05.commit multicast frees a
//                      stage in both CTAs
//   acc_full[2]        (each CTA, count 1)   tcgen05.commit multicast: tile i is complete in both CTAs' TMEM
//   acc_empty[2]       (leader, count 16)    one arrival per epilogue warp of BOTH CTAs: buffer drained
// Epilogue: EIGHT warps per CTA (two per TMEM lane quarter, each taking half of the tile's 32-column chunks); a warp
// hands its share of the buffer back as soon as its last tcgen05.ld has returned; the residual of the NEXT chunk
// is fetched while the current one is converted; output goes through shared memory and TMA (a warp packs 32 rows x
// 32 columns of fp16 into a 2 KB staging buffer and one lane issues cp.async.bulk.tensor — rows >= M and columns
// >= N are clipped by the tensor map, whole 64-byte row segments reach L2 instead of 16-byte pieces).
// Launched WITHOUT programmatic stream serialization and never triggering its dependents early: a pair whose
// cta_group::2 TMEM allocation is pending must not share its SMs with a foreign tensor-memory CTA (measured: the
// one-tile-per-launch pair mode of round 1 dead-locked inside the full step for exactly that reason).
// Measured on B200 (profiles/r02_*): +7 % on the eight-frame step over the single-CTA tiles.
// ------------------------------------------------------------------------------------------------
constexpr int kPairqThreads = 320;                    // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kPairqEpiWarps = 8;
constexpr int kOutBox = 32;                           // TMA-store box: 32 rows x 32 fp16 columns
constexpr int kOutBufBytes = kOutBox * kOutBox * 2;   // 2 KB

template <int BN, int STAGES>
struct PairqSmem {
  using S = GemmSmem<BN, STAGES, true>;
  static constexpr int kRing = STAGES * S::kStageBytes;               // multiple of 1024
  static constexpr int kOut = kPairqEpiWarps * 2 * kOutBufBytes;      // 32 KB
  static constexpr int kTotal = kRing + kOut + 1024;
};

// this thread's residual for output columns [col0, col0 + 32) of `row` (N % 8 == 0 is a launch condition)
__device__ __forceinline__ void pairq_load_res(const GemmKParams& p, long long row, bool row_ok, int col0,
                                               uint4 (&dst)[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (row_ok && col0 + q * 8 + 8 <= p.n) dst[q] = *reinterpret_cast<const uint4*>(p.residual + row * p.ldr + col0 + q * 8);
    else dst[q] = make_uint4(0u, 0u, 0u, 0u);
  }
}

template <int BN, bool GEGLU, int kStages>
__global__ void __launch_bounds__(kPairqThreads, 1)
    gemm_pair_kernel(const __grid_constant__ GemmKParams p, const __grid_constant__ CUtensorMap tmD) {
  using S = GemmSmem<BN, kStages, true>;
  using Q = PairqSmem<BN, kStages>;
  // UMMA N <= 256: a 320-wide tile (BN = 320, the full width of the 64x64 level: every A tile is loaded exactly once)
  // is TWO 160-wide MMAs per K step into adjacent accumulator columns.  320 columns leave no room for a second
  // accumulator buffer, so that tile is single-buffered (the epilogue of a long-K tile is a few % of its main loop).
  constexpr int kParts = (BN > 256) ? 2 : 1;
  constexpr int kPartN = BN / kParts;
  constexpr int kBufs = (BN > 256) ? 1 : 2;
  static_assert(BN % 32 == 0 && kPartN <= 256 && kPartN % 16 == 0 && BN <= 512, "tile width");
  static_assert(((kPartN / 2) * 128) % 1024 == 0, "each part's B rows start on a swizzle-atom boundary");
  static_assert(!GEGLU || BN % 64 == 0, "GEGLU tiles hold (value, gate) chunk pairs");
  static_assert(Q::kRing % 1024 == 0, "the staging buffers follow the ring and need 128-byte alignment");
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages];
  __shared__ __align__(8) uint64_t empty_bar[kStages];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kAccStride = 256;  // columns between the two accumulator buffers

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 2 * kPairqEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(&tmem_base_smem, kTmemCols);
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t pair_rank = cluster_ctarank();
  const int n_clusters = gridDim.x >> 1;
  const int cluster_id = blockIdx.x >> 1;
  const int m_pairs = (p.m + 2 * kBM - 1) / (2 * kBM);
  const int n_tiles = (p.n + BN - 1) / BN;
  const int total_tiles = m_pairs * n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;  // K chunks issued so far (ring position)
      for (int t = cluster_id; t < total_tiles; t += n_clusters) {
        const int mp = t % m_pairs, nt = t / m_pairs;
        const int m0 = (2 * mp + static_cast<int>(pair_rank)) * kBM;
        const int n0 = nt * BN;
        int b0 = 0, y0 = 0, x0 = 0;
        if (p.conv) {
          b0 = m0 / p.hw;
          y0 = (p.hw >= kBM) ? (m0 % p.hw) / p.w : 0;
          x0 = (p.hw >= kBM) ? (m0 % p.hw) % p.w : 0;
        }
        for (int kc = 0; kc < p.k_chunks; ++kc, ++g) {
          const int s = g % kStages;
          const uint32_t ph = (g / kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          if (pair_rank == 0) mbar_expect_tx(&full_bar[s], 2 * S::kStageBytes);
          const uint32_t fb = dsmem_map(smem_u32(&full_bar[s]), 0);
          if (p.conv) {
            const int tap = kc / p.chunks_per_tap;
            const int cc = kc - tap * p.chunks_per_tap;
            const int kh = tap / 3, kw = tap - kh * 3;
            tma_load_4d_pair(sa, &p.tmA, fb, cc * kBK, p.cs * x0 + kw - 1, p.cs * y0 + kh - 1, b0);
          } else if (kc < p.k1_chunks) {
            tma_load_2d_pair(sa, &p.tmA, fb, kc * kBK, m0);
          } else {
            tma_load_2d_pair(sa, &p.tmA2, fb, (kc - p.k1_chunks) * kBK, m0);
          }
#pragma unroll
          for (int h = 0; h < kParts; ++h)  // this CTA's half of each part's B rows (tmB's box is kPartN / 2 rows)
            tma_load_2d_pair(sb + h * (kPartN / 2) * 128, &p.tmB, fb, kc * kBK,
                             n0 + h * kPartN + static_cast<int>(pair_rank) * (kPartN / 2));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && pair_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * kBM, kPartN);
      uint32_t g = 0;
      int i = 0;
      for (int t = cluster_id; t < total_tiles; t += n_clusters, ++i) {
        const int buf = i % kBufs;
        mbar_wait(&acc_empty[buf], ((i / kBufs) & 1) ^ 1);  // all 16 epilogue warps of the pair have drained this buffer
        tc_fence_after_sync();
        const uint32_t tacc = tmem_base + buf * kAccStride;
        for (int kc = 0; kc < p.k_chunks; ++kc, ++g) {
          const int s = g % kStages;
          const uint32_t ph = (g / kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + s * S::kStageBytes);
          const uint32_t b_addr = a_addr + S::kABytes;
          const uint64_t da = umma_desc_k_sw128(a_addr);
          const uint64_t db = umma_desc_k_sw128(b_addr);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
#pragma unroll
            for (int h = 0; h < kParts; ++h)
              umma_f16_ss_pair(tacc + h * kPartN, da + 2 * k, db + h * (((kPartN / 2) * 128) >> 4) + 2 * k, idesc,
                               (kc | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[s]);
        }
        umma_commit_pair(&acc_full[buf]);
      }
    }
  } else {
    // ---------------- epilogue warps 2..9 of both CTAs ----------------
    const int ew = warp - 2;
    const int gq = warp & 3;     // TMEM lane quarter this warp may access (hardware rule: warp % 4)
    const int half = ew >> 2;    // which half of the tile's column chunks this warp drains
    uint8_t* obuf = smem + Q::kRing + ew * (2 * kOutBufBytes);
    const uint32_t acc_empty_leader0 = dsmem_map(smem_u32(&acc_empty[0]), 0);
    const uint32_t acc_empty_leader1 = dsmem_map(smem_u32(&acc_empty[1]), 0);
    constexpr int kUnits = GEGLU ? BN / 64 : BN / 32;   // 32 OUTPUT columns each
    constexpr int kUnitsLo = (kUnits + 1) / 2;
    const int u_begin = half ? kUnitsLo : 0;
    const int u_end = half ? kUnits : kUnitsLo;
    constexpr int kAccColsPerUnit = GEGLU ? 64 : 32;
    const bool have_res = !GEGLU && (p.residual != nullptr);
    uint32_t ob = 0;  // staging-buffer parity, runs on across tiles
    uint4 rcur[4], rnxt[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) rcur[q] = rnxt[q] = make_uint4(0u, 0u, 0u, 0u);
    if (have_res && cluster_id < total_tiles) {  // residual of the first tile's first chunk
      const int mp = cluster_id % m_pairs, nt = cluster_id / m_pairs;
      const long long row = static_cast<long long>((2 * mp + static_cast<int>(pair_rank)) * kBM) + gq * 32 + lane;
      pairq_load_res(p, row, row < p.m, nt * BN + u_begin * 32, rcur);
    }
    int i = 0;
    for (int t = cluster_id; t < total_tiles; t += n_clusters, ++i) {
      const int buf = i % kBufs;
      const int mp = t % m_pairs, nt = t / m_pairs;
      const int m0 = (2 * mp + static_cast<int>(pair_rank)) * kBM;
      const int n0 = nt * BN;
      const int row0 = m0 + gq * 32;                       // first row of this warp's slab
      const long long row = static_cast<long long>(row0) + lane;
      const bool row_ok = row < p.m;
      const long long brow = (p.bias_batch_stride != 0) ? (row / p.rows_per_batch) : 0;
      mbar_wait(&acc_full[buf], (i / kBufs) & 1);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + buf * kAccStride + (static_cast<uint32_t>(gq * 32) << 16);
      bool arrived = false;
#pragma unroll 1
      for (int u = u_begin; u < u_end; ++u) {
        const int col0 = n0 + u * kAccColsPerUnit;          // first accumulator column of this unit (global N index)
        if (col0 >= p.n) break;                             // warp-uniform
        const bool last = (u + 1 == u_end) || (col0 + kAccColsPerUnit >= p.n);
        uint4 o4[4];
        if constexpr (!GEGLU) {
          uint32_t r[32];
          tmem_ld_x32(taddr + u * 32, r);
          if (have_res && !last) pairq_load_res(p, row, row_ok, col0 + 32, rnxt);
          tmem_wait_ld();
          if (last) {  // this warp's share of the buffer is in registers: hand it back to the MMA thread now
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(buf == 0 ? acc_empty_leader0 : acc_empty_leader1);
            arrived = true;
          }
          const int ncols = min(32, p.n - col0);
          const float* bp = (p.bias != nullptr && row_ok) ? p.bias + brow * p.bias_batch_stride + col0 : nullptr;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(r[q * 8 + e]);
            if (bp != nullptr && q * 8 + 8 <= ncols) {
              const float4 b0 = *reinterpret_cast<const float4*>(bp + q * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(bp + q * 8 + 4);
              o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w;
              o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
            }
            if (have_res) {
              const __half2* h2 = reinterpret_cast<const __half2*>(&rcur[q]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h2[e]);
                o[2 * e] += f.x;
                o[2 * e + 1] += f.y;
              }
            }
            o4[q].x = pack_half2(o[0], o[1]);
            o4[q].y = pack_half2(o[2], o[3]);
            o4[q].z = pack_half2(o[4], o[5]);
            o4[q].w = pack_half2(o[6], o[7]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) rcur[q] = rnxt[q];
        } else {
          uint32_t rv[32], rg[32];
          tmem_ld_x32(taddr + u * 64, rv);
          tmem_ld_x32(taddr + u * 64 + 32, rg);
          tmem_wait_ld();
          if (last) {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(buf == 0 ? acc_empty_leader0 : acc_empty_leader1);
            arrived = true;
          }
          const float* bp = (p.bias != nullptr && row_ok) ? p.bias + brow * p.bias_batch_stride + col0 : nullptr;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int j = q * 8 + e;
              float a = __uint_as_float(rv[j]);
              float gt = __uint_as_float(rg[j]);
              if (bp != nullptr) {
                a += bp[j];
                gt += bp[32 + j];
              }
              o[e] = a * gelu_erf_poly_f(gt);
            }
            o4[q].x = pack_half2(o[0], o[1]);
            o4[q].y = pack_half2(o[2], o[3]);
            o4[q].z = pack_half2(o[4], o[5]);
            o4[q].w = pack_half2(o[6], o[7]);
          }
        }
        // registers -> staging buffer -> TMA store.  The buffer was last used two stores ago: at most one
        // younger bulk group may still be reading shared memory.
        uint8_t* sbuf = obuf + ob * kOutBufBytes;
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        uint4* srow = reinterpret_cast<uint4*>(sbuf + lane * (kOutBox * 2));
#pragma unroll
        for (int q = 0; q < 4; ++q) srow[q] = o4[q];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && row0 < p.m) {
          tma_store_2d(&tmD, sbuf, GEGLU ? (col0 >> 1) : col0, row0);
          tma_store_commit();
        }
        ob ^= 1u;
      }
      if (!arrived) {  // no column of this warp's half lies inside N
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(buf == 0 ? acc_empty_leader0 : acc_empty_leader1);
      }
      if (have_res && t + n_clusters < total_tiles) {  // residual of the next tile's first chunk
        const int tn = t + n_clusters;
        const int mpn = tn % m_pairs, ntn = tn / m_pairs;
        const long long rown = static_cast<long long>((2 * mpn + static_cast<int>(pair_rank)) * kBM) + gq * 32 + lane;
        pairq_load_res(p, rown, rown < p.m, ntn * BN + u_begin * 32, rcur);
      }
    }
    if (lane == 0) tma_store_wait_all();  // the staging buffers must outlive the stores that read them
  }

  tc_fence_before_sync();
  cluster_sync_all();  // the leader's MMAs read the partner's shared memory and write its TMEM
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// split-K second pass: sum of the fp32 partial slabs ws[splits][M][N] -> bias/residual -> fp16 D
__global__ void splitk_finalize_kernel(GemmKParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const long long slab = static_cast<long long>(p.m) * p.n;
  if ((p.n & 3) == 0) {
    const long long total4 = slab >> 2;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long e = i << 2;
      const long long row = e / p.n;
      const int col = static_cast<int>(e - row * p.n);
      float4 a = *reinterpret_cast<const float4*>(p.ws + e);
      for (int s = 1; s < p.splits; ++s) {
        const float4 b = *reinterpret_cast<const float4*>(p.ws + s * slab + e);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      if (p.bias != nullptr) {
        const long long brow = (p.bias_batch_stride != 0) ? (row / p.rows_per_batch) : 0;
        const float* bp = p.bias + brow * p.bias_batch_stride + col;
        a.x += bp[0]; a.y += bp[1]; a.z += bp[2]; a.w += bp[3];
      }
      if (p.residual != nullptr) {
        const __half2* rp = reinterpret_cast<const __half2*>(p.residual + row * p.ldr + col);
        const float2 r0 = __half22float2(rp[0]), r1 = __half22float2(rp[1]);
        a.x += r0.x; a.y += r0.y; a.z += r1.x; a.w += r1.y;
      }
      uint2 o;
      o.x = pack_half2(a.x, a.y);
      o.y = pack_half2(a.z, a.w);
      *reinterpret_cast<uint2*>(p.d + row * p.ldd + col) = o;
    }
    return;
  }
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < slab;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / p.n;
    const int col = static_cast<int>(i - row * p.n);
    float v = 0.f;
    for (int s = 0; s < p.splits; ++s) v += p.ws[s * slab + i];
    if (p.bias != nullptr) {
      const long long brow = (p.bias_batch_stride != 0) ? (row / p.rows_per_batch) : 0;
      v += p.bias[brow * p.bias_batch_stride + col];
    }
    if (p.residual != nullptr) v += __half2float(p.residual[row * p.ldr + col]);
    p.d[row * p.ldd + col] = __float2half_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static int make_tmap_f16_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                            const uint32_t* elem_strides = nullptr) {
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
      return MDB_ERR_CUDA;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("TMA base address %p is not 16-byte aligned", base);
    return MDB_ERR_INVALID;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    if (gstr[i] % 16 != 0) {
      set_error("TMA stride %d (%llu bytes) is not a multiple of 16", i, (unsigned long long)gstr[i]);
      return MDB_ERR_INVALID;
    }
  }
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu,%llu box %u,%u)", (int)r, rank,
              (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0), bx[0], rank > 1 ? bx[1] : 0);
    return MDB_ERR_CUDA;
  }
  return MDB_OK;
}

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  return make_tmap_f16_sw(out, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int make_tmap_f16_plain(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap_f16_sw(out, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

void count_launch(int n = 1);

// launch heuristics (mdb_set_tuning): defaults selected by the B200 measurements under profiles/
static int g_pair_min_tiles = 128;  // smallest grid, in 128-row tile equivalents, that goes to gemm_pair_kernel
static int g_bn80_below = 100;      // N % 160 == 0 layers with fewer 160-wide CTAs than this use 80-wide tiles
constexpr int kLongKChunks = 64;    // ... unless K >= 4096 (automatic split-K): then 160-wide tiles and split K

template <int BN, bool GEGLU, int STAGES>
static int launch_gemm(const GemmKParams& kp, dim3 grid, cudaStream_t st) {
  const unsigned cluster_z = kp.cluster_reduce ? static_cast<unsigned>(kp.splits) : 1u;
  static bool attr_set = false;
  auto kern = gemm_tc_kernel<BN, GEGLU, STAGES>;
  constexpr int kSmem = GemmSmem<BN, STAGES, false>::kTotal;
  if (!attr_set) {
    MDB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set = true;
  }
  MDB_CHECK_CUDA(launch_pdl_cluster(kern, grid, dim3(kGemmThreads), kSmem, st, cluster_z, kp));
  count_launch();
  return MDB_OK;
}

template <int BN, bool GEGLU, int STAGES>
static int launch_gemm_pair(const GemmKParams& kp, const CUtensorMap& tmD, int total_tiles, cudaStream_t st) {
  static bool attr_set = false;
  auto kern = gemm_pair_kernel<BN, GEGLU, STAGES>;
  constexpr int kSmem = PairqSmem<BN, STAGES>::kTotal;
  static_assert(kSmem <= 227 * 1024, "shared memory budget");
  if (!attr_set) {
    MDB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set = true;
  }
  const int clusters = total_tiles < 74 ? total_tiles : 74;  // one pair per TPC (148 SMs)
  // no programmatic stream serialization: the pair kernel starts only after its predecessor has completed
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kPairqThreads);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MDB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, kp, tmD));
  count_launch();
  return MDB_OK;
}

int get_gemm_tuning(int key) { return key == MDB_TUNE_GEMM_PAIR_MIN_TILES ? g_pair_min_tiles : g_bn80_below; }
void set_gemm_tuning(int key, int value) {
  if (key == MDB_TUNE_GEMM_PAIR_MIN_TILES) g_pair_min_tiles = value;
  else g_bn80_below = value;
}

}  // namespace mdb

using namespace mdb;

extern "C" int mdb_gemm_f16(const mdb_gemm_desc* g, mdb_stream_t stream) {
  MDB_REQUIRE(g != nullptr, "mdb_gemm_f16: null descriptor");
  MDB_REQUIRE(g->m > 0 && g->n > 0 && g->k > 0, "mdb_gemm_f16: bad shape m=%d n=%d k=%d", g->m, g->n, g->k);
  MDB_REQUIRE(g->k % kBK == 0, "mdb_gemm_f16: K=%d must be a multiple of 64", g->k);
  MDB_REQUIRE(g->a && g->b && g->d, "mdb_gemm_f16: null operand");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool geglu = g->epilogue == MDB_EPI_GEGLU;
  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.d = static_cast<__half*>(g->d);
  kp.ldd = g->ldd;
  kp.bias = g->bias;
  kp.bias_batch_stride = g->bias_batch_stride;
  kp.rows_per_batch = g->rows_per_batch > 0 ? g->rows_per_batch : 1;
  kp.residual = static_cast<const __half*>(g->residual);
  kp.ldr = g->ldr;
  kp.m = g->m;
  kp.n = g->n;
  kp.k_chunks = g->k / kBK;
  kp.conv = g->conv;
  MDB_REQUIRE(g->ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(g->d) & 15) == 0,
              "mdb_gemm_f16: D must be 16B aligned with ldd %% 8 == 0 (ldd=%lld)", (long long)g->ldd);
  if (g->bias) {
    MDB_REQUIRE((reinterpret_cast<uintptr_t>(g->bias) & 15) == 0 && g->bias_batch_stride % 4 == 0,
                "mdb_gemm_f16: bias must be 16B aligned with bias_batch_stride %% 4 == 0");
  }
  if (g->residual) {
    MDB_REQUIRE(g->ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(g->residual) & 15) == 0,
                "mdb_gemm_f16: residual must be 16B aligned with ldr %% 8 == 0");
    MDB_REQUIRE(!geglu, "mdb_gemm_f16: residual is not supported with the GEGLU epilogue");
  }

  if (g->ln_u != nullptr) {
    MDB_REQUIRE(!geglu, "mdb_gemm_f16: LayerNorm fusion is not available with the GEGLU epilogue (N / 128 CTAs per row "
                        "block would each recompute the row statistics: measured slower than the LayerNorm kernel)");
    MDB_REQUIRE(!g->conv && g->a2 == nullptr && (reinterpret_cast<uintptr_t>(g->ln_u) & 15) == 0,
                "mdb_gemm_f16: LayerNorm fusion needs a plain single-source A whose K is the normalised width");
    kp.ln_u = g->ln_u;
    kp.ln_eps = g->ln_eps;
  }
  int rc;
  if (g->conv) {
    MDB_REQUIRE(g->a2 == nullptr, "mdb_gemm_f16: conv mode takes a single source");
    MDB_REQUIRE(g->c % kBK == 0 && g->k == 9 * g->c, "mdb_gemm_f16: conv needs c %% 64 == 0 and k == 9c (c=%d k=%d)",
                g->c, g->k);
    // conv == 1: stride 1; conv == 2: stride 2 (Downsample.op, openaimodel.py:175): the output pixel (y, x) reads input
    // (2y + kh - 1, 2x + kw - 1) — the same shifted boxes with TMA element strides of 2 along w and h, no im2col buffer
    const int cs = g->conv == 2 ? 2 : 1;
    const int ho = (g->h - 1) / cs + 1, wo = (g->w - 1) / cs + 1;
    MDB_REQUIRE(g->m == g->nb * ho * wo, "mdb_gemm_f16: conv m != nb*ho*wo");
    const int hw = ho * wo;  // OUTPUT pixels per image: tiles are cut over these
    uint32_t box[4];
    if (wo > kBM) {
      // rows wider than the tile (the VAE's 256- and 512-pixel levels): a tile is 128 consecutive pixels of ONE row
      MDB_REQUIRE(wo % kBM == 0 && cs == 1, "mdb_gemm_f16: conv rows wider than 128 pixels need 128 | w and stride 1 (w=%d)", g->w);
      box[0] = kBK; box[1] = kBM; box[2] = 1; box[3] = 1;
    } else if (hw >= kBM) {
      MDB_REQUIRE(kBM % wo == 0 && hw % kBM == 0,
                  "mdb_gemm_f16: conv tile needs w | 128 and 128 | h*w (output h=%d w=%d)", ho, wo);
      box[0] = kBK; box[1] = cs * wo; box[2] = cs * (kBM / wo); box[3] = 1;
    } else {
      MDB_REQUIRE(kBM % hw == 0, "mdb_gemm_f16: conv tile needs h*w | 128 (output h=%d w=%d)", ho, wo);
      box[0] = kBK; box[1] = cs * wo; box[2] = cs * ho; box[3] = kBM / hw;
    }
    MDB_REQUIRE(box[1] <= 256 && box[2] <= 256, "mdb_gemm_f16: conv TMA box too large");
    uint64_t dims[4] = {(uint64_t)g->c, (uint64_t)g->w, (uint64_t)g->h, (uint64_t)g->nb};
    uint64_t str[3] = {(uint64_t)g->c * 2, (uint64_t)g->c * g->w * 2, (uint64_t)g->c * g->h * g->w * 2};
    const uint32_t estr[4] = {1u, (uint32_t)cs, (uint32_t)cs, 1u};
    rc = make_tmap_f16_sw(&kp.tmA, g->a, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, cs == 2 ? estr : nullptr);
    if (rc) return rc;
    kp.chunks_per_tap = g->c / kBK;
    kp.w = wo;
    kp.hw = hw;
    kp.cs = cs;
    kp.k1_chunks = kp.k_chunks;
  } else {
    const int k1 = g->a2 ? g->k1 : g->k;
    MDB_REQUIRE(k1 % kBK == 0 && k1 > 0 && k1 <= g->k, "mdb_gemm_f16: k1=%d must be a multiple of 64 within K", k1);
    uint32_t box[2] = {kBK, kBM};
    uint64_t dims[2] = {(uint64_t)k1, (uint64_t)g->m};
    uint64_t str[1] = {(uint64_t)g->lda * 2};
    rc = make_tmap_f16(&kp.tmA, g->a, 2, dims, str, box);
    if (rc) return rc;
    if (g->a2) {
      uint64_t dims2[2] = {(uint64_t)(g->k - k1), (uint64_t)g->m};
      uint64_t str2[1] = {(uint64_t)g->lda2 * 2};
      rc = make_tmap_f16(&kp.tmA2, g->a2, 2, dims2, str2, box);
      if (rc) return rc;
    }
    kp.k1_chunks = k1 / kBK;
  }

  // ---- which kernel ----
  // Large grids (>= g_pair_min_tiles 128-row tile equivalents, no split-K, at least two M tiles): the persistent
  // CTA-pair kernel with 256 x {320, 256, 160, 128} tiles — widest first: fewest L2 -> SM bytes per flop; the
  // 320-wide tile (two 160-wide MMAs per K step, single accumulator buffer) only for long K.
  const int m_tiles = (g->m + kBM - 1) / kBM;
  bool pair = g->splits <= 1 && m_tiles >= 2 && g->n % 8 == 0 && g->ln_u == nullptr;  // (the pair kernel has no LN fusion)
  int bn = 0;
  if (pair) {
    if (geglu) bn = (g->n % 256 == 0) ? 256 : 0;
    else if (g->n % 320 == 0 && kp.k_chunks >= 16) bn = 320;
    else if (g->n % 256 == 0) bn = 256;
    else if (g->n % 160 == 0) bn = 160;
    else bn = 128;
    // ... and enough work per launch: the pair kernel owns its SMs (one ~200 KB CTA each, no PDL, nothing of the other
    // stream beside it), which only pays when the grid is large AND the K loop is not a handful of chunks
    // (measured: 8192x320x320 at one frame is faster on the single-CTA tiles, 4096x1280x1280 at eight on the pair)
    const long long eq = (long long)((m_tiles + 1) / 2) * 2 * ((g->n + bn - 1) / (bn ? bn : 1));
    if (bn == 0 || eq < (long long)g_pair_min_tiles || eq * kp.k_chunks < 16ll * g_pair_min_tiles) pair = false;
  }
  if (pair) {
    // bn chosen above
  } else if (geglu) {
    MDB_REQUIRE(g->n % 128 == 0, "mdb_gemm_f16: GEGLU needs N %% 128 == 0 (N=%d)", g->n);
    bn = 128;
  } else if (g->n % 160 == 0) {
    // 160-wide tiles unless that leaves most of the 148 SMs idle; then (short K) halve the tile width, or (long K:
    // the weight-streaming 3x3 convs of the 8x8 ... 32x32 levels at one frame) keep the wide tile and split K —
    // see auto_splits below and profiles/r02_deepk_microbench.md
    const long long tiles160 = (long long)m_tiles * (g->n / 160) * (g->splits > 1 ? g->splits : 1);
    const bool wide_split = g->splits == 0 && kp.k_chunks >= kLongKChunks && tiles160 * 2 <= 148;
    bn = (tiles160 < g_bn80_below && !wide_split) ? 80 : 160;
  } else {
    bn = 128;
  }
  {
    // a pair CTA stages half of the B rows (of each 160-wide part of a 320-wide tile)
    uint32_t box[2] = {kBK, (uint32_t)(pair ? (bn == 320 ? 80 : bn / 2) : bn)};
    uint64_t dims[2] = {(uint64_t)g->k, (uint64_t)g->n};
    uint64_t str[1] = {(uint64_t)g->ldb * 2};
    rc = make_tmap_f16(&kp.tmB, g->b, 2, dims, str, box);
    if (rc) return rc;
  }

  int splits = g->splits > 1 ? g->splits : 1;
  if (g->ln_u != nullptr) splits = 1;  // the correction is applied by the CTA that holds the whole K range
  if (g->splits == 0 && !geglu && !pair && g->ln_u == nullptr) {
    // automatic split-K: a power of two up to 8 (reduced inside a thread-block cluster through DSMEM) that brings
    // the grid to about one CTA per SM while every split keeps at least 16 K chunks (measured on B200 with cold
    // weights: below that the cluster reduction costs more than the extra CTAs gain)
    const long long tiles = (long long)m_tiles * ((g->n + bn - 1) / bn);
    if (tiles < 100)
      for (int c = 8; c >= 2; c >>= 1)
        if (tiles * c <= 148 && kp.k_chunks / c >= 16) {
          splits = c;
          break;
        }
  }
  if (splits > kp.k_chunks) splits = kp.k_chunks;
  if (geglu || pair) splits = 1;
  kp.chunks_per_split = (kp.k_chunks + splits - 1) / splits;
  splits = (kp.k_chunks + kp.chunks_per_split - 1) / kp.chunks_per_split;  // no empty splits
  kp.splits = splits;
  kp.ws = g->splitk_ws;
  // 2, 4 or 8 splits: the partners form a thread-block cluster and reduce through DSMEM (one kernel);
  // other counts go through the global fp32 workspace + finalize kernel.
  kp.cluster_reduce = (!geglu && (splits == 2 || splits == 4 || splits == 8)) ? 1 : 0;
  if (splits > 1 && !kp.cluster_reduce) {
    MDB_REQUIRE(g->splitk_ws != nullptr, "mdb_gemm_f16: splits > 1 needs splitk_ws");
  }

  dim3 grid(m_tiles, (g->n + bn - 1) / bn, splits);
  const bool deep = (long long)grid.x * grid.y * grid.z <= 148 && kp.chunks_per_split >= 12;
  if (pair) {
    // output tensor map of the TMA-store epilogue: [M][N_out] fp16, 32 x 32 boxes, dense (no swizzle)
    CUtensorMap tmD;
    uint32_t boxd[2] = {(uint32_t)kOutBox, (uint32_t)kOutBox};
    uint64_t dimsd[2] = {(uint64_t)(geglu ? g->n / 2 : g->n), (uint64_t)g->m};
    uint64_t strd[1] = {(uint64_t)g->ldd * 2};
    rc = make_tmap_f16_plain(&tmD, g->d, 2, dimsd, strd, boxd);
    if (rc) return rc;
    const int total_tiles = ((m_tiles + 1) / 2) * (int)grid.y;
    if (geglu) return launch_gemm_pair<256, true, 5>(kp, tmD, total_tiles, st);
    if (bn == 320) return launch_gemm_pair<320, false, 5>(kp, tmD, total_tiles, st);
    if (bn == 160) return launch_gemm_pair<160, false, 6>(kp, tmD, total_tiles, st);
    if (bn == 256) return launch_gemm_pair<256, false, 5>(kp, tmD, total_tiles, st);
    return launch_gemm_pair<128, false, 6>(kp, tmD, total_tiles, st);
  }
  if (geglu) rc = launch_gemm<128, true, 3>(kp, grid, st);
  else if (bn == 160) rc = deep ? launch_gemm<160, false, 6>(kp, grid, st) : launch_gemm<160, false, 3>(kp, grid, st);
  else if (bn == 80) rc = deep ? launch_gemm<80, false, 8>(kp, grid, st) : launch_gemm<80, false, 3>(kp, grid, st);
  else rc = deep ? launch_gemm<128, false, 6>(kp, grid, st) : launch_gemm<128, false, 3>(kp, grid, st);
  if (rc) return rc;
  if (splits > 1 && !kp.cluster_reduce) {
    const long long total = ((long long)g->m * g->n + 3) / 4;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    MDB_CHECK_CUDA(launch_pdl(splitk_finalize_kernel, dim3(blocks), dim3(256), 0, st, kp));
    count_launch();
  }
  return MDB_OK;
}
