// One tcgen05 kernel for the three gene-wide Dense products whose reduction or output dimension is
// the gene axis (all with a 64-wide partner dimension):
//
//   encoder forward  (K1)  A1[B x 64]  += X[B x G] . W1[G x 64]                      DO_B
//   head backward    (K4)  dWh[64 x G] += H3^T . dZ,  dH3[B x 64] += dZ . Wh^T,  db = colsum(dZ)   DO_A + DO_B + COLSUM
//   encoder backward (K5)  dW1[G x 64] += X^T . dA1                                   DO_A
//
// "Z" is the cells x genes bf16 operand (X or dZ).  A 128-cell x 128-gene tile of Z is brought into
// shared memory ONCE by TMA (two SWIZZLE_128B boxes of 64 genes) and feeds both products:
//   (b) as a K-major  A operand (M = 128 cells, K = genes) against W  [64 x genes]  (K-major B), and
//   (a) as an MN-major A operand (M = 128 genes, K = cells) against H [cells x 64] (MN-major B),
// i.e. the same bytes with two different UMMA descriptors, so Z is read from HBM once per step for
// both gradients.  Accumulators live in TMEM: one 128x64 fp32 tile per gene block
// of the item for (a), a double-buffered 128x64 tile per cell block for (b).  (b) leaves through a
// swizzled staging tile and a TMA reduce-add (split over gene ranges); (a) is added to the gradient
// buffer by the epilogue warps at the end of the item; the per-gene column sums (bias gradient) are one
// more tcgen05 product of the same tile with a constant ones operand (Z^T . 1, N = 16).
//
// Warp roles (256 threads): 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-7 epilogue.
#include <cstdio>
#include <vector>
#include "engine.h"
#include "tc_common.cuh"

namespace dca {
namespace tc {
int g_gg_prefetch = 0;      // dca_set_tunable("gg_prefetch", 0 | 1)
int g_gg_profile = 0;       // dca_set_tunable("gg_profile", 1): per-role wait-time counters, printed after every launch (diagnosis; DCA_GRAPH=0)
int g_gg_flat = 1;          // dca_set_tunable("gg_flat", 0 | 1): flat unit partition of the backward kernels (every SM busy)
namespace gg {

constexpr int kThreads = 256;
constexpr int kZStages = 3;
constexpr uint32_t kZBytes = 128 * 128 * 2;          // 2 boxes x [128 cells x 64 genes] bf16
constexpr uint32_t kWBytes = 64 * 128 * 2;           // 2 boxes x [64 feats x 64 genes] bf16
constexpr uint32_t kHBytes = 128 * 64 * 2;           // [128 cells x 64 feats] bf16
constexpr uint32_t kOutBytes = 2 * 128 * 32 * 4;     // two [128 x 32] fp32 staging tiles
constexpr int kMaxGb = 4;                            // gene blocks (of 128) per item: 4 x 64 TMEM columns
constexpr uint32_t kTmemCols = 512;

struct Params {
  int B, G, n_heads;
  int n_cb, n_gb;                 // cell blocks (128), gene blocks per head (128)
  int gb_per_item, cb_per_item;
  int gene_ranges, cell_splits;   // per head
  int total_items;
  unsigned long long* dbg;        // gg_profile: [grid][16] cycle counters (nullptr = off)
  int flat, total_units;          // flat: units = (head, gene range, cell block) in that order, an equal contiguous run per CTA
  float* dW[3]; int64_t dW_ld; int dW_transposed;   // (a): transposed -> dW[f*ld + g] (Keras [64 x G]); else dW[g*ld + f]
  float* db[3];                                     // column sums of Z per head (COLSUM)
  const __nv_bfloat16* Zp[3]; int64_t ldz; int prefetch;   // raw Z pointers: L2 prefetch of whole row segments ahead of the TMA boxes
};

__device__ __forceinline__ void gg_prefetch_l2(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(gptr)), "r"(bytes) : "memory");
}

// diagnosis: accumulate the cycles a single-thread role spends inside one wait / section into its counter slot
#define GG_TIMED(slot, stmt)                                             \
  do {                                                                   \
    if (p.dbg) { const long long t0_ = clock64(); stmt; dbg_acc[slot] += (unsigned long long)(clock64() - t0_); } \
    else { stmt; }                                                       \
  } while (0)

template <bool DO_A, bool DO_B, bool COLSUM>
__global__ void __launch_bounds__(kThreads, 1)
gene_gemm_kernel(const __grid_constant__ CUtensorMap map_z0, const __grid_constant__ CUtensorMap map_z1,
                 const __grid_constant__ CUtensorMap map_z2, const __grid_constant__ CUtensorMap map_h,
                 const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_w1,
                 const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_o, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET (not by an integer round trip): the pointer keeps the shared address space, so the
  // staging stores / bias loads compile to STS / LDS instead of generic ST / LD
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr uint32_t kStage = kZBytes + (DO_B ? kWBytes : 0);
  uint8_t* s_z = smem;                                          // [kZStages][Z | W]
  uint8_t* s_h = s_z + kZStages * kStage;                       // [2][H]
  uint8_t* s_o = s_h + (DO_A ? 2 * kHBytes : 0);                // staging for (b)
  // COLSUM: a [128 cells x 64] bf16 tile of ones in H's layout -- the second N chunk of the dW product's B operand (N = 64 + 16):
  // the 16 extra accumulator columns are the column sums of Z (every element equal, so the swizzle pattern does not matter)
  uint8_t* s_ones = s_o + ((DO_B || (DO_A && !DO_B)) ? kOutBytes : 0);
  __shared__ uint64_t z_full[kZStages], z_empty[kZStages], h_full[2], h_empty[2], dh_full[2], dh_empty[2], dw_full, dw_empty;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_unit;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    s_unit = -1;
    for (int i = 0; i < kZStages; ++i) { mbar_init(&z_full[i], 1); mbar_init(&z_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&h_full[i], 1); mbar_init(&h_empty[i], 1); mbar_init(&dh_full[i], 1); mbar_init(&dh_empty[i], 4); }
    mbar_init(&dw_full, 1); mbar_init(&dw_empty, 4);
    fence_barrier_init();
    tma_prefetch_desc(&map_z0); tma_prefetch_desc(&map_h); tma_prefetch_desc(&map_w0); tma_prefetch_desc(&map_o);
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, kTmemCols);
  if (COLSUM) {   // bf16 1.0 = 0x3F80
    for (int i = threadIdx.x; i < (int)kHBytes / 4; i += kThreads) reinterpret_cast<uint32_t*>(s_ones)[i] = 0x3F803F80u;
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  constexpr uint32_t kDwCols = COLSUM ? 80 : 64;   // per gene block: 64 dW columns (+ 16 column-sum columns)
  const uint32_t tm_dw = tmem;                 // kMaxGb x kDwCols columns
  const uint32_t tm_dh = tmem + kMaxGb * kDwCols;   // 2 x 64 columns

  struct Item { int head, gb0, gb1, cb0, cb1; };
  auto decode = [&](int it) {
    Item x;
    const int cs = it % p.cell_splits; const int r = it / p.cell_splits;
    const int gr = r % p.gene_ranges; x.head = r / p.gene_ranges;
    x.gb0 = gr * p.gb_per_item; x.gb1 = min(p.n_gb, x.gb0 + p.gb_per_item);
    x.cb0 = cs * p.cb_per_item; x.cb1 = min(p.n_cb, x.cb0 + p.cb_per_item);
    return x;
  };
  // Item sequence of this CTA.  Strided: items blockIdx.x, +gridDim.x, ...  Flat: the CTA owns units [u0, u1) of the
  // (head, gene range, cell block) order; an item is the part of ONE gene range inside that run, so that a CTA switches
  // dW accumulators (and flushes them) at most (run length / n_cb) + 1 times and every SM gets the same number of tiles.
  struct Cursor { int u, u1; };
  auto first = [&]() {
    Cursor c;
    if (p.flat) { c.u = (int)((long long)blockIdx.x * p.total_units / gridDim.x); c.u1 = (int)((long long)(blockIdx.x + 1) * p.total_units / gridDim.x); }
    else { c.u = blockIdx.x; c.u1 = p.total_items; }
    return c;
  };
  auto next = [&](Cursor& c, Item& x) {
    if (c.u >= c.u1) return false;
    if (!p.flat) { x = decode(c.u); c.u += gridDim.x; return true; }
    const int r = c.u / p.n_cb, cb0 = c.u % p.n_cb, n = min(p.n_cb - cb0, c.u1 - c.u);
    const int gr = r % p.gene_ranges; x.head = r / p.gene_ranges;
    x.gb0 = gr * p.gb_per_item; x.gb1 = min(p.n_gb, x.gb0 + p.gb_per_item);
    x.cb0 = cb0; x.cb1 = cb0 + n; c.u += n;
    return true;
  };

  if (warp == 0) {
    // ===================================================== TMA producer (whole warp walks the loops, one elected lane issues)
    const bool leader = elect_one();
    {
      unsigned long long dbg_acc[4] = {0, 0, 0, 0};          // 0 wait z_empty, 1 wait h_empty, 2 total
      const long long t_begin = clock64();
      uint32_t zi = 0, hi = 0;
      int unit = 0;
      Cursor cur = first(); Item x;
      while (next(cur, x)) {
        const CUtensorMap* mz = x.head == 0 ? &map_z0 : (x.head == 1 ? &map_z1 : &map_z2);
        for (int cb = x.cb0; cb < x.cb1; ++cb) {
          if (DO_A) {
            const uint32_t hs = hi & 1, hp = (hi >> 1) & 1; ++hi;
            GG_TIMED(1, mbar_wait(&h_empty[hs], hp ^ 1));
            if (leader) { mbar_expect_tx(&h_full[hs], kHBytes); tma_load_2d(s_h + hs * kHBytes, &map_h, 0, cb * 128, &h_full[hs]); }
          }
          for (int gb = x.gb0; gb < x.gb1; ++gb) {
            if (((gb - x.gb0) & 3) == 0) { if (leader) *reinterpret_cast<volatile int*>(&s_unit) = unit; ++unit; }      // progress mark for the prefetch warp
            const uint32_t st = zi % kZStages, ph = (zi / kZStages) & 1; ++zi;
            GG_TIMED(0, mbar_wait(&z_empty[st], ph ^ 1));
            uint8_t* dst = s_z + st * kStage;
            if (leader) {
              mbar_expect_tx(&z_full[st], kStage);
              tma_load_2d(dst, mz, gb * 128, cb * 128, &z_full[st]);
              tma_load_2d(dst + kZBytes / 2, mz, gb * 128 + 64, cb * 128, &z_full[st]);
              if (DO_B) {
                const CUtensorMap* mw = x.head == 0 ? &map_w0 : (x.head == 1 ? &map_w1 : &map_w2);
                if (DO_A) {      // head backward: W = Keras [64 x G] (K-major B): two [64 feats x 64 genes] boxes
                  tma_load_2d(dst + kZBytes, mw, gb * 128, 0, &z_full[st]);
                  tma_load_2d(dst + kZBytes + kWBytes / 2, mw, gb * 128 + 64, 0, &z_full[st]);
                } else {         // encoder forward: W = Keras [G x 64] (MN-major B): one [128 genes x 64 feats] box
                  tma_load_2d(dst + kZBytes, mw, 0, gb * 128, &z_full[st]);
                }
              }
            }
          }
        }
      }
      if (p.dbg && leader) { dbg_acc[2] = (unsigned long long)(clock64() - t_begin); for (int i = 0; i < 3; ++i) p.dbg[blockIdx.x * 16 + i] = dbg_acc[i]; }
    }
  } else if (warp == 3 && p.prefetch) {
    // ===================================================== L2 prefetch (optional, dca_set_tunable "gg_prefetch")
    // A SWIZZLE_128B box is 128 rows x 128 bytes, i.e. every box touches 128 DRAM pages for 128 bytes each.  This warp
    // walks the producer's sequence ONE unit ahead (unit = up to four gene blocks of one cell block = one contiguous
    // segment of up to 1 KB per row) and prefetches the unit's row segments into L2 with bulk prefetches (lane = row
    // mod 32), paced by the producer's progress mark in shared memory.
    int unit = 0;
    auto prefetch_unit = [&](int head, int cb, int g0, int g1) {              // genes [g0*128, g1*128) of cell block cb
      const int c0 = g0 * 128, c1 = min(g1 * 128, p.G);
      if (c1 <= c0) return;
      const uint32_t bytes = (uint32_t)(c1 - c0) * 2u;
      const __nv_bfloat16* zp = p.Zp[head];
      for (int r = lane; r < 128; r += 32) {
        const int row = cb * 128 + r;
        if (row < p.B) gg_prefetch_l2(zp + (int64_t)row * p.ldz + c0, bytes);
      }
    };
    Cursor cur = first(); Item x;
    while (next(cur, x)) {
      for (int cb = x.cb0; cb < x.cb1; ++cb)
        for (int g4 = x.gb0; g4 < x.gb1; g4 += 4, ++unit) {
          if (unit > 0) {                                                      // unit 0 is not worth prefetching (its loads are already in flight)
            long long spins = 0;
            while (*reinterpret_cast<volatile int*>(&s_unit) < unit - 1) { __nanosleep(64); if (++spins > (1ll << 26)) break; }
            prefetch_unit(x.head, cb, g4, min(g4 + 4, x.gb1));
          }
        }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    // The WHOLE warp walks the loops (so that smem addresses, descriptors and barrier phases are warp-uniform values the
    // compiler keeps in uniform registers); only the tcgen05.mma / commit instructions are issued by one lane.  With the
    // loops inside `if (lane == 0)` every descriptor went through a R2UR waterfall loop: ~75 cycles per MMA, 1800 per tile
    // of the head backward (24 MMAs) -- the issuing thread, not HBM, bounded the kernel (gg_profile).
    const bool leader = elect_one();
    {
      constexpr uint32_t idesc_b = make_idesc_bf16(128, 64, 0, DO_A ? 0 : 1);   // Z K-major x W (K-major [64xG] | MN-major [Gx64])
      constexpr uint32_t idesc_a = make_idesc_bf16(128, kDwCols, 1, 1);     // Z MN-major x [H | ones] MN-major
      uint32_t zi = 0, hi = 0, di = 0, wi = 0;
      unsigned long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // 0 wait z_full, 1 wait h_full, 2 wait dh_empty, 3 wait dw_empty, 4 total, 5 tiles
      const long long t_begin = clock64();
      Cursor cur = first(); Item x;
      for (; next(cur, x); ++wi) {
        if (DO_A) { GG_TIMED(3, mbar_wait(&dw_empty, (wi & 1) ^ 1)); tcgen05_fence_after(); }
        for (int cb = x.cb0; cb < x.cb1; ++cb) {
          uint32_t hs = 0, ds = 0;
          if (DO_A) { hs = hi & 1; const uint32_t hp = (hi >> 1) & 1; ++hi; GG_TIMED(1, mbar_wait(&h_full[hs], hp)); }
          if (DO_B) { ds = di & 1; const uint32_t dp = (di >> 1) & 1; ++di; GG_TIMED(2, mbar_wait(&dh_empty[ds], dp ^ 1)); }
          for (int gb = x.gb0; gb < x.gb1; ++gb) {
            const uint32_t st = zi % kZStages, ph = (zi / kZStages) & 1; ++zi;
            GG_TIMED(0, mbar_wait(&z_full[st], ph));
            ++dbg_acc[5];
            tcgen05_fence_after();
            const uint32_t zb = smem_u32(s_z + st * kStage);
            if (DO_B) {
              const uint32_t wb = zb + kZBytes;
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (leader) umma_bf16(tm_dh + ds * 64, make_smem_desc(zb + h * (kZBytes / 2) + k * 32, 0, 1024),
                            DO_A ? make_smem_desc(wb + h * (kWBytes / 2) + k * 32, 0, 1024)
                                 : make_smem_desc(wb + (h * 4 + k) * 2048, 0, 1024), idesc_b,
                            (gb > x.gb0 || h > 0 || k > 0) ? 1u : 0u);
            }
            if (DO_A) {
              const uint32_t hb = smem_u32(s_h + hs * kHBytes);
              // COLSUM: N = 80 -- the B operand's second 64-element chunk (leading-dimension offset) is the ones tile, so the
              // same MMA that forms dW also leaves db[g] = sum over the tile's cells of Z[cell][g] in 16 more columns
              const uint32_t ones_off = smem_u32(s_ones) - hb;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (leader) umma_bf16(tm_dw + (gb - x.gb0) * kDwCols, make_smem_desc(zb + k * 2048, kZBytes / 2, 1024),
                          make_smem_desc(hb + k * 2048, COLSUM ? ones_off : 0u, 1024), idesc_a, (cb > x.cb0 || k > 0) ? 1u : 0u);
            }
            if (leader) umma_commit(&z_empty[st]);
          }
          if (DO_A) if (leader) umma_commit(&h_empty[hs]);
          if (DO_B) if (leader) umma_commit(&dh_full[ds]);
        }
        if (DO_A) if (leader) umma_commit(&dw_full);
      }
      if (p.dbg && leader) { dbg_acc[4] = (unsigned long long)(clock64() - t_begin); for (int i = 0; i < 6; ++i) p.dbg[blockIdx.x * 16 + 4 + i] = dbg_acc[i]; }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================================================== epilogue
    const int quarter = warp & 3;
    // ONE elected lane of warp 4 issues every TMA reduce-add and owns their bulk groups (commit / wait are per thread)
    bool epi_leader = false;
    if (warp == 4) epi_leader = elect_one();
    uint32_t di = 0, wi = 0;
    unsigned long long dbg_acc[6] = {0, 0, 0, 0, 0, 0};       // 0 wait dh_full, 1 dH flush, 2 wait dw_full, 3 dW flush, 4 total
    const long long t_begin = clock64();
    Cursor cur = first(); Item x;
    for (; next(cur, x); ++wi) {
      if (DO_B) {
        for (int cb = x.cb0; cb < x.cb1; ++cb) {
          const uint32_t ds = di & 1, dp = (di >> 1) & 1; ++di;
          GG_TIMED(0, mbar_wait(&dh_full[ds], dp));
          const long long t_fl = clock64();
          tcgen05_fence_after();
          if (epi_leader) bulk_wait_read<0>();          // previous reduce has read the staging tiles
          named_barrier_sync(3, 128);
          const int row = quarter * 32 + lane;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tm_dh + ((uint32_t)(quarter * 32) << 16) + ds * 64 + c * 32, v);
            tmem_ld_wait();
            uint8_t* tile = s_o + c * (kOutBytes / 2);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<uint4*>(tile + row * 128 + ((q ^ (row & 7)) << 4)) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          }
          tcgen05_fence_before();
          fence_proxy_async_smem();
          named_barrier_sync(3, 128);
          if (lane == 0) mbar_arrive(&dh_empty[ds]);
          if (epi_leader) {
            tma_reduce_add_2d(&map_o, 0, cb * 128, s_o);
            tma_reduce_add_2d(&map_o, 32, cb * 128, s_o + kOutBytes / 2);
            bulk_commit();
          }
          dbg_acc[1] += (unsigned long long)(clock64() - t_fl);
        }
      }
      if (DO_A) {
        GG_TIMED(2, mbar_wait(&dw_full, wi & 1));
        const long long t_fw = clock64();
        tcgen05_fence_after();
        float* dst = p.dW[x.head];
        if (!DO_B && !p.dW_transposed) {
          // [128 genes x 64] accumulator == row-major block of dW[G x 64]: swizzled staging + TMA reduce-add
          for (int gb = x.gb0; gb < x.gb1; ++gb) {
            if (epi_leader) bulk_wait_read<0>();
            named_barrier_sync(3, 128);
            const int row = quarter * 32 + lane;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t v[32];
              tmem_ld_32x32(tm_dw + ((uint32_t)(quarter * 32) << 16) + (gb - x.gb0) * kDwCols + c * 32, v);
              tmem_ld_wait();
              uint8_t* tile = s_o + c * (kOutBytes / 2);
#pragma unroll
              for (int q = 0; q < 8; ++q)
                *reinterpret_cast<uint4*>(tile + row * 128 + ((q ^ (row & 7)) << 4)) = make_uint4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            }
            fence_proxy_async_smem();
            named_barrier_sync(3, 128);
            if (epi_leader) {
              tma_reduce_add_2d(&map_o, 0, gb * 128, s_o);
              tma_reduce_add_2d(&map_o, 32, gb * 128, s_o + kOutBytes / 2);
              bulk_commit();
            }
          }
        } else
        for (int gb = x.gb0; gb < x.gb1; ++gb) {
          const int g = gb * 128 + quarter * 32 + lane;
          if (COLSUM) {
            uint32_t v[32];
            tmem_ld_32x32(tm_dw + ((uint32_t)(quarter * 32) << 16) + (gb - x.gb0) * kDwCols + 64, v);
            tmem_ld_wait();
            if (g < p.G && p.db[x.head]) atomicAdd(p.db[x.head] + g, __uint_as_float(v[0]));
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tm_dw + ((uint32_t)(quarter * 32) << 16) + (gb - x.gb0) * kDwCols + c * 32, v);
            tmem_ld_wait();
            if (g < p.G) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int f = c * 32 + j;
                float* a = p.dW_transposed ? dst + (int64_t)f * p.dW_ld + g : dst + (int64_t)g * p.dW_ld + f;
                atomicAdd(a, __uint_as_float(v[j]));
              }
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dw_empty);
        dbg_acc[3] += (unsigned long long)(clock64() - t_fw);
      }
    }
    if (epi_leader) bulk_wait<0>();
    if (p.dbg && epi_leader) { dbg_acc[4] = (unsigned long long)(clock64() - t_begin); for (int i = 0; i < 5; ++i) p.dbg[blockIdx.x * 16 + 10 + i] = dbg_acc[i]; }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, kTmemCols);
}

template <bool DO_A, bool DO_B, bool COLSUM>
constexpr uint32_t smem_bytes() {
  return kZStages * (kZBytes + (DO_B ? kWBytes : 0)) + (DO_A ? 2 * kHBytes : 0) + kOutBytes + (COLSUM ? kHBytes : 0) + 1024;
}

}  // namespace gg

// Z: bf16 [B x G] per head (ldz elements); H: bf16 [B x 64]; W[i]: bf16 Keras-layout kernels -- [G x 64] for mode 1 (the
// encoder kernel), [64 x G] per head for mode 3; out_b: fp32 [B x 64] (+=).
// mode: 1 = DO_B (K1), 2 = DO_A (K5), 3 = DO_A|DO_B|COLSUM (K4).
int gene_gemm_tc(int mode, const __nv_bfloat16* const Z[3], int64_t ldz, int B, int G, int n_heads,
                 const __nv_bfloat16* H, const __nv_bfloat16* const W[3], float* out_b, float* const dW[3], int64_t dW_ld,
                 int dW_transposed, float* const db[3], int sm_count, cudaStream_t s) {
  using namespace gg;
  if (ldz % 8 != 0) { set_error("gene_gemm_tc: ldz must be a multiple of 8 (16-byte TMA stride)"); return DCA_ERR_BAD_ARG; }
  const bool do_a = mode & 2, do_b = mode & 1;
  CUtensorMap mz[3], mh, mw[3], mo;
  for (int i = 0; i < 3; ++i) DCA_TRY(make_tensor_map_2d(&mz[i], Z[i < n_heads ? i : 0], 2, 1, (uint64_t)B, (uint64_t)G, (uint64_t)ldz, 128, 64, 1));
  if (do_a) DCA_TRY(make_tensor_map_2d(&mh, H, 2, 1, (uint64_t)B, 64, 64, 128, 64, 1)); else mh = mz[0];
  if (do_b) {
    for (int i = 0; i < 3; ++i) {
      const __nv_bfloat16* w = W[i < n_heads ? i : 0];
      if (do_a) DCA_TRY(make_tensor_map_2d(&mw[i], w, 2, 1, 64, (uint64_t)G, (uint64_t)G, 64, 64, 1));
      else DCA_TRY(make_tensor_map_2d(&mw[i], w, 2, 1, (uint64_t)G, 64, 64, 128, 64, 1));
    }
    DCA_TRY(make_tensor_map_2d(&mo, out_b, 4, 0, (uint64_t)B, 64, 64, 128, 32, 1));
  } else {
    mw[0] = mw[1] = mw[2] = mz[0]; mo = mz[0];
    if (do_a && !dW_transposed) {
      if (dW_ld != 64) { set_error("gene_gemm_tc: non-transposed dW needs ld == 64"); return DCA_ERR_BAD_ARG; }
      DCA_TRY(make_tensor_map_2d(&mo, dW[0], 4, 0, (uint64_t)G, 64, 64, 128, 32, 1));
    }
  }
  Params p{};
  p.B = B; p.G = G; p.n_heads = n_heads;
  p.n_cb = cdiv(B, 128); p.n_gb = cdiv(G, 128);
  // Work decomposition: items = (head, gene range, cell range), processed by a persistent grid.
  if (do_a && do_b) {
    // head backward: gene ranges limited by TMEM (kMaxGb accumulators); cells split so that one wave of fat items fills the
    // SMs.  Search the small space for the shortest makespan in tile units: rounds x (tiles per item + one tile-equivalent
    // per cell block for the dH reduce-add the item issues + a flush / start-up term per gene block -- the (a) outputs
    // leave by atomics, every extra cell split repeats the dW flush, every extra gene range the dH traffic).
    int best_gpi = 1, best_splits = 1; long long best_cost = -1;
    for (int gpi = 1; gpi <= kMaxGb; ++gpi) {
      const int ranges = cdiv(p.n_gb, gpi) * n_heads;
      for (int splits = 1; splits <= (p.n_cb < 8 ? p.n_cb : 8); ++splits) {
        const int cbpi = cdiv(p.n_cb, splits), cs = cdiv(p.n_cb, cbpi);
        const long long items = (long long)ranges * cs, rounds = (items + sm_count - 1) / sm_count;
        const long long cost = rounds * ((long long)gpi * cbpi + cbpi + 2ll * gpi);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_gpi = gpi; best_splits = splits; }
      }
    }
    p.gb_per_item = best_gpi; p.gene_ranges = cdiv(p.n_gb, best_gpi);
    p.cb_per_item = cdiv(p.n_cb, best_splits); p.cell_splits = cdiv(p.n_cb, p.cb_per_item);
  } else if (do_a) {
    // encoder backward: the same search (per item: tiles + half a tile per cell block for the H tile it loads + the dW flush
    // by TMA reduce-add per gene block); TMEM holds up to kMaxGb accumulators, so one H tile serves four gene blocks
    int best_gpi = 1, best_splits = 1; long long best_cost = -1;
    for (int gpi = 1; gpi <= kMaxGb; ++gpi) {
      const int ranges = cdiv(p.n_gb, gpi) * n_heads;
      for (int splits = 1; splits <= (p.n_cb < 8 ? p.n_cb : 8); ++splits) {
        const int cbpi = cdiv(p.n_cb, splits), cs = cdiv(p.n_cb, cbpi);
        const long long items = (long long)ranges * cs, rounds = (items + sm_count - 1) / sm_count;
        const long long cost = rounds * (2ll * gpi * cbpi + cbpi + 4ll * gpi);      // in half tiles
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_gpi = gpi; best_splits = splits; }
      }
    }
    p.gb_per_item = best_gpi; p.gene_ranges = cdiv(p.n_gb, best_gpi);
    p.cb_per_item = cdiv(p.n_cb, best_splits); p.cell_splits = cdiv(p.n_cb, p.cb_per_item);
  } else {
    // encoder forward: one cell block per item, genes split so that there are ~4 items per SM
    p.cb_per_item = 1; p.cell_splits = p.n_cb;
    int gsplits = cdiv(4 * sm_count, p.n_cb); if (gsplits < 1) gsplits = 1; if (gsplits > p.n_gb) gsplits = p.n_gb;
    while (gsplits > 1 && cdiv(p.n_gb, gsplits) < 4) --gsplits;
    p.gb_per_item = cdiv(p.n_gb, gsplits); p.gene_ranges = cdiv(p.n_gb, p.gb_per_item);
  }
  p.total_items = p.gene_ranges * p.cell_splits * n_heads;
  if (do_a && g_gg_flat) {
    // Flat partition (backward kernels): units = (head, gene range of gpi blocks, cell block); every CTA takes an equal
    // contiguous run.  gpi by a small cost model in tile units per CTA: tiles + one per cell block (H tile / dH reduce-add)
    // + two per gene block for each dW flush (one per gene range the run touches).
    int best_gpi = 1; long long best_cost = -1;
    for (int gpi = 1; gpi <= kMaxGb; ++gpi) {
      const long long units = (long long)cdiv(p.n_gb, gpi) * n_heads * p.n_cb;
      const long long ctas = units < sm_count ? units : sm_count, per = (units + ctas - 1) / ctas;
      const long long cost = per * gpi + per + ((per + p.n_cb - 1) / p.n_cb + 1) * 2 * gpi;
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_gpi = gpi; }
    }
    p.flat = 1; p.gb_per_item = best_gpi; p.gene_ranges = cdiv(p.n_gb, best_gpi);
    p.cb_per_item = p.n_cb; p.cell_splits = 1;
    p.total_units = p.gene_ranges * n_heads * p.n_cb;
    p.total_items = p.total_units;          // grid size below
  }
  for (int i = 0; i < 3; ++i) { p.dW[i] = dW ? dW[i] : nullptr; p.db[i] = db ? db[i] : nullptr; p.Zp[i] = Z[i < n_heads ? i : 0]; }
  p.ldz = ldz; p.prefetch = p.flat ? 0 : g_gg_prefetch;
  if (g_gg_profile) {
    static unsigned long long* dbg_buf = nullptr;
    if (!dbg_buf) DCA_CUDA_OK(cudaMalloc(&dbg_buf, sizeof(unsigned long long) * 16 * 1024));
    DCA_CUDA_OK(cudaMemsetAsync(dbg_buf, 0, sizeof(unsigned long long) * 16 * 1024, s));
    p.dbg = dbg_buf;
  }
  p.dW_ld = dW_ld; p.dW_transposed = dW_transposed;
  const int grid = p.total_items < sm_count ? p.total_items : sm_count;
#define DCA_GG_LAUNCH(A, Bb, Cc)                                                                                       \
  do {                                                                                                                 \
    static bool attr = false;                                                                                          \
    constexpr uint32_t sm = smem_bytes<A, Bb, Cc>();                                                                   \
    if (!attr) { DCA_CUDA_OK(cudaFuncSetAttribute(gene_gemm_kernel<A, Bb, Cc>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); attr = true; } \
    gene_gemm_kernel<A, Bb, Cc><<<grid, kThreads, sm, s>>>(mz[0], mz[1], mz[2], mh, mw[0], mw[1], mw[2], mo, p);                         \
  } while (0)
  if (mode == 1) DCA_GG_LAUNCH(false, true, false);
  else if (mode == 2) DCA_GG_LAUNCH(true, false, false);
  else if (mode == 3) DCA_GG_LAUNCH(true, true, true);
  else { set_error("gene_gemm_tc: bad mode %d", mode); return DCA_ERR_BAD_ARG; }
#undef DCA_GG_LAUNCH
  DCA_LAUNCH_CHECK();
  if (p.dbg) {          // diagnosis only: synchronises the stream
    std::vector<unsigned long long> h((size_t)grid * 16);
    DCA_CUDA_OK(cudaStreamSynchronize(s));
    DCA_CUDA_OK(cudaMemcpy(h.data(), p.dbg, h.size() * 8, cudaMemcpyDeviceToHost));
    double a[16] = {0}; double mx_total = 0;
    for (int b = 0; b < grid; ++b) { for (int i = 0; i < 16; ++i) a[i] += (double)h[(size_t)b * 16 + i] / grid; if ((double)h[(size_t)b * 16 + 8] > mx_total) mx_total = (double)h[(size_t)b * 16 + 8]; }
    fprintf(stderr, "[gg_profile mode %d grid %d flat %d gpi %d] kcycles avg/CTA  producer: wait_z_empty %.0f wait_h_empty %.0f total %.0f | "
                    "mma: wait_z_full %.0f wait_h_full %.0f wait_dh_empty %.0f wait_dw_empty %.0f total %.0f (max %.0f) tiles %.1f | "
                    "epilogue: wait_dh_full %.0f dh_flush %.0f wait_dw_full %.0f dw_flush %.0f total %.0f\n",
            mode, grid, p.flat, p.gb_per_item, a[0] / 1e3, a[1] / 1e3, a[2] / 1e3, a[4] / 1e3, a[5] / 1e3, a[6] / 1e3, a[7] / 1e3, a[8] / 1e3,
            mx_total / 1e3, a[9], a[10] / 1e3, a[11] / 1e3, a[12] / 1e3, a[13] / 1e3, a[14] / 1e3);
  }
  return DCA_OK;
}

}  // namespace tc
}  // namespace dca

// ------------------------------------------------------------------------------------ C ABI (tests / profiling)
using namespace dca;
extern "C" int dca_tc_gene_gemm(int32_t mode, const void* Z0, const void* Z1, const void* Z2, int64_t ldz, int32_t batch,
                                int32_t genes, int32_t n_heads, const void* H, const void* W, float* out_b, float* dW0,
                                float* dW1, float* dW2, int64_t dW_ld, int32_t dW_transposed, float* db0, float* db1,
                                float* db2, void* stream) {
  if (!Z0 || batch <= 0 || genes <= 0 || n_heads < 1 || n_heads > 3) { set_error("dca_tc_gene_gemm: bad argument"); return DCA_ERR_BAD_ARG; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const __nv_bfloat16* Z[3] = {(const __nv_bfloat16*)Z0, (const __nv_bfloat16*)Z1, (const __nv_bfloat16*)Z2};
  float* dW[3] = {dW0, dW1, dW2};
  float* db[3] = {db0, db1, db2};
  const __nv_bfloat16* Wp[3];
  for (int i = 0; i < 3; ++i) Wp[i] = (const __nv_bfloat16*)W + (mode == 3 ? (size_t)(i < n_heads ? i : 0) * 64 * genes : 0);
  return tc::gene_gemm_tc(mode, Z, ldz, batch, genes, n_heads, (const __nv_bfloat16*)H, Wp, out_b, dW,
                          dW_ld, dW_transposed, db, sms, (cudaStream_t)stream);
}
