// swapnet_b200 — tcgen05 implicit-GEMM kernels (sm_100a).
//
// Every dense contraction of the SwapNet hot path (reference call sites:
// modules/layers.py:15,31,131-138  Conv2d 4x4 s2 / ConvTranspose2d 4x4 s2 /
// reflect-pad Conv2d 3x3; modules/swapnet_modules.py:85-90 upsample+pad+conv
// head; modules/discriminators.py:111-131 PatchGAN convs; and their autograd
// dgrad / wgrad) is lowered by the host onto ONE generic contraction:
//
//   "tap GEMM" (conv mode, K-major operands)
//     D[(n,h,w), j] = sum_{t < ntaps} sum_{c < k_per_tap}
//                       A[n, h + dh_t, w + dw_t, (hp_t), c_off_t + c] * Wp[j, kb_off_t + c]
//   A is an NHWC activation tensor carried as split-bf16 planes (hi, lo); the
//   128-row M tile is a th x tw x nb patch of GEMM rows, so the A tile of one
//   (tap, 64-channel chunk) is a single 5-D TMA box whose out-of-bounds part is
//   zero-filled by the TMA unit (= the conv zero padding).  Stride-2 gathers go
//   through a 2x2 "parity view" of the same memory (dims c', w/2, h%2, h/2, n).
//
//   "wgrad GEMM" (MN-major operands)
//     G[i, j] (+)= sum_{pixels (n,h,w)} X[n, h + dh, w + dw, i] * Y[n, h + dh', w + dw', j]
//   both operands are activation patches; the reduction runs over pixels.
//
// Arithmetic: bf16 tensor-core MMAs (tcgen05.mma kind::f16) with fp32
// accumulation in TMEM.  NSPLIT = 3 evaluates hi*hi + lo*hi + hi*lo, i.e. an
// fp32-faithful product (~2^-16 relative) as the 1e-3 fp32 parity bar requires;
// NSPLIT = 1 is the single-pass bf16 fast mode.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA
// issuer, warps 2..5 = epilogue (TMEM -> registers -> global).
#include <cstdlib>

#include "common.cuh"
#include "plan.h"

namespace {

constexpr int kBlockM = 128;
constexpr int kTileRing = 4;
constexpr int kTileBytes = 16384;  // 128 rows x 128 B (one operand plane of one stage)
constexpr int kThreads = 192;

template <int NSPLIT>
struct Cfg {
  static constexpr int kPlanes = NSPLIT == 3 ? 2 : 1;
  static constexpr int kStages = NSPLIT == 3 ? 3 : 6;
  static constexpr int kStageBytes = kPlanes * 2 * kTileBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;  // + alignment slack
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == SN_ACT_TANH) return tanhf(v);
  return v;
}

// Column sums over a warp's 32 rows of a 32 x 16 register tile (x[j] = this lane's value of column j): a butterfly that
// halves the number of live values at every exchange (8 + 4 + 2 + 1 + 1 = 16 shuffles instead of 16 x 5).  On return
// x[0] holds the total of column (lane >> 1) — both lanes of a pair hold the same total.
__device__ __forceinline__ void col_sums16(float (&x)[16], int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool up = lane & 16;
    const float send = up ? x[j] : x[j + 8];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
    x[j] = (up ? x[j + 8] : x[j]) + recv;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool up = lane & 8;
    const float send = up ? x[j] : x[j + 4];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
    x[j] = (up ? x[j + 4] : x[j]) + recv;
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const bool up = lane & 4;
    const float send = up ? x[j] : x[j + 2];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
    x[j] = (up ? x[j + 2] : x[j]) + recv;
  }
  {
    const bool up = lane & 2;
    const float send = up ? x[0] : x[1];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
    x[0] = (up ? x[1] : x[0]) + recv;
  }
  x[0] += __shfl_xor_sync(0xffffffffu, x[0], 1);
}

// ============================================================================
// conv mode
// ============================================================================
// CW = channels per A row (64 / 32 / 16): a template parameter so that the single MMA-issuing thread
// carries no runtime index arithmetic (measured: runtime div/mod there costs 30 % of the kernel).
template <int NSPLIT, int CW>
__global__ void __launch_bounds__(kThreads, 1)
tap_gemm_kernel(const __grid_constant__ TapGemmParams p, const int m_tiles, const int n_tiles, const int total_tiles) {
  // Persistent: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (M tile fastest, then N tile,
  // then output-parity phase).  The smem ring keeps running across tiles and the accumulator is double
  // buffered in TMEM (2 x 128 columns), so the epilogue of tile i overlaps the main loop of tile i + 1 and
  // the pipeline prologue is paid once per CTA instead of once per tile.
  using C = Cfg<NSPLIT>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C::kStages];
  __shared__ __align__(8) uint64_t empty_bar[C::kStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];    // MMA -> epilogue: accumulator b complete
  __shared__ __align__(8) uint64_t tempty_bar[2];   // epilogue -> MMA: accumulator b drained (4 warps)
  // dynamic tile schedule (p.tile_counter != null): the producer draws tile numbers from a device counter and hands them
  // to the MMA and epilogue warps through a 4-deep ring, so a CTA that becomes resident late (SMs held by NCCL or by the
  // weight-gradient stream's CTAs) finds only the tiles nobody has taken yet instead of a fixed 1/gridDim share
  __shared__ __align__(8) uint64_t tr_full[kTileRing];
  __shared__ __align__(8) uint64_t tr_empty[kTileRing];
  __shared__ int tile_ring[kTileRing];
  __shared__ uint32_t tmem_base_smem;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool dyn = p.tile_counter != nullptr;

  constexpr int cw = CW;                    // channels per A row: 64, 32 or 16
  constexpr int tps = 64 / cw;              // taps sharing one 64-deep stage (1, 2 or 4)
  // output parity phase z (merged 4-phase launches) owns taps [z*tpp, (z+1)*tpp)
  const int tpp = p.ntaps / p.nphase;
  const int k_iters = cw == 64 ? tpp * p.chunks : tpp / tps;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 4);
    }
    for (int r = 0; r < kTileRing; ++r) {
      mbar_init(&tr_full[r], 1);
      mbar_init(&tr_empty[r], 5);     // the MMA thread + the four epilogue warps
    }
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    for (int pl = 0; pl < C::kPlanes; ++pl) {
      tma_prefetch_desc(&p.tmA[pl]);
      tma_prefetch_desc(&p.tmB[pl]);
    }
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t stage_tx = C::kPlanes * (p.a_rows * 128 + p.block_n * 128);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x;; ++tcount) {
        int next = tile + (int)gridDim.x;
        if (dyn) {
          // the first tile is static; every further one a ticket.  The ticket for the NEXT tile is drawn now (its latency
          // hides behind this tile's loads), the tile number goes to the other warps through the ring
          if (tile < total_tiles) next = (int)gridDim.x + atomicAdd(p.tile_counter, 1);
          const uint32_t slot = tcount % kTileRing, rph = (tcount / kTileRing) & 1;
          mbar_wait(&tr_empty[slot], rph ^ 1);
          tile_ring[slot] = tile < total_tiles ? tile : -1;
          mbar_arrive(&tr_full[slot]);
        }
        if (tile >= total_tiles) break;
        const int mt = tile % m_tiles;
        const int rest = tile / m_tiles;
        const int ncol0 = (rest % n_tiles) * p.block_n;
        const int tap0 = (rest / n_tiles) * tpp;
        const int w0 = (mt % p.tiles_w) * p.tw;
        const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.th;
        const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.nb;
        for (int k = 0; k < k_iters; ++k, ++it) {
          const uint32_t s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], stage_tx);
          uint8_t* st = smem + s * C::kStageBytes;
          if (cw == 64) {
            const int t = k / p.chunks;
            const int ch = k - t * p.chunks;
            const TapDesc tap = p.taps[tap0 + t];
            if (p.a_merged) {   // dims (c, w, h, n, plane): hi tile, then lo tile
              tma_load_5d(&p.tmA[0], &full_bar[s], st, tap.c_off + ch * 64, w0 + tap.dw, h0 + tap.dh, n0, 0);
            } else {
#pragma unroll
              for (int pl = 0; pl < C::kPlanes; ++pl)
                tma_load_5d(&p.tmA[pl], &full_bar[s], st + pl * kTileBytes, tap.c_off + ch * 64,
                            w0 + tap.dw, tap.hp, h0 + tap.dh, n0);
            }
            if (p.b_merged) {   // dims (k, row, plane)
              tma_load_3d(&p.tmB[0], &full_bar[s], st + C::kPlanes * kTileBytes, tap.kb_off + ch * 64, ncol0, 0);
            } else {
#pragma unroll
              for (int pl = 0; pl < C::kPlanes; ++pl)
                tma_load_2d(&p.tmB[pl], &full_bar[s], st + (C::kPlanes + pl) * kTileBytes,
                            tap.kb_off + ch * 64, ncol0);
            }
          } else {
            // narrow operand: 64/cw taps, each a [128 rows][cw channels] sub-tile, fill one stage; the
            // weights of consecutive taps are contiguous in K, so B is still one 64-deep box
            constexpr int sub = 128 * cw * 2;
#pragma unroll
            for (int j = 0; j < tps; ++j) {
              const TapDesc tap = p.taps[tap0 + k * tps + j];
#pragma unroll
              for (int pl = 0; pl < C::kPlanes; ++pl)
                tma_load_5d(&p.tmA[pl], &full_bar[s], st + pl * kTileBytes + j * sub, tap.c_off,
                            w0 + tap.dw, tap.hp, h0 + tap.dh, n0);
            }
            const int kb = p.taps[tap0 + k * tps].kb_off;
            if (p.b_merged) {
              tma_load_3d(&p.tmB[0], &full_bar[s], st + C::kPlanes * kTileBytes, kb, ncol0, 0);
            } else {
#pragma unroll
              for (int pl = 0; pl < C::kPlanes; ++pl)
                tma_load_2d(&p.tmB[pl], &full_bar[s], st + (C::kPlanes + pl) * kTileBytes, kb, ncol0);
            }
          }
        }
        tile = next;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_16(kBlockM, p.block_n, p.a_fmt, p.b_fmt, 0, 0);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x;; tile += gridDim.x, ++tcount) {
        if (dyn) {
          const uint32_t slot = tcount % kTileRing, rph = (tcount / kTileRing) & 1;
          mbar_wait(&tr_full[slot], rph);
          tile = tile_ring[slot];
          mbar_arrive(&tr_empty[slot]);
          if (tile < 0) break;
        } else if (tile >= total_tiles) {
          break;
        }
        const uint32_t b = tcount & 1;
        mbar_wait(&tempty_bar[b], ((tcount >> 1) & 1) ^ 1);   // the epilogue has drained accumulator b
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + b * 128;
        uint32_t acc = 0;
        for (int k = 0; k < k_iters; ++k, ++it) {
          const uint32_t s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t st = smem_base + s * C::kStageBytes;
          // K-major, rows of cw channels: SWIZZLE_128B/64B/32B, SBO = 8 rows, LBO unused (1)
          constexpr uint32_t a_layout = cw >= 64 ? 2u : (cw == 32 ? 4u : 6u);
          constexpr uint32_t a_sbo = 8u * cw * 2u;
          const uint64_t a_hi = umma_smem_desc(st, 16, a_sbo, a_layout);
          const uint64_t b_hi = umma_smem_desc(st + C::kPlanes * kTileBytes, 16, 1024);
          const uint64_t a_lo = umma_smem_desc(st + p.a_lo_off, 16, a_sbo, a_layout);
          const uint64_t b_lo = umma_smem_desc(st + C::kPlanes * kTileBytes + p.b_lo_off, 16, 1024);
          constexpr uint32_t sub16 = (uint32_t)(128 * cw * 2) >> 4;  // sub-tile stride in 16-B units
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {  // 4 x (UMMA_K = 16 elements = 32 B) per 64-deep stage
            const uint64_t badv = (uint64_t)(kk * 2);
            // A: k-step kk lives in sub-tile (16kk / cw), at byte offset ((16kk) % cw) * 2 of its rows
            const uint64_t aadv = (uint64_t)(((kk * 16) / cw) * sub16 + (((kk * 16) % cw) >> 3));
            umma_bf16(tmem_d, a_hi + aadv, b_hi + badv, idesc, acc);
            acc = 1;
            if (NSPLIT == 3) {
              umma_bf16(tmem_d, a_lo + aadv, b_hi + badv, idesc, 1);
              umma_bf16(tmem_d, a_hi + aadv, b_lo + badv, idesc, 1);
            }
          }
          umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
        }
        umma_commit(&tfull_bar[b]);
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    const int w_i = row % p.tw;
    const int h_i = (row / p.tw) % p.th;
    const int n_i = row / (p.tw * p.th);
    const float oscale = p.b_scale ? p.b_scale[1] : 1.f;  // undo the power-of-two weight scale (exact)
    uint32_t tcount = 0;
    for (int tile = blockIdx.x;; tile += gridDim.x, ++tcount) {
      if (dyn) {
        const uint32_t slot = tcount % kTileRing, rph = (tcount / kTileRing) & 1;
        mbar_wait(&tr_full[slot], rph);
        tile = tile_ring[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&tr_empty[slot]);
        if (tile < 0) break;
      } else if (tile >= total_tiles) {
        break;
      }
      const int mt = tile % m_tiles;
      const int rest = tile / m_tiles;
      const int ncol0 = (rest % n_tiles) * p.block_n;
      const int z = rest / n_tiles;
      const int gw = (mt % p.tiles_w) * p.tw + w_i;
      const int gh = ((mt / p.tiles_w) % p.tiles_h) * p.th + h_i;
      const int gn = (mt / (p.tiles_w * p.tiles_h)) * p.nb + n_i;
      const bool valid = (row < p.a_rows) && (gw < p.m_w) && (gh < p.m_h) && (gn < p.m_n);
      const int ph_h = p.nphase == 4 ? (z >> 1) : 0, ph_w = p.nphase == 4 ? (z & 1) : 0;
      float* optr = p.out + (long long)gn * p.out_sn + (long long)(gh * p.omh + p.ooh + ph_h) * p.out_sh +
                    (long long)(gw * p.omw + p.oow + ph_w) * p.out_sw + ncol0;
      const uint32_t b = tcount & 1;
      mbar_wait(&tfull_bar[b], (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + b * 128 + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_d + (uint32_t)c0, r);
        tmem_ld_wait();
        if (c0 + 16 >= p.block_n) {   // last read of this accumulator: hand it back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[b]);
        }
        if (valid && p.stack_slot) {
          // phase-stacked head: column -> (phase, channel); every phase lands on its own output pixel
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = ncol0 + c0 + j;
            const int ph = col / p.stack_slot, c = col - ph * p.stack_slot;
            if (ph < 4 && c < p.stack_c) {
              float v = __uint_as_float(r[j]) * oscale;
              if (p.bias) v += p.bias[c];
              float* o = p.out + (long long)gn * p.out_sn + (long long)(gh * p.omh + (ph >> 1)) * p.out_sh +
                         (long long)(gw * p.omw + (ph & 1)) * p.out_sw + c;
              *o = apply_act(v, p.act);
            }
          }
        } else if (p.stats && ncol0 + c0 >= p.n_valid) {
          // a 16-column chunk beyond the last output channel (n_valid % 16 == 0 in this mode): nothing to store or sum
        } else if (p.stats) {
          // InstanceNorm statistics fused into the producer (layers.py:17,33,134): every tile row belongs to image gn
          // (nb == 1, checked by the host), so the warp's 32 rows reduce to per-column sums of y and y^2 — one fp64
          // atomic pair per column per warp into stats[n][c] = (sum, sum of squares), finalised by stats_finalize.
          float v[16], sq[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float t = __uint_as_float(r[j]) * oscale;
            if (p.bias) t += p.bias[ncol0 + c0 + j];
            v[j] = valid ? t : 0.f;
          }
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(optr + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) sq[j] = v[j] * v[j];
          col_sums16(v, lane);
          col_sums16(sq, lane);
          if ((lane & 1) == 0) {
            const int tn = (mt / (p.tiles_w * p.tiles_h)) * p.nb;      // the tile's image (uniform over the CTA)
            double* st = p.stats + ((long long)tn * p.n_valid + ncol0 + c0 + (lane >> 1)) * 2;
            atomicAdd(st, (double)v[0]);
            atomicAdd(st + 1, (double)sq[0]);
          }
        } else if (valid) {
          if (p.vec4 && ncol0 + c0 + 16 <= p.n_valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 v;
              v.x = __uint_as_float(r[j + 0]) * oscale;
              v.y = __uint_as_float(r[j + 1]) * oscale;
              v.z = __uint_as_float(r[j + 2]) * oscale;
              v.w = __uint_as_float(r[j + 3]) * oscale;
              if (p.bias) {
                const float4 bb = *reinterpret_cast<const float4*>(p.bias + ncol0 + c0 + j);
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
              }
              if (p.act) {
                v.x = apply_act(v.x, p.act); v.y = apply_act(v.y, p.act);
                v.z = apply_act(v.z, p.act); v.w = apply_act(v.w, p.act);
              }
              *reinterpret_cast<float4*>(optr + c0 + j) = v;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = ncol0 + c0 + j;
              if (col < p.n_valid) {
                float v = __uint_as_float(r[j]) * oscale;
                if (p.bias) v += p.bias[col];
                optr[c0 + j] = apply_act(v, p.act);
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
  if (dyn && threadIdx.x == 0) {
    // the last CTA to finish re-arms the counters for the next launch of this plan (every CTA's tickets are drawn
    // before it gets here; launches of one plan never overlap)
    __threadfence();
    if (atomicAdd(p.tile_counter + 1, 1) == (int)gridDim.x - 1) {
      p.tile_counter[0] = 0;
      p.tile_counter[1] = 0;
      __threadfence();
    }
  }
}

// ============================================================================
// wgrad mode
// ============================================================================
template <int NSPLIT>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_gemm_kernel(const __grid_constant__ WgradParams p) {
  using C = Cfg<NSPLIT>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C::kStages];
  __shared__ __align__(8) uint64_t empty_bar[C::kStages];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tile = blockIdx.x / p.n_tiles;
  const int n_tile = blockIdx.x - m_tile * p.n_tiles;
  // grid.y = tap (wide Y) or tap group (narrow Y: the group's taps are column blocks of one accumulator
  // and share the X tile, which is then read once per pixel tile instead of once per tap)
  const bool grouped = p.ngroups > 0;
  const int tap_i = grouped ? p.gstart[blockIdx.y] : blockIdx.y;
  const int gsize = grouped ? p.gsize[blockIdx.y] : 1;
  const int m0 = m_tile * kBlockM;
  const int ncol0 = n_tile * p.block_n;
  const int total = p.tiles_w * p.tiles_h * p.tiles_n;
  const int kt0 = (int)(((long long)total * blockIdx.z) / gridDim.z);
  const int kt1 = (int)(((long long)total * (blockIdx.z + 1)) / gridDim.z);
  const int k_iters = kt1 - kt0;
  const int ycw = p.y_chunk;                               // channels per Y row: 64, 32, 16
  const int y_blocks = ycw == 64 ? p.block_n / 64 : gsize; // narrow Y: one ycw-wide atom per grouped tap
  const int y_block_bytes = 64 * ycw * 2;                  // 64 pixels x ycw channels

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0 && lane == 0) {
    for (int pl = 0; pl < C::kPlanes; ++pl) {
      tma_prefetch_desc(&p.tmX[pl]);
      tma_prefetch_desc(&p.tmY[pl]);
    }
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (k_iters > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const TapDesc xt = p.xtaps[tap_i];
        const uint32_t stage_tx = C::kPlanes * (2 * 8192 + y_blocks * y_block_bytes);
        // rot_mode 0: every CTA walks the pixel tiles in the same order (a narrow wavefront shared through L2);
        // 1: staggered per tap (blockIdx.y); 2: staggered per CTA
        int rot = 0;
        if (p.rot_mode == 1) rot = (int)(((long long)k_iters * blockIdx.y) / gridDim.y);
        else if (p.rot_mode == 2)
          rot = (int)(((long long)k_iters * ((blockIdx.x + blockIdx.y * gridDim.x) % 148)) / 148);
        for (int it = 0; it < k_iters; ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], stage_tx);
          int kr = it + rot;                       // per-CTA rotation of the pixel-tile order (see p.rot_mode)
          if (kr >= k_iters) kr -= k_iters;
          const int kt = kt0 + kr;
          const int tw_i = kt % p.tiles_w;
          const int th_i = (kt / p.tiles_w) % p.tiles_h;
          const int tn_i = kt / (p.tiles_w * p.tiles_h);
          const int w0 = tw_i * p.tw, h0 = th_i * p.th, n0 = tn_i * p.nb;
          uint8_t* st = smem + s * C::kStageBytes;
          if (p.x_merged) {   // one box per 64-channel block: [hi 8 KB][lo 8 KB]
            for (int b = 0; b < 2; ++b)
              tma_load_5d(&p.tmX[0], &full_bar[s], st + b * 16384, xt.c_off + m0 + b * 64, w0 + xt.dw, h0 + xt.dh,
                          n0, 0);
          } else {
#pragma unroll
            for (int pl = 0; pl < C::kPlanes; ++pl)
              for (int b = 0; b < 2; ++b)
                tma_load_5d(&p.tmX[pl], &full_bar[s], st + pl * kTileBytes + b * 8192,
                            xt.c_off + m0 + b * 64, w0 + xt.dw, xt.hp, h0 + xt.dh, n0);
          }
          if (p.y_merged) {
            const TapDesc yt = p.ytaps[tap_i];
            for (int b = 0; b < y_blocks; ++b)
              tma_load_5d(&p.tmY[0], &full_bar[s], st + C::kPlanes * kTileBytes + b * 16384,
                          yt.c_off + ncol0 + b * 64, w0 + yt.dw, h0 + yt.dh, n0, 0);
          } else {
#pragma unroll
            for (int pl = 0; pl < C::kPlanes; ++pl)
              for (int b = 0; b < y_blocks; ++b) {
                const TapDesc yt = p.ytaps[grouped ? tap_i + b : tap_i];
                tma_load_5d(&p.tmY[pl], &full_bar[s],
                            st + (C::kPlanes + pl) * kTileBytes + b * y_block_bytes,
                            yt.c_off + (grouped ? 0 : ncol0 + b * 64), w0 + yt.dw, yt.hp, h0 + yt.dh, n0);
              }
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_16(kBlockM, p.block_n, p.x_fmt, p.y_fmt, 1, 1);
        uint32_t acc = 0;
        for (int it = 0; it < k_iters; ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t st = smem_base + s * C::kStageBytes;
          // MN-major SWIZZLE_128B: LBO = stride between 64-channel blocks (8192 B),
          // SBO = stride between 8-pixel groups (1024 B)
          const uint32_t y_layout = umma_layout_of_chunk(ycw);
          const uint32_t y_sbo = 8u * ycw * 2u;              // 8 pixel rows of the narrow / full atom
          // merged planes: the 64-channel blocks are [hi 8 KB][lo 8 KB] pairs, 16 KB apart
          const uint32_t x_lbo = p.x_merged ? 16384u : 8192u;
          const uint32_t x_lo_off = p.x_merged ? 8192u : (uint32_t)kTileBytes;
          const uint32_t y_lbo = p.y_merged ? 16384u : (uint32_t)y_block_bytes;
          const uint32_t y_lo_off = p.y_merged ? 8192u : (uint32_t)kTileBytes;
          const uint64_t x_hi = umma_smem_desc(st, x_lbo, 1024);
          // LBO = stride between the N atoms (64-channel blocks, or the grouped taps' narrow blocks)
          const uint64_t y_hi = umma_smem_desc(st + C::kPlanes * kTileBytes, y_lbo, y_sbo, y_layout);
          const uint64_t x_lo = umma_smem_desc(st + x_lo_off, x_lbo, 1024);
          const uint64_t y_lo = umma_smem_desc(st + C::kPlanes * kTileBytes + y_lo_off, y_lbo, y_sbo, y_layout);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x 16 pixels; 16 pixel rows = 2048 B (X), 16 * ycw * 2 B (Y)
            const uint64_t xadv = (uint64_t)(k * 128);
            const uint64_t yadv = (uint64_t)((k * 16 * ycw * 2) >> 4);
            umma_bf16(tmem_base, x_hi + xadv, y_hi + yadv, idesc, acc);
            acc = 1;
            if (NSPLIT == 3) {
              umma_bf16(tmem_base, x_lo + xadv, y_hi + yadv, idesc, 1);
              umma_bf16(tmem_base, x_hi + xadv, y_lo + yadv, idesc, 1);
            }
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&accum_bar);
      }
    } else {
      const int q = warp & 3;
      const int row = m0 + q * 32 + lane;
      const bool valid = row < p.rows_valid;
      float* orow = p.out + (long long)row * p.s_row;
      mbar_wait(&accum_bar, 0);
      tc_fence_after();
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        if (valid) {
          if (!grouped) {
            float* optr = orow + p.tap_off[tap_i];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = ncol0 + c0 + j;
              if (col < p.cols_valid)
                atomicAdd(optr + (long long)col * p.s_col, __uint_as_float(r[j]));
            }
          } else {
            const int b = c0 / ycw;               // which grouped tap this 16-column chunk belongs to
            if (b < gsize) {
              float* optr = orow + p.tap_off[tap_i + b];
              const int cbase = c0 - b * ycw;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (cbase + j < p.cols_valid)
                  atomicAdd(optr + (long long)(cbase + j) * p.s_col, __uint_as_float(r[j]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

}  // namespace

// ============================================================================
// host side
// ============================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// 5-D map over a split-bf16 NHWC plane.  dims (c', w, hp, h, n); see file header.
// plane_stride > 0 (non-parity only): one extra outermost dimension of 2 planes (hi, lo) `plane_stride`
// bytes apart -> dims (c, w, h, n, plane); otherwise dims (c', w, h parity, h, n).
int sn_make_act_map(CUtensorMap* tm, const void* base, int N, int H, int W, int C, int pitch,
                    int parity, int box_w, int box_h, int box_n, int chunk = 64, long long plane_stride = 0) {
  PFN_encodeTiled enc = get_encode_fn();
  SN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
  SN_REQUIRE(pitch % 8 == 0 && ((uintptr_t)base % 16) == 0,
             "activation plane needs 16-B aligned base and pitch %% 8 == 0 (pitch=%d)", pitch);
  SN_REQUIRE(box_w >= 1 && box_h >= 1 && box_n >= 1 && box_w <= 256 && box_h <= 256 && box_n <= 256,
             "bad TMA box %d x %d x %d", box_w, box_h, box_n);
  cuuint64_t dims[5];
  cuuint64_t strides[4];
  const cuuint64_t e = 2;  // bytes per bf16
  cuuint32_t box[5] = {(cuuint32_t)chunk, (cuuint32_t)box_w, 1, (cuuint32_t)box_h, (cuuint32_t)box_n};
  if (plane_stride > 0) {
    SN_REQUIRE(plane_stride % 16 == 0, "merged planes: 16-B aligned plane stride");
    if (!parity) {
      dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = N; dims[4] = 2;
      strides[0] = (cuuint64_t)pitch * e;
      strides[1] = (cuuint64_t)W * pitch * e;
    } else {
      // parity view with BOTH parities folded into the channel coordinate: c' = hp*W*pitch + pw*pitch + c
      // (overlapping dimensions are fine for loads), dims (c', w/2, h/2, n, plane)
      SN_REQUIRE(H % 2 == 0 && W % 2 == 0, "parity view needs even H, W (got %d x %d)", H, W);
      dims[0] = (cuuint64_t)(W + 1) * pitch + C; dims[1] = W / 2; dims[2] = H / 2; dims[3] = N; dims[4] = 2;
      strides[0] = (cuuint64_t)2 * pitch * e;
      strides[1] = (cuuint64_t)2 * W * pitch * e;
    }
    strides[2] = (cuuint64_t)H * W * pitch * e;
    strides[3] = (cuuint64_t)plane_stride;
    box[2] = (cuuint32_t)box_h; box[3] = (cuuint32_t)box_n; box[4] = 2;
  } else if (!parity) {
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = N;
    strides[0] = (cuuint64_t)pitch * e;
    strides[1] = (cuuint64_t)W * pitch * e;
    strides[2] = (cuuint64_t)W * pitch * e;
    strides[3] = (cuuint64_t)H * W * pitch * e;
  } else {
    SN_REQUIRE(H % 2 == 0 && W % 2 == 0, "parity view needs even H, W (got %d x %d)", H, W);
    dims[0] = (cuuint64_t)pitch + C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
    strides[0] = (cuuint64_t)2 * pitch * e;
    strides[1] = (cuuint64_t)W * pitch * e;
    strides[2] = (cuuint64_t)2 * W * pitch * e;
    strides[3] = (cuuint64_t)H * W * pitch * e;
  }
  SN_REQUIRE(chunk == 64 || chunk == 32 || chunk == 16, "row chunk must be 64, 32 or 16 channels (got %d)", chunk);
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle swz = chunk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : chunk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SN_REQUIRE(r == CUDA_SUCCESS,
             "cuTensorMapEncodeTiled(act) failed: %d (N=%d H=%d W=%d C=%d pitch=%d parity=%d box=%dx%dx%d)",
             (int)r, N, H, W, C, pitch, parity, box_w, box_h, box_n);
  return SN_OK;
}

// 2-D map over packed weights [rows][k_total] bf16, K contiguous.
int sn_make_weight_map(CUtensorMap* tm, const void* base, int rows, long long k_total, int box_rows,
                       long long plane_stride = 0) {
  PFN_encodeTiled enc = get_encode_fn();
  SN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
  SN_REQUIRE(k_total % 64 == 0 && ((uintptr_t)base % 16) == 0, "packed weights need K %% 64 == 0");
  cuuint64_t dims[3] = {(cuuint64_t)k_total, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)k_total * 2, (cuuint64_t)plane_stride};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, plane_stride > 0 ? 3 : 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed: %d (rows=%d K=%lld box_rows=%d)",
             (int)r, rows, k_total, box_rows);
  return SN_OK;
}

static void pick_patch(int m_h, int m_w, int rows, int* th, int* tw, int* nb) {
  // rows = 128 (conv mode) or 64 (wgrad mode); dims are powers of two except the
  // PatchGAN 63/62 planes, which take the largest patch and rely on masking.
  int w = 16;
  while (w > 1 && w / 2 >= m_w) w /= 2;
  if (w > rows) w = rows;
  int h = rows / w;
  while (h > 1 && h / 2 >= m_h) h /= 2;
  *tw = w;
  *th = h;
  *nb = rows / (w * h);
}

// conv mode: any th x tw x nb patch with <= 128 GEMM rows works (rows the TMA box does not write
// are never stored), so pick the patch that wastes the fewest rows on this plane — e.g. 7 x 17 on the
// 34 x 34 padded resblock-gradient grid (90 % instead of 60 % with 8 x 16).  Ties: squarer patch
// (smaller halo re-read across taps).
static void pick_patch_conv(int m_n, int m_h, int m_w, int* th, int* tw, int* nb) {
  double best_eff = -1.0;
  int best_perim = 1 << 30, best_rows = 0, best_area = 0;
  const long long useful = (long long)m_n * m_h * m_w;
  for (int w = 1; w <= m_w && w <= 128; ++w) {
    for (int h = 1; h <= m_h && h * w <= 128; ++h) {
      // planes of >= 128 pixels: purely spatial patches (halo locality); smaller planes: whole
      // planes of several images per tile
      int n = 1;
      if (m_h * m_w < 128) {
        if (w != m_w || h != m_h) continue;
        n = 128 / (w * h);
        if (n > m_n) n = m_n;
        if (n > 256) n = 256;
      }
      const long long tiles = (long long)((m_w + w - 1) / w) * ((m_h + h - 1) / h) * ((m_n + n - 1) / n);
      const double eff = (double)useful / (double)(tiles * 128);
      const int rows = w * h * n, perim = w + h, area = w * h;
      // ties: more rows per tile, then the larger spatial patch (fewer images per tile), then the squarer one
      const bool better =
          eff > best_eff + 1e-9 ||
          (eff > best_eff - 1e-9 &&
           (rows > best_rows ||
            (rows == best_rows && (area > best_area || (area == best_area && (perim < best_perim || (perim == best_perim && w > *tw)))))));
      if (better) {
        best_eff = eff; best_rows = rows; best_perim = perim; best_area = area;
        *tw = w; *th = h; *nb = n;
      }
    }
  }
}

static int g_smem_attr_done[2][2] = {{0, 0}, {0, 0}};

int sn_tap_gemm_plan_init(TapGemmPlan* plan, const sn_tap_gemm_desc* d) {
  memset(plan, 0, sizeof(*plan));
  TapGemmParams& p = plan->p;
  SN_REQUIRE(d->nsplit == 1 || d->nsplit == 3, "nsplit must be 1 or 3");
  SN_REQUIRE(d->ntaps >= 1 && d->ntaps <= SN_MAX_TAPS, "ntaps out of range: %d", d->ntaps);
  const int a_chunk = d->a_chunk ? d->a_chunk : 64;
  SN_REQUIRE(a_chunk == 64 || a_chunk == 32 || a_chunk == 16, "a_chunk must be 64, 32 or 16");
  if (a_chunk == 64) {
    SN_REQUIRE(d->k_per_tap > 0 && d->k_per_tap % 64 == 0, "k_per_tap must be a multiple of 64");
  } else {
    SN_REQUIRE(d->k_per_tap == a_chunk && d->ntaps % (64 / a_chunk) == 0,
               "narrow operand: k_per_tap must equal a_chunk (%d) and ntaps (%d) be a multiple of %d", a_chunk,
               d->ntaps, 64 / a_chunk);
  }
  p.a_chunk = a_chunk;
  SN_REQUIRE(d->block_n >= 16 && d->block_n <= 128 && d->block_n % 16 == 0,
             "block_n must be a multiple of 16 in [16,128]");
  SN_REQUIRE(d->a_hi && d->b_hi && d->out, "null operand");
  SN_REQUIRE(d->nsplit == 1 || (d->a_lo && d->b_lo), "nsplit=3 needs lo planes");
  SN_REQUIRE(d->a_fmt == d->b_fmt, "A and B of one tcgen05.mma must share a 16-bit format (a=%d b=%d)",
             d->a_fmt, d->b_fmt);
  int th, tw, nb;
  pick_patch_conv(d->m_n, d->m_h, d->m_w, &th, &tw, &nb);
  p.a_rows = th * tw * nb;
  p.tw = tw; p.th = th; p.nb = nb;
  p.tiles_w = (d->m_w + tw - 1) / tw;
  p.tiles_h = (d->m_h + th - 1) / th;
  p.tiles_n = (d->m_n + nb - 1) / nb;
  p.m_w = d->m_w; p.m_h = d->m_h; p.m_n = d->m_n;
  p.ntaps = d->ntaps;
  p.chunks = a_chunk == 64 ? d->k_per_tap / 64 : 1;
  for (int t = 0; t < d->ntaps; ++t) {
    p.taps[t].c_off = d->taps[t].c_off;
    p.taps[t].kb_off = d->taps[t].kb_off;
    p.taps[t].dw = (short)d->taps[t].dw;
    p.taps[t].dh = (short)d->taps[t].dh;
    p.taps[t].hp = (short)d->taps[t].hp;
  }
  p.block_n = d->block_n;
  p.n_valid = d->n_valid;
  p.out = d->out;
  p.out_sn = d->out_sn; p.out_sh = d->out_sh; p.out_sw = d->out_sw;
  p.omh = d->out_mul_h; p.ooh = d->out_off_h; p.omw = d->out_mul_w; p.oow = d->out_off_w;
  p.nphase = d->nphase == 4 ? 4 : 1;
  SN_REQUIRE(d->nphase == 0 || d->nphase == 1 || d->nphase == 4, "nphase must be 1 or 4");
  if (p.nphase == 4) {
    SN_REQUIRE(d->ntaps % 4 == 0, "4-phase launch: ntaps must split into 4 equal groups");
    const int tpp = d->ntaps / 4;
    SN_REQUIRE(a_chunk == 64 || tpp % (64 / a_chunk) == 0, "4-phase launch: taps per phase must fill whole stages");
  }
  p.bias = d->bias;
  p.stack_slot = d->stack_slot; p.stack_c = d->stack_c;
  p.stats = nullptr;
  if (d->stack_slot > 0)
    SN_REQUIRE(d->nphase <= 1 && d->n_valid == 4 * d->stack_slot && d->stack_c >= 1 && d->stack_c <= d->stack_slot &&
                   d->out_mul_h == 2 && d->out_mul_w == 2,
               "phase-stacked output: n_valid = 4*stack_slot, nphase 1, out_mul 2 (slot=%d c=%d n_valid=%d)",
               d->stack_slot, d->stack_c, d->n_valid);
  p.b_scale = d->b_scale;
  p.a_fmt = d->a_fmt;
  p.b_fmt = d->b_fmt;
  p.act = d->act;
  p.vec4 = ((uintptr_t)d->out % 16 == 0) && (d->out_sn % 4 == 0) && (d->out_sh % 4 == 0) &&
           (d->out_sw % 4 == 0) && (!d->bias || (uintptr_t)d->bias % 16 == 0);
  // fused InstanceNorm statistics: whole tile inside one image, 16-column chunks fully valid, plain vectorised stores
  if (d->stats && nb == 1 && p.vec4 && d->n_valid % 16 == 0 && d->act == 0 && d->stack_slot == 0 && d->m_n * 1 >= 1)
    p.stats = d->stats;
  plan->stats_bytes = p.stats ? sizeof(double) * 2 * (size_t)d->m_n * d->n_valid : 0;
  p.tile_counter = nullptr;
  {
    static int dyn_ok = -1;
    if (dyn_ok < 0) {
      const char* e = getenv("SN_TAP_STATIC_TILES");   // A/B switch: the static tile striding of round 1
      dyn_ok = (e && e[0] == '1') ? 0 : 1;
    }
    if (dyn_ok) {
      SN_CHECK_CUDA(cudaMalloc(&p.tile_counter, 2 * sizeof(int)));
      SN_CHECK_CUDA(cudaMemset(p.tile_counter, 0, 2 * sizeof(int)));
    }
  }
  int rc;
  const void* a_pl[2] = {d->a_hi, d->a_lo};
  const void* b_pl[2] = {d->b_hi, d->b_lo};
  static int merge_ok = -1;
  if (merge_ok < 0) {
    const char* e = getenv("SN_NO_MERGED_PLANES");   // A/B switch
    merge_ok = (e && e[0] == '1') ? 0 : 1;
  }
  const long long a_ps = d->nsplit == 3 ? (const char*)d->a_lo - (const char*)d->a_hi : 0;
  const long long b_ps = d->nsplit == 3 ? (const char*)d->b_lo - (const char*)d->b_hi : 0;
  // the lo tile must start on a swizzle-atom boundary (8 rows x 128 B)
  p.a_merged = merge_ok && a_chunk == 64 && a_ps > 0 && a_ps % 16 == 0 && p.a_rows % 8 == 0;
  if (p.a_merged && d->a_parity)   // h parity moves into the channel coordinate of the merged parity map
    for (int t = 0; t < d->ntaps; ++t) p.taps[t].c_off += p.taps[t].hp * d->a_w * d->a_pitch;
  p.b_merged = merge_ok && b_ps > 0 && b_ps % 16 == 0 && d->block_n % 8 == 0;
  p.a_lo_off = p.a_merged ? p.a_rows * 128 : kTileBytes;
  p.b_lo_off = p.b_merged ? d->block_n * 128 : kTileBytes;
  for (int pl = 0; pl < (d->nsplit == 3 ? 2 : 1); ++pl) {
    if (!(p.a_merged && pl == 1)) {
      rc = sn_make_act_map(&p.tmA[pl], a_pl[pl], d->a_n, d->a_h, d->a_w, d->a_c, d->a_pitch,
                           d->a_parity, tw, th, nb, a_chunk, p.a_merged ? a_ps : 0);
      if (rc) return rc;
    }
    if (!(p.b_merged && pl == 1)) {
      rc = sn_make_weight_map(&p.tmB[pl], b_pl[pl], d->b_rows, d->b_k, d->block_n, p.b_merged ? b_ps : 0);
      if (rc) return rc;
    }
  }
  if (p.a_merged) p.tmA[1] = p.tmA[0];
  if (p.b_merged) p.tmB[1] = p.tmB[0];
  plan->nsplit = d->nsplit;
  plan->grid = dim3(p.tiles_w * p.tiles_h * p.tiles_n, (d->n_valid + d->block_n - 1) / d->block_n, p.nphase);
  return SN_OK;
}

template <int NSPLIT, int CW>
static int launch_tap(const TapGemmPlan* plan, cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    SN_CHECK_CUDA(cudaFuncSetAttribute(tap_gemm_kernel<NSPLIT, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg<NSPLIT>::kSmemBytes));
    attr_done = true;
  }
  // persistent launch: at most one CTA per SM; plan->grid = (M tiles, N tiles, phases)
  static int sms = 0;
  static int one_tile_per_cta = -1;
  if (sms == 0) {
    int dev = 0;
    SN_CHECK_CUDA(cudaGetDevice(&dev));
    SN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const char* e = getenv("SN_TAP_ONE_TILE_PER_CTA");   // A/B switch: the non-persistent schedule
    one_tile_per_cta = (e && e[0] == '1') ? 1 : 0;
  }
  const int m_tiles = (int)plan->grid.x, n_tiles = (int)plan->grid.y;
  const int total = m_tiles * n_tiles * (int)plan->grid.z;
  if (plan->p.stats) SN_CHECK_CUDA(cudaMemsetAsync(plan->p.stats, 0, plan->stats_bytes, stream));
  const int ctas = one_tile_per_cta ? total : (total < sms ? total : sms);
  tap_gemm_kernel<NSPLIT, CW><<<ctas, kThreads, Cfg<NSPLIT>::kSmemBytes, stream>>>(plan->p, m_tiles, n_tiles, total);
  SN_CHECK_CUDA(cudaGetLastError());
  return SN_OK;
}

int sn_tap_gemm_plan_launch(const TapGemmPlan* plan, cudaStream_t stream) {
  const int cw = plan->p.a_chunk;
  if (plan->nsplit == 3) {
    if (cw == 64) return launch_tap<3, 64>(plan, stream);
    if (cw == 32) return launch_tap<3, 32>(plan, stream);
    return launch_tap<3, 16>(plan, stream);
  }
  if (cw == 64) return launch_tap<1, 64>(plan, stream);
  if (cw == 32) return launch_tap<1, 32>(plan, stream);
  return launch_tap<1, 16>(plan, stream);
}

int sn_wgrad_plan_init(WgradPlan* plan, const sn_wgrad_desc* d, int sm_count) {
  memset(plan, 0, sizeof(*plan));
  WgradParams& p = plan->p;
  SN_REQUIRE(d->nsplit == 1 || d->nsplit == 3, "nsplit must be 1 or 3");
  SN_REQUIRE(d->ntaps >= 1 && d->ntaps <= SN_MAX_TAPS, "ntaps out of range: %d", d->ntaps);
  const int y_chunk = d->y_chunk ? d->y_chunk : 64;
  SN_REQUIRE(y_chunk == 64 || y_chunk == 32 || y_chunk == 16, "y_chunk must be 64, 32 or 16");
  p.y_chunk = y_chunk;
  p.ngroups = 0;
  int grid_y = d->ntaps;
  if (y_chunk == 64) {
    SN_REQUIRE(d->block_n == 64 || d->block_n == 128, "wgrad block_n must be 64 or 128 for 64-channel Y rows");
  } else {
    SN_REQUIRE(d->cols_valid <= y_chunk, "narrow Y: cols_valid must fit one %d-channel block", y_chunk);
    if (d->ngroups > 0) {
      int maxg = 0, covered = 0;
      for (int g = 0; g < d->ngroups; ++g) {
        SN_REQUIRE(d->group_size[g] >= 1 && d->group_start[g] >= 0 &&
                       d->group_start[g] + d->group_size[g] <= d->ntaps, "bad tap group %d", g);
        if (d->group_size[g] > maxg) maxg = d->group_size[g];
        covered += d->group_size[g];
        p.gstart[g] = (short)d->group_start[g];
        p.gsize[g] = (short)d->group_size[g];
      }
      SN_REQUIRE(covered == d->ntaps && maxg * y_chunk <= 128 && d->block_n == maxg * y_chunk,
                 "tap groups must cover all taps and block_n == max group * y_chunk <= 128");
      p.ngroups = d->ngroups;
      grid_y = d->ngroups;
    } else {
      SN_REQUIRE(d->block_n == y_chunk, "narrow Y without groups: block_n must equal y_chunk");
    }
  }
  SN_REQUIRE(d->x_hi && d->y_hi && d->out, "null operand");
  SN_REQUIRE(d->nsplit == 1 || (d->x_lo && d->y_lo), "nsplit=3 needs lo planes");
  SN_REQUIRE(d->x_fmt == d->y_fmt, "X and Y of one tcgen05.mma must share a 16-bit format (x=%d y=%d)",
             d->x_fmt, d->y_fmt);
  int th, tw, nb;
  pick_patch(d->m_h, d->m_w, 64, &th, &tw, &nb);
  p.tw = tw; p.th = th; p.nb = nb;
  p.tiles_w = (d->m_w + tw - 1) / tw;
  p.tiles_h = (d->m_h + th - 1) / th;
  p.tiles_n = (d->m_n + nb - 1) / nb;
  p.ntaps = d->ntaps;
  for (int t = 0; t < d->ntaps; ++t) {
    p.xtaps[t].c_off = d->xtaps[t].c_off; p.xtaps[t].dw = (short)d->xtaps[t].dw;
    p.xtaps[t].dh = (short)d->xtaps[t].dh; p.xtaps[t].hp = (short)d->xtaps[t].hp;
    p.ytaps[t].c_off = d->ytaps[t].c_off; p.ytaps[t].dw = (short)d->ytaps[t].dw;
    p.ytaps[t].dh = (short)d->ytaps[t].dh; p.ytaps[t].hp = (short)d->ytaps[t].hp;
    p.tap_off[t] = d->tap_off[t];
  }
  p.block_n = d->block_n;
  p.m_tiles = (d->rows_valid + kBlockM - 1) / kBlockM;
  p.n_tiles = y_chunk == 64 ? (d->cols_valid + d->block_n - 1) / d->block_n : 1;
  p.rows_valid = d->rows_valid;
  p.cols_valid = d->cols_valid;
  p.out = d->out;
  p.s_row = d->s_row;
  p.s_col = d->s_col;
  p.x_fmt = d->x_fmt;
  p.y_fmt = d->y_fmt;
  int rc;
  const void* x_pl[2] = {d->x_hi, d->x_lo};
  const void* y_pl[2] = {d->y_hi, d->y_lo};
  {
    const char* e = getenv("SN_NO_MERGED_PLANES");
    const bool merge_ok = !(e && e[0] == '1');
    const long long x_ps = d->nsplit == 3 ? (const char*)d->x_lo - (const char*)d->x_hi : 0;
    const long long y_ps = d->nsplit == 3 ? (const char*)d->y_lo - (const char*)d->y_hi : 0;
    p.x_merged = merge_ok && x_ps > 0 && x_ps % 16 == 0 && tw * th * nb == 64;
    p.y_merged = merge_ok && y_ps > 0 && y_ps % 16 == 0 && tw * th * nb == 64 && y_chunk == 64 && d->ngroups == 0;
    for (int t = 0; t < d->ntaps; ++t) {   // h parity -> channel coordinate of the merged parity maps
      if (p.x_merged && d->x_parity) p.xtaps[t].c_off += p.xtaps[t].hp * d->x_w * d->x_pitch;
      if (p.y_merged && d->y_parity) p.ytaps[t].c_off += p.ytaps[t].hp * d->y_w * d->y_pitch;
    }
    for (int pl = 0; pl < (d->nsplit == 3 ? 2 : 1); ++pl) {
      if (!(p.x_merged && pl == 1)) {
        rc = sn_make_act_map(&p.tmX[pl], x_pl[pl], d->x_n, d->x_h, d->x_w, d->x_c, d->x_pitch,
                             d->x_parity, tw, th, nb, 64, p.x_merged ? x_ps : 0);
        if (rc) return rc;
      }
      if (!(p.y_merged && pl == 1)) {
        rc = sn_make_act_map(&p.tmY[pl], y_pl[pl], d->y_n, d->y_h, d->y_w, d->y_c, d->y_pitch,
                             d->y_parity, tw, th, nb, y_chunk, p.y_merged ? y_ps : 0);
        if (rc) return rc;
      }
    }
    if (p.x_merged) p.tmX[1] = p.tmX[0];
    if (p.y_merged) p.tmY[1] = p.tmY[0];
  }
  plan->nsplit = d->nsplit;
  const int base_ctas = p.m_tiles * p.n_tiles * grid_y;
  const int total = p.tiles_w * p.tiles_h * p.tiles_n;
  p.rot_mode = 0;
  if (const char* e = getenv("SN_WGRAD_ROT")) p.rot_mode = atoi(e);
  int ks = d->ksplit;
  if (const char* e = getenv("SN_WGRAD_KSPLIT")) {   // experiment override
    if (atoi(e) > 0) ks = atoi(e);
  }
  if (ks <= 0) {  // aim for ~3 waves, at least 8 k-iterations per CTA
    ks = (3 * sm_count + base_ctas - 1) / base_ctas;
    int max_ks = total / 8;
    if (max_ks < 1) max_ks = 1;
    if (ks > max_ks) ks = max_ks;
    if (ks < 1) ks = 1;
  }
  if (ks > total) ks = total;
  plan->grid = dim3(p.m_tiles * p.n_tiles, grid_y, ks);
  return SN_OK;
}

int sn_wgrad_plan_launch(const WgradPlan* plan, cudaStream_t stream) {
  const int idx = plan->nsplit == 3 ? 1 : 0;
  if (plan->nsplit == 3) {
    if (!g_smem_attr_done[1][idx]) {
      SN_CHECK_CUDA(cudaFuncSetAttribute(wgrad_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg<3>::kSmemBytes));
      g_smem_attr_done[1][idx] = 1;
    }
    wgrad_gemm_kernel<3><<<plan->grid, kThreads, Cfg<3>::kSmemBytes, stream>>>(plan->p);
  } else {
    if (!g_smem_attr_done[1][idx]) {
      SN_CHECK_CUDA(cudaFuncSetAttribute(wgrad_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg<1>::kSmemBytes));
      g_smem_attr_done[1][idx] = 1;
    }
    wgrad_gemm_kernel<1><<<plan->grid, kThreads, Cfg<1>::kSmemBytes, stream>>>(plan->p);
  }
  SN_CHECK_CUDA(cudaGetLastError());
  return SN_OK;
}
