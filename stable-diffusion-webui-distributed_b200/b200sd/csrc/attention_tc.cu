// attention_tc.cu — flash-attention forward (non-causal) on tcgen05 for sm_100a.
//
//   O[b, s, h, :] = softmax(Q[b, s, h, :] . K[b, :, h, :]^T * scale) . V[b, :, h, :]
//
// Covers the UNet's self-attention (S = 4096/1024/256/64, d = 40/80/160) and cross-attention (77 context
// tokens) — upstream ldm CrossAttention (SURVEY.md §8 a-ext x6, x7; not in /root/reference).
//
// Layout: Q/K/V rows are tokens; every head owns d_pad (multiple of 64) consecutive halfs, the first d of
// which are data and the rest zero (the projection GEMM produces this directly from zero-padded weight
// rows), so each 64-half chunk of a tile is exactly one TMA SWIZZLE_128B box.  O is written unpadded
// ([b, s, h*d]) because it feeds the out-projection GEMM as a plain K-major A operand.
//
// One CTA = one 128-row Q tile of one (batch, head), kv consumed in tiles of 64.  320 threads:
//   warp 0    TMA producer (Q once; K ring, V ring)
//   warp 1    TMEM allocator + single-thread tcgen05.mma issuer:  S[j%2] = Q K_j^T (M128 x N<=64 x K=d),
//             O (+)= P[j%2] V_j (M128 x N=d_pad x K<=64, V consumed MN-major straight from its TMA tile)
//   warps 2-9 softmax.  Thread pair == query row: TMEM lane = (warp%4)*32 + lane, and the two warps of a lane
//             quarter split the 64 kv columns of the tile in halves.
// S and P are DOUBLE BUFFERED (S in TMEM, P in shared memory): Q K_{j+2}^T is issued as soon as softmax j has
// consumed its S buffer, and P_j V_j runs while softmax j+1 is already exponentiating — in steady state the softmax
// warps never wait for the tensor pipe, which matters because for d = 40 this kernel is bound by the exp (MUFU)
// rate, not by the MMAs (16 exps per 96 MMA-FLOPs).
// Softmax is single pass ("lazy max"): P = exp2(S*scale - m_used) uses the running maximum of earlier tiles while the
// tile's own maximum is tracked; only if that exceeds m_used by more than 2^8 (P would leave fp16's comfortable
// range) is the tile redone after rescaling O in TMEM — a vote between the two warps of a row quarter, rare after
// the first tile.  Each S element is read from TMEM once, the TMEM load of the next 16 columns is in flight while
// the current 16 are exponentiated, and with a ones column in V (v_ones_col) the row sums come out of the P.V MMA
// instead of CUDA-core adds.  d <= 64: 80 KB smem + 256 TMEM columns per CTA -> two CTAs per SM.
#include "tc_common.cuh"
#include "b200sd_internal.h"

namespace b200sd {

constexpr int kSoftmaxWarps = 8;
constexpr int kSoftmaxThreads = 32 * kSoftmaxWarps;
constexpr int kAttnThreads = 64 + kSoftmaxThreads;
constexpr int kQTile = 128;
constexpr int kKv = 64;                              // kv rows per tile
constexpr uint32_t kQChunkBytes = kQTile * 128;      // 128 rows x 64 halfs
constexpr uint32_t kKvChunkBytes = kKv * 128;        // 64 rows x 64 halfs
constexpr uint32_t kPBytes = kQTile * kKv * 2;       // one K-major SWIZZLE_128B atom: 128 rows x 64 halfs
constexpr int kMaxRing = 4;

struct AttnParams {
  int B, heads, Sq, Skv, d, d_pad;
  int d16;           // d rounded up to 16 (MMA K of Q.K^T; O columns that carry data)
  int dpv;           // MMA N of P.V = d_pad (whole 64-wide MN-major swizzle atoms; pad columns of V are zero)
  int chunks;        // d_pad / 64
  int k_stages, v_stages;  // K / V ring depths (<= kMaxRing), as deep as shared memory allows
  int tmem_cols;
  int l_col;         // >= 0: V carries a ones column at l_col (== d) and O[:, l_col] is the softmax denominator
  int resc_cols;     // O columns touched by a rescale (multiple of 16, covers l_col)
  float scale_log2;  // softmax scale * log2(e)
  void* O;
  long long ldo;
  int is_bf16;
};

struct __align__(16) AttnShared {
  uint64_t q_full;
  uint64_t k_full[kMaxRing], k_empty[kMaxRing];
  uint64_t v_full[kMaxRing], v_empty[kMaxRing];
  uint64_t s_full[2], p_full[2], o_full;
  uint32_t tmem_base;
  uint32_t pad;
  float xch[2][kQTile];  // row maxima / sums exchanged between the two column halves
};

template <bool kBf16>
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  if constexpr (kBf16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// named barriers 1..4: the two warps (64 threads) that share one TMEM lane quarter, i.e. the thread pairs of 32 rows
__device__ __forceinline__ void pair_bar_sync(int quarter) {
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}
__device__ __forceinline__ bool pair_bar_or(int quarter, bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.u32 q, %1, 0;\n\t"
      "bar.red.or.pred p, %2, 64, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(out)
      : "r"(static_cast<uint32_t>(pred)), "r"(quarter + 1)
      : "memory");
  return out != 0;
}

// maximum of my 32 columns [col0, col0+32) of the S tile (raw logits); kv columns >= nvalid ignored
template <bool kFull>
__device__ __forceinline__ float half_row_max(uint32_t tmem_row, int col0, int nvalid) {
  float mx = -INFINITY;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int cbase = col0 + s * 16;
    if (!kFull && cbase >= nvalid) break;
    uint32_t v[16];
    tmem_ld_x16(tmem_row + cbase, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float a = __uint_as_float(v[i]), b2 = __uint_as_float(v[i + 1]);
      if (!kFull) {
        if (cbase + i >= nvalid) a = -INFINITY;
        if (cbase + i + 1 >= nvalid) b2 = -INFINITY;
      }
      mx = fmax3(mx, a, b2);
    }
  }
  return mx;
}

// exponentiate 16 columns, track their maximum / sum, write them as two 16-byte chunks of the swizzled P atom
template <bool kFull, bool kBf16, bool kSum>
__device__ __forceinline__ void softmax16(const uint32_t (&v)[16], uint8_t* sPj, int r, int cbase, int nvalid,
                                          float scale_log2, float m_used, float& tile_max, float& lsum) {
  uint32_t pk[8];
  float mx = tile_max, acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    float a = __uint_as_float(v[i]), b2 = __uint_as_float(v[i + 1]);
    float ea = fast_exp2(fmaf(a, scale_log2, -m_used));
    float eb = fast_exp2(fmaf(b2, scale_log2, -m_used));
    if (!kFull) {
      if (cbase + i >= nvalid) { a = -INFINITY; ea = 0.f; }
      if (cbase + i + 1 >= nvalid) { b2 = -INFINITY; eb = 0.f; }
    }
    mx = fmax3(mx, a, b2);
    if constexpr (kSum) acc += ea + eb;
    pk[i >> 1] = pack_h2<kBf16>(ea, eb);
  }
  tile_max = mx;
  if constexpr (kSum) lsum += acc;
  const uint32_t chunk16 = static_cast<uint32_t>(cbase >> 3);  // 8 halfs per 16-byte chunk
  *reinterpret_cast<uint4*>(sPj + sw128_offset(r, chunk16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4*>(sPj + sw128_offset(r, chunk16 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}

// One pass over my 32 columns: the TMEM load of the second 16 is in flight while the first 16 are processed.
template <bool kFull, bool kBf16, bool kSum>
__device__ __forceinline__ void softmax_half(uint32_t tmem_row, uint8_t* sPj, int r, int col0, int nvalid,
                                             float scale_log2, float m_used, float& tile_max, float& lsum) {
  uint32_t va[16], vb[16];
  tmem_ld_x16(tmem_row + col0, va);
  tmem_ld_wait();
  tmem_ld_x16(tmem_row + col0 + 16, vb);
  softmax16<kFull, kBf16, kSum>(va, sPj, r, col0, nvalid, scale_log2, m_used, tile_max, lsum);
  tmem_ld_wait();
  softmax16<kFull, kBf16, kSum>(vb, sPj, r, col0 + 16, nvalid, scale_log2, m_used, tile_max, lsum);
}

template <bool kBf16, bool kSum>
__device__ __forceinline__ void softmax_warps(const AttnParams& p, AttnShared* sh, uint8_t* sP, uint32_t tmem_base,
                                              int warp, int lane, int q0, int head, int b, int nkv) {
  const int quarter = warp & 3;
  const int half = (warp - 2) >> 2;   // which 32-column half of the kv tile is mine
  const int r = quarter * 32 + lane;  // query row in the tile == TMEM lane
  const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
  const uint32_t o_row = tmem_base + 128 + lane_base;
  const int col0 = half * 32;
  float m_used = -INFINITY;  // scaled log2 domain
  float l = 0.f;
  for (int j = 0; j < nkv; ++j) {
    const int buf = j & 1;
    const int nvalid = min(kKv, p.Skv - j * kKv);
    const bool full = nvalid == kKv;
    const uint32_t s_row = tmem_base + buf * kKv + lane_base;
    uint8_t* sPj = sP + buf * kPBytes;
    mbar_wait(&sh->s_full[buf], (j >> 1) & 1, 17);
    tc_fence_after();
    if (j == 0) {
      const float mx = full ? half_row_max<true>(s_row, col0, nvalid) : half_row_max<false>(s_row, col0, nvalid);
      sh->xch[half][r] = mx;
      pair_bar_sync(quarter);
      m_used = fmaxf(sh->xch[0][r], sh->xch[1][r]) * p.scale_log2;
    }
    float tile_max = -INFINITY, lsum = 0.f;
    if (full) softmax_half<true, kBf16, kSum>(s_row, sPj, r, col0, nvalid, p.scale_log2, m_used, tile_max, lsum);
    else      softmax_half<false, kBf16, kSum>(s_row, sPj, r, col0, nvalid, p.scale_log2, m_used, tile_max, lsum);
    const float tm = tile_max * p.scale_log2;
    if (pair_bar_or(quarter, tm > m_used + 8.0f)) {
      // rare path: the running maximum moved by more than 2^8 for some row of this quarter
      sh->xch[half][r] = tm;
      pair_bar_sync(quarter);
      const float m_new = fmax3(m_used, sh->xch[0][r], sh->xch[1][r]);
      const float alpha = fast_exp2(m_used - m_new);
      if (j > 0) {
        mbar_wait(&sh->o_full, (j - 1) & 1, 18);  // P.V of the previous tile must have landed in O
        tc_fence_after();
        for (int c = half; c < p.resc_cols / 16; c += 2) {
          uint32_t o[16];
          tmem_ld_x16(o_row + c * 16, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_x16(o_row + c * 16, o);
        }
        tmem_st_wait();
      }
      l *= alpha;
      m_used = m_new;
      tile_max = -INFINITY;
      lsum = 0.f;
      if (full) softmax_half<true, kBf16, kSum>(s_row, sPj, r, col0, nvalid, p.scale_log2, m_used, tile_max, lsum);
      else      softmax_half<false, kBf16, kSum>(s_row, sPj, r, col0, nvalid, p.scale_log2, m_used, tile_max, lsum);
    }
    l += lsum;
    // Observe every phase of o_full (P.V of the previous tile: issued a whole softmax ago, normally complete) so that
    // parity waits on it can never alias an older phase.  It must happen BEFORE this tile's arrive: P.V of this tile
    // cannot complete (and flip the phase again) until all 256 arrivals are in.
    if (j > 0) mbar_wait(&sh->o_full, (j - 1) & 1, 20);
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(&sh->p_full[buf]);
  }
  // ---- epilogue: O / l -> global (the pair splits the 16-column chunks of O) ----
  mbar_wait(&sh->o_full, (nkv - 1) & 1, 19);
  tc_fence_after();
  if constexpr (kSum) {
    pair_bar_sync(quarter);  // both threads of every pair are past their last xch read
    sh->xch[half][r] = l;
    pair_bar_sync(quarter);
    l = sh->xch[0][r] + sh->xch[1][r];
  } else {
    uint32_t o[16];
    tmem_ld_x16(o_row + (p.l_col / 16) * 16, o);
    tmem_ld_wait();
    l = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i == (p.l_col & 15)) l = __uint_as_float(o[i]);
  }
  const float inv_l = 1.0f / l;
  const int srow = q0 + r;
  const bool valid = srow < p.Sq;
  uint8_t* orow = reinterpret_cast<uint8_t*>(p.O) +
                  ((static_cast<long long>(b) * p.Sq + (valid ? srow : 0)) * p.ldo + static_cast<long long>(head) * p.d) * 2;
  for (int c = half; c < p.d16 / 16; c += 2) {
    uint32_t o[16];
    tmem_ld_x16(o_row + c * 16, o);
    tmem_ld_wait();
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      h[i] = pack_h2<kBf16>(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
    if (valid) {
      if (c * 16 + 8 <= p.d) *reinterpret_cast<uint4*>(orow + c * 32) = make_uint4(h[0], h[1], h[2], h[3]);
      if (c * 16 + 16 <= p.d) *reinterpret_cast<uint4*>(orow + c * 32 + 16) = make_uint4(h[4], h[5], h[6], h[7]);
    }
  }
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t q_bytes = static_cast<uint32_t>(p.chunks) * kQChunkBytes;
  const uint32_t kv_bytes = static_cast<uint32_t>(p.chunks) * kKvChunkBytes;
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + q_bytes;  // 2 x 16 KB
  uint8_t* sK = sP + 2 * kPBytes;
  uint8_t* sV = sK + p.k_stages * kv_bytes;
  AttnShared* sh = reinterpret_cast<AttnShared*>(sV + p.v_stages * kv_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQTile;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int nkv = (p.Skv + kKv - 1) / kKv;
  const int col0 = head * p.d_pad;

  // K / V loads are issued by one thread that polls both rings, so a full V ring never delays a K load (or vice
  // versa); the first loads go out before the block-wide sync, overlapping the TMEM allocation.
  int next_k = 0, next_v = 0;
  auto issue_k = [&](int j) {
    const int st = j % p.k_stages;
    mbar_arrive_expect_tx(&sh->k_full[st], kv_bytes);
    for (int c = 0; c < p.chunks; ++c)
      tma_load_3d(sK + st * kv_bytes + c * kKvChunkBytes, &tmK, &sh->k_full[st], col0 + c * 64, j * kKv, b);
  };
  auto issue_v = [&](int j) {
    const int st = j % p.v_stages;
    mbar_arrive_expect_tx(&sh->v_full[st], kv_bytes);
    for (int c = 0; c < p.chunks; ++c)
      tma_load_3d(sV + st * kv_bytes + c * kKvChunkBytes, &tmV, &sh->v_full[st], col0 + c * 64, j * kKv, b);
  };
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&sh->q_full, 1);
    for (int s = 0; s < kMaxRing; ++s) {
      mbar_init(&sh->k_full[s], 1);
      mbar_init(&sh->k_empty[s], 1);
      mbar_init(&sh->v_full[s], 1);
      mbar_init(&sh->v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sh->s_full[s], 1);
      mbar_init(&sh->p_full[s], kSoftmaxThreads);
    }
    mbar_init(&sh->o_full, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(&sh->q_full, q_bytes);
    for (int c = 0; c < p.chunks; ++c) tma_load_3d(sQ + c * kQChunkBytes, &tmQ, &sh->q_full, col0 + c * 64, q0, b);
    for (; next_k < nkv && next_k < p.k_stages; ++next_k) issue_k(next_k);
    for (; next_v < nkv && next_v < p.v_stages; ++next_v) issue_v(next_v);
  }
  if (warp == 1) tmem_alloc(&sh->tmem_base, static_cast<uint32_t>(p.tmem_cols));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;  // S0 @ +0, S1 @ +64 (64 fp32 columns each), O @ +128 (dpv columns)
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    // ------------------------------------ TMA producer ------------------------------------
    if (lane == 0) {
      uint32_t spins = 0;
      while (next_k < nkv || next_v < nkv) {
        bool progress = false;
        if (next_k < nkv && mbar_try_wait(&sh->k_empty[next_k % p.k_stages], ((next_k / p.k_stages) & 1) ^ 1u)) {
          issue_k(next_k++);
          progress = true;
        }
        if (next_v < nkv && mbar_try_wait(&sh->v_empty[next_v % p.v_stages], ((next_v / p.v_stages) & 1) ^ 1u)) {
          issue_v(next_v++);
          progress = true;
        }
        if (progress) spins = 0;
        else if (++spins > B200SD_SPIN_LIMIT) {
          printf("b200sd: attention producer timeout block=(%d,%d,%d) k=%d v=%d\n", blockIdx.x, blockIdx.y, blockIdx.z,
                 next_k, next_v);
          __trap();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ---------------------------------------
    if (lane == 0) {
      const bool bf = p.is_bf16 != 0;
      const int ksteps_qk = p.d16 / 16;
      const uint32_t aQ = smem_u32(sQ);
      auto issue_qk = [&](int j) {  // S[j&1] = Q K_j^T
        const int st = j & 1;                // S buffer
        const int ks = j % p.k_stages;       // K ring slot
        const int nvalid = min(kKv, p.Skv - j * kKv);
        const int n16 = (nvalid + 15) & ~15;
        mbar_wait(&sh->k_full[ks], (j / p.k_stages) & 1, 14);
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16(128, n16, bf, false, false);
        const uint32_t aK = smem_u32(sK + ks * kv_bytes);
        for (int k = 0; k < ksteps_qk; ++k) {
          const uint32_t offq = static_cast<uint32_t>(k >> 2) * kQChunkBytes + static_cast<uint32_t>(k & 3) * 32u;
          const uint32_t offk = static_cast<uint32_t>(k >> 2) * kKvChunkBytes + static_cast<uint32_t>(k & 3) * 32u;
          umma_f16_ss(tmem_base + st * kKv, make_sdesc_sw128(aQ + offq, 16, 1024), make_sdesc_sw128(aK + offk, 16, 1024),
                      idesc, k != 0 ? 1u : 0u);
        }
        umma_commit(&sh->k_empty[ks]);
        umma_commit(&sh->s_full[st]);
      };
      mbar_wait(&sh->q_full, 0, 13);
      issue_qk(0);
      if (nkv > 1) issue_qk(1);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int nvalid = min(kKv, p.Skv - j * kKv);
        const int n16 = (nvalid + 15) & ~15;
        // ---- O (+)= P_j V_j ----
        const int vs = j % p.v_stages;
        mbar_wait(&sh->p_full[st], ph, 15);
        mbar_wait(&sh->v_full[vs], (j / p.v_stages) & 1, 16);
        tc_fence_after();
        {
          const uint32_t idesc = make_idesc_f16(128, p.dpv, bf, false, true);  // B (= V) is MN-major
          const uint32_t aP = smem_u32(sP + st * kPBytes);
          const uint32_t aV = smem_u32(sV + vs * kv_bytes);
          const int ksteps_pv = n16 / 16;
          for (int k = 0; k < ksteps_pv; ++k) {
            const uint32_t offP = static_cast<uint32_t>(k) * 32u;    // 16 halfs inside the 128-byte swizzle row
            const uint32_t offV = static_cast<uint32_t>(k) * 2048u;  // 16 kv rows x 128 B
            umma_f16_ss(tmem_O, make_sdesc_sw128(aP + offP, 16, 1024), make_sdesc_sw128(aV + offV, kKvChunkBytes, 1024),
                        idesc, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&sh->v_empty[vs]);
          umma_commit(&sh->o_full);
        }
        // ---- softmax j has released S[st]: refill it two tiles ahead ----
        if (j + 2 < nkv) issue_qk(j + 2);
      }
    }
  } else {
    const bool sum_here = p.l_col < 0;
    if (p.is_bf16) {
      if (sum_here) softmax_warps<true, true>(p, sh, sP, tmem_base, warp, lane, q0, head, b, nkv);
      else          softmax_warps<true, false>(p, sh, sP, tmem_base, warp, lane, q0, head, b, nkv);
    } else {
      if (sum_here) softmax_warps<false, true>(p, sh, sP, tmem_base, warp, lane, q0, head, b, nkv);
      else          softmax_warps<false, false>(p, sh, sP, tmem_base, warp, lane, q0, head, b, nkv);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------------
static int g_attn_max_smem = 0;
static bool g_attn_dev_ready[64] = {};

int attention_tc(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, void* O,
                 long long ldo, int B, int heads, int Sq, int Skv, int d, int d_pad, float scale, int v_ones_col,
                 int is_bf16, cudaStream_t stream) {
  if (B <= 0 || heads <= 0 || Sq <= 0) return B200SD_OK;
  if (Skv <= 0 || d <= 0 || d % 8 != 0 || d_pad % 64 != 0 || d_pad < d || d > 240) return B200SD_ERR_INVALID;
  if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8) return B200SD_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(K) | reinterpret_cast<uintptr_t>(V) |
       reinterpret_cast<uintptr_t>(O)) & 15)
    return B200SD_ERR_INVALID;
  if (ldq < static_cast<long long>(heads) * d_pad || ldk < static_cast<long long>(heads) * d_pad ||
      ldv < static_cast<long long>(heads) * d_pad || ldo < static_cast<long long>(heads) * d)
    return B200SD_ERR_INVALID;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return B200SD_ERR_CUDA;
    if (dev < 0 || dev >= 64) return B200SD_ERR_UNSUPPORTED;
    if (!g_attn_dev_ready[dev]) {  // per-device opt-in to large dynamic smem; first call must be outside capture
      int smem = 0;
      if (cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
        return B200SD_ERR_CUDA;
      if (cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return B200SD_ERR_CUDA;
      g_attn_max_smem = smem;
      g_attn_dev_ready[dev] = true;
    }
  }
  AttnParams p{};
  p.B = B; p.heads = heads; p.Sq = Sq; p.Skv = Skv; p.d = d; p.d_pad = d_pad;
  p.d16 = (d + 15) & ~15;
  p.chunks = d_pad / 64;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.O = O; p.ldo = ldo; p.is_bf16 = is_bf16;
  p.dpv = d_pad;
  if (p.dpv > 256) return B200SD_ERR_UNSUPPORTED;
  if (v_ones_col && d >= d_pad) return B200SD_ERR_INVALID;  // needs a free pad column
  p.l_col = v_ones_col ? d : -1;
  p.resc_cols = v_ones_col ? ((d + 1 + 15) & ~15) : p.d16;
  p.tmem_cols = (128 + p.dpv <= 256) ? 256 : 512;
  // ring depths: two CTAs per SM when a Q tile + both P buffers + 4 K + 3 V tiles fit in half an SM (d <= 64),
  // otherwise whatever one CTA can hold
  const size_t kvt = static_cast<size_t>(p.chunks) * kKvChunkBytes;
  const size_t fixed = 1024 + static_cast<size_t>(p.chunks) * kQChunkBytes + 2 * kPBytes + sizeof(AttnShared) + 64;
  const size_t half_sm = static_cast<size_t>(g_attn_max_smem) / 2 - 1024;
  size_t budget = (p.tmem_cols <= 256 && fixed + 4 * kvt <= half_sm) ? half_sm : static_cast<size_t>(g_attn_max_smem);
  if (fixed + 2 * kvt > budget) return B200SD_ERR_UNSUPPORTED;
  int total = static_cast<int>((budget - fixed) / kvt);
  if (total > 2 * kMaxRing - 1) total = 2 * kMaxRing - 1;
  p.k_stages = (total + 1) / 2;
  p.v_stages = total / 2;
  if (p.v_stages < 1) return B200SD_ERR_UNSUPPORTED;
  const size_t smem = fixed + static_cast<size_t>(p.k_stages + p.v_stages) * kvt;
  CUtensorMap tmQ, tmK, tmV;
  const uint32_t es[3] = {1, 1, 1};
  int rc;
  {
    const uint32_t box[3] = {64, kQTile, 1};
    const uint64_t dims[3] = {static_cast<uint64_t>(heads) * d_pad, static_cast<uint64_t>(Sq), static_cast<uint64_t>(B)};
    const uint64_t st[2] = {static_cast<uint64_t>(ldq) * 2, static_cast<uint64_t>(ldq) * 2 * Sq};
    if ((rc = make_tmap_sw128(&tmQ, Q, 3, dims, st, box, es)) != B200SD_OK) return rc;
  }
  const uint32_t kvbox[3] = {64, kKv, 1};
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(heads) * d_pad, static_cast<uint64_t>(Skv), static_cast<uint64_t>(B)};
    const uint64_t st[2] = {static_cast<uint64_t>(ldk) * 2, static_cast<uint64_t>(ldk) * 2 * Skv};
    if ((rc = make_tmap_sw128(&tmK, K, 3, dims, st, kvbox, es)) != B200SD_OK) return rc;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(heads) * d_pad, static_cast<uint64_t>(Skv), static_cast<uint64_t>(B)};
    const uint64_t st[2] = {static_cast<uint64_t>(ldv) * 2, static_cast<uint64_t>(ldv) * 2 * Skv};
    if ((rc = make_tmap_sw128(&tmV, V, 3, dims, st, kvbox, es)) != B200SD_OK) return rc;
  }
  dim3 grid((Sq + kQTile - 1) / kQTile, heads, B);
  attention_tc_kernel<<<grid, kAttnThreads, smem, stream>>>(tmQ, tmK, tmV, p);
  return cudaGetLastError() == cudaSuccess ? B200SD_OK : B200SD_ERR_CUDA;
}

}  // namespace b200sd
