// Implicit-GEMM first convolution on tensor cores: the uint8 NCHW frames ARE the operand.
//
// AtariNet's conv1 (monobeast.py:560, 8x8 stride 4 over [N,4,84,84] uint8) as a patch-matrix GEMM reads a
// [N*400, 256] bf16 matrix that is 7x larger than the frames it was gathered from; materialising it costs
// more HBM time than the product itself (profiles/launches_r1_summary.txt: im2col 229 us + GEMM 110 us +
// wgrad 112 us, all bound by the 531 MB patch matrix).  Here producer warps gather each patch row straight
// from the frame (L1/L2-resident: every input byte is reused by 4 patches), convert u8 -> bf16 exactly
// (pixel values 0..255 are integers <= 2^8; the 1/255 stays in the epilogue) and write it into shared
// memory in the SWIZZLE_128B layout the UMMA descriptors expect, so the patch matrix never exists in HBM:
//
//   forward : act[m, o] = relu(scale * sum_k patch[m, k] * W[o, k] + b[o])     A gathered (K-major), B = W resident
//   wgrad   : dW[o, k]  = scale * sum_m dY[m, o] * patch[m, k]                 A = dY^T via TMA (MN-major),
//                                                                             B gathered (MN-major), split over m
//
// One kernel-height row of a patch (KW = 8 pixels of one channel) is 8 bytes in the frame and becomes one
// 16-byte swizzle chunk; a (patch, channel) pair is one 128-byte swizzle row (KH*KW = 64 values).  The same
// physical rows serve as a K-major A tile (forward: 128 patches x 64 k) and as an MN-major B box (wgrad:
// 64 patches x 64 k-values) - only the descriptor differs.
// Requirements (checked by conv_u8_implicit_applicable): KH = KW = 8, C = 4, O = 32, W % 4 == 0, S % 4 == 0.
#include "conv_implicit.cuh"

#include "gemm_simt.cuh"  // splitk_reduce_kernel / GemmEpilogue
#include "tc_common.cuh"

namespace tb {

using namespace tcd;

namespace {

constexpr int kProducers = 256;                 // 8 gather warps
constexpr int kConvThreads = 192 + kProducers;  // warp 0: TMA, 1: MMA issue, 2-5: epilogue, 6-13: gather
constexpr int kC = 4;                           // input channels = k-blocks of 64 (= KH*KW)
constexpr int kO = 32;                          // output channels
constexpr int kStages = 4;

struct ConvGeom {
  int H, W, S, OH, OW;
  int64_t M;  // patches = N*OH*OW
};

// bytes (b0,b1) / (b2,b3) of w -> two exact bf16 values packed as bf16x2: 0x4B0000vv is the float 2^23 + v
__device__ __forceinline__ uint32_t u8pair_to_bf16x2_lo(uint32_t w) {
  const float f0 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650)) - 8388608.0f;
  const float f1 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7651)) - 8388608.0f;
  return __byte_perm(__float_as_uint(f0), __float_as_uint(f1), 0x7632);
}
__device__ __forceinline__ uint32_t u8pair_to_bf16x2_hi(uint32_t w) {
  const float f0 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7652)) - 8388608.0f;
  const float f1 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7653)) - 8388608.0f;
  return __byte_perm(__float_as_uint(f0), __float_as_uint(f1), 0x7632);
}
// 8 pixels (two words) -> one 16-byte chunk at a shared-memory address
__device__ __forceinline__ void store_chunk(uint32_t saddr, uint32_t w0, uint32_t w1) {
  const uint32_t a = u8pair_to_bf16x2_lo(w0), b = u8pair_to_bf16x2_hi(w0), c = u8pair_to_bf16x2_lo(w1),
                 d = u8pair_to_bf16x2_hi(w1);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// patch index -> byte offset of its top-left pixel in channel 0 (or -1 past the end)
__device__ __forceinline__ int64_t patch_origin(const ConvGeom& g, int64_t m) {
  if (m >= g.M) return -1;
  const int per = g.OH * g.OW;
  const int64_t n = m / per;
  const int rem = int(m - n * per);
  const int oy = rem / g.OW, ox = rem - oy * g.OW;
  return (n * kC * g.H + int64_t(oy) * g.S) * g.W + int64_t(ox) * g.S;
}

// ---- forward ---------------------------------------------------------------------------------------
// Persistent over 128-patch tiles.  Stage s of the A ring holds the tile's k-block of channel c (ring
// position runs on), filled by all 256 gather threads: thread p owns patch row p & 127 and kernel rows
// [4*(p>>7), +4).  The next tile's 32 words are loaded before the current tile is converted, so one
// global-load latency is exposed per tile, not per stage.
__global__ void __launch_bounds__(kConvThreads, 1)
conv_u8_fwd_implicit_kernel(const uint8_t* __restrict__ frame, const __grid_constant__ CUtensorMap tmB, TcEpilogue ep,
                            ConvGeom g, int tiles_m) {
  constexpr uint32_t B_BYTES = kO * kBlockK * 2;  // 4 KB per k-block
  constexpr uint32_t TMEM_COLS = 64;              // two 32-column accumulator buffers
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_addr(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + kStages * kABytes;
  const uint32_t bars = sB + kC * B_BYTES;  // full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], wfull
  const uint32_t tmem_slot = bars + 8 * (2 * kStages + 5);
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (kStages + s); };
  auto tmem_full = [&](int b) { return bars + 8u * (2 * kStages + b); };
  auto tmem_empty = [&](int b) { return bars + 8u * (2 * kStages + 2 + b); };
  const uint32_t wfull = bars + 8u * (2 * kStages + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full(s), kProducers); mbar_init(empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 4); }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    if (lane == 0) {  // the whole weight matrix [32, 256] stays resident: one box per k-block
      mbar_expect_tx(wfull, kC * B_BYTES);
      for (int c = 0; c < kC; ++c) tma_load_2d(sB + c * B_BYTES, &tmB, wfull, c * kBlockK, 0);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(kO >> 3) << 17) | (uint32_t(kBlockM >> 4) << 24);
      mbar_wait(wfull, 0);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < tiles_m; tile += gridDim.x, ++it) {
        const int ab = it & 1;
        mbar_wait(tmem_empty(ab), ((it >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + uint32_t(ab * kO);
        for (int c = 0; c < kC; ++c) {
          mbar_wait(full(stage), phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16(tacc, make_smem_desc(sA + stage * kABytes + k * 32), make_smem_desc(sB + c * B_BYTES + k * 32), idesc,
                      (c | k) != 0 ? 1u : 0u);
          umma_commit(empty(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tmem_full(ab));
      }
    }
  } else if (warp < 6) {
    const int quarter = warp & 3;
    float bias_r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias_r[j] = __ldg(ep.bias + j);
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles_m; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      mbar_wait(tmem_full(ab), (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t r = int64_t(tile) * kBlockM + quarter * 32 + lane;
      uint32_t v[32];
      tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(ab * kO), v);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty(ab));  // values are in registers: free the accumulator early
      if (r < g.M) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float x0 = __uint_as_float(v[j]) * ep.scale + bias_r[j];
          float x1 = __uint_as_float(v[j + 1]) * ep.scale + bias_r[j + 1];
          if (ep.relu) { x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f); }
          __nv_bfloat162 p = __floats2bfloat162_rn(x0, x1);
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&p);
        }
        uint4* c = reinterpret_cast<uint4*>(ep.C16 + r * ep.ldc16);
#pragma unroll
        for (int j = 0; j < 4; ++j) c[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      }
    }
  } else {
    const int p = threadIdx.x - 192;
    const int row = p & 127, kh0 = (p >> 7) * 4;
    const uint32_t row_off = uint32_t(row >> 3) * 1024u + uint32_t(row & 7) * 128u;
    const int rr = row & 7;
    uint32_t cur[kC][4][2], nxt[kC][4][2];
    auto load_tile = [&](int tile, uint32_t (&buf)[kC][4][2]) {
      const int64_t org = patch_origin(g, int64_t(tile) * kBlockM + row);
#pragma unroll
      for (int c = 0; c < kC; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (org >= 0) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(frame + org + (int64_t(c) * g.H + kh0 + j) * g.W);
            buf[c][j][0] = __ldg(src); buf[c][j][1] = __ldg(src + 1);
          } else {
            buf[c][j][0] = 0u; buf[c][j][1] = 0u;
          }
        }
    };
    int stage = 0; uint32_t phase = 0;
    if (int(blockIdx.x) < tiles_m) load_tile(blockIdx.x, cur);
    for (int tile = blockIdx.x; tile < tiles_m; tile += gridDim.x) {
      const int nt = tile + gridDim.x;
      if (nt < tiles_m) load_tile(nt, nxt);
#pragma unroll
      for (int c = 0; c < kC; ++c) {
        mbar_wait(empty(stage), phase ^ 1);
        const uint32_t dst = sA + stage * kABytes + row_off;
#pragma unroll
        for (int j = 0; j < 4; ++j) store_chunk(dst + (uint32_t((kh0 + j) ^ rr) << 4), cur[c][j][0], cur[c][j][1]);
        fence_async_smem();
        mbar_arrive(full(stage));
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
#pragma unroll
      for (int c = 0; c < kC; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) { cur[c][j][0] = nxt[c][j][0]; cur[c][j][1] = nxt[c][j][1]; }
    }
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- weight gradient -------------------------------------------------------------------------------
// D[o (padded to 128), k = 256] accumulates over this CTA's slice of 64-patch blocks; stage = A: two
// 64(o) x 64(patch) TMA boxes of dY (columns >= 32 are out of bounds -> zero), B: four 64(patch) x 64(k)
// gathered boxes (one per channel).  Gather thread p owns patch row p & 63 of channel p >> 6.
__global__ void __launch_bounds__(kConvThreads, 1)
conv_u8_wgrad_implicit_kernel(const uint8_t* __restrict__ frame, const __grid_constant__ CUtensorMap tmA, ConvGeom g,
                              float* __restrict__ partial, int total_kb, int per) {
  constexpr uint32_t B_BYTES = kC * 8192;  // 32 KB
  constexpr uint32_t TMEM_COLS = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_addr(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + kStages * kABytes;
  const uint32_t bars = sB + kStages * B_BYTES;  // full[kStages], empty[kStages], tmem_full
  const uint32_t tmem_slot = bars + 8 * (2 * kStages + 1);
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (kStages + s); };
  const uint32_t tmem_full = bars + 8u * (2 * kStages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb0 = blockIdx.x * per;
  const int kb1 = (kb0 + per < total_kb) ? kb0 + per : total_kb;
  const int num_kb = kb1 > kb0 ? kb1 - kb0 : 0;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full(s), kProducers + 1); mbar_init(empty(s), 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < num_kb; ++i) {
        const int kc = (kb0 + i) * kBlockK;  // first patch of the block
        mbar_wait(empty(stage), phase ^ 1);
        mbar_expect_tx(full(stage), kABytes);
        tma_load_2d(sA + stage * kABytes, &tmA, full(stage), 0, kc);
        tma_load_2d(sA + stage * kABytes + 8192, &tmA, full(stage), 64, kc);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && num_kb > 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (uint32_t(256 >> 3) << 17) |
                                 (uint32_t(kBlockM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full(stage), phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)
          umma_bf16(tmem_base, make_smem_desc_mn(sA + stage * kABytes + k * 2048), make_smem_desc_mn(sB + stage * B_BYTES + k * 2048),
                    idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(empty(stage));
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full);
    }
  } else if (warp < 6) {
    if ((warp & 3) == 0) {  // TMEM lanes 0..31 = the 32 real output channels
      if (num_kb > 0) {
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      float* pz = partial + (int64_t(blockIdx.x) * kO + lane) * (kC * 64);
#pragma unroll 1
      for (int c0 = 0; c0 < kC * 64; c0 += 32) {
        uint32_t v[32];
        if (num_kb > 0) {
          tmem_ld32(tmem_base + uint32_t(c0), v);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(pz + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                 __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
  } else {
    const int p = threadIdx.x - 192;
    const int row = p & 63, c = p >> 6;
    const uint32_t row_off = uint32_t(c) * 8192u + uint32_t(row >> 3) * 1024u + uint32_t(row & 7) * 128u;
    const int rr = row & 7;
    uint32_t cur[8][2], nxt[8][2];
    auto load_block = [&](int kb, uint32_t (&buf)[8][2]) {
      const int64_t org = patch_origin(g, int64_t(kb) * kBlockK + row);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (org >= 0) {
          const uint32_t* src = reinterpret_cast<const uint32_t*>(frame + org + (int64_t(c) * g.H + j) * g.W);
          buf[j][0] = __ldg(src); buf[j][1] = __ldg(src + 1);
        } else {
          buf[j][0] = 0u; buf[j][1] = 0u;
        }
      }
    };
    int stage = 0; uint32_t phase = 0;
    if (num_kb > 0) load_block(kb0, cur);
    for (int i = 0; i < num_kb; ++i) {
      if (i + 1 < num_kb) load_block(kb0 + i + 1, nxt);
      mbar_wait(empty(stage), phase ^ 1);
      const uint32_t dst = sB + stage * B_BYTES + row_off;
#pragma unroll
      for (int j = 0; j < 8; ++j) store_chunk(dst + (uint32_t(j ^ rr) << 4), cur[j][0], cur[j][1]);
      fence_async_smem();
      mbar_arrive(full(stage));
      if (++stage == kStages) { stage = 0; phase ^= 1; }
#pragma unroll
      for (int j = 0; j < 8; ++j) { cur[j][0] = nxt[j][0]; cur[j][1] = nxt[j][1]; }
    }
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

ConvGeom make_geom(int64_t N, int H, int W, int S) {
  ConvGeom g;
  g.H = H; g.W = W; g.S = S; g.OH = (H - 8) / S + 1; g.OW = (W - 8) / S + 1;
  g.M = N * g.OH * g.OW;
  return g;
}

}  // namespace

bool conv_u8_implicit_applicable(int C, int H, int W, int KH, int KW, int S, int O) {
  const char* e = getenv("TB_CONV1_IMPLICIT");
  if (e && e[0] == '0') return false;
  return C == kC && KH == 8 && KW == 8 && O == kO && (W % 4) == 0 && (S % 4) == 0 && H >= 8 && W >= 8;
}

int conv_u8_fwd_implicit(const uint8_t* frame, const void* w_bf16, int64_t N, int H, int W, int S, const TcEpilogue& ep,
                         cudaStream_t stream) {
  TB_REQUIRE(frame && w_bf16 && ep.C16 && ep.bias && ep.ldc16 == kO, "conv_u8_fwd_implicit: bad arguments");
  TB_REQUIRE((reinterpret_cast<uintptr_t>(frame) & 3) == 0 && (reinterpret_cast<uintptr_t>(ep.C16) & 15) == 0,
             "conv_u8_fwd_implicit: unaligned pointer");
  if (N == 0) return 0;
  ProfScope prof(ep.tag, stream);
  const ConvGeom g = make_geom(N, H, W, S);
  CUtensorMap mb;
  int rc = make_map(&mb, w_bf16, kO, kC * 64, kC * 64, kO);
  if (rc) return rc;
  constexpr size_t smem = 1024 + kStages * kABytes + kC * (kO * kBlockK * 2) + 8 * (2 * kStages + 5) + 16;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(conv_u8_fwd_implicit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    TB_REQUIRE(e == cudaSuccess, "conv_u8_fwd_implicit: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  const int64_t tiles = (g.M + kBlockM - 1) / kBlockM;
  TB_REQUIRE(tiles < (int64_t(1) << 31), "conv_u8_fwd_implicit: too many tiles");
  const int64_t grid = tiles < kNumSMsB200 ? tiles : kNumSMsB200;
  conv_u8_fwd_implicit_kernel<<<(unsigned)grid, kConvThreads, smem, stream>>>(frame, mb, ep, g, int(tiles));
  return check_launch("conv_u8_fwd_implicit_kernel");
}

int conv_u8_wgrad_implicit(const void* dy_bf16, const uint8_t* frame, int64_t N, int H, int W, int S, float* dW, float scale,
                           float* partial, int64_t partial_floats, const char* tag, cudaStream_t stream) {
  TB_REQUIRE(frame && dy_bf16 && dW && partial, "conv_u8_wgrad_implicit: null pointer");
  TB_REQUIRE((reinterpret_cast<uintptr_t>(frame) & 3) == 0, "conv_u8_wgrad_implicit: unaligned frame pointer");
  ProfScope prof(tag, stream);
  const ConvGeom g = make_geom(N, H, W, S);
  const int64_t total_kb = (g.M + kBlockK - 1) / kBlockK;
  TB_REQUIRE(total_kb >= 1 && total_kb < (int64_t(1) << 31), "conv_u8_wgrad_implicit: bad size");
  int64_t grid = total_kb < kNumSMsB200 ? total_kb : kNumSMsB200;
  const int64_t per = (total_kb + grid - 1) / grid;
  grid = (total_kb + per - 1) / per;
  TB_REQUIRE(grid * kO * kC * 64 <= partial_floats, "conv_u8_wgrad_implicit: partial buffer too small");
  CUtensorMap ma;  // dY [M, 32] with the patch index as the reduction (row) index: MN-major boxes of 64 x 64
  int rc = make_map(&ma, dy_bf16, g.M, kO, kO, kBlockK, 64);
  if (rc) return rc;
  constexpr size_t smem = 1024 + kStages * (kABytes + size_t(kC) * 8192) + 8 * (2 * kStages + 1) + 16;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(conv_u8_wgrad_implicit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    TB_REQUIRE(e == cudaSuccess, "conv_u8_wgrad_implicit: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  conv_u8_wgrad_implicit_kernel<<<(unsigned)grid, kConvThreads, smem, stream>>>(frame, ma, g, partial, int(total_kb), int(per));
  rc = check_launch("conv_u8_wgrad_implicit_kernel");
  if (rc) return rc;
  GemmEpilogue rep;
  rep.scale = scale;
  const int64_t total = int64_t(kO) * kC * 64;
  splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(partial, dW, kO, kC * 64, kC * 64, int(grid), rep);
  return check_launch("splitk_reduce_kernel");
}

}  // namespace tb
