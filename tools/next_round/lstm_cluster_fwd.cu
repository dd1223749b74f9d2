// DRAFT (untested, not part of the library build): one LSTM layer's forward recurrence on ONE 16-CTA cluster with the
// hidden state exchanged through distributed shared memory instead of global memory + a grid barrier.
// See NOTES_NEXT.md ("LSTM recurrence on ONE 16-CTA cluster").  Compile check only:
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v -c tools/next_round/lstm_cluster_fwd.cu -o /tmp/x.o
//
// Layout: CTA r of 16 owns hidden units [34 r, 34 r + 34) (H <= 544), i.e. 136 gate columns ordered n = 4*u + gate so
// that the four gates of a unit sit in adjacent accumulator columns (one lane-pair exchange instead of a shared-memory
// transpose).  17 warps: warp w multiplies the whole h tile (M = 32, K = 544) by n-tile w.  W_hh slice [136][552] bf16
// (150 KB) and the double-buffered h tile [2][32][552] bf16 (71 KB) live in shared memory.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tbx {

constexpr int kCluster = 16;
constexpr int kU = 34;                 // units per CTA (even: 68-byte column blocks are 4-byte aligned)
constexpr int kHp = kCluster * kU;     // 544 padded hidden size = K
constexpr int kKs = kHp + 8;           // shared-memory row stride (conflict-free ldmatrix)
constexpr int kNloc = 4 * kU;          // 136 gate columns per CTA = 17 n8 tiles
constexpr int kWarps = kNloc / 8;      // 17
constexpr int kThreads = kWarps * 32;  // 544

struct ClusterFwdArgs {
  const float* w_hh;     // [4H, H] fp32, torch gate order i,f,g,o
  float* gates;          // [T1*B, 4H]: in = hoisted x-projection + biases, out = activated gates
  float* hs; float* cs;  // [T1*B, H]
  float* cm;             // [T1*B, H]  c_{t-1} * notdone_t (row 0 pre-initialised)
  __nv_bfloat16* hmq;    // [T1*B, Hq] masked recurrent inputs for the weight-gradient GEMM (row block 0 pre-initialised)
  const float* h_init;   // [B, H]
  const float* nd;       // [T1*B]
  int T1, B, H, Hq;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t caddr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(caddr), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1) lstm_cluster_fwd_kernel(ClusterFwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(smem);                              // [136][kKs]
  __nv_bfloat16* Ht = Ws + kNloc * kKs;                                                    // [2][32][kKs]
  __nv_bfloat16* Hb = Ht + 2 * 32 * kKs;                                                   // [32][kU] this step's h block
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int H = a.H, B = a.B;
  const int rank = int(cluster_rank());
  const int u0 = rank * kU;
  const int rows = B < 32 ? B : 32;
  // weights: local column n = 4*u + g  <->  W_hh row g*H + (u0 + u); zero beyond H
  for (int i = tid; i < kNloc * kHp; i += kThreads) {
    const int n = i / kHp, k = i - n * kHp;
    const int u = n >> 2, g = n & 3;
    float v = 0.f;
    if (u0 + u < H && k < H) v = a.w_hh[(int64_t(g) * H + u0 + u) * H + k];
    Ws[n * kKs + k] = __float2bfloat16_rn(v);
  }
  // initial state tile (raw h; the done-mask is applied to the product): every CTA builds the full tile itself
  for (int i = tid; i < 32 * kHp; i += kThreads) {
    const int r = i / kHp, k = i - r * kHp;
    float v = 0.f;
    if (r < rows && k < H) v = a.h_init[int64_t(r) * H + k];
    Ht[r * kKs + k] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  cluster_sync_all();
  // accumulator ownership: rows ra = lane>>2 (+8, +16, +24), columns c0 = (lane&3)*2, c0+1 of n-tile `wrp`:
  // unit ul = wrp*2 + ((lane&3)>>1), gates (0,1) on even (lane&1) lanes, (2,3) on odd lanes
  const int ul = wrp * 2 + ((lane & 3) >> 1);
  const int ug = u0 + ul;
  const bool uok = ug < H;
  const int gpair = (lane & 1) * 2;
  float c_prev[4] = {0.f, 0.f, 0.f, 0.f};  // even lanes: cell state of (row ra + 8*j, unit ul)
  for (int t = 0; t < a.T1; ++t) {
    const __nv_bfloat16* Hc = Ht + (t & 1) * 32 * kKs;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 2
    for (int ks = 0; ks < kHp / 16; ++ks) {
      uint32_t bq[2];
      ldmatrix_x2(bq, Ws + (wrp * 8 + (lane & 7)) * kKs + ks * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t af[4];
        ldmatrix_x4(af, Hc + (mt * 16 + (lane & 15)) * kKs + ks * 16 + (lane >> 4) * 8);
        mma16816(acc[mt], af, bq);
      }
    }
    // gates of (row, unit): pre-activation = x-projection + notdone_t * recurrent product
    float hv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = (lane >> 2) + 8 * j;
      const int mt = j >> 1, hi = (j & 1) * 2;
      float g0 = 0.f, g1 = 0.f;
      const bool ok = uok && row < rows;
      const int64_t gr = (int64_t(t) * B + row) * 4 * H;
      if (ok) {
        const float ndv = a.nd[int64_t(t) * B + row];
        g0 = a.gates[gr + int64_t(gpair) * H + ug] + ndv * acc[mt][hi];
        g1 = a.gates[gr + int64_t(gpair + 1) * H + ug] + ndv * acc[mt][hi + 1];
        if (gpair == 0) { g0 = sigm(g0); g1 = sigm(g1); } else { g0 = tanhf(g0); g1 = sigm(g1); }
        a.gates[gr + int64_t(gpair) * H + ug] = g0;
        a.gates[gr + int64_t(gpair + 1) * H + ug] = g1;
      }
      const float o0 = __shfl_xor_sync(0xffffffffu, g0, 1), o1 = __shfl_xor_sync(0xffffffffu, g1, 1);
      float hnew = 0.f;
      if ((lane & 1) == 0 && ok) {  // even lane: (i, f) own, (g, o) from the odd neighbour
        const int64_t o = (int64_t(t) * B + row) * H + ug;
        const float cm = a.cm[o];
        const float cn = g1 * cm + g0 * o0;
        hnew = o1 * tanhf(cn);
        c_prev[j] = cn;
        a.cs[o] = cn; a.hs[o] = hnew;
        if (t + 1 < a.T1) {
          const float ndn = a.nd[int64_t(t + 1) * B + row];
          a.cm[o + int64_t(B) * H] = cn * ndn;
          a.hmq[(int64_t(t + 1) * B + row) * a.Hq + ug] = __float2bfloat16_rn(hnew * ndn);
        }
      }
      hv[j] = hnew;
    }
    if ((lane & 1) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) Hb[((lane >> 2) + 8 * j) * kU + ul] = __float2bfloat16_rn(hv[j]);
    }
    __syncthreads();
    if (t + 1 < a.T1) {
      // broadcast this CTA's [32][34] block into the next-step tile of all 16 CTAs: 17 words per row, one word per thread
      const int row = tid / (kU / 2), wcol = tid - row * (kU / 2);
      const uint32_t v = *reinterpret_cast<const uint32_t*>(Hb + row * kU + wcol * 2);
      const uint32_t local = smem_u32(Ht + ((t + 1) & 1) * 32 * kKs + row * kKs + u0 + wcol * 2);
#pragma unroll
      for (int p = 0; p < kCluster; ++p) st_cluster_u32(map_to_rank(local, p), v);
    }
    cluster_sync_all();
  }
  (void)c_prev;
}

// host: cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
//       cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
//       smem = (136*552 + 2*32*552 + 32*34) * 2 bytes = 223 KB; grid = 16 CTAs (one cluster), block = 544 threads
}  // namespace tbx
