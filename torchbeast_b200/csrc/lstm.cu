// Stacked LSTM over the unroll with per-step done-reset, forward + BPTT (sm_100a).
//
// Replaces the reference's 81 seq_len-1 nn.LSTM calls and their autograd graph
// (/root/reference/torchbeast/monobeast.py:603-611, polybeast_learner.py:241-249):
//     for t: state *= notdone_t;  out_t, state = LSTM(x_t, state)         (all layers per t)
// Re-ordered layer-major: layer l only depends on layer l-1's outputs, so for each layer the
// input projection of ALL T+1 steps is hoisted into one [N,In]x[In,4H] GEMM, the recurrence runs
// as T+1 fused step kernels (recurrent product + gate non-linearities + state update + next
// step's masked state, W_hh slice and h tile staged in shared memory), and BPTT accumulates
// the weight gradients with two [4H,N]x[N,*] GEMMs after the time loop instead of per step.
// torch.nn.LSTM conventions: gate order i,f,g,o; weight_ih [4H,In], weight_hh [4H,H]; two biases.
#include "lstm.cuh"

#include "gemm_simt.cuh"
#include "net_kernels.cuh"

namespace tb {

static inline int padded_h(int H) {
  int hp = (H + 3) & ~3;
  if (((hp / 4) & 1) == 0) hp += 4;  // odd number of 16-byte groups per row: conflict-free LDS.128
  return hp;
}

size_t lstm_ws_bytes(int64_t T1, int64_t B, int In, int H, int layers) {
  return lstm_ws(nullptr, T1, B, In, H, layers).bytes;
}

LstmWs lstm_ws(void* base, int64_t T1, int64_t B, int In, int H, int layers) {
  LstmWs w;
  size_t off = 0;
  auto takef = [&](int64_t n) {
    float* p = base ? reinterpret_cast<float*>(static_cast<char*>(base) + off) : nullptr;
    off += (size_t(n) * sizeof(float) + 255) & ~size_t(255);
    return p;
  };
  const int64_t N = T1 * B;
  for (int l = 0; l < kLstmMaxLayers; ++l) {
    LstmLayerWs& L = w.layer[l];
    if (l < layers) {
      L.gates = takef(N * 4 * H); L.hs = takef(N * H); L.cs = takef(N * H); L.hm = takef(N * H);
      L.cm = takef(N * H); L.dgates = takef(N * 4 * H); L.bsum = takef(4 * H); L.w_hh_t = takef(int64_t(H) * 4 * H);
    } else {
      L = LstmLayerWs();
    }
  }
  w.dh = takef(B * H); w.dc = takef(B * H);
  w.dx_mid = takef(N * (H > In ? H : In));
  w.bytes = off;
  return w;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void add2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + b[i];
}

// ---------------------------------------------------------------------------------------
// forward step: one CTA = 4 hidden units (x 4 gates = 16 rows of W_hh) x a 32-row batch tile.
// 128 threads: lane = batch row, warp q = gate; each thread owns gate q of 4 units.
// ---------------------------------------------------------------------------------------
constexpr int kStepUnits = 4;
constexpr int kStepThreads = 128;

struct StepArgs {
  const float* h_prev; const float* c_prev;  // [B,H] (previous step's h/c or the initial state)
  const float* nd;                           // [B] notdone_t
  const float* w_hh;                         // [4H,H]
  float* gates;                              // [B,4H] in: x-projection + biases; out: activated gates
  float* hs; float* cs; float* hm; float* cm;  // [B,H] this step's outputs
  int B, H, Hp;
};

__global__ void __launch_bounds__(kStepThreads) lstm_step_fwd_kernel(StepArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                                  // [16][Hp]
  float* Xs = smem + 16 * a.Hp;                      // [32][Hp]  masked h_prev tile
  __shared__ float act_s[4][kStepUnits][33];
  const int tid = threadIdx.x, lane = tid & 31, q = tid >> 5;
  const int H = a.H, Hp = a.Hp;
  const int j0 = blockIdx.x * kStepUnits;
  const int b0 = blockIdx.y * 32;
  // stage W_hh rows (g*H + j0 + u) and the masked recurrent input
  for (int idx = tid; idx < 16 * Hp; idx += kStepThreads) {
    const int r = idx / Hp, k = idx % Hp;
    const int g = r >> 2, u = r & 3;
    float v = 0.0f;
    if (k < H && j0 + u < H) v = __ldg(a.w_hh + (int64_t(g) * H + j0 + u) * H + k);
    Ws[idx] = v;
  }
  for (int idx = tid; idx < 32 * Hp; idx += kStepThreads) {
    const int r = idx / Hp, k = idx % Hp;
    float v = 0.0f;
    if (k < H && b0 + r < a.B) v = a.h_prev[int64_t(b0 + r) * H + k] * a.nd[b0 + r];
    Xs[idx] = v;
  }
  __syncthreads();
  float acc[kStepUnits] = {0.f, 0.f, 0.f, 0.f};
  const float4* x4 = reinterpret_cast<const float4*>(Xs + lane * Hp);
  const float4* w4 = reinterpret_cast<const float4*>(Ws + (q * 4) * Hp);
  const int k4n = Hp / 4;
#pragma unroll 2
  for (int k4 = 0; k4 < k4n; ++k4) {
    const float4 x = x4[k4];
#pragma unroll
    for (int u = 0; u < kStepUnits; ++u) {
      const float4 w = w4[u * k4n + k4];
      acc[u] = fmaf(x.x, w.x, acc[u]); acc[u] = fmaf(x.y, w.y, acc[u]);
      acc[u] = fmaf(x.z, w.z, acc[u]); acc[u] = fmaf(x.w, w.w, acc[u]);
    }
  }
  const int b = b0 + lane;
#pragma unroll
  for (int u = 0; u < kStepUnits; ++u) {
    float v = 0.0f;
    if (b < a.B && j0 + u < H) {
      float* gp = a.gates + int64_t(b) * 4 * H + int64_t(q) * H + j0 + u;
      const float pre = *gp + acc[u];
      v = (q == 2) ? tanhf(pre) : sigmoidf_(pre);
      *gp = v;
    }
    act_s[q][u][lane] = v;
  }
  __syncthreads();
  // state update: thread = (batch row lane, unit q)
  const int u = q;
  if (b < a.B && j0 + u < H) {
    const int64_t o = int64_t(b) * H + j0 + u;
    const float nd = a.nd[b];
    const float cmv = a.c_prev[o] * nd;
    const float ig = act_s[0][u][lane], fg = act_s[1][u][lane], gg = act_s[2][u][lane], og = act_s[3][u][lane];
    const float c = fg * cmv + ig * gg;
    a.cs[o] = c;
    a.hs[o] = og * tanhf(c);
    a.cm[o] = cmv;
    a.hm[o] = a.h_prev[o] * nd;
  }
}

// ---------------------------------------------------------------------------------------
// backward step, pointwise part: dgates_t (pre-activation) and the cell-state carry.
//   dh = dy_t + dh_raw * nd_next   (dh_raw = dgates_{t+1} . W_hh, masked by notdone_{t+1})
// ---------------------------------------------------------------------------------------
struct BwdPointArgs {
  const float* dy; const float* dh_raw; const float* nd_next;  // dh_raw/nd_next null at the last step
  const float* nd;                                             // notdone_t
  const float* gates; const float* cs; const float* cm;        // forward saves at t
  float* dc;                                                   // [B,H] carry in/out (dL/dc_t in, dL/dc_{t-1} out)
  float* dgates;                                               // [B,4H]
  int B, H; int first;                                         // first != 0: carries are zero (t = T1-1)
};

__global__ void lstm_step_bwd_point_kernel(BwdPointArgs a) {
  const int64_t total = int64_t(a.B) * a.H;
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = int(i / a.H), j = int(i % a.H);
  const int64_t g0 = int64_t(b) * 4 * a.H + j;
  const float ig = a.gates[g0], fg = a.gates[g0 + a.H], gg = a.gates[g0 + 2 * a.H], og = a.gates[g0 + 3 * a.H];
  float dh = a.dy[i];
  float dc = 0.0f;
  if (!a.first) {
    dh += a.dh_raw[i] * a.nd_next[b];
    dc = a.dc[i];
  }
  const float tc = tanhf(a.cs[i]);
  const float d_o = dh * tc;
  dc += dh * og * (1.0f - tc * tc);
  const float d_i = dc * gg, d_f = dc * a.cm[i], d_g = dc * ig;
  a.dgates[g0] = d_i * ig * (1.0f - ig);
  a.dgates[g0 + a.H] = d_f * fg * (1.0f - fg);
  a.dgates[g0 + 2 * a.H] = d_g * (1.0f - gg * gg);
  a.dgates[g0 + 3 * a.H] = d_o * og * (1.0f - og);
  a.dc[i] = dc * fg * a.nd[b];
}

static int launch_step_fwd(const StepArgs& a, cudaStream_t st) {
  static bool attr_set = false;
  const size_t smem = size_t(48) * a.Hp * sizeof(float);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(lstm_step_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    TB_REQUIRE(e == cudaSuccess, "lstm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  TB_REQUIRE(smem <= 200 * 1024, "lstm: hidden size %d too large for the step kernel", a.H);
  dim3 grid((a.H + kStepUnits - 1) / kStepUnits, (a.B + 31) / 32);
  lstm_step_fwd_kernel<<<grid, kStepThreads, smem, st>>>(a);
  return check_launch("lstm_step_fwd_kernel");
}

#define TB_TRY(expr)        \
  do {                      \
    int _rc = (expr);       \
    if (_rc) return _rc;    \
  } while (0)

int lstm_forward(const float* x, const float* notdone, const float* h0, const float* c0, const LstmParams& p,
                 int64_t T1, int64_t B, int In, int H, int layers, LstmWs& ws, float* y, float* hN, float* cN,
                 float* splitk, cudaStream_t st) {
  TB_REQUIRE(layers >= 1 && layers <= kLstmMaxLayers, "lstm: 1..%d layers", kLstmMaxLayers);
  const int64_t N = T1 * B;
  const float* xin = x;
  int in_dim = In;
  for (int l = 0; l < layers; ++l) {
    LstmLayerWs& L = ws.layer[l];
    float* hs = (l == layers - 1) ? y : L.hs;
    add2_kernel<<<(4 * H + 255) / 256, 256, 0, st>>>(p.b_ih[l], p.b_hh[l], L.bsum, 4 * H);
    TB_TRY(check_launch("add2_kernel"));
    GemmEpilogue ep; ep.bias = L.bsum; ep.tag = "lstm_xproj_fwd";
    TB_TRY((gemm_simt<float, float, false, true>(xin, p.w_ih[l], L.gates, N, 4 * H, in_dim, in_dim, in_dim, 4 * H, ep, 1,
                                                  nullptr, st)));
    ProfScope prof("lstm_recurrence_fwd", st);
    for (int64_t t = 0; t < T1; ++t) {
      StepArgs a;
      a.h_prev = (t == 0) ? h0 + int64_t(l) * B * H : hs + (t - 1) * B * H;
      a.c_prev = (t == 0) ? c0 + int64_t(l) * B * H : L.cs + (t - 1) * B * H;
      a.nd = notdone + t * B;
      a.w_hh = p.w_hh[l];
      a.gates = L.gates + t * B * 4 * H;
      a.hs = hs + t * B * H; a.cs = L.cs + t * B * H; a.hm = L.hm + t * B * H; a.cm = L.cm + t * B * H;
      a.B = int(B); a.H = H; a.Hp = padded_h(H);
      TB_TRY(launch_step_fwd(a, st));
    }
    cudaError_t e = cudaMemcpyAsync(hN + int64_t(l) * B * H, hs + (T1 - 1) * B * H, sizeof(float) * B * H,
                                    cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(cN + int64_t(l) * B * H, L.cs + (T1 - 1) * B * H, sizeof(float) * B * H,
                          cudaMemcpyDeviceToDevice, st);
    TB_REQUIRE(e == cudaSuccess, "lstm: state copy: %s", cudaGetErrorString(e));
    xin = hs;
    in_dim = H;
  }
  (void)splitk;
  return 0;
}

static int splits_for(int64_t M, int64_t N, int64_t K, int64_t scratch_floats) {
  const int64_t bm = (N <= 32) ? 128 : (M <= 64 ? 64 : 128), bn = (N <= 32) ? 32 : 64;
  const int64_t tiles = ((M + bm - 1) / bm) * ((N + bn - 1) / bn);
  int64_t s = (2 * kNumSMsB200 + tiles - 1) / tiles;
  const int64_t ktiles = (K + kGemmBK - 1) / kGemmBK;
  if (s > ktiles / 4) s = ktiles / 4;
  if (s * M * N > scratch_floats) s = scratch_floats / (M * N);
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return int(s);
}

int lstm_backward(const float* dy, const float* x, const float* notdone, const LstmParams& p, const LstmGrads& g,
                  int64_t T1, int64_t B, int In, int H, int layers, LstmWs& ws, float* dx, float* splitk,
                  float* colsum_scratch, cudaStream_t st) {
  const int64_t N = T1 * B;
  const int64_t scratch = int64_t(8) << 20;  // == kSplitKScratchFloats (atarinet.cu)
  const float* dyl = dy;
  for (int l = layers - 1; l >= 0; --l) {
    LstmLayerWs& L = ws.layer[l];
    const float* xin = (l == 0) ? x : ws.layer[l - 1].hs;
    const int in_dim = (l == 0) ? In : H;
    float* dxl = (l == 0) ? dx : ws.dx_mid;
    {
    ProfScope prof("lstm_recurrence_bwd", st);
    for (int64_t t = T1 - 1; t >= 0; --t) {
      BwdPointArgs a;
      a.dy = dyl + t * B * H;
      a.first = (t == T1 - 1);
      a.dh_raw = ws.dh; a.nd_next = a.first ? nullptr : notdone + (t + 1) * B;
      a.nd = notdone + t * B;
      a.gates = L.gates + t * B * 4 * H; a.cs = L.cs + t * B * H; a.cm = L.cm + t * B * H;
      a.dc = ws.dc; a.dgates = L.dgates + t * B * 4 * H; a.B = int(B); a.H = H;
      const int64_t total = B * H;
      lstm_step_bwd_point_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
      TB_TRY(check_launch("lstm_step_bwd_point_kernel"));
      if (t > 0) {
        // dh_raw[B,H] = dgates_t[B,4H] . W_hh[4H,H]   (masked by notdone_t when consumed at t-1)
        GemmEpilogue ep; ep.tag = "lstm_step_dh";
        const int s = splits_for(B, H, 4 * H, scratch);
        TB_TRY((gemm_simt<float, float, false, false>(a.dgates, p.w_hh[l], ws.dh, B, H, 4 * H, 4 * H, H, H, ep, s,
                                                       splitk, st)));
      }
    }
    }
    // parameter gradients over all steps at once
    GemmEpilogue ep; ep.tag = "lstm_wgrad";
    int s = splits_for(4 * H, H, N, scratch);
    TB_TRY((gemm_simt<float, float, true, false>(L.dgates, L.hm, g.w_hh[l], 4 * H, H, N, 4 * H, H, H, ep, s, splitk, st)));
    s = splits_for(4 * H, in_dim, N, scratch);
    TB_TRY((gemm_simt<float, float, true, false>(L.dgates, xin, g.w_ih[l], 4 * H, in_dim, N, 4 * H, in_dim, in_dim, ep, s,
                                                  splitk, st)));
    TB_TRY(colsum(L.dgates, g.b_ih[l], N, 4 * H, 4 * H, colsum_scratch, st));
    cudaError_t e = cudaMemcpyAsync(g.b_hh[l], g.b_ih[l], sizeof(float) * 4 * H, cudaMemcpyDeviceToDevice, st);
    TB_REQUIRE(e == cudaSuccess, "lstm: bias grad copy: %s", cudaGetErrorString(e));
    // gradient w.r.t. this layer's input: dx[N,in] = dgates[N,4H] . W_ih[4H,in]
    ep.tag = "lstm_xproj_dgrad";
    TB_TRY((gemm_simt<float, float, false, false>(L.dgates, p.w_ih[l], dxl, N, in_dim, 4 * H, 4 * H, in_dim, in_dim, ep, 1,
                                                   nullptr, st)));
    dyl = dxl;
  }
  return 0;
}

}  // namespace tb
