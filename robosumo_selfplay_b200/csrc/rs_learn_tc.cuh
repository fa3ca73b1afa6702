// rs_learn_tc.cuh -- tensor-core (tcgen05, kind::tf32, TMEM accumulators) versions of the batched MLP kernels of
// rs_learn.cuh: policy/value inference and the PPO2 minibatch forward/backward.
//
// Per 128-sample tile and per net, the three GEMMs whose operands are already K-contiguous run on the 5th-gen tensor cores:
//     F1  H1pre[128,64] = X[128,K0] * W0[K0,64]      (K0 = D padded to a multiple of 8, zero filled)
//     F2  H2pre[128,64] = H1[128,64] * W1[64,64]
//     B2  dH1  [128,64] = dZ2[128,64] * W1^T
// one elected thread issues tcgen05.mma, the accumulator lives in TMEM and comes back with tcgen05.ld as "thread t = sample
// row t" for the bias / ReLU / mask epilogues.  The weight-gradient products (in^T * dz, K = the sample index) stay on the
// FP32 pipe: their operands would need MN-major (transposed) descriptors, which this build does not use.
// Operand tiles use the no-swizzle core-matrix layout of rs_tc.cuh.  Numerics: tf32 inputs (10-bit mantissa), fp32 accumulate.
#pragma once
#include "rs_learn.cuh"
#include "rs_tc.cuh"

namespace rsl {

using rstc::tile_off;

struct TcTile {
    float *xs, *h1, *h2, *w0t, *w1t, *w1n, *b0, *b1, *wh, *bh, *dout;
    int K0;
};
__host__ __device__ inline int tc_k0(int D) { return (D + 7) & ~7; }
__host__ __device__ inline size_t tc_tile_bytes(int D) {
    const int K0 = tc_k0(D);
    return sizeof(float) * ((size_t)RSL_TILE * K0 + 2 * RSL_TILE * RSL_H + (size_t)RSL_H * K0 + 2 * RSL_H * RSL_H + 2 * RSL_H + RSL_H * RSL_HW + RSL_HW + RSL_TILE * RSL_DS) + 1024;
}
__device__ inline TcTile tc_carve(float* base, int D) {
    TcTile t; t.K0 = tc_k0(D);
    t.xs = base; base += RSL_TILE * t.K0;
    t.h1 = base; base += RSL_TILE * RSL_H;
    t.h2 = base; base += RSL_TILE * RSL_H;
    t.w0t = base; base += RSL_H * t.K0;
    t.w1t = base; base += RSL_H * RSL_H;
    t.w1n = base; base += RSL_H * RSL_H;
    t.b0 = base; base += RSL_H; t.b1 = base; base += RSL_H;
    t.wh = base; base += RSL_H * RSL_HW; t.bh = base; base += RSL_HW;
    t.dout = base;
    return t;
}
// weights of one net into operand tiles: W0^T [64][K0] and W1^T [64][64] (B operands of F1 / F2), W1 [64][64] (B operand of B2)
__device__ inline void tc_stage_net(const TcTile& t, const float* __restrict__ p, int D, int w0, int b0, int w1, int b1, int wh, int bh, int out) {
    // global loads are issued four float4 deep before the dependent shared-memory stores: the staging is latency-bound on L2
    // (64 dependent load->store round trips per thread were 30 % of the tile kernel)
    const int nv0 = t.K0 * RSL_H / 4;                                  // W0 is [k][64]: 16 float4 per input row k
    for (int i0 = threadIdx.x; i0 < nv0; i0 += 4 * blockDim.x) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * blockDim.x, k = i >> 4, n4 = (i & 15) * 4;
            v[u] = (i < nv0 && k < D) ? __ldg(reinterpret_cast<const float4*>(p + w0 + k * RSL_H + n4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * blockDim.x, k = i >> 4, n4 = (i & 15) * 4;
            if (i < nv0) {
                t.w0t[tile_off(n4, k, t.K0)] = v[u].x; t.w0t[tile_off(n4 + 1, k, t.K0)] = v[u].y;
                t.w0t[tile_off(n4 + 2, k, t.K0)] = v[u].z; t.w0t[tile_off(n4 + 3, k, t.K0)] = v[u].w;
            }
        }
    }
    const int nv1 = RSL_H * RSL_H / 4;
    for (int i0 = threadIdx.x; i0 < nv1; i0 += 4 * blockDim.x) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int i = i0 + u * blockDim.x; v[u] = i < nv1 ? __ldg(reinterpret_cast<const float4*>(p + w1 + 4 * i)) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * blockDim.x, k = i >> 4, n4 = (i & 15) * 4;
            if (i < nv1) {
                const float w[4] = { v[u].x, v[u].y, v[u].z, v[u].w };
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    t.w1t[tile_off(n4 + j, k, RSL_H)] = w[j];          // [out j][in i]
                    t.w1n[tile_off(k, n4 + j, RSL_H)] = w[j];          // [in i][out j]
                }
            }
        }
    }
    for (int i = threadIdx.x; i < RSL_H; i += blockDim.x) { t.b0[i] = p[b0 + i]; t.b1[i] = p[b1 + i]; }
    for (int i = threadIdx.x; i < RSL_H * RSL_HW; i += blockDim.x) { int r = i / RSL_HW, c = i % RSL_HW; t.wh[i] = c < out ? p[wh + r * out + c] : 0.f; }
    if (threadIdx.x < RSL_HW) t.bh[threadIdx.x] = threadIdx.x < out ? p[bh + threadIdx.x] : 0.f;
}
__device__ inline void tc_stage_x(const TcTile& t, const float* __restrict__ X, size_t ldx, const int* __restrict__ idx, int row0, int n, int D) {
    // warp w gathers rows w, w + nw, ...; the row indices of the whole warp are fetched with one load and four rows (16 loads per
    // lane) are in flight before the first shared-memory store
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int my_row = warp + nw * lane, my_g = row0 + my_row;
    long long my_src = -1;
    if (my_row < RSL_TILE && my_g < n) my_src = (long long)(idx ? idx[my_g] : my_g) * (long long)ldx;
    for (int l0 = 0; l0 * nw + warp < RSL_TILE; l0 += 4) {
        float v[4][4];
        long long off[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            off[u] = __shfl_sync(0xffffffffu, my_src, (l0 + u) & 31);
            const bool ok = (l0 + u) * nw + warp < RSL_TILE && off[u] >= 0;
#pragma unroll
            for (int j = 0; j < 4; j++) { const int k = lane + 32 * j; v[u][j] = (ok && k < D) ? __ldg(X + off[u] + k) : 0.f; }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int r = (l0 + u) * nw + warp;
            if (r < RSL_TILE) {
#pragma unroll
                for (int j = 0; j < 4; j++) { const int k = lane + 32 * j; if (k < t.K0) t.xs[tile_off(r, k, t.K0)] = v[u][j]; }
            }
        }
    }
}
// The tensor-core kernels run RSL_TC_THREADS = 256 threads on a 128-row tile: thread t works on row t % 128 and on the column
// half t / 128 (accumulator columns 32 ch .. 32 ch + 31) -- eight warps instead of four to hide the shared-memory, TMEM and
// L2 latencies of the epilogues, transposes and staging loops; scalar per-row work (loss, outputs) is done by the ch == 0 threads.
#define RSL_TC_THREADS 256
struct TcCtx { uint32_t tmem; uint64_t* bar; uint32_t phase; };

// D[tmem] = A[128 x K] * B^T with B given as [64 x K] tile (both K-major); called by all threads, issued by thread 0
__device__ __forceinline__ void tc_gemm(TcCtx& c, const float* A, const float* Bt, int K) {
    rstc::fence_async_smem();
    rstc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        rstc::tc_fence_after();
        const uint32_t sbo = (K / 4) * 128;
        rstc::umma_tf32(c.tmem, rstc::smem_u32(A), 128, sbo, 256, rstc::smem_u32(Bt), 128, sbo, 256, rstc::make_idesc(128, 64, 0, 0), K / 8, c.bar);
    }
    rstc::mbar_wait(c.bar, c.phase);
    c.phase ^= 1u;
    rstc::tc_fence_after();
}
// epilogue of F1 / F2: row r = threadIdx.x of the accumulator -> relu(acc + bias) into an operand tile [128][64]
__device__ __forceinline__ void tc_relu_to_tile(const TcCtx& c, const float* bias, float* tile) {
    const int r = threadIdx.x & 127, c0 = (threadIdx.x >> 7) * 32;
    float o[32];
    rstc::tmem_ld32(c.tmem, c0, o);
#pragma unroll
    for (int q = 0; q < 32; q += 4) {
        float4 v = make_float4(fmaxf(o[q] + bias[c0 + q], 0.f), fmaxf(o[q + 1] + bias[c0 + q + 1], 0.f),
                               fmaxf(o[q + 2] + bias[c0 + q + 2], 0.f), fmaxf(o[q + 3] + bias[c0 + q + 3], 0.f));
        *reinterpret_cast<float4*>(tile + tile_off(r, c0 + q, RSL_H)) = v;
    }
}
// forward of the staged net: H1, H2 tiles filled; head outputs of this thread's row in out8
__device__ __forceinline__ void tc_net_forward(TcCtx& c, const TcTile& t, float* outh) {
    tc_gemm(c, t.xs, t.w0t, t.K0);
    tc_relu_to_tile(c, t.b0, t.h1);
    tc_gemm(c, t.h1, t.w1t, RSL_H);
    tc_relu_to_tile(c, t.b1, t.h2);
    // heads: each thread sums its own column half of H2 (its own writes: no barrier needed); ch == 1 hands its partial to ch == 0
    const int r = threadIdx.x & 127, ch = threadIdx.x >> 7;
#pragma unroll
    for (int q = 0; q < RSL_HW; q++) outh[q] = ch ? 0.f : t.bh[q];
    for (int k = 32 * ch; k < 32 * ch + 32; k += 4) {
        const float4 a = *reinterpret_cast<const float4*>(t.h2 + tile_off(r, k, RSL_H));
        const float av[4] = { a.x, a.y, a.z, a.w };
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int q = 0; q < RSL_HW; q++) outh[q] = fmaf(av[u], t.wh[(k + u) * RSL_HW + q], outh[q]);
    }
    if (ch) {
#pragma unroll
        for (int q = 0; q < RSL_HW; q++) t.dout[r * RSL_DS + q] = outh[q];
    }
    __syncthreads();
    if (!ch) {
#pragma unroll
        for (int q = 0; q < RSL_HW; q++) outh[q] += t.dout[r * RSL_DS + q];
    }
    __syncthreads();                                    // dout is reused by the caller
}
__device__ __forceinline__ void tc_begin(TcCtx& c, uint64_t* bar, uint32_t* slot) {
    if (threadIdx.x == 0) rstc::mbar_init(bar, 1);
    if (threadIdx.x < 32) rstc::tmem_alloc(slot, 64);
    rstc::tc_fence_before();
    __syncthreads();
    rstc::tc_fence_after();
    c.tmem = *slot; c.bar = bar; c.phase = 0;
}
__device__ __forceinline__ void tc_end(TcCtx& c) {
    rstc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) rstc::tmem_dealloc(c.tmem, 64);
}

// ---- inference (same job contract as k_mlp_forward) ----
__global__ void __launch_bounds__(RSL_TC_THREADS) k_mlp_forward_tc(MlpJobs J, int D, int A, int n) {
    extern __shared__ __align__(16) float smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const Layout L = make_layout(D, A);
    TcTile t = tc_carve(smem, D);
    TcCtx c;
    tc_begin(c, &bar, &slot);
    const int job = blockIdx.y;
    const float* __restrict__ params = J.params[job];
    float* __restrict__ mean = J.mean[job];
    float* __restrict__ value = J.value[job];
    const int row0 = blockIdx.x * RSL_TILE, g = row0 + (threadIdx.x & 127);
    const bool writer = threadIdx.x < 128 && g < n;
    tc_stage_x(t, J.X[job], J.ldx[job], nullptr, row0, n, D);
    float o[RSL_HW];
    if (mean) {
        tc_stage_net(t, params, D, L.pi_w0, L.pi_b0, L.pi_w1, L.pi_b1, L.pi_w, L.pi_b, A);
        tc_net_forward(c, t, o);
        if (writer) for (int q = 0; q < A; q++) mean[(size_t)g * A + q] = o[q];
        rstc::tc_fence_before();
        __syncthreads();
    }
    if (value) {
        tc_stage_net(t, params, D, L.vf_w0, L.vf_b0, L.vf_w1, L.vf_b1, L.vf_w, L.vf_b, 1);
        tc_net_forward(c, t, o);
        if (writer) value[g] = o[0];
    }
    tc_end(c);
}

// ---- gradients on the FP32 pipe, operands in core-matrix tiles: out[M x 64] = in^T dz over the tile rows ----
__device__ inline void tc_grad_weight(const float* __restrict__ in, int C, int M, const float* __restrict__ dz, float* __restrict__ out, int rows) {
    const int i0 = (threadIdx.x >> 3) * 8, j0 = (threadIdx.x & 7) * 8;
    if (i0 >= M) return;
    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) acc[a][b] = 0.f;
    for (int r = 0; r < rows; r++) {
        const float4 a0 = *reinterpret_cast<const float4*>(in + tile_off(r, i0, C)), a1 = *reinterpret_cast<const float4*>(in + tile_off(r, i0 + 4, C));
        const float4 b0 = *reinterpret_cast<const float4*>(dz + tile_off(r, j0, RSL_H)), b1 = *reinterpret_cast<const float4*>(dz + tile_off(r, j0 + 4, RSL_H));
        const float av[8] = { a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w }, bv[8] = { b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w };
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 8; b++) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 8; a++) if (i0 + a < M)
#pragma unroll
        for (int b = 0; b < 8; b++) out[(i0 + a) * RSL_H + j0 + b] = acc[a][b];
}
__device__ inline void tc_grad_bias(const float* __restrict__ dz, float* __restrict__ out, int rows) {
    if (threadIdx.x < RSL_H) {
        float s = 0.f;
        for (int r = 0; r < rows; r++) s += dz[tile_off(r, threadIdx.x, RSL_H)];
        out[threadIdx.x] = s;
    }
}
// ---- weight gradients on the tensor core:  dW[f][j] = sum_s in[s][f] * dz[s][j]  (K = sample index) -------------------------
// kind::tf32 operands must be K-major (an MN-major descriptor returns zeros, rs_tc_selftest), so both operands are copied
// TRANSPOSED into scratch tiles -- inT [feature][sample] over the dead W0^T tile, dzT [unit][sample] over the dead W1^T tile --
// 64 samples at a time (two K-halves accumulated in TMEM).  M = 128 feature rows are always issued; rows beyond the valid
// ones read stale but finite shared memory and only produce accumulator rows nobody reads.  One extra row of ones makes its
// accumulator row the bias gradient (column sums of dz).  Requires K0 == 128 (scratch extent) and a spare row (`ones_row`).
__device__ __forceinline__ void tc_gemm_acc(TcCtx& c, const float* A, const float* Bt, int K, bool accumulate) {
    rstc::fence_async_smem();
    rstc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        rstc::tc_fence_after();
        const uint32_t sbo = (K / 4) * 128;
        rstc::umma_tf32_acc(c.tmem, rstc::smem_u32(A), 128, sbo, 256, rstc::smem_u32(Bt), 128, sbo, 256, rstc::make_idesc(128, 64, 0, 0), K / 8, accumulate, c.bar);
    }
    rstc::mbar_wait(c.bar, c.phase);
    c.phase ^= 1u;
    rstc::tc_fence_after();
}
__device__ inline void tc_wgrad(TcCtx& c, const TcTile& t, const float* __restrict__ in, int Cin, int nf, int nf_valid, int ones_row,
                                const float* __restrict__ dz, float* __restrict__ gw, float* __restrict__ gb) {
    float* inT = t.w0t;
    float* dzT = t.w0t + 128 * 64;
    // transposing copies: a warp moves an 8-feature x 4-sample patch per instruction, lane = 4 * (feature % 8) + (sample % 4), so that
    // the 32 destination words are consecutive (conflict-free stores; the loads are 2-way conflicted at worst)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int fi = lane >> 2, si = lane & 3;
    for (int half = 0; half < 2; half++) {
        for (int p = warp; p < (nf >> 3) * 16; p += nw) {           // patches: feature group fg (8 features) x sample group sg (4 samples)
            const int fg = p >> 4, sg = p & 15, f = fg * 8 + fi;
            const float v = in[tile_off(half * 64 + sg * 4 + si, f, Cin)];
            inT[fg * 512 + sg * 32 + lane] = f == ones_row ? 1.f : v;
        }
        if (ones_row >= nf) { for (int sl = threadIdx.x; sl < 64; sl += blockDim.x) inT[tile_off(ones_row, sl, 64)] = 1.f; }
        for (int p = warp; p < (RSL_H >> 3) * 16; p += nw) {
            const int fg = p >> 4, sg = p & 15;
            dzT[fg * 512 + sg * 32 + lane] = dz[tile_off(half * 64 + sg * 4 + si, fg * 8 + fi, RSL_H)];
        }
        tc_gemm_acc(c, inT, dzT, 64, half > 0);
    }
    // read the accumulator out through shared memory (the transposed tiles are dead now) so that the global writes are coalesced:
    // a thread owns an accumulator ROW, and row-per-thread stores to the row-major partial hit 32 different 256-byte rows at once
    const int r = threadIdx.x & 127, c0 = (threadIdx.x >> 7) * 32;
    float o[32];
    rstc::tmem_ld32(c.tmem, c0, o);
    float* stg = t.w0t;                                   // [128][65] floats: spans the W0^T and W1^T operand tiles (48 KB)
#pragma unroll
    for (int q = 0; q < 32; q++) stg[r * 65 + c0 + q] = o[q];
    __syncthreads();
    for (int i = threadIdx.x; i < nf_valid * RSL_H; i += blockDim.x) gw[i] = stg[(i >> 6) * 65 + (i & 63)];
    if (threadIdx.x < RSL_H) gb[threadIdx.x] = stg[ones_row * 65 + threadIdx.x];
    __syncthreads();
}

// dZ1 = relu'(H1) * (dZ2 * W1^T): GEMM B2 on the tensor core, mask epilogue, written in place of H1
__device__ __forceinline__ void tc_backprop_hidden(TcCtx& c, const TcTile& t) {
    tc_gemm(c, t.h2, t.w1n, RSL_H);                     // A = dZ2 (in the h2 tile), B rows = input unit i, K = output unit j
    const int r = threadIdx.x & 127, c0 = (threadIdx.x >> 7) * 32;
    float o[32];
    rstc::tmem_ld32(c.tmem, c0, o);
#pragma unroll
    for (int q = 0; q < 32; q += 4) {
        float4* p = reinterpret_cast<float4*>(t.h1 + tile_off(r, c0 + q, RSL_H));
        float4 h = *p;
        *p = make_float4(h.x > 0.f ? o[q] : 0.f, h.y > 0.f ? o[q + 1] : 0.f, h.z > 0.f ? o[q + 2] : 0.f, h.w > 0.f ? o[q + 3] : 0.f);
    }
}

// one block = 128 samples: same contract as k_ppo_tile (per-block partial gradients and stat partials)
__global__ void __launch_bounds__(RSL_TC_THREADS) k_ppo_tile_tc(PPOArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const Layout L = make_layout(a.D, a.A);
    const int D = a.D, A = a.A;
    TcTile t = tc_carve(smem, D);
    TcCtx c;
    tc_begin(c, &bar, &slot);
    const int row0 = blockIdx.x * RSL_TILE, r = threadIdx.x & 127, ch = threadIdx.x >> 7, g = row0 + r;
    const int rows = min(RSL_TILE, a.n - row0);
    const bool live = ch == 0 && g < a.n;          // the per-sample scalar work (loss, statistics) belongs to the ch == 0 thread of a row
    const int s = live ? (a.idx ? a.idx[g] : g) : 0;
    float* gp = a.gpart + (size_t)blockIdx.x * L.P;
    // (no zero fill of the partial: every one of its P entries is written exactly once below)
    tc_stage_x(t, a.obs, (size_t)D, a.idx, row0, a.n, D);
    // ---------------- policy net ----------------
    tc_stage_net(t, a.params, D, L.pi_w0, L.pi_b0, L.pi_w1, L.pi_b1, L.pi_w, L.pi_b, A);
    float mu[RSL_HW];
    tc_net_forward(c, t, mu);
    const double mean = a.adv_sums[0] / a.adv_count;
    const double var = fmax(a.adv_sums[1] / a.adv_count - mean * mean, 0.0);
    const float adv_mean = (float)mean, adv_std = (float)sqrt(var);
    float st_pg = 0.f, st_kl = 0.f, st_clip = 0.f, st_vf = 0.f;
    float dmu[RSL_HW], dls[RSL_HW];
#pragma unroll
    for (int q = 0; q < RSL_HW; q++) { dmu[q] = 0.f; dls[q] = 0.f; }
    if (live) {
        float ls[RSL_HW], z[RSL_HW], nl = 0.f, lsum = 0.f;
        for (int q = 0; q < A; q++) {
            ls[q] = a.params[L.logstd + q];
            z[q] = (a.actions[(size_t)s * A + q] - mu[q]) * expf(-ls[q]);
            nl += z[q] * z[q]; lsum += ls[q];
        }
        nl = 0.5f * nl + 0.9189385332046727f * (float)A + lsum;
        const float old = a.old_nlp[s];
        float ratio = expf(old - nl);
        const bool isnan_ = ratio != ratio;
        if (isnan_) ratio = 2.f;
        const float adv = (a.returns[s] - a.values[s] - adv_mean) / (adv_std + 1e-8f);
        const float w = a.weights ? a.weights[s] : 1.f;
        const float rc = fminf(fmaxf(ratio, 1.f - a.cliprange), 1.f + a.cliprange);
        const float pg1 = -adv * ratio, pg2 = -adv * rc;
        st_pg = w * fmaxf(pg1, pg2);
        st_kl = nl - old;
        st_clip = fabsf(ratio - 1.f) > a.cliprange ? 1.f : 0.f;
        if (a.log_ratio) a.log_ratio[g] = old - nl;
        float gnl = 0.f;
        if (!isnan_) {
            const bool inside = (ratio >= 1.f - a.cliprange) && (ratio <= 1.f + a.cliprange);
            if (pg1 >= pg2) gnl = adv * ratio; else if (inside) gnl = adv * ratio;
        }
        gnl *= w * a.inv_n;
        for (int q = 0; q < A; q++) { dmu[q] = gnl * (-z[q] * expf(-ls[q])); dls[q] = gnl * (1.f - z[q] * z[q]); }
    }
    if (!ch) for (int q = 0; q < RSL_HW; q++) t.dout[r * RSL_DS + q] = dmu[q];
    __syncthreads();
    if (ch) for (int q = 0; q < RSL_HW; q++) dmu[q] = t.dout[r * RSL_DS + q];      // the other column half of the row needs dmu for dZ2
    for (int o = threadIdx.x; o < RSL_H * A; o += blockDim.x) {
        const int k = o / A, q = o - k * A;
        float acc = 0.f;
        for (int u = 0; u < rows; u++) acc = fmaf(t.h2[tile_off(u, k, RSL_H)], t.dout[u * RSL_DS + q], acc);
        gp[L.pi_w + o] = acc;
    }
    if (threadIdx.x < A) { float acc = 0.f; for (int u = 0; u < rows; u++) acc += t.dout[u * RSL_DS + threadIdx.x]; gp[L.pi_b + threadIdx.x] = acc; }
    __syncthreads();
    if (!ch) for (int q = 0; q < RSL_HW; q++) t.dout[r * RSL_DS + q] = dls[q];
    __syncthreads();
    if (threadIdx.x < A) { float acc = 0.f; for (int u = 0; u < rows; u++) acc += t.dout[u * RSL_DS + threadIdx.x]; gp[L.logstd + threadIdx.x] = acc; }
    // dZ2 = relu'(H2) * (dmu Wp^T), in place of H2 (own row)
    for (int k = 32 * ch; k < 32 * ch + 32; k++) {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < RSL_HW; q++) acc = fmaf(dmu[q], t.wh[k * RSL_HW + q], acc);
        float* p = t.h2 + tile_off(r, k, RSL_H);
        *p = *p > 0.f ? acc : 0.f;
    }
    __syncthreads();
    const bool tcw = (t.K0 == 128) && (D < 128);       // tensor-core weight gradients need the 128-wide scratch and a spare feature row
    if (tcw) tc_wgrad(c, t, t.h1, RSL_H, RSL_H, RSL_H, RSL_H, t.h2, gp + L.pi_w1, gp + L.pi_b1);
    else { tc_grad_weight(t.h1, RSL_H, RSL_H, t.h2, gp + L.pi_w1, rows); tc_grad_bias(t.h2, gp + L.pi_b1, rows); }
    tc_backprop_hidden(c, t);                                              // dZ1 in place of H1 (tensor core)
    __syncthreads();
    if (tcw) tc_wgrad(c, t, t.xs, t.K0, t.K0, D, D, t.h1, gp + L.pi_w0, gp + L.pi_b0);
    else { tc_grad_weight(t.xs, t.K0, D, t.h1, gp + L.pi_w0, rows); tc_grad_bias(t.h1, gp + L.pi_b0, rows); }
    rstc::tc_fence_before();
    __syncthreads();
    // ---------------- value net ----------------
    tc_stage_net(t, a.params, D, L.vf_w0, L.vf_b0, L.vf_w1, L.vf_b1, L.vf_w, L.vf_b, 1);
    float vo[RSL_HW];
    tc_net_forward(c, t, vo);
    float dv = 0.f;
    if (live) {
        const float err = vo[0] - a.returns[s];
        st_vf = 0.5f * err * err;
        dv = a.vf_coef * err * a.inv_n;
    }
    if (!ch) t.dout[r * RSL_DS] = dv;
    __syncthreads();
    if (ch) dv = t.dout[r * RSL_DS];
    if (threadIdx.x < RSL_H) { float acc = 0.f; for (int u = 0; u < rows; u++) acc = fmaf(t.h2[tile_off(u, threadIdx.x, RSL_H)], t.dout[u * RSL_DS], acc); gp[L.vf_w + threadIdx.x] = acc; }
    if (threadIdx.x == 64) { float acc = 0.f; for (int u = 0; u < rows; u++) acc += t.dout[u * RSL_DS]; gp[L.vf_b] = acc; }
    __syncthreads();
    for (int k = 32 * ch; k < 32 * ch + 32; k++) { float* p = t.h2 + tile_off(r, k, RSL_H); *p = *p > 0.f ? dv * t.wh[k * RSL_HW] : 0.f; }
    __syncthreads();
    if (tcw) tc_wgrad(c, t, t.h1, RSL_H, RSL_H, RSL_H, RSL_H, t.h2, gp + L.vf_w1, gp + L.vf_b1);
    else { tc_grad_weight(t.h1, RSL_H, RSL_H, t.h2, gp + L.vf_w1, rows); tc_grad_bias(t.h2, gp + L.vf_b1, rows); }
    tc_backprop_hidden(c, t);
    __syncthreads();
    if (tcw) tc_wgrad(c, t, t.xs, t.K0, t.K0, D, D, t.h1, gp + L.vf_w0, gp + L.vf_b0);
    else { tc_grad_weight(t.xs, t.K0, D, t.h1, gp + L.vf_w0, rows); tc_grad_bias(t.h1, gp + L.vf_b0, rows); }
    // ---------------- stat partials ----------------
    float v4[4] = { st_pg, st_vf, st_kl, st_clip };
    __syncthreads();
    for (int k = 0; k < 4; k++) {
        float x = v4[k];
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && threadIdx.x < 128) t.dout[k * 4 + (threadIdx.x >> 5)] = x;
    }
    __syncthreads();
    if (threadIdx.x < 4) a.spart[(size_t)blockIdx.x * 8 + threadIdx.x] = t.dout[threadIdx.x * 4] + t.dout[threadIdx.x * 4 + 1] + t.dout[threadIdx.x * 4 + 2] + t.dout[threadIdx.x * 4 + 3];
    tc_end(c);
}

}  // namespace rsl
