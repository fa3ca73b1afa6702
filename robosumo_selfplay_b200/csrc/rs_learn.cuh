// rs_learn.cuh -- learner-side kernels: policy/value MLP inference for the rollout, Gaussian
// sampling + neglogp, V-trace backward scan, and the PPO2 clipped-loss minibatch step
// (forward + backward of both MLPs, deterministic gradient reduction, global-norm clip,
// TF-style Adam).
//
// Replaces the TF1 graph of the reference:
//   network      baselines/baselines/common/models.py:93-101 (mlp, 2 x 64, activation on every layer)
//   fc           baselines/baselines/a2c/utils.py:58-63
//   heads        policies.py:50,70-71 ; DiagGaussianPd baselines/baselines/common/distributions.py:227-251
//   step/value   policies.py:84-128
//   loss/update  model.py:51-139,179-213
//   V-trace      runner.py:166-200
//
// Flat parameter layout = order of tf.trainable_variables(scope) of the reference (verified on
// /root/reference/model.ckpt): pi.fc0.w[D,64] b[64] pi.fc1.w[64,64] b[64] vf.fc0.w[D,64] b[64]
// vf.fc1.w[64,64] b[64] pi.w[64,A] pi.b[A] logstd[1,A] vf.w[64,1] vf.b[1]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rsl {

#define RSL_H 64           // hidden width (defaults.py:24 num_hidden=64)
#define RSL_TILE 128       // samples per block
#define RSL_HP 65          // padded row stride of hidden tiles (conflict-free column access)
#define RSL_HW 16          // head width the kernels pad to (action dim <= 16: ant 8, bug 12, spider 16; value head uses column 0)
#define RSL_DS (RSL_HW + 1)

struct Layout {
    int D, A, P;
    int pi_w0, pi_b0, pi_w1, pi_b1, vf_w0, vf_b0, vf_w1, vf_b1, pi_w, pi_b, logstd, vf_w, vf_b;
};
__host__ __device__ inline Layout make_layout(int D, int A) {
    Layout L; L.D = D; L.A = A;
    int o = 0;
    L.pi_w0 = o; o += D * RSL_H; L.pi_b0 = o; o += RSL_H; L.pi_w1 = o; o += RSL_H * RSL_H; L.pi_b1 = o; o += RSL_H;
    L.vf_w0 = o; o += D * RSL_H; L.vf_b0 = o; o += RSL_H; L.vf_w1 = o; o += RSL_H * RSL_H; L.vf_b1 = o; o += RSL_H;
    L.pi_w = o; o += RSL_H * A; L.pi_b = o; o += A; L.logstd = o; o += A; L.vf_w = o; o += RSL_H; L.vf_b = o; o += 1;
    L.P = o;
    return L;
}

// ---- shared-memory plan of one 128-sample tile -------------------------------------------------
//   xs  [128][Dp]   input rows (Dp = D rounded up to a multiple of 4, +1 if that is a multiple of 32)
//   h1  [128][65], h2 [128][65]
//   w0  [D][64], w1 [64][64], wh [64][OUTp], biases
struct Tile {
    float *xs, *h1, *h2, *w0, *b0, *w1, *b1, *wh, *bh, *dout;      // w0 may point to global memory (stage_w0)
    int Dp;
};
__host__ __device__ inline int row_stride(int D) { int p = (D + 3) & ~3; if ((p & 31) == 0) p += 4; return p | 1; }
// first-layer weights are staged in shared memory when they fit next to the input tile; for the wide observations of the
// six- and eight-legged bodies (D = 165 / 209) they are read through L1 instead (every thread reads the same address)
__host__ __device__ inline bool stage_w0(int D) { return D <= 136; }
__host__ __device__ inline size_t tile_bytes(int D, int A) {
    return sizeof(float) * ((size_t)RSL_TILE * row_stride(D) + 2 * RSL_TILE * RSL_HP + (stage_w0(D) ? (size_t)D * RSL_H : 0) + RSL_H + RSL_H * RSL_H + RSL_H
                            + RSL_H * RSL_HW + RSL_HW + RSL_TILE * RSL_DS);
}
__device__ inline Tile carve(float* base, int D) {
    Tile t; t.Dp = row_stride(D);
    t.xs = base; base += RSL_TILE * t.Dp;
    t.h1 = base; base += RSL_TILE * RSL_HP;
    t.h2 = base; base += RSL_TILE * RSL_HP;
    t.w0 = base; base += stage_w0(D) ? D * RSL_H : 0; t.b0 = base; base += RSL_H;
    t.w1 = base; base += RSL_H * RSL_H; t.b1 = base; base += RSL_H;
    t.wh = base; base += RSL_H * RSL_HW; t.bh = base; base += RSL_HW;
    t.dout = base;
    return t;
}

// stage one net (trunk + head with `out` columns, out <= 8) into shared memory
__device__ inline void stage_net(Tile& t, const float* __restrict__ p, int D, int w0, int b0, int w1, int b1, int wh, int bh, int out) {
    if (stage_w0(D)) { for (int i = threadIdx.x; i < D * RSL_H; i += blockDim.x) t.w0[i] = p[w0 + i]; }
    else t.w0 = const_cast<float*>(p + w0);
    for (int i = threadIdx.x; i < RSL_H * RSL_H; i += blockDim.x) t.w1[i] = p[w1 + i];
    for (int i = threadIdx.x; i < RSL_H; i += blockDim.x) { t.b0[i] = p[b0 + i]; t.b1[i] = p[b1 + i]; }
    for (int i = threadIdx.x; i < RSL_H * RSL_HW; i += blockDim.x) { int r = i / RSL_HW, c = i % RSL_HW; t.wh[i] = c < out ? p[wh + r * out + c] : 0.f; }
    if (threadIdx.x < RSL_HW) t.bh[threadIdx.x] = threadIdx.x < out ? p[bh + threadIdx.x] : 0.f;
}
// stage input rows; idx == nullptr -> rows row0..row0+127 of X (row stride ldx); rows >= n are zero
__device__ inline void stage_x(const Tile& t, const float* __restrict__ X, size_t ldx, const int* __restrict__ idx, int row0, int n, int D) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < RSL_TILE; r += nw) {
        int g = row0 + r;
        const float* src = nullptr;
        if (g < n) src = X + (size_t)(idx ? idx[g] : g) * ldx;
        for (int k = lane; k < D; k += 32) t.xs[r * t.Dp + k] = src ? src[k] : 0.f;
    }
}
// one thread = one sample row: out[64] = relu(in[row][:K] * W[K][64] + b)
__device__ __forceinline__ void dense_relu_row(const float* __restrict__ in, int K, const float* __restrict__ W, const float* __restrict__ b, float* __restrict__ out_row, int act = 0) {
    float acc[RSL_H];
#pragma unroll
    for (int j = 0; j < RSL_H; j++) acc[j] = b[j];
    for (int k = 0; k < K; k++) {
        const float a = in[k];
        const float4* w4 = reinterpret_cast<const float4*>(W + k * RSL_H);
#pragma unroll
        for (int j = 0; j < RSL_H / 4; j++) {
            float4 w = w4[j];
            acc[4*j] = fmaf(a, w.x, acc[4*j]); acc[4*j+1] = fmaf(a, w.y, acc[4*j+1]);
            acc[4*j+2] = fmaf(a, w.z, acc[4*j+2]); acc[4*j+3] = fmaf(a, w.w, acc[4*j+3]);
        }
    }
#pragma unroll
    for (int j = 0; j < RSL_H; j++) out_row[j] = act ? tanhf(acc[j]) : fmaxf(acc[j], 0.f);      // relu (models.py:93-101) | tanh (policy_zoo/policy.py:52,64)
}
// forward of the staged net for this thread's row; head outputs (8 padded) returned in out8
__device__ __forceinline__ void net_forward_row(const Tile& t, int D, float* outh, int act = 0) {
    const int r = threadIdx.x;
    dense_relu_row(t.xs + r * t.Dp, D, t.w0, t.b0, t.h1 + r * RSL_HP, act);
    dense_relu_row(t.h1 + r * RSL_HP, RSL_H, t.w1, t.b1, t.h2 + r * RSL_HP, act);
#pragma unroll
    for (int c = 0; c < RSL_HW; c++) outh[c] = t.bh[c];
    for (int k = 0; k < RSL_H; k++) {
        const float a = t.h2[r * RSL_HP + k];
#pragma unroll
        for (int c = 0; c < RSL_HW; c++) outh[c] = fmaf(a, t.wh[k * RSL_HW + c], outh[c]);
    }
}

// ---- inference: mean [n, A] and/or value [n]; up to 4 independent jobs per launch (blockIdx.y), e.g. the four
//      (policy, observation) combinations of one rollout step (runner.py:62-97) ----
struct MlpJobs {
    const float* params[4]; const float* X[4]; float* mean[4]; float* value[4];
    size_t ldx[4];
    int act[4];          // hidden activation: 0 relu, 1 tanh
};
__global__ void __launch_bounds__(RSL_TILE) k_mlp_forward(MlpJobs J, int D, int A, int n) {
    extern __shared__ __align__(16) float smem[];
    const Layout L = make_layout(D, A);
    Tile t = carve(smem, D);
    const int job = blockIdx.y;
    const float* __restrict__ params = J.params[job];
    float* __restrict__ mean = J.mean[job];
    float* __restrict__ value = J.value[job];
    const int row0 = blockIdx.x * RSL_TILE, g = row0 + threadIdx.x;
    stage_x(t, J.X[job], J.ldx[job], nullptr, row0, n, D);
    float o[RSL_HW];
    if (mean) {
        stage_net(t, params, D, L.pi_w0, L.pi_b0, L.pi_w1, L.pi_b1, L.pi_w, L.pi_b, A);
        __syncthreads();
        net_forward_row(t, D, o, J.act[job]);
        if (g < n) for (int c = 0; c < A; c++) mean[(size_t)g * A + c] = o[c];
        __syncthreads();
    }
    if (value) {
        stage_net(t, params, D, L.vf_w0, L.vf_b0, L.vf_w1, L.vf_b1, L.vf_w, L.vf_b, 1);
        __syncthreads();
        net_forward_row(t, D, o, J.act[job]);
        if (g < n) value[g] = o[0];
    }
}

// ---- rollout epilogue: sample actions, neglogp under both policies, write one trajectory step ----
__device__ __forceinline__ void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float neglogp(const float* a, const float* mu, const float* logstd, int A) {
    float s = 0.f, ls = 0.f;
    for (int c = 0; c < A; c++) { float z = (a[c] - mu[c]) * expf(-logstd[c]); s += z * z; ls += logstd[c]; }
    return 0.5f * s + 0.9189385332046727f * (float)A + ls;     // distributions.py:238-241
}
// thread per env.  mu00 = pi0(o0), mu10 = pi1(o0), mu11 = pi1(o1), mu01 = pi0(o1): [E, A] each.
// outputs (all [E,...]): actions [E,2,A] for env.step; act0/act1 [E,A]; nlp[2][E]; opp_nlp[2][E]
__global__ void k_rollout_sample(int E, int A, const float* __restrict__ logstd0, const float* __restrict__ logstd1,
                                 const float* __restrict__ mu00, const float* __restrict__ mu10, const float* __restrict__ mu11,
                                 const float* __restrict__ mu01, uint64_t seed, uint32_t tick, int deterministic,
                                 float* __restrict__ actions, float* __restrict__ nlp0, float* __restrict__ nlp1,
                                 float* __restrict__ opp_nlp0, float* __restrict__ opp_nlp1) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    float a0[16], a1[16], ls0[16], ls1[16];
    for (int c = 0; c < A; c++) { ls0[c] = logstd0[c]; ls1[c] = logstd1[c]; }
    for (int c = 0; c < A; c += 2) {
        uint32_t r[4];
        philox((uint32_t)e, tick, (uint32_t)c, 0x52534c31u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        float n[4];
        for (int k = 0; k < 2; k++) {
            float u1 = ((float)(r[2*k] >> 8) + 0.5f) * (1.f / 16777216.f), u2 = ((float)(r[2*k+1] >> 8) + 0.5f) * (1.f / 16777216.f);
            float rad = sqrtf(-2.f * __logf(u1)), s, co;
            __sincosf(6.283185307179586f * u2, &s, &co);
            n[2*k] = rad * co; n[2*k+1] = rad * s;
        }
        if (deterministic) { n[0] = n[1] = n[2] = n[3] = 0.f; }
        a0[c] = mu00[(size_t)e * A + c] + __expf(ls0[c]) * n[0];
        a1[c] = mu11[(size_t)e * A + c] + __expf(ls1[c]) * n[1];
        if (c + 1 < A) {
            a0[c+1] = mu00[(size_t)e * A + c + 1] + __expf(ls0[c+1]) * n[2];
            a1[c+1] = mu11[(size_t)e * A + c + 1] + __expf(ls1[c+1]) * n[3];
        }
    }
    for (int c = 0; c < A; c++) { actions[(size_t)e * 2 * A + c] = a0[c]; actions[(size_t)e * 2 * A + A + c] = a1[c]; }
    nlp0[e] = neglogp(a0, mu00 + (size_t)e * A, ls0, A);          // -log pi0(a0|o0)        runner.py:67,84
    opp_nlp0[e] = neglogp(a0, mu10 + (size_t)e * A, ls1, A);      // -log pi1(a0|o0)        runner.py:85
    opp_nlp1[e] = neglogp(a1, mu11 + (size_t)e * A, ls1, A);      // -log pi1(a1|o1)        runner.py:87
    nlp1[e] = neglogp(a1, mu01 + (size_t)e * A, ls0, A);          // -log pi0(a1|o1)        runner.py:90
}
// neglogp of given actions under (mu, logstd): PolicyWithValue.action_probability (policies.py:107-108)
__global__ void k_neglogp(int n, int A, const float* __restrict__ act, const float* __restrict__ mu, const float* __restrict__ logstd, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ls[16];
    for (int c = 0; c < A; c++) ls[c] = logstd[c];
    out[i] = neglogp(act + (size_t)i * A, mu + (size_t)i * A, ls, A);
}

// ---- V-trace (runner.py:166-200): thread per (agent, env); arrays are [2][T][E] ----
__global__ void k_vtrace(int T, int E, float gamma, float lam, float rho_bar, float c_bar,
                         const float* __restrict__ rewards, const float* __restrict__ values, const uint8_t* __restrict__ dones,
                         const float* __restrict__ nlp, const float* __restrict__ opp_nlp,
                         const float* __restrict__ last_values /*[2][E]*/, const uint8_t* __restrict__ last_dones /*[E][2]*/,
                         float* __restrict__ returns, float* __restrict__ ratios /*[3][T][E]: off_policy, off_env, total*/) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * E) return;
    const int agt = i / E, e = i - agt * E;
    const size_t TE = (size_t)T * E;
    // the reference accumulates this recursion in float64 (1.0 - bool -> float64, runner.py:188-196) and stores float32
    double acc = 0.0;
    float nextv = last_values[(size_t)agt * E + e];
    double nextnt = 1.0 - (double)last_dones[2 * e + agt];
    for (int t = T - 1; t >= 0; t--) {
        const size_t o = (size_t)t * E + e;
        float rho = 1.f, cc = 1.f;
        if (agt == 1 || ratios) {
            float offp = expf(opp_nlp[TE + o] - nlp[TE + o]);        // exp(opp_nlp[1] - nlp[1])    runner.py:170
            float offe = expf(nlp[o] - opp_nlp[o]);                  // exp(nlp[0] - opp_nlp[0])    runner.py:171
            float ratio = offp * offe;
            if (ratios && agt == 0) { ratios[o] = offp; ratios[TE + o] = offe; ratios[2 * TE + o] = ratio; }
            // np.clip / np.minimum propagate a NaN ratio (runner.py:176-181); fminf would silently return the bound
            if (agt == 1) { rho = ratio != ratio ? ratio : fminf(ratio, rho_bar); cc = ratio != ratio ? ratio : fminf(ratio, c_bar); }
        }
        cc *= lam;
        const float v = values[(size_t)agt * TE + o], r = rewards[(size_t)agt * TE + o];
        // gamma * nextvalues is a float32 product in the reference (python float x float32 array), the rest is float64
        const double delta = (double)rho * ((double)r + (double)(gamma * nextv) * nextnt - (double)v);
        acc = delta + (double)gamma * nextnt * (double)cc * acc;
        returns[(size_t)agt * TE + o] = (float)((double)v + acc);
        nextv = v;
        nextnt = 1.0 - (double)dones[(size_t)agt * TE + o];
    }
}

// ---- PPO minibatch ------------------------------------------------------------------------------
// pass 0: advantage moments of the minibatch (model.py:182-185): sums[0] = sum adv, sums[1] = sum adv^2 (double)
__global__ void k_adv_moments(const int* __restrict__ idx, int n, const float* __restrict__ returns, const float* __restrict__ values, double* __restrict__ sums) {
    double s = 0, s2 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int g = idx ? idx[i] : i;
        double a = (double)returns[g] - (double)values[g];
        s += a; s2 += a * a;
    }
    for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    __shared__ double sh[2][32];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = s; sh[1][w] = s2; }
    __syncthreads();
    if (w == 0) {
        s = l < (blockDim.x >> 5) ? sh[0][l] : 0; s2 = l < (blockDim.x >> 5) ? sh[1][l] : 0;
        for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        if (l == 0) { atomicAdd(&sums[0], s); atomicAdd(&sums[1], s2); }
    }
}

// C[M x 64] (global partial, += semantics: written once per block) = A^T B over the tile rows:
// A = in tile [128][lda] (columns 0..M-1), B = dz tile [128][65].  Thread owns an 8x8 (M padded) sub-tile.
__device__ inline void grad_weight_tile(const float* __restrict__ in, int lda, int M, const float* __restrict__ dz, float* __restrict__ out, int rows) {
    // 128 threads: tile rows i0 = (tid / 8) * 8 over M (up to 128), cols j0 = (tid % 8) * 8
    for (int base = 0; base < M; base += 128) {
        const int i0 = base + (threadIdx.x >> 3) * 8, j0 = (threadIdx.x & 7) * 8;
        float acc[8][8];
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 8; b++) acc[a][b] = 0.f;
        if (i0 < M) {
            for (int r = 0; r < rows; r++) {
                float av[8], bv[8];
#pragma unroll
                for (int a = 0; a < 8; a++) av[a] = (i0 + a < M) ? in[r * lda + i0 + a] : 0.f;
#pragma unroll
                for (int b = 0; b < 8; b++) bv[b] = dz[r * RSL_HP + j0 + b];
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
    }
}
// column sums of a dz tile -> out[64]
__device__ inline void grad_bias_tile(const float* __restrict__ dz, float* __restrict__ out, int rows) {
    if (threadIdx.x < RSL_H) {
        float s = 0.f;
        for (int r = 0; r < rows; r++) s += dz[r * RSL_HP + threadIdx.x];
        out[threadIdx.x] = s;
    }
}
// thread per row: din[row][i] = relu'(hin[row][i]) * sum_j dz[row][j] W[i][j]   for i < 64, written in place of hin
__device__ __forceinline__ void backprop_row(const float* __restrict__ dz_row, const float* __restrict__ W /*[64][64]*/, float* __restrict__ hin_row) {
    float d[RSL_H];
#pragma unroll
    for (int j = 0; j < RSL_H; j++) d[j] = dz_row[j];
    for (int i = 0; i < RSL_H; i++) {
        const float4* w4 = reinterpret_cast<const float4*>(W + i * RSL_H);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int j = 0; j < RSL_H / 4; j++) { float4 w = w4[j]; s0 = fmaf(d[4*j], w.x, s0); s1 = fmaf(d[4*j+1], w.y, s1); s2 = fmaf(d[4*j+2], w.z, s2); s3 = fmaf(d[4*j+3], w.w, s3); }
        hin_row[i] = hin_row[i] > 0.f ? (s0 + s1) + (s2 + s3) : 0.f;
    }
}

struct PPOArgs {
    const float* params; int D, A;
    const float *obs, *actions, *returns, *values, *old_nlp, *weights;    // flat sample arrays [N,...]
    const int* idx; int n;                                                 // minibatch = idx[0..n)
    const double* adv_sums;                                                // from k_adv_moments (already all-reduced)
    double adv_count;                                                      // global minibatch size
    float cliprange, ent_coef, vf_coef, inv_n;                             // inv_n = 1 / global minibatch size
    float* gpart;       // [nblocks][P] per-block gradient partials
    float* spart;       // [nblocks][8] per-block stat partials: pg, vf, approxkl, clipfrac (sums)
    float* log_ratio;   // [n] or NULL
};

// one block = 128 samples: forward + backward of both nets, per-block partial gradients (no atomics)
__global__ void __launch_bounds__(RSL_TILE) k_ppo_tile(PPOArgs a) {
    extern __shared__ __align__(16) float smem[];
    const Layout L = make_layout(a.D, a.A);
    const int D = a.D, A = a.A;
    Tile t = carve(smem, D);
    const int row0 = blockIdx.x * RSL_TILE, r = threadIdx.x, g = row0 + r;
    const int rows = min(RSL_TILE, a.n - row0);
    const bool live = g < a.n;
    const int s = live ? (a.idx ? a.idx[g] : g) : 0;
    float* gp = a.gpart + (size_t)blockIdx.x * L.P;
    for (int i = threadIdx.x; i < L.P; i += blockDim.x) gp[i] = 0.f;
    stage_x(t, a.obs, (size_t)D, a.idx, row0, a.n, D);
    // ---------------- policy net ----------------
    stage_net(t, a.params, D, L.pi_w0, L.pi_b0, L.pi_w1, L.pi_b1, L.pi_w, L.pi_b, A);
    __syncthreads();
    float mu[RSL_HW];
    net_forward_row(t, D, mu);
    const double mean = a.adv_sums[0] / a.adv_count;
    const double var = fmax(a.adv_sums[1] / a.adv_count - mean * mean, 0.0);
    const float adv_mean = (float)mean, adv_std = (float)sqrt(var);
    float st_pg = 0.f, st_kl = 0.f, st_clip = 0.f, st_vf = 0.f;
    float dmu[RSL_HW], dls[RSL_HW];
#pragma unroll
    for (int c = 0; c < RSL_HW; c++) { dmu[c] = 0.f; dls[c] = 0.f; }
    if (live) {
        float ls[RSL_HW], z[RSL_HW], nl = 0.f, lsum = 0.f;
        for (int c = 0; c < A; c++) {
            ls[c] = a.params[L.logstd + c];
            z[c] = (a.actions[(size_t)s * A + c] - mu[c]) * expf(-ls[c]);
            nl += z[c] * z[c]; lsum += ls[c];
        }
        nl = 0.5f * nl + 0.9189385332046727f * (float)A + lsum;
        const float old = a.old_nlp[s];
        float ratio = expf(old - nl);
        const bool isnan_ = ratio != ratio;
        if (isnan_) ratio = 2.f;                                            // model.py:96
        const float adv = (a.returns[s] - a.values[s] - adv_mean) / (adv_std + 1e-8f);
        const float w = a.weights ? a.weights[s] : 1.f;
        const float rc = fminf(fmaxf(ratio, 1.f - a.cliprange), 1.f + a.cliprange);
        const float pg1 = -adv * ratio, pg2 = -adv * rc;
        st_pg = w * fmaxf(pg1, pg2);
        st_kl = nl - old;
        st_clip = fabsf(ratio - 1.f) > a.cliprange ? 1.f : 0.f;
        if (a.log_ratio) a.log_ratio[g] = old - nl;
        // d loss / d neglogp
        float gnl = 0.f;
        if (!isnan_) {
            const bool inside = (ratio >= 1.f - a.cliprange) && (ratio <= 1.f + a.cliprange);
            if (pg1 >= pg2) gnl = adv * ratio; else if (inside) gnl = adv * ratio;
        }
        gnl *= w * a.inv_n;
        for (int c = 0; c < A; c++) {
            dmu[c] = gnl * (-z[c] * expf(-ls[c]));            // d nlp / d mu = -(a - mu) / sigma^2
            dls[c] = gnl * (1.f - z[c] * z[c]);               // d nlp / d logstd
        }
    }
    // head grads: dWp[64][A] = h2^T dmu ; dbp ; dlogstd (column sums via the dout tile)
    for (int c = 0; c < RSL_HW; c++) t.dout[r * RSL_DS + c] = dmu[c];
    
    __syncthreads();
    for (int o = threadIdx.x; o < RSL_H * A; o += blockDim.x) {
        const int k = o / A, c = o - k * A;
        float acc = 0.f;
        for (int q = 0; q < rows; q++) acc = fmaf(t.h2[q * RSL_HP + k], t.dout[q * RSL_DS + c], acc);
        gp[L.pi_w + o] = acc;
    }
    if (threadIdx.x < A) { float acc = 0.f; for (int q = 0; q < rows; q++) acc += t.dout[q * RSL_DS + threadIdx.x]; gp[L.pi_b + threadIdx.x] = acc; }
    __syncthreads();
    for (int c = 0; c < RSL_HW; c++) t.dout[r * RSL_DS + c] = dls[c];
    __syncthreads();
    if (threadIdx.x < A) { float acc = 0.f; for (int q = 0; q < rows; q++) acc += t.dout[q * RSL_DS + threadIdx.x]; gp[L.logstd + threadIdx.x] = acc; }
    // dz2 = relu'(h2) * (dmu Wp^T), in place of h2
    {
        float* h2r = t.h2 + r * RSL_HP;
        for (int k = 0; k < RSL_H; k++) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < RSL_HW; c++) acc = fmaf(dmu[c], t.wh[k * RSL_HW + c], acc);
            h2r[k] = h2r[k] > 0.f ? acc : 0.f;
        }
    }
    __syncthreads();
    grad_weight_tile(t.h1, RSL_HP, RSL_H, t.h2, gp + L.pi_w1, rows);
    grad_bias_tile(t.h2, gp + L.pi_b1, rows);
    __syncthreads();
    backprop_row(t.h2 + r * RSL_HP, t.w1, t.h1 + r * RSL_HP);               // dz1 in place of h1
    __syncthreads();
    grad_weight_tile(t.xs, t.Dp, D, t.h1, gp + L.pi_w0, rows);
    grad_bias_tile(t.h1, gp + L.pi_b0, rows);
    __syncthreads();
    // ---------------- value net ----------------
    stage_net(t, a.params, D, L.vf_w0, L.vf_b0, L.vf_w1, L.vf_b1, L.vf_w, L.vf_b, 1);
    __syncthreads();
    float vo[RSL_HW];
    net_forward_row(t, D, vo);
    float dv = 0.f;
    if (live) {
        const float err = vo[0] - a.returns[s];
        st_vf = 0.5f * err * err;
        dv = a.vf_coef * err * a.inv_n;                                      // d (vf_coef * .5 mean(err^2)) / dv
    }
    t.dout[r * RSL_DS] = dv;
    __syncthreads();
    if (threadIdx.x < RSL_H) { float acc = 0.f; for (int q = 0; q < rows; q++) acc = fmaf(t.h2[q * RSL_HP + threadIdx.x], t.dout[q * RSL_DS], acc); gp[L.vf_w + threadIdx.x] = acc; }
    if (threadIdx.x == 64) { float acc = 0.f; for (int q = 0; q < rows; q++) acc += t.dout[q * RSL_DS]; gp[L.vf_b] = acc; }
    __syncthreads();
    {
        float* h2r = t.h2 + r * RSL_HP;
        for (int k = 0; k < RSL_H; k++) h2r[k] = h2r[k] > 0.f ? dv * t.wh[k * RSL_HW] : 0.f;
    }
    __syncthreads();
    grad_weight_tile(t.h1, RSL_HP, RSL_H, t.h2, gp + L.vf_w1, rows);
    grad_bias_tile(t.h2, gp + L.vf_b1, rows);
    __syncthreads();
    backprop_row(t.h2 + r * RSL_HP, t.w1, t.h1 + r * RSL_HP);
    __syncthreads();
    grad_weight_tile(t.xs, t.Dp, D, t.h1, gp + L.vf_w0, rows);
    grad_bias_tile(t.h1, gp + L.vf_b0, rows);
    // ---------------- stat partials ----------------
    float v4[4] = { st_pg, st_vf, st_kl, st_clip };
    __syncthreads();
    for (int k = 0; k < 4; k++) {
        float x = v4[k];
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0) t.dout[k * 4 + (threadIdx.x >> 5)] = x;
    }
    __syncthreads();
    if (threadIdx.x < 4) a.spart[(size_t)blockIdx.x * 8 + threadIdx.x] = t.dout[threadIdx.x * 4] + t.dout[threadIdx.x * 4 + 1] + t.dout[threadIdx.x * 4 + 2] + t.dout[threadIdx.x * 4 + 3];
}

// trajectory writes of one rollout step (runner.py:70-72,95-100): before the env step the observation and done flags the policies
// acted on, after it the reward terms and the episode records
__global__ void k_traj_pre(int E, int D, int t, int T, const float* __restrict__ obs, const uint8_t* __restrict__ done,
                           float* __restrict__ mb_obs, uint8_t* __restrict__ mb_dones) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // over [E][2][D]
    if (i >= (long long)E * 2 * D) return;
    const int k = (int)(i % D), a = (int)((i / D) & 1);
    const long long e = i / (2 * D);
    mb_obs[(((long long)a * T + t) * E + e) * D + k] = obs[i];
    if (k == 0) mb_dones[((long long)a * T + t) * E + e] = done[2 * e + a];
}
__global__ void k_traj_post(int E, int t, int T, const float* __restrict__ info, const uint8_t* __restrict__ done, const float* __restrict__ epi,
                            float* __restrict__ mb_shaping, float* __restrict__ mb_main, uint8_t* __restrict__ ep_done, float* __restrict__ ep_info) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    for (int a = 0; a < 2; a++) {
        mb_shaping[((long long)a * T + t) * E + e] = info[(e * 2 + a) * RS_INFO_DIM + 6];
        mb_main[((long long)a * T + t) * E + e] = info[(e * 2 + a) * RS_INFO_DIM + 3];
    }
    ep_done[(long long)t * E + e] = done[2 * e];
    for (int k = 0; k < 3; k++) ep_info[((long long)t * E + e) * 3 + k] = epi[e * 3 + k];
}

// [pg_loss, vf_loss, entropy, approxkl, clipfrac] (model.py:137): means of the four stat sums stored behind the gradient, and the
// DiagGaussian entropy sum(logstd + 0.5 log(2 pi e)) (distributions.py:244-245) of the parameters as they are now
__global__ void k_ppo_stats(const float* __restrict__ grad_stats, const float* __restrict__ params, int P, int logstd_off, int A,
                            double inv_n, double* __restrict__ out) {
    if (threadIdx.x == 0) {
        double ent = 0;
        for (int i = 0; i < A; i++) ent += (double)params[logstd_off + i] + 1.4189385332046727;      // 0.5 * log(2 pi e)
        out[0] = (double)grad_stats[P] * inv_n; out[1] = (double)grad_stats[P + 1] * inv_n; out[2] = ent;
        out[3] = (double)grad_stats[P + 2] * inv_n; out[4] = (double)grad_stats[P + 3] * inv_n;
    }
}

// ---- data-parallel minibatch schedule on the device ---------------------------------------------------------------------------
// Every rank holds the same GLOBAL permutation of an epoch (the reference's np.random.shuffle, replayed bit-exactly on the host
// and uploaded once per epoch).  Minibatch m is the slice perm[m * nbt, (m + 1) * nbt); this kernel keeps, in order, the entries
// that fall in this rank's sample range [lo, hi) as LOCAL indices: out_idx[m][0 .. counts[m]).  One block per minibatch, stable
// compaction by ballot / prefix sums (replaces the NumPy boolean masks over the global index array of round 1).
__global__ void __launch_bounds__(1024) k_epoch_split(const int* __restrict__ perm, long long N, int nbt, long long lo, long long hi,
                                                      int* __restrict__ out_idx, int* __restrict__ counts) {
    __shared__ int wsum[32];
    __shared__ int base_s;
    const int m = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long s0 = (long long)m * nbt, s1 = s0 + nbt < N ? s0 + nbt : N;
    int* out = out_idx + (long long)m * nbt;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (long long p0 = s0; p0 < s1; p0 += blockDim.x) {
        const long long p = p0 + threadIdx.x;
        long long v = -1;
        if (p < s1) v = perm[p];
        const bool in = v >= lo && v < hi;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const int rank = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) wsum[w] = __popc(bal);
        __syncthreads();
        int off = 0;
        for (int q = 0; q < w; q++) off += wsum[q];
        const int base = base_s;
        if (in) out[base + off + rank] = (int)(v - lo);
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < (int)(blockDim.x >> 5); q++) t += wsum[q]; base_s = base + t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[m] = base_s;
}
// advantage moments of ALL minibatches of an epoch in one launch (returns and values do not change during the update, so the
// per-minibatch normalisation of model.py:182-185 needs no launch -- and, data-parallel, no collective -- per minibatch):
// sums[m] = (sum adv, sum adv^2) over idx[m * cap + 0 .. n_m), n_m = counts[m] or, without counts, the slice length.
// One block per minibatch, fixed reduction order (deterministic).
__global__ void __launch_bounds__(1024) k_adv_moments_multi(const int* __restrict__ idx, const int* __restrict__ counts, int cap, long long n_total,
                                                            const float* __restrict__ returns, const float* __restrict__ values, double* __restrict__ sums) {
    const int m = blockIdx.x;
    long long n = counts ? counts[m] : (n_total - (long long)m * cap < cap ? n_total - (long long)m * cap : cap);
    const int* ix = idx ? idx + (long long)m * cap : nullptr;
    double s = 0, s2 = 0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const long long g = ix ? ix[i] : (long long)m * cap + i;
        const double a = (double)returns[g] - (double)values[g];
        s += a; s2 += a * a;
    }
    for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    __shared__ double sh[2][32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = s; sh[1][w] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0, t2 = 0;
        for (int q = 0; q < (int)(blockDim.x >> 5); q++) { t += sh[0][q]; t2 += sh[1][q]; }
        sums[2 * m] = t; sums[2 * m + 1] = t2;
    }
}

// ---- fused minibatch epilogue: two launches instead of four --------------------------------------------------------------------
// (1) per-block partials -> grad_stats [P + 4] in block order (deterministic) and, by block 0, the entropy of the PRE-update
//     parameters into stats5[2] (model.py:137 evaluates it in the same session.run as the train op);
// (2) after the data-parallel all-reduce (if any): global norm, clip, TF-style Adam and the other four statistics.
//     Every block recomputes the sum of squares over the whole gradient in the same fixed order (24.5 k numbers: cheaper than a
//     separate single-block launch plus its launch gap, and deterministic).
__global__ void k_grad_reduce2(const float* __restrict__ gpart, const float* __restrict__ spart, int nblocks, int P,
                               float* __restrict__ grad, const float* __restrict__ params, int logstd_off, int A, float ent_share,
                               double* __restrict__ stats5) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) {
        float s = 0.f;
        for (int b = 0; b < nblocks; b++) s += gpart[(size_t)b * P + i];
        // - ent_coef * d entropy / d logstd (= 1 per component): this rank's share n_local / n_global of it, so that the all-reduced
        // gradient carries the term exactly once and grad_stats is the final gradient on one GPU
        if (i >= logstd_off && i < logstd_off + A) s -= ent_share;
        grad[i] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x < 4) {
        float s = 0.f;
        for (int b = 0; b < nblocks; b++) s += spart[(size_t)b * 8 + threadIdx.x];
        grad[P + threadIdx.x] = s;
    }
    if (stats5 && blockIdx.x == 0 && threadIdx.x == 32) {
        double ent = 0;
        for (int q = 0; q < A; q++) ent += (double)params[logstd_off + q] + 1.4189385332046727;      // 0.5 * log(2 pi e)
        stats5[2] = ent;
    }
}
__global__ void __launch_bounds__(256) k_adam2(float* __restrict__ params, float* __restrict__ m, float* __restrict__ v, float* __restrict__ grad,
                                               int P, float max_grad_norm, float lr_t, float b1, float b2,
                                               float eps, float* __restrict__ gnorm_out, double inv_n, double* __restrict__ stats5) {
    __shared__ double sh[8];
    double sq = 0;
    for (int j = threadIdx.x; j < P; j += blockDim.x) { const float g = grad[j]; sq += (double)g * (double)g; }
    for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = sq;
    __syncthreads();
    double tot = 0;
    for (int q = 0; q < 8; q++) tot += sh[q];
    const float gn = (float)sqrt(tot);
    float scale = 1.f;
    if (max_grad_norm > 0.f) scale = max_grad_norm / fmaxf(gn, max_grad_norm);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        if (gnorm_out) *gnorm_out = gn;
        if (stats5) { stats5[0] = (double)grad[P] * inv_n; stats5[1] = (double)grad[P + 1] * inv_n; stats5[3] = (double)grad[P + 2] * inv_n; stats5[4] = (double)grad[P + 3] * inv_n; }
    }
    if (i >= P) return;
    const float g = grad[i] * scale;
    const float mi = b1 * m[i] + (1.f - b1) * g;
    const float vi = b2 * v[i] + (1.f - b2) * g * g;
    m[i] = mi; v[i] = vi;
    params[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

}  // namespace rsl
