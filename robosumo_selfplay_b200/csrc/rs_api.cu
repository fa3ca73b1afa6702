// rs_api.cu -- CUDA kernels (one warp per env pair) and the C ABI of include/rs_b200.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <type_traits>
#ifndef RS_WPB
#define RS_WPB 28  // max warps (env pairs) per block; one block per SM: 28 Ant slabs of 8.2 KB fill the 227 KB (E = 4096 -> one wave)
#endif
// Re-alignment of the warps of a block (instruction-cache locality vs. waiting for the slowest warp), see simulate() in rs_core.h:
//   RS_SYNC_MODE 7 (default) once per trip = [evaluation start] + [one Newton iteration] (the per-warp progress machine),
//                1 once per forward evaluation (round-1 behaviour), 0 never.
// Measured and dropped in round 1: per substep, named-barrier groups of 4-14 warps, per phase, a one-evaluation sliding window.
#ifndef RS_SYNC_MODE
#define RS_SYNC_MODE 7
#endif
#if RS_SYNC_MODE != 7
#define RS_TRIP_MACHINE 0
#endif
#ifdef RS_EXPERIMENT_CLOCK
__device__ long long rs_dbg[4200 * 128];
extern "C" int rs_debug_read(void* dst, size_t bytes) { return (int)cudaMemcpyFromSymbol(dst, rs_dbg, bytes); }
#endif
__device__ __forceinline__ long long rs_clock() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
#if defined(__CUDA_ARCH__)
#if RS_SYNC_MODE == 7
#define RS_TRIP_ANY(p) __syncthreads_or(p)
#define RS_TRIP_SYNC() __syncthreads()
#endif
#if RS_SYNC_MODE != 0
#define RS_EVAL_SYNC() __syncthreads()      // forward(): the per-evaluation path (six- and eight-legged bodies, rs_forward_debug)
#endif
#ifdef RS_EXPERIMENT_CLOCK
#define RS_ACC(i) { __syncwarp(); long long rs_now = rs_clock(); c.acc[i] += rs_now - c.tlast; c.tlast = rs_now; }
#define RS_CLOCK_BEGIN() { __syncwarp(); if ((threadIdx.x & 31) == 0 && c.evk < 20) rs_dbg[(size_t)c.env * 128 + 3 * c.evk] = rs_clock(); }
#define RS_CLOCK_MARK(i) { if (i == 2) { __syncwarp(); if ((threadIdx.x & 31) == 0 && c.evk < 20) rs_dbg[(size_t)c.env * 128 + 3 * c.evk + 1] = rs_clock(); } }
#define RS_CLOCK_END() { __syncwarp(); if ((threadIdx.x & 31) == 0 && c.evk < 20) { rs_dbg[(size_t)c.env * 128 + 3 * c.evk + 2] = rs_clock(); rs_dbg[(size_t)c.env * 128 + 64 + c.evk] = c.s->niter | (c.s->coupled << 8) | (c.s->ncon << 16); } c.evk++; }
#endif
#endif
#define RS_LOCKSTEP 1
#include "rs_env.h"
#include "rs_learn.cuh"
#include "rs_tc.cuh"
#include "rs_learn_tc.cuh"

using namespace rs;

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches(0);
static int fail(int code, const char* fmt, const char* detail) {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
#define CUDA_OK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return fail(RS_ERR_CUDA, "CUDA error: %s", cudaGetErrorString(_e)); } while (0)

struct EnvDev {
    int E;
    float *qpos, *qvel, *warm, *ep_ret, *ep_dret;
    int* pred;                         // active-set prediction carried across steps: [E][4 + MAXC/2] = limit masks (non-zero, negative side, loaded), nprev, cprev pairs
    int* latch;                        // [2]: OR of every status bit any env raised since the last rs_status_latch(clear) -- auto-reset does not clear it; number of env-steps that raised one
    int* next;                         // [2] pair counters of the persistent k_step (this launch: next[tick], cleared for the next one)
    int tick, persistent;
    int *ep_step, *status, *diag;      // diag[E][4]: Newton iterations, coupled evaluations, contacts summed over the last env step, max iterations of one evaluation
    unsigned int* ep_count;
    const rs_agent_model* am;
    EnvParams P;
    float h;
    int max_newton;
};

struct rs_env {
    rs_config cfg;
    int LA, LB, nq, nv, nu, obsA, obsB;
    EnvDev d;
    rs_agent_model* d_am;
    // staging for the host-buffer entry point
    float *h_act, *h_obs, *h_rew, *h_info, *h_epi;
    uint8_t* h_done;
    float *s_act, *s_obs, *s_rew, *s_info, *s_epi;
    uint8_t* s_done;
    cudaStream_t stream;
    size_t smem;
    int wpb;      // warps (env pairs) per block: as many slabs as fit in one SM's shared memory, at most RS_WPB
    int sms;      // multiprocessors of the device (grid of the persistent k_step)
};

template <int LA, int LB>
__device__ __forceinline__ Slab<LA, LB>* warp_setup(Ctx<LA, LB>& c, const EnvDev& d, rs_agent_model* sm_am, unsigned char* smem_raw) {
    typedef Slab<LA, LB> S;
    for (int i = threadIdx.x; i < (int)(2 * sizeof(rs_agent_model) / 4); i += blockDim.x) ((int*)sm_am)[i] = ((const int*)d.am)[i];
    __syncthreads();
    S* s = reinterpret_cast<S*>(smem_raw) + __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // warp-uniform on purpose: cheaper to keep / rebuild than a per-lane value (-1.7 % step time)
    c.s = s; c.am = sm_am; c.h = d.h; c.max_newton = d.max_newton;
    return s;
}

template <int LA, int LB>
__device__ __forceinline__ void load_state(Ctx<LA, LB>& c, const EnvDev& d, int e) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    RS_LANE_LOOP(i, S::NQ) { s.q[i] = d.qpos[(size_t)e * S::NQ + i]; }
    RS_LANE_LOOP(i, S::NV) { s.v[i] = d.qvel[(size_t)e * S::NV + i]; s.x[i] = d.warm[(size_t)e * S::NV + i]; }
    // active sets at the end of the previous step (limit rows as three bit masks, contact rows by key): the first evaluation of
    // this step predicts from them exactly as later evaluations predict from their predecessor
    const int* pr = d.pred + (size_t)e * (4 + S::MAXC / 2);
    const int nz = pr[0], neg = pr[1], loaded = pr[2];
    RS_LANE_LOOP(j, S::NU) { s.lsgn[j] = ((nz >> j) & 1) ? (((neg >> j) & 1) ? -1.f : 1.f) : 0.f; s.ljar[j] = ((loaded >> j) & 1) ? -1.f : 1.f; }
    RS_LANE_LOOP(k, S::MAXC / 2) { const int w = pr[4 + k]; s.cprev[2 * k] = (unsigned short)(w & 0xFFFF); s.cprev[2 * k + 1] = (unsigned short)((unsigned)w >> 16); }
    if (RS_LANE0) { s.status = d.status[e]; s.ncon = 0; s.niter = 0; s.tot_iter = 0; s.tot_coupled = 0; s.tot_ncon = 0; s.max_iter = 0; s.nprev = pr[3]; }
    RS_SYNC();
}
template <int LA, int LB>
__device__ __forceinline__ void clear_prediction(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    RS_LANE_LOOP(j, S::NU) { s.lsgn[j] = 0.f; s.ljar[j] = 1.f; }
    if (RS_LANE0) s.nprev = 0;
    RS_SYNC();
}
template <int LA, int LB>
__device__ __forceinline__ void store_state(Ctx<LA, LB>& c, const EnvDev& d, int e) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    RS_LANE_LOOP(i, S::NQ) { d.qpos[(size_t)e * S::NQ + i] = s.q[i]; }
    RS_LANE_LOOP(i, S::NV) { d.qvel[(size_t)e * S::NV + i] = s.v[i]; d.warm[(size_t)e * S::NV + i] = s.x[i]; }
    static_assert(S::NU <= 32, "limit masks are one word");
    const int lane = threadIdx.x & 31;
    const float sg = lane < S::NU ? s.lsgn[lane] : 0.f, jr = lane < S::NU ? s.ljar[lane] : 1.f;
    const unsigned nz = __ballot_sync(0xffffffffu, sg != 0.f), neg = __ballot_sync(0xffffffffu, sg < 0.f), loaded = __ballot_sync(0xffffffffu, jr < 0.f);
    int* pr = d.pred + (size_t)e * (4 + S::MAXC / 2);
    if (lane == 0) { pr[0] = (int)nz; pr[1] = (int)neg; pr[2] = (int)loaded; pr[3] = s.nprev; }
    RS_LANE_LOOP(k, S::MAXC / 2) { pr[4 + k] = (int)((unsigned)s.cprev[2 * k] | ((unsigned)s.cprev[2 * k + 1] << 16)); }
}
template <int LA, int LB>
__device__ __forceinline__ void set_act(Ctx<LA, LB>& c, const float* ctrl) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    RS_LANE_LOOP(u, S::NU) {
        float x = ctrl[u];
        x = fminf(fmaxf(x, -1.f), 1.f);                       // ctrllimited, ctrlrange +-1
        s.act[u] = c.am[u >= 2 * LA ? 1 : 0].gear * x;
    }
    RS_SYNC();
}
__device__ __forceinline__ float tsfeat(int step) { return (float)(-1.0 + 2.0 * (double)step / 500.0); }

template <int LA, int LB>
__global__ void __launch_bounds__(32 * RS_WPB) k_reset(EnvDev d, const uint8_t* mask, float* obs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ rs_agent_model sm_am[2];
    typedef Slab<LA, LB> S;
    Ctx<LA, LB> c;
    warp_setup(c, d, sm_am, smem_raw);
    int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= d.E) return;
    if (mask && !mask[e]) return;
    S& s = *c.s;
    unsigned int ep = d.ep_count[e] + 1;
    env_reset_state(c, d.P, (uint32_t)e, ep);
    clear_prediction(c);
    store_state(c, d, e);
    if (RS_LANE0) { d.ep_count[e] = ep; d.ep_step[e] = 0; d.ep_ret[e] = 0.f; d.ep_dret[e] = 0.f; d.status[e] = 0; }
    const int OD = (7 + 2*LA) + (6 + 2*LA) + 6 * (1 + 3*LA) + 14 + (7 + 2*LB) + (6 + 2*LB) + 6 * (1 + 3*LB) + 14;
    if (obs) env_write_obs(c, obs + (size_t)e * OD, -1.f);
    (void)s;
}

template <int LA, int LB>
__global__ void __launch_bounds__(32 * RS_WPB) k_set_state(EnvDev d, const float* qpos, const float* qvel, float* obs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ rs_agent_model sm_am[2];
    typedef Slab<LA, LB> S;
    Ctx<LA, LB> c;
    warp_setup(c, d, sm_am, smem_raw);
    int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= d.E) return;
    S& s = *c.s;
    RS_LANE_LOOP(i, S::NQ) { s.q[i] = qpos[(size_t)e * S::NQ + i]; }
    RS_LANE_LOOP(i, S::NV) { s.v[i] = qvel[(size_t)e * S::NV + i]; s.x[i] = 0.f; }
    RS_SYNC();
    RS_LANE_LOOP(a, 2) {
        float* qq = s.q + c.qadr(a);
        float n = sqrtf(qq[3]*qq[3] + qq[4]*qq[4] + qq[5]*qq[5] + qq[6]*qq[6]);
        float inv = n > 1e-12f ? 1.f / n : 1.f;
        qq[3] *= inv; qq[4] *= inv; qq[5] *= inv; qq[6] *= inv;
    }
    RS_SYNC();
    clear_prediction(c);
    store_state(c, d, e);
    if (RS_LANE0) d.status[e] = 0;
    const int OD = (7 + 2*LA) + (6 + 2*LA) + 6 * (1 + 3*LA) + 14 + (7 + 2*LB) + (6 + 2*LB) + 6 * (1 + 3*LB) + 14;
    if (obs) env_write_obs(c, obs + (size_t)e * OD, tsfeat(d.ep_step[e]));
}

// everything of an env step behind the physics, for the pair in this warp's slab: mj_checkPos / mj_checkVel analogue, rewards and
// flags, episode bookkeeping, auto-reset, observation
template <int LA, int LB>
__device__ __forceinline__ void pair_finish(Ctx<LA, LB>& c, const EnvDev& d, int e, const float* before, const float* __restrict__ actions,
                                            float* __restrict__ obs, float* __restrict__ rew, uint8_t* __restrict__ done,
                                            float* __restrict__ info, float* __restrict__ episode, int auto_reset, bool live) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    const int lane = threadIdx.x & 31;
    const float* act = actions + (size_t)e * S::NU;
    {
        bool bad = false;
        RS_LANE_LOOP(i, S::NQ) { if (!isfinite(s.q[i])) bad = true; }
        RS_LANE_LOOP(i, S::NV) { if (!isfinite(s.v[i])) bad = true; }
        if (__any_sync(0xffffffffu, bad)) { if (lane == 0) s.status |= RS_STATUS_NAN; }
        RS_SYNC();
    }
    int num_steps = d.ep_step[e] + 1;
    StepOut o;
    env_rewards(c, d.P, before, act, num_steps, d.h * d.P.frame_skip, &o);
    float er = d.ep_ret[e] + o.rew[0], edr = d.ep_dret[e] + o.info[0][6];
    if (!live) return;
    if (lane == 0) { d.diag[4 * e] = s.tot_iter; d.diag[4 * e + 1] = s.tot_coupled; d.diag[4 * e + 2] = s.tot_ncon; d.diag[4 * e + 3] = s.max_iter; }
#ifdef RS_EXPERIMENT_CLOCK
    if (lane == 0) { rs_dbg[(size_t)e * 128 + 61] = rs_clock(); for (int i = 0; i < 6; i++) rs_dbg[(size_t)e * 128 + 90 + i] = c.acc[i]; }
#endif
    if (lane == 0) {
        rew[2 * e] = o.rew[0]; rew[2 * e + 1] = o.rew[1];
        done[2 * e] = (uint8_t)o.done[0]; done[2 * e + 1] = (uint8_t)o.done[1];
        if (episode) { episode[3 * e] = er; episode[3 * e + 1] = edr; episode[3 * e + 2] = (float)num_steps; }
    }
    if (info && lane < 2 * RS_INFO_DIM) info[(size_t)e * 2 * RS_INFO_DIM + lane] = o.info[lane / RS_INFO_DIM][lane % RS_INFO_DIM];
    const int OD = (7 + 2*LA) + (6 + 2*LA) + 6 * (1 + 3*LA) + 14 + (7 + 2*LB) + (6 + 2*LB) + 6 * (1 + 3*LB) + 14;
    RS_SYNC();
    int st = s.status;
    if (lane == 0) { const int fresh = st & ~d.status[e] & 7; if (fresh) { atomicOr(d.latch, fresh); atomicAdd(d.latch + 1, 1); } }      // bits raised during THIS step
    if (o.done[0] && auto_reset) {
        unsigned int ep = d.ep_count[e] + 1;
        env_reset_state(c, d.P, (uint32_t)e, ep);
        clear_prediction(c);
        store_state(c, d, e);
        if (lane == 0) { d.ep_count[e] = ep; d.ep_step[e] = 0; d.ep_ret[e] = 0.f; d.ep_dret[e] = 0.f; d.status[e] = st & RS_STATUS_CONTACT_FULL; }
        env_write_obs(c, obs + (size_t)e * OD, -1.f);
    } else {
        store_state(c, d, e);
        if (lane == 0) { d.ep_step[e] = num_steps; d.ep_ret[e] = er; d.ep_dret[e] = edr; d.status[e] = st | (o.done[0] ? 8 : 0); }
        env_write_obs(c, obs + (size_t)e * OD, tsfeat(num_steps));
    }
    RS_SYNC();
}

template <int LA, int LB>
__device__ __forceinline__ void pair_begin(Ctx<LA, LB>& c, const EnvDev& d, int e, const float* __restrict__ actions, float* before) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    load_state(c, d, e);
    set_act(c, actions + (size_t)e * S::NU);
    before[0] = s.q[c.qadr(0)]; before[1] = s.q[c.qadr(0) + 1]; before[2] = s.q[c.qadr(1)]; before[3] = s.q[c.qadr(1) + 1];
    RS_SYNC();
#ifdef RS_EXPERIMENT_CLOCK
    c.evk = 0; c.env = e; for (int i = 0; i < 6; i++) c.acc[i] = 0; c.tlast = rs_clock(); if ((threadIdx.x & 31) == 0) rs_dbg[(size_t)e * 128 + 60] = rs_clock();
#endif
}

// Ant pairs (the per-trip re-alignment of simulate_trips): PERSISTENT blocks, one per SM.  A warp that finishes its pair writes the
// pair's outputs and takes the next pair off a device-wide counter while the rest of the block carries on, so with several pairs per
// warp slot nobody idles until the slowest pair of a block of 28 is through (that wait was a third of the block time); with at
// most one pair per slot (E <= SMs x 28) it degenerates to one pair per warp.  The counter of the NEXT launch is cleared here
// (launches of one env are stream-ordered).  Other morphologies: one pair per warp, re-aligned per evaluation.
template <int LA, int LB>
__global__ void __launch_bounds__(32 * RS_WPB) k_step(EnvDev d, const float* __restrict__ actions, float* __restrict__ obs,
                                                       float* __restrict__ rew, uint8_t* __restrict__ done,
                                                       float* __restrict__ info, float* __restrict__ episode, int auto_reset) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ rs_agent_model sm_am[2];
    typedef Slab<LA, LB> S;
    Ctx<LA, LB> c;
    warp_setup(c, d, sm_am, smem_raw);
    float before[4];
    if constexpr (RS_TRIP_MACHINE && LA + LB <= RS_TRIP_MAX_LEGS) {
        if (blockIdx.x == 0 && threadIdx.x == 0) d.next[d.tick ^ 1] = 0;
        int e = -1;
        simulate_trips(c, d.P.frame_skip, [&](bool finish) -> bool {
            if (finish) pair_finish(c, d, e, before, actions, obs, rew, done, info, episode, auto_reset, true);
            int ne = 0;
            if (d.persistent) {
                if ((threadIdx.x & 31) == 0) ne = atomicAdd(d.next + d.tick, 1);
                ne = __shfl_sync(0xffffffffu, ne, 0);
            } else ne = finish ? d.E : (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));      // one wave: one pair per warp
            if (ne >= d.E) return false;
            e = ne;
            pair_begin(c, d, e, actions, before);
            return true;
        });
    } else {
        int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
#ifdef RS_LOCKSTEP
        const bool live = e < d.E;       // surplus warps of the last block shadow the last env (they must reach the block barriers)
        if (!live) e = d.E - 1;
#else
        const bool live = true;
        if (e >= d.E) return;
#endif
        pair_begin(c, d, e, actions, before);
        simulate(c, d.P.frame_skip);
        pair_finish(c, d, e, before, actions, obs, rew, done, info, episode, auto_reset, live);
    }
}

template <int LA, int LB>
__global__ void __launch_bounds__(32 * RS_WPB) k_forward_debug(EnvDev d, const float* ctrl, float* qacc, int* ncon, int* niter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ rs_agent_model sm_am[2];
    typedef Slab<LA, LB> S;
    Ctx<LA, LB> c;
    warp_setup(c, d, sm_am, smem_raw);
    int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const bool live = e < d.E;
    if (!live) e = d.E - 1;
    S& s = *c.s;
    load_state(c, d, e);
    set_act(c, ctrl + (size_t)e * S::NU);
    forward(c);
    if (!live) return;
    RS_LANE_LOOP(i, S::NV) { qacc[(size_t)e * S::NV + i] = s.x[i]; }
    if (RS_LANE0) { ncon[e] = s.ncon; niter[e] = s.niter; }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// morphology pairs: legs per agent in {4 (ant), 6 (bug), 8 (spider)}, all nine combinations (robosumo/__init__.py:8-105)
#ifdef RS_DEV_ANT_ONLY      /* developer builds: only Ant-vs-Ant (the full build takes ~85 s) */
#define RS_FOR_PAIRS(X) X(4, 4)
#elif defined(RS_DEV_SAME_ONLY) /* developer builds: the three same-morphology pairs */
#define RS_FOR_PAIRS(X) X(4, 4) X(6, 6) X(8, 8)
#else
#define RS_FOR_PAIRS(X) X(4, 4) X(4, 6) X(4, 8) X(6, 4) X(6, 6) X(6, 8) X(8, 4) X(8, 6) X(8, 8)
#endif
template <typename F> static int dispatch(const rs_env* h, F f) {
#define RS_CASE(A, B) if (h->LA == A && h->LB == B) return f(std::integral_constant<int, A>(), std::integral_constant<int, B>());
    RS_FOR_PAIRS(RS_CASE)
#undef RS_CASE
    return fail(RS_ERR_UNSUPPORTED, "unsupported morphology pair (%s)", "legs");
}
static size_t slab_bytes(int LA, int LB) {
#define RS_CASE(A, B) if (LA == A && LB == B) return sizeof(Slab<A, B>);
    RS_FOR_PAIRS(RS_CASE)
#undef RS_CASE
    return 0;
}

extern "C" {

int rs_agent_model_size(void) { return (int)sizeof(rs_agent_model); }
const char* rs_last_error(void) { return g_err; }
long long rs_launch_count(void) { return g_launches.load(); }

int rs_create(const rs_config* cfg, const rs_agent_model* agents, rs_env** out) {
    if (!cfg || !agents || !out || cfg->num_envs <= 0) return fail(RS_ERR_ARG, "rs_create: bad argument%s", "");
    CUDA_OK(cudaSetDevice(cfg->device));
    rs_env* h = new rs_env();
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->LA = agents[0].L; h->LB = agents[1].L;
    h->nq = agents[0].nq + agents[1].nq; h->nv = agents[0].nv + agents[1].nv; h->nu = agents[0].nu + agents[1].nu;
    h->obsA = agents[0].nq + agents[0].nv + 6 * (1 + 3 * agents[0].L) + 14;
    h->obsB = agents[1].nq + agents[1].nv + 6 * (1 + 3 * agents[1].L) + 14;
    size_t sb = slab_bytes(h->LA, h->LB);
    if (!sb) { delete h; return fail(RS_ERR_UNSUPPORTED, "unsupported morphology pair%s", ""); }
    h->wpb = (int)((227 * 1024 - 2 * sizeof(rs_agent_model)) / sb);
    if (h->wpb > RS_WPB) h->wpb = RS_WPB;
    if (h->wpb < 1) { delete h; return fail(RS_ERR_UNSUPPORTED, "slab does not fit in shared memory%s", ""); }
    // the dynamic shared-memory limit is an attribute of the KERNEL (template instance), not of this handle: opt in to what the
    // largest block of this morphology needs, so that a small env created later cannot lower the limit under a large one
    const size_t smem_max = sb * h->wpb;
    {   // fewer pairs than one wave can hold: spread them over all SMs (fewer warps per block = less issue contention) instead
        // of filling some SMs and leaving others idle.  (With two or more waves an even split measured slower: the block
        // scheduler backfills finished SMs anyway.)
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
        h->sms = sms;
        const long long E = cfg->num_envs;
        if (E <= (long long)sms * h->wpb) { const int even = (int)((E + sms - 1) / sms); if (even >= 1 && even < h->wpb) h->wpb = even; }
    }
    h->smem = sb * h->wpb;
    const int E = cfg->num_envs;
    CUDA_OK(cudaMalloc(&h->d_am, 2 * sizeof(rs_agent_model)));
    CUDA_OK(cudaMemcpy(h->d_am, agents, 2 * sizeof(rs_agent_model), cudaMemcpyHostToDevice));
    EnvDev& d = h->d;
    d.E = E; d.am = h->d_am; d.h = cfg->timestep; d.max_newton = cfg->newton_iters > 0 ? cfg->newton_iters : 16;
    d.P.frame_skip = cfg->frame_skip; d.P.timestep_limit = cfg->timestep_limit; d.P.ring_limit = cfg->ring_limit;
    d.P.init_pos_noise = cfg->init_pos_noise; d.P.init_vel_noise = cfg->init_vel_noise;
    d.P.seed_lo = (uint32_t)cfg->seed; d.P.seed_hi = (uint32_t)(cfg->seed >> 32);
    CUDA_OK(cudaMalloc(&d.qpos, sizeof(float) * E * h->nq)); CUDA_OK(cudaMalloc(&d.qvel, sizeof(float) * E * h->nv));
    const int LTh = h->LA + h->LB;
    const size_t pred_bytes = sizeof(int) * (size_t)E * (4 + (LTh <= 8 ? 24 : (LTh <= 12 ? 32 : 40)) / 2);      // Slab::MAXC
    CUDA_OK(cudaMalloc(&d.pred, pred_bytes)); CUDA_OK(cudaMemset(d.pred, 0, pred_bytes));
    CUDA_OK(cudaMalloc(&d.warm, sizeof(float) * E * h->nv)); CUDA_OK(cudaMalloc(&d.ep_ret, sizeof(float) * E));
    CUDA_OK(cudaMalloc(&d.ep_dret, sizeof(float) * E)); CUDA_OK(cudaMalloc(&d.ep_step, sizeof(int) * E));
    CUDA_OK(cudaMalloc(&d.status, sizeof(int) * E)); CUDA_OK(cudaMalloc(&d.ep_count, sizeof(unsigned int) * E));
    CUDA_OK(cudaMalloc(&d.diag, sizeof(int) * E * 4)); CUDA_OK(cudaMemset(d.diag, 0, sizeof(int) * E * 4));
    CUDA_OK(cudaMalloc(&d.latch, sizeof(int) * 2)); CUDA_OK(cudaMemset(d.latch, 0, sizeof(int) * 2));
    CUDA_OK(cudaMalloc(&d.next, sizeof(int) * 2)); CUDA_OK(cudaMemset(d.next, 0, sizeof(int) * 2)); d.tick = 0;
    CUDA_OK(cudaMemset(d.qpos, 0, sizeof(float) * E * h->nq)); CUDA_OK(cudaMemset(d.qvel, 0, sizeof(float) * E * h->nv));
    CUDA_OK(cudaMemset(d.warm, 0, sizeof(float) * E * h->nv)); CUDA_OK(cudaMemset(d.ep_ret, 0, sizeof(float) * E));
    CUDA_OK(cudaMemset(d.ep_dret, 0, sizeof(float) * E)); CUDA_OK(cudaMemset(d.ep_step, 0, sizeof(int) * E));
    CUDA_OK(cudaMemset(d.status, 0, sizeof(int) * E)); CUDA_OK(cudaMemset(d.ep_count, 0, sizeof(unsigned int) * E));
    int rc = dispatch(h, [&](auto la, auto lb) {
        constexpr int A = decltype(la)::value, B = decltype(lb)::value;
        CUDA_OK(cudaFuncSetAttribute(k_step<A, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        CUDA_OK(cudaFuncSetAttribute(k_reset<A, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        CUDA_OK(cudaFuncSetAttribute(k_set_state<A, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        CUDA_OK(cudaFuncSetAttribute(k_forward_debug<A, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        return RS_OK;
    });
    if (rc) { delete h; return rc; }
    // host staging
    const int OD = h->obsA + h->obsB;
    CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_OK(cudaMallocHost(&h->h_act, sizeof(float) * E * h->nu)); CUDA_OK(cudaMallocHost(&h->h_obs, sizeof(float) * E * OD));
    CUDA_OK(cudaMallocHost(&h->h_rew, sizeof(float) * E * 2)); CUDA_OK(cudaMallocHost(&h->h_info, sizeof(float) * E * 2 * RS_INFO_DIM));
    CUDA_OK(cudaMallocHost(&h->h_epi, sizeof(float) * E * 3)); CUDA_OK(cudaMallocHost(&h->h_done, E * 2));
    CUDA_OK(cudaMalloc(&h->s_act, sizeof(float) * E * h->nu)); CUDA_OK(cudaMalloc(&h->s_obs, sizeof(float) * E * OD));
    CUDA_OK(cudaMalloc(&h->s_rew, sizeof(float) * E * 2)); CUDA_OK(cudaMalloc(&h->s_info, sizeof(float) * E * 2 * RS_INFO_DIM));
    CUDA_OK(cudaMalloc(&h->s_epi, sizeof(float) * E * 3)); CUDA_OK(cudaMalloc(&h->s_done, E * 2));
    *out = h;
    return RS_OK;
}

void rs_destroy(rs_env* h) {
    if (!h) return;
    cudaFree(h->d_am); cudaFree(h->d.qpos); cudaFree(h->d.qvel); cudaFree(h->d.warm); cudaFree(h->d.pred); cudaFree(h->d.ep_ret);
    cudaFree(h->d.ep_dret); cudaFree(h->d.ep_step); cudaFree(h->d.diag); cudaFree(h->d.latch); cudaFree(h->d.next); cudaFree(h->d.status); cudaFree(h->d.ep_count);
    cudaFreeHost(h->h_act); cudaFreeHost(h->h_obs); cudaFreeHost(h->h_rew); cudaFreeHost(h->h_info); cudaFreeHost(h->h_epi); cudaFreeHost(h->h_done);
    cudaFree(h->s_act); cudaFree(h->s_obs); cudaFree(h->s_rew); cudaFree(h->s_info); cudaFree(h->s_epi); cudaFree(h->s_done);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int rs_dims(const rs_env* h, int* nq, int* nv, int* nu, int* obs_a, int* obs_b, int* act_a, int* act_b) {
    if (!h) return fail(RS_ERR_ARG, "rs_dims: null handle%s", "");
    if (nq) *nq = h->nq; if (nv) *nv = h->nv; if (nu) *nu = h->nu;
    if (obs_a) *obs_a = h->obsA; if (obs_b) *obs_b = h->obsB;
    if (act_a) *act_a = 2 * h->LA; if (act_b) *act_b = 2 * h->LB;
    return RS_OK;
}

#define GRID(h) dim3(((h)->d.E + (h)->wpb - 1) / (h)->wpb), dim3(32 * (h)->wpb), (h)->smem, (cudaStream_t)stream

int rs_reset(rs_env* h, const uint8_t* mask, float* obs, void* stream) {
    if (!h) return fail(RS_ERR_ARG, "rs_reset: null handle%s", "");
    return dispatch(h, [&](auto la, auto lb) {
        k_reset<decltype(la)::value, decltype(lb)::value><<<GRID(h)>>>(h->d, mask, obs);
        g_launches++;
        CUDA_OK(cudaGetLastError());
        return RS_OK;
    });
}

int rs_set_state(rs_env* h, const float* qpos, const float* qvel, float* obs, void* stream) {
    if (!h || !qpos || !qvel) return fail(RS_ERR_ARG, "rs_set_state: bad argument%s", "");
    return dispatch(h, [&](auto la, auto lb) {
        k_set_state<decltype(la)::value, decltype(lb)::value><<<GRID(h)>>>(h->d, qpos, qvel, obs);
        g_launches++;
        CUDA_OK(cudaGetLastError());
        return RS_OK;
    });
}

int rs_get_state(rs_env* h, float* qpos, float* qvel, int* ep_step, int* status, void* stream) {
    if (!h) return fail(RS_ERR_ARG, "rs_get_state: null handle%s", "");
    cudaStream_t st = (cudaStream_t)stream;
    const int E = h->d.E;
    if (qpos) CUDA_OK(cudaMemcpyAsync(qpos, h->d.qpos, sizeof(float) * E * h->nq, cudaMemcpyDeviceToDevice, st));
    if (qvel) CUDA_OK(cudaMemcpyAsync(qvel, h->d.qvel, sizeof(float) * E * h->nv, cudaMemcpyDeviceToDevice, st));
    if (ep_step) CUDA_OK(cudaMemcpyAsync(ep_step, h->d.ep_step, sizeof(int) * E, cudaMemcpyDeviceToDevice, st));
    if (status) CUDA_OK(cudaMemcpyAsync(status, h->d.status, sizeof(int) * E, cudaMemcpyDeviceToDevice, st));
    return RS_OK;
}

int rs_step(rs_env* h, const float* actions, float* obs, float* rew, uint8_t* done, float* info, float* episode,
            int auto_reset, void* stream) {
    if (!h || !actions || !obs || !rew || !done) return fail(RS_ERR_ARG, "rs_step: bad argument%s", "");
    return dispatch(h, [&](auto la, auto lb) {
        constexpr int A = decltype(la)::value, B = decltype(lb)::value;
        if (RS_TRIP_MACHINE && A + B <= RS_TRIP_MAX_LEGS) {      // persistent blocks, pairs handed out by a device counter
            int grid = (h->d.E + h->wpb - 1) / h->wpb;
            h->d.persistent = grid > h->sms ? 1 : 0;
            if (grid > h->sms) grid = h->sms;
            k_step<A, B><<<grid, 32 * h->wpb, h->smem, (cudaStream_t)stream>>>(h->d, actions, obs, rew, done, info, episode, auto_reset);
            h->d.tick ^= 1;
        } else
            k_step<A, B><<<GRID(h)>>>(h->d, actions, obs, rew, done, info, episode, auto_reset);
        g_launches++;
        CUDA_OK(cudaGetLastError());
        return RS_OK;
    });
}

// true when `p` is page-locked host memory (cudaHostAlloc / cudaHostRegister / torch pin_memory): the copy engine can use it directly
static bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

int rs_step_host(rs_env* h, const float* actions, float* obs, float* rew, uint8_t* done, float* info, float* episode,
                 int auto_reset) {
    if (!h || !actions || !obs || !rew || !done) return fail(RS_ERR_ARG, "rs_step_host: bad argument%s", "");
    const int E = h->d.E, OD = h->obsA + h->obsB;
    // caller buffers that are page-locked are used as they are; pageable ones go through the handle's pinned staging
    const float* a_src = actions;
    if (!is_pinned(actions)) { memcpy(h->h_act, actions, sizeof(float) * E * h->nu); a_src = h->h_act; }
    CUDA_OK(cudaMemcpyAsync(h->s_act, a_src, sizeof(float) * E * h->nu, cudaMemcpyHostToDevice, h->stream));
    int rc = rs_step(h, h->s_act, h->s_obs, h->s_rew, h->s_done, h->s_info, h->s_epi, auto_reset, h->stream);
    if (rc) return rc;
    struct Out { void* user; void* stage; const void* dev; size_t bytes; bool direct; };
    Out outs[5] = {
        { obs, h->h_obs, h->s_obs, sizeof(float) * E * OD, false },
        { rew, h->h_rew, h->s_rew, sizeof(float) * E * 2, false },
        { done, h->h_done, h->s_done, (size_t)E * 2, false },
        { info, h->h_info, h->s_info, sizeof(float) * E * 2 * RS_INFO_DIM, false },
        { episode, h->h_epi, h->s_epi, sizeof(float) * E * 3, false } };
    for (Out& o : outs) {
        if (!o.user) continue;
        o.direct = is_pinned(o.user);
        CUDA_OK(cudaMemcpyAsync(o.direct ? o.user : o.stage, o.dev, o.bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_OK(cudaStreamSynchronize(h->stream));
    for (Out& o : outs) if (o.user && !o.direct) memcpy(o.user, o.stage, o.bytes);
    return RS_OK;
}

int rs_get_diag(rs_env* h, int* diag, void* stream) {
    if (!h || !diag) return fail(RS_ERR_ARG, "rs_get_diag: bad argument%s", "");
    CUDA_OK(cudaMemcpyAsync(diag, h->d.diag, sizeof(int) * h->d.E * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return RS_OK;
}

/* OR of the status bits raised by any env since the last clear (out[0]) and how many env-steps raised one (out[1]); HOST
 * pointer, synchronises with `stream`.  Auto-reset wipes the per-env status word, not this latch: the caller checks it once per
 * rollout, where mujoco-py's warning callback would have raised MujocoException inside the worker (builder.py:351-369). */
int rs_status_latch(rs_env* h, int* out2_host, int clear, void* stream) {
    if (!h || !out2_host) return fail(RS_ERR_ARG, "rs_status_latch: bad argument%s", "");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_OK(cudaMemcpyAsync(out2_host, h->d.latch, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
    if (clear) CUDA_OK(cudaMemsetAsync(h->d.latch, 0, sizeof(int) * 2, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return RS_OK;
}
/* env.seed(s) of the reference (run.py:73-83: env i is seeded seed + i; mujoco_env.py:82-84): re-keys the Philox streams that
 * draw the reset states; env e keeps its own stream (seed, e).  Takes effect at the next reset / auto-reset. */
int rs_seed(rs_env* h, unsigned long long seed) {
    if (!h) return fail(RS_ERR_ARG, "rs_seed: null handle%s", "");
    h->cfg.seed = seed;
    h->d.P.seed_lo = (uint32_t)seed; h->d.P.seed_hi = (uint32_t)(seed >> 32);
    return RS_OK;
}

int rs_forward_debug(rs_env* h, const float* ctrl, float* qacc, int* ncon, int* niter, void* stream) {
    if (!h || !ctrl || !qacc || !ncon || !niter) return fail(RS_ERR_ARG, "rs_forward_debug: bad argument%s", "");
    return dispatch(h, [&](auto la, auto lb) {
        k_forward_debug<decltype(la)::value, decltype(lb)::value><<<GRID(h)>>>(h->d, ctrl, qacc, ncon, niter);
        g_launches++;
        CUDA_OK(cudaGetLastError());
        return RS_OK;
    });
}


// ------------------------------------------------------------------------------------------
// learner-side entry points (rs_learn.cuh)
// ------------------------------------------------------------------------------------------
static int ensure_smem(const void* fn, size_t bytes) {
    static std::atomic<size_t> cur_fwd(0), cur_ppo(0);
    std::atomic<size_t>& cur = (fn == (const void*)rsl::k_mlp_forward) ? cur_fwd : cur_ppo;
    if (bytes > cur.load()) {
        CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cur.store(bytes);
    }
    return RS_OK;
}

int rs_param_count(int obs_dim, int act_dim) { return rsl::make_layout(obs_dim, act_dim).P; }

static int mlp_forward_jobs(const rsl::MlpJobs& J, int njobs, int obs_dim, int act_dim, int n, int precision, void* stream) {
    if (precision == 1 && rsl::tc_tile_bytes(obs_dim) > 227 * 1024) precision = 0;     // wide observations: FP32-pipe kernel
    dim3 grid((n + RSL_TILE - 1) / RSL_TILE, njobs);
    if (precision == 1) {
        size_t sm = rsl::tc_tile_bytes(obs_dim);
        static std::atomic<size_t> cur(0);
        if (sm > cur.load()) { CUDA_OK(cudaFuncSetAttribute(rsl::k_mlp_forward_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); cur.store(sm); }
        rsl::k_mlp_forward_tc<<<grid, RSL_TC_THREADS, sm, (cudaStream_t)stream>>>(J, obs_dim, act_dim, n);
    } else {
        size_t sm = rsl::tile_bytes(obs_dim, act_dim);
        if (sm > 227 * 1024) return fail(RS_ERR_UNSUPPORTED, "rs_mlp_forward: obs_dim too large for one tile%s", "");
        int rc = ensure_smem((const void*)rsl::k_mlp_forward, sm); if (rc) return rc;
        rsl::k_mlp_forward<<<grid, RSL_TILE, sm, (cudaStream_t)stream>>>(J, obs_dim, act_dim, n);
    }
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return RS_OK;
}

int rs_mlp_forward(const float* params, int obs_dim, int act_dim, const float* obs, long long ld, int n, float* mean, float* value,
                   int precision, void* stream) {
    if (!params || !obs || n <= 0 || act_dim > RSL_HW || (!mean && !value)) return fail(RS_ERR_ARG, "rs_mlp_forward: bad argument%s", "");
    rsl::MlpJobs J;
    memset(&J, 0, sizeof(J));
    J.params[0] = params; J.X[0] = obs; J.ldx[0] = (size_t)ld; J.mean[0] = mean; J.value[0] = value;
    return mlp_forward_jobs(J, 1, obs_dim, act_dim, n, precision, stream);
}

int rs_mlp_forward_multi(const rs_mlp_job* jobs, int njobs, int obs_dim, int act_dim, int n, int precision, void* stream) {
    if (!jobs || njobs < 1 || njobs > 4 || n <= 0 || act_dim > RSL_HW) return fail(RS_ERR_ARG, "rs_mlp_forward_multi: bad argument%s", "");
    rsl::MlpJobs J;
    memset(&J, 0, sizeof(J));
    for (int i = 0; i < njobs; i++) {
        if (!jobs[i].params || !jobs[i].obs || (!jobs[i].mean && !jobs[i].value)) return fail(RS_ERR_ARG, "rs_mlp_forward_multi: bad job%s", "");
        J.params[i] = jobs[i].params; J.X[i] = jobs[i].obs; J.ldx[i] = (size_t)jobs[i].obs_row_stride; J.mean[i] = jobs[i].mean; J.value[i] = jobs[i].value; J.act[i] = jobs[i].activation;
        if (jobs[i].activation) precision = 0;      // tanh nets (policy_zoo opponents) run on the FP32-pipe kernel
    }
    return mlp_forward_jobs(J, njobs, obs_dim, act_dim, n, precision, stream);
}

int rs_rollout_sample(int E, int act_dim, const float* logstd0, const float* logstd1, const float* mu00, const float* mu10,
                      const float* mu11, const float* mu01, unsigned long long seed, unsigned int tick, int deterministic,
                      float* actions, float* nlp0, float* nlp1, float* opp_nlp0, float* opp_nlp1, void* stream) {
    if (E <= 0 || act_dim > 16 || !actions) return fail(RS_ERR_ARG, "rs_rollout_sample: bad argument%s", "");
    rsl::k_rollout_sample<<<(E + 127) / 128, 128, 0, (cudaStream_t)stream>>>(E, act_dim, logstd0, logstd1, mu00, mu10, mu11, mu01, seed, tick,
                                                                             deterministic, actions, nlp0, nlp1, opp_nlp0, opp_nlp1);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return RS_OK;
}

int rs_rollout(rs_env* h, int T, const rs_rollout_io* io, int precision, unsigned long long seed, unsigned int tick0, int deterministic,
               void* stream) {
    if (!h || !io || T <= 0 || !io->params0 || !io->params1 || !io->obs || !io->rew || !io->done || !io->info || !io->episode || !io->mb_obs ||
        !io->mb_actions || !io->mb_values || !io->mb_nlp || !io->mb_opp_nlp || !io->mb_dones || !io->mb_shaping || !io->mb_main || !io->ep_done ||
        !io->ep_info || !io->scratch)
        return fail(RS_ERR_ARG, "rs_rollout: bad argument%s", "");
    if (h->obsA != h->obsB || h->LA != h->LB) return fail(RS_ERR_UNSUPPORTED, "rs_rollout: the two agents must share one observation / action layout%s", "");
    const int E = h->d.E, D = h->obsA, A = h->nu / 2;
    const rsl::Layout L = rsl::make_layout(D, A);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t TE = (size_t)T * E;
    float* mu00 = io->scratch; float* mu10 = mu00 + (size_t)E * A; float* mu11 = mu10 + (size_t)E * A; float* mu01 = mu11 + (size_t)E * A;
    for (int t = 0; t < T; t++) {
        rsl::MlpJobs J;
        memset(&J, 0, sizeof(J));
        // models[0].step(obs[:,0]) | models[1].action_probability(obs[:,0], a0) | models[1].step(obs[:,1]) | models[0].value + action_probability(obs[:,1], a1)
        const float* prm[4] = { io->params0, io->params1, io->params1, io->params0 };
        float* means[4] = { mu00, mu10, mu11, mu01 };
        float* vals[4] = { io->mb_values + (size_t)t * E, nullptr, nullptr, io->mb_values + TE + (size_t)t * E };
        for (int j = 0; j < 4; j++) { J.params[j] = prm[j]; J.X[j] = io->obs + (j >= 2 ? D : 0); J.ldx[j] = (size_t)2 * D; J.mean[j] = means[j]; J.value[j] = vals[j]; J.act[j] = 0; }
        int rc = mlp_forward_jobs(J, 4, D, A, E, precision, stream);
        if (rc) return rc;
        const long long n_obs = (long long)E * 2 * D;
        rsl::k_traj_pre<<<(unsigned)((n_obs + 255) / 256), 256, 0, st>>>(E, D, t, T, io->obs, io->done, io->mb_obs, io->mb_dones);
        g_launches++;
        float* act = io->mb_actions + (size_t)t * E * 2 * A;
        rc = rs_rollout_sample(E, A, io->params0 + L.logstd, io->params1 + L.logstd, mu00, mu10, mu11, mu01, seed, tick0 + (unsigned)t, deterministic, act,
                               io->mb_nlp + (size_t)t * E, io->mb_nlp + TE + (size_t)t * E, io->mb_opp_nlp + (size_t)t * E, io->mb_opp_nlp + TE + (size_t)t * E, stream);
        if (rc) return rc;
        rc = rs_step(h, act, io->obs, io->rew, io->done, io->info, io->episode, 1, stream);
        if (rc) return rc;
        rsl::k_traj_post<<<(E + 127) / 128, 128, 0, st>>>(E, t, T, io->info, io->done, io->episode, io->mb_shaping, io->mb_main, io->ep_done, io->ep_info);
        g_launches++;
    }
    CUDA_OK(cudaGetLastError());
    return RS_OK;
}

int rs_neglogp(int n, int act_dim, const float* act, const float* mu, const float* logstd, float* out, void* stream) {
    if (n <= 0 || act_dim > 16) return fail(RS_ERR_ARG, "rs_neglogp: bad argument%s", "");
    rsl::k_neglogp<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, act_dim, act, mu, logstd, out);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return RS_OK;
}

int rs_vtrace(int T, int E, float gamma, float lam, float rho_bar, float c_bar, const float* rewards, const float* values,
              const uint8_t* dones, const float* nlp, const float* opp_nlp, const float* last_values, const uint8_t* last_dones,
              float* returns, float* ratios, void* stream) {
    if (T <= 0 || E <= 0 || !returns) return fail(RS_ERR_ARG, "rs_vtrace: bad argument%s", "");
    rsl::k_vtrace<<<(2 * E + 127) / 128, 128, 0, (cudaStream_t)stream>>>(T, E, gamma, lam, rho_bar, c_bar, rewards, values, dones, nlp, opp_nlp,
                                                                        last_values, last_dones, returns, ratios);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return RS_OK;
}

int rs_adv_moments(const int* idx, int n, const float* returns, const float* values, double* sums, void* stream) {
    if (n < 0 || !sums) return fail(RS_ERR_ARG, "rs_adv_moments: bad argument%s", "");
    CUDA_OK(cudaMemsetAsync(sums, 0, 2 * sizeof(double), (cudaStream_t)stream));
    if (n > 0) {
        int blocks = (n + 255) / 256; if (blocks > 296) blocks = 296;
        rsl::k_adv_moments<<<blocks, 256, 0, (cudaStream_t)stream>>>(idx, n, returns, values, sums);
        g_launches++;
        CUDA_OK(cudaGetLastError());
    }
    return RS_OK;
}

long long rs_ppo_workspace_floats(int obs_dim, int act_dim, int max_minibatch) {
    long long nb = (max_minibatch + RSL_TILE - 1) / RSL_TILE;
    return nb * (rsl::make_layout(obs_dim, act_dim).P + 8) + 16;
}

/* local part of one PPO minibatch: grad_stats[P + 4] = sum over the n local samples of d loss / d theta (already divided by
 * global_n) followed by the 4 stat sums (pg, vf, approxkl, clipfrac).  adv_sums[2] are the (global) advantage moments. */
int rs_ppo_grad(const float* params, int obs_dim, int act_dim, const float* obs, const float* actions, const float* returns,
                const float* values, const float* old_nlp, const float* weights, const int* idx, int n, long long global_n,
                const double* adv_sums, float cliprange, float ent_coef, float vf_coef, float* workspace, float* grad_stats,
                float* log_ratio, double* stats5, int precision, void* stream) {
    if (!params || !obs || !grad_stats || !workspace || act_dim > RSL_HW || n < 0 || global_n <= 0) return fail(RS_ERR_ARG, "rs_ppo_grad: bad argument%s", "");
    const rsl::Layout L = rsl::make_layout(obs_dim, act_dim);
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {      // this rank holds no sample of the minibatch: zero contribution (the entropy still comes from the parameters)
        rsl::k_grad_reduce2<<<(L.P + 255) / 256, 256, 0, st>>>(workspace, workspace, 0, L.P, grad_stats, params, L.logstd, act_dim, 0.f, stats5);
        g_launches++;
        CUDA_OK(cudaGetLastError());
        return RS_OK;
    }
    if (precision == 1 && rsl::tc_tile_bytes(obs_dim) > 227 * 1024) precision = 0;     // wide observations: FP32-pipe kernel
    size_t sm = precision == 1 ? rsl::tc_tile_bytes(obs_dim) : rsl::tile_bytes(obs_dim, act_dim);
    if (sm > 227 * 1024) return fail(RS_ERR_UNSUPPORTED, "rs_ppo_grad: obs_dim too large for one tile%s", "");
    if (precision == 1) {
        static std::atomic<size_t> cur(0);
        if (sm > cur.load()) { CUDA_OK(cudaFuncSetAttribute(rsl::k_ppo_tile_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); cur.store(sm); }
    } else {
        int rc = ensure_smem((const void*)rsl::k_ppo_tile, sm); if (rc) return rc;
    }
    const int nb = (n + RSL_TILE - 1) / RSL_TILE;
    rsl::PPOArgs a;
    a.params = params; a.D = obs_dim; a.A = act_dim; a.obs = obs; a.actions = actions; a.returns = returns; a.values = values;
    a.old_nlp = old_nlp; a.weights = weights; a.idx = idx; a.n = n; a.adv_sums = adv_sums; a.adv_count = (double)global_n;
    a.cliprange = cliprange; a.ent_coef = ent_coef; a.vf_coef = vf_coef; a.inv_n = 1.0f / (float)global_n;
    a.gpart = workspace; a.spart = workspace + (size_t)nb * L.P; a.log_ratio = log_ratio;
    if (precision == 1) rsl::k_ppo_tile_tc<<<nb, RSL_TC_THREADS, sm, st>>>(a);
    else rsl::k_ppo_tile<<<nb, RSL_TILE, sm, st>>>(a);
    rsl::k_grad_reduce2<<<(L.P + 255) / 256, 256, 0, st>>>(a.gpart, a.spart, nb, L.P, grad_stats, params, L.logstd, act_dim,
                                                          (float)((double)ent_coef * (double)n / (double)global_n), stats5);
    g_launches += 2;
    CUDA_OK(cudaGetLastError());
    return RS_OK;
}

/* global-norm clip and TF-style Adam on the (all-reduced) gradient (entropy term already in it, see rs_ppo_grad), ONE launch; also finalises the statistics
 * [pg_loss, vf_loss, ., approxkl, clipfrac] from the stat sums behind the gradient when stats5 != NULL.  step_t counts from 1. */
int rs_adam_step(float* params, float* m, float* v, float* grad, int obs_dim, int act_dim, float ent_coef, float max_grad_norm,
                 float lr, long long step_t, float beta1, float beta2, float eps, float* gnorm_out, long long global_n, double* stats5,
                 void* stream) {
    if (!params || !m || !v || !grad || step_t < 1 || (stats5 && global_n < 1)) return fail(RS_ERR_ARG, "rs_adam_step: bad argument%s", "");
    const rsl::Layout L = rsl::make_layout(obs_dim, act_dim);
    cudaStream_t st = (cudaStream_t)stream;
    const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)step_t)) / (1.0 - pow((double)beta1, (double)step_t));
    rsl::k_adam2<<<(L.P + 255) / 256, 256, 0, st>>>(params, m, v, grad, L.P, max_grad_norm, (float)lr_t, beta1, beta2, eps,
                                                    gnorm_out, stats5 ? 1.0 / (double)global_n : 0.0, stats5);
    g_launches += 1;
    CUDA_OK(cudaGetLastError());
    return RS_OK;
}

int rs_ppo_stats(const float* grad_stats, const float* params, int obs_dim, int act_dim, long long global_n, double* stats5, void* stream) {
    if (!grad_stats || !params || !stats5 || global_n < 1) return fail(RS_ERR_ARG, "rs_ppo_stats: bad argument%s", "");
    const rsl::Layout L = rsl::make_layout(obs_dim, act_dim);
    rsl::k_ppo_stats<<<1, 32, 0, (cudaStream_t)stream>>>(grad_stats, params, L.P, L.logstd, act_dim, 1.0 / (double)global_n, stats5);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return RS_OK;
}

int rs_ppo_minibatch_step(float* params, float* m, float* v, int obs_dim, int act_dim, const float* obs, const float* actions,
                          const float* returns, const float* values, const float* old_nlp, const float* weights, const int* idx, int n,
                          float cliprange, float ent_coef, float vf_coef, float max_grad_norm, float lr, long long step_t,
                          float* workspace, float* grad_stats, double* adv_sums, int moments_ready, float* gnorm_out, double* stats5,
                          float* log_ratio, int precision, void* stream) {
    int rc = RS_OK;
    if (!moments_ready) { rc = rs_adv_moments(idx, n, returns, values, adv_sums, stream); if (rc) return rc; }
    rc = rs_ppo_grad(params, obs_dim, act_dim, obs, actions, returns, values, old_nlp, weights, idx, n, n, adv_sums, cliprange, ent_coef,
                     vf_coef, workspace, grad_stats, log_ratio, stats5, precision, stream);
    if (rc) return rc;
    return rs_adam_step(params, m, v, grad_stats, obs_dim, act_dim, ent_coef, max_grad_norm, lr, step_t, 0.9f, 0.999f, 1e-5f, gnorm_out, n, stats5, stream);
}

}  // extern "C"

// ---- gradient all-reduce over NVLink peer memory ---------------------------------------------------------------------------
#define RS_PEER_MAX 16
struct rs_peer {
    int rank, world, device;
    long long nfloats, step;
    float* buf;                     // [2][nfloats] this rank's data, double-buffered by the parity of the step
    int* flags;                     // [RS_PEER_MAX] step counters written by the peers, [RS_PEER_MAX] = error latch
    float* peer_buf[RS_PEER_MAX];   // the same buffers of every rank (own pointer for the own rank)
    int* peer_flags[RS_PEER_MAX];
    bool opened[RS_PEER_MAX];
};
struct PeerArgs { const float* buf[RS_PEER_MAX]; int* flags[RS_PEER_MAX]; int* my_flags; int rank, world; int step; long long n, off; };

__global__ void __launch_bounds__(256) k_peer_allreduce(PeerArgs a, float* __restrict__ out) {
    // (1) this rank's data is complete (the kernels that wrote it precede this one on the stream): tell every peer
    if (blockIdx.x == 0 && threadIdx.x < a.world) {
        __threadfence_system();
        *((volatile int*)(a.flags[threadIdx.x] + a.rank)) = a.step;
    }
    // (2) wait for every peer's data of this step (the counters are monotonic, so a peer that is already a step ahead passes too)
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < a.world) {
        const volatile int* f = (const volatile int*)(a.my_flags + threadIdx.x);
        const long long t0 = clock64();
        if (*((const volatile int*)(a.my_flags + RS_PEER_MAX)) != 0) ok = 0;      // an earlier wait already failed: the replicas are out of step, do not spin again
        else while (*f < a.step) {
            __nanosleep(64);
            if (clock64() - t0 > (1ll << 34)) { ok = 0; atomicExch(a.my_flags + RS_PEER_MAX, a.step); break; }      // ~9 s: give up, latch
        }
    }
    __syncthreads();
    __threadfence_system();
    if (!ok) return;
    // (3) sum over the ranks in rank order; peer memory is read around the caches (it changes under this GPU's feet)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) {
        float acc = 0.f;
        for (int r = 0; r < a.world; r++) acc += __ldcv(a.buf[r] + a.off + i);
        out[i] = acc;
    }
}

extern "C" {
int rs_peer_handle_bytes(void) { return 2 * (int)sizeof(cudaIpcMemHandle_t); }
int rs_peer_create(int rank, int world, long long nfloats, int device, rs_peer** out) {
    if (!out || world < 1 || world > RS_PEER_MAX || rank < 0 || rank >= world || nfloats <= 0) return fail(RS_ERR_ARG, "rs_peer_create: bad argument%s", "");
    CUDA_OK(cudaSetDevice(device));
    rs_peer* p = new rs_peer();
    memset(p, 0, sizeof(*p));
    p->rank = rank; p->world = world; p->device = device; p->nfloats = nfloats; p->step = 0;
    CUDA_OK(cudaMalloc(&p->buf, sizeof(float) * 2 * nfloats)); CUDA_OK(cudaMemset(p->buf, 0, sizeof(float) * 2 * nfloats));
    CUDA_OK(cudaMalloc(&p->flags, sizeof(int) * (RS_PEER_MAX + 1))); CUDA_OK(cudaMemset(p->flags, 0, sizeof(int) * (RS_PEER_MAX + 1)));
    p->peer_buf[rank] = p->buf; p->peer_flags[rank] = p->flags;
    *out = p;
    return RS_OK;
}
int rs_peer_export(rs_peer* p, void* handle_out) {
    if (!p || !handle_out) return fail(RS_ERR_ARG, "rs_peer_export: bad argument%s", "");
    cudaIpcMemHandle_t h[2];
    CUDA_OK(cudaIpcGetMemHandle(&h[0], p->buf)); CUDA_OK(cudaIpcGetMemHandle(&h[1], p->flags));
    memcpy(handle_out, h, sizeof(h));
    return RS_OK;
}
int rs_peer_connect(rs_peer* p, const void* all_handles) {
    if (!p || !all_handles) return fail(RS_ERR_ARG, "rs_peer_connect: bad argument%s", "");
    CUDA_OK(cudaSetDevice(p->device));
    for (int r = 0; r < p->world; r++) {
        if (r == p->rank) continue;
        cudaIpcMemHandle_t h[2];
        memcpy(h, (const char*)all_handles + (size_t)r * sizeof(h), sizeof(h));
        CUDA_OK(cudaIpcOpenMemHandle((void**)&p->peer_buf[r], h[0], cudaIpcMemLazyEnablePeerAccess));
        CUDA_OK(cudaIpcOpenMemHandle((void**)&p->peer_flags[r], h[1], cudaIpcMemLazyEnablePeerAccess));
        p->opened[r] = true;
    }
    return RS_OK;
}
float* rs_peer_send_buffer(rs_peer* p) { return p ? p->buf + ((p->step + 1) & 1) * p->nfloats : nullptr; }
int rs_peer_allreduce(rs_peer* p, float* out, long long nfloats, void* stream) {
    if (!p || !out || nfloats <= 0 || nfloats > p->nfloats) return fail(RS_ERR_ARG, "rs_peer_allreduce: bad argument%s", "");
    p->step++;
    PeerArgs a;
    for (int r = 0; r < p->world; r++) { a.buf[r] = p->peer_buf[r]; a.flags[r] = p->peer_flags[r]; }
    a.my_flags = p->flags; a.rank = p->rank; a.world = p->world; a.step = (int)p->step; a.n = nfloats; a.off = (p->step & 1) * p->nfloats;
    k_peer_allreduce<<<(unsigned)((nfloats + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, out);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return RS_OK;
}
int rs_peer_error(rs_peer* p) {
    if (!p) return 0;
    int e = 0;
    if (cudaMemcpy(&e, p->flags + RS_PEER_MAX, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return e;
}
void rs_peer_destroy(rs_peer* p) {
    if (!p) return;
    cudaDeviceSynchronize();
    for (int r = 0; r < p->world; r++) if (p->opened[r]) { cudaIpcCloseMemHandle(p->peer_buf[r]); cudaIpcCloseMemHandle(p->peer_flags[r]); }
    cudaFree(p->buf); cudaFree(p->flags);
    delete p;
}
}  // extern "C" (peer)

extern "C" {
/* data-parallel minibatch step in ONE call (4 launches, no host work in between): local gradient into this rank's peer buffer ->
 * rs_peer_allreduce -> clip + Adam + statistics.  adv_sums holds the GLOBAL advantage moments of the minibatch (dist.EpochSchedule). */
int rs_ppo_minibatch_step_peer(rs_peer* peer, float* params, float* m, float* v, int obs_dim, int act_dim, const float* obs, const float* actions,
                               const float* returns, const float* values, const float* old_nlp, const float* weights, const int* idx, int n,
                               long long global_n, float cliprange, float ent_coef, float vf_coef, float max_grad_norm, float lr, long long step_t,
                               float* workspace, float* grad_stats, const double* adv_sums, float* gnorm_out, double* stats5, float* log_ratio,
                               int precision, void* stream) {
    if (!peer) return fail(RS_ERR_ARG, "rs_ppo_minibatch_step_peer: null peer%s", "");
    const rsl::Layout L = rsl::make_layout(obs_dim, act_dim);
    int rc = rs_ppo_grad(params, obs_dim, act_dim, obs, actions, returns, values, old_nlp, weights, idx, n, global_n, adv_sums, cliprange, ent_coef,
                         vf_coef, workspace, rs_peer_send_buffer(peer), log_ratio, stats5, precision, stream);
    if (rc) return rc;
    rc = rs_peer_allreduce(peer, grad_stats, L.P + 4, stream);
    if (rc) return rc;
    return rs_adam_step(params, m, v, grad_stats, obs_dim, act_dim, ent_coef, max_grad_norm, lr, step_t, 0.9f, 0.999f, 1e-5f, gnorm_out, global_n, stats5, stream);
}
}  // extern "C"
extern "C" {
/* data-parallel minibatch schedule: local index lists of every minibatch of an epoch from the global permutation (device) */
int rs_epoch_split(const int* perm, long long n_global, int nbatch_train, long long lo, long long hi, int* out_idx, int* counts, void* stream) {
    if (!perm || !out_idx || !counts || n_global <= 0 || nbatch_train <= 0 || hi < lo) return fail(RS_ERR_ARG, "rs_epoch_split: bad argument%s", "");
    const int nmb = (int)((n_global + nbatch_train - 1) / nbatch_train);
    rsl::k_epoch_split<<<nmb, 1024, 0, (cudaStream_t)stream>>>(perm, n_global, nbatch_train, lo, hi, out_idx, counts);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return RS_OK;
}
/* advantage moments of all minibatches of an epoch: sums[nmb][2]; idx [nmb][cap] or NULL (identity), counts [nmb] or NULL (full slices of n_total) */
int rs_adv_moments_multi(const int* idx, const int* counts, int nmb, int cap, long long n_total, const float* returns, const float* values,
                         double* sums, void* stream) {
    if (nmb <= 0 || cap <= 0 || !returns || !values || !sums) return fail(RS_ERR_ARG, "rs_adv_moments_multi: bad argument%s", "");
    rsl::k_adv_moments_multi<<<nmb, 1024, 0, (cudaStream_t)stream>>>(idx, counts, cap, n_total, returns, values, sums);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return RS_OK;
}

/* host-side data-parallel cut of a global permutation (int64, as np.random.shuffle leaves it): the entries inside [lo, hi) as local
 * int32 indices, in order, concatenated over the minibatches, and how many each minibatch of nbatch_train entries holds */
}  // extern "C"
template <typename T>
static int epoch_split_host_core(const T* perm, long long n_global, int nbatch_train, long long lo, long long hi, int* local, int* counts) {
    long long w = 0;
    int m = 0;
    const long long cap = hi - lo;
    for (long long s0 = 0; s0 < n_global; s0 += nbatch_train, m++) {
        const long long s1 = s0 + nbatch_train < n_global ? s0 + nbatch_train : n_global;
        const long long w0 = w;
        if (w + (s1 - s0) <= cap) {      // no overflow possible inside this minibatch: tight branch-free loop
            for (long long i = s0; i < s1; i++) { const long long v = (long long)perm[i]; local[w] = (int)(v - lo); w += (v >= lo) & (v < hi); }
        } else {
            for (long long i = s0; i < s1; i++) {
                const long long v = (long long)perm[i];
                if (v >= lo && v < hi) { if (w >= cap) return fail(RS_ERR_ARG, "rs_epoch_split_host: not a permutation%s", ""); local[w++] = (int)(v - lo); }
            }
        }
        counts[m] = (int)(w - w0);
    }
    return RS_OK;
}
extern "C" {
int rs_epoch_split_host(const void* perm, int elem_bytes, long long n_global, int nbatch_train, long long lo, long long hi, int* local, int* counts) {
    if (!perm || !local || !counts || n_global <= 0 || nbatch_train <= 0 || hi < lo || (elem_bytes != 4 && elem_bytes != 8)) return fail(RS_ERR_ARG, "rs_epoch_split_host: bad argument%s", "");
    return elem_bytes == 4 ? epoch_split_host_core((const int32_t*)perm, n_global, nbatch_train, lo, hi, local, counts)
                           : epoch_split_host_core((const int64_t*)perm, n_global, nbatch_train, lo, hi, local, counts);
}

// ---- legacy NumPy shuffle replay (host) -----------------------------------------------------------------------------------
static inline void mt19937_refill(uint32_t* mt) {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, A = 0x9908b0dfu;
    int kk;
    for (kk = 0; kk < 624 - 397; kk++) { uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
    for (; kk < 623; kk++) { uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u); }
    uint32_t y = (mt[623] & UPPER) | (mt[0] & LOWER);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
}
// tempered outputs of the current 624-word block, produced by one vectorisable loop per refill
struct MtStream {
    uint32_t* key; int pos; uint32_t out[624];
    void temper(int from) {
        for (int i = from; i < 624; i++) {
            uint32_t y = key[i];
            y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
            out[i] = y;
        }
    }
    inline uint32_t next() {
        if (pos >= 624) { mt19937_refill(key); pos = 0; temper(0); }
        return out[pos++];
    }
};
}  // extern "C"
template <typename T>
static int legacy_shuffle_core(uint32_t* key, int* pos, T* x, long long n) {
    if (!key || !pos || !x || n < 0 || *pos < 0 || *pos > 624) return fail(RS_ERR_ARG, "rs_legacy_shuffle: bad argument%s", "");
    if (n > 0x100000000LL) return fail(RS_ERR_UNSUPPORTED, "rs_legacy_shuffle: n beyond the 32-bit interval path%s", "");
    MtStream g; g.key = key; g.pos = *pos;
    g.temper(g.pos < 624 ? g.pos : 624);
    // the swap partners depend on the generator only, not on the data: draw them a block ahead and prefetch, so that the random
    // accesses into x (4-64 MB) overlap instead of paying one cache miss per element
    const int B = 64;
    long long jj[B];
    uint64_t mask = 0;                                       // smallest 2^k - 1 >= i; it only shrinks as i counts down
    if (n > 1) { mask = (uint64_t)(n - 1); mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32; }
    for (long long i = n - 1; i >= 1; i -= B) {
        const int m = (int)(i >= B ? B : i);                 // partners for i, i-1, ..., i-m+1
        for (int q = 0; q < m; q++) {
            const uint64_t top = (uint64_t)(i - q);
            if (top <= (mask >> 1)) mask >>= 1;
            uint64_t j;
            do { j = (uint64_t)g.next() & mask; } while (j > top);
            jj[q] = (long long)j;
            __builtin_prefetch(x + j, 1, 0);
        }
        for (int q = 0; q < m; q++) { const long long a = i - q, b = jj[q]; const T tmp = x[a]; x[a] = x[b]; x[b] = tmp; }
    }
    *pos = g.pos;
    return RS_OK;
}
extern "C" {
int rs_legacy_shuffle(uint32_t* key, int* pos, int64_t* x, long long n) { return legacy_shuffle_core<int64_t>(key, pos, x, n); }
/* the same permutation on an int32 array (n < 2^31): half the bytes under the random accesses, which is what the routine waits for */
int rs_legacy_shuffle32(uint32_t* key, int* pos, int32_t* x, long long n) {
    if (n >= 0x7fffffffLL) return fail(RS_ERR_UNSUPPORTED, "rs_legacy_shuffle32: n does not fit int32%s", "");
    return legacy_shuffle_core<int32_t>(key, pos, x, n);
}

/* tcgen05 descriptor/layout self-test: D[128,64] = op(A) * op(B) through kind::tf32 UMMA (see rs_tc.cuh) */
int rs_tc_selftest(const int* prm13, const float* A, const float* B, float* D, void* stream) {
    if (!prm13 || !A || !B || !D) return fail(RS_ERR_ARG, "rs_tc_selftest: bad argument%s", "");
    rstc::SelfTestParams P;
    for (int i = 0; i < 13; i++) P.v[i] = prm13[i];
    if (P.v[0] * P.v[1] > 128 * 128 || P.v[2] * P.v[3] > 128 * 64) return fail(RS_ERR_ARG, "rs_tc_selftest: tile too large%s", "");
    const size_t sm = sizeof(float) * (128 * 128 + 128 * 64);
    static bool attr = false;
    if (!attr) { CUDA_OK(cudaFuncSetAttribute(rstc::k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); attr = true; }
    rstc::k_tc_selftest<<<1, 128, sm, (cudaStream_t)stream>>>(P, A, B, D);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return RS_OK;
}

}  // extern "C"
