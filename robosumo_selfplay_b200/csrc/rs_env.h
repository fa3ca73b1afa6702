// rs_env.h -- fused environment logic around the physics of rs_core.h, one warp per env pair:
// action clipping, SumoEnv._step rewards / win / lose / draw flags (robosumo/robosumo/envs/
// sumo.py:120-202), Agent.get_obs assembly (agents.py:190-214), the wrapper's episode
// bookkeeping and timestep feature (sumo_env.py:40-72), reset_model's distribution
// (sumo.py:232-253) with Philox instead of the per-process NumPy RandomState, and the worker's
// auto-reset (subproc_vec_env.py:12-16).
#pragma once
#include "rs_core.h"

namespace rs {

// ---- Philox4x32-10 (counter-based RNG; env e, episode n, draw i -> independent block) ----
RS_HD void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
RS_HD float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0,1)

struct EnvParams {
    int frame_skip, timestep_limit;
    float ring_limit, init_pos_noise, init_vel_noise;
    uint32_t seed_lo, seed_hi;
};

// reset_model: writes s.q, s.v (quaternions normalised as mj_forward would), zero warm start
template <int LA, int LB>
RS_HD void env_reset_state(Ctx<LA, LB>& c, const EnvParams& P, uint32_t env_id, uint32_t episode) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    uint32_t o[4];
    philox4x32(env_id, episode, 0xFFFFFFFFu, 0u, P.seed_lo, P.seed_hi, o);
    const float phi = 6.283185307179586f * u01(o[0]);
    RS_LANE_LOOP(i, S::NQ) {
        int a = i >= c.qadr(1) ? 1 : 0, k = i - c.qadr(a);
        float base = 0.f;
        float ang = phi + (a ? 3.14159265358979f : 0.f);
        if (k == 0) base = 1.15f * cosf(ang); else if (k == 1) base = 1.15f * sinf(ang); else if (k == 2) base = 1.25f;
        else if (k == 3) base = 1.f;
        uint32_t r[4];
        philox4x32(env_id, episode, (uint32_t)i, 1u, P.seed_lo, P.seed_hi, r);
        s.q[i] = base + P.init_pos_noise * (2.f * u01(r[0]) - 1.f);
    }
    RS_LANE_LOOP(i, S::NV) {
        uint32_t r[4];
        philox4x32(env_id, episode, (uint32_t)i, 2u, P.seed_lo, P.seed_hi, r);
        float u1 = u01(r[0]), u2 = u01(r[1]);
        s.v[i] = P.init_vel_noise * sqrtf(-2.f * logf(u1)) * cosf(6.283185307179586f * u2);
        s.x[i] = 0.f;
    }
    RS_SYNC();
    RS_LANE_LOOP(a, 2) {
        float* qq = s.q + c.qadr(a);
        float n = sqrtf(qq[3]*qq[3] + qq[4]*qq[4] + qq[5]*qq[5] + qq[6]*qq[6]);
        float inv = n > 1e-12f ? 1.f / n : 1.f;
        qq[3] *= inv; qq[4] *= inv; qq[5] *= inv; qq[6] *= inv;
    }
    RS_SYNC();
}

// Agent.get_obs for both agents -> obs[obsA + obsB]; tsfeat = value of the last element
template <int LA, int LB>
RS_HD void env_write_obs(Ctx<LA, LB>& c, float* obs, float tsfeat) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    const int dimA = (7 + 2*LA) + (6 + 2*LA) + 6 * (1 + 3*LA) + 14;
    const int dimB = (7 + 2*LB) + (6 + 2*LB) + 6 * (1 + 3*LB) + 14;
    RS_LANE_LOOP(i, dimA + dimB) {
        int a = i >= dimA ? 1 : 0, k = a ? i - dimA : i;
        int nq = 7 + 2 * c.L(a), nv = 6 + 2 * c.L(a), nf = 6 * (1 + 3 * c.L(a));
        float val = 0.f;
        if (k < nq) { val = s.q[c.qadr(a) + k]; if (k == 2) val += c.am[a].adjust_z; }
        else if (k < nq + nv) val = s.v[c.vadr(a) + k - nq];
        else if (k < nq + nv + nf) val = 0.f;                 // cfrc_ext: never computed by MuJoCo >= 2.0 here
        else if (k < nq + nv + nf + 7) { int j = k - nq - nv - nf; val = s.q[c.qadr(1 - a) + j]; if (j == 2) val += c.am[1 - a].adjust_z; }
        else if (k < nq + nv + nf + 13) val = 0.f;
        else val = tsfeat;
        obs[i] = val;
    }
}

RS_HD bool out_of_ring(float x, float y, float z, float ring) {
    return (z < 0.29f) || (fmaxf(fabsf(x), fabsf(y)) >= ring);
}

// SumoEnv._step reward block + wrapper bookkeeping.  Executed by every lane redundantly on
// uniform data; results returned in registers.  before[4] = (x0,y0,x1,y1) before the step.
struct StepOut { float rew[2]; int done[2]; float info[2][RS_INFO_DIM]; };
template <int LA, int LB>
RS_HD void env_rewards(const Ctx<LA, LB>& c, const EnvParams& P, const float* before, const float* actions,
                       int num_steps /* already incremented */, float dt, StepOut* o) {
    typedef Slab<LA, LB> S;
    const S& s = *c.s;
    float px[2], py[2], pz[2];
    for (int a = 0; a < 2; a++) { px[a] = s.q[c.qadr(a)]; py[a] = s.q[c.qadr(a) + 1]; pz[a] = s.q[c.qadr(a) + 2] + c.am[a].adjust_z; }
    bool out[2] = { out_of_ring(px[0], py[0], pz[0], P.ring_limit), out_of_ring(px[1], py[1], pz[1], P.ring_limit) };
    bool timeout = num_steps > P.timestep_limit;
    for (int a = 0; a < 2; a++) {
        int opp = 1 - a;
        int na = 2 * c.L(a), u0 = a ? 2 * LA : 0;
        float ss = 0.f;
        for (int k = 0; k < na; k++) ss += actions[u0 + k] * actions[u0 + k];
        float ctrl = -0.1f * ss;
        float lose = out[a] ? -2000.f : 0.f, win = out[opp] ? 2000.f : 0.f;
        float main_r = win + lose + (timeout ? -1000.f : 0.f);
        float mvx = (px[a] - before[2*a]) / dt, mvy = (py[a] - before[2*a + 1]) / dt;
        float dx = px[opp] - before[2*a], dy = py[opp] - before[2*a + 1];
        float dn = sqrtf(dx*dx + dy*dy);
        float move = fmaxf((mvx * dx + mvy * dy) / dn, 0.f) * 0.1f;
        float push = -10.f * expf(-sqrtf(px[opp]*px[opp] + py[opp]*py[opp]));
        float shaping = ctrl + push + move;
        o->rew[a] = main_r + shaping;
        o->done[a] = (out[a] || out[opp] || timeout) ? 1 : 0;
        o->info[a][0] = ctrl; o->info[a][1] = lose; o->info[a][2] = win; o->info[a][3] = main_r;
        o->info[a][4] = move; o->info[a][5] = push; o->info[a][6] = shaping;
        o->info[a][7] = (out[opp] ? 1.f : 0.f);
    }
    // sumo_env.py:62-65: timeout flag for both agents when agent 0's main reward is exactly the draw penalty
    if (o->done[0] && o->info[0][3] == -1000.f) { o->info[0][7] += 2.f; o->info[1][7] += 2.f; }
}

}  // namespace rs
