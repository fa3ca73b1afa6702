// Host emulation of the kernel source (rs_core.h / rs_env.h) for CPU unit tests.
// TEST HARNESS ONLY: compiled by tests/emu/build.py into tests/_emu/libemu.so and loaded only by
// tests.  The product package never loads it -- its ops require the CUDA library.
#include <string.h>
#include <stdlib.h>
#include "../../robosumo_selfplay_b200/csrc/rs_env.h"

using namespace rs;

template <int LA, int LB>
static int run_forward(const rs_agent_model* am, float h, int max_newton, const float* q, const float* v, const float* ctrl,
                       float* qacc, float* Mout, float* tau, int* ncon, int* niter, float* con, float* qnorm) {
    typedef Slab<LA, LB> S;
    S* s = (S*)calloc(1, sizeof(S));
    Ctx<LA, LB> c; c.s = s; c.am = am; c.h = h; c.max_newton = max_newton;
    memcpy(s->q, q, sizeof(float) * S::NQ); memcpy(s->v, v, sizeof(float) * S::NV);
    for (int a = 0; a < 2; a++) for (int k = 0; k < 2 * c.L(a); k++) {
        int u = (a ? 2 * LA : 0) + k; float x = ctrl[u]; x = x < -1.f ? -1.f : (x > 1.f ? 1.f : x); s->act[u] = am[a].gear * x;
    }
    forward(c);
    memcpy(qacc, s->x, sizeof(float) * S::NV);
    if (Mout) dense_M(c, Mout);
    if (tau) memset(tau, 0, sizeof(float) * S::NV);
    if (qnorm) memcpy(qnorm, s->q, sizeof(float) * S::NQ);
    *ncon = s->ncon; *niter = s->niter;
    if (con) for (int k = 0; k < s->ncon; k++) { con[8*k] = 0.f; memcpy(con + 8*k + 1, s->cpos[k], 12); memcpy(con + 8*k + 4, s->cn[k], 12); con[8*k+7] = (float)(s->bA(k) * 100 + s->bB(k)); }
    int st = s->status; free(s); return st;
}

template <int LA, int LB>
static int run_step(const rs_agent_model* am, float h, int max_newton, float* q, float* v, float* warm, const float* ctrl, int nsub) {
    typedef Slab<LA, LB> S;
    S* s = (S*)calloc(1, sizeof(S));
    Ctx<LA, LB> c; c.s = s; c.am = am; c.h = h; c.max_newton = max_newton;
    memcpy(s->q, q, sizeof(float) * S::NQ); memcpy(s->v, v, sizeof(float) * S::NV); memcpy(s->x, warm, sizeof(float) * S::NV);
    for (int a = 0; a < 2; a++) for (int k = 0; k < 2 * c.L(a); k++) {
        int u = (a ? 2 * LA : 0) + k; float x = ctrl[u]; x = x < -1.f ? -1.f : (x > 1.f ? 1.f : x); s->act[u] = am[a].gear * x;
    }
    simulate(c, nsub);
    memcpy(q, s->q, sizeof(float) * S::NQ); memcpy(v, s->v, sizeof(float) * S::NV); memcpy(warm, s->x, sizeof(float) * S::NV);
    int st = s->status; free(s); return st;
}

extern "C" {
int emu_forward(const rs_agent_model* am, float h, int max_newton, const float* q, const float* v, const float* ctrl,
                float* qacc, float* Mout, float* tau, int* ncon, int* niter, float* con, float* qnorm) {
    int LA = am[0].L, LB = am[1].L;
    if (LA == 4 && LB == 4) return run_forward<4, 4>(am, h, max_newton, q, v, ctrl, qacc, Mout, tau, ncon, niter, con, qnorm);
    if (LA == 6 && LB == 6) return run_forward<6, 6>(am, h, max_newton, q, v, ctrl, qacc, Mout, tau, ncon, niter, con, qnorm);
    if (LA == 8 && LB == 8) return run_forward<8, 8>(am, h, max_newton, q, v, ctrl, qacc, Mout, tau, ncon, niter, con, qnorm);
    if (LA == 4 && LB == 6) return run_forward<4, 6>(am, h, max_newton, q, v, ctrl, qacc, Mout, tau, ncon, niter, con, qnorm);
    if (LA == 8 && LB == 4) return run_forward<8, 4>(am, h, max_newton, q, v, ctrl, qacc, Mout, tau, ncon, niter, con, qnorm);
    return -1;
}
int emu_step(const rs_agent_model* am, float h, int max_newton, float* q, float* v, float* warm, const float* ctrl, int nsub) {
    int LA = am[0].L, LB = am[1].L;
    if (LA == 4 && LB == 4) return run_step<4, 4>(am, h, max_newton, q, v, warm, ctrl, nsub);
    if (LA == 6 && LB == 6) return run_step<6, 6>(am, h, max_newton, q, v, warm, ctrl, nsub);
    if (LA == 8 && LB == 8) return run_step<8, 8>(am, h, max_newton, q, v, warm, ctrl, nsub);
    if (LA == 4 && LB == 6) return run_step<4, 6>(am, h, max_newton, q, v, warm, ctrl, nsub);
    if (LA == 8 && LB == 4) return run_step<8, 4>(am, h, max_newton, q, v, warm, ctrl, nsub);
    return -1;
}
int emu_slab_bytes(int LA, int LB) {
    if (LA == 4 && LB == 4) return (int)sizeof(Slab<4, 4>);
    if (LA == 6 && LB == 6) return (int)sizeof(Slab<6, 6>);
    if (LA == 8 && LB == 8) return (int)sizeof(Slab<8, 8>);
    return -1;
}
}
