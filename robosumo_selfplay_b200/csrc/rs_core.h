// rs_core.h -- articulated-body + contact physics of one RoboSumo env pair, one warp per pair.
//
// Replaces what the reference gets from MuJoCo 2.1 through mujoco-py for this path
// (robosumo/robosumo/envs/mujoco_env.py:104-129 -> mujoco-py/mujoco_py/mjsim.pyx:101-129;
//  stage list mujoco-py/mujoco_py/pxd/mujoco.pxd:208-327; options assets/tatami.xml:3-6).
//
// Design (B200-first, not a MuJoCo port):
//  * one warp owns one env pair; all per-pair state lives in a shared-memory slab for the
//    whole env step (5 substeps x 4 RK stages = 20 forward evaluations), HBM is touched only
//    at the start (state + action) and the end (state + obs + reward);
//  * the body tree is exploited structurally: a floating torso group (torso + welded leg
//    stubs, pre-merged at model-compile time) with L two-link legs -> closed-form spatial
//    recursion per leg, no generic tree walk;
//  * the constraint Jacobian is never stored: J*x is evaluated through body twists and
//    J^T*f through body wrenches; H = M + J^T D J is assembled contact by contact from three
//    direction Jacobians;
//  * the soft-constraint problem is solved in the primal (Newton with exact line search on
//    the one-sided quadratic rows of the pyramidal cone), which is the reference's configured
//    solver, warm-started from the previous stage; the Newton system keeps M's arrowhead form
//    (arrow_solve) and contacts between the two agents enter as a low-rank correction of it
//    (woodbury_solve);
//  * the warps of a block re-align once per "trip" (evaluation start + one Newton iteration,
//    simulate_trips) so that they share the instruction stream without waiting for each
//    other's extra iterations; with more pairs than warp slots the blocks are persistent and
//    a warp takes its next pair off a device counter (rs_api.cu).
//
// Code is written as phases of independent "items" (RS_LANE_LOOP) separated by warp syncs
// and communicating only through the slab, so that the same source also compiles as a
// host emulation for the CPU unit tests (tests/emu).  The product builds the CUDA path only.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/rs_b200.h"

#if defined(__CUDACC__)
#define RS_HD __host__ __device__ __forceinline__
#else
#define RS_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define RS_LANE_LOOP(i, n) for (int i = (int)(threadIdx.x & 31); i < (n); i += 32)
#define RS_SYNC() __syncwarp()
#define RS_ATOMIC_ADDF(p, v) atomicAdd((p), (v))
#define RS_ATOMIC_INC(p) atomicAdd((p), 1)
#define RS_ATOMIC_OR(p, v) atomicOr((p), (v))
#define RS_LANE0 ((threadIdx.x & 31) == 0)
#ifndef RS_RCP_EXACT
#define RS_RCP(x) rs_rcp_approx(x)      /* pivot reciprocals of the linear solves: MUFU.RCP (1 ulp) instead of the 12-instruction IEEE sequence on the dependent chain of every pivot */
#else
#define RS_RCP(x) __frcp_rn(x)
#endif
#ifdef RS_RSQRT_EXACT
#define RS_RSQRT(x) (1.0f / sqrtf(x))
#else
#define RS_RSQRT(x) rs_rsqrt_nr(x)      /* MUFU.RSQ + one Newton step (<= 1 ulp) instead of IEEE sqrt followed by IEEE divide (~20 dependent instructions per normalisation) */
#endif
#ifdef RS_SINCOS_EXACT
#define RS_SINCOS(x, s, c) sincosf((x), (s), (c))
#else
#define RS_SINCOS(x, s, c) rs_sincos((x), (s), (c))      /* joint angles and half rotation angles, |x| < ~100: two-step reduction + fdlibm kernels, no slow path */
#endif
#ifdef RS_DIV_EXACT
#define RS_DIV(a, b) ((a) / (b))
#else
#define RS_DIV(a, b) ((a) * rs_rcp_nr(b))                 /* MUFU.RCP + one Newton step (<= 1 ulp) instead of the IEEE division sequence */
#endif
#define RS_UNROLL1 _Pragma("unroll 1")
#define RS_COLD __device__ __forceinline__
#define RS_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#define RS_LIKELY(x) __builtin_expect(!!(x), 1)
#define RS_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define RS_LANE_LOOP(i, n) for (int i = 0; i < (n); i++)
#define RS_SYNC()
#define RS_ATOMIC_ADDF(p, v) (*(p) += (v))
#define RS_ATOMIC_INC(p) ((*(p))++)
#define RS_ATOMIC_OR(p, v) (*(p) |= (v))
#define RS_LANE0 (true)
#define RS_RCP(x) (1.0f / (x))
#define RS_RSQRT(x) (1.0f / sqrtf(x))
#define RS_SINCOS(x, s, c) sincosf((x), (s), (c))
#define RS_DIV(a, b) ((a) / (b))
#define RS_UNROLL1
#define RS_COLD static inline
#define RS_WARP_ANY(p) (p)
#define RS_LIKELY(x) (x)
#define RS_UNLIKELY(x) (x)
#endif


// scene constants (assets/tatami.xml, utils.py:64-88)
#define RS_FLOOR_Z (-0.025f)
#define RS_BOX_HX 2.3f
#define RS_BOX_HZ 0.25f
#define RS_BOX_CZ 0.25f
#define RS_RAIL 2.0f
#define RS_RAIL_Z 0.5f
#define RS_RAIL_R 0.03f
#define RS_MARGIN 0.01f
#define RS_MU 1.0f
#define RS_GRAV 9.81f
// solref (0.02, 1), solimp (0.9, 0.95, 0.001, 0.5, 2)  [MuJoCo defaults]
#define RS_DMIN 0.9f
#define RS_DMAX 0.95f
#define RS_WIDTH 0.001f

namespace rs {

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float rs_rsqrt_nr(float x) {
    float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    const float e = fmaf(-x * y, y, 1.0f);       // 1 - x y^2
    return fmaf(0.5f * y, e, y);
}
__device__ __forceinline__ float rs_rcp_nr(float x) {
    float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return fmaf(y, fmaf(-x, y, 1.0f), y);
}
// sin and cos together for moderate arguments (|x| < ~100, where k * pi/2 in two pieces is exact enough): ~30 instructions, <= 1-2 ulp
__device__ __forceinline__ void rs_sincos(float x, float* sn, float* cs) {
    const float k = rintf(x * 0.636619772f);
    float r = fmaf(k, -1.57079637e+0f, x);
    r = fmaf(k, 4.37113883e-8f, r);
    const float z = r * r;
    const float sp = fmaf(z, fmaf(z, fmaf(z, 2.7557314297e-06f, -1.9841270114e-04f), 8.3333337680e-03f), -1.6666667163e-01f);
    const float sr = fmaf(r * z, sp, r);
    const float cp = fmaf(z, fmaf(z, fmaf(z, -2.7557314297e-07f, 2.4801587642e-05f), -1.3888889225e-03f), 4.1666667908e-02f);
    const float cr = fmaf(z * z, cp, fmaf(-0.5f, z, 1.0f));
    const int q = (int)k & 3;
    const float s0 = (q & 1) ? cr : sr, c0 = (q & 1) ? sr : cr;
    *sn = (q & 2) ? -s0 : s0;
    *cs = ((q + 1) & 2) ? -c0 : c0;
}
__device__ __forceinline__ float rs_rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif

struct V3 { float x, y, z; };
RS_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
RS_HD V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
RS_HD void st3(float* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
RS_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
RS_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
RS_HD V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
RS_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RS_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
RS_HD float norm(V3 a) { return sqrtf(dot(a, a)); }
RS_HD V3 normalized(V3 a, float* len) {
    float d2 = dot(a, a);
    if (d2 < 1e-24f) { *len = 0.f; return v3(1.f, 0.f, 0.f); }
    float s = RS_RSQRT(d2);
    *len = d2 * s;
    return s * a;
}
// R (row-major 3x3) * v  and  R^T * v
RS_HD V3 mulR(const float* R, V3 v) { return v3(R[0]*v.x + R[1]*v.y + R[2]*v.z, R[3]*v.x + R[4]*v.y + R[5]*v.z, R[6]*v.x + R[7]*v.y + R[8]*v.z); }
RS_HD V3 mulRT(const float* R, V3 v) { return v3(R[0]*v.x + R[3]*v.y + R[6]*v.z, R[1]*v.x + R[4]*v.y + R[7]*v.z, R[2]*v.x + R[5]*v.y + R[8]*v.z); }
// Rodrigues rotation of v about unit axis a by angle (s = sin, c = cos)
RS_HD V3 rot(V3 a, float s, float c, V3 v) { return c * v + s * cross(a, v) + ((1.f - c) * dot(a, v)) * a; }

template <int LA, int LB>
struct Slab {
    enum {
        LT = LA + LB, NQA = 7 + 2 * LA, NVA = 6 + 2 * LA, NVB = 6 + 2 * LB, NQ = 14 + 2 * LT, NV = 12 + 2 * LT, NU = 2 * LT,
        NVP = (NV | 1), NB = 2 + 2 * LT, NG = 2 + 3 * LT, NGA = 1 + 3 * LA, NGB = 1 + 3 * LB,
        MAXC = (LT <= 8 ? 24 : (LT <= 12 ? 32 : 40)),     // contact capacity per pair (24 Ant pairs, 32 with a Bug, 40 with a Spider: its 3-sigma-action soak filled 32)
        // H storage: two per-agent blocks (row stride n+1) when the agents are uncoupled, one dense NV x NVP matrix when an
        // inter-agent contact couples them.  The dense form spills over the arrays that are dead between build_H and the
        // end of the linear solve (tw .. ljd below), so only HDED floats are dedicated to it.
        BSA = NVA * (NVA + 1), BSB = NVB * (NVB + 1), HBD = BSA + BSB, HFULL = NV * NVP,
        ALIAS = NB * 12 + LT * 6 + 2 * NV + MAXC * 4 + NU,
        HDED = (HFULL - ALIAS > HBD ? HFULL - ALIAS : HBD)
    };
    float q[NQ], v[NV], q0[NQ], v0[NV], vsum[NV], asum[NV];
    float x[NV];      // qacc: Newton iterate, warm start across stages / steps
    float r[NV];      // M x - qfrc_smooth  (dynamics() leaves -qfrc_smooth here, solve() adds M x)
    float d[NV];      // Newton direction
    float act[NU];    // gear * clip(ctrl)
    float Rt[2][9];   // torso rotation matrices
    float org[NB][3]; // body origins: [a] torso a, [2+g] hip of leg g, [2+LT+g] ankle of leg g
    float tip[LT][3];
    float axa[LT][3];               // ankle axes in world (hip axes are fixed in the torso: recomputed as R_t * ax_hip)
    // joint-space inertia, arrowhead form: root 6x6 per agent, root-leg coupling 6x2 and leg 2x2 (hh, ha, aa) per leg
    float Mr[2][36], Mc[LT][12], Ml[LT][3];
    float H[HDED];
    // ---- alias zone (order matters: the dense H extends over it) ----
    float tw[NB][6];                // body twists (omega, v at body origin)
    float wr[NB][6];                // body wrenches (torque about body origin, force)
    float legF[LT][6];              // leg subtree (torque, force)
    float Md[NV];
    float jtf[NV];
    float cjd[MAXC][4];
    float ljd[NU];
    // ---- contacts: position, normal, first tangent (second = n x t1), D, rows ----
    float cpos[MAXC][3], cn[MAXC][3], ct1[MAXC][3], cD[MAXC];
    float caref[MAXC][4], cjar[MAXC][4];          // (dynamics() borrows these two as link-inertia scratch)
    int cbody[MAXC];                               // (bA + 1) | (bB + 1) << 8 | key << 16 | active rows << 28 ; body -1 = world
    unsigned short cprev[MAXC];                    // contacts of the previous evaluation: key << 4 | final active rows
    int nprev;
    // joint limits (one potential row per hinge)
    float lsgn[NU], lD[NU], laref[NU], ljar[NU];
    float scr[64];                  // per-contact direction Jacobians: idx(16 as float) + 3 x 16
    int ncon, status, niter, same;
    int coupled;                              // bit 0: an inter-agent contact couples the two agents, bit 1: an intra-agent (leg-leg) contact; 0: H keeps M's arrowhead form
    int lmask, pmask, pvalid;                 // limit rows: active set in use / predicted from the previous evaluation / prediction valid
    int tot_iter, tot_coupled, tot_ncon, max_iter;      // diagnostics accumulated over one env step
    int wood_m, wood_k[2];                    // inter-agent contacts with active rows in this iteration (count, first two indices)
    RS_HD int bA(int k) const { return (cbody[k] & 255) - 1; }
    RS_HD int bB(int k) const { return ((cbody[k] >> 8) & 255) - 1; }
    RS_HD int ckey(int k) const { return (cbody[k] >> 16) & 4095; }
    RS_HD int cact(int k) const { return (cbody[k] >> 28) & 15; }
    RS_HD void set_cact(int k, int bits) { cbody[k] = (cbody[k] & 0x0FFFFFFF) | (bits << 28); }
    // index of H(ir, ic); in block-diagonal mode ir and ic belong to the same agent
    RS_HD int hidx(int ir, int ic) const {
        if (coupled & 1) return ir * NVP + ic;
        return ir >= NVA ? BSA + (ir - NVA) * (NVB + 1) + (ic - NVA) : ir * (NVA + 1) + ic;
    }
};

template <int LA, int LB>
struct Ctx {
    typedef Slab<LA, LB> S;
    S* s;
    const rs_agent_model* am;   // [2]
    float h;                    // timestep
    int max_newton;
#ifdef RS_EXPERIMENT_CLOCK
    int env; long long acc[6], tlast;  // developer instrumentation: per-evaluation timestamps, see rs_api.cu
#endif
    int evk;                    // evaluations started by this warp (sliding-window re-alignment, device only)
    RS_HD int L(int a) const { return a ? LB : LA; }
    RS_HD int qadr(int a) const { return a ? S::NQA : 0; }
    RS_HD int vadr(int a) const { return a ? S::NVA : 0; }
    RS_HD int leg0(int a) const { return a ? LA : 0; }
    RS_HD int agent_of_leg(int g) const { return g >= LA ? 1 : 0; }
    RS_HD int hipdof(int g) const { int a = agent_of_leg(g); return vadr(a) + 6 + 2 * (g - leg0(a)); }
    RS_HD int hipq(int g) const { int a = agent_of_leg(g); return qadr(a) + 7 + 2 * (g - leg0(a)); }
    RS_HD V3 hip_axis(int g) const { int a = agent_of_leg(g); return mulR(s->Rt[a], ld3(am[a].ax_hip[g - leg0(a)])); }
    RS_HD int bhip(int g) const { return 2 + g; }
    RS_HD int bank(int g) const { return 2 + S::LT + g; }
};

// ------------------------------------------------------------------------------------------
// kinematics (mj_kinematics): normalises the free-joint quaternions in q in place
// ------------------------------------------------------------------------------------------
template <int LA, int LB>
RS_HD void fk(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    RS_LANE_LOOP(a, 2) {
        float* qq = s.q + c.qadr(a);
        float w = qq[3], x = qq[4], y = qq[5], z = qq[6];
        float n2 = w * w + x * x + y * y + z * z;
        if (n2 < 1e-24f) { w = 1.f; x = y = z = 0.f; } else { float inv = RS_RSQRT(n2); w *= inv; x *= inv; y *= inv; z *= inv; }
        qq[3] = w; qq[4] = x; qq[5] = y; qq[6] = z;
        float* R = s.Rt[a];
        R[0] = w*w + x*x - y*y - z*z; R[1] = 2.f*(x*y - w*z);         R[2] = 2.f*(x*z + w*y);
        R[3] = 2.f*(x*y + w*z);       R[4] = w*w - x*x + y*y - z*z;   R[5] = 2.f*(y*z - w*x);
        R[6] = 2.f*(x*z - w*y);       R[7] = 2.f*(y*z + w*x);         R[8] = w*w - x*x - y*y + z*z;
        s.org[a][0] = qq[0]; s.org[a][1] = qq[1]; s.org[a][2] = qq[2];
    }
    RS_SYNC();
    RS_LANE_LOOP(g, S::LT) {
        int a = c.agent_of_leg(g), l = g - c.leg0(a);
        const rs_agent_model& m = c.am[a];
        const float* R = s.Rt[a];
        V3 pt = ld3(s.org[a]);
        float qh = s.q[c.hipq(g)], qa = s.q[c.hipq(g) + 1];
        float sh, ch, sa, ca;
        RS_SINCOS(qh, &sh, &ch); RS_SINCOS(qa, &sa, &ca);
        V3 axh = ld3(m.ax_hip[l]), axa = ld3(m.ax_ank[l]);
        V3 ph = pt + mulR(R, ld3(m.r_hip[l]));
        V3 pa = ph + mulR(R, rot(axh, sh, ch, ld3(m.r_ank[l])));
        V3 tipl = rot(axh, sh, ch, rot(axa, sa, ca, ld3(m.e_ank[l])));
        st3(s.org[c.bhip(g)], ph);
        st3(s.org[c.bank(g)], pa);
        st3(s.tip[g], pa + mulR(R, tipl));
        st3(s.axa[g], mulR(R, rot(axh, sh, ch, axa)));
    }
    RS_SYNC();
}

// spatial inertia of a capsule body about origin O (world axes), applied to a motion (w, vO):
// returns momentum-like pair (n about O, f)
struct Cap { float m, ip, ia; V3 c, u; };   // mass, perpendicular / axial inertia, com (rel. O), axis
RS_HD void applyI(const Cap& b, V3 w, V3 vO, V3* n, V3* f) {
    V3 vc = vO + cross(w, b.c);
    *f = b.m * vc;
    V3 Iw = b.ip * w + ((b.ia - b.ip) * dot(b.u, w)) * b.u;
    *n = Iw + cross(b.c, *f);
}
// accumulate the 10 spatial-inertia numbers of a capsule about O
RS_HD void accI(const Cap& b, float* I10) {
    I10[0] += b.m;
    I10[1] += b.m * b.c.x; I10[2] += b.m * b.c.y; I10[3] += b.m * b.c.z;
    float cc = dot(b.c, b.c), k = b.ia - b.ip;
    I10[4] += b.ip + k * b.u.x * b.u.x + b.m * (cc - b.c.x * b.c.x);
    I10[5] += k * b.u.x * b.u.y - b.m * b.c.x * b.c.y;
    I10[6] += k * b.u.x * b.u.z - b.m * b.c.x * b.c.z;
    I10[7] += b.ip + k * b.u.y * b.u.y + b.m * (cc - b.c.y * b.c.y);
    I10[8] += k * b.u.y * b.u.z - b.m * b.c.y * b.c.z;
    I10[9] += b.ip + k * b.u.z * b.u.z + b.m * (cc - b.c.z * b.c.z);
}
// body force of RNE: f = I a + v x* (I v)
RS_HD void bodyForce(const Cap& b, V3 w, V3 vO, V3 al, V3 aO, V3* n, V3* f) {
    V3 hn, hf, an, af;
    applyI(b, w, vO, &hn, &hf);
    applyI(b, al, aO, &an, &af);
    *n = an + cross(w, hn) + cross(vO, hf);
    *f = af + cross(w, hf);
}

// ------------------------------------------------------------------------------------------
// joint-space inertia M (mj_crb) and smooth forces tau = passive + actuator - bias (mj_rne,
// mj_passive, mj_fwdActuation).  Reference point for agent a: its torso origin (an inertial
// point coincident with it at this instant), world-aligned axes.
// ------------------------------------------------------------------------------------------
template <int LA, int LB>
RS_HD void dynamics(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    // Lane-parallel formulation: the per-leg work is cut into three equal-shaped items (the same instruction stream on different
    // data, so a warp runs 3 * LT lanes instead of LT): j = 0 the ankle link with the ankle motion s_a, j = 1 the hip link with the
    // hip motion s_h, j = 2 the ankle link with s_h (the cross term of the hip column).  Items 0 / 1 also carry their link's
    // spatial inertia about O and its RNE force.  Results are parked in scratch that is dead during dynamics (the alias zone and
    // caref | cjar) and combined by a second pass (lane = joint) and a summation pass (lane = agent x component).
    float* X = &s.tw[0][0];                 // [LT][3][6]  I * s      (n about O, f)
    float* F = X + 18 * S::LT;              // [LT][2][6]  RNE force of the link (n about O, f): [0] ankle link, [1] hip link
    float* LI = &s.caref[0][0];             // [LT][2][10] spatial inertia of the link about O
    static_assert(30 * S::LT <= S::ALIAS, "dynamics scratch must fit in the alias zone");
    static_assert(20 * S::LT <= 8 * S::MAXC, "link inertia scratch must fit in caref + cjar");
    RS_LANE_LOOP(i, 3 * S::LT) {
        const int g = i / 3, j = i - 3 * g;
        const int a = c.agent_of_leg(g), l = g - c.leg0(a);
        const rs_agent_model& m = c.am[a];
        const float* R = s.Rt[a];
        const V3 O = ld3(s.org[a]), ph = ld3(s.org[c.bhip(g)]) - O, pa = ld3(s.org[c.bank(g)]) - O, pt = ld3(s.tip[g]) - O;
        const V3 axh = c.hip_axis(g), axa = ld3(s.axa[g]);
        const bool hipl = (j == 1);          // this item's link
        float len;
        Cap b;
        b.m = hipl ? m.m_hip[l] : m.m_ank[l]; b.ip = hipl ? m.ip_hip[l] : m.ip_ank[l]; b.ia = hipl ? m.ia_hip[l] : m.ia_ank[l];
        const V3 p0 = hipl ? ph : pa, p1 = hipl ? pa : pt;
        b.c = 0.5f * (p0 + p1); b.u = normalized(p1 - p0, &len);
        // motion subspaces about O: rotation about axis through point p  ->  (axis, axis x (O - p)) = (axis, -axis x p)
        const V3 sa_v = cross(pa, axa), sh_v = cross(ph, axh);
        V3 n, f;
        applyI(b, j == 0 ? axa : axh, j == 0 ? sa_v : sh_v, &n, &f);
        st3(X + 18 * g + 6 * j, n); st3(X + 18 * g + 6 * j + 3, f);
        if (j < 2) {
            float I10[10];
            for (int k = 0; k < 10; k++) I10[k] = 0.f;
            accI(b, I10);
            for (int k = 0; k < 10; k++) LI[20 * g + 10 * j + k] = I10[k];
            // ---- RNE with qacc = 0 ----
            const float* vv = s.v + c.vadr(a);
            const int dh = c.hipdof(g);
            const V3 wt = mulR(R, v3(vv[3], vv[4], vv[5])), vt = v3(vv[0], vv[1], vv[2]);
            V3 at_lin = cross(vt, wt);  at_lin.z += RS_GRAV;        // spatial accel of torso: (0, -w x v - g)
            const float qdh = s.v[dh], qda = hipl ? 0.f : s.v[dh + 1];
            // hip body: v = v_t + s_h qd ; a = a_t + (v_t x s_h) qd
            const V3 wh = wt + qdh * axh, vh = vt + qdh * sh_v;
            const V3 alh = qdh * cross(wt, axh);
            const V3 ah = at_lin + qdh * (cross(wt, sh_v) + cross(vt, axh));
            // ankle body: a = a_h + (v_h x s_a) qd   (qda = 0 for the hip item: the same expressions return the hip body's motion)
            const V3 wa = wh + qda * axa, vaO = vh + qda * sa_v;
            const V3 ala = alh + qda * cross(wh, axa);
            const V3 aa = ah + qda * (cross(wh, sa_v) + cross(vh, axa));
            V3 nF, fF;
            bodyForce(b, wa, vaO, ala, aa, &nF, &fF);
            st3(F + 12 * g + 6 * j, nF); st3(F + 12 * g + 6 * j + 3, fF);
        }
    }
    RS_SYNC();
    // lane = joint (leg g, col 0 = hip, 1 = ankle): the joint's column of M and its bias force
    RS_LANE_LOOP(i, 2 * S::LT) {
        const int g = i >> 1, col = i & 1;
        const int a = c.agent_of_leg(g), l = g - c.leg0(a);
        const rs_agent_model& m = c.am[a];
        const float* R = s.Rt[a];
        const V3 O = ld3(s.org[a]), ph = ld3(s.org[c.bhip(g)]) - O, pa = ld3(s.org[c.bank(g)]) - O;
        const V3 axh = c.hip_axis(g), axa = ld3(s.axa[g]);
        const V3 sa_v = cross(pa, axa), sh_v = cross(ph, axh);
        const float* x = X + 18 * g; const float* fo = F + 12 * g;
        // column: ankle = I_ank s_a ; hip = (I_hip + I_ank) s_h.  bias: ankle = its link's force ; hip = the leg subtree's
        V3 n = ld3(x + (col ? 0 : 6)), f = ld3(x + (col ? 3 : 9)), nF = ld3(fo), fF = ld3(fo + 3);
        if (!col) { n = n + ld3(x + 12); f = f + ld3(x + 15); nF = nF + ld3(fo + 6); fF = fF + ld3(fo + 9); }
        const V3 ax = col ? axa : axh, sv = col ? sa_v : sh_v;
        s.Ml[g][2 * col] = dot(ax, n) + dot(sv, f) + m.armature;                 // Mhh / Maa
        if (col) s.Ml[g][1] = dot(axh, n) + dot(sh_v, f);                        // Mha
        const V3 rr = mulRT(R, n);
        const float colv[6] = { f.x, f.y, f.z, rr.x, rr.y, rr.z };
        for (int k = 0; k < 6; k++) s.Mc[g][2 * k + col] = colv[k];
        const int dof = c.hipdof(g) + col, ua = (a ? 2 * LA : 0) + 2 * l + col;
        s.r[dof] = m.damping * s.v[dof] - s.act[ua] + dot(ax, nF) + dot(sv, fF);      // -qfrc_smooth
    }
    // lane = agent x component: sums over the agent's legs of the link inertias (10) and the subtree forces (6) -> s.scr[a][16]
    RS_LANE_LOOP(i, 32) {
        const int a = i >> 4, k = i & 15;
        float acc = 0.f;
        RS_UNROLL1
        for (int l = 0; l < c.L(a); l++) {
            const int g = c.leg0(a) + l;
            acc += k < 10 ? LI[20 * g + k] + LI[20 * g + 10 + k] : F[12 * g + (k - 10)] + F[12 * g + 6 + (k - 10)];
        }
        s.scr[i] = acc;
    }
    RS_SYNC();
    // root blocks and root bias, lane = agent x body axis k: column k of the lin-ang and ang-ang blocks (the torso group is rigid, so
    // its own part of the ang-ang block is the constant body-frame inertia about the torso origin; only the legs' part is rotated),
    // then the torso group's RNE (computed by the three lanes of an agent alike, stored by the first)
    RS_LANE_LOOP(i, 6) {
        const int a = i >= 3 ? 1 : 0, k = i - 3 * a;
        const rs_agent_model& m = c.am[a];
        const float* R = s.Rt[a];
        const int va = c.vadr(a);
        const float* sum = s.scr + 16 * a;                       // legs: m, m c (3), I about O (6), subtree torque (3), force (3)
        const V3 cL = ld3(m.cT), cT = mulR(R, cL);               // torso-group com, body / world axes
        const float mt = m.mT + sum[0];
        const V3 mc = m.mT * cT + ld3(sum + 1);
        float* Mr = s.Mr[a];
        for (int j = 0; j < 3; j++) Mr[k * 6 + j] = (j == k) ? mt : 0.f;
        const V3 ek = v3(R[k], R[3 + k], R[6 + k]);
        const V3 f = cross(ek, mc);
        const V3 nl = v3(sum[4] * ek.x + sum[5] * ek.y + sum[6] * ek.z, sum[5] * ek.x + sum[7] * ek.y + sum[8] * ek.z, sum[6] * ek.x + sum[8] * ek.y + sum[9] * ek.z);
        const float ck = k == 0 ? cL.x : (k == 1 ? cL.y : cL.z), cc = dot(cL, cL);
        V3 nb = v3(m.IT[k], m.IT[3 + k], m.IT[6 + k]) + m.mT * (v3(k == 0 ? cc : 0.f, k == 1 ? cc : 0.f, k == 2 ? cc : 0.f) - ck * cL) + mulRT(R, nl);
        Mr[0 * 6 + 3 + k] = f.x; Mr[(3 + k) * 6 + 0] = f.x;
        Mr[1 * 6 + 3 + k] = f.y; Mr[(3 + k) * 6 + 1] = f.y;
        Mr[2 * 6 + 3 + k] = f.z; Mr[(3 + k) * 6 + 2] = f.z;
        Mr[3 * 6 + 3 + k] = nb.x; Mr[4 * 6 + 3 + k] = nb.y; Mr[5 * 6 + 3 + k] = nb.z;
        // torso group RNE (generic full-inertia body force; I_world w = R (IT w_body))
        const float* vv = s.v + va;
        const V3 wb = v3(vv[3], vv[4], vv[5]);
        const V3 wt = mulR(R, wb), vt = v3(vv[0], vv[1], vv[2]);
        V3 at_lin = cross(vt, wt);  at_lin.z += RS_GRAV;
        const V3 vc = vt + cross(wt, cT);
        const V3 hf = m.mT * vc;
        const V3 hn = mulR(R, mulR(m.IT, wb)) + cross(cT, hf);
        const V3 af = m.mT * at_lin;               // alpha = 0
        const V3 an = cross(cT, af);
        const V3 nT = an + cross(wt, hn) + cross(vt, hf) + ld3(sum + 10);
        const V3 fT = af + cross(wt, hf) + ld3(sum + 13);
        const V3 nbT = mulRT(R, nT);
        if (k == 0) {
            s.r[va + 0] = fT.x; s.r[va + 1] = fT.y; s.r[va + 2] = fT.z;            // -qfrc_smooth (root: bias only)
            s.r[va + 3] = nbT.x; s.r[va + 4] = nbT.y; s.r[va + 5] = nbT.z;
        }
    }
    RS_SYNC();
}

// ------------------------------------------------------------------------------------------
// collision (mj_collision restricted to the pairs this scene can produce)
// ------------------------------------------------------------------------------------------
RS_HD V3 make_frame_y(V3 n, V3 yhint) {   // mju_makeFrame: second axis (third = n x second)
    V3 y = yhint;
    if (dot(y, y) < 0.25f) { y = (n.y < 0.5f && n.y > -0.5f) ? v3(0.f, 1.f, 0.f) : v3(0.f, 0.f, 1.f); }
    float len;
    return normalized(y - dot(n, y) * n, &len);
}

template <int LA, int LB>
RS_HD void add_contact(Ctx<LA, LB>& c, int bA, int bB, float dist, V3 pos, V3 n, V3 yhint, float tran, int key) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    if (!(dist < RS_MARGIN)) return;
    int k = RS_ATOMIC_INC(&s.ncon);
    if (k >= S::MAXC) return;      // counted, dropped: status flag raised by the caller
    s.cbody[k] = (bA + 1) | ((bB + 1) << 8) | ((key & 4095) << 16);     // key: stable identity of this geom pair across evaluations
    s.cD[k] = dist; s.caref[k][0] = tran;      // parked here until make_constraints turns them into D and aref
    if (bA >= 0) RS_ATOMIC_OR(&s.coupled, ((bA < 2 ? bA : c.agent_of_leg((bA - 2) % S::LT)) != (bB < 2 ? bB : c.agent_of_leg((bB - 2) % S::LT))) ? 1 : 2);
    st3(s.cpos[k], pos);
    st3(s.cn[k], n); st3(s.ct1[k], make_frame_y(n, yhint));
}

// sphere (centre cA, radius rA, side A) against sphere (cB, rB, side B): normal A -> B
template <int LA, int LB>
RS_HD void sph_sph(Ctx<LA, LB>& c, int bA, int bB, V3 cA, float rA, V3 cB, float rB, float tran, int key) {
    float len;
    V3 n = normalized(cB - cA, &len);
    float dist = len - rA - rB;
    if (dist < RS_MARGIN) add_contact(c, bA, bB, dist, cA + (rA + 0.5f * dist) * n, n, v3(0, 0, 0), tran, key);
}
RS_HD V3 seg_nearest(V3 e0, V3 e1, V3 p) {
    V3 d = e1 - e0;
    float dd = dot(d, d);
    float t = dd > 1e-20f ? RS_DIV(dot(p - e0, d), dd) : 0.f;
    t = fminf(fmaxf(t, 0.f), 1.f);
    return e0 + t * d;
}
// closest points of two segments given as centre/axis/half-length (mjc_CapsuleCapsule parametrisation)
RS_HD void seg_seg(V3 p1, V3 a1, float l1, V3 p2, V3 a2, float l2, V3* o1, V3* o2) {
    V3 dif = p1 - p2;
    float mb = -dot(a1, a2), u = -dot(a1, dif), v = dot(a2, dif);
    float det = 1.f - mb * mb;
    float x1, x2;
    if (fabsf(det) >= 1e-6f) {
        float idet = RS_RCP(det);
        x1 = (u - mb * v) * idet; x2 = (v - mb * u) * idet;
        if (x1 > l1) { x1 = l1; x2 = v - mb * l1; } else if (x1 < -l1) { x1 = -l1; x2 = v + mb * l1; }
        if (x2 > l2) { x2 = l2; x1 = u - mb * l2; } else if (x2 < -l2) { x2 = -l2; x1 = u + mb * l2; }
        if (x1 > l1) x1 = l1; else if (x1 < -l1) x1 = -l1;
    } else {   // (near-)parallel: centre of the overlap interval
        x1 = fminf(fmaxf(u, -l1), l1);
        x2 = fminf(fmaxf(v - mb * x1, -l2), l2);
    }
    *o1 = p1 + x1 * a1; *o2 = p2 + x2 * a2;
}

// geom i of the pair: 0/1 torso spheres, then per global leg g: aux, hip, ankle capsules
template <int LA, int LB>
RS_HD void geom_of(const Ctx<LA, LB>& c, int i, V3* e0, V3* e1, float* r, int* body, float* iw, int* agent, bool* sphere) {
    typedef Slab<LA, LB> S;
    const S& s = *c.s;
    if (i < 2) {
        *e0 = *e1 = ld3(s.org[i]); *r = c.am[i].torso_r; *body = i; *iw = c.am[i].iw_torso; *agent = i; *sphere = true;
        return;
    }
    int k = i - 2, g = k / 3, kind = k - 3 * g;
    int a = c.agent_of_leg(g), l = g - c.leg0(a);
    *agent = a; *sphere = false; *r = c.am[a].leg_r;
    if (kind == 0) { *e0 = ld3(s.org[a]); *e1 = ld3(s.org[c.bhip(g)]); *body = a; *iw = c.am[a].iw_aux[l]; }
    else if (kind == 1) { *e0 = ld3(s.org[c.bhip(g)]); *e1 = ld3(s.org[c.bank(g)]); *body = c.bhip(g); *iw = c.am[a].iw_hip[l]; }
    else { *e0 = ld3(s.org[c.bank(g)]); *e1 = ld3(s.tip[g]); *body = c.bank(g); *iw = c.am[a].iw_ank[l]; }
}

template <int LA, int LB>
RS_HD void sphere_vs_world(Ctx<LA, LB>& c, int body, V3 p, float r, float iw, V3 yhint, int key, bool floor_too) {
    // floor plane z = RS_FLOOR_Z, normal +z
    if (floor_too) {
        float dist = p.z - RS_FLOOR_Z - r;
        if (dist < RS_MARGIN) add_contact(c, -1, body, dist, v3(p.x, p.y, p.z - r - 0.5f * dist), v3(0.f, 0.f, 1.f), yhint, iw, key);
    }
    // tatami box: centre (0,0,RS_BOX_CZ), half (RS_BOX_HX, RS_BOX_HX, RS_BOX_HZ)
    {
        V3 loc = v3(p.x, p.y, p.z - RS_BOX_CZ);
        V3 cl = v3(fminf(fmaxf(loc.x, -RS_BOX_HX), RS_BOX_HX), fminf(fmaxf(loc.y, -RS_BOX_HX), RS_BOX_HX), fminf(fmaxf(loc.z, -RS_BOX_HZ), RS_BOX_HZ));
        bool outside = (cl.x != loc.x) || (cl.y != loc.y) || (cl.z != loc.z);
        if (outside) {
            float len;
            V3 n = normalized(loc - cl, &len);     // from box to sphere
            float dist = len - r;
            if (dist < RS_MARGIN) add_contact(c, -1, body, dist, v3(cl.x, cl.y, cl.z + RS_BOX_CZ) + (0.5f * dist) * n, n, v3(0, 0, 0), iw, key + 1);
        } else {
            float px = RS_BOX_HX - fabsf(loc.x), py = RS_BOX_HX - fabsf(loc.y), pz = RS_BOX_HZ - fabsf(loc.z);
            V3 n; float best;
            if (px <= py && px <= pz) { best = px; n = v3(loc.x >= 0.f ? 1.f : -1.f, 0.f, 0.f); }
            else if (py <= pz) { best = py; n = v3(0.f, loc.y >= 0.f ? 1.f : -1.f, 0.f); }
            else { best = pz; n = v3(0.f, 0.f, loc.z >= 0.f ? 1.f : -1.f); }
            float dist = -best - r;
            add_contact(c, -1, body, dist, v3(loc.x, loc.y, loc.z + RS_BOX_CZ) + (-r - 0.5f * dist) * n, n, v3(0, 0, 0), iw, key + 1);
        }
    }
}

// conservative filter for capsule_vs_box_edges: can the segment e0-e1 pass within R of one of the tatami's top or vertical edges?
// (seen along a top edge, the segment's bounding rectangle grown by R must contain the edge; likewise from above for a vertical one)
RS_HD bool near_box_edge(V3 e0, V3 e1, float R) {
    const float top = RS_BOX_CZ + RS_BOX_HZ, bot = RS_BOX_CZ - RS_BOX_HZ;
    const float xl = fminf(e0.x, e1.x) - R, xh = fmaxf(e0.x, e1.x) + R, yl = fminf(e0.y, e1.y) - R, yh = fmaxf(e0.y, e1.y) + R;
    const float zl = fminf(e0.z, e1.z) - R, zh = fmaxf(e0.z, e1.z) + R;
    const bool sx = (xl <= RS_BOX_HX && xh >= RS_BOX_HX) || (xl <= -RS_BOX_HX && xh >= -RS_BOX_HX);
    const bool sy = (yl <= RS_BOX_HX && yh >= RS_BOX_HX) || (yl <= -RS_BOX_HX && yh >= -RS_BOX_HX);
    return (zl <= top && zh >= top && (sx || sy)) || (sx && sy && zl <= top && zh >= bot);
}

// capsule INTERIOR (centre pc, unit axis ax, half length hl, radius r) against the tatami's four top and four vertical edges: the
// closest edge whose nearest capsule point is interior to the segment and outside the box.  true = within the margin; *n points
// from the box to the capsule, *be is the point on the edge.
RS_COLD bool capsule_vs_box_edges(V3 pc, V3 ax, float hl, float r, float* dist, V3* be_out, V3* n_out) {
    float best = 1e30f; V3 bg = pc, be = pc;
    const float R = r + RS_MARGIN, top = RS_BOX_CZ + RS_BOX_HZ;
    const float hx = hl * fabsf(ax.x) + R, hy = hl * fabsf(ax.y) + R, hz = hl * fabsf(ax.z) + R;      // half extents of the bounding box grown by R
    RS_UNROLL1
    for (int k = 0; k < 8; k++) {
        // cheap reject, the same bound as near_box_edge for this one edge
        if (k < 4) { const float sg = (k & 2) ? -RS_BOX_HX : RS_BOX_HX; if (fabsf(((k & 1) ? pc.x : pc.y) - sg) > ((k & 1) ? hx : hy) || fabsf(pc.z - top) > hz) continue; }
        else if (fabsf(pc.x - ((k & 1) ? RS_BOX_HX : -RS_BOX_HX)) > hx || fabsf(pc.y - ((k & 2) ? RS_BOX_HX : -RS_BOX_HX)) > hy || fabsf(pc.z - RS_BOX_CZ) > hz + RS_BOX_HZ) continue;
        // k < 4: top edges (y = +-HX along x, x = +-HX along y at z = top); k >= 4: the four vertical edges
        V3 ep, ea; float el;
        if (k < 4) { const float sg = (k & 2) ? -1.f : 1.f; ep = (k & 1) ? v3(sg * RS_BOX_HX, 0.f, RS_BOX_CZ + RS_BOX_HZ) : v3(0.f, sg * RS_BOX_HX, RS_BOX_CZ + RS_BOX_HZ);
                     ea = (k & 1) ? v3(0.f, 1.f, 0.f) : v3(1.f, 0.f, 0.f); el = RS_BOX_HX; }
        else { ep = v3((k & 1) ? RS_BOX_HX : -RS_BOX_HX, (k & 2) ? RS_BOX_HX : -RS_BOX_HX, RS_BOX_CZ); ea = v3(0.f, 0.f, 1.f); el = RS_BOX_HZ; }
        if (fabsf(dot(ax, ea)) > 1.f - 1e-6f) continue;            // parallel: the endpoint tests cover it
        V3 pg, pe;
        seg_seg(pc, ax, hl, ep, ea, el, &pg, &pe);
        if (fabsf(dot(pg - pc, ax)) >= hl * (1.f - 1e-6f)) continue;   // an endpoint is closest
        if (fabsf(pg.x) < RS_BOX_HX && fabsf(pg.y) < RS_BOX_HX && fabsf(pg.z - RS_BOX_CZ) < RS_BOX_HZ) continue;   // inside: face push-out case
        const float dd = norm(pg - pe) - r;
        if (dd < best) { best = dd; bg = pg; be = pe; }
    }
    if (!(best < RS_MARGIN)) return false;
    float l2;
    *n_out = normalized(bg - be, &l2);
    *be_out = be; *dist = best;
    return l2 > 1e-12f;
}

template <int LA, int LB>
RS_HD void collide(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    if (RS_LANE0) { s.ncon = 0; s.coupled = 0; }
    RS_SYNC();
    // bounding data of every geom (centre, axis, half length, radius), parked in the H array (dead until build_H): the pair
    // loops below cull on it and only rebuild the geoms of the few pairs that pass
    static_assert(10 * S::NG <= S::HDED, "geom cache must fit in the H array");
    float* gc = s.H;
    bool rim = false;
    // --- agent geoms against the world (floor plane, tatami box, four border rails) ---
    RS_LANE_LOOP(i, S::NG) {
        V3 e0, e1; float r, iw; int body, agent; bool sph;
        geom_of(c, i, &e0, &e1, &r, &body, &iw, &agent, &sph);
        float len = 0.f;
        V3 ax = sph ? v3(0, 0, 0) : normalized(e1 - e0, &len);
        { float* G = gc + 10 * i; st3(G, 0.5f * (e0 + e1)); st3(G + 3, ax); G[6] = 0.5f * len; G[7] = r; }     // broad-phase cache
        // MuJoCo's capsule geom frame has z = from - to; the plane-capsule routine emits the +z end first
        // keys: geom i owns 12 slots: end0 {floor, box}, end1 {floor, box}, rails 4..7; pair contacts start at 12 * NG
        sphere_vs_world(c, body, e0, r, iw, ax, 12 * i, true);
        if (!sph) sphere_vs_world(c, body, e1, r, iw, ax, 12 * i + 2, true);
        // the rim of the arena (border rails at +-2.0, tatami edges at +-2.3): flagged here, tested in the cold pass below
        rim = rim || ((fmaxf(fmaxf(fabsf(e0.x), fabsf(e1.x)), fmaxf(fabsf(e0.y), fabsf(e1.y))) + r + RS_MARGIN + RS_RAIL_R >= RS_RAIL)
                      && fminf(e0.z, e1.z) - r - RS_MARGIN - RS_RAIL_R < RS_RAIL_Z);
    }
    // Cold pass, entered only by pairs with a geom at the rim (they are about to be done, sumo.py:149-150), so that the hot
    // instruction footprint of an evaluation does not carry it:
    //  * border rails: thin cylinders treated as capsules of radius RS_RAIL_R, key slots 4..7;
    //  * capsule INTERIOR against the tatami's top and vertical edges (the distance from a segment to the box is attained at an
    //    endpoint -- the sphere tests above -- or on an edge): a leg straddling the edge at |x| = 2.3 rests on it instead of
    //    sinking through until an endpoint touches.  At most one contact per capsule (the closest edge), key slot 8.
    if (RS_UNLIKELY(RS_WARP_ANY(rim))) {
        RS_LANE_LOOP(i, S::NG) {
            V3 e0, e1; float r, iw; int body, agent; bool sph;
            geom_of(c, i, &e0, &e1, &r, &body, &iw, &agent, &sph);
            const float R = r + RS_MARGIN + RS_RAIL_R;
            if (fmaxf(fmaxf(fabsf(e0.x), fabsf(e1.x)), fmaxf(fabsf(e0.y), fabsf(e1.y))) + R < RS_RAIL || fminf(e0.z, e1.z) - R >= RS_RAIL_Z) continue;
            float len = 0.f;
            const V3 ax = sph ? v3(0, 0, 0) : normalized(e1 - e0, &len);
            RS_UNROLL1
            for (int k = 0; k < 4; k++) {
                // top: y=+2 along x; right: x=+2 along y; bottom: y=-2 along x; left: x=-2 along y
                const float reach = (k & 1) ? ((k & 2) ? -fminf(e0.x, e1.x) : fmaxf(e0.x, e1.x)) : ((k & 2) ? -fminf(e0.y, e1.y) : fmaxf(e0.y, e1.y));
                if (reach + R < RS_RAIL) continue;      // this rail is out of reach
                V3 rp = (k == 0) ? v3(0.f, RS_RAIL, RS_RAIL_Z) : (k == 1) ? v3(RS_RAIL, 0.f, RS_RAIL_Z) : (k == 2) ? v3(0.f, -RS_RAIL, RS_RAIL_Z) : v3(-RS_RAIL, 0.f, RS_RAIL_Z);
                V3 ra = (k & 1) ? v3(0.f, 1.f, 0.f) : v3(1.f, 0.f, 0.f);
                V3 pg, pr;
                if (sph) { pg = e0; pr = seg_nearest(rp - RS_RAIL * ra, rp + RS_RAIL * ra, e0); }
                else seg_seg(0.5f * (e0 + e1), ax, 0.5f * len, rp, ra, RS_RAIL, &pg, &pr);
                sph_sph(c, -1, body, pr, RS_RAIL_R, pg, r, iw, 12 * i + 4 + k);
            }
            if (sph || !near_box_edge(e0, e1, r + RS_MARGIN)) continue;
            V3 be, n; float best;
            if (capsule_vs_box_edges(0.5f * (e0 + e1), ax, 0.5f * len, r, &best, &be, &n)) add_contact(c, -1, body, best, be + (0.5f * best) * n, n, v3(0, 0, 0), iw, 12 * i + 8);
        }
    }
    RS_SYNC();
    // --- agent 0 geoms against agent 1 geoms ---
    {
        V3 dt = ld3(s.org[1]) - ld3(s.org[0]);
        // (a tighter bound from the actual geom extents was measured: it skips the pair loop more often but does not pay for its atomics)
        float reach = c.am[0].reach + c.am[1].reach + RS_MARGIN;
#ifdef RS_EXPERIMENT_NO_PAIRS
        reach = 0.f;      // developer experiment: what would the step cost without inter-agent contacts?
#endif
        if (dot(dt, dt) < reach * reach) {
            RS_LANE_LOOP(p, S::NGA * S::NGB) {
                int ia = p / S::NGB, ib = p - ia * S::NGB;
                // pair geom indices: agent 0 torso = 0, its leg geoms 2..2+3LA ; agent 1 torso = 1, leg geoms after
                int gi = ia == 0 ? 0 : 1 + ia, gj = ib == 0 ? 1 : 1 + 3 * LA + ib;
                const float* GA = gc + 10 * gi; const float* GB = gc + 10 * gj;
                V3 ca = ld3(GA), cb = ld3(GB);
                float bound = GA[6] + GB[6] + GA[7] + GB[7] + RS_MARGIN;
                V3 dc = cb - ca;
                if (dot(dc, dc) < bound * bound) {
                    V3 a0, a1, b0, b1; float rA, rB, iwA, iwB; int bA, bB, agA, agB; bool sA, sB;
                    geom_of(c, gi, &a0, &a1, &rA, &bA, &iwA, &agA, &sA);
                    geom_of(c, gj, &b0, &b1, &rB, &bB, &iwB, &agB, &sB);
                    V3 pA, pB;
                    if (sA && sB) { pA = a0; pB = b0; }
                    else if (sA) { pA = a0; pB = seg_nearest(b0, b1, a0); }
                    else if (sB) { pB = b0; pA = seg_nearest(a0, a1, b0); }
                    else seg_seg(ca, ld3(GA + 3), GA[6], cb, ld3(GB + 3), GB[6], &pA, &pB);
#ifdef RS_EXPERIMENT_NO_PAIR_CONTACTS
                    if (c.max_newton < 0)      // developer experiment: run the pair tests, drop their contacts
#endif
                    sph_sph(c, bA, bB, pA, rA, pB, rB, iwA + iwB, 12 * S::NG + p);
                }
            }
        }
    }
    // --- geoms of one agent against each other (mj_collision's parent/child and weld filters applied): only for the
    //     six- and eight-legged bodies; the Ant's legs cannot reach each other (tests/test_oracle.py) ---
    RS_UNROLL1
    for (int a = 0; a < 2; a++) {
        const int L = c.L(a);
        if (L <= 4) continue;
        const int npair = 2 * L * (L - 1), nank = L * (L + 1);      // 4 * C(L,2) leg-leg combinations, ankle vs torso group
        RS_LANE_LOOP(p, npair + nank) {
            int gi, gj;
            if (p < npair) {
                int pi = p >> 2, combo = p & 3, l = 0;
                while (pi >= L - 1 - l) { pi -= L - 1 - l; l++; }
                int m = l + 1 + pi;
                gi = 2 + 3 * (c.leg0(a) + l) + 1 + (combo >> 1);
                gj = 2 + 3 * (c.leg0(a) + m) + 1 + (combo & 1);
            } else {
                int q = p - npair, l = q / (L + 1), t = q - l * (L + 1);
                gi = t == 0 ? a : 2 + 3 * (c.leg0(a) + t - 1);      // torso sphere or a welded stub capsule
                gj = 2 + 3 * (c.leg0(a) + l) + 2;                   // ankle capsule
            }
            const float* GA = gc + 10 * gi; const float* GB = gc + 10 * gj;
            V3 ca = ld3(GA), cb = ld3(GB);
            float bound = GA[6] + GB[6] + GA[7] + GB[7] + RS_MARGIN;
            V3 dc = cb - ca;
            if (dot(dc, dc) < bound * bound) {
                V3 a0, a1, b0, b1; float rA, rB, iwA, iwB; int bA, bB, agA, agB; bool sA, sB;
                geom_of(c, gi, &a0, &a1, &rA, &bA, &iwA, &agA, &sA);
                geom_of(c, gj, &b0, &b1, &rB, &bB, &iwB, &agB, &sB);
                V3 pA, pB;
                if (sA) { pA = a0; pB = seg_nearest(b0, b1, a0); }
                else seg_seg(ca, ld3(GA + 3), GA[6], cb, ld3(GB + 3), GB[6], &pA, &pB);
                sph_sph(c, bA, bB, pA, rA, pB, rB, iwA + iwB, 12 * S::NG + S::NGA * S::NGB + a * 512 + p);
            }
        }
    }
    RS_SYNC();
    if (s.ncon > S::MAXC) { if (RS_LANE0) { s.ncon = S::MAXC; s.status |= RS_STATUS_CONTACT_FULL; } }
    RS_SYNC();
    // the atomic counter hands out slots in arrival order; sort the list by key (rank = number of smaller keys, staged through
    // the H array) so that every later sum over contacts runs in a fixed order and a step is bit-reproducible
    if (s.ncon > 1) {
        static_assert(12 * S::MAXC <= S::HDED, "contact staging must fit in the H array");
        RS_LANE_LOOP(k, s.ncon) {
            const int key = s.ckey(k);
            int rank = 0;
            RS_UNROLL1
            for (int j = 0; j < s.ncon; j++) rank += s.ckey(j) < key ? 1 : 0;
            float* G = s.H + 12 * rank;
            st3(G, ld3(s.cpos[k])); st3(G + 3, ld3(s.cn[k])); st3(G + 6, ld3(s.ct1[k]));
            G[9] = s.cD[k]; G[10] = s.caref[k][0]; ((int*)G)[11] = s.cbody[k];
        }
        RS_SYNC();
        RS_LANE_LOOP(k, s.ncon) {
            const float* G = s.H + 12 * k;
            st3(s.cpos[k], ld3(G)); st3(s.cn[k], ld3(G + 3)); st3(s.ct1[k], ld3(G + 6));
            s.cD[k] = G[9]; s.caref[k][0] = G[10]; s.cbody[k] = ((const int*)G)[11];
        }
        RS_SYNC();
    }
}

// ------------------------------------------------------------------------------------------
// twists: body velocities generated by a generalized vector `vec` (J x without storing J)
// ------------------------------------------------------------------------------------------
template <int LA, int LB>
RS_HD void twists(Ctx<LA, LB>& c, const float* vec) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    RS_LANE_LOOP(g, S::LT) {
        int a = c.agent_of_leg(g);
        const float* vv = vec + c.vadr(a);
        V3 wt = mulR(s.Rt[a], v3(vv[3], vv[4], vv[5])), vt = v3(vv[0], vv[1], vv[2]);
        if (g == c.leg0(a)) { st3(s.tw[a], wt); st3(s.tw[a] + 3, vt); }
        V3 pt = ld3(s.org[a]), ph = ld3(s.org[c.bhip(g)]), pa = ld3(s.org[c.bank(g)]);
        int dh = c.hipdof(g);
        V3 vh = vt + cross(wt, ph - pt);
        V3 wh = wt + vec[dh] * c.hip_axis(g);
        V3 vk = vh + cross(wh, pa - ph);
        V3 wk = wh + vec[dh + 1] * ld3(s.axa[g]);
        st3(s.tw[c.bhip(g)], wh); st3(s.tw[c.bhip(g)] + 3, vh);
        st3(s.tw[c.bank(g)], wk); st3(s.tw[c.bank(g)] + 3, vk);
    }
    RS_SYNC();
}
template <int LA, int LB>
RS_HD V3 point_vel(const Slab<LA, LB>& s, int b, V3 p) {
    if (b < 0) return v3(0, 0, 0);
    return ld3(s.tw[b] + 3) + cross(ld3(s.tw[b]), p - ld3(s.org[b]));
}
// rows of J * vec for every contact (4 pyramid rows) and every limit row, after twists(vec)
template <int LA, int LB>
RS_HD void rows_of(Ctx<LA, LB>& c, const float* vec, float (*cout)[4], float* lout) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    RS_LANE_LOOP(k, s.ncon) {
        V3 p = ld3(s.cpos[k]);
        V3 rel = point_vel(s, s.bB(k), p) - point_vel(s, s.bA(k), p);
        V3 n = ld3(s.cn[k]), t1 = ld3(s.ct1[k]);
        float un = dot(n, rel), u1 = dot(t1, rel), u2 = dot(cross(n, t1), rel);
        cout[k][0] = un + RS_MU * u1; cout[k][1] = un - RS_MU * u1; cout[k][2] = un + RS_MU * u2; cout[k][3] = un - RS_MU * u2;
    }
    RS_LANE_LOOP(j, S::NU) {
        int g = j >> 1;
        lout[j] = s.lsgn[j] * vec[c.hipdof(g) + (j & 1)];
    }
    RS_SYNC();
}

// ------------------------------------------------------------------------------------------
// constraint parameters (mj_makeConstraint + mj_makeImpedance): limits, pyramidal contacts
// ------------------------------------------------------------------------------------------
RS_HD float impedance(float pos_minus_margin) {
    float x = fabsf(pos_minus_margin) * (1.0f / RS_WIDTH);
    float y;
    if (x >= 1.f) y = 1.f; else if (x <= 0.5f) y = 2.f * x * x; else y = 1.f - 2.f * (1.f - x) * (1.f - x);
    return RS_DMIN + y * (RS_DMAX - RS_DMIN);
}
template <int LA, int LB>
RS_HD void make_constraints(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    const float tc = fmaxf(0.02f, 2.f * c.h);
    const float Kc = 1.f / (RS_DMAX * RS_DMAX * tc * tc);
    // limits first: sign and D, then J v through rows_of.  The prediction masks (bit j = limit row j, one row per lane) are built with
    // warp votes on the device -- a lane-per-bit atomicOr on one shared word serialises its 16 lanes
#if defined(__CUDA_ARCH__)
    static_assert(S::NU <= 32, "limit masks are one word, one lane per row");
    bool pv_lane = false, pm_lane = false;
#else
    if (RS_LANE0) { s.pmask = 0; s.pvalid = 0; }
    RS_SYNC();
#endif
    RS_LANE_LOOP(j, S::NU) {
        int g = j >> 1, a = c.agent_of_leg(g), l = g - c.leg0(a), isank = j & 1;
        const rs_agent_model& m = c.am[a];
        float qv = s.q[c.hipq(g) + isank];
        float lo = isank ? m.lo_ank[l] : m.lo_hip[l], hi = isank ? m.hi_ank[l] : m.hi_hip[l];
        float sgn = 0.f, pos = 0.f;
        if (qv - lo < 0.f) { sgn = 1.f; pos = qv - lo; } else if (hi - qv < 0.f) { sgn = -1.f; pos = hi - qv; }
        // the row's state at the previous evaluation's solution predicts its state now far better than the sign of the
        // residual at the warm start does (aref moves by B * dvel between RK stages, more than |jar| of a loaded row)
#if defined(__CUDA_ARCH__)
        if (sgn != 0.f && s.lsgn[j] == sgn) { pv_lane = true; pm_lane = s.ljar[j] < 0.f; }
#else
        if (sgn != 0.f && s.lsgn[j] == sgn) { RS_ATOMIC_OR(&s.pvalid, 1 << j); if (s.ljar[j] < 0.f) RS_ATOMIC_OR(&s.pmask, 1 << j); }
#endif
        s.lsgn[j] = sgn;
        float imp = impedance(pos);
        float diag = isank ? m.iwd_ank[l] : m.iwd_hip[l];
        float R = fmaxf(1e-15f, RS_DIV((1.f - imp) * diag, imp));
        s.lD[j] = sgn != 0.f ? RS_DIV(1.f, R) : 0.f;
        s.laref[j] = -Kc * imp * pos;           // position term only (see the note on aref below)
    }
#if defined(__CUDA_ARCH__)
    { const unsigned bv = __ballot_sync(0xffffffffu, pv_lane), bm = __ballot_sync(0xffffffffu, pm_lane); if (RS_LANE0) { s.pvalid = (int)bv; s.pmask = (int)bm; } }
#endif
    // aref = -B (J v) - K imp pos.  Only its position term is stored: the residual the solver starts from,
    //   jar = J x0 - aref = J (x0 + B v) + K imp pos,
    // takes ONE pass of twists + rows_of over the vector x0 + B v (solve(), first pass) instead of one for J v here and one for J x0 there
    RS_LANE_LOOP(k, s.ncon) {
        float pm = s.cD[k] - RS_MARGIN;
        float imp = impedance(pm);
        float diag = s.caref[k][0] * (1.f + RS_MU * RS_MU);
        float R = fmaxf(1e-15f, RS_DIV((1.f - imp) * diag, imp));
        R = 2.f * RS_MU * RS_MU * R;
        s.cD[k] = RS_DIV(1.f, R);
        for (int r = 0; r < 4; r++) s.caref[k][r] = -Kc * imp * pm;
    }
    RS_SYNC();
}
RS_HD float solref_damping(float h) {        // B of mj_makeImpedance for solref (0.02, 1): 2 / (dmax * timeconst)
    const float tc = fmaxf(0.02f, 2.f * h);
    return 2.f / (RS_DMAX * tc);
}

// ------------------------------------------------------------------------------------------
// J^T f through body wrenches
// ------------------------------------------------------------------------------------------
template <int LA, int LB>
RS_HD void jt_forces(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    // every body gathers the contact forces that act on it (no atomics: fixed summation order, bit-reproducible)
    RS_LANE_LOOP(b, S::NB) {
        V3 T = v3(0.f, 0.f, 0.f), F = v3(0.f, 0.f, 0.f);
        const V3 o = ld3(s.org[b]);
        RS_UNROLL1
        for (int k = 0; k < s.ncon; k++) {
            const int cb = s.cbody[k], act = (cb >> 28) & 15;
            const bool onB = ((cb >> 8) & 255) - 1 == b;
            if ((!onB && (cb & 255) - 1 != b) || act == 0) continue;
            const float D = s.cD[k];
            const float f0 = (act & 1) ? -D * s.cjar[k][0] : 0.f, f1 = (act & 2) ? -D * s.cjar[k][1] : 0.f;
            const float f2 = (act & 4) ? -D * s.cjar[k][2] : 0.f, f3 = (act & 8) ? -D * s.cjar[k][3] : 0.f;
            const V3 cn_ = ld3(s.cn[k]), ct_ = ld3(s.ct1[k]);
            V3 Fc = (f0 + f1 + f2 + f3) * cn_ + (RS_MU * (f0 - f1)) * ct_ + (RS_MU * (f2 - f3)) * cross(cn_, ct_);
            if (!onB) Fc = (-1.f) * Fc;
            T = T + cross(ld3(s.cpos[k]) - o, Fc);
            F = F + Fc;
        }
        st3(s.wr[b], T); st3(s.wr[b] + 3, F);
    }
    RS_SYNC();
    RS_LANE_LOOP(g, S::LT) {
        int dh = c.hipdof(g);
        V3 ph = ld3(s.org[c.bhip(g)]), pa = ld3(s.org[c.bank(g)]);
        V3 Ta = ld3(s.wr[c.bank(g)]), Fa = ld3(s.wr[c.bank(g)] + 3);
        V3 Fh = ld3(s.wr[c.bhip(g)] + 3) + Fa;
        V3 Th = ld3(s.wr[c.bhip(g)]) + Ta + cross(pa - ph, Fa);
        float fl_h = ((s.lmask >> (2 * g)) & 1) ? -s.lD[2 * g] * s.ljar[2 * g] : 0.f;
        float fl_a = ((s.lmask >> (2 * g + 1)) & 1) ? -s.lD[2 * g + 1] * s.ljar[2 * g + 1] : 0.f;
        s.jtf[dh] = dot(c.hip_axis(g), Th) + s.lsgn[2 * g] * fl_h;
        s.jtf[dh + 1] = dot(ld3(s.axa[g]), Ta) + s.lsgn[2 * g + 1] * fl_a;
        st3(s.legF[g], Th); st3(s.legF[g] + 3, Fh);     // about the hip origin
    }
    RS_SYNC();
    RS_LANE_LOOP(a, 2) {
        V3 pt = ld3(s.org[a]);
        V3 T = ld3(s.wr[a]), F = ld3(s.wr[a] + 3);
        for (int l = 0; l < c.L(a); l++) {
            int g = c.leg0(a) + l;
            V3 Fg = ld3(s.legF[g] + 3);
            T = T + ld3(s.legF[g]) + cross(ld3(s.org[c.bhip(g)]) - pt, Fg);
            F = F + Fg;
        }
        V3 Tb = mulRT(s.Rt[a], T);
        int va = c.vadr(a);
        s.jtf[va] = F.x; s.jtf[va + 1] = F.y; s.jtf[va + 2] = F.z;
        s.jtf[va + 3] = Tb.x; s.jtf[va + 4] = Tb.y; s.jtf[va + 5] = Tb.z;
    }
    RS_SYNC();
}

// ------------------------------------------------------------------------------------------
// H = M + J^T D_active J, assembled contact by contact
// ------------------------------------------------------------------------------------------
template <int LA, int LB>
RS_HD V3 side_col(const Ctx<LA, LB>& c, int b, int k, V3 p, bool skip_root, int* idx) {
    // k-th (0..7) dof of the chain of body b: its index and the vector jc with  J(dir) = jc . dir  for a force direction
    // `dir` applied at point p (translation: e_k; rotation about axis u through o: u x (p - o))
    typedef Slab<LA, LB> S;
    const S& s = *c.s;
    *idx = -1;
    if (b < 0) return v3(0.f, 0.f, 0.f);
    int a, g = -1, depth = 0;
    if (b < 2) a = b;
    else if (b < 2 + S::LT) { g = b - 2; a = c.agent_of_leg(g); depth = 1; }
    else { g = b - 2 - S::LT; a = c.agent_of_leg(g); depth = 2; }
    const int va = c.vadr(a);
    if (k < 6 && skip_root) return v3(0.f, 0.f, 0.f);      // both bodies hang off the same floating base: root columns cancel exactly
    if (k < 3) { *idx = va + k; return v3(k == 0 ? 1.f : 0.f, k == 1 ? 1.f : 0.f, k == 2 ? 1.f : 0.f); }
    V3 u, o;
    if (k < 6) { const float* R = s.Rt[a]; const int e = k - 3; u = v3(R[e], R[3 + e], R[6 + e]); o = ld3(s.org[a]); *idx = va + k; }
    else if (k == 6 && depth >= 1) { u = c.hip_axis(g); o = ld3(s.org[c.bhip(g)]); *idx = c.hipdof(g); }
    else if (k == 7 && depth >= 2) { u = ld3(s.axa[g]); o = ld3(s.org[c.bank(g)]); *idx = c.hipdof(g) + 1; }
    else return v3(0.f, 0.f, 0.f);
    return cross(u, p - o);
}

template <int LA, int LB>
RS_HD void build_H(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    float* H = s.H;
    const int hn = (s.coupled & 1) ? (int)S::HFULL : (int)S::HBD;
    RS_LANE_LOOP(i, hn) { H[i] = 0.f; }
    RS_SYNC();
    // scatter the arrowhead inertia, one row per lane, and the active limit rows (diagonal)
    RS_LANE_LOOP(i, S::NV) {
        const int a = i >= S::NVA ? 1 : 0, va = c.vadr(a), k = i - va;
        if (k < 6) {
            RS_UNROLL1
            for (int j = 0; j < 6; j++) H[s.hidx(i, va + j)] = s.Mr[a][k * 6 + j];
            RS_UNROLL1
            for (int l = 0; l < c.L(a); l++) {
                const int g = c.leg0(a) + l, dh = va + 6 + 2 * l;
                H[s.hidx(i, dh)] = s.Mc[g][2 * k]; H[s.hidx(i, dh + 1)] = s.Mc[g][2 * k + 1];
            }
        } else {
            const int l = (k - 6) >> 1, isank = (k - 6) & 1, g = c.leg0(a) + l, dh = va + 6 + 2 * l, j = 2 * g + isank;
            RS_UNROLL1
            for (int kk = 0; kk < 6; kk++) H[s.hidx(i, va + kk)] = s.Mc[g][2 * kk + isank];
            const float lim = ((s.lmask >> j) & 1) ? s.lD[j] : 0.f;
            H[s.hidx(i, dh)] = isank ? s.Ml[g][1] : s.Ml[g][0] + lim;
            H[s.hidx(i, dh + 1)] = isank ? s.Ml[g][2] + lim : s.Ml[g][1];
        }
    }
    RS_SYNC();
    const int ncon = s.ncon;
    for (int k = 0; k < ncon; k++) {
        const int act = s.cact(k);
        float a0 = (float)(act & 1), a1 = (float)((act >> 1) & 1), a2 = (float)((act >> 2) & 1), a3 = (float)((act >> 3) & 1);
        float na = a0 + a1 + a2 + a3;
        if (na == 0.f) continue;       // uniform across the warp (shared data)
        float D = s.cD[k];
        float cnn = D * na, cn1 = RS_MU * D * (a0 - a1), c11 = RS_MU * RS_MU * D * (a0 + a1);
        float cn2 = RS_MU * D * (a2 - a3), c22 = RS_MU * RS_MU * D * (a2 + a3);
        float* sc = s.scr;
        const int bA = s.bA(k), bB = s.bB(k);
        RS_LANE_LOOP(e, 16) {
            int side = e >> 3, kk = e & 7;
            int b = side ? bB : bA, bo = side ? bA : bB;
            float sg = side ? 1.f : -1.f;
            V3 p = ld3(s.cpos[k]), n = ld3(s.cn[k]), t1 = ld3(s.ct1[k]);
            int idx;
            const bool same = bo >= 0 && b >= 0 && ((bo < 2 ? bo : c.agent_of_leg((bo - 2) % S::LT)) == (b < 2 ? b : c.agent_of_leg((b - 2) % S::LT)));
            const V3 jc = sg * side_col(c, b, kk, p, same, &idx);
            sc[e] = (float)idx; sc[16 + e] = dot(jc, n); sc[32 + e] = dot(jc, t1); sc[48 + e] = dot(jc, cross(n, t1));
        }
        RS_SYNC();
        int lo = bA < 0 ? 8 : 0;       // world side contributes nothing
        int n = 16 - lo;
        RS_LANE_LOOP(e, n * n) {
            int r = lo + e / n, cc = lo + e % n;
            int ir = (int)sc[r], ic = (int)sc[cc];
            if (ir >= 0 && ic >= 0) {
                float nr = sc[16 + r], nc = sc[16 + cc], t1r = sc[32 + r], t1c = sc[32 + cc], t2r = sc[48 + r], t2c = sc[48 + cc];
                float val = cnn * nr * nc + cn1 * (nr * t1c + t1r * nc) + c11 * t1r * t1c + cn2 * (nr * t2c + t2r * nc) + c22 * t2r * t2c;
                H[s.hidx(ir, ic)] += val;
            }
        }
        RS_SYNC();
    }
    RS_SYNC();
}

// Solve  H d = -g  (g in s.d on entry) by Gauss-Jordan elimination without pivoting (H is SPD).
// Row j lives with lane j (rows are smem-resident, stride NVP is odd -> conflict-free); every lane
// eliminates the pivot column from its own row, rows above the pivot included, so there is no
// back-substitution.  One warp barrier per pivot.  When no inter-agent contact couples the agents
// H is block diagonal and both per-agent blocks are eliminated in the same pass (half the pivots,
// half the columns).
#if defined(__CUDA_ARCH__)
// The same elimination with the matrix rows in REGISTERS (N <= 32: lane j owns row j, fully unrolled so that every register
// index is static) and the pivot row broadcast by warp shuffles: no shared-memory round trip and no barrier per pivot.  A pair
// with an inter-agent contact is the straggler of its block (its warp runs alone while 27 others wait at the evaluation
// barrier), so what counts here is the latency of ONE warp: 28 pivots cost ~2 k cycles this way against ~28 k through shared
// memory (tools/pair_cost_profile.py).  Same operations in the same order as the loop below: bit-identical results.
template <int N>
__device__ __forceinline__ void gj_rows_in_registers(const float* Hb, const int ld, float* d) {
    const int j = threadIdx.x & 31;
    float row[N + 1];
#pragma unroll
    for (int cc = 0; cc < N; cc++) row[cc] = j < N ? Hb[j * ld + cc] : 0.f;
    row[N] = j < N ? -d[j] : 0.f;
    float diag = 1.f;
#pragma unroll
    for (int k = 0; k < N; k++) {
        const float pk = __shfl_sync(0xffffffffu, row[k], k);
        const float f = (j == k) ? 0.f : row[k] * RS_RCP(fmaxf(pk, 1e-12f));
        if (j == k) diag = pk;
#pragma unroll
        for (int cc = k + 1; cc <= N; cc++) row[cc] = fmaf(-f, __shfl_sync(0xffffffffu, row[cc], k), row[cc]);
    }
    __syncwarp();
    if (j < N) d[j] = row[N] * RS_RCP(fmaxf(diag, 1e-12f));
    __syncwarp();
}
#endif

template <int LA, int LB>
RS_HD void chol_solve(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    const bool bd = !(s.coupled & 1);
#if defined(__CUDA_ARCH__) && !defined(RS_NO_REG_GJ)
    if (!bd) {
        if constexpr (S::NV <= 32) { gj_rows_in_registers<S::NV>(s.H, S::NVP, s.d); return; }
    } else if constexpr (LA > 4 || LB > 4) {        // intra-agent contacts exist only for the six- and eight-legged bodies
        gj_rows_in_registers<S::NVA>(s.H, S::NVA + 1, s.d);
        gj_rows_in_registers<S::NVB>(s.H + S::BSA, S::NVB + 1, s.d + S::NVA);
        return;
    }
#endif
    const int nk = bd ? (S::NVA > S::NVB ? S::NVA : S::NVB) : S::NV;
    RS_LANE_LOOP(j, S::NV) { s.d[j] = -s.d[j]; }
    RS_SYNC();
    for (int k = 0; k < nk; k++) {
        RS_LANE_LOOP(j, S::NV) {
            const int b0 = (bd && j >= S::NVA) ? S::NVA : 0;
            const int bn = bd ? (j >= S::NVA ? S::NVB : S::NVA) : S::NV;
            const int pk = b0 + k;
            if (k < bn && j != pk) {
                const float* prow = s.H + s.hidx(pk, b0);       // pivot row, columns of this block
                float* jrow = s.H + s.hidx(j, b0);
                float f = jrow[k] * RS_RCP(fmaxf(prow[k], 1e-12f));
                for (int cc = k + 1; cc < bn; cc++) jrow[cc] = fmaf(-f, prow[cc], jrow[cc]);
                s.d[j] = fmaf(-f, s.d[pk], s.d[j]);
            }
        }
        RS_SYNC();
    }
    RS_LANE_LOOP(j, S::NV) { s.d[j] = s.d[j] * RS_RCP(fmaxf(s.H[s.hidx(j, j)], 1e-12f)); }
    RS_SYNC();
}

// ---- arrowhead fast path -----------------------------------------------------------------------------------------------
// With world contacts only, every row of J touches one floating base and at most one leg, so H = M + J^T D J keeps M's
// arrowhead form per agent:  [ A  B ; B^T  Dg ]  with A 6x6, B_g 6x2 and Dg_g 2x2 per leg.  It is assembled in that compact
// form (same layout as Mr | Mc | Ml) and solved by eliminating the legs first:
//   S = A - sum_g B_g Dg_g^-1 B_g^T,   S d_t = rhs_t - sum_g B_g Dg_g^-1 rhs_g,   d_g = Dg_g^-1 rhs_g - (B_g Dg_g^-1)^T d_t.
template <int LA, int LB>
struct Arrow {
    typedef Slab<LA, LB> S;
    enum { A0 = 0, B0 = 72, D0 = 72 + 12 * S::LT, NCOPY = 72 + 15 * S::LT, W0 = NCOPY, Y0 = W0 + 12 * S::LT, S0 = Y0 + 2 * S::LT, END = S0 + 84 };
    static_assert((int)END <= (int)S::HDED, "compact arrowhead H must fit in the H array");
    static_assert(offsetof(S, Ml) - offsetof(S, Mr) == (72 + 12 * S::LT) * sizeof(float), "Mr | Mc | Ml must be contiguous");
};

template <int LA, int LB>
RS_HD void build_H_arrow(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    typedef Arrow<LA, LB> AR;
    S& s = *c.s;
    float* H = s.H;
    RS_LANE_LOOP(i, (int)AR::NCOPY) { H[i] = (&s.Mr[0][0])[i]; }
    RS_SYNC();
    RS_LANE_LOOP(j, S::NU) { if ((s.lmask >> j) & 1) H[AR::D0 + 3 * (j >> 1) + 2 * (j & 1)] += s.lD[j]; }
    const int ncon = s.ncon;
    // The cost of this routine per contact is what makes a block wait for its slowest warp (the warps re-align per Newton
    // iteration), so the contacts are handled side by side: one pass computes the direction Jacobians of ALL contacts (lane =
    // contact x chain dof: J(dir) = jc . dir for the 8 dofs of the touched body's chain, zero where the chain has no such dof),
    // parked in the alias zone (dead between the gradient and the end of the linear solve); then one 32-lane pass per ACTIVE
    // contact adds its 8 x 8 block  J^T D J  (lane = row r x column pair) into A | B | Dg.
    enum { CHUNK = S::ALIAS / 24 < 16 ? S::ALIAS / 24 : 16 };
    float* SC = &s.tw[0][0];
    RS_UNROLL1
    for (int k0 = 0; k0 < ncon; k0 += CHUNK) {
        const int nk = ncon - k0 < CHUNK ? ncon - k0 : (int)CHUNK;
        RS_LANE_LOOP(i, 8 * nk) {
            const int k = k0 + (i >> 3), e = i & 7;
            if (s.bA(k) >= 0) continue;                           // inter-agent contact: low-rank correction, woodbury_solve()
            const int b = s.bB(k);                                // the other side is the world
            const V3 p = ld3(s.cpos[k]), n = ld3(s.cn[k]), t1 = ld3(s.ct1[k]);
            int idx;
            const V3 jc = side_col(c, b, e, p, false, &idx);
            float* o = SC + 3 * i;
            o[0] = dot(jc, n); o[1] = dot(jc, t1); o[2] = dot(jc, cross(n, t1));      // (jc = 0 where idx < 0)
        }
        RS_SYNC();
        RS_UNROLL1
        for (int k = k0; k < k0 + nk; k++) {
            const int act = s.cact(k);
            if (act == 0 || s.bA(k) >= 0) continue;        // uniform across the warp (shared data)
            const float a0 = (float)(act & 1), a1 = (float)((act >> 1) & 1), a2 = (float)((act >> 2) & 1), a3 = (float)((act >> 3) & 1);
            const float D = s.cD[k];
            const float cnn = D * (a0 + a1 + a2 + a3), cn1 = RS_MU * D * (a0 - a1), c11 = RS_MU * RS_MU * D * (a0 + a1);
            const float cn2 = RS_MU * D * (a2 - a3), c22 = RS_MU * RS_MU * D * (a2 + a3);
            const int b = s.bB(k);
            const int g = b < 2 ? 0 : (b - 2) % S::LT, a = b < 2 ? b : c.agent_of_leg(g);
            const int ndof = b < 2 ? 6 : (b < 2 + S::LT ? 7 : 8);                     // dofs of the chain: root, + hip, + ankle
            const float* sc = SC + 24 * (k - k0);
            RS_LANE_LOOP(e, 32) {
                const int r = e >> 2, c0 = (e & 3) << 1;
                if (r < ndof && c0 < ndof) {
                    const float nr = sc[3 * r], t1r = sc[3 * r + 1], t2r = sc[3 * r + 2];
                    for (int cc = c0; cc < c0 + 2; cc++) {
                        if (cc >= ndof || (r >= 6 && cc < r)) continue;      // the leg rows only hold their own 2 x 2 upper triangle
                        const float nc = sc[3 * cc], t1c = sc[3 * cc + 1], t2c = sc[3 * cc + 2];
                        const float val = nr * (cnn * nc + cn1 * t1c + cn2 * t2c) + t1r * (cn1 * nc + c11 * t1c) + t2r * (cn2 * nc + c22 * t2c);
                        if (r < 6) { if (cc < 6) H[AR::A0 + a * 36 + r * 6 + cc] += val; else H[AR::B0 + 12 * g + 2 * r + (cc - 6)] += val; }
                        else H[AR::D0 + 3 * g + (r - 6) + (cc - 6)] += val;
                    }
                }
            }
            RS_SYNC();
        }
    }
    RS_SYNC();
}

// s.d = -H^-1 s.d for the compact arrowhead H of build_H_arrow
template <int LA, int LB>
RS_HD void arrow_solve(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    typedef Arrow<LA, LB> AR;
    S& s = *c.s;
    float* H = s.H;
    RS_LANE_LOOP(g, S::LT) {
        const float* Dg = H + AR::D0 + 3 * g; const float* B = H + AR::B0 + 12 * g;
        float* W = H + AR::W0 + 12 * g;
        const float idet = RS_RCP(fmaxf(Dg[0] * Dg[2] - Dg[1] * Dg[1], 1e-20f));
        const float i00 = Dg[2] * idet, i01 = -Dg[1] * idet, i11 = Dg[0] * idet;
        for (int k = 0; k < 6; k++) { W[2 * k] = B[2 * k] * i00 + B[2 * k + 1] * i01; W[2 * k + 1] = B[2 * k] * i01 + B[2 * k + 1] * i11; }
        const int dh = c.hipdof(g);
        const float gh = -s.d[dh], ga = -s.d[dh + 1];
        H[AR::Y0 + 2 * g] = i00 * gh + i01 * ga; H[AR::Y0 + 2 * g + 1] = i01 * gh + i11 * ga;
    }
    RS_SYNC();
#if defined(__CUDA_ARCH__) && !defined(RS_NO_REG_GJ)
    // Schur complement onto the two floating bases and its 6x6 elimination with the rows in registers: lane 6 a + r owns row r
    // of agent a (7 numbers), the pivot row travels by warp shuffle -- no shared-memory round trip and no barrier per pivot
    {
        const int lane = threadIdx.x & 31, a = lane >= 6 ? 1 : 0, r = lane - 6 * a;
        const bool own = lane < 12;
        float row[7];
#pragma unroll
        for (int cc = 0; cc < 7; cc++) row[cc] = 0.f;
        if (own) {
            const int l0 = c.leg0(a), va = c.vadr(a);
#pragma unroll
            for (int cc = 0; cc < 6; cc++) row[cc] = H[AR::A0 + a * 36 + r * 6 + cc];
            row[6] = -s.d[va + r];
            RS_UNROLL1
            for (int l = 0; l < c.L(a); l++) {
                const float* W = H + AR::W0 + 12 * (l0 + l); const float* B = H + AR::B0 + 12 * (l0 + l); const float* Y = H + AR::Y0 + 2 * (l0 + l);
                const float w0 = W[2 * r], w1 = W[2 * r + 1];
#pragma unroll
                for (int cc = 0; cc < 6; cc++) row[cc] -= w0 * B[2 * cc] + w1 * B[2 * cc + 1];
                row[6] -= B[2 * r] * Y[0] + B[2 * r + 1] * Y[1];
            }
        }
        float diag = 1.f;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int src = own ? 6 * a + k : lane;
            const float pk = __shfl_sync(0xffffffffu, row[k], src);
            const float f = (r == k) ? 0.f : row[k] * RS_RCP(fmaxf(pk, 1e-12f));
            if (r == k) diag = pk;
#pragma unroll
            for (int cc = k + 1; cc < 7; cc++) row[cc] = fmaf(-f, __shfl_sync(0xffffffffu, row[cc], src), row[cc]);
        }
        if (own) s.d[c.vadr(a) + r] = row[6] * RS_RCP(fmaxf(diag, 1e-12f));
    }
    RS_SYNC();
#else
    // Schur complement onto the floating base, upper triangle + right-hand side (27 entries per agent), mirrored
    RS_LANE_LOOP(e, 54) {
        const int a = e >= 27 ? 1 : 0, t = e - 27 * a;
        // t -> (r, cc) with r <= cc <= 6: rows of length 7, 6, 5, 4, 3, 2 start at t = 0, 7, 13, 18, 22, 25
        const int r = (t >= 7) + (t >= 13) + (t >= 18) + (t >= 22) + (t >= 25);
        const int cc = t - (7 * r - ((r * (r - 1)) >> 1)) + r;
        float acc;
        const int l0 = c.leg0(a), va = c.vadr(a);
        if (cc < 6) {
            acc = H[AR::A0 + a * 36 + r * 6 + cc];
            RS_UNROLL1
            for (int l = 0; l < c.L(a); l++) {
                const float* W = H + AR::W0 + 12 * (l0 + l); const float* B = H + AR::B0 + 12 * (l0 + l);
                acc -= W[2 * r] * B[2 * cc] + W[2 * r + 1] * B[2 * cc + 1];
            }
            H[AR::S0 + a * 42 + r * 7 + cc] = acc; H[AR::S0 + a * 42 + cc * 7 + r] = acc;
        } else {
            acc = -s.d[va + r];
            RS_UNROLL1
            for (int l = 0; l < c.L(a); l++) {
                const float* Y = H + AR::Y0 + 2 * (l0 + l); const float* B = H + AR::B0 + 12 * (l0 + l);
                acc -= B[2 * r] * Y[0] + B[2 * r + 1] * Y[1];
            }
            H[AR::S0 + a * 42 + r * 7 + 6] = acc;
        }
    }
    RS_SYNC();
    // 6x6 Gauss-Jordan per agent, row per lane (12 lanes)
    for (int k = 0; k < 6; k++) {
        RS_LANE_LOOP(j, 12) {
            const int a = j >= 6 ? 1 : 0, jr = j - 6 * a;
            if (jr != k) {
                const float* prow = H + AR::S0 + a * 42 + k * 7;
                float* jrow = H + AR::S0 + a * 42 + jr * 7;
                const float f = jrow[k] * RS_RCP(fmaxf(prow[k], 1e-12f));
                for (int cc = k + 1; cc < 7; cc++) jrow[cc] = fmaf(-f, prow[cc], jrow[cc]);
            }
        }
        RS_SYNC();
    }
    RS_LANE_LOOP(j, 12) {
        const int a = j >= 6 ? 1 : 0, jr = j - 6 * a;
        float* row = H + AR::S0 + a * 42 + jr * 7;
        const float v = row[6] * RS_RCP(fmaxf(row[jr], 1e-12f));
        row[6] = v; s.d[c.vadr(a) + jr] = v;
    }
    RS_SYNC();
#endif
    RS_LANE_LOOP(g, S::LT) {
        const int a = c.agent_of_leg(g), dh = c.hipdof(g);
        const float* W = H + AR::W0 + 12 * g;
        float dh_ = H[AR::Y0 + 2 * g], da_ = H[AR::Y0 + 2 * g + 1];
        for (int k = 0; k < 6; k++) { const float dt = s.d[c.vadr(a) + k]; dh_ -= W[2 * k] * dt; da_ -= W[2 * k + 1] * dt; }
        s.d[dh] = dh_; s.d[dh + 1] = da_;
    }
    RS_SYNC();
}

// ---- inter-agent contacts as a low-rank correction of the arrowhead solve -------------------------------------------------
// A contact between the two agents couples them: H = H0 + U^T W U with H0 the arrowhead matrix of everything else, U the three
// direction Jacobians (n, t1, t2; 28-vectors with the 8 chain dofs of each touched body) of the m such contacts with active rows
// and W the 3 x 3 weights of their active pyramid rows.  The dense 28 x 29 elimination this used to take made the pair the
// straggler of its block in every iteration (and streamed 4.5 k instructions through the instruction cache for one warp).  Here:
//   X = H0^-1 [ -g | U^T ]   one arrowhead solve with 1 + 3 m right-hand sides (arrow_solve_multi),
//   G = D^-1 + P^T (U H0^-1 U^T) P   over the 4 m pyramid rows (P: row -> n +- mu t; inactive rows are identity rows), SPD,
//   d = y - Z P G^-1 P^T (U y)   (Woodbury).
// m <= 2 takes this path; more simultaneous inter-agent contacts fall back to the dense elimination.
#ifndef RS_WOODBURY
#define RS_WOODBURY 1
#endif
// H0 X = R in place for NR vectors of NV floats (H0 = the compact arrowhead matrix of build_H_arrow)
template <int LA, int LB, int NR>
RS_HD void arrow_solve_multi(Ctx<LA, LB>& c, float* R) {
    typedef Slab<LA, LB> S;
    typedef Arrow<LA, LB> AR;
    S& s = *c.s;
    float* H = s.H;
    RS_LANE_LOOP(g, S::LT) {
        const float* Dg = H + AR::D0 + 3 * g; const float* B = H + AR::B0 + 12 * g;
        float* W = H + AR::W0 + 12 * g;
        const float idet = RS_RCP(fmaxf(Dg[0] * Dg[2] - Dg[1] * Dg[1], 1e-20f));
        const float i00 = Dg[2] * idet, i01 = -Dg[1] * idet, i11 = Dg[0] * idet;
        for (int k = 0; k < 6; k++) { W[2 * k] = B[2 * k] * i00 + B[2 * k + 1] * i01; W[2 * k + 1] = B[2 * k] * i01 + B[2 * k + 1] * i11; }
    }
    RS_LANE_LOOP(i, S::LT * NR) {       // Y = Dg^-1 rhs_g, over the leg entries of the vector
        const int g = i / NR, j = i - g * NR;
        const float* Dg = H + AR::D0 + 3 * g;
        const float idet = RS_RCP(fmaxf(Dg[0] * Dg[2] - Dg[1] * Dg[1], 1e-20f));
        const float i00 = Dg[2] * idet, i01 = -Dg[1] * idet, i11 = Dg[0] * idet;
        float* v = R + j * S::NV + c.hipdof(g);
        const float gh = v[0], ga = v[1];
        v[0] = i00 * gh + i01 * ga; v[1] = i01 * gh + i11 * ga;
    }
    RS_SYNC();
#if defined(__CUDA_ARCH__)
    {   // Schur complement onto the two floating bases, 6 x (6 + NR) per agent with the rows in registers (lane 6 a + r)
        const int lane = threadIdx.x & 31, a = lane >= 6 ? 1 : 0, r = lane - 6 * a;
        const bool own = lane < 12;
        float row[6 + NR];
#pragma unroll
        for (int cc = 0; cc < 6 + NR; cc++) row[cc] = 0.f;
        if (own) {
            const int l0 = c.leg0(a), va = c.vadr(a);
#pragma unroll
            for (int cc = 0; cc < 6; cc++) row[cc] = H[AR::A0 + a * 36 + r * 6 + cc];
#pragma unroll
            for (int j = 0; j < NR; j++) row[6 + j] = R[j * S::NV + va + r];
            RS_UNROLL1
            for (int l = 0; l < c.L(a); l++) {
                const float* W = H + AR::W0 + 12 * (l0 + l); const float* B = H + AR::B0 + 12 * (l0 + l);
                const float w0 = W[2 * r], w1 = W[2 * r + 1], b0 = B[2 * r], b1 = B[2 * r + 1];
#pragma unroll
                for (int cc = 0; cc < 6; cc++) row[cc] -= w0 * B[2 * cc] + w1 * B[2 * cc + 1];
                const float* Y = R + va + 6 + 2 * l;
#pragma unroll
                for (int j = 0; j < NR; j++) row[6 + j] -= b0 * Y[j * S::NV] + b1 * Y[j * S::NV + 1];
            }
        }
        float diag = 1.f;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int src = own ? 6 * a + k : lane;
            const float pk = __shfl_sync(0xffffffffu, row[k], src);
            const float f = (r == k) ? 0.f : row[k] * RS_RCP(fmaxf(pk, 1e-12f));
            if (r == k) diag = pk;
#pragma unroll
            for (int cc = k + 1; cc < 6 + NR; cc++) row[cc] = fmaf(-f, __shfl_sync(0xffffffffu, row[cc], src), row[cc]);
        }
        if (own) {
            const float id = RS_RCP(fmaxf(diag, 1e-12f));
#pragma unroll
            for (int j = 0; j < NR; j++) R[j * S::NV + c.vadr(a) + r] = row[6 + j] * id;
        }
    }
#else
    for (int a = 0; a < 2; a++) {       // host emulation: the same elimination on a local copy
        float Sx[6][6 + NR];
        const int l0 = c.leg0(a), va = c.vadr(a);
        for (int r = 0; r < 6; r++) {
            for (int cc = 0; cc < 6; cc++) Sx[r][cc] = H[AR::A0 + a * 36 + r * 6 + cc];
            for (int j = 0; j < NR; j++) Sx[r][6 + j] = R[j * S::NV + va + r];
            for (int l = 0; l < c.L(a); l++) {
                const float* W = H + AR::W0 + 12 * (l0 + l); const float* B = H + AR::B0 + 12 * (l0 + l);
                for (int cc = 0; cc < 6; cc++) Sx[r][cc] -= W[2 * r] * B[2 * cc] + W[2 * r + 1] * B[2 * cc + 1];
                const float* Y = R + va + 6 + 2 * l;
                for (int j = 0; j < NR; j++) Sx[r][6 + j] -= B[2 * r] * Y[j * S::NV] + B[2 * r + 1] * Y[j * S::NV + 1];
            }
        }
        for (int k = 0; k < 6; k++) {
            float prow[6 + NR];
            for (int cc = 0; cc < 6 + NR; cc++) prow[cc] = Sx[k][cc];
            for (int r = 0; r < 6; r++) {
                if (r == k) continue;
                const float f = Sx[r][k] * RS_RCP(fmaxf(prow[k], 1e-12f));
                for (int cc = k + 1; cc < 6 + NR; cc++) Sx[r][cc] = fmaf(-f, prow[cc], Sx[r][cc]);
            }
        }
        for (int r = 0; r < 6; r++) for (int j = 0; j < NR; j++) R[j * S::NV + va + r] = Sx[r][6 + j] * RS_RCP(fmaxf(Sx[r][r], 1e-12f));
    }
#endif
    RS_SYNC();
    RS_LANE_LOOP(i, S::LT * NR) {       // back-substitution: d_g = Y - (B Dg^-1)^T d_t
        const int g = i / NR, j = i - g * NR, a = c.agent_of_leg(g);
        const float* W = H + AR::W0 + 12 * g;
        float* v = R + j * S::NV + c.hipdof(g);
        const float* dt = R + j * S::NV + c.vadr(a);
        float dh_ = v[0], da_ = v[1];
        for (int k = 0; k < 6; k++) { dh_ -= W[2 * k] * dt[k]; da_ -= W[2 * k + 1] * dt[k]; }
        v[0] = dh_; v[1] = da_;
    }
    RS_SYNC();
}

// s.d = -H^-1 s.d with H = (arrowhead H0 already in s.H) + the m = s.wood_m inter-agent contacts s.wood_k[]
template <int LA, int LB, int M>
RS_HD void woodbury_solve(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    enum { K = 3 * M, NR = 1 + K, NA = 4 * M };
    // scratch (alias zone, dead between the gradient and the end of the linear solve; build_H_arrow is done with it)
    float* R = &s.tw[0][0];                    // [NR][NV]: -g, then the rows of U (afterwards y and Z)
    float* UC = R + NR * S::NV;                // [M][16][4]: dof index, J n, J t1, J t2 of the 2 x 8 chain dofs (A side negated)
    float* SM = UC + 64 * M;                   // [K][K + 1]: U Z | U y
    float* G = s.H + Arrow<LA, LB>::S0;        // [NA][NA + 1]  (the tail of the H array, past the compact arrowhead matrix, W and Y)
    static_assert(NR * S::NV + 64 * M + K * (K + 1) <= (int)S::ALIAS, "woodbury scratch must fit in the alias zone");
    static_assert((int)Arrow<LA, LB>::S0 + NA * (NA + 1) <= (int)S::HDED, "woodbury G must fit behind the compact arrowhead matrix");
    RS_LANE_LOOP(i, 16 * M) {
        const int q = i >> 4, e = i & 15, k = s.wood_k[q];
        const int b = e < 8 ? s.bA(k) : s.bB(k);
        const V3 p = ld3(s.cpos[k]), n = ld3(s.cn[k]), t1 = ld3(s.ct1[k]);
        int idx;
        const V3 jc = (e < 8 ? -1.f : 1.f) * side_col(c, b, e & 7, p, false, &idx);
        float* o = UC + 4 * i;
        o[0] = (float)idx; o[1] = dot(jc, n); o[2] = dot(jc, t1); o[3] = dot(jc, cross(n, t1));
    }
    RS_LANE_LOOP(i, NR * S::NV) { R[i] = i < S::NV ? -s.d[i] : 0.f; }
    RS_SYNC();
    RS_LANE_LOOP(i, 16 * M) {
        const int q = i >> 4;
        const float* o = UC + 4 * i;
        const int idx = (int)o[0];
        if (idx >= 0) { for (int dir = 0; dir < 3; dir++) R[(1 + 3 * q + dir) * S::NV + idx] = o[1 + dir]; }
    }
    RS_SYNC();
    arrow_solve_multi<LA, LB, NR>(c, R);
    // U Z and U y: entry (i, j) = row i of U . solved vector j (j = K: y)
    RS_LANE_LOOP(t, K * (K + 1)) {
        const int i = t / (K + 1), j = t - i * (K + 1), q = i / 3, dir = i - 3 * q;
        const float* x = R + (j < K ? 1 + j : 0) * S::NV;
        const float* o = UC + 64 * q;
        float acc = 0.f;
        for (int e = 0; e < 16; e++) { const int idx = (int)o[4 * e]; acc += idx >= 0 ? o[4 * e + 1 + dir] * x[idx] : 0.f; }      // (unrolled: the loads overlap)
        SM[t] = acc;
    }
    RS_SYNC();
    // G = D^-1 + P^T S P over the pyramid rows (row (q, r): n + sg t1|t2); inactive rows are identity rows with zero right-hand side
    RS_LANE_LOOP(t, NA * (NA + 1)) {
        const int al = t / (NA + 1), be = t - al * (NA + 1);
        const int qa = al >> 2, ra = al & 3, acta = (s.cact(s.wood_k[qa]) >> ra) & 1;
        const int ia0 = 3 * qa, ia1 = 3 * qa + 1 + (ra >> 1);
        const float sa = (ra & 1) ? -RS_MU : RS_MU;
        float val;
        if (be == NA) val = acta ? SM[ia0 * (K + 1) + K] + sa * SM[ia1 * (K + 1) + K] : 0.f;
        else {
            const int qb = be >> 2, rb = be & 3, actb = (s.cact(s.wood_k[qb]) >> rb) & 1;
            const int ib0 = 3 * qb, ib1 = 3 * qb + 1 + (rb >> 1);
            const float sb = (rb & 1) ? -RS_MU : RS_MU;
            if (acta && actb)
                val = SM[ia0 * (K + 1) + ib0] + sb * SM[ia0 * (K + 1) + ib1] + sa * SM[ia1 * (K + 1) + ib0] + sa * sb * SM[ia1 * (K + 1) + ib1]
                      + (al == be ? RS_DIV(1.f, s.cD[s.wood_k[qa]]) : 0.f);
            else val = al == be ? 1.f : 0.f;
        }
        G[t] = val;
    }
    RS_SYNC();
    RS_UNROLL1
    for (int k = 0; k < NA; k++) {       // Gauss-Jordan, lane = row (SPD: no pivoting)
        RS_LANE_LOOP(al, NA) {
            if (al != k) {
                const float* prow = G + k * (NA + 1); float* arow = G + al * (NA + 1);
                const float f = arow[k] * RS_RCP(fmaxf(prow[k], 1e-20f));
                for (int cc = k + 1; cc <= NA; cc++) arow[cc] = fmaf(-f, prow[cc], arow[cc]);
            }
        }
        RS_SYNC();
    }
    // d = y - Z P w
    RS_LANE_LOOP(i, S::NV) {
        float acc = R[i];
        RS_UNROLL1
        for (int q = 0; q < M; q++) {
            float w[4];
            for (int r = 0; r < 4; r++) w[r] = G[(4 * q + r) * (NA + 1) + NA] * RS_RCP(fmaxf(G[(4 * q + r) * (NA + 1) + 4 * q + r], 1e-20f));
            acc -= (w[0] + w[1] + w[2] + w[3]) * R[(1 + 3 * q) * S::NV + i] + RS_MU * (w[0] - w[1]) * R[(2 + 3 * q) * S::NV + i] + RS_MU * (w[2] - w[3]) * R[(3 + 3 * q) * S::NV + i];
        }
        s.d[i] = acc;
    }
    RS_SYNC();
}

// out = M * vec (+ add) using the arrowhead structure of M
template <int LA, int LB>
RS_HD void mat_vec(Ctx<LA, LB>& c, const float* vec, float* out, const float* add) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    RS_LANE_LOOP(i, S::NV) {
        const int a = i >= S::NVA ? 1 : 0, va = c.vadr(a), k = i - va;
        float acc = add ? add[i] : 0.f;
        if (k < 6) {
            for (int j = 0; j < 6; j++) acc += s.Mr[a][k * 6 + j] * vec[va + j];
            RS_UNROLL1
            for (int l = 0; l < c.L(a); l++) {
                const int g = c.leg0(a) + l;
                acc += s.Mc[g][2 * k] * vec[va + 6 + 2 * l] + s.Mc[g][2 * k + 1] * vec[va + 7 + 2 * l];
            }
        } else {
            const int l = (k - 6) >> 1, isank = (k - 6) & 1, g = c.leg0(a) + l;
            for (int kk = 0; kk < 6; kk++) acc += s.Mc[g][2 * kk + isank] * vec[va + kk];
            const float vh = vec[va + 6 + 2 * l], vk = vec[va + 7 + 2 * l];
            acc += isank ? s.Ml[g][1] * vh + s.Ml[g][2] * vk : s.Ml[g][0] * vh + s.Ml[g][1] * vk;
        }
        out[i] = acc;
    }
    RS_SYNC();
}
// dense copy of M (tests / debugging)
template <int LA, int LB>
RS_HD void dense_M(Ctx<LA, LB>& c, float* out /*[NV*NV]*/) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    for (int i = 0; i < S::NV * S::NV; i++) out[i] = 0.f;
    for (int a = 0; a < 2; a++) {
        int va = c.vadr(a);
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) out[(va + i) * S::NV + va + j] = s.Mr[a][i * 6 + j];
        for (int l = 0; l < c.L(a); l++) {
            int g = c.leg0(a) + l, dh = va + 6 + 2 * l;
            for (int k = 0; k < 6; k++) for (int w = 0; w < 2; w++) { out[(va + k) * S::NV + dh + w] = s.Mc[g][2 * k + w]; out[(dh + w) * S::NV + va + k] = s.Mc[g][2 * k + w]; }
            out[dh * S::NV + dh] = s.Ml[g][0]; out[dh * S::NV + dh + 1] = s.Ml[g][1]; out[(dh + 1) * S::NV + dh] = s.Ml[g][1]; out[(dh + 1) * S::NV + dh + 1] = s.Ml[g][2];
        }
    }
}

// phi'(alpha) and phi''(alpha) of the line search (uniform result)
template <int LA, int LB>
RS_HD void dphi(Ctx<LA, LB>& c, float alpha, float p0, float p1, float* d1, float* d2) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    const int nc4 = 4 * s.ncon;
    const float* jar = &s.cjar[0][0]; const float* jdv = &s.cjd[0][0];
#if defined(__CUDA_ARCH__)
    float a1 = 0.f, a2 = 0.f;
    RS_UNROLL1
    for (int i = threadIdx.x & 31; i < nc4; i += 32) {
        float jd = jdv[i], j = jar[i] + alpha * jd;
        if (j < 0.f) { float D = s.cD[i >> 2]; a1 += D * j * jd; a2 += D * jd * jd; }
    }
    RS_UNROLL1
    for (int i = threadIdx.x & 31; i < S::NU; i += 32) {
        float jd = s.ljd[i], j = s.ljar[i] + alpha * jd;
        if (s.lsgn[i] != 0.f && j < 0.f) { float D = s.lD[i]; a1 += D * j * jd; a2 += D * jd * jd; }
    }
    RS_UNROLL1
    for (int o = 16; o; o >>= 1) { a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
    *d1 = p0 + alpha * p1 + a1; *d2 = p1 + a2;
#else
    float a1 = 0.f, a2 = 0.f;
    for (int i = 0; i < nc4; i++) {
        float jd = jdv[i], j = jar[i] + alpha * jd;
        if (j < 0.f) { float D = s.cD[i >> 2]; a1 += D * j * jd; a2 += D * jd * jd; }
    }
    for (int i = 0; i < S::NU; i++) {
        float jd = s.ljd[i], j = s.ljar[i] + alpha * jd;
        if (s.lsgn[i] != 0.f && j < 0.f) { float D = s.lD[i]; a1 += D * j * jd; a2 += D * jd * jd; }
    }
    *d1 = p0 + alpha * p1 + a1; *d2 = p1 + a2;
#endif
}

// p0 = d . r and p1 = d . (M d) of the line search (uniform result)
template <int LA, int LB>
RS_HD void ls_dots(Ctx<LA, LB>& c, float* p0, float* p1) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    float a = 0.f, b = 0.f;
    RS_LANE_LOOP(i, S::NV) { a += s.d[i] * s.r[i]; b += s.d[i] * s.Md[i]; }
#if defined(__CUDA_ARCH__)
    RS_UNROLL1
    for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
#endif
    *p0 = a; *p1 = b;
}
// row i of the line search: contact rows first (4 per contact), then the limit rows; false = the row does not exist
template <int LA, int LB>
RS_HD bool ls_row(const Slab<LA, LB>& s, int i, int nc4, float* jar, float* jd, float* D) {
    typedef Slab<LA, LB> S;
    if (i < nc4) { *jar = (&s.cjar[0][0])[i]; *jd = (&s.cjd[0][0])[i]; *D = s.cD[i >> 2]; return true; }
    const int j = i - nc4;
    *jar = s.ljar[j]; *jd = s.ljd[j]; *D = s.lD[j];
    return s.lsgn[j] != 0.f;
}
// Exact line search along s.d: phi'(alpha) = p0 + alpha p1 + sum_rows D (jar + alpha jd) jd [jar + alpha jd < 0] is increasing and
// piecewise linear with one breakpoint -jar / jd per row.  Every lane takes breakpoints and evaluates phi' there (a loop over the
// rows; the row data are warp-uniform shared-memory reads), the warp keeps the largest breakpoint with phi' < 0 and the smallest
// with phi' >= 0, and one Newton step from the middle of that linear piece lands on the root: one pass instead of the up to 20
// bracketing / safeguarded-Newton evaluations (each a warp reduction) of round 1.
template <int LA, int LB>
RS_HD float line_search(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    float p0, p1;
    ls_dots(c, &p0, &p1);
    const int nc4 = 4 * s.ncon, R = nc4 + S::NU;
    float lo = 0.f, hi = 3.0e38f;
    RS_LANE_LOOP(i, R) {
        float jar, jd, D;
        if (!ls_row(s, i, nc4, &jar, &jd, &D) || jd == 0.f) continue;
        const float a = -RS_DIV(jar, jd);
        if (!(a > 0.f && a < 1.0e30f)) continue;
        float acc = p0 + a * p1;
        RS_UNROLL1
        for (int r = 0; r < R; r++) {
            float jr, jdr, Dr;
            if (!ls_row(s, r, nc4, &jr, &jdr, &Dr)) continue;
            const float j = jr + a * jdr;
            if (j < 0.f) acc += Dr * j * jdr;
        }
        if (acc < 0.f) lo = fmaxf(lo, a); else hi = fminf(hi, a);
    }
#if defined(__CUDA_ARCH__)
    RS_UNROLL1
    for (int o = 16; o; o >>= 1) { lo = fmaxf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fminf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
#endif
    // a point inside the linear piece (lo, hi) that holds the root, as close to lo as is safely past that breakpoint: the Newton
    // step from it is exact, and evaluating near the small end keeps am - d1 / d2 free of cancellation when hi is astronomically
    // large (a row with a tiny jd puts a breakpoint at 1e9)
    const float am = lo + fminf(0.5f * (hi - lo), 1e-3f * fmaxf(lo, 1.f));
    float d1, d2;
    dphi(c, am, p0, p1, &d1, &d2);
    float alpha = am - RS_DIV(d1, d2);
    alpha = fminf(fmaxf(alpha, lo), hi);
#ifdef RS_LS_DEBUG
    { float e1, e2; dphi(c, alpha, p0, p1, &e1, &e2); float z1, z2; dphi(c, 0.f, p0, p1, &z1, &z2);
      printf("  ls: R %d p0 %g p1 %g phi'(0) %g lo %g hi %g am %g d1 %g d2 %g alpha %g phi'(alpha) %g\n", R, p0, p1, z1, lo, hi, am, d1, d2, alpha, e1); }
#endif
    return alpha;
}

// ------------------------------------------------------------------------------------------
// mj_fwdConstraint: primal Newton, exact line search, warm start from s.x.  Split in three so that the warps of a block can
// re-align per Newton ITERATION (simulate() below): solve_first (residuals at the warm start), solve_iter (one iteration,
// returns true when the active set survived the full step, i.e. the point is the exact optimum), solve_finish.
// ------------------------------------------------------------------------------------------
#ifndef RS_ARROW
#define RS_ARROW 1     // 0: always use the generic (block-diagonal / dense) assembly and elimination
#endif
#ifndef RS_ACC
#define RS_ACC(i)
#endif
#ifndef RS_SOLVE_TRACE
#define RS_SOLVE_TRACE(s, it)     // host-emulation analysis hook (tools), empty in every build of the product
#endif
// residuals at the warm start x0: jar = J x0 - aref, r = M x0 - tau, and the predicted active sets
template <int LA, int LB>
RS_HD void solve_first(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    RS_ACC(5);
    // residual pass: rows of J (x0 + B v), see make_constraints (s.d is free until the first Newton direction)
    const float Bc = solref_damping(c.h);
    RS_LANE_LOOP(i, S::NV) { s.d[i] = s.x[i] + Bc * s.v[i]; }
    RS_SYNC();
    twists(c, s.d);
    rows_of(c, s.d, s.cjar, s.ljar);
    RS_ACC(3);
    mat_vec(c, s.x, s.r, s.r);                         // r = M x0 - qfrc_smooth
    RS_LANE_LOOP(k, s.ncon) {
        // rows of the same geom pair at the previous evaluation's solution predict this evaluation's active rows
        int bits = 16;
        const int key = s.ckey(k);
        for (int q = 0; q < s.nprev; q++) if ((s.cprev[q] >> 4) == key) bits = s.cprev[q] & 15;
        int sign = 0;
        RS_UNROLL1
        for (int r = 0; r < 4; r++) { s.cjar[k][r] -= s.caref[k][r]; if (s.cjar[k][r] < 0.f) sign |= 1 << r; }
        s.set_cact(k, bits < 16 ? bits : sign);
    }
#if defined(__CUDA_ARCH__)
    {   // one vote instead of a zeroing pass, a barrier and per-lane atomics
        const int j = threadIdx.x & 31, pvd = s.pvalid;
        bool neg = false;
        if (j < S::NU) { const float v = s.lsgn[j] != 0.f ? s.ljar[j] - s.laref[j] : 1.f; s.ljar[j] = v; neg = v < 0.f && !((pvd >> j) & 1); }
        const unsigned b = __ballot_sync(0xffffffffu, neg);
        if (j == 0) s.lmask = (s.pmask & pvd) | (int)b;
    }
    RS_SYNC();
#else
    RS_LANE_LOOP(j, S::NU) { s.ljar[j] = s.lsgn[j] != 0.f ? s.ljar[j] - s.laref[j] : 1.f; }
    if (RS_LANE0) s.lmask = s.pmask & s.pvalid;
    RS_SYNC();
    RS_LANE_LOOP(j, S::NU) { if (s.ljar[j] < 0.f && !((s.pvalid >> j) & 1)) RS_ATOMIC_OR(&s.lmask, 1 << j); }
    RS_SYNC();
#endif
}

// one Newton iteration from the current point (s.x, residuals s.r / s.cjar / s.ljar, active sets); true = converged
template <int LA, int LB>
RS_HD bool solve_iter(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    RS_ACC(5);
    // coupling that matters for THIS iteration's H: only contacts with active rows contribute J^T D J, so an inter- or
    // intra-agent contact that is inside the margin but not loaded does not force the dense paths
#if defined(__CUDA_ARCH__)
    {
        unsigned cpl = 0;
        RS_LANE_LOOP(k, s.ncon) {
            const int bA = s.bA(k), bB = s.bB(k);
            if (bA >= 0 && s.cact(k) != 0)
                cpl |= ((bA < 2 ? bA : c.agent_of_leg((bA - 2) % S::LT)) != (bB < 2 ? bB : c.agent_of_leg((bB - 2) % S::LT))) ? 1u : 2u;
        }
        cpl = __reduce_or_sync(0xffffffffu, cpl);
        if (RS_LANE0) s.coupled = (int)cpl;
    }
    RS_SYNC();
#else
    if (RS_LANE0) s.coupled = 0;
    RS_SYNC();
    RS_LANE_LOOP(k, s.ncon) {
        const int bA = s.bA(k), bB = s.bB(k);
        if (bA >= 0 && s.cact(k) != 0)
            RS_ATOMIC_OR(&s.coupled, ((bA < 2 ? bA : c.agent_of_leg((bA - 2) % S::LT)) != (bB < 2 ? bB : c.agent_of_leg((bB - 2) % S::LT))) ? 1 : 2);
    }
    RS_SYNC();
#endif
    if (RS_UNLIKELY(s.coupled == 1)) {      // inter-agent contacts only: list the active ones (fixed order) for the low-rank path
        if (RS_LANE0) {
            int m = 0;
            for (int k = 0; k < s.ncon; k++) if (s.bA(k) >= 0 && s.cact(k) != 0) { if (m < 2) s.wood_k[m] = k; m++; }
            s.wood_m = m;
        }
        RS_SYNC();
    }
    jt_forces(c);
    RS_LANE_LOOP(i, S::NV) { s.d[i] = s.r[i] - s.jtf[i]; }     // gradient
    RS_SYNC();
    RS_ACC(0);
    const bool wood = RS_WOODBURY && RS_ARROW && s.coupled == 1 && s.wood_m <= 2;
    if (RS_LIKELY(RS_ARROW && (s.coupled == 0 || wood))) build_H_arrow(c); else build_H(c);
    RS_ACC(1);
    if (RS_LIKELY(RS_ARROW && s.coupled == 0)) arrow_solve(c);      // s.d = -H^-1 grad
    else if (wood) { if (s.wood_m == 1) woodbury_solve<LA, LB, 1>(c); else woodbury_solve<LA, LB, 2>(c); }
    else chol_solve(c);
    RS_ACC(2);
    twists(c, s.d);
    rows_of(c, s.d, s.cjd, s.ljd);
    RS_ACC(3);
    // does the active set survive the full step?
    if (RS_LANE0) s.same = 1;
    RS_SYNC();
    RS_LANE_LOOP(k, s.ncon) {
        const int act = s.cact(k);
        for (int r = 0; r < 4; r++) if ((((act >> r) & 1) != 0) != (s.cjar[k][r] + s.cjd[k][r] < 0.f)) s.same = 0;
    }
    RS_LANE_LOOP(j, S::NU) {
        if (s.lsgn[j] != 0.f && ((((s.lmask >> j) & 1) != 0) != (s.ljar[j] + s.ljd[j] < 0.f))) s.same = 0;
    }
    RS_SYNC();
    const int same = s.same;
    int predicted = 0;
    if (!same) {
        // the full step leaves the active set: the iteration goes on from x + alpha d and needs M d (for r and the line search).
        // Was the set used for this iteration the sign set at the current point?  (not when it came from the prediction)
        if (RS_LANE0) s.pvalid = 0;
        RS_SYNC();
        RS_LANE_LOOP(j, S::NU) { if (s.lsgn[j] != 0.f && ((((s.lmask >> j) & 1) != 0) != (s.ljar[j] < 0.f))) s.pvalid = 1; }
        RS_LANE_LOOP(k, s.ncon) {
            const int act = s.cact(k);
            for (int r = 0; r < 4; r++) if ((((act >> r) & 1) != 0) != (s.cjar[k][r] < 0.f)) s.pvalid = 1;
        }
        RS_SYNC();
        predicted = s.pvalid;
        mat_vec(c, s.d, s.Md, (const float*)0);
    }
    float alpha = 1.f;
    if (RS_UNLIKELY(!same && !predicted)) alpha = line_search(c);
#ifdef RS_LS_DEBUG
    printf(" iter: same %d predicted %d alpha %g ncon %d coupled %d\n", same, predicted, alpha, s.ncon, s.coupled);
#endif
    if (same) { RS_LANE_LOOP(i, S::NV) { s.x[i] += s.d[i]; } }       // converged: r is not needed any more, M d was never formed
    else { RS_LANE_LOOP(i, S::NV) { s.x[i] += alpha * s.d[i]; s.r[i] += alpha * s.Md[i]; } }
    RS_LANE_LOOP(k, s.ncon) {
        int sign = 0;
        for (int r = 0; r < 4; r++) { s.cjar[k][r] += alpha * s.cjd[k][r]; if (s.cjar[k][r] < 0.f) sign |= 1 << r; }
        s.set_cact(k, sign);             // sign set at the new point
    }
#if defined(__CUDA_ARCH__)
    {
        const int j = threadIdx.x & 31;
        bool neg = false;
        if (j < S::NU && s.lsgn[j] != 0.f) { const float v = s.ljar[j] + alpha * s.ljd[j]; s.ljar[j] = v; neg = v < 0.f; }
        const unsigned b = __ballot_sync(0xffffffffu, neg);
        if (j == 0) s.lmask = (int)b;                                                                      // sign set at the new point
    }
    RS_SYNC();
#else
    RS_LANE_LOOP(j, S::NU) { if (s.lsgn[j] != 0.f) s.ljar[j] += alpha * s.ljd[j]; }
    if (RS_LANE0) s.lmask = 0;
    RS_SYNC();
    RS_LANE_LOOP(j, S::NU) { if (s.lsgn[j] != 0.f && s.ljar[j] < 0.f) RS_ATOMIC_OR(&s.lmask, 1 << j); }     // sign set at the new point
    RS_SYNC();
#endif
    RS_ACC(4);
    return same != 0;
}

// the active sets at the solution become the next evaluation's prediction; diagnostics
template <int LA, int LB>
RS_HD void solve_finish(Ctx<LA, LB>& c, int it, bool conv) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    RS_SOLVE_TRACE(s, it);
    RS_LANE_LOOP(k, s.ncon) { s.cprev[k] = (unsigned short)((s.ckey(k) << 4) | s.cact(k)); }
    if (RS_LANE0) s.nprev = s.ncon;
    if (RS_LANE0) { s.niter = it; s.tot_iter += it; s.tot_coupled += s.coupled & 1; s.tot_ncon += s.ncon; if (it > s.max_iter) s.max_iter = it; if (!conv) s.status |= RS_STATUS_NEWTON_MAXIT; }
    RS_SYNC();
}

template <int LA, int LB>
RS_HD void solve(Ctx<LA, LB>& c) {
    solve_first(c);
    bool conv = false;
    int it = 0;
    while (!conv && it < c.max_newton) { conv = solve_iter(c); it++; }
    solve_finish(c, it, conv);
}

// one forward evaluation: qacc(q, v) into s.x
#ifndef RS_EVAL_SYNC
#define RS_EVAL_SYNC()
#endif
#ifndef RS_CLOCK_BEGIN
#define RS_CLOCK_BEGIN()      // developer instrumentation: busy cycles of a warp between the block-wide re-alignments
#define RS_CLOCK_END()
#define RS_CLOCK_MARK(i)
#endif
#ifndef RS_SUBSTEP_SYNC
#define RS_SUBSTEP_SYNC()
#endif
// everything of a forward evaluation in front of the Newton iterations
template <int LA, int LB>
RS_HD void eval_begin(Ctx<LA, LB>& c) {
    RS_CLOCK_BEGIN();
    fk(c);
    dynamics(c);
    RS_CLOCK_MARK(0);
    collide(c);
    RS_CLOCK_MARK(1);
    make_constraints(c);
    RS_CLOCK_MARK(2);
    solve_first(c);
}
template <int LA, int LB>
RS_HD void forward(Ctx<LA, LB>& c) {
    RS_EVAL_SYNC();     // optional block-wide re-alignment of the warps (instruction-cache locality)
    eval_begin(c);
    bool conv = false;
    int it = 0;
    while (!conv && it < c.max_newton) { conv = solve_iter(c); it++; }
    solve_finish(c, it, conv);
    RS_CLOCK_MARK(3);
    RS_CLOCK_END();
}

// mj_integratePos from q0 with velocity vel over dt into s.q
template <int LA, int LB>
RS_HD void integrate_pos(Ctx<LA, LB>& c, const float* vel, float dt) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    RS_LANE_LOOP(a, 2) {
        int qa = c.qadr(a), va = c.vadr(a);
        for (int k = 0; k < 3; k++) s.q[qa + k] = s.q0[qa + k] + dt * vel[va + k];
        V3 w = v3(vel[va + 3], vel[va + 4], vel[va + 5]);
        float wn;
        V3 ax = normalized(w, &wn);
        float ang = dt * wn, sh, ch;
        if (wn < 1e-12f) ang = 0.f;
        RS_SINCOS(0.5f * ang, &sh, &ch);
        float bw = ch, bx = sh * ax.x, by = sh * ax.y, bz = sh * ax.z;
        float aw = s.q0[qa + 3], ax_ = s.q0[qa + 4], ay = s.q0[qa + 5], az = s.q0[qa + 6];
        float w_ = aw*bw - ax_*bx - ay*by - az*bz, x_ = aw*bx + ax_*bw + ay*bz - az*by;
        float y_ = aw*by - ax_*bz + ay*bw + az*bx, z_ = aw*bz + ax_*by - ay*bx + az*bw;
        float n2 = w_*w_ + x_*x_ + y_*y_ + z_*z_, inv = n2 > 1e-24f ? RS_RSQRT(n2) : 1.f;
        s.q[qa + 3] = w_ * inv; s.q[qa + 4] = x_ * inv; s.q[qa + 5] = y_ * inv; s.q[qa + 6] = z_ * inv;
    }
    RS_LANE_LOOP(g, S::LT) {
        int qh = c.hipq(g), dh = c.hipdof(g);
        s.q[qh] = s.q0[qh] + dt * vel[dh]; s.q[qh + 1] = s.q0[qh + 1] + dt * vel[dh + 1];
    }
    RS_SYNC();
}

// do_simulation: nsub x mj_step with RK4 (mujoco_env.py:125-129; mj_RungeKutta N=4)
// RK4 bookkeeping after the evaluation of stage `st` (0..3) of a substep: accumulate, move to the next stage's state or finish the substep
template <int LA, int LB>
RS_HD void rk_after_eval(Ctx<LA, LB>& c, int st) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    const float h = c.h;
    if (st == 0) {   // mj_forward normalised qpos in place: carry that into the substep origin
        RS_LANE_LOOP(a, 2) { for (int k = 3; k < 7; k++) s.q0[c.qadr(a) + k] = s.q[c.qadr(a) + k]; }
    }
    const float B = (st == 0 || st == 3) ? (1.f / 6.f) : (1.f / 3.f);
    const float A = (st == 2) ? 1.f : 0.5f;
    RS_LANE_LOOP(i, S::NV) { s.vsum[i] += B * s.v[i]; s.asum[i] += B * s.x[i]; }
    RS_SYNC();
    if (st < 3) {
        integrate_pos(c, s.v, h * A);
        RS_LANE_LOOP(i, S::NV) { s.v[i] = s.v0[i] + h * A * s.x[i]; }
        RS_SYNC();
    } else {
        integrate_pos(c, s.vsum, h);
        RS_LANE_LOOP(i, S::NV) { s.v[i] = s.v0[i] + h * s.asum[i]; }
        RS_SYNC();
    }
}
template <int LA, int LB>
RS_HD void substep_begin(Ctx<LA, LB>& c) {
    typedef Slab<LA, LB> S;
    S& s = *c.s;
    RS_LANE_LOOP(i, S::NQ) { s.q0[i] = s.q[i]; }
    RS_LANE_LOOP(i, S::NV) { s.v0[i] = s.v[i]; s.vsum[i] = 0.f; s.asum[i] = 0.f; }
    RS_SYNC();
}

#ifndef RS_TRIP_ANY
#define RS_TRIP_ANY(p) (p)      // device: __syncthreads_or -- the block-wide vote doubles as the barrier that re-aligns the warps
#define RS_TRIP_SYNC()
#endif
#ifndef RS_TRIP_MAX_LEGS
#define RS_TRIP_MAX_LEGS 12     // pairs with at most this many legs (Ant and Bug pairs, Ant-Spider) re-align per trip and run persistent blocks; the larger ones per evaluation (Bug-Bug 3.91 -> 3.65 ms, Spider-Spider equal)
#endif
#ifndef RS_TRIP_MACHINE
#define RS_TRIP_MACHINE 1      // 0: re-align per evaluation for every morphology (round-1 behaviour)
#endif
// The warps of a block must run the same code at the same time (the kernel is several times the SM's instruction cache), but the
// pairs need different numbers of Newton iterations per evaluation (1 for 70 %, 2 for 25 %, 3..6 for the rest).  Re-aligning once
// per EVALUATION made every block wait for its slowest pair twenty times per step (the 46 % barrier stall of round 1).  Here the
// block re-aligns once per TRIP = [start of an evaluation, for the warps that begin one] + [one Newton iteration, every warp]:
// a pair that needs an extra iteration simply starts its next evaluation one trip later, nobody waits for its whole tail, and a
// block is finished after max-over-pairs(total iterations of the step) trips instead of sum-over-evaluations(max-over-pairs).
// next_pair(finish): finish the pair in the slab (finish = true: write its outputs) and load the next one; false = none left.
// (Measured and dropped: changing pairs in the first half of the next trip, beside the evaluation starts of the other warps -- equal.)
template <int LA, int LB, typename NextPair>
RS_HD void simulate_trips(Ctx<LA, LB>& c, int nsub, NextPair next_pair) {
    int sub = 0, st = 0, it = 0;
    bool fresh = true, done = !next_pair(false);
    if (nsub <= 0) { while (!done) done = !next_pair(true); }
    if (!done) substep_begin(c);
    while (RS_TRIP_ANY(!done)) {
        if (!done && fresh) { eval_begin(c); it = 0; fresh = false; }
        RS_TRIP_SYNC();
        if (!done) {
            const bool conv = solve_iter(c);
            it++;
            if (conv || it >= c.max_newton) {
                solve_finish(c, it, conv);
                RS_CLOCK_MARK(3);
                RS_CLOCK_END();
                rk_after_eval(c, st);
                fresh = true;
                if (++st == 4) {
                    st = 0;
                    if (++sub == nsub) { sub = 0; done = !next_pair(true); }
                    if (!done) substep_begin(c);
                }
            }
        }
    }
}
template <int LA, int LB>
RS_HD void simulate(Ctx<LA, LB>& c, int nsub) {
    // measured at E = 4096 (tools/bench_morphologies.py): Ant 1.65 -> 1.61 ms per trip, Bug 5.01 -> 5.06 (equal), Spider 7.56 -> 8.11:
    // the larger bodies have 19 / 15 warps per block (bigger slabs) and less to gain from not waiting, so they keep the
    // per-evaluation re-alignment
    if constexpr (RS_TRIP_MACHINE && LA + LB <= RS_TRIP_MAX_LEGS) {
        bool first = true;
        simulate_trips(c, nsub, [&](bool) -> bool { const bool r = first; first = false; return r; });       // the one pair already in the slab
    } else {
    for (int sub = 0; sub < nsub; sub++) {
        RS_SUBSTEP_SYNC();
        substep_begin(c);
        for (int st = 0; st < 4; st++) {
            forward(c);
            rk_after_eval(c, st);
        }
    }
    }
}

}  // namespace rs
