/* ORACLE -- test infrastructure only (see oracle/README.md).  NOT product code.
 *
 * Double-precision CPU restatement of the physics the reference obtains from
 * MuJoCo 2.1.0 for one RoboSumo scene:
 *
 *   reference call sites: robosumo/robosumo/envs/mujoco_env.py:104-129 (reset/forward/
 *   set_state/do_simulation) -> mujoco-py/mujoco_py/mjsim.pyx:91-129 (mj_resetData,
 *   mj_forward, mj_step) -> libmujoco210 (third-party binary, NOT in /root/reference).
 *   Stage order restated: mujoco-py/mujoco_py/pxd/mujoco.pxd:208-327 (mj_fwdPosition,
 *   mj_fwdVelocity, mj_fwdActuation, mj_fwdAcceleration, mj_fwdConstraint, mj_RungeKutta),
 *   options tatami.xml:3 (RK4, dt=0.01), defaults tatami.xml:6.
 *
 * PARITY STATUS: UNPINNED.  MuJoCo is absent from the build container and the
 * reference holds no golden trajectory for this path; the algorithm below follows
 * MuJoCo's published computation (soft constraints with solref/solimp impedance,
 * pyramidal friction cones, primal Newton solver on the convex cost, RK4 with a full
 * forward evaluation per stage).  Narrow phase, shared with the CUDA path and documented in
 * capsule-vs-box = the two endpoint spheres plus the closest box edge against the
 * capsule interior (exact distance outside the box; MuJoCo's mjc_CapsuleBox emits its own
 * multi-contact set); the thin border cylinders are treated as capsules (MuJoCo uses its
 * convex-convex routine for those pairs; they differ at the cylinders' flat ends only).
 *
 * Formulation is deliberately different from the CUDA kernels (generic body tree,
 * explicit per-body Jacobians, dense matrices) so that agreement is evidence.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>

#define MAXCON 256
#define MAXEFC 1200
#define MINVAL 1e-15
enum { G_PLANE = 0, G_SPHERE = 2, G_CAPSULE = 3, G_CYLINDER = 5, G_BOX = 6 };
enum { J_FREE = 0, J_HINGE = 3 };

typedef struct {
    int nq, nv, nu, nbody, njnt, ngeom;
    int *body_parent, *body_jntadr, *body_jntnum, *body_dofadr, *body_dofnum, *body_weldid;
    int *jnt_type, *jnt_qposadr, *jnt_dofadr, *jnt_bodyid, *jnt_limited;
    int *dof_bodyid, *dof_jntid;
    int *geom_type, *geom_bodyid, *geom_contype, *geom_conaffinity, *geom_condim;
    int *act_jntid;
    double timestep, gravity[3];
    double *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia;
    double *jnt_pos, *jnt_axis, *jnt_range, *jnt_margin;
    double *dof_armature, *dof_damping;
    double *geom_pos, *geom_quat, *geom_size, *geom_margin, *geom_friction;
    double *act_gear, *act_ctrlrange, *qpos0;
    /* derived at qpos0 (mj_setConst [M]) */
    double *body_invweight0, *dof_invweight0;
    int *ibuf; double *dbuf;
} Model;

typedef struct {
    double dist, pos[3], frame[9], margin, mu;
    int g1, g2;
} Contact;

typedef struct {
    double *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis, *gpos, *gmat;
    double *M, *H, *L, *bias, *smooth, *qacc_smooth, *qacc, *jacp, *jacr, *tmp, *grad, *dir;
    double *w, *alpha, *ao;
    Contact con[MAXCON];
    int ncon, nefc, nlimit;
    double *J, *efc_pos, *efc_margin, *efc_diag, *efc_R, *efc_D, *efc_aref, *efc_jar, *efc_jd, *efc_force;
    int newton_iters;
} Data;

/* ---------- small math ---------- */
static double dot3(const double* a, const double* b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
static void cross3(double* r, const double* a, const double* b) {
    double x = a[1]*b[2] - a[2]*b[1], y = a[2]*b[0] - a[0]*b[2], z = a[0]*b[1] - a[1]*b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
static double norm3(const double* a) { return sqrt(dot3(a, a)); }
static double normalize3(double* a) {
    double n = norm3(a);
    if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return 0; }
    a[0] /= n; a[1] /= n; a[2] /= n; return n;
}
static void quat_mul(double* r, const double* a, const double* b) {
    double w = a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3];
    double x = a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2];
    double y = a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1];
    double z = a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0];
    r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat_normalize(double* q) {
    double n = sqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3]);
    if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void quat2mat(double* m, const double* q) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    m[0] = w*w + x*x - y*y - z*z; m[1] = 2*(x*y - w*z);         m[2] = 2*(x*z + w*y);
    m[3] = 2*(x*y + w*z);         m[4] = w*w - x*x + y*y - z*z; m[5] = 2*(y*z - w*x);
    m[6] = 2*(x*z - w*y);         m[7] = 2*(y*z + w*x);         m[8] = w*w - x*x - y*y + z*z;
}
static void mat_vec(double* r, const double* m, const double* v) {
    double x = m[0]*v[0] + m[1]*v[1] + m[2]*v[2], y = m[3]*v[0] + m[4]*v[1] + m[5]*v[2],
           z = m[6]*v[0] + m[7]*v[1] + m[8]*v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static void matT_vec(double* r, const double* m, const double* v) {
    double x = m[0]*v[0] + m[3]*v[1] + m[6]*v[2], y = m[1]*v[0] + m[4]*v[1] + m[7]*v[2],
           z = m[2]*v[0] + m[5]*v[1] + m[8]*v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static void axisangle2quat(double* q, const double* axis, double ang) {
    double s = sin(ang / 2);
    q[0] = cos(ang / 2); q[1] = axis[0]*s; q[2] = axis[1]*s; q[3] = axis[2]*s;
}

/* ---------- model / data allocation ---------- */
Model* orc_model_create(const int* ints, const double* dbls) {
    Model* m = (Model*)calloc(1, sizeof(Model));
    m->nq = ints[0]; m->nv = ints[1]; m->nu = ints[2]; m->nbody = ints[3]; m->njnt = ints[4]; m->ngeom = ints[5];
    int nb = m->nbody, nj = m->njnt, nv = m->nv, ng = m->ngeom, nu = m->nu;
    int ni = 6 + 6*nb + 5*nj + 2*nv + 5*ng + nu;
    m->ibuf = (int*)malloc(sizeof(int) * ni);
    memcpy(m->ibuf, ints, sizeof(int) * ni);
    int* p = m->ibuf + 6;
    m->body_parent = p; p += nb; m->body_jntadr = p; p += nb; m->body_jntnum = p; p += nb;
    m->body_dofadr = p; p += nb; m->body_dofnum = p; p += nb; m->body_weldid = p; p += nb;
    m->jnt_type = p; p += nj; m->jnt_qposadr = p; p += nj; m->jnt_dofadr = p; p += nj;
    m->jnt_bodyid = p; p += nj; m->jnt_limited = p; p += nj;
    m->dof_bodyid = p; p += nv; m->dof_jntid = p; p += nv;
    m->geom_type = p; p += ng; m->geom_bodyid = p; p += ng; m->geom_contype = p; p += ng;
    m->geom_conaffinity = p; p += ng; m->geom_condim = p; p += ng;
    m->act_jntid = p; p += nu;
    int nd = 4 + 18*nb + 9*nj + 2*nv + 14*ng + 3*nu + m->nq;
    m->dbuf = (double*)malloc(sizeof(double) * (nd + 2*nb + nv));
    memcpy(m->dbuf, dbls, sizeof(double) * nd);
    double* d = m->dbuf;
    m->timestep = d[0]; m->gravity[0] = d[1]; m->gravity[1] = d[2]; m->gravity[2] = d[3]; d += 4;
    m->body_pos = d; d += 3*nb; m->body_quat = d; d += 4*nb; m->body_ipos = d; d += 3*nb;
    m->body_iquat = d; d += 4*nb; m->body_mass = d; d += nb; m->body_inertia = d; d += 3*nb;
    m->jnt_pos = d; d += 3*nj; m->jnt_axis = d; d += 3*nj; m->jnt_range = d; d += 2*nj; m->jnt_margin = d; d += nj;
    m->dof_armature = d; d += nv; m->dof_damping = d; d += nv;
    m->geom_pos = d; d += 3*ng; m->geom_quat = d; d += 4*ng; m->geom_size = d; d += 3*ng;
    m->geom_margin = d; d += ng; m->geom_friction = d; d += 3*ng;
    m->act_gear = d; d += nu; m->act_ctrlrange = d; d += 2*nu; m->qpos0 = d; d += m->nq;
    m->body_invweight0 = d; d += 2*nb; m->dof_invweight0 = d; d += nv;
    for (int i = 0; i < nb; i++) if (m->body_jntnum[i] > 1) { fprintf(stderr, "oracle: >1 joint per body unsupported\n"); return NULL; }
    for (int i = 0; i < 3*nj; i++) if (m->jnt_pos[i] != 0.0) { fprintf(stderr, "oracle: jnt_pos != 0 unsupported\n"); return NULL; }
    return m;
}
void orc_model_free(Model* m) { if (m) { free(m->ibuf); free(m->dbuf); free(m); } }

static Data* data_create(const Model* m) {
    Data* d = (Data*)calloc(1, sizeof(Data));
    int nb = m->nbody, nj = m->njnt, nv = m->nv, ng = m->ngeom;
    d->xpos = calloc(3*nb, 8); d->xquat = calloc(4*nb, 8); d->xmat = calloc(9*nb, 8); d->xipos = calloc(3*nb, 8);
    d->ximat = calloc(9*nb, 8); d->xanchor = calloc(3*nj, 8); d->xaxis = calloc(3*nj, 8);
    d->gpos = calloc(3*ng, 8); d->gmat = calloc(9*ng, 8);
    d->M = calloc(nv*nv, 8); d->H = calloc(nv*nv, 8); d->L = calloc(nv*nv, 8);
    d->bias = calloc(nv, 8); d->smooth = calloc(nv, 8); d->qacc_smooth = calloc(nv, 8); d->qacc = calloc(nv, 8);
    d->jacp = calloc(3*nv, 8); d->jacr = calloc(3*nv, 8); d->tmp = calloc(nv, 8); d->grad = calloc(nv, 8); d->dir = calloc(nv, 8);
    d->w = calloc(3*nb, 8); d->alpha = calloc(3*nb, 8); d->ao = calloc(3*nb, 8);
    d->J = calloc((size_t)MAXEFC * nv, 8);
    d->efc_pos = calloc(MAXEFC, 8); d->efc_margin = calloc(MAXEFC, 8); d->efc_diag = calloc(MAXEFC, 8);
    d->efc_R = calloc(MAXEFC, 8); d->efc_D = calloc(MAXEFC, 8); d->efc_aref = calloc(MAXEFC, 8);
    d->efc_jar = calloc(MAXEFC, 8); d->efc_jd = calloc(MAXEFC, 8); d->efc_force = calloc(MAXEFC, 8);
    return d;
}
static void data_free(Data* d) {
    free(d->xpos); free(d->xquat); free(d->xmat); free(d->xipos); free(d->ximat); free(d->xanchor); free(d->xaxis);
    free(d->gpos); free(d->gmat); free(d->M); free(d->H); free(d->L); free(d->bias); free(d->smooth);
    free(d->qacc_smooth); free(d->qacc); free(d->jacp); free(d->jacr); free(d->tmp); free(d->grad); free(d->dir);
    free(d->w); free(d->alpha); free(d->ao); free(d->J); free(d->efc_pos); free(d->efc_margin); free(d->efc_diag);
    free(d->efc_R); free(d->efc_D); free(d->efc_aref); free(d->efc_jar); free(d->efc_jd); free(d->efc_force); free(d);
}

/* ---------- kinematics (mj_kinematics [M]); normalises free-joint quaternions in qpos in place ---------- */
static void kinematics(const Model* m, Data* d, double* qpos) {
    d->xpos[0] = d->xpos[1] = d->xpos[2] = 0;
    d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
    quat2mat(d->xmat, d->xquat);
    for (int i = 1; i < m->nbody; i++) {
        int par = m->body_parent[i];
        double* xp = d->xpos + 3*i; double* xq = d->xquat + 4*i;
        int jn = m->body_jntnum[i], ja = m->body_jntadr[i];
        if (jn == 1 && m->jnt_type[ja] == J_FREE) {
            int qa = m->jnt_qposadr[ja];
            quat_normalize(qpos + qa + 3);
            memcpy(xp, qpos + qa, 24); memcpy(xq, qpos + qa + 3, 32);
            memcpy(d->xanchor + 3*ja, xp, 24);
            d->xaxis[3*ja] = 0; d->xaxis[3*ja+1] = 0; d->xaxis[3*ja+2] = 1;
        } else {
            double v[3];
            mat_vec(v, d->xmat + 9*par, m->body_pos + 3*i);
            for (int k = 0; k < 3; k++) xp[k] = d->xpos[3*par + k] + v[k];
            quat_mul(xq, d->xquat + 4*par, m->body_quat + 4*i);
            if (jn == 1) {   /* hinge, jnt_pos == 0: anchor = body origin */
                double mat[9], q[4], r[4];
                quat2mat(mat, xq);
                memcpy(d->xanchor + 3*ja, xp, 24);
                mat_vec(d->xaxis + 3*ja, mat, m->jnt_axis + 3*ja);
                axisangle2quat(q, m->jnt_axis + 3*ja, qpos[m->jnt_qposadr[ja]] - m->qpos0[m->jnt_qposadr[ja]]);
                quat_mul(r, xq, q); memcpy(xq, r, 32);
            }
        }
        quat_normalize(xq);
        quat2mat(d->xmat + 9*i, xq);
        double v[3], q[4];
        mat_vec(v, d->xmat + 9*i, m->body_ipos + 3*i);
        for (int k = 0; k < 3; k++) d->xipos[3*i + k] = xp[k] + v[k];
        quat_mul(q, xq, m->body_iquat + 4*i);
        quat2mat(d->ximat + 9*i, q);
    }
    for (int g = 0; g < m->ngeom; g++) {
        int b = m->geom_bodyid[g];
        double v[3], q[4];
        mat_vec(v, d->xmat + 9*b, m->geom_pos + 3*g);
        for (int k = 0; k < 3; k++) d->gpos[3*g + k] = d->xpos[3*b + k] + v[k];
        quat_mul(q, d->xquat + 4*b, m->geom_quat + 4*g);
        quat2mat(d->gmat + 9*g, q);
    }
}

/* point Jacobian of body b at world point p: jacp, jacr are 3 x nv (mj_jac [M]) */
static void jac(const Model* m, const Data* d, int b, const double* p, double* jacp, double* jacr) {
    int nv = m->nv;
    memset(jacp, 0, sizeof(double) * 3 * nv); memset(jacr, 0, sizeof(double) * 3 * nv);
    for (int body = b; body != 0; body = m->body_parent[body]) {
        if (m->body_jntnum[body] == 0) continue;
        int j = m->body_jntadr[body], da = m->jnt_dofadr[j];
        if (m->jnt_type[j] == J_FREE) {
            for (int k = 0; k < 3; k++) jacp[k*nv + da + k] = 1.0;
            double r[3] = { p[0] - d->xpos[3*body], p[1] - d->xpos[3*body+1], p[2] - d->xpos[3*body+2] };
            for (int k = 0; k < 3; k++) {
                double ax[3] = { d->xmat[9*body + k], d->xmat[9*body + 3 + k], d->xmat[9*body + 6 + k] }, c[3];
                cross3(c, ax, r);
                for (int i = 0; i < 3; i++) { jacr[i*nv + da + 3 + k] = ax[i]; jacp[i*nv + da + 3 + k] = c[i]; }
            }
        } else {
            const double* ax = d->xaxis + 3*j;
            double r[3] = { p[0] - d->xanchor[3*j], p[1] - d->xanchor[3*j+1], p[2] - d->xanchor[3*j+2] }, c[3];
            cross3(c, ax, r);
            for (int i = 0; i < 3; i++) { jacr[i*nv + da] = ax[i]; jacp[i*nv + da] = c[i]; }
        }
    }
}

/* joint-space inertia (mj_crb result [M]) and bias forces (mj_rne, flg_acc=0 [M]) by explicit body sums */
static void inertia_and_bias(const Model* m, Data* d, const double* qvel, int want_bias) {
    int nv = m->nv, nb = m->nbody;
    memset(d->M, 0, sizeof(double) * nv * nv);
    memset(d->bias, 0, sizeof(double) * nv);
    /* velocity / bias-acceleration recursion (qacc = 0) */
    memset(d->w, 0, 24 * nb); memset(d->alpha, 0, 24 * nb); memset(d->ao, 0, 24 * nb);
    for (int i = 1; i < nb; i++) {
        int par = m->body_parent[i];
        double *w = d->w + 3*i, *al = d->alpha + 3*i, *ao = d->ao + 3*i;
        const double *wp = d->w + 3*par, *alp = d->alpha + 3*par, *aop = d->ao + 3*par;
        int jn = m->body_jntnum[i], ja = m->body_jntadr[i];
        if (jn == 1 && m->jnt_type[ja] == J_FREE) {
            mat_vec(w, d->xmat + 9*i, qvel + m->jnt_dofadr[ja] + 3);   /* body-frame angular velocity */
            /* alpha = 0, ao = 0 when qacc = 0 */
        } else {
            double r[3] = { d->xpos[3*i] - d->xpos[3*par], d->xpos[3*i+1] - d->xpos[3*par+1], d->xpos[3*i+2] - d->xpos[3*par+2] };
            double c1[3], c2[3];
            cross3(c1, alp, r); cross3(c2, wp, r); cross3(c2, wp, c2);
            for (int k = 0; k < 3; k++) { ao[k] = aop[k] + c1[k] + c2[k]; w[k] = wp[k]; al[k] = alp[k]; }
            if (jn == 1) {
                const double* ax = d->xaxis + 3*ja; double qd = qvel[m->jnt_dofadr[ja]], c[3];
                cross3(c, wp, ax);
                for (int k = 0; k < 3; k++) { w[k] += ax[k]*qd; al[k] += c[k]*qd; }
            }
        }
    }
    for (int b = 1; b < nb; b++) {
        double mass = m->body_mass[b];
        const double* R = d->ximat + 9*b; const double* I = m->body_inertia + 3*b;
        double Iw[9];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
            Iw[3*i + j] = R[3*i]*I[0]*R[3*j] + R[3*i+1]*I[1]*R[3*j+1] + R[3*i+2]*I[2]*R[3*j+2];
        jac(m, d, b, d->xipos + 3*b, d->jacp, d->jacr);
        for (int i = 0; i < nv; i++) {
            double jpi[3] = { d->jacp[i], d->jacp[nv+i], d->jacp[2*nv+i] };
            double jri[3] = { d->jacr[i], d->jacr[nv+i], d->jacr[2*nv+i] };
            if (jpi[0] == 0 && jpi[1] == 0 && jpi[2] == 0 && jri[0] == 0 && jri[1] == 0 && jri[2] == 0) continue;
            double Ijr[3]; matT_vec(Ijr, Iw, jri);   /* Iw symmetric */
            for (int j = 0; j < nv; j++) {
                double jpj[3] = { d->jacp[j], d->jacp[nv+j], d->jacp[2*nv+j] };
                double jrj[3] = { d->jacr[j], d->jacr[nv+j], d->jacr[2*nv+j] };
                d->M[i*nv + j] += mass * dot3(jpi, jpj) + dot3(Ijr, jrj);
            }
        }
        if (want_bias) {
            const double *w = d->w + 3*b, *al = d->alpha + 3*b, *ao = d->ao + 3*b;
            double r[3] = { d->xipos[3*b] - d->xpos[3*b], d->xipos[3*b+1] - d->xpos[3*b+1], d->xipos[3*b+2] - d->xpos[3*b+2] };
            double c1[3], c2[3], F[3], N[3], Iwv[3], Ial[3];
            cross3(c1, al, r); cross3(c2, w, r); cross3(c2, w, c2);
            for (int k = 0; k < 3; k++) F[k] = mass * (ao[k] + c1[k] + c2[k] - m->gravity[k]);
            mat_vec(Iwv, Iw, w); mat_vec(Ial, Iw, al); cross3(N, w, Iwv);
            for (int k = 0; k < 3; k++) N[k] += Ial[k];
            for (int i = 0; i < nv; i++)
                d->bias[i] += d->jacp[i]*F[0] + d->jacp[nv+i]*F[1] + d->jacp[2*nv+i]*F[2]
                            + d->jacr[i]*N[0] + d->jacr[nv+i]*N[1] + d->jacr[2*nv+i]*N[2];
        }
    }
    for (int i = 0; i < nv; i++) d->M[i*nv + i] += m->dof_armature[i];
}

/* dense Cholesky A = L L^T (lower), returns 0 on success */
static int cholesky(const double* A, double* L, int n) {
    memset(L, 0, sizeof(double) * n * n);
    for (int j = 0; j < n; j++) {
        double s = A[j*n + j];
        for (int k = 0; k < j; k++) s -= L[j*n + k] * L[j*n + k];
        if (s <= 0) return -1;
        L[j*n + j] = sqrt(s);
        for (int i = j + 1; i < n; i++) {
            double t = A[i*n + j];
            for (int k = 0; k < j; k++) t -= L[i*n + k] * L[j*n + k];
            L[i*n + j] = t / L[j*n + j];
        }
    }
    return 0;
}
static void chol_solve(const double* L, double* x, int n) {   /* in place */
    for (int i = 0; i < n; i++) { double s = x[i]; for (int k = 0; k < i; k++) s -= L[i*n + k]*x[k]; x[i] = s / L[i*n + i]; }
    for (int i = n - 1; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < n; k++) s -= L[k*n + i]*x[k]; x[i] = s / L[i*n + i]; }
}

/* ---------- narrow phase.  Normal points from geom1 to geom2 (types ordered) ---------- */
static void make_frame(double* f) {   /* mju_makeFrame [M] */
    normalize3(f);
    if (norm3(f + 3) < 0.5) {
        f[3] = f[4] = f[5] = 0;
        if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
    }
    double t = dot3(f, f + 3);
    for (int k = 0; k < 3; k++) f[3 + k] -= t * f[k];
    normalize3(f + 3);
    cross3(f + 6, f, f + 3);
}
static int add_con(Data* d, int g1, int g2, double dist, const double* pos, const double* n, const double* yhint, double margin) {
    if (dist >= margin || d->ncon >= MAXCON) return 0;
    Contact* c = &d->con[d->ncon++];
    c->g1 = g1; c->g2 = g2; c->dist = dist; c->margin = margin;
    memcpy(c->pos, pos, 24); memcpy(c->frame, n, 24);
    if (yhint) memcpy(c->frame + 3, yhint, 24); else c->frame[3] = c->frame[4] = c->frame[5] = 0;
    make_frame(c->frame);
    return 1;
}
static void sphere_sphere(Data* d, int g1, int g2, const double* c1, double r1, const double* c2, double r2, double margin) {
    double n[3] = { c2[0]-c1[0], c2[1]-c1[1], c2[2]-c1[2] }, pos[3];
    double len = normalize3(n), dist = len - r1 - r2;
    for (int k = 0; k < 3; k++) pos[k] = c1[k] + n[k] * (r1 + 0.5*dist);
    add_con(d, g1, g2, dist, pos, n, NULL, margin);
}
static void plane_sphere(Data* d, int gp, int gs, const double* c, double r, double margin, const double* yhint) {
    const double* pp = d->gpos + 3*gp; const double* pm = d->gmat + 9*gp;
    double n[3] = { pm[2], pm[5], pm[8] }, diff[3] = { c[0]-pp[0], c[1]-pp[1], c[2]-pp[2] }, pos[3];
    double dist = dot3(diff, n) - r;
    for (int k = 0; k < 3; k++) pos[k] = c[k] - n[k] * (r + 0.5*dist);
    add_con(d, gp, gs, dist, pos, n, yhint, margin);
}
/* nearest point of a capsule/cylinder axis segment (geom g) to point c */
static void seg_nearest(const Data* d, const Model* m, int g, const double* c, double* out) {
    const double* p = d->gpos + 3*g; const double* R = d->gmat + 9*g;
    double ax[3] = { R[2], R[5], R[8] }, diff[3] = { c[0]-p[0], c[1]-p[1], c[2]-p[2] };
    double x = dot3(diff, ax), h = m->geom_size[3*g + 1];
    if (x > h) x = h; if (x < -h) x = -h;
    for (int k = 0; k < 3; k++) out[k] = p[k] + ax[k]*x;
}
static void capsule_capsule(Data* d, const Model* m, int g1, int g2, double margin) {
    const double *p1 = d->gpos + 3*g1, *p2 = d->gpos + 3*g2, *R1 = d->gmat + 9*g1, *R2 = d->gmat + 9*g2;
    double a1[3] = { R1[2], R1[5], R1[8] }, a2[3] = { R2[2], R2[5], R2[8] };
    double l1 = m->geom_size[3*g1 + 1], l2 = m->geom_size[3*g2 + 1], r1 = m->geom_size[3*g1], r2 = m->geom_size[3*g2];
    double dif[3] = { p1[0]-p2[0], p1[1]-p2[1], p1[2]-p2[2] };
    double ma = 1.0, mb = -dot3(a1, a2), mc = 1.0, u = -dot3(a1, dif), v = dot3(a2, dif);
    double det = ma*mc - mb*mb;
    if (fabs(det) >= 1e-12) {
        double x1 = (mc*u - mb*v) / det, x2 = (ma*v - mb*u) / det;
        if (x1 > l1) { x1 = l1; x2 = (v - mb*l1) / mc; } else if (x1 < -l1) { x1 = -l1; x2 = (v + mb*l1) / mc; }
        if (x2 > l2) { x2 = l2; x1 = (u - mb*l2) / ma; } else if (x2 < -l2) { x2 = -l2; x1 = (u + mb*l2) / ma; }
        if (x1 > l1) x1 = l1; else if (x1 < -l1) x1 = -l1;
        double v1[3], v2[3];
        for (int k = 0; k < 3; k++) { v1[k] = p1[k] + a1[k]*x1; v2[k] = p2[k] + a2[k]*x2; }
        sphere_sphere(d, g1, g2, v1, r1, v2, r2, margin);
    } else {   /* parallel axes: endpoints of capsule 1 against segment 2 */
        for (int s = -1; s <= 1; s += 2) {
            double e[3], q[3];
            for (int k = 0; k < 3; k++) e[k] = p1[k] + a1[k]*l1*s;
            seg_nearest(d, m, g2, e, q);
            sphere_sphere(d, g1, g2, e, r1, q, r2, margin);
        }
    }
}
static void sphere_box(Data* d, const Model* m, int gs, int gb, const double* c, double r, double margin) {
    const double* bp = d->gpos + 3*gb; const double* R = d->gmat + 9*gb; const double* h = m->geom_size + 3*gb;
    double diff[3] = { c[0]-bp[0], c[1]-bp[1], c[2]-bp[2] }, loc[3], cl[3];
    matT_vec(loc, R, diff);
    int outside = 0;
    for (int k = 0; k < 3; k++) { cl[k] = loc[k]; if (cl[k] > h[k]) { cl[k] = h[k]; outside = 1; } else if (cl[k] < -h[k]) { cl[k] = -h[k]; outside = 1; } }
    double nl[3], dist, posl[3];
    if (outside) {
        for (int k = 0; k < 3; k++) nl[k] = loc[k] - cl[k];
        double len = normalize3(nl);
        dist = len - r;
        for (int k = 0; k < 3; k++) posl[k] = cl[k] + nl[k]*0.5*dist;
    } else {   /* centre inside the box: push out through the nearest face */
        int kb = 0; double best = 1e300;
        for (int k = 0; k < 3; k++) { double pen = h[k] - fabs(loc[k]); if (pen < best) { best = pen; kb = k; } }
        nl[0] = nl[1] = nl[2] = 0; nl[kb] = loc[kb] >= 0 ? 1 : -1;
        dist = -best - r;
        for (int k = 0; k < 3; k++) posl[k] = loc[k] + nl[k]*(-r - 0.5*dist);
    }
    double n[3], pos[3];
    mat_vec(n, R, nl); mat_vec(pos, R, posl);
    for (int k = 0; k < 3; k++) { pos[k] += bp[k]; n[k] = -n[k]; }   /* from sphere (geom1) to box (geom2) */
    add_con(d, gs, gb, dist, pos, n, NULL, margin);
}
/* Capsule INTERIOR against the box edges.  The distance from a segment to a convex box is attained at a segment endpoint (the two
 * endpoint-sphere tests) or, with an interior point of the segment, on a box EDGE; so those two plus this test give the exact
 * distance whenever the axis is outside the box (a leg straddling the tatami edge, tatami.xml:21, utils.py:64-68).  At most one
 * contact: the closest edge.  [M] MuJoCo's own mjc_CapsuleBox is a different multi-contact routine; see DESIGN.md section 4. */
static void capsule_box_edges(Data* d, const Model* m, int gc, int gb, double margin) {
    const double* bp = d->gpos + 3*gb; const double* R = d->gmat + 9*gb; const double* h = m->geom_size + 3*gb;
    const double* cp = d->gpos + 3*gc; const double* Rc = d->gmat + 9*gc;
    double r = m->geom_size[3*gc], l1 = m->geom_size[3*gc + 1];
    double diff[3] = { cp[0]-bp[0], cp[1]-bp[1], cp[2]-bp[2] }, p1[3], a1w[3] = { Rc[2], Rc[5], Rc[8] }, a1[3];
    matT_vec(p1, R, diff); matT_vec(a1, R, a1w);                 /* capsule centre and axis in the box frame */
    double best = 1e300, bpg[3] = {0, 0, 0}, bpe[3] = {0, 0, 0};
    for (int ax = 0; ax < 3; ax++) for (int s1 = -1; s1 <= 1; s1 += 2) for (int s2 = -1; s2 <= 1; s2 += 2) {
        int u = (ax + 1) % 3, w = (ax + 2) % 3;
        double p2[3] = {0, 0, 0}, a2[3] = {0, 0, 0}, l2 = h[ax];
        p2[u] = s1 * h[u]; p2[w] = s2 * h[w]; a2[ax] = 1.0;
        double dif[3] = { p1[0]-p2[0], p1[1]-p2[1], p1[2]-p2[2] };
        double mb = -dot3(a1, a2), uu = -dot3(a1, dif), vv = dot3(a2, dif), det = 1.0 - mb*mb, x1, x2;
        if (fabs(det) < 1e-12) continue;                        /* parallel: the endpoint tests cover it */
        x1 = (uu - mb*vv) / det; x2 = (vv - mb*uu) / det;
        if (x1 > l1) { x1 = l1; x2 = vv - mb*l1; } else if (x1 < -l1) { x1 = -l1; x2 = vv + mb*l1; }
        if (x2 > l2) { x2 = l2; x1 = uu - mb*l2; } else if (x2 < -l2) { x2 = -l2; x1 = uu + mb*l2; }
        if (x1 > l1) x1 = l1; else if (x1 < -l1) x1 = -l1;
        if (fabs(x1) >= l1 * (1.0 - 1e-6)) continue;            /* an endpoint is closest: already tested as a sphere */
        double pg[3], pe[3];
        for (int k = 0; k < 3; k++) { pg[k] = p1[k] + a1[k]*x1; pe[k] = p2[k] + a2[k]*x2; }
        if (fabs(pg[0]) < h[0] && fabs(pg[1]) < h[1] && fabs(pg[2]) < h[2]) continue;   /* axis point inside the box: face push-out case */
        double dd[3] = { pg[0]-pe[0], pg[1]-pe[1], pg[2]-pe[2] }, len = sqrt(dot3(dd, dd));
        if (len - r < best) { best = len - r; memcpy(bpg, pg, 24); memcpy(bpe, pe, 24); }
    }
    if (best >= margin) return;
    double nl[3] = { bpe[0]-bpg[0], bpe[1]-bpg[1], bpe[2]-bpg[2] }, posl[3], n[3], pos[3];   /* from capsule (geom1) to box (geom2) */
    if (normalize3(nl) < 1e-12) return;
    for (int k = 0; k < 3; k++) posl[k] = bpe[k] - nl[k]*0.5*best;
    mat_vec(n, R, nl); mat_vec(pos, R, posl);
    for (int k = 0; k < 3; k++) pos[k] += bp[k];
    add_con(d, gc, gb, best, pos, n, NULL, margin);
}
static void cap_ends(const Data* d, const Model* m, int g, double* e0, double* e1) {
    const double* p = d->gpos + 3*g; const double* R = d->gmat + 9*g; double h = m->geom_size[3*g + 1];
    for (int k = 0; k < 3; k++) { e0[k] = p[k] + R[3*k + 2]*h; e1[k] = p[k] - R[3*k + 2]*h; }
}

static void collide_pair(const Model* m, Data* d, int ga, int gb) {
    int g1 = ga, g2 = gb;
    if (m->geom_type[g1] > m->geom_type[g2]) { g1 = gb; g2 = ga; }
    int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]);
    double r1 = m->geom_size[3*g1], r2 = m->geom_size[3*g2];
    double e0[3], e1[3], q[3];
    if (t1 == G_PLANE && t2 == G_SPHERE) plane_sphere(d, g1, g2, d->gpos + 3*g2, r2, margin, NULL);
    else if (t1 == G_PLANE && t2 == G_CAPSULE) {
        const double* R = d->gmat + 9*g2; double ax[3] = { R[2], R[5], R[8] };
        cap_ends(d, m, g2, e0, e1);
        plane_sphere(d, g1, g2, e0, r2, margin, ax); plane_sphere(d, g1, g2, e1, r2, margin, ax);
    }
    else if (t1 == G_SPHERE && t2 == G_SPHERE) sphere_sphere(d, g1, g2, d->gpos + 3*g1, r1, d->gpos + 3*g2, r2, margin);
    else if (t1 == G_SPHERE && (t2 == G_CAPSULE || t2 == G_CYLINDER)) {
        seg_nearest(d, m, g2, d->gpos + 3*g1, q);
        sphere_sphere(d, g1, g2, d->gpos + 3*g1, r1, q, r2, margin);
    }
    else if (t1 == G_CAPSULE && (t2 == G_CAPSULE || t2 == G_CYLINDER)) capsule_capsule(d, m, g1, g2, margin);
    else if (t1 == G_SPHERE && t2 == G_BOX) sphere_box(d, m, g1, g2, d->gpos + 3*g1, r1, margin);
    else if (t1 == G_CAPSULE && t2 == G_BOX) {
        cap_ends(d, m, g1, e0, e1);
        sphere_box(d, m, g1, g2, e0, r1, margin); sphere_box(d, m, g1, g2, e1, r1, margin);
        capsule_box_edges(d, m, g1, g2, margin);
    }
}

static void collision(const Model* m, Data* d) {   /* mj_collision filters [M] */
    d->ncon = 0;
    for (int g1 = 0; g1 < m->ngeom; g1++) for (int g2 = g1 + 1; g2 < m->ngeom; g2++) {
        int b1 = m->geom_bodyid[g1], b2 = m->geom_bodyid[g2];
        int w1 = m->body_weldid[b1], w2 = m->body_weldid[b2];
        if (w1 == w2) continue;
        int pw1 = m->body_weldid[m->body_parent[w1]], pw2 = m->body_weldid[m->body_parent[w2]];
        if (w1 != 0 && w2 != 0 && (w1 == pw2 || w2 == pw1)) continue;
        if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]))) continue;
        collide_pair(m, d, g1, g2);
    }
}

/* ---------- constraints (mj_makeConstraint / mj_makeImpedance [M]) ---------- */
static const double SOLREF[2] = { 0.02, 1.0 };
static const double SOLIMP[5] = { 0.9, 0.95, 0.001, 0.5, 2.0 };

static void impedance_row(const Model* m, Data* d, int i, const double* qvel_J) {
    /* i: row; uses efc_pos, efc_margin, efc_diag; qvel_J = (J qvel)[i] */
    double timeconst = fmax(SOLREF[0], 2.0 * m->timestep), dampratio = SOLREF[1];   /* refsafe */
    double dmin = SOLIMP[0], dmax = SOLIMP[1], width = SOLIMP[2], mid = SOLIMP[3], power = SOLIMP[4];
    double x = fabs(d->efc_pos[i] - d->efc_margin[i]) / width, y;
    if (x >= 1) y = 1; else if (x <= 0) y = 0;
    else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
    else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
    double imp = dmin + y * (dmax - dmin);
    d->efc_R[i] = fmax(MINVAL, (1 - imp) * d->efc_diag[i] / imp);
    double K = 1.0 / fmax(MINVAL, dmax*dmax * timeconst*timeconst * dampratio*dampratio);
    double B = 2.0 / fmax(MINVAL, dmax * timeconst);
    d->efc_aref[i] = -B * qvel_J[i] - K * imp * (d->efc_pos[i] - d->efc_margin[i]);
}

static void make_constraint(const Model* m, Data* d, const double* qpos, const double* qvel) {
    int nv = m->nv, n = 0;
    /* joint limits */
    for (int j = 0; j < m->njnt; j++) {
        if (!m->jnt_limited[j] || m->jnt_type[j] != J_HINGE) continue;
        double q = qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
        for (int side = 0; side < 2; side++) {
            double dist = side == 0 ? q - m->jnt_range[2*j] : m->jnt_range[2*j + 1] - q;
            if (dist < margin && n < MAXEFC) {
                memset(d->J + (size_t)n*nv, 0, sizeof(double) * nv);
                d->J[(size_t)n*nv + m->jnt_dofadr[j]] = side == 0 ? 1.0 : -1.0;
                d->efc_pos[n] = dist; d->efc_margin[n] = margin;
                d->efc_diag[n] = m->dof_invweight0[m->jnt_dofadr[j]];
                n++;
            }
        }
    }
    d->nlimit = n;
    /* contacts: pyramidal cone, condim 3 -> 4 rows */
    double* jp2 = (double*)malloc(sizeof(double) * 6 * nv); double* jr = jp2 + 3*nv;
    for (int c = 0; c < d->ncon; c++) {
        Contact* con = &d->con[c];
        int b1 = m->geom_bodyid[con->g1], b2 = m->geom_bodyid[con->g2];
        con->mu = fmax(m->geom_friction[3*con->g1], m->geom_friction[3*con->g2]);
        if (n + 4 > MAXEFC) break;
        jac(m, d, b2, con->pos, jp2, jr);
        jac(m, d, b1, con->pos, d->jacp, d->jacr);
        for (int i = 0; i < 3*nv; i++) jp2[i] -= d->jacp[i];
        double tran = m->body_invweight0[2*b1] + m->body_invweight0[2*b2];
        for (int r = 0; r < 4; r++) {
            const double* t = con->frame + 3 * (1 + r / 2);
            double sgn = (r % 2 == 0) ? 1.0 : -1.0, dir[3];
            for (int k = 0; k < 3; k++) dir[k] = con->frame[k] + sgn * con->mu * t[k];
            double* row = d->J + (size_t)(n + r) * nv;
            for (int i = 0; i < nv; i++) row[i] = dir[0]*jp2[i] + dir[1]*jp2[nv + i] + dir[2]*jp2[2*nv + i];
            d->efc_pos[n + r] = con->dist; d->efc_margin[n + r] = con->margin;
            d->efc_diag[n + r] = tran + con->mu*con->mu*tran;
        }
        n += 4;
    }
    free(jp2);
    d->nefc = n;
    /* impedance, R, aref */
    for (int i = 0; i < n; i++) {
        double s = 0; const double* row = d->J + (size_t)i*nv;
        for (int k = 0; k < nv; k++) s += row[k]*qvel[k];
        d->efc_jd[i] = s;
    }
    for (int i = 0; i < n; i++) impedance_row(m, d, i, d->efc_jd);
    for (int c = 0, i = d->nlimit; i + 3 < n; c++, i += 4) {   /* pyramid regularisation: Rpy = 2 mu^2 R(first row) */
        double mu = d->con[c].mu, Rpy = 2 * mu*mu * d->efc_R[i];
        for (int r = 0; r < 4; r++) d->efc_R[i + r] = Rpy;
    }
    for (int i = 0; i < n; i++) d->efc_D[i] = 1.0 / d->efc_R[i];
}

/* ---------- primal Newton on the convex cost (mj_solNewton semantics [M], pyramidal rows = one-sided quadratics) ---------- */
static double cost_and_state(const Model* m, Data* d, const double* x, double* grad) {
    int nv = m->nv, n = d->nefc;
    double cost = 0;
    for (int i = 0; i < nv; i++) {
        double s = 0;
        for (int j = 0; j < nv; j++) s += d->M[i*nv + j] * (x[j] - d->qacc_smooth[j]);
        d->tmp[i] = s;
    }
    for (int i = 0; i < nv; i++) cost += 0.5 * d->tmp[i] * (x[i] - d->qacc_smooth[i]);
    for (int i = 0; i < n; i++) {
        const double* row = d->J + (size_t)i*nv; double s = 0;
        for (int k = 0; k < nv; k++) s += row[k]*x[k];
        d->efc_jar[i] = s - d->efc_aref[i];
        if (d->efc_jar[i] < 0) { d->efc_force[i] = -d->efc_D[i]*d->efc_jar[i]; cost += 0.5*d->efc_D[i]*d->efc_jar[i]*d->efc_jar[i]; }
        else d->efc_force[i] = 0;
    }
    if (grad) {
        for (int k = 0; k < nv; k++) grad[k] = d->tmp[k];
        for (int i = 0; i < n; i++) if (d->efc_force[i] != 0) {
            const double* row = d->J + (size_t)i*nv;
            for (int k = 0; k < nv; k++) grad[k] -= row[k]*d->efc_force[i];
        }
    }
    return cost;
}

static void solve_constraints(const Model* m, Data* d, const double* warm) {
    int nv = m->nv, n = d->nefc;
    d->newton_iters = 0;
    if (n == 0) { memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv); return; }
    double* x = d->qacc;
    /* warmstart: the better of qacc_warmstart and qacc_smooth */
    memcpy(x, d->qacc_smooth, sizeof(double) * nv);
    if (warm) {
        double c0 = cost_and_state(m, d, d->qacc_smooth, NULL), c1 = cost_and_state(m, d, warm, NULL);
        if (c1 < c0) memcpy(x, warm, sizeof(double) * nv);
    }
    for (int it = 0; it < 100; it++) {
        d->newton_iters = it + 1;
        cost_and_state(m, d, x, d->grad);
        double g2 = 0; for (int k = 0; k < nv; k++) g2 += d->grad[k]*d->grad[k];
        if (sqrt(g2) < 1e-11) break;
        memcpy(d->H, d->M, sizeof(double) * nv * nv);
        for (int i = 0; i < n; i++) if (d->efc_jar[i] < 0) {
            const double* row = d->J + (size_t)i*nv; double D = d->efc_D[i];
            for (int a = 0; a < nv; a++) if (row[a] != 0) for (int b = 0; b < nv; b++) d->H[a*nv + b] += D*row[a]*row[b];
        }
        if (cholesky(d->H, d->L, nv)) break;
        for (int k = 0; k < nv; k++) d->dir[k] = -d->grad[k];
        chol_solve(d->L, d->dir, nv);
        /* exact line search on the piecewise-quadratic: phi'(a) = p0 + a*p1 + sum_{active(a)} D (jar+a jd) jd */
        double p0 = 0, p1 = 0;
        for (int i = 0; i < nv; i++) {
            double s = 0; for (int j = 0; j < nv; j++) s += d->M[i*nv + j]*d->dir[j];
            p1 += s*d->dir[i]; p0 += d->tmp[i]*d->dir[i];
        }
        for (int i = 0; i < n; i++) {
            const double* row = d->J + (size_t)i*nv; double s = 0;
            for (int k = 0; k < nv; k++) s += row[k]*d->dir[k];
            d->efc_jd[i] = s;
        }
        double lo = 0, hi = 1;
        #define DPHI(a, out) { double _s = p0 + (a)*p1; for (int i = 0; i < n; i++) { double j_ = d->efc_jar[i] + (a)*d->efc_jd[i]; if (j_ < 0) _s += d->efc_D[i]*j_*d->efc_jd[i]; } out = _s; }
        double dlo, dhi; DPHI(lo, dlo); DPHI(hi, dhi);
        int guard = 0;
        while (dhi < 0 && guard++ < 60) { lo = hi; dlo = dhi; hi *= 2; DPHI(hi, dhi); }
        double a = hi;
        if (dlo < 0 && dhi >= 0) {
            for (int k = 0; k < 200; k++) {
                /* safeguarded: exact root of the local linear piece, else bisection */
                double h2 = p1, da; DPHI(a, da);
                for (int i = 0; i < n; i++) { double j_ = d->efc_jar[i] + a*d->efc_jd[i]; if (j_ < 0) h2 += d->efc_D[i]*d->efc_jd[i]*d->efc_jd[i]; }
                if (da < 0) { lo = a; } else { hi = a; }
                if (fabs(da) < 1e-14 * (fabs(p0) + 1e-300) || hi - lo < 1e-16) break;
                double an = a - da / h2;
                if (!(an > lo && an < hi)) an = 0.5 * (lo + hi);
                a = an;
            }
        } else if (dlo >= 0) a = 0;
        #undef DPHI
        if (a == 0) break;
        for (int k = 0; k < nv; k++) x[k] += a * d->dir[k];
    }
    cost_and_state(m, d, x, NULL);
}

/* ---------- forward dynamics: qacc(qpos, qvel, ctrl) ---------- */
static void forward(const Model* m, Data* d, double* qpos, const double* qvel, const double* ctrl, const double* warm) {
    int nv = m->nv;
    kinematics(m, d, qpos);
    inertia_and_bias(m, d, qvel, 1);
    collision(m, d);
    make_constraint(m, d, qpos, qvel);
    for (int i = 0; i < nv; i++) d->smooth[i] = -m->dof_damping[i]*qvel[i] - d->bias[i];   /* passive - bias */
    for (int a = 0; a < m->nu; a++) {
        double c = ctrl[a], lo = m->act_ctrlrange[2*a], hi = m->act_ctrlrange[2*a + 1];
        if (c < lo) c = lo; if (c > hi) c = hi;
        d->smooth[m->jnt_dofadr[m->act_jntid[a]]] += m->act_gear[a] * c;
    }
    cholesky(d->M, d->L, nv);
    memcpy(d->qacc_smooth, d->smooth, sizeof(double) * nv);
    chol_solve(d->L, d->qacc_smooth, nv);
    solve_constraints(m, d, warm);
}

static void integrate_pos(const Model* m, double* qpos, const double* qvel, double h) {   /* mj_integratePos [M] */
    for (int j = 0; j < m->njnt; j++) {
        int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
        if (m->jnt_type[j] == J_FREE) {
            for (int k = 0; k < 3; k++) qpos[qa + k] += h * qvel[da + k];
            double w[3] = { qvel[da+3], qvel[da+4], qvel[da+5] };
            double ang = h * normalize3(w), q[4], r[4];
            axisangle2quat(q, w, ang);
            quat_mul(r, qpos + qa + 3, q); quat_normalize(r);
            memcpy(qpos + qa + 3, r, 32);
        } else qpos[qa] += h * qvel[da];
    }
}

/* one mj_step with RK4 (mj_RungeKutta, N=4 [M]); warm = qacc_warmstart (in/out) */
static void step_rk4(const Model* m, Data* d, double* qpos, double* qvel, const double* ctrl, double* warm) {
    int nq = m->nq, nv = m->nv; double h = m->timestep;
    static const double A[3] = { 0.5, 0.5, 1.0 }, B[4] = { 1.0/6, 1.0/3, 1.0/3, 1.0/6 };
    double *q0 = malloc(8*nq), *v0 = malloc(8*nv), *X = malloc(8*4*nv), *F = malloc(8*4*nv), *q = malloc(8*nq), *v = malloc(8*nv);
    forward(m, d, qpos, qvel, ctrl, warm);              /* mj_forward at the start state */
    memcpy(warm, d->qacc, 8*nv);                        /* qacc_warmstart saved after the main forward */
    memcpy(q0, qpos, 8*nq); memcpy(v0, qvel, 8*nv);
    memcpy(X, qvel, 8*nv); memcpy(F, d->qacc, 8*nv);
    for (int i = 1; i < 4; i++) {
        memcpy(q, q0, 8*nq);
        integrate_pos(m, q, X + (i-1)*nv, h * A[i-1]);
        for (int k = 0; k < nv; k++) v[k] = v0[k] + h * A[i-1] * F[(i-1)*nv + k];
        forward(m, d, q, v, ctrl, warm);
        memcpy(X + i*nv, v, 8*nv); memcpy(F + i*nv, d->qacc, 8*nv);
    }
    for (int k = 0; k < nv; k++) {
        double dx = 0, df = 0;
        for (int i = 0; i < 4; i++) { dx += B[i]*X[i*nv + k]; df += B[i]*F[i*nv + k]; }
        v[k] = dx; qvel[k] = v0[k] + h*df;
    }
    memcpy(qpos, q0, 8*nq);
    integrate_pos(m, qpos, v, h);
    free(q0); free(v0); free(X); free(F); free(q); free(v);
}

/* ---------- exported API ---------- */
/* mj_setConst [M]: body_invweight0 / dof_invweight0 at qpos0 */
void orc_set_const(Model* m) {
    Data* d = data_create(m);
    int nv = m->nv;
    double* qpos = malloc(8 * m->nq), *qvel = calloc(nv, 8), *A = malloc(8*6*nv), *col = malloc(8*nv);
    memcpy(qpos, m->qpos0, 8 * m->nq);
    kinematics(m, d, qpos);
    inertia_and_bias(m, d, qvel, 0);
    cholesky(d->M, d->L, nv);
    for (int b = 0; b < m->nbody; b++) {
        m->body_invweight0[2*b] = m->body_invweight0[2*b + 1] = 0;
        if (m->body_weldid[b] == 0) continue;
        jac(m, d, b, d->xipos + 3*b, d->jacp, d->jacr);
        double tr = 0, ro = 0;
        for (int r = 0; r < 6; r++) {
            const double* row = r < 3 ? d->jacp + r*nv : d->jacr + (r-3)*nv;
            memcpy(col, row, 8*nv); chol_solve(d->L, col, nv);
            double s = 0; for (int k = 0; k < nv; k++) s += row[k]*col[k];
            if (r < 3) tr += s; else ro += s;
        }
        m->body_invweight0[2*b] = tr / 3; m->body_invweight0[2*b + 1] = ro / 3;
    }
    for (int i = 0; i < nv; i++) {
        memset(col, 0, 8*nv); col[i] = 1; chol_solve(d->L, col, nv);
        m->dof_invweight0[i] = col[i];
    }
    for (int j = 0; j < m->njnt; j++) if (m->jnt_type[j] == J_FREE) {
        int da = m->jnt_dofadr[j];
        double a = (m->dof_invweight0[da] + m->dof_invweight0[da+1] + m->dof_invweight0[da+2]) / 3;
        double b = (m->dof_invweight0[da+3] + m->dof_invweight0[da+4] + m->dof_invweight0[da+5]) / 3;
        for (int k = 0; k < 3; k++) { m->dof_invweight0[da + k] = a; m->dof_invweight0[da + 3 + k] = b; }
    }
    free(qpos); free(qvel); free(A); free(col); data_free(d);
}
void orc_get_invweight(const Model* m, double* body_iw, double* dof_iw) {
    memcpy(body_iw, m->body_invweight0, 8 * 2 * m->nbody); memcpy(dof_iw, m->dof_invweight0, 8 * m->nv);
}

/* mj_forward: fills qacc and optional diagnostics. qpos is normalised in place. */
int orc_forward(const Model* m, double* qpos, const double* qvel, const double* ctrl,
                double* qacc, double* Mout, double* bias, double* qacc_smooth,
                double* con_out /* MAXCON x 8: dist,pos3,normal3,geompair */, int* nefc, int* iters,
                double* geom_xpos, double* body_xpos) {
    Data* d = data_create(m);
    forward(m, d, qpos, qvel, ctrl, NULL);
    int nv = m->nv, ncon = d->ncon;
    if (qacc) memcpy(qacc, d->qacc, 8*nv);
    if (Mout) memcpy(Mout, d->M, 8*nv*nv);
    if (bias) memcpy(bias, d->bias, 8*nv);
    if (qacc_smooth) memcpy(qacc_smooth, d->qacc_smooth, 8*nv);
    if (con_out) for (int c = 0; c < ncon; c++) {
        double* o = con_out + 8*c; o[0] = d->con[c].dist; memcpy(o + 1, d->con[c].pos, 24); memcpy(o + 4, d->con[c].frame, 24);
        o[7] = d->con[c].g1 * 1000 + d->con[c].g2;
    }
    if (nefc) *nefc = d->nefc;
    if (iters) *iters = d->newton_iters;
    if (geom_xpos) memcpy(geom_xpos, d->gpos, 8*3*m->ngeom);
    if (body_xpos) memcpy(body_xpos, d->xpos, 8*3*m->nbody);
    data_free(d);
    return ncon;
}

/* do_simulation: nsub x mj_step (mujoco_env.py:125-129).  Returns max ncon seen. warm: nv doubles in/out (may be NULL). */
int orc_step(const Model* m, double* qpos, double* qvel, const double* ctrl, int nsub, double* warm) {
    Data* d = data_create(m);
    double* w = calloc(m->nv, 8);
    if (warm) memcpy(w, warm, 8 * m->nv);
    int maxcon = 0;
    for (int s = 0; s < nsub; s++) { step_rk4(m, d, qpos, qvel, ctrl, w); if (d->ncon > maxcon) maxcon = d->ncon; }
    if (warm) memcpy(warm, w, 8 * m->nv);
    free(w); data_free(d);
    return maxcon;
}

/* batch of independent envs, pthreads over envs (cpu baseline) */
typedef struct { const Model* m; int E, nsub, tid, nthreads; double *qpos, *qvel, *warm; const double* ctrl; } BatchArg;
static void* batch_worker(void* p) {
    BatchArg* a = (BatchArg*)p; const Model* m = a->m;
    for (int e = a->tid; e < a->E; e += a->nthreads)
        orc_step(m, a->qpos + (size_t)e * m->nq, a->qvel + (size_t)e * m->nv, a->ctrl + (size_t)e * m->nu, a->nsub,
                 a->warm ? a->warm + (size_t)e * m->nv : NULL);
    return NULL;
}
void orc_step_batch(const Model* m, int E, double* qpos, double* qvel, const double* ctrl, int nsub, double* warm, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256]; BatchArg args[256];
    for (int t = 0; t < nthreads; t++) {
        args[t] = (BatchArg){ m, E, nsub, t, nthreads, qpos, qvel, warm, ctrl };
        pthread_create(&th[t], NULL, batch_worker, &args[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* set_state + mj_forward as far as the observation needs it: normalise quaternions (mujoco_env.py:110-119) */
void orc_normalize_qpos(const Model* m, double* qpos) {
    for (int j = 0; j < m->njnt; j++) if (m->jnt_type[j] == J_FREE) quat_normalize(qpos + m->jnt_qposadr[j] + 3);
}
