/*
 * bio_oracle.c -- CPU restatement (plain C, fp64) of the env-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import,
 * link or execute this file; it is used by tests/, by
 * __graft_entry__.smoke() as the checker and by bench.py's cpu_baseline /
 * --impl reference legs.
 *
 * PARITY UNPINNED for the dynamics as a whole: the arithmetic of this path
 * lives in opensim==4.1 (reference setup.py:7;
 * mitkof6/opensim-core@bindings_timestepper, reference
 * scripts/build_opensim-core:32) and its bundled Simbody, neither of which is
 * under /root/reference nor installable offline, and the reference ships no
 * golden vectors.  The physics below restates the published models (Millard
 * 2013 muscle, Simbody Hunt-Crossley contact, OpenSim CoordinateLimitForce,
 * SIMM splines, rigid-body dynamics); it is pinned by the first-principles
 * tests in tests/test_oracle_physics.py and, for forward kinematics, mass
 * distribution, rigid-body inverse dynamics, moment arms and the active
 * force-length-velocity curves, by the OpenSim-produced artefacts the
 * reference holds (tests/test_reference_artefacts.py: ScaleTool static pose to
 * 5e-7 m, IK solution, force plates, StaticOptimization).  Contact, limit
 * forces, tendon compliance, activation dynamics and the integrator have no
 * OpenSim-produced number anywhere and stay unpinned.
 * Everything that IS reference source (observation layout, reward,
 * termination, reset, action pre-processing) follows the cited lines.
 *
 * Formulation (deliberately different from the CUDA kernels' so that the
 * parity tests compare two derivations): world-frame spatial algebra about
 * a reference point O (the root body origin), composite-rigid-body mass
 * matrix + recursive Newton-Euler bias + dense Cholesky solve.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/bio_b200.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define MAXB BIO_MAX_BODIES
#define MAXD BIO_MAX_DOF
#define MAXM BIO_MAX_MUSCLES
#define MAXA BIO_MAX_ACT

typedef struct OrcEval {
    /* derivatives */
    double udot[MAXD];
    double adot[MAXM];
    double lmdot[MAXM];
    /* muscle read-outs */
    double tendon_force[MAXM];
    double fiber_force[MAXM];
    double active_fiber_force[MAXM];
    double path_len[MAXM];
    double path_vel[MAXM];
    /* forces */
    double contact[2][6];
    double limit_force[BIO_MAX_LIMITS];
    double mass_matrix[MAXD][MAXD];
    double bias[MAXD];
    /* kinematics for obs */
    double obs_pos[BIO_MAX_OBSPTS][3];
    double obs_vel[BIO_MAX_OBSPTS][3];
    double com_pos[3];
    double com_vel[3];
} OrcEval;

typedef struct OrcEnv {
    double q[MAXD];
    double u[MAXD];
    double act[MAXM];
    double lm[MAXM];
    double last_action[MAXA];
    double history[BIO_MAX_HORIZON][MAXA];
    double old_px;
    double ep_return;
    int32_t hist_pos;
    int32_t istep;
    int32_t first;
    int32_t ep_len;
    int64_t episode;
} OrcEnv;

/* ------------------------------------------------------------------ vec */
static inline void cross3(const double* a, const double* b, double* o) {
    double x = a[1] * b[2] - a[2] * b[1];
    double y = a[2] * b[0] - a[0] * b[2];
    double z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline double dot3(const double* a, const double* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static inline void matvec3(const double* R, const double* v, double* o) {
    double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void matmul3(const double* A, const double* B, double* C) {
    double t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(C, t, sizeof t);
}
static void rot_axis_angle(const double* a, double ang, double* R) {
    double c = cos(ang), s = sin(ang), v = 1 - c;
    R[0] = c + a[0] * a[0] * v;        R[1] = a[0] * a[1] * v - a[2] * s; R[2] = a[0] * a[2] * v + a[1] * s;
    R[3] = a[1] * a[0] * v + a[2] * s; R[4] = c + a[1] * a[1] * v;        R[5] = a[1] * a[2] * v - a[0] * s;
    R[6] = a[2] * a[0] * v - a[1] * s; R[7] = a[2] * a[1] * v + a[0] * s; R[8] = c + a[2] * a[2] * v;
}
/* spatial motion cross product  V x S  (both [w; v]) */
static void crm(const double* V, const double* S, double* o) {
    double a[3], b[3], c[3];
    cross3(V, S, a);
    cross3(V, S + 3, b);
    cross3(V + 3, S, c);
    o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
    o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}

/* -------------------------------------------------------------- functions */
/* SimmSpline / LinearFunction / Constant value, first and second derivative. */
void orc_func_eval(const BioModelTables* m, int f, double x, double* y, double* d1, double* d2) {
    int kind = m->func_kind[f];
    if (kind == BIO_FUNC_CONST) { *y = m->func_c[f][0]; *d1 = 0; *d2 = 0; return; }
    if (kind == BIO_FUNC_LINEAR) { *y = m->func_c[f][0] * x + m->func_c[f][1]; *d1 = m->func_c[f][0]; *d2 = 0; return; }
    int kb = m->func_knot_begin[f], n = m->func_knot_count[f];
    const double* kx = m->knot_x + kb;
    const double (*kc)[4] = m->knot_c + kb;
    if (x <= kx[0]) { *y = kc[0][0] + kc[0][1] * (x - kx[0]); *d1 = kc[0][1]; *d2 = 0; return; }
    if (x >= kx[n - 1]) { *y = kc[n - 1][0] + kc[n - 1][1] * (x - kx[n - 1]); *d1 = kc[n - 1][1]; *d2 = 0; return; }
    int i = 0;
    while (i + 1 < n - 1 && x >= kx[i + 1]) i++;
    double dx = x - kx[i];
    *y = kc[i][0] + dx * (kc[i][1] + dx * (kc[i][2] + dx * kc[i][3]));
    *d1 = kc[i][1] + dx * (2 * kc[i][2] + 3 * dx * kc[i][3]);
    *d2 = 2 * kc[i][2] + 6 * dx * kc[i][3];
}

/* Tabulated Millard curve (uniform cubic Hermite, linear extrapolation). */
void orc_curve_eval(const BioModelTables* m, int c, double x, double* y, double* dy) {
    const int n = BIO_CURVE_N;
    double x0 = m->curve_x0[c], x1 = m->curve_x1[c];
    double h = (x1 - x0) / n;
    const double (*tab)[2] = m->curve_tab[c];
    if (x < x0) { *dy = tab[0][1] / h; *y = tab[0][0] + *dy * (x - x0); return; }
    if (x > x1) { *dy = tab[n][1] / h; *y = tab[n][0] + *dy * (x - x1); return; }
    double t = (x - x0) / h;
    int i = (int)floor(t);
    if (i < 0) i = 0;
    if (i > n - 1) i = n - 1;
    double s = t - i, s2 = s * s, s3 = s2 * s;
    double y0 = tab[i][0], m0 = tab[i][1], y1 = tab[i + 1][0], m1 = tab[i + 1][1];
    *y = (2 * s3 - 3 * s2 + 1) * y0 + (s3 - 2 * s2 + s) * m0 + (-2 * s3 + 3 * s2) * y1 + (s3 - s2) * m1;
    *dy = ((6 * s2 - 6 * s) * y0 + (3 * s2 - 4 * s + 1) * m0 + (-6 * s2 + 6 * s) * y1 + (3 * s2 - 2 * s) * m1) / h;
}

/* ------------------------------------------------------------ kinematics */
typedef struct Kin {
    double O[3];
    double R[MAXB][9];
    double r[MAXB][3];       /* body origin relative to O */
    double V[MAXB][6];       /* spatial velocity about O */
    double Avp[MAXB][6];     /* velocity-product acceleration (+ gravity as base accel) */
    double S[MAXD][6];       /* effective motion axis of each dof */
} Kin;

static void kinematics(const BioModelTables* m, const double* q, const double* u, Kin* k) {
    memset(k->S, 0, sizeof k->S);
    for (int b = 0; b < m->n_bodies; b++) {
        int p = m->body_parent[b];
        double Rp[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, R[9], r[3], V[6] = {0}, A[6] = {0};
        if (p >= 0) {
            memcpy(Rp, k->R[p], sizeof Rp);
            matvec3(Rp, m->body_joint_loc[b], r);
            for (int i = 0; i < 3; i++) r[i] += k->r[p][i];
            memcpy(V, k->V[p], sizeof V);
            memcpy(A, k->Avp[p], sizeof A);
        } else {
            for (int i = 0; i < 3; i++) { r[i] = m->body_joint_loc[b][i]; A[3 + i] = -m->gravity[i]; }
        }
        memcpy(R, Rp, sizeof R);
        int root_open = (p < 0);
        int ab = m->body_axis_begin[b], ac = m->body_axis_count[b];
        for (int a = ab; a < ab + ac; a++) {
            int d = m->axis_dof[a];
            double s, ds, dds;
            orc_func_eval(m, m->axis_func[a], d >= 0 ? q[d] : 0.0, &s, &ds, &dds);
            double S[6], Sd[6], aw[3];
            if (m->axis_kind[a] == BIO_AXIS_TRANS) {
                matvec3(Rp, m->axis_vec[a], aw);
                S[0] = S[1] = S[2] = 0; S[3] = aw[0]; S[4] = aw[1]; S[5] = aw[2];
                for (int i = 0; i < 3; i++) r[i] += aw[i] * s;
            } else {
                if (root_open) { /* translations of the root are done: fix O here */
                    memcpy(k->O, r, sizeof r); r[0] = r[1] = r[2] = 0; root_open = 0;
                }
                matvec3(R, m->axis_vec[a], aw);
                S[0] = aw[0]; S[1] = aw[1]; S[2] = aw[2];
                cross3(r, aw, S + 3);
                double Rk[9];
                rot_axis_angle(m->axis_vec[a], s, Rk);
                matmul3(R, Rk, R);
            }
            if (d >= 0) {
                double qd = u[d], sd = ds * qd;
                crm(V, S, Sd);
                for (int i = 0; i < 6; i++) {
                    k->S[d][i] += ds * S[i];
                    A[i] += S[i] * (dds * qd * qd) + Sd[i] * sd;
                }
                for (int i = 0; i < 6; i++) V[i] += S[i] * sd;
            }
        }
        if (root_open) { memcpy(k->O, r, sizeof r); r[0] = r[1] = r[2] = 0; }
        memcpy(k->R[b], R, sizeof R);
        memcpy(k->r[b], r, sizeof r);
        memcpy(k->V[b], V, sizeof V);
        memcpy(k->Avp[b], A, sizeof A);
    }
}

static inline void point_pos(const Kin* k, int b, const double* loc, double* x) {
    matvec3(k->R[b], loc, x);
    for (int i = 0; i < 3; i++) x[i] += k->r[b][i];
}
static inline void point_vel(const Kin* k, int b, const double* x, double* v) {
    cross3(k->V[b], x, v);
    for (int i = 0; i < 3; i++) v[i] += k->V[b][3 + i];
}
static inline void add_force(double W[][6], int b, const double* x, const double* f) {
    double n[3];
    cross3(x, f, n);
    for (int i = 0; i < 3; i++) { W[b][i] += n[i]; W[b][3 + i] += f[i]; }
}

/* ---------------------------------------------------------------- muscle */
typedef struct MusOut { double T, Ffib, Fact, lmdot, adot; } MusOut;

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* Millard2012EquilibriumMuscle with elastic tendon and fibre damping. */
static void muscle_dynamics(const BioModelTables* m, int i, double L, double a, double lm, double e,
                            int newton_iters, double h_imp, MusOut* o) {
    double fiso = m->mus_fiso[i], lopt = m->mus_lopt[i], lts = m->mus_lts[i];
    double h = m->mus_height[i], beta = m->mus_beta[i], amin = m->mus_amin[i];
    double lmc = lm < m->mus_lm_min[i] ? m->mus_lm_min[i] : lm;
    double lat = sqrt(lmc * lmc - h * h);       /* fibre length along tendon */
    double cosa = lat / lmc;
    double lt = L - lat;
    double fal, dfal, fpe, dfpe, ft, dft, fv, dfv;
    orc_curve_eval(m, 0, lmc / lopt, &fal, &dfal);
    orc_curve_eval(m, 2, lmc / lopt, &fpe, &dfpe);
    orc_curve_eval(m, 3, lt / lts, &ft, &dft);
    double ac = clampd(a, amin, 1.0);
    double vn = 0.0, gain = 1.0, fsum = 0.0, derr = 1.0;
    fv = 1.0; dfv = 0.0;
    for (int it = 0; it < newton_iters; it++) {
        orc_curve_eval(m, 1, vn, &fv, &dfv);
        fsum = ac * fal * fv + fpe + beta * vn;
        double err = fsum * cosa - ft;
        derr = (ac * fal * dfv + beta) * cosa;
        double delta = -err / derr;
        vn += delta;
        if (fabs(delta) < 1e-12) break;
    }
    if (h_imp > 0) {
        /* Linearly implicit fibre-length update of the stated scheme: the fibre velocity of
           the damped equilibrium g(vn, lm) = 0 falls steeply with the fibre length where the passive element and
           the tendon are stretched, lambda = d(lmdot)/d(lm) = -vmax lopt g_lm / g_vn (measured down to -4.2e3 1/s
           for the glutei of the 3D models, h lambda = -2.1 at h = 0.5 ms: beyond explicit Euler).  The integrator
           advances lm by h lmdot / (1 - h lambda); g_lm and g_vn (= derr) are taken at the last Newton iterate:
             g_lm = (a fal' fv + fpe') cos(alpha) / lopt + (a fal fv + fpe + beta vn) sin^2(alpha) / (lm cos(alpha))
                    + ft' / (lts cos(alpha)),   clamped at >= 0 (descending limb of the active curve) */
        double glm = (ac * dfal * fv + dfpe) / lopt * cosa + fsum * (1.0 - cosa * cosa) / (lmc * cosa)
                     + dft / (lts * cosa);
        if (glm < 0) glm = 0;
        gain = derr / (derr + h_imp * m->mus_vmax[i] * lopt * glm);
    }
    if (lm <= m->mus_lm_min[i] && vn < 0) vn = 0.0;   /* clamped fibre cannot shorten */
    orc_curve_eval(m, 1, vn, &fv, &dfv);
    o->Fact = fiso * ac * fal * fv;
    o->Ffib = fiso * (ac * fal * fv + fpe + beta * vn);
    o->T = fiso * ft;
    o->lmdot = vn * m->mus_vmax[i] * lopt * gain;
    double ec = clampd(e, amin, 1.0);
    double tau = ec > ac ? m->mus_tact[i] * (0.5 + 1.5 * ac) : m->mus_tdeact[i] / (0.5 + 1.5 * ac);
    o->adot = (ec - ac) / tau;
}

/* Static fibre equilibrium (equilibrateMuscles, opensim_wrapper.py:290):
   activation kept, fibre velocity zero. Bisection on the bracket
   [lm_min, fibre length at which the tendon goes slack]. */
double orc_equilibrium_lm(const BioModelTables* m, int i, double L, double a) {
    double lopt = m->mus_lopt[i], lts = m->mus_lts[i], h = m->mus_height[i];
    double lo = m->mus_lm_min[i];
    double ac = clampd(a, m->mus_amin[i], 1.0);
    double slack = L - lts;
    if (slack <= 0) return lo;
    double hi = sqrt(slack * slack + h * h);
    if (hi <= lo) return lo;
    double glo;
    {
        double lat = sqrt(lo * lo - h * h), fal, d, fpe, ft;
        orc_curve_eval(m, 0, lo / lopt, &fal, &d);
        orc_curve_eval(m, 2, lo / lopt, &fpe, &d);
        orc_curve_eval(m, 3, (L - lat) / lts, &ft, &d);
        glo = (ac * fal + fpe) * (lat / lo) - ft;
    }
    if (glo >= 0) return lo;
    for (int it = 0; it < 60; it++) {
        double mid = 0.5 * (lo + hi);
        double lat = sqrt(mid * mid - h * h), fal, d, fpe, ft;
        orc_curve_eval(m, 0, mid / lopt, &fal, &d);
        orc_curve_eval(m, 2, mid / lopt, &fpe, &d);
        orc_curve_eval(m, 3, (L - lat) / lts, &ft, &d);
        double g = (ac * fal + fpe) * (lat / mid) - ft;
        if (g < 0) lo = mid; else hi = mid;
    }
    return 0.5 * (lo + hi);
}

/* Path geometry of muscle i (GeometryPath without wrap objects): active
   points, segment unit vectors, length and lengthening speed. */
typedef struct PathGeom {
    int np;
    int body[8], dof[8];          /* dof >= 0: moving point governed by that coordinate */
    double x[8][3], e[8][3], dw[8][3];
    double L, Ld;
} PathGeom;

static void muscle_geom(const BioModelTables* m, const Kin* k, const double* q, const double* u, int i,
                        PathGeom* g) {
    int pb = m->mus_pt_begin[i], pc = m->mus_pt_count[i];
    double vp[3] = {0, 0, 0};
    g->np = 0; g->L = 0; g->Ld = 0;
    for (int p = pb; p < pb + pc; p++) {
        int b = m->pt_body[p], kind = m->pt_kind[p], d = m->pt_dof[p];
        double loc[3], dloc[3] = {0, 0, 0};
        if (kind == BIO_PT_CONDITIONAL) {
            double v = q[d];
            if (!(v >= m->pt_range[p][0] - 1e-5 && v <= m->pt_range[p][1] + 1e-5)) continue;
        }
        if (kind == BIO_PT_MOVING) {
            for (int c = 0; c < 3; c++) {
                double d2;
                orc_func_eval(m, m->pt_func[p][c], q[d], &loc[c], &dloc[c], &d2);
            }
        } else {
            memcpy(loc, m->pt_loc[p], sizeof loc);
        }
        int n = g->np;
        double* x = g->x[n];
        double v[3];
        point_pos(k, b, loc, x);
        point_vel(k, b, x, v);
        g->body[n] = b; g->dof[n] = -1;
        g->dw[n][0] = g->dw[n][1] = g->dw[n][2] = 0;
        if (kind == BIO_PT_MOVING) {
            matvec3(k->R[b], dloc, g->dw[n]);
            for (int c = 0; c < 3; c++) v[c] += g->dw[n][c] * u[d];
            g->dof[n] = d;
        }
        if (n > 0) {
            double* e = g->e[n - 1];
            const double* xp = g->x[n - 1];
            for (int c = 0; c < 3; c++) e[c] = x[c] - xp[c];
            double len = sqrt(dot3(e, e));
            for (int c = 0; c < 3; c++) e[c] /= len;
            g->L += len;
            g->Ld += e[0] * (v[0] - vp[0]) + e[1] * (v[1] - vp[1]) + e[2] * (v[2] - vp[2]);
        }
        memcpy(vp, v, sizeof v);
        g->np = n + 1;
    }
}

/* Tension T along the path: equal and opposite forces at the ends of every
   segment, plus the generalized force of moving points (f . R_b dloc/dq). */
static void muscle_apply(const PathGeom* g, double T, double W[][6], double* Q) {
    for (int s = 0; s + 1 < g->np; s++) {
        double f[3] = {T * g->e[s][0], T * g->e[s][1], T * g->e[s][2]}, fn[3] = {-f[0], -f[1], -f[2]};
        add_force(W, g->body[s], g->x[s], f);
        add_force(W, g->body[s + 1], g->x[s + 1], fn);
        if (g->dof[s] >= 0) Q[g->dof[s]] += dot3(f, g->dw[s]);
        if (g->dof[s + 1] >= 0) Q[g->dof[s + 1]] += dot3(fn, g->dw[s + 1]);
    }
}

void orc_path_lengths(const BioModelTables* m, const double* q, const double* u, double* L, double* Ld) {
    Kin k;
    PathGeom g;
    kinematics(m, q, u, &k);
    for (int i = 0; i < m->n_muscles; i++) { muscle_geom(m, &k, q, u, i, &g); L[i] = g.L; Ld[i] = g.Ld; }
}

/* quintic smooth step (SimTK::Function::Step) */
static double step5(double x) {
    if (x <= 0) return 0;
    if (x >= 1) return 1;
    return x * x * x * (10 + x * (6 * x - 15));
}

/* ------------------------------------------------------------------------
 * One evaluation of the system dynamics at (q,u,act,lm) with zero-order-hold
 * controls: everything Manager.integrate evaluates per RHS
 * (opensim_wrapper.py:299-301) plus the read-outs of calc_* (:118-259).
 * ext_force: perturbation force on obs point ext_pt (env2D.py:83-100).
 * ---------------------------------------------------------------------- */
void orc_eval_h(const BioModelTables* m, int newton_iters, const double* q, const double* u,
                const double* act, const double* lm, const double* ctrl,
                const double* ext_force, int ext_pt, double h_imp, OrcEval* o) {
    Kin k;
    int nb = m->n_bodies, nd = m->n_dof;
    double W[MAXB][6];
    double Q[MAXD];
    /* h_imp > 0: velocity-Jacobian B of the dissipative forces (contact normal
       dissipation, friction, limit damping), B = sum J^T D J, D diagonal and
       >= 0; the solve becomes (M + h B) udot = rhs (linearly implicit in u). */
    double B[MAXD][MAXD];
    memset(B, 0, sizeof B);
    memset(W, 0, sizeof W);
    memset(Q, 0, sizeof Q);
    memset(o->contact, 0, sizeof o->contact);
    kinematics(m, q, u, &k);

    /* muscles */
    for (int i = 0; i < m->n_muscles; i++) {
        PathGeom g;
        MusOut mo;
        muscle_geom(m, &k, q, u, i, &g);
        muscle_dynamics(m, i, g.L, act[i], lm[i], ctrl[i], newton_iters, h_imp, &mo);
        muscle_apply(&g, mo.T, W, Q);
        o->path_len[i] = g.L; o->path_vel[i] = g.Ld;
        o->tendon_force[i] = mo.T; o->fiber_force[i] = mo.Ffib; o->active_fiber_force[i] = mo.Fact;
        o->lmdot[i] = mo.lmdot; o->adot[i] = mo.adot;
    }
    /* coordinate actuators */
    if (m->is_torque)
        for (int i = 0; i < m->n_act; i++)
            if (m->act_dof[i] >= 0) Q[m->act_dof[i]] += ctrl[i];
    /* Hunt-Crossley spheres on the half-space y<0 */
    for (int s = 0; s < m->n_spheres; s++) {
        int b = m->sph_body[s];
        double xc[3];
        point_pos(&k, b, m->sph_loc[s], xc);
        double R = m->sph_radius[s];
        double depth = R - (xc[1] + k.O[1]);
        if (depth <= 0) continue;
        double p[3] = {xc[0], -0.5 * depth - k.O[1], xc[2]}, v[3];
        point_vel(&k, b, p, v);
        double vn = -v[1];
        double kk = m->sph_k[s];
        double fH = (4.0 / 3.0) * kk * depth * sqrt(R * kk * depth);
        double f = fH * (1 + 1.5 * m->sph_c[s] * vn);
        if (f <= 0) continue;
        double F[3] = {0, f, 0};
        double vs = sqrt(v[0] * v[0] + v[2] * v[2]);
        if (vs != 0) {
            double vrel = vs / m->sph_vt[s];
            double ff = f * ((vrel < 1 ? vrel : 1) * (m->sph_ud[s] + 2 * (m->sph_us[s] - m->sph_ud[s]) / (1 + vrel * vrel))
                             + m->sph_uv[s] * vs);
            F[0] = -ff * v[0] / vs;
            F[2] = -ff * v[2] / vs;
        }
        add_force(W, b, p, F);
        if (h_imp > 0) {
            double vrel = vs / m->sph_vt[s];
            double gs = (vrel < 1 ? 1.0 / m->sph_vt[s] : 1.0 / vs) *
                            (m->sph_ud[s] + 2 * (m->sph_us[s] - m->sph_ud[s]) / (1 + vrel * vrel)) + m->sph_uv[s];
            double D[3] = {f * gs, 1.5 * m->sph_c[s] * fH, f * gs};
            int chain[MAXD], nc = 0, last = -1;
            for (int j = nd - 1; j >= 0 && last < 0; j--) {
                /* deepest dof whose body is b or an ancestor of b */
                int bb = b;
                while (bb >= 0 && last < 0) { if (m->dof_body[j] == bb) last = j; bb = m->body_parent[bb]; }
            }
            for (int j = last; j >= 0; j--) if ((m->dof_anc_mask[last] >> j) & 1u) chain[nc++] = j;
            double col[MAXD][3];
            for (int a = 0; a < nc; a++) {
                const double* S = k.S[chain[a]];
                cross3(S, p, col[a]);
                for (int c = 0; c < 3; c++) col[a][c] += S[3 + c];
            }
            for (int a = 0; a < nc; a++)
                for (int bq = 0; bq < nc; bq++)
                    B[chain[a]][chain[bq]] += D[0] * col[a][0] * col[bq][0] + D[1] * col[a][1] * col[bq][1] +
                                              D[2] * col[a][2] * col[bq][2];
        }
        double pa[3] = {p[0] + k.O[0], p[1] + k.O[1], p[2] + k.O[2]}, n[3];
        cross3(pa, F, n);
        int g = m->sph_group[s];
        for (int i = 0; i < 3; i++) { o->contact[g][i] += F[i]; o->contact[g][3 + i] += n[i]; }
    }
    /* coordinate limit forces */
    for (int l = 0; l < m->n_limits; l++) {
        int d = m->lim_dof[l];
        double w = m->lim_w[l], qq = q[d];
        double sup = step5((qq - m->lim_qup[l]) / w);
        double slo = 1.0 - step5((qq - (m->lim_qlo[l] - w)) / w);
        double f = -m->lim_kup[l] * sup * (qq - m->lim_qup[l]) + m->lim_klo[l] * slo * (m->lim_qlo[l] - qq)
                   - m->lim_damp[l] * (sup + slo) * u[d];
        o->limit_force[l] = f;
        Q[d] += f;
        if (h_imp > 0) B[d][d] += m->lim_damp[l] * (sup + slo);
    }
    /* perturbation */
    if (ext_force && ext_pt >= 0) {
        double x[3];
        point_pos(&k, m->obs_body[ext_pt], m->obs_loc[ext_pt], x);
        add_force(W, m->obs_body[ext_pt], x, ext_force);
    }

    /* spatial inertias about O in ground axes: mass, h = m*c, I_O */
    double Im[MAXB], Ih[MAXB][3], II[MAXB][9];
    double F[MAXB][6];
    double mtot = 0, msum[3] = {0, 0, 0}, psum[3] = {0, 0, 0};
    for (int b = 0; b < nb; b++) {
        double c[3], Ib[9], t[9], Rt[9];
        point_pos(&k, b, m->body_com[b], c);
        const double* i6 = m->body_inertia[b];
        double Ic[9] = {i6[0], i6[3], i6[4], i6[3], i6[1], i6[5], i6[4], i6[5], i6[2]};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rt[3 * i + j] = k.R[b][3 * j + i];
        matmul3(k.R[b], Ic, t);
        matmul3(t, Rt, Ib);
        double mb = m->body_mass[b], cc = dot3(c, c);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                Ib[3 * i + j] += mb * ((i == j ? cc : 0.0) - c[i] * c[j]);
        Im[b] = mb;
        for (int i = 0; i < 3; i++) Ih[b][i] = mb * c[i];
        memcpy(II[b], Ib, sizeof Ib);
        /* f = I*Avp + V x* (I V) - W */
        double IV[6], IA[6], t1[3], t2[3];
        const double* V = k.V[b]; const double* A = k.Avp[b];
        matvec3(Ib, V, IV); cross3(Ih[b], V + 3, t1);
        cross3(Ih[b], V, t2);
        for (int i = 0; i < 3; i++) { IV[i] += t1[i]; IV[3 + i] = mb * V[3 + i] - t2[i]; }
        matvec3(Ib, A, IA); cross3(Ih[b], A + 3, t1);
        cross3(Ih[b], A, t2);
        for (int i = 0; i < 3; i++) { IA[i] += t1[i]; IA[3 + i] = mb * A[3 + i] - t2[i]; }
        double c1[3], c2[3], c3[3];
        cross3(V, IV, c1); cross3(V + 3, IV + 3, c2); cross3(V, IV + 3, c3);
        for (int i = 0; i < 3; i++) {
            F[b][i] = IA[i] + c1[i] + c2[i] - W[b][i];
            F[b][3 + i] = IA[3 + i] + c3[i] - W[b][3 + i];
        }
        /* whole-body centre of mass */
        double vc[3];
        point_vel(&k, b, c, vc);
        mtot += mb;
        for (int i = 0; i < 3; i++) { msum[i] += mb * c[i]; psum[i] += mb * vc[i]; }
    }
    for (int i = 0; i < 3; i++) { o->com_pos[i] = msum[i] / mtot + k.O[i]; o->com_vel[i] = psum[i] / mtot; }
    /* composite inertias and subtree forces (children have larger indices) */
    for (int b = nb - 1; b > 0; b--) {
        int p = m->body_parent[b];
        if (p < 0) continue;
        Im[p] += Im[b];
        for (int i = 0; i < 3; i++) Ih[p][i] += Ih[b][i];
        for (int i = 0; i < 9; i++) II[p][i] += II[b][i];
        for (int i = 0; i < 6; i++) F[p][i] += F[b][i];
    }
    /* mass matrix and bias */
    double M[MAXD][MAXD];
    memset(M, 0, sizeof M);
    for (int i = 0; i < nd; i++) {
        int b = m->dof_body[i];
        const double* S = k.S[i];
        double IS[6], t1[3], t2[3];
        matvec3(II[b], S, IS); cross3(Ih[b], S + 3, t1); cross3(Ih[b], S, t2);
        for (int c = 0; c < 3; c++) { IS[c] += t1[c]; IS[3 + c] = Im[b] * S[3 + c] - t2[c]; }
        for (int j = 0; j <= i; j++) {
            if (!((m->dof_anc_mask[i] >> j) & 1u)) continue;
            double v = 0;
            for (int c = 0; c < 6; c++) v += k.S[j][c] * IS[c];
            M[i][j] = M[j][i] = v;
        }
        double bi = 0;
        for (int c = 0; c < 6; c++) bi += S[c] * F[b][c];
        o->bias[i] = bi - Q[i];
    }
    for (int i = 0; i < nd; i++) for (int j = 0; j < nd; j++) o->mass_matrix[i][j] = M[i][j];
    if (h_imp > 0)
        for (int i = 0; i < nd; i++) for (int j = 0; j < nd; j++) M[i][j] += h_imp * B[i][j];
    /* Cholesky solve M udot = -bias */
    double Lc[MAXD][MAXD];
    memset(Lc, 0, sizeof Lc);
    for (int i = 0; i < nd; i++) {
        for (int j = 0; j <= i; j++) {
            double s = M[i][j];
            for (int c = 0; c < j; c++) s -= Lc[i][c] * Lc[j][c];
            if (i == j) Lc[i][i] = sqrt(s); else Lc[i][j] = s / Lc[j][j];
        }
    }
    double y[MAXD];
    for (int i = 0; i < nd; i++) {
        double s = -o->bias[i];
        for (int c = 0; c < i; c++) s -= Lc[i][c] * y[c];
        y[i] = s / Lc[i][i];
    }
    for (int i = nd - 1; i >= 0; i--) {
        double s = y[i];
        for (int c = i + 1; c < nd; c++) s -= Lc[c][i] * o->udot[c];
        o->udot[i] = s / Lc[i][i];
    }
    /* obs points */
    for (int p = 0; p < m->n_obspts; p++) {
        double x[3];
        point_pos(&k, m->obs_body[p], m->obs_loc[p], x);
        point_vel(&k, m->obs_body[p], x, o->obs_vel[p]);
        for (int i = 0; i < 3; i++) o->obs_pos[p][i] = x[i] + k.O[i];
    }
}

void orc_eval(const BioModelTables* m, int newton_iters, const double* q, const double* u,
              const double* act, const double* lm, const double* ctrl,
              const double* ext_force, int ext_pt, OrcEval* o) {
    orc_eval_h(m, newton_iters, q, u, act, lm, ctrl, ext_force, ext_pt, 0.0, o);
}

/* --------------------------------------------------------------- helpers */
uint64_t orc_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
/* counter-based draw keyed by (seed, global env index, episode, stream) */
uint64_t orc_rand(uint64_t seed, uint64_t env, uint64_t episode, uint64_t stream) {
    uint64_t h = orc_splitmix64(seed ^ orc_splitmix64(env));
    h = orc_splitmix64(h + episode);
    return orc_splitmix64(h ^ (stream * 0xD6E8FEB86659FD93ull));
}

/* Perturbation force at time t (env2D.py:83-100): 100 knots over 10 s,
   piecewise constant, +-50 N where fmod(t_knot,2) > thresh. */
static double perturb_force(const BioTaskConfig* c, uint64_t seed, uint64_t env, double t) {
    if (!c->perturb) return 0.0;
    double dtk = 10.0 / 99.0;
    int kidx = (int)ceil(t / dtk - 1e-12);
    if (kidx < 0) kidx = 0;
    if (kidx > 99) kidx = 99;
    double tk = kidx * dtk;
    if (!(fmod(tk, 2.0) > c->perturb_thresh)) return 0.0;
    if (c->perturb_negative_only) return -c->perturb_force;
    return (orc_rand(seed, env, (uint64_t)kidx, 7) & 1) ? c->perturb_force : -c->perturb_force;
}

static void eval_env_h(const BioModelTables* m, const BioTaskConfig* c, const double* q, const double* u,
                       const double* act, const double* lm, const double* ctrl, double t,
                       uint64_t seed, uint64_t env, double h_imp, OrcEval* ev) {
    double fx[3] = {perturb_force(c, seed, env, t), 0, 0};
    orc_eval_h(m, c->newton_iters, q, u, act, lm, ctrl, c->perturb ? fx : NULL,
               c->perturb ? c->perturb_obspt : -1, h_imp, ev);
}
static void eval_env(const BioModelTables* m, const BioTaskConfig* c, const double* q, const double* u,
                     const double* act, const double* lm, const double* ctrl, double t,
                     uint64_t seed, uint64_t env, OrcEval* ev) {
    eval_env_h(m, c, q, u, act, lm, ctrl, t, seed, env, 0.0, ev);
}

typedef struct Deriv { double qd[MAXD], ud[MAXD], ad[MAXM], ld[MAXM]; } Deriv;

static void post_clamp(const BioModelTables* m, OrcEnv* e) {
    for (int i = 0; i < m->n_muscles; i++) {
        e->act[i] = clampd(e->act[i], m->mus_amin[i], 1.0);
        if (e->lm[i] < m->mus_lm_min[i]) e->lm[i] = m->mus_lm_min[i];
    }
}

/* Error-controlled Runge-Kutta-Merson over one control step: the stand-in for the reference's
   opensim.Manager default integrator at its accuracy of 1e-3 (muscle_walking_imitation_env2D.py:41-42,
   opensim_wrapper.py:287-301).  CPU baseline only ("restated reference algorithm, not OpenSim"); the
   number of right-hand-side evaluations is returned for the report. */
#define ORC_ADAPTIVE_ACCURACY 1e-3
static void pack_state(const BioModelTables* m, const OrcEnv* e, double* y) {
    int nd = m->n_dof, nm = m->n_muscles;
    for (int i = 0; i < nd; i++) { y[i] = e->q[i]; y[nd + i] = e->u[i]; }
    for (int i = 0; i < nm; i++) { y[2 * nd + i] = e->act[i]; y[2 * nd + nm + i] = e->lm[i]; }
}
static void unpack_state(const BioModelTables* m, const double* y, OrcEnv* e) {
    int nd = m->n_dof, nm = m->n_muscles;
    for (int i = 0; i < nd; i++) { e->q[i] = y[i]; e->u[i] = y[nd + i]; }
    for (int i = 0; i < nm; i++) { e->act[i] = y[2 * nd + i]; e->lm[i] = y[2 * nd + nm + i]; }
}
static void rhs_packed(const BioModelTables* m, const BioTaskConfig* c, const double* y, const double* ctrl, double t,
                       uint64_t seed, uint64_t env, double* dy) {
    int nd = m->n_dof, nm = m->n_muscles;
    OrcEnv st;
    OrcEval ev;
    unpack_state(m, y, &st);
    post_clamp(m, &st);
    eval_env(m, c, st.q, st.u, st.act, st.lm, ctrl, t, seed, env, &ev);
    for (int i = 0; i < nd; i++) { dy[i] = st.u[i]; dy[nd + i] = ev.udot[i]; }
    for (int i = 0; i < nm; i++) { dy[2 * nd + i] = ev.adot[i]; dy[2 * nd + nm + i] = ev.lmdot[i]; }
}
int orc_last_adaptive_evals = 0;
static void integrate_adaptive(const BioModelTables* m, const BioTaskConfig* c, OrcEnv* e, const double* ctrl,
                               uint64_t seed, uint64_t env) {
    enum { NY = 2 * MAXD + 2 * MAXM };
    int n = 2 * m->n_dof + 2 * m->n_muscles, evals = 0;
    double y[NY], k1[NY], k2[NY], k3[NY], k4[NY], k5[NY], yt[NY];
    double t = e->istep * c->dt, tend = t + c->dt, h = c->dt / 4;
    pack_state(m, e, y);
    while (tend - t > 1e-12 && evals < 20000) {
        if (h > tend - t) h = tend - t;
        rhs_packed(m, c, y, ctrl, t, seed, env, k1);
        for (int i = 0; i < n; i++) yt[i] = y[i] + h / 3 * k1[i];
        rhs_packed(m, c, yt, ctrl, t + h / 3, seed, env, k2);
        for (int i = 0; i < n; i++) yt[i] = y[i] + h / 6 * (k1[i] + k2[i]);
        rhs_packed(m, c, yt, ctrl, t + h / 3, seed, env, k3);
        for (int i = 0; i < n; i++) yt[i] = y[i] + h / 8 * (k1[i] + 3 * k3[i]);
        rhs_packed(m, c, yt, ctrl, t + h / 2, seed, env, k4);
        for (int i = 0; i < n; i++) yt[i] = y[i] + h * (0.5 * k1[i] - 1.5 * k3[i] + 2 * k4[i]);
        rhs_packed(m, c, yt, ctrl, t + h, seed, env, k5);
        evals += 5;
        double err = 0;
        for (int i = 0; i < n; i++) {
            double ei = fabs(h / 30 * (2 * k1[i] - 9 * k3[i] + 8 * k4[i] - k5[i]));
            double sc = fabs(y[i]) > 0.1 ? fabs(y[i]) : 0.1;     /* relative, with a floor for states near zero */
            if (ei / sc > err) err = ei / sc;
        }
        if (err <= ORC_ADAPTIVE_ACCURACY || h < 1e-7) {
            for (int i = 0; i < n; i++) y[i] += h / 6 * (k1[i] + 4 * k4[i] + k5[i]);
            t += h;
            OrcEnv st = *e;
            unpack_state(m, y, &st);
            post_clamp(m, &st);
            pack_state(m, &st, y);
        }
        double fac = err > 0 ? 0.9 * pow(ORC_ADAPTIVE_ACCURACY / err, 0.25) : 4.0;
        if (fac < 0.2) fac = 0.2;
        if (fac > 4.0) fac = 4.0;
        h *= fac;
    }
    unpack_state(m, y, e);
    post_clamp(m, e);
    orc_last_adaptive_evals = evals;
}

/* Advance one env by dt with n_substeps fixed steps of the stated scheme. */
static void integrate(const BioModelTables* m, const BioTaskConfig* c, OrcEnv* e, const double* ctrl,
                      uint64_t seed, uint64_t env) {
    int nd = m->n_dof, nm = m->n_muscles;
    if (c->integrator == BIO_INT_ADAPTIVE_RKM) { integrate_adaptive(m, c, e, ctrl, seed, env); return; }
    double h = c->dt / c->n_substeps;
    double t0 = e->istep * c->dt;
    OrcEval ev;
    for (int s = 0; s < c->n_substeps; s++) {
        double t = t0 + s * h;
        if (c->integrator == BIO_INT_SEMI_IMPLICIT_EULER || c->integrator == BIO_INT_IMPLICIT_DAMPING) {
            eval_env_h(m, c, e->q, e->u, e->act, e->lm, ctrl, t, seed, env,
                       c->integrator == BIO_INT_IMPLICIT_DAMPING ? h : 0.0, &ev);
            for (int i = 0; i < nd; i++) { e->u[i] += h * ev.udot[i]; e->q[i] += h * e->u[i]; }
            for (int i = 0; i < nm; i++) { e->act[i] += h * ev.adot[i]; e->lm[i] += h * ev.lmdot[i]; }
        } else if (c->integrator == BIO_INT_RK2_MIDPOINT) {
            OrcEnv mid = *e;
            eval_env(m, c, e->q, e->u, e->act, e->lm, ctrl, t, seed, env, &ev);
            for (int i = 0; i < nd; i++) { mid.q[i] = e->q[i] + 0.5 * h * e->u[i]; mid.u[i] = e->u[i] + 0.5 * h * ev.udot[i]; }
            for (int i = 0; i < nm; i++) { mid.act[i] = e->act[i] + 0.5 * h * ev.adot[i]; mid.lm[i] = e->lm[i] + 0.5 * h * ev.lmdot[i]; }
            post_clamp(m, &mid);
            eval_env(m, c, mid.q, mid.u, mid.act, mid.lm, ctrl, t + 0.5 * h, seed, env, &ev);
            for (int i = 0; i < nd; i++) { e->q[i] += h * mid.u[i]; e->u[i] += h * ev.udot[i]; }
            for (int i = 0; i < nm; i++) { e->act[i] += h * ev.adot[i]; e->lm[i] += h * ev.lmdot[i]; }
        } else { /* classic RK4 */
            Deriv kk[4];
            OrcEnv st = *e;
            const double cs[4] = {0, 0.5, 0.5, 1.0};
            for (int r = 0; r < 4; r++) {
                if (r > 0) {
                    for (int i = 0; i < nd; i++) { st.q[i] = e->q[i] + cs[r] * h * kk[r - 1].qd[i]; st.u[i] = e->u[i] + cs[r] * h * kk[r - 1].ud[i]; }
                    for (int i = 0; i < nm; i++) { st.act[i] = e->act[i] + cs[r] * h * kk[r - 1].ad[i]; st.lm[i] = e->lm[i] + cs[r] * h * kk[r - 1].ld[i]; }
                    post_clamp(m, &st);
                }
                eval_env(m, c, st.q, st.u, st.act, st.lm, ctrl, t + cs[r] * h, seed, env, &ev);
                for (int i = 0; i < nd; i++) { kk[r].qd[i] = st.u[i]; kk[r].ud[i] = ev.udot[i]; }
                for (int i = 0; i < nm; i++) { kk[r].ad[i] = ev.adot[i]; kk[r].ld[i] = ev.lmdot[i]; }
            }
            for (int i = 0; i < nd; i++) {
                e->q[i] += h / 6 * (kk[0].qd[i] + 2 * kk[1].qd[i] + 2 * kk[2].qd[i] + kk[3].qd[i]);
                e->u[i] += h / 6 * (kk[0].ud[i] + 2 * kk[1].ud[i] + 2 * kk[2].ud[i] + kk[3].ud[i]);
            }
            for (int i = 0; i < nm; i++) {
                e->act[i] += h / 6 * (kk[0].ad[i] + 2 * kk[1].ad[i] + 2 * kk[2].ad[i] + kk[3].ad[i]);
                e->lm[i] += h / 6 * (kk[0].ld[i] + 2 * kk[1].ld[i] + 2 * kk[2].ld[i] + kk[3].ld[i]);
            }
        }
        post_clamp(m, e);
    }
}

static int ref_row(const BioTaskConfig* c, const BioRefTables* ref, int idx) {
    if (c->ref_mirror && c->cycle > 0 && idx > c->cycle) idx = 2 * c->cycle - idx;
    if (idx < 0) idx = 0;
    if (idx > ref->n_rows - 1) idx = ref->n_rows - 1;
    return idx;
}

/* Observation vector (env2D.py:158-230, App. C layout). Returns length. */
static int write_obs(const BioModelTables* m, const BioTaskConfig* c, const BioRefTables* ref,
                     const OrcEnv* e, const OrcEval* ev, double* obs) {
    int o = 0;
    double ph = (double)e->istep / c->cycle;
    obs[o++] = ph - floor(ph);
    double pel[3] = {0, 0, 0};
    int has[3] = {0, 0, 0};
    for (int i = 0; i < m->n_coords; i++) {
        int pt = m->coord_pelvis_trans[i];
        if (pt) { pel[pt - 1] = e->q[m->coord_dof[i]]; has[pt - 1] = 1; }
    }
    for (int i = 0; i < m->n_coords; i++)
        if (!m->coord_pelvis_trans[i]) obs[o++] = m->coord_dof[i] >= 0 ? e->q[m->coord_dof[i]] : m->coord_const[i];
    for (int i = 0; i < m->n_coords; i++) obs[o++] = m->coord_dof[i] >= 0 ? e->u[m->coord_dof[i]] : 0.0;
    for (int i = 0; i < m->n_coords; i++) obs[o++] = m->coord_dof[i] >= 0 ? ev->udot[m->coord_dof[i]] : 0.0;
    if (c->use_target_obs) {
        int row = ref_row(c, ref, e->istep + 1);
        for (int i = 0; i < m->n_coords; i++)
            if (m->coord_pelvis_trans[i] != 1) obs[o++] = ref->q[(size_t)row * ref->n_coords + i];
        for (int i = 0; i < m->n_coords; i++)
            if (m->coord_pelvis_trans[i] != 1) obs[o++] = ref->u[(size_t)row * ref->n_coords + i];
    }
    for (int p = 0; p < c->n_obs_bodies; p++)
        for (int i = 0; i < 3; i++) obs[o++] = ev->obs_pos[p][i] - (has[i] ? pel[i] : 0.0);
    for (int i = 0; i < 3; i++) obs[o++] = ev->com_pos[i] - (has[i] ? pel[i] : 0.0);
    for (int p = 0; p < c->n_obs_body_vel; p++)
        for (int i = 0; i < 3; i++) obs[o++] = ev->obs_vel[p][i];
    for (int i = 0; i < 3; i++) obs[o++] = ev->com_vel[i];
    for (int i = 0; i < m->n_muscles; i++) { obs[o++] = e->act[i]; obs[o++] = e->lm[i]; obs[o++] = ev->lmdot[i]; }
    if (c->use_grf) {
        double weight = fabs(m->total_mass * m->gravity[1]);
        double moment = weight * c->height;
        for (int g = 0; g < 2; g++) {
            for (int i = 0; i < 3; i++) obs[o++] = ev->contact[g][i] / weight;
            for (int i = 0; i < 3; i++) obs[o++] = ev->contact[g][3 + i] / moment;
        }
    }
    return o;
}

int orc_obs_dim(const BioModelTables* m, const BioTaskConfig* c) {
    int np = 0, ntx = 0;
    for (int i = 0; i < m->n_coords; i++) { if (m->coord_pelvis_trans[i]) np++; if (m->coord_pelvis_trans[i] == 1) ntx++; }
    int d = 1 + (m->n_coords - np) + 2 * m->n_coords;
    if (c->use_target_obs) d += 2 * (m->n_coords - ntx);
    d += 3 * (c->n_obs_bodies + 1) + 3 * (c->n_obs_body_vel + 1) + 3 * m->n_muscles;
    if (c->use_grf) d += 12;
    return d;
}

static void controls_zero(const BioModelTables* m, double* ctrl) {
    for (int i = 0; i < m->n_act; i++) ctrl[i] = 0.0;
}

/* Reference-state reset (env2D.py:133-156; opensim_wrapper.py:287-332). */
void orc_reset_env(const BioModelTables* m, const BioTaskConfig* c, const BioRefTables* ref,
                   OrcEnv* e, uint64_t seed, uint64_t env, double* obs) {
    int idx = 0;
    if (!c->test_mode && c->reset_max_index > 0)
        idx = (int)(orc_rand(seed, env, (uint64_t)e->episode, 1) % (uint64_t)(c->reset_max_index + 1));
    if (idx > ref->n_rows - 1) idx = ref->n_rows - 1;
    e->istep = idx;
    e->first = 1;
    e->ep_return = 0; e->ep_len = 0;
    for (int i = 0; i < m->n_coords; i++) {
        int d = m->coord_dof[i];
        if (d < 0) continue;
        e->q[d] = ref->q[(size_t)idx * ref->n_coords + i];
        e->u[d] = ref->u[(size_t)idx * ref->n_coords + i];
    }
    double L[MAXM], Ld[MAXM];
    orc_path_lengths(m, e->q, e->u, L, Ld);
    for (int i = 0; i < m->n_muscles; i++) {
        e->act[i] = m->mus_default_act[i];
        e->lm[i] = orc_equilibrium_lm(m, i, L[i], e->act[i]);
    }
    memset(e->last_action, 0, sizeof e->last_action);
    memset(e->history, 0, sizeof e->history);
    e->hist_pos = 0;
    /* old_pos_pelvisx is NOT cleared by the reference (App. E.6) */
    if (obs) {
        double ctrl[MAXA];
        OrcEval ev;
        controls_zero(m, ctrl);
        eval_env(m, c, e->q, e->u, e->act, e->lm, ctrl, e->istep * c->dt, seed, env, &ev);
        write_obs(m, c, ref, e, &ev, obs);
    }
}

void orc_init_env(OrcEnv* e) { memset(e, 0, sizeof *e); }

static double body_mse(const double* cur, const double* des) {
    double s = 0;
    for (int i = 0; i < 3; i++) { double d = cur[i] - des[i]; s += d * d; }
    return s / 3.0;
}

/* One control step of one env: OsimEnv.step (opensim_environment.py:100-113)
   with the subclass step (env2D.py:115-131 / torque env2D.py:117-149),
   get_reward (:267-358) and is_done (:237-265).
   Returns the done-reason flags (0 = not done). */
int orc_step_env(const BioModelTables* m, const BioTaskConfig* c, const BioRefTables* ref, OrcEnv* e,
                 uint64_t seed, uint64_t env, const double* action_in, double* obs, double* reward,
                 double* terms, OrcEval* ev_out) {
    int na = m->n_act, nm = m->n_muscles;
    double action[MAXA], ctrl[MAXA], curr[MAXA];
    int nan = 0;
    for (int i = 0; i < na; i++) if (isnan(action_in[i])) nan = 1;
    for (int i = 0; i < na; i++) action[i] = nan ? 0.0 : action_in[i];
    OrcEval ev;
    if (c->use_pd && !nan) {
        /* PD law on the pre-step observation (torque env2D.py:125-139) */
        double tau[MAXA];
        for (int i = 0; i < c->n_pd; i++) {
            int cx = c->pd_x_coord[i], cv = c->pd_v_coord[i];
            double x = m->coord_dof[cx] >= 0 ? e->q[m->coord_dof[cx]] : m->coord_const[cx];
            double v = m->coord_dof[cv] >= 0 ? e->u[m->coord_dof[cv]] : 0.0;
            tau[i] = c->pd_kp[i] * (action[i] - x) + c->pd_kv[i] * (-v);
        }
        for (int i = 0; i < na; i++) action[i] = i < c->n_pd ? tau[i] : 0.0;
    }
    if (e->first) {
        for (int i = 0; i < na; i++) e->last_action[i] = action[i];
        for (int hh = 0; hh < c->horizon; hh++) for (int i = 0; i < na; i++) e->history[hh][i] = action[i];
        e->hist_pos = 0;
        e->first = 0;
    }
    for (int i = 0; i < na; i++) e->history[e->hist_pos][i] = action[i];
    e->hist_pos = (e->hist_pos + 1) % c->horizon;
    for (int i = 0; i < na; i++) {
        double s = 0;
        for (int hh = 0; hh < c->horizon; hh++) s += e->history[hh][i];
        curr[i] = s / c->horizon;
    }
    /* actuate: clip to [min_control, max_control] (opensim_wrapper.py:92-107) */
    for (int i = 0; i < na; i++) ctrl[i] = clampd(c->feed_mean_action ? curr[i] : action[i], m->act_min[i], m->act_max[i]);
    /* integrate (opensim_wrapper.py:299-301) */
    integrate(m, c, e, ctrl, seed, env);
    e->istep += 1;
    eval_env(m, c, e->q, e->u, e->act, e->lm, ctrl, e->istep * c->dt, seed, env, &ev);
    if (ev_out) *ev_out = ev;
    if (obs) write_obs(m, c, ref, e, &ev, obs);

    /* reward */
    int row = ref_row(c, ref, e->istep);
    const double* qd = ref->q + (size_t)row * ref->n_coords;
    double qerr = 0, px = 0, py = 0;
    for (int i = 0; i < m->n_coords; i++) {
        int d = m->coord_dof[i];
        double v = d >= 0 ? e->q[d] : m->coord_const[i];
        double dd = v - qd[i];
        qerr += dd * dd;
        if (m->coord_pelvis_trans[i] == 1) px = v;
        if (m->coord_pelvis_trans[i] == 2) py = v;
    }
    qerr /= m->n_coords;
    double com_err = body_mse(ev.com_pos, ref->com_pos + (size_t)row * 3);
    double position_r = exp(-30.0 * qerr);
    double com_r = exp(-20.0 * com_err);
    double foot[2];
    for (int s = 0; s < 2; s++) {
        double sum = 0;
        for (int j = 0; j < 4; j++)
            sum += body_mse(ev.obs_pos[c->rew_obspt[s][j]],
                            ref->body_pos + ((size_t)row * ref->n_refbodies + c->rew_refbody[s][j]) * 3);
        foot[s] = 0.5 * exp(-20.0 * sum);
    }
    double foot_r = foot[0], foot_l = foot[1];
    double effort, a_error = 0;
    if (c->effort_torque) {
        double s = 0;
        for (int i = 0; i < na; i++) s += curr[i] * curr[i];
        effort = sqrt(s) / (c->max_actuation * na * na);
    } else {
        double s = 0, total = 1.51 * m->total_mass;
        for (int i = 0; i < nm; i++) {
            s += e->act[i] * e->act[i];
            double l = m->mus_slow_twitch[i], ex = ctrl[i], ac = e->act[i];
            double fa = 40 * l * sin(0.5 * M_PI * ex) + 133 * (1 - l) * (1 - cos(0.5 * M_PI * ex));
            double fm = 74 * l * sin(0.5 * M_PI * ac) + 111 * (1 - l) * (1 - cos(0.5 * M_PI * ac));
            double ln = e->lm[i] / m->mus_lopt[i], v = ev.lmdot[i], g = 0;
            if (ln < 0.5) g = 0.5; else if (ln < 1.0) g = ln; else if (ln < 1.5) g = -2 * ln + 3;
            double es = 0.25 * ev.fiber_force[i] * -v, ew = ev.active_fiber_force[i] * -v;
            total += m->mus_cot_mass[i] * fa + m->mus_cot_mass[i] * g * fm + (es > 0 ? es : 0) + (ew > 0 ? ew : 0);
        }
        a_error = exp(-2 * sqrt(s));
        effort = total / (20.0 * nm * nm);
    }
    double prog = (c->effort_use_dy ? py : px) - e->old_px + 1.0;
    double effort_r = exp(-effort / (prog > 1.0 ? prog : 1.0));
    double dn = 0;
    for (int i = 0; i < na; i++) { double d = curr[i] - e->last_action[i]; dn += d * d; }
    double action_r = exp(-c->action_r_scale * sqrt(dn));
    double imit = position_r * com_r;
    if (c->reward_use_feet) imit *= (foot_l + foot_r);
    double rew = (0.5 + c->w_imitate) * imit + c->w_effort * effort_r + c->w_action * action_r;
    for (int i = 0; i < na; i++) e->last_action[i] = curr[i];
    e->old_px = c->effort_use_dy ? py : px;
    if (terms) {
        terms[0] = position_r; terms[1] = com_r; terms[2] = foot_l; terms[3] = foot_r;
        if (c->n_reward_terms > 4) terms[4] = a_error;
    }
    /* termination */
    int done = 0;
    double maxlim = 0, maxacc = 0;
    for (int l = 0; l < m->n_limits; l++) if (fabs(ev.limit_force[l]) > maxlim) maxlim = fabs(ev.limit_force[l]);
    for (int i = 0; i < m->n_dof; i++) if (fabs(ev.udot[i]) > maxacc) maxacc = fabs(ev.udot[i]);
    int finite = isfinite(rew);
    for (int i = 0; i < m->n_dof; i++) finite = finite && isfinite(e->q[i]) && isfinite(e->u[i]) && isfinite(ev.udot[i]);
    if (!finite) {
        done = BIO_DONE_NONFINITE; rew = 0;
        if (terms) for (int i = 0; i < c->n_reward_terms; i++) terms[i] = 0;
    }
    else if (ev.obs_pos[c->term_obspt][1] < c->term_height) done = BIO_DONE_HEIGHT;
    else if (maxlim > c->term_limit_force) done = BIO_DONE_LIMIT_FORCE;
    else if (maxacc > c->term_acc) done = BIO_DONE_ACCEL;
    else if (e->istep >= c->n_steps) done = BIO_DONE_HORIZON;
    else if (c->term_feet_cross && ev.obs_pos[c->feet_obspt[0]][2] - ev.obs_pos[c->feet_obspt[1]][2] < 0) done = BIO_DONE_FEET_CROSS;
    *reward = rew;
    e->ep_return += rew; e->ep_len += 1;
    return done;
}

/* Batch step with auto-reset over a slice of envs (callers thread over slices).
   actions [N][n_act], obs [N][obs_dim], reward [N], done [N], terms [N][n_terms]. */
void orc_batch_step(const BioModelTables* m, const BioTaskConfig* c, const BioRefTables* ref, OrcEnv* envs,
                    int n, uint64_t seed, int64_t env_offset, const double* actions, double* obs,
                    double* reward, uint8_t* done, double* terms, int* reasons) {
    int od = orc_obs_dim(m, c);
    for (int i = 0; i < n; i++) {
        double tbuf[8];
        int r = orc_step_env(m, c, ref, &envs[i], seed, (uint64_t)(env_offset + i), actions + (size_t)i * m->n_act,
                             obs + (size_t)i * od, &reward[i], tbuf, NULL);
        if (terms) for (int k = 0; k < c->n_reward_terms; k++) terms[(size_t)i * c->n_reward_terms + k] = tbuf[k];
        done[i] = r != 0;
        if (reasons) reasons[i] = r;
        if (r && c->auto_reset) {
            envs[i].episode += 1;
            orc_reset_env(m, c, ref, &envs[i], seed, (uint64_t)(env_offset + i), obs + (size_t)i * od);
        }
    }
}

/* The same over all envs on `threads` OpenMP threads (the CPU arms of bench.py; built with -fopenmp in the
   "fast" build only, plain loop otherwise). */
void orc_batch_step_mt(const BioModelTables* m, const BioTaskConfig* c, const BioRefTables* ref, OrcEnv* envs,
                       int n, uint64_t seed, int64_t env_offset, const double* actions, double* obs,
                       double* reward, uint8_t* done, double* terms, int* reasons, int threads) {
    int od = orc_obs_dim(m, c);
    const int chunk = 8;
    const int nchunks = (n + chunk - 1) / chunk;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int k = 0; k < nchunks; k++) {
        const int i0 = k * chunk, cnt = (i0 + chunk <= n) ? chunk : n - i0;
        orc_batch_step(m, c, ref, envs + i0, cnt, seed, env_offset + i0, actions + (size_t)i0 * m->n_act,
                       obs + (size_t)i0 * od, reward + i0, done + i0,
                       terms ? terms + (size_t)i0 * c->n_reward_terms : NULL, reasons ? reasons + i0 : NULL);
    }
}

int orc_openmp(void) {
#ifdef _OPENMP
    return 1;
#else
    return 0;
#endif
}

void orc_batch_reset(const BioModelTables* m, const BioTaskConfig* c, const BioRefTables* ref, OrcEnv* envs,
                     int n, uint64_t seed, int64_t env_offset, double* obs, int bump_episode) {
    int od = orc_obs_dim(m, c);
    for (int i = 0; i < n; i++) {
        if (bump_episode) envs[i].episode += 1;
        orc_reset_env(m, c, ref, &envs[i], seed, (uint64_t)(env_offset + i), obs ? obs + (size_t)i * od : NULL);
    }
}

uint64_t orc_sizeof_env(void) { return sizeof(OrcEnv); }
uint64_t orc_sizeof_eval(void) { return sizeof(OrcEval); }
uint64_t orc_sizeof_model_tables(void) { return sizeof(BioModelTables); }
uint64_t orc_sizeof_task_config(void) { return sizeof(BioTaskConfig); }
