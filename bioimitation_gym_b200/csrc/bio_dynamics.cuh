// bio_dynamics.cuh -- per-env dynamics evaluation on the device.
//
// One thread owns one env.  Replaces, per right-hand-side evaluation, what
// OpenSim/Simbody compute inside Manager.integrate (reference
// opensim_wrapper.py:299-301): forward kinematics of the CustomJoint/PinJoint
// tree, GeometryPath lengths and force application, the
// Millard2012EquilibriumMuscle damped-equilibrium fibre velocity and
// activation ODE, HuntCrossleyForce sphere/half-space contact,
// CoordinateLimitForce, CoordinateActuator torques, and the multibody forward
// dynamics (composite-rigid-body mass matrix, recursive Newton-Euler bias,
// sparse L^T D L solve along the kinematic tree).
//
// Spatial quantities are expressed in ground axes about a per-evaluation
// reference point O = origin of the root body, so fp32 never subtracts large
// absolute positions.
#pragma once
#include <math.h>

#include "bio_model.cuh"

namespace bio {

#define BIO_DEV __host__ __device__ __forceinline__

// sin and cos of a float to fp32 rounding level for |x| up to a few hundred radians (joint angles): two-term
// Cody-Waite reduction by pi/2, then the single-precision minimax polynomials on [-pi/4, pi/4] (Cephes
// coefficients).  Replaces sincosf, whose large-argument path costs three times the instructions.
static __host__ __device__ __forceinline__ void sincos_f32(float x, float* sn, float* cs) {
    const float k = rintf(x * 0.636619772f);
    float r = fmaf(k, -1.57079637e+0f, x);
    r = fmaf(k, 4.37113883e-8f, r);
    const float z = r * r;
    const float ps = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
    const float sr = fmaf(r * z, ps, r);
    const float pc = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
    const float cr = fmaf(z * z, pc, fmaf(-0.5f, z, 1.0f));
    const int n = (int)k;
    const float s0 = (n & 1) ? cr : sr, c0 = (n & 1) ? sr : cr;
    *sn = (n & 2) ? -s0 : s0;
    *cs = ((n + 1) & 2) ? -c0 : c0;
}

template <typename T> struct Num;
template <> struct Num<float> {
    static BIO_DEV float sqrt(float x) { return sqrtf(x); }
    // fp32 production build: quotient / reciprocal / square root of a positive number from the
    // special-function unit (<= 2 ulp) instead of the IEEE sequences with their slow-path calls, and in the
    // flush-to-zero forms: __fdividef / rsqrtf wrap the same MUFU instruction in a denormal-range rescaling
    // (FSETP + two predicated FMUL: 5 instructions per quotient instead of 2; ncu: 69 instructions per
    // evaluation and warp) that the O(1e-3 .. 1e4) operands of the dynamics never need; the host emulation keeps
    // plain arithmetic
#ifdef __CUDA_ARCH__
    static BIO_DEV float rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static BIO_DEV float rcp(float b) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b)); return r; }
    static BIO_DEV float div(float a, float b) { return a * rcp(b); }
    static BIO_DEV float sqrt_pos(float x) { return x * rsqrt(x); }
    // square root of a number that may be zero, special-function unit only (no slow-path call: keeps the caller a leaf)
    static BIO_DEV float sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
    static BIO_DEV float rsqrt(float x) { return 1.0f / sqrtf(x); }
    static BIO_DEV float div(float a, float b) { return a / b; }
    static BIO_DEV float rcp(float b) { return 1.0f / b; }
    static BIO_DEV float sqrt_pos(float x) { return sqrtf(x); }
    static BIO_DEV float sqrt_fast(float x) { return sqrtf(x); }
#endif
    static BIO_DEV float abs(float x) { return fabsf(x); }
    static BIO_DEV float floor(float x) { return floorf(x); }
    static BIO_DEV float exp(float x) { return expf(x); }
    static BIO_DEV void sincos(float x, float* s, float* c) { sincos_f32(x, s, c); }
    static BIO_DEV float fmod(float x, float y) { return fmodf(x, y); }
    static BIO_DEV float ceil(float x) { return ceilf(x); }
    // Newton on the normalised fibre velocity stops once a correction is below this; the iteration converges
    // quadratically, so the iterate it stops on is off by ~(correction)^2 <= 1e-7, fp32 rounding level
    static BIO_DEV float newton_tol() { return 3e-4f; }
    static constexpr int bisect_iters = 30;
};
template <> struct Num<double> {
    static BIO_DEV double sqrt(double x) { return ::sqrt(x); }
    static BIO_DEV double rsqrt(double x) { return 1.0 / ::sqrt(x); }
    static BIO_DEV double div(double a, double b) { return a / b; }
    static BIO_DEV double rcp(double b) { return 1.0 / b; }
    static BIO_DEV double sqrt_pos(double x) { return ::sqrt(x); }
    static BIO_DEV double sqrt_fast(double x) { return ::sqrt(x); }
    static BIO_DEV double abs(double x) { return ::fabs(x); }
    static BIO_DEV double floor(double x) { return ::floor(x); }
    static BIO_DEV double exp(double x) { return ::exp(x); }
    static BIO_DEV void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
    static BIO_DEV double fmod(double x, double y) { return ::fmod(x, y); }
    static BIO_DEV double ceil(double x) { return ::ceil(x); }
    static BIO_DEV double newton_tol() { return 1e-12; }
    static constexpr int bisect_iters = 60;
};

template <typename T> BIO_DEV void cross3(const T* a, const T* b, T* o) {
    T x = a[1] * b[2] - a[2] * b[1];
    T y = a[2] * b[0] - a[0] * b[2];
    T z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
template <typename T> BIO_DEV T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> BIO_DEV void matvec3(const T* R, const T* v, T* o) {
    T x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    T y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    T z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
template <typename T> BIO_DEV T clampv(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }

// 4 / 2 consecutive scalars as one shared-memory transaction (16-byte aligned arrays of the work
// buffer and of the model block); plain element access on the host
template <typename T> BIO_DEV void ld4(const T* p, T& a, T& b, T& c, T& d) {
#ifdef __CUDA_ARCH__
    if (sizeof(T) == 4) { const float4 v = *reinterpret_cast<const float4*>(p); a = v.x; b = v.y; c = v.z; d = v.w; }
    else {
        const double2 v = *reinterpret_cast<const double2*>(p), w = *reinterpret_cast<const double2*>(p + 2);
        a = v.x; b = v.y; c = w.x; d = w.y;
    }
#else
    a = p[0]; b = p[1]; c = p[2]; d = p[3];
#endif
}
template <typename T> BIO_DEV void st4(T* p, T a, T b, T c, T d) {
#ifdef __CUDA_ARCH__
    if (sizeof(T) == 4) *reinterpret_cast<float4*>(p) = make_float4((float)a, (float)b, (float)c, (float)d);
    else { *reinterpret_cast<double2*>(p) = make_double2(a, b); *reinterpret_cast<double2*>(p + 2) = make_double2(c, d); }
#else
    p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#endif
}
template <typename T> BIO_DEV void ld2(const T* p, T& a, T& b) {
#ifdef __CUDA_ARCH__
    if (sizeof(T) == 4) { const float2 v = *reinterpret_cast<const float2*>(p); a = v.x; b = v.y; }
    else { const double2 v = *reinterpret_cast<const double2*>(p); a = v.x; b = v.y; }
#else
    a = p[0]; b = p[1];
#endif
}
template <typename T> BIO_DEV void st2(T* p, T a, T b) {
#ifdef __CUDA_ARCH__
    if (sizeof(T) == 4) *reinterpret_cast<float2*>(p) = make_float2((float)a, (float)b);
    else *reinterpret_cast<double2*>(p) = make_double2(a, b);
#else
    p[0] = a; p[1] = b;
#endif
}

// index of the lowest set bit (mask != 0)
BIO_DEV int lowest_bit(int mask) {
#ifdef __CUDA_ARCH__
    return __ffs(mask) - 1;
#else
    return __builtin_ctz((unsigned)mask);
#endif
}

// OpenSim Function of one coordinate: value and first/second derivative.  hint (optional): spline
// interval of the previous evaluation of this function; the search starts there (the coordinates
// move little between substeps) and the interval found is written back.
template <typename T>
BIO_DEV void func_eval(const DevModel<T>& m, int f, T x, T& y, T& d1, T& d2, int8_t* hint = nullptr) {
    const int kind = m.func_kind[f];
    if (kind == BIO_FUNC_CONST) { y = m.func_c[f][0]; d1 = T(0); d2 = T(0); return; }
    if (kind == BIO_FUNC_LINEAR) { y = m.func_c[f][0] * x + m.func_c[f][1]; d1 = m.func_c[f][0]; d2 = T(0); return; }
    const int kb = m.func_knot_begin[f], n = m.func_knot_count[f];
    if (x <= m.knot_x[kb]) {
        d1 = m.knot_c[kb][1]; y = m.knot_c[kb][0] + d1 * (x - m.knot_x[kb]); d2 = T(0); return;
    }
    if (x >= m.knot_x[kb + n - 1]) {
        d1 = m.knot_c[kb + n - 1][1]; y = m.knot_c[kb + n - 1][0] + d1 * (x - m.knot_x[kb + n - 1]); d2 = T(0); return;
    }
    int i;
    if (hint) {
        i = *hint;
        i = i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
        while (i > 0 && x < m.knot_x[kb + i]) i--;
    } else {
        // bucketed start (host table), then at most a few forward steps
        int bk = (int)((x - m.knot_x[kb]) * m.func_bucket_inv[f]);
        bk = bk < 0 ? 0 : (bk > 15 ? 15 : bk);
        i = m.func_bucket[f][bk];
    }
    while (i + 1 < n - 1 && x >= m.knot_x[kb + i + 1]) i++;
    if (hint) *hint = (int8_t)i;
    const T dx = x - m.knot_x[kb + i];
    const T c0 = m.knot_c[kb + i][0], c1 = m.knot_c[kb + i][1], c2 = m.knot_c[kb + i][2], c3 = m.knot_c[kb + i][3];
    y = c0 + dx * (c1 + dx * (c2 + dx * c3));
    d1 = c1 + dx * (T(2) * c2 + T(3) * dx * c3);
    d2 = T(2) * c2 + T(6) * dx * c3;
}

// Tabulated Millard curve: value and slope.  Outside the table the curve continues linearly with
// the end slope; written without branches: the cubic is evaluated at the clamped abscissa and the
// overshoot is added along its slope there.
template <typename T>
BIO_DEV void curve_eval(const DevModel<T>& m, int c, T x, T& y, T& dy) {
    const T ih = m.curve_inv_h[c];
    const T t = (x - m.curve_x0[c]) * ih;
    const T tc = t < T(0) ? T(0) : (t > T(BIO_CURVE_N) ? T(BIO_CURVE_N) : t);
    int i = (int)tc;
    i = i > BIO_CURVE_N - 1 ? BIO_CURVE_N - 1 : i;
    const T s = tc - T(i);
    if constexpr (sizeof(T) == 4) {
        T c0, c1, c2, c3;                        // monomial form of the interval's cubic (DevModel::curve_q)
        ld4(m.curve_q[c][i], c0, c1, c2, c3);
        const T dyt = c1 + s * (T(2) * c2 + T(3) * s * c3);
        y = c0 + s * (c1 + s * (c2 + s * c3)) + dyt * (t - tc);
        dy = dyt * ih;
    } else {
        const T y0 = m.curve_tab[c][i][0], m0 = m.curve_tab[c][i][1];
        const T y1 = m.curve_tab[c][i + 1][0], m1 = m.curve_tab[c][i + 1][1];
        // cubic Hermite in monomial (Horner) form: same polynomial as the basis form
        const T dl = y1 - y0;
        const T c2 = T(3) * dl - T(2) * m0 - m1, c3 = m0 + m1 - T(2) * dl;
        const T dyt = m0 + s * (T(2) * c2 + T(3) * s * c3);
        y = y0 + s * (m0 + s * (c2 + s * c3)) + dyt * (t - tc);
        dy = dyt * ih;
    }
}
template <typename T>
BIO_DEV T curve_value(const DevModel<T>& m, int c, T x) { T y, d; curve_eval(m, c, x, y, d); return y; }

// Linearly implicit fibre-length update of the stated scheme (h_imp > 0, substep evaluations only): the integrator
// advances lm by h lmdot / (1 - h lambda), lambda = d(lmdot)/d(lm) = -vmax lopt g_lm / g_vn of the damped equilibrium
// g(vn, lm) = 0 at the last Newton iterate (same recipe as muscle_dynamics of the oracle, see there):
//   g_lm = (a fal' fv + fpe') cos(alpha) / lopt + fsum sin^2(alpha) / (lm cos(alpha)) + ft' / (lts cos(alpha)) >= 0
// Returns 1 / (1 - h lambda), the factor on lmdot.  With lat = lm cos(alpha) (fibre length along the tendon) and
// h2 = (lm sin(alpha))^2 (constant fibre height squared): sin^2 / (lm cos) = h2 / (lm^2 lat), 1 / cos = lm / lat.
template <typename T>
BIO_DEV T fibre_gain(const T h_imp, const T vmax_lopt, const T g_vn, const T dact_pe, const T inv_lopt, const T cosa,
                     const T fsum, const T lmc, const T lat, const T h2, const T dft, const T inv_lts) {
    const T ilat = Num<T>::rcp(lat), ilm = cosa * ilat;             // 1 / lat, 1 / lm
    T glm = dact_pe * inv_lopt * cosa + (fsum * h2 * ilm * ilm + dft * inv_lts * lmc) * ilat;
    glm = glm < T(0) ? T(0) : glm;
    return Num<T>::div(g_vn, g_vn + h_imp * vmax_lopt * glm);
}

template <typename T>
struct Kin {
    T O[3];
    T R[BIO_MAX_BODIES][9];
    T r[BIO_MAX_BODIES][3];
    T V[BIO_MAX_BODIES][6];
    T A[BIO_MAX_BODIES][6];
    T S[BIO_MAX_DOF][6];
};

template <typename T>
__device__ void kinematics(const DevModel<T>& m, const T* q, const T* u, Kin<T>& k) {
    const int nb = m.n_bodies, nd = m.n_dof;
    for (int d = 0; d < nd; d++)
        for (int c = 0; c < 6; c++) k.S[d][c] = T(0);
    for (int b = 0; b < nb; b++) {
        const int p = m.body_parent[b];
        T Rp[9], R[9], r[3], V[6], A[6];
        if (p >= 0) {
            for (int c = 0; c < 9; c++) Rp[c] = k.R[p][c];
            matvec3(Rp, m.body_joint_loc[b], r);
            for (int c = 0; c < 3; c++) r[c] += k.r[p][c];
            for (int c = 0; c < 6; c++) { V[c] = k.V[p][c]; A[c] = k.A[p][c]; }
        } else {
            Rp[0] = T(1); Rp[1] = T(0); Rp[2] = T(0); Rp[3] = T(0); Rp[4] = T(1); Rp[5] = T(0);
            Rp[6] = T(0); Rp[7] = T(0); Rp[8] = T(1);
            for (int c = 0; c < 3; c++) { r[c] = m.body_joint_loc[b][c]; V[c] = V[3 + c] = T(0); A[c] = T(0); A[3 + c] = -m.gravity[c]; }
        }
        for (int c = 0; c < 9; c++) R[c] = Rp[c];
        bool root_open = p < 0;
        const int ab = m.body_axis_begin[b], ae = ab + m.body_axis_count[b];
        for (int a = ab; a < ae; a++) {
            const int d = m.axis_dof[a];
            T s, ds, dds;
            func_eval(m, m.axis_func[a], d >= 0 ? q[d] : T(0), s, ds, dds);
            T S[6], aw[3];
            if (m.axis_kind[a] == BIO_AXIS_TRANS) {
                matvec3(Rp, m.axis_vec[a], aw);
                S[0] = S[1] = S[2] = T(0); S[3] = aw[0]; S[4] = aw[1]; S[5] = aw[2];
                for (int c = 0; c < 3; c++) r[c] += aw[c] * s;
            } else {
                if (root_open) { for (int c = 0; c < 3; c++) { k.O[c] = r[c]; r[c] = T(0); } root_open = false; }
                matvec3(R, m.axis_vec[a], aw);
                S[0] = aw[0]; S[1] = aw[1]; S[2] = aw[2];
                cross3(r, aw, S + 3);
                // R <- R * Rot(axis, s)   (Rodrigues)
                T sn, cs;
                Num<T>::sincos(s, &sn, &cs);
                const T ax = m.axis_vec[a][0], ay = m.axis_vec[a][1], az = m.axis_vec[a][2], vv = T(1) - cs;
                const T K[9] = {cs + ax * ax * vv, ax * ay * vv - az * sn, ax * az * vv + ay * sn,
                                ay * ax * vv + az * sn, cs + ay * ay * vv, ay * az * vv - ax * sn,
                                az * ax * vv - ay * sn, az * ay * vv + ax * sn, cs + az * az * vv};
                T Rn[9];
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++)
                        Rn[3 * i + j] = R[3 * i] * K[j] + R[3 * i + 1] * K[3 + j] + R[3 * i + 2] * K[6 + j];
                for (int c = 0; c < 9; c++) R[c] = Rn[c];
            }
            if (d >= 0) {
                const T qd = u[d], sd = ds * qd, acc = dds * qd * qd;
                // Sdot = V x S (spatial motion cross product)
                T c1[3], c2[3], c3[3];
                cross3(V, S, c1); cross3(V, S + 3, c2); cross3(V + 3, S, c3);
                for (int c = 0; c < 3; c++) {
                    k.S[d][c] += ds * S[c];
                    k.S[d][3 + c] += ds * S[3 + c];
                    A[c] += S[c] * acc + c1[c] * sd;
                    A[3 + c] += S[3 + c] * acc + (c2[c] + c3[c]) * sd;
                }
                for (int c = 0; c < 6; c++) V[c] += S[c] * sd;
            }
        }
        if (root_open) { for (int c = 0; c < 3; c++) { k.O[c] = r[c]; r[c] = T(0); } }
        for (int c = 0; c < 9; c++) k.R[b][c] = R[c];
        for (int c = 0; c < 3; c++) k.r[b][c] = r[c];
        for (int c = 0; c < 6; c++) { k.V[b][c] = V[c]; k.A[b][c] = A[c]; }
    }
}

template <typename T> BIO_DEV void point_pos(const Kin<T>& k, int b, const T* loc, T* x) {
    matvec3(k.R[b], loc, x);
    for (int c = 0; c < 3; c++) x[c] += k.r[b][c];
}
template <typename T> BIO_DEV void point_vel(const Kin<T>& k, int b, const T* x, T* v) {
    cross3(k.V[b], x, v);
    for (int c = 0; c < 3; c++) v[c] += k.V[b][3 + c];
}
template <typename T> BIO_DEV void add_force(T W[][6], int b, const T* x, const T* f) {
    T n[3];
    cross3(x, f, n);
    for (int c = 0; c < 3; c++) { W[b][c] += n[c]; W[b][3 + c] += f[c]; }
}

// Static fibre equilibrium (model.equilibrateMuscles, opensim_wrapper.py:290).
template <typename T>
__device__ T equilibrium_lm(const DevModel<T>& m, int i, T L, T a) {
    const T lopt = m.mus_lopt[i], lts = m.mus_lts[i], h = m.mus_height[i];
    T lo = m.mus_lm_min[i];
    const T ac = clampv(a, m.mus_amin[i], T(1));
    const T slack = L - lts;
    if (slack <= T(0)) return lo;
    T hi = Num<T>::sqrt(slack * slack + h * h);
    if (hi <= lo) return lo;
    {
        const T lat = Num<T>::sqrt(lo * lo - h * h);
        const T g = (ac * curve_value(m, 0, lo / lopt) + curve_value(m, 2, lo / lopt)) * (lat / lo) -
                    curve_value(m, 3, (L - lat) / lts);
        if (g >= T(0)) return lo;
    }
    for (int it = 0; it < Num<T>::bisect_iters; it++) {
        const T mid = T(0.5) * (lo + hi);
        const T lat = Num<T>::sqrt(mid * mid - h * h);
        const T g = (ac * curve_value(m, 0, mid / lopt) + curve_value(m, 2, mid / lopt)) * (lat / mid) -
                    curve_value(m, 3, (L - lat) / lts);
        if (g < T(0)) lo = mid; else hi = mid;
    }
    return T(0.5) * (lo + hi);
}

template <typename T> BIO_DEV T step5(T x) {
    if (x <= T(0)) return T(0);
    if (x >= T(1)) return T(1);
    return x * x * x * (T(10) + x * (T(6) * x - T(15)));
}

// Outputs of one evaluation.  The "readout" members are only filled when
// FULL (the end-of-step evaluation that feeds obs / reward / done).
template <typename T>
struct EvalOut {
    T udot[BIO_MAX_DOF];
    T adot[BIO_MAX_MUSCLES];
    T lmdot[BIO_MAX_MUSCLES];
    // readouts
    T fiber_force[BIO_MAX_MUSCLES];
    T active_fiber_force[BIO_MAX_MUSCLES];
    T contact[2][6];
    T max_limit;
    T obs_pos[BIO_MAX_OBSPTS][3];
    T obs_vel[BIO_MAX_OBSPTS][3];
    T com_pos[3], com_vel[3];
};

// Optional debug sinks (bio_eval_debug): row pointers of this env or nullptr.
template <typename T>
struct DebugRow {
    T* tendon_force; T* path_len; T* path_vel; T* limit_force; T* mass_matrix; T* bias;
};

#define BIO_MAX_MUSCLE_PTS 8

// Sparse L^T D L of the joint-space inertia H (packed lower triangle, only the tree-coupled entries
// H[i(i+1)/2 + j], j an ancestor-or-self of i, are stored and touched), in place: D on the diagonal,
// L_kq,i below it.  Elimination runs from the leaves, so no fill-in appears (Featherstone's L^T D L).
template <typename T>
__device__ void ltdl_factor(const DevModel<T>& m, T* H) {
    for (int kq = m.n_dof - 1; kq >= 0; kq--) {
        const T dk = H[kq * (kq + 1) / 2 + kq];
        for (int i = m.dof_parent[kq]; i >= 0; i = m.dof_parent[i]) {
            const T a = H[kq * (kq + 1) / 2 + i] / dk;
            for (int j = i; j >= 0; j = m.dof_parent[j]) H[i * (i + 1) / 2 + j] -= a * H[kq * (kq + 1) / 2 + j];
            H[kq * (kq + 1) / 2 + i] = a;
        }
    }
}
// x <- (L^T D L)^-1 x with the factor of ltdl_factor
template <typename T>
__device__ void ltdl_solve(const DevModel<T>& m, const T* H, T* x) {
    const int nd = m.n_dof;
    for (int i = nd - 1; i >= 0; i--)
        for (int j = m.dof_parent[i]; j >= 0; j = m.dof_parent[j]) x[j] -= H[i * (i + 1) / 2 + j] * x[i];
    for (int i = 0; i < nd; i++) x[i] /= H[i * (i + 1) / 2 + i];
    for (int i = 0; i < nd; i++) {
        T v = x[i];
        for (int j = m.dof_parent[i]; j >= 0; j = m.dof_parent[j]) v -= H[i * (i + 1) / 2 + j] * x[j];
        x[i] = v;
    }
}
// y = H x for the packed tree-sparse symmetric H (entries off the tree coupling are zero)
template <typename T>
__device__ void tree_sym_matvec(const DevModel<T>& m, const T* H, const T* x, T* y) {
    const int nd = m.n_dof;
    for (int i = 0; i < nd; i++) y[i] = H[i * (i + 1) / 2 + i] * x[i];
    for (int i = 0; i < nd; i++)
        for (int j = m.dof_parent[i]; j >= 0; j = m.dof_parent[j]) {
            const T v = H[i * (i + 1) / 2 + j];
            y[i] += v * x[j];
            y[j] += v * x[i];
        }
}

template <typename T, bool FULL>
__device__ void eval_dynamics(const DevModel<T>& m, int newton_iters, const T* q, const T* u, const T* act,
                              const T* lm, const T* ctrl, T ext_fx, int ext_pt, T h_imp, EvalOut<T>& o,
                              const DebugRow<T>* dbg) {
    Kin<T> k;
    const int nb = m.n_bodies, nd = m.n_dof, nm = m.n_muscles;
    T W[BIO_MAX_BODIES][6];
    T Q[BIO_MAX_DOF];
    for (int b = 0; b < nb; b++)
        for (int c = 0; c < 6; c++) W[b][c] = T(0);
    for (int d = 0; d < nd; d++) Q[d] = T(0);
    kinematics(m, q, u, k);

    // ---- muscles: path geometry -> Millard equilibrium -> point forces ----
    for (int i = 0; i < nm; i++) {
        T px[BIO_MAX_MUSCLE_PTS][3], pdw[BIO_MAX_MUSCLE_PTS][3];
        int pbody[BIO_MAX_MUSCLE_PTS], pdof[BIO_MAX_MUSCLE_PTS];
        int np = 0;
        const int pb = m.mus_pt_begin[i], pe = pb + m.mus_pt_count[i];
        for (int p = pb; p < pe; p++) {
            const int kind = m.pt_kind[p], d = m.pt_dof[p], b = m.pt_body[p];
            T loc[3], dloc[3] = {T(0), T(0), T(0)};
            if (kind == BIO_PT_CONDITIONAL) {
                const T v = q[d];
                if (!(v >= m.pt_range[p][0] - T(1e-5) && v <= m.pt_range[p][1] + T(1e-5))) continue;
            }
            if (kind == BIO_PT_MOVING) {
                T d2;
                for (int c = 0; c < 3; c++) func_eval(m, m.pt_func[p][c], q[d], loc[c], dloc[c], d2);
                matvec3(k.R[b], dloc, pdw[np]);
                pdof[np] = d;
            } else {
                for (int c = 0; c < 3; c++) loc[c] = m.pt_loc[p][c];
                pdof[np] = -1;
            }
            point_pos(k, b, loc, px[np]);
            pbody[np] = b;
            np++;
        }
        T L = T(0), Ld = T(0);
        T e[BIO_MAX_MUSCLE_PTS][3];
        for (int s = 0; s + 1 < np; s++) {
            const T dx = px[s + 1][0] - px[s][0], dy = px[s + 1][1] - px[s][1], dz = px[s + 1][2] - px[s][2];
            const T len = Num<T>::sqrt(dx * dx + dy * dy + dz * dz);
            const T il = T(1) / len;
            e[s][0] = dx * il; e[s][1] = dy * il; e[s][2] = dz * il;
            L += len;
        }
        if (dbg && dbg->path_vel) {
            for (int s = 0; s + 1 < np; s++) {
                T v0[3], v1[3];
                point_vel(k, pbody[s], px[s], v0);
                point_vel(k, pbody[s + 1], px[s + 1], v1);
                if (pdof[s] >= 0) for (int c = 0; c < 3; c++) v0[c] += pdw[s][c] * u[pdof[s]];
                if (pdof[s + 1] >= 0) for (int c = 0; c < 3; c++) v1[c] += pdw[s + 1][c] * u[pdof[s + 1]];
                Ld += e[s][0] * (v1[0] - v0[0]) + e[s][1] * (v1[1] - v0[1]) + e[s][2] * (v1[2] - v0[2]);
            }
            dbg->path_vel[i] = Ld;
        }
        if (dbg && dbg->path_len) dbg->path_len[i] = L;

        // Millard2012EquilibriumMuscle, elastic tendon + fibre damping
        const T fiso = m.mus_fiso[i], lopt = m.mus_lopt[i], h = m.mus_height[i], beta = m.mus_beta[i];
        const T amin = m.mus_amin[i], lmin = m.mus_lm_min[i];
        const T lmi = lm[i];
        const T lmc = lmi < lmin ? lmin : lmi;
        const T lat = Num<T>::sqrt(lmc * lmc - h * h);
        const T cosa = lat / lmc;
        T fal, fpe, ft, fv, dfv, dfal, dfpe, dft;
        curve_eval(m, 0, lmc / lopt, fal, dfal);
        curve_eval(m, 2, lmc / lopt, fpe, dfpe);
        curve_eval(m, 3, (L - lat) / m.mus_lts[i], ft, dft);
        const T ac = clampv(act[i], amin, T(1));
        const T afal = ac * fal;
        T vn = T(0), gain = T(1);
        for (int it = 0; it < newton_iters; it++) {
            curve_eval(m, 1, vn, fv, dfv);
            const T fsum = afal * fv + fpe + beta * vn;
            const T err = fsum * cosa - ft;
            const T derr = (afal * dfv + beta) * cosa;
            const T delta = -err / derr;
            if (h_imp > T(0))
                gain = fibre_gain(h_imp, m.mus_vmax[i] * lopt, derr, ac * dfal * fv + dfpe, T(1) / lopt, cosa, fsum, lmc, lat,
                                  h * h, dft, T(1) / m.mus_lts[i]);
            vn += delta;
            if (Num<T>::abs(delta) < Num<T>::newton_tol()) break;
        }
        if (lmi <= lmin && vn < T(0)) vn = T(0);
        o.lmdot[i] = vn * m.mus_vmax[i] * lopt * gain;
        const T ec = clampv(ctrl[i], amin, T(1));
        const T tau = ec > ac ? m.mus_tact[i] * (T(0.5) + T(1.5) * ac) : m.mus_tdeact[i] / (T(0.5) + T(1.5) * ac);
        o.adot[i] = (ec - ac) / tau;
        const T tension = fiso * ft;
        if (FULL) {
            curve_eval(m, 1, vn, fv, dfv);
            o.active_fiber_force[i] = fiso * afal * fv;
            o.fiber_force[i] = fiso * (afal * fv + fpe + beta * vn);
        }
        if (dbg && dbg->tendon_force) dbg->tendon_force[i] = tension;
        // equal and opposite forces along every segment
        for (int s = 0; s + 1 < np; s++) {
            const T f[3] = {tension * e[s][0], tension * e[s][1], tension * e[s][2]};
            const T fn[3] = {-f[0], -f[1], -f[2]};
            add_force(W, pbody[s], px[s], f);
            add_force(W, pbody[s + 1], px[s + 1], fn);
            if (pdof[s] >= 0) Q[pdof[s]] += dot3(f, pdw[s]);
            if (pdof[s + 1] >= 0) Q[pdof[s + 1]] += dot3(fn, pdw[s + 1]);
        }
    }
    // ---- coordinate actuators ----
    if (m.is_torque)
        for (int i = 0; i < m.n_act; i++)
            if (m.act_dof[i] >= 0) Q[m.act_dof[i]] += ctrl[i];
    // ---- Hunt-Crossley spheres on the ground half-space ----
    if (FULL)
        for (int g = 0; g < 2; g++)
            for (int c = 0; c < 6; c++) o.contact[g][c] = T(0);
    // h_imp > 0 (BIO_INT_IMPLICIT_DAMPING): diagonal velocity-Jacobian D >= 0 of every
    // active contact (friction, normal dissipation) kept for (M + h J^T D J) udot = rhs
    T cpos[BIO_MAX_SPHERES][3], cD[BIO_MAX_SPHERES][2];
    bool cact[BIO_MAX_SPHERES];
    for (int s = 0; s < m.n_spheres; s++) {
        cact[s] = false;
        const int b = m.sph_body[s];
        T xc[3];
        point_pos(k, b, m.sph_loc[s], xc);
        const T rad = m.sph_radius[s];
        const T depth = rad - (xc[1] + k.O[1]);
        if (depth <= T(0)) continue;
        const T p[3] = {xc[0], T(-0.5) * depth - k.O[1], xc[2]};
        T v[3];
        point_vel(k, b, p, v);
        const T vn = -v[1];
        const T kk = m.sph_k[s];
        const T fH = T(4.0 / 3.0) * kk * depth * Num<T>::sqrt(rad * kk * depth);
        const T f = fH * (T(1) + T(1.5) * m.sph_c[s] * vn);
        if (f <= T(0)) continue;
        T F[3] = {T(0), f, T(0)};
        const T vs = Num<T>::sqrt(v[0] * v[0] + v[2] * v[2]);
        if (vs != T(0)) {
            const T vrel = vs / m.sph_vt[s];
            const T ff = f * ((vrel < T(1) ? vrel : T(1)) * (m.sph_ud[s] + T(2) * (m.sph_us[s] - m.sph_ud[s]) / (T(1) + vrel * vrel)) +
                              m.sph_uv[s] * vs);
            F[0] = -ff * v[0] / vs;
            F[2] = -ff * v[2] / vs;
        }
        add_force(W, b, p, F);
        if (h_imp > T(0)) {
            const T vrel = vs / m.sph_vt[s];
            const T gs = (vrel < T(1) ? T(1) / m.sph_vt[s] : T(1) / vs) *
                             (m.sph_ud[s] + T(2) * (m.sph_us[s] - m.sph_ud[s]) / (T(1) + vrel * vrel)) + m.sph_uv[s];
            cact[s] = true;
            cD[s][0] = f * gs;
            cD[s][1] = T(1.5) * m.sph_c[s] * fH;
            for (int c = 0; c < 3; c++) cpos[s][c] = p[c];
        }
        if (FULL) {
            const T pa[3] = {p[0] + k.O[0], p[1] + k.O[1], p[2] + k.O[2]};
            T n[3];
            cross3(pa, F, n);
            const int g = m.sph_group[s];
            for (int c = 0; c < 3; c++) { o.contact[g][c] += F[c]; o.contact[g][3 + c] += n[c]; }
        }
    }
    // ---- coordinate limit forces ----
    T maxlim = T(0);
    T limD[BIO_MAX_LIMITS];
    for (int l = 0; l < m.n_limits; l++) {
        const int d = m.lim_dof[l];
        const T w = m.lim_w[l], qq = q[d];
        const T sup = step5((qq - m.lim_qup[l]) / w);
        const T slo = T(1) - step5((qq - (m.lim_qlo[l] - w)) / w);
        const T f = -m.lim_kup[l] * sup * (qq - m.lim_qup[l]) + m.lim_klo[l] * slo * (m.lim_qlo[l] - qq) -
                    m.lim_damp[l] * (sup + slo) * u[d];
        Q[d] += f;
        limD[l] = m.lim_damp[l] * (sup + slo);
        const T af = Num<T>::abs(f);
        maxlim = af > maxlim ? af : maxlim;
        if (dbg && dbg->limit_force) dbg->limit_force[l] = f;
    }
    if (FULL) o.max_limit = maxlim;
    // ---- perturbation force on an obs point ----
    if (ext_pt >= 0) {
        T x[3];
        const T fx[3] = {ext_fx, T(0), T(0)};
        point_pos(k, m.obs_body[ext_pt], m.obs_loc[ext_pt], x);
        add_force(W, m.obs_body[ext_pt], x, fx);
    }

    // ---- spatial inertias about O, body forces, composites ----
    T Im[BIO_MAX_BODIES], Ih[BIO_MAX_BODIES][3], II[BIO_MAX_BODIES][6];  // II: xx yy zz xy xz yz
    T F[BIO_MAX_BODIES][6];
    T mtot = T(0), msum[3] = {T(0), T(0), T(0)}, psum[3] = {T(0), T(0), T(0)};
    for (int b = 0; b < nb; b++) {
        T c[3];
        point_pos(k, b, m.body_com[b], c);
        const T* R = k.R[b];
        const T* i6 = m.body_inertia[b];
        // Iw = R * Ic * R^T (symmetric)
        T t[9];
        for (int r_ = 0; r_ < 3; r_++) {
            t[3 * r_ + 0] = R[3 * r_] * i6[0] + R[3 * r_ + 1] * i6[3] + R[3 * r_ + 2] * i6[4];
            t[3 * r_ + 1] = R[3 * r_] * i6[3] + R[3 * r_ + 1] * i6[1] + R[3 * r_ + 2] * i6[5];
            t[3 * r_ + 2] = R[3 * r_] * i6[4] + R[3 * r_ + 1] * i6[5] + R[3 * r_ + 2] * i6[2];
        }
        const T mb = m.body_mass[b], cc = dot3(c, c);
        T I6[6];
        I6[0] = t[0] * R[0] + t[1] * R[1] + t[2] * R[2] + mb * (cc - c[0] * c[0]);
        I6[1] = t[3] * R[3] + t[4] * R[4] + t[5] * R[5] + mb * (cc - c[1] * c[1]);
        I6[2] = t[6] * R[6] + t[7] * R[7] + t[8] * R[8] + mb * (cc - c[2] * c[2]);
        I6[3] = t[0] * R[3] + t[1] * R[4] + t[2] * R[5] - mb * c[0] * c[1];
        I6[4] = t[0] * R[6] + t[1] * R[7] + t[2] * R[8] - mb * c[0] * c[2];
        I6[5] = t[3] * R[6] + t[4] * R[7] + t[5] * R[8] - mb * c[1] * c[2];
        const T hh[3] = {mb * c[0], mb * c[1], mb * c[2]};
        Im[b] = mb;
        for (int j = 0; j < 3; j++) Ih[b][j] = hh[j];
        for (int j = 0; j < 6; j++) II[b][j] = I6[j];
        const T* V = k.V[b];
        const T* A = k.A[b];
        // I*V and I*A
        T IV[6], IA[6], t1[3], t2[3];
        IV[0] = I6[0] * V[0] + I6[3] * V[1] + I6[4] * V[2];
        IV[1] = I6[3] * V[0] + I6[1] * V[1] + I6[5] * V[2];
        IV[2] = I6[4] * V[0] + I6[5] * V[1] + I6[2] * V[2];
        cross3(hh, V + 3, t1); cross3(hh, V, t2);
        for (int j = 0; j < 3; j++) { IV[j] += t1[j]; IV[3 + j] = mb * V[3 + j] - t2[j]; }
        IA[0] = I6[0] * A[0] + I6[3] * A[1] + I6[4] * A[2];
        IA[1] = I6[3] * A[0] + I6[1] * A[1] + I6[5] * A[2];
        IA[2] = I6[4] * A[0] + I6[5] * A[1] + I6[2] * A[2];
        cross3(hh, A + 3, t1); cross3(hh, A, t2);
        for (int j = 0; j < 3; j++) { IA[j] += t1[j]; IA[3 + j] = mb * A[3 + j] - t2[j]; }
        T c1[3], c2[3], c3[3];
        cross3(V, IV, c1); cross3(V + 3, IV + 3, c2); cross3(V, IV + 3, c3);
        for (int j = 0; j < 3; j++) {
            F[b][j] = IA[j] + c1[j] + c2[j] - W[b][j];
            F[b][3 + j] = IA[3 + j] + c3[j] - W[b][3 + j];
        }
        if (FULL) {
            T vc[3];
            point_vel(k, b, c, vc);
            mtot += mb;
            for (int j = 0; j < 3; j++) { msum[j] += hh[j]; psum[j] += mb * vc[j]; }
        }
    }
    if (FULL) {
        const T im = T(1) / mtot;
        for (int j = 0; j < 3; j++) { o.com_pos[j] = msum[j] * im + k.O[j]; o.com_vel[j] = psum[j] * im; }
    }
    for (int b = nb - 1; b > 0; b--) {
        const int p = m.body_parent[b];
        if (p < 0) continue;
        Im[p] += Im[b];
        for (int j = 0; j < 3; j++) Ih[p][j] += Ih[b][j];
        for (int j = 0; j < 6; j++) { II[p][j] += II[b][j]; F[p][j] += F[b][j]; }
    }
    // ---- joint-space inertia (only tree-coupled entries) and bias ----
    T H[BIO_MAX_DOF * (BIO_MAX_DOF + 1) / 2];
    T rhs[BIO_MAX_DOF];
    for (int i = 0; i < nd; i++) {
        const int b = m.dof_body[i];
        const T* S = k.S[i];
        T IS[6], t1[3], t2[3];
        IS[0] = II[b][0] * S[0] + II[b][3] * S[1] + II[b][4] * S[2];
        IS[1] = II[b][3] * S[0] + II[b][1] * S[1] + II[b][5] * S[2];
        IS[2] = II[b][4] * S[0] + II[b][5] * S[1] + II[b][2] * S[2];
        cross3(Ih[b], S + 3, t1); cross3(Ih[b], S, t2);
        for (int j = 0; j < 3; j++) { IS[j] += t1[j]; IS[3 + j] = Im[b] * S[3 + j] - t2[j]; }
        for (int j = i; j >= 0; j = m.dof_parent[j]) {
            T v = T(0);
            for (int c = 0; c < 6; c++) v += k.S[j][c] * IS[c];
            H[i * (i + 1) / 2 + j] = v;
        }
        T bi = T(0);
        for (int c = 0; c < 6; c++) bi += S[c] * F[b][c];
        rhs[i] = Q[i] - bi;
        if (dbg && dbg->bias) dbg->bias[i] = bi - Q[i];
    }
    if (dbg && dbg->mass_matrix) {
        for (int i = 0; i < nd; i++)
            for (int j = 0; j < nd; j++) dbg->mass_matrix[i * nd + j] = T(0);
        for (int i = 0; i < nd; i++)
            for (int j = i; j >= 0; j = m.dof_parent[j]) {
                const T v = H[i * (i + 1) / 2 + j];
                dbg->mass_matrix[i * nd + j] = v;
                dbg->mass_matrix[j * nd + i] = v;
            }
    }
    if (h_imp > T(0)) {
        for (int l = 0; l < m.n_limits; l++) {
            const int d = m.lim_dof[l];
            H[d * (d + 1) / 2 + d] += h_imp * limD[l];
        }
        for (int s = 0; s < m.n_spheres; s++) {
            if (!cact[s]) continue;
            T col[BIO_MAX_DOF][3];
            const int last = m.body_last_dof[m.sph_body[s]];
            for (int j = last; j >= 0; j = m.dof_parent[j]) {
                cross3(k.S[j], cpos[s], col[j]);
                for (int c = 0; c < 3; c++) col[j][c] += k.S[j][3 + c];
            }
            const T dt_ = h_imp * cD[s][0], dn_ = h_imp * cD[s][1];
            for (int i = last; i >= 0; i = m.dof_parent[i])
                for (int j = i; j >= 0; j = m.dof_parent[j])
                    H[i * (i + 1) / 2 + j] += dt_ * (col[i][0] * col[j][0] + col[i][2] * col[j][2]) + dn_ * col[i][1] * col[j][1];
        }
    }
    // ---- sparse L^T D L factorisation along the tree, then solve ----
    ltdl_factor(m, H);
    ltdl_solve(m, H, rhs);
    for (int i = 0; i < nd; i++) o.udot[i] = rhs[i];
    // ---- read-outs for obs / reward / done ----
    if (FULL) {
        for (int p = 0; p < m.n_obspts; p++) {
            T x[3];
            point_pos(k, m.obs_body[p], m.obs_loc[p], x);
            point_vel(k, m.obs_body[p], x, o.obs_vel[p]);
            for (int c = 0; c < 3; c++) o.obs_pos[p][c] = x[c] + k.O[c];
        }
    }
}

// path lengths only (used by the reset-table precompute)
template <typename T>
__device__ void path_lengths(const DevModel<T>& m, const T* q, T* L) {
    Kin<T> k;
    T u0[BIO_MAX_DOF];
    for (int d = 0; d < m.n_dof; d++) u0[d] = T(0);
    kinematics(m, q, u0, k);
    for (int i = 0; i < m.n_muscles; i++) {
        T prev[3] = {T(0), T(0), T(0)};
        bool have = false;
        T len = T(0);
        const int pb = m.mus_pt_begin[i], pe = pb + m.mus_pt_count[i];
        for (int p = pb; p < pe; p++) {
            const int kind = m.pt_kind[p], d = m.pt_dof[p];
            T loc[3];
            if (kind == BIO_PT_CONDITIONAL) {
                const T v = q[d];
                if (!(v >= m.pt_range[p][0] - T(1e-5) && v <= m.pt_range[p][1] + T(1e-5))) continue;
            }
            if (kind == BIO_PT_MOVING) {
                T d1, d2;
                for (int c = 0; c < 3; c++) func_eval(m, m.pt_func[p][c], q[d], loc[c], d1, d2);
            } else {
                for (int c = 0; c < 3; c++) loc[c] = m.pt_loc[p][c];
            }
            T x[3];
            point_pos(k, m.pt_body[p], loc, x);
            if (have) {
                const T dx = x[0] - prev[0], dy = x[1] - prev[1], dz = x[2] - prev[2];
                len += Num<T>::sqrt(dx * dx + dy * dy + dz * dz);
            }
            for (int c = 0; c < 3; c++) prev[c] = x[c];
            have = true;
        }
        L[i] = len;
    }
}

}  // namespace bio
