// bio_coop.cuh -- cooperative step kernel: G lanes of a warp own one env.
//
// The thread-per-env kernel (bio_kernels.cuh) leaves a B200 almost idle at the
// batch sizes the envs are used with (4096 envs = 128 warps on 592 warp
// schedulers) and keeps its per-env working set in local memory.  Here a
// half-warp (2D models, <=16 muscles) or a full warp (3D models) steps one env:
//   lane = muscle   path geometry, Millard equilibrium, activation ODE
//   lane = body     kinematics per tree level, wrench gather, spatial inertia
//   lane = sphere / limit / dof / H-entry for contact, limits, joint-space inertia
// The per-env working set lives in shared memory (EnvWork), phases are separated
// by __syncwarp(), the model block is shared by the CTA.  Same physics and same
// operation results as eval_dynamics() in bio_dynamics.cuh; parity against the
// CPU oracle is tested in fp64 through this kernel.
#pragma once
#include "bio_kernels.cuh"

namespace bio {

template <int CLS> struct CoopCls;
template <> struct CoopCls<0> { enum { G = 16, ND = 10, NM = 16, NP = 48, NAX = 16 }; };
template <> struct CoopCls<1> { enum { G = 32, ND = 16, NM = 24, NP = 80, NAX = 24 }; };

#define COOP_MAXOBS 12

// working arrays of the general (spatial) evaluation, coop_eval.  Arrays whose lifetimes do not overlap share
// storage (a warp of the 3D kernels owns 6.1 KB, so that 28 warps fit one SM next to the model block).
template <typename T, typename C>
struct WorkGeneral {
    // pose of every body: row i of the rotation with component i of the position about O in its fourth column
    // (three 16-byte reads per use instead of twelve scalar ones: the lanes of a phase share a few bodies, so a
    // 16-byte read of 22 lanes is one or two shared-memory wavefronts where the scalar reads were twelve)
    alignas(16) T Rr[BIO_MAX_BODIES][12];
    // spatial velocity [0..5] and bias acceleration [6..11] of every body (about O, ground axes), 16-byte rows
    alignas(16) T VA[BIO_MAX_BODIES][12];
    alignas(16) T S[C::ND][8];                     // motion vector of every dof, [6..7] unused (16-byte reads)
    union {
        struct {
            // joint-space path (coop_solve): spatial inertia about O and body force per body: [0] mass, [1..3] m*c,
            // [4..9] I (xx yy zz xy xz yz), [10..15] force (n; f); turned into composite / subtree sums in place
            T BI[BIO_MAX_BODIES][16];
            union {
                T mv[8][6];                            // phases A..C: moving path points, location [0..2], d/dq [3..5]
                alignas(16) T IS[C::ND][8];            // phase G: I^c_body(i) * S_i, [6..7] unused
                T Lw[C::ND * (C::ND + 1) / 2];         // phase H: factor L of L^T D L (apart from H: no write-after-read barrier)
            };
            union {
                struct { T ax_s[C::NAX], ax_ds[C::NAX], ax_dds[C::NAX]; };   // phases A, B: joint functions and derivatives
                T H[C::ND * (C::ND + 1) / 2];                                // phases G, H: joint-space inertia
            };
            T rhs[C::ND];
        };
        // articulated-body path (p3_aba): the 6 x 6 spatial inertia of every body about O by columns, force as the
        // seventh column: [c * 6 + r]; written in phase E, when mv and the joint functions are dead
        alignas(16) T BIc[BIO_MAX_BODIES][42];
    };
    T Q[C::ND], limDd[C::ND];
    T mq[8];                                       // generalized force of the moving path points
};

// working arrays of the planar program, coop_eval_planar
template <typename T, typename C>
struct WorkPlanar {
    alignas(16) T ax[C::NAX][4];                   // displacement, ds/dq, ds/dq * qdot, d2s/dq2 * qdot^2
    alignas(16) T axr[C::NAX][2];                  // cos, sin of the signed rotation angle
    alignas(16) T pose[BIO_MAX_BODIES][4];         // cos, sin, x, y (about O, ground axes)
    alignas(16) T V[BIO_MAX_BODIES][4], A[BIO_MAX_BODIES][4];  // (w, vx, vy), bias acceleration likewise
    alignas(16) T S[C::ND][4];                     // motion vector of every dof
    // spatial inertia about O (symmetric 3 x 3 in ww wx wy xx xy yy order = I0..I5) and force per body, by columns
    // padded to four: (I0 I1 I2 -) (I1 I3 I4 -) (I2 I4 I5 -) (p0 p1 p2 -); rows of 20 words, so that the body lanes
    // of phase E store to distinct banks (16-word rows: 4-way conflicts, 24 excess wavefronts per evaluation)
    alignas(16) T bI[BIO_MAX_BODIES][20];
    T Qf[C::ND], Ld[C::ND];                        // generalized force and h * limit damping per dof
    alignas(16) T mv[P2_MAXMOV][8];                // moving points: location [0..2], d/dq [4..6]
    alignas(16) T sphI[BIO_MAX_SPHERES][8];        // h * contact damping of a sphere as an inertia about O: ww wx wy xx yy
    T mq[P2_MAXMOV];                               // their generalized force
};
// articulated-body pass of the planar program, behind the read-outs of a full evaluation (the wrench sources are dead
// by then): cooperative pass (p2_aba_coop) W = per dof U / D [0..2], -u / D [3]; one-lane-per-chain pass
// (p2_phase_f / _g, host emulation): brx = chain -> root articulated inertia [0..5] and force [6..8] of its first
// body, brk = per chain dof U / D [0..2], u / D [3]
template <typename T, typename C>
struct WorkPlanarAba {
    T pad_[128];
    alignas(16) T W[C::ND][4];
    alignas(16) T brx[P2_MAXBR][12];
    alignas(16) T brk[P2_MAXBR][12];
};

template <typename T, typename C>
struct WorkReadout { T obs_pos[COOP_MAXOBS][3], obs_vel[COOP_MAXOBS][3], comp[BIO_MAX_BODIES][6]; };   // full evaluation
static_assert(sizeof(WorkReadout<double, CoopCls<1>>) <= 128 * sizeof(double), "read-outs overlap the articulated-body arrays");
template <typename T>
struct WorkSources { alignas(16) T w[P2_MAXSRC][4]; };                                                 // planar program
#define COOP_MAXSRC6 48      // wrench sources of a 3D model (one per body a muscle touches; the reference's models: 46)
// general evaluation: moment about O and the x force [a: n0 n1 n2 f0], y / z force [b: f1 f2] -- two arrays of
// 16- and 8-byte records instead of one of 32 bytes: a quarter fewer shared-memory bytes per source on both sides
// (the 3D kernels keep the shared-memory pipe 72 % busy), and the muscle lanes, which store records 2..3 sources
// apart, spread over the banks (ncu: 44 wavefronts per evaluation for the six 16-byte stores of 22 lanes with the
// 32-byte records, 18 would do)
template <typename T>
struct WorkSources6 { alignas(16) T a[COOP_MAXSRC6][4]; alignas(16) T b[COOP_MAXSRC6][2]; };

// Size class 0 (half-warp per env) runs the planar program only, so its buffer holds no arrays of the
// general evaluation; class 1 (warp per env) keeps both.
template <typename T, int CLS> struct WorkUnions;
template <typename T> struct WorkUnions<T, 0> {
    typedef CoopCls<0> C;
    struct { WorkPlanar<T, C> p; } k;
    union { WorkReadout<T, C> out; WorkSources<T> src; WorkPlanarAba<T, C> pa; } x;
};
template <typename T> struct WorkUnions<T, 1> {
    typedef CoopCls<1> C;
    union { WorkGeneral<T, C> g; WorkPlanar<T, C> p; } k;
    union {
        struct { T col[BIO_MAX_SPHERES][C::ND][3]; } jac;             // phase G (implicit damping)
        WorkReadout<T, C> out;
        WorkSources<T> src;
        WorkSources6<T> src6;                                        // phases C..E
        // articulated-body pass, behind the read-outs of a full evaluation: per dof U / D [0..5] and -u / D [6] for the
        // way back; exchange of U between the lanes of a chain (double-buffered by step parity)
        struct { T pad_[128]; alignas(16) T W[C::ND][8]; alignas(16) T Ux[2][2][8]; } aba;
        WorkPlanarAba<T, C> pa;
    } x;
    // arguments of the running evaluation (coop_eval re-reads them where it uses them, see there):
    // flags = Newton iterations | (perturbed observation point + 1) << 8 | full evaluation << 16
    // (flags bits 24..31: substep counter of the stated integrator's loop; t0: time at the start of the control step)
    struct { T fx, himp; int32_t flags; T t0; } ev;
};

template <typename T, int CLS>
struct alignas(16) EnvWorkBody : WorkUnions<T, CLS> {
    typedef CoopCls<CLS> C;
    T q[C::ND], u[C::ND], act[C::NM], lm[C::NM];          // state of the current evaluation
    alignas(16) T O[4];
    T sphx[BIO_MAX_SPHERES][3], sphF[BIO_MAX_SPHERES][3], sphD[BIO_MAX_SPHERES][2];
    T limf[BIO_MAX_LIMITS], limD[BIO_MAX_LIMITS];
    T udot[C::ND], adot[C::NM], lmdot[C::NM];
    // fibre forces of the full evaluation (the last one of a step); during the substeps the same storage holds the
    // last change of the Newton root and the number of solves of this step (planar program, reset at step start)
    union { T ffib[C::NM]; T vnd[C::NM]; };
    union { T fact[C::NM]; T vna[C::NM]; };
    // Newton warm start: last normalised fibre velocity.  After a FULL evaluation (the last one of a step: the warm
    // start is reset before the next solve) the slot holds the tendon force, for BioStepExtra::tendon_force
    T vn[C::NM];
    T ctrl[C::NM];
    T com_pos[3], com_vel[3];
    int8_t knot_hint[P2_MAXTASK];                          // last spline interval per phase-A task
    T contact[2][6];
    T max_limit;
};

// Two envs share a warp in class 0 and touch the same fields at the same time: the buffer size is
// an odd multiple of 64 bytes, so that the second env's copy starts 16 banks away from the first's.
template <size_t PAD> struct PadBytes { unsigned char bank_pad[PAD]; };
template <> struct PadBytes<0> {};
template <typename T, int CLS> struct EnvWorkSize {
    static constexpr size_t raw = sizeof(EnvWorkBody<T, CLS>);
    static constexpr size_t want = CLS == 0 ? (((raw + 63) / 64) | 1) * 64 : ((raw + 15) / 16) * 16 + 16;
};
template <typename T, int CLS>
struct alignas(16) EnvWork : EnvWorkBody<T, CLS>, PadBytes<EnvWorkSize<T, CLS>::want - EnvWorkSize<T, CLS>::raw> {};
static_assert(sizeof(EnvWork<float, 0>) == EnvWorkSize<float, 0>::want && sizeof(EnvWork<float, 1>) == EnvWorkSize<float, 1>::want,
              "work buffer padding");

// Two envs sharing a warp (G = 16) run in lockstep: every branch that contains a barrier, shuffle or ballot is
// taken by the whole warp (an auto-reset of one env makes the other repeat its evaluation, see the step
// kernel), so all of them name the full warp with a compile-time mask -- a run-time mask costs a MATCH/REDUX/
// VOTE/branch sequence in front of every shuffle group.  Shuffles stay inside the env through their width;
// ballots are cut to the env's lanes by group_ballot.
template <int G> __device__ __forceinline__ unsigned group_mask() { return 0xffffffffu; }

// pose helpers (WorkGeneral::Rr)
template <typename T> __device__ __forceinline__ void pose_load(const T* Rr, T* R, T* r) {
    ld4(Rr, R[0], R[1], R[2], r[0]); ld4(Rr + 4, R[3], R[4], R[5], r[1]); ld4(Rr + 8, R[6], R[7], R[8], r[2]);
}
template <typename T> __device__ __forceinline__ void pose_store(T* Rr, const T* R, const T* r) {
    st4(Rr, R[0], R[1], R[2], r[0]); st4(Rr + 4, R[3], R[4], R[5], r[1]); st4(Rr + 8, R[6], R[7], R[8], r[2]);
}
// R loc + r: a point of the body in ground axes about O
template <typename T> __device__ __forceinline__ void pose_point(const T* Rr, const T* loc, T* x) {
    T R[9], r[3];
    pose_load(Rr, R, r);
    x[0] = R[0] * loc[0] + R[1] * loc[1] + R[2] * loc[2] + r[0];
    x[1] = R[3] * loc[0] + R[4] * loc[1] + R[5] * loc[2] + r[1];
    x[2] = R[6] * loc[0] + R[7] * loc[1] + R[8] * loc[2] + r[2];
}
template <int G> __device__ __forceinline__ void gsync() { __syncwarp(); }
template <int G> __device__ __forceinline__ unsigned group_ballot(bool pred) {
    const unsigned b = __ballot_sync(0xffffffffu, pred);
    return G == 32 ? b : (b >> ((threadIdx.x & 31) / G * G)) & ((1u << (G & 31)) - 1u);
}

// re-read of a shared-memory word at the point of use (never kept in a register across phases)
template <typename U> __device__ __forceinline__ U ldv(const U& x) { return *reinterpret_cast<const volatile U*>(&x); }

// column K of a row-major 3x3 (times a sign), and R <- R * Rot(e_K, angle) as a mix of the two other columns
template <typename T, int K>
__device__ __forceinline__ void axis_col(const T* R, T sg, T* aw) {
    aw[0] = sg * R[K]; aw[1] = sg * R[3 + K]; aw[2] = sg * R[6 + K];
}
template <typename T, int I1, int I2>
__device__ __forceinline__ void rot_cols(T* R, T cs, T sn) {
#pragma unroll
    for (int row = 0; row < 3; row++) {
        const T a1 = R[3 * row + I1], a2 = R[3 * row + I2];
        R[3 * row + I1] = cs * a1 + sn * a2;
        R[3 * row + I2] = cs * a2 - sn * a1;
    }
}

}  // namespace bio
#include "bio_coop_planar.cuh"
#include "bio_coop_spatial.cuh"
namespace bio {

// ---------------------------------------------------------------------------
// One evaluation of the dynamics of the env in E (state in E.q/u/act/lm,
// controls in E.ctrl).  Results: E.udot, E.adot, E.lmdot and, when full, the
// read-outs for obs / reward / done.  All G lanes of the env must call this.
// ---------------------------------------------------------------------------
template <typename T, int CLS, bool FAST>
__device__ __noinline__ void coop_eval(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    typedef CoopCls<CLS> C;
    constexpr int G = C::G;
    auto& K = E.k.g;
    // The scalars that live through the whole evaluation -- its arguments (E.ev, written by coop_eval_any), the
    // model's counts, the path switch -- stay in shared memory and are re-read where they are used.  Kept in
    // registers they are what the compiler spills at 72 registers per thread, and with 232 KB of the SM's 256 KB
    // configured as shared memory a spill reload is an L2 round trip (ncu: ~15 reload sites per evaluation, 14 % of
    // the stall samples of the 896-thread shape).  Every phase takes its own copies (one read per phase, not one per
    // use inside its loops).  No call to a libdevice slow path inside: the function is a leaf and keeps its return
    // address in a register.
#define EV_NB ldv(m.n_bodies)
#define EV_ND ldv(m.n_dof)
#define EV_NM ldv(m.n_muscles)
#define EV_ABA (G == 32 && ldv(m.prog.aba_ok) != 0)   /* root-plus-chains model: articulated-body pass instead of phases F..H */
#define EV_FULL (((ldv(E.ev.flags) >> 16) & 1) != 0)
#define EV_H_IMP ldv(E.ev.himp)
#define EV_EXT_FX ldv(E.ev.fx)
#define EV_EXT_PT (((ldv(E.ev.flags) >> 8) & 255) - 1)
#define EV_NEWTON_ITERS (ldv(E.ev.flags) & 255)

    // ---- phase A: joint functions of the coordinates, and the location functions of moving path
    // points (task n_axes + 3 k + c: component c of moving point k); the spline interval of the
    // previous evaluation is the search hint ----
    // (FAST: the instantiation for models that take the packed phase A, the kinematics scan, the three-lane phase E
    // and the articulated-body pass -- every shipped 3D model -- holds nothing else: the hot loop's text shrinks from
    // ~100 KB with the other paths' code in between to what it executes; ncu had 16 % of the instruction-cache
    // requests missing and 0.8 no-instruction stall cycles per issued instruction)
    bool a_done = false;
    if constexpr (G == 32) {
        if (FAST || ldv(m.prog.atask_ok) != 0) { p3_phase_a<T, CLS>(m, E, lane); a_done = true; }
    }
    if constexpr (!FAST)
    if (!a_done)
    for (int a = lane; a < m.n_axes + 3 * m.n_moving; a += G) {
        T s, ds, dds;
        if (a < m.n_axes) {
            const int d = m.axis_dof[a];
            func_eval(m, m.axis_func[a], d >= 0 ? E.q[d] : T(0), s, ds, dds, a < P2_MAXTASK ? &E.knot_hint[a] : nullptr);
            K.ax_s[a] = s; K.ax_ds[a] = ds; K.ax_dds[a] = dds;
        } else {
            const int k = (a - m.n_axes) / 3, c = (a - m.n_axes) % 3, p = m.moving_pt[k];
            func_eval(m, m.pt_func[p][c], E.q[m.pt_dof[p]], s, ds, dds, a < P2_MAXTASK ? &E.knot_hint[a] : nullptr);
            K.mv[k][c] = s; K.mv[k][3 + c] = ds;
        }
    }
    gsync<G>();

    // ---- phase B: kinematics; root-plus-chains models on a full warp: prefix scans over the steps of
    // every chain (bio_coop_spatial.cuh), else one tree level at a time, lane = body of the level ----
    if (FAST || (G == 32 && m.prog.chain_ok)) {
        if constexpr (G == 32) p3_phase_b_scan<T, CLS>(m, E, lane);
        gsync<G>();
    } else
    if constexpr (!FAST)
    for (int lev = 0; lev < m.n_levels; lev++) {
        const int lb = m.level_begin[lev] + lane;
        if (lb < m.level_begin[lev + 1]) {
            const int b = m.level_body[lb], p = m.body_parent[b];
            T Rp[9], R[9], r[3], V[6], A[6];
            if (p >= 0) {
                T rp[3];
                pose_load(K.Rr[p], Rp, rp);
                matvec3(Rp, m.body_joint_loc[b], r);
                for (int c = 0; c < 3; c++) r[c] += rp[c];
                for (int c = 0; c < 6; c++) { V[c] = K.VA[p][c]; A[c] = K.VA[p][6 + c]; }
            } else {
                Rp[0] = T(1); Rp[1] = T(0); Rp[2] = T(0); Rp[3] = T(0); Rp[4] = T(1); Rp[5] = T(0);
                Rp[6] = T(0); Rp[7] = T(0); Rp[8] = T(1);
                for (int c = 0; c < 3; c++) { r[c] = m.body_joint_loc[b][c]; V[c] = V[3 + c] = T(0); A[c] = T(0); A[3 + c] = -m.gravity[c]; }
            }
            for (int c = 0; c < 9; c++) R[c] = Rp[c];
            bool root_open = p < 0;
            const int ab = m.body_axis_begin[b], ae = ab + m.body_axis_count[b];
            int dprev = -1;
            T Sd[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
            for (int a = ab; a < ae; a++) {
                const int d = m.axis_dof[a], code = m.axis_code[a];
                const T s = K.ax_s[a], ds = K.ax_ds[a], dds = K.ax_dds[a];
                T S[6], aw[3];
                if (m.axis_kind[a] == BIO_AXIS_TRANS) {
                    if (code != 0) {
                        const int kx = (code > 0 ? code : -code) - 1;
                        const T sg = code > 0 ? T(1) : T(-1);
                        if (kx == 0) axis_col<T, 0>(Rp, sg, aw);
                        else if (kx == 1) axis_col<T, 1>(Rp, sg, aw);
                        else axis_col<T, 2>(Rp, sg, aw);
                    } else {
                        matvec3(Rp, m.axis_vec[a], aw);
                    }
                    S[0] = S[1] = S[2] = T(0); S[3] = aw[0]; S[4] = aw[1]; S[5] = aw[2];
                    for (int c = 0; c < 3; c++) r[c] += aw[c] * s;
                } else {
                    if (root_open) { for (int c = 0; c < 3; c++) { E.O[c] = r[c]; r[c] = T(0); } root_open = false; }
                    T sn, cs;
                    if (code != 0) {
                        // rotation about a coordinate axis of the current frame: two columns mix
                        const int kx = (code > 0 ? code : -code) - 1;
                        const T sg = code > 0 ? T(1) : T(-1);
                        Num<T>::sincos(sg * s, &sn, &cs);
                        if (kx == 0) { axis_col<T, 0>(R, sg, aw); rot_cols<T, 1, 2>(R, cs, sn); }
                        else if (kx == 1) { axis_col<T, 1>(R, sg, aw); rot_cols<T, 2, 0>(R, cs, sn); }
                        else { axis_col<T, 2>(R, sg, aw); rot_cols<T, 0, 1>(R, cs, sn); }
                    } else {
                        matvec3(R, m.axis_vec[a], aw);
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
                    S[0] = aw[0]; S[1] = aw[1]; S[2] = aw[2];
                    cross3(r, aw, S + 3);
                }
                if (d >= 0) {
                    if (d != dprev) {
                        if (dprev >= 0) for (int c = 0; c < 6; c++) K.S[dprev][c] = Sd[c];
                        for (int c = 0; c < 6; c++) Sd[c] = T(0);
                        dprev = d;
                    }
                    const T qd = E.u[d], sd = ds * qd, acc = dds * qd * qd;
                    T c1[3], c2[3], c3[3];
                    cross3(V, S, c1); cross3(V, S + 3, c2); cross3(V + 3, S, c3);
                    for (int c = 0; c < 3; c++) {
                        Sd[c] += ds * S[c];
                        Sd[3 + c] += ds * S[3 + c];
                        A[c] += S[c] * acc + c1[c] * sd;
                        A[3 + c] += S[3 + c] * acc + (c2[c] + c3[c]) * sd;
                    }
                    for (int c = 0; c < 6; c++) V[c] += S[c] * sd;
                }
            }
            if (dprev >= 0) for (int c = 0; c < 6; c++) K.S[dprev][c] = Sd[c];
            if (root_open) { for (int c = 0; c < 3; c++) { E.O[c] = r[c]; r[c] = T(0); } }
            pose_store(K.Rr[b], R, r);
            for (int c = 0; c < 6; c++) { K.VA[b][c] = V[c]; K.VA[b][6 + c] = A[c]; }
        }
        gsync<G>();
    }

    // ---- phase C: lane = muscle ----
    if (lane < EV_NM) {
        const int i = lane;
        const PlanarProg<T>& pr = m.prog;
        const T h_imp = EV_H_IMP;
        const int evf = ldv(E.ev.flags), newton_iters = evf & 255;
        const bool full = ((evf >> 16) & 1) != 0;
        T L = T(0);
        // compiled paths (PlanarProg::mc_seg; bio_create gives a 3D model without them to the thread-per-env kernel):
        // constant length of the variant selected by the conditional points plus its live segments; a live segment
        // that crosses bodies leaves +-(x cross e, e) on the two bodies
        T Wv[P2_MAXSLOT][6];
        T mqu = T(0);
        int mov = -1;
#pragma unroll
        for (int sl = 0; sl < P2_MAXSLOT; sl++)
#pragma unroll
            for (int c = 0; c < 6; c++) Wv[sl][c] = T(0);
        int var = 0;
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
            const int p = pr.mc_cond[i][cc];
            if (p >= 0) {
                const T v = E.q[(pr.pt_info[p] >> 6) & 31];
                if (v >= m.pt_range[p][0] - T(1e-5) && v <= m.pt_range[p][1] + T(1e-5)) var |= 1 << cc;
            }
        }
        L = pr.mc_len0[i][var];
        T mdw[3] = {T(0), T(0), T(0)};           // moving point: R_b d(location)/dq
        for (int j = 0; j < pr.mc_nlive; j++) {
            const uint32_t seg = pr.mc_seg[i][var][j];
            if (!(seg >> 31)) continue;
            T xe[2][3];
            int slot2[2];
            bool mv2[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int p = (seg >> (8 * e)) & 255u;
                const int info = pr.pt_info[p];
                const int b = info & 15;
                T loc[3], Rb[9], rb[3];
                pose_load(K.Rr[b], Rb, rb);
                mv2[e] = ((info >> 4) & 3) == BIO_PT_MOVING;
                if (mv2[e]) {
                    mov = (info >> 13) & 7;
                    T dl[3];
                    for (int c = 0; c < 3; c++) { loc[c] = K.mv[mov][c]; dl[c] = K.mv[mov][3 + c]; }
                    matvec3(Rb, dl, mdw);
                } else {
                    T l3;
                    ld4(pr.pt_xyz[p], loc[0], loc[1], loc[2], l3);
                }
                matvec3(Rb, loc, xe[e]);
                for (int c = 0; c < 3; c++) xe[e][c] += rb[c];
                slot2[e] = (info >> 11) & 3;
            }
            const T dx = xe[1][0] - xe[0][0], dy = xe[1][1] - xe[0][1], dz = xe[1][2] - xe[0][2];
            const T d2 = dx * dx + dy * dy + dz * dz;
            const T il = Num<T>::rsqrt(d2);
            L += d2 * il;
            const T ev[3] = {dx * il, dy * il, dz * il};
            const T msg = (mv2[0] ? T(1) : T(0)) - (mv2[1] ? T(1) : T(0));
            mqu += msg * dot3(ev, mdw);
            if (slot2[0] != slot2[1]) {
                T nn[3];
                cross3(xe[0], ev, nn);
#pragma unroll
                for (int sl = 0; sl < P2_MAXSLOT; sl++) {
                    const T sgn = sl == slot2[0] ? T(1) : (sl == slot2[1] ? T(-1) : T(0));
#pragma unroll
                    for (int c = 0; c < 3; c++) { Wv[sl][c] += sgn * nn[c]; Wv[sl][3 + c] += sgn * ev[c]; }
                }
            }
        }
        // per-muscle constants: three 16-byte reads (PlanarProg::mus_k), reciprocals from the host
        T fiso, lopt, inv_lopt, h2, beta, amin, lmin, inv_lts, vmax_lopt, inv_tact, inv_tdeact, k11;
        ld4(pr.mus_k[i], fiso, lopt, inv_lopt, h2);
        ld4(pr.mus_k[i] + 4, beta, amin, lmin, inv_lts);
        ld4(pr.mus_k[i] + 8, vmax_lopt, inv_tact, inv_tdeact, k11);
        const T lmi = E.lm[i];
        const T lmc = lmi < lmin ? lmin : lmi;
        const T lat = Num<T>::sqrt_pos(lmc * lmc - h2);
        const T cosa = Num<T>::div(lat, lmc);
        T fal, fpe, ft, fv, dfv, dfal, dfpe, dft;
        const T lnorm = lmc * inv_lopt;
        curve_eval(m, 0, lnorm, fal, dfal);
        curve_eval(m, 2, lnorm, fpe, dfpe);
        curve_eval(m, 3, (L - lat) * inv_lts, ft, dft);
        const T ac = clampv(E.act[i], amin, T(1));
        const T afal = ac * fal;
        // Newton on the damped-equilibrium residual, warm-started from the root of the previous
        // evaluation of this step.  The residual is increasing in vn, convex for vn < 0 and concave
        // for vn > 0, so Newton converges monotonically from 0 and from any point between 0 and the
        // root; an iterate that would cross 0 is put on 0 (globally convergent).
        // (start: extrapolated from the last two roots of this control step, see p2_phase_c)
        const T vlast = E.vn[i], age = E.vna[i];
        T vn = vlast + (age >= T(2) ? E.vnd[i] : T(0));
        T fsum = T(0), derr = T(1);
        for (int it = 0; it < newton_iters; it++) {
            curve_eval(m, 1, vn, fv, dfv);
            fsum = afal * fv + fpe + beta * vn;
            const T err = fsum * cosa - ft;
            derr = (afal * dfv + beta) * cosa;
            const T delta = -Num<T>::div(err, derr);
            const T vnew = vn + delta;
            const bool crossed = vnew * vn < T(0);
            vn = crossed ? T(0) : vnew;
            if (!crossed && Num<T>::abs(delta) < Num<T>::newton_tol()) break;
        }
        E.vn[i] = vn;
        E.vnd[i] = vn - vlast;
        E.vna[i] = age + T(1);
        if (lmi <= lmin && vn < T(0)) vn = T(0);
        // linearly implicit fibre-length update (fibre_gain): lmdot carries the factor 1 / (1 - h lambda) in the
        // substep evaluations (h_imp > 0); the full evaluation (h_imp = 0) reports the fibre velocity itself
        T gain = T(1);
        if (h_imp > T(0))
            gain = fibre_gain(h_imp, vmax_lopt, derr, ac * dfal * fv + dfpe, inv_lopt, cosa, fsum, lmc, lat, h2, dft, inv_lts);
        E.lmdot[i] = vn * vmax_lopt * gain;
        // activation ODE: adot = (e - a) / tau, tau = tact (0.5 + 1.5 a) rising, tdeact / (0.5 + 1.5 a) falling
        const T ec = clampv(E.ctrl[i], amin, T(1));
        const T wa = T(0.5) + T(1.5) * ac;
        E.adot[i] = (ec - ac) * (ec > ac ? inv_tact * Num<T>::rcp(wa) : inv_tdeact * wa);
        const T tension = fiso * ft;
        if (full) {
            curve_eval(m, 1, vn, fv, dfv);
            E.fact[i] = fiso * afal * fv;
            E.ffib[i] = fiso * (afal * fv + fpe + beta * vn);
            E.vn[i] = tension;                // read-out slot, see EnvWorkBody::vn
        }
        {   // wrench sources of this muscle: one per body it touches
            const int s0 = pr.mus_src0[i], ns = pr.mus_src0[i + 1] - s0;
#pragma unroll
            for (int sl = 0; sl < P2_MAXSLOT; sl++) {
                if (sl < ns) {
                    st4(E.x.src6.a[s0 + sl], tension * Wv[sl][0], tension * Wv[sl][1], tension * Wv[sl][2], tension * Wv[sl][3]);
                    st2(E.x.src6.b[s0 + sl], tension * Wv[sl][4], tension * Wv[sl][5]);
                }
            }
            if (mov >= 0) K.mq[mov] = tension * mqu;
        }
    }
    // ---- phase D: lane = contact sphere | coordinate limit ----
    if (lane < m.n_spheres) {
        const int s = lane, b = m.sph_body[s];
        // per-sphere constants: three 16-byte reads (PlanarProg::sph_k)
        T loc[3], rad, kk, c15, ud, us2, uv, vt, inv_vt, k11;
        ld4(m.prog.sph_k[s], loc[0], loc[1], loc[2], rad);
        ld4(m.prog.sph_k[s] + 4, kk, c15, ud, us2);
        ld4(m.prog.sph_k[s] + 8, uv, vt, inv_vt, k11);
        T xc[3];
        pose_point(K.Rr[b], loc, xc);
        const T depth = rad - (xc[1] + E.O[1]);
        T F[3] = {T(0), T(0), T(0)}, D0 = T(0), D1 = T(0);
        T p[3] = {xc[0], T(-0.5) * depth - E.O[1], xc[2]};
        if (depth > T(0)) {
            T v[3];
            T Vb[6];
            ld4(K.VA[b], Vb[0], Vb[1], Vb[2], Vb[3]); ld2(K.VA[b] + 4, Vb[4], Vb[5]);
            cross3(Vb, p, v);
            for (int c = 0; c < 3; c++) v[c] += Vb[3 + c];
            const T vn = -v[1];
            const T fH = T(4.0 / 3.0) * kk * depth * Num<T>::sqrt_pos(rad * kk * depth);
            const T f = fH * (T(1) + c15 * vn);
            if (f > T(0)) {
                F[1] = f;
                const T vs = Num<T>::sqrt_fast(v[0] * v[0] + v[2] * v[2]);
                const T vrel = vs * inv_vt;
                const T strib = ud + Num<T>::div(us2, T(1) + vrel * vrel);
                if (vs != T(0)) {
                    const T ff = f * ((vrel < T(1) ? vrel : T(1)) * strib + uv * vs);
                    const T fs = -Num<T>::div(ff, vs);
                    F[0] = fs * v[0];
                    F[2] = fs * v[2];
                }
                D0 = f * ((vrel < T(1) ? inv_vt : Num<T>::rcp(vs)) * strib + uv);
                D1 = c15 * fH;
            }
        }
        for (int c = 0; c < 3; c++) { E.sphx[s][c] = p[c]; E.sphF[s][c] = F[c]; }
        E.sphD[s][0] = D0; E.sphD[s][1] = D1;
    } else if (lane - m.n_spheres < m.n_limits) {
        const int l = lane - m.n_spheres, d = m.lim_dof[l];
        // per-limit constants: two 16-byte reads (PlanarProg::lim_k)
        T qup, qlo, kup, klo, damp, inv_w, w, k7;
        ld4(m.prog.lim_k[l], qup, qlo, kup, klo);
        ld4(m.prog.lim_k[l] + 4, damp, inv_w, w, k7);
        const T qq = E.q[d];
        const T sup = step5((qq - qup) * inv_w);
        const T slo = T(1) - step5((qq - (qlo - w)) * inv_w);
        E.limf[l] = -kup * sup * (qq - qup) + klo * slo * (qlo - qq) - damp * (sup + slo) * E.u[d];
        E.limD[l] = damp * (sup + slo);
    }
    gsync<G>();

    // ---- phase E.  Root-plus-chains models on a full warp (articulated-body path): three lanes per body
    // (p3_phase_e, bio_coop_spatial.cuh).  Else: lane = body (wrench gather, inertia, body force) | dof ----
    bool e_done = false;
    if constexpr (G == 32) {
        if (FAST || EV_ABA) { p3_phase_e<T, CLS, FAST>(m, E, lane); e_done = true; }
    }
    if constexpr (!FAST)
    if (!e_done) {
    // wrench of the path points on every body: on a full warp four lanes share the point list of a
    // body (the pelvis carries a third of all points), quad butterfly, then the body lane fetches the sum
    T Wn[3] = {T(0), T(0), T(0)}, Wf[3] = {T(0), T(0), T(0)};
    if (EV_NM > 0) {
        constexpr int PARTS = G == 32 ? 4 : 1;
        const int gb = lane / PARTS, part = lane % PARTS;
        if (gb < EV_NB) {
            for (int k = m.prog.inc_begin[gb] + part; k < m.prog.inc_begin[gb + 1]; k += PARTS) {
                T w0, w1, w2, w3, w4, w5;
                const int e = m.prog.inc_src[k];
                ld4(E.x.src6.a[e], w0, w1, w2, w3);
                ld2(E.x.src6.b[e], w4, w5);
                Wn[0] += w0; Wn[1] += w1; Wn[2] += w2; Wf[0] += w3; Wf[1] += w4; Wf[2] += w5;
            }
        }
        if (PARTS > 1) {
            const unsigned mask = group_mask<G>();
#pragma unroll
            for (int c = 0; c < 3; c++) {
                Wn[c] += __shfl_xor_sync(mask, Wn[c], 1, G); Wn[c] += __shfl_xor_sync(mask, Wn[c], 2, G);
                Wf[c] += __shfl_xor_sync(mask, Wf[c], 1, G); Wf[c] += __shfl_xor_sync(mask, Wf[c], 2, G);
            }
#pragma unroll
            for (int c = 0; c < 3; c++) {
                Wn[c] = __shfl_sync(mask, Wn[c], (lane * PARTS) & (G - 1), G);
                Wf[c] = __shfl_sync(mask, Wf[c], (lane * PARTS) & (G - 1), G);
            }
        }
    }
    if (lane < EV_NB) {
        const int b = lane;
        for (int smask = m.prog.body_sph_mask[b]; smask; smask &= smask - 1) {
            const int s = lowest_bit(smask);
            if (E.sphF[s][1] == T(0)) continue;
            T n[3];
            cross3(E.sphx[s], E.sphF[s], n);
            for (int c = 0; c < 3; c++) { Wn[c] += n[c]; Wf[c] += E.sphF[s][c]; }
        }
        const int ext_pt = EV_EXT_PT;
        if (ext_pt >= 0 && m.obs_body[ext_pt] == b) {
            T x[3], n[3];
            const T fx[3] = {EV_EXT_FX, T(0), T(0)};
            pose_point(K.Rr[b], m.obs_loc[ext_pt], x);
            cross3(x, fx, n);
            for (int c = 0; c < 3; c++) { Wn[c] += n[c]; Wf[c] += fx[c]; }
        }
        T cpos[3], R[9], rb[3];
        pose_load(K.Rr[b], R, rb);
        matvec3(R, m.body_com[b], cpos);
        for (int c = 0; c < 3; c++) cpos[c] += rb[c];
        const T* i6 = m.body_inertia[b];
        T t[9];
        for (int r_ = 0; r_ < 3; r_++) {
            t[3 * r_ + 0] = R[3 * r_] * i6[0] + R[3 * r_ + 1] * i6[3] + R[3 * r_ + 2] * i6[4];
            t[3 * r_ + 1] = R[3 * r_] * i6[3] + R[3 * r_ + 1] * i6[1] + R[3 * r_ + 2] * i6[5];
            t[3 * r_ + 2] = R[3 * r_] * i6[4] + R[3 * r_ + 1] * i6[5] + R[3 * r_ + 2] * i6[2];
        }
        const T mb = m.body_mass[b], cc = dot3(cpos, cpos);
        T I6[6];
        I6[0] = t[0] * R[0] + t[1] * R[1] + t[2] * R[2] + mb * (cc - cpos[0] * cpos[0]);
        I6[1] = t[3] * R[3] + t[4] * R[4] + t[5] * R[5] + mb * (cc - cpos[1] * cpos[1]);
        I6[2] = t[6] * R[6] + t[7] * R[7] + t[8] * R[8] + mb * (cc - cpos[2] * cpos[2]);
        I6[3] = t[0] * R[3] + t[1] * R[4] + t[2] * R[5] - mb * cpos[0] * cpos[1];
        I6[4] = t[0] * R[6] + t[1] * R[7] + t[2] * R[8] - mb * cpos[0] * cpos[2];
        I6[5] = t[3] * R[6] + t[4] * R[7] + t[5] * R[8] - mb * cpos[1] * cpos[2];
        const T hh[3] = {mb * cpos[0], mb * cpos[1], mb * cpos[2]};
        T V[6], A[6];
        for (int c = 0; c < 6; c++) { V[c] = K.VA[b][c]; A[c] = K.VA[b][6 + c]; }
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
        K.BI[b][0] = mb;
        for (int j = 0; j < 3; j++) {
            K.BI[b][1 + j] = hh[j];
            K.BI[b][10 + j] = IA[j] + c1[j] + c2[j] - Wn[j];
            K.BI[b][13 + j] = IA[3 + j] + c3[j] - Wf[j];
        }
        for (int j = 0; j < 6; j++) K.BI[b][4 + j] = I6[j];
    } else if (lane - EV_NB < EV_ND) {
        const int d = lane - EV_NB;
        const T h_imp = EV_H_IMP;
        T qf = T(0), ld = T(0);
        if (m.gdof_ok) {                         // host lists: the (<= 2) limits / moving points and the actuator of the dof
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int l = m.gdof_lim[d][j], p = m.gdof_movpt[d][j];
                if (l >= 0) { qf += E.limf[l]; ld += E.limD[l]; }
                if (p >= 0) qf += K.mq[m.pt_mov[p]];
            }
            const int a = m.gdof_act[d];
            if (a >= 0) qf += E.ctrl[a];
        } else {
            for (int l = 0; l < m.n_limits; l++) if (m.lim_dof[l] == d) { qf += E.limf[l]; ld += E.limD[l]; }
            for (int k = 0; k < m.n_moving; k++) { const int p = m.moving_pt[k]; if (m.pt_dof[p] == d) qf += K.mq[k]; }
            if (m.is_torque) for (int a = 0; a < m.n_act; a++) if (m.act_dof[a] == d) qf += E.ctrl[a];
        }
        K.limDd[d] = h_imp * ld; K.Q[d] = qf;
    }
    }
    gsync<G>();

    // ---- full evaluation read-outs (pt region is dead now) ----
    if (EV_FULL) {
        const int nb = EV_NB;
        if (lane < nb) {
            const int b = lane;
            T cpos[3], vc[3];
            pose_point(K.Rr[b], m.body_com[b], cpos);
            cross3(K.VA[b], cpos, vc);
            for (int c = 0; c < 3; c++) { E.x.out.comp[b][c] = m.body_mass[b] * cpos[c]; E.x.out.comp[b][3 + c] = m.body_mass[b] * (vc[c] + K.VA[b][3 + c]); }
        } else if (lane - nb < m.n_obspts) {
            const int p = lane - nb, b = m.obs_body[p];
            T x[3], v[3];
            pose_point(K.Rr[b], m.obs_loc[p], x);
            cross3(K.VA[b], x, v);
            for (int c = 0; c < 3; c++) { E.x.out.obs_pos[p][c] = x[c] + E.O[c]; E.x.out.obs_vel[p][c] = v[c] + K.VA[b][3 + c]; }
        }
        gsync<G>();
        if (lane < 3) {
            T ms = T(0), ps = T(0);
            for (int b = 0; b < nb; b++) { ms += E.x.out.comp[b][lane]; ps += E.x.out.comp[b][3 + lane]; }
            const T im = Num<T>::rcp(m.total_mass);
            E.com_pos[lane] = ms * im + E.O[lane];
            E.com_vel[lane] = ps * im;
        } else if (lane < 5) {
            const int g = lane - 3;
            T w[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
            for (int s = 0; s < m.n_spheres; s++) {
                if (m.sph_group[s] != g) continue;
                const T pa[3] = {E.sphx[s][0] + E.O[0], E.sphx[s][1] + E.O[1], E.sphx[s][2] + E.O[2]};
                T n[3];
                cross3(pa, E.sphF[s], n);
                for (int c = 0; c < 3; c++) { w[c] += E.sphF[s][c]; w[3 + c] += n[c]; }
            }
            for (int c = 0; c < 6; c++) E.contact[g][c] = w[c];
        } else if (lane == 5) {
            T mx = T(0);
            for (int l = 0; l < m.n_limits; l++) { const T a = Num<T>::abs(E.limf[l]); mx = a > mx ? a : mx; }
            E.max_limit = mx;
        }
    }

    // ---- root-plus-chains models on a full warp: articulated-body pass (bio_coop_spatial.cuh) ----
    if constexpr (G == 32) {
        if (FAST || EV_ABA) { p3_aba<T, CLS>(m, E, lane, EV_H_IMP); return; }
    }
    if constexpr (!FAST) {

    // ---- phase F: composite inertias / subtree forces.  Root-plus-chains models on a full warp: lane =
    // (chain, value) runs the suffix sum down its chain, then 16 lanes add the chain heads to the root;
    // else task = (body of the level, value) with a barrier per tree level ----
    if (G == 32 && m.prog.chain_ok && m.prog.n_branches >= 1) {
        const int cch = lane >> 4, v = lane & 15;
        if (cch < m.prog.n_branches) {
            T acc = T(0);
            for (int k = m.prog.gch_nb[cch] - 1; k >= 0; k--) {
                const int b = m.prog.gch_body[cch][k];
                acc += K.BI[b][v];
                K.BI[b][v] = acc;
            }
        }
        gsync<G>();
        if (lane < 16) {
            T acc = K.BI[m.prog.root_body][lane];
            for (int l = 0; l < m.prog.n_branches; l++) acc += K.BI[m.prog.gch_body[l][0]][lane];
            K.BI[m.prog.root_body][lane] = acc;
        }
        gsync<G>();
    } else
    for (int lev = m.n_levels - 2; lev >= 0; lev--) {
        const int cnt = (m.level_begin[lev + 1] - m.level_begin[lev]) * 16;
        for (int tsk = lane; tsk < cnt; tsk += G) {
            const int b = m.level_body[m.level_begin[lev] + (tsk >> 4)], v = tsk & 15;
            T acc = K.BI[b][v];
            for (int k = m.child_begin[b]; k < m.child_begin[b + 1]; k++) acc += K.BI[m.child_list[k]][v];
            K.BI[b][v] = acc;
        }
        gsync<G>();
    }

    // ---- phase G: I^c S per dof, contact Jacobian columns, then one lane per coupled (i,j) entry ----
    const int nd = EV_ND;
    const T h_imp = EV_H_IMP;
    if (lane < nd) {
        const int i = lane, b = m.dof_body[i];
        const T* S = K.S[i];
        const T* B = K.BI[b];
        T IS[6], t1[3], t2[3];
        IS[0] = B[4] * S[0] + B[7] * S[1] + B[8] * S[2];
        IS[1] = B[7] * S[0] + B[5] * S[1] + B[9] * S[2];
        IS[2] = B[8] * S[0] + B[9] * S[1] + B[6] * S[2];
        cross3(B + 1, S + 3, t1); cross3(B + 1, S, t2);
        for (int c = 0; c < 3; c++) { K.IS[i][c] = IS[c] + t1[c]; K.IS[i][3 + c] = B[0] * S[3 + c] - t2[c]; }
        T bi = T(0);
        for (int c = 0; c < 6; c++) bi += S[c] * B[10 + c];
        K.rhs[i] = K.Q[i] - bi;
    }
    unsigned act_mask = 0u;      // active contacts of this env (same value on every lane)
    if (h_imp > T(0)) {
        act_mask = group_ballot<G>(lane < m.n_spheres && E.sphD[lane < m.n_spheres ? lane : 0][1] > T(0));
        // contact Jacobian columns col[s][d] = w_d x p_s + v_d for the dofs on the sphere's chain
        if (act_mask)
            for (int tsk = lane; tsk < m.jc_n; tsk += G) {      // (sphere, dof on its chain)
                const int s = m.jc_s[tsk], d = m.jc_d[tsk];
                if (!((act_mask >> s) & 1u)) continue;
                T cv[3];
                cross3(K.S[d], E.sphx[s], cv);
                for (int c = 0; c < 3; c++) E.x.jac.col[s][d][c] = cv[c] + K.S[d][3 + c];
            }
    }
    gsync<G>();
    for (int e = lane; e < m.n_entries; e += G) {
        const int i = m.ent_i[e], j = m.ent_j[e];
        T sj[8], is[8];
        ld4(K.S[j], sj[0], sj[1], sj[2], sj[3]); ld4(K.S[j] + 4, sj[4], sj[5], sj[6], sj[7]);
        ld4(K.IS[i], is[0], is[1], is[2], is[3]); ld4(K.IS[i] + 4, is[4], is[5], is[6], is[7]);
        T v = T(0);
#pragma unroll
        for (int c = 0; c < 6; c++) v += sj[c] * is[c];
        if (h_imp > T(0)) {
            unsigned mm = act_mask & m.ent_sph[e];               // active spheres whose chain holds i (and j)
            while (mm) {
                const int s = __ffs(mm) - 1;
                mm &= mm - 1u;
                const T* ci = E.x.jac.col[s][i];
                const T* cj = E.x.jac.col[s][j];
                v += h_imp * (E.sphD[s][0] * (ci[0] * cj[0] + ci[2] * cj[2]) + E.sphD[s][1] * ci[1] * cj[1]);
            }
            if (i == j) v += K.limDd[i];
        }
        K.H[i * (i + 1) / 2 + j] = v;
    }
    gsync<G>();

    // ---- phase H: sparse L^T D L along the tree + solve (bio_coop_planar.cuh) ----
    coop_solve<T, CLS>(m, E, lane);
    }   // !FAST
}
#undef EV_NB
#undef EV_ND
#undef EV_NM
#undef EV_ABA
#undef EV_FULL
#undef EV_H_IMP
#undef EV_EXT_FX
#undef EV_EXT_PT
#undef EV_NEWTON_ITERS

// planar (2D) models take the planar program; size class 0 holds nothing else (bio_create puts a model
// there only when the host could build the program)
template <typename T, int CLS>
__device__ __forceinline__ void coop_eval_any(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane,
                                              const int newton_iters, const T ext_fx, const int ext_pt, const T h_imp,
                                              const bool full, const int sub_next = 0) {
    if constexpr (CLS == 0) {
        if ((m.prog.coop_aba & 2) != 0) coop_eval_planar<T, CLS, true>(m, E, lane, newton_iters, ext_fx, ext_pt, h_imp, full);
        else coop_eval_planar<T, CLS, false>(m, E, lane, newton_iters, ext_fx, ext_pt, h_imp, full);
    } else {
        // every lane stores the same words, so no barrier is needed before a lane reads them back
        E.ev.fx = ext_fx; E.ev.himp = h_imp;
        E.ev.flags = (newton_iters & 255) | ((ext_pt + 1) << 8) | ((full ? 1 : 0) << 16) | (sub_next << 24);
        if (m.prog.ok) coop_eval_planar<T, CLS>(m, E, lane, newton_iters, ext_fx, ext_pt, h_imp, full);
        else if ((ldv(m.prog.aba_ok) & 2) != 0) coop_eval<T, CLS, true>(m, E, lane);     // see coop_eval: FAST
        else coop_eval<T, CLS, false>(m, E, lane);
    }
}

// The warp's work buffer and the model block from the thread index alone (read through a volatile asm, so that the
// compiler does not merge them with the pointers the step kernel holds): what the substep loop of the 3D kernels
// uses in place of values that would otherwise be spilled around the call of the evaluation and reloaded from L2.
template <typename T, int CLS>
__device__ __forceinline__ EnvWork<T, CLS>& coop_fresh_work(unsigned& tid) {
    extern __shared__ __align__(16) unsigned char smem[];
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    return reinterpret_cast<EnvWork<T, CLS>*>(smem + ((sizeof(DevModel<T>) + 15) / 16) * 16)[tid / CoopCls<CLS>::G];
}
template <typename T>
__device__ __forceinline__ const DevModel<T>& coop_fresh_model() {
    extern __shared__ __align__(16) unsigned char smem[];
    return *reinterpret_cast<const DevModel<T>*>(smem);
}

// state <-> work buffer helpers (lane d < nd owns a dof, lane k < nm owns a muscle)
template <typename T, int CLS>
__device__ __forceinline__ void coop_clamp(const DevModel<T>& m, EnvWork<T, CLS>& E, int lane) {
    if (lane < m.n_muscles) {
        E.act[lane] = clampv(E.act[lane], m.mus_amin[lane], T(1));
        if (E.lm[lane] < m.mus_lm_min[lane]) E.lm[lane] = m.mus_lm_min[lane];
    }
}

template <typename T, int CLS>
__device__ void coop_integrate(const DevModel<T>& m, const DevTask<T>& c, EnvWork<T, CLS>& E, int lane, int istep,
                               unsigned long long seed, unsigned long long env) {
    constexpr int G = CoopCls<CLS>::G;
    const int nd = m.n_dof, nm = m.n_muscles;
    if constexpr (CLS == 1 && sizeof(T) == 4) {
        if (c.integrator == BIO_INT_SEMI_IMPLICIT_EULER || c.integrator == BIO_INT_IMPLICIT_DAMPING) {
            // Substep loop of the stated scheme for the 3D kernels (72 registers per thread at 896 threads): nothing
            // of the loop lives in a register across the call of the evaluation -- the counter and the start time
            // sit in E.ev, the pointers are re-derived from the thread index, the rest comes from the constant bank
            // -- so nothing is spilled and reloaded per substep (a reload is an L2 round trip here, see coop_eval).
            E.ev.t0 = T(istep) * c.dt;
            E.ev.flags = 0;
            for (;;) {
                unsigned tid;
                EnvWork<T, CLS>& W = coop_fresh_work<T, CLS>(tid);
                const int sub = (int)((unsigned)ldv(W.ev.flags) >> 24);
                if (sub >= c.n_substeps) break;
                const T fx = c.perturb ? perturb_force(c, seed, env, ldv(W.ev.t0) + T(sub) * c.h_sub) : T(0);
                coop_eval_any<T, CLS>(coop_fresh_model<T>(), W, (int)(tid % G), c.newton_iters, fx,
                                      c.perturb ? c.perturb_obspt : -1,
                                      c.integrator == BIO_INT_IMPLICIT_DAMPING ? c.h_sub : T(0), false, sub + 1);
                unsigned tid2;
                EnvWork<T, CLS>& V = coop_fresh_work<T, CLS>(tid2);
                const DevModel<T>& mm = coop_fresh_model<T>();
                const int ln = (int)(tid2 % G);
                const T hs = c.h_sub;
                if (ln < ldv(mm.n_dof)) { const T un = V.u[ln] + hs * V.udot[ln]; V.u[ln] = un; V.q[ln] += hs * un; }
                if (ln < ldv(mm.n_muscles)) {          // explicit Euler with the clamps of coop_clamp
                    const T lmn = V.lm[ln] + hs * V.lmdot[ln], lmin = mm.mus_lm_min[ln];
                    V.act[ln] = clampv(V.act[ln] + hs * V.adot[ln], mm.mus_amin[ln], T(1));
                    V.lm[ln] = lmn < lmin ? lmin : lmn;
                }
                gsync<G>();
            }
            return;
        }
    }
    const T h = c.h_sub;
    const T t0 = T(istep) * c.dt;
    const int ext_pt = c.perturb ? c.perturb_obspt : -1;
    const bool isd = lane < nd, ism = lane < nm;
    for (int sub = 0; sub < c.n_substeps; sub++) {
        const T t = t0 + T(sub) * h;
        if (c.integrator == BIO_INT_SEMI_IMPLICIT_EULER || c.integrator == BIO_INT_IMPLICIT_DAMPING) {
            coop_eval_any<T, CLS>(m, E, lane, c.newton_iters, perturb_force(c, seed, env, t), ext_pt,
                              c.integrator == BIO_INT_IMPLICIT_DAMPING ? h : T(0), false);
            if (isd) { const T un = E.u[lane] + h * E.udot[lane]; E.u[lane] = un; E.q[lane] += h * un; }
            if (ism) {                           // explicit Euler with the clamps of coop_clamp
                const T lmn = E.lm[lane] + h * E.lmdot[lane], lmin = m.mus_lm_min[lane];
                E.act[lane] = clampv(E.act[lane] + h * E.adot[lane], m.mus_amin[lane], T(1));
                E.lm[lane] = lmn < lmin ? lmin : lmn;
            }
            gsync<G>();
            continue;
        } else if (c.integrator == BIO_INT_RK2_MIDPOINT) {
            const T q0 = isd ? E.q[lane] : T(0), u0 = isd ? E.u[lane] : T(0);
            const T a0 = ism ? E.act[lane] : T(0), l0 = ism ? E.lm[lane] : T(0);
            coop_eval_any<T, CLS>(m, E, lane, c.newton_iters, perturb_force(c, seed, env, t), ext_pt, T(0), false);
            const T hh = T(0.5) * h;
            if (isd) { E.q[lane] = q0 + hh * u0; E.u[lane] = u0 + hh * E.udot[lane]; }
            if (ism) { E.act[lane] = a0 + hh * E.adot[lane]; E.lm[lane] = l0 + hh * E.lmdot[lane]; }
            coop_clamp(m, E, lane);
            gsync<G>();
            coop_eval_any<T, CLS>(m, E, lane, c.newton_iters, perturb_force(c, seed, env, t + hh), ext_pt, T(0), false);
            if (isd) { const T um = E.u[lane]; E.q[lane] = q0 + h * um; E.u[lane] = u0 + h * E.udot[lane]; }
            if (ism) { E.act[lane] = a0 + h * E.adot[lane]; E.lm[lane] = l0 + h * E.lmdot[lane]; }
        } else {  // classic RK4
            const T q0 = isd ? E.q[lane] : T(0), u0 = isd ? E.u[lane] : T(0);
            const T a0 = ism ? E.act[lane] : T(0), l0 = ism ? E.lm[lane] : T(0);
            T aq = T(0), au = T(0), aa = T(0), al = T(0);
            for (int r = 0; r < 4; r++) {
                const T wgt = (r == 0 || r == 3) ? T(1) : T(2);
                const T cn = r == 2 ? T(1) : T(0.5);
                coop_eval_any<T, CLS>(m, E, lane, c.newton_iters,
                                  perturb_force(c, seed, env, t + (r == 0 ? T(0) : (r == 3 ? h : T(0.5) * h))), ext_pt,
                                  T(0), false);
                if (isd) { aq += wgt * E.u[lane]; au += wgt * E.udot[lane]; }
                if (ism) { aa += wgt * E.adot[lane]; al += wgt * E.lmdot[lane]; }
                if (r < 3) {
                    if (isd) { const T us = E.u[lane]; E.q[lane] = q0 + cn * h * us; E.u[lane] = u0 + cn * h * E.udot[lane]; }
                    if (ism) { E.act[lane] = a0 + cn * h * E.adot[lane]; E.lm[lane] = l0 + cn * h * E.lmdot[lane]; }
                    coop_clamp(m, E, lane);
                    gsync<G>();
                }
            }
            const T h6 = h / T(6);
            if (isd) { E.q[lane] = q0 + h6 * aq; E.u[lane] = u0 + h6 * au; }
            if (ism) { E.act[lane] = a0 + h6 * aa; E.lm[lane] = l0 + h6 * al; }
        }
        coop_clamp(m, E, lane);
        gsync<G>();
    }
}

// sum over the G lanes of the env (butterfly, same order on every lane)
template <typename T, int G>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(group_mask<G>(), v, o, G);
    return v;
}
template <typename T, int G>
__device__ __forceinline__ T group_max(T v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) { const T w = __shfl_xor_sync(group_mask<G>(), v, o, G); v = w > v ? w : v; }
    return v;
}

// Observation row of the env from the host-built slot descriptors (build_obs_desc): the source array is
// picked with selects and read with one indexed load, so the lanes of a group (every one on a different kind of
// slot) do not serialise over a switch; reference rows come from global memory under a predicate.
template <typename T, int CLS>
__device__ void coop_write_obs(const DevModel<T>& m, const DevTask<T>& c, const EnvWork<T, CLS>& E, int lane, int istep,
                               T* orow) {
    constexpr int G = CoopCls<CLS>::G;
    T pel[3];
#pragma unroll
    for (int k = 0; k < 3; k++) pel[k] = m.pel_dof[k] >= 0 ? E.q[m.pel_dof[k]] : T(0);
    const int row_next = ref_row(c, istep + 1);
    const T ph = T(istep) / T(c.cycle);
    const T phase = ph - Num<T>::floor(ph);
    const T* pos = &E.x.out.obs_pos[0][0];
    const T* vel = &E.x.out.obs_vel[0][0];
    const T* con = &E.contact[0][0];
    for (int o = lane; o < c.obs_dim; o += G) {
        const int desc = m.obs_desc[o], kind = desc >> 16, psel = (desc >> 12) & 3, idx = desc & 0xfff;
        const T cst = m.obs_cst[o];
        const T* base = kind == 1 ? E.q : kind == 2 ? E.u : kind == 3 ? E.udot : kind == 6 ? pos : kind == 7 ? E.com_pos
                      : kind == 8 ? vel : kind == 9 ? E.com_vel : kind == 10 ? E.act : kind == 11 ? E.lm
                      : kind == 12 ? E.lmdot : kind == 13 ? con : E.q;
        const bool in_work = kind != 0 && kind != 4 && kind != 5 && kind != 14;
        T v = base[in_work ? idx : 0];
        v -= psel == 1 ? pel[0] : psel == 2 ? pel[1] : psel == 3 ? pel[2] : T(0);
        if (kind == 13) v *= cst;
        if (kind == 4) v = c.ref_q[(size_t)row_next * c.ref_coords + idx];
        if (kind == 5) v = c.ref_u[(size_t)row_next * c.ref_coords + idx];
        if (kind == 0) v = phase;
        if (kind == 14) v = cst;
        orow[o] = v;
    }
}

// Size class 0: the observation rows of the warp's two envs (adjacent in memory) are staged in the planar
// work arrays of their envs, idle after the evaluation; the whole warp writes them, 16 bytes per lane, so that
// a caller's page-locked host buffer receives full lines.  rows: bit e set = row of env e of the warp.
template <typename T, int CLS>
__device__ __forceinline__ void coop_flush_obs(EnvWork<T, CLS>* works, T* __restrict__ obs, int od, int i0, int n,
                                               unsigned rows) {
    constexpr int EPW = 32 / CoopCls<CLS>::G;
    __syncwarp();                              // the staged rows are complete
    const int wl = threadIdx.x & 31;
    if (i0 + 1 >= n) rows &= 1u;
    const T* s0 = reinterpret_cast<const T*>(&works[(threadIdx.x >> 5) * EPW].k.p);
    const T* s1 = reinterpret_cast<const T*>(&works[(threadIdx.x >> 5) * EPW + (EPW > 1 ? 1 : 0)].k.p);
    T* base = obs + (size_t)i0 * od;
    const int lo = (rows & 1u) ? 0 : od, hi = (rows & 2u) ? 2 * od : od;     // floats [lo, hi) of the two-row block
    if (sizeof(T) == 4 && (reinterpret_cast<size_t>(base) & 15) == 0) {
        for (int v = wl + (lo >> 2); 4 * v < hi; v += 32) {
            T x[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { const int k = 4 * v + j; x[j] = k < od ? s0[k] : (k < 2 * od ? s1[k - od] : T(0)); }
            if (4 * v >= lo && 4 * v + 3 < hi) st4(base + 4 * v, x[0], x[1], x[2], x[3]);
            else for (int j = 0; j < 4; j++) { const int k = 4 * v + j; if (k >= lo && k < hi) base[k] = x[j]; }
        }
    } else {
        for (int k = lo + wl; k < hi; k += 32) base[k] = k < od ? s0[k] : s1[k - od];
    }
    __syncwarp();
}

// Persistent launch: one CTA per SM, the model block is staged once per SM, and every warp walks the work
// items (one item = the 32/G envs of a warp) item = blockIdx + gridDim * (warp + warps_per_cta * k), so
// that the items are dealt round-robin over the SMs first.  Two launch shapes per fp32 instantiation:
//   LO  512 threads = 16 warps at 128 registers: least latency per item.  4096 2D envs on 148 SMs: warps 0..13
//       of every SM are busy, i.e. 4,4,3,3 warps on the four schedulers (two CTAs of 8 warps left 4,4,4,2
//       and staged the model twice; measured 5 % slower).
//   HI  more items in flight for batches of several items per warp:
//       2D (size class 0): 640 threads = 20 warps at 96 registers (measured: 16384 envs +18 %, 131072 envs +7 %,
//       4096 envs -3 %); 3D (size class 1): 896 threads = 28 warps at 72 registers -- the 3D work buffer is 6.1 KB
//       per warp, so 28 of them fit next to the model block; 8192 envs are 55.4 items per SM: two rounds of 28
//       warps instead of three rounds of 20.
// bio_create picks by the number of item rounds per SM (COOP_SHAPE_COST); fp64 runs 256 threads.
#ifndef BIO_COOP_LO
#define BIO_COOP_LO 512
#endif
#ifndef BIO_COOP_HI3D
#define BIO_COOP_HI3D 896
#endif
#ifndef BIO_COOP_HI2D
#define BIO_COOP_HI2D 640
#endif
#define COOP_THREADS_LO(T) (sizeof(T) == 4 ? BIO_COOP_LO : 256)
#define COOP_THREADS_HI(T, CLS) (sizeof(T) == 4 ? ((CLS) == 1 ? BIO_COOP_HI3D : BIO_COOP_HI2D) : 256)
// time of one round of items with the HI shape relative to the LO shape
#define COOP_SHAPE_COST(CLS) ((CLS) == 1 ? 1.45 : 1.12)

template <typename T, int CLS, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
bio_coop_step_kernel(const DevModel<T>* __restrict__ gm, const DevTask<T> c, const EnvState<T> st, int n,
                     unsigned long long seed, long long env_offset, const T* __restrict__ actions,
                     T* __restrict__ obs, T* __restrict__ reward, uint8_t* __restrict__ done, T* __restrict__ terms,
                     double* __restrict__ stats) {
    typedef CoopCls<CLS> C;
    constexpr int G = C::G;
    constexpr int EPW = 32 / G;                 // envs per warp
    extern __shared__ __align__(16) unsigned char smem[];
#ifdef BIO_PHASE_CLOCK
    long long sk_t[10];
    int sk_n = 0;
#define SK_CLK() do { if (sk_n < 10) sk_t[sk_n++] = clock64(); } while (0)
#else
#define SK_CLK() do { } while (0)
#endif
    SK_CLK();                                   // 0: kernel entry
    const DevModel<T>& m = stage_model(gm, smem);
    SK_CLK();                                   // 1: model staged
    EnvWork<T, CLS>* works = reinterpret_cast<EnvWork<T, CLS>*>(smem + ((sizeof(DevModel<T>) + 15) / 16) * 16);
    const int slot = threadIdx.x / G, lane = threadIdx.x % G;
    const int warp = threadIdx.x >> 5, warps_per_cta = blockDim.x >> 5;
    const int n_items = (n + EPW - 1) / EPW;
    for (int item = blockIdx.x + gridDim.x * warp; item < n_items; item += gridDim.x * warps_per_cta) {
    const int i = item * EPW + ((threadIdx.x & 31) / G);
    // all lanes of an env take the same branch; envs of a warp may differ only in i >= n
    const bool valid = i < n;
    const int ii = valid ? i : n - 1;          // out-of-range groups shadow the last env, stores are masked
    EnvWork<T, CLS>& E = works[slot];
    const unsigned long long env = (unsigned long long)(env_offset + ii);
    const int na = m.n_act, nd = m.n_dof, nm = m.n_muscles, Hh = c.horizon;
    const bool isd = lane < nd, ism = lane < nm, isa = lane < na;

    // every global load of the step's inputs is issued here, before anything waits on one of them: state,
    // bookkeeping, the action and the action history (one memory round trip; the L2 is cold in the benchmark)
    const T q_in = isd ? st.q[sx(st.aos, lane, ii, nd, n)] : T(0), u_in = isd ? st.u[sx(st.aos, lane, ii, nd, n)] : T(0);
    const T act_in = ism ? st.act[sx(st.aos, lane, ii, nm, n)] : T(0), lm_in = ism ? st.lm[sx(st.aos, lane, ii, nm, n)] : T(0);
    const T action_in = isa ? actions[(size_t)ii * na + lane] : T(0);
    const T last_in = isa ? st.last_action[sx(st.aos, lane, ii, na, n)] : T(0);
    T hv[BIO_MAX_HORIZON];
#pragma unroll
    for (int hh = 0; hh < BIO_MAX_HORIZON; hh++)
        hv[hh] = (isa && hh < Hh) ? st.history[sx(st.aos, hh * na + lane, ii, Hh * na, n)] : T(0);
    int istep = st.istep[ii];
    int hist_pos = st.hist_pos[ii];
    const bool first = st.first[ii] != 0;
    if (isd) { E.q[lane] = q_in; E.u[lane] = u_in; }
    if (ism) { E.act[lane] = act_in; E.lm[lane] = lm_in; }
    if (lane == 0) {   // bookkeeping read after the last evaluation: into L2 now
        asm volatile("prefetch.global.L2 [%0];" ::"l"(st.old_px + ii));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(st.ep_return + ii));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(st.ep_len + ii));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(st.episode + ii));
    }

    // ---- action pre-processing: lane = actuator ----
    T action = action_in;
    const bool lane_nan = action != action;
    const unsigned gmask = group_mask<G>();
    const bool nan = group_ballot<G>(lane_nan) != 0u;
    gsync<G>();   // E.q / E.u visible for the PD law
    if (nan) {
        action = T(0);
    } else if (c.use_pd) {
        T tau = T(0);
        if (lane < c.n_pd) {
            const int cx = c.pd_x_coord[lane], cv = c.pd_v_coord[lane];
            const T x = m.coord_dof[cx] >= 0 ? E.q[m.coord_dof[cx]] : m.coord_const[cx];
            const T v = m.coord_dof[cv] >= 0 ? E.u[m.coord_dof[cv]] : T(0);
            tau = c.pd_kp[lane] * (action - x) + c.pd_kv[lane] * (-v);
        }
        action = tau;
    }
    T last_action = T(0), curr = T(0);
    if (first) hist_pos = 0;
    if (isa) {
        // the first step of an episode fills the history with its action
        if (first) {
#pragma unroll
            for (int hh = 0; hh < BIO_MAX_HORIZON; hh++) hv[hh] = action;
        }
        last_action = first ? action : last_in;
        T sum = T(0);
#pragma unroll
        for (int hh = 0; hh < BIO_MAX_HORIZON; hh++) {
            if (hh < Hh) {
                sum += hh == hist_pos ? action : hv[hh];
                if (valid && (first || hh == hist_pos)) st.history[sx(st.aos, hh * na + lane, ii, Hh * na, n)] = action;
            }
        }
        curr = sum / T(Hh);
        E.ctrl[lane] = clampv(c.feed_mean_action ? curr : action, m.act_min[lane], m.act_max[lane]);
    }
    hist_pos = (hist_pos + 1) % Hh;
    if (ism) { E.vn[lane] = T(0); E.vna[lane] = T(0); }   // every control step starts its Newton solves from 0 (results do not depend on history)
    for (int t = lane; t < P2_MAXTASK; t += G) E.knot_hint[t] = 0;   // spline search hints (any start gives the same interval)
    gsync<G>();

    // ---- integrate one control step, evaluate at the new state ----
    SK_CLK();                                   // 2: state loaded, actions pre-processed
    coop_integrate<T, CLS>(m, c, E, lane, istep, seed, env);
    SK_CLK();                                   // 3: substeps done
    istep += 1;
    {   // reference rows read after the evaluation (reward: this row, observation: the next one): into L1 now
        const size_t r0 = (size_t)ref_row(c, istep), r1 = (size_t)ref_row(c, istep + 1);
        if (lane < c.ref_coords) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(c.ref_q + r0 * c.ref_coords + lane));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(c.ref_q + r1 * c.ref_coords + lane));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(c.ref_u + r1 * c.ref_coords + lane));
        }
        if (lane < 8) {
            const int sd = lane / 4, j = lane % 4;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(c.ref_body_pos + (r0 * c.ref_bodies + c.rew_refbody[sd][j]) * 3));
        }
        if (lane == 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(c.ref_com_pos + r0 * 3));
    }
    const int ext_pt = c.perturb ? c.perturb_obspt : -1;
    coop_eval_any<T, CLS>(m, E, lane, c.newton_iters, perturb_force(c, seed, env, T(istep) * c.dt), ext_pt, T(0), true);
    gsync<G>();
    // Size class 0: the observation rows of the warp's two envs are adjacent in memory; they are staged in the
    // (now idle) planar work arrays and written by the whole warp at the end, 16 bytes per lane (full 128-byte
    // lines for a caller's page-locked host buffer).  Size class 1: a warp writes its one row directly.
    static_assert(CLS != 0 || sizeof(E.k.p) >= 256 * sizeof(T), "observation stage does not fit the planar work arrays");
    SK_CLK();                                   // 4: full evaluation done
    T* orow = CLS == 0 ? reinterpret_cast<T*>(&E.k.p) : obs + (size_t)ii * c.obs_dim;
    if (valid || CLS == 0) coop_write_obs<T, CLS>(m, c, E, lane, istep, orow);

    SK_CLK();                                   // 5: observation row staged / written
    // ---- reward (env2D.py:267-358): lane-parallel partial sums ----
    const int row = ref_row(c, istep);
    T qpart = T(0);
    for (int k = lane; k < m.n_coords; k += G) {
        const int d = m.coord_dof[k];
        const T v = d >= 0 ? E.q[d] : m.coord_const[k];
        const T dd = v - c.ref_q[(size_t)row * c.ref_coords + k];
        qpart += dd * dd;
    }
    const T qerr = group_sum<T, G>(qpart) / T(m.n_coords);
    const T px = m.pel_dof[0] >= 0 ? E.q[m.pel_dof[0]] : T(0), py = m.pel_dof[1] >= 0 ? E.q[m.pel_dof[1]] : T(0);
    const T com_err = body_mse(E.com_pos, c.ref_com_pos + (size_t)row * 3);
    const T position_r = Num<T>::exp(T(-30) * qerr);
    const T com_r = Num<T>::exp(T(-20) * com_err);
    // feet terms: lanes 0..7 each take one (side, body)
    T fpart = T(0);
    if (lane < 8) {
        const int sd = lane / 4, j = lane % 4;
        fpart = body_mse(E.x.out.obs_pos[c.rew_obspt[sd][j]],
                         c.ref_body_pos + ((size_t)row * c.ref_bodies + c.rew_refbody[sd][j]) * 3);
    }
    // sum of the 4 bodies of each side: butterfly over 4 lanes
    fpart += __shfl_xor_sync(gmask, fpart, 1, G);
    fpart += __shfl_xor_sync(gmask, fpart, 2, G);
    const T foot_r = T(0.5) * Num<T>::exp(T(-20) * __shfl_sync(gmask, fpart, 0, G));
    const T foot_l = T(0.5) * Num<T>::exp(T(-20) * __shfl_sync(gmask, fpart, 4, G));
    T effort, a_error = T(0);
    if (c.effort_torque) {
        const T s2 = group_sum<T, G>(isa ? curr * curr : T(0));
        effort = Num<T>::sqrt(s2) / (c.max_actuation * T(na * na));
    } else {
        T a2 = T(0), cot = T(0);
        if (ism) {
            const int k = lane;
            const T hp = T(1.5707963267948966);
            a2 = E.act[k] * E.act[k];
            const T l = m.mus_slow_twitch[k];
            T se, ce, sa, ca;
            Num<T>::sincos(hp * E.ctrl[k], &se, &ce);
            Num<T>::sincos(hp * E.act[k], &sa, &ca);
            const T fa = T(40) * l * se + T(133) * (T(1) - l) * (T(1) - ce);
            const T fm = T(74) * l * sa + T(111) * (T(1) - l) * (T(1) - ca);
            const T ln = E.lm[k] / m.mus_lopt[k], v = E.lmdot[k];
            T g = T(0);
            if (ln < T(0.5)) g = T(0.5); else if (ln < T(1)) g = ln; else if (ln < T(1.5)) g = T(-2) * ln + T(3);
            const T es = T(0.25) * E.ffib[k] * -v, ew = E.fact[k] * -v;
            cot = m.mus_cot_mass[k] * fa + m.mus_cot_mass[k] * g * fm + (es > T(0) ? es : T(0)) + (ew > T(0) ? ew : T(0));
        }
        const T total = T(1.51) * m.total_mass + group_sum<T, G>(cot);
        a_error = Num<T>::exp(T(-2) * Num<T>::sqrt(group_sum<T, G>(a2)));
        effort = total / (T(20) * T(nm * nm));
    }
    const T old_px = st.old_px[ii];
    const T progress_coord = c.effort_use_dy ? py : px;
    const T prog = progress_coord - old_px + T(1);
    const T effort_r = Num<T>::exp(-effort / (prog > T(1) ? prog : T(1)));
    const T dact = isa ? curr - last_action : T(0);
    const T action_r = Num<T>::exp(-c.action_r_scale * Num<T>::sqrt(group_sum<T, G>(dact * dact)));
    T imit = position_r * com_r;
    if (c.reward_use_feet) imit *= (foot_l + foot_r);
    T rew = (T(0.5) + c.w_imitate) * imit + c.w_effort * effort_r + c.w_action * action_r;

    // ---- termination (env2D.py:237-265) ----
    T acc = T(0);
    bool fin = true;
    if (isd) {
        acc = Num<T>::abs(E.udot[lane]);
        fin = isfinite(E.q[lane]) && isfinite(E.u[lane]) && isfinite(E.udot[lane]);
    }
    const T maxacc = group_max<T, G>(acc);
    const bool finite = isfinite(rew) && (group_ballot<G>(!fin) == 0u);
    int reason = 0;
    if (!finite) { reason = BIO_DONE_NONFINITE; rew = T(0); }
    else if (E.x.out.obs_pos[c.term_obspt][1] < c.term_height) reason = BIO_DONE_HEIGHT;
    else if (E.max_limit > c.term_limit_force) reason = BIO_DONE_LIMIT_FORCE;
    else if (maxacc > c.term_acc) reason = BIO_DONE_ACCEL;
    else if (istep >= c.n_steps) reason = BIO_DONE_HORIZON;
    else if (c.term_feet_cross && E.x.out.obs_pos[c.feet_obspt[0]][2] - E.x.out.obs_pos[c.feet_obspt[1]][2] < T(0)) reason = BIO_DONE_FEET_CROSS;

    T ep_return = st.ep_return[ii] + rew;
    int ep_len = st.ep_len[ii] + 1;
    long long episode = st.episode[ii];
    int first_next = 0;
    if (valid && lane == 0) {
        reward[ii] = rew;
        done[ii] = reason != 0;
        if (terms) {
            T* tr = terms + (size_t)ii * c.n_reward_terms;
            const bool z = reason == BIO_DONE_NONFINITE;
            tr[0] = z ? T(0) : position_r; tr[1] = z ? T(0) : com_r; tr[2] = z ? T(0) : foot_l; tr[3] = z ? T(0) : foot_r;
            if (c.n_reward_terms > 4) tr[4] = z ? T(0) : a_error;
        }
        if (stats) {
            if (nan) atomicAdd(&stats[10], 1.0);
            if (reason) {
                atomicAdd(&stats[1], 1.0);
                atomicAdd(&stats[2], (double)ep_return);
                atomicAdd(&stats[3], (double)ep_len);
                int bit = 0;
                while (!((reason >> bit) & 1)) bit++;
                atomicAdd(&stats[4 + bit], 1.0);
            }
        }
    }
    SK_CLK();                                   // 6: reward, termination, outputs of lane 0
    if (c.ex.any) {   // optional extra outputs (BioStepExtra) of the end-of-step evaluation, before any reset
        const StepExtra<T>& x = c.ex;
        if (valid) {
            if (x.done_reason && lane == 0) x.done_reason[ii] = reason;
            if (x.udot && isd) x.udot[(size_t)ii * nd + lane] = E.udot[lane];
            if (ism) {
                if (x.tendon_force) x.tendon_force[(size_t)ii * nm + lane] = E.vn[lane];
                if (x.fiber_force) x.fiber_force[(size_t)ii * nm + lane] = E.ffib[lane];
                if (x.fiber_vel) x.fiber_vel[(size_t)ii * nm + lane] = E.lmdot[lane];
            }
            if (x.contact && lane < 12) x.contact[(size_t)ii * 12 + lane] = E.contact[lane / 6][lane % 6];
            if (x.limit_force && lane < m.n_limits) x.limit_force[(size_t)ii * m.n_limits + lane] = E.limf[lane];
        }
        if (x.terminal_obs) {
            if constexpr (CLS != 0) {           // a warp owns one env: its row goes out once more
                if (valid && reason != 0) coop_write_obs<T, CLS>(m, c, E, lane, istep, x.terminal_obs + (size_t)ii * c.obs_dim);
            }
        }
    }
    gsync<G>();   // everyone is done with E.x.out before a reset overwrites it
    // Size class 0: both staged rows go out now (the whole warp writes: 16 bytes per lane)
    if constexpr (CLS == 0) {
        coop_flush_obs<T, CLS>(works, obs, c.obs_dim, item * EPW, n, 3u);
        if (c.ex.terminal_obs) {                // rows of the envs that finished in this step
            const unsigned fin = __ballot_sync(0xffffffffu, valid && reason != 0);
            const unsigned rows = ((fin & 0xffffu) ? 1u : 0u) | ((fin >> 16) ? 2u : 0u);
            if (rows) coop_flush_obs<T, CLS>(works, c.ex.terminal_obs, c.obs_dim, item * EPW, n, rows);
        }
    }
    // Auto-reset.  The branch is taken by the whole warp (see group_mask): when only one of its two envs
    // resets, the other one repeats the evaluation of its current state and drops the results.
    const bool do_reset = reason != 0 && c.auto_reset;
    const unsigned reset_bits = __ballot_sync(0xffffffffu, do_reset);
    if (reset_bits) {
        if (do_reset) {
            episode += 1;
            int idx = 0;
            if (!c.test_mode && c.reset_max_index > 0)
                idx = (int)(bio_rand(seed, env, (unsigned long long)episode, 1) % (unsigned long long)(c.reset_max_index + 1));
            idx = idx > c.ref_rows - 1 ? c.ref_rows - 1 : idx;
            for (int k = lane; k < m.n_coords; k += G) {
                const int d = m.coord_dof[k];
                if (d < 0) continue;
                E.q[d] = c.ref_q[(size_t)idx * c.ref_coords + k];
                E.u[d] = c.ref_u[(size_t)idx * c.ref_coords + k];
            }
            if (ism) { E.act[lane] = m.mus_default_act[lane]; E.lm[lane] = c.ref_lm0[(size_t)idx * nm + lane]; }
            if (isa) E.ctrl[lane] = T(0);
            istep = idx;
            first_next = 1;
            ep_return = T(0);
            ep_len = 0;
        }
        // Newton start of both envs of the warp: 0 (the slots hold the read-outs of the full evaluation)
        if (ism) { E.vn[lane] = T(0); E.vna[lane] = T(0); }
        gsync<G>();
        coop_eval_any<T, CLS>(m, E, lane, c.newton_iters, perturb_force(c, seed, env, T(istep) * c.dt), ext_pt, T(0), true);
        gsync<G>();
        if (do_reset && (valid || CLS == 0)) coop_write_obs<T, CLS>(m, c, E, lane, istep, orow);
        if constexpr (CLS == 0) {
            // rows of the envs that were reset (bit e: env e of the warp)
            const unsigned rows = ((reset_bits & 0xffffu) ? 1u : 0u) | ((reset_bits >> 16) ? 2u : 0u);
            coop_flush_obs<T, CLS>(works, obs, c.obs_dim, item * EPW, n, rows);
        }
    }
    SK_CLK();                                   // 7: rows flushed, auto-reset handled
    // ---- write back ----
    if (valid) {
        if (isd) { st.q[sx(st.aos, lane, ii, nd, n)] = E.q[lane]; st.u[sx(st.aos, lane, ii, nd, n)] = E.u[lane]; }
        if (ism) { st.act[sx(st.aos, lane, ii, nm, n)] = E.act[lane]; st.lm[sx(st.aos, lane, ii, nm, n)] = E.lm[lane]; }
        if (isa) st.last_action[sx(st.aos, lane, ii, na, n)] = first_next ? T(0) : curr;
        if (lane == 0) {
            st.old_px[ii] = progress_coord;
            st.istep[ii] = istep;
            st.first[ii] = first_next;
            st.hist_pos[ii] = hist_pos;
            st.ep_return[ii] = ep_return;
            st.ep_len[ii] = ep_len;
            st.episode[ii] = episode;
        }
    }
    gsync<G>();   // the warp's work slots are reused by its next item
    SK_CLK();                                   // 8: state written back
#ifdef BIO_PHASE_CLOCK
    if (threadIdx.x == 0 && blockIdx.x == 0)
        printf("step cycles: stage %lld  load+actions %lld  substeps %lld  full eval %lld  obs %lld  reward+done %lld  flush+reset %lld  write back %lld\n",
               sk_t[1] - sk_t[0], sk_t[2] - sk_t[1], sk_t[3] - sk_t[2], sk_t[4] - sk_t[3], sk_t[5] - sk_t[4], sk_t[6] - sk_t[5],
               sk_t[7] - sk_t[6], sk_t[8] - sk_t[7]);
#endif
    }
    if (stats && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats[0], (double)n);
}

}  // namespace bio
