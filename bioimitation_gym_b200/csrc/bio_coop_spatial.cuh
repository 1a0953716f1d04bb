// bio_coop_spatial.cuh -- kinematics of the general (spatial, 3D) evaluation as warp scans.
//
// Same idea as p2_phase_b_scan for the planar models: the walk root joint -> leaf of a chain is a
// sequence of associative updates, so with lane = (chain, step) -- 16 steps per chain, two chains per
// warp -- every quantity is a prefix over the steps of a chain:
//   orientation   product of unit quaternions (one per rotation axis; translations are the identity)
//   position      sum of the displacements (joint offsets turned by the orientation before the step,
//                 translations along the axes of the parent frame), started at the step that fixes O
//   velocity      sum of (motion vector * rate), spatial 6-vectors about O in ground axes
//   bias accel.   sum of motion vector * d2s/dq2 q'^2 + (velocity before the step) x (motion vector) * rate
// 4 shuffle rounds per prefix instead of a level-by-level walk with a barrier per tree level.  Results
// go to the arrays of the general evaluation (K.Rr, K.VA, K.S, E.O), so the other phases of
// coop_eval are unchanged.  Used when the host found the root-plus-chains shape (PlanarProg::chain_ok).
#pragma once

namespace bio {

template <typename T>
__device__ __forceinline__ T shfl_up16(const unsigned mask, const T v, const int off) { return __shfl_up_sync(mask, v, off, 16); }
template <typename T>
__device__ __forceinline__ T shfl_at16(const unsigned mask, const T v, const int src) { return __shfl_sync(mask, v, src, 16); }

// v turned by the unit quaternion (w, x, y, z)
template <typename T>
__device__ __forceinline__ void quat_rotate(const T w, const T x, const T y, const T z, const T* v, T* o) {
    const T tx = T(2) * (y * v[2] - z * v[1]), ty = T(2) * (z * v[0] - x * v[2]), tz = T(2) * (x * v[1] - y * v[0]);
    o[0] = v[0] + w * tx + (y * tz - z * ty);
    o[1] = v[1] + w * ty + (z * tx - x * tz);
    o[2] = v[2] + w * tz + (x * ty - y * tx);
}

template <typename T, int CLS>
__device__ __forceinline__ void p3_phase_b_scan(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    constexpr int G = CoopCls<CLS>::G;
    static_assert(G == 32, "the spatial scan uses two 16-lane segments");
    const PlanarProg<T>& pr = m.prog;
    auto& K = E.k.g;
    const unsigned mask = group_mask<G>();
    const int l = lane >> 4, i = lane & 15;
    const bool live = l < pr.n_branches && i < pr.ch_n[l];
    int code = 0, sc = i;
    T j[3] = {T(0), T(0), T(0)}, n[3] = {T(1), T(0), T(0)};
    T st = T(0), ds = T(0), sd = T(0), acc = T(0);
    T qw = T(1), qx = T(0), qy = T(0), qz = T(0);
    if (l < pr.n_branches) sc = pr.ch_scan[l][i];
    if (live) {
        code = pr.ch_code[l][i];
        const int a = code & 255, d = (code >> 12) & 31;
        T j3;
        ld4(pr.ch_j[l][i], j[0], j[1], j[2], j3);
        n[0] = m.axis_vec[a][0]; n[1] = m.axis_vec[a][1]; n[2] = m.axis_vec[a][2];
        const T s = K.ax_s[a], qd = d != 31 ? E.u[d] : T(0);
        ds = K.ax_ds[a];
        sd = ds * qd;
        acc = K.ax_dds[a] * qd * qd;
        if (code & P2_F_ROT) {
            T sh, ch;
            Num<T>::sincos(T(0.5) * s, &sh, &ch);
            qw = ch; qx = sh * n[0]; qy = sh * n[1]; qz = sh * n[2];
        } else {
            st = s;
        }
    }
    const bool is_rot = (code & P2_F_ROT) != 0;
    const int first = sc & 15, o_step = (sc >> 4) & 15, in_dof = (sc >> 8) & 3;
    // orientation after every step: q_i = q_(i-off) * q_i
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
        const T aw = shfl_up16(mask, qw, off), ax = shfl_up16(mask, qx, off);
        const T ay = shfl_up16(mask, qy, off), az = shfl_up16(mask, qz, off);
        if (i >= off) {
            const T w = aw * qw - ax * qx - ay * qy - az * qz;
            const T x = aw * qx + ax * qw + ay * qz - az * qy;
            const T y = aw * qy - ax * qz + ay * qw + az * qx;
            const T z = aw * qz + ax * qy - ay * qx + az * qw;
            qw = w; qx = x; qy = y; qz = z;
        }
    }
    // orientation before the step, and of the parent body (before the first step of this step's body)
    T ew = shfl_up16(mask, qw, 1), ex = shfl_up16(mask, qx, 1), ey = shfl_up16(mask, qy, 1), ez = shfl_up16(mask, qz, 1);
    if (i == 0) { ew = T(1); ex = ey = ez = T(0); }
    const T pw = shfl_at16(mask, ew, first), px = shfl_at16(mask, ex, first);
    const T py = shfl_at16(mask, ey, first), pz = shfl_at16(mask, ez, first);
    // axis in ground axes: a rotation turns about the axis of the current frame, a translation runs along
    // the axis of the parent frame; joint offset of the first step of a body in ground axes
    T aw[3], jw[3];
    quat_rotate(is_rot ? ew : pw, is_rot ? ex : px, is_rot ? ey : py, is_rot ? ez : pz, n, aw);
    quat_rotate(ew, ex, ey, ez, j, jw);
    // displacement of the step; steps up to the one that fixes O build O, the later ones the position about O
    T r[3], o[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const T d = jw[c] + aw[c] * st;
        r[c] = i > o_step ? d : T(0);
        o[c] = i > o_step ? T(0) : d;
    }
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const T a0 = shfl_up16(mask, r[c], off), a1 = shfl_up16(mask, o[c], off);
            if (i >= off) { r[c] += a0; o[c] += a1; }
        }
    }
    // motion vector about O: rotation (a, r x a), translation (0, a)
    T S[6];
    if (is_rot) {
        S[0] = aw[0]; S[1] = aw[1]; S[2] = aw[2];
        cross3(r, aw, S + 3);
    } else {
        S[0] = S[1] = S[2] = T(0);
        S[3] = aw[0]; S[4] = aw[1]; S[5] = aw[2];
    }
    // velocity after / before the step
    T V[6], Ve[6];
#pragma unroll
    for (int c = 0; c < 6; c++) V[c] = S[c] * sd;
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
#pragma unroll
        for (int c = 0; c < 6; c++) {
            const T a0 = shfl_up16(mask, V[c], off);
            if (i >= off) V[c] += a0;
        }
    }
#pragma unroll
    for (int c = 0; c < 6; c++) Ve[c] = V[c] - S[c] * sd;
    // bias acceleration: S * acc + (V x S) * rate, angular: w x S_w, linear: w x S_v + v x S_w
    T A[6], c1[3], c2[3], c3[3];
    cross3(Ve, S, c1); cross3(Ve, S + 3, c2); cross3(Ve + 3, S, c3);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        A[c] = S[c] * acc + c1[c] * sd;
        A[3 + c] = S[3 + c] * acc + (c2[c] + c3[c]) * sd;
    }
    if (i == 0) { A[3] -= m.gravity[0]; A[4] -= m.gravity[1]; A[5] -= m.gravity[2]; }
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
#pragma unroll
        for (int c = 0; c < 6; c++) {
            const T a0 = shfl_up16(mask, A[c], off);
            if (i >= off) A[c] += a0;
        }
    }
    // motion vector of the dof: sum over its (<= 3, consecutive) axes
    T Sd[6];
#pragma unroll
    for (int c = 0; c < 6; c++) {
        Sd[c] = ds * S[c];
        const T p1 = shfl_up16(mask, Sd[c], 1), p2 = shfl_up16(mask, Sd[c], 2);
        if (in_dof >= 1) Sd[c] += p1;
        if (in_dof >= 2) Sd[c] += p2;
    }
    // publish (both chains repeat the root joint and store the same values)
    if (live) {
        if (code & P2_F_SPUB) {
            const int d = (code >> 12) & 31;
            st4(K.S[d], Sd[0], Sd[1], Sd[2], Sd[3]); st2(K.S[d] + 4, Sd[4], Sd[5]);
        }
        if (i == o_step) { E.O[0] = o[0]; E.O[1] = o[1]; E.O[2] = o[2]; }
        if (code & P2_F_LAST) {
            const int b = (code >> 8) & 15;
            const T xx = qx * qx, yy = qy * qy, zz = qz * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz;
            const T wx = qw * qx, wy = qw * qy, wz = qw * qz;
            T* R = K.Rr[b];                        // rows of the rotation, position in the fourth column
            st4(R, T(1) - T(2) * (yy + zz), T(2) * (xy - wz), T(2) * (xz + wy), r[0]);
            st4(R + 4, T(2) * (xy + wz), T(1) - T(2) * (xx + zz), T(2) * (yz - wx), r[1]);
            st4(R + 8, T(2) * (xz - wy), T(2) * (yz + wx), T(1) - T(2) * (xx + yy), r[2]);
            st4(K.VA[b], V[0], V[1], V[2], V[3]); st4(K.VA[b] + 4, V[4], V[5], A[0], A[1]);
            st4(K.VA[b] + 8, A[2], A[3], A[4], A[5]);
        }
    }
}

// ---------------------------------------------------------------------------
// Phase A of the spatial evaluation from the packed task descriptors (PlanarProg::at_i4 / at_f4, the tables of
// p2_phase_a): lane = joint-axis function | location function of a moving path point, one 2 x 16-byte descriptor
// per task (kind, destination, dof, knot range, coefficients) instead of seven dependent index loads, the spline
// interval of the previous evaluation as the search hint.  Outputs: value and two derivatives of every axis
// function (the half-angle quaternions are formed in the scan), location and d/dq of the moving points.
// ---------------------------------------------------------------------------
template <typename T, int CLS>
__device__ __forceinline__ void p3_phase_a(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    const PlanarProg<T>& pr = m.prog;
    auto& K = E.k.g;
    for (int t = lane; t < pr.n_atasks; t += 32) {
        const int4 ti = *reinterpret_cast<const int4*>(pr.at_i4[t]);
        T c0, c1, add, f3;
        ld4(pr.at_f4[t], c0, c1, add, f3);
        const int kind = ti.x & 3, dst = ti.x >> 8, d = ti.y;
        const T x = d >= 0 ? E.q[d] : T(0);
        T s, ds, dds = T(0);
        if (kind == BIO_FUNC_SPLINE) {
            // SimmSpline: cubic piece of the interval [knot i, knot i + 1), straight lines beyond the end knots
            const int kb = ti.z, n = ti.w;
            int i = E.knot_hint[t];
            i = i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
            while (i > 0 && x < m.knot_x[kb + i]) i--;
            while (i + 1 < n - 1 && x >= m.knot_x[kb + i + 1]) i++;
            E.knot_hint[t] = (int8_t)i;
            const bool below = x <= c0, above = x >= c1;
            i = above ? n - 1 : (below ? 0 : i);
            T k0, k1, k2, k3;
            ld4(m.knot_c[kb + i], k0, k1, k2, k3);
            if (below || above) { k2 = T(0); k3 = T(0); }
            const T dx = x - m.knot_x[kb + i];
            s = k0 + dx * (k1 + dx * (k2 + dx * k3));
            ds = k1 + dx * (T(2) * k2 + T(3) * dx * k3);
            dds = T(2) * k2 + T(6) * dx * k3;
        } else {
            const bool lin = kind == BIO_FUNC_LINEAR;
            s = lin ? c0 * x + c1 : c0;
            ds = lin ? c0 : T(0);
        }
        if (dst < 64) {
            K.ax_s[dst] = s; K.ax_ds[dst] = ds; K.ax_dds[dst] = dds;
        } else {
            const int k = (dst - 64) / 3, c = (dst - 64) % 3;
            K.mv[k][c] = s; K.mv[k][3 + c] = ds;
        }
    }
}

// ---------------------------------------------------------------------------
// Phase E of the articulated-body path: body inertias and forces with THREE lanes per body (lane = 4 * body + part;
// the wrench gather uses all four).  The muscle wrench sources of the body are dealt over its four lanes, its
// (<= 4) contact spheres one per lane, the perturbation force goes to lane 3; a quad butterfly leaves the total
// wrench in every lane.  Lane `part` < 3 then owns row `part` of the body's inertia about O (R I_b R^T + the
// parallel-axis term), the matching component of I V and I A (the three lanes exchange I V by shuffles for the
// V x* I V term), and writes columns `part` and 3 + `part` of the 6 x 6 spatial inertia and components
// `part`, 3 + `part` of the body force (BIc, read by p3_aba).  Components are picked with the unit vector e_p of
// the lane: (h x v)_p = (e_p x h) . v, and e_p x h is also the lower half of column p.  lane = dof afterwards:
// generalized force and limit damping next to the motion vector.  (7 busy lanes -> 21; ncu: phase E was 11 % of the
// instructions at 13 threads per instruction.)
// ---------------------------------------------------------------------------
template <typename T, int CLS, bool FAST = false>
__device__ __forceinline__ void p3_phase_e(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    const PlanarProg<T>& pr = m.prog;
    auto& K = E.k.g;
    const unsigned mask = 0xffffffffu;
    const int b = lane >> 2, part = lane & 3;        // b < BIO_MAX_BODIES: every lane reads valid rows
    const bool live = b < ldv(m.n_bodies);
    T Wn0 = T(0), Wn1 = T(0), Wn2 = T(0), Wf0 = T(0), Wf1 = T(0), Wf2 = T(0);
    if (live) {
        if (ldv(m.n_muscles) > 0) {
            for (int k = pr.inc_begin[b] + part; k < pr.inc_begin[b + 1]; k += 4) {
                T w0, w1, w2, w3, w4, w5;
                const int e = pr.inc_src[k];
                ld4(E.x.src6.a[e], w0, w1, w2, w3);
                ld2(E.x.src6.b[e], w4, w5);
                Wn0 += w0; Wn1 += w1; Wn2 += w2; Wf0 += w3; Wf1 += w4; Wf2 += w5;
            }
        }
        const unsigned sp = (pr.sph_pk[b] >> (8 * part)) & 255u;
        if (sp != 255u && E.sphF[sp][1] != T(0)) {
            T n[3];
            cross3(E.sphx[sp], E.sphF[sp], n);
            Wn0 += n[0]; Wn1 += n[1]; Wn2 += n[2];
            Wf0 += E.sphF[sp][0]; Wf1 += E.sphF[sp][1]; Wf2 += E.sphF[sp][2];
        }
        const int ext_pt = ((ldv(E.ev.flags) >> 8) & 255) - 1;
        if (part == 3 && ext_pt >= 0 && m.obs_body[ext_pt] == b) {
            T x[3], n[3];
            const T fx[3] = {ldv(E.ev.fx), T(0), T(0)};
            pose_point(K.Rr[b], m.obs_loc[ext_pt], x);
            cross3(x, fx, n);
            Wn0 += n[0]; Wn1 += n[1]; Wn2 += n[2]; Wf0 += fx[0];
        }
    }
    Wn0 += __shfl_xor_sync(mask, Wn0, 1); Wn1 += __shfl_xor_sync(mask, Wn1, 1); Wn2 += __shfl_xor_sync(mask, Wn2, 1);
    Wf0 += __shfl_xor_sync(mask, Wf0, 1); Wf1 += __shfl_xor_sync(mask, Wf1, 1); Wf2 += __shfl_xor_sync(mask, Wf2, 1);
    Wn0 += __shfl_xor_sync(mask, Wn0, 2); Wn1 += __shfl_xor_sync(mask, Wn1, 2); Wn2 += __shfl_xor_sync(mask, Wn2, 2);
    Wf0 += __shfl_xor_sync(mask, Wf0, 2); Wf1 += __shfl_xor_sync(mask, Wf1, 2); Wf2 += __shfl_xor_sync(mask, Wf2, 2);
    {
        const int p = part < 3 ? part : 0;
        const T ux = p == 0 ? T(1) : T(0), uy = p == 1 ? T(1) : T(0), uz = p == 2 ? T(1) : T(0);
        T R0, R1, R2, R3, R4, R5, R6, R7, R8, q0, q1, q2, r0, r1, r2, qp;
        ld4(K.Rr[b], R0, R1, R2, q0); ld4(K.Rr[b] + 4, R3, R4, R5, q1); ld4(K.Rr[b] + 8, R6, R7, R8, q2);
        ld4(K.Rr[b] + 4 * p, r0, r1, r2, qp);        // row p of the pose
        const T cm0 = m.body_com[b][0], cm1 = m.body_com[b][1], cm2 = m.body_com[b][2];
        const T cx = R0 * cm0 + R1 * cm1 + R2 * cm2 + q0;
        const T cy = R3 * cm0 + R4 * cm1 + R5 * cm2 + q1;
        const T cz = R6 * cm0 + R7 * cm1 + R8 * cm2 + q2;
        const T cp = r0 * cm0 + r1 * cm1 + r2 * cm2 + qp;
        const T mb = m.body_mass[b], mcc = mb * (cx * cx + cy * cy + cz * cz), mcp = mb * cp;
        const T* i6 = m.body_inertia[b];
        const T i0 = i6[0], i1 = i6[1], i2 = i6[2], i3 = i6[3], i4 = i6[4], i5 = i6[5];
        const T t0 = r0 * i0 + r1 * i3 + r2 * i4, t1 = r0 * i3 + r1 * i1 + r2 * i5, t2 = r0 * i4 + r1 * i5 + r2 * i2;
        // row p of the inertia about O
        const T Ig0 = t0 * R0 + t1 * R1 + t2 * R2 - mcp * cx + ux * mcc;
        const T Ig1 = t0 * R3 + t1 * R4 + t2 * R5 - mcp * cy + uy * mcc;
        const T Ig2 = t0 * R6 + t1 * R7 + t2 * R8 - mcp * cz + uz * mcc;
        const T hx = mb * cx, hy = mb * cy, hz = mb * cz;
        const T g0 = uy * hz - uz * hy, g1 = uz * hx - ux * hz, g2 = ux * hy - uy * hx;      // e_p x h
        T V0, V1, V2, V3, V4, V5, A0, A1, A2, A3, A4, A5;
        ld4(K.VA[b], V0, V1, V2, V3); ld4(K.VA[b] + 4, V4, V5, A0, A1); ld4(K.VA[b] + 8, A2, A3, A4, A5);
        // component p of I V and I A: angular Ig . w + (h x v)_p, linear m v_p - (h x w)_p
        const T IVn = Ig0 * V0 + Ig1 * V1 + Ig2 * V2 + (g0 * V3 + g1 * V4 + g2 * V5);
        const T IVf = mb * (ux * V3 + uy * V4 + uz * V5) - (g0 * V0 + g1 * V1 + g2 * V2);
        const T IAn = Ig0 * A0 + Ig1 * A1 + Ig2 * A2 + (g0 * A3 + g1 * A4 + g2 * A5);
        const T IAf = mb * (ux * A3 + uy * A4 + uz * A5) - (g0 * A0 + g1 * A1 + g2 * A2);
        // all of I V from the three lanes of the body; component p of w x (I V)_n + v x (I V)_f and of w x (I V)_f
        const T n0 = __shfl_sync(mask, IVn, 0, 4), n1 = __shfl_sync(mask, IVn, 1, 4), n2 = __shfl_sync(mask, IVn, 2, 4);
        const T f0 = __shfl_sync(mask, IVf, 0, 4), f1 = __shfl_sync(mask, IVf, 1, 4), f2 = __shfl_sync(mask, IVf, 2, 4);
        const T a0 = uy * V2 - uz * V1, a1 = uz * V0 - ux * V2, a2 = ux * V1 - uy * V0;      // e_p x w
        const T b0 = uy * V5 - uz * V4, b1 = uz * V3 - ux * V5, b2 = ux * V4 - uy * V3;      // e_p x v
        const T fn = IAn + (a0 * n0 + a1 * n1 + a2 * n2) + (b0 * f0 + b1 * f1 + b2 * f2) - (ux * Wn0 + uy * Wn1 + uz * Wn2);
        const T ff = IAf + (a0 * f0 + a1 * f1 + a2 * f2) - (ux * Wf0 + uy * Wf1 + uz * Wf2);
        if (live && part < 3) {
            T* o = K.BIc[b] + 6 * p;
            st2(o, Ig0, Ig1); st2(o + 2, Ig2, g0); st2(o + 4, g1, g2);
            st2(o + 18, -g0, -g1); st2(o + 20, -g2, mb * ux); st2(o + 22, mb * uy, mb * uz);
            K.BIc[b][36 + p] = fn; K.BIc[b][39 + p] = ff;
        }
    }
    // lane = dof: generalized force of its (<= 2) limits / moving points and its actuator, h * limit damping
    if (lane < ldv(m.n_dof)) {
        const int d = lane;
        T qf = T(0), ld = T(0);
        if (FAST || m.gdof_ok) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int l = m.gdof_lim[d][j], pt = m.gdof_movpt[d][j];
                if (l >= 0) { qf += E.limf[l]; ld += E.limD[l]; }
                if (pt >= 0) qf += K.mq[m.pt_mov[pt]];
            }
            const int a = m.gdof_act[d];
            if (a >= 0) qf += E.ctrl[a];
        } else {
            for (int l = 0; l < m.n_limits; l++) if (m.lim_dof[l] == d) { qf += E.limf[l]; ld += E.limD[l]; }
            for (int k = 0; k < m.n_moving; k++) { const int pt = m.moving_pt[k]; if (m.pt_dof[pt] == d) qf += K.mq[k]; }
            if (m.is_torque) for (int a = 0; a < m.n_act; a++) if (m.act_dof[a] == d) qf += E.ctrl[a];
        }
        K.S[d][6] = qf; K.S[d][7] = ldv(E.ev.himp) * ld;   // next to the motion vector: one read per elimination step
    }
}

// ---------------------------------------------------------------------------
// Articulated-body pass of the spatial evaluation (replaces the composite inertias, the joint-space
// matrix and its sparse L^T D L for root-plus-chains models; PlanarProg::aba_*).
//
// Every spatial quantity is expressed in ground axes about the common point O, so handing an inertia or a
// force to the parent body is a plain sum.  A half-warp owns a chain; lane c < 6 of it keeps column c of
// the chain's articulated inertia I^A (symmetric 6 x 6) in registers, lane 6 the force p^A as a seventh
// column.  Per dof d, from the leaf (motion vector S, generalized force Q, h * limit damping Ld):
//   U = I^A S          lane c: U_c = S . column_c  (lane 6: S . p^A)       -- no reduction across lanes
//   D = S . U + Ld     every lane, from the U exchanged through shared memory
//   column_c -= U (U_c - [c = 6] Q) / D            i.e.  I^a = I^A - U U^T / D,  p^a = p^A + U (Q - S . p^A) / D
// and the coefficients (U / D, -(Q - S . p^A) / D) stay in shared memory for the way back.  A body joins
// the chain by adding its columns (phase E wrote them: BIc); the implicit contact damping h J^T D J of a foot's
// active spheres is added to the foot's columns on the fly (J = [-(p x), 1]: column c is (p x g, g) with
// g = diag(D0, D1, D0) (e_c x p) for c < 3, diag(D0, D1, D0) e_(c-3) else).  Both half-warps then take the
// root body plus the two chain heads and eliminate the root's dofs (redundantly), and every lane walks
// the accelerations back down: qdd = u / D - (U / D) . a,  a += S qdd.
// The same block elimination from the leaves as coop_solve and the oracle's Cholesky, in another order of
// operations; one __syncwarp per dof, no shuffles.
// ---------------------------------------------------------------------------
template <typename T, int CLS>
__device__ __forceinline__ void p3_aba(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane, const T h_imp) {
    const PlanarProg<T>& pr = m.prog;
    auto& K = E.k.g;
    auto& X = E.x.aba;
    const int grp = lane >> 4;
    const int cc = (lane & 15) < 6 ? (lane & 15) : 6;    // lanes 6..15 of a half-warp all carry the force column
    const bool fcol = cc == 6;
    // active contacts of this env (same value on every lane)
    unsigned act = 0u;
    if (h_imp > T(0)) act = __ballot_sync(0xffffffffu, lane < m.n_spheres && E.sphD[lane < m.n_spheres ? lane : 0][1] > T(0));
    // unit motion of column cc: angular part ua (c < 3), linear part ub (3 <= c < 6)
    const T ua0 = cc == 0 ? T(1) : T(0), ua1 = cc == 1 ? T(1) : T(0), ua2 = cc == 2 ? T(1) : T(0);
    const T ub0 = cc == 3 ? T(1) : T(0), ub1 = cc == 4 ? T(1) : T(0), ub2 = cc == 5 ? T(1) : T(0);
    T col[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    auto add_body = [&](const int b) {
        const T* src = &K.BIc[b][cc * 6];
        T v0, v1, v2, v3, v4, v5;
        ld2(src, v0, v1); ld2(src + 2, v2, v3); ld2(src + 4, v4, v5);
        col[0] += v0; col[1] += v1; col[2] += v2; col[3] += v3; col[4] += v4; col[5] += v5;
        unsigned mm = act & (unsigned)pr.body_sph_mask[b];
        while (mm) {
            const int s = __ffs(mm) - 1;
            mm &= mm - 1u;
            const T px = E.sphx[s][0], py = E.sphx[s][1], pz = E.sphx[s][2];
            const T d0 = h_imp * E.sphD[s][0], d1 = h_imp * E.sphD[s][1];
            // velocity of the contact point for the unit motion, times the damping
            const T gx = d0 * (ua1 * pz - ua2 * py + ub0);
            const T gy = d1 * (ua2 * px - ua0 * pz + ub1);
            const T gz = d0 * (ua0 * py - ua1 * px + ub2);
            col[0] += py * gz - pz * gy; col[1] += pz * gx - px * gz; col[2] += px * gy - py * gx;
            col[3] += gx; col[4] += gy; col[5] += gz;
        }
    };
    T* ux = X.Ux[0][grp];                                // exchange buffers, swapped every step
    T* uy = X.Ux[1][grp];
    auto eliminate = [&](const int d, const bool on) {
        T s0, s1, s2, s3, s4, s5, s6, s7;
        ld4(K.S[d], s0, s1, s2, s3); ld4(K.S[d] + 4, s4, s5, s6, s7);
        const T Uc = s0 * col[0] + s1 * col[1] + s2 * col[2] + s3 * col[3] + s4 * col[4] + s5 * col[5];
        ux[cc] = Uc;
        __syncwarp();
        T u0, u1, u2, u3, u4, u5, u6, u7;
        ld4(ux, u0, u1, u2, u3); ld4(ux + 4, u4, u5, u6, u7);
        { T* t = ux; ux = uy; uy = t; }
        // S[d][6] = generalized force Q, S[d][7] = h * limit damping of the dof (phase E)
        const T D = s0 * u0 + s1 * u1 + s2 * u2 + s3 * u3 + s4 * u4 + s5 * u5 + s7;
        T coef = (Uc - (fcol ? s6 : T(0))) * Num<T>::rcp(D);
        if (on) X.W[d][cc] = coef; else coef = T(0);
        col[0] -= u0 * coef; col[1] -= u1 * coef; col[2] -= u2 * coef;
        col[3] -= u3 * coef; col[4] -= u4 * coef; col[5] -= u5 * coef;
    };
    // chains, leaf -> root (a chain with fewer dofs idles through the last steps)
    const int nst = pr.aba_nsteps;
    for (int s = 0; s < nst; s++) {
        const int code = pr.aba_step[grp][s];
        const bool on = code != 255;
        if (on && (code >> 4)) add_body((code >> 4) - 1);
        eliminate(on ? (code & 15) : 0, on);
    }
    // chain heads to shared memory (in place of the first body of the chain), then the root
    if (grp < pr.n_branches) {
        T* dst = &K.BIc[pr.gch_body[grp][0]][cc * 6];
        st2(dst, col[0], col[1]); st2(dst + 2, col[2], col[3]); st2(dst + 4, col[4], col[5]);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 6; r++) col[r] = T(0);
    add_body(pr.root_body);
    for (int l = 0; l < pr.n_branches; l++) {
        const T* src = &K.BIc[pr.gch_body[l][0]][cc * 6];
        T v0, v1, v2, v3, v4, v5;
        ld2(src, v0, v1); ld2(src + 2, v2, v3); ld2(src + 4, v4, v5);
        col[0] += v0; col[1] += v1; col[2] += v2; col[3] += v3; col[4] += v4; col[5] += v5;
    }
    const int nroot = pr.aba_nroot;
    // way back: every lane carries the spatial acceleration beyond the bias term
    T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0), a4 = T(0), a5 = T(0);
    if (pr.aba_freeroot) {
        // Free root joint (three translations along the ground axes, then three rotations about axes through O; no
        // limit or moving point on its dofs, actuators on the rotations only: PlanarProg::aba_fr): its motion vectors
        // span all of R^6, so the
        // root's spatial acceleration solves  I^A a = -p^A  directly -- with I^A = [[A, B], [B^T, M]] two symmetric
        // 3 x 3 inverses (the mass block M and the Schur complement A - B M^-1 B^T) instead of six elimination and six
        // way-back steps with their barriers -- and the coordinates follow as  qdd_trans = a_v,  qdd_rot = R^-1 a_w
        // with R = the rotation axes.  Every lane needs all of I^A and p^A: the columns go through shared memory once
        // (in place of the root body's own columns, which every lane has consumed).
        __syncwarp();
        T* g = K.BIc[pr.root_body];
        st2(g + cc * 6, col[0], col[1]); st2(g + cc * 6 + 2, col[2], col[3]); st2(g + cc * 6 + 4, col[4], col[5]);
        __syncwarp();
        // (16-byte reads of the 42 words; the root is body 0, its block starts on a 16-byte boundary; u_ = unused rows
        // 3..5 of the first three columns, the transpose of B)
        T A00, A10, A20, A01, A11, A21, A02, A12, A22;              // A[i][j] = column j, row i
        T B00, B10, B20, M00, M10, M20, B01, B11, B21, M01, M11, M21, B02, B12, B22, M02, M12, M22;
        T n0, n1, n2, f0, f1, f2, u0_, u1_, u2_;
        ld4(g, A00, A10, A20, u0_); ld4(g + 4, u1_, u2_, A01, A11); ld4(g + 8, A21, u0_, u1_, u2_);
        ld4(g + 12, A02, A12, A22, u0_); ld4(g + 16, u1_, u2_, B00, B10); ld4(g + 20, B20, M00, M10, M20);
        ld4(g + 24, B01, B11, B21, M01); ld4(g + 28, M11, M21, B02, B12); ld4(g + 32, B22, M02, M12, M22);
        ld4(g + 36, n0, n1, n2, f0); ld2(g + 40, f1, f2);
        // Minv = M^-1 (symmetric, cofactors)
        T c00 = M11 * M22 - M12 * M12, c01 = M02 * M12 - M01 * M22, c02 = M01 * M12 - M02 * M11;
        T c11 = M00 * M22 - M02 * M02, c12 = M01 * M02 - M00 * M12, c22 = M00 * M11 - M01 * M01;
        T idet = Num<T>::rcp(M00 * c00 + M01 * c01 + M02 * c02);
        const T m00 = c00 * idet, m01 = c01 * idet, m02 = c02 * idet, m11 = c11 * idet, m12 = c12 * idet, m22 = c22 * idet;
        // X = B Minv
        const T X00 = B00 * m00 + B01 * m01 + B02 * m02, X01 = B00 * m01 + B01 * m11 + B02 * m12, X02 = B00 * m02 + B01 * m12 + B02 * m22;
        const T X10 = B10 * m00 + B11 * m01 + B12 * m02, X11 = B10 * m01 + B11 * m11 + B12 * m12, X12 = B10 * m02 + B11 * m12 + B12 * m22;
        const T X20 = B20 * m00 + B21 * m01 + B22 * m02, X21 = B20 * m01 + B21 * m11 + B22 * m12, X22 = B20 * m02 + B21 * m12 + B22 * m22;
        // Schur complement S = A - X B^T (symmetric) and right-hand side  -(n - X f)
        const T S00 = A00 - (X00 * B00 + X01 * B01 + X02 * B02), S01 = A01 - (X00 * B10 + X01 * B11 + X02 * B12);
        const T S02 = A02 - (X00 * B20 + X01 * B21 + X02 * B22), S11 = A11 - (X10 * B10 + X11 * B11 + X12 * B12);
        const T S12 = A12 - (X10 * B20 + X11 * B21 + X12 * B22), S22 = A22 - (X20 * B20 + X21 * B21 + X22 * B22);
        // rotation axes e0 e1 e2 (columns of R), the rows of R^-1 as cross products / det, and the generalized
        // torques Q of the rotations (actuators of the torque models; phase E left them next to the motion vectors):
        // R^T (A a_w + B a_v + n) = Q, so the angular right-hand side is  R^-T Q - n
        const int4 fr = *reinterpret_cast<const int4*>(pr.aba_fr);       // dofs: tx | ty << 8 | tz << 16, r0 | r1 << 8 | r2 << 16
        const int dr0 = fr.y & 255, dr1 = (fr.y >> 8) & 255, dr2 = (fr.y >> 16) & 255;
        T e00, e10, e20, e01, e11, e21, e02, e12, e22, pad;
        ld4(K.S[dr0], e00, e10, e20, pad); ld4(K.S[dr1], e01, e11, e21, pad); ld4(K.S[dr2], e02, e12, e22, pad);
        const T Q0 = K.S[dr0][6], Q1 = K.S[dr1][6], Q2 = K.S[dr2][6];
        const T x0 = e11 * e22 - e21 * e12, x1 = e21 * e02 - e01 * e22, x2 = e01 * e12 - e11 * e02;      // e1 x e2
        const T y0 = e12 * e20 - e22 * e10, y1 = e22 * e00 - e02 * e20, y2 = e02 * e10 - e12 * e00;      // e2 x e0
        const T z0 = e10 * e21 - e20 * e11, z1 = e20 * e01 - e00 * e21, z2 = e00 * e11 - e10 * e01;      // e0 x e1
        const T id3 = Num<T>::rcp(e00 * x0 + e10 * x1 + e20 * x2);
        const T q0 = (Q0 * x0 + Q1 * y0 + Q2 * z0) * id3, q1 = (Q0 * x1 + Q1 * y1 + Q2 * z1) * id3;
        const T q2 = (Q0 * x2 + Q1 * y2 + Q2 * z2) * id3;
        const T r0 = q0 - (n0 - (X00 * f0 + X01 * f1 + X02 * f2)), r1 = q1 - (n1 - (X10 * f0 + X11 * f1 + X12 * f2));
        const T r2 = q2 - (n2 - (X20 * f0 + X21 * f1 + X22 * f2));
        c00 = S11 * S22 - S12 * S12; c01 = S02 * S12 - S01 * S22; c02 = S01 * S12 - S02 * S11;
        c11 = S00 * S22 - S02 * S02; c12 = S01 * S02 - S00 * S12; c22 = S00 * S11 - S01 * S01;
        idet = Num<T>::rcp(S00 * c00 + S01 * c01 + S02 * c02);
        a0 = (c00 * r0 + c01 * r1 + c02 * r2) * idet;
        a1 = (c01 * r0 + c11 * r1 + c12 * r2) * idet;
        a2 = (c02 * r0 + c12 * r1 + c22 * r2) * idet;
        // a_v = -Minv (f + B^T a_w)
        const T t0 = f0 + (B00 * a0 + B10 * a1 + B20 * a2), t1 = f1 + (B01 * a0 + B11 * a1 + B21 * a2);
        const T t2 = f2 + (B02 * a0 + B12 * a1 + B22 * a2);
        a3 = -(m00 * t0 + m01 * t1 + m02 * t2);
        a4 = -(m01 * t0 + m11 * t1 + m12 * t2);
        a5 = -(m02 * t0 + m12 * t1 + m22 * t2);
        // coordinates: translations read a_v off, rotations solve [e0 e1 e2] qdd = a_w
        if (lane == 0) {
            E.udot[fr.x & 255] = a3; E.udot[(fr.x >> 8) & 255] = a4; E.udot[(fr.x >> 16) & 255] = a5;
            E.udot[dr0] = (x0 * a0 + x1 * a1 + x2 * a2) * id3;
            E.udot[dr1] = (y0 * a0 + y1 * a1 + y2 * a2) * id3;
            E.udot[dr2] = (z0 * a0 + z1 * a1 + z2 * a2) * id3;
        }
    } else {
        for (int k = 0; k < nroot; k++) eliminate(pr.aba_root[k], true);
    }
    __syncwarp();
    auto back = [&](const int d, const bool store) {
        T w0, w1, w2, w3, w4, w5, w6, w7, s0, s1, s2, s3, s4, s5, s6, s7;
        ld4(X.W[d], w0, w1, w2, w3); ld4(X.W[d] + 4, w4, w5, w6, w7);
        ld4(K.S[d], s0, s1, s2, s3); ld4(K.S[d] + 4, s4, s5, s6, s7);
        const T qdd = -w6 - (w0 * a0 + w1 * a1 + w2 * a2 + w3 * a3 + w4 * a4 + w5 * a5);
        if (store) E.udot[d] = qdd;
        a0 += s0 * qdd; a1 += s1 * qdd; a2 += s2 * qdd; a3 += s3 * qdd; a4 += s4 * qdd; a5 += s5 * qdd;
    };
    if (!pr.aba_freeroot) for (int k = nroot - 1; k >= 0; k--) back(pr.aba_root[k], lane == 0);
    for (int s = nst - 1; s >= 0; s--) {
        const int code = pr.aba_step[grp][s];
        if (code != 255) back(code & 15, (lane & 15) == 0);
    }
    __syncwarp();
}

}  // namespace bio
