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
// go to the arrays of the general evaluation (K.R, K.r, K.V, K.A, K.S, E.O), so the other phases of
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
#pragma unroll
            for (int c = 0; c < 6; c++) K.S[d][c] = Sd[c];
        }
        if (i == o_step) { E.O[0] = o[0]; E.O[1] = o[1]; E.O[2] = o[2]; }
        if (code & P2_F_LAST) {
            const int b = (code >> 8) & 15;
            const T xx = qx * qx, yy = qy * qy, zz = qz * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz;
            const T wx = qw * qx, wy = qw * qy, wz = qw * qz;
            T* R = K.R[b];
            R[0] = T(1) - T(2) * (yy + zz); R[1] = T(2) * (xy - wz); R[2] = T(2) * (xz + wy);
            R[3] = T(2) * (xy + wz); R[4] = T(1) - T(2) * (xx + zz); R[5] = T(2) * (yz - wx);
            R[6] = T(2) * (xz - wy); R[7] = T(2) * (yz + wx); R[8] = T(1) - T(2) * (xx + yy);
#pragma unroll
            for (int c = 0; c < 3; c++) K.r[b][c] = r[c];
#pragma unroll
            for (int c = 0; c < 6; c++) { K.V[b][c] = V[c]; K.A[b][c] = A[c]; }
        }
    }
}

}  // namespace bio
