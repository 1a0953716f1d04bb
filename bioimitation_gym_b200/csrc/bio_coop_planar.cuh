// bio_coop_planar.cuh -- joint-space solve of the general evaluation, and the
// evaluation of PLANAR models by the planar program (PlanarProg, bio_model.cuh).
//
// Planar model: every rotation about z, every translation in x/y (the reference's 2D
// gait models).  The dynamics are those of the (w_z, v_x, v_y) part of every spatial
// vector; the spatial inertia is a symmetric 3x3.  The tree is a root body with up to
// three dofs that carries chains of single-dof bodies, so one lane per chain can keep a
// whole chain in registers:
//
//   phase A  lane = elementary axis | moving-point function: joint functions, sin/cos
//   phase B  lane = chain: kinematics root -> leaf (the root joint is recomputed by every
//            chain lane instead of being exchanged)
//   phase C  lane = muscle: path geometry as crossing segments (a segment between two
//            points of one body does no work on the tree), tendon force, damped-equilibrium
//            fibre velocity (Newton, warm-started), activation ODE; the muscle leaves one
//            tension-scaled wrench per body it touches ("wrench source")
//   phase D  lane = contact sphere | coordinate limit (spheres are wrench sources too)
//   phase E  lane = body: gather of the wrench sources acting on the body, its spatial inertia and
//            force (contact damping h J^T D J enters as an inertia of the foot) | lane = dof:
//            generalized force of limits, moving points and actuators
//   phase F  lane = chain: articulated-body pass leaf -> root (the chain's L^T D L and its Schur
//            complement on the root dofs, one dof at a time)
//   phase G  lane = chain: root 3x3 solve (redundantly per lane), accelerations back down the chain
//
// Same equations as the general evaluation (coop_eval) and the CPU oracle; phases talk through
// shared memory only, so the host emulation in tests/emul can run them lane by lane.
#pragma once

namespace bio {

// Sparse L^T D L of K.H along the tree with the forward substitution fused in,
// then the back substitution by tree depth.  In: K.H (tree-coupled entries),
// K.rhs.  Out: E.udot.  (general evaluation)
template <typename T, int CLS>
__device__ __forceinline__ void coop_solve(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    constexpr int G = CoopCls<CLS>::G;
    constexpr int NIT = CLS == 0 ? 1 : 2;       // pairs per step <= G * NIT (checked at create)
    auto& K = E.k.g;
    const int nd = m.n_dof;
    for (int st = 0; st < nd; st++) {
        const int pb = m.lt_step_begin[st], pe = m.lt_step_begin[st + 1];
        if (pb == pe) continue;                  // root-most dof: nothing to eliminate
        const int k = nd - 1 - st;
        const T inv = Num<T>::rcp(K.H[k * (k + 1) / 2 + k]);
        const T bk = K.rhs[k];                   // final z_k: every descendant step is done
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            if (it > 0 && pe - pb <= it * G) break;      // (warp-uniform) no pair left for this round
            const int p = pb + lane + it * G;
            if (p < pe) {
                const uint32_t pk = m.lt_pack[p];
                const int ij = pk & 255u, ki = (pk >> 8) & 255u, kj = (pk >> 16) & 255u;
                const T a = K.H[ki] * inv;
                K.H[ij] -= a * K.H[kj];
                if (pk & 0x80000000u) {          // diagonal pair (i,i): owns L_ki and the rhs update of i
                    K.Lw[ki] = a;
                    K.rhs[(pk >> 24) & 15u] -= a * bk;
                }
            }
        }
        gsync<G>();
    }
    // L x = D^-1 z by columns: lane i keeps w_i in a register; when x_j is final (all its
    // ancestors are < j) it is broadcast by shuffle and every descendant i subtracts L_ij x_j
    T wv = T(0);
    int row = 0;
    uint32_t anc = 0u;
    if (lane < nd) {
        row = lane * (lane + 1) / 2;
        wv = Num<T>::div(K.rhs[lane], K.H[row + lane]);
        anc = m.dof_anc_mask[lane];
    }
    for (int j = 0; j < nd - 1; j++) {
        const T xj = __shfl_sync(group_mask<G>(), wv, j, G);
        if (lane > j && ((anc >> j) & 1u)) wv -= K.Lw[row + j] * xj;
    }
    if (lane < nd) E.udot[lane] = wv;
    gsync<G>();
}

// ---------------------------------------------------------------------------
// planar program
// ---------------------------------------------------------------------------
template <typename T> BIO_DEV void rot2(T c, T s, T x, T y, T& ox, T& oy) { ox = c * x - s * y; oy = s * x + c * y; }

// ---- phase A: joint functions of the coordinates (and the functions of moving path points) ----
template <typename T, int CLS>
BIO_DEV void p2_phase_a(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    constexpr int G = CoopCls<CLS>::G;
    const PlanarProg<T>& pr = m.prog;
    auto& K = E.k.p;
    for (int t = lane; t < pr.n_atasks; t += G) {
        // one packed descriptor per task (PlanarProg::at_i4 / at_f4): the loads below do not wait on each other
        const int4 ti = *reinterpret_cast<const int4*>(pr.at_i4[t]);
        T c0, c1, add, f3;
        ld4(pr.at_f4[t], c0, c1, add, f3);
        const int kind = ti.x & 3, dst = ti.x >> 8, d = ti.y;
        const T x = d >= 0 ? E.q[d] : T(0), qd = d >= 0 ? E.u[d] : T(0);
        T s, ds, dds = T(0);
        if (kind == BIO_FUNC_SPLINE) {
            // SimmSpline: cubic piece of the interval [knot i, knot i + 1), straight lines beyond the end knots;
            // the interval of the previous evaluation is the search hint (the coordinates move little)
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
            T sn = T(0), cs = T(1);
            if (ti.x & 4) Num<T>::sincos((ti.x & 8) ? -s : s, &sn, &cs);
            // displacement along the axis (translations only), rates
            st4(K.ax[dst], (ti.x & 4) ? T(0) : s, ds, ds * qd, dds * qd * qd);
            st2(K.axr[dst], cs, sn);
        } else {
            const int k = (dst - 64) / 3, c = (dst - 64) % 3;
            K.mv[k][c] = s + add;
            K.mv[k][4 + c] = ds;
        }
    }
}

// ---- phase B: lane l < n_branches walks the root joint and then its chain, one elementary axis
// per step (lane 0 publishes the root body; the other lanes only need its frame).  Every step is the
// same arithmetic: a translation has cos = 1, sin = 0 and a displacement, a rotation has no
// displacement, a constant axis has zero rates (exact no-ops) ----
template <typename T, int CLS>
BIO_DEV void p2_phase_b(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    const PlanarProg<T>& pr = m.prog;
    auto& K = E.k.p;
    const int nbr = pr.n_branches > 0 ? pr.n_branches : 1;
    if (lane >= nbr) return;
    T c = T(1), s = T(0), rx = T(0), ry = T(0);
    T w = T(0), vx = T(0), vy = T(0);
    T aw = T(0), ax = -m.gravity[0], ay = -m.gravity[1];
    T cp = T(1), sp = T(0);                      // frame of the parent body: translations run along its axes
    T Sw = T(0), Sx = T(0), Sy = T(0);
    const int n = pr.ch_n[lane];
#pragma unroll 2
    for (int i = 0; i < n; i++) {
        const int code = pr.ch_code[lane][i];
        const int a = code & 255;
        T jx, jy, rw, tA, tB, t3, st, ds, sd, acc, cs, sn;
        ld2(pr.ch_j[lane][i], jx, jy);
        ld4(pr.ax_k[a], rw, tA, tB, t3);
        ld4(K.ax[a], st, ds, sd, acc);
        ld2(K.axr[a], cs, sn);
        // first axis of a body: move to its joint location (zero otherwise)
        rx += c * jx - s * jy; ry += s * jx + c * jy;
        const bool first = (code & P2_F_FIRST) != 0, opre = (code & P2_F_OPRE) != 0, opost = (code & P2_F_OPOST) != 0;
        cp = first ? c : cp; sp = first ? s : sp;
        // the chain lanes repeat the root joint and store the same values
        if (opre) st4(E.O, rx, ry, T(0), T(0));
        rx = opre ? T(0) : rx; ry = opre ? T(0) : ry;
        // motion vector of the axis: translation along x / y of the parent frame, or rotation about
        // +-z through the current origin
        const T kx = tA * cp + tB * sp + rw * ry, ky = tA * sp - tB * cp - rw * rx;
        rx += kx * st; ry += ky * st;
        const T cn = c * cs - s * sn, snn = s * cs + c * sn;
        c = cn; s = snn;
        // V x S (planar): angular part 0, linear = w * (-S_vy, S_vx) + S_w * (V_vy, -V_vx)
        const T cx = -w * ky + rw * vy, cy = w * kx - rw * vx;
        const T keep = (code & P2_F_SRESET) ? T(0) : T(1);
        Sw = keep * Sw + ds * rw; Sx = keep * Sx + ds * kx; Sy = keep * Sy + ds * ky;
        aw += rw * acc;
        ax += kx * acc + cx * sd;
        ay += ky * acc + cy * sd;
        w += rw * sd; vx += kx * sd; vy += ky * sd;
        if (code & P2_F_SPUB) st4(K.S[(code >> 12) & 31], Sw, Sx, Sy, T(0));
        if (opost) st4(E.O, rx, ry, T(0), T(0));
        rx = opost ? T(0) : rx; ry = opost ? T(0) : ry;
        if (code & P2_F_LAST) {                  // last axis of its body: publish the frame
            const int b = (code >> 8) & 15;
            st4(K.pose[b], c, s, rx, ry);
            st4(K.V[b], w, vx, vy, T(0));
            st4(K.A[b], aw, ax, ay, T(0));
        }
    }
}

// ---- phase B as warp scans: lane = (chain, step), 8 steps per chain.  The walk above is a chain of
// associative updates, so every quantity is a prefix over the steps of a chain:
//   orientation   product of unit complex numbers (cos, sin)
//   position      sum of the displacements (joint offsets turned by the orientation before the step,
//                 translations along the parent frame), started at the step that fixes the origin O
//   velocity      sum of motion vector * rate;  bias acceleration: sum of motion vector * d2s/dq2 q'^2
//                 + (velocity before the step) x (motion vector) * rate
// 3 shuffle rounds per prefix instead of 8 dependent steps.  Same results as the serial walk up to the
// association order of the sums. ----
template <typename T>
__device__ __forceinline__ T shfl_up8(const unsigned mask, const T v, const int off) { return __shfl_up_sync(mask, v, off, 8); }
template <typename T>
__device__ __forceinline__ T shfl_at8(const unsigned mask, const T v, const int src) { return __shfl_sync(mask, v, src, 8); }

template <typename T, int CLS>
__device__ __forceinline__ void p2_phase_b_scan(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    constexpr int G = CoopCls<CLS>::G;
    const PlanarProg<T>& pr = m.prog;
    auto& K = E.k.p;
    const unsigned mask = group_mask<G>();
    const int l = lane >> 3, i = lane & 7;
    const bool live = l < pr.n_branches && i < pr.ch_n[l];
    int code = 0, sc = i;
    T jx = T(0), jy = T(0), rw = T(0), tA = T(0), tB = T(0), t3, st = T(0), ds = T(0), sd = T(0), acc = T(0);
    T c = T(1), s = T(0);
    if (l < pr.n_branches) sc = pr.ch_scan[l][i];
    if (live) {
        code = pr.ch_code[l][i];
        const int a = code & 255;
        ld2(pr.ch_j[l][i], jx, jy);
        ld4(pr.ax_k[a], rw, tA, tB, t3);
        ld4(K.ax[a], st, ds, sd, acc);
        ld2(K.axr[a], c, s);
    }
    const int first = sc & 15, o_step = (sc >> 4) & 15, in_dof = (sc >> 8) & 3;
    // orientation after every step (inclusive prefix product) and before it
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
        const T pc = shfl_up8(mask, c, off), ps = shfl_up8(mask, s, off);
        if (i >= off) { const T cn = c * pc - s * ps, sn = s * pc + c * ps; c = cn; s = sn; }
    }
    T ce = shfl_up8(mask, c, 1), se = shfl_up8(mask, s, 1);
    if (i == 0) { ce = T(1); se = T(0); }
    // frame of the parent body = orientation before the first step of this step's body
    const T cp = shfl_at8(mask, ce, first), sp = shfl_at8(mask, se, first);
    // displacement of the step; steps up to the one that fixes O build O, the later ones the position about O
    const T tx = tA * cp + tB * sp, ty = tA * sp - tB * cp;
    const T dx = ce * jx - se * jy + tx * st, dy = se * jx + ce * jy + ty * st;
    T rx = i > o_step ? dx : T(0), ry = i > o_step ? dy : T(0);
    T ox = i > o_step ? T(0) : dx, oy = i > o_step ? T(0) : dy;
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
        const T a0 = shfl_up8(mask, rx, off), a1 = shfl_up8(mask, ry, off);
        const T a2 = shfl_up8(mask, ox, off), a3 = shfl_up8(mask, oy, off);
        if (i >= off) { rx += a0; ry += a1; ox += a2; oy += a3; }
    }
    // motion vector (a rotation turns about the current origin; its own displacement is zero)
    const T kx = tx + rw * ry, ky = ty - rw * rx;
    // velocity after / before the step
    T w = rw * sd, vx = kx * sd, vy = ky * sd;
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
        const T a0 = shfl_up8(mask, w, off), a1 = shfl_up8(mask, vx, off), a2 = shfl_up8(mask, vy, off);
        if (i >= off) { w += a0; vx += a1; vy += a2; }
    }
    const T we = w - rw * sd, vxe = vx - kx * sd, vye = vy - ky * sd;
    // bias acceleration: V x S (planar): angular part 0, linear = w * (-S_vy, S_vx) + S_w * (V_vy, -V_vx)
    const T cx = -we * ky + rw * vye, cy = we * kx - rw * vxe;
    T aw = rw * acc, ax = kx * acc + cx * sd, ay = ky * acc + cy * sd;
    if (i == 0) { ax -= m.gravity[0]; ay -= m.gravity[1]; }
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
        const T a0 = shfl_up8(mask, aw, off), a1 = shfl_up8(mask, ax, off), a2 = shfl_up8(mask, ay, off);
        if (i >= off) { aw += a0; ax += a1; ay += a2; }
    }
    // motion vector of the dof: sum over its (<= 3, consecutive) axes
    T Sw = ds * rw, Sx = ds * kx, Sy = ds * ky;
    {
        const T p1w = shfl_up8(mask, Sw, 1), p1x = shfl_up8(mask, Sx, 1), p1y = shfl_up8(mask, Sy, 1);
        const T p2w = shfl_up8(mask, Sw, 2), p2x = shfl_up8(mask, Sx, 2), p2y = shfl_up8(mask, Sy, 2);
        if (in_dof >= 1) { Sw += p1w; Sx += p1x; Sy += p1y; }
        if (in_dof >= 2) { Sw += p2w; Sx += p2x; Sy += p2y; }
    }
    // publish (the chains repeat the root joint and store the same values)
    if (live) {
        if (code & P2_F_SPUB) st4(K.S[(code >> 12) & 31], Sw, Sx, Sy, T(0));
        if (i == o_step) st4(E.O, ox, oy, T(0), T(0));
        if (code & P2_F_LAST) {
            const int b = (code >> 8) & 15;
            st4(K.pose[b], c, s, rx, ry);
            st4(K.V[b], w, vx, vy, T(0));
            st4(K.A[b], aw, ax, ay, T(0));
        }
    }
}

// ---- phase C: lane = muscle ----
template <typename T, int CLS, bool FAST = false>
BIO_DEV void p2_phase_c(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane, const int newton_iters,
                        const T h_imp, const bool full) {
    const PlanarProg<T>& pr = m.prog;
    auto& K = E.k.p;
    if (lane >= m.n_muscles) return;
    const int i = lane;
    // path geometry: length, and the wrench of every segment that crosses from one body to another
    // (+w on the body of its first point, -w on the other); compiled form, else a streaming pass over the
    // active path points
    T W[P2_MAXSLOT][3];
#pragma unroll
    for (int sl = 0; sl < P2_MAXSLOT; sl++) W[sl][0] = W[sl][1] = W[sl][2] = T(0);
    T xp = T(0), yp = T(0), zp = T(0), ex = T(0), ey = T(0), ez = T(0), L = T(0);
    T mdx = T(0), mdy = T(0), mdz = T(0);        // moving point: d(location)/dq in ground axes
    T mqu = T(0);                                // and (e_out - e_in) . d(location)/dq
    int mov = -1;
    if (FAST || pr.path_ok) {
        // compiled path: variant from the muscle's conditional points, constant length of the variant, then
        // its live segments (same count on every lane: no divergence between muscles)
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
        for (int j = 0; j < pr.mc_nlive; j++) {
            const uint32_t seg = pr.mc_seg[i][var][j];
            if (!(seg >> 31)) continue;
            T pxy[2][3];
            int slot2[2];
            bool mv2[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int p = (seg >> (8 * e)) & 255u;
                const int info = pr.pt_info[p];
                const int b = info & 15;
                T c, s, ox, oy, lx, ly, lz, l3;
                ld4(K.pose[b], c, s, ox, oy);
                mv2[e] = ((info >> 4) & 3) == BIO_PT_MOVING;
                if (mv2[e]) {
                    T dlx, dly, dl3;
                    mov = (info >> 13) & 7;
                    ld4(K.mv[mov], lx, ly, lz, l3);
                    ld4(K.mv[mov] + 4, dlx, dly, mdz, dl3);
                    rot2(c, s, dlx, dly, mdx, mdy);
                } else {
                    ld4(pr.pt_xyz[p], lx, ly, lz, l3);
                }
                rot2(c, s, lx, ly, pxy[e][0], pxy[e][1]);
                pxy[e][0] += ox; pxy[e][1] += oy; pxy[e][2] = lz;
                slot2[e] = (info >> 11) & 3;
            }
            const T dx = pxy[1][0] - pxy[0][0], dy = pxy[1][1] - pxy[0][1], dz = pxy[1][2] - pxy[0][2];
            const T d2 = dx * dx + dy * dy + dz * dz;
            const T il = Num<T>::rsqrt(d2);
            L += d2 * il;
            ex = dx * il; ey = dy * il; ez = dz * il;
            const T msg = (mv2[0] ? T(1) : T(0)) - (mv2[1] ? T(1) : T(0));
            mqu += msg * (ex * mdx + ey * mdy + ez * mdz);
            if (slot2[0] != slot2[1]) {
                const T wn = pxy[0][0] * ey - pxy[0][1] * ex;
#pragma unroll
                for (int sl = 0; sl < P2_MAXSLOT; sl++) {
                    const T sgn = sl == slot2[0] ? T(1) : (sl == slot2[1] ? T(-1) : T(0));
                    W[sl][0] += sgn * wn; W[sl][1] += sgn * ex; W[sl][2] += sgn * ey;
                }
            }
        }
    } else {
    int prev_slot = -1;
    bool prev_moving = false;
    const int pb = m.mus_pt_begin[i], pe = pb + m.mus_pt_count[i];
    for (int p = pb; p < pe; p++) {
        const int info = pr.pt_info[p];
        const int b = info & 15, kind = (info >> 4) & 3, slot = (info >> 11) & 3;
        if (kind == BIO_PT_CONDITIONAL) {
            const T v = E.q[(info >> 6) & 31];
            if (!(v >= m.pt_range[p][0] - T(1e-5) && v <= m.pt_range[p][1] + T(1e-5))) continue;
        }
        T c, s, ox, oy, lx, ly, lz, l3;
        ld4(K.pose[b], c, s, ox, oy);
        if (kind == BIO_PT_MOVING) {
            T dlx, dly, dl3;
            mov = (info >> 13) & 7;
            ld4(K.mv[mov], lx, ly, lz, l3);
            ld4(K.mv[mov] + 4, dlx, dly, mdz, dl3);
            rot2(c, s, dlx, dly, mdx, mdy);
        } else {
            ld4(pr.pt_xyz[p], lx, ly, lz, l3);
        }
        T x, y;
        rot2(c, s, lx, ly, x, y);
        x += ox; y += oy;
        const T z = lz;
        if (prev_slot >= 0) {
            const T dx = x - xp, dy = y - yp, dz = z - zp;
            const T d2 = dx * dx + dy * dy + dz * dz;
            const T il = Num<T>::rsqrt(d2);
            L += d2 * il;
            ex = dx * il; ey = dy * il; ez = dz * il;
            // the segment pulls its first point along +e and its second along -e
            const T msg = (prev_moving ? T(1) : T(0)) - (kind == BIO_PT_MOVING ? T(1) : T(0));
            mqu += msg * (ex * mdx + ey * mdy + ez * mdz);
            if (slot != prev_slot) {
                const T wn = xp * ey - yp * ex;
#pragma unroll
                for (int sl = 0; sl < P2_MAXSLOT; sl++) {
                    const T sgn = sl == prev_slot ? T(1) : (sl == slot ? T(-1) : T(0));
                    W[sl][0] += sgn * wn; W[sl][1] += sgn * ex; W[sl][2] += sgn * ey;
                }
            }
        }
        prev_moving = kind == BIO_PT_MOVING;
        xp = x; yp = y; zp = z; prev_slot = slot;
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
    curve_eval(m, 3, (L - lat) * inv_lts, ft, dft);
    const T tension = fiso * ft;
    {   // wrench sources of this muscle: one per body it touches
        const int s0 = pr.mus_src0[i], ns = pr.mus_src0[i + 1] - s0;
#pragma unroll
        for (int sl = 0; sl < P2_MAXSLOT; sl++) {
            if (sl < ns) {
                st4(E.x.src.w[s0 + sl], tension * W[sl][0], tension * W[sl][1], tension * W[sl][2], T(0));
            }
        }
        // generalized force of the moving point: f . R_b dloc/dq
        if (mov >= 0) K.mq[mov] = tension * mqu;
    }
    const T lnorm = lmc * inv_lopt;
    curve_eval(m, 0, lnorm, fal, dfal);
    curve_eval(m, 2, lnorm, fpe, dfpe);
    const T ac = clampv(E.act[i], amin, T(1));
    const T afal = ac * fal;
    // Newton on the damped-equilibrium residual, warm-started from the root of the previous
    // evaluation of this step.  The residual is increasing in vn, convex for vn < 0 and concave for
    // vn > 0, so Newton converges monotonically from 0 and from any point between 0 and the root;
    // an iterate that would cross 0 is put on 0, which makes the iteration globally convergent.
    // ... From the third solve of a control step on the start is extrapolated from the last two roots (the state
    // moves 0.5 ms per substep): the first correction then is usually below the stop criterion.  A start beyond the
    // root lands between 0 and the root after one step (tangent of a concave / convex branch), so the
    // iteration stays globally convergent.
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
    // linearly implicit fibre-length update (fibre_gain): lmdot carries the factor 1 / (1 - h lambda) in the substep
    // evaluations (h_imp > 0); the full evaluation (h_imp = 0) reports the fibre velocity itself
    T gain = T(1);
    if (h_imp > T(0)) gain = fibre_gain(h_imp, vmax_lopt, derr, ac * dfal * fv + dfpe, inv_lopt, cosa, fsum, lmc, lat, h2, dft, inv_lts);
    E.lmdot[i] = vn * vmax_lopt * gain;
    // activation ODE: adot = (e - a) / tau, tau = tact (0.5 + 1.5 a) rising, tdeact / (0.5 + 1.5 a) falling
    const T ec = clampv(E.ctrl[i], amin, T(1));
    const T wa = T(0.5) + T(1.5) * ac;
    E.adot[i] = (ec - ac) * (ec > ac ? inv_tact * Num<T>::rcp(wa) : inv_tdeact * wa);
    if (full) {
        curve_eval(m, 1, vn, fv, dfv);
        E.fact[i] = fiso * afal * fv;
        E.ffib[i] = fiso * (afal * fv + fpe + beta * vn);
        E.vn[i] = tension;                        // read-out slot, see EnvWorkBody::vn
    }
}

// ---- phase D: lane = contact sphere | coordinate limit ----
template <typename T, int CLS>
BIO_DEV void p2_phase_d(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane, const T h_imp) {
    const PlanarProg<T>& pr = m.prog;
    auto& K = E.k.p;
    if (lane < m.n_spheres) {
        const int s = lane, b = m.sph_body[s];
        // per-sphere constants: three 16-byte reads (PlanarProg::sph_k)
        T lx, ly, zc, rad, kk, c15, ud, us2, uv, vt, inv_vt, k11;
        ld4(pr.sph_k[s], lx, ly, zc, rad);
        ld4(pr.sph_k[s] + 4, kk, c15, ud, us2);
        ld4(pr.sph_k[s] + 8, uv, vt, inv_vt, k11);
        T xc, yc, pc, ps, pox, poy;
        ld4(K.pose[b], pc, ps, pox, poy);
        rot2(pc, ps, lx, ly, xc, yc);
        xc += pox; yc += poy;
        const T depth = rad - (yc + E.O[1]);
        T Fx = T(0), Fy = T(0), D0 = T(0), D1 = T(0);
        const T py = T(-0.5) * depth - E.O[1];
        if (depth > T(0)) {
            T bw, bvx, bvy, b3;
            ld4(K.V[b], bw, bvx, bvy, b3);
            const T vx = bvx - bw * py, vy = bvy + bw * xc;
            const T vn = -vy;
            const T fH = T(4.0 / 3.0) * kk * depth * Num<T>::sqrt_pos(rad * kk * depth);
            const T f = fH * (T(1) + c15 * vn);
            if (f > T(0)) {
                Fy = f;
                const T vs = Num<T>::abs(vx);
                const T vrel = vs * inv_vt;
                const T strib = ud + Num<T>::div(us2, T(1) + vrel * vrel);
                // friction opposes the slip: -ff * vx / |vx|
                const T ff = f * ((vrel < T(1) ? vrel : T(1)) * strib + uv * vs);
                Fx = vx > T(0) ? -ff : (vx < T(0) ? ff : T(0));
                D0 = f * ((vrel < T(1) ? inv_vt : Num<T>::rcp(vs)) * strib + uv);
                D1 = c15 * fH;
            }
        }
        E.sphx[s][0] = xc; E.sphx[s][1] = py; E.sphx[s][2] = zc;
        E.sphF[s][0] = Fx; E.sphF[s][1] = Fy; E.sphF[s][2] = T(0);
        E.sphD[s][0] = D0; E.sphD[s][1] = D1;
        st4(E.x.src.w[pr.sph_src0 + s], xc * Fy - py * Fx, Fx, Fy, T(0));
        // implicit contact damping h J^T D J, J = [[-py, 1, 0], [xc, 0, 1]], as an inertia of the sphere's body
        const T d0 = h_imp * D0, d1 = h_imp * D1;
        st4(K.sphI[s], d0 * py * py + d1 * xc * xc, -d0 * py, d1 * xc, d0);
        K.sphI[s][4] = d1;
    } else if (lane - m.n_spheres < m.n_limits) {
        const int l = lane - m.n_spheres, d = m.lim_dof[l];
        // per-limit constants: two 16-byte reads (PlanarProg::lim_k)
        T qup, qlo, kup, klo, damp, inv_w, w, k7;
        ld4(pr.lim_k[l], qup, qlo, kup, klo);
        ld4(pr.lim_k[l] + 4, damp, inv_w, w, k7);
        const T qq = E.q[d];
        const T sup = step5((qq - qup) * inv_w);
        const T slo = T(1) - step5((qq - (qlo - w)) * inv_w);
        E.limf[l] = -kup * sup * (qq - qup) + klo * slo * (qlo - qq) - damp * (sup + slo) * E.u[d];
        E.limD[l] = damp * (sup + slo);
    }
}

// ---- phase E: lane b < n_bodies: gather of the wrench sources acting on body b, then its spatial
// inertia about O and force (inertial - applied); the contact damping of its active spheres,
// h J^T D J with J = [[-py, 1, 0], [px, 0, 1]], is part of the inertia.
// lane n_bodies + d: generalized force on dof d from limits, moving path points and actuators ----
template <typename T, int CLS, bool FAST = false>
BIO_DEV void p2_phase_e(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane, const T h_imp, const T ext_fx,
                        const int ext_pt) {
    const PlanarProg<T>& pr = m.prog;
    auto& K = E.k.p;
    if (lane < m.n_bodies) {
        const int b = lane;
        T Wn = T(0), Wx = T(0), Wy = T(0);
        if (FAST || pr.inc8_ok) {
            // source indices of the body in 8 bytes: the (<= 8) loads below are issued together
            const uint32_t p0 = pr.inc_pk[b][0], p1 = pr.inc_pk[b][1];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const unsigned e = ((k < 4 ? p0 : p1) >> (8 * (k & 3))) & 255u;
                if (e != 255u) {
                    T s0, s1, s2, s3;
                    ld4(E.x.src.w[e], s0, s1, s2, s3);
                    Wn += s0; Wx += s1; Wy += s2;
                }
            }
        } else {
#pragma unroll 4
            for (int k = pr.inc_begin[b]; k < pr.inc_begin[b + 1]; k++) {
                T s0, s1, s2, s3;
                ld4(E.x.src.w[pr.inc_src[k]], s0, s1, s2, s3);
                Wn += s0; Wx += s1; Wy += s2;
            }
        }
        T c, s, ox, oy, cx, cy, comx, comy, mb, izz;
        ld4(K.pose[b], c, s, ox, oy);
        ld4(pr.body_k[b], comx, comy, mb, izz);
        rot2(c, s, comx, comy, cx, cy);
        cx += ox; cy += oy;
        const T hx = mb * cx, hy = mb * cy;
        T Iww = izz + mb * (cx * cx + cy * cy), Iwx = -hy, Iwy = hx, Ixx = mb, Iyy = mb;
        T w, vx, vy, aw, ax, ay, pad3;
        ld4(K.V[b], w, vx, vy, pad3);
        ld4(K.A[b], aw, ax, ay, pad3);
        const T px = mb * vx - hy * w, py = mb * vy + hx * w;
        const T IAn = Iww * aw + hx * ay - hy * ax, IAx = mb * ax - hy * aw, IAy = mb * ay + hx * aw;
        if (ext_pt >= 0 && m.obs_body[ext_pt] == b) {
            T x, y;
            rot2(c, s, m.obs_loc[ext_pt][0], m.obs_loc[ext_pt][1], x, y);
            y += oy;
            Wn += -y * ext_fx;
            Wx += ext_fx;
        }
        T* o = K.bI[b];
        const T fn = IAn + (vx * py - vy * px) - Wn;   // V x* (I V): n = v x p, f = w z x p
        const T ffx = IAx - w * py - Wx;
        const T ffy = IAy + w * px - Wy;
        if (h_imp > T(0)) {
            if (FAST || pr.inc8_ok) {            // spheres of the body in 4 bytes (zero terms for a sphere out of contact)
                const uint32_t sp4 = pr.sph_pk[b];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned sp = (sp4 >> (8 * k)) & 255u;
                    if (sp != 255u) {
                        T a0, a1, a2, a3;
                        ld4(K.sphI[sp], a0, a1, a2, a3);
                        Iww += a0; Iwx += a1; Iwy += a2; Ixx += a3; Iyy += K.sphI[sp][4];
                    }
                }
            } else {
                int mask = pr.body_sph_mask[b];
                while (mask) {
                    const int sp = lowest_bit(mask);
                    mask &= mask - 1;
                    T a0, a1, a2, a3;
                    ld4(K.sphI[sp], a0, a1, a2, a3);
                    Iww += a0; Iwx += a1; Iwy += a2; Ixx += a3; Iyy += K.sphI[sp][4];
                }
            }
        }
        st4(o, Iww, Iwx, Iwy, T(0));
        st4(o + 4, Iwx, Ixx, T(0), T(0));
        st4(o + 8, Iwy, T(0), Iyy, T(0));
        st4(o + 12, fn, ffx, ffy, T(0));
    } else if (lane - m.n_bodies < m.n_dof) {
        const int d = lane - m.n_bodies;
        T qf = T(0), ld = T(0);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int l = pr.dof_lim[d][j];
            if (l >= 0) { qf += E.limf[l]; ld += E.limD[l]; }
            const int k = pr.dof_mov[d][j];
            if (k >= 0) qf += K.mq[k];
        }
        const int a = pr.dof_act[d];
        if (a >= 0) qf += E.ctrl[a];
        K.Qf[d] = qf;
        K.S[d][3] = qf;                           // next to the motion vector (p2_aba_coop: one read per step)
        K.Ld[d] = h_imp * ld;
    }
}

// ---- phase F: lane l < n_branches: articulated-body pass leaf -> root of the chain.  Every spatial
// quantity is expressed in ground axes about the common point O, so handing an inertia or a force to the
// parent is a plain sum.  Per body (from the leaf): I^A = I + I^a(child), p^A = p + p^a(child); with its dof
// (motion vector S, generalized force Q, implicit damping Ld on the diagonal):
//   U = I^A S,  D = S.U + Ld,  u = Q - S.p^A,  I^a = I^A - U U^T / D,  p^a = p^A + U u / D
// (the L^T D L elimination of the chain block and its Schur complement on the root, one dof at a time).
// The chain hands (I^a, p^a) of its first body to the root and keeps (U / D, u / D) for the way back. ----
template <typename T, int CLS>
BIO_DEV void p2_phase_f(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    const PlanarProg<T>& pr = m.prog;
    auto& K = E.k.p;
    if (lane >= pr.n_branches) return;
    const int l = lane;
    const int4 bb = *reinterpret_cast<const int4*>(pr.br_i8[l]), dd = *reinterpret_cast<const int4*>(pr.br_i8[l] + 4);
    const int bk[3] = {bb.x, bb.y, bb.z}, dk[3] = {dd.x, dd.y, dd.z};
    T Ia[6] = {T(0), T(0), T(0), T(0), T(0), T(0)}, pa[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int k = P2_MAXCB - 1; k >= 0; k--) {
        const int b = bk[k];
        const int d = dk[k];
        if (b >= 0) {
            T v0, v1, v2, v3, v4, v5, v6, v7, v8, pd;
            ld4(K.bI[b], v0, v1, v2, pd);
            ld4(K.bI[b] + 4, pd, v3, v4, pd);
            ld4(K.bI[b] + 8, pd, pd, v5, pd);
            ld4(K.bI[b] + 12, v6, v7, v8, pd);
            Ia[0] += v0; Ia[1] += v1; Ia[2] += v2; Ia[3] += v3; Ia[4] += v4; Ia[5] += v5;
            pa[0] += v6; pa[1] += v7; pa[2] += v8;
        }
        if (d >= 0) {
            T S0, S1, S2, s3;
            ld4(K.S[d], S0, S1, S2, s3);
            const T U0 = Ia[0] * S0 + Ia[1] * S1 + Ia[2] * S2;
            const T U1 = Ia[1] * S0 + Ia[3] * S1 + Ia[4] * S2;
            const T U2 = Ia[2] * S0 + Ia[4] * S1 + Ia[5] * S2;
            const T Dinv = Num<T>::rcp(S0 * U0 + S1 * U1 + S2 * U2 + K.Ld[d]);
            const T ud = (K.Qf[d] - (S0 * pa[0] + S1 * pa[1] + S2 * pa[2])) * Dinv;
            const T W0 = U0 * Dinv, W1 = U1 * Dinv, W2 = U2 * Dinv;
            Ia[0] -= U0 * W0; Ia[1] -= U0 * W1; Ia[2] -= U0 * W2;
            Ia[3] -= U1 * W1; Ia[4] -= U1 * W2; Ia[5] -= U2 * W2;
            pa[0] += U0 * ud; pa[1] += U1 * ud; pa[2] += U2 * ud;
            st4(E.x.pa.brk[l] + 4 * k, W0, W1, W2, ud);
        }
    }
    st4(E.x.pa.brx[l], Ia[0], Ia[1], Ia[2], Ia[3]);
    st4(E.x.pa.brx[l] + 4, Ia[4], Ia[5], pa[0], pa[1]);
    E.x.pa.brx[l][8] = pa[2];
}

// ---- phase G: root solve on the articulated inertia (every chain lane repeats it), then the chain's way
// back: qdd = u / D - (U / D) . a(parent),  a(body) = a(parent) + S qdd ----
template <typename T, int CLS>
BIO_DEV void p2_phase_g(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    const PlanarProg<T>& pr = m.prog;
    auto& K = E.k.p;
    const int nbr = pr.n_branches > 0 ? pr.n_branches : 1;
    if (lane >= nbr) return;
    const int4 ri = *reinterpret_cast<const int4*>(pr.root_i4);
    const int rdof[3] = {ri.x, ri.y, ri.z};
    T a[9], pd;
    ld4(K.bI[ri.w], a[0], a[1], a[2], pd);
    ld4(K.bI[ri.w] + 4, pd, a[3], a[4], pd);
    ld4(K.bI[ri.w] + 8, pd, pd, a[5], pd);
    ld4(K.bI[ri.w] + 12, a[6], a[7], a[8], pd);
#pragma unroll
    for (int l = 0; l < P2_MAXBR; l++) {
        if (l < pr.n_branches) {
            T v[8];
            ld4(E.x.pa.brx[l], v[0], v[1], v[2], v[3]);
            ld4(E.x.pa.brx[l] + 4, v[4], v[5], v[6], v[7]);
#pragma unroll
            for (int e = 0; e < 8; e++) a[e] += v[e];
            a[8] += E.x.pa.brx[l][8];
        }
    }
    T Sr[3][3], IS[3][3], H[3][3], rhs[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const bool has = rdof[r] >= 0;
        const int d = has ? rdof[r] : 0;
        T s3;
        ld4(K.S[d], Sr[r][0], Sr[r][1], Sr[r][2], s3);
        if (!has) Sr[r][0] = Sr[r][1] = Sr[r][2] = T(0);
        IS[r][0] = a[0] * Sr[r][0] + a[1] * Sr[r][1] + a[2] * Sr[r][2];
        IS[r][1] = a[1] * Sr[r][0] + a[3] * Sr[r][1] + a[4] * Sr[r][2];
        IS[r][2] = a[2] * Sr[r][0] + a[4] * Sr[r][1] + a[5] * Sr[r][2];
    }
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const bool has = rdof[r] >= 0;
        const int d = has ? rdof[r] : 0;
#pragma unroll
        for (int c = 0; c <= r; c++) H[r][c] = Sr[c][0] * IS[r][0] + Sr[c][1] * IS[r][1] + Sr[c][2] * IS[r][2];
        H[r][r] += has ? K.Ld[d] : T(1);         // no dof: identity row, zero right-hand side
        rhs[r] = (has ? K.Qf[d] : T(0)) - (Sr[r][0] * a[6] + Sr[r][1] * a[7] + Sr[r][2] * a[8]);
    }
    const T i0 = Num<T>::rcp(H[0][0]);
    const T l10 = H[1][0] * i0, l20 = H[2][0] * i0;
    const T i1 = Num<T>::rcp(H[1][1] - l10 * H[1][0]);
    const T t21 = H[2][1] - l20 * H[1][0];
    const T l21 = t21 * i1;
    const T i2 = Num<T>::rcp(H[2][2] - l20 * H[2][0] - l21 * t21);
    const T y1 = rhs[1] - l10 * rhs[0], y2 = rhs[2] - l20 * rhs[0] - l21 * y1;
    T ar[3];
    ar[2] = y2 * i2;
    ar[1] = y1 * i1 - l21 * ar[2];
    ar[0] = rhs[0] * i0 - l10 * ar[1] - l20 * ar[2];
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 3; r++) if (rdof[r] >= 0) E.udot[rdof[r]] = ar[r];
    }
    if (lane < pr.n_branches) {
        // acceleration of the parent body beyond its bias term
        T a0 = Sr[0][0] * ar[0] + Sr[1][0] * ar[1] + Sr[2][0] * ar[2];
        T a1 = Sr[0][1] * ar[0] + Sr[1][1] * ar[1] + Sr[2][1] * ar[2];
        T a2 = Sr[0][2] * ar[0] + Sr[1][2] * ar[1] + Sr[2][2] * ar[2];
        const int4 dd = *reinterpret_cast<const int4*>(pr.br_i8[lane] + 4);
        const int dk[3] = {dd.x, dd.y, dd.z};
#pragma unroll
        for (int k = 0; k < P2_MAXCB; k++) {
            const int d = dk[k];
            if (d >= 0) {
                T W0, W1, W2, ud, S0, S1, S2, s3;
                ld4(E.x.pa.brk[lane] + 4 * k, W0, W1, W2, ud);
                ld4(K.S[d], S0, S1, S2, s3);
                const T qdd = ud - (W0 * a0 + W1 * a1 + W2 * a2);
                E.udot[d] = qdd;
                a0 += S0 * qdd; a1 += S1 * qdd; a2 += S2 * qdd;
            }
        }
    }
}

// ---- phases F and G as one cooperative pass (device; p2_phase_f / _g are the same elimination with one lane per
// chain, kept for the host emulation and BIO_PLANAR_SERIAL_ABA=1).  Four lanes own a chain: lane c < 3 keeps column c
// of the chain's articulated inertia (symmetric 3 x 3) in registers, lane 3 the force as a fourth column.  Per dof,
// from the leaf:  U_c = S . column_c (lane 3: S . p^A);  the four U go round by shuffles;  D = S . U + Ld;
// column_c -= U (U_c - [c = 3] Q) / D  (I^a = I^A - U U^T / D,  p^a = p^A + U (Q - S . p^A) / D);  the coefficients
// (U / D, -(Q - S . p^A) / D) stay in shared memory for the way back.  The two chains swap their heads by a shuffle,
// every group adds the root body and eliminates the root's dofs, and every lane walks the accelerations back down. ----
template <typename T, int CLS, bool FAST = false>
__device__ __forceinline__ void p2_aba_coop(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    constexpr int G = CoopCls<CLS>::G;
    const PlanarProg<T>& pr = m.prog;
    auto& K = E.k.p;
    auto& X = E.x.pa;
    const unsigned mask = group_mask<G>();
    const int l = lane >> 2, c = lane & 3;
    const bool live = l < pr.n_branches;
    const int lc = live ? l : 0;
    const int4 bb = *reinterpret_cast<const int4*>(pr.br_i8[lc]), dd = *reinterpret_cast<const int4*>(pr.br_i8[lc] + 4);
    const int bk[3] = {bb.x, bb.y, bb.z}, dk[3] = {dd.x, dd.y, dd.z};
    const int4 ri = *reinterpret_cast<const int4*>(pr.root_i4);
    const int rdof[3] = {ri.x, ri.y, ri.z};
    T c0 = T(0), c1 = T(0), c2 = T(0);
    auto add_body = [&](const int b) {
        T v0, v1, v2, v3;
        ld4(K.bI[b] + 4 * c, v0, v1, v2, v3);
        c0 += v0; c1 += v1; c2 += v2;
    };
    auto eliminate = [&](const int d, const bool on) {
        T S0, S1, S2, Q;
        ld4(K.S[d], S0, S1, S2, Q);
        const T Ld = K.Ld[d];
        const T Uc = S0 * c0 + S1 * c1 + S2 * c2;
        const T U0 = __shfl_sync(mask, Uc, 0, 4), U1 = __shfl_sync(mask, Uc, 1, 4), U2 = __shfl_sync(mask, Uc, 2, 4);
        const T D = S0 * U0 + S1 * U1 + S2 * U2 + Ld;
        T coef = (Uc - (c == 3 ? Q : T(0))) * Num<T>::rcp(D);
        if (on) X.W[d][c] = coef; else coef = T(0);
        c0 -= U0 * coef; c1 -= U1 * coef; c2 -= U2 * coef;
    };
#pragma unroll
    for (int k = P2_MAXCB - 1; k >= 0; k--) {
        const int b = bk[k], d = dk[k];
        if (live && b >= 0) add_body(b);
        eliminate(d >= 0 ? d : 0, live && d >= 0);
    }
    if (!live) { c0 = T(0); c1 = T(0); c2 = T(0); }
    // chain heads: the two chain groups (lanes 0..3, 4..7) swap theirs, then everybody holds the root's columns
    {
        const T o0 = __shfl_xor_sync(mask, c0, 4, 8), o1 = __shfl_xor_sync(mask, c1, 4, 8), o2 = __shfl_xor_sync(mask, c2, 4, 8);
        c0 += o0; c1 += o1; c2 += o2;
        add_body(ri.w);
    }
    // way back: qdd = u / D - (U / D) . a(parent), a(body) = a(parent) + S qdd
    T a0 = T(0), a1 = T(0), a2 = T(0);
    if (FAST || pr.root_ident) {
        // Free planar root (translations along +x, +y of the ground, then a rotation about +z through O; no limit or
        // moving point on its dofs: PlanarProg::root_id4): its motion vectors are the unit vectors of (w, x, y), so the
        // root's acceleration solves  I^A a = Q - p^A  directly (one symmetric 3 x 3 inverse instead of three
        // elimination and three way-back steps) and the coordinates ARE its components.
        const int4 rid = *reinterpret_cast<const int4*>(pr.root_id4);
        const T I00 = __shfl_sync(mask, c0, 0, 4), I01 = __shfl_sync(mask, c1, 0, 4), I02 = __shfl_sync(mask, c2, 0, 4);
        const T I11 = __shfl_sync(mask, c1, 1, 4), I12 = __shfl_sync(mask, c2, 1, 4), I22 = __shfl_sync(mask, c2, 2, 4);
        const T p0 = __shfl_sync(mask, c0, 3, 4), p1 = __shfl_sync(mask, c1, 3, 4), p2 = __shfl_sync(mask, c2, 3, 4);
        const T r0 = K.S[rid.x][3] - p0, r1 = K.S[rid.y][3] - p1, r2 = K.S[rid.z][3] - p2;
        const T k00 = I11 * I22 - I12 * I12, k01 = I02 * I12 - I01 * I22, k02 = I01 * I12 - I02 * I11;
        const T k11 = I00 * I22 - I02 * I02, k12 = I01 * I02 - I00 * I12, k22 = I00 * I11 - I01 * I01;
        const T idet = Num<T>::rcp(I00 * k00 + I01 * k01 + I02 * k02);
        a0 = (k00 * r0 + k01 * r1 + k02 * r2) * idet;
        a1 = (k01 * r0 + k11 * r1 + k12 * r2) * idet;
        a2 = (k02 * r0 + k12 * r1 + k22 * r2) * idet;
        if (lane == 0) { E.udot[rid.x] = a0; E.udot[rid.y] = a1; E.udot[rid.z] = a2; }
    } else {
#pragma unroll
        for (int r = 2; r >= 0; r--) eliminate(rdof[r] >= 0 ? rdof[r] : 0, l == 0 && rdof[r] >= 0);
    }
    __syncwarp();
    auto back = [&](const int d, const bool store) {
        T W0, W1, W2, W3, S0, S1, S2, s3;
        ld4(X.W[d], W0, W1, W2, W3);
        ld4(K.S[d], S0, S1, S2, s3);
        const T qdd = -W3 - (W0 * a0 + W1 * a1 + W2 * a2);
        if (store) E.udot[d] = qdd;
        a0 += S0 * qdd; a1 += S1 * qdd; a2 += S2 * qdd;
    };
    if (!FAST && !pr.root_ident) {
#pragma unroll
        for (int r = 0; r < 3; r++) if (rdof[r] >= 0) back(rdof[r], lane == 0);
    }
    if (live) {
#pragma unroll
        for (int k = 0; k < P2_MAXCB; k++) if (dk[k] >= 0) back(dk[k], c == 0);
    }
}

// ---- full evaluation read-outs (two steps with a barrier in between) ----
template <typename T, int CLS>
BIO_DEV void p2_readout_1(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    auto& K = E.k.p;
    const int nb = m.n_bodies;
    if (lane < nb) {
        const int b = lane;
        T cx, cy;
        rot2(K.pose[b][0], K.pose[b][1], m.body_com[b][0], m.body_com[b][1], cx, cy);
        cx += K.pose[b][2]; cy += K.pose[b][3];
        const T mb = m.body_mass[b];
        E.x.out.comp[b][0] = mb * cx; E.x.out.comp[b][1] = mb * cy; E.x.out.comp[b][2] = mb * (m.body_com[b][2] + m.body_z[b]);
        E.x.out.comp[b][3] = mb * (K.V[b][1] - K.V[b][0] * cy);
        E.x.out.comp[b][4] = mb * (K.V[b][2] + K.V[b][0] * cx);
        E.x.out.comp[b][5] = T(0);
    } else if (lane - nb < m.n_obspts) {
        const int p = lane - nb, b = m.obs_body[p];
        T x, y;
        rot2(K.pose[b][0], K.pose[b][1], m.obs_loc[p][0], m.obs_loc[p][1], x, y);
        x += K.pose[b][2]; y += K.pose[b][3];
        E.x.out.obs_pos[p][0] = x + E.O[0]; E.x.out.obs_pos[p][1] = y + E.O[1];
        E.x.out.obs_pos[p][2] = m.obs_loc[p][2] + m.body_z[b];
        E.x.out.obs_vel[p][0] = K.V[b][1] - K.V[b][0] * y;
        E.x.out.obs_vel[p][1] = K.V[b][2] + K.V[b][0] * x;
        E.x.out.obs_vel[p][2] = T(0);
    }
}

template <typename T, int CLS>
BIO_DEV void p2_readout_2(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    const int nb = m.n_bodies;
    if (lane < 3) {
        T ms = T(0), ps = T(0);
#pragma unroll 1
        for (int b = 0; b < nb; b++) { ms += E.x.out.comp[b][lane]; ps += E.x.out.comp[b][3 + lane]; }
        const T im = T(1) / m.total_mass;
        E.com_pos[lane] = ms * im + E.O[lane];
        E.com_vel[lane] = ps * im;
    } else if (lane < 5) {
        const int g = lane - 3;
        T w[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
#pragma unroll 1
        for (int s = 0; s < m.n_spheres; s++) {
            if (m.sph_group[s] != g) continue;
            const T pa[3] = {E.sphx[s][0] + E.O[0], E.sphx[s][1] + E.O[1], E.sphx[s][2]};
            T n[3];
            cross3(pa, E.sphF[s], n);
            for (int c = 0; c < 3; c++) { w[c] += E.sphF[s][c]; w[3 + c] += n[c]; }
        }
        for (int c = 0; c < 6; c++) E.contact[g][c] = w[c];
    } else if (lane == 5) {
        T mx = T(0);
#pragma unroll 1
        for (int l = 0; l < m.n_limits; l++) { const T a = Num<T>::abs(E.limf[l]); mx = a > mx ? a : mx; }
        E.max_limit = mx;
    }
}

// -DBIO_PHASE_CLOCK (profiling build, tools/phase_clock.sh): thread 0 of CTA 0 accumulates the cycles of every
// phase of its env and prints them after each full evaluation
#ifdef BIO_PHASE_CLOCK
#include <stdio.h>
#define P2_CLK(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long t_ = clock64(); s_clk[k] += t_ - t_prev; t_prev = t_; } } while (0)
#else
#define P2_CLK(k) do { } while (0)
#endif

// FAST: the instantiation for models that take the scan kinematics, the compiled muscle paths, the packed source
// lists, the cooperative articulated-body pass and the direct root solve (every shipped 2D model on a half-warp);
// it holds no other path's code, so the text of the hot loop is what the loop executes (instruction cache).
template <typename T, int CLS, bool FAST = false>
__device__ __noinline__ void coop_eval_planar(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane,
                                              const int newton_iters, const T ext_fx, const int ext_pt, const T h_imp,
                                              const bool full) {
    constexpr int G = CoopCls<CLS>::G;
#ifdef BIO_PHASE_CLOCK
    __shared__ long long s_clk[12];
    __shared__ long long s_last;
    long long t_prev = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0 && s_last != 0) s_clk[9] += t_prev - s_last;   // between evaluations
#endif
    p2_phase_a<T, CLS>(m, E, lane);
    gsync<G>();
    P2_CLK(0);
    if (G == 16 && (FAST || m.prog.scan_ok)) p2_phase_b_scan<T, CLS>(m, E, lane);
    else p2_phase_b<T, CLS>(m, E, lane);
    gsync<G>();
    P2_CLK(1);
    p2_phase_c<T, CLS, FAST>(m, E, lane, newton_iters, h_imp, full);
    P2_CLK(2);
    p2_phase_d<T, CLS>(m, E, lane, h_imp);
    gsync<G>();
    P2_CLK(3);
    p2_phase_e<T, CLS, FAST>(m, E, lane, h_imp, ext_fx, ext_pt);
    gsync<G>();
    P2_CLK(4);
#ifdef __CUDA_ARCH__
    if (FAST || m.prog.coop_aba) {
        p2_aba_coop<T, CLS, FAST>(m, E, lane);
        P2_CLK(5);
    } else
#endif
    if constexpr (!FAST) {
        p2_phase_f<T, CLS>(m, E, lane);
        gsync<G>();
        P2_CLK(5);
        p2_phase_g<T, CLS>(m, E, lane);
    }
    if (full) {
        p2_readout_1<T, CLS>(m, E, lane);
        gsync<G>();
        p2_readout_2<T, CLS>(m, E, lane);
    }
    gsync<G>();
    P2_CLK(6);
#ifdef BIO_PHASE_CLOCK
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        s_clk[10] += 1;
        if (full) {
            printf("phase cycles over %lld evaluations: A %lld B %lld C %lld D %lld E %lld F %lld G(+readout) %lld between %lld\n",
                   s_clk[10], s_clk[0], s_clk[1], s_clk[2], s_clk[3], s_clk[4], s_clk[5], s_clk[6], s_clk[9]);
            for (int k = 0; k < 12; k++) s_clk[k] = 0;
            s_last = 0;
        } else {
            s_last = clock64();
        }
    }
#endif
}

}  // namespace bio
