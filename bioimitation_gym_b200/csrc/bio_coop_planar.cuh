// bio_coop_planar.cuh -- joint-space solve shared by both evaluations, and the
// evaluation specialised for PLANAR models (every rotation about z, every
// translation in x/y: the reference's 2D gait models).
//
// For such a model the dynamics are exactly those of the (w_z, v_x, v_y) part of
// every spatial vector: moments about x / y and forces along z act on directions
// that have no degree of freedom.  Poses are (cos, sin, x, y) with a constant z
// per body (z still enters path lengths and the reported contact moments), spatial
// vectors have 3 components, the spatial inertia 4.  Same formulas as coop_eval
// otherwise; chosen at run time by DevModel::planar.
#pragma once

namespace bio {

// Sparse L^T D L of E.H along the tree with the forward substitution fused in,
// then the back substitution by tree depth.  In: E.H (tree-coupled entries),
// E.rhs.  Out: E.udot.
template <typename T, int CLS>
__device__ __forceinline__ void coop_solve(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane) {
    constexpr int G = CoopCls<CLS>::G;
    constexpr int NIT = CLS == 0 ? 1 : 2;       // pairs per step <= G * NIT (checked at create)
    const int nd = m.n_dof;
    for (int st = 0; st < nd; st++) {
        const int pb = m.lt_step_begin[st], pe = m.lt_step_begin[st + 1];
        if (pb == pe) continue;                  // root-most dof: nothing to eliminate
        const int k = nd - 1 - st;
        const T inv = T(1) / E.H[k * (k + 1) / 2 + k];
        const T bk = E.rhs[k];                   // final z_k: every descendant step is done
        T keep_a[NIT];
        int keep_ki[NIT];
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            keep_ki[it] = -1;
            const int p = pb + lane + it * G;
            if (p < pe) {
                const uint32_t pk = m.lt_pack[p];
                const int ij = pk & 255u, ki = (pk >> 8) & 255u, kj = (pk >> 16) & 255u;
                const T a = E.H[ki] * inv;
                E.H[ij] -= a * E.H[kj];
                if (pk & 0x80000000u) {          // diagonal pair (i,i): owns L_ki and the rhs update of i
                    keep_a[it] = a;
                    keep_ki[it] = ki;
                    E.rhs[(pk >> 24) & 15u] -= a * bk;
                }
            }
        }
        gsync<G>();
#pragma unroll
        for (int it = 0; it < NIT; it++)
            if (keep_ki[it] >= 0) E.H[keep_ki[it]] = keep_a[it];   // row k is not read by later steps
    }
    gsync<G>();
    // L x = D^-1 z by columns: lane i keeps w_i in a register; when x_j is final (all its
    // ancestors are < j) it is broadcast by shuffle and every descendant i subtracts L_ij x_j
    T wv = T(0);
    int row = 0;
    uint32_t anc = 0u;
    if (lane < nd) {
        row = lane * (lane + 1) / 2;
        wv = E.rhs[lane] / E.H[row + lane];
        anc = m.dof_anc_mask[lane];
    }
    for (int j = 0; j < nd - 1; j++) {
        const T xj = __shfl_sync(group_mask<G>(), wv, j, G);
        if (lane > j && ((anc >> j) & 1u)) wv -= E.H[row + j] * xj;
    }
    if (lane < nd) E.udot[lane] = wv;
    gsync<G>();
}

// planar helpers: R2 = [c -s; s c]
template <typename T> BIO_DEV void rot2(T c, T s, T x, T y, T& ox, T& oy) { ox = c * x - s * y; oy = s * x + c * y; }

template <typename T, int CLS>
__device__ __noinline__ void coop_eval_planar(const DevModel<T>& m, EnvWork<T, CLS>& E, const int lane,
                                              const int newton_iters, const T ext_fx, const int ext_pt, const T h_imp,
                                              const bool full) {
    typedef CoopCls<CLS> C;
    constexpr int G = C::G;
    const int nb = m.n_bodies, nd = m.n_dof, nm = m.n_muscles;
    // layout inside the general arrays:
    //   R[b] = (c, s)   r[b] = (x, y, z_const)   V[b] = (w, vx, vy)   A[b] likewise   S[d] = (w, vx, vy)
    //   BI[b] = (m, hx, hy, Izz, n, fx, fy)

    // ---- phase A: joint functions ----
    for (int a = lane; a < m.n_axes; a += G) {
        const int d = m.axis_dof[a];
        T s, ds, dds;
        func_eval(m, m.axis_func[a], d >= 0 ? E.q[d] : T(0), s, ds, dds);
        E.ax_s[a] = s; E.ax_ds[a] = ds; E.ax_dds[a] = dds;
    }
    gsync<G>();

    // ---- phase B: planar kinematics by tree level ----
    for (int lev = 0; lev < m.n_levels; lev++) {
        const int lb = m.level_begin[lev] + lane;
        if (lb < m.level_begin[lev + 1]) {
            const int b = m.level_body[lb], p = m.body_parent[b];
            T cp, sp, rx, ry, w, vx, vy, aw_, ax_, ay_;
            if (p >= 0) {
                cp = E.R[p][0]; sp = E.R[p][1];
                rot2(cp, sp, m.body_joint_loc[b][0], m.body_joint_loc[b][1], rx, ry);
                rx += E.r[p][0]; ry += E.r[p][1];
                w = E.V[p][0]; vx = E.V[p][1]; vy = E.V[p][2];
                aw_ = E.A[p][0]; ax_ = E.A[p][1]; ay_ = E.A[p][2];
            } else {
                cp = T(1); sp = T(0);
                rx = m.body_joint_loc[b][0]; ry = m.body_joint_loc[b][1];
                w = vx = vy = T(0);
                aw_ = T(0); ax_ = -m.gravity[0]; ay_ = -m.gravity[1];
            }
            T c = cp, s_ = sp;
            const int ab = m.body_axis_begin[b], ae = ab + m.body_axis_count[b];
            T Sw = T(0), Sx = T(0), Sy = T(0);
            for (int a = ab; a < ae; a++) {
                const int desc = m.axis_desc[a], d = (desc >> 3) & 31;
                const T s = E.ax_s[a];
                const T sg = (desc & 2) ? T(-1) : T(1);
                T kw, kx, ky;                     // this axis' motion vector
                if (!(desc & 1)) {                // translation along x or y of the parent frame
                    kw = T(0);
                    kx = sg * ((desc & 4) ? -sp : cp);
                    ky = sg * ((desc & 4) ? cp : sp);
                    rx += kx * s; ry += ky * s;
                } else {                          // rotation about +-z through the current origin
                    if (desc & 256) { E.O[0] = rx; E.O[1] = ry; E.O[2] = T(0); rx = ry = T(0); }
                    kw = sg; kx = sg * ry; ky = -sg * rx;
                    T sn, cs;
                    Num<T>::sincos(sg * s, &sn, &cs);
                    const T cn = c * cs - s_ * sn, snn = s_ * cs + c * sn;
                    c = cn; s_ = snn;
                }
                if (d != 31) {
                    const T ds = E.ax_ds[a], qd = E.u[d], sd = ds * qd, acc = E.ax_dds[a] * qd * qd;
                    if (desc & 512) Sw = Sx = Sy = T(0);
                    // V x S (planar): angular part 0, linear = w * (-S_vy, S_vx) + S_w * (V_vy, -V_vx)
                    const T cx = -w * ky + kw * vy, cy = w * kx - kw * vx;
                    Sw += ds * kw; Sx += ds * kx; Sy += ds * ky;
                    aw_ += kw * acc;
                    ax_ += kx * acc + cx * sd;
                    ay_ += ky * acc + cy * sd;
                    w += kw * sd; vx += kx * sd; vy += ky * sd;
                    if (desc & 1024) { E.S[d][0] = Sw; E.S[d][1] = Sx; E.S[d][2] = Sy; }
                }
                if (desc & 2048) { E.O[0] = rx; E.O[1] = ry; E.O[2] = T(0); rx = ry = T(0); }
            }
            E.R[b][0] = c; E.R[b][1] = s_;
            E.r[b][0] = rx; E.r[b][1] = ry; E.r[b][2] = m.body_z[b];
            E.V[b][0] = w; E.V[b][1] = vx; E.V[b][2] = vy;
            E.A[b][0] = aw_; E.A[b][1] = ax_; E.A[b][2] = ay_;
        }
        gsync<G>();
    }

    // ---- phase C: lane = muscle (path geometry is 3-D: points keep their constant z) ----
    if (lane < nm) {
        const int i = lane;
        // one streaming pass over the path points: positions, segment unit vectors and the
        // length; ptf[p] first holds the direction sum (e_out - e_in) and is scaled by the
        // tension once it is known (inactive points keep zero position and force)
        int pmov = -1, prev = -1;
        T mdloc[3] = {T(0), T(0), T(0)};
        T xp = T(0), yp = T(0), zp = T(0), ex = T(0), ey = T(0), ez = T(0), L = T(0);
        const int pb = m.mus_pt_begin[i], pe = pb + m.mus_pt_count[i];
        for (int p = pb; p < pe; p++) {
            const int kind = m.pt_kind[p], d = m.pt_dof[p], b = m.pt_body[p];
            T loc[3];
            if (kind == BIO_PT_CONDITIONAL) {
                const T v = E.q[d];
                if (!(v >= m.pt_range[p][0] - T(1e-5) && v <= m.pt_range[p][1] + T(1e-5))) {
                    for (int c = 0; c < 3; c++) { E.x.pt.ptf[p][c] = T(0); E.x.pt.ptx[p][c] = T(0); }
                    continue;
                }
            }
            if (kind == BIO_PT_MOVING) {
                T d2;
                for (int c = 0; c < 3; c++) func_eval(m, m.pt_func[p][c], E.q[d], loc[c], mdloc[c], d2);
                pmov = p;
            } else {
                for (int c = 0; c < 3; c++) loc[c] = m.pt_loc[p][c];
            }
            T x, y;
            rot2(E.R[b][0], E.R[b][1], loc[0], loc[1], x, y);
            x += E.r[b][0]; y += E.r[b][1];
            const T z = loc[2] + E.r[b][2];
            E.x.pt.ptx[p][0] = x; E.x.pt.ptx[p][1] = y; E.x.pt.ptx[p][2] = z;
            if (prev >= 0) {
                const T dx = x - xp, dy = y - yp, dz = z - zp;
                const T d2 = dx * dx + dy * dy + dz * dz;
                const T il = Num<T>::rsqrt(d2);
                L += d2 * il;
                const T nx = dx * il, ny = dy * il, nz = dz * il;
                E.x.pt.ptf[prev][0] = nx - ex; E.x.pt.ptf[prev][1] = ny - ey; E.x.pt.ptf[prev][2] = nz - ez;
                ex = nx; ey = ny; ez = nz;
            }
            xp = x; yp = y; zp = z; prev = p;
        }
        if (prev >= 0) { E.x.pt.ptf[prev][0] = -ex; E.x.pt.ptf[prev][1] = -ey; E.x.pt.ptf[prev][2] = -ez; }
        const T fiso = m.mus_fiso[i], lopt = m.mus_lopt[i], h = m.mus_height[i], beta = m.mus_beta[i];
        const T amin = m.mus_amin[i], lmin = m.mus_lm_min[i];
        const T lmi = E.lm[i];
        const T lmc = lmi < lmin ? lmin : lmi;
        const T lat = Num<T>::sqrt(lmc * lmc - h * h);
        const T cosa = lat / lmc;
        T fal, fpe, ft, fv, dfv, dtmp;
        curve_eval(m, 0, lmc / lopt, fal, dtmp);
        curve_eval(m, 2, lmc / lopt, fpe, dtmp);
        curve_eval(m, 3, (L - lat) / m.mus_lts[i], ft, dtmp);
        const T ac = clampv(E.act[i], amin, T(1));
        const T afal = ac * fal;
        // Warm start from the root of the previous evaluation of this step.  The residual is
        // monotone and sigmoid-shaped (convex for vn<0, concave for vn>0), so Newton is only
        // guaranteed from points between 0 and the root: anything else restarts from 0.
        T vn = E.vn[i];
        const T e0 = (afal + fpe) * cosa - ft;    // residual at vn = 0 (f_V(0) = 1)
        for (int it = 0; it < newton_iters; it++) {
            curve_eval(m, 1, vn, fv, dfv);
            const T err = (afal * fv + fpe + beta * vn) * cosa - ft;
            if (it == 0 && vn != T(0) && !(vn * e0 < T(0) && err * e0 > T(0))) { vn = T(0); continue; }
            const T derr = (afal * dfv + beta) * cosa;
            const T delta = -err / derr;
            vn += delta;
            if (Num<T>::abs(delta) < Num<T>::newton_tol()) break;
        }
        E.vn[i] = vn;
        if (lmi <= lmin && vn < T(0)) vn = T(0);
        E.lmdot[i] = vn * m.mus_vmax[i] * lopt;
        const T ec = clampv(E.ctrl[i], amin, T(1));
        const T tau = ec > ac ? m.mus_tact[i] * (T(0.5) + T(1.5) * ac) : m.mus_tdeact[i] / (T(0.5) + T(1.5) * ac);
        E.adot[i] = (ec - ac) / tau;
        const T tension = fiso * ft;
        if (full) {
            curve_eval(m, 1, vn, fv, dfv);
            E.fact[i] = fiso * afal * fv;
            E.ffib[i] = fiso * (afal * fv + fpe + beta * vn);
        }
        for (int p = pb; p < pe; p++)
            for (int c = 0; c < 3; c++) E.x.pt.ptf[p][c] *= tension;
        if (pmov >= 0) {
            const int b = m.pt_body[pmov];
            T dwx, dwy;
            rot2(E.R[b][0], E.R[b][1], mdloc[0], mdloc[1], dwx, dwy);
            E.x.pt.ptq[pmov] = E.x.pt.ptf[pmov][0] * dwx + E.x.pt.ptf[pmov][1] * dwy + E.x.pt.ptf[pmov][2] * mdloc[2];
        }
    }
    // ---- phase D: lane = contact sphere | coordinate limit ----
    if (lane < m.n_spheres) {
        const int s = lane, b = m.sph_body[s];
        T xc, yc;
        rot2(E.R[b][0], E.R[b][1], m.sph_loc[s][0], m.sph_loc[s][1], xc, yc);
        xc += E.r[b][0]; yc += E.r[b][1];
        const T zc = m.sph_loc[s][2] + E.r[b][2];
        const T rad = m.sph_radius[s];
        const T depth = rad - (yc + E.O[1]);
        T Fx = T(0), Fy = T(0), D0 = T(0), D1 = T(0);
        const T py = T(-0.5) * depth - E.O[1];
        if (depth > T(0)) {
            const T vx = E.V[b][1] - E.V[b][0] * py, vy = E.V[b][2] + E.V[b][0] * xc;
            const T vn = -vy;
            const T kk = m.sph_k[s];
            const T fH = T(4.0 / 3.0) * kk * depth * Num<T>::sqrt(rad * kk * depth);
            const T f = fH * (T(1) + T(1.5) * m.sph_c[s] * vn);
            if (f > T(0)) {
                Fy = f;
                const T vs = Num<T>::abs(vx);
                const T vrel = vs / m.sph_vt[s];
                const T strib = m.sph_ud[s] + T(2) * (m.sph_us[s] - m.sph_ud[s]) / (T(1) + vrel * vrel);
                if (vs != T(0)) {
                    const T ff = f * ((vrel < T(1) ? vrel : T(1)) * strib + m.sph_uv[s] * vs);
                    Fx = -ff * vx / vs;
                }
                D0 = f * ((vrel < T(1) ? T(1) / m.sph_vt[s] : T(1) / vs) * strib + m.sph_uv[s]);
                D1 = T(1.5) * m.sph_c[s] * fH;
            }
        }
        E.sphx[s][0] = xc; E.sphx[s][1] = py; E.sphx[s][2] = zc;
        E.sphF[s][0] = Fx; E.sphF[s][1] = Fy; E.sphF[s][2] = T(0);
        E.sphD[s][0] = D0; E.sphD[s][1] = D1;
    } else if (lane - m.n_spheres < m.n_limits) {
        const int l = lane - m.n_spheres, d = m.lim_dof[l];
        const T w = m.lim_w[l], qq = E.q[d];
        const T sup = step5((qq - m.lim_qup[l]) / w);
        const T slo = T(1) - step5((qq - (m.lim_qlo[l] - w)) / w);
        E.limf[l] = -m.lim_kup[l] * sup * (qq - m.lim_qup[l]) + m.lim_klo[l] * slo * (m.lim_qlo[l] - qq) -
                    m.lim_damp[l] * (sup + slo) * E.u[d];
        E.limD[l] = m.lim_damp[l] * (sup + slo);
    }
    gsync<G>();

    // ---- phase E: lane = body (planar wrench, inertia, body force) | dof (generalized forces) ----
    if (lane < nb) {
        const int b = lane;
        T Wn = T(0), Wx = T(0), Wy = T(0);
        for (int k = m.body_pt_begin[b]; k < m.body_pt_begin[b] + m.body_pt_count[b]; k++) {
            const int p = m.body_pt_list[k];
            const T fx = E.x.pt.ptf[p][0], fy = E.x.pt.ptf[p][1];   // inactive points: zero force, zero position
            Wn += E.x.pt.ptx[p][0] * fy - E.x.pt.ptx[p][1] * fx;
            Wx += fx; Wy += fy;
        }
        for (int s = 0; s < m.n_spheres; s++) {
            if (m.sph_body[s] != b || E.sphF[s][1] == T(0)) continue;
            Wn += E.sphx[s][0] * E.sphF[s][1] - E.sphx[s][1] * E.sphF[s][0];
            Wx += E.sphF[s][0]; Wy += E.sphF[s][1];
        }
        if (ext_pt >= 0 && m.obs_body[ext_pt] == b) {
            T x, y;
            rot2(E.R[b][0], E.R[b][1], m.obs_loc[ext_pt][0], m.obs_loc[ext_pt][1], x, y);
            y += E.r[b][1];
            Wn += -y * ext_fx;
            Wx += ext_fx;
        }
        T cx, cy;
        rot2(E.R[b][0], E.R[b][1], m.body_com[b][0], m.body_com[b][1], cx, cy);
        cx += E.r[b][0]; cy += E.r[b][1];
        const T mb = m.body_mass[b];
        const T Izz = m.body_inertia[b][2] + mb * (cx * cx + cy * cy);
        const T hx = mb * cx, hy = mb * cy;
        const T w = E.V[b][0], vx = E.V[b][1], vy = E.V[b][2];
        const T aw_ = E.A[b][0], ax_ = E.A[b][1], ay_ = E.A[b][2];
        // momentum (L; p) and I*A
        const T Lz = Izz * w + hx * vy - hy * vx, px = mb * vx - hy * w, py = mb * vy + hx * w;
        const T IAn = Izz * aw_ + hx * ay_ - hy * ax_, IAx = mb * ax_ - hy * aw_, IAy = mb * ay_ + hx * aw_;
        (void)Lz;
        E.BI[b][0] = mb; E.BI[b][1] = hx; E.BI[b][2] = hy; E.BI[b][3] = Izz;
        E.BI[b][4] = IAn + (vx * py - vy * px) - Wn;      // V x* (I V): n = v x p, f = w z x p
        E.BI[b][5] = IAx - w * py - Wx;
        E.BI[b][6] = IAy + w * px - Wy;
    } else if (lane - nb < nd) {
        const int d = lane - nb;
        T qf = T(0), ld = T(0);
        for (int l = 0; l < m.n_limits; l++) if (m.lim_dof[l] == d) { qf += E.limf[l]; ld += E.limD[l]; }
        E.limDd[d] = ld;
        for (int k = 0; k < m.n_moving; k++) { const int p = m.moving_pt[k]; if (m.pt_dof[p] == d) qf += E.x.pt.ptq[p]; }
        if (m.is_torque) for (int a = 0; a < m.n_act; a++) if (m.act_dof[a] == d) qf += E.ctrl[a];
        E.Q[d] = qf;
    }
    gsync<G>();

    // ---- full evaluation read-outs ----
    if (full) {
        if (lane < nb) {
            const int b = lane;
            T cx, cy;
            rot2(E.R[b][0], E.R[b][1], m.body_com[b][0], m.body_com[b][1], cx, cy);
            cx += E.r[b][0]; cy += E.r[b][1];
            const T mb = m.body_mass[b];
            E.x.out.comp[b][0] = mb * cx; E.x.out.comp[b][1] = mb * cy; E.x.out.comp[b][2] = mb * (m.body_com[b][2] + E.r[b][2]);
            E.x.out.comp[b][3] = mb * (E.V[b][1] - E.V[b][0] * cy);
            E.x.out.comp[b][4] = mb * (E.V[b][2] + E.V[b][0] * cx);
            E.x.out.comp[b][5] = T(0);
        } else if (lane - nb < m.n_obspts) {
            const int p = lane - nb, b = m.obs_body[p];
            T x, y;
            rot2(E.R[b][0], E.R[b][1], m.obs_loc[p][0], m.obs_loc[p][1], x, y);
            x += E.r[b][0]; y += E.r[b][1];
            E.x.out.obs_pos[p][0] = x + E.O[0]; E.x.out.obs_pos[p][1] = y + E.O[1];
            E.x.out.obs_pos[p][2] = m.obs_loc[p][2] + E.r[b][2];
            E.x.out.obs_vel[p][0] = E.V[b][1] - E.V[b][0] * y;
            E.x.out.obs_vel[p][1] = E.V[b][2] + E.V[b][0] * x;
            E.x.out.obs_vel[p][2] = T(0);
        }
        gsync<G>();
        if (lane < 3) {
            T ms = T(0), ps = T(0);
            for (int b = 0; b < nb; b++) { ms += E.x.out.comp[b][lane]; ps += E.x.out.comp[b][3 + lane]; }
            const T im = T(1) / m.total_mass;
            E.com_pos[lane] = ms * im + E.O[lane];
            E.com_vel[lane] = ps * im;
        } else if (lane < 5) {
            const int g = lane - 3;
            T w[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
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
            for (int l = 0; l < m.n_limits; l++) { const T a = Num<T>::abs(E.limf[l]); mx = a > mx ? a : mx; }
            E.max_limit = mx;
        }
    }

    // ---- phase F: composites; task = (body of the level, one of 7 values) ----
    for (int lev = m.n_levels - 2; lev >= 0; lev--) {
        const int cnt = (m.level_begin[lev + 1] - m.level_begin[lev]) * 8;
        for (int tsk = lane; tsk < cnt; tsk += G) {
            const int b = m.level_body[m.level_begin[lev] + (tsk >> 3)], v = tsk & 7;
            if (v == 7) continue;
            T acc = E.BI[b][v];
            for (int k = m.child_begin[b]; k < m.child_begin[b + 1]; k++) acc += E.BI[m.child_list[k]][v];
            E.BI[b][v] = acc;
        }
        gsync<G>();
    }

    // ---- phase G: I^c S per dof, contact Jacobian columns, entries ----
    if (lane < nd) {
        const int i = lane;
        const T* B = E.BI[m.dof_body[i]];
        const T sw = E.S[i][0], sx = E.S[i][1], sy = E.S[i][2];
        E.IS[i][0] = B[3] * sw + B[1] * sy - B[2] * sx;
        E.IS[i][1] = B[0] * sx - B[2] * sw;
        E.IS[i][2] = B[0] * sy + B[1] * sw;
        E.rhs[i] = E.Q[i] - (sw * B[4] + sx * B[5] + sy * B[6]);
    }
    unsigned act_mask = 0u;
    if (h_imp > T(0)) {
        for (int s = 0; s < m.n_spheres; s++) if (E.sphD[s][1] > T(0)) act_mask |= 1u << s;
        if (act_mask)
            for (int tsk = lane; tsk < m.jc_n; tsk += G) {      // (sphere, dof on its chain)
                const int s = m.jc_s[tsk], d = m.jc_d[tsk];
                if (!((act_mask >> s) & 1u)) continue;
                E.x.jac.col[s][d][0] = E.S[d][1] - E.S[d][0] * E.sphx[s][1];
                E.x.jac.col[s][d][1] = E.S[d][2] + E.S[d][0] * E.sphx[s][0];
            }
    }
    gsync<G>();
    for (int e = lane; e < m.n_entries; e += G) {
        const int i = m.ent_i[e], j = m.ent_j[e];
        T v = E.S[j][0] * E.IS[i][0] + E.S[j][1] * E.IS[i][1] + E.S[j][2] * E.IS[i][2];
        if (h_imp > T(0)) {
            unsigned mm = act_mask & m.ent_sph[e];               // active spheres whose chain holds i (and j)
            while (mm) {
                const int s = __ffs(mm) - 1;
                mm &= mm - 1u;
                v += h_imp * (E.sphD[s][0] * E.x.jac.col[s][i][0] * E.x.jac.col[s][j][0] +
                              E.sphD[s][1] * E.x.jac.col[s][i][1] * E.x.jac.col[s][j][1]);
            }
            if (i == j) v += h_imp * E.limDd[i];
        }
        E.H[i * (i + 1) / 2 + j] = v;
    }
    gsync<G>();

    // ---- phase H ----
    coop_solve<T, CLS>(m, E, lane);
}

}  // namespace bio
