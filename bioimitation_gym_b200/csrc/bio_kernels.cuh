// bio_kernels.cuh -- the per-control-step kernels (one thread = one env).
//
// K1 bio_step_kernel : action pre-processing (env2D.py:115-131, torque PD
//     torque env2D.py:125-139) -> actuate/clip (opensim_wrapper.py:92-107) ->
//     S fixed substeps of the stated scheme (replaces Manager.integrate,
//     opensim_wrapper.py:299-301) -> end-of-step evaluation -> observation
//     (env2D.py:158-230) -> reward (:267-358) -> done (:237-265) -> in-kernel
//     auto-reset of finished envs (:133-156).
// K2 bio_reset_kernel, K3 bio_eval_kernel (debug / inverse-dynamics operator
//     set), plus the reset-table precompute and the state transposes.
#pragma once
#include "bio_dynamics.cuh"

namespace bio {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// counter-based draw keyed by (seed, global env index, episode, stream):
// results do not depend on how envs are sharded over GPUs.
__device__ __forceinline__ unsigned long long bio_rand(unsigned long long seed, unsigned long long env,
                                                       unsigned long long episode, unsigned long long stream) {
    unsigned long long h = splitmix64(seed ^ splitmix64(env));
    h = splitmix64(h + episode);
    return splitmix64(h ^ (stream * 0xD6E8FEB86659FD93ull));
}

template <typename T>
__device__ const DevModel<T>& stage_model(const DevModel<T>* gm, unsigned char* smem) {
    const int4* src = reinterpret_cast<const int4*>(gm);
    int4* dst = reinterpret_cast<int4*>(smem);
    constexpr int n16 = (int)(sizeof(DevModel<T>) / 16);
    static_assert(sizeof(DevModel<T>) % 16 == 0, "DevModel must be a multiple of 16 bytes");
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
    return *reinterpret_cast<const DevModel<T>*>(smem);
}

// perturbation force at time t (env2D.py:83-100): 100 knots over 10 s,
// piecewise constant, +-F where fmod(t_knot, 2) > thresh
template <typename T>
__device__ __noinline__ T perturb_force_on(T force, T thresh, int negative_only, unsigned long long seed,
                                           unsigned long long env, T t) {
    const double dtk = 10.0 / 99.0;
    int kidx = (int)ceil((double)t / dtk - 1e-12);
    kidx = kidx < 0 ? 0 : (kidx > 99 ? 99 : kidx);
    const double tk = kidx * dtk;
    if (!(fmod(tk, 2.0) > (double)thresh)) return T(0);
    if (negative_only) return -force;
    return (bio_rand(seed, env, (unsigned long long)kidx, 7) & 1ull) ? force : -force;
}
template <typename T>
__device__ __forceinline__ T perturb_force(const DevTask<T>& c, unsigned long long seed, unsigned long long env, T t) {
    // the fp64 knot arithmetic stays out of line (scalars only: the task block stays in the parameter space)
    return c.perturb ? perturb_force_on<T>(c.perturb_force, c.perturb_thresh, c.perturb_negative_only, seed, env, t) : T(0);
}

template <typename T>
struct Local {
    T q[BIO_MAX_DOF], u[BIO_MAX_DOF], act[BIO_MAX_MUSCLES], lm[BIO_MAX_MUSCLES];
};

template <typename T>
__device__ void clamp_muscle_state(const DevModel<T>& m, T* act, T* lm) {
    for (int i = 0; i < m.n_muscles; i++) {
        act[i] = clampv(act[i], m.mus_amin[i], T(1));
        if (lm[i] < m.mus_lm_min[i]) lm[i] = m.mus_lm_min[i];
    }
}

template <typename T, bool FULL>
__device__ void eval_env(const DevModel<T>& m, const DevTask<T>& c, const Local<T>& s, const T* ctrl, T t,
                         unsigned long long seed, unsigned long long env, EvalOut<T>& ev, T h_imp = T(0)) {
    const T fx = perturb_force(c, seed, env, t);
    eval_dynamics<T, FULL>(m, c.newton_iters, s.q, s.u, s.act, s.lm, ctrl, fx, c.perturb ? c.perturb_obspt : -1, h_imp,
                           ev, (const DebugRow<T>*)nullptr);
}

// dt of fixed-step integration, h = dt / n_substeps
template <typename T>
__device__ void integrate(const DevModel<T>& m, const DevTask<T>& c, Local<T>& s, const T* ctrl, int istep,
                          unsigned long long seed, unsigned long long env, EvalOut<T>& ev) {
    const int nd = m.n_dof, nm = m.n_muscles;
    const T h = c.dt / T(c.n_substeps);
    const T t0 = T(istep) * c.dt;
    for (int sub = 0; sub < c.n_substeps; sub++) {
        const T t = t0 + T(sub) * h;
        if (c.integrator == BIO_INT_SEMI_IMPLICIT_EULER || c.integrator == BIO_INT_IMPLICIT_DAMPING) {
            eval_env<T, false>(m, c, s, ctrl, t, seed, env, ev, c.integrator == BIO_INT_IMPLICIT_DAMPING ? h : T(0));
            for (int i = 0; i < nd; i++) { s.u[i] += h * ev.udot[i]; s.q[i] += h * s.u[i]; }
            for (int i = 0; i < nm; i++) { s.act[i] += h * ev.adot[i]; s.lm[i] += h * ev.lmdot[i]; }
        } else if (c.integrator == BIO_INT_RK2_MIDPOINT) {
            Local<T> mid;
            eval_env<T, false>(m, c, s, ctrl, t, seed, env, ev);
            const T hh = T(0.5) * h;
            for (int i = 0; i < nd; i++) { mid.q[i] = s.q[i] + hh * s.u[i]; mid.u[i] = s.u[i] + hh * ev.udot[i]; }
            for (int i = 0; i < nm; i++) { mid.act[i] = s.act[i] + hh * ev.adot[i]; mid.lm[i] = s.lm[i] + hh * ev.lmdot[i]; }
            clamp_muscle_state(m, mid.act, mid.lm);
            eval_env<T, false>(m, c, mid, ctrl, t + hh, seed, env, ev);
            for (int i = 0; i < nd; i++) { s.q[i] += h * mid.u[i]; s.u[i] += h * ev.udot[i]; }
            for (int i = 0; i < nm; i++) { s.act[i] += h * ev.adot[i]; s.lm[i] += h * ev.lmdot[i]; }
        } else {  // classic RK4
            Local<T> st, acc;
            for (int i = 0; i < nd; i++) { st.q[i] = s.q[i]; st.u[i] = s.u[i]; acc.q[i] = T(0); acc.u[i] = T(0); }
            for (int i = 0; i < nm; i++) { st.act[i] = s.act[i]; st.lm[i] = s.lm[i]; acc.act[i] = T(0); acc.lm[i] = T(0); }
            for (int r = 0; r < 4; r++) {
                const T wgt = (r == 0 || r == 3) ? T(1) : T(2);
                const T cn = r == 2 ? T(1) : T(0.5);  // node of the NEXT stage
                eval_env<T, false>(m, c, st, ctrl, t + (r == 0 ? T(0) : (r == 3 ? h : T(0.5) * h)), seed, env, ev);
                for (int i = 0; i < nd; i++) {
                    acc.q[i] += wgt * st.u[i]; acc.u[i] += wgt * ev.udot[i];
                }
                for (int i = 0; i < nm; i++) { acc.act[i] += wgt * ev.adot[i]; acc.lm[i] += wgt * ev.lmdot[i]; }
                if (r < 3) {
                    // next stage state uses THIS stage's derivative
                    T uq[BIO_MAX_DOF];
                    for (int i = 0; i < nd; i++) uq[i] = st.u[i];
                    for (int i = 0; i < nd; i++) { st.q[i] = s.q[i] + cn * h * uq[i]; st.u[i] = s.u[i] + cn * h * ev.udot[i]; }
                    for (int i = 0; i < nm; i++) { st.act[i] = s.act[i] + cn * h * ev.adot[i]; st.lm[i] = s.lm[i] + cn * h * ev.lmdot[i]; }
                    clamp_muscle_state(m, st.act, st.lm);
                }
            }
            const T h6 = h / T(6);
            for (int i = 0; i < nd; i++) { s.q[i] += h6 * acc.q[i]; s.u[i] += h6 * acc.u[i]; }
            for (int i = 0; i < nm; i++) { s.act[i] += h6 * acc.act[i]; s.lm[i] += h6 * acc.lm[i]; }
        }
        clamp_muscle_state(m, s.act, s.lm);
    }
}

template <typename T>
__device__ int ref_row(const DevTask<T>& c, int idx) {
    if (c.ref_mirror && c.cycle > 0 && idx > c.cycle) idx = 2 * c.cycle - idx;
    idx = idx < 0 ? 0 : idx;
    return idx > c.ref_rows - 1 ? c.ref_rows - 1 : idx;
}

// Observation row (env2D.py:158-230; SURVEY App. C layout)
template <typename T>
__device__ void write_obs(const DevModel<T>& m, const DevTask<T>& c, const Local<T>& s, const EvalOut<T>& ev,
                          int istep, T* obs) {
    int o = 0;
    const T ph = T(istep) / T(c.cycle);
    obs[o++] = ph - Num<T>::floor(ph);
    T pel[3] = {T(0), T(0), T(0)};
    for (int i = 0; i < m.n_coords; i++) {
        const int pt = m.coord_pelvis_trans[i];
        if (pt) pel[pt - 1] = s.q[m.coord_dof[i]];
    }
    for (int i = 0; i < m.n_coords; i++)
        if (!m.coord_pelvis_trans[i]) obs[o++] = m.coord_dof[i] >= 0 ? s.q[m.coord_dof[i]] : m.coord_const[i];
    for (int i = 0; i < m.n_coords; i++) obs[o++] = m.coord_dof[i] >= 0 ? s.u[m.coord_dof[i]] : T(0);
    for (int i = 0; i < m.n_coords; i++) obs[o++] = m.coord_dof[i] >= 0 ? ev.udot[m.coord_dof[i]] : T(0);
    if (c.use_target_obs) {
        const int row = ref_row(c, istep + 1);
        for (int i = 0; i < m.n_coords; i++)
            if (m.coord_pelvis_trans[i] != 1) obs[o++] = c.ref_q[(size_t)row * c.ref_coords + i];
        for (int i = 0; i < m.n_coords; i++)
            if (m.coord_pelvis_trans[i] != 1) obs[o++] = c.ref_u[(size_t)row * c.ref_coords + i];
    }
    if (!m.has_tz) pel[2] = T(0);
    for (int p = 0; p < c.n_obs_bodies; p++)
        for (int j = 0; j < 3; j++) obs[o++] = ev.obs_pos[p][j] - pel[j];
    for (int j = 0; j < 3; j++) obs[o++] = ev.com_pos[j] - pel[j];
    for (int p = 0; p < c.n_obs_body_vel; p++)
        for (int j = 0; j < 3; j++) obs[o++] = ev.obs_vel[p][j];
    for (int j = 0; j < 3; j++) obs[o++] = ev.com_vel[j];
    for (int i = 0; i < m.n_muscles; i++) { obs[o++] = s.act[i]; obs[o++] = s.lm[i]; obs[o++] = ev.lmdot[i]; }
    if (c.use_grf) {
        const T weight = Num<T>::abs(m.total_mass * m.gravity[1]);
        const T iw = T(1) / weight, im = T(1) / (weight * c.height);
        for (int g = 0; g < 2; g++) {
            for (int j = 0; j < 3; j++) obs[o++] = ev.contact[g][j] * iw;
            for (int j = 0; j < 3; j++) obs[o++] = ev.contact[g][3 + j] * im;
        }
    }
}

// Reference-state reset of env i (env2D.py:133-156): reference row ->
// coordinates and speeds, default activation, tabulated static fibre
// equilibrium.  Leaves old_px untouched (reference quirk).
template <typename T>
__device__ int reset_state(const DevModel<T>& m, const DevTask<T>& c, Local<T>& s, unsigned long long seed,
                           unsigned long long env, long long episode) {
    int idx = 0;
    if (!c.test_mode && c.reset_max_index > 0)
        idx = (int)(bio_rand(seed, env, (unsigned long long)episode, 1) % (unsigned long long)(c.reset_max_index + 1));
    idx = idx > c.ref_rows - 1 ? c.ref_rows - 1 : idx;
    for (int i = 0; i < m.n_coords; i++) {
        const int d = m.coord_dof[i];
        if (d < 0) continue;
        s.q[d] = c.ref_q[(size_t)idx * c.ref_coords + i];
        s.u[d] = c.ref_u[(size_t)idx * c.ref_coords + i];
    }
    for (int i = 0; i < m.n_muscles; i++) {
        s.act[i] = m.mus_default_act[i];
        s.lm[i] = c.ref_lm0[(size_t)idx * m.n_muscles + i];
    }
    return idx;
}

template <typename T>
__device__ void load_state(const DevModel<T>& m, const EnvState<T>& st, int i, int n, Local<T>& s) {
    for (int d = 0; d < m.n_dof; d++) { s.q[d] = st.q[sx(st.aos, d, i, m.n_dof, n)]; s.u[d] = st.u[sx(st.aos, d, i, m.n_dof, n)]; }
    for (int k = 0; k < m.n_muscles; k++) { s.act[k] = st.act[sx(st.aos, k, i, m.n_muscles, n)]; s.lm[k] = st.lm[sx(st.aos, k, i, m.n_muscles, n)]; }
}
template <typename T>
__device__ void store_state(const DevModel<T>& m, const EnvState<T>& st, int i, int n, const Local<T>& s) {
    for (int d = 0; d < m.n_dof; d++) { st.q[sx(st.aos, d, i, m.n_dof, n)] = s.q[d]; st.u[sx(st.aos, d, i, m.n_dof, n)] = s.u[d]; }
    for (int k = 0; k < m.n_muscles; k++) { st.act[sx(st.aos, k, i, m.n_muscles, n)] = s.act[k]; st.lm[sx(st.aos, k, i, m.n_muscles, n)] = s.lm[k]; }
}

template <typename T>
__device__ T body_mse(const T* cur, const T* des) {
    const T a = cur[0] - des[0], b = cur[1] - des[1], cc = cur[2] - des[2];
    return (a * a + b * b + cc * cc) / T(3);
}

template <typename T>
__global__ void bio_step_kernel(const DevModel<T>* __restrict__ gm, const DevTask<T> c, const EnvState<T> st, int n,
                                unsigned long long seed, long long env_offset, const T* __restrict__ actions,
                                T* __restrict__ obs, T* __restrict__ reward, uint8_t* __restrict__ done,
                                T* __restrict__ terms, double* __restrict__ stats) {
    extern __shared__ __align__(16) unsigned char smem[];
    const DevModel<T>& m = stage_model(gm, smem);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long env = (unsigned long long)(env_offset + i);
    const int na = m.n_act, nm = m.n_muscles, H = c.horizon;
    Local<T> s;
    load_state(m, st, i, n, s);
    int istep = st.istep[i];

    // ---- action pre-processing ----
    T action[BIO_MAX_ACT], curr[BIO_MAX_ACT], ctrl[BIO_MAX_ACT];
    bool nan = false;
    for (int j = 0; j < na; j++) { action[j] = actions[(size_t)i * na + j]; nan = nan || (action[j] != action[j]); }
    if (nan) {
        for (int j = 0; j < na; j++) action[j] = T(0);
        if (stats) atomicAdd(&stats[10], 1.0);
    } else if (c.use_pd) {
        T tau[BIO_MAX_ACT];
        for (int j = 0; j < c.n_pd; j++) {
            const int cx = c.pd_x_coord[j], cv = c.pd_v_coord[j];
            const T x = m.coord_dof[cx] >= 0 ? s.q[m.coord_dof[cx]] : m.coord_const[cx];
            const T v = m.coord_dof[cv] >= 0 ? s.u[m.coord_dof[cv]] : T(0);
            tau[j] = c.pd_kp[j] * (action[j] - x) + c.pd_kv[j] * (-v);
        }
        for (int j = 0; j < na; j++) action[j] = j < c.n_pd ? tau[j] : T(0);
    }
    int hist_pos = st.hist_pos[i];
    const bool first = st.first[i] != 0;
    T last_action[BIO_MAX_ACT];
    if (first) {
        for (int j = 0; j < na; j++) {
            last_action[j] = action[j];
            for (int hh = 0; hh < H; hh++) st.history[sx(st.aos, hh * na + j, i, H * na, n)] = action[j];
        }
        hist_pos = 0;
    } else {
        for (int j = 0; j < na; j++) last_action[j] = st.last_action[sx(st.aos, j, i, na, n)];
    }
    for (int j = 0; j < na; j++) st.history[sx(st.aos, hist_pos * na + j, i, H * na, n)] = action[j];
    hist_pos = (hist_pos + 1) % H;
    for (int j = 0; j < na; j++) {
        T sum = T(0);
        for (int hh = 0; hh < H; hh++) sum += st.history[sx(st.aos, hh * na + j, i, H * na, n)];
        curr[j] = sum / T(H);
        ctrl[j] = clampv(c.feed_mean_action ? curr[j] : action[j], m.act_min[j], m.act_max[j]);
    }

    // ---- integrate one control step, evaluate at the new state ----
    EvalOut<T> ev;
    integrate(m, c, s, ctrl, istep, seed, env, ev);
    istep += 1;
    eval_env<T, true>(m, c, s, ctrl, T(istep) * c.dt, seed, env, ev);
    T* orow = obs + (size_t)i * c.obs_dim;
    write_obs(m, c, s, ev, istep, orow);

    // ---- reward (env2D.py:267-358) ----
    const int row = ref_row(c, istep);
    T qerr = T(0), px = T(0), py = T(0);
    for (int k = 0; k < m.n_coords; k++) {
        const int d = m.coord_dof[k];
        const T v = d >= 0 ? s.q[d] : m.coord_const[k];
        const T dd = v - c.ref_q[(size_t)row * c.ref_coords + k];
        qerr += dd * dd;
        if (m.coord_pelvis_trans[k] == 1) px = v;
        if (m.coord_pelvis_trans[k] == 2) py = v;
    }
    qerr /= T(m.n_coords);
    const T com_err = body_mse(ev.com_pos, c.ref_com_pos + (size_t)row * 3);
    const T position_r = Num<T>::exp(T(-30) * qerr);
    const T com_r = Num<T>::exp(T(-20) * com_err);
    T foot[2];
    for (int sd = 0; sd < 2; sd++) {
        T sum = T(0);
        for (int j = 0; j < 4; j++)
            sum += body_mse(ev.obs_pos[c.rew_obspt[sd][j]],
                            c.ref_body_pos + ((size_t)row * c.ref_bodies + c.rew_refbody[sd][j]) * 3);
        foot[sd] = T(0.5) * Num<T>::exp(T(-20) * sum);
    }
    const T foot_r = foot[0], foot_l = foot[1];
    T effort, a_error = T(0);
    if (c.effort_torque) {
        T sum = T(0);
        for (int j = 0; j < na; j++) sum += curr[j] * curr[j];
        effort = Num<T>::sqrt(sum) / (c.max_actuation * T(na * na));
    } else {
        T sum = T(0), total = T(1.51) * m.total_mass;
        const T hp = T(1.5707963267948966);
        for (int k = 0; k < nm; k++) {
            sum += s.act[k] * s.act[k];
            const T l = m.mus_slow_twitch[k];
            T se, ce, sa, ca;
            Num<T>::sincos(hp * ctrl[k], &se, &ce);
            Num<T>::sincos(hp * s.act[k], &sa, &ca);
            const T fa = T(40) * l * se + T(133) * (T(1) - l) * (T(1) - ce);
            const T fm = T(74) * l * sa + T(111) * (T(1) - l) * (T(1) - ca);
            const T ln = s.lm[k] / m.mus_lopt[k], v = ev.lmdot[k];
            T g = T(0);
            if (ln < T(0.5)) g = T(0.5); else if (ln < T(1)) g = ln; else if (ln < T(1.5)) g = T(-2) * ln + T(3);
            const T es = T(0.25) * ev.fiber_force[k] * -v, ew = ev.active_fiber_force[k] * -v;
            total += m.mus_cot_mass[k] * fa + m.mus_cot_mass[k] * g * fm + (es > T(0) ? es : T(0)) + (ew > T(0) ? ew : T(0));
        }
        a_error = Num<T>::exp(T(-2) * Num<T>::sqrt(sum));
        effort = total / (T(20) * T(nm * nm));
    }
    const T old_px = st.old_px[i];
    const T progress_coord = c.effort_use_dy ? py : px;
    const T prog = progress_coord - old_px + T(1);
    const T effort_r = Num<T>::exp(-effort / (prog > T(1) ? prog : T(1)));
    T dn = T(0);
    for (int j = 0; j < na; j++) { const T d = curr[j] - last_action[j]; dn += d * d; }
    const T action_r = Num<T>::exp(-c.action_r_scale * Num<T>::sqrt(dn));
    T imit = position_r * com_r;
    if (c.reward_use_feet) imit *= (foot_l + foot_r);
    T rew = (T(0.5) + c.w_imitate) * imit + c.w_effort * effort_r + c.w_action * action_r;
    if (terms) {
        T* tr = terms + (size_t)i * c.n_reward_terms;
        tr[0] = position_r; tr[1] = com_r; tr[2] = foot_l; tr[3] = foot_r;
        if (c.n_reward_terms > 4) tr[4] = a_error;
    }
    // ---- termination (env2D.py:237-265) ----
    T maxacc = T(0);
    bool finite = isfinite(rew);
    for (int d = 0; d < m.n_dof; d++) {
        const T a = Num<T>::abs(ev.udot[d]);
        maxacc = a > maxacc ? a : maxacc;
        finite = finite && isfinite(s.q[d]) && isfinite(s.u[d]) && isfinite(ev.udot[d]);
    }
    int reason = 0;
    if (!finite) {
        reason = BIO_DONE_NONFINITE; rew = T(0);
        if (terms) for (int j = 0; j < c.n_reward_terms; j++) terms[(size_t)i * c.n_reward_terms + j] = T(0);
    }
    else if (ev.obs_pos[c.term_obspt][1] < c.term_height) reason = BIO_DONE_HEIGHT;
    else if (ev.max_limit > c.term_limit_force) reason = BIO_DONE_LIMIT_FORCE;
    else if (maxacc > c.term_acc) reason = BIO_DONE_ACCEL;
    else if (istep >= c.n_steps) reason = BIO_DONE_HORIZON;
    else if (c.term_feet_cross && ev.obs_pos[c.feet_obspt[0]][2] - ev.obs_pos[c.feet_obspt[1]][2] < T(0)) reason = BIO_DONE_FEET_CROSS;
    reward[i] = rew;
    done[i] = reason != 0;
    if (c.ex.any) {      // optional extra outputs (BioStepExtra), before a reset overwrites the state
        const StepExtra<T>& x = c.ex;
        if (x.done_reason) x.done_reason[i] = reason;
        if (x.terminal_obs && reason) for (int o = 0; o < c.obs_dim; o++) x.terminal_obs[(size_t)i * c.obs_dim + o] = orow[o];
        if (x.udot) for (int d = 0; d < m.n_dof; d++) x.udot[(size_t)i * m.n_dof + d] = ev.udot[d];
        if (x.fiber_force) for (int k = 0; k < nm; k++) x.fiber_force[(size_t)i * nm + k] = ev.fiber_force[k];
        if (x.fiber_vel) for (int k = 0; k < nm; k++) x.fiber_vel[(size_t)i * nm + k] = ev.lmdot[k];
        if (x.contact) for (int k = 0; k < 12; k++) x.contact[(size_t)i * 12 + k] = ev.contact[k / 6][k % 6];
        if (x.tendon_force || x.limit_force) {   // not kept by the evaluation: one more pass with the debug sinks
            DebugRow<T> row;
            row.path_len = row.path_vel = row.mass_matrix = row.bias = nullptr;
            row.tendon_force = x.tendon_force ? x.tendon_force + (size_t)i * nm : nullptr;
            row.limit_force = x.limit_force ? x.limit_force + (size_t)i * m.n_limits : nullptr;
            EvalOut<T> e2;
            eval_dynamics<T, true>(m, c.newton_iters, s.q, s.u, s.act, s.lm, ctrl, perturb_force(c, seed, env, T(istep) * c.dt),
                                   c.perturb ? c.perturb_obspt : -1, T(0), e2, &row);
        }
    }
    T ep_return = st.ep_return[i] + rew;
    int ep_len = st.ep_len[i] + 1;
    long long episode = st.episode[i];
    int first_next = 0;
    if (reason) {
        if (stats) {
            atomicAdd(&stats[1], 1.0);
            atomicAdd(&stats[2], (double)ep_return);
            atomicAdd(&stats[3], (double)ep_len);
            int bit = 0;
            while (!((reason >> bit) & 1)) bit++;
            atomicAdd(&stats[4 + bit], 1.0);
        }
        if (c.auto_reset) {
            episode += 1;
            istep = reset_state(m, c, s, seed, env, episode);
            first_next = 1;
            ep_return = T(0);
            ep_len = 0;
            for (int j = 0; j < na; j++) ctrl[j] = T(0);
            eval_env<T, true>(m, c, s, ctrl, T(istep) * c.dt, seed, env, ev);
            write_obs(m, c, s, ev, istep, orow);
        }
    }
    // ---- write back ----
    store_state(m, st, i, n, s);
    for (int j = 0; j < na; j++) st.last_action[sx(st.aos, j, i, na, n)] = first_next ? T(0) : curr[j];
    st.old_px[i] = progress_coord;
    st.istep[i] = istep;
    st.first[i] = first_next;
    st.hist_pos[i] = hist_pos;
    st.ep_return[i] = ep_return;
    st.ep_len[i] = ep_len;
    st.episode[i] = episode;
    if (stats && threadIdx.x == 0) {
        const int rem = n - blockIdx.x * blockDim.x;
        atomicAdd(&stats[0], (double)(rem < (int)blockDim.x ? rem : (int)blockDim.x));
    }
}

template <typename T>
__global__ void bio_reset_kernel(const DevModel<T>* __restrict__ gm, const DevTask<T> c, const EnvState<T> st, int n,
                                 unsigned long long seed, long long env_offset, const uint8_t* __restrict__ mask,
                                 T* __restrict__ obs, int bump_episode) {
    extern __shared__ __align__(16) unsigned char smem[];
    const DevModel<T>& m = stage_model(gm, smem);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (mask && !mask[i]) return;
    const unsigned long long env = (unsigned long long)(env_offset + i);
    Local<T> s;
    long long episode = st.episode[i] + (bump_episode ? 1 : 0);
    const int istep = reset_state(m, c, s, seed, env, episode);
    store_state(m, st, i, n, s);
    st.istep[i] = istep;
    st.first[i] = 1;
    st.hist_pos[i] = 0;
    st.ep_return[i] = T(0);
    st.ep_len[i] = 0;
    st.episode[i] = episode;
    for (int j = 0; j < m.n_act; j++) st.last_action[sx(st.aos, j, i, m.n_act, n)] = T(0);
    if (obs) {
        T ctrl[BIO_MAX_ACT];
        for (int j = 0; j < m.n_act; j++) ctrl[j] = T(0);
        EvalOut<T> ev;
        eval_env<T, true>(m, c, s, ctrl, T(istep) * c.dt, seed, env, ev);
        write_obs(m, c, s, ev, istep, obs + (size_t)i * c.obs_dim);
    }
}

// Static fibre equilibrium of every muscle at every reference row: makes the
// reset a table read instead of 4x equilibrateMuscles (opensim_wrapper.py:287-332).
template <typename T>
__global__ void bio_lm0_kernel(const DevModel<T>* __restrict__ gm, const T* __restrict__ ref_q, int rows,
                               int ref_coords, T* __restrict__ lm0) {
    extern __shared__ __align__(16) unsigned char smem[];
    const DevModel<T>& m = stage_model(gm, smem);
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    T q[BIO_MAX_DOF], L[BIO_MAX_MUSCLES];
    for (int i = 0; i < m.n_coords; i++)
        if (m.coord_dof[i] >= 0) q[m.coord_dof[i]] = ref_q[(size_t)r * ref_coords + i];
    path_lengths(m, q, L);
    for (int i = 0; i < m.n_muscles; i++) lm0[(size_t)r * m.n_muscles + i] = equilibrium_lm(m, i, L[i], m.mus_default_act[i]);
}

template <typename T>
struct DebugOut {
    T* udot; T* tendon_force; T* fiber_force; T* fiber_vel; T* act_dot; T* path_len; T* path_vel;
    T* contact; T* limit_force; T* mass_matrix; T* bias;
};

template <typename T>
__global__ void bio_eval_kernel(const DevModel<T>* __restrict__ gm, const DevTask<T> c, const EnvState<T> st, int n,
                                unsigned long long seed, long long env_offset, const T* __restrict__ controls,
                                const DebugOut<T> out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const DevModel<T>& m = stage_model(gm, smem);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Local<T> s;
    load_state(m, st, i, n, s);
    T ctrl[BIO_MAX_ACT];
    for (int j = 0; j < m.n_act; j++) ctrl[j] = controls ? controls[(size_t)i * m.n_act + j] : T(0);
    const int nd = m.n_dof, nm = m.n_muscles;
    DebugRow<T> row;
    row.tendon_force = out.tendon_force ? out.tendon_force + (size_t)i * nm : nullptr;
    row.path_len = out.path_len ? out.path_len + (size_t)i * nm : nullptr;
    row.path_vel = out.path_vel ? out.path_vel + (size_t)i * nm : nullptr;
    row.limit_force = out.limit_force ? out.limit_force + (size_t)i * m.n_limits : nullptr;
    row.mass_matrix = out.mass_matrix ? out.mass_matrix + (size_t)i * nd * nd : nullptr;
    row.bias = out.bias ? out.bias + (size_t)i * nd : nullptr;
    EvalOut<T> ev;
    const T t = T(st.istep[i]) * c.dt;
    const T fx = perturb_force(c, seed, (unsigned long long)(env_offset + i), t);
    eval_dynamics<T, true>(m, c.newton_iters, s.q, s.u, s.act, s.lm, ctrl, fx, c.perturb ? c.perturb_obspt : -1, T(0), ev,
                           &row);
    if (out.udot) for (int d = 0; d < nd; d++) out.udot[(size_t)i * nd + d] = ev.udot[d];
    if (out.fiber_force) for (int k = 0; k < nm; k++) out.fiber_force[(size_t)i * nm + k] = ev.fiber_force[k];
    if (out.fiber_vel) for (int k = 0; k < nm; k++) out.fiber_vel[(size_t)i * nm + k] = ev.lmdot[k];
    if (out.act_dot) for (int k = 0; k < nm; k++) out.act_dot[(size_t)i * nm + k] = ev.adot[k];
    if (out.contact) for (int k = 0; k < 12; k++) out.contact[(size_t)i * 12 + k] = ev.contact[k / 6][k % 6];
}

// Inverse-dynamics operator set (bio_id_apply; inverse_dynamics.cpp:65-197) at the current state of every env:
// one dynamics evaluation gives the joint-space inertia and the bias, the operators then run on the packed
// tree-sparse matrix with the factorisation of the step path (ltdl_factor / ltdl_solve).
template <typename T>
__global__ void bio_id_kernel(const DevModel<T>* __restrict__ gm, const DevTask<T> c, const EnvState<T> st, int n,
                              unsigned long long seed, long long env_offset, int op, const T* __restrict__ x,
                              const T* __restrict__ controls, const T* __restrict__ shift, T* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const DevModel<T>& m = stage_model(gm, smem);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nd = m.n_dof;
    Local<T> s;
    load_state(m, st, i, n, s);
    T ctrl[BIO_MAX_ACT];
    for (int j = 0; j < m.n_act; j++) ctrl[j] = controls ? controls[(size_t)i * m.n_act + j] : T(0);
    T Md[BIO_MAX_DOF * BIO_MAX_DOF], bias[BIO_MAX_DOF];
    DebugRow<T> row;
    row.tendon_force = row.path_len = row.path_vel = row.limit_force = nullptr;
    row.mass_matrix = Md;
    row.bias = bias;
    EvalOut<T> ev;
    const T fx = perturb_force(c, seed, (unsigned long long)(env_offset + i), T(st.istep[i]) * c.dt);
    eval_dynamics<T, false>(m, c.newton_iters, s.q, s.u, s.act, s.lm, ctrl, fx, c.perturb ? c.perturb_obspt : -1, T(0), ev,
                            &row);
    T H[BIO_MAX_DOF * (BIO_MAX_DOF + 1) / 2], v[BIO_MAX_DOF], y[BIO_MAX_DOF];
    for (int a = 0; a < nd; a++)
        for (int b = a; b >= 0; b = m.dof_parent[b]) H[a * (a + 1) / 2 + b] = Md[a * nd + b];
    for (int a = 0; a < nd; a++) v[a] = x[(size_t)i * nd + a];
    if (op == BIO_ID_MULTIPLY_M || op == BIO_ID_RESIDUAL) {
        tree_sym_matvec(m, H, v, y);
        if (op == BIO_ID_RESIDUAL) for (int a = 0; a < nd; a++) y[a] += bias[a];
    } else {
        if (op == BIO_ID_SOLVE_SHIFTED && shift)
            for (int a = 0; a < nd; a++) H[a * (a + 1) / 2 + a] += shift[(size_t)i * nd + a];
        ltdl_factor(m, H);
        ltdl_solve(m, H, v);
        for (int a = 0; a < nd; a++) y[a] = v[a];
    }
    for (int a = 0; a < nd; a++) out[(size_t)i * nd + a] = y[a];
}

// [N][k] row-major <-> SoA [k][N]
template <typename T>
__global__ void bio_transpose_kernel(const T* __restrict__ src, T* __restrict__ dst, int n, int k, int to_soa) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * k) return;
    const int i = (int)(idx / k), j = (int)(idx % k);
    if (to_soa) dst[(size_t)j * n + i] = src[idx];
    else dst[idx] = src[(size_t)j * n + i];
}

}  // namespace bio
