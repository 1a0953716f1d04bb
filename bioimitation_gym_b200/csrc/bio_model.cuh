// bio_model.cuh -- device-side mirror of BioModelTables / BioTaskConfig in the
// kernel's scalar type, and the host conversion.  The model block is copied to
// shared memory once per CTA; all warps read it with uniform (broadcast)
// addresses.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <stdint.h>

#include <utility>

#include "../../include/bio_b200.h"

namespace bio {

// ---------------------------------------------------------------------------
// Planar program (bio_coop_planar.cuh): host-built schedule for models whose tree is a
// root body (<= 3 planar dofs) carrying chains of single-dof bodies (the reference's 2D gait
// models: pelvis+torso, and femur / tibia / foot per leg).  One lane walks one chain, so the
// kinematics, the composite inertias, the joint-space block of the chain and its elimination
// stay in that lane's registers; muscles hand their tension-scaled unit wrenches to the bodies
// through a list of wrench sources.
// ---------------------------------------------------------------------------
#define P2_MAXBR 2        // chains hanging from the root
#define P2_MAXSTEP 16     // elementary axes on the walk root joint -> leaf of one chain
#define P2_MAXCB 3        // bodies per chain
#define P2_MAXAX 3        // elementary axes per joint
#define P2_MAXSLOT 3      // distinct bodies one muscle touches
#define P2_MAXSRC 64      // wrench sources: (muscle, slot) pairs, then spheres
#define P2_MAXMOV 4       // moving path points
#define P2_MAXTASK (BIO_MAX_AXES + 3 * P2_MAXMOV)
#define P2_MAXVAR 4       // variants of a muscle path: on/off states of its (<= 2) conditional points
#define P2_MAXABA 12      // dofs of one chain (articulated-body schedule)
#define P2_MAXLIVE 2      // live segments of a variant (they cross bodies or touch a moving point)
#define P2_F_FIRST (1 << 17)     // first axis of its body
#define P2_F_LAST (1 << 18)      // last axis of its body: publish the body frame
#define P2_F_ROOT (1 << 19)      // axis of the root joint (recomputed by every chain lane, published by lane 0)
#define P2_F_SRESET (1 << 20)    // first axis of its dof
#define P2_F_SPUB (1 << 21)      // last axis of its dof: publish the motion vector
#define P2_F_OPRE (1 << 22)      // first rotation of the root: the origin O is fixed before this axis
#define P2_F_OPOST (1 << 23)     // root without rotation: O is fixed after this (last) axis
#define P2_F_ROT (1 << 24)       // rotation axis (else translation)

template <typename T>
struct alignas(16) PlanarProg {
    int32_t ok, n_branches, n_atasks, n_src;
    int32_t root_body, root_ndof, sph_src0, scan_ok;   // scan_ok: planar chain walk as warp scans (<= 8 steps per chain)
    int32_t chain_ok, pad_[3];                         // chain lists valid (any model: root + <= 2 chains of <= 16 steps)
    int32_t gch_nb[P2_MAXBR];                          // bodies of every chain, root side first (any chain_ok model)
    int32_t gch_body[P2_MAXBR][BIO_MAX_BODIES];
    int32_t root_dof[4];
    int32_t br_nb[P2_MAXBR];
    int32_t br_body[P2_MAXBR][P2_MAXCB];
    int32_t br_dof[P2_MAXBR][P2_MAXCB];        // dof of the body's joint (-1: none)
    int32_t body_sph_mask[BIO_MAX_BODIES];     // contact spheres carried by the body
    // chain walk: root axes, then the axes of the chain's bodies, one step per elementary axis:
    // code = axis | body<<8 | dof<<12 | flags (P2_F_*); ch_j = joint location on the first axis of a
    // body (else 0); ax_k = (rw, tA, tB): motion vector (rw, tA*cp + tB*sp + rw*ry, tA*sp - tB*cp - rw*rx)
    int32_t ch_n[P2_MAXBR];
    int32_t ch_scan[P2_MAXBR][P2_MAXSTEP];     // scan form: first step of the body | O step<<4 | axes before it in its dof<<8
    int32_t ch_code[P2_MAXBR][P2_MAXSTEP];
    alignas(16) T ch_j[P2_MAXBR][P2_MAXSTEP][4];
    alignas(16) T ax_k[BIO_MAX_AXES][4];
    // phase A tasks, splines first (the lanes of the last round then only see cheap functions):
    // at_dst < 64: elementary axis at_dst; else component (at_dst - 64) % 3 of moving point (at_dst - 64) / 3
    int32_t at_func[P2_MAXTASK], at_dof[P2_MAXTASK], at_dst[P2_MAXTASK];
    T at_add[P2_MAXTASK];                      // constant added to the value (body z of a moving point)
    // the same tasks packed for one 16-byte read each: at_i4 = kind | rotation << 2 | negated angle << 3 | dst << 8,
    // dof (-1: none), first knot, knot count;  at_f4 = Constant / LinearFunction coefficients c0, c1 (spline: first
    // and last knot abscissa), at_add, 0
    alignas(16) int32_t at_i4[P2_MAXTASK][4];
    alignas(16) T at_f4[P2_MAXTASK][4];
    // path points: location in the body frame with z already in ground axes (planar: constant)
    alignas(16) T pt_xyz[BIO_MAX_PATHPTS][4];
    int32_t pt_info[BIO_MAX_PATHPTS];          // body | kind<<4 | dof<<6 | slot<<11 | moving index<<13
    int32_t mus_src0[BIO_MAX_MUSCLES + 1];     // first wrench source of the muscle (one per slot); [n_muscles] = end
    int32_t mov_dof[P2_MAXMOV];
    // compiled paths: a segment between two fixed points of one body has a constant length and does no work on
    // the tree, so per variant (state of the muscle's conditional points) the path is a constant length plus
    // its live segments: mc_seg = first point | second point << 8 | 1 << 31
    int32_t path_ok, mc_nlive, a2_cheap, gpath_ok;  // a2_cheap: phase-A tasks past the first 16 are cheap; gpath_ok: general evaluation uses compiled paths     // path_ok: every muscle has this form; mc_nlive: most live segments of any variant
    int32_t mc_cond[BIO_MAX_MUSCLES][2];       // conditional points of the muscle (-1: none)
    uint32_t mc_seg[BIO_MAX_MUSCLES][P2_MAXVAR][P2_MAXLIVE];
    T mc_len0[BIO_MAX_MUSCLES][P2_MAXVAR];
    // per-lane constants of phases C and D packed for 16-byte reads, reciprocals taken on the host:
    //   mus_k[i]: fiso, lopt, 1/lopt, height^2 | beta, amin, lm_min, 1/lts | vmax*lopt, 1/tact, 1/tdeact, 0
    //   sph_k[s]: loc x, loc y, z in ground axes, radius | k, 1.5 c, ud, 2 (us - ud) | uv, vt, 1/vt, 0
    //   lim_k[l]: qup, qlo, kup, klo | damping, 1/w, w, 0
    // chain / root index lists packed for 16-byte reads: br_i8[l] = body of the chain's k-th body (-1: none) x 3, nb |
    // its dof (-1: none) x 3, 0;  root_i4 = the root's dofs (-1: none) x 3, root body
    alignas(16) T body_k[BIO_MAX_BODIES][4];   // centre of mass x, y (body frame), mass, inertia about z
    // wrench sources (8 bytes: up to 8 source indices, 255 = none) and contact spheres (4 bytes, 255 = none) of every
    // body, so that phase E issues all its loads at once; inc8_ok = 0: a body has more, walk inc_src / body_sph_mask
    alignas(8) uint32_t inc_pk[BIO_MAX_BODIES][2];
    uint32_t sph_pk[BIO_MAX_BODIES];
    int32_t inc8_ok, atask_ok, inc8_pad_[2];       // atask_ok: spatial evaluation runs phase A from the packed tasks
    alignas(16) int32_t br_i8[P2_MAXBR][8];
    alignas(16) int32_t root_i4[4];
    alignas(16) T mus_k[BIO_MAX_MUSCLES][12];
    alignas(16) T sph_k[BIO_MAX_SPHERES][12];
    alignas(16) T lim_k[BIO_MAX_LIMITS][8];
    // generalized-force inputs per dof
    int8_t dof_lim[BIO_MAX_DOF][2], dof_mov[BIO_MAX_DOF][2], dof_act[BIO_MAX_DOF];
    // wrench sources acting on every body
    int32_t inc_begin[BIO_MAX_BODIES + 1];
    uint8_t inc_src[3 * BIO_MAX_MUSCLES + BIO_MAX_SPHERES + 8];
    // articulated-body schedule of the spatial (3D) evaluation (p3_aba, bio_coop_spatial.cuh): per chain the dofs
    // in elimination order, leaf first: dof | (body whose inertia joins before this dof + 1) << 4, 255 = no step;
    // then the root's dofs, last first
    int32_t aba_ok, aba_nsteps, aba_nroot, coop_aba;   // coop_aba: planar program runs phases F, G as p2_aba_coop
    uint8_t aba_step[P2_MAXBR][P2_MAXABA];
    uint8_t aba_root[8];
    // free root joint (p3_aba): dofs of the translations along x, y, z (one byte each) | dofs of the three rotations
    // in axis order; aba_freeroot = the root is such a joint with no limit / moving point on its dofs (actuators on
    // the rotations are fine)
    alignas(16) int32_t aba_fr[4];
    int32_t aba_freeroot, root_ident, aba_fr_pad_[2];
    // free planar root (p2_aba_coop): dofs of the rotation about +z and of the translations along +x, +y; root_ident =
    // the root is such a joint (unit rates, no limit / moving point on its dofs)
    alignas(16) int32_t root_id4[4];
};

template <typename T>
struct alignas(16) DevModel {
    int32_t n_bodies, n_dof, n_axes, n_muscles, n_act, n_pathpts, n_spheres, n_limits;
    int32_t n_funcs, n_knots, n_obspts, n_coords, is_torque, has_tz, max_pts_per_muscle, pad;
    T gravity[3];
    T total_mass;
    int32_t body_parent[BIO_MAX_BODIES];
    int32_t body_axis_begin[BIO_MAX_BODIES];
    int32_t body_axis_count[BIO_MAX_BODIES];
    T body_mass[BIO_MAX_BODIES];
    T body_com[BIO_MAX_BODIES][3];
    T body_inertia[BIO_MAX_BODIES][6];
    T body_joint_loc[BIO_MAX_BODIES][3];
    int32_t axis_kind[BIO_MAX_AXES];
    int32_t axis_dof[BIO_MAX_AXES];
    int32_t axis_func[BIO_MAX_AXES];
    T axis_vec[BIO_MAX_AXES][3];
    int32_t dof_body[BIO_MAX_DOF];
    int32_t dof_parent[BIO_MAX_DOF];   // nearest ancestor dof on the root path (-1: none)
    int32_t body_last_dof[BIO_MAX_BODIES];  // deepest dof of the body's own joint
    uint32_t dof_anc_mask[BIO_MAX_DOF];
    int32_t func_kind[BIO_MAX_FUNCS];
    int32_t func_knot_begin[BIO_MAX_FUNCS];
    int32_t func_knot_count[BIO_MAX_FUNCS];
    T func_c[BIO_MAX_FUNCS][2];
    T knot_x[BIO_MAX_KNOTS];
    alignas(16) T knot_c[BIO_MAX_KNOTS][4];
    T mus_fiso[BIO_MAX_MUSCLES];
    T mus_lopt[BIO_MAX_MUSCLES];
    T mus_lts[BIO_MAX_MUSCLES];
    T mus_vmax[BIO_MAX_MUSCLES];
    T mus_tact[BIO_MAX_MUSCLES];
    T mus_tdeact[BIO_MAX_MUSCLES];
    T mus_amin[BIO_MAX_MUSCLES];
    T mus_beta[BIO_MAX_MUSCLES];
    T mus_default_act[BIO_MAX_MUSCLES];
    T mus_height[BIO_MAX_MUSCLES];
    T mus_lm_min[BIO_MAX_MUSCLES];
    T mus_cot_mass[BIO_MAX_MUSCLES];
    T mus_slow_twitch[BIO_MAX_MUSCLES];
    int32_t mus_pt_begin[BIO_MAX_MUSCLES];
    int32_t mus_pt_count[BIO_MAX_MUSCLES];
    int32_t pt_body[BIO_MAX_PATHPTS];
    int32_t pt_kind[BIO_MAX_PATHPTS];
    int32_t pt_dof[BIO_MAX_PATHPTS];
    int32_t pt_func[BIO_MAX_PATHPTS][3];
    T pt_loc[BIO_MAX_PATHPTS][3];
    T pt_range[BIO_MAX_PATHPTS][2];
    int32_t sph_body[BIO_MAX_SPHERES];
    int32_t sph_group[BIO_MAX_SPHERES];
    T sph_loc[BIO_MAX_SPHERES][3];
    T sph_radius[BIO_MAX_SPHERES];
    T sph_k[BIO_MAX_SPHERES];
    T sph_c[BIO_MAX_SPHERES];
    T sph_us[BIO_MAX_SPHERES];
    T sph_ud[BIO_MAX_SPHERES];
    T sph_uv[BIO_MAX_SPHERES];
    T sph_vt[BIO_MAX_SPHERES];
    int32_t lim_dof[BIO_MAX_LIMITS];
    T lim_kup[BIO_MAX_LIMITS];
    T lim_qup[BIO_MAX_LIMITS];
    T lim_klo[BIO_MAX_LIMITS];
    T lim_qlo[BIO_MAX_LIMITS];
    T lim_damp[BIO_MAX_LIMITS];
    T lim_w[BIO_MAX_LIMITS];
    int32_t act_dof[BIO_MAX_ACT];
    T act_min[BIO_MAX_ACT];
    T act_max[BIO_MAX_ACT];
    int32_t obs_body[BIO_MAX_OBSPTS];
    T obs_loc[BIO_MAX_OBSPTS][3];
    int32_t coord_dof[BIO_MAX_COORDS];
    int32_t coord_pelvis_trans[BIO_MAX_COORDS];
    int32_t pel_dof[4];                      // dof of pelvis_tx / ty / tz (-1: none), from coord_pelvis_trans
    // generalized-force inputs of every dof for the general evaluation: limits, moving path points (point
    // index) and the actuator acting on it (-1: none); gdof_ok = 0: more than two per dof, scan the lists instead
    int32_t gdof_ok, gdof_pad_[3];
    int8_t gdof_lim[BIO_MAX_DOF][2], gdof_movpt[BIO_MAX_DOF][2], gdof_act[BIO_MAX_DOF];
    T coord_const[BIO_MAX_COORDS];
    // schedules of the cooperative kernel (bio_coop.cuh), built by convert_model
    int32_t n_levels, n_moving, n_entries, pad2;
    int32_t body_level[BIO_MAX_BODIES];
    int32_t body_pt_begin[BIO_MAX_BODIES];
    int32_t body_pt_count[BIO_MAX_BODIES];
    int32_t body_pt_list[BIO_MAX_PATHPTS];
    int32_t moving_pt[8];
    int8_t pt_mov[BIO_MAX_PATHPTS];          // index of a moving path point in moving_pt (-1: not moving)
    int32_t ent_i[BIO_MAX_DOF * (BIO_MAX_DOF + 1) / 2];
    int32_t ent_j[BIO_MAX_DOF * (BIO_MAX_DOF + 1) / 2];
#define BIO_COOP_MAX_OBS_DIM 256            /* larger observation rows take the thread-per-env kernel */
    int32_t obs_desc[BIO_COOP_MAX_OBS_DIM];                   // (kind << 16) | pelvis selector << 12 | index per observation slot (build_obs_desc)
    T obs_cst[BIO_COOP_MAX_OBS_DIM];                          // constant of the slot: locked coordinate value (kind 14), contact scale (kind 13)
    // spline search: 16 uniform buckets per function -> first candidate knot
    T func_bucket_inv[BIO_MAX_FUNCS];
    int8_t func_bucket[BIO_MAX_FUNCS][16];
    // L^T D L schedule: step s handles k = n_dof-1-s; one lane per (i,j) pair of proper ancestors of k
    int32_t lt_step_begin[BIO_MAX_DOF + 1];
    // solve by tree depth: dofs of one depth are independent
    int32_t n_depths;
    int32_t dof_depth[BIO_MAX_DOF];
    // children of every body (composite inertia gather)
    int32_t child_begin[BIO_MAX_BODIES + 1];
    int32_t child_list[BIO_MAX_BODIES];
    int32_t level_begin[BIO_MAX_BODIES + 1];  // bodies grouped by tree level
    int32_t level_body[BIO_MAX_BODIES];
    int32_t axis_code[BIO_MAX_AXES];          // +-(k+1): axis is +-e_k, 0: general direction
    uint32_t lt_pack[320];                    // ij | ki<<8 | kj<<16 | i<<24 | diag<<31 per (i,j) pair
    int32_t lt_max_pairs;                     // largest number of pairs in one step
    int32_t planar;                           // 1: all rotations about z, translations in x/y (2D models)
    int32_t pad3;
    T body_z[BIO_MAX_BODIES];                 // planar models: constant z of every body origin
    // implicit damping: (sphere, dof on its chain) tasks and, per H entry, the spheres that touch it
    int32_t jc_n;
    int32_t pad4[3];
    uint8_t jc_s[64], jc_d[64];
    uint8_t ent_sph[BIO_MAX_DOF * (BIO_MAX_DOF + 1) / 2 + 8];
    // packed per-axis descriptor (planar kinematics): bit0 rotation, bit1 negative axis, bit2 translation
    // along y, bits3-7 dof (31: constant), bit8 first rotation of the root (fixes O), bit9 first axis of
    // its dof, bit10 last axis of its dof, bit11 root without rotation: fix O after this (last) axis
    int32_t axis_desc[BIO_MAX_AXES];
    // curves: uniform cubic Hermite, rows (y, h*dy/dx)
    T curve_x0[BIO_N_CURVES];
    T curve_inv_h[BIO_N_CURVES];
    T curve_x1[BIO_N_CURVES];
    // tabulated Millard curves on BIO_CURVE_N uniform intervals.  fp64: cubic Hermite pairs (value, slope * h) at
    // the knots.  fp32: the cubic of every interval in monomial form c0..c3 (s in [0, 1]), one 16-byte read per
    // evaluation instead of two 8-byte ones and no coefficient arithmetic in the kernel.
    T curve_tab[BIO_N_CURVES][sizeof(T) == 4 ? 1 : BIO_CURVE_N + 1][2];
    alignas(16) T curve_q[BIO_N_CURVES][sizeof(T) == 4 ? BIO_CURVE_N : 1][4];
    PlanarProg<T> prog;
};

// optional extra outputs of the step kernels (BioStepExtra), typed; `any` = some member is set
template <typename T>
struct StepExtra {
    T* terminal_obs; int32_t* done_reason; T* udot; T* tendon_force; T* fiber_force; T* fiber_vel; T* contact;
    T* limit_force;
    int32_t any, pad;
};

template <typename T>
struct DevTask {
    int32_t n_substeps, integrator, newton_iters, horizon, feed_mean_action, use_pd, n_pd;
    int32_t test_mode, cycle, n_steps, reset_max_index, ref_mirror, auto_reset;
    int32_t term_obspt, term_feet_cross, feet_obspt[2];
    int32_t reward_use_feet, effort_torque, effort_use_dy, n_reward_terms;
    int32_t rew_obspt[2][4], rew_refbody[2][4];
    int32_t use_target_obs, use_grf, n_obs_bodies, n_obs_body_vel, obs_dim;
    int32_t perturb, perturb_obspt, perturb_negative_only;
    int32_t pd_x_coord[BIO_MAX_ACT], pd_v_coord[BIO_MAX_ACT];
    T pd_kp[BIO_MAX_ACT], pd_kv[BIO_MAX_ACT];
    T dt, term_height, term_limit_force, term_acc;
    T h_sub, h_pad_;   // dt / n_substeps, divided on the host in T arithmetic (the quotient the kernels used to form)
    T w_imitate, w_effort, w_action, action_r_scale, max_actuation, height;
    T perturb_thresh, perturb_force;
    // reference tables (device)
    int32_t ref_rows, ref_coords, ref_bodies, pad;
    const T* ref_q;
    const T* ref_u;
    const T* ref_body_pos;
    const T* ref_com_pos;
    const T* ref_lm0;     // [rows][n_muscles] static fibre equilibrium per reference row
    StepExtra<T> ex;      // all null unless bio_set_step_extra attached buffers
};

// Per-env state.  Every array holds K variables per env (q, u: n_dof; act, lm: n_muscles; last_action: n_act;
// history: horizon * n_act) in one of two layouts, picked at create time by the kernel that steps the handle:
//   aos = 1  [n_envs][K]  cooperative kernels: the lanes of an env's group read / write consecutive words
//   aos = 0  [K][n_envs]  thread-per-env kernels: the lanes of a warp (consecutive envs) read consecutive words
// sx() is the element index of variable v of env i.
template <typename T>
struct EnvState {
    int32_t aos, pad_;
    T* q;
    T* u;
    T* act;
    T* lm;
    T* last_action;
    T* history;      // [horizon][n_act][N]
    T* old_px;
    T* ep_return;
    int32_t* istep;
    int32_t* first;
    int32_t* hist_pos;
    int32_t* ep_len;
    long long* episode;
};

__host__ __device__ __forceinline__ size_t sx(const int aos, const int v, const int i, const int K, const int n) {
    return aos ? (size_t)i * K + v : (size_t)v * n + i;
}

#define BIO_CP(field)                                                                  \
    do {                                                                               \
        const size_t n_ = sizeof(d.field) / sizeof(T);                                 \
        static_assert(sizeof(s.field) / sizeof(double) == sizeof(d.field) / sizeof(T), \
                      "field size mismatch: " #field);                                 \
        const double* sp_ = (const double*)&s.field;                                   \
        T* dp_ = (T*)&d.field;                                                         \
        for (size_t i_ = 0; i_ < n_; i_++) dp_[i_] = (T)sp_[i_];                       \
    } while (0)
#define BIO_CPI(field)                                                         \
    do {                                                                       \
        static_assert(sizeof(s.field) == sizeof(d.field), "int field " #field); \
        memcpy(&d.field, &s.field, sizeof(d.field));                           \
    } while (0)

template <typename T> void build_planar_prog(const BioModelTables& s, DevModel<T>& d);

template <typename T>
void convert_model(const BioModelTables& s, DevModel<T>& d) {
    memset(&d, 0, sizeof(d));
    d.n_bodies = s.n_bodies; d.n_dof = s.n_dof; d.n_axes = s.n_axes; d.n_muscles = s.n_muscles;
    d.n_act = s.n_act; d.n_pathpts = s.n_pathpts; d.n_spheres = s.n_spheres; d.n_limits = s.n_limits;
    d.n_funcs = s.n_funcs; d.n_knots = s.n_knots; d.n_obspts = s.n_obspts; d.n_coords = s.n_coords;
    d.is_torque = s.is_torque;
    d.has_tz = 0;
    for (int i = 0; i < s.n_coords; i++) if (s.coord_pelvis_trans[i] == 3) d.has_tz = 1;
    d.gdof_ok = 1;
    for (int dd = 0; dd < BIO_MAX_DOF; dd++) { d.gdof_lim[dd][0] = d.gdof_lim[dd][1] = d.gdof_movpt[dd][0] = d.gdof_movpt[dd][1] = d.gdof_act[dd] = -1; }
    auto gput = [&](int8_t* two, int v) { if (two[0] < 0) two[0] = (int8_t)v; else if (two[1] < 0) two[1] = (int8_t)v; else d.gdof_ok = 0; };
    for (int l = 0; l < s.n_limits; l++) gput(d.gdof_lim[s.lim_dof[l]], l);
    for (int p = 0; p < s.n_pathpts && p < 128; p++) if (s.pt_kind[p] == BIO_PT_MOVING && s.pt_dof[p] >= 0) gput(d.gdof_movpt[s.pt_dof[p]], p);
    if (s.is_torque)
        for (int a = 0; a < s.n_act; a++) {
            if (s.act_dof[a] < 0) continue;
            if (d.gdof_act[s.act_dof[a]] >= 0) d.gdof_ok = 0;
            d.gdof_act[s.act_dof[a]] = (int8_t)a;
        }
    for (int k = 0; k < 4; k++) d.pel_dof[k] = -1;
    for (int i = 0; i < s.n_coords; i++)
        if (s.coord_pelvis_trans[i] >= 1 && s.coord_pelvis_trans[i] <= 3) d.pel_dof[s.coord_pelvis_trans[i] - 1] = s.coord_dof[i];
    d.max_pts_per_muscle = 0;
    for (int i = 0; i < s.n_muscles; i++)
        if (s.mus_pt_count[i] > d.max_pts_per_muscle) d.max_pts_per_muscle = s.mus_pt_count[i];
    BIO_CP(gravity);
    d.total_mass = (T)s.total_mass;
    BIO_CPI(body_parent); BIO_CPI(body_axis_begin); BIO_CPI(body_axis_count);
    BIO_CP(body_mass); BIO_CP(body_com); BIO_CP(body_inertia); BIO_CP(body_joint_loc);
    BIO_CPI(axis_kind); BIO_CPI(axis_dof); BIO_CPI(axis_func); BIO_CP(axis_vec);
    BIO_CPI(dof_body); BIO_CPI(dof_anc_mask);
    for (int i = 0; i < BIO_MAX_DOF; i++) {
        d.dof_parent[i] = -1;
        for (int j = i - 1; j >= 0; j--)
            if ((s.dof_anc_mask[i] >> j) & 1u) { d.dof_parent[i] = j; break; }
    }
    for (int b = 0; b < BIO_MAX_BODIES; b++) {
        d.body_last_dof[b] = -1;
        for (int j = 0; j < s.n_dof; j++) if (s.dof_body[j] == b) d.body_last_dof[b] = j;
    }
    // tree levels, per-body path-point lists, moving points, coupled (i,j) entries
    d.n_levels = 0;
    for (int b = 0; b < s.n_bodies; b++) {
        d.body_level[b] = s.body_parent[b] < 0 ? 0 : d.body_level[s.body_parent[b]] + 1;
        if (d.body_level[b] + 1 > d.n_levels) d.n_levels = d.body_level[b] + 1;
    }
    {
        int k = 0;
        for (int b = 0; b < s.n_bodies; b++) {
            d.body_pt_begin[b] = k;
            for (int p = 0; p < s.n_pathpts; p++) if (s.pt_body[p] == b) d.body_pt_list[k++] = p;
            d.body_pt_count[b] = k - d.body_pt_begin[b];
        }
        d.n_moving = 0;
        for (int p = 0; p < BIO_MAX_PATHPTS; p++) d.pt_mov[p] = -1;
        for (int p = 0; p < s.n_pathpts; p++)
            if (s.pt_kind[p] == BIO_PT_MOVING && d.n_moving < 8) { d.pt_mov[p] = (int8_t)d.n_moving; d.moving_pt[d.n_moving++] = p; }
        d.n_entries = 0;
        for (int i = 0; i < s.n_dof; i++)
            for (int j = 0; j <= i; j++)
                if ((s.dof_anc_mask[i] >> j) & 1u) { d.ent_i[d.n_entries] = i; d.ent_j[d.n_entries] = j; d.n_entries++; }
    }
    BIO_CPI(func_kind); BIO_CPI(func_knot_begin); BIO_CPI(func_knot_count);
    BIO_CP(func_c); BIO_CP(knot_x); BIO_CP(knot_c);
    for (int f = 0; f < s.n_funcs; f++) {
        if (s.func_kind[f] != BIO_FUNC_SPLINE) continue;
        const int kb = s.func_knot_begin[f], n = s.func_knot_count[f];
        const double x0 = s.knot_x[kb], x1 = s.knot_x[kb + n - 1], bw = (x1 - x0) / 16.0;
        d.func_bucket_inv[f] = (T)(1.0 / bw);
        for (int b = 0; b < 16; b++) {
            // last knot at or before the bucket start, shifted one bucket down so
            // that rounding of the bucket index in the scalar type cannot skip a knot
            const double xs = x0 + (b > 0 ? b - 1 : 0) * bw;
            int i = 0;
            while (i + 1 < n - 1 && s.knot_x[kb + i + 1] <= xs) i++;
            d.func_bucket[f][b] = (int8_t)i;
        }
        // a spline whose pieces are all the same constant (the z of the reference's moving path points) is
        // evaluated as a Constant: same value and zero derivatives, no knot search
        bool flat = true;
        for (int i = 0; i < n; i++)
            flat = flat && s.knot_c[kb + i][0] == s.knot_c[kb][0] && s.knot_c[kb + i][1] == 0.0 &&
                   s.knot_c[kb + i][2] == 0.0 && s.knot_c[kb + i][3] == 0.0;
        if (flat) { d.func_kind[f] = BIO_FUNC_CONST; d.func_c[f][0] = (T)s.knot_c[kb][0]; d.func_c[f][1] = T(0); }
    }
    {   // L^T D L schedule
        auto tri = [](int i, int j) { return i * (i + 1) / 2 + j; };
        int np = 0;
        for (int st = 0; st < s.n_dof; st++) {
            const int k = s.n_dof - 1 - st;
            d.lt_step_begin[st] = np;
            for (int i = k - 1; i >= 0; i--) {
                if (!((s.dof_anc_mask[k] >> i) & 1u)) continue;
                for (int j = i; j >= 0; j--) {
                    if (!((s.dof_anc_mask[k] >> j) & 1u)) continue;
                    if (np < 320) {
                        d.lt_pack[np] = (uint32_t)tri(i, j) | ((uint32_t)tri(k, i) << 8) | ((uint32_t)tri(k, j) << 16) |
                                        ((uint32_t)i << 24) | (i == j ? 0x80000000u : 0u);
                    }
                    np++;
                }
            }
        }
        d.lt_step_begin[s.n_dof] = np;
        d.lt_max_pairs = 0;
        for (int st = 0; st < s.n_dof; st++)
            if (d.lt_step_begin[st + 1] - d.lt_step_begin[st] > d.lt_max_pairs)
                d.lt_max_pairs = d.lt_step_begin[st + 1] - d.lt_step_begin[st];
        // planar model: every rotation about +-z, every translation along +-x / +-y, no gravity in z
        d.planar = s.gravity[2] == 0.0 ? 1 : 0;
        for (int a = 0; a < s.n_axes; a++) {
            const double ax = s.axis_vec[a][0], ay = s.axis_vec[a][1], az = s.axis_vec[a][2];
            if (s.axis_kind[a] == BIO_AXIS_ROT) { if (!(ax == 0.0 && ay == 0.0 && (az == 1.0 || az == -1.0))) d.planar = 0; }
            else if (!(az == 0.0 && ((ay == 0.0 && (ax == 1.0 || ax == -1.0)) || (ax == 0.0 && (ay == 1.0 || ay == -1.0))))) d.planar = 0;
        }
        for (int b = 0; b < s.n_bodies; b++) {
            const int ab = s.body_axis_begin[b], ae = ab + s.body_axis_count[b];
            bool seen_rot = false;
            for (int a = ab; a < ae; a++) {
                const int dof = s.axis_dof[a];
                const bool rot = s.axis_kind[a] == BIO_AXIS_ROT;
                const double comp = rot ? s.axis_vec[a][2] : (s.axis_vec[a][0] != 0.0 ? s.axis_vec[a][0] : s.axis_vec[a][1]);
                int w = (rot ? 1 : 0) | (comp < 0 ? 2 : 0) | ((!rot && s.axis_vec[a][0] == 0.0) ? 4 : 0) |
                        ((dof >= 0 ? dof : 31) << 3);
                if (rot && !seen_rot && s.body_parent[b] < 0) w |= 1 << 8;
                if (rot) seen_rot = true;
                if (dof >= 0 && (a == ab || s.axis_dof[a - 1] != dof)) w |= 1 << 9;
                if (dof >= 0 && (a == ae - 1 || s.axis_dof[a + 1] != dof)) w |= 1 << 10;
                if (a == ae - 1 && !seen_rot && s.body_parent[b] < 0) w |= 1 << 11;
                d.axis_desc[a] = w;
            }
        }
        d.jc_n = 0;
        for (int sp = 0; sp < s.n_spheres; sp++) {
            const int last = d.body_last_dof[s.sph_body[sp]];
            for (int j = 0; j <= last && last >= 0; j++)
                if ((s.dof_anc_mask[last] >> j) & 1u) {
                    if (d.jc_n < 64) { d.jc_s[d.jc_n] = (uint8_t)sp; d.jc_d[d.jc_n] = (uint8_t)j; }
                    d.jc_n++;
                }
        }
        for (int e = 0; e < d.n_entries; e++) {
            unsigned msk = 0;
            for (int sp = 0; sp < s.n_spheres; sp++) {
                const int last = d.body_last_dof[s.sph_body[sp]];
                if (last >= 0 && ((s.dof_anc_mask[last] >> d.ent_i[e]) & 1u)) msk |= 1u << sp;  // i on the chain => j too
            }
            d.ent_sph[e] = (uint8_t)msk;
        }
        for (int b = 0; b < s.n_bodies; b++)
            d.body_z[b] = (T)((s.body_parent[b] >= 0 ? (double)d.body_z[s.body_parent[b]] : 0.0) + s.body_joint_loc[b][2]);
        d.n_depths = 0;
        for (int i = 0; i < s.n_dof; i++) {
            int dep = 0;
            for (int j = 0; j < i; j++) if ((s.dof_anc_mask[i] >> j) & 1u) dep++;
            d.dof_depth[i] = dep;
            if (dep + 1 > d.n_depths) d.n_depths = dep + 1;
        }
        int nc = 0;
        for (int b = 0; b < s.n_bodies; b++) {
            d.child_begin[b] = nc;
            for (int c = b + 1; c < s.n_bodies; c++) if (s.body_parent[c] == b) d.child_list[nc++] = c;
        }
        d.child_begin[s.n_bodies] = nc;
        int nl2 = 0;
        for (int lev = 0; lev < d.n_levels; lev++) {
            d.level_begin[lev] = nl2;
            for (int b = 0; b < s.n_bodies; b++) if (d.body_level[b] == lev) d.level_body[nl2++] = b;
        }
        d.level_begin[d.n_levels] = nl2;
        for (int a = 0; a < s.n_axes; a++) {
            d.axis_code[a] = 0;
            for (int k = 0; k < 3; k++) {
                const double v = s.axis_vec[a][k], o1 = s.axis_vec[a][(k + 1) % 3], o2 = s.axis_vec[a][(k + 2) % 3];
                if (o1 == 0.0 && o2 == 0.0 && (v == 1.0 || v == -1.0)) d.axis_code[a] = v > 0 ? (k + 1) : -(k + 1);
            }
        }
    }
    BIO_CP(mus_fiso); BIO_CP(mus_lopt); BIO_CP(mus_lts); BIO_CP(mus_vmax); BIO_CP(mus_tact);
    BIO_CP(mus_tdeact); BIO_CP(mus_amin); BIO_CP(mus_beta); BIO_CP(mus_default_act);
    BIO_CP(mus_height); BIO_CP(mus_lm_min); BIO_CP(mus_cot_mass); BIO_CP(mus_slow_twitch);
    BIO_CPI(mus_pt_begin); BIO_CPI(mus_pt_count);
    BIO_CPI(pt_body); BIO_CPI(pt_kind); BIO_CPI(pt_dof); BIO_CPI(pt_func);
    BIO_CP(pt_loc); BIO_CP(pt_range);
    BIO_CPI(sph_body); BIO_CPI(sph_group);
    BIO_CP(sph_loc); BIO_CP(sph_radius); BIO_CP(sph_k); BIO_CP(sph_c); BIO_CP(sph_us);
    BIO_CP(sph_ud); BIO_CP(sph_uv); BIO_CP(sph_vt);
    BIO_CPI(lim_dof);
    BIO_CP(lim_kup); BIO_CP(lim_qup); BIO_CP(lim_klo); BIO_CP(lim_qlo); BIO_CP(lim_damp); BIO_CP(lim_w);
    BIO_CPI(act_dof); BIO_CP(act_min); BIO_CP(act_max);
    BIO_CPI(obs_body); BIO_CP(obs_loc);
    BIO_CPI(coord_dof); BIO_CPI(coord_pelvis_trans); BIO_CP(coord_const);
    BIO_CP(curve_x0); BIO_CP(curve_x1);
    for (int c = 0; c < BIO_N_CURVES; c++) {
        if (sizeof(T) == 4) {
            for (int i = 0; i < BIO_CURVE_N; i++) {
                const double y0 = s.curve_tab[c][i][0], m0 = s.curve_tab[c][i][1];
                const double y1 = s.curve_tab[c][i + 1][0], m1 = s.curve_tab[c][i + 1][1], dl = y1 - y0;
                d.curve_q[c][sizeof(T) == 4 ? i : 0][0] = (T)y0;
                d.curve_q[c][sizeof(T) == 4 ? i : 0][1] = (T)m0;
                d.curve_q[c][sizeof(T) == 4 ? i : 0][2] = (T)(3.0 * dl - 2.0 * m0 - m1);
                d.curve_q[c][sizeof(T) == 4 ? i : 0][3] = (T)(m0 + m1 - 2.0 * dl);
            }
        } else {
            for (int i = 0; i <= BIO_CURVE_N; i++)
                for (int k = 0; k < 2; k++) d.curve_tab[c][sizeof(T) == 4 ? 0 : i][k] = (T)s.curve_tab[c][i][k];
        }
    }
    for (int c = 0; c < BIO_N_CURVES; c++)
        d.curve_inv_h[c] = (T)((double)BIO_CURVE_N / (s.curve_x1[c] - s.curve_x0[c]));
    build_planar_prog(s, d);
}

// Compiled muscle paths (see PlanarProg::mc_seg); any muscle that does not fit leaves path_ok = 0 and the
// kernels stream over the path points instead.
template <typename T>
void compile_paths(const BioModelTables& s, PlanarProg<T>& pr) {
    pr.path_ok = 1;
    pr.mc_nlive = 0;
    { const char* z = getenv("BIO_PLANAR_STREAM_PATHS"); if (z && z[0] == '1') pr.path_ok = 0; }   // tests: take the fallback
    for (int i = 0; i < s.n_muscles && pr.path_ok; i++) {
        const int pb = s.mus_pt_begin[i], pe = pb + s.mus_pt_count[i];
        int n_cond = 0;
        pr.mc_cond[i][0] = pr.mc_cond[i][1] = -1;
        for (int p = pb; p < pe; p++)
            if (s.pt_kind[p] == BIO_PT_CONDITIONAL) { if (n_cond >= 2) { pr.path_ok = 0; break; } pr.mc_cond[i][n_cond++] = p; }
        if (pe > 255) pr.path_ok = 0;
        for (int var = 0; var < P2_MAXVAR && pr.path_ok; var++) {
            double len0 = 0.0;
            int n_live = 0, prev = -1;
            for (int l = 0; l < P2_MAXLIVE; l++) pr.mc_seg[i][var][l] = 0u;
            for (int p = pb; p < pe; p++) {
                if (s.pt_kind[p] == BIO_PT_CONDITIONAL) {
                    const int c = pr.mc_cond[i][0] == p ? 0 : 1;
                    if (!((var >> c) & 1)) continue;
                }
                if (prev >= 0) {
                    const bool fixed = s.pt_kind[p] != BIO_PT_MOVING && s.pt_kind[prev] != BIO_PT_MOVING;
                    if (fixed && s.pt_body[p] == s.pt_body[prev]) {
                        double d2 = 0.0;
                        for (int c = 0; c < 3; c++) { const double dd = s.pt_loc[p][c] - s.pt_loc[prev][c]; d2 += dd * dd; }
                        len0 += sqrt(d2);
                    } else {
                        if (n_live >= P2_MAXLIVE) { pr.path_ok = 0; break; }
                        pr.mc_seg[i][var][n_live++] = (uint32_t)prev | ((uint32_t)p << 8) | 0x80000000u;
                    }
                }
                prev = p;
            }
            pr.mc_len0[i][var] = (T)len0;
            if (n_live > pr.mc_nlive) pr.mc_nlive = n_live;
        }
    }
}

// Muscle geometry of the general (spatial) evaluation in the compiled form: point table, one wrench source
// per (muscle, body it touches) and the sources acting on every body.  gpath_ok = 0: stream over the points.
template <typename T>
void build_general_paths(const BioModelTables& s, DevModel<T>& d) {
    PlanarProg<T>& pr = d.prog;
    pr.gpath_ok = 0;
    if (s.n_muscles < 1 || s.n_pathpts > BIO_MAX_PATHPTS) return;
    int n_src = 0, src_body[P2_MAXSRC];
    for (int i = 0; i < s.n_muscles; i++) {
        int slot_body[P2_MAXSLOT], n_slot = 0, n_moving = 0;
        pr.mus_src0[i] = n_src;
        for (int p = s.mus_pt_begin[i]; p < s.mus_pt_begin[i] + s.mus_pt_count[i]; p++) {
            const int b = s.pt_body[p];
            int slot = -1;
            for (int k = 0; k < n_slot; k++) if (slot_body[k] == b) slot = k;
            if (slot < 0) { if (n_slot >= P2_MAXSLOT) return; slot = n_slot; slot_body[n_slot++] = b; }
            if (s.pt_kind[p] == BIO_PT_MOVING) { n_moving++; if (d.pt_mov[p] < 0) return; }
            const int dof = s.pt_dof[p] >= 0 ? s.pt_dof[p] : 31;
            pr.pt_info[p] = b | (s.pt_kind[p] << 4) | (dof << 6) | (slot << 11) | ((d.pt_mov[p] >= 0 ? d.pt_mov[p] : 0) << 13);
            for (int c = 0; c < 3; c++) pr.pt_xyz[p][c] = (T)s.pt_loc[p][c];
            pr.pt_xyz[p][3] = T(0);
        }
        if (n_moving > 1) return;
        for (int k = 0; k < n_slot; k++) { if (n_src >= P2_MAXSRC) return; src_body[n_src++] = slot_body[k]; }
    }
    pr.mus_src0[s.n_muscles] = n_src;
    compile_paths(s, pr);
    if (!pr.path_ok) return;
    int k = 0;
    for (int b = 0; b < s.n_bodies; b++) {
        pr.inc_begin[b] = k;
        for (int e = 0; e < n_src; e++)
            if (src_body[e] == b) { if (k >= (int)sizeof(pr.inc_src)) return; pr.inc_src[k++] = (uint8_t)e; }
    }
    for (int b = s.n_bodies; b <= BIO_MAX_BODIES; b++) pr.inc_begin[b] = k;
    pr.n_src = n_src;
    pr.gpath_ok = 1;
}

// Planar program of a model (see PlanarProg); prog.ok = 0 when the model does not have the
// root-plus-chains shape, and the general cooperative evaluation is used instead.
template <typename T>
void build_planar_prog(const BioModelTables& s, DevModel<T>& d) {
    PlanarProg<T>& pr = d.prog;
    memset(&pr, 0, sizeof(pr));
    for (int b = 0; b < s.n_bodies && b < BIO_MAX_BODIES; b++)       // contact spheres of every body (any model)
        for (int sp = 0; sp < s.n_spheres && sp < BIO_MAX_SPHERES; sp++) if (s.sph_body[sp] == b) pr.body_sph_mask[b] |= 1 << sp;
    // per-muscle / per-sphere / per-limit constants packed for 16-byte reads, reciprocals from the host (any model:
    // the planar program and the spatial evaluation both read them; a planar model's sphere carries the constant z
    // of its body)
    for (int i = 0; i < s.n_muscles && i < BIO_MAX_MUSCLES; i++) {
        const double k12[12] = {s.mus_fiso[i], s.mus_lopt[i], 1.0 / s.mus_lopt[i], s.mus_height[i] * s.mus_height[i],
                                s.mus_beta[i], s.mus_amin[i], s.mus_lm_min[i], 1.0 / s.mus_lts[i],
                                s.mus_vmax[i] * s.mus_lopt[i], 1.0 / s.mus_tact[i], 1.0 / s.mus_tdeact[i], 0.0};
        for (int c = 0; c < 12; c++) pr.mus_k[i][c] = (T)k12[c];
    }
    for (int sp = 0; sp < s.n_spheres && sp < BIO_MAX_SPHERES; sp++) {
        const double zb = d.planar ? (double)d.body_z[s.sph_body[sp]] : 0.0;
        const double k12[12] = {s.sph_loc[sp][0], s.sph_loc[sp][1], s.sph_loc[sp][2] + zb, s.sph_radius[sp],
                                s.sph_k[sp], 1.5 * s.sph_c[sp], s.sph_ud[sp], 2.0 * (s.sph_us[sp] - s.sph_ud[sp]),
                                s.sph_uv[sp], s.sph_vt[sp], 1.0 / s.sph_vt[sp], 0.0};
        for (int c = 0; c < 12; c++) pr.sph_k[sp][c] = (T)k12[c];
    }
    for (int l = 0; l < s.n_limits && l < BIO_MAX_LIMITS; l++) {
        const double k8[8] = {s.lim_qup[l], s.lim_qlo[l], s.lim_kup[l], s.lim_klo[l], s.lim_damp[l], 1.0 / s.lim_w[l], s.lim_w[l], 0.0};
        for (int c = 0; c < 8; c++) pr.lim_k[l][c] = (T)k8[c];
    }
    // ---- stage 1 (any model): root body 0 carrying <= P2_MAXBR unbranched chains; the walk root joint ->
    // leaf of every chain as a list of elementary-axis steps (used by the planar program and by the scan
    // form of the spatial kinematics) ----
    if (s.n_bodies < 1 || s.n_bodies > BIO_MAX_BODIES || s.body_parent[0] >= 0) return;
    int n_child[BIO_MAX_BODIES] = {0}, child[BIO_MAX_BODIES];
    for (int b = 1; b < s.n_bodies; b++) {
        const int p = s.body_parent[b];
        if (p < 0 || p >= b) return;                      // one root, parents first
        if (p != 0) { if (n_child[p]) return; child[p] = b; }
        n_child[p]++;
    }
    int chain_body[P2_MAXBR][BIO_MAX_BODIES], chain_nb[P2_MAXBR] = {0};
    for (int b = 1; b < s.n_bodies; b++) {
        if (s.body_parent[b] != 0) continue;
        if (pr.n_branches >= P2_MAXBR) { pr.n_branches = 0; return; }
        const int l = pr.n_branches++;
        for (int cur = b;; cur = child[cur]) {
            chain_body[l][chain_nb[l]++] = cur;
            if (!n_child[cur]) break;
        }
    }
    pr.root_body = 0;
    pr.chain_ok = 1;
    for (int l = 0; l < P2_MAXBR; l++) {
        pr.gch_nb[l] = l < pr.n_branches ? chain_nb[l] : 0;
        for (int k = 0; k < BIO_MAX_BODIES; k++) pr.gch_body[l][k] = (l < pr.n_branches && k < chain_nb[l]) ? chain_body[l][k] : 0;
    }
    for (int l = 0; l < (pr.n_branches > 0 ? pr.n_branches : 1); l++) {
        int n = 0;
        auto add_joint = [&](int b, bool root) {
            const int ab = s.body_axis_begin[b], cnt = s.body_axis_count[b];
            if (cnt < 1) pr.chain_ok = 0;                 // every body needs at least one step to publish its frame
            for (int j = 0; j < cnt; j++) {
                if (n >= P2_MAXSTEP) { pr.chain_ok = 0; return; }
                const int a = ab + j, desc = d.axis_desc[a], dof = (desc >> 3) & 31;
                pr.ch_code[l][n] = a | (b << 8) | (dof << 12) | (j == 0 ? P2_F_FIRST : 0) | (j == cnt - 1 ? P2_F_LAST : 0) |
                                   (root ? P2_F_ROOT : 0) | ((desc & 512) ? P2_F_SRESET : 0) |
                                   ((desc & 1024) ? P2_F_SPUB : 0) | ((desc & 256) ? P2_F_OPRE : 0) |
                                   ((desc & 2048) ? P2_F_OPOST : 0) | (s.axis_kind[a] == BIO_AXIS_ROT ? P2_F_ROT : 0);
                for (int c = 0; c < 3; c++) pr.ch_j[l][n][c] = j == 0 ? (T)s.body_joint_loc[b][c] : T(0);
                pr.ch_j[l][n][3] = T(0);
                n++;
            }
        };
        add_joint(0, true);
        for (int k = 0; k < chain_nb[l] && pr.n_branches > 0; k++) add_joint(chain_body[l][k], false);
        pr.ch_n[l] = n;
        int o_step = -1, first = 0, in_dof = 0;
        for (int i = 0; i < n; i++)
            if (pr.ch_code[l][i] & (P2_F_OPRE | P2_F_OPOST)) o_step = i;
        if (o_step < 0) { pr.chain_ok = 0; o_step = 0; }
        for (int i = 0; i < P2_MAXSTEP; i++) {
            if (i >= n) { pr.ch_scan[l][i] = (i & 15) | (o_step << 4); continue; }
            const int code = pr.ch_code[l][i];
            if (code & P2_F_FIRST) first = i;
            in_dof = (code & P2_F_SRESET) || ((code >> 12) & 31) == 31 ? 0 : in_dof + 1;
            if (in_dof > 2) pr.chain_ok = 0;
            pr.ch_scan[l][i] = first | (o_step << 4) | (in_dof << 8);
        }
    }
    if (!pr.chain_ok) return;
    // phase A tasks (planar program: p2_phase_a; spatial evaluation: p3_phase_a)
    int n_mov = 0, mov_of_pt[BIO_MAX_PATHPTS];
    auto build_atasks = [&](const bool planar_z) -> bool {
        n_mov = 0;
        for (int a = 0; a < s.n_axes; a++) { pr.at_func[a] = s.axis_func[a]; pr.at_dof[a] = s.axis_dof[a]; pr.at_add[a] = T(0); pr.at_dst[a] = a; }
        for (int p = 0; p < s.n_pathpts; p++) {
            mov_of_pt[p] = -1;
            if (s.pt_kind[p] != BIO_PT_MOVING) continue;
            if (n_mov >= P2_MAXMOV) return false;
            for (int c = 0; c < 3; c++) {
                const int t = s.n_axes + 3 * n_mov + c;
                pr.at_func[t] = s.pt_func[p][c];
                pr.at_dof[t] = s.pt_dof[p];
                pr.at_add[t] = (planar_z && c == 2) ? d.body_z[s.pt_body[p]] : T(0);
                pr.at_dst[t] = 64 + 3 * n_mov + c;
            }
            pr.mov_dof[n_mov] = s.pt_dof[p];
            mov_of_pt[p] = n_mov++;
        }
        pr.n_atasks = s.n_axes + 3 * n_mov;
        {   // stable insertion sort by cost: splines, then rotations (sin / cos), then the rest
            auto rank = [&](int t) {
                if (d.func_kind[pr.at_func[t]] == BIO_FUNC_SPLINE) return 0;
                if (pr.at_dst[t] < 64 && (d.axis_desc[pr.at_dst[t]] & 1)) return 1;
                return d.func_kind[pr.at_func[t]] == BIO_FUNC_LINEAR ? 2 : 3;
            };
            for (int i = 1; i < pr.n_atasks; i++)
                for (int j = i; j > 0 && rank(j) < rank(j - 1); j--) {
                    std::swap(pr.at_func[j], pr.at_func[j - 1]); std::swap(pr.at_dof[j], pr.at_dof[j - 1]);
                    std::swap(pr.at_add[j], pr.at_add[j - 1]); std::swap(pr.at_dst[j], pr.at_dst[j - 1]);
                }
            pr.a2_cheap = 1;
            for (int t = 16; t < pr.n_atasks; t++) if (rank(t) < 2) pr.a2_cheap = 0;
            for (int t = 0; t < pr.n_atasks; t++) {
                const int f = pr.at_func[t], kind = d.func_kind[f], dst = pr.at_dst[t];
                const int desc = dst < 64 ? d.axis_desc[dst] : 0;
                const int kb = s.func_knot_begin[f], n = s.func_knot_count[f];
                pr.at_i4[t][0] = kind | ((desc & 1) << 2) | (((desc >> 1) & 1) << 3) | (dst << 8);
                pr.at_i4[t][1] = pr.at_dof[t];
                pr.at_i4[t][2] = kind == BIO_FUNC_SPLINE ? kb : 0;
                pr.at_i4[t][3] = kind == BIO_FUNC_SPLINE ? n : 0;
                pr.at_f4[t][0] = kind == BIO_FUNC_SPLINE ? (T)s.knot_x[kb] : d.func_c[f][0];
                pr.at_f4[t][1] = kind == BIO_FUNC_SPLINE ? (T)s.knot_x[kb + n - 1] : d.func_c[f][1];
                pr.at_f4[t][2] = pr.at_add[t];
                pr.at_f4[t][3] = T(0);
            }
        }
        return true;
    };
    // ---- articulated-body schedule of the spatial evaluation: every chain from its leaf, then the root ----
    if (!d.planar) {
        bool ok = true;
        for (int b = 0; b < s.n_bodies; b++) {           // contact spheres of every body in 4 bytes (p3_phase_e)
            unsigned char sp4[4] = {255, 255, 255, 255};
            int ns = 0;
            for (int sp = 0; sp < s.n_spheres; sp++)
                if (s.sph_body[sp] == b) { if (ns < 4) sp4[ns] = (unsigned char)sp; ns++; }
            if (ns > 4) ok = false;
            pr.sph_pk[b] = sp4[0] | (sp4[1] << 8) | (sp4[2] << 16) | ((uint32_t)sp4[3] << 24);
        }
        int nmax = 0, nroot = 0;
        memset(pr.aba_step, 255, sizeof(pr.aba_step));
        for (int l = 0; l < pr.n_branches; l++) {
            int n = 0;
            for (int k = chain_nb[l] - 1; k >= 0; k--) {
                const int b = chain_body[l][k];
                bool first = true;
                for (int dd = s.n_dof - 1; dd >= 0; dd--) {
                    if (s.dof_body[dd] != b) continue;
                    if (n >= P2_MAXABA || dd > 15) { ok = false; break; }
                    pr.aba_step[l][n++] = (uint8_t)(dd | ((first ? b + 1 : 0) << 4));
                    first = false;
                }
                if (first) ok = false;                    // a body without a dof of its own does not occur after weld merging
            }
            if (n > nmax) nmax = n;
        }
        for (int dd = s.n_dof - 1; dd >= 0; dd--)
            if (s.dof_body[dd] == 0) { if (nroot < 8) pr.aba_root[nroot] = (uint8_t)dd; nroot++; }
        if (nroot > 8) ok = false;
        { const char* z = getenv("BIO_NO_ABA"); if (z && z[0] == '1') ok = false; }   // tests: joint-space L^T D L instead
        pr.aba_nsteps = nmax; pr.aba_nroot = nroot > 8 ? 8 : nroot; pr.aba_ok = ok ? 1 : 0;
        pr.atask_ok = (build_atasks(false) && n_mov == d.n_moving) ? 1 : 0;   // packed phase-A tasks (p3_phase_a)
        // bit 1 of aba_ok: the model takes the packed phase A, the scan, the three-lane phase E with the host's dof
        // lists and the articulated-body pass -- coop_eval's FAST instantiation (BIO_NO_FAST3D=1: the general one)
        if (pr.aba_ok && pr.atask_ok && pr.chain_ok && d.gdof_ok && !getenv("BIO_NO_FAST3D")) pr.aba_ok |= 2;
        {   // free root joint: translations along +x, +y, +z of the ground (unit rate) before three rotations (unit rate)
            bool fr = ok && s.body_axis_count[0] == 6 && nroot == 6 && d.gdof_ok;
            int dt[3] = {-1, -1, -1}, drot[3] = {-1, -1, -1};
            const int ab = s.body_axis_begin[0];
            auto unit_rate = [&](int a) {
                const int f = s.axis_func[a];
                return s.func_kind[f] == BIO_FUNC_LINEAR && s.func_c[f][0] == 1.0 && s.func_c[f][1] == 0.0;
            };
            for (int j = 0; j < 3 && fr; j++) {
                const int a = ab + j, dd = s.axis_dof[a];
                if (s.axis_kind[a] != BIO_AXIS_TRANS || dd < 0 || !unit_rate(a)) { fr = false; break; }
                int comp = -1;
                for (int k = 0; k < 3; k++)
                    if (s.axis_vec[a][k] == 1.0 && s.axis_vec[a][(k + 1) % 3] == 0.0 && s.axis_vec[a][(k + 2) % 3] == 0.0) comp = k;
                if (comp < 0 || dt[comp] >= 0) { fr = false; break; }
                dt[comp] = dd;
            }
            for (int j = 0; j < 3 && fr; j++) {
                const int a = ab + 3 + j, dd = s.axis_dof[a];
                if (s.axis_kind[a] != BIO_AXIS_ROT || dd < 0 || !unit_rate(a)) { fr = false; break; }
                drot[j] = dd;
            }
            for (int j = 0; j < 3 && fr; j++) {
                for (int k = 0; k < 3; k++) if (dt[j] == drot[k] || (k != j && (drot[j] == drot[k] || dt[j] == dt[k]))) fr = false;
                for (int dd : {dt[j], drot[j]}) {
                    if (dd < 0 || dd > 15) { fr = false; continue; }
                    if (d.gdof_lim[dd][0] >= 0 || d.gdof_lim[dd][1] >= 0 || d.gdof_movpt[dd][0] >= 0 || d.gdof_movpt[dd][1] >= 0)
                        fr = false;
                }
                if (dt[j] >= 0 && dt[j] <= 15 && d.gdof_act[dt[j]] >= 0) fr = false;   // actuators: rotations only
            }
            { const char* z = getenv("BIO_NO_FREEROOT"); if (z && z[0] == '1') fr = false; }   // tests: dof-by-dof root
            pr.aba_freeroot = fr ? 1 : 0;
            if (fr) {
                pr.aba_fr[0] = dt[0] | (dt[1] << 8) | (dt[2] << 16);
                pr.aba_fr[1] = drot[0] | (drot[1] << 8) | (drot[2] << 16);
            }
        }
    }
    // ---- stage 2 (planar models): the planar program ----
    if (!d.planar) { build_general_paths(s, d); return; }
    auto joint_dofs = [&](int b, int* dofs) {             // distinct dofs of the joint, in axis order
        int n = 0;
        for (int a = s.body_axis_begin[b]; a < s.body_axis_begin[b] + s.body_axis_count[b]; a++) {
            const int dof = s.axis_dof[a];
            if (dof < 0) continue;
            bool seen = false;
            for (int k = 0; k < n; k++) seen = seen || dofs[k] == dof;
            if (!seen) { if (n < 4) dofs[n] = dof; n++; }
        }
        return n;
    };
    { int dofs[4] = {-1, -1, -1, -1}; pr.root_ndof = joint_dofs(0, dofs); if (pr.root_ndof > 3) return;
      for (int k = 0; k < 4; k++) pr.root_dof[k] = k < pr.root_ndof ? dofs[k] : -1; }
    for (int l = 0; l < pr.n_branches; l++) {
        if (chain_nb[l] > P2_MAXCB) return;
        for (int k = 0; k < P2_MAXCB; k++) {
            int dofs[4] = {-1, -1, -1, -1};
            if (k < chain_nb[l] && joint_dofs(chain_body[l][k], dofs) > 1) return;
            pr.br_body[l][k] = k < chain_nb[l] ? chain_body[l][k] : -1;
            pr.br_dof[l][k] = k < chain_nb[l] ? dofs[0] : -1;
        }
        pr.br_nb[l] = chain_nb[l];
    }
    if (s.n_spheres > BIO_MAX_SPHERES) return;
    for (int a = 0; a < s.n_axes; a++) {
        const int desc = d.axis_desc[a];
        const T sg = (desc & 2) ? T(-1) : T(1);
        pr.ax_k[a][0] = (desc & 1) ? sg : T(0);
        pr.ax_k[a][1] = (!(desc & 1) && !(desc & 4)) ? sg : T(0);
        pr.ax_k[a][2] = (!(desc & 1) && (desc & 4)) ? -sg : T(0);
        pr.ax_k[a][3] = T(0);
    }
    pr.scan_ok = pr.n_branches >= 1;
    for (int l = 0; l < pr.n_branches; l++) if (pr.ch_n[l] > 8) pr.scan_ok = 0;
    if (!build_atasks(true)) return;
    // path points, muscle slots and wrench sources
    int n_src = 0;
    int src_body[P2_MAXSRC];
    for (int i = 0; i < s.n_muscles; i++) {
        int slot_body[P2_MAXSLOT], n_slot = 0, n_moving = 0;
        pr.mus_src0[i] = n_src;
        for (int p = s.mus_pt_begin[i]; p < s.mus_pt_begin[i] + s.mus_pt_count[i]; p++) {
            const int b = s.pt_body[p];
            int slot = -1;
            for (int k = 0; k < n_slot; k++) if (slot_body[k] == b) slot = k;
            if (slot < 0) { if (n_slot >= P2_MAXSLOT) return; slot = n_slot; slot_body[n_slot++] = b; }
            if (s.pt_kind[p] == BIO_PT_MOVING) n_moving++;
            const int dof = s.pt_dof[p] >= 0 ? s.pt_dof[p] : 31;
            pr.pt_info[p] = b | (s.pt_kind[p] << 4) | (dof << 6) | (slot << 11) | ((mov_of_pt[p] >= 0 ? mov_of_pt[p] : 0) << 13);
            pr.pt_xyz[p][0] = (T)s.pt_loc[p][0];
            pr.pt_xyz[p][1] = (T)s.pt_loc[p][1];
            pr.pt_xyz[p][2] = (T)(s.pt_loc[p][2] + (double)d.body_z[b]);
            pr.pt_xyz[p][3] = T(0);
        }
        if (n_moving > 1) return;
        for (int k = 0; k < n_slot; k++) { if (n_src >= P2_MAXSRC) return; src_body[n_src++] = slot_body[k]; }
    }
    pr.mus_src0[s.n_muscles] = n_src;
    compile_paths(s, pr);
    for (int b = 0; b < s.n_bodies; b++) {
        pr.body_k[b][0] = (T)s.body_com[b][0]; pr.body_k[b][1] = (T)s.body_com[b][1];
        pr.body_k[b][2] = (T)s.body_mass[b]; pr.body_k[b][3] = (T)s.body_inertia[b][2];
    }
    static_assert(P2_MAXCB == 3, "br_i8 packs three bodies per chain");
    for (int l = 0; l < P2_MAXBR; l++) {
        for (int k = 0; k < 3; k++) {
            const bool has = l < pr.n_branches && k < pr.br_nb[l];
            pr.br_i8[l][k] = has ? pr.br_body[l][k] : -1;
            pr.br_i8[l][4 + k] = has ? pr.br_dof[l][k] : -1;
        }
        pr.br_i8[l][3] = l < pr.n_branches ? pr.br_nb[l] : 0;
        pr.br_i8[l][7] = 0;
    }
    for (int r = 0; r < 3; r++) pr.root_i4[r] = r < pr.root_ndof ? pr.root_dof[r] : -1;
    pr.root_i4[3] = pr.root_body;
    pr.sph_src0 = n_src;
    for (int sp = 0; sp < s.n_spheres; sp++) { if (n_src >= P2_MAXSRC) return; src_body[n_src++] = s.sph_body[sp]; }
    pr.n_src = n_src;
    {
        int k = 0;
        for (int b = 0; b < s.n_bodies; b++) {
            pr.inc_begin[b] = k;
            for (int e = 0; e < n_src; e++)
                if (src_body[e] == b) { if (k >= (int)sizeof(pr.inc_src)) return; pr.inc_src[k++] = (uint8_t)e; }
        }
        for (int b = s.n_bodies; b <= BIO_MAX_BODIES; b++) pr.inc_begin[b] = k;
    }
    // per-dof generalized-force inputs
    for (int dd = 0; dd < BIO_MAX_DOF; dd++) { pr.dof_lim[dd][0] = pr.dof_lim[dd][1] = pr.dof_mov[dd][0] = pr.dof_mov[dd][1] = pr.dof_act[dd] = -1; }
    for (int l = 0; l < s.n_limits; l++) {
        const int dd = s.lim_dof[l];
        if (pr.dof_lim[dd][0] < 0) pr.dof_lim[dd][0] = (int8_t)l; else if (pr.dof_lim[dd][1] < 0) pr.dof_lim[dd][1] = (int8_t)l; else return;
    }
    for (int k = 0; k < n_mov; k++) {
        const int dd = pr.mov_dof[k];
        if (pr.dof_mov[dd][0] < 0) pr.dof_mov[dd][0] = (int8_t)k; else if (pr.dof_mov[dd][1] < 0) pr.dof_mov[dd][1] = (int8_t)k; else return;
    }
    if (s.is_torque)
        for (int a = 0; a < s.n_act; a++) {
            const int dd = s.act_dof[a];
            if (dd < 0) continue;
            if (pr.dof_act[dd] >= 0) return;
            pr.dof_act[dd] = (int8_t)a;
        }
    pr.inc8_ok = 1;
    for (int b = 0; b < s.n_bodies; b++) {
        unsigned char src8[8], sp4[4];
        memset(src8, 255, sizeof(src8)); memset(sp4, 255, sizeof(sp4));
        const int cnt = pr.inc_begin[b + 1] - pr.inc_begin[b];
        if (cnt > 8) pr.inc8_ok = 0;
        for (int k = 0; k < cnt && k < 8; k++) src8[k] = pr.inc_src[pr.inc_begin[b] + k];
        int ns = 0;
        for (int sp = 0; sp < s.n_spheres; sp++)
            if (s.sph_body[sp] == b) { if (ns < 4) sp4[ns] = (unsigned char)sp; ns++; }
        if (ns > 4) pr.inc8_ok = 0;
        pr.inc_pk[b][0] = src8[0] | (src8[1] << 8) | (src8[2] << 16) | ((uint32_t)src8[3] << 24);
        pr.inc_pk[b][1] = src8[4] | (src8[5] << 8) | (src8[6] << 16) | ((uint32_t)src8[7] << 24);
        pr.sph_pk[b] = sp4[0] | (sp4[1] << 8) | (sp4[2] << 16) | ((uint32_t)sp4[3] << 24);
    }
    {   // free planar root: +x and +y translations of the ground (unit rate), then a rotation about +z (unit rate)
        bool fr = s.body_axis_count[0] == 3 && pr.root_ndof == 3;
        int dw = -1, dx = -1, dy = -1;
        const int ab = s.body_axis_begin[0];
        auto unit_rate = [&](int a) {
            const int f = s.axis_func[a];
            return s.func_kind[f] == BIO_FUNC_LINEAR && s.func_c[f][0] == 1.0 && s.func_c[f][1] == 0.0;
        };
        for (int j = 0; j < 3 && fr; j++) {
            const int a = ab + j, dd = s.axis_dof[a];
            const double vx = s.axis_vec[a][0], vy = s.axis_vec[a][1], vz = s.axis_vec[a][2];
            if (dd < 0 || !unit_rate(a)) { fr = false; break; }
            if (j < 2) {
                if (s.axis_kind[a] != BIO_AXIS_TRANS) { fr = false; break; }
                if (vx == 1.0 && vy == 0.0 && vz == 0.0 && dx < 0) dx = dd;
                else if (vx == 0.0 && vy == 1.0 && vz == 0.0 && dy < 0) dy = dd;
                else fr = false;
            } else {
                if (s.axis_kind[a] != BIO_AXIS_ROT || !(vx == 0.0 && vy == 0.0 && vz == 1.0)) fr = false;
                dw = dd;
            }
        }
        if (fr && (dw == dx || dw == dy || dx == dy || dw < 0 || dx < 0 || dy < 0)) fr = false;
        if (fr)
            for (int dd : {dw, dx, dy})
                if (pr.dof_lim[dd][0] >= 0 || pr.dof_lim[dd][1] >= 0 || pr.dof_mov[dd][0] >= 0 || pr.dof_mov[dd][1] >= 0) fr = false;
        { const char* z = getenv("BIO_NO_FREEROOT"); if (z && z[0] == '1') fr = false; }   // tests: dof-by-dof root
        pr.root_ident = fr ? 1 : 0;
        pr.root_id4[0] = fr ? dw : 0; pr.root_id4[1] = fr ? dx : 0; pr.root_id4[2] = fr ? dy : 0; pr.root_id4[3] = 0;
    }
    pr.ok = 1;
    pr.coop_aba = 1;
    { const char* z = getenv("BIO_PLANAR_SERIAL_ABA"); if (z && z[0] == '1') pr.coop_aba = 0; }   // tests: one lane per chain
    // bit 1 of coop_aba: coop_eval_planar's FAST instantiation (scan kinematics, compiled paths, packed source lists,
    // cooperative pass with the direct root solve); BIO_NO_FAST2D=1: the general one
    if (pr.coop_aba && pr.scan_ok && pr.path_ok && pr.inc8_ok && pr.root_ident && !getenv("BIO_NO_FAST2D")) pr.coop_aba |= 2;
}

// Observation layout as a descriptor per slot (same order as write_obs in
// bio_kernels.cuh; reference env2D.py:158-230).
template <typename T>
int build_obs_desc(const BioModelTables& s, const BioTaskConfig& t, DevModel<T>& d) {
    // One descriptor per observation slot, resolved as far as the host can: kind 0 gait phase; 1 / 2 / 3 q / u /
    // udot of dof idx; 4 / 5 reference q / u of coordinate idx (next row); 6 / 8 position / velocity component idx
    // of the observed body origins; 7 / 9 centre-of-mass position / velocity; 10 / 11 / 12 activation, fibre
    // length, fibre velocity of muscle idx; 13 contact wrench component idx times obs_cst (1 / body weight,
    // 1 / (weight * height) for moments); 14 the constant obs_cst (locked coordinates).  Pelvis selector 1..3:
    // subtract pelvis_tx / ty / tz (positions are reported relative to the pelvis).
    int o = 0;
    auto put = [&](int kind, int idx, int pel = 0, double cst = 0.0) {
        if (o < BIO_COOP_MAX_OBS_DIM) { d.obs_desc[o] = (kind << 16) | (pel << 12) | idx; d.obs_cst[o] = (T)cst; }
        o++;
    };
    auto coord = [&](int kind, int i) {
        if (s.coord_dof[i] >= 0) put(kind, s.coord_dof[i]);
        else put(14, 0, 0, kind == 1 ? s.coord_const[i] : 0.0);
    };
    put(0, 0);
    for (int i = 0; i < s.n_coords; i++) if (!s.coord_pelvis_trans[i]) coord(1, i);
    for (int i = 0; i < s.n_coords; i++) coord(2, i);
    for (int i = 0; i < s.n_coords; i++) coord(3, i);
    if (t.use_target_obs) {
        for (int i = 0; i < s.n_coords; i++) if (s.coord_pelvis_trans[i] != 1) put(4, i);
        for (int i = 0; i < s.n_coords; i++) if (s.coord_pelvis_trans[i] != 1) put(5, i);
    }
    for (int p = 0; p < t.n_obs_bodies; p++) for (int j = 0; j < 3; j++) put(6, p * 3 + j, 1 + j);
    for (int j = 0; j < 3; j++) put(7, j, 1 + j);
    for (int p = 0; p < t.n_obs_body_vel; p++) for (int j = 0; j < 3; j++) put(8, p * 3 + j);
    for (int j = 0; j < 3; j++) put(9, j);
    for (int i = 0; i < s.n_muscles; i++) { put(10, i); put(11, i); put(12, i); }
    const double weight = fabs(s.total_mass * s.gravity[1]);
    if (t.use_grf) for (int g = 0; g < 2; g++) for (int j = 0; j < 6; j++) put(13, g * 6 + j, 0, 1.0 / (j < 3 ? weight : weight * t.height));
    return o;
}

template <typename T>
void convert_task(const BioTaskConfig& s, DevTask<T>& d) {
    memset(&d, 0, sizeof(d));
    d.n_substeps = s.n_substeps; d.integrator = s.integrator; d.newton_iters = s.newton_iters;
    d.horizon = s.horizon; d.feed_mean_action = s.feed_mean_action; d.use_pd = s.use_pd; d.n_pd = s.n_pd;
    d.test_mode = s.test_mode; d.cycle = s.cycle; d.n_steps = s.n_steps;
    d.reset_max_index = s.reset_max_index; d.ref_mirror = s.ref_mirror; d.auto_reset = s.auto_reset;
    d.term_obspt = s.term_obspt; d.term_feet_cross = s.term_feet_cross;
    d.feet_obspt[0] = s.feet_obspt[0]; d.feet_obspt[1] = s.feet_obspt[1];
    d.reward_use_feet = s.reward_use_feet; d.effort_torque = s.effort_torque;
    d.effort_use_dy = s.effort_use_dy; d.n_reward_terms = s.n_reward_terms;
    BIO_CPI(rew_obspt); BIO_CPI(rew_refbody);
    d.use_target_obs = s.use_target_obs; d.use_grf = s.use_grf; d.n_obs_bodies = s.n_obs_bodies;
    d.n_obs_body_vel = s.n_obs_body_vel; d.obs_dim = s.obs_dim;
    d.perturb = s.perturb; d.perturb_obspt = s.perturb_obspt;
    d.perturb_negative_only = s.perturb_negative_only;
    BIO_CPI(pd_x_coord); BIO_CPI(pd_v_coord); BIO_CP(pd_kp); BIO_CP(pd_kv);
    d.dt = (T)s.dt; d.h_sub = d.dt / (T)d.n_substeps; d.h_pad_ = T(0); d.term_height = (T)s.term_height; d.term_limit_force = (T)s.term_limit_force;
    d.term_acc = (T)s.term_acc; d.w_imitate = (T)s.w_imitate; d.w_effort = (T)s.w_effort;
    d.w_action = (T)s.w_action; d.action_r_scale = (T)s.action_r_scale;
    d.max_actuation = (T)s.max_actuation; d.height = (T)s.height;
    d.perturb_thresh = (T)s.perturb_thresh; d.perturb_force = (T)s.perturb_force;
}

}  // namespace bio
