// coop_emul.cu -- HOST emulation of the planar program of the cooperative step kernel
// (bioimitation_gym_b200/csrc/bio_coop_planar.cuh).  TEST INFRASTRUCTURE ONLY: the phases of
// the evaluation are __host__ __device__ functions that exchange data through the per-env work
// buffer only, so running them lane by lane, phase by phase on the CPU executes exactly the
// code the GPU runs (minus the warp barriers).  tests/test_planar_program.py compares the
// result with the CPU oracle; nothing in the product path calls this.
#include <string.h>

#include "../../bioimitation_gym_b200/csrc/bio_coop.cuh"

namespace {

template <typename T>
int run(const BioModelTables* s, int newton_iters, const double* q, const double* u, const double* act,
        const double* lm, const double* ctrl, const double* vn0, double h_imp, double ext_fx, int ext_pt, double* udot,
        double* adot, double* lmdot, double* misc) {
    using namespace bio;
    typedef EnvWork<T, 0> Work;
    constexpr int G = CoopCls<0>::G;
    DevModel<T>* m = new DevModel<T>();
    convert_model(*s, *m);
    if (!m->prog.ok) { delete m; return -1; }
    if (s->n_dof > CoopCls<0>::ND || s->n_muscles > CoopCls<0>::NM || s->n_axes > CoopCls<0>::NAX) { delete m; return -2; }
    Work* E = new Work();
    memset(E, 0, sizeof(Work));
    for (int i = 0; i < s->n_dof; i++) { E->q[i] = (T)q[i]; E->u[i] = (T)u[i]; }
    for (int i = 0; i < s->n_muscles; i++) { E->act[i] = (T)act[i]; E->lm[i] = (T)lm[i]; E->vn[i] = vn0 ? (T)vn0[i] : T(0); }
    for (int i = 0; i < s->n_act; i++) E->ctrl[i] = (T)ctrl[i];
    for (int l = 0; l < G; l++) p2_phase_a<T, 0>(*m, *E, l);
    for (int l = 0; l < G; l++) p2_phase_b<T, 0>(*m, *E, l);
    for (int l = 0; l < G; l++) { p2_phase_c<T, 0>(*m, *E, l, newton_iters, (T)h_imp, true); p2_phase_d<T, 0>(*m, *E, l, (T)h_imp); }
    for (int l = 0; l < G; l++) p2_phase_e<T, 0>(*m, *E, l, (T)h_imp, (T)ext_fx, ext_pt);
    for (int l = 0; l < G; l++) p2_phase_f<T, 0>(*m, *E, l);
    for (int l = 0; l < G; l++) p2_phase_g<T, 0>(*m, *E, l);
    for (int l = 0; l < G; l++) p2_readout_1<T, 0>(*m, *E, l);
    for (int l = 0; l < G; l++) p2_readout_2<T, 0>(*m, *E, l);
    for (int i = 0; i < s->n_dof; i++) udot[i] = (double)E->udot[i];
    for (int i = 0; i < s->n_muscles; i++) { adot[i] = (double)E->adot[i]; lmdot[i] = (double)E->lmdot[i]; }
    // misc: com_pos[3], com_vel[3], contact[12], max_limit, fiber force[nm], active fiber force[nm]
    int o = 0;
    for (int c = 0; c < 3; c++) misc[o++] = (double)E->com_pos[c];
    for (int c = 0; c < 3; c++) misc[o++] = (double)E->com_vel[c];
    for (int g = 0; g < 2; g++) for (int c = 0; c < 6; c++) misc[o++] = (double)E->contact[g][c];
    misc[o++] = (double)E->max_limit;
    for (int i = 0; i < s->n_muscles; i++) misc[o++] = (double)E->ffib[i];
    for (int i = 0; i < s->n_muscles; i++) misc[o++] = (double)E->fact[i];
    for (int p = 0; p < s->n_obspts; p++) for (int c = 0; c < 3; c++) misc[o++] = (double)E->x.out.obs_pos[p][c];
    for (int p = 0; p < s->n_obspts; p++) for (int c = 0; c < 3; c++) misc[o++] = (double)E->x.out.obs_vel[p][c];
    delete E;
    delete m;
    return 0;
}

}  // namespace

extern "C" int emul_planar_eval(const BioModelTables* s, int precision, int newton_iters, const double* q,
                                const double* u, const double* act, const double* lm, const double* ctrl,
                                const double* vn0, double h_imp, double ext_fx, int ext_pt, double* udot, double* adot, double* lmdot,
                                double* misc) {
    if (precision == BIO_PREC_F32)
        return run<float>(s, newton_iters, q, u, act, lm, ctrl, vn0, h_imp, ext_fx, ext_pt, udot, adot, lmdot, misc);
    return run<double>(s, newton_iters, q, u, act, lm, ctrl, vn0, h_imp, ext_fx, ext_pt, udot, adot, lmdot, misc);
}

// sizes that decide how many CTAs fit on an SM (bytes): model block and per-env work buffer
extern "C" void emul_sizes(int64_t* out) {
    out[0] = (int64_t)sizeof(bio::DevModel<float>);
    out[1] = (int64_t)sizeof(bio::EnvWork<float, 0>);
    out[2] = (int64_t)sizeof(bio::EnvWork<float, 1>);
    out[3] = (int64_t)sizeof(bio::DevModel<double>);
    out[4] = (int64_t)sizeof(bio::EnvWork<double, 0>);
    out[5] = (int64_t)sizeof(bio::EnvWork<double, 1>);
    out[6] = (int64_t)sizeof(bio::PlanarProg<float>);
}

// which fast paths the host-built program enables for a model: ok (planar program), scan_ok (planar
// kinematics as warp scans), chain_ok (chain lists: spatial kinematics as warp scans), n_branches,
// steps of chain 0 / 1, phase-A tasks, wrench sources
extern "C" void emul_prog_info(const BioModelTables* s, int32_t* out) {
    bio::DevModel<float>* m = new bio::DevModel<float>();
    bio::convert_model(*s, *m);
    out[0] = m->prog.ok; out[1] = m->prog.scan_ok; out[2] = m->prog.chain_ok; out[3] = m->prog.n_branches;
    out[4] = m->prog.ch_n[0]; out[5] = m->prog.ch_n[1]; out[6] = m->prog.n_atasks; out[7] = m->prog.n_src;
    out[8] = m->prog.path_ok; out[9] = m->prog.mc_nlive; out[10] = m->prog.a2_cheap;
    delete m;
}
