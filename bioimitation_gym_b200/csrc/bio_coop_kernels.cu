// bio_coop_kernels.cu -- instantiates the cooperative step kernel for the scalar type BIO_T
// and the size class BIO_CLS (0: half-warp per env, planar gait models; 1: warp per env, 3D),
// in both launch shapes (see COOP_THREADS_LO / _HI in bio_coop.cuh).
#include "bio_launch.cuh"

#if !defined(BIO_T) || !defined(BIO_CLS)
#error "compile with -DBIO_T=float|double -DBIO_CLS=0|1"
#endif

namespace bio {

namespace {
constexpr int LO = COOP_THREADS_LO(BIO_T), HI = COOP_THREADS_HI(BIO_T, BIO_CLS);
}

template <>
cudaError_t coop_set_smem<BIO_T, BIO_CLS>(int threads, int smem) {
    if (threads == HI && HI != LO)
        return cudaFuncSetAttribute(bio_coop_step_kernel<BIO_T, BIO_CLS, HI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return cudaFuncSetAttribute(bio_coop_step_kernel<BIO_T, BIO_CLS, LO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

template <>
int coop_ctas_per_sm<BIO_T, BIO_CLS>(int threads, int smem) {
    int n = 0;
    cudaError_t e;
    if (threads == HI && HI != LO)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bio_coop_step_kernel<BIO_T, BIO_CLS, HI>, HI, (size_t)smem);
    else
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bio_coop_step_kernel<BIO_T, BIO_CLS, LO>, LO, (size_t)smem);
    return e == cudaSuccess ? n : 0;
}

template <>
void launch_coop<BIO_T, BIO_CLS>(int threads, int grid, size_t smem, cudaStream_t s, const DevModel<BIO_T>* gm,
                                 const DevTask<BIO_T>& c, const EnvState<BIO_T>& st, int n, unsigned long long seed,
                                 long long env_offset, const BIO_T* actions, BIO_T* obs, BIO_T* reward, uint8_t* done,
                                 BIO_T* terms, double* stats) {
    if (threads == HI && HI != LO)
        bio_coop_step_kernel<BIO_T, BIO_CLS, HI><<<grid, HI, smem, s>>>(gm, c, st, n, seed, env_offset, actions, obs,
                                                                        reward, done, terms, stats);
    else
        bio_coop_step_kernel<BIO_T, BIO_CLS, LO><<<grid, LO, smem, s>>>(gm, c, st, n, seed, env_offset, actions, obs,
                                                                        reward, done, terms, stats);
}

}  // namespace bio
