// bio_coop_kernels.cu -- instantiates the cooperative step kernel for the scalar type BIO_T
// and the size class BIO_CLS (0: half-warp per env, planar gait models; 1: warp per env, 3D).
#include "bio_launch.cuh"

#if !defined(BIO_T) || !defined(BIO_CLS)
#error "compile with -DBIO_T=float|double -DBIO_CLS=0|1"
#endif

namespace bio {

template <>
cudaError_t coop_set_smem<BIO_T, BIO_CLS>(int smem) {
    return cudaFuncSetAttribute(bio_coop_step_kernel<BIO_T, BIO_CLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

template <>
int coop_ctas_per_sm<BIO_T, BIO_CLS>(int smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bio_coop_step_kernel<BIO_T, BIO_CLS>, COOP_THREADS(BIO_T),
                                                      (size_t)smem) != cudaSuccess)
        return 0;
    return n;
}

template <>
void launch_coop<BIO_T, BIO_CLS>(int grid, size_t smem, cudaStream_t s, const DevModel<BIO_T>* gm,
                                 const DevTask<BIO_T>& c, const EnvState<BIO_T>& st, int n, unsigned long long seed,
                                 long long env_offset, const BIO_T* actions, BIO_T* obs, BIO_T* reward, uint8_t* done,
                                 BIO_T* terms, double* stats) {
    bio_coop_step_kernel<BIO_T, BIO_CLS><<<grid, COOP_THREADS(BIO_T), smem, s>>>(gm, c, st, n, seed, env_offset, actions,
                                                                                obs, reward, done, terms, stats);
}

}  // namespace bio
