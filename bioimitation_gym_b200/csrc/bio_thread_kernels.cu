// bio_thread_kernels.cu -- instantiates the thread-per-env kernels (step, reset, equilibrium
// table, debug evaluation, state transpose) for the scalar type BIO_T.
#include "bio_launch.cuh"

#ifndef BIO_T
#error "compile with -DBIO_T=float or -DBIO_T=double"
#endif

namespace bio {

template <>
cudaError_t thread_set_smem<BIO_T>(int smem) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(bio_step_kernel<BIO_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    if ((e = cudaFuncSetAttribute(bio_reset_kernel<BIO_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    if ((e = cudaFuncSetAttribute(bio_eval_kernel<BIO_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    if ((e = cudaFuncSetAttribute(bio_id_kernel<BIO_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))) return e;
    return cudaFuncSetAttribute(bio_lm0_kernel<BIO_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

template <>
void launch_step<BIO_T>(int grid, int block, size_t smem, cudaStream_t s, const DevModel<BIO_T>* gm,
                        const DevTask<BIO_T>& c, const EnvState<BIO_T>& st, int n, unsigned long long seed,
                        long long env_offset, const BIO_T* actions, BIO_T* obs, BIO_T* reward, uint8_t* done,
                        BIO_T* terms, double* stats) {
    bio_step_kernel<BIO_T><<<grid, block, smem, s>>>(gm, c, st, n, seed, env_offset, actions, obs, reward, done, terms,
                                                     stats);
}

template <>
void launch_reset<BIO_T>(int grid, int block, size_t smem, cudaStream_t s, const DevModel<BIO_T>* gm,
                         const DevTask<BIO_T>& c, const EnvState<BIO_T>& st, int n, unsigned long long seed,
                         long long env_offset, const uint8_t* mask, BIO_T* obs, int bump) {
    bio_reset_kernel<BIO_T><<<grid, block, smem, s>>>(gm, c, st, n, seed, env_offset, mask, obs, bump);
}

template <>
void launch_lm0<BIO_T>(int grid, int block, size_t smem, const DevModel<BIO_T>* gm, const BIO_T* ref_q, int rows,
                       int n_coords, BIO_T* lm0) {
    bio_lm0_kernel<BIO_T><<<grid, block, smem>>>(gm, ref_q, rows, n_coords, lm0);
}

template <>
void launch_eval<BIO_T>(int grid, int block, size_t smem, cudaStream_t s, const DevModel<BIO_T>* gm,
                        const DevTask<BIO_T>& c, const EnvState<BIO_T>& st, int n, unsigned long long seed,
                        long long env_offset, const BIO_T* controls, const DebugOut<BIO_T>& d) {
    bio_eval_kernel<BIO_T><<<grid, block, smem, s>>>(gm, c, st, n, seed, env_offset, controls, d);
}

template <>
void launch_id<BIO_T>(int grid, int block, size_t smem, cudaStream_t s, const DevModel<BIO_T>* gm, const DevTask<BIO_T>& c,
                      const EnvState<BIO_T>& st, int n, unsigned long long seed, long long env_offset, int op,
                      const BIO_T* x, const BIO_T* controls, const BIO_T* shift, BIO_T* out) {
    bio_id_kernel<BIO_T><<<grid, block, smem, s>>>(gm, c, st, n, seed, env_offset, op, x, controls, shift, out);
}

template <>
void launch_transpose<BIO_T>(unsigned grid, int block, cudaStream_t s, const BIO_T* src, BIO_T* dst, int n, int k,
                             int to_soa) {
    bio_transpose_kernel<BIO_T><<<grid, block, 0, s>>>(src, dst, n, k, to_soa);
}

}  // namespace bio
