// bio_launch.cuh -- host-side launchers, one translation unit per (scalar type[, size class]).
//
// The kernels are templates; instantiating all of them in one translation unit makes
// the build serial (minutes).  bio_capi.cu only sees these declarations; the definitions
// and explicit instantiations live in bio_thread_kernels.cu (-DBIO_T=float|double) and
// bio_coop_kernels.cu (-DBIO_T=... -DBIO_CLS=0|1), compiled in parallel by build().
#pragma once
#include <cuda_runtime.h>

#include "bio_coop.cuh"

namespace bio {

template <typename T> cudaError_t thread_set_smem(int smem);
template <typename T>
void launch_step(int grid, int block, size_t smem, cudaStream_t s, const DevModel<T>* gm, const DevTask<T>& c,
                 const EnvState<T>& st, int n, unsigned long long seed, long long env_offset, const T* actions, T* obs,
                 T* reward, uint8_t* done, T* terms, double* stats);
template <typename T>
void launch_reset(int grid, int block, size_t smem, cudaStream_t s, const DevModel<T>* gm, const DevTask<T>& c,
                  const EnvState<T>& st, int n, unsigned long long seed, long long env_offset, const uint8_t* mask,
                  T* obs, int bump);
template <typename T>
void launch_lm0(int grid, int block, size_t smem, const DevModel<T>* gm, const T* ref_q, int rows, int n_coords, T* lm0);
template <typename T>
void launch_eval(int grid, int block, size_t smem, cudaStream_t s, const DevModel<T>* gm, const DevTask<T>& c,
                 const EnvState<T>& st, int n, unsigned long long seed, long long env_offset, const T* controls,
                 const DebugOut<T>& d);
template <typename T>
void launch_id(int grid, int block, size_t smem, cudaStream_t s, const DevModel<T>* gm, const DevTask<T>& c,
               const EnvState<T>& st, int n, unsigned long long seed, long long env_offset, int op, const T* x,
               const T* controls, const T* shift, T* out);
template <typename T>
void launch_transpose(unsigned grid, int block, cudaStream_t s, const T* src, T* dst, int n, int k, int to_soa);

// cooperative kernel, launch shape `threads` (COOP_THREADS_LO / COOP_THREADS_HI)
template <typename T, int CLS> cudaError_t coop_set_smem(int threads, int smem);
template <typename T, int CLS> int coop_ctas_per_sm(int threads, int smem);   // resident CTAs per SM at this shared-memory size
template <typename T, int CLS>
void launch_coop(int threads, int grid, size_t smem, cudaStream_t s, const DevModel<T>* gm, const DevTask<T>& c,
                 const EnvState<T>& st, int n, unsigned long long seed, long long env_offset, const T* actions, T* obs,
                 T* reward, uint8_t* done, T* terms, double* stats);

}  // namespace bio
