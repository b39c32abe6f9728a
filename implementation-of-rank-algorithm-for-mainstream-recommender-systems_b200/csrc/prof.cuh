// prof.cuh — optional phase timing of a kernel (development aid; compiled out unless -DRK_PROFILE).
// Thread 0 of every CTA accumulates clock64() deltas per phase in shared memory and adds them to
// g_rk_prof at exit; scripts/phase_profile.py builds a profiling copy of the library and prints
// the shares.  Slot 11 = total CTA cycles, 12/13 = free counters.
#pragma once
#ifdef RK_PROFILE
namespace rk { extern __device__ unsigned long long g_rk_prof[16]; }
#define PROF_DECL __shared__ unsigned long long prof_acc[16]; long long prof_t0 = clock64(); const long long prof_start = prof_t0; \
    if (threadIdx.x < 16) prof_acc[threadIdx.x] = 0; __syncthreads();
#define PROF(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); prof_acc[i] += t_ - prof_t0; prof_t0 = t_; } } while (0)
#define PROF_COUNT(i) do { if (threadIdx.x == 0) prof_acc[i] += 1; } while (0)
#define PROF_END do { if (threadIdx.x == 0) { prof_acc[11] = clock64() - prof_start; for (int i_ = 0; i_ < 16; ++i_) atomicAdd(&rk::g_rk_prof[i_], prof_acc[i_]); } } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_COUNT(i)
#define PROF_END
#endif
