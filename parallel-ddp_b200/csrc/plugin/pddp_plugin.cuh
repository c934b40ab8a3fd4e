// pddp_plugin.cuh -- the plant plug-in surface of libpddp: what a plant author's header may rely on.
//
// A plant is a header in the style of the reference's plants/{dynamics,cost}_*.cuh, compiled into its own translation unit
// (csrc/plant_tu.cu) together with the solver kernels that call it.  plant_tu.cu includes the header INSIDE a namespace of its own
// (several plants share one library), after <cuda_runtime.h> and <math.h>: a plant header that needs other system headers has them
// included first with -include.  It defines
//
//     NUM_POS, STATE_SIZE, CONTROL_SIZE                                            (config.cuh:21-61)
//     initI<T>(T *s_I), initT<T>(T *s_T)                                           (dynamics_arm.cuh:71,351; host, 36*NUM_POS floats each)
//     dynamics<T>(s_qdd, s_x, s_u, d_I, d_Tbody, s_eePos = nullptr, reps = 1, s_eeVel = nullptr)      (dynamics_arm.cuh:2095-2097)
//     dynamicsGradient<T>(s_dqdd, s_qdd, s_x, s_u, d_I, d_Tbody)                   (dynamics_arm.cuh:2165-2167)
//     costFunc<T>(xk, uk, xgk, k, Q1, Q2, R, QF1, QF2)  -> T                        (cost_arm.cuh:128-130)
//     costGrad<T>(Hk, gk, xk, uk, xgk, k, ld_H, Q1, Q2, R, QF1, QF2)               (cost_arm.cuh:156-158)
//
// with the reference's names, argument order and meaning.  Calling convention (the reference's): every function is
// called by ALL threads of the thread block with block-uniform pointer arguments into shared or global memory; it strides its own
// loops with singleLoopVals / doubleLoopVals, may declare __shared__ statics, may call hd__syncthreads(), and leaves its results in
// the caller's buffers.  The cooperating group is the thread block, and here a thread block is ONE warp (32 x 1 threads; the
// reference launches 8 x 7), so hd__syncthreads() costs a warp barrier and many independent knots / trajectories share an SM.
// Results must not depend on the block shape -- each output element is produced by one thread -- and for the reference's plants
// they do not.
//
// Compile-time constants of config.cuh that became run-time configuration (pddp_config) are visible to plant code under their
// reference names as expressions: NUM_TIME_STEPS (the horizon of the handle whose kernel is running).  dt, the cost weights and the model arrays arrive as arguments, as in the reference.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pddp_plugin {
// horizon of the running kernel: set by every plug-in kernel before it calls plant code (a static __shared__ int per block)
__device__ __forceinline__ int &rt_num_time_steps(){ __shared__ int v; return v; }
}
namespace pddp_plugin { static int host_num_time_steps = 0; }    // host instantiations of plant code (tests only): set by the caller
#ifdef __CUDA_ARCH__
#define NUM_TIME_STEPS (::pddp_plugin::rt_num_time_steps())
#else
#define NUM_TIME_STEPS (::pddp_plugin::host_num_time_steps)
#endif
#ifndef EE_COST
#define EE_COST 0
#endif

/* loop bounds and barrier of the cooperating group (utils/cudaUtils.h:65-88) */
__host__ __device__ __forceinline__ void doubleLoopVals(int *starty, int *dy, int *startx, int *dx){
#ifdef __CUDA_ARCH__
    *starty = threadIdx.y; *dy = blockDim.y; *startx = threadIdx.x; *dx = blockDim.x;
#else
    *starty = 0; *dy = 1; *startx = 0; *dx = 1;
#endif
}
__host__ __device__ __forceinline__ void singleLoopVals(int *start, int *delta){
#ifdef __CUDA_ARCH__
    *start = threadIdx.x + threadIdx.y*blockDim.x; *delta = blockDim.x*blockDim.y;
#else
    *start = 0; *delta = 1;
#endif
}
__host__ __device__ __forceinline__ void hd__syncthreads(){
#ifdef __CUDA_ARCH__
    __syncthreads();
#endif
}
