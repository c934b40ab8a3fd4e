// plant_ptx_wrap.cu -- TEST INFRASTRUCTURE (input of oracle/tools/ptx2c.py).  Straight-line instantiations of a plug-in plant's functions
// and of the integrators' gradient assembly, compiled to PTX only: sin / cos / pow are left as calls to opaque functions so that the
// arithmetic around them stays in one basic block and can be read off instruction by instruction.
#define sin pddp_opq_sin
#define cos pddp_opq_cos
#define pow pddp_opq_pow
#ifdef PDDP_WRAP_CONTEXT
#define OPQ_ATTR            /* like the inlined library code next to shared-memory stores: not shared across them */
#else
#define OPQ_ATTR __attribute__((const))
#endif
extern __device__ __noinline__ float pddp_opq_sin(float) OPQ_ATTR;
extern __device__ __noinline__ float pddp_opq_cos(float) OPQ_ATTR;
extern __device__ __noinline__ double pddp_opq_sin(double) OPQ_ATTR;
extern __device__ __noinline__ double pddp_opq_cos(double) OPQ_ATTR;
extern __device__ __noinline__ double pddp_opq_pow(double, int) OPQ_ATTR;
extern __device__ __noinline__ double pddp_opq_pow(float, int) OPQ_ATTR;
#include <cuda_runtime.h>
// the group is one thread here: loops run serially, exactly the arithmetic one thread of the real kernels performs
#ifdef PDDP_WRAP_RUNTIME_LOOPS       // loop bounds the compiler cannot see through (the kernels' are thread indices): keeps the (row, column) selects dynamic
__device__ int pddp_loop_vals[4];
__host__ __device__ __forceinline__ void doubleLoopVals(int *starty, int *dy, int *startx, int *dx){ *starty = pddp_loop_vals[0]; *dy = pddp_loop_vals[1]; *startx = pddp_loop_vals[2]; *dx = pddp_loop_vals[3]; }
#else
__host__ __device__ __forceinline__ void doubleLoopVals(int *starty, int *dy, int *startx, int *dx){ *starty = 0; *dy = 1; *startx = 0; *dx = 1; }
#endif
__host__ __device__ __forceinline__ void singleLoopVals(int *start, int *delta){ *start = 0; *delta = 1; }
__host__ __device__ __forceinline__ void hd__syncthreads(){ }
#define NUM_TIME_STEPS 1000000
#ifndef PDDP_WRAP_CONTEXT
#define threadIdx pddp_fake_tid
struct { int x, y; } __device__ const pddp_fake_tid = {0, 0};
#endif
#include PDDP_PLANT_HEADER
#undef threadIdx
extern "C" __global__ void k_dynamics(float *__restrict__ x, float *__restrict__ u, float *__restrict__ qdd){ dynamics<float>(qdd, x, u, nullptr, nullptr); }
// with s_qdd, as every caller in the solver passes it (_integratorGradient): the compiler then shares sub-expressions between the
// acceleration and its gradient, which changes what it can fuse
extern "C" __global__ void k_gradient(float *__restrict__ x, float *__restrict__ u, float *__restrict__ qdd, float *__restrict__ dqdd){ __builtin_assume(qdd != nullptr); dynamicsGradient<float>(dqdd, qdd, x, u, nullptr, nullptr); }
#ifdef PDDP_WRAP_INTEGRATORS
#include "../../parallel-ddp_b200/csrc/plugin/integrators.cuh"
extern "C" __global__ void k_step1(float *xn, float *x, float *u, float *qdd, float dt){ _integrator<float,1>(xn, x, u, qdd, nullptr, nullptr, dt); }
extern "C" __global__ void k_step2(float *xn, float *x, float *u, float *qdd, float dt){ _integrator<float,2>(xn, x, u, qdd, nullptr, nullptr, dt); }
extern "C" __global__ void k_step3(float *xn, float *x, float *u, float *qdd, float dt){ _integrator<float,3>(xn, x, u, qdd, nullptr, nullptr, dt); }
extern "C" __global__ void k_grad1(float *AB, float *x, float *u, float *qdd, float *dqdd, float dt){ _integratorGradient<float,1>(AB, x, u, qdd, dqdd, nullptr, nullptr, dt, STATE_SIZE); }
extern "C" __global__ void k_grad2(float *AB, float *x, float *u, float *qdd, float *dqdd, float dt){ _integratorGradient<float,2>(AB, x, u, qdd, dqdd, nullptr, nullptr, dt, STATE_SIZE); }
extern "C" __global__ void k_grad3(float *AB, float *x, float *u, float *qdd, float *dqdd, float dt){ _integratorGradient<float,3>(AB, x, u, qdd, dqdd, nullptr, nullptr, dt, STATE_SIZE); }
#endif
#ifdef PDDP_WRAP_CONTEXT
// dynamicsGradient in the memory context of the solver's kernels: shared-memory arguments, the acceleration requested with it, one thread doing
// the work.  (Which products the compiler fuses depends on what it may share between the two halves, i.e. on this context.)
extern "C" __global__ void k_gradient_ctx(const float *x, const float *u, float *qdd, float *dqdd){
    __shared__ float s_x[STATE_SIZE], s_u[CONTROL_SIZE], s_qdd[NUM_POS], s_dqdd[NUM_POS*(STATE_SIZE+CONTROL_SIZE)];
    const int l = threadIdx.x;
    if (l < STATE_SIZE){ s_x[l] = x[l]; } if (l < CONTROL_SIZE){ s_u[l] = u[l]; }
    __syncthreads();
    dynamicsGradient<float>(s_dqdd, s_qdd, s_x, s_u, nullptr, nullptr);
    __syncthreads();
    if (l < NUM_POS){ qdd[l] = s_qdd[l]; }
    for (int i = l; i < NUM_POS*(STATE_SIZE+CONTROL_SIZE); i += blockDim.x){ dqdd[i] = s_dqdd[i]; }
}
#endif
