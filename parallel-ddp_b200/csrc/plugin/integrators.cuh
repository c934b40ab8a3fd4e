// plugin/integrators.cuh -- one-step integrators and their analytic gradients on top of a plant's dynamics / dynamicsGradient
// (the reference's utils/integrators.cuh: Euler :24-53, Midpoint :56-121, RK3 :123-233), for the plug-in calling convention of
// pddp_plugin.cuh (all threads of the block call, block-uniform pointers).  The reference picks the rule with the INTEGRATOR
// macro; here it is the second template argument, because pddp_config.integrator is a run-time choice and the plug-in kernels are
// instantiated for all three.  _integrator<T>(...) / _integratorGradient<T>(...) with the reference's argument lists are kept as
// the INTEGRATOR-macro forms at the end.
//
//   x = [q ; qd],  xd = [qd ; qdd(x,u)],  d xd / d(x,u) = [0 I 0 ; dqdd]      (dqdd: NUM_POS x (STATE_SIZE + CONTROL_SIZE), column-major)
//
// The RK3 gradient evaluates the dynamics at the stage states exactly as the reference writes them (integrators.cuh:181-182,
// 190-191: the velocity half of both stage states starts from q, not qd) -- these expressions define its AB matrices.
#pragma once
#define DIM_AB_r STATE_SIZE
#define DIM_AB_c (STATE_SIZE + CONTROL_SIZE)

// entry (r, c) of d xd / d(x,u) given dqdd (integrators.cuh:15-17)
template <typename T>
__host__ __device__ __forceinline__
T dqdd2dxd(T *dqdd, int r, int c){ return r < NUM_POS ? static_cast<T>(r + NUM_POS == c ? 1 : 0) : dqdd[(c-1)*NUM_POS + r]; }

namespace pddp_integ {
// y = a + h*b on both halves of the state: positions advance with `vel`, velocities with `acc`
template <typename T>
__host__ __device__ __forceinline__
void advance(T *y, const T *a, T h, const T *vel, const T *acc){
    int first, step; singleLoopVals(&first, &step);
    for (int i = first; i < NUM_POS; i += step){
        y[i] = a[i] + h*vel[i];
        y[i+NUM_POS] = a[i+NUM_POS] + h*acc[i];
    }
}
}

template <typename T, int INTEG>
__host__ __device__ __forceinline__
void _integrator(T *s_xkp1, T *s_x, T *s_u, T *s_qdd, T *d_I, T *d_Tbody, T dt, T *s_eePos = nullptr, T *s_eeVel = nullptr){
    static_assert(INTEG >= 1 && INTEG <= 3, "INTEGRATOR 1 Euler | 2 Midpoint | 3 RK3");
    int first, step; singleLoopVals(&first, &step);
    if (INTEG == 1){
        dynamics<T>(s_qdd, s_x, s_u, d_I, d_Tbody, s_eePos, 1, s_eeVel);
        hd__syncthreads();
        pddp_integ::advance<T>(s_xkp1, s_x, dt, s_x + NUM_POS, s_qdd);
    }
    else if (INTEG == 2){
        #ifdef __CUDA_ARCH__
        __shared__ T s_mid[STATE_SIZE];
        #else
        T s_mid[STATE_SIZE];
        #endif
        dynamics<T>(s_qdd, s_x, s_u, d_I, d_Tbody, s_eePos, 1, s_eeVel);
        hd__syncthreads();
        for (int i = first; i < NUM_POS; i += step){
            s_mid[i] = s_x[i] + static_cast<T>(0.5)*dt*s_x[i+NUM_POS];
            s_mid[i+NUM_POS] = s_x[i+NUM_POS] + static_cast<T>(0.5)*dt*s_qdd[i];
        }
        hd__syncthreads();
        dynamics<T>(s_qdd, s_mid, s_u, d_I, d_Tbody);
        hd__syncthreads();
        // the positions advance with the INITIAL velocity (integrators.cuh:78), the velocities with the midpoint acceleration
        pddp_integ::advance<T>(s_xkp1, s_x, dt, s_x + NUM_POS, s_qdd);
    }
    else {
        #ifdef __CUDA_ARCH__
        __shared__ T s_xb[STATE_SIZE]; __shared__ T s_xc[STATE_SIZE]; __shared__ T s_ab[NUM_POS]; __shared__ T s_ac[NUM_POS];
        #else
        T s_xb[STATE_SIZE], s_xc[STATE_SIZE], s_ab[NUM_POS], s_ac[NUM_POS];
        #endif
        dynamics<T>(s_qdd, s_x, s_u, d_I, d_Tbody, s_eePos, 1, s_eeVel);
        hd__syncthreads();
        for (int i = first; i < NUM_POS; i += step){
            s_xb[i] = s_x[i] + static_cast<T>(0.5)*dt*s_x[i+NUM_POS];
            s_xb[i+NUM_POS] = s_x[i+NUM_POS] + static_cast<T>(0.5)*dt*s_qdd[i];
        }
        hd__syncthreads();
        dynamics<T>(s_ab, s_xb, s_u, d_I, d_Tbody);
        hd__syncthreads();
        for (int i = first; i < NUM_POS; i += step){
            s_xc[i] = s_x[i] + dt*(static_cast<T>(2)*s_xb[i+NUM_POS] - s_x[i+NUM_POS]);
            s_xc[i+NUM_POS] = s_x[i+NUM_POS] + dt*(static_cast<T>(2)*s_ab[i] - s_qdd[i]);
        }
        hd__syncthreads();
        dynamics<T>(s_ac, s_xc, s_u, d_I, d_Tbody);
        hd__syncthreads();
        for (int i = first; i < NUM_POS; i += step){
            s_xkp1[i] = s_x[i] + (dt/static_cast<T>(6))*(s_x[i+NUM_POS] + static_cast<T>(4)*s_xb[i+NUM_POS] + s_xc[i+NUM_POS]);
            s_xkp1[i+NUM_POS] = s_x[i+NUM_POS] + (dt/static_cast<T>(6))*(s_qdd[i] + static_cast<T>(4)*s_ab[i] + s_ac[i]);
        }
    }
}

// ABk (STATE_SIZE x (STATE_SIZE + CONTROL_SIZE), leading dimension ld_AB) = d x_{k+1} / d (x_k, u_k); s_qdd, s_dqdd receive the
// acceleration and its gradient at (s_x, s_u)
template <typename T, int INTEG>
__host__ __device__ __forceinline__
void _integratorGradient(T *ABk, T *s_x, T *s_u, T *s_qdd, T *s_dqdd, T *d_I, T *d_Tbody, T dt, int ld_AB){
    static_assert(INTEG >= 1 && INTEG <= 3, "INTEGRATOR 1 Euler | 2 Midpoint | 3 RK3");
    int first, step; singleLoopVals(&first, &step);
    int cy, dy, rx, dx; doubleLoopVals(&cy, &dy, &rx, &dx);
    constexpr int ND = NUM_POS*(STATE_SIZE + CONTROL_SIZE);
    if (INTEG == 1){
        dynamicsGradient<T>(s_dqdd, s_qdd, s_x, s_u, d_I, d_Tbody);
        hd__syncthreads();
        for (int c = cy; c < DIM_AB_c; c += dy){ for (int r = rx; r < DIM_AB_r; r += dx){
            ABk[c*ld_AB + r] = static_cast<T>(c == r ? 1 : 0) + dt*dqdd2dxd(s_dqdd, r, c);
        }}
    }
    else if (INTEG == 2){
        #ifdef __CUDA_ARCH__
        __shared__ T s_mid[STATE_SIZE]; __shared__ T s_am[NUM_POS]; __shared__ T s_dm[ND];
        #else
        T s_mid[STATE_SIZE], s_am[NUM_POS], s_dm[ND];
        #endif
        dynamicsGradient<T>(s_dqdd, s_qdd, s_x, s_u, d_I, d_Tbody);
        hd__syncthreads();
        for (int i = first; i < NUM_POS; i += step){
            s_mid[i] = s_x[i] + static_cast<T>(0.5)*dt*s_x[i+NUM_POS];
            s_mid[i+NUM_POS] = s_x[i+NUM_POS] + static_cast<T>(0.5)*dt*s_qdd[i];
        }
        hd__syncthreads();
        dynamicsGradient<T>(s_dm, s_am, s_mid, s_u, d_I, d_Tbody);
        hd__syncthreads();
        // chain rule through the half step: AB = A_mid (I + h/2 [A B]_start) + [0, h/2 B_mid]   (integrators.cuh:104-119)
        for (int c = cy; c < DIM_AB_c; c += dy){ for (int r = rx; r < DIM_AB_r; r += dx){
            T acc = 0;
            for (int i = 0; i < DIM_AB_r; i++){
                T left = static_cast<T>(r == i ? 1 : 0) + static_cast<T>(0.5)*dt*dqdd2dxd(s_dm, r, i);
                T right = static_cast<T>(c == i ? 1 : 0) + static_cast<T>(0.5)*dt*dqdd2dxd(s_dqdd, i, c);
                acc += left * right;
            }
            ABk[c*ld_AB + r] = acc + (c < STATE_SIZE ? static_cast<T>(0) : static_cast<T>(0.5)*dt*dqdd2dxd(s_dm, r, c));
        }}
    }
    else {
        #ifdef __CUDA_ARCH__
        __shared__ T s_G1[DIM_AB_r*DIM_AB_c]; __shared__ T s_G2[DIM_AB_r*DIM_AB_c];
        __shared__ T s_xb[STATE_SIZE]; __shared__ T s_xc[STATE_SIZE];
        __shared__ T s_ab[NUM_POS]; __shared__ T s_ac[NUM_POS]; __shared__ T s_db[ND]; __shared__ T s_dc[ND];
        #else
        T s_G1[DIM_AB_r*DIM_AB_c], s_G2[DIM_AB_r*DIM_AB_c], s_xb[STATE_SIZE], s_xc[STATE_SIZE], s_ab[NUM_POS], s_ac[NUM_POS], s_db[ND], s_dc[ND];
        #endif
        dynamicsGradient<T>(s_dqdd, s_qdd, s_x, s_u, d_I, d_Tbody);
        hd__syncthreads();
        // stage states as the reference's gradient forms them (integrators.cuh:181-182): both halves start from q
        for (int i = first; i < NUM_POS; i += step){
            s_xb[i] = s_x[i] + static_cast<T>(0.5)*dt*s_x[i+NUM_POS];
            s_xb[i+NUM_POS] = s_x[i] + static_cast<T>(0.5)*dt*s_qdd[i];
        }
        hd__syncthreads();
        dynamicsGradient<T>(s_db, s_ab, s_xb, s_u, d_I, d_Tbody);
        hd__syncthreads();
        for (int i = first; i < NUM_POS; i += step){          // integrators.cuh:190-191
            s_xc[i] = s_x[i] + dt*s_x[i+NUM_POS] + static_cast<T>(2)*dt*s_xb[i+NUM_POS];
            s_xc[i+NUM_POS] = s_x[i] + dt*s_qdd[i] + static_cast<T>(2)*dt*s_ab[i];
        }
        hd__syncthreads();
        dynamicsGradient<T>(s_dc, s_ac, s_xc, s_u, d_I, d_Tbody);
        hd__syncthreads();
        // G1 = [0, B_b] + A_b ([I, 0] + (h/2) [A B]_start)        (integrators.cuh:196-207)
        for (int c = cy; c < DIM_AB_c; c += dy){ for (int r = rx; r < DIM_AB_r; r += dx){
            T acc = 0;
            #pragma unroll
            for (int i = 0; i < DIM_AB_r; i++){ acc += dqdd2dxd(s_db, r, i)*(static_cast<T>(0.5)*dt*dqdd2dxd(s_dqdd, i, c) + static_cast<T>(c == i ? 1 : 0)); }
            s_G1[r + DIM_AB_r*c] = acc + (c < STATE_SIZE ? static_cast<T>(0) : dqdd2dxd(s_db, r, c));
        }}
        hd__syncthreads();
        // G2 = [0, B_c] + A_c ([I, 0] + 2h G1 - h [A B]_start)       (integrators.cuh:209-221)
        for (int c = cy; c < DIM_AB_c; c += dy){ for (int r = rx; r < DIM_AB_r; r += dx){
            T acc = 0;
            #pragma unroll
            for (int i = 0; i < DIM_AB_r; i++){ acc += dqdd2dxd(s_dc, r, i)*(static_cast<T>(2)*dt*s_G1[c*DIM_AB_r + i] - dt*dqdd2dxd(s_dqdd, i, c) + static_cast<T>(c == i ? 1 : 0)); }
            s_G2[r + DIM_AB_r*c] = acc + (c < STATE_SIZE ? static_cast<T>(0) : dqdd2dxd(s_dc, r, c));
        }}
        hd__syncthreads();
        // AB = [I, 0] + (h/6) [A B]_start + (2h/3) G1 + (h/6) G2      (integrators.cuh:223-231)
        for (int c = cy; c < DIM_AB_c; c += dy){ for (int r = rx; r < DIM_AB_r; r += dx){
            ABk[r + ld_AB*c] = (dt/static_cast<T>(6))*dqdd2dxd(s_dqdd, r, c) + (static_cast<T>(2)*dt/static_cast<T>(3))*s_G1[r + DIM_AB_r*c] +
                               (dt/static_cast<T>(6))*s_G2[r + DIM_AB_r*c] + static_cast<T>(r == c ? 1 : 0);
        }}
    }
}

#ifdef INTEGRATOR
// the reference's forms: the rule comes from the INTEGRATOR macro (config.cuh:78-80)
template <typename T> __host__ __device__ __forceinline__
void _integrator(T *s_xkp1, T *s_x, T *s_u, T *s_qdd, T *d_I, T *d_Tbody, T dt, T *s_eePos = nullptr, T *s_eeVel = nullptr){ _integrator<T, INTEGRATOR>(s_xkp1, s_x, s_u, s_qdd, d_I, d_Tbody, dt, s_eePos, s_eeVel); }
template <typename T> __host__ __device__ __forceinline__
void _integratorGradient(T *ABk, T *s_x, T *s_u, T *s_qdd, T *s_dqdd, T *d_I, T *d_Tbody, T dt, int ld_AB){ _integratorGradient<T, INTEGRATOR>(ABk, s_x, s_u, s_qdd, s_dqdd, d_I, d_Tbody, dt, ld_AB); }
#endif
