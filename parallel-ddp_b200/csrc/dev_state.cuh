// dev_state.cuh -- device-side state of a batch of problems and the pieces shared by the Kuka kernels (kernels.cuh) and the
// plug-in kernels (plugin/plugin_kernels.cuh): TMA / mbarrier helpers, the forward sweep (plant independent, templated on the
// state size) and the shift helpers of the receding-horizon load step.  Everything here is a template or an inline device
// function, so the header can be part of several translation units.
#pragma once
#include "pddp_math.cuh"

namespace pddp {

// Knot strides (floats) of AB, H and g in HBM: the 14x21 / 21x21 / 21 tiles padded to 16-byte multiples so that one
// cp.async.bulk (TMA) moves a whole tile; the tile interior keeps the reference's column-major layout.
constexpr int AB_STRIDE = 296, H_STRIDE = 444, G_STRIDE = 24;

struct DevState {
    // sizes
    int B, N, A, M, n, m, max_iter;
    int npos;                      // NUM_POS
    int ab_stride, h_stride, g_stride;   // knot strides (floats) of AB, H, g: the padded Kuka strides above, dense for the plug-in plants
    int integrator;                // INTEGRATOR 1 Euler | 2 Midpoint | 3 RK3 (plug-in plants; the Kuka kernels are Euler)
    int iter_cap;                  // iteration limit of this solve (<= max_iter; runiLQR_MPC_GPU's max_iter argument)
    float dt, tol_cost, two_tol;
    float rho_min, rho_max, rho_factor, inv_rho_factor, exp_red_min, exp_red_max, max_defect;
    float Q1, Q2, R, QF1, QF2;
    int use_limits; float Q_PL, Q_VL, R_TL;      // USE_LIMITS_FLAG (plants/cost_arm.cuh:11-94), joint-space cost of the arm
    int smooth_abs; float sa_alpha, sa_alpha2;   // USE_SMOOTH_ABS (plants/cost_arm.cuh:218-220,242-252), end-effector cost
    // model constants (device)
    const float *I, *Tbody, *alpha;
    // trajectories
    float *x, *u, *d;              // candidates [B][A][N][n|m|n]
    float *xp, *xp2, *up, *dp;     // accepted / previous accepted [B][N][.]
    float *AB, *H, *g;
    float *Pbuf[2], *pbuf[2];      // ping-pong per problem: the backward pass of iteration `iter` writes P = Pbuf[iter & 1] and seeds its
                                   // blocks from Pp = Pbuf[(iter & 1) ^ 1].  iter[b] only advances when problem b goes on to another
                                   // iteration, which is exactly when the reference copies P to Pp (nextIterationSetupGPU)
    float *KT, *du, *ApBK, *Bdu;
    float *xGoal;                  // [B][n]
    float *costk;                  // per-knot costs [B][A][N]
    float *J, *dT, *dJexp;         // [B][A], [B][A], [B][2M]
    // per-problem solver scalars
    float *rho, *drho, *prevJ, *dJ, *z;
    int *iter, *alphaIndex, *ignore_defect, *done, *accepted, *final_src;
    float *Jout; int *alphaOut;    // [B][max_iter+1]
    int *n_active;                 // [1]
    float grav;                    // gravity constant of the plant (9.81; 0 in the reference's MPC_MODE)
    int skip_unchanged;            // opt-in: skip the gradient refresh of a problem whose line search was rejected
    int rolled_out;                // this solve started with loadVarsGPU's forward rollout
    long long *dbg;                // [4096] stage clocks of CTA 0 (only written by -DPDDP_BP_TRACE builds)
    // end-effector cost (EE_COST 1, plants/cost_arm.cuh:204-389): xGoal[b][0..5] is the goal pose, costk[b][a][0..M-1] the
    // simulation's per-interval cost partials (fpHelpers.cuh:299)
    int ee;
    const int *cost_shift;         // [B] or null: finalCostShift of runiLQR_MPC_GPU (MPCHelpers.cuh:876) -- the pose terms take their final weights from knot N-1-shift on
    int *init_knot;                // [B] EE_COST quirk of the receding-horizon path, see select_kernel mode 1 (0 everywhere else)
    const float *xTarget;          // [B][n] or null: the nominal-state terms measure x from it (receding-horizon path, MPCHelpers.cuh:900)
    float Q_EE1, Q_EE2, QF_EE1, QF_EE2, R_EE, Q_xdEE, QF_xdEE, Q_xEE, QF_xEE;
    // step-size sharding of the line search over ranks (pddp_alpha_shard_init): this rank sweeps / simulates candidates
    // a_first .. a_first + a_cnt - 1 (all of them without sharding); the per-candidate (J, defect) pairs travel through xchg_send ->
    // ncclAllGather -> xchg_recv [rank][B][a_cnt][2], the accepted candidate through acc_buf [B][N*(2n+m)] (ncclAllReduce of its bits)
    int a_first, a_cnt;
    float *xchg_send, *xchg_recv; int *acc_buf;
};

// ------------------------------------------------------------------------------------------------------------------
// TMA (1-D bulk copy) + mbarrier helpers
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p){ return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count){ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes){
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar){
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity){
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// polling wait with back-off for the producer warp: a tight try_wait loop would keep its scheduler's shared-memory queue busy
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long *bar, unsigned parity){
    unsigned ok = 0;
    while (true){
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok){ break; }
        __nanosleep(400);
    }
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar){ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory"); }
// named barrier over the first `count` threads' warps (the producer warp of a CTA does not take part)
__device__ __forceinline__ void bar_sync(int id, int count){ asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory"); }

// ------------------------------------------------------------------------------------------------------------------
// forward sweep: x_a[k+1] = xp[k+1] + ( -alpha_a Bdu_k + (A-BK)_k (x_a[k]-xp[k]) + [boundary] d_k )
// grid = B*splits CTAs of 32*A/splits threads; the whole (A-BK) sequence of the problem is staged in shared memory once.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SWEEP_CH = 16;         // knots per slice
constexpr int SWEEP_SLOTS = 4;       // slices resident in shared memory
template <int n>
struct __align__(16) SweepSlot { float A[SWEEP_CH*n*n], B[SWEEP_CH*n], xp[SWEEP_CH*n], d[SWEEP_CH*n]; };
template <int n>
__global__ void sweep_kernel(DevState S, int splits, int b0){
    extern __shared__ __align__(16) unsigned char sw_raw[];
    __shared__ unsigned long long full[SWEEP_SLOTS];
    SweepSlot<n> *slots = reinterpret_cast<SweepSlot<n>*>(sw_raw);
    // `splits` CTAs share one problem (each takes A/splits step sizes) so that a small batch still covers the SMs
    const int b = b0 + blockIdx.x / splits, a0 = S.a_first + (blockIdx.x % splits)*(S.a_cnt / splits), N = S.N, NBF = N / S.M;
    if (S.done[b]){ return; }
    // The problem's sequence ((A - BK), B du, xp, d per knot) streams through a ring of SWEEP_SLOTS slices of SWEEP_CH knots,
    // each filled by four TMA bulk copies on its own mbarrier.  The recursion starts as soon as the first slice has landed; a
    // slice is refilled (CTA barrier, then one thread issues) once every warp has moved two slices past it, so any horizon
    // N <= 1024 fits in 60 KB of shared memory.
    const int nslices = N / SWEEP_CH;
    const float *gA = S.ApBK + (size_t)b*N*n*n, *gB = S.Bdu + (size_t)b*N*n, *gxp = S.xp + (size_t)b*N*n, *gd = S.dp + (size_t)b*N*n;
    auto issue = [&](int c){
        SweepSlot<n> &sl = slots[c % SWEEP_SLOTS]; unsigned long long *bar = &full[c % SWEEP_SLOTS];
        constexpr unsigned bytesA = SWEEP_CH*n*n*4, bytesV = SWEEP_CH*n*4;
        mbar_expect_tx(bar, bytesA + 3*bytesV);
        tma_load_1d(sl.A, gA + (size_t)c*SWEEP_CH*n*n, bytesA, bar);
        tma_load_1d(sl.B, gB + (size_t)c*SWEEP_CH*n, bytesV, bar);
        tma_load_1d(sl.xp, gxp + (size_t)c*SWEEP_CH*n, bytesV, bar);
        tma_load_1d(sl.d, gd + (size_t)c*SWEEP_CH*n, bytesV, bar);
    };
    if (threadIdx.x == 0){
        for (int c = 0; c < SWEEP_SLOTS; c++){ mbar_init(&full[c], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int c = 0; c < SWEEP_SLOTS && c < nslices; c++){ issue(c); }
    }
    __syncthreads();
    const int a = a0 + (threadIdx.x >> 5), l = threadIdx.x & 31, lc = l < n ? l : n - 1;   // blockDim = 32 * (a_cnt / splits): every warp has a step size
    const float alpha = S.alpha[a];
    float *gx = S.x + ((size_t)b*S.A + a)*N*n;
    mbar_wait(&full[0], 0);
    // lane l < n carries component l (the other lanes replay component n-1 and store nothing).  The recursion is one dependent chain
    // per knot -- (x - xp) -> broadcast -> n fused multiply-adds -> three additions -- so everything that does not depend on x_a[k]
    // (the column entries of A - BK, B du, the defect, xp[k+1]) is fetched BEFORE the chain, and xp[k] is the register that held
    // xp[k+1] one knot earlier: no shared-memory load sits on the chain.
    float xpk = slots[0].xp[lc];
    float xk = xpk;                                  // x_a[0] = xp[0]
    if (l < n){ gx[l] = xk; }
    int to_boundary = NBF;                   // steps until k+1 is a shooting-interval boundary
    for (int c = 0; c < nslices; c++){
        const SweepSlot<n> &sl = slots[c % SWEEP_SLOTS];
        const bool last = (c == nslices - 1);
        // slice c-2 is behind every warp of the CTA once all of them are here: hand its slot to slice c-2+SWEEP_SLOTS
        if (c >= 2 && c - 2 + SWEEP_SLOTS < nslices){
            __syncthreads();
            if (threadIdx.x == 0){ issue(c - 2 + SWEEP_SLOTS); }
        }
        const int kend = last ? SWEEP_CH - 1 : SWEEP_CH;
        for (int kk = 0; kk < kend; kk++){
            const int k = c*SWEEP_CH + kk;
            // x_p[k+1] of the last knot of a slice lives in the next slice
            const float *xpn;
            if (kk == SWEEP_CH - 1){ mbar_wait(&full[(c+1) % SWEEP_SLOTS], ((c+1) / SWEEP_SLOTS) & 1); xpn = slots[(c+1) % SWEEP_SLOTS].xp; }
            else { xpn = sl.xp + (kk+1)*n; }
            const float *Ak = sl.A + kk*n*n + lc;
            float ar[n];
            #pragma unroll
            for (int i = 0; i < n; i++){ ar[i] = Ak[n*i]; }
            const float Bk = sl.B[kk*n + lc], dk = sl.d[kk*n + lc], xn = xpn[lc];
            const float dx = SUB(xk, xpk);
            float dxs[n];
            #pragma unroll
            for (int i = 0; i < n; i++){ dxs[i] = __shfl_sync(FULL, dx, i); }
            float val = 0.f;
            #pragma unroll
            for (int i = 0; i < n; i++){ val = FMA(ar[i], dxs[i], val); }
            const bool onb = (--to_boundary == 0);
            if (onb){ to_boundary = NBF; }
            const float tt = ADD(FMA(-alpha, Bk, val), onb ? dk : 0.f);
            xk = ADD(xn, tt); xpk = xn;
            if (l < n){ gx[(k+1)*n + l] = xk; }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// receding horizon (MPCHelpers.cuh): load step of runiLQR_MPC_GPU, one CTA per problem
// ------------------------------------------------------------------------------------------------------------------
struct MpcState {
    float *cx, *cu, *cd;           // [B][N][.] the current plan with its defects (h_d_x/u/d[alphaIndex] of the reference)
    float *x_old, *u_old, *KT_old; // the shifted previous plan, restored when a solve takes no step
    const float *xActual;          // [B][n]
    const int *shift, *clear;      // [B]
};
// shiftAndCopy (MPCHelpers.cuh:425-453): A[k] <- A[min(shift + k, dimN-1)] for k < dimN-1 (zero past the end with `flag`), B likewise.
// In place, without the reference's scratch copy, by `nthr` threads of the CTA (this thread is number `tid` of them; they meet at named
// barrier `bar`): a slab of 8*nthr entries is read into registers, the barrier passes, the slab is written.  Sources lie at or ahead of their
// destinations and slabs go front to back, so a slab's reads can only meet the writes of the same slab -- which the barrier orders;
// the clamped source (the last knot) is never written.  Eight loads in flight per thread instead of one.
__device__ __forceinline__ void mpc_bar(int bar, int nthr){ asm volatile("bar.sync %0, %1;" :: "r"(bar), "r"(nthr) : "memory"); }
__device__ __forceinline__ void mpc_shift_part(float *A, int shift, int sz, int dimN, bool flag, float *B, int tid, int nthr, int bar){
    constexpr int R = 8;
    const int cnt = (dimN - 1)*sz;
    for (int base = 0; base < cnt; base += R*nthr){
        float v[R];
        #pragma unroll
        for (int r = 0; r < R; r++){
            const int idx = base + r*nthr + tid; v[r] = 0.f;
            if (idx < cnt){
                const int k = idx / sz, i = idx - k*sz; int ksrc = shift + k; if (ksrc > dimN - 1){ ksrc = dimN - 1; }
                if (!(flag && ksrc >= dimN - 1)){ v[r] = A[(size_t)ksrc*sz + i]; }
            }
        }
        mpc_bar(bar, nthr);
        #pragma unroll
        for (int r = 0; r < R; r++){ const int idx = base + r*nthr + tid; if (idx < cnt){ A[idx] = v[r]; if (B){ B[idx] = v[r]; } } }
    }
}
// slices of the plain loops for a subset of the CTA's threads
__device__ __forceinline__ void mpc_swap_part(float *A, float *B, int cnt, int tid, int nthr){ for (int i = tid; i < cnt; i += nthr){ const float v = A[i]; A[i] = B[i]; B[i] = v; } }
__device__ __forceinline__ void mpc_zero_part(float *A, int cnt, int tid, int nthr){ for (int i = tid; i < cnt; i += nthr){ A[i] = 0.f; } }
__device__ __forceinline__ void mpc_copy_part(float *D, const float *A, int cnt, int tid, int nthr){ for (int i = tid; i < cnt; i += nthr){ D[i] = A[i]; } }
__device__ __forceinline__ void mpc_swap(float *A, float *B, int cnt){ for (int i = threadIdx.x; i < cnt; i += blockDim.x){ const float v = A[i]; A[i] = B[i]; B[i] = v; } }
__device__ __forceinline__ void mpc_zero(float *A, int cnt){ for (int i = threadIdx.x; i < cnt; i += blockDim.x){ A[i] = 0.f; } }
__device__ __forceinline__ void mpc_copy(float *D, const float *A, int cnt){ for (int i = threadIdx.x; i < cnt; i += blockDim.x){ D[i] = A[i]; } }

} // namespace pddp
