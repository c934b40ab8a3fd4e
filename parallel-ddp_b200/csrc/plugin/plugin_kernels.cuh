// plugin_kernels.cuh -- the solver kernels of a plug-in plant (any STATE_SIZE / CONTROL_SIZE, any of the three integrators).
//
// Included by csrc/plant_tu.cu AFTER the plant header and plugin/integrators.cuh: everything here that touches the plant calls it
// through the reference's plug-in names (dynamics, dynamicsGradient via _integrator / _integratorGradient, costFunc, costGrad), with a
// thread block of one warp as the cooperating group (pddp_plugin.cuh).  The solver arithmetic around those calls is written with
// explicit rounding intrinsics (pddp_math.cuh), so this translation unit can be compiled with nvcc's default contraction -- the
// plant author's code is then compiled exactly as the reference compiles its plants -- without changing the solver's roundings.
//
//   plug_bp_kernel      backPassKern + the rho retry of backwardPassGPU      bpHelpers.cuh:337-420,484-517 (incl. :96-188 dim1 / dim4)
//   sweep_kernel<n>     forwardSweepKern                                     dev_state.cuh (shared with the Kuka path)
//   plug_sim_kernel     forwardSimKern + computeControlKT + per-knot costs   fpHelpers.cuh:200-301,132-152
//   plug_nis_kernel     hand-over + integratorGradientKern + costGradientHessianKern   nisInitHelpers.cuh:203-221,44-93,245-279
//   plug_mpc_*          loadVarsGPU_MPC incl. rolloutMPC                     MPCHelpers.cuh:602-657,524-556
// Line search / accept-reject (select_kernel), reset and store kernels are plant independent and live in the main library.
//
// This file has no include guard and no namespace of its own: plant_tu.cu includes it inside the plant's namespace (after
// dev_state.cuh and include/pddp_plant.h), so that several plants can live in one library.

constexpr int PN = STATE_SIZE, PM = CONTROL_SIZE, PNP = NUM_POS, PNM = STATE_SIZE + CONTROL_SIZE;
static_assert(PN == 2*PNP, "x = [q ; qd]");
static_assert(PN <= 32 && PM <= 32, "one lane per state / control entry");
#define PLUG_BLOCK dim3(32, 1, 1)
#ifndef PDDP_MAX_RHO_RETRIES
#define PDDP_MAX_RHO_RETRIES 200      // the reference retries for ever (IGNORE_MAX_ROX_EXIT 1, bpHelpers.cuh:504): a device kernel must end
#endif

__device__ __forceinline__ int plug_tid(){ return threadIdx.x + threadIdx.y*blockDim.x; }
__device__ __forceinline__ int plug_nthreads(){ return blockDim.x*blockDim.y; }
#define LFOR(i, cnt) for (int i = (int)(threadIdx.x & 31); i < (cnt); i += 32)

// ------------------------------------------------------------------------------------------------------------------
// backward pass: one CTA per problem, one warp per time block (M warps).  A block whose Huu is not positive definite
// (1-D and 4-D inverses only) reports failure; then rho goes up and ALL blocks run again (backwardPassGPU :497-511) --
// inside the kernel, per problem, without a host round trip.
// ------------------------------------------------------------------------------------------------------------------
struct PlugBpWs {
    float P[PN*PN], p[PN], AB[PN*PNM], AB2[PN*PNM], H[PNM*PNM], g[PNM], K[PM*PN], du[PM], Huu[(2*PM*PM > 32 ? 2*PM*PM : 32) + 2], dJ[2*PM], dx[PN], nP[PN*PN], np[PN];
};

// Huu^-1 into w.Huu[PM*PM ...] (column-major), 0 on success / 1 on "not positive definite"
__device__ __forceinline__ int plug_inv_huu(PlugBpWs &w, int l){
    constexpr int nm = PNM, oHUU = PN*PNM + PN;
    if (PM == 4){
        // invHuu_dim4 (bpHelpers.cuh:130-188): adjugate of the 4x4, 1/det tested for > 0
        float *adj = w.Huu, *M4 = w.Huu + 16;
        LFOR(e, 16){ const int kx = e % 4, ky = e / 4; M4[ky*4 + kx] = w.H[oHUU + kx + nm*ky]; }
        __syncwarp();
        LFOR(e, 16){
            const int kx = e % 4, ky = e / 4;
            const int r0 = (kx+1)%4, c0 = (ky+1)%4, r1 = (r0+1)%4, c1 = (c0+1)%4, r2 = (r1+1)%4, c2 = (c1+1)%4;
            const float C0 = M4[c0*4+r0], C1 = M4[c0*4+r1], C2 = M4[c0*4+r2], C3 = M4[c1*4+r0], C4 = M4[c1*4+r1], C5 = M4[c1*4+r2], C6 = M4[c2*4+r0], C7 = M4[c2*4+r1], C8 = M4[c2*4+r2];
            // C0 C4 C8 + C3 C7 C2 + C6 C1 C5 - C2 C4 C6 - C5 C7 C0 - C8 C1 C3, contracted as nvcc contracts the reference's expression
            float cdet = FMA(MUL(C0, C4), C8, MUL(MUL(C3, C7), C2));
            cdet = FMA(MUL(C6, C1), C5, cdet);
            cdet = FMA(-MUL(C2, C4), C6, cdet);
            cdet = FMA(-MUL(C5, C7), C0, cdet);
            cdet = FMA(-MUL(C8, C1), C3, cdet);
            adj[ky*4 + kx] = ((kx + ky) % 2) ? -cdet : cdet;
        }
        __syncwarp();
        const float det = FMA(adj[3], M4[3], FMA(adj[2], M4[2], FMA(adj[0], M4[0], MUL(adj[1], M4[1]))));
        const float val = DIV(1.0f, det);
        if (val <= 0.0f){ return 1; }
        float inv = 0.f; const int kx = l % 4, ky = (l / 4) % 4;
        if (l < 16){ inv = MUL(val, adj[ky*4 + kx]); }
        __syncwarp();
        if (l < 16){ M4[kx*4 + ky] = inv; }
        __syncwarp();
        return 0;
    } else {
        // invHuu (bpHelpers.cuh:190-204): un-pivoted Gauss-Jordan on [Huu | I], never reports failure
        LFOR(e, PM*PM){ const int kx = e % PM, ky = e / PM; w.Huu[kx + PM*ky] = w.H[oHUU + kx + nm*ky]; w.Huu[PM*PM + ky*PM + kx] = (kx == ky) ? 1.0f : 0.0f; }
        __syncwarp();
        gauss_jordan_group<PM, 32>(w.Huu);
        return 0;
    }
}

__global__ void plug_bp_kernel(DevState S, int b0){
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_err[8]; __shared__ float s_rho, s_drho;
    constexpr int n = PN, m = PM, nm = PNM, oHXU = n*nm, oHUU = n*nm + n, oGU = n, oB = n*n;
    const int b = b0 + blockIdx.x, block = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (S.done[b]){ return; }
    PlugBpWs &w = reinterpret_cast<PlugBpWs*>(smem_raw)[block];
    const int cur = S.iter[b] & 1, N = S.N, NBB = N / S.M, a = S.alphaIndex[b];
    const size_t bN = (size_t)b*N;
    // the trajectory the pass linearises around is the accepted one (xp, dp): the reference reads x[alphaIndex], d[alphaIndex], which
    // are copies of it at this point (memcpyCurrAKern)
    (void)a;
    const float *gx = S.xp + bN*n, *gxp2 = S.xp2 + bN*n, *gd = S.dp + bN*n;
    const float *gAB = S.AB + bN*S.ab_stride, *gH = S.H + bN*S.h_stride, *gg = S.g + bN*S.g_stride;
    float *gP = S.Pbuf[cur] + bN*n*n, *gp = S.pbuf[cur] + bN*n; const float *gPp = S.Pbuf[cur^1] + bN*n*n, *gpp = S.pbuf[cur^1] + bN*n;
    if (threadIdx.x == 0){ s_rho = S.rho[b]; s_drho = S.drho[b]; }
    __syncthreads();
    for (int attempt = 0; ; attempt++){
        const float rho = s_rho;
        int fail = 0;
        int ks = NBB*(block+1) - 1, iterCount;
        LFOR(e, 2*m){ w.dJ[e] = 0.f; }
        if (ks == N - 1){
            // final block (:362-367): Hxx[N-1] -> P[N-2], gx[N-1] -> p[N-2]
            LFOR(e, n*n){ const int kx = e % n, ky = e / n; const float v = gH[(size_t)ks*S.h_stride + kx + nm*ky]; gP[(size_t)(ks-1)*n*n + e] = v; w.P[e] = v; }
            LFOR(e, n){ const float v = gg[(size_t)ks*S.g_stride + e]; gp[(ks-1)*n + e] = v; w.p[e] = v; }
            ks--; iterCount = NBB - 2;
        } else {
            // other blocks (:369,376): last iteration's P, p at the block's right edge, p shifted to the new linearisation point
            iterCount = NBB - 1;
            LFOR(e, n*n){ w.P[e] = gPp[(size_t)ks*n*n + e]; }
            LFOR(e, n){ w.dx[e] = SUB(gx[(ks+1)*n + e], gxp2[(ks+1)*n + e]); }
            __syncwarp();
            LFOR(r, n){ float val = 0.f; for (int j = 0; j < n; j++){ val = FMA(w.P[r + n*j], w.dx[j], val); } w.p[r] = FMA(1.0f, val, gpp[ks*n + r]); }
        }
        __syncwarp();
        for (int iter = iterCount; iter >= 0; iter--, ks--){
            const float *bH = gH + (size_t)ks*S.h_stride, *bg = gg + (size_t)ks*S.g_stride, *bd = gd + ks*n;
            LFOR(e, n*nm){ w.AB[e] = gAB[(size_t)ks*S.ab_stride + e]; }
            __syncwarp();
            // backprop (:37-93): AB2 = AB'(P + rho I[u rows])
            LFOR(e, n*nm){
                const int kx = e % nm, ky = e / nm; float val = 0.f;
                for (int j = 0; j < n; j++){ val = FMA(w.AB[kx*n + j], ADD(w.P[ky*n + j], (kx >= n && ky == j) ? rho : 0.0f), val); }
                w.AB2[ky*nm + kx] = val;
            }
            // p += P d on the block-local defect boundary (:67-81)
            float pnew = 0.f;
            if (l < n){
                float val = 0.f;
                if (S.M > 1 && (((iter+1) % NBB) == 0) && iter < N-1){ for (int j = 0; j < n; j++){ val = FMA(bd[j], ADD(w.P[l + j*n], 0.0f), val); } }
                pnew = ADD(w.p[l], val);
            }
            __syncwarp();
            if (l < n){ w.p[l] = pnew; }
            __syncwarp();
            // H = AB2*AB + H_cost (the product lands transposed, cudaUtils.h:547-585); g = AB'p + g_cost (:86-87)
            LFOR(e, nm*nm){
                const int kx = e % nm, ky = e / nm; float val = 0.f;
                for (int j = 0; j < n; j++){ val = FMA(w.AB2[ky + nm*j], w.AB[kx*n + j], val); }
                w.H[kx + nm*ky] = FMA(1.0f, val, MUL(1.0f, bH[kx + nm*ky]));
            }
            LFOR(kx, nm){ float val = 0.f; for (int j = 0; j < n; j++){ val = FMA(w.p[j], w.AB[kx*n + j], val); } w.g[kx] = FMA(1.0f, val, MUL(1.0f, bg[kx])); }
            __syncwarp();
            float *bKT = S.KT + (bN + ks)*n*m, *bdu = S.du + (bN + ks)*m;
            if (m == 1){
                // computeKTdu_dim1 (:96-128): Huu must be positive
                if (w.H[oHUU] <= 0.0f){ fail = 1; break; }
                const float val = DIV(1.0f, ADD(w.H[oHUU], 0.0f));
                LFOR(ky, n){ const float k_ = MUL(w.H[oGU + ky*nm], val); w.K[ky*m] = k_; bKT[ky] = k_; }
                if (l == 0){ const float d_ = MUL(w.g[oGU], val); w.du[0] = d_; bdu[0] = d_; }
            } else {
                if (plug_inv_huu(w, l)){ fail = 1; break; }
                const float *Hinv = w.Huu + m*m;
                // computeKTdu (:206-220)
                LFOR(e, n*m){
                    const int kx = e % m, ky = e / m; float val = 0.f;
                    for (int j = 0; j < m; j++){ val = FMA(Hinv[kx + m*j], w.H[oGU + ky*nm + j], val); }
                    w.K[kx + ky*m] = MUL(1.0f, val);
                }
                LFOR(r, m){ float val = 0.f; for (int j = 0; j < m; j++){ val = FMA(Hinv[r + m*j], w.g[oGU + j], val); } w.du[r] = ADD(MUL(1.0f, val), 0.0f); }
                __syncwarp();
                LFOR(e, n*m){ const int kx = e % n, ky = e / n; bKT[kx + n*ky] = w.K[ky + m*kx]; }
                LFOR(r, m){ bdu[r] = w.du[r]; }
            }
            __syncwarp();
            // computeCTG (:223-276), skipped for the very first knot (:396)
            if (iter != 0 || block != 0){
                LFOR(e, n*m){
                    const int kx = e % n, ky = e / n; float val = 0.f;
                    for (int j = 0; j < m; j++){ val = FMA(w.K[kx*m + j], w.H[oHUU + ky*nm + j], val); }
                    w.AB2[kx + ky*n] = SUB(val, w.H[oHXU + kx + nm*ky]);
                }
                __syncwarp();
                LFOR(e, n*n){
                    const int kx = e % n, ky = e / n; float val = 0.f;
                    for (int j = 0; j < m; j++){ val = ADD(val, FMA(w.AB2[kx + n*j], w.K[ky*m + j], -MUL(w.K[kx*m + j], w.H[oGU + ky*nm + j]))); }
                    w.nP[e] = ADD(w.H[kx + ky*nm], val);
                }
                LFOR(kx, n){
                    float val = 0.f;
                    for (int j = 0; j < m; j++){ val = ADD(val, FMA(w.du[j], w.AB2[kx + n*j], -MUL(w.K[kx*m + j], w.g[oGU + j]))); }
                    w.np[kx] = ADD(w.g[kx], val);
                }
                __syncwarp();
                LFOR(e, n*n){ const float v = w.nP[e]; w.P[e] = v; gP[(size_t)(ks-1)*n*n + e] = v; }
                LFOR(e, n){ const float v = w.np[e]; w.p[e] = v; gp[(ks-1)*n + e] = v; }
            }
            // computeFSVars (:279-312)
            if (S.M > 1){
                float *bA = S.ApBK + (bN + ks)*n*n, *bB = S.Bdu + (bN + ks)*n;
                LFOR(e, n*n){
                    const int kx = e % n, ky = e / n; float val = 0.f;
                    for (int j = 0; j < m; j++){ val = FMA(w.AB[oB + kx + n*j], w.K[ky*m + j], val); }
                    bA[e] = SUB(w.AB[kx + n*ky], val);
                }
                LFOR(kx, n){ float val = 0.f; for (int j = 0; j < m; j++){ val = FMA(w.AB[oB + kx + n*j], w.du[j], val); } bB[kx] = val; }
            }
            // computeExpRed (:315-334)
            LFOR(ind, m){
                float dot = 0.f; for (int j = 0; j < m; j++){ dot = FMA(w.H[oHUU + ind + nm*j], w.du[j], dot); }
                w.dJ[ind] = FMA(w.du[ind], w.g[oGU + ind], w.dJ[ind]); w.dJ[m + ind] = FMA(w.du[ind], dot, w.dJ[m + ind]);
            }
            __syncwarp();
        }
        __syncwarp();
        if (l == 0){
            s_err[block] = fail;
            if (!fail){
                float e0 = w.dJ[0], e1 = w.dJ[m];
                for (int j = 1; j < m; j++){ e0 = ADD(e0, w.dJ[j]); e1 = ADD(e1, w.dJ[m + j]); }
                S.dJexp[(size_t)b*2*S.M + 2*block] = e0; S.dJexp[(size_t)b*2*S.M + 2*block + 1] = e1;
            }
        }
        __syncthreads();
        int any = 0; for (int q = 0; q < S.M; q++){ any |= s_err[q]; }
        if (!any || attempt >= PDDP_MAX_RHO_RETRIES){ break; }
        __syncthreads();
        if (threadIdx.x == 0){
            // host arithmetic of the reference (:498-499): never fused
            const float drho = fmaxf(MUL(s_drho, S.rho_factor), S.rho_factor); s_drho = drho; s_rho = fminf(MUL(s_rho, drho), S.rho_max);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0){ S.rho[b] = s_rho; S.drho[b] = s_drho; }
}

// ------------------------------------------------------------------------------------------------------------------
// forward simulation: one CTA (one warp, the plug-in's cooperating group) per (problem, candidate, shooting interval)
// ------------------------------------------------------------------------------------------------------------------
template <int INTEG>
__global__ void plug_sim_kernel(DevState S, int b0, int n_cand){
    constexpr int n = PN, m = PM;
    __shared__ float s_x[PN], s_u[PM], s_qdd[PNP], s_xn[PN], s_dx[PN], s_xg[PN];
    const int w = blockIdx.x % S.M, a = (blockIdx.x / S.M) % n_cand, b = b0 + blockIdx.x / (S.M*n_cand), l = plug_tid();
    if (S.done[b]){ return; }
    if (l == 0){ ::pddp_plugin::rt_num_time_steps() = S.N; }
    if (l < n){ s_xg[l] = S.xGoal[b*n + l]; }
    const int N = S.N, NBF = N / S.M, kStart = w*NBF, iters = (w < S.M - 1) ? NBF : NBF - 1;
    const float alpha = S.alpha[a], dt = S.dt;
    float *gx = S.x + ((size_t)b*S.A + a)*N*n, *gu = S.u + ((size_t)b*S.A + a)*N*m, *gdd = S.d + ((size_t)b*S.A + a)*N*n;
    float *gc = S.costk + ((size_t)b*S.A + a)*N;
    const float *gxp = S.xp + (size_t)b*N*n, *gup = S.up + (size_t)b*N*m, *gKT = S.KT + (size_t)b*N*n*m, *gdu = S.du + (size_t)b*N*m;
    // state at the start of the interval: left there by the sweep; the first interval starts at xp[0]
    if (l < n){
        if (w == 0){ const float v = gxp[l]; s_x[l] = v; gx[l] = v; }
        else { s_x[l] = gx[kStart*n + l]; }
    }
    __syncthreads();
    for (int kk = 0; kk < iters; kk++){
        const int k = kStart + kk;
        if (l < n){ s_dx[l] = SUB(s_x[l], gxp[k*n + l]); }
        __syncthreads();
        // u = up - (alpha du + K dx)          (computeControlKT fpHelpers.cuh:210-219)
        if (l < m){
            const float *KTk = gKT + (size_t)k*n*m; float Kdx = 0.f;
            for (int c = 0; c < n; c++){ Kdx = FMA(KTk[c + l*n], s_dx[c], Kdx); }
            const float uu = SUB(gup[k*m + l], FMA(alpha, gdu[k*m + l], Kdx));
            s_u[l] = uu; gu[k*m + l] = uu;
        }
        __syncthreads();
        _integrator<float, INTEG>(s_xn, s_x, s_u, s_qdd, const_cast<float*>(S.I), const_cast<float*>(S.Tbody), dt);
        __syncthreads();
        if (kk < NBF - 1){
            if (l < n){ const float v = s_xn[l]; s_x[l] = v; gx[(k+1)*n + l] = v; }
        } else if (w < S.M - 1){
            // last step of a non-final interval: defect against the next interval's start state (fpHelpers.cuh:255-258)
            if (l < n){ gdd[((w+1)*NBF-1)*n + l] = SUB(s_xn[l], gx[(k+1)*n + l]); }
        }
        __syncthreads();
    }
    // u[N-1] is never simulated and stays the accepted one
    if (w == S.M - 1 && l < m){ gu[(N-1)*m + l] = gup[(N-1)*m + l]; }
    __syncthreads();
    // per-knot costs of this interval (costKern fpHelpers.cuh:132-152: one thread per knot calls the plant's costFunc)
    const int kEnd = (w == S.M - 1) ? N : kStart + NBF;
    for (int k = kStart + l; k < kEnd; k += plug_nthreads()){ gc[k] = costFunc<float>(gx + k*n, gu + k*m, s_xg, k, S.Q1, S.Q2, S.R, S.QF1, S.QF2); }
}

// per-knot costs of the initial trajectory (initAlgGPU's costKern, nisInitHelpers.cuh:385) into candidate slot 0
__global__ void plug_init_cost_kernel(DevState S){
    const int b = blockIdx.x, n = PN, m = PM;
    if (plug_tid() == 0){ ::pddp_plugin::rt_num_time_steps() = S.N; }
    __syncthreads();
    for (int k = plug_tid(); k < S.N; k += plug_nthreads()){
        S.costk[((size_t)b*S.A + 0)*S.N + k] = costFunc<float>(S.xp + ((size_t)b*S.N + k)*n, S.up + ((size_t)b*S.N + k)*m, S.xGoal + b*n, k, S.Q1, S.Q2, S.R, S.QF1, S.QF2);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// next-iteration setup: one CTA per (problem, knot).  mode as nis_kernel (kernels.cuh): 0 iteration, 1 initialisation,
// 2 initialisation after loadVarsGPU's forward rollout.
// ------------------------------------------------------------------------------------------------------------------
template <int INTEG>
__global__ void plug_nis_kernel(DevState S, int mode, int b0, int nb){
    constexpr int n = PN, m = PM, nm = PNM;
    __shared__ float s_x[PN], s_u[PM], s_qdd[PNP], s_dqdd[PNP*PNM];
    const int N = S.N, b = b0 + blockIdx.x / N, k = blockIdx.x % N, l = plug_tid();
    if (b >= b0 + nb || S.done[b]){ return; }
    if (l == 0){ ::pddp_plugin::rt_num_time_steps() = N; }
    float *gxp = S.xp + ((size_t)b*N + k)*n, *gup = S.up + ((size_t)b*N + k)*m, *gdp = S.dp + ((size_t)b*N + k)*n, *gxp2 = S.xp2 + ((size_t)b*N + k)*n;
    const bool acc = (mode == 2) || ((mode == 0) && S.accepted[b]);
    const int a = S.alphaIndex[b];
    const float *cx = S.x + (((size_t)b*S.A + a)*N + k)*n, *cu = S.u + (((size_t)b*S.A + a)*N + k)*m, *cd = S.d + (((size_t)b*S.A + a)*N + k)*n;
    if (l < n){
        const float xold = gxp[l];
        const float xv = acc ? cx[l] : xold; s_x[l] = xv;
        gxp2[l] = (mode == 2) ? xv : xold;                         // xp2 <- xp (fpHelpers.cuh:371; nisInitHelpers.cuh:378-379 at initialisation)
        if (acc){ gxp[l] = xv; if (((k+1) % (N / S.M)) == 0 && k < N-1){ gdp[l] = cd[l]; } }
    }
    if (l < m){ const float uv = acc ? cu[l] : gup[l]; s_u[l] = uv; if (acc){ gup[l] = uv; } }
    if (S.skip_unchanged && mode == 0 && !acc){ return; }
    __syncthreads();
    // costGradientHessianKern (nisInitHelpers.cuh:44-93): one thread per knot calls the plant's costGrad
    if (l == 0){
        costGrad<float>(S.H + ((size_t)b*N + k)*S.h_stride, S.g + ((size_t)b*N + k)*S.g_stride, s_x, s_u, S.xGoal + b*n, k, nm, S.Q1, S.Q2, S.R, S.QF1, S.QF2);
    }
    if (k == N - 1){ return; }
    // integratorGradientKern (nisInitHelpers.cuh:203-221)
    _integratorGradient<float, INTEG>(S.AB + ((size_t)b*N + k)*S.ab_stride, s_x, s_u, s_qdd, s_dqdd, const_cast<float*>(S.I), const_cast<float*>(S.Tbody), S.dt, n);
}

// ------------------------------------------------------------------------------------------------------------------
// receding horizon, load step (loadVarsGPU_MPC MPCHelpers.cuh:602-657): shifts, then the open-loop rollout from the measured
// state over the whole horizon (rolloutMPC :524-556, FULL_ROLLOUT 1), then the hand-over.  One CTA (one warp) per problem.
// ------------------------------------------------------------------------------------------------------------------
template <int INTEG>
__global__ void plug_mpc_load_kernel(DevState S, MpcState Q){
    constexpr int n = PN, m = PM;
    __shared__ float s_x[PN], s_u[PM], s_qdd[PNP], s_xn[PN];
    const int b = blockIdx.x, N = S.N, shift = Q.shift[b], l = plug_tid(); const bool clear = Q.clear[b] != 0;
    if (l == 0){ S.init_knot[b] = 0; ::pddp_plugin::rt_num_time_steps() = N; }
    float *cx = Q.cx + (size_t)b*N*n, *cu = Q.cu + (size_t)b*N*m, *cd = Q.cd + (size_t)b*N*n;
    float *xp = S.xp + (size_t)b*N*n, *up = S.up + (size_t)b*N*m, *dp = S.dp + (size_t)b*N*n, *KT = S.KT + (size_t)b*N*n*m;
    float *P0 = S.Pbuf[0] + (size_t)b*N*n*n, *P1 = S.Pbuf[1] + (size_t)b*N*n*n, *p0 = S.pbuf[0] + (size_t)b*N*n, *p1 = S.pbuf[1] + (size_t)b*N*n;
    if (!clear && (S.iter[b] & 1) == 0){ mpc_swap(P0, P1, N*n*n); mpc_swap(p0, p1, N*n); __syncthreads(); }      // see mpc_load_kernel (kernels.cuh)
    if (shift > 0){
        // register-slab shifts (dev_state.cuh): eight loads in flight per thread, no scratch pass
        const int T = plug_nthreads();
        mpc_shift_part(cx, shift, n, N, false, xp, l, T, 0);
        mpc_shift_part(cd, shift, n, N, false, nullptr, l, T, 0);
        if (!clear){
            mpc_shift_part(cu, shift, m, N-1, true, up, l, T, 0);
            mpc_shift_part(KT, shift, n*m, N-1, true, nullptr, l, T, 0);
            mpc_shift_part(P0, shift, n*n, N, false, nullptr, l, T, 0); mpc_shift_part(P1, shift, n*n, N, false, nullptr, l, T, 0);
            mpc_shift_part(p0, shift, n, N, false, nullptr, l, T, 0); mpc_shift_part(p1, shift, n, N, false, nullptr, l, T, 0);
        }
        __syncthreads();
    }
    if (clear){ mpc_zero(cu, N*m); mpc_zero(KT, N*n*m); mpc_zero(P0, N*n*n); mpc_zero(P1, N*n*n); mpc_zero(p0, N*n); mpc_zero(p1, N*n); }
    mpc_zero(S.du + (size_t)b*N*m, N*m); mpc_zero(S.dT + (size_t)b*S.A, S.A);
    __syncthreads();
    if (l < n){ const float v = Q.xActual[b*n + l]; s_x[l] = v; cx[l] = v; }
    for (int k = 0; k < N-1; k++){
        if (l < m){ s_u[l] = cu[k*m + l]; }
        __syncthreads();
        _integrator<float, INTEG>(s_xn, s_x, s_u, s_qdd, const_cast<float*>(S.I), const_cast<float*>(S.Tbody), S.dt);
        __syncthreads();
        if (l < n){ const float v = s_xn[l]; s_x[l] = v; cx[(k+1)*n + l] = v; }
        __syncthreads();
    }
    mpc_copy(Q.x_old + (size_t)b*N*n, xp, N*n); mpc_copy(Q.u_old + (size_t)b*N*m, up, N*m); mpc_copy(Q.KT_old + (size_t)b*N*n*m, KT, N*n*m);
    __syncthreads();
    mpc_copy(xp, cx, N*n); mpc_copy(up, cu, N*m); mpc_copy(dp, cd, N*n);
}

// ------------------------------------------------------------------------------------------------------------------
// plug-in unit kernels: the plant's dynamics and the integrator gradient on independent samples
// ------------------------------------------------------------------------------------------------------------------
__global__ void plug_unit_dynamics_kernel(const float *I, const float *Tbody, const float *x, const float *u, int nsamp, float *qdd){
    __shared__ float s_x[PN], s_u[PM], s_qdd[PNP];
    const int l = plug_tid();
    for (int k = blockIdx.x; k < nsamp; k += gridDim.x){
        if (l < PN){ s_x[l] = x[(size_t)k*PN + l]; } if (l < PM){ s_u[l] = u[(size_t)k*PM + l]; }
        __syncthreads();
        dynamics<float>(s_qdd, s_x, s_u, const_cast<float*>(I), const_cast<float*>(Tbody));
        __syncthreads();
        if (l < PNP){ qdd[(size_t)k*PNP + l] = s_qdd[l]; }
        __syncthreads();
    }
}
template <int INTEG>
__global__ void plug_unit_gradient_kernel(const float *I, const float *Tbody, const float *x, const float *u, int nsamp, float dt, float *AB, float *qdd, float *xnext){
    __shared__ float s_x[PN], s_u[PM], s_qdd[PNP], s_dqdd[PNP*PNM], s_xn[PN];
    const int l = plug_tid();
    for (int k = blockIdx.x; k < nsamp; k += gridDim.x){
        if (l < PN){ s_x[l] = x[(size_t)k*PN + l]; } if (l < PM){ s_u[l] = u[(size_t)k*PM + l]; }
        __syncthreads();
        _integratorGradient<float, INTEG>(AB + (size_t)k*PN*PNM, s_x, s_u, s_qdd, s_dqdd, const_cast<float*>(I), const_cast<float*>(Tbody), dt, PN);
        __syncthreads();
        if (qdd && l < PNP){ qdd[(size_t)k*PNP + l] = s_qdd[l]; }
        __syncthreads();
        if (xnext){
            _integrator<float, INTEG>(s_xn, s_x, s_u, s_qdd, const_cast<float*>(I), const_cast<float*>(Tbody), dt);
            __syncthreads();
            if (l < PN){ xnext[(size_t)k*PN + l] = s_xn[l]; }
            __syncthreads();
        }
    }
}
__global__ void plug_unit_cost_kernel(DevState S, const float *x, const float *u, const float *xg, const int *knot, int nsamp, float *J, float *H, float *g){
    if (plug_tid() == 0){ ::pddp_plugin::rt_num_time_steps() = S.N; }
    __syncthreads();
    for (int k = blockIdx.x*plug_nthreads() + plug_tid(); k < nsamp; k += gridDim.x*plug_nthreads()){
        float xs[PN], us[PM];
        for (int i = 0; i < PN; i++){ xs[i] = x[(size_t)k*PN + i]; } for (int i = 0; i < PM; i++){ us[i] = u[(size_t)k*PM + i]; }
        J[k] = costFunc<float>(xs, us, const_cast<float*>(xg), knot[k], S.Q1, S.Q2, S.R, S.QF1, S.QF2);
        costGrad<float>(H + (size_t)k*PNM*PNM, g + (size_t)k*PNM, xs, us, const_cast<float*>(xg), knot[k], PNM, S.Q1, S.Q2, S.R, S.QF1, S.QF2);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side: the operations table a plant translation unit exports (include/pddp_plant.h)
// ------------------------------------------------------------------------------------------------------------------
#define PLUG_BY_INTEG(KERNEL, GRID, BLOCK, SMEM, ST, ...) do { \
        if (S.integrator == 1){ KERNEL<1><<<GRID, BLOCK, SMEM, ST>>>(__VA_ARGS__); } \
        else if (S.integrator == 2){ KERNEL<2><<<GRID, BLOCK, SMEM, ST>>>(__VA_ARGS__); } \
        else { KERNEL<3><<<GRID, BLOCK, SMEM, ST>>>(__VA_ARGS__); } } while (0)

static int plug_launch_bp(const void *state, void *stream, int b0, int nb){
    const DevState &S = *static_cast<const DevState*>(state);
    plug_bp_kernel<<<nb, 32*S.M, S.M*sizeof(PlugBpWs), static_cast<cudaStream_t>(stream)>>>(S, b0);
    return (int)cudaGetLastError();
}
static int plug_launch_sweep(const void *state, void *stream, int b0, int nb, int num_sms){
    const DevState &S = *static_cast<const DevState*>(state);
    int splits = 1; while (nb*splits*2 <= num_sms && (S.A % (splits*2)) == 0){ splits *= 2; }
    sweep_kernel<PN><<<nb*splits, 32*(S.A/splits), SWEEP_SLOTS*sizeof(SweepSlot<PN>), static_cast<cudaStream_t>(stream)>>>(S, splits, b0);
    return (int)cudaGetLastError();
}
static int plug_launch_sim(const void *state, void *stream, int b0, int nb, int n_cand){
    const DevState &S = *static_cast<const DevState*>(state);
    PLUG_BY_INTEG(plug_sim_kernel, nb*n_cand*S.M, PLUG_BLOCK, 0, static_cast<cudaStream_t>(stream), S, b0, n_cand);
    return (int)cudaGetLastError();
}
static int plug_launch_init_cost(const void *state, void *stream){
    const DevState &S = *static_cast<const DevState*>(state);
    plug_init_cost_kernel<<<S.B, PLUG_BLOCK, 0, static_cast<cudaStream_t>(stream)>>>(S);
    return (int)cudaGetLastError();
}
static int plug_launch_nis(const void *state, void *stream, int mode, int b0, int nb){
    const DevState &S = *static_cast<const DevState*>(state);
    PLUG_BY_INTEG(plug_nis_kernel, nb*S.N, PLUG_BLOCK, 0, static_cast<cudaStream_t>(stream), S, mode, b0, nb);
    return (int)cudaGetLastError();
}
static int plug_launch_mpc_load(const void *state, const void *mpc, void *stream){
    const DevState &S = *static_cast<const DevState*>(state); const MpcState &Q = *static_cast<const MpcState*>(mpc);
    PLUG_BY_INTEG(plug_mpc_load_kernel, S.B, PLUG_BLOCK, 0, static_cast<cudaStream_t>(stream), S, Q);
    return (int)cudaGetLastError();
}
static int plug_unit_dynamics(const void *state, void *stream, const float *d_x, const float *d_u, int nsamp, float *d_qdd){
    const DevState &S = *static_cast<const DevState*>(state);
    plug_unit_dynamics_kernel<<<nsamp < 2048 ? nsamp : 2048, PLUG_BLOCK, 0, static_cast<cudaStream_t>(stream)>>>(S.I, S.Tbody, d_x, d_u, nsamp, d_qdd);
    return (int)cudaGetLastError();
}
static int plug_unit_gradient(const void *state, void *stream, const float *d_x, const float *d_u, int nsamp, float *d_AB, float *d_qdd, float *d_xnext){
    const DevState &S = *static_cast<const DevState*>(state);
    PLUG_BY_INTEG(plug_unit_gradient_kernel, nsamp < 2048 ? nsamp : 2048, PLUG_BLOCK, 0, static_cast<cudaStream_t>(stream), S.I, S.Tbody, d_x, d_u, nsamp, S.dt, d_AB, d_qdd, d_xnext);
    return (int)cudaGetLastError();
}
static int plug_unit_cost(const void *state, void *stream, const float *d_x, const float *d_u, const float *d_xg, const int *d_knot, int nsamp, float *d_J, float *d_H, float *d_g){
    const DevState &S = *static_cast<const DevState*>(state);
    plug_unit_cost_kernel<<<(nsamp + 31)/32, PLUG_BLOCK, 0, static_cast<cudaStream_t>(stream)>>>(S, d_x, d_u, d_xg, d_knot, nsamp, d_J, d_H, d_g);
    return (int)cudaGetLastError();
}
static void plug_init_model(float *I, float *Tbody){ initI<float>(I); initT<float>(Tbody); }
static int plug_prepare(int max_M){
    if (cudaFuncSetAttribute(plug_bp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(max_M*sizeof(PlugBpWs))) != cudaSuccess){ return (int)cudaGetLastError(); }
    if (cudaFuncSetAttribute(sweep_kernel<PN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SWEEP_SLOTS*sizeof(SweepSlot<PN>))) != cudaSuccess){ return (int)cudaGetLastError(); }
    return 0;
}

