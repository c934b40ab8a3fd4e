// kernels.cuh -- the batched iLQR hot path as sm_100a kernels.
//
// One iteration (reference call order DDPWrappers.cuh:52-106) is five launches on one stream, no host sync:
//   bp_kernel      backward Riccati-like recursion, one CTA per (problem, time block)       [bpHelpers.cuh:337-420]
//   sweep_kernel   linearised multiple-shooting sweep, one CTA per problem, one warp per alpha [fpHelpers.cuh:17-63]
//   sim_kernel     nonlinear forward sim + control update + per-knot cost + defects,
//                  one warp per (problem, alpha, shooting interval)                         [fpHelpers.cuh:200-301,132-152]
//   select_kernel  cost tree-sum, defect max, line search, accept/reject, rho schedule      [fpHelpers.cuh:94-111,374-408; nisInitHelpers.cuh:487-518]
//   nis_kernel     trajectory hand-over + integrator gradient AB + cost gradient g (H once), one warp per (problem, knot)
//                                                                                          [nisInitHelpers.cuh:203-221,44-93,245-279]
// Data layout in HBM: the reference's per-problem layouts (column-major tiles, [k][col][row]) with a leading batch
// dimension; candidates x,u,d carry [batch][alpha][N][.].  The reference's per-iteration copies disappear:
//   Pp<-P / pp<-p     -> ping-pong of two buffers (cur flips every iteration)
//   alpha broadcast   -> sweep/sim read the accepted trajectory (xp,up,dp) as the base of every candidate
//   reject restore    -> nothing to restore (xp,up,dp are only overwritten on accept)
#pragma once
#include "pddp_math.cuh"
#include "plant_kuka.cuh"

namespace pddp {

struct DevState {
    // sizes
    int B, N, A, M, n, m, max_iter;
    float dt, tol_cost, two_tol;
    float rho_min, rho_max, rho_factor, inv_rho_factor, exp_red_min, exp_red_max, max_defect;
    float Q1, Q2, R, QF1, QF2;
    // model constants (device)
    const float *I, *Tbody, *alpha;
    // trajectories
    float *x, *u, *d;              // candidates [B][A][N][n|m|n]
    float *xp, *xp2, *up, *dp;     // accepted / previous accepted [B][N][.]
    float *AB, *H, *g;
    float *Pbuf[2], *pbuf[2];      // ping-pong: P = Pbuf[cur], Pp = Pbuf[cur^1]
    float *KT, *du, *ApBK, *Bdu;
    float *xGoal;                  // [B][n]
    float *costk;                  // per-knot costs [B][A][N]
    float *J, *dT, *dJexp;         // [B][A], [B][A], [B][2M]
    // per-problem solver scalars
    float *rho, *drho, *prevJ, *dJ, *z;
    int *iter, *alphaIndex, *ignore_defect, *done, *accepted, *final_src;
    float *Jout; int *alphaOut;    // [B][max_iter+1]
    int *n_active;                 // [1]
};

// ------------------------------------------------------------------------------------------------------------------
// async staging helpers (LDGSTS): tiles are not 16-byte aligned in the reference layout (14*21 floats), so 4-byte copies
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gsrc){
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit(){ asm volatile("cp.async.commit_group;\n" ::); }
template <int N_> __device__ __forceinline__ void cp_async_wait(){ asm volatile("cp.async.wait_group %0;\n" :: "n"(N_)); }

// ------------------------------------------------------------------------------------------------------------------
// backward pass
// ------------------------------------------------------------------------------------------------------------------
// One CTA of BP_THREADS (16 warps) per (problem, time block).  Every knot is five barrier-separated stages; inside a
// stage each thread owns at most one output element (a 14- or 7-term fused chain), its indices fixed before the knot
// loop.  The 7x7 Huu inverse is the serial bottleneck of a knot, so warp 0 computes Huu first and eliminates it in
// registers (shuffles) while warps 1..15 assemble the other 392 entries of H and g.  The inputs of knot k-1 (AB, H, g,
// d: 3.1 KB) are in flight (LDGSTS) while knot k is processed.
constexpr int BP_THREADS = 512;
template <int n, int m>
struct BpSmem {
    static constexpr int nm = n + m;
    float AB[2][n*nm];      // double-buffered inputs of the knot being processed / prefetched
    float Hc[2][nm*nm];
    float gc[2][nm];
    float dk[2][n];
    float P[n*n], p[n];
    float AB2[n*nm];        // AB'(P+rho) then K'Huu - Hxu
    float H[nm*nm], g[nm];
    float K[m*n], du[m];
    float Huu[2*m*m];
    float dx[n];
    float dJ[2*m];
};

template <int n, int m>
__device__ __forceinline__ void bp_prefetch(BpSmem<n,m> &s, int buf, const float *gAB, const float *gH, const float *gg, const float *gd){
    constexpr int nm = n + m;
    const int t = threadIdx.x;
    if (t < n*nm){ cp_async4(&s.AB[buf][t], gAB + t); }
    if (t < nm*nm){ cp_async4(&s.Hc[buf][t], gH + t); }
    if (t >= nm*nm && t < nm*nm + nm){ cp_async4(&s.gc[buf][t - nm*nm], gg + t - nm*nm); }
    if (t >= nm*nm + nm && t < nm*nm + nm + n){ cp_async4(&s.dk[buf][t - nm*nm - nm], gd + t - nm*nm - nm); }
    cp_async_commit();
}

template <int n, int m>
__global__ void __launch_bounds__(BP_THREADS) bp_kernel(DevState S, int cur, int b0){
    constexpr int nm = n + m, oHXU = n*nm, oHUU = n*nm + n, oGU = n, oB = n*n;
    static_assert(nm*nm + nm + n <= BP_THREADS && n*nm + n <= BP_THREADS, "one element per thread");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BpSmem<n,m> &s = *reinterpret_cast<BpSmem<n,m>*>(smem_raw);
    const int b = b0 + blockIdx.x / S.M, block = blockIdx.x % S.M, t = threadIdx.x;
    if (S.done[b]){ return; }
    const int N = S.N, NBB = N / S.M;
    const float rho = S.rho[b];
    float *gP = S.Pbuf[cur] + (size_t)b*N*n*n, *gp = S.pbuf[cur] + (size_t)b*N*n;
    const float *gPp = S.Pbuf[cur^1] + (size_t)b*N*n*n, *gpp = S.pbuf[cur^1] + (size_t)b*N*n;
    const float *gAB = S.AB + (size_t)b*N*n*nm, *gH = S.H + (size_t)b*N*nm*nm, *gg = S.g + (size_t)b*N*nm;
    const float *gd = S.dp + (size_t)b*N*n, *gx = S.xp + (size_t)b*N*n, *gx2 = S.xp2 + (size_t)b*N*n;
    float *gKT = S.KT + (size_t)b*N*n*m, *gdu = S.du + (size_t)b*N*m, *gApBK = S.ApBK + (size_t)b*N*n*n, *gBdu = S.Bdu + (size_t)b*N*n;

    // ---- per-thread roles, fixed for the whole block of knots
    // stage A: element t of AB2 (t < n*nm); threads n*nm .. n*nm+n-1 update p
    const int a_ky = t / nm, a_kx = t % nm;
    // stage B: warp 0 -> Huu (2 per lane) then the elimination; threads 32.. -> the other H entries, then g
    int b_kx = -1, b_ky = 0;
    {
        const int item = t - 32;
        if (item >= 0 && item < n*nm){ b_ky = item / nm; b_kx = item % nm; }                               // columns 0..n-1, all rows
        else if (item >= n*nm && item < n*nm + m*n){ const int q = item - n*nm; b_ky = n + q / n; b_kx = q % n; } // columns n.., rows 0..n-1
    }
    const int b_g = (t - 32 - (nm*nm - m*m));          // 0..nm-1 -> g
    // stage C: K element t (t < n*m): kx = row of K (m), ky = column (n); threads 128.. -> du
    const int c_ky = t / m, c_kx = t % m;
    // stage D: [0, n*m) T;  [128, 128+n*n) ApBK;  [352, 352+n) Bdu;  [384, 384+m) expected reduction;  [400, 400+n*m) KT store; [500,500+m) du store
    const int d_ky = t / n, d_kx = t % n;                       // T (ky < m) and KT store
    const int d2 = t - 128, d2_ky = d2 / n, d2_kx = d2 % n;     // ApBK
    // stage E: P element t (t < n*n); threads 256.. -> p
    const int e_ky = t / n, e_kx = t % n;

    int ks = NBB*(block+1) - 1, iterCount;
    float dJ0 = 0.f, dJ1 = 0.f;          // thread 384+ind: running sums of du*gu and du*(Huu du) (bpHelpers.cuh:327-328)
    if (ks == N - 1){
        // final block: Hxx[N-1] -> P[N-2], gx[N-1] -> p[N-2]  (bpHelpers.cuh:362-367)
        if (t < n*n){ const int kx = t % n, ky = t / n; float v = MUL(1.0f, gH[(size_t)ks*nm*nm + kx + nm*ky]); s.P[t] = v; gP[(size_t)(ks-1)*n*n + t] = v; }
        if (t >= 256 && t < 256 + n){ const int r = t - 256; float v = MUL(1.0f, gg[ks*nm + r]); s.p[r] = v; gp[(ks-1)*n + r] = v; }
        ks--; iterCount = NBB - 2;
        __syncthreads();
    } else {
        // other blocks: start from the previous iteration's P,p at the block boundary, shifted to the new linearisation point
        iterCount = NBB - 1;
        if (t < n*n){ s.P[t] = gPp[(size_t)ks*n*n + t]; }
        if (t >= 256 && t < 256 + n){ const int r = t - 256; s.dx[r] = SUB(gx[(ks+1)*n + r], gx2[(ks+1)*n + r]); }
        __syncthreads();
        if (t < n){
            float val = 0.f;
            #pragma unroll
            for (int j = 0; j < n; j++){ val = FMA(s.P[t + n*j], s.dx[j], val); }
            s.p[t] = FMA(1.0f, val, gpp[ks*n + t]);
        }
        __syncthreads();
    }
    bp_prefetch<n,m>(s, 0, gAB + (size_t)ks*n*nm, gH + (size_t)ks*nm*nm, gg + ks*nm, gd + ks*n);
    int buf = 0;
    #pragma unroll 1
    for (int iter = iterCount; iter >= 0; iter--, ks--, buf ^= 1){
        if (iter > 0){ bp_prefetch<n,m>(s, buf^1, gAB + (size_t)(ks-1)*n*nm, gH + (size_t)(ks-1)*nm*nm, gg + (ks-1)*nm, gd + (ks-1)*n); cp_async_wait<1>(); }
        else { cp_async_wait<0>(); }
        __syncthreads();
        const float *sAB = s.AB[buf], *bH = s.Hc[buf], *bg = s.gc[buf], *bd = s.dk[buf];
        // ---- stage A: AB2 = AB'(P + rho I[u rows]);  p += P d on the block-local defect boundary (bpHelpers.cuh:54-81)
        if (t < n*nm){
            float val = 0.f;
            #pragma unroll
            for (int j = 0; j < n; j++){ val = FMA(sAB[a_kx*n+j], ADD(s.P[a_ky*n+j], (a_kx >= n && a_ky == j) ? rho : 0.f), val); }
            s.AB2[a_ky*nm+a_kx] = val;
        } else if (t < n*nm + n){
            const int r = t - n*nm; float val = 0.f;
            if (S.M > 1 && (((iter+1) % NBB) == 0) && iter < N-1){
                #pragma unroll
                for (int j = 0; j < n; j++){ val = FMA(bd[j], ADD(s.P[r + j*n], 0.f), val); }
            }
            s.p[r] = ADD(s.p[r], val);
        }
        __syncthreads();
        // ---- stage B: H = (AB2 AB)' + H_cost, g = AB'p + g_cost; warp 0 does Huu first and inverts it meanwhile
        if (t < 32){
            #pragma unroll
            for (int q = 0; q < 2; q++){
                const int e = t + 32*q;
                if (e < m*m){
                    const int ky = n + e / m, kx = n + e % m; float val = 0.f;
                    #pragma unroll
                    for (int j = 0; j < n; j++){ val = FMA(s.AB2[ky+nm*j], sAB[kx*n+j], val); }
                    const float h = FMA(1.0f, val, MUL(1.0f, bH[kx+nm*ky]));
                    s.H[kx+nm*ky] = h;
                    s.Huu[(kx-n) + m*(ky-n)] = MUL(1.0f, h); s.Huu[m*m + (ky-n)*m + (kx-n)] = (kx == ky) ? 1.f : 0.f;
                }
            }
            __syncwarp();
            gauss_jordan_warp_reg<m>(s.Huu);
        } else {
            if (b_kx >= 0){
                float val = 0.f;
                #pragma unroll
                for (int j = 0; j < n; j++){ val = FMA(s.AB2[b_ky+nm*j], sAB[b_kx*n+j], val); }
                s.H[b_kx+nm*b_ky] = FMA(1.0f, val, MUL(1.0f, bH[b_kx+nm*b_ky]));
            } else if (b_g >= 0 && b_g < nm){
                float val = 0.f;
                #pragma unroll
                for (int j = 0; j < n; j++){ val = FMA(s.p[j], sAB[b_g*n+j], val); }
                s.g[b_g] = FMA(1.0f, val, MUL(1.0f, bg[b_g]));
            }
        }
        __syncthreads();
        const float *Hinv = &s.Huu[m*m];
        // ---- stage C: K = Huu^-1 Hux, du = Huu^-1 gu (bpHelpers.cuh:206-220)
        if (t < n*m){
            float val = 0.f;
            #pragma unroll
            for (int j = 0; j < m; j++){ val = FMA(Hinv[c_kx+m*j], s.H[oGU + c_ky*nm + j], val); }
            s.K[c_kx+c_ky*m] = MUL(1.0f, val);
        } else if (t >= 128 && t < 128 + m){
            const int r = t - 128; float val = 0.f;
            #pragma unroll
            for (int j = 0; j < m; j++){ val = FMA(Hinv[r+m*j], s.g[oGU+j], val); }
            s.du[r] = ADD(MUL(1.0f, val), 0.f);
        }
        __syncthreads();
        // ---- stage D: T = K'Huu - Hxu (into AB2), A-BK, B du, expected reduction, KT/du to HBM
        const bool do_ctg = (iter != 0 || block != 0);
        if (t < n*m){
            if (do_ctg){
                float val = 0.f;
                #pragma unroll
                for (int j = 0; j < m; j++){ val = FMA(s.K[d_kx*m+j], s.H[oHUU+d_ky*nm+j], val); }
                s.AB2[d_kx+d_ky*n] = SUB(val, s.H[oHXU+d_kx+nm*d_ky]);
            }
        } else if (t >= 128 && t < 128 + n*n){
            if (S.M > 1){
                float val = 0.f;
                #pragma unroll
                for (int j = 0; j < m; j++){ val = FMA(sAB[oB+d2_kx+n*j], s.K[d2_ky*m+j], val); }
                gApBK[(size_t)ks*n*n + d2_kx + n*d2_ky] = SUB(sAB[d2_kx+n*d2_ky], val);
            }
        } else if (t >= 352 && t < 352 + n){
            if (S.M > 1){
                const int kx = t - 352; float val = 0.f;
                #pragma unroll
                for (int j = 0; j < m; j++){ val = FMA(sAB[oB+kx+n*j], s.du[j], val); }
                gBdu[ks*n + kx] = val;
            }
        } else if (t >= 384 && t < 384 + m){
            const int ind = t - 384; float dot = 0.f;
            #pragma unroll
            for (int j = 0; j < m; j++){ dot = FMA(s.H[oHUU+ind+nm*j], s.du[j], dot); }
            dJ0 = FMA(s.du[ind], s.g[oGU+ind], dJ0); dJ1 = FMA(s.du[ind], dot, dJ1);
        } else if (t >= 400 && t < 400 + n*m){
            const int e = t - 400, ky = e / n, kx = e % n;
            gKT[(size_t)ks*n*m + kx + n*ky] = s.K[ky + m*kx];
        } else if (t >= 500 && t < 500 + m){
            gdu[ks*m + t - 500] = s.du[t - 500];
        }
        __syncthreads();
        // ---- stage E: cost-to-go of the previous knot (bpHelpers.cuh:223-276)
        if (do_ctg){
            if (t < n*n){
                float val = 0.f;
                #pragma unroll
                for (int j = 0; j < m; j++){ val = ADD(val, FMA(s.AB2[e_kx+n*j], s.K[e_ky*m+j], -MUL(s.K[e_kx*m+j], s.H[oGU+e_ky*nm+j]))); }
                const float v = ADD(s.H[e_kx+e_ky*nm], val);
                s.P[t] = v; gP[(size_t)(ks-1)*n*n + t] = v;
            } else if (t >= 256 && t < 256 + n){
                const int r = t - 256; float val = 0.f;
                #pragma unroll
                for (int j = 0; j < m; j++){ val = ADD(val, FMA(s.du[j], s.AB2[r+n*j], -MUL(s.K[r*m+j], s.g[oGU+j]))); }
                const float v = ADD(s.g[r], val);
                s.p[r] = v; gp[(ks-1)*n + r] = v;
            }
        }
        // the next iteration's top-of-loop barrier orders stage E's writes before stage A's reads
    }
    // ---- expected cost reduction of this block: thread 0 sums the m per-thread partials in order (bpHelpers.cuh:416)
    if (t >= 384 && t < 384 + m){ s.dJ[t-384] = dJ0; s.dJ[m + t-384] = dJ1; }
    __syncthreads();
    if (t == 0){
        float a0 = s.dJ[0], a1 = s.dJ[m];
        for (int j = 1; j < m; j++){ a0 = ADD(a0, s.dJ[j]); a1 = ADD(a1, s.dJ[m+j]); }
        S.dJexp[(size_t)b*2*S.M + 2*block] = a0; S.dJexp[(size_t)b*2*S.M + 2*block + 1] = a1;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// forward sweep: x_a[k+1] = xp[k+1] + ( -alpha_a Bdu_k + (A-BK)_k (x_a[k]-xp[k]) + [boundary] d_k )
// grid = B*splits CTAs of 32*A/splits threads; the whole (A-BK) sequence of the problem is staged in shared memory once.
// ------------------------------------------------------------------------------------------------------------------
template <int n>
__global__ void sweep_kernel(DevState S, int splits, int b0){
    extern __shared__ __align__(16) float sw[];
    // `splits` CTAs share one problem (each takes A/splits step sizes) so that a small batch still covers the SMs
    const int b = b0 + blockIdx.x / splits, a0 = (blockIdx.x % splits)*(S.A / splits), N = S.N, NBF = N / S.M;
    if (S.done[b]){ return; }
    float *sA = sw;                         // [N-1][n*n]
    float *sB = sA + (size_t)(N-1)*n*n;     // [N-1][n]
    float *sxp = sB + (size_t)(N-1)*n;      // [N][n]
    float *sd = sxp + (size_t)N*n;          // [N][n] (only boundary knots are read)
    {
        const float4 *src = reinterpret_cast<const float4*>(S.ApBK + (size_t)b*N*n*n); float4 *dst = reinterpret_cast<float4*>(sA);
        for (int i = threadIdx.x; i < (N-1)*n*n/4; i += blockDim.x){ dst[i] = src[i]; }
        const float *gB = S.Bdu + (size_t)b*N*n, *gxp = S.xp + (size_t)b*N*n, *gd = S.dp + (size_t)b*N*n;
        for (int i = threadIdx.x; i < (N-1)*n; i += blockDim.x){ sB[i] = gB[i]; }
        for (int i = threadIdx.x; i < N*n; i += blockDim.x){ sxp[i] = gxp[i]; sd[i] = gd[i]; }
    }
    __syncthreads();
    const int a = a0 + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (a >= S.A){ return; }
    const float alpha = S.alpha[a];
    float *gx = S.x + ((size_t)b*S.A + a)*N*n;
    float xk = (l < n) ? sxp[l] : 0.f;      // x_a[0] = xp[0]
    if (l < n){ gx[l] = xk; }
    for (int k = 0; k < N-1; k++){
        const float *Ak = sA + (size_t)k*n*n;
        float dx = (l < n) ? SUB(xk, sxp[k*n + l]) : 0.f;
        float val = 0.f;
        #pragma unroll
        for (int i = 0; i < n; i++){ float dxi = __shfl_sync(FULL, dx, i); if (l < n){ val = FMA(Ak[l + n*i], dxi, val); } }
        if (l < n){
            const bool onb = (((k+1) % NBF) == 0) && (k < N-1);
            float tt = ADD(FMA(-alpha, sB[k*n+l], val), onb ? sd[k*n+l] : 0.f);
            xk = ADD(sxp[(k+1)*n + l], tt);
            gx[(k+1)*n + l] = xk;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// forward simulation + per-knot cost + defects
// grid = B*A/2 CTAs of 32*M threads: warp w simulates shooting interval w of TWO candidates (b, a) and (b, a+1), one per
// half-warp (SIM_LANES = 16 lanes cooperate on one trajectory; both halves run the same instruction stream).
// ------------------------------------------------------------------------------------------------------------------
constexpr int SIM_LANES = 16;
struct SimGroupSmem {
    kuka::FwdWsT<false> ws;
    float x[16], u[8], dx[16], qdd[8], xn[16], KT[kuka::NX*kuka::NU + 2];
};

// joint-space quadratic cost of one knot (plants/cost_arm.cuh:128-153), evaluated by one lane
__device__ __forceinline__ float cost_knot(const float *x, const float *u, const float *xg, bool final_knot, const DevState &S){
    float cost = 0.f;
    if (final_knot){
        for (int i = 0; i < kuka::NX; i++){ float dl = SUB(x[i], xg[i]); cost = FMA(MUL(i < kuka::NB ? S.QF1 : S.QF2, dl), dl, cost); }
    } else {
        for (int i = 0; i < kuka::NX; i++){ float dl = SUB(x[i], xg[i]); cost = FMA(MUL(i < kuka::NB ? S.Q1 : S.Q2, dl), dl, cost); }
        for (int i = 0; i < kuka::NU; i++){ cost = FMA(MUL(S.R, u[i]), u[i], cost); }
    }
    return MUL(0.5f, cost);
}

__global__ void sim_kernel(DevState S, int b0){
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int n = kuka::NX, m = kuka::NU, LANES = SIM_LANES, GPW = 32 / SIM_LANES;
    float *sI = reinterpret_cast<float*>(smem_raw);            // 252
    float *sTb = sI + 36*kuka::NB;                             // 252
    float *sxg = sTb + 36*kuka::NB;                            // 16
    SimGroupSmem *gsm = reinterpret_cast<SimGroupSmem*>(sxg + 16);
    const int apb = (S.A + GPW - 1) / GPW;                     // CTAs per problem
    const int b = b0 + blockIdx.x / apb;
    if (S.done[b]){ return; }
    for (int i = threadIdx.x; i < 36*kuka::NB; i += blockDim.x){ sI[i] = S.I[i]; sTb[i] = S.Tbody[i]; }
    if (threadIdx.x < n){ sxg[threadIdx.x] = S.xGoal[b*n + threadIdx.x]; }
    __syncthreads();
    const int w = threadIdx.x >> 5, grp = (threadIdx.x & 31) / LANES, l = threadIdx.x & (LANES-1);
    if (w >= S.M){ return; }
    int a = (blockIdx.x % apb)*GPW + grp;
    const bool live = a < S.A;                                 // odd A: the last half-warp replays candidate A-1 without storing
    if (!live){ a = S.A - 1; }
    SimGroupSmem &s = gsm[w*GPW + grp];
    kuka::init_ws<LANES>(s.ws, nullptr, sTb);
    const int N = S.N, NBF = N / S.M, kStart = w*NBF, iters = (w < S.M - 1) ? NBF : NBF - 1;
    const float alpha = S.alpha[a], dt = S.dt;
    float *gx = S.x + ((size_t)b*S.A + a)*N*n, *gu = S.u + ((size_t)b*S.A + a)*N*m, *gdd = S.d + ((size_t)b*S.A + a)*N*n;
    float *gc = S.costk + ((size_t)b*S.A + a)*N;
    const float *gxp = S.xp + (size_t)b*N*n, *gup = S.up + (size_t)b*N*m, *gKT = S.KT + (size_t)b*N*n*m, *gdu = S.du + (size_t)b*N*m;
    // state at the start of the interval (left there by the sweep)
    if (l < n){ s.x[l] = gx[kStart*n + l]; }
    // prefetch registers for knot kStart
    constexpr int KTQ = (n*m + LANES - 1) / LANES;
    float rKT[KTQ], rdu = 0.f, rxp = 0.f, rup = 0.f;
    #pragma unroll
    for (int q = 0; q < KTQ; q++){ const int i = l + LANES*q; rKT[q] = (i < n*m) ? gKT[(size_t)kStart*n*m + i] : 0.f; }
    if (l < m){ rdu = gdu[kStart*m + l]; rup = gup[kStart*m + l]; }
    if (l < n){ rxp = gxp[kStart*n + l]; }
    __syncwarp();
    #pragma unroll 1
    for (int kk = 0; kk < iters; kk++){
        const int k = kStart + kk;
        // stage this knot's feedback data, start fetching the next knot's
        #pragma unroll
        for (int q = 0; q < KTQ; q++){ const int i = l + LANES*q; if (i < n*m){ s.KT[i] = rKT[q]; } }
        if (l < n){ s.dx[l] = SUB(s.x[l], rxp); }
        const float du_k = rdu, up_k = rup;
        if (kk + 1 < iters){
            #pragma unroll
            for (int q = 0; q < KTQ; q++){ const int i = l + LANES*q; rKT[q] = (i < n*m) ? gKT[(size_t)(k+1)*n*m + i] : 0.f; }
            if (l < m){ rdu = gdu[(k+1)*m + l]; rup = gup[(k+1)*m + l]; }
            if (l < n){ rxp = gxp[(k+1)*n + l]; }
        }
        __syncwarp();
        // u = up - (alpha du + K dx)          (fpHelpers.cuh:210-219)
        if (l < m){
            float Kdx = 0.f;
            #pragma unroll
            for (int c = 0; c < n; c++){ Kdx = FMA(s.KT[c + l*n], s.dx[c], Kdx); }
            const float uu = SUB(up_k, FMA(alpha, du_k, Kdx));
            s.u[l] = uu; if (live){ gu[k*m + l] = uu; }
        }
        __syncwarp();
        kuka::forward<LANES, false>(s.ws, nullptr, sI, s.x, s.u, s.qdd);
        // Euler step (integrators.cuh:31-35)
        if (l < kuka::NB){ s.xn[l] = FMA(dt, s.x[l+kuka::NB], s.x[l]); s.xn[l+kuka::NB] = FMA(dt, s.qdd[l], s.x[l+kuka::NB]); }
        __syncwarp();
        if (kk < NBF - 1){
            if (l < n){ const float v = s.xn[l]; s.x[l] = v; if (live){ gx[(k+1)*n + l] = v; } }
        } else if (w < S.M - 1){
            // last step of a non-final interval: defect against the next interval's start state (fpHelpers.cuh:255-258)
            if (l < n && live){ gdd[((w+1)*NBF-1)*n + l] = SUB(s.xn[l], gx[(k+1)*n + l]); }
        }
        __syncwarp();
    }
    // u[N-1] is never simulated and stays the accepted one
    if (w == S.M - 1 && live && l < m){ gu[(N-1)*m + l] = gup[(N-1)*m + l]; }
    __syncwarp();
    // per-knot costs of this interval, one knot per lane (the trajectory was just written by this group; the final knot
    // belongs to the last interval)
    if (live){
        const int kEnd = (w == S.M - 1) ? N : kStart + NBF;
        for (int k = kStart + l; k < kEnd; k += LANES){ gc[k] = cost_knot(gx + k*n, gu + k*m, sxg, k == N - 1, S); }
    }
}

// per-knot costs of the initial trajectory (initAlgGPU's costKern<<<1,N>>>, nisInitHelpers.cuh:385) into candidate slot 0
__global__ void init_cost_kernel(DevState S){
    const int b = blockIdx.x, k = threadIdx.x, n = S.n, m = S.m;
    if (k >= S.N){ return; }
    const float *x = S.xp + ((size_t)b*S.N + k)*n, *u = S.up + ((size_t)b*S.N + k)*m;
    S.costk[((size_t)b*S.A + 0)*S.N + k] = cost_knot(x, u, S.xGoal + b*n, k == S.N - 1, S);
}

// ------------------------------------------------------------------------------------------------------------------
// selection: J[a] (reduceSum tree order), dT[a], line search, accept/reject, rho schedule.  grid = B CTAs, one warp per alpha.
// mode 0: iteration;  mode 1: initialisation (prevJ, Jout[0], alphaOut[0])
// ------------------------------------------------------------------------------------------------------------------
__global__ void select_kernel(DevState S, int mode, int b0){
    extern __shared__ __align__(16) float ssel[];          // [A][N] costs + [A] J + [A] dT
    const int b = b0 + blockIdx.x, N = S.N, A = S.A, n = S.n;
    if (S.done[b]){ return; }
    const int a = threadIdx.x >> 5, l = threadIdx.x & 31;
    float *sJ = ssel + (size_t)A*N, *sdT = sJ + A;
    const int nA = (mode == 1) ? 1 : A;
    if (a < nA){
        float *v = ssel + (size_t)a*N; const float *gc = S.costk + ((size_t)b*A + a)*N;
        for (int i = l; i < N; i += 32){ v[i] = ADD(0.f, gc[i]); }
        __syncwarp();
        for (int st = N/2; st >= 2; st >>= 1){
            for (int i = l; i < st; i += 32){ v[i] = ADD(v[i], v[i+st]); }
            __syncwarp();
        }
        if (l == 0){ sJ[a] = ADD(v[0], v[1]); }
        // defect: max over the M-1 interval boundaries of the L1 norm (fpHelpers.cuh:94-111)
        float dmax = 0.f;
        if (mode == 0 && l < S.M - 1){
            const int NBF = N / S.M; const float *dk = S.d + (((size_t)b*A + a)*N + (l+1)*NBF - 1)*n;
            float acc = 0.f; for (int c = 0; c < n; c++){ acc = ADD(acc, fabsf(dk[c])); }
            dmax = acc;
        }
        for (int o = 16; o >= 1; o >>= 1){ dmax = fmaxf(dmax, __shfl_xor_sync(FULL, dmax, o)); }
        if (l == 0){ sdT[a] = dmax; }
    }
    __syncthreads();
    if (threadIdx.x != 0){ return; }
    float *Jout = S.Jout + (size_t)b*(S.max_iter+1); int *alphaOut = S.alphaOut + (size_t)b*(S.max_iter+1);
    if (mode == 1){
        float pj = ADD(sJ[0], S.two_tol);                  // nisInitHelpers.cuh:393
        S.prevJ[b] = pj; Jout[0] = SUB(pj, S.two_tol); alphaOut[0] = -1;
        return;
    }
    for (int i = 0; i < A; i++){ S.J[(size_t)b*A + i] = sJ[i]; S.dT[(size_t)b*A + i] = sdT[i]; }
    // expected reduction summed over the time blocks in order (fpHelpers.cuh:376)
    float *dJe = S.dJexp + (size_t)b*2*S.M;
    float e0 = dJe[0], e1 = dJe[1];
    for (int i = 1; i < S.M; i++){ e0 = ADD(e0, dJe[2*i]); e1 = ADD(e1, dJe[2*i+1]); }
    dJe[0] = e0; dJe[1] = e1;
    // line search (fpHelpers.cuh:395-408): host arithmetic in the reference, so nothing is fused here
    const float prevJ = S.prevJ[b];
    float dJ = -1.f, z = 0.f; int alphaIndex = S.alphaIndex[b], ignore = S.ignore_defect[b];
    for (int i = 0; i < A; i++){
        float cdJ = SUB(prevJ, sJ[i]); bool JFlag = cdJ >= 0.f && cdJ > dJ;
        float al = S.alpha[i];
        float den = ADD(MUL(al, e0), MUL(MUL(MUL(0.5f, al), al), e1));
        float cz = DIV(cdJ, den); bool zFlag = (S.exp_red_min < cz && cz < S.exp_red_max);
        bool dFlag = (S.M == 1 || ignore) ? true : (sdT[i] < S.max_defect);
        if (JFlag && zFlag && dFlag){ if (sdT[i] < S.max_defect){ ignore = 0; } alphaIndex = i; dJ = cdJ; z = cz; }
    }
    S.z[b] = z; S.ignore_defect[b] = ignore;
    // accept / reject (nisInitHelpers.cuh:493-516)
    int iter = S.iter[b]; float rho = S.rho[b], drho = S.drho[b]; int done = 0, accepted;
    if (dJ < 0.f){
        drho = fmaxf(MUL(drho, S.rho_factor), S.rho_factor); rho = fminf(MUL(rho, drho), S.rho_max);
        alphaIndex = 0; alphaOut[iter] = -1; Jout[iter] = prevJ; accepted = 0;
    } else {
        drho = fminf(DIV(drho, S.rho_factor), S.inv_rho_factor); rho = fmaxf(MUL(rho, drho), S.rho_min);
        dJ = DIV(dJ, prevJ); S.prevJ[b] = sJ[alphaIndex]; alphaOut[iter] = alphaIndex; Jout[iter] = sJ[alphaIndex]; accepted = 1;
        if (dJ < S.tol_cost){ done = 1; }
    }
    if (!done){ if (iter == S.max_iter){ done = 1; } else { iter += 1; } }
    S.dJ[b] = dJ; S.rho[b] = rho; S.drho[b] = drho; S.alphaIndex[b] = alphaIndex; S.accepted[b] = accepted; S.iter[b] = iter;
    if (done){ S.done[b] = 1; S.final_src[b] = accepted ? alphaIndex : -1; atomicSub(S.n_active, 1); }
}

// ------------------------------------------------------------------------------------------------------------------
// next-iteration setup: one warp per (problem, knot).  Hands the accepted candidate over to (xp,up,dp), keeps the
// previous one in xp2, and refreshes AB (analytic Euler gradient), g (and H when write_H) at the accepted trajectory.
// mode 1 = initialisation: the trajectory is already in xp/up, xp2 <- xp.
// ------------------------------------------------------------------------------------------------------------------
constexpr int NIS_LANES = 32;
struct NisGroupSmem {
    kuka::FwdWs ws; kuka::GradWs gs;
    float x[16], u[8], qdd[8], dqdd[3*kuka::NB*kuka::NB + 1];
};
constexpr int NIS_WARPS = 1;

__global__ void __launch_bounds__(32*NIS_WARPS) nis_kernel(DevState S, int mode, int write_H, int b0, int nb){
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int n = kuka::NX, m = kuka::NU, nm = n + m, np = kuka::NB, LANES = NIS_LANES, GPW = 32 / NIS_LANES;
    float *sI = reinterpret_cast<float*>(smem_raw); float *sTb = sI + 36*kuka::NB;
    NisGroupSmem *gsm = reinterpret_cast<NisGroupSmem*>(sTb + 36*kuka::NB);
    const int w = threadIdx.x >> 5, grp = (threadIdx.x & 31) / LANES, l = threadIdx.x & (LANES-1);
    const int gk = (blockIdx.x*NIS_WARPS + w)*GPW + grp, N = S.N;        // N is even: the groups of one warp share the problem
    const int b = b0 + gk / N, k = gk % N;
    for (int i = threadIdx.x; i < 36*kuka::NB; i += blockDim.x){ sI[i] = S.I[i]; sTb[i] = S.Tbody[i]; }
    __syncthreads();
    if (b >= b0 + nb || S.done[b]){ return; }
    NisGroupSmem &s = gsm[w*GPW + grp];
    kuka::init_ws<LANES>(s.ws, &s.gs, sTb);
    float *gxp = S.xp + ((size_t)b*N + k)*n, *gup = S.up + ((size_t)b*N + k)*m, *gdp = S.dp + ((size_t)b*N + k)*n, *gxp2 = S.xp2 + ((size_t)b*N + k)*n;
    const bool acc = (mode == 0) && S.accepted[b];
    const int a = S.alphaIndex[b];
    const float *cx = S.x + (((size_t)b*S.A + a)*N + k)*n, *cu = S.u + (((size_t)b*S.A + a)*N + k)*m, *cd = S.d + (((size_t)b*S.A + a)*N + k)*n;
    if (l < n){
        const float xold = gxp[l]; gxp2[l] = xold;                 // xp2 <- xp (fpHelpers.cuh:371 / nisInitHelpers.cuh:379)
        const float xv = acc ? cx[l] : xold; s.x[l] = xv;
        if (acc){ gxp[l] = xv; gdp[l] = cd[l]; }
    }
    if (l < m){ const float uv = acc ? cu[l] : gup[l]; s.u[l] = uv; if (acc){ gup[l] = uv; } }
    __syncwarp();
    // cost gradient (plants/cost_arm.cuh:156-202)
    const float *xg = S.xGoal + b*n;
    float *gg = S.g + ((size_t)b*N + k)*nm;
    const bool fin = (k == N - 1);
    for (int e = l; e < nm; e += LANES){
        if (e < n){ gg[e] = MUL(fin ? (e < np ? S.QF1 : S.QF2) : (e < np ? S.Q1 : S.Q2), SUB(s.x[e], xg[e])); }
        else { gg[e] = fin ? 0.f : MUL(S.R, s.u[e-n]); }
    }
    if (write_H){
        float *gH = S.H + ((size_t)b*N + k)*nm*nm;
        for (int e = l; e < nm*nm; e += LANES){
            const int i = e / nm, j = e % nm; float v = 0.f;
            if (fin){ if (i < n && j < n){ v = (i != j) ? 0.f : (i < np ? S.QF1 : S.QF2); } }
            else { v = (i != j) ? 0.f : (i < np ? S.Q1 : (i < n ? S.Q2 : S.R)); }
            gH[e] = v;
        }
    }
    // integrator gradient AB = [I 0] + dt [0 I 0 ; dqdd]   (integrators.cuh:15-17,38-53); the final knot has none
    // (its group still runs the collective code so that both halves of a warp stay in lockstep, but stores nothing)
    kuka::gradient<LANES>(s.ws, s.gs, sI, s.x, s.u, s.qdd, s.dqdd);
    if (fin){ return; }
    float *gAB = S.AB + ((size_t)b*N + k)*n*nm;
    const float dt = S.dt;
    for (int e = l; e < n*nm; e += LANES){
        const int ky = e / n, kx = e % n;
        const float dxd = kx < np ? ((kx + np == ky) ? 1.f : 0.f) : s.dqdd[(ky-1)*np + kx];
        gAB[e] = FMA(dt, dxd, (ky == kx) ? 1.f : 0.f);
    }
}

// final trajectory of every problem (storeVarsGPU, nisInitHelpers.cuh:739-750): accepted candidate of the last iteration or xp/up
__global__ void store_kernel(DevState S, float *x_out, float *u_out, int *iters_out){
    const int b = blockIdx.x, N = S.N, n = S.n, m = S.m;
    const int src = S.final_src[b];
    const float *sx = (src >= 0) ? S.x + ((size_t)b*S.A + src)*N*n : S.xp + (size_t)b*N*n;
    const float *su = (src >= 0) ? S.u + ((size_t)b*S.A + src)*N*m : S.up + (size_t)b*N*m;
    for (int i = threadIdx.x; i < N*n; i += blockDim.x){ x_out[(size_t)b*N*n + i] = sx[i]; }
    for (int i = threadIdx.x; i < N*m; i += blockDim.x){ u_out[(size_t)b*N*m + i] = su[i]; }
    if (threadIdx.x == 0 && iters_out){ iters_out[b] = S.iter[b]; }
}

// ------------------------------------------------------------------------------------------------------------------
// plug-in unit kernels (one group per sample, same group widths as the production kernels)
// ------------------------------------------------------------------------------------------------------------------
__global__ void unit_dynamics_kernel(const float *I, const float *Tbody, const float *x, const float *u, int nsamp, float *qdd){
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LANES = SIM_LANES, GPW = 32 / SIM_LANES;
    float *sI = reinterpret_cast<float*>(smem_raw); float *sTb = sI + 36*kuka::NB;
    SimGroupSmem *gsm = reinterpret_cast<SimGroupSmem*>(sTb + 36*kuka::NB);
    const int grp = (threadIdx.x & 31) / LANES, l = threadIdx.x & (LANES-1);
    for (int i = threadIdx.x; i < 36*kuka::NB; i += blockDim.x){ sI[i] = I[i]; sTb[i] = Tbody[i]; }
    __syncthreads();
    SimGroupSmem &s = gsm[grp];
    kuka::init_ws<LANES>(s.ws, nullptr, sTb);
    for (int k0 = blockIdx.x*GPW; k0 < nsamp; k0 += gridDim.x*GPW){
        const int k = k0 + grp < nsamp ? k0 + grp : nsamp - 1;         // tail: replay the last sample
        if (l < kuka::NX){ s.x[l] = x[k*kuka::NX + l]; } if (l < kuka::NU){ s.u[l] = u[k*kuka::NU + l]; }
        __syncwarp();
        kuka::forward<LANES, false>(s.ws, nullptr, sI, s.x, s.u, s.qdd);
        if (l < kuka::NB){ qdd[k*kuka::NB + l] = s.qdd[l]; }
        __syncwarp();
    }
}
__global__ void unit_gradient_kernel(const float *I, const float *Tbody, const float *x, const float *u, int nsamp, float dt, float *AB, float *qdd){
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int n = kuka::NX, nm = kuka::NX + kuka::NU, np = kuka::NB, LANES = NIS_LANES, GPW = 32 / NIS_LANES;
    float *sI = reinterpret_cast<float*>(smem_raw); float *sTb = sI + 36*kuka::NB;
    NisGroupSmem *gsm = reinterpret_cast<NisGroupSmem*>(sTb + 36*kuka::NB);
    const int grp = (threadIdx.x & 31) / LANES, l = threadIdx.x & (LANES-1);
    for (int i = threadIdx.x; i < 36*kuka::NB; i += blockDim.x){ sI[i] = I[i]; sTb[i] = Tbody[i]; }
    __syncthreads();
    NisGroupSmem &s = gsm[grp];
    kuka::init_ws<LANES>(s.ws, &s.gs, sTb);
    for (int k0 = blockIdx.x*GPW; k0 < nsamp; k0 += gridDim.x*GPW){
        const int k = k0 + grp < nsamp ? k0 + grp : nsamp - 1;
        if (l < kuka::NX){ s.x[l] = x[k*kuka::NX + l]; } if (l < kuka::NU){ s.u[l] = u[k*kuka::NU + l]; }
        __syncwarp();
        kuka::gradient<LANES>(s.ws, s.gs, sI, s.x, s.u, s.qdd, s.dqdd);
        for (int e = l; e < n*nm; e += LANES){
            const int ky = e / n, kx = e % n;
            const float dxd = kx < np ? ((kx + np == ky) ? 1.f : 0.f) : s.dqdd[(ky-1)*np + kx];
            AB[(size_t)k*n*nm + e] = FMA(dt, dxd, (ky == kx) ? 1.f : 0.f);
        }
        if (l < np){ qdd[k*np + l] = s.qdd[l]; }
        __syncwarp();
    }
}

} // namespace pddp
