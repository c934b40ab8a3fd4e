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
//   Pp<-P / pp<-p     -> ping-pong of two buffers per problem, indexed by the parity of its iteration counter
//   alpha broadcast   -> sweep/sim read the accepted trajectory (xp,up,dp) as the base of every candidate
//   reject restore    -> nothing to restore (xp,up,dp are only overwritten on accept)
#pragma once
#include "pddp_math.cuh"
#include "plant_kuka.cuh"

#include "dev_state.cuh"
#include "bp_warp.cuh"
namespace pddp {

// ------------------------------------------------------------------------------------------------------------------
// backward pass
// ------------------------------------------------------------------------------------------------------------------
// One CTA per (problem, time block): BP_THREADS compute threads (8 warps) plus one producer warp.  The per-knot inputs
// (AB, H, g: 3 KB) stream through a BP_STAGES-deep ring in shared memory filled by TMA bulk copies (cp.async.bulk +
// mbarrier) that the producer warp keeps BP_STAGES knots ahead; it never joins the compute warps' named barrier.
// Every knot is five barrier-separated stages on the critical path
//     A   AB2 = AB'(P + rho I)                       147 threads, 2 outputs each
//     B   Huu (warp 0, directly in the register layout of the elimination) -> 7x7 Gauss-Jordan in registers (shuffles),
//         while warps 1..7 assemble the other 392 entries of H and g
//     C   K = Huu^-1 Hux, du = Huu^-1 gu
//     D   T = K'Huu - Hxu;  beside it A-BK, B du, KT, du -> HBM and the expected-reduction sums
//     E   P, p of the previous knot
// The stages are bound by shared-memory wavefronts and dependent-issue latency, not by FP32 rate, so each thread owns a
// small tile of outputs that share operand vectors, all operands are fetched with 8/16-byte loads before the first FMA
// (7-vectors live in rows padded to 8 floats), and lanes of a warp read either consecutive or identical rows
// (conflict-free or broadcast).  The order of the fused multiply-adds inside every output is the reference's.
constexpr int BP_THREADS = 256;              // compute threads; one more warp only feeds the TMA ring
constexpr int BP_CTA = BP_THREADS + 32;
constexpr int BP_STAGES = 4;
// 7-vectors (rows of K, T, Hux, Huu, Hinv) are read as two 16-byte loads; rows 12 floats apart put the 8 lanes of a quarter
// warp on 8 disjoint bank quads (rows 8 floats apart collide two by two)
constexpr int BP_RS = 12;
template <int n, int m>
struct __align__(16) BpSmem {
    static constexpr int nm = n + m;
    float AB[BP_STAGES][AB_STRIDE];      // ring of knot inputs
    float Hc[BP_STAGES][H_STRIDE];
    float gc[BP_STAGES][G_STRIDE];
    unsigned long long full[BP_STAGES], empty[BP_STAGES];
    float P[n*n];                        // P[ky*n + j]
    float Pr[n*n];                       // P + rho on the diagonal: the (P + rho I) operand of the u-rows of AB'(.) (bpHelpers.cuh:62)
    float p[n + 2];
    float p0[n + 2];                     // p as stage E wrote it (stage A of the next knot adds the defect term in place)
    float AB2[nm*n + 2];                 // AB2[kx*n + ky] = sum_j AB[kx*n+j] (P+rho)[ky*n+j]   (kx < nm, ky < n)
    float H[nm*nm + 3];                  // H[kx + nm*ky]
    float g[nm + 3];
    float Hux[n*BP_RS];                      // Hux[ky*BP_RS + j] = H[(n+j) + nm*ky], rows of 7 in BP_RS floats
    float Huu[m*BP_RS];                      // Huu[c*BP_RS + l]  = H[(n+l) + nm*(n+c)]
    float Hinv[m*BP_RS];                     // Hinv[l*BP_RS + j] = (Huu^-1)(l, j)
    float K[n*BP_RS];                        // K[ky*BP_RS + kx]  = K(kx, ky)  (reference: K[kx + ky*m])
    float T[n*BP_RS];                        // T[kx*BP_RS + j]   = (K'Huu - Hxu)(kx, j)
    float du[8];
    float dx[n + 2];
    float dJ[2*m + 2];
};

#ifdef PDDP_BP_TRACE
// stage clocks of thread 0 are kept in registers and written once per knot (a store in front of a stage's loads would delay them)
#define BP_TRACE(slot) do { tr[slot] = clock64(); } while (0)
#define BP_TRACE_DECL long long tr[12] = {0,0,0,0,0,0,0,0,0,0,0,0}
#define BP_TRACE_FLUSH() do { if (tracer && iter >= iterCount - 7){ for (int q_ = 0; q_ < 12; q_++){ S.dbg[(iterCount - iter)*12 + q_] = tr[q_]; } } } while (0)
#else
#define BP_TRACE(slot) do { } while (0)
#define BP_TRACE_DECL
#define BP_TRACE_FLUSH() do { } while (0)
#endif
#ifndef PDDP_BP_SKIP
#define PDDP_BP_SKIP 0      // timing ablations only (tools/bp_ablate.sh): bit q set = stage q's body is left out
#endif

// K-term fused chain val = sum_j a[j*sa] * b[j*sb], j ascending, operands loaded first
template <int K>
__device__ __forceinline__ float chain(const float *a, int sa, const float *b, int sb){
    float x[K], y[K];
    #pragma unroll
    for (int j = 0; j < K; j++){ x[j] = a[j*sa]; y[j] = b[j*sb]; }
    float val = 0.f;
    #pragma unroll
    for (int j = 0; j < K; j++){ val = FMA(x[j], y[j], val); }
    return val;
}
// 14 contiguous floats at an 8-byte aligned shared address / 8 floats at a 16-byte aligned one
__device__ __forceinline__ void ld14(float (&x)[14], const float *p){
    const float2 *q = reinterpret_cast<const float2*>(p);
    #pragma unroll
    for (int i = 0; i < 7; i++){ const float2 v = q[i]; x[2*i] = v.x; x[2*i+1] = v.y; }
}
__device__ __forceinline__ void ld8(float (&x)[8], const float *p){
    const float4 *q = reinterpret_cast<const float4*>(p);
    const float4 a = q[0], b = q[1];
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
// keeps the operand loads of a tile ahead of its first FMA (all LDS in flight together)
#define SCHED_FENCE() asm volatile("" ::: "memory")
template <int K, int KP>
__device__ __forceinline__ float dotr(const float (&x)[KP], const float (&y)[KP]){
    float val = 0.f;
    #pragma unroll
    for (int j = 0; j < K; j++){ val = FMA(x[j], y[j], val); }
    return val;
}

template <int n, int m>
__global__ void __launch_bounds__(BP_CTA, 2) bp_kernel(DevState S, int b0){
    static_assert(n == 14 && m == 7, "the thread maps of the backward pass are laid out for the 14-state, 7-control arm");
    constexpr int nm = n + m, oB = n*n;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BpSmem<n,m> &s = *reinterpret_cast<BpSmem<n,m>*>(smem_raw);
    const int b = b0 + blockIdx.x / S.M, block = blockIdx.x % S.M, t = threadIdx.x;
    if (S.done[b]){ return; }
    const int cur = S.iter[b] & 1;                                // this iteration's P buffer (see DevState::Pbuf)
    const int N = S.N, NBB = N / S.M;
    const float rho = S.rho[b];
    const size_t bN = (size_t)b*N;

    int ks = NBB*(block+1) - 1, iterCount;
    const bool last_block = (ks == N - 1);
    if (last_block){ ks--; iterCount = NBB - 2; } else { iterCount = NBB - 1; }
    const int nknots = iterCount + 1, ks0 = ks;                   // knots ks0, ks0-1, ... ks0-nknots+1
    // ---- ring setup; warp 8 is the producer: it keeps BP_STAGES knots of (AB, H, g) in flight and never joins a barrier again
    if (t == 0){
        for (int q = 0; q < BP_STAGES; q++){ mbar_init(&s.full[q], 1); mbar_init(&s.empty[q], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t >= BP_THREADS){
        if (t == BP_THREADS){
            for (int i = 0; i < nknots; i++){       // knot number i (processing order) goes to slot i % BP_STAGES
                const int slot = i % BP_STAGES; const size_t k = bN + ks0 - i;
                if (i >= BP_STAGES){ mbar_wait_sleep(&s.empty[slot], ((i / BP_STAGES) - 1) & 1); }
                mbar_expect_tx(&s.full[slot], (AB_STRIDE + H_STRIDE + G_STRIDE)*4);
                tma_load_1d(s.AB[slot], S.AB + k*AB_STRIDE, AB_STRIDE*4, &s.full[slot]);
                tma_load_1d(s.Hc[slot], S.H + k*H_STRIDE, H_STRIDE*4, &s.full[slot]);
                tma_load_1d(s.gc[slot], S.g + k*G_STRIDE, G_STRIDE*4, &s.full[slot]);
            }
        }
        return;
    }
    float dJ0 = 0.f, dJ1 = 0.f;          // threads 210..216: running sums of du*gu and du*(Huu du) (bpHelpers.cuh:327-328)
    if (last_block){
        // final block: Hxx[N-1] -> P[N-2], gx[N-1] -> p[N-2]  (bpHelpers.cuh:362-367)
        const size_t kN = bN + N - 1; float *gP = S.Pbuf[cur] + (kN-1)*n*n, *gp = S.pbuf[cur] + (kN-1)*n;
        if (t < n*n){ const int kx = t % n, ky = t / n; const float v = MUL(1.0f, S.H[kN*H_STRIDE + kx + nm*ky]); s.P[t] = v; s.Pr[t] = (kx == ky) ? ADD(v, rho) : v; gP[t] = v; }
        if (t >= 224 && t < 224 + n){ const int r = t - 224; const float v = MUL(1.0f, S.g[kN*G_STRIDE + r]); s.p[r] = v; gp[r] = v; }
    } else {
        // other blocks: start from the previous iteration's P,p at the block boundary, shifted to the new linearisation point
        const float *gPp = S.Pbuf[cur^1] + (bN + ks)*n*n;
        if (t < n*n){ const float v = gPp[t]; s.P[t] = v; s.Pr[t] = (t % n == t / n) ? ADD(v, rho) : v; }
        if (t >= 224 && t < 224 + n){ const int r = t - 224; s.dx[r] = SUB(S.xp[(bN + ks + 1)*n + r], S.xp2[(bN + ks + 1)*n + r]); }
        bar_sync(1, BP_THREADS);
        if (t < n){ s.p[t] = FMA(1.0f, chain<n>(&s.P[t], n, s.dx, 1), S.pbuf[cur^1][(bN + ks)*n + t]); }
    }
    // ---- per-thread tile coordinates of every stage (fixed for the whole block)
    // stage A, t < 56: column groups {0-2},{3-5},{6-8},{9-11},{12,13} | {14-16},{17-19},{20} x 7 row pairs
    const int a_g = t & 7, a_kp = (t >> 3) % 7;
    const int a_kx0 = a_g < 5 ? 3*a_g : n + 3*(a_g - 5), a_cnt = (a_g == 4) ? 2 : (a_g == 7 ? 1 : 3);
    const int a_kx1 = a_cnt > 1 ? a_kx0 + 1 : a_kx0, a_kx2 = a_cnt > 2 ? a_kx0 + 2 : a_kx0;
    const int u = t - 32;                                         // stage B, warps 1..7
    const int b_kxt = (u >= 0 ? u : 0) % 7, b_kp = ((u >= 0 ? u : 0) / 7) % 7;      // region 1: rows < n, all columns: (column triple, row pair), u < 49
    const int b2 = (u >= 64 && u < 113) ? u - 64 : 0;             // region 2: rows >= n, columns < n (column pairs)
    const int b2_kxp = b2 % 7, b2_ky = n + b2 / 7;
    const int hl = min(t & 7, 6), hg = (t & 31) >> 3;             // Huu in warp 0: lane = l + 8*(column pair)
    const int c_kx = t % m, c_ky = (t / m) % n;                   // stage C, t < 98
    const int d_kx = t % n, d_ky = (t / n) % m;                   // stage D, T: t < 98
    const int dv = (t >= 32 && t < 130) ? t - 32 : 0;             // deferred A-BK: (kx, ky pair)
    const int d2_kx = dv % n, d2_kyp = dv / n;
    const int te = (t >= 32) ? t - 32 : 0;                        // stage E: warps 1..7 (warp 0 is the slowest to leave stage D)
    const int e_kxp = te % 7, e_kyp = (te / 7) % 7;               // 2 x 2 tiles of P, te < 49
    // Results of a knot go to HBM one knot late, from warps 1..7 while warp 0 eliminates the next Huu: a barrier does not
    // release before the global stores issued in front of it have been performed, which would put a memory round trip on
    // the critical path of stages D and E.  K, du, P, p0 and the knot's ring slot are all still intact at that point.
    auto flush_outputs = [&](size_t kp, const float *sABp, bool with_P){
        const int u_ = t - 32;
        if (u_ < 0){ return; }
        if (S.M > 1){
            if (u_ < 98){
                float bb[8], k0[8], k1[8];
                #pragma unroll
                for (int j = 0; j < m; j++){ bb[j] = sABp[oB + d2_kx + n*j]; }
                bb[7] = 0.f;
                ld8(k0, s.K + (2*d2_kyp)*BP_RS); ld8(k1, s.K + (2*d2_kyp+1)*BP_RS);
                const float a0 = sABp[d2_kx + n*(2*d2_kyp)], a1 = sABp[d2_kx + n*(2*d2_kyp+1)];
                SCHED_FENCE();
                S.ApBK[kp*n*n + d2_kx + n*(2*d2_kyp)] = SUB(a0, dotr<m>(bb, k0));
                S.ApBK[kp*n*n + d2_kx + n*(2*d2_kyp+1)] = SUB(a1, dotr<m>(bb, k1));
            } else if (u_ >= 196 && u_ < 196 + n){
                const int kx = u_ - 196; S.Bdu[kp*n + kx] = chain<m>(&sABp[oB+kx], n, s.du, 1);
            }
        }
        if (u_ >= 98 && u_ < 196){ const int e = u_ - 98, kx = e % n, ky = e / n; S.KT[kp*n*m + e] = s.K[kx*BP_RS + ky]; }
        if (u_ >= 196 + n && u_ < 196 + n + m){ const int r = u_ - (196 + n); S.du[kp*m + r] = s.du[r]; }
        if (with_P){
            if (u_ < n*n){ S.Pbuf[cur][(kp-1)*n*n + u_] = s.P[u_]; }
            else if (u_ < n*n + n){ S.pbuf[cur][(kp-1)*n + u_ - n*n] = s.p0[u_ - n*n]; }
        }
    };
#ifdef PDDP_BP_TRACE
    int tracer; asm volatile("mov.u32 %0, %1;" : "=r"(tracer) : "r"((int)(blockIdx.x == 0 && threadIdx.x == PDDP_TRACE_THREAD)));
#endif
    bar_sync(1, BP_THREADS);
    #pragma unroll 1
    for (int iter = iterCount, i = 0; iter >= 0; iter--, ks--, i++){
        const int slot = i % BP_STAGES;
        BP_TRACE_DECL;
        BP_TRACE(0);
        mbar_wait(&s.full[slot], (i / BP_STAGES) & 1);
        BP_TRACE(1);
        const float *sAB = s.AB[slot], *bH = s.Hc[slot], *bg = s.gc[slot];
        const size_t kk = bN + ks;
        const bool boundary = S.M > 1 && iter == NBB - 1;      // block-local defect-boundary test of the reference (bpHelpers.cuh:73)
        // ---- stage A: AB2 = AB'(P + rho I[u rows]);  p += P d on the block-local defect boundary (bpHelpers.cuh:54-81)
        if (PDDP_BP_SKIP & 1){ } else
        if (t < 56){
            // tile = (group of up to 3 columns kx of AB) x (2 rows ky of P): 5 operand vectors for 6 outputs.  The column
            // groups do not straddle the x/u boundary: x-rows use P (the reference adds +0, an identity), u-rows use P + rho
            // on the diagonal.
            float x0[14], x1[14], x2[14], y0[14], y1[14];
            const float *Pq = (a_kx0 >= n) ? s.Pr : s.P;
            ld14(x0, sAB + a_kx0*n); ld14(x1, sAB + a_kx1*n); ld14(x2, sAB + a_kx2*n);
            ld14(y0, Pq + (2*a_kp)*n); ld14(y1, Pq + (2*a_kp+1)*n);
            SCHED_FENCE();
            float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f, v20 = 0.f, v21 = 0.f;
            #pragma unroll
            for (int j = 0; j < n; j++){
                v00 = FMA(x0[j], y0[j], v00); v01 = FMA(x0[j], y1[j], v01);
                v10 = FMA(x1[j], y0[j], v10); v11 = FMA(x1[j], y1[j], v11);
                v20 = FMA(x2[j], y0[j], v20); v21 = FMA(x2[j], y1[j], v21);
            }
            *reinterpret_cast<float2*>(&s.AB2[a_kx0*n + 2*a_kp]) = make_float2(v00, v01);
            if (a_cnt > 1){ *reinterpret_cast<float2*>(&s.AB2[a_kx1*n + 2*a_kp]) = make_float2(v10, v11); }
            if (a_cnt > 2){ *reinterpret_cast<float2*>(&s.AB2[a_kx2*n + 2*a_kp]) = make_float2(v20, v21); }
        } else if (t >= 224 && t < 224 + n){
            const int r = t - 224; float val = 0.f;
            if (boundary){ val = chain<n>(S.dp + kk*n, 1, &s.P[r], n); }
            s.p[r] = ADD(s.p[r], val);
        }
        bar_sync(1, BP_THREADS);
        BP_TRACE(2);
        // ---- stage B: H = (AB2 AB)' + H_cost, g = AB'p + g_cost (bpHelpers.cuh:83-118); Huu^-1 (invHelpers.cuh)
        if ((PDDP_BP_SKIP & 4) && t < 32){ } else if ((PDDP_BP_SKIP & 2) && t >= 32){ } else
        if (t < 32){
            // lane l + 8*q computes Huu(l, 2q) and Huu(l, 2q+1); the rows are then gathered on lanes 0..6
            float x[14], y0[14], y1[14];
            const int c0 = min(2*hg, 6), c1 = min(2*hg + 1, 6);
            ld14(x, sAB + (n + hl)*n); ld14(y0, s.AB2 + (n + c0)*n); ld14(y1, s.AB2 + (n + c1)*n);
            const float q0 = bH[(n + hl) + nm*(n + c0)], q1 = bH[(n + hl) + nm*(n + c1)];
            SCHED_FENCE();
            float v0 = 0.f, v1 = 0.f;
            #pragma unroll
            for (int j = 0; j < n; j++){ v0 = FMA(y0[j], x[j], v0); v1 = FMA(y1[j], x[j], v1); }
            const float h0 = FMA(1.0f, v0, MUL(1.0f, q0)), h1 = FMA(1.0f, v1, MUL(1.0f, q1));
            if ((t & 7) < 7){
                s.H[(n + hl) + nm*(n + c0)] = h0; s.Huu[c0*BP_RS + hl] = h0;
                if (2*hg + 1 < m){ s.H[(n + hl) + nm*(n + c1)] = h1; s.Huu[c1*BP_RS + hl] = h1; }
            }
            float a[2*m];
            #pragma unroll
            for (int c = 0; c < m; c++){ a[c] = MUL(1.0f, __shfl_sync(FULL, (c & 1) ? h1 : h0, (t & 7) + 8*(c >> 1))); a[m + c] = (t == c) ? 1.f : 0.f; }
            BP_TRACE(3);
            float a_in[2*m];
            #pragma unroll
            for (int c = 0; c < 2*m; c++){ a_in[c] = a[c]; }
            if (!gauss_jordan_rows<m, 32, true>(a, t)){
                #pragma unroll
                for (int c = 0; c < 2*m; c++){ a[c] = a_in[c]; }
                gauss_jordan_rows<m, 32, false>(a, t);
            }
            if (t < m){
                *reinterpret_cast<float4*>(&s.Hinv[t*BP_RS]) = make_float4(a[m], a[m+1], a[m+2], a[m+3]);
                *reinterpret_cast<float4*>(&s.Hinv[t*BP_RS + 4]) = make_float4(a[m+4], a[m+5], a[m+6], 0.f);
            }
            BP_TRACE(4);
        } else if (u < 49){
            // region 1 (rows ky < n, all columns): tile = 2 rows of AB2 x 3 columns of AB
            float x0[14], x1[14], x2[14], y0[14], y1[14];
            const int kx0 = 3*b_kxt, r0 = 2*b_kp;
            ld14(x0, sAB + kx0*n); ld14(x1, sAB + (kx0+1)*n); ld14(x2, sAB + (kx0+2)*n);
            ld14(y0, s.AB2 + r0*n); ld14(y1, s.AB2 + (r0+1)*n);
            float q[6];
            #pragma unroll
            for (int c = 0; c < 3; c++){ q[c] = bH[kx0 + c + nm*r0]; q[3+c] = bH[kx0 + c + nm*(r0+1)]; }
            SCHED_FENCE();
            float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            #pragma unroll
            for (int j = 0; j < n; j++){
                v[0] = FMA(y0[j], x0[j], v[0]); v[1] = FMA(y0[j], x1[j], v[1]); v[2] = FMA(y0[j], x2[j], v[2]);
                v[3] = FMA(y1[j], x0[j], v[3]); v[4] = FMA(y1[j], x1[j], v[4]); v[5] = FMA(y1[j], x2[j], v[5]);
            }
            #pragma unroll
            for (int c = 0; c < 3; c++){
                const float h0 = FMA(1.0f, v[c], MUL(1.0f, q[c])), h1 = FMA(1.0f, v[3+c], MUL(1.0f, q[3+c]));
                s.H[kx0 + c + nm*r0] = h0; s.H[kx0 + c + nm*(r0+1)] = h1;
                if (kx0 + c >= n){ s.Hux[r0*BP_RS + kx0 + c - n] = h0; s.Hux[(r0+1)*BP_RS + kx0 + c - n] = h1; }
            }
        } else if (u >= 64 && u < 113){
            float y[14], x0[14], x1[14];
            ld14(y, s.AB2 + b2_ky*n); ld14(x0, sAB + (2*b2_kxp)*n); ld14(x1, sAB + (2*b2_kxp+1)*n);
            const float q0 = bH[2*b2_kxp + nm*b2_ky], q1 = bH[2*b2_kxp + 1 + nm*b2_ky];
            SCHED_FENCE();
            float v0 = 0.f, v1 = 0.f;
            #pragma unroll
            for (int j = 0; j < n; j++){ v0 = FMA(y[j], x0[j], v0); v1 = FMA(y[j], x1[j], v1); }
            s.H[2*b2_kxp + nm*b2_ky] = FMA(1.0f, v0, MUL(1.0f, q0)); s.H[2*b2_kxp + 1 + nm*b2_ky] = FMA(1.0f, v1, MUL(1.0f, q1));
        } else if (u >= 128 && u < 128 + nm){
            const int r = u - 128;
            float x[14], y[14];
            ld14(x, s.p); ld14(y, sAB + r*n);
            s.g[r] = FMA(1.0f, dotr<n>(x, y), MUL(1.0f, bg[r]));
        }
        if (i > 0){ flush_outputs(kk + 1, s.AB[(i - 1) % BP_STAGES], true); }
        bar_sync(1, BP_THREADS);
        BP_TRACE(5);
        // ---- stage C: K = Huu^-1 Hux, du = Huu^-1 gu (bpHelpers.cuh:206-220)
        if (PDDP_BP_SKIP & 8){ } else
        if (t < n*m){
            float hi[8], hx[8];
            ld8(hi, s.Hinv + c_kx*BP_RS); ld8(hx, s.Hux + c_ky*BP_RS);
            s.K[c_ky*BP_RS + c_kx] = MUL(1.0f, dotr<m>(hi, hx));
        } else if (t >= 128 && t < 128 + m){
            const int r = t - 128;
            float hi[8], gu[8];
            ld8(hi, s.Hinv + r*BP_RS);
            #pragma unroll
            for (int j = 0; j < m; j++){ gu[j] = s.g[n + j]; }
            gu[7] = 0.f;
            s.du[r] = ADD(MUL(1.0f, dotr<m>(hi, gu)), 0.f);
        }
        bar_sync(1, BP_THREADS);
        BP_TRACE(6);
        // ---- stage D: T = K'Huu - Hxu; A-BK, B du, KT, du to HBM; expected reduction (bpHelpers.cuh:223-240,282-331)
        const bool do_ctg = (iter != 0 || block != 0);
        if (PDDP_BP_SKIP & 16){ } else
        if (t < n*m){
            if (do_ctg){
                float k[8], h[8];
                ld8(k, s.K + d_kx*BP_RS); ld8(h, s.Huu + d_ky*BP_RS);
                const float hxu = s.H[d_kx + nm*(n + d_ky)];
                s.T[d_kx*BP_RS + d_ky] = SUB(dotr<m>(k, h), hxu);
            }
        } else if (t >= 196 + n && t < 196 + n + m){
            const int ind = t - (196 + n);
            float h[m];
            #pragma unroll
            for (int j = 0; j < m; j++){ h[j] = s.Huu[j*BP_RS + ind]; }
            float dot = 0.f;
            #pragma unroll
            for (int j = 0; j < m; j++){ dot = FMA(h[j], s.du[j], dot); }
            dJ0 = FMA(s.du[ind], s.g[n + ind], dJ0); dJ1 = FMA(s.du[ind], dot, dJ1);
        }
        bar_sync(1, BP_THREADS);
        BP_TRACE(7);
        // ---- stage E: cost-to-go of the previous knot (bpHelpers.cuh:242-276)
        if (do_ctg && !(PDDP_BP_SKIP & 32)){
            if (t >= 32 && t < 32 + 49){
                // tile = 2 x 2 entries of P: rows kx0, kx0+1 of T and K, rows ky0, ky0+1 of K and Hux
                const int kx0 = 2*e_kxp, ky0 = 2*e_kyp;
                float a0[8], a1[8], kx_0[8], kx_1[8], ky_0[8], ky_1[8], h0[8], h1[8];
                ld8(a0, s.T + kx0*BP_RS); ld8(a1, s.T + (kx0+1)*BP_RS); ld8(kx_0, s.K + kx0*BP_RS); ld8(kx_1, s.K + (kx0+1)*BP_RS);
                ld8(ky_0, s.K + ky0*BP_RS); ld8(ky_1, s.K + (ky0+1)*BP_RS); ld8(h0, s.Hux + ky0*BP_RS); ld8(h1, s.Hux + (ky0+1)*BP_RS);
                const float x00 = s.H[kx0 + ky0*nm], x10 = s.H[kx0 + 1 + ky0*nm], x01 = s.H[kx0 + (ky0+1)*nm], x11 = s.H[kx0 + 1 + (ky0+1)*nm];
                SCHED_FENCE();
                float v00 = 0.f, v10 = 0.f, v01 = 0.f, v11 = 0.f;            // v(kx, ky)
                #pragma unroll
                for (int j = 0; j < m; j++){
                    v00 = ADD(v00, FMA(a0[j], ky_0[j], -MUL(kx_0[j], h0[j]))); v10 = ADD(v10, FMA(a1[j], ky_0[j], -MUL(kx_1[j], h0[j])));
                    v01 = ADD(v01, FMA(a0[j], ky_1[j], -MUL(kx_0[j], h1[j]))); v11 = ADD(v11, FMA(a1[j], ky_1[j], -MUL(kx_1[j], h1[j])));
                }
                const float p00 = ADD(x00, v00), p10 = ADD(x10, v10), p01 = ADD(x01, v01), p11 = ADD(x11, v11);
                *reinterpret_cast<float2*>(&s.P[ky0*n + kx0]) = make_float2(p00, p10);
                *reinterpret_cast<float2*>(&s.P[(ky0+1)*n + kx0]) = make_float2(p01, p11);
                // P + rho I: the diagonal entries of this tile are (kx0, ky0) and (kx0+1, ky0+1) when kx0 == ky0
                const bool dg = (kx0 == ky0);
                *reinterpret_cast<float2*>(&s.Pr[ky0*n + kx0]) = make_float2(dg ? ADD(p00, rho) : p00, p10);
                *reinterpret_cast<float2*>(&s.Pr[(ky0+1)*n + kx0]) = make_float2(p01, dg ? ADD(p11, rho) : p11);
            } else if (t >= 228 && t < 228 + n){
                const int r = t - 228;
                float a[8], k2[8];
                ld8(a, s.T + r*BP_RS); ld8(k2, s.K + r*BP_RS);
                float val = 0.f;
                #pragma unroll
                for (int j = 0; j < m; j++){ val = ADD(val, FMA(s.du[j], a[j], -MUL(k2[j], s.g[n + j]))); }
                const float v = ADD(s.g[r], val);
                s.p[r] = v; s.p0[r] = v;
            }
        }
        BP_TRACE(8);
        bar_sync(1, BP_THREADS);      // P,p of the next knot are complete
        BP_TRACE(9);
        BP_TRACE_FLUSH();
        if (t == 0 && i > 0){ mbar_arrive(&s.empty[(i - 1) % BP_STAGES]); }      // the previous knot's ring slot (its deferred stores are done) goes back to the producer
    }
    flush_outputs(bN + ks + 1, s.AB[(nknots - 1) % BP_STAGES], block != 0);      // the last knot of the block (block 0 ends without a cost-to-go)
    // ---- expected cost reduction of this block: thread 0 sums the m per-thread partials in order (bpHelpers.cuh:416)
    if (t >= 196 + n && t < 196 + n + m){ s.dJ[t - (196 + n)] = dJ0; s.dJ[m + t - (196 + n)] = dJ1; }
    bar_sync(1, BP_THREADS);
    if (t == 0){
        float a0 = s.dJ[0], a1 = s.dJ[m];
        for (int j = 1; j < m; j++){ a0 = ADD(a0, s.dJ[j]); a1 = ADD(a1, s.dJ[m+j]); }
        S.dJexp[(size_t)b*2*S.M + 2*block] = a0; S.dJexp[(size_t)b*2*S.M + 2*block + 1] = a1;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// forward simulation + per-knot cost + defects
// grid = B*(A/2)*M CTAs of one warp: the warp simulates shooting interval w of TWO candidates (b, a) and (b, a+1), one per
// half-warp (SIM_LANES = 16 lanes cooperate on one trajectory; both halves run the same instruction stream).
// ------------------------------------------------------------------------------------------------------------------
constexpr int SIM_LANES = 16;
struct SimGroupData {
    kuka::FwdWsT<false> ws;
    __align__(16) float dx[16];
    float x[16], u[8], qdd[8], xn[16];
    float ee[8];               // tool pose of the current knot (EE_COST)
};
// The two groups of a warp run the same instruction on their own workspaces: the workspaces are padded to an odd multiple
// of 16 banks so that a 16-lane unit-stride access of one group never meets the other group's banks.
constexpr int SIM_GROUP_FLOATS = (int)(sizeof(SimGroupData) / 4), SIM_PAD = (48 - SIM_GROUP_FLOATS % 32) % 32;
struct SimGroupSmem : SimGroupData { float pad[SIM_PAD ? SIM_PAD : 32]; };
static_assert((sizeof(SimGroupSmem) / 4) % 32 == 16, "group workspaces must sit 16 banks apart");

// USE_LIMITS_FLAG (plants/cost_arm.cuh:11-94): qr * quadPen<dLevel>(val, limit) of entry `ind` of [x; u] -- zero inside 0.8 x the
// iiwa14's joint / velocity / torque limit, 0.5 delta^2 (dLevel 0) or +-delta (dLevel 1) beyond it, delta = |val| - limit.  The
// reference forms 0.5*delta*delta in double: both products are exact there, so its float result is MUL(MUL(0.5f, delta), delta).
// Rounding of the reference's device build (SASS of its costKern / costGradientHessianKern built with USE_LIMITS_FLAG 1): every penalty is
// rounded on its own and selected against 0, so nothing is contracted into the sums -- except that `cost = 0.5*cost; cost += term_0` is one
// fma(cost, 0.5, term_0); the other terms, and the gradient's `gk[i] += term`, are plain additions.
__device__ __forceinline__ float limit_of(int ind){
    constexpr int np = kuka::NB, n = kuka::NX;
    if (ind < np){ return (float)(ind == 6 ? 3.05432619099 * 0.8 : ((ind & 1) ? 2.09439510239 * 0.8 : 2.96705972839 * 0.8)); }
    if (ind < n){ const int j = ind - np; return (float)(j > 4 ? 2.356194 * 0.8 : (j == 4 ? 2.268928 * 0.8 : (j == 3 ? 1.308996 * 0.8 : (j == 2 ? 1.745329 * 0.8 : 1.483529 * 0.8)))); }
    return (float)(300.0 * 0.8);
}
template <int DLEVEL>
__device__ __forceinline__ float limit_pen(float val, int ind){
    const float delta = SUB(fabsf(val), limit_of(ind));
    return delta < 0.f ? 0.f : (DLEVEL == 0 ? MUL(MUL(0.5f, delta), delta) : (DLEVEL == 1 ? (val < 0.f ? -delta : delta) : 1.f));
}
__device__ __forceinline__ float limit_weight(int ind, const DevState &S){ return ind < kuka::NB ? S.Q_PL : (ind < kuka::NX ? S.Q_VL : S.R_TL); }
// joint-space quadratic cost of one knot (plants/cost_arm.cuh:128-153), evaluated by one lane
__device__ __forceinline__ float cost_knot(const float *x, const float *u, const float *xg, bool final_knot, const DevState &S){
    float cost = 0.f;
    if (final_knot){
        for (int i = 0; i < kuka::NX; i++){ float dl = SUB(x[i], xg[i]); cost = FMA(MUL(i < kuka::NB ? S.QF1 : S.QF2, dl), dl, cost); }
    } else {
        for (int i = 0; i < kuka::NX; i++){ float dl = SUB(x[i], xg[i]); cost = FMA(MUL(i < kuka::NB ? S.Q1 : S.Q2, dl), dl, cost); }
        for (int i = 0; i < kuka::NU; i++){ cost = FMA(MUL(S.R, u[i]), u[i], cost); }
    }
    if (S.use_limits){
        cost = FMA(cost, 0.5f, MUL(limit_weight(0, S), limit_pen<0>(x[0], 0)));
        for (int i = 1; i < kuka::NX; i++){ cost = ADD(cost, MUL(limit_weight(i, S), limit_pen<0>(x[i], i))); }
        if (!final_knot){ for (int i = 0; i < kuka::NU; i++){ cost = ADD(cost, MUL(limit_weight(kuka::NX + i, S), limit_pen<0>(u[i], kuka::NX + i))); } }
        return cost;
    }
    return MUL(0.5f, cost);
}

// end-effector cost terms (plants/cost_arm.cuh; USE_EE_VEL_COST, USE_LIMITS_FLAG, USE_SMOOTH_ABS 0, no xTarget, timeShift 0)
// eeCost :207-223
__device__ __forceinline__ float ee_pose_cost(const float *ee, const float *goal, bool fin, const DevState &S){
    float cost = 0.f;
    #pragma unroll
    for (int i = 0; i < 6; i++){
        const float dl = SUB(ee[i], goal[i]), Q = fin ? (i < 3 ? S.QF_EE1 : S.QF_EE2) : (i < 3 ? S.Q_EE1 : S.Q_EE2);
        cost = FMA(MUL(MUL(0.5f, Q), dl), dl, cost);
    }
    if (S.smooth_abs){ cost = SUB(sqrtf(FMA(2.f, cost, S.sa_alpha2)), S.sa_alpha); }      // USE_SMOOTH_ABS, cost_arm.cuh:218-220
    return cost;
}
// `cost += nominalStateCost(...)` :263-270
__device__ __forceinline__ float ee_add_nominal(const float *x, const float *xt, int ind, bool fin, float cost, const DevState &S){
    const float Qq = fin ? S.QF_xEE : S.Q_xEE, Qqd = fin ? S.QF_xdEE : S.Q_xdEE;
    const float dq = xt ? SUB(x[ind], xt[ind]) : x[ind], dqd = xt ? SUB(x[ind + kuka::NB], xt[ind + kuka::NB]) : x[ind + kuka::NB];
    return FMA(0.5f, FMA(MUL(Qq, dq), dq, MUL(MUL(Qqd, dqd), dqd)), cost);
}
// USE_LIMITS_FLAG inside the end-effector cost (:289-291,310-312): joint `ind` adds the penalties of its angle, its velocity and its torque
// (the torque also on the final knot, unlike the joint-space cost)
__device__ __forceinline__ float ee_add_limits(const float *x, const float *u, int ind, float cost, const DevState &S){
    cost = ADD(cost, MUL(S.Q_PL, limit_pen<0>(x[ind], ind)));
    cost = ADD(cost, MUL(S.Q_VL, limit_pen<0>(x[ind + kuka::NB], ind + kuka::NB)));
    return ADD(cost, MUL(S.R_TL, limit_pen<0>(u[ind], ind + kuka::NX)));
}
// joint `ind`'s share of one knot (split costFunc :283-303)
__device__ __forceinline__ float ee_cost_share(int ind, const float *ee, const float *goal, const float *x, const float *xt, const float *u, bool fin, bool fin_pose, const DevState &S){
    float cost = 0.f;
    if (ind == 0){ cost = ADD(cost, ee_pose_cost(ee, goal, fin_pose, S)); }
    const float Rk = fin ? 0.f : S.R_EE;
    cost = FMA(MUL(MUL(0.5f, Rk), u[ind]), u[ind], cost);
    cost = ee_add_nominal(x, xt, ind, fin, cost, S);
    if (S.use_limits){ cost = ee_add_limits(x, u, ind, cost, S); }
    return cost;
}

// LANES = 16: two candidates per warp, the throughput shape (27 % faster than 32 lanes once every scheduler has several warps).
// LANES = 32: one candidate per warp, the latency shape for small batches (fewer passes per phase: 13 % faster per knot when the
// whole launch fits the machine with one warp per scheduler).  Same arithmetic either way; the host picks (launch_sim_any).
template <bool EE, int LANES>
__global__ void __launch_bounds__(32, 14) sim_kernel(DevState S, int b0, int n_cand, int a_first){
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int n = kuka::NX, m = kuka::NU, GPW = 32 / LANES;
    float *sTb = reinterpret_cast<float*>(smem_raw);           // 252
    float *sxg = sTb + 36*kuka::NB;                            // 16
    SimGroupSmem *gsm = reinterpret_cast<SimGroupSmem*>(sxg + 16);
    // one warp per CTA: the launch is a single wave bound by the shared-memory pipe of the fullest SM, and 2048 one-warp CTAs
    // spread over 148 SMs as 13 or 14 warps each where 512 four-warp CTAs left 12 or 16
    const int apb = (n_cand + GPW - 1) / GPW;                  // candidate groups per problem (n_cand = A, or 1 for the initial rollout)
    const int w = blockIdx.x % S.M, pg = blockIdx.x / S.M;     // shooting interval; (problem, candidate group)
    const int b = b0 + pg / apb;
    if (S.done[b]){ return; }
    for (int i = threadIdx.x; i < 36*kuka::NB; i += 32){ sTb[i] = S.Tbody[i]; }
    if (threadIdx.x < n){ sxg[threadIdx.x] = S.xGoal[b*n + threadIdx.x]; }
    __syncwarp();
    const int grp = threadIdx.x / LANES, l = threadIdx.x & (LANES-1);
    float Ib[36]; kuka::load_body_inertia<LANES>(S.I, Ib);     // lane 2 j + h: inertia of body j, in registers for the whole kernel
    int a = (pg % apb)*GPW + grp;
    const bool live = a < n_cand;                              // odd count: the last half-warp replays the last candidate without storing
    if (!live){ a = n_cand - 1; }
    a += a_first;                                              // candidates a_first .. a_first + n_cand - 1 (step-size sharding)
    SimGroupSmem &s = gsm[grp];
    kuka::init_ws<LANES>(s.ws, nullptr, sTb, S.grav);
    const kuka::FwdIdx<LANES> fix = kuka::make_fwd_idx<LANES>();
    // EE: every interval runs NBF steps -- the last one also evaluates knot N-1 for its pose cost (fpHelpers.cuh:235)
    const int N = S.N, NBF = N / S.M, kStart = w*NBF, iters = (EE || w < S.M - 1) ? NBF : NBF - 1;
    float cacc = 0.f;                                          // EE: lane l < 7 carries s_cost[l]
    const float alpha = S.alpha[a], dt = S.dt;
    float *gx = S.x + ((size_t)b*S.A + a)*N*n, *gu = S.u + ((size_t)b*S.A + a)*N*m, *gdd = S.d + ((size_t)b*S.A + a)*N*n;
    float *gc = S.costk + ((size_t)b*S.A + a)*N;
    const float *gxp = S.xp + (size_t)b*N*n, *gup = S.up + (size_t)b*N*m, *gKT = S.KT + (size_t)b*N*n*m, *gdu = S.du + (size_t)b*N*m;
    // state at the start of the interval: left there by the sweep; the first interval always starts at xp[0] (with a single
    // interval there is no sweep at all, fpHelpers.cuh:17-63 is skipped for M_BLOCKS_F = 1)
    if (l < n){
        if (w == 0){ const float v = gxp[l]; s.x[l] = v; if (live){ gx[l] = v; } }
        else { s.x[l] = gx[kStart*n + l]; }
    }
    // prefetch registers for knot kStart: lane l < m holds row l of the gain K (14 values) -- the row it multiplies with x - xp
    const int lm = l < m ? l : m - 1;
    float kt[n], rdu = 0.f, rxp = 0.f, rup = 0.f;
    auto load_gain = [&](int kq){
        const float2 *src = reinterpret_cast<const float2*>(gKT + (size_t)kq*n*m + lm*n);
        #pragma unroll
        for (int c = 0; c < n; c += 2){ const float2 v = src[c >> 1]; kt[c] = v.x; kt[c+1] = v.y; }
    };
    load_gain(kStart);
    rdu = gdu[kStart*m + lm]; rup = gup[kStart*m + lm];
    rxp = gxp[kStart*n + (l < n ? l : n - 1)];
    __syncwarp();
    #pragma unroll 1
    for (int kk = 0; kk < iters; kk++){
        const int k = kStart + kk;
        if (l < n){ s.dx[l] = SUB(s.x[l], rxp); }
        const float du_k = rdu, up_k = rup;
        // unconditional, index-clamped loads straight into the loop-carried registers: a predicated load would be
        // followed by a register move that waits for it, which exposes the global-memory latency in every step
        const int kn = (kk + 1 < iters) ? k + 1 : k;
        rdu = gdu[kn*m + lm]; rup = gup[kn*m + lm];
        rxp = gxp[kn*n + (l < n ? l : n - 1)];
        __syncwarp();
        // u = up - (alpha du + K dx)          (fpHelpers.cuh:210-219)
        {
            float dxv[16];
            #pragma unroll
            for (int c = 0; c < 16; c += 4){ const float4 v = *reinterpret_cast<const float4*>(&s.dx[c]); dxv[c] = v.x; dxv[c+1] = v.y; dxv[c+2] = v.z; dxv[c+3] = v.w; }
            float Kdx = 0.f;
            #pragma unroll
            for (int c = 0; c < n; c++){ Kdx = FMA(kt[c], dxv[c], Kdx); }
            const float uu = SUB(up_k, FMA(alpha, du_k, Kdx));
            if (l < m){ s.u[l] = uu; if (live){ gu[k*m + l] = uu; } }
        }
        load_gain(kn);                                          // the next knot's row, a whole dynamics evaluation ahead of its use
        __syncwarp();
        kuka::forward_sim<LANES>(s.ws, Ib, s.x, s.u, s.qdd, fix, EE ? s.ee : nullptr);
        if (EE){
            // running / final cost of this knot, not on the knots that close a defect (fpHelpers.cuh:259-265)
            if (l < kuka::NB && (kk < NBF - 1 || w == S.M - 1)){ cacc = ADD(cacc, ee_cost_share(l, s.ee, sxg, s.x, S.xTarget ? S.xTarget + (size_t)b*n : nullptr, s.u, k == N - 1, k >= N - 1 - (S.cost_shift ? S.cost_shift[b] : 0), S)); }
        }
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
    if (EE){
        // cost partial of this (interval, candidate): s_cost[0] + ... + s_cost[6] in order (fpHelpers.cuh:299)
        float J = __shfl_sync(FULL, cacc, 0, LANES);
        #pragma unroll
        for (int i = 1; i < kuka::NB; i++){ J = ADD(J, __shfl_sync(FULL, cacc, i, LANES)); }
        if (l == 0 && live){ gc[w] = J; }
        return;
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
    // mode 2 / 3 (step-size sharding): 2 = this rank's candidates only, their (J, defect) pairs go to the exchange buffer and the kernel
    // ends; 3 = the pairs of all candidates come back from the all-gather and the selection below runs, identically on every rank
    const int wi = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int a = (mode == 2) ? S.a_first + wi : wi;
    float *sJ = ssel + (size_t)A*N, *sdT = sJ + A;
    const int nA = (mode == 1) ? 1 : (mode == 2 ? S.a_first + S.a_cnt : A);
    if (mode == 3){
        if (threadIdx.x < A){
            const int r = threadIdx.x / S.a_cnt, i = threadIdx.x % S.a_cnt;
            const float *pr = S.xchg_recv + (((size_t)r*S.B + b)*S.a_cnt + i)*2;
            sJ[threadIdx.x] = pr[0]; sdT[threadIdx.x] = pr[1];
        }
    } else if (a < nA){
        float *v = ssel + (size_t)a*N; const float *gc = S.costk + ((size_t)b*A + a)*N;
        if (S.ee && (mode == 0 || S.rolled_out)){
            // costKern<T,0> (fpHelpers.cuh:165-172): the simulation's per-interval partials, summed in interval order
            if (l == 0){ float J = 0.f; for (int i = 0; i < S.M; i++){ J = ADD(J, gc[i]); } sJ[a] = J; }
        } else {
            for (int i = l; i < N; i += 32){ v[i] = ADD(0.f, gc[i]); }
            __syncwarp();
            for (int st = N/2; st >= 2; st >>= 1){
                for (int i = l; i < st; i += 32){ v[i] = ADD(v[i], v[i+st]); }
                __syncwarp();
            }
            if (l == 0){ sJ[a] = ADD(v[0], v[1]); }
        }
        // defect: max over the M-1 interval boundaries of the L1 norm (fpHelpers.cuh:94-111)
        float dmax = 0.f;
        if ((mode == 0 || mode == 2) && l < S.M - 1){
            const int NBF = N / S.M; const float *dk = S.d + (((size_t)b*A + a)*N + (l+1)*NBF - 1)*n;
            float acc = 0.f; for (int c = 0; c < n; c++){ acc = ADD(acc, fabsf(dk[c])); }
            dmax = acc;
        }
        for (int o = 16; o >= 1; o >>= 1){ dmax = fmaxf(dmax, __shfl_xor_sync(FULL, dmax, o)); }
        if (l == 0){ sdT[a] = dmax; }
    }
    __syncthreads();
    if (mode == 2){
        if (l == 0 && wi < S.a_cnt){ float *ps = S.xchg_send + ((size_t)b*S.a_cnt + wi)*2; ps[0] = sJ[a]; ps[1] = sdT[a]; }
        return;
    }
    if (threadIdx.x != 0){ return; }
    float *Jout = S.Jout + (size_t)b*(S.max_iter+1); int *alphaOut = S.alphaOut + (size_t)b*(S.max_iter+1);
    if (mode == 1){
        // Reference behaviour under EE_COST: costGradientHessianKern leaves the per-knot costs in d_JT[0..N-1], costKern<T,1> puts
        // their sum into d_JT[0] only, and initAlgGPU reads d_JT[*alphaIndex] (nisInitHelpers.cuh:388-391).  runiLQR_GPU starts with
        // alphaIndex = 0; the receding-horizon wrapper keeps the previous solve's index, so when that is not 0 its "initial cost" is
        // the cost of knot alphaIndex alone.  Reproduced as is (init_knot = that index, set by mpc_load_kernel).
        const int ik = (S.ee && !S.rolled_out) ? S.init_knot[b] : 0;
        if (ik != 0){ sJ[0] = S.costk[((size_t)b*A + 0)*N + ik]; }
        float pj = ADD(sJ[0], S.two_tol);                  // nisInitHelpers.cuh:393
        S.prevJ[b] = pj; Jout[0] = SUB(pj, S.two_tol); alphaOut[0] = S.rolled_out ? 0 : -1;     // nisInitHelpers.cuh:363
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
    if (!done){ if (iter == S.iter_cap){ done = 1; } else { iter += 1; } }
    S.dJ[b] = dJ; S.rho[b] = rho; S.drho[b] = drho; S.alphaIndex[b] = alphaIndex; S.accepted[b] = accepted; S.iter[b] = iter;
    if (done){ S.done[b] = 1; S.final_src[b] = accepted ? alphaIndex : -1; atomicSub(S.n_active, 1); }
}

// ------------------------------------------------------------------------------------------------------------------
// next-iteration setup: one warp per (problem, knot).  Hands the accepted candidate over to (xp,up,dp), keeps the
// previous one in xp2, and refreshes AB (analytic Euler gradient), g (and H when write_H) at the accepted trajectory.
// mode 1 = initialisation: the trajectory is already in xp/up, xp2 <- xp.  mode 2 = initialisation after the forward
// rollout of loadVarsGPU: candidate 0 is handed over first and becomes xp, xp2, up, dp.
// ------------------------------------------------------------------------------------------------------------------
constexpr int NIS_LANES = 32;
struct NisGroupSmem {
    kuka::FwdWs ws; kuka::GradWs gs;
    float x[16], u[8], qdd[8], dqdd[3*kuka::NB*kuka::NB + 1];
    float ee[8], dee[6*kuka::NB + 2];      // tool pose and its derivative (EE_COST)
};
#ifndef PDDP_NIS_WARPS
#define PDDP_NIS_WARPS 1
#endif
#ifndef PDDP_NIS_STAGE
#define PDDP_NIS_STAGE 0      // 1: body inertias / joint frames staged in shared memory per CTA, 0: read through L1 (measured faster: one more CTA per SM)
#endif
constexpr int NIS_WARPS = PDDP_NIS_WARPS;
constexpr int NIS_CONST_FLOATS = PDDP_NIS_STAGE ? 2*36*kuka::NB : 0;

__global__ void __launch_bounds__(32*NIS_WARPS) nis_kernel(DevState S, int mode, int write_H, int b0, int nb){
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int n = kuka::NX, m = kuka::NU, nm = n + m, np = kuka::NB, LANES = NIS_LANES, GPW = 32 / NIS_LANES;
    NisGroupSmem *gsm = reinterpret_cast<NisGroupSmem*>(reinterpret_cast<float*>(smem_raw) + NIS_CONST_FLOATS);
    const int w = threadIdx.x >> 5, grp = (threadIdx.x & 31) / LANES, l = threadIdx.x & (LANES-1);
    const int gk = (blockIdx.x*NIS_WARPS + w)*GPW + grp, N = S.N;        // N is even: the groups of one warp share the problem
    const int b = b0 + gk / N, k = gk % N;
#if PDDP_NIS_STAGE
    float *sI = reinterpret_cast<float*>(smem_raw); float *sTb = sI + 36*kuka::NB;
    for (int i = threadIdx.x; i < 36*kuka::NB; i += blockDim.x){ sI[i] = S.I[i]; sTb[i] = S.Tbody[i]; }
    __syncthreads();
#else
    const float *sI = S.I, *sTb = S.Tbody;
#endif
    if (b >= b0 + nb){ return; }
    // Everything this knot reads from global memory is requested before the first wait: the per-problem scalars (read one after the
    // other behind their branches they were two dependent trips to L2), the constant joint transforms, the current trajectory; the
    // accepted candidate's state follows as soon as its index is here and travels while the workspace is initialised.  (These waits
    // were 8 % of the kernel's stall samples.)
    NisGroupSmem &s = gsm[w*GPW + grp];
    float *gxp = S.xp + ((size_t)b*N + k)*n, *gup = S.up + ((size_t)b*N + k)*m, *gdp = S.dp + ((size_t)b*N + k)*n, *gxp2 = S.xp2 + ((size_t)b*N + k)*n;
    const int done_b = S.done[b], accepted_b = S.accepted[b], a = S.alphaIndex[b];
    constexpr int TBQ = (16*kuka::NB + LANES - 1)/LANES;
    float tbr[TBQ];
    #pragma unroll
    for (int q = 0; q < TBQ; q++){ const int e = l + q*LANES, ec = e < 16*kuka::NB ? e : 16*kuka::NB - 1; tbr[q] = sTb[36*(ec >> 4) + (ec & 15)]; }
    const float xold = gxp[l < n ? l : n-1], uold = gup[l < m ? l : m-1];
    if (done_b){ return; }
    const bool acc = (mode == 2) || ((mode == 0) && accepted_b);         // mode 2: initialisation after a forward rollout
    const float *cx = S.x + (((size_t)b*S.A + a)*N + k)*n, *cu = S.u + (((size_t)b*S.A + a)*N + k)*m, *cd = S.d + (((size_t)b*S.A + a)*N + k)*n;
    const bool on_boundary = ((k+1) % (N / S.M)) == 0 && k < N-1;
    const float xcand = cx[l < n ? l : n-1], ucand = cu[l < m ? l : m-1], dcand = cd[l < n ? l : n-1];
    // kuka::init_ws with the joint transforms from the registers above
    if (l == 0){ s.ws.grav = S.grav; }
    #pragma unroll
    for (int q = 0; q < TBQ; q++){ const int e = l + q*LANES; if (e < 16*kuka::NB){ s.ws.Tb[e] = tbr[q]; s.gs.dTb[e] = 0.f; } }
    if (l < n){
        const float xv = acc ? xcand : xold; s.x[l] = xv;
        gxp2[l] = (mode == 2) ? xv : xold;                         // xp2 <- xp (fpHelpers.cuh:371); initAlgGPU sets both to the start trajectory (nisInitHelpers.cuh:378-379)
        // defects: a candidate's array is the broadcast copy of dp (memcpyCurrAKern, nisInitHelpers.cuh:22-32) with the interval-
        // boundary entries rewritten by the simulation, so only those change hands; the others are never read by a solve but the
        // receding-horizon shift moves them onto boundaries later
        if (acc){ gxp[l] = xv; if (on_boundary){ gdp[l] = dcand; } }
    }
    if (l < m){ const float uv = acc ? ucand : uold; s.u[l] = uv; if (acc){ gup[l] = uv; } }
    // opt-in (pddp_set_skip_unchanged): after a rejected line search the trajectory, hence AB, H and g, are unchanged; the
    // reference recomputes them all the same (nisInitHelpers.cuh:245-279) and so does the default here
    if (S.skip_unchanged && mode == 0 && !acc){ return; }
    __syncwarp();
    // cost gradient (plants/cost_arm.cuh:156-202)
    const float *xg = S.xGoal + b*n;
    float *gg = S.g + ((size_t)b*N + k)*G_STRIDE;
    const bool fin = (k == N - 1);
    // joint-space cost; the end-effector cost needs the tool pose and follows the gradient below
    if (!S.ee){
        for (int e = l; e < nm; e += LANES){
            float gv;
            if (e < n){ gv = MUL(fin ? (e < np ? S.QF1 : S.QF2) : (e < np ? S.Q1 : S.Q2), SUB(s.x[e], xg[e])); }
            else { gv = fin ? 0.f : MUL(S.R, s.u[e-n]); }
            if (S.use_limits && (e < n || !fin)){ gv = ADD(gv, MUL(limit_weight(e, S), limit_pen<1>(e < n ? s.x[e] : s.u[e-n], e))); }      // cost_arm.cuh:176-179,197-200
            gg[e] = gv;
        }
    }
    if (write_H && !S.ee){
        float *gH = S.H + ((size_t)b*N + k)*H_STRIDE;
        for (int e = l; e < nm*nm; e += LANES){
            const int i = e / nm, j = e % nm; float v = 0.f;
            if (fin){ if (i < n && j < n){ v = (i != j) ? 0.f : (i < np ? S.QF1 : S.QF2); } }
            else { v = (i != j) ? 0.f : (i < np ? S.Q1 : (i < n ? S.Q2 : S.R)); }
            gH[e] = v;
        }
    }
    // integrator gradient AB = [I 0] + dt [0 I 0 ; dqdd]   (integrators.cuh:15-17,38-53); the final knot has none
    // (its group still runs the collective code so that both halves of a warp stay in lockstep, but stores nothing)
    kuka::gradient<LANES>(s.ws, s.gs, sI, s.x, s.u, s.qdd, s.dqdd, S.ee ? s.ee : nullptr, S.ee ? s.dee : nullptr);
    if (S.ee){
        // costGrad with the pose terms (plants/cost_arm.cuh:328-388): g, and the whole Hessian at every knot -- Gauss-Newton on the
        // pose (unweighted, as the reference has it) plus the diagonal weights
        const float Rk = fin ? 0.f : S.R_EE;
        const float *xt = S.xTarget ? S.xTarget + (size_t)b*n : nullptr;
        const bool finp = k >= N - 1 - (S.cost_shift ? S.cost_shift[b] : 0);       // pose terms: final weights (finalCostShift)
        for (int r = l; r < nm; r += LANES){
            float val = 0.f;
            if (r < np){
                float v2 = 0.f;
                #pragma unroll
                for (int i = 0; i < 6; i++){
                    const float dl = SUB(s.ee[i], xg[i]), Q = finp ? (i < 3 ? S.QF_EE1 : S.QF_EE2) : (i < 3 ? S.Q_EE1 : S.Q_EE2);
                    v2 = FMA(MUL(Q, dl), s.dee[r*6+i], v2);
                }
                if (S.smooth_abs){                                   // USE_SMOOTH_ABS, cost_arm.cuh:242-252: d/dq of sqrt(2 c + alpha^2)
                    float v3 = 0.f;
                    #pragma unroll
                    for (int i = 0; i < 6; i++){
                        const float dl = SUB(s.ee[i], xg[i]), Q = finp ? (i < 3 ? S.QF_EE1 : S.QF_EE2) : (i < 3 ? S.Q_EE1 : S.Q_EE2);
                        v3 = FMA(MUL(Q, dl), dl, v3);
                    }
                    v2 = DIV(v2, sqrtf(ADD(v3, S.sa_alpha2)));
                }
                val = ADD(val, v2);
            }
            if (r < n){ val = ADD(val, MUL((r < np) ? (fin ? S.QF_xEE : S.Q_xEE) : (fin ? S.QF_xdEE : S.Q_xdEE), xt ? SUB(s.x[r], xt[r]) : s.x[r])); }   // the reference's build leaves this product unfused (pinned by its GPU unit dump)
            else { val = FMA(Rk, s.u[r-n], val); }
            if (S.use_limits){ val = ADD(val, MUL(limit_weight(r, S), limit_pen<1>(r < n ? s.x[r] : s.u[r-n], r))); }      // cost_arm.cuh:341-343
            gg[r] = val;
        }
        float *gH = S.H + ((size_t)b*N + k)*H_STRIDE;
        // The reference rewrites all (n+m)^2 entries at every knot of every iteration; only the 7 x 7 pose block and the diagonal ever change
        // (the other entries are written as +0 by the initialisation, write_H = 1), so an iteration rewrites those 63 entries
        const int cnt = write_H ? nm*nm : np*np + (nm - np);
        for (int i = l; i < cnt; i += LANES){
            const int e = write_H ? i : (i < np*np ? (i / np)*nm + (i % np) : (i - np*np + np)*(nm + 1));
            const int cc = e / nm, r = e % nm; float val = 0.f;
            if (r < np && cc < np){
                #pragma unroll
                for (int j = 0; j < 6; j++){ val = FMA(s.dee[r*6+j], s.dee[cc*6+j], val); }
            }
            if (r == cc){
                val = ADD(val, (r < np) ? (fin ? S.QF_xEE : S.Q_xEE) : (r < n ? (fin ? S.QF_xdEE : S.Q_xdEE) : Rk));
                if (S.use_limits){ val = ADD(val, MUL(limit_weight(r, S), limit_pen<2>(r < n ? s.x[r] : s.u[r-n], r))); }      // cost_arm.cuh:374-376
            }
            gH[e] = val;
        }
        // initialisation without a rollout: the knot's cost (costGrad's d_JT, :383-388; single-valued costFunc :306-325)
        if (mode == 1 && l == 0){
            float cost = 0.f;
            for (int ind = 0; ind < np; ind++){
                if (ind == 0){ cost = ADD(cost, ee_pose_cost(s.ee, xg, finp, S)); }
                cost = FMA(MUL(MUL(0.5f, Rk), s.u[ind]), s.u[ind], cost);
                cost = ee_add_nominal(s.x, xt, ind, fin, cost, S);
                if (S.use_limits){ cost = ee_add_limits(s.x, s.u, ind, cost, S); }
            }
            S.costk[((size_t)b*S.A + 0)*N + k] = cost;
        }
    }
    if (fin){ return; }
    float *gAB = S.AB + ((size_t)b*N + k)*AB_STRIDE;
    const float dt = S.dt;
    // entry (row kx, column ky): 1[ky == kx] + dt * (kx < np ? 1[ky == kx + np] : dqdd[kx - np, ky]); a lane takes the pair of entries
    // (r, ky), (np + r, ky) of dqdd's element j = ky*np + r: half the trips of a loop over all n*nm entries, no index arithmetic per half
    for (int j = l; j < np*nm; j += LANES){
        const int ky = j / np, r = j - ky*np;
        gAB[ky*n + r] = FMA(dt, (r + np == ky) ? 1.f : 0.f, (ky == r) ? 1.f : 0.f);
        gAB[ky*n + np + r] = FMA(dt, s.dqdd[j], (ky == np + r) ? 1.f : 0.f);
    }
}

// step-size sharding: the accepted candidate travels from the rank that simulated it to all others as an all-reduce (integer sum) of
// its bit patterns -- every other rank contributes zeros, so the values arrive bit for bit.  pack: owner -> acc_buf, others zero;
// unpack (after the all-reduce): acc_buf -> the candidate's slot of x, u, d on every rank.
__global__ void ashard_pack_kernel(DevState S){
    const int b = blockIdx.x, N = S.N, n = S.n, m = S.m, a = S.alphaIndex[b], per = N*(2*n + m);
    int *dst = S.acc_buf + (size_t)b*per;
    const bool mine = S.accepted[b] && a >= S.a_first && a < S.a_first + S.a_cnt;
    const int *sx = reinterpret_cast<const int*>(S.x + ((size_t)b*S.A + a)*N*n), *su = reinterpret_cast<const int*>(S.u + ((size_t)b*S.A + a)*N*m),
              *sd = reinterpret_cast<const int*>(S.d + ((size_t)b*S.A + a)*N*n);
    for (int i = threadIdx.x; i < per; i += blockDim.x){
        int v = 0;
        if (mine){ v = i < N*n ? sx[i] : (i < N*(n+m) ? su[i - N*n] : sd[i - N*(n+m)]); }
        dst[i] = v;
    }
}
__global__ void ashard_unpack_kernel(DevState S){
    const int b = blockIdx.x, N = S.N, n = S.n, m = S.m, a = S.alphaIndex[b], per = N*(2*n + m);
    if (!S.accepted[b] || (a >= S.a_first && a < S.a_first + S.a_cnt)){ return; }
    const int *src = S.acc_buf + (size_t)b*per;
    int *dx = reinterpret_cast<int*>(S.x + ((size_t)b*S.A + a)*N*n), *du = reinterpret_cast<int*>(S.u + ((size_t)b*S.A + a)*N*m), *dd = reinterpret_cast<int*>(S.d + ((size_t)b*S.A + a)*N*n);
    for (int i = threadIdx.x; i < per; i += blockDim.x){
        const int v = src[i];
        if (i < N*n){ dx[i] = v; } else if (i < N*(n+m)){ du[i - N*n] = v; } else { dd[i - N*(n+m)] = v; }
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

// rolloutMPC<NUM_TIME_STEPS> (MPCHelpers.cuh:524-556): open loop from the measured state over the whole horizon, by one warp.  Out of
// line on purpose: inlined, the compiler laid the copy code of the other warps into the middle of this loop, whose body then
// spanned 44 KB -- more than the 32 KB instruction cache behind a lone warp -- instead of the 28 KB it has in the simulation kernel.
__device__ __noinline__ void mpc_rollout(const float *I, const float *Tbody, float grav, float dt, unsigned char *smem_raw, const float *xActual, float *cx, const float *cu, int N){
    constexpr int n = kuka::NX, m = kuka::NU, LANES = 32;
    float *sI = reinterpret_cast<float*>(smem_raw); float *sTb = sI + 36*kuka::NB;
    SimGroupSmem &s = *reinterpret_cast<SimGroupSmem*>(sTb + 36*kuka::NB);
    const int l = threadIdx.x;
    for (int i = l; i < 36*kuka::NB; i += 32){ sI[i] = I[i]; sTb[i] = Tbody[i]; }
    __syncwarp();
    kuka::init_ws<LANES>(s.ws, nullptr, sTb, grav);
    const kuka::FwdIdx<LANES> fix = kuka::make_fwd_idx<LANES>();
    float Ib[36]; kuka::load_body_inertia<LANES>(I, Ib);
    if (l < n){ const float v = xActual[l]; s.x[l] = v; cx[l] = v; }
    for (int k = 0; k < N-1; k++){
        if (l < m){ s.u[l] = cu[k*m + l]; }
        __syncwarp();
        kuka::forward_sim<LANES>(s.ws, Ib, s.x, s.u, s.qdd, fix);
        if (l < kuka::NB){ s.xn[l] = FMA(dt, s.x[l+kuka::NB], s.x[l]); s.xn[l+kuka::NB] = FMA(dt, s.qdd[l], s.x[l+kuka::NB]); }
        __syncwarp();
        if (l < n){ const float v = s.xn[l]; s.x[l] = v; cx[(k+1)*n + l] = v; }
        __syncwarp();
    }
}
// loadVarsGPU_MPC (MPCHelpers.cuh:602-657) followed by the hand-over initAlgGPU does (xp, up, dp <- current plan)
__global__ void mpc_load_kernel(DevState S, MpcState Q){
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int n = kuka::NX, m = kuka::NU;
    const int b = blockIdx.x, N = S.N, shift = Q.shift[b]; const bool clear = Q.clear[b] != 0;
    if (threadIdx.x == 0){ S.init_knot[b] = S.ee ? S.alphaIndex[b] : 0; }      // the slot index the reference's plan lives in (the reset that follows zeroes alphaIndex)
    float *cx = Q.cx + (size_t)b*N*n, *cu = Q.cu + (size_t)b*N*m, *cd = Q.cd + (size_t)b*N*n;
    float *xp = S.xp + (size_t)b*N*n, *up = S.up + (size_t)b*N*m, *dp = S.dp + (size_t)b*N*n, *KT = S.KT + (size_t)b*N*n*m;
    float *P0 = S.Pbuf[0] + (size_t)b*N*n*n, *P1 = S.Pbuf[1] + (size_t)b*N*n*n, *p0 = S.pbuf[0] + (size_t)b*N*n, *p1 = S.pbuf[1] + (size_t)b*N*n;
    // The plan itself (x, d, u: small) is shifted by the whole CTA; after that warp 0 integrates the open-loop rollout -- 127 dependent
    // dynamics evaluations, the long pole of the step -- while warps 1-7 move the big arrays (gains, both cost-to-go buffers) under it.
    const int t = threadIdx.x, T = blockDim.x;
    if (shift > 0){
        mpc_shift_part(cx, shift, n, N, false, xp, t, T, 0);
        mpc_shift_part(cd, shift, n, N, false, nullptr, t, T, 0);
        if (!clear){ mpc_shift_part(cu, shift, m, N-1, true, up, t, T, 0); }
    }
    if (clear){ mpc_zero(cu, N*m); }
    __syncthreads();
    if (t < 32){
        mpc_rollout(S.I, S.Tbody, S.grav, S.dt, smem_raw, Q.xActual + b*n, cx, cu, N);
    } else {
        const int w = t - 32, W = T - 32;
        // The solve that follows restarts the iteration counter at 1: its first backward pass overwrites Pbuf[1] and seeds its blocks
        // from Pbuf[0].  The reference seeds them from the (shifted) Pp, the OLDER buffer, so the newer one -- Pbuf[iter & 1] of the
        // previous solve, whose parity differs between problems that stopped at different iterations -- has to sit in slot 1.
        if (!clear && (S.iter[b] & 1) == 0){ mpc_swap_part(P0, P1, N*n*n, w, W); mpc_swap_part(p0, p1, N*n, w, W); mpc_bar(1, W); }
        if (shift > 0 && !clear){
            mpc_shift_part(KT, shift, n*m, N-1, true, nullptr, w, W, 1);
            mpc_shift_part(P0, shift, n*n, N, false, nullptr, w, W, 1); mpc_shift_part(P1, shift, n*n, N, false, nullptr, w, W, 1);
            mpc_shift_part(p0, shift, n, N, false, nullptr, w, W, 1); mpc_shift_part(p1, shift, n, N, false, nullptr, w, W, 1);
        }
        if (clear){ mpc_zero_part(KT, N*n*m, w, W); mpc_zero_part(P0, N*n*n, w, W); mpc_zero_part(P1, N*n*n, w, W); mpc_zero_part(p0, N*n, w, W); mpc_zero_part(p1, N*n, w, W); }
        mpc_zero_part(S.du + (size_t)b*N*m, N*m, w, W); mpc_zero_part(S.dT + (size_t)b*S.A, S.A, w, W);
        mpc_bar(1, W);
        mpc_copy_part(Q.x_old + (size_t)b*N*n, xp, N*n, w, W); mpc_copy_part(Q.u_old + (size_t)b*N*m, up, N*m, w, W);
        mpc_copy_part(Q.KT_old + (size_t)b*N*n*m, KT, N*n*m, w, W); mpc_copy_part(dp, cd, N*n, w, W);
    }
    __syncthreads();
    mpc_copy(xp, cx, N*n); mpc_copy(up, cu, N*m);
}
// storeVarsGPU_MPC (:755-776) on the device side.  A solve counts as successful when one of its iterations accepted a step size
// above zero, or when the failure counter was still at zero before the step (`publish_anyway`: it reaches one either way) -- the
// scan of MPCHelpers.cuh:987-991, done here so that the step needs no host round trip in the middle: the final
// trajectory then becomes the current plan, else the shifted previous plan and gains come back.  Everything the host wants back
// is packed behind `pack` -- x | u | KT | alphaOut | Jout | iterations | success -- for a single device-to-host copy.
constexpr int MPC_STORE_SPLIT = 8;      // CTAs per problem: the hand-back of a single arm is a copy of 60 KB, too slow for one CTA's loads in flight
__global__ void mpc_store_kernel(DevState S, MpcState Q, const int *publish_anyway, float *pack){
    const int b = blockIdx.x / MPC_STORE_SPLIT, t0 = (blockIdx.x % MPC_STORE_SPLIT)*blockDim.x + threadIdx.x, TT = MPC_STORE_SPLIT*blockDim.x, N = S.N, n = S.n, m = S.m, B = S.B, L = S.max_iter + 1;
    float *x_out = pack, *u_out = x_out + (size_t)B*N*n, *KT_out = u_out + (size_t)B*N*m;
    int *a_out = reinterpret_cast<int*>(KT_out + (size_t)B*N*n*m); float *J_out = reinterpret_cast<float*>(a_out + (size_t)B*L);
    int *it_out = reinterpret_cast<int*>(J_out + (size_t)B*L), *success = it_out + B;
    __shared__ int succ_s;
    const int its = S.iter[b];
    if (threadIdx.x == 0){
        int s = publish_anyway[b]; for (int i = 1; i <= its; i++){ if (S.alphaOut[(size_t)b*L + i] > 0){ s = 1; } }
        succ_s = s; if (t0 == 0){ success[b] = s; it_out[b] = its; }
    }
    for (int i = t0; i < L; i += TT){ a_out[(size_t)b*L + i] = S.alphaOut[(size_t)b*L + i]; J_out[(size_t)b*L + i] = S.Jout[(size_t)b*L + i]; }
    __syncthreads();
    float *cx = Q.cx + (size_t)b*N*n, *cu = Q.cu + (size_t)b*N*m, *cd = Q.cd + (size_t)b*N*n, *KT = S.KT + (size_t)b*N*n*m;
    const int src = S.final_src[b];
    const float *sx = (src >= 0) ? S.x + ((size_t)b*S.A + src)*N*n : S.xp + (size_t)b*N*n;
    const float *su = (src >= 0) ? S.u + ((size_t)b*S.A + src)*N*m : S.up + (size_t)b*N*m;
    // defects of the current plan: a candidate's array is the broadcast copy of dp (memcpyCurrAKern) with its interval-boundary
    // entries rewritten by the simulation; the other entries are never read by a solve but they shift onto boundaries later
    const float *sdc = S.d + ((size_t)b*S.A + (src >= 0 ? src : 0))*N*n, *sdp = S.dp + (size_t)b*N*n; const int NBF = N / S.M;
    auto sd_at = [&](int i){ const int k = i / n; const bool onb = (((k+1) % NBF) == 0) && (k < N-1); return (src >= 0 && onb) ? sdc[i] : sdp[i]; };
    if (succ_s){
        for (int i = t0; i < N*n; i += TT){ const float v = sx[i]; cx[i] = v; x_out[(size_t)b*N*n + i] = v; cd[i] = sd_at(i); }
        for (int i = t0; i < N*m; i += TT){ const float v = su[i]; cu[i] = v; u_out[(size_t)b*N*m + i] = v; }
        for (int i = t0; i < N*n*m; i += TT){ KT_out[(size_t)b*N*n*m + i] = KT[i]; }
    } else {
        for (int i = t0; i < N*n; i += TT){ cx[i] = Q.x_old[(size_t)b*N*n + i]; cd[i] = sd_at(i); }
        for (int i = t0; i < N*m; i += TT){ cu[i] = Q.u_old[(size_t)b*N*m + i]; }
        for (int i = t0; i < N*n*m; i += TT){ KT[i] = Q.KT_old[(size_t)b*N*n*m + i]; }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// plug-in unit kernels (one group per sample, same group widths as the production kernels)
// ------------------------------------------------------------------------------------------------------------------
__global__ void unit_dynamics_kernel(const float *I, const float *Tbody, float grav, const float *x, const float *u, int nsamp, float *qdd){
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LANES = SIM_LANES, GPW = 32 / SIM_LANES;
    float *sI = reinterpret_cast<float*>(smem_raw); float *sTb = sI + 36*kuka::NB;
    SimGroupSmem *gsm = reinterpret_cast<SimGroupSmem*>(sTb + 36*kuka::NB);
    const int grp = (threadIdx.x & 31) / LANES, l = threadIdx.x & (LANES-1);
    for (int i = threadIdx.x; i < 36*kuka::NB; i += blockDim.x){ sI[i] = I[i]; sTb[i] = Tbody[i]; }
    __syncthreads();
    SimGroupSmem &s = gsm[grp];
    kuka::init_ws<LANES>(s.ws, nullptr, sTb, grav);
    const kuka::FwdIdx<LANES> fix = kuka::make_fwd_idx<LANES>();
    float Ib[36]; kuka::load_body_inertia<LANES>(I, Ib);
    for (int k0 = blockIdx.x*GPW; k0 < nsamp; k0 += gridDim.x*GPW){
        const int k = k0 + grp < nsamp ? k0 + grp : nsamp - 1;         // tail: replay the last sample
        if (l < kuka::NX){ s.x[l] = x[k*kuka::NX + l]; } if (l < kuka::NU){ s.u[l] = u[k*kuka::NU + l]; }
        __syncwarp();
        kuka::forward_sim<LANES>(s.ws, Ib, s.x, s.u, s.qdd, fix);
        if (l < kuka::NB){ qdd[k*kuka::NB + l] = s.qdd[l]; }
        __syncwarp();
    }
}
__global__ void unit_gradient_kernel(const float *I, const float *Tbody, float grav, const float *x, const float *u, int nsamp, float dt, float *AB, float *qdd){
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int n = kuka::NX, nm = kuka::NX + kuka::NU, np = kuka::NB, LANES = NIS_LANES, GPW = 32 / NIS_LANES;
    float *sI = reinterpret_cast<float*>(smem_raw); float *sTb = sI + 36*kuka::NB;
    NisGroupSmem *gsm = reinterpret_cast<NisGroupSmem*>(sTb + 36*kuka::NB);
    const int grp = (threadIdx.x & 31) / LANES, l = threadIdx.x & (LANES-1);
    for (int i = threadIdx.x; i < 36*kuka::NB; i += blockDim.x){ sI[i] = I[i]; sTb[i] = Tbody[i]; }
    __syncthreads();
    NisGroupSmem &s = gsm[grp];
    kuka::init_ws<LANES>(s.ws, &s.gs, sTb, grav);
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
