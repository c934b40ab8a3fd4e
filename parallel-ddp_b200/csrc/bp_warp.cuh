// bp_warp.cuh -- backward pass as barrier-free warp chains (the throughput shape of bpHelpers.cuh:337-420 for the 14-state, 7-control arm).
//
// One WARP owns one (problem, time block) chain of N/M knots; a CTA is BPW_WARPS independent warps that never meet at a barrier.
// Per warp: a two-slot ring of the knot inputs (AB, H, g: 3 KB) filled by TMA bulk copies on the warp's own mbarrier -- the next
// knot is in flight while the current one is computed -- and 5 KB of working matrices, so 20 chains are resident per SM (the
// block-cooperative bp_kernel of kernels.cuh holds 2 CTAs = 2 chains per SM and spends its time in five CTA barriers per knot).
// Every stage is a set of register tiles: a lane loads the few operand vectors its tile shares (8 / 16-byte shared loads; lanes of
// one tile row read the same address, which the hardware broadcasts) and runs the fused chains of its outputs in the reference's
// order j = 0 .. 13 / 0 .. 6.  Stages are separated by __syncwarp() only.  The arithmetic -- every product, sum and fusion --
// is that of bp_kernel, statement by statement: results are bit-identical (tests/test_gpu_parity.py runs both).
//
//   A   AB2 = AB'(P + rho I[u rows])                         56 tiles of 3 x 2      (bpHelpers.cuh:54-66)
//   B   H = (AB2 AB)' + H_cost, g = AB'p + g_cost            49 + 49 tiles, 21       (:83-118)
//       Huu in the register layout of the elimination -> 7 x 7 Gauss-Jordan by shuffles (invHelpers / cudaUtils.h:236-264)
//   C   K = Huu^-1 Hux, du = Huu^-1 gu                                              (:206-220)
//   D   T = K'Huu - Hxu, expected reduction                                          (:223-240, 315-334)
//   E   P, p of the previous knot (49 tiles of 2 x 2)                                (:242-276)
//   F   A - BK, B du, KT, du, P, p -> HBM                                           (:279-312)
#pragma once
#include "dev_state.cuh"

namespace pddp {

constexpr int BPW_WARPS = 4;             // chains per CTA
constexpr int BPW_RS = 12;               // row stride of the 7-vectors (Hux, Huu, Hinv, K, T): two 16-byte loads per row
struct __align__(16) BpWarpSmem {
    float AB[2][AB_STRIDE], Hc[2][H_STRIDE], gc[2][G_STRIDE];      // ring of knot inputs; H is assembled in place in Hc
    float P[196], Pr[196];                                         // P and P + rho on the diagonal (the operand of the u-rows of AB'(.))
    float AB2[296];                                                // AB2[kx*14 + ky]; T[kx*BPW_RS + j] lives here from stage D on
    float Hux[14*BPW_RS], K[14*BPW_RS], Huu[7*BPW_RS], Hinv[7*BPW_RS];
    float g[24], p[16], dx[16], du[8], pad[8];
    unsigned long long full[2];
};
static_assert(sizeof(BpWarpSmem) % 16 == 0, "warp workspaces must keep 16-byte alignment");

__device__ __forceinline__ void bpw_ld14(float (&x)[14], const float *p){
    const float2 *q = reinterpret_cast<const float2*>(p);
    #pragma unroll
    for (int i = 0; i < 7; i++){ const float2 v = q[i]; x[2*i] = v.x; x[2*i+1] = v.y; }
}
__device__ __forceinline__ void bpw_ld8(float (&x)[8], const float *p){
    const float4 *q = reinterpret_cast<const float4*>(p);
    const float4 a = q[0], b = q[1];
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
template <int K, int KP>
__device__ __forceinline__ float bpw_dot(const float (&x)[KP], const float (&y)[KP]){
    float val = 0.f;
    #pragma unroll
    for (int j = 0; j < K; j++){ val = FMA(x[j], y[j], val); }
    return val;
}

#ifndef PDDP_BPW_MINCTA
#define PDDP_BPW_MINCTA 5
#endif
#ifndef PDDP_BPW_UNROLL
#define PDDP_BPW_UNROLL 1
#endif
#define BPW_PRAGMA2(x) _Pragma(#x)
#define BPW_PRAGMA(x) BPW_PRAGMA2(x)
#define BPW_UNROLL BPW_PRAGMA(unroll PDDP_BPW_UNROLL)
__global__ void __launch_bounds__(32*BPW_WARPS, PDDP_BPW_MINCTA) bp_warp_kernel(DevState S, int b0, int nchains){
    constexpr int n = 14, m = 7, nm = 21, oB = n*n;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const int chain = blockIdx.x*BPW_WARPS + w;
    if (chain >= nchains){ return; }
    BpWarpSmem &s = reinterpret_cast<BpWarpSmem*>(smem_raw)[w];
    const int b = b0 + chain / S.M, block = chain % S.M;
    if (S.done[b]){ return; }
    const int cur = S.iter[b] & 1;
    const int N = S.N, NBB = N / S.M;
    const float rho = S.rho[b];
    const size_t bN = (size_t)b*N;
    int ks = NBB*(block+1) - 1, iterCount;
    const bool last_block = (ks == N - 1);
    if (last_block){ ks--; iterCount = NBB - 2; } else { iterCount = NBB - 1; }
    const int nknots = iterCount + 1, ks0 = ks;
    auto issue = [&](int i){            // knot number i (processing order) into slot i & 1, by lane 0
        const int slot = i & 1; const size_t k = bN + ks0 - i;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the slot was last written in place (H assembly) through the generic proxy
        mbar_expect_tx(&s.full[slot], (AB_STRIDE + H_STRIDE + G_STRIDE)*4);
        tma_load_1d(s.AB[slot], S.AB + k*AB_STRIDE, AB_STRIDE*4, &s.full[slot]);
        tma_load_1d(s.Hc[slot], S.H + k*H_STRIDE, H_STRIDE*4, &s.full[slot]);
        tma_load_1d(s.gc[slot], S.g + k*G_STRIDE, G_STRIDE*4, &s.full[slot]);
    };
    if (l == 0){
        mbar_init(&s.full[0], 1); mbar_init(&s.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        issue(0);
    }
    // ---- cost-to-go at the right edge of the block
    float *gPcur = S.Pbuf[cur], *gpcur = S.pbuf[cur];
    if (last_block){
        // final block: Hxx[N-1] -> P[N-2], gx[N-1] -> p[N-2]  (bpHelpers.cuh:362-367)
        const size_t kN = bN + N - 1;
        for (int e = l; e < n*n; e += 32){
            const int kx = e % n, ky = e / n; const float v = MUL(1.0f, S.H[kN*H_STRIDE + kx + nm*ky]);
            s.P[e] = v; s.Pr[e] = (kx == ky) ? ADD(v, rho) : v; gPcur[(kN-1)*n*n + e] = v;
        }
        if (l < n){ const float v = MUL(1.0f, S.g[kN*G_STRIDE + l]); s.p[l] = v; gpcur[(kN-1)*n + l] = v; }
    } else {
        // other blocks: the previous iteration's P, p at the block boundary, p shifted to the new linearisation point (:369,376)
        const float *gPp = S.Pbuf[cur^1] + (bN + ks)*n*n;
        for (int e = l; e < n*n; e += 32){ const float v = gPp[e]; s.P[e] = v; s.Pr[e] = (e % n == e / n) ? ADD(v, rho) : v; }
        if (l < n){ s.dx[l] = SUB(S.xp[(bN + ks + 1)*n + l], S.xp2[(bN + ks + 1)*n + l]); }
        __syncwarp();
        if (l < n){
            float val = 0.f;
            #pragma unroll
            for (int j = 0; j < n; j++){ val = FMA(s.P[l + n*j], s.dx[j], val); }
            s.p[l] = FMA(1.0f, val, S.pbuf[cur^1][(bN + ks)*n + l]);
        }
    }
    __syncwarp();
    float dJ0 = 0.f, dJ1 = 0.f;               // lanes 0..6: running sums of du*gu and du*(Huu du) (bpHelpers.cuh:327-328)
    const int hl = min(l & 7, 6), hg = l >> 3;                      // Huu: lane = l' + 8*(column pair)
    #pragma unroll 1
    for (int iter = iterCount, i = 0; iter >= 0; iter--, ks--, i++){
        const int slot = i & 1;
        if (l == 0 && i + 1 < nknots){ issue(i + 1); }              // the other slot was released by the __syncwarp that ended knot i-1
        mbar_wait(&s.full[slot], (i >> 1) & 1);
        const float *sAB = s.AB[slot], *bg = s.gc[slot]; float *sH = s.Hc[slot];
        const size_t kk = bN + ks;
        const bool boundary = S.M > 1 && iter == NBB - 1;           // block-local defect-boundary test of the reference (bpHelpers.cuh:73)
        // ---- stage A: AB2 = AB'(P + rho I[u rows]); p += P d on the boundary
        BPW_UNROLL
        for (int tile = l; tile < 56; tile += 32){
            const int a_g = tile & 7, a_kp = tile >> 3;
            const int a_kx0 = a_g < 5 ? 3*a_g : n + 3*(a_g - 5), a_cnt = (a_g == 4) ? 2 : (a_g == 7 ? 1 : 3);
            const int a_kx1 = a_cnt > 1 ? a_kx0 + 1 : a_kx0, a_kx2 = a_cnt > 2 ? a_kx0 + 2 : a_kx0;
            float x0[14], x1[14], x2[14], y0[14], y1[14];
            const float *Pq = (a_kx0 >= n) ? s.Pr : s.P;
            bpw_ld14(x0, sAB + a_kx0*n); bpw_ld14(x1, sAB + a_kx1*n); bpw_ld14(x2, sAB + a_kx2*n);
            bpw_ld14(y0, Pq + (2*a_kp)*n); bpw_ld14(y1, Pq + (2*a_kp+1)*n);
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
        }
        if (l < n){
            float val = 0.f;
            if (boundary){
                const float *gd = S.dp + kk*n;
                #pragma unroll
                for (int j = 0; j < n; j++){ val = FMA(gd[j], s.P[l + n*j], val); }
            }
            s.p[l] = ADD(s.p[l], val);
        }
        __syncwarp();
        // ---- stage B: H = (AB2 AB)' + H_cost (in place in the ring slot), g = AB'p + g_cost
        // region 1 (rows ky < n, all columns): tile = 2 rows of AB2 x 3 columns of AB
        BPW_UNROLL
        for (int tile = l; tile < 49; tile += 32){
            const int kx0 = 3*(tile % 7), r0 = 2*(tile / 7);
            float x0[14], x1[14], x2[14], y0[14], y1[14];
            bpw_ld14(x0, sAB + kx0*n); bpw_ld14(x1, sAB + (kx0+1)*n); bpw_ld14(x2, sAB + (kx0+2)*n);
            bpw_ld14(y0, s.AB2 + r0*n); bpw_ld14(y1, s.AB2 + (r0+1)*n);
            float q[6];
            #pragma unroll
            for (int c = 0; c < 3; c++){ q[c] = sH[kx0 + c + nm*r0]; q[3+c] = sH[kx0 + c + nm*(r0+1)]; }
            float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            #pragma unroll
            for (int j = 0; j < n; j++){
                v[0] = FMA(y0[j], x0[j], v[0]); v[1] = FMA(y0[j], x1[j], v[1]); v[2] = FMA(y0[j], x2[j], v[2]);
                v[3] = FMA(y1[j], x0[j], v[3]); v[4] = FMA(y1[j], x1[j], v[4]); v[5] = FMA(y1[j], x2[j], v[5]);
            }
            #pragma unroll
            for (int c = 0; c < 3; c++){
                const float h0 = FMA(1.0f, v[c], MUL(1.0f, q[c])), h1 = FMA(1.0f, v[3+c], MUL(1.0f, q[3+c]));
                sH[kx0 + c + nm*r0] = h0; sH[kx0 + c + nm*(r0+1)] = h1;
                if (kx0 + c >= n){ s.Hux[r0*BPW_RS + kx0 + c - n] = h0; s.Hux[(r0+1)*BPW_RS + kx0 + c - n] = h1; }
            }
        }
        // region 2 (rows ky >= n, columns < n): tile = 1 row of AB2 x 2 columns of AB
        BPW_UNROLL
        for (int tile = l; tile < 49; tile += 32){
            const int kxp = tile % 7, ky = n + tile / 7;
            float y[14], x0[14], x1[14];
            bpw_ld14(y, s.AB2 + ky*n); bpw_ld14(x0, sAB + (2*kxp)*n); bpw_ld14(x1, sAB + (2*kxp+1)*n);
            const float q0 = sH[2*kxp + nm*ky], q1 = sH[2*kxp + 1 + nm*ky];
            float v0 = 0.f, v1 = 0.f;
            #pragma unroll
            for (int j = 0; j < n; j++){ v0 = FMA(y[j], x0[j], v0); v1 = FMA(y[j], x1[j], v1); }
            sH[2*kxp + nm*ky] = FMA(1.0f, v0, MUL(1.0f, q0)); sH[2*kxp + 1 + nm*ky] = FMA(1.0f, v1, MUL(1.0f, q1));
        }
        if (l < nm){
            float x[14], y[14];
            bpw_ld14(x, s.p); bpw_ld14(y, sAB + l*n);
            s.g[l] = FMA(1.0f, bpw_dot<n>(x, y), MUL(1.0f, bg[l]));
        }
        // Huu: lane l' + 8*q computes Huu(l', 2q) and Huu(l', 2q+1); the rows are then gathered on lanes 0..6 for the elimination
        {
            float x[14], y0[14], y1[14];
            const int c0 = min(2*hg, 6), c1 = min(2*hg + 1, 6);
            bpw_ld14(x, sAB + (n + hl)*n); bpw_ld14(y0, s.AB2 + (n + c0)*n); bpw_ld14(y1, s.AB2 + (n + c1)*n);
            const float q0 = sH[(n + hl) + nm*(n + c0)], q1 = sH[(n + hl) + nm*(n + c1)];
            float v0 = 0.f, v1 = 0.f;
            #pragma unroll
            for (int j = 0; j < n; j++){ v0 = FMA(y0[j], x[j], v0); v1 = FMA(y1[j], x[j], v1); }
            const float h0 = FMA(1.0f, v0, MUL(1.0f, q0)), h1 = FMA(1.0f, v1, MUL(1.0f, q1));
            __syncwarp();                         // every lane has read its Hc operands of the u x u block before it is overwritten
            if ((l & 7) < 7){
                sH[(n + hl) + nm*(n + c0)] = h0; s.Huu[c0*BPW_RS + hl] = h0;
                if (2*hg + 1 < m){ sH[(n + hl) + nm*(n + c1)] = h1; s.Huu[c1*BPW_RS + hl] = h1; }
            }
            float a[2*m];
            #pragma unroll
            for (int c = 0; c < m; c++){ a[c] = MUL(1.0f, __shfl_sync(FULL, (c & 1) ? h1 : h0, (l & 7) + 8*(c >> 1))); a[m + c] = (l == c) ? 1.f : 0.f; }
            gauss_jordan_rows<m>(a, l);
            if (l < m){
                *reinterpret_cast<float4*>(&s.Hinv[l*BPW_RS]) = make_float4(a[m], a[m+1], a[m+2], a[m+3]);
                *reinterpret_cast<float4*>(&s.Hinv[l*BPW_RS + 4]) = make_float4(a[m+4], a[m+5], a[m+6], 0.f);
            }
        }
        __syncwarp();
        // ---- stage C: K = Huu^-1 Hux, du = Huu^-1 gu
        BPW_UNROLL
        for (int e = l; e < n*m; e += 32){
            const int c_kx = e % m, c_ky = e / m;
            float hi[8], hx[8];
            bpw_ld8(hi, s.Hinv + c_kx*BPW_RS); bpw_ld8(hx, s.Hux + c_ky*BPW_RS);
            s.K[c_ky*BPW_RS + c_kx] = MUL(1.0f, bpw_dot<m>(hi, hx));
        }
        if (l < m){
            float hi[8], gu[8];
            bpw_ld8(hi, s.Hinv + l*BPW_RS);
            #pragma unroll
            for (int j = 0; j < m; j++){ gu[j] = s.g[n + j]; }
            gu[7] = 0.f;
            s.du[l] = ADD(MUL(1.0f, bpw_dot<m>(hi, gu)), 0.f);
        }
        __syncwarp();
        // ---- stage D: T = K'Huu - Hxu (into the AB2 workspace); expected reduction
        const bool do_ctg = (iter != 0 || block != 0);
        float *sT = s.AB2;
        if (do_ctg){
            BPW_UNROLL
            for (int e = l; e < n*m; e += 32){
                const int d_kx = e % n, d_ky = e / n;
                float k[8], h[8];
                bpw_ld8(k, s.K + d_kx*BPW_RS); bpw_ld8(h, s.Huu + d_ky*BPW_RS);
                const float hxu = sH[d_kx + nm*(n + d_ky)];
                sT[d_kx*BPW_RS + d_ky] = SUB(bpw_dot<m>(k, h), hxu);
            }
        }
        if (l < m){
            float dot = 0.f;
            #pragma unroll
            for (int j = 0; j < m; j++){ dot = FMA(s.Huu[j*BPW_RS + l], s.du[j], dot); }
            dJ0 = FMA(s.du[l], s.g[n + l], dJ0); dJ1 = FMA(s.du[l], dot, dJ1);
        }
        __syncwarp();
        // ---- stage F (before E: it reads K, du and the knot's AB, none of which E changes): A - BK, B du, KT, du -> HBM
        if (S.M > 1){
            BPW_UNROLL
            for (int e = l; e < 98; e += 32){
                const int kx = e % n, kyp = e / n;
                float bb[8], k0[8], k1[8];
                #pragma unroll
                for (int j = 0; j < m; j++){ bb[j] = sAB[oB + kx + n*j]; }
                bb[7] = 0.f;
                bpw_ld8(k0, s.K + (2*kyp)*BPW_RS); bpw_ld8(k1, s.K + (2*kyp+1)*BPW_RS);
                const float a0 = sAB[kx + n*(2*kyp)], a1 = sAB[kx + n*(2*kyp+1)];
                S.ApBK[kk*n*n + kx + n*(2*kyp)] = SUB(a0, bpw_dot<m>(bb, k0));
                S.ApBK[kk*n*n + kx + n*(2*kyp+1)] = SUB(a1, bpw_dot<m>(bb, k1));
            }
            if (l < n){
                float val = 0.f;
                #pragma unroll
                for (int j = 0; j < m; j++){ val = FMA(sAB[oB + l + n*j], s.du[j], val); }
                S.Bdu[kk*n + l] = val;
            }
        }
        for (int e = l; e < n*m; e += 32){ const int kx = e % n, ky = e / n; S.KT[kk*n*m + e] = s.K[kx*BPW_RS + ky]; }
        if (l < m){ S.du[kk*m + l] = s.du[l]; }
        // ---- stage E: cost-to-go of the previous knot
        if (do_ctg){
            float pv = 0.f;
            if (l < n){             // p first: it reads g and T, and nothing below changes them
                float a[8], k2[8];
                bpw_ld8(a, sT + l*BPW_RS); bpw_ld8(k2, s.K + l*BPW_RS);
                float val = 0.f;
                #pragma unroll
                for (int j = 0; j < m; j++){ val = ADD(val, FMA(s.du[j], a[j], -MUL(k2[j], s.g[n + j]))); }
                pv = ADD(s.g[l], val);
            }
            BPW_UNROLL
            for (int tile = l; tile < 49; tile += 32){
                const int kx0 = 2*(tile % 7), ky0 = 2*(tile / 7);
                float a0[8], a1[8], kx_0[8], kx_1[8], ky_0[8], ky_1[8], h0[8], h1[8];
                bpw_ld8(a0, sT + kx0*BPW_RS); bpw_ld8(a1, sT + (kx0+1)*BPW_RS); bpw_ld8(kx_0, s.K + kx0*BPW_RS); bpw_ld8(kx_1, s.K + (kx0+1)*BPW_RS);
                bpw_ld8(ky_0, s.K + ky0*BPW_RS); bpw_ld8(ky_1, s.K + (ky0+1)*BPW_RS); bpw_ld8(h0, s.Hux + ky0*BPW_RS); bpw_ld8(h1, s.Hux + (ky0+1)*BPW_RS);
                const float x00 = sH[kx0 + ky0*nm], x10 = sH[kx0 + 1 + ky0*nm], x01 = sH[kx0 + (ky0+1)*nm], x11 = sH[kx0 + 1 + (ky0+1)*nm];
                float v00 = 0.f, v10 = 0.f, v01 = 0.f, v11 = 0.f;
                #pragma unroll
                for (int j = 0; j < m; j++){
                    v00 = ADD(v00, FMA(a0[j], ky_0[j], -MUL(kx_0[j], h0[j]))); v10 = ADD(v10, FMA(a1[j], ky_0[j], -MUL(kx_1[j], h0[j])));
                    v01 = ADD(v01, FMA(a0[j], ky_1[j], -MUL(kx_0[j], h1[j]))); v11 = ADD(v11, FMA(a1[j], ky_1[j], -MUL(kx_1[j], h1[j])));
                }
                const float p00 = ADD(x00, v00), p10 = ADD(x10, v10), p01 = ADD(x01, v01), p11 = ADD(x11, v11);
                *reinterpret_cast<float2*>(&s.P[ky0*n + kx0]) = make_float2(p00, p10);
                *reinterpret_cast<float2*>(&s.P[(ky0+1)*n + kx0]) = make_float2(p01, p11);
                const bool dg = (kx0 == ky0);
                *reinterpret_cast<float2*>(&s.Pr[ky0*n + kx0]) = make_float2(dg ? ADD(p00, rho) : p00, p10);
                *reinterpret_cast<float2*>(&s.Pr[(ky0+1)*n + kx0]) = make_float2(p01, dg ? ADD(p11, rho) : p11);
                float *gP = gPcur + (kk-1)*n*n;
                *reinterpret_cast<float2*>(&gP[ky0*n + kx0]) = make_float2(p00, p10);
                *reinterpret_cast<float2*>(&gP[(ky0+1)*n + kx0]) = make_float2(p01, p11);
            }
            if (l < n){ s.p[l] = pv; gpcur[(kk-1)*n + l] = pv; }
        }
        __syncwarp();
    }
    // ---- expected cost reduction of this block: the m per-lane partials summed in order (bpHelpers.cuh:416)
    float a0 = __shfl_sync(FULL, dJ0, 0), a1 = __shfl_sync(FULL, dJ1, 0);
    #pragma unroll
    for (int j = 1; j < m; j++){ a0 = ADD(a0, __shfl_sync(FULL, dJ0, j)); a1 = ADD(a1, __shfl_sync(FULL, dJ1, j)); }
    if (l == 0){ S.dJexp[(size_t)b*2*S.M + 2*block] = a0; S.dJexp[(size_t)b*2*S.M + 2*block + 1] = a1; }
}

} // namespace pddp
