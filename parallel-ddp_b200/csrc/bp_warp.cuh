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
constexpr int BPW_P = 20;                // row pitch of P, P + rho I and AB2: a 14-vector is four 16-byte loads (two per half); 20 words put the
                                         // rows r, r+1, ... r+7 on disjoint bank quads, so lanes that read consecutive rows do not conflict
struct __align__(16) BpWarpSmem {
    float AB[2][AB_STRIDE], Hc[2][H_STRIDE], gc[2][G_STRIDE];      // ring of knot inputs; H is assembled in place in Hc
    float P[14*BPW_P], Pr[14*BPW_P];                               // P[ky][kx] and P + rho on the diagonal (the operand of the u-rows of AB'(.))
    float AB2[21*BPW_P];                                           // AB2[kx][ky]; T[kx*BPW_RS + j] lives here from stage D on
    float Hux[14*BPW_RS], K[14*BPW_RS], Huu[7*BPW_RS], Hinv[7*BPW_RS];
    float g[24], p[16], dx[16], du[8], pad[8];
    unsigned long long full[2];
};
static_assert(sizeof(BpWarpSmem) % 16 == 0, "warp workspaces must keep 16-byte alignment");

__device__ __forceinline__ void bpw_ld8(float (&x)[8], const float *p){
    const float4 *q = reinterpret_cast<const float4*>(p);
    const float4 a = q[0], b = q[1];
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
// eight (HALF = 0: entries 0..7) or six (HALF = 1: entries 8..13) floats of a 14-vector stored with 8-byte alignment (the rows of AB)
template <int HALF>
__device__ __forceinline__ void bpw_ldh64(float (&x)[8], const float *p){
    const float2 *q = reinterpret_cast<const float2*>(p + 8*HALF);
    #pragma unroll
    for (int i = 0; i < (HALF ? 3 : 4); i++){ const float2 v = q[i]; x[2*i] = v.x; x[2*i+1] = v.y; }
    if (HALF){ x[6] = 0.f; x[7] = 0.f; }
}
template <int K, int KP>
__device__ __forceinline__ float bpw_dot(const float (&x)[KP], const float (&y)[KP]){
    float val = 0.f;
    #pragma unroll
    for (int j = 0; j < K; j++){ val = FMA(x[j], y[j], val); }
    return val;
}

#ifndef PDDP_BPW_MINCTA
#define PDDP_BPW_MINCTA 4
#endif
__global__ void __launch_bounds__(32*BPW_WARPS, PDDP_BPW_MINCTA) bp_warp_kernel(DevState S, int b0, int nchains){
    constexpr int n = 14, m = 7, nm = 21, oB = n*n, PP = BPW_P;
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
            s.P[ky*PP + kx] = v; s.Pr[ky*PP + kx] = (kx == ky) ? ADD(v, rho) : v; gPcur[(kN-1)*n*n + e] = v;
        }
        if (l < n){ const float v = MUL(1.0f, S.g[kN*G_STRIDE + l]); s.p[l] = v; gpcur[(kN-1)*n + l] = v; }
    } else {
        // other blocks: the previous iteration's P, p at the block boundary, p shifted to the new linearisation point (:369,376)
        const float *gPp = S.Pbuf[cur^1] + (bN + ks)*n*n;
        for (int e = l; e < n*n; e += 32){ const int kx = e % n, ky = e / n; const float v = gPp[e]; s.P[ky*PP + kx] = v; s.Pr[ky*PP + kx] = (kx == ky) ? ADD(v, rho) : v; }
        if (l < n){ s.dx[l] = SUB(S.xp[(bN + ks + 1)*n + l], S.xp2[(bN + ks + 1)*n + l]); }
        __syncwarp();
        if (l < n){
            float val = 0.f;
            #pragma unroll
            for (int j = 0; j < n; j++){ val = FMA(s.P[l + PP*j], s.dx[j], val); }
            s.p[l] = FMA(1.0f, val, S.pbuf[cur^1][(bN + ks)*n + l]);
        }
    }
    __syncwarp();
    float dJ0 = 0.f, dJ1 = 0.f;               // lanes 0..6: running sums of du*gu and du*(Huu du) (bpHelpers.cuh:327-328)
    // ---- tile coordinates (fixed per lane)
    // stage A: (up to 3 columns kx of AB, not straddling the x / u boundary) x (rows ky = q, q+4, q+8, q+12 of P): 8 x 4 = 32 tiles.
    // Row groups are interleaved: at any one load the four groups read four CONSECUTIVE rows (disjoint banks), never rows 4 apart
    const int a_g = l & 7, a_q = l >> 3;
    const int a_kx0 = a_g < 5 ? 3*a_g : n + 3*(a_g - 5), a_cnt = (a_g == 4) ? 2 : (a_g == 7 ? 1 : 3);
    const int a_kx1 = a_cnt > 1 ? a_kx0 + 1 : a_kx0, a_kx2 = a_cnt > 2 ? a_kx0 + 2 : a_kx0;
    const int a_ky0 = a_q, a_ky1 = a_q + 4, a_ky2 = a_q + 8, a_ky3 = a_q < 2 ? a_q + 12 : a_q;                  // rows 12, 13 exist for q = 0, 1 only
    // stage B: (rows r = g, g+6, g+12, g+18 of AB2) x (5 columns kx of AB): 6 x 5 = 30 tiles of the 21 x 21 matrix H (rows interleaved as above)
    const int h_rg = min(l / 5, 5), h_cg = l % 5;
    const int h_r0 = h_rg, h_rc = h_rg < 3 ? 4 : 3, h_k0 = 5*h_cg, h_kc = h_cg < 4 ? 5 : 1;
    // stage E: (2 rows kx) x (4 rows ky) of P: 7 x 4 = 28 tiles
    const int e_kx0 = 2*(l % 7), e_q = min(l / 7, 3), e_ky0 = 4*e_q, e_kc = e_q < 3 ? 4 : 2;
    #pragma unroll 1
    for (int iter = iterCount, i = 0; iter >= 0; iter--, ks--, i++){
        const int slot = i & 1;
        if (l == 0 && i + 1 < nknots){ issue(i + 1); }              // the other slot was released by the __syncwarp that ended knot i-1
        mbar_wait(&s.full[slot], (i >> 1) & 1);
        const float *sAB = s.AB[slot], *bg = s.gc[slot]; float *sH = s.Hc[slot];
        const size_t kk = bN + ks;
        const bool boundary = S.M > 1 && iter == NBB - 1;           // block-local defect-boundary test of the reference (bpHelpers.cuh:73)
        // ---- stage A: AB2 = AB'(P + rho I[u rows]); p += P d on the boundary
        {
            const float *Pq = (a_kx0 >= n) ? s.Pr : s.P;
            float v[3][4];
            #pragma unroll
            for (int c = 0; c < 3; c++){
                #pragma unroll
                for (int r = 0; r < 4; r++){ v[c][r] = 0.f; }
            }
            #define BPW_A_HALF(HALF) { \
                float x0[8], x1[8], x2[8], y0[8], y1[8], y2[8], y3[8]; \
                bpw_ldh64<HALF>(x0, sAB + a_kx0*n); bpw_ldh64<HALF>(x1, sAB + a_kx1*n); bpw_ldh64<HALF>(x2, sAB + a_kx2*n); \
                bpw_ld8(y0, Pq + a_ky0*PP + 8*HALF); bpw_ld8(y1, Pq + a_ky1*PP + 8*HALF); bpw_ld8(y2, Pq + a_ky2*PP + 8*HALF); bpw_ld8(y3, Pq + a_ky3*PP + 8*HALF); \
                _Pragma("unroll") \
                for (int j = 0; j < (HALF ? 6 : 8); j++){ \
                    v[0][0] = FMA(x0[j], y0[j], v[0][0]); v[0][1] = FMA(x0[j], y1[j], v[0][1]); v[0][2] = FMA(x0[j], y2[j], v[0][2]); v[0][3] = FMA(x0[j], y3[j], v[0][3]); \
                    v[1][0] = FMA(x1[j], y0[j], v[1][0]); v[1][1] = FMA(x1[j], y1[j], v[1][1]); v[1][2] = FMA(x1[j], y2[j], v[1][2]); v[1][3] = FMA(x1[j], y3[j], v[1][3]); \
                    v[2][0] = FMA(x2[j], y0[j], v[2][0]); v[2][1] = FMA(x2[j], y1[j], v[2][1]); v[2][2] = FMA(x2[j], y2[j], v[2][2]); v[2][3] = FMA(x2[j], y3[j], v[2][3]); \
                } }
            BPW_A_HALF(0) BPW_A_HALF(1)
            #undef BPW_A_HALF
            #pragma unroll
            for (int c = 0; c < 3; c++){
                if (c < a_cnt){
                    float *o = &s.AB2[(a_kx0 + c)*PP];
                    o[a_ky0] = v[c][0]; o[a_ky1] = v[c][1]; o[a_ky2] = v[c][2];
                    if (a_q < 2){ o[a_ky3] = v[c][3]; }
                }
            }
        }
        if (l < n){
            float val = 0.f;
            if (boundary){
                const float *gd = S.dp + kk*n;
                #pragma unroll
                for (int j = 0; j < n; j++){ val = FMA(gd[j], s.P[l + PP*j], val); }
            }
            s.p[l] = ADD(s.p[l], val);
        }
        __syncwarp();
        // ---- stage B: H = (AB2 AB)' + H_cost (in place in the ring slot), g = AB'p + g_cost
        if (l < 30){
            float v[4][5];
            #pragma unroll
            for (int r = 0; r < 4; r++){
                #pragma unroll
                for (int c = 0; c < 5; c++){ v[r][c] = 0.f; }
            }
            const int r1 = h_r0 + 6, r2 = h_r0 + 12, r3 = h_rc > 3 ? h_r0 + 18 : h_r0;
            const int k1 = h_kc > 1 ? h_k0 + 1 : h_k0, k2 = h_kc > 1 ? h_k0 + 2 : h_k0, k3 = h_kc > 1 ? h_k0 + 3 : h_k0, k4 = h_kc > 1 ? h_k0 + 4 : h_k0;
            #define BPW_H_HALF(HALF) { \
                float y0[8], y1[8], y2[8], y3[8], x0[8], x1[8], x2[8], x3[8], x4[8]; \
                bpw_ld8(y0, s.AB2 + h_r0*PP + 8*HALF); bpw_ld8(y1, s.AB2 + r1*PP + 8*HALF); bpw_ld8(y2, s.AB2 + r2*PP + 8*HALF); bpw_ld8(y3, s.AB2 + r3*PP + 8*HALF); \
                bpw_ldh64<HALF>(x0, sAB + h_k0*n); bpw_ldh64<HALF>(x1, sAB + k1*n); bpw_ldh64<HALF>(x2, sAB + k2*n); bpw_ldh64<HALF>(x3, sAB + k3*n); bpw_ldh64<HALF>(x4, sAB + k4*n); \
                _Pragma("unroll") \
                for (int j = 0; j < (HALF ? 6 : 8); j++){ \
                    v[0][0] = FMA(y0[j], x0[j], v[0][0]); v[0][1] = FMA(y0[j], x1[j], v[0][1]); v[0][2] = FMA(y0[j], x2[j], v[0][2]); v[0][3] = FMA(y0[j], x3[j], v[0][3]); v[0][4] = FMA(y0[j], x4[j], v[0][4]); \
                    v[1][0] = FMA(y1[j], x0[j], v[1][0]); v[1][1] = FMA(y1[j], x1[j], v[1][1]); v[1][2] = FMA(y1[j], x2[j], v[1][2]); v[1][3] = FMA(y1[j], x3[j], v[1][3]); v[1][4] = FMA(y1[j], x4[j], v[1][4]); \
                    v[2][0] = FMA(y2[j], x0[j], v[2][0]); v[2][1] = FMA(y2[j], x1[j], v[2][1]); v[2][2] = FMA(y2[j], x2[j], v[2][2]); v[2][3] = FMA(y2[j], x3[j], v[2][3]); v[2][4] = FMA(y2[j], x4[j], v[2][4]); \
                    v[3][0] = FMA(y3[j], x0[j], v[3][0]); v[3][1] = FMA(y3[j], x1[j], v[3][1]); v[3][2] = FMA(y3[j], x2[j], v[3][2]); v[3][3] = FMA(y3[j], x3[j], v[3][3]); v[3][4] = FMA(y3[j], x4[j], v[3][4]); \
                } }
            BPW_H_HALF(0) BPW_H_HALF(1)
            #undef BPW_H_HALF
            #pragma unroll
            for (int r = 0; r < 4; r++){
                if (r < h_rc){
                    const int rr = h_r0 + 6*r;
                    #pragma unroll
                    for (int c = 0; c < 5; c++){
                        if (c < h_kc){
                            const int kx = h_k0 + c;
                            const float h = FMA(1.0f, v[r][c], MUL(1.0f, sH[kx + nm*rr]));
                            sH[kx + nm*rr] = h;
                            if (kx >= n){ if (rr < n){ s.Hux[rr*BPW_RS + kx - n] = h; } else { s.Huu[(rr - n)*BPW_RS + kx - n] = h; } }
                        }
                    }
                }
            }
        }
        if (l < nm){
            float val = 0.f;
            #pragma unroll
            for (int j = 0; j < n; j++){ val = FMA(s.p[j], sAB[l*n + j], val); }
            s.g[l] = FMA(1.0f, val, MUL(1.0f, bg[l]));
        }
        __syncwarp();
        // Huu^-1: row l of [Huu | I] on lane l < 7, 7 x 7 Gauss-Jordan by shuffles
        {
            float a[2*m];
            #pragma unroll
            for (int c = 0; c < m; c++){ a[c] = MUL(1.0f, s.Huu[c*BPW_RS + min(l, m - 1)]); a[m + c] = (l == c) ? 1.f : 0.f; }
            gauss_jordan_rows<m>(a, l);
            if (l < m){
                *reinterpret_cast<float4*>(&s.Hinv[l*BPW_RS]) = make_float4(a[m], a[m+1], a[m+2], a[m+3]);
                *reinterpret_cast<float4*>(&s.Hinv[l*BPW_RS + 4]) = make_float4(a[m+4], a[m+5], a[m+6], 0.f);
            }
        }
        __syncwarp();
        // ---- stage C: K = Huu^-1 Hux, du = Huu^-1 gu
        #pragma unroll
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
            #pragma unroll
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
            #pragma unroll
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
        #pragma unroll
        for (int e = l; e < n*m; e += 32){ const int kx = e % n, ky = e / n; S.KT[kk*n*m + e] = s.K[kx*BPW_RS + ky]; }
        if (l < m){ S.du[kk*m + l] = s.du[l]; }
        // ---- stage E: cost-to-go of the previous knot: tile = rows kx0, kx0+1 of T and K x up to 4 rows ky of K and Hux
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
            if (l < 28){
                float a0[8], a1[8], kx_0[8], kx_1[8];
                bpw_ld8(a0, sT + e_kx0*BPW_RS); bpw_ld8(a1, sT + (e_kx0+1)*BPW_RS); bpw_ld8(kx_0, s.K + e_kx0*BPW_RS); bpw_ld8(kx_1, s.K + (e_kx0+1)*BPW_RS);
                float *gP = gPcur + (kk-1)*n*n;
                #pragma unroll
                for (int r = 0; r < 4; r++){
                    if (r < e_kc){
                        const int ky = e_ky0 + r;
                        float ky_[8], h_[8];
                        bpw_ld8(ky_, s.K + ky*BPW_RS); bpw_ld8(h_, s.Hux + ky*BPW_RS);
                        const float x0 = sH[e_kx0 + ky*nm], x1 = sH[e_kx0 + 1 + ky*nm];
                        float v0 = 0.f, v1 = 0.f;
                        #pragma unroll
                        for (int j = 0; j < m; j++){
                            v0 = ADD(v0, FMA(a0[j], ky_[j], -MUL(kx_0[j], h_[j]))); v1 = ADD(v1, FMA(a1[j], ky_[j], -MUL(kx_1[j], h_[j])));
                        }
                        const float p0 = ADD(x0, v0), p1 = ADD(x1, v1);
                        *reinterpret_cast<float2*>(&s.P[ky*PP + e_kx0]) = make_float2(p0, p1);
                        *reinterpret_cast<float2*>(&s.Pr[ky*PP + e_kx0]) = make_float2(e_kx0 == ky ? ADD(p0, rho) : p0, e_kx0 + 1 == ky ? ADD(p1, rho) : p1);
                        *reinterpret_cast<float2*>(&gP[ky*n + e_kx0]) = make_float2(p0, p1);
                    }
                }
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
