// epilogue.cu -- the fused ALM pass: ONE sweep over a 16-row tile performs (src/robustPCA.jl:188-222)
//   E  = soft_th((D - A) + Y/mu, lambda/mu) [max(.,0)]      :188-191
//   W  = (D - E) + Y/mu                                      :192
//   T  = W V_r                     (DMMA, K = n)             <- replaces U_r S_r of svd! (:194)
//   A' = T diag(f) V_r'            (DMMA, K = svp)           :205-213   f = 1 - 1/(mu s_i)  (1 when nukeA=false)
//   A' = max(A',0)                                           :217-219
//   Z  = (D - A') - E ;  Y' = Y + mu Z ;  ||Z||_F^2          :221-222 (+ Frobenius bracket of the :225 stop test)
// A and Y are read once from HBM (the second touch in phase 4 hits L2: the tile was fetched moments earlier by the
// same CTA) and written once.  W/T never leave the SM.
//
// Layout per CTA (256 threads, persistent over row tiles):
//   Wsm [NP][20]  W tile, column-major with padded row stride 20 (R = 16 rows): DMMA fragment loads are conflict free
//   Vsm [NP][20]  chunk of 16 right singular vectors, Vsm[j][c] = V[j, c0 + c]
//   Tpart[64][20] partial T of the 4 K-quarters, Tsm[16][20] reduced & scaled T
//   the finished A' tile is staged back through Wsm to restore coalesced global access.
// MODE_U variant: writes U[:, c] = (W v_c) / s_c instead of A' (left singular vectors of the returned SVD, :238).
#include "kernels.h"

namespace tlsq {

namespace {

constexpr int ER = 16;        // rows per tile
constexpr int ERS = ER + 4;   // padded row stride of Wsm (20: 20 mod 16 == 4 -> conflict-free fragment loads)
constexpr int EVS = 20;       // padded stride of Vsm / Tsm / Tpart rows

template <int C, bool HANKEL, bool MODE_U>
__global__ void __launch_bounds__(256, 2)
epilogue_kernel(const EpiArgs a, int ntiles) {
    constexpr int NP = 64 * C;
    extern __shared__ double sm[];
    double* Wsm = sm;                      // NP * ERS
    double* Vsm = Wsm + NP * ERS;          // NP * EVS
    double* Tpart = Vsm + NP * EVS;        // 64 * EVS
    double* Tsm = Tpart + 64 * EVS;        // 16 * EVS
    __shared__ double zred[8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int svp = *a.svp;
    const int nchunks = (svp + 15) / 16;
    const int erow = tid & (ER - 1), ecol0 = tid >> 4;     // element-wise mapping: 16 rows x 16 column groups
    const int N = (int)a.N;
    bool v_resident = false;
    double zz = 0.0;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row = (int64_t)tile * ER + erow;
        const bool rok = row < a.M;

        // ---- phase 1: element-wise E-step, W -> shared --------------------------------------------------------
        for (int col = ecol0; col < NP; col += 16) {
            double w = 0.0;
            if (rok && col < N) {
                const double d = src_at<HANKEL>(a.D, row, col);
                const int64_t off = (int64_t)col * a.ldw + row;
                double e;
                alm_ew(d, __ldg(a.Ap + off), __ldg(a.Yp + off), a.im, a.eps, a.nonnegE, e, w);
            }
            Wsm[col * ERS + erow] = w;
        }

        double acc[2][C][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < C; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

        for (int ch = 0; ch < nchunks; ++ch) {
            const int c0 = ch * 16;
            if (!v_resident) {
                __syncthreads();           // previous users of Vsm (phase 3 of the previous chunk / tile) are done
                for (int idx = tid; idx < NP * 16; idx += 256) {
                    const int j = idx % NP, c = idx / NP;
                    double v = 0.0;
                    if (j < N && c0 + c < svp) v = __ldg(a.Vs + (int64_t)(c0 + c) * N + j);
                    Vsm[j * EVS + c] = v;
                }
                v_resident = (nchunks == 1);
            }
            __syncthreads();               // Wsm and Vsm ready

            // ---- phase 2: T(16 x 16) = W(16 x NP) * Vchunk(NP x 16); warp = (m-tile, K-quarter) ---------------
            {
                const int mi = warp & 1, ks = warp >> 1;
                const int jbeg = ks * (NP / 4), jend = jbeg + NP / 4;
                double tacc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll 4
                for (int j0 = jbeg; j0 < jend; j0 += 4) {
                    const double av = Wsm[(j0 + t) * ERS + 8 * mi + g];
                    const double b0 = Vsm[(j0 + t) * EVS + g];
                    const double b1 = Vsm[(j0 + t) * EVS + 8 + g];
                    dmma884(tacc[0][0], tacc[0][1], av, b0);
                    dmma884(tacc[1][0], tacc[1][1], av, b1);
                }
                double* tp = Tpart + (ks * 16 + 8 * mi + g) * EVS;
                tp[2 * t] = tacc[0][0];
                tp[2 * t + 1] = tacc[0][1];
                tp[8 + 2 * t] = tacc[1][0];
                tp[8 + 2 * t + 1] = tacc[1][1];
            }
            __syncthreads();
            {   // reduce the 4 K-quarters (fixed order), scale by the shrink factor
                const int i = tid & 15, c = tid >> 4;
                double s = Tpart[i * EVS + c];
                s += Tpart[(16 + i) * EVS + c];
                s += Tpart[(32 + i) * EVS + c];
                s += Tpart[(48 + i) * EVS + c];
                const double f = (c0 + c < svp) ? __ldg(a.fvec + c0 + c) : 0.0;
                if (MODE_U) {
                    const int64_t r = (int64_t)tile * ER + i;
                    if (r < a.M && c0 + c < svp) a.Uout[(int64_t)(c0 + c) * a.ldw + r] = s * f;
                } else {
                    Tsm[i * EVS + c] = s * f;
                }
            }
            if (!MODE_U) {
                __syncthreads();
                // ---- phase 3: A'(16 x NP) += T'(16 x 16) * Vchunk'(16 x NP); warp owns 8*C columns ----------------
#pragma unroll
                for (int k0 = 0; k0 < 16; k0 += 4) {
                    double av[2];
                    av[0] = Tsm[g * EVS + k0 + t];
                    av[1] = Tsm[(8 + g) * EVS + k0 + t];
#pragma unroll
                    for (int ni = 0; ni < C; ++ni) {
                        const double bv = Vsm[(8 * C * warp + 8 * ni + g) * EVS + k0 + t];
                        dmma884(acc[0][ni][0], acc[0][ni][1], av[0], bv);
                        dmma884(acc[1][ni][0], acc[1][ni][1], av[1], bv);
                    }
                }
            }
        }
        if (MODE_U) { __syncthreads(); continue; }

        __syncthreads();                   // all reads of Wsm (phase 2) and Tsm/Vsm (phase 3) done
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < C; ++ni) {
                const int i = 8 * mi + g;
                const int j = 8 * C * warp + 8 * ni + 2 * t;
                Wsm[j * ERS + i] = acc[mi][ni][0];
                Wsm[(j + 1) * ERS + i] = acc[mi][ni][1];
            }
        __syncthreads();

        // ---- phase 4: clamp, residual, dual update (coalesced; D/A/Y re-read from L2) ---------------------------
        if (rok && a.raw_only) {
            for (int col = ecol0; col < N; col += 16) a.An[(int64_t)col * a.ldw + row] = Wsm[col * ERS + erow];
        } else if (rok) {
            for (int col = ecol0; col < N; col += 16) {
                const double d = src_at<HANKEL>(a.D, row, col);
                const int64_t off = (int64_t)col * a.ldw + row;
                const double ap = __ldg(a.Ap + off);
                const double yp = __ldg(a.Yp + off);
                double e, w;
                alm_ew(d, ap, yp, a.im, a.eps, a.nonnegE, e, w);
                double an = Wsm[col * ERS + erow];
                if (a.nonnegA) an = (__double_as_longlong(an) > 0) ? an : 0.0;       // A .= max.(A, 0)   :218
                const double z = __dsub_rn(__dsub_rn(d, an), e);                      // @. Z = D - A - E  :221
                const double yn = __dadd_rn(yp, __dmul_rn(a.mu, z));                  // @. Y = Y + mu*Z   :222
                zz = fma(z, z, zz);
                a.An[off] = an;
                a.Yn[off] = yn;
                if (a.Eout) a.Eout[off] = e;
                if (a.Zout) a.Zout[off] = z;
                if (a.Wn) {                                                           // SVT input of iteration k+1
                    double e2, w2;
                    alm_ew(d, an, yn, a.im_next, a.eps_next, a.nonnegE, e2, w2);
                    a.Wn[off] = w2;
                }
            }
        }
        __syncthreads();                   // Wsm is rewritten by phase 1 of the next tile
    }

    if (!MODE_U) {
        zz = warp_sum(zz);
        if (lane == 0) zred[warp] = zz;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += zred[w];
            atomicAdd(a.zz, s);
        }
    }
}

template <int C>
cudaError_t launch_c(const EpiArgs& a, bool hankel, bool mode_u, int sm_count, cudaStream_t st) {
    constexpr int NP = 64 * C;
    const size_t smem = (size_t)(NP * ERS + NP * EVS + 64 * EVS + 16 * EVS) * sizeof(double);
    const int ntiles = (int)((a.M + ER - 1) / ER);
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    int grid = sm_count * per_sm;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    cudaError_t e;
#define TLSQ_EPI_LAUNCH(H, U)                                                                               \
    do {                                                                                                    \
        auto kern = epilogue_kernel<C, H, U>;                                                               \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
        if (e != cudaSuccess) return e;                                                                     \
        kern<<<grid, 256, smem, st>>>(a, ntiles);                                                           \
    } while (0)
    if (hankel) { if (mode_u) TLSQ_EPI_LAUNCH(true, true); else TLSQ_EPI_LAUNCH(true, false); }
    else        { if (mode_u) TLSQ_EPI_LAUNCH(false, true); else TLSQ_EPI_LAUNCH(false, false); }
#undef TLSQ_EPI_LAUNCH
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_epilogue(const EpiArgs& a, bool hankel, bool mode_u, int sm_count, cudaStream_t st,
                            int64_t* launches) {
    if (launches) *launches += 1;
    if (a.N <= 64) return launch_c<1>(a, hankel, mode_u, sm_count, st);
    if (a.N <= 128) return launch_c<2>(a, hankel, mode_u, sm_count, st);
    if (a.N <= 256) return launch_c<4>(a, hankel, mode_u, sm_count, st);
    return launch_c<8>(a, hankel, mode_u, sm_count, st);
}

}  // namespace tlsq
