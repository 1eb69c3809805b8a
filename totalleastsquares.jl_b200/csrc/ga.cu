// ga.cu -- Grassmann averages (rpca_ga, src/robustPCA.jl:255-316) as ONE streaming sweep of X per iteration.
//
// Reference iteration (:291-295): w_n = sign(u_n'q) ||x_n||  (sweep 1 of U),  mu = sum_n w_n u_n / sum_n w_n
// (sweep 2 of U, mu! :308-316),  q = mu/||mu||.  Because w_n u_n = sign(u_n'q) x_n the normalised copy U is never
// needed, and because the tile's rows of mu are complete as soon as the tile has been summed over n, the dot
// products x_n'mu that decide the NEXT iteration's signs (sign(u_n'q_new) = sign(x_n'mu)) are accumulated in the
// same sweep while the tile is still in shared memory.  One HBM read of X per iteration (the reference does two).
//
// ga_sweep_kernel<MODE>: persistent CTAs, row tiles of 32 rows x 256 columns staged in shared memory [col][32].
//   phase A (fused with the load): thread (row, column group) accumulates sum_n s_n X[row,n]
//   phase B: thread n accumulates sum_i X[i,n] mu_i over the tile with a skewed (conflict-free) row order.
#include "kernels.h"

namespace tlsq {

namespace {

constexpr int GA_R = 32;      // rows per tile
constexpr int GA_NC = 256;    // columns per chunk

template <int MODE>
__global__ void __launch_bounds__(256, 3)
ga_sweep_kernel(const double* __restrict__ X, int64_t d, int64_t N, int64_t ld, double* __restrict__ vec,
                const double* __restrict__ s, const double* __restrict__ sumw, double* __restrict__ t,
                int ntiles) {
    extern __shared__ double sm[];
    double* Xs = sm;                       // GA_NC * GA_R
    double* sv = Xs + GA_NC * GA_R;        // GA_NC   signs of the current chunk
    double* red = sv + GA_NC;              // 8 * 32  partial row sums
    double* mus = red + 8 * GA_R;          // 32      mu (or q) rows of the tile
    __shared__ double mmred[8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunks = (int)((N + GA_NC - 1) / GA_NC);
    const double inv_sumw_den = (MODE == GA_PASS) ? *sumw : 1.0;
    double tacc = 0.0;                     // column accumulator (single-chunk fast path)
    double mm = 0.0;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row = (int64_t)tile * GA_R + lane;
        const bool rok = row < d;

        if (MODE == GA_PASS) {
            // ---- phase A over all chunks: v_row = sum_n s_n X[row, n] -------------------------------------------
            double vpart = 0.0;
            for (int ch = 0; ch < nchunks; ++ch) {
                const int64_t cbase = (int64_t)ch * GA_NC;
                __syncthreads();
                if (cbase + tid < N) sv[tid] = s[cbase + tid]; else sv[tid] = 0.0;
                __syncthreads();
#pragma unroll 8
                for (int k = 0; k < 32; ++k) {
                    const int c = warp + 8 * k;
                    double x = 0.0;
                    if (rok && cbase + c < N) x = __ldg(X + (cbase + c) * ld + row);
                    if (nchunks == 1) Xs[c * GA_R + lane] = x;
                    vpart = fma(sv[c], x, vpart);
                }
            }
            red[warp * GA_R + lane] = vpart;
            __syncthreads();
            if (warp == 0) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) v += red[w * GA_R + lane];
                const double m = v / inv_sumw_den;                  // s ./= ws   (:315)
                mus[lane] = rok ? m : 0.0;
                if (rok) { vec[row] = m; mm = fma(m, m, mm); }
            }
            __syncthreads();
        } else if (MODE == GA_DOTS) {
            __syncthreads();
            if (warp == 0) mus[lane] = rok ? vec[row] : 0.0;
            __syncthreads();
        }

        // ---- phase B: t[n] += sum_i X[i,n] * (mu_i | q_i | X[i,n]) -------------------------------------------------
        for (int ch = 0; ch < nchunks; ++ch) {
            const int64_t cbase = (int64_t)ch * GA_NC;
            if (!(MODE == GA_PASS && nchunks == 1)) {
                __syncthreads();
#pragma unroll 8
                for (int k = 0; k < 32; ++k) {
                    const int c = warp + 8 * k;
                    double x = 0.0;
                    if (rok && cbase + c < N) x = __ldg(X + (cbase + c) * ld + row);
                    Xs[c * GA_R + lane] = x;
                }
                __syncthreads();
            }
            double acc = 0.0;
            const int n = tid;
#pragma unroll 8
            for (int i = 0; i < GA_R; ++i) {
                const int ii = (i + n) & (GA_R - 1);
                const double x = Xs[n * GA_R + ii];
                if (MODE == GA_NORMS) acc = fma(x, x, acc);
                else acc = fma(x, mus[ii], acc);
            }
            if (nchunks == 1) tacc += acc;
            else if (cbase + n < N) atomicAdd(t + cbase + n, acc);
        }
        __syncthreads();
    }

    if (nchunks == 1 && tid < N) atomicAdd(t + tid, tacc);
    if (MODE == GA_PASS) {
        // only warp 0 holds mm
        if (warp == 0) {
            mm = warp_sum(mm);
            if (lane == 0) atomicAdd(t + N, mm);
        }
    }
    (void)mmred;
}

__global__ void __launch_bounds__(256)
ga_signs_kernel(const double* __restrict__ t, const double* __restrict__ norms2, int64_t N,
                double* __restrict__ s, double* __restrict__ sumw) {
    __shared__ double red[8];
    double local = 0.0;
    for (int64_t n = threadIdx.x; n < N; n += blockDim.x) {
        const double tn = t[n];
        const double nrm = sqrt(norms2[n]);                                   // Xnorms[n]   (:265)
        double sg = (tn > 0.0) ? 1.0 : ((tn < 0.0) ? -1.0 : 0.0);            // sign(U[:,n]'q)   (:292)
        if (!(nrm > 0.0)) sg = 0.0;
        s[n] = sg;
        local += sg * nrm;                                                    // ws += w[n]   (:312)
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
        *sumw = v;
    }
}

__global__ void __launch_bounds__(256)
ga_update_kernel(const double* __restrict__ mu, const double* __restrict__ mm, int64_t d, double* __restrict__ q,
                 double* __restrict__ dq2) {
    const double nrm = sqrt(*mm);
    double local = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += (int64_t)gridDim.x * blockDim.x) {
        const double qn = mu[i] / nrm;                                        // q .= mu ./ norm(mu)   (:295)
        const double df = qn - q[i];
        local = fma(df, df, local);                                           // dq   (:296)
        q[i] = qn;                                                            // qold .= q   (:302)
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0 && local != 0.0) atomicAdd(dq2, local);
}

__global__ void __launch_bounds__(256)
vec_sumsq_kernel(const double* __restrict__ v, int64_t d, double* __restrict__ ss) {
    double local = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += (int64_t)gridDim.x * blockDim.x)
        local = fma(v[i], v[i], local);
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) atomicAdd(ss, local);
}

__global__ void __launch_bounds__(256)
vec_scale_rsqrt_kernel(const double* __restrict__ v, const double* __restrict__ ss, int64_t d,
                       double* __restrict__ out) {
    const double nrm = sqrt(*ss);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = v[i] / nrm;                                                  // q ./= norm(q)   (:287)
}

__global__ void __launch_bounds__(256)
ga_deflate_kernel(double* __restrict__ X, int64_t d, int64_t N, int64_t ld, const double* __restrict__ q,
                  const double* __restrict__ xs) {
    const int64_t total = d * N;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx % d, n = idx / d;
        const int64_t off = n * ld + i;
        X[off] = fma(-q[i], xs[n], X[off]);                                   // mul!(X, q, Xs1, -1, 1)   (:272)
    }
}

inline int vec_grid(int64_t total, int sm_count) {
    int64_t want = (total + 255) / 256;
    int64_t cap = (int64_t)sm_count * 8;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}


// ----------------------------------------------------------------------------------------------------------
// Robust entry-wise averages for rpca_ga (src/robustPCA.jl:323-333 entrywise_trimmed_mean, :349-357
// entrywise_median): per ROW j of the normalised data U[j,n] = X[j,n]/||x_n|| a sort over the N observations.
// A block stages 32 consecutive rows x N columns in shared memory with coalesced loads (lanes along rows); every
// warp then sorts its rows one after the other with a bitonic network on (key, index) pairs -- the index breaks ties
// like the reference's stable sortperm.
//   kind 1: keys = U[j,:]; s[j] = sum_{n in I} w_n U[j,n] / sum_{n in I} w_n, I = sorted[lo:hi)
//   kind 2: keys = w .* U[j,:]; m = sorted[N/2 - 1]; s[j] = sign(w_m) U[j,m]
// w_n = sgn_n * sqrt(n2_n) (sgn from the previous sweep's dot products, :291-293).
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ga_robust_kernel(const double* __restrict__ X, int64_t d, int N, int NP2, int RT, int64_t ld,
                 const double* __restrict__ sgn, const double* __restrict__ n2, int kind, int lo, int hi,
                 double* __restrict__ out) {
    extern __shared__ double smr[];
    const int TS = RT + 1;                                // RT rows per tile (32, or 8 for long observation axes)
    double* tile = smr;                                   // [N][RT + 1]  (padded: rows along the fast index)
    double* wv = tile + (size_t)N * TS;                   // [N] weights
    double* inv = wv + N;                                 // [N] 1 / ||x_n||
    double* keys = inv + N;                               // [8 warps][NP2]
    int* idxs = reinterpret_cast<int*>(keys + 8 * NP2);   // [8 warps][NP2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int n = tid; n < N; n += 256) {
        const double nr = sqrt(n2[n]);
        wv[n] = sgn[n] * nr;
        inv[n] = nr;
    }
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    for (int64_t r0 = (int64_t)blockIdx.x * RT; r0 < d; r0 += (int64_t)gridDim.x * RT) {
        __syncthreads();
        for (int idx = tid; idx < N * RT; idx += 256) {
            const int n = idx / RT, r = idx % RT;
            const int64_t row = r0 + r;
            tile[n * TS + r] = row < d ? __ldg(X + (int64_t)n * ld + row) : 0.0;
        }
        __syncthreads();
        double* kk = keys + warp * NP2;
        int* ii = idxs + warp * NP2;
        for (int rr = warp; rr < RT; rr += 8) {
            const int64_t row = r0 + rr;
            if (row >= d) break;
            for (int n = lane; n < NP2; n += 32) {
                double u = INF;
                if (n < N) {
                    u = tile[n * TS + rr] / inv[n];                              // U[j,n] = X[j,n] / Xnorms[n]   :266
                    if (kind == 2) u = wv[n] * u;
                }
                kk[n] = u;
                ii[n] = n;
            }
            __syncwarp();
            for (int k = 2; k <= NP2; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = lane; t < NP2 / 2; t += 32) {
                        const int a = 2 * t - (t & (j - 1));                      // lower index of the pair (bit j clear)
                        const int b = a + j;
                        const bool up = (a & k) == 0;
                        const double ka = kk[a], kb = kk[b];
                        const int ia = ii[a], ib = ii[b];
                        const bool gt = (ka > kb) || (ka == kb && ia > ib);
                        if (gt == up) { kk[a] = kb; kk[b] = ka; ii[a] = ib; ii[b] = ia; }
                    }
                    __syncwarp();
                }
            if (kind == 1) {
                double num = 0.0, den = 0.0;
                for (int p = lo + lane; p < hi; p += 32) {
                    const double w = wv[ii[p]];
                    num = fma(w, kk[p], num);
                    den += w;
                }
                num = warp_sum(num);
                den = warp_sum(den);
                if (lane == 0) out[row] = num / den;
            } else if (lane == 0) {
                const int m = ii[N / 2 - 1];
                const double w = wv[m];
                const double u = tile[m * TS + rr] / inv[m];
                out[row] = (w > 0.0 ? 1.0 : (w < 0.0 ? -1.0 : 0.0)) * u;
            }
            __syncwarp();
        }
    }
}

}  // namespace

cudaError_t launch_ga_sweep(GaMode mode, const double* X, int64_t d, int64_t N, int64_t ld, double* vec,
                            const double* s, const double* sumw, double* t, int sm_count, cudaStream_t st,
                            int64_t* launches) {
    const int ntiles = (int)((d + GA_R - 1) / GA_R);
    int grid = sm_count * 3;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    const size_t smem = (size_t)(GA_NC * GA_R + GA_NC + 8 * GA_R + GA_R) * sizeof(double);
    cudaError_t e;
#define TLSQ_GA_LAUNCH(MODE)                                                                          \
    do {                                                                                              \
        auto kern = ga_sweep_kernel<MODE>;                                                            \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
        if (e != cudaSuccess) return e;                                                               \
        kern<<<grid, 256, smem, st>>>(X, d, N, ld, vec, s, sumw, t, ntiles);                          \
    } while (0)
    switch (mode) {
        case GA_NORMS: TLSQ_GA_LAUNCH(GA_NORMS); break;
        case GA_DOTS:  TLSQ_GA_LAUNCH(GA_DOTS); break;
        default:       TLSQ_GA_LAUNCH(GA_PASS); break;
    }
#undef TLSQ_GA_LAUNCH
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_ga_signs(const double* t, const double* norms2, int64_t N, double* s, double* sumw,
                            cudaStream_t st, int64_t* launches) {
    ga_signs_kernel<<<1, 256, 0, st>>>(t, norms2, N, s, sumw);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_ga_update(const double* mu, const double* mm, int64_t d, double* q, double* dq2, int sm_count,
                             cudaStream_t st, int64_t* launches) {
    ga_update_kernel<<<vec_grid(d, sm_count), 256, 0, st>>>(mu, mm, d, q, dq2);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_vec_sumsq(const double* v, int64_t d, double* ss, int sm_count, cudaStream_t st,
                             int64_t* launches) {
    vec_sumsq_kernel<<<vec_grid(d, sm_count), 256, 0, st>>>(v, d, ss);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_vec_scale_rsqrt(const double* v, const double* ss, int64_t d, double* out, int sm_count,
                                   cudaStream_t st, int64_t* launches) {
    vec_scale_rsqrt_kernel<<<vec_grid(d, sm_count), 256, 0, st>>>(v, ss, d, out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_ga_deflate(double* X, int64_t d, int64_t N, int64_t ld, const double* q, const double* xs,
                              int sm_count, cudaStream_t st, int64_t* launches) {
    ga_deflate_kernel<<<vec_grid(d * N, sm_count), 256, 0, st>>>(X, d, N, ld, q, xs);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_ga_robust(const double* X, int64_t d, int64_t N, int64_t ld, const double* sgn, const double* n2,
                             int kind, double P, double* out, int sm_count, cudaStream_t st, int64_t* launches) {
    int np2 = 2;
    while (np2 < N) np2 <<= 1;
    int rt = 32;
    auto need = [&](int rows) {
        return ((size_t)N * (rows + 1) + 2 * (size_t)N + 8 * (size_t)np2) * sizeof(double) + 8 * (size_t)np2 * sizeof(int);
    };
    if (need(rt) > (size_t)220 * 1024) rt = 8;
    const size_t smem = need(rt);
    if (smem > (size_t)220 * 1024) return cudaErrorInvalidValue;          // N > ~1024 observations
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(ga_robust_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = smem;
    }
    const int lo = (int)floor(P * (double)N), hi = (int)floor((1.0 - P) * (double)N);      // :325
    int64_t blocks = (d + rt - 1) / rt;
    if (blocks > (int64_t)sm_count * 2) blocks = (int64_t)sm_count * 2;
    ga_robust_kernel<<<(unsigned)blocks, 256, smem, st>>>(X, d, (int)N, np2, rt, ld, sgn, n2, kind, lo, hi, out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace tlsq
