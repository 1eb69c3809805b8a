// eig_fast.cu -- fast path of the per-iteration n x n eigenproblem (n <= 256): the ALM step needs only the svp
// dominant eigenpairs of G = W'W plus the exact COUNT svp = #{sigma_i >= 1/mu} (src/robustPCA.jl:198, full-SVD
// semantics).  Diagonalising the ~n-svp "bulk" directions with a full Jacobi costs ~10 sweeps every iteration
// (measured: the bulk is not warm-startable, it changes completely between ALM iterations), so instead:
//
//  1. Warm-started block subspace iteration on a b = 32 column block:  X = G Q  (DMMA GEMM, atb_kernel), then a
//     one-sided Jacobi of the b columns of X inside ONE CTA (16 warps = 16 column pairs, warp-shuffle dot products,
//     rotations in registers, tournament exchange through shared memory) with Q rotated alongside.  This is a
//     Rayleigh-Ritz step on span(Q); theta_i = q_i'G q_i, residual r_i = ||G q_i - theta_i q_i||; the next basis is
//     the normalised X (one power step).  Stops when all wanted pairs (theta_i >= tau^2, plus the first unwanted
//     one) have r_i <= tol * theta_1.
//  2. Certificate of the count.  Ritz values are lower bounds (theta_i <= lambda_i), so at least svp eigenvalues are
//     >= tau^2.  For "at most": G_w = G - Q_w Theta_w Q_w' differs from G by a PSD rank-svp term, hence
//     lambda_{svp+1}(G) <= lambda_max(G_w) (Weyl), and lambda_max(G_w) <= ||G_w^(2^k)||_F^(1/2^k), evaluated by k
//     Frobenius-normalised squarings (DMMA, the same atb_kernel).  After k = 6 squarings the bound is within
//     n^(1/128) = 4.4 % of lambda_max.  If bound < tau^2 the count is proven; otherwise (or if step 1 does not
//     converge, e.g. when a wanted singular value sits at the edge of the bulk) a device flag routes the iteration
//     through the full Jacobi of eig.cu.  Every kernel looks at the flags itself -- there is no host round trip.
#include "kernels.h"

namespace tlsq {

namespace {

// ---------------------------------------------------------------------------------------------------
// C (na x nb) = (sa*A)' (sb*B),  A: K x na (lda), B: K x nb (ldb), column-major; 32 x 32 output tile per CTA,
// 8 warps = 2 (m) x 4 (n), warp tile 16 x 8 (2 DMMA accumulators); K staged in 64-row chunks, [col][68] padded.
// symmetric != 0: A == B, only tiles with ti <= tj are computed and mirrored; fro2 (optional) accumulates ||C||_F^2.
// run control: the kernel exits unless (flags[0] == want_conv) && (flags[1] == 0)   (flags may be null).
// ---------------------------------------------------------------------------------------------------
constexpr int AT = 32, AK = 64, AKS = AK + 4;

__global__ void __launch_bounds__(256)
atb_kernel(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb, int K, int na, int nb,
           const double* __restrict__ scale2, double* __restrict__ C, int ldc, int symmetric,
           double* __restrict__ fro2, const int* __restrict__ flags, int want_conv,
           const double* __restrict__ cert_f2 = nullptr, int cert_j = 0, double cert_thr = 0.0,
           const double* __restrict__ cert_theta = nullptr) {
    if (flags && !(flags[0] == want_conv && flags[1] == 0)) return;
    if (cert_f2) {
        // squaring j of the count certificate: stop as soon as the bound from the chain so far already proves it
        // (lambda_max <= f_0 u_0, u_j = 1, u_i = sqrt(f_{i+1} u_{i+1}); every CTA evaluates the same finished numbers)
        double u = 1.0;
        for (int j = cert_j - 1; j >= 0; --j) u = sqrt(sqrt(cert_f2[j + 1]) * u);
        const double bound = sqrt(cert_f2[0]) * u;
        const double thr = cert_theta ? cert_theta[0] : cert_thr;
        if (bound < thr * (1.0 - 1e-10)) return;
    }
    const int ti = blockIdx.x, tj = blockIdx.y;
    if (symmetric && ti > tj) return;
    __shared__ double As[AT * AKS];
    __shared__ double Bs[AT * AKS];
    __shared__ double red[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 1, wn = warp >> 1;
    double sc = 1.0;
    if (scale2) {                                   // operands are scaled by 1/sqrt(*scale2) each (Frobenius normalisation)
        const double s2 = *scale2;
        sc = s2 > 0.0 ? 1.0 / s2 : 0.0;             // product of the two operand scales
    }
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    const int lr = tid & 63, lc0 = tid >> 6;        // loader: 64 rows x 4 column groups
    for (int k0 = 0; k0 < K; k0 += AK) {
        const bool rok = k0 + lr < K;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c = lc0 + 4 * q;
            const int ca = ti * AT + c, cb = tj * AT + c;
            As[c * AKS + lr] = (rok && ca < na) ? A[(int64_t)ca * lda + k0 + lr] : 0.0;
            Bs[c * AKS + lr] = (rok && cb < nb) ? B[(int64_t)cb * ldb + k0 + lr] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < AK; kk += 4) {
            const double a0 = As[(16 * wm + g) * AKS + kk + t];
            const double a1 = As[(16 * wm + 8 + g) * AKS + kk + t];
            const double b0 = Bs[(8 * wn + g) * AKS + kk + t];
            dmma884(acc[0][0], acc[0][1], a0, b0);
            dmma884(acc[1][0], acc[1][1], a1, b0);
        }
        __syncthreads();
    }
    double f2 = 0.0;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = ti * AT + 16 * wm + 8 * mi + g;
            const int j = tj * AT + 8 * wn + 2 * t + e;
            if (i < na && j < nb) {
                const double v = acc[mi][e] * sc;
                C[(int64_t)j * ldc + i] = v;
                if (symmetric && ti != tj) C[(int64_t)i * ldc + j] = v;
                f2 = fma(v, v, f2);
            }
        }
    if (fro2) {
        if (symmetric && ti != tj) f2 *= 2.0;
        f2 = warp_sum(f2);
        if (lane == 0) red[warp] = f2;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w];
            atomicAdd(fro2, s);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// one Rayleigh-Ritz step of the block subspace iteration, single CTA, 16 warps <-> 16 column pairs (b = 32)
// ---------------------------------------------------------------------------------------------------
constexpr int SIB = 32;          // block width

template <int E>
__device__ __forceinline__ bool rotate_pair(double (&xp)[E], double (&xq)[E], double (&vp)[E], double (&vq)[E],
                                            double tol, double wanted2) {
    double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        alpha = fma(xp[e], xp[e], alpha);
        beta = fma(xq[e], xq[e], beta);
        gamma = fma(xp[e], xq[e], gamma);
    }
    alpha = warp_sum(alpha);
    beta = warp_sum(beta);
    gamma = warp_sum(gamma);
    double c, s;
    const double tl = (alpha < wanted2 && beta < wanted2) ? 1.0e-6 : tol;
    if (!jacobi_cs(alpha, beta, gamma, tl, c, s)) return false;        // already orthogonal (or a null column)
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const double a = xp[e], b = xq[e];
        xp[e] = fma(c, a, -s * b);
        xq[e] = fma(s, a, c * b);
        const double va = vp[e], vb = vq[e];
        vp[e] = fma(c, va, -s * vb);
        vq[e] = fma(s, va, c * vb);
    }
    return true;
}

// flags: [0] conv, [1] need_full, [2] svp, [3] SI steps used, [4] certified, [5] jacobi sweeps of the last SI step
template <int E, int BW>
__global__ void __launch_bounds__(16 * BW, 1)
si_jacobi_kernel(const double* __restrict__ X, double* __restrict__ Q, int n, double tau2, int min_wanted,
                 double tol_jac, double tol_res, int last_step, double* __restrict__ theta_out,
                 double* __restrict__ Qout, int* __restrict__ flags) {
    if (flags[0] == 1 || flags[1] == 1) return;
    constexpr int LEN = 32 * E;
    extern __shared__ double sm[];                 // SIB slots x [X part | Q part]
    __shared__ double th[BW], rr[BW], nx[BW], ths[BW];
    __shared__ int order[BW];
    __shared__ int s_svp, s_conv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;      // warp = seat 0..15
    constexpr int m = BW / 2;
    double xp[E], xq[E], vp[E], vq[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = lane + 32 * e;
        const bool ok = i < n;
        xp[e] = ok ? X[(int64_t)(2 * warp) * n + i] : 0.0;
        xq[e] = ok ? X[(int64_t)(2 * warp + 1) * n + i] : 0.0;
        vp[e] = ok ? Q[(int64_t)(2 * warp) * n + i] : 0.0;
        vq[e] = ok ? Q[(int64_t)(2 * warp + 1) * n + i] : 0.0;
    }
    // pairs of two UNWANTED columns (|x|^2 < tau^4, i.e. Ritz value below tau^2) only need to stay well conditioned
    const double wanted2 = tau2 < 1.0e150 ? 0.25 * tau2 * tau2 : 0.0;      // opnorm mode: tight everywhere
    int sweeps = 0;
    for (; sweeps < 30; ++sweeps) {
        int rotated = 0;
        for (int round = 0; round < 2 * m - 1; ++round) {
            rotated |= rotate_pair<E>(xp, xq, vp, vq, tol_jac, wanted2) ? 1 : 0;
            const int s = warp;
            int ts, tw, bs, bw;
            if (s == 0) { ts = 0; tw = 0; } else if (s == m - 1) { ts = m - 1; tw = 1; } else { ts = s + 1; tw = 0; }
            if (s == 0) { bs = 1; bw = 0; } else { bs = s - 1; bw = 1; }
            double* dt = sm + (size_t)(2 * ts + tw) * 2 * LEN;
            double* db = sm + (size_t)(2 * bs + bw) * 2 * LEN;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                dt[lane + 32 * e] = xp[e];
                dt[LEN + lane + 32 * e] = vp[e];
                db[lane + 32 * e] = xq[e];
                db[LEN + lane + 32 * e] = vq[e];
            }
            __syncthreads();
            const double* pt = sm + (size_t)(2 * s) * 2 * LEN;
            const double* pb = sm + (size_t)(2 * s + 1) * 2 * LEN;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                xp[e] = pt[lane + 32 * e];
                vp[e] = pt[LEN + lane + 32 * e];
                xq[e] = pb[lane + 32 * e];
                vq[e] = pb[LEN + lane + 32 * e];
            }
            __syncthreads();
        }
        if (!__syncthreads_or(rotated)) { ++sweeps; break; }
    }
    // Ritz values / residuals of the two columns of this warp (columns stay in their slots: index 2*warp + which)
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        double dot = 0.0, xx = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const double x = which ? xq[e] : xp[e];
            const double v = which ? vq[e] : vp[e];
            dot = fma(v, x, dot);
            xx = fma(x, x, xx);
        }
        dot = warp_sum(dot);
        xx = warp_sum(xx);
        double r2 = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const double x = which ? xq[e] : xp[e];
            const double v = which ? vq[e] : vp[e];
            const double df = x - dot * v;
            r2 = fma(df, df, r2);
        }
        r2 = warp_sum(r2);
        if (lane == 0) {
            th[2 * warp + which] = dot;
            rr[2 * warp + which] = sqrt(r2);
            nx[2 * warp + which] = sqrt(xx);
        }
    }
    __syncthreads();
    if (tid < BW) {
        const double lj = th[tid];
        int rank = 0;
        for (int k = 0; k < BW; ++k) rank += (th[k] > lj) || (th[k] == lj && k < tid);
        order[rank] = tid;
        ths[rank] = lj;
    }
    __syncthreads();
    if (tid == 0) {
        int svp = 0;
        for (int i = 0; i < BW; ++i) svp += (ths[i] >= tau2) ? 1 : 0;
        // only the WANTED pairs must be converged: Ritz values are lower bounds, so the count is >= svp whatever the
        // state of the unwanted ones (they track the flat bulk and converge arbitrarily slowly); "<= svp" is proven
        // by the squaring certificate.
        if (svp < min_wanted) svp = min_wanted;       // top-k mode (opnorm: the dominant pair regardless of tau)
        int conv = (svp <= BW - 2) ? 1 : 0;
        const int need = svp;
        const double lim = tol_res * fabs(ths[0]);
        for (int i = 0; i < need; ++i) conv &= (rr[order[i]] <= lim) ? 1 : 0;
        s_svp = svp;
        s_conv = conv;
        flags[3] += 1;
        flags[5] = sweeps;
        if (conv) { flags[0] = 1; flags[2] = svp; }
        else if (svp > BW - 2 || last_step) flags[1] = 1;
    }
    __syncthreads();
    const int conv = s_conv;
    // outputs: sorted Ritz vectors (used when converged) / normalised power step (next basis)
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        const int col = 2 * warp + which;
        int rank = 0;
        for (int k = 0; k < BW; ++k) rank = (order[k] == col) ? k : rank;
        const double nrm = nx[col];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = lane + 32 * e;
            if (i < n) {
                const double x = which ? xq[e] : xp[e];
                const double v = which ? vq[e] : vp[e];
                if (conv) Qout[(int64_t)rank * n + i] = v;
                else Q[(int64_t)col * n + i] = nrm > 0.0 ? x / nrm : v;
            }
        }
        if (conv && lane == 0) theta_out[rank] = th[col];
    }
}

// Gd = G - sum_{c < svp} theta_c q_c q_c'  (+ ||Gd||_F^2), runs only when the subspace iteration converged
__global__ void __launch_bounds__(256)
deflate_kernel(const double* __restrict__ G, int n, const double* __restrict__ Qs, const double* __restrict__ theta,
               const int* __restrict__ flags, double* __restrict__ Gd, double* __restrict__ fro2) {
    if (!(flags[0] == 1 && flags[1] == 0)) return;
    const int svp = flags[2];
    double f2 = 0.0;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (int64_t)n * n) {
        const int i = (int)(idx % n), j = (int)(idx / n);
        double v = G[idx];
        for (int c = 0; c < svp; ++c) v = fma(-theta[c] * Qs[(int64_t)c * n + i], Qs[(int64_t)c * n + j], v);
        Gd[idx] = v;
        f2 = v * v;
    }
    f2 = warp_sum(f2);
    if ((threadIdx.x & 31) == 0 && f2 != 0.0) atomicAdd(fro2, f2);
}

// certificate + outputs of the fast path (single CTA)
__global__ void __launch_bounds__(256)
fast_finish_kernel(int n, int nsq, const double* __restrict__ f2, double tau2, int top1, double tau, int nukeA,
                   const double* __restrict__ theta, const double* __restrict__ Qs, double* __restrict__ Qb,
                   double* __restrict__ Vs, double* __restrict__ lam, double* __restrict__ sigma,
                   double* __restrict__ fvec, int* __restrict__ svp_out, int* __restrict__ flags, int bw) {
    __shared__ int ok;
    if (threadIdx.x == 0) {
        int good = (flags[0] == 1 && flags[1] == 0) ? 1 : 0;
        if (good) {
            // lambda_max(Gd) <= f_0 * u_0,  u_k = 1, u_j = sqrt(f_{j+1} u_{j+1})
            // (the squaring chain stops early once it has proven the count: use the part that was computed)
            int ne = 0;
            while (ne < nsq && f2[ne + 1] > 0.0) ++ne;
            double u = 1.0;
            for (int j = ne - 1; j >= 0; --j) u = sqrt(sqrt(f2[j + 1]) * u);
            const double bound = sqrt(f2[0]) * u;
            // ALM step: nothing left above tau^2.  opnorm mode: nothing left above theta_1, i.e. theta_1 IS lambda_max.
            const double thr = top1 ? theta[0] : tau2;
            good = (bound < thr * (1.0 - 1e-10)) ? 1 : 0;
            if (!(bound == bound)) good = 0;
        }
        if (!good) flags[1] = 1;
        flags[4] = good;
        ok = good;
    }
    __syncthreads();
    if (!ok) return;
    const int svp = flags[2];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double l = i < bw ? theta[i] : 0.0;
        const double sg = l > 0.0 ? sqrt(l) : 0.0;
        lam[i] = l;
        sigma[i] = sg;
        const bool keep = i < svp;
        fvec[i] = keep ? (nukeA ? (sg - tau) / sg : 1.0) : 0.0;
    }
    for (int idx = threadIdx.x; idx < n * bw; idx += blockDim.x) {
        const double v = Qs[idx];
        Vs[idx] = v;
        Qb[idx] = v;
    }
    if (threadIdx.x == 0) *svp_out = svp;
}

// Qb = Vs[:, 0:32]  after a full Jacobi (only when the fallback ran)
__global__ void copy_block_kernel(const double* __restrict__ Vs, int n, double* __restrict__ Qb,
                                  const int* __restrict__ flags) {
    if (flags && flags[1] == 0) return;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * SIB; idx += gridDim.x * blockDim.x)
        Qb[idx] = Vs[idx];
}

// ||G||_F^2 of an n x n matrix
__global__ void __launch_bounds__(256)
fro2_kernel(const double* __restrict__ G, int64_t nn, double* __restrict__ out) {
    double f2 = 0.0;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nn; idx += (int64_t)gridDim.x * blockDim.x)
        f2 = fma(G[idx], G[idx], f2);
    f2 = warp_sum(f2);
    if ((threadIdx.x & 31) == 0 && f2 != 0.0) atomicAdd(out, f2);
}

// bounds[0] <= lambda_max(G) <= bounds[1] from the Frobenius norms of the squaring chain:
//   upper: u_k = 1,          u_j = sqrt(f_{j+1} u_{j+1});   lower: l_k = 1/sqrt(n), l_j = sqrt(f_{j+1} l_{j+1})
__global__ void lmax_bounds_kernel(const double* __restrict__ f2, int nsq, int n, double* __restrict__ bounds) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double u = 1.0, l = 1.0 / sqrt((double)n);
    for (int j = nsq - 1; j >= 0; --j) {
        const double f = sqrt(f2[j + 1]);
        u = sqrt(f * u);
        l = sqrt(f * l);
    }
    const double f0 = sqrt(f2[0]);
    bounds[0] = f0 * l;
    bounds[1] = f0 * u;
}

__global__ void init_block_kernel(double* __restrict__ Qb, int n) {      // Qb = [e_1 ... e_32]
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * SIB; idx += gridDim.x * blockDim.x)
        Qb[idx] = (idx % n == idx / n) ? 1.0 : 0.0;
}

template <int E, int BW>
cudaError_t launch_si(const double* X, double* Q, int n, double tau2, int min_wanted, double tol_jac, double tol_res,
                      int last, double* theta, double* Qout, int* flags, cudaStream_t st) {
    const size_t smem = (size_t)BW * 2 * 32 * E * sizeof(double);
    auto kern = si_jacobi_kernel<E, BW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<1, 16 * BW, smem, st>>>(X, Q, n, tau2, min_wanted, tol_jac, tol_res, last, theta, Qout, flags);
    return cudaGetLastError();
}

}  // namespace

size_t eig_fast_work_doubles(int n) {
    // Qb, Qwork, Qout, X (n x 32 each) + theta (32) + Ca, Cb (n x n) + f2 (16) + flags (8 ints -> 8 doubles)
    return (size_t)4 * n * SIB + SIB + (size_t)2 * n * n + 16 + 8 + 16;
}

bool eig_fast_supported(int n) { return n > 64 && n <= 512; }
int eig_fast_max_block(int n) { return n <= 256 ? 32 : 16; }   // shared memory: block x 2 x n doubles must fit

EigFastWork eig_fast_carve(double* base, int n) {
    EigFastWork w;
    w.Qb = base; base += (size_t)n * SIB;
    w.Qwork = base; base += (size_t)n * SIB;
    w.Qout = base; base += (size_t)n * SIB;
    w.X = base; base += (size_t)n * SIB;
    w.theta = base; base += SIB;
    w.Ca = base; base += (size_t)n * n;
    w.Cb = base; base += (size_t)n * n;
    w.f2 = base; base += 16;
    w.flags = reinterpret_cast<int*>(base);
    return w;
}

cudaError_t launch_lmax_bounds(const double* G, int n, double* Ca, double* Cb, double* f2, double* bounds,
                               cudaStream_t st, int64_t* launches) {
    constexpr int NSQ = 10;      // bracket ratio n^(1/2^(NSQ+1)): 0.27 % at n = 256
    cudaError_t e;
    if ((e = cudaMemsetAsync(f2, 0, 16 * sizeof(double), st)) != cudaSuccess) return e;
    fro2_kernel<<<32, 256, 0, st>>>(G, (int64_t)n * n, f2);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const dim3 gs((n + AT - 1) / AT, (n + AT - 1) / AT);
    const double* cin = G;
    double* cout = Ca;
    for (int j = 0; j < NSQ; ++j) {
        atb_kernel<<<gs, 256, 0, st>>>(cin, n, cin, n, n, n, n, f2 + j, cout, n, 1, f2 + j + 1, nullptr, 0);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        cin = cout;
        cout = (cout == Ca) ? Cb : Ca;
    }
    lmax_bounds_kernel<<<1, 32, 0, st>>>(f2, NSQ, n, bounds);
    if (launches) *launches += NSQ + 2;
    return cudaGetLastError();
}

cudaError_t launch_copy_block(const double* Vs, int n, double* Qb, const int* flags, cudaStream_t st,
                              int64_t* launches) {
    copy_block_kernel<<<8, 256, 0, st>>>(Vs, n, Qb, flags);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_init_block(double* Qb, int n, cudaStream_t st, int64_t* launches) {
    init_block_kernel<<<8, 256, 0, st>>>(Qb, n);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_eig_fast(const double* G, int n, double tau, int nukeA, EigFastWork w, double* lam, double* Vs,
                            double* sigma, double* fvec, int* svp, cudaStream_t st, int64_t* launches, int top1,
                            int bw, int max_steps) {
    if (bw != 16) bw = SIB;
    if (n > 256) bw = 16;
    // subspace-iteration steps attempted before falling back (launches after convergence exit at once, but each still
    // costs a launch slot: the caller passes last iteration's step count + a margin)
    const int NSI = max_steps < 2 ? 2 : (max_steps > 16 ? 16 : max_steps);
    constexpr int NSQ = 6;       // squarings of the certificate: bound within n^(1/128) of lambda_max
    cudaError_t e;
    const double tau2 = top1 ? 1.0e300 : tau * tau;
    const int min_wanted = top1 ? 1 : 0;
    const double tol_jac = 1.0e-15 * sqrt((double)n);
    const double tol_res = 4.0e-14;
    if ((e = cudaMemsetAsync(w.f2, 0, (16 + 8) * sizeof(double), st)) != cudaSuccess) return e;   // f2 + flags
    if ((e = cudaMemcpyAsync(w.Qwork, w.Qb, (size_t)n * bw * 8, cudaMemcpyDeviceToDevice, st)) != cudaSuccess)
        return e;
    const dim3 gx((n + AT - 1) / AT, 1);
    for (int it = 0; it < NSI; ++it) {
        // X = G' Q = G Q   (skipped once converged / failed)
        atb_kernel<<<gx, 256, 0, st>>>(G, n, w.Qwork, n, n, n, bw, nullptr, w.X, n, 0, nullptr, w.flags, 0);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        const int last = (it == NSI - 1) ? 1 : 0;
        if (bw == 16) {
            if (n <= 128) e = launch_si<4, 16>(w.X, w.Qwork, n, tau2, min_wanted, tol_jac, tol_res, last, w.theta, w.Qout, w.flags, st);
            else if (n <= 256) e = launch_si<8, 16>(w.X, w.Qwork, n, tau2, min_wanted, tol_jac, tol_res, last, w.theta, w.Qout, w.flags, st);
            else e = launch_si<16, 16>(w.X, w.Qwork, n, tau2, min_wanted, tol_jac, tol_res, last, w.theta, w.Qout, w.flags, st);
        } else {
            if (n <= 128) e = launch_si<4, 32>(w.X, w.Qwork, n, tau2, min_wanted, tol_jac, tol_res, last, w.theta, w.Qout, w.flags, st);
            else e = launch_si<8, 32>(w.X, w.Qwork, n, tau2, min_wanted, tol_jac, tol_res, last, w.theta, w.Qout, w.flags, st);
        }
        if (e != cudaSuccess) return e;
    }
    deflate_kernel<<<(unsigned)(((int64_t)n * n + 255) / 256), 256, 0, st>>>(G, n, w.Qout, w.theta, w.flags, w.Ca,
                                                                            w.f2);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const dim3 gs((n + AT - 1) / AT, (n + AT - 1) / AT);
    double* cin = w.Ca;
    double* cout = w.Cb;
    for (int j = 0; j < NSQ; ++j) {
        // C_{j+1} = (C_j / f_j)' (C_j / f_j),  f2[j+1] = ||C_{j+1}||_F^2
        atb_kernel<<<gs, 256, 0, st>>>(cin, n, cin, n, n, n, n, w.f2 + j, cout, n, 1, w.f2 + j + 1, w.flags, 1,
                                       w.f2, j, tau2, top1 ? w.theta : nullptr);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        double* tmp = cin; cin = cout; cout = tmp;
    }
    fast_finish_kernel<<<1, 256, 0, st>>>(n, NSQ, w.f2, tau2, top1, tau, nukeA, w.theta, w.Qout, w.Qb, Vs, lam, sigma, fvec,
                                          svp, w.flags, bw);
    if (launches) *launches += 2 * NSI + 1 + NSQ + 1;
    return cudaGetLastError();
}

}  // namespace tlsq
