// elementwise.cu -- streaming helper kernels: max|D| (src/robustPCA.jl:178), Y = D/dual & A = 0 (:174-181),
// E recomputation for the returned value (:188-191,238), transpose (M < N inputs), hankel / unhankel
// (:76-92, :28-39, :53-68).
#include "kernels.h"

namespace tlsq {

namespace {

// max |x_i| over a contiguous array (dense D with ld == M, or the signal span of an implicit Hankel matrix: every
// sample y[0 .. (M-1) lag + N - 1] appears in H because lag <= N, src/robustPCA.jl:80)
__global__ void __launch_bounds__(256)
maxabs_linear_kernel(const double* __restrict__ x, int64_t total, double* __restrict__ out) {
    double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; idx + 3 * stride < total; idx += 4 * stride) {
        m0 = fmax(m0, fabs(__ldg(x + idx)));
        m1 = fmax(m1, fabs(__ldg(x + idx + stride)));
        m2 = fmax(m2, fabs(__ldg(x + idx + 2 * stride)));
        m3 = fmax(m3, fabs(__ldg(x + idx + 3 * stride)));
    }
    for (; idx < total; idx += stride) m0 = fmax(m0, fabs(__ldg(x + idx)));
    double m = warp_max(fmax(fmax(m0, m1), fmax(m2, m3)));
    // non-negative doubles order like their bit patterns
    if ((threadIdx.x & 31) == 0)
        atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(m));
}

// dense D with a leading dimension > M: thread <-> row, loop over the columns (no 64-bit divisions)
__global__ void __launch_bounds__(256)
maxabs_kernel(const MatSrc D, int64_t M, int64_t N, double* __restrict__ out) {
    double m = 0.0;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < M; row += (int64_t)gridDim.x * blockDim.x)
        for (int64_t col = 0; col < N; ++col) m = fmax(m, fabs(__ldg(D.p + col * D.ld + row)));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0)
        atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(m));
}

// Y = D / dual (:181), A = 0, and (optionally) the first SVT input W_1; thread <-> row, coalesced column sweeps
template <bool HANKEL>
__global__ void __launch_bounds__(256)
init_ya_kernel(const MatSrc D, int64_t M, int64_t N, double dual, double* __restrict__ Y, double* __restrict__ A,
               double* __restrict__ W, double im, double eps, int nonnegE, int64_t ldy) {
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < M; row += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll 4
        for (int64_t col = 0; col < N; ++col) {
            const double d = src_at<HANKEL>(D, row, col);
            const double y = __ddiv_rn(d, dual);                         // Y ./= dual_norm   (:181)
            Y[col * ldy + row] = y;
            if (A) A[col * M + row] = 0.0;
            if (W) {
                double e, w;
                alm_ew(d, 0.0, y, im, eps, nonnegE, e, w);               // first SVT input   (:188-192)
                W[col * M + row] = w;
            }
        }
    }
}

template <bool HANKEL>
__global__ void __launch_bounds__(256)
compute_e_kernel(const MatSrc D, int64_t M, int64_t N, const double* __restrict__ A, const double* __restrict__ Y,
                 double im, double eps, int nonnegE, double* __restrict__ E) {
    const int64_t total = M * N;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx % M, col = idx / M;
        double e, w;
        alm_ew(src_at<HANKEL>(D, row, col), A[idx], Y[idx], im, eps, nonnegE, e, w);
        E[idx] = e;
    }
}

__global__ void __launch_bounds__(256)
transpose_kernel(const double* __restrict__ in, int64_t M, int64_t N, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t i = i0 + tx, j = j0 + ty + 8 * r;
        if (i < M && j < N) tile[ty + 8 * r][tx] = in[j * M + i];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t j = j0 + tx, i = i0 + ty + 8 * r;
        if (i < M && j < N) out[i * N + j] = tile[tx][ty + 8 * r];
    }
}

__global__ void __launch_bounds__(256)
hankel_kernel(const double* __restrict__ x, int64_t K, int64_t L, int64_t lag, double* __restrict__ H) {
    const int64_t total = K * L;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = idx % K, l = idx / K;
        H[idx] = x[k * lag + l];                                     // X[k, l] = x[(k-1)lag + l]   (:87-88)
    }
}

// y[t] = mean over {(k,l): k*lag + l == t} of A[k,l]; entries are visited for increasing l (column-major order of
// the reference's accumulation, :62-65); for lag == 1 this is the anti-diagonal mean (:28-39).
__global__ void __launch_bounds__(256)
unhankel_kernel(const double* __restrict__ A, int64_t K, int64_t L, int64_t lag, int64_t Ns,
                double* __restrict__ y) {
    for (int64_t tt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tt < Ns;
         tt += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        int64_t cnt = 0;
        for (int64_t l = 0; l < L; ++l) {
            const int64_t rem = tt - l;
            if (rem < 0) break;
            if (rem % lag) continue;
            const int64_t k = rem / lag;
            if (k >= K) continue;
            s += A[l * K + k];
            ++cnt;
        }
        y[tt] = s / (double)(cnt > 0 ? cnt : 1);                     // y ./= max.(counts, 1)   (:66)
    }
}

// sharded unhankel: this rank owns Hankel rows [r0, r0 + Kl); sums/counts of the anti-diagonals it touches
__global__ void __launch_bounds__(256)
unhankel_partial_kernel(const double* __restrict__ A, int64_t r0, int64_t Kl, int64_t L, int64_t lag, int64_t Ns,
                        double* __restrict__ sum, double* __restrict__ cnt) {
    const int64_t t_lo = r0 * lag, t_hi = (r0 + Kl - 1) * lag + L;          // [t_lo, t_hi)
    for (int64_t tt = t_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tt < t_hi && tt < Ns;
         tt += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        int64_t c = 0;
        for (int64_t l = 0; l < L; ++l) {
            const int64_t rem = tt - l;
            if (rem < 0) break;
            if (rem % lag) continue;
            const int64_t k = rem / lag - r0;
            if (k < 0 || k >= Kl) continue;
            s += A[l * Kl + k];
            ++c;
        }
        sum[tt] = s;
        cnt[tt] = (double)c;
    }
}

__global__ void __launch_bounds__(256)
unhankel_divide_kernel(const double* __restrict__ sum, const double* __restrict__ cnt, int64_t Ns, double* __restrict__ y) {
    for (int64_t tt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tt < Ns; tt += (int64_t)gridDim.x * blockDim.x) {
        const double c = cnt[tt];
        y[tt] = sum[tt] / (c > 0.0 ? c : 1.0);                                 // y ./= max.(counts, 1)   (:66)
    }
}

// Anti-diagonal sums of A = clamp(T V') over the local Hankel rows [r0, r0 + Kl), lag 1, WITHOUT materialising A
// (unhankel of src/robustPCA.jl:28-39 applied to the factored iterate):
//   sum[k] = sum_{j} clamp(T[k - j - r0, :] . V[j, :]),   0 <= k - j - r0 < Kl
// A block owns 256 consecutive samples k; the 256 + n - 1 rows of T it needs are staged in shared memory.
__global__ void __launch_bounds__(256)
unhankel_factors_kernel(const double* __restrict__ T, int64_t ldt, const double* __restrict__ V, int rp, int nonnegA,
                        int64_t r0, int64_t Kl, int n, int64_t Ns, double* __restrict__ sum) {
    extern __shared__ double sm[];
    const int span = 256 + n - 1;
    double* Vsm = sm;                        // [n][rp]
    double* Tsm = sm + (size_t)n * rp;       // [rp][span]
    for (int idx = threadIdx.x; idx < n * rp; idx += 256) {
        const int j = idx % n, c = idx / n;
        Vsm[j * rp + c] = __ldg(V + (int64_t)c * n + j);
    }
    const int64_t nblk = (Ns + 255) / 256;
    for (int64_t b = blockIdx.x; b < nblk; b += gridDim.x) {
        const int64_t kb = b * 256;
        const int64_t ilo = kb - r0 - (n - 1);
        __syncthreads();
        for (int idx = threadIdx.x; idx < rp * span; idx += 256) {
            const int q = idx % span, c = idx / span;
            const int64_t i = ilo + q;
            Tsm[c * span + q] = (i >= 0 && i < Kl) ? __ldg(T + (int64_t)c * ldt + i) : 0.0;
        }
        __syncthreads();
        const int64_t k = kb + threadIdx.x;
        if (k < Ns) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) {
                const int q = (int)threadIdx.x + (n - 1) - j;
                const int64_t i = ilo + q;
                if (i < 0 || i >= Kl) continue;
                double av = 0.0;
                for (int c = 0; c < rp; ++c) av = fma(Tsm[c * span + q], Vsm[j * rp + c], av);
                if (nonnegA) av = (__double_as_longlong(av) > 0) ? av : 0.0;
                s += av;
            }
            sum[k] = s;
        }
    }
}

// y[k] = sum[k] / #{(i, j): i + j = k, 0 <= i < K, 0 <= j < n}   (lag 1; the count is known in closed form)
__global__ void __launch_bounds__(256)
unhankel_divide_count_kernel(const double* __restrict__ sum, int64_t K, int64_t n, int64_t Ns, double* __restrict__ y) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < Ns; k += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = k;
        if (n - 1 < c) c = n - 1;
        if (K - 1 < c) c = K - 1;
        if (K + n - 2 - k < c) c = K + n - 2 - k;
        c = c < 0 ? 1 : c + 1;
        y[k] = sum[k] / (double)c;
    }
}

inline int stream_grid(int64_t total, int sm_count) {
    int64_t want = (total + 255) / 256;
    int64_t cap = (int64_t)sm_count * 8;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

}  // namespace

cudaError_t launch_maxabs(const MatSrc& D, bool hankel, int64_t M, int64_t N, double* out, int sm_count,
                          cudaStream_t st, int64_t* launches) {
    if (hankel || D.ld == M) {
        const int64_t total = hankel ? (M - 1) * D.ld + N : M * N;
        maxabs_linear_kernel<<<stream_grid(total / 4 + 1, sm_count), 256, 0, st>>>(D.p, total, out);
    } else {
        maxabs_kernel<<<stream_grid(M, sm_count), 256, 0, st>>>(D, M, N, out);
    }
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_init_ya(const MatSrc& D, bool hankel, int64_t M, int64_t N, double dual, double* Y, double* A,
                           double* W, double im, double eps, int nonnegE, int sm_count, cudaStream_t st,
                           int64_t* launches, int64_t ldy) {
    const int grid = stream_grid(M, sm_count);
    if (ldy <= 0) ldy = M;
    if (hankel) init_ya_kernel<true><<<grid, 256, 0, st>>>(D, M, N, dual, Y, A, W, im, eps, nonnegE, ldy);
    else init_ya_kernel<false><<<grid, 256, 0, st>>>(D, M, N, dual, Y, A, W, im, eps, nonnegE, ldy);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_compute_e(const MatSrc& D, bool hankel, int64_t M, int64_t N, const double* A, const double* Y,
                             double im, double eps, int nonnegE, double* E, int sm_count, cudaStream_t st,
                             int64_t* launches) {
    const int grid = stream_grid(M * N, sm_count);
    if (hankel) compute_e_kernel<true><<<grid, 256, 0, st>>>(D, M, N, A, Y, im, eps, nonnegE, E);
    else compute_e_kernel<false><<<grid, 256, 0, st>>>(D, M, N, A, Y, im, eps, nonnegE, E);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_transpose(const double* in, int64_t M, int64_t N, double* out, cudaStream_t st,
                             int64_t* launches) {
    dim3 grid((unsigned)((M + 31) / 32), (unsigned)((N + 31) / 32));
    transpose_kernel<<<grid, 256, 0, st>>>(in, M, N, out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_hankel(const double* x, int64_t K, int64_t L, int64_t lag, double* H, cudaStream_t st,
                          int64_t* launches) {
    const int grid = stream_grid(K * L, 148);
    hankel_kernel<<<grid, 256, 0, st>>>(x, K, L, lag, H);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_unhankel(const double* A, int64_t K, int64_t L, int64_t lag, int64_t Ns, double* y,
                            cudaStream_t st, int64_t* launches) {
    const int grid = stream_grid(Ns, 148);
    unhankel_kernel<<<grid, 256, 0, st>>>(A, K, L, lag, Ns, y);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_unhankel_partial(const double* A, int64_t r0, int64_t Kl, int64_t L, int64_t lag, int64_t Ns,
                                    double* sum, double* cnt, cudaStream_t st, int64_t* launches) {
    const int64_t span = (Kl - 1) * lag + L;
    unhankel_partial_kernel<<<stream_grid(span, 148), 256, 0, st>>>(A, r0, Kl, L, lag, Ns, sum, cnt);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_unhankel_divide(const double* sum, const double* cnt, int64_t Ns, double* y, cudaStream_t st,
                                   int64_t* launches) {
    unhankel_divide_kernel<<<stream_grid(Ns, 148), 256, 0, st>>>(sum, cnt, Ns, y);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_unhankel_factors(const double* T, int64_t ldt, const double* V, int svp, int nonnegA, int64_t r0,
                                    int64_t Kl, int64_t n, int64_t Ns, double* sum, int sm_count, cudaStream_t st,
                                    int64_t* launches) {
    if (svp < 1) return cudaMemsetAsync(sum, 0, (size_t)Ns * sizeof(double), st);     // A = 0
    const int rp = svp;
    const size_t smem = ((size_t)n * rp + (size_t)rp * (256 + n - 1)) * sizeof(double);
    if (smem > (size_t)220 * 1024) return cudaErrorInvalidValue;
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(unhankel_factors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = smem;
    }
    int64_t blocks = (Ns + 255) / 256;
    if (blocks > (int64_t)sm_count * 4) blocks = (int64_t)sm_count * 4;
    unhankel_factors_kernel<<<(unsigned)blocks, 256, smem, st>>>(T, ldt, V, rp, nonnegA, r0, Kl, (int)n, Ns, sum);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_unhankel_divide_count(const double* sum, int64_t K, int64_t n, int64_t Ns, double* y, cudaStream_t st,
                                         int64_t* launches) {
    unhankel_divide_count_kernel<<<stream_grid(Ns, 148), 256, 0, st>>>(sum, K, n, Ns, y);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace tlsq
