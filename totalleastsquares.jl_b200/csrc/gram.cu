// gram.cu -- FP64 tensor-core (DMMA) SYRK:  G = W'W with W formed on the fly from (D, A, Y).
//
// Replaces the LAPACK dgesdd calls of the reference (opnorm :177,225 and svd! :194): only the n x n Gram of
// the tall M x n SVT input is ever needed.  Work decomposition:
//   grid.x = upper-triangular 64x64 output blocks (bi <= bj), grid.y = split over 32-row K tiles.
//   CTA: 256 threads = 8 warps in a 4 x 2 layout, warp tile 16 (i) x 32 (j) = 2 x 4 DMMA.8x8x4 accumulators.
//   Per K tile the two 32 x 64 column panels of W are computed element-wise (soft-threshold etc.) straight from
//   global memory into padded shared memory ([col][36] doubles: the 8-byte fragment loads of a half-warp hit
//   16 distinct banks), then 8 k-steps of DMMA.  Partial blocks go to a workspace and are summed in a fixed
//   order by gram_reduce_kernel (deterministic, and every rank of a sharded run sees identical bits after the
//   all-reduce).
#include "kernels.h"

namespace tlsq {

namespace {

constexpr int GB = kGramBlk;
constexpr int GR = kGramRows;
constexpr int GRS = GR + 4;

template <int MODE, bool HANKEL>
__device__ __forceinline__ double gram_elem(const GramSrc& s, int64_t row, int64_t col) {
    const double d = src_at<HANKEL>(s.D, row, col);
    if (MODE == GRAM_D) return d;
    const int64_t off = col * s.ldw + row;
    const double a = __ldg(s.A + off);
    const double y = __ldg(s.Y + off);
    double e, w;
    alm_ew(d, a, y, s.im, s.eps, s.nonnegE, e, w);
    if (MODE == GRAM_W) return w;
    const double a2 = __ldg(s.A2 + off);
    return __dsub_rn(__dsub_rn(d, a2), e);          // @. Z = D - A - E   (src/robustPCA.jl:221)
}

template <int MODE, bool HANKEL>
__global__ void __launch_bounds__(256, 3)
gram_kernel(const GramSrc s, double* __restrict__ partial, int nb, int ntiles) {
    __shared__ double Wi[GB * GRS];
    __shared__ double Wj[GB * GRS];

    // decode (bi, bj), bi <= bj, from the linear upper-triangular block index
    int bi = 0, bj = 0;
    {
        int rem = blockIdx.x;
        for (bi = 0; bi < nb; ++bi) {
            const int cnt = nb - bi;
            if (rem < cnt) { bj = bi + rem; break; }
            rem -= cnt;
        }
    }
    const bool diag = (bi == bj);
    const double* Wjp = diag ? Wi : Wj;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wi = warp & 3, wj = warp >> 2;
    // on diagonal blocks skip warp tiles that lie entirely below the diagonal
    const bool active = !(diag && (16 * wi >= 32 * wj + 32));

    double acc[2][4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    const int lrow = lane, lcol0 = warp;     // loader: a warp reads 32 consecutive rows of one column

    for (int tile = blockIdx.y; tile < ntiles; tile += gridDim.y) {
        const int64_t row = (int64_t)tile * GR + lrow;
        const bool rok = row < s.M;
        double wv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int64_t col = (int64_t)bi * GB + lcol0 + 8 * q;
            wv[q] = (rok && col < s.N) ? gram_elem<MODE, HANKEL>(s, row, col) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) Wi[(lcol0 + 8 * q) * GRS + lrow] = wv[q];
        if (!diag) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int64_t col = (int64_t)bj * GB + lcol0 + 8 * q;
                wv[q] = (rok && col < s.N) ? gram_elem<MODE, HANKEL>(s, row, col) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) Wj[(lcol0 + 8 * q) * GRS + lrow] = wv[q];
        }
        __syncthreads();
        if (active) {
#pragma unroll
            for (int k0 = 0; k0 < GR; k0 += 4) {
                double a[2], b[4];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) a[mi] = Wi[(16 * wi + 8 * mi + g) * GRS + k0 + t];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = Wjp[(32 * wj + 8 * ni + g) * GRS + k0 + t];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
        }
        __syncthreads();
    }

    double* P = partial + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * (GB * GB);
    if (active) {
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int il = 16 * wi + 8 * mi + g;
                const int jl = 32 * wj + 8 * ni + 2 * t;
                P[jl * GB + il] = acc[mi][ni][0];
                P[(jl + 1) * GB + il] = acc[mi][ni][1];
            }
    }
}

// G[i,j] = G[j,i] = sum_split partial[split][blk(i,j)][jl][il]   for i <= j  (fixed summation order)
__global__ void gram_reduce_kernel(const double* __restrict__ partial, int nsplit, int nblk, int nb, int B, int N,
                                   double* __restrict__ G) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * N) return;
    const int i = (int)(idx % N), j = (int)(idx / N);
    if (i > j) return;
    const int bi = i / B, bj = j / B, il = i % B, jl = j % B;
    const int blk = bi * nb - (bi * (bi - 1)) / 2 + (bj - bi);
    const double* p = partial + (int64_t)blk * (B * B) + jl * B + il;
    double sum = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) sum += p[(int64_t)sp * nblk * (B * B)];
    G[(int64_t)j * N + i] = sum;
    G[(int64_t)i * N + j] = sum;
}

template <int MODE>
cudaError_t launch_mode(const GramSrc& s, bool hankel, const GramPlan& plan, double* partial, cudaStream_t st) {
    dim3 grid(plan.nblk, plan.nsplit), block(256);
    if (hankel)
        gram_kernel<MODE, true><<<grid, block, 0, st>>>(s, partial, plan.nb, plan.ntiles);
    else
        gram_kernel<MODE, false><<<grid, block, 0, st>>>(s, partial, plan.nb, plan.ntiles);
    return cudaGetLastError();
}

}  // namespace

GramPlan gram_plan(int64_t M, int64_t N, int sm_count) {
    GramPlan p;
    p.nb = (int)((N + GB - 1) / GB);
    p.nblk = p.nb * (p.nb + 1) / 2;
    p.ntiles = (int)((M + GR - 1) / GR);
    if (p.ntiles < 1) p.ntiles = 1;
    int slots = sm_count * 3;                     // 3 resident CTAs per SM (launch bounds)
    int ns = (slots + p.nblk - 1) / p.nblk;
    if (ns > p.ntiles) ns = p.ntiles;
    if (ns < 1) ns = 1;
    p.nsplit = ns;
    p.partial_bytes = (size_t)p.nsplit * p.nblk * GB * GB * sizeof(double);
    return p;
}

cudaError_t launch_gram(const GramSrc& s, GramMode mode, bool hankel, const GramPlan& plan, double* partial,
                        double* G, cudaStream_t st, int64_t* launches) {
    cudaError_t e;
    switch (mode) {
        case GRAM_D: e = launch_mode<GRAM_D>(s, hankel, plan, partial, st); break;
        case GRAM_W: e = launch_mode<GRAM_W>(s, hankel, plan, partial, st); break;
        default:     e = launch_mode<GRAM_Z>(s, hankel, plan, partial, st); break;
    }
    if (e != cudaSuccess) return e;
    if (launches) *launches += 2;
    return launch_gram_reduce(partial, plan.nsplit, plan.nblk, plan.nb, GB, (int)s.N, G, st);
}

cudaError_t launch_gram_reduce(const double* partial, int nsplit, int nblk, int nb, int B, int N, double* G,
                               cudaStream_t st) {
    const int64_t nn = (int64_t)N * N;
    gram_reduce_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(partial, nsplit, nblk, nb, B, N, G);
    return cudaGetLastError();
}

}  // namespace tlsq
