// stream.cu -- the ALM epilogue as two streaming kernels (used when the SVT input W_k is materialised and svp <= 32):
//
//   alm_stream_kernel  every thread owns one matrix row:
//                      loop 1  T[i,:] = (W_k[i,:] V_r) .* f      streaming read of W, svp FMAs per element
//                      loop 2  one coalesced pass over D, A_{k-1}, Y_{k-1}:
//                        A_k = clamp(sum_c T[i,c] V[j,c])                         src/robustPCA.jl:205-219
//                        E_k = soft_th((D - A_{k-1}) + Y_{k-1}/mu, lambda/mu)     :188-191
//                        Z = (D - A_k) - E_k ;  Y_k = Y_{k-1} + mu Z ;  ||Z||_F^2 :221-222, 225
//                        W_{k+1} = (D - E_{k+1}) + Y_k/mu_{k+1}                   :188-192 of the NEXT iteration
//                      and writes A_k, Y_k, W_{k+1} (E_k only on request).
// Compared with the fused tile kernel (epilogue.cu) this moves one extra S of traffic (W is read by tproj) but both
// kernels are pure streams with every thread owning one row: 128-byte coalesced column segments, no shared-memory
// staging of the data, no intra-CTA phases -- the memory system sees a STREAM-like access pattern.
// The low-rank reconstruction costs svp FP64 FMAs per element, which the otherwise idle FP64 pipe absorbs.
#include "kernels.h"

namespace tlsq {

namespace {

template <int RP, bool HANKEL>
__global__ void __launch_bounds__(128)
alm_stream_kernel(const EpiArgs a, const double* __restrict__ W, int svp) {
    extern __shared__ double Vsm[];            // [N][RP]
    __shared__ double fsm[RP > 0 ? RP : 1];
    const int N = (int)a.N;
    if (threadIdx.x < RP) fsm[threadIdx.x] = (int)threadIdx.x < svp ? __ldg(a.fvec + threadIdx.x) : 0.0;
    for (int idx = threadIdx.x; idx < N * RP; idx += blockDim.x) {
        const int j = idx % N, c = idx / N;
        Vsm[j * RP + c] = c < svp ? __ldg(a.Vs + (int64_t)c * N + j) : 0.0;
    }
    __syncthreads();
    double zz = 0.0;
    constexpr int UB = 4;                      // columns whose loads are issued together
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < a.M;
         row += (int64_t)gridDim.x * blockDim.x) {
        // ---- T[row, :] = f .* (W[row, :] V_r): one streaming read of the materialised SVT input ----------------
        double tr[RP > 0 ? RP : 1];
#pragma unroll
        for (int c = 0; c < RP; ++c) tr[c] = 0.0;
        if (RP > 0) {
            constexpr int UW = 8;
            for (int j0 = 0; j0 < N; j0 += UW) {
                double wv[UW];
#pragma unroll
                for (int u = 0; u < UW; ++u) {
                    const int j = j0 + u;
                    wv[u] = j < N ? __ldg(W + (int64_t)j * a.ldw + row) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < UW; ++u) {
                    const int j = j0 + u < N ? j0 + u : N - 1;
                    const double* v = Vsm + j * RP;
#pragma unroll
                    for (int c = 0; c < RP; ++c) tr[c] = fma(wv[u], v[c], tr[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < RP; ++c) tr[c] *= fsm[c];
        }
        for (int j0 = 0; j0 < N; j0 += UB) {
            double dv[UB], av[UB], yv[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int j = j0 + u;
                if (j < N) {
                    const int64_t off = (int64_t)j * a.ldw + row;
                    dv[u] = src_at<HANKEL>(a.D, row, j);
                    av[u] = __ldg(a.Ap + off);
                    yv[u] = __ldg(a.Yp + off);
                }
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int j = j0 + u;
                if (j < N) {
                    const int64_t off = (int64_t)j * a.ldw + row;
                    const double d = dv[u], ap = av[u], yp = yv[u];
                    double an = 0.0;
                    const double* v = Vsm + j * RP;
#pragma unroll
                    for (int c = 0; c < RP; ++c) an = fma(tr[c], v[c], an);
                    if (a.nonnegA) an = (__double_as_longlong(an) > 0) ? an : 0.0;   // A .= max.(A, 0)   :218
                    double e, w;
                    alm_ew(d, ap, yp, a.im, a.eps, a.nonnegE, e, w);
                    const double z = __dsub_rn(__dsub_rn(d, an), e);                  // @. Z = D - A - E  :221
                    const double yn = __dadd_rn(yp, __dmul_rn(a.mu, z));              // @. Y = Y + mu*Z   :222
                    zz = fma(z, z, zz);
                    a.An[off] = an;
                    a.Yn[off] = yn;
                    if (a.Eout) a.Eout[off] = e;
                    if (a.Zout) a.Zout[off] = z;
                    double e2, w2;
                    alm_ew(d, an, yn, a.im_next, a.eps_next, a.nonnegE, e2, w2);      // SVT input of iteration k+1
                    a.Wn[off] = w2;
                }
            }
        }
    }
    __shared__ double red[4];
    zz = warp_sum(zz);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = zz;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(a.zz, red[0] + red[1] + red[2] + red[3]);
}

template <int RP>
cudaError_t launch_stream_rp(const EpiArgs& a, const double* W, int svp, bool hankel, int sm_count, cudaStream_t st) {
    const size_t smem = (size_t)a.N * RP * sizeof(double);
    cudaError_t e;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(alm_stream_kernel<RP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(alm_stream_kernel<RP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int64_t blocks = (a.M + 127) / 128;
    int64_t cap = (int64_t)sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (hankel) alm_stream_kernel<RP, true><<<(unsigned)blocks, 128, smem, st>>>(a, W, svp);
    else alm_stream_kernel<RP, false><<<(unsigned)blocks, 128, smem, st>>>(a, W, svp);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_stream_epilogue(const EpiArgs& a, const double* W, double* T, int svp, bool hankel, int sm_count,
                                   cudaStream_t st, int64_t* launches) {
    cudaError_t e;
    const int N = (int)a.N;
    const int rp = (svp + 7) & ~7;
    (void)T;
    switch (rp) {
        case 0:  e = launch_stream_rp<0>(a, W, svp, hankel, sm_count, st); break;
        case 8:  e = launch_stream_rp<8>(a, W, svp, hankel, sm_count, st); break;
        case 16: e = launch_stream_rp<16>(a, W, svp, hankel, sm_count, st); break;
        case 24: e = launch_stream_rp<24>(a, W, svp, hankel, sm_count, st); break;
        default: e = launch_stream_rp<32>(a, W, svp, hankel, sm_count, st); break;
    }
    if (launches) *launches += 1;
    return e;
}

}  // namespace tlsq
