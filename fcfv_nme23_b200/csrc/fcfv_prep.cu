// Space-filling-curve ELEMENT order (mesh preparation, off the hot path).
//
// The assembly / residual kernels gather the records of the elements next to consecutive faces, so they want
// elements that are close in the plane to be close in memory.  Mesh generators usually deliver that (Triangle's
// divide-and-conquer output, the row-major structured generator); a mesh that does not (elements listed in random
// order) runs 3x slower.  This file computes the remedy: elements sorted along a Hilbert curve through their
// centroids.  It is the only place in the library that calls library device code (CUB reductions and radix sort).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cuda_runtime.h>

#include "fcfv_prep.h"

namespace fcfv {
namespace {

constexpr int kHilbertBits = 20;     // per coordinate: 2^20 x 2^20 cells, 40-bit keys

// index of cell (x, y) along the Hilbert curve of a 2^bits x 2^bits grid (most significant quadrant first)
__host__ __device__ inline unsigned long long hilbert_index(unsigned x, unsigned y, int bits) {
    unsigned long long d = 0;
    for (unsigned s = 1u << (bits - 1); s > 0; s >>= 1) {
        const unsigned rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += (unsigned long long)s * s * ((3u * rx) ^ ry);
        x &= s - 1; y &= s - 1;
        if (ry == 0) {
            if (rx == 1) { x = s - 1 - x; y = s - 1 - y; }
            const unsigned t = x; x = y; y = t;
        }
    }
    return d;
}

// box = xmin, xmax, ymin, ymax.  q = min(2^bits - 1, trunc((x - xmin) * (2^bits / (xmax - xmin)))), 0 for a flat box.
__global__ void k_hilbert_keys(int n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ box,
                               unsigned long long* __restrict__ key, int* __restrict__ idx) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double cells = (double)(1u << kHilbertBits);
    const double wx = box[1] - box[0], wy = box[3] - box[2];
    const double sx = wx > 0.0 ? cells / wx : 0.0, sy = wy > 0.0 ? cells / wy : 0.0;
    long long qx = (long long)((x[e] - box[0]) * sx), qy = (long long)((y[e] - box[2]) * sy);
    const long long top = (1LL << kHilbertBits) - 1;
    qx = qx < 0 ? 0 : (qx > top ? top : qx);
    qy = qy < 0 ? 0 : (qy > top ? top : qy);
    key[e] = hilbert_index((unsigned)qx, (unsigned)qy, kHilbertBits);
    idx[e] = e;
}

__global__ void k_widen_1based(int n, const int* __restrict__ in, long long* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = (long long)in[e] + 1;
}

size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace

// old_of_new (n, 1-based, device): element k of the new order is element old_of_new[k] of the old one; ties (same
// cell) keep their old order (the radix sort is stable).  tmp == nullptr: only *need is written.
cudaError_t hilbert_element_order(cudaStream_t st, int n, const double* d_x, const double* d_y, long long* d_old_of_new,
                                  void* tmp, size_t tmp_bytes, size_t* need) {
    size_t red = 0, srt = 0;
    cudaError_t rc = cub::DeviceReduce::Min(nullptr, red, d_x, (double*)nullptr, n, st);
    if (rc != cudaSuccess) return rc;
    rc = cub::DeviceRadixSort::SortPairs(nullptr, srt, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                         (const int*)nullptr, (int*)nullptr, n, 0, 2 * kHilbertBits, st);
    if (rc != cudaSuccess) return rc;
    const size_t o_box = 0, o_k0 = align256(4 * sizeof(double)), o_k1 = o_k0 + align256(sizeof(unsigned long long) * (size_t)n),
                 o_i0 = o_k1 + align256(sizeof(unsigned long long) * (size_t)n), o_i1 = o_i0 + align256(sizeof(int) * (size_t)n),
                 o_cub = o_i1 + align256(sizeof(int) * (size_t)n), total = o_cub + align256(red > srt ? red : srt);
    if (need) *need = total;
    if (!tmp) return cudaSuccess;
    if (tmp_bytes < total) return cudaErrorInvalidValue;
    if (n == 0) return cudaSuccess;
    char* base = (char*)tmp;
    double* box = (double*)(base + o_box);
    auto* k0 = (unsigned long long*)(base + o_k0); auto* k1 = (unsigned long long*)(base + o_k1);
    int* i0 = (int*)(base + o_i0); int* i1 = (int*)(base + o_i1);
    void* ctmp = base + o_cub;
    size_t cb = red;
    if ((rc = cub::DeviceReduce::Min(ctmp, cb, d_x, box + 0, n, st)) != cudaSuccess) return rc;
    if ((rc = cub::DeviceReduce::Max(ctmp, cb, d_x, box + 1, n, st)) != cudaSuccess) return rc;
    if ((rc = cub::DeviceReduce::Min(ctmp, cb, d_y, box + 2, n, st)) != cudaSuccess) return rc;
    if ((rc = cub::DeviceReduce::Max(ctmp, cb, d_y, box + 3, n, st)) != cudaSuccess) return rc;
    k_hilbert_keys<<<(n + 255) / 256, 256, 0, st>>>(n, d_x, d_y, box, k0, i0);
    cb = srt;
    if ((rc = cub::DeviceRadixSort::SortPairs(ctmp, cb, k0, k1, i0, i1, n, 0, 2 * kHilbertBits, st)) != cudaSuccess) return rc;
    k_widen_1based<<<(n + 255) / 256, 256, 0, st>>>(n, i1, d_old_of_new);
    return cudaGetLastError();
}

}  // namespace fcfv
