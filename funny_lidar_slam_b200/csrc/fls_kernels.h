// fls_kernels.h — launch interfaces of the residual kernels (K1 p2plane/iVox, K2 NDT, K3 ICP).
#pragma once
#include "fls_common.cuh"
#include "fls_ivox.cuh"

namespace fls {

static constexpr int kP2PlaneBlock = 128;

struct P2PlaneArgs {
    const float4* __restrict__ src;  // body-frame scan, packed float4
    int n;
    IvoxView map;
    double plane_thres;
    GnState* state;
    float4* __restrict__ rec0;  // persistent per-point record: J0..J3
    float4* __restrict__ rec1;  //                              J4, J5, |d|, 1
    unsigned char* __restrict__ flags;
    double* __restrict__ partials;  // [grid][kAccStride]
};

int p2plane_grid(int n);
void launch_p2plane_iter(const P2PlaneArgs& a, cudaStream_t st);
void launch_ivox_knn_test(const IvoxView& map, const float4* d_q, int n, float4* d_out, int* d_found, cudaStream_t st);

}  // namespace fls
