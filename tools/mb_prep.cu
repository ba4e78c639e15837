// Micro-benchmark (development aid): times hess_fused and form_q in isolation, one CTA per matrix.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I paladin_b200/csrc tools/mb_prep.cu -o tools/mb_prep
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "pb_hess.cuh"

__global__ void __launch_bounds__( 256 ) k_hess( int n, double* A, size_t stride, double* tau ) {
  extern __shared__ double sm[];
  pb::hess_fused<16>( n, 0, n - 1, A + blockIdx.x * stride, n, tau + (size_t)blockIdx.x * n, sm );
}
__global__ void __launch_bounds__( 256 ) k_formq( int n, const double* A, size_t stride, const double* tau, double* Q ) {
  pb::form_q<16>( n, 0, n - 1, A + blockIdx.x * stride, n, tau + (size_t)blockIdx.x * n, Q + blockIdx.x * stride, n );
}
__global__ void k_clock( long long* out ) { long long t0 = clock64(); __nanosleep( 1000000 ); out[0] = clock64() - t0; }

int main( int argc, char** argv ) {
  const int n = 512, cnt = argc > 1 ? atoi( argv[1] ) : 148;
  const size_t stride = (size_t)n * n + 528;
  std::vector<double> h( stride * cnt );
  srand( 1 );
  for ( auto& x : h ) x = rand() / (double)RAND_MAX - 0.5;
  double *A, *Q, *tau;
  cudaMalloc( &A, sizeof( double ) * stride * cnt );
  cudaMalloc( &Q, sizeof( double ) * stride * cnt );
  cudaMalloc( &tau, sizeof( double ) * n * cnt );
  cudaEvent_t e0, e1;
  cudaEventCreate( &e0 );
  cudaEventCreate( &e1 );
  const size_t smem = sizeof( double ) * pb::hess_smem_doubles( n );
  cudaFuncSetAttribute( k_hess, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem );
  for ( int rep = 0; rep < 3; ++rep ) {
    cudaMemcpy( A, h.data(), sizeof( double ) * stride * cnt, cudaMemcpyHostToDevice );
    float ms1, ms2;
    cudaEventRecord( e0 );
    k_hess<<<cnt, 256, smem>>>( n, A, stride, tau );
    cudaEventRecord( e1 );
    cudaEventSynchronize( e1 );
    cudaEventElapsedTime( &ms1, e0, e1 );
    cudaEventRecord( e0 );
    k_formq<<<cnt, 256>>>( n, A, stride, tau, Q );
    cudaEventRecord( e1 );
    cudaEventSynchronize( e1 );
    cudaEventElapsedTime( &ms2, e0, e1 );
    printf( "cnt %d  hess_fused %.2f ms   form_q %.2f ms   (%s)\n", cnt, ms1, ms2, cudaGetErrorString( cudaGetLastError() ) );
  }
  return 0;
}
