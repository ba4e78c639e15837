// Micro-benchmark (development aid): the warp-level Francis iteration (pb_small.cuh::lahqr) in isolation.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I paladin_b200/csrc tools/mb_lahqr.cu -o tools/mb_lahqr
// Prints the dependent-chain latencies the iteration is made of (DFMA, larfg3_fast, a shared-memory round trip) and the
// cycles per Francis step of lahqr on random Hessenberg matrices of order 32 / 48 / 64, one matrix per warp, for 1..8
// warps per SM.
#include <cstdio>
#include <cstdlib>
#include <vector>
__constant__ int mb_do_count;
__device__ unsigned long long mb_steps[4];
#define PB_AED_COUNT( what ) ( ( mb_do_count && ( threadIdx.x & 31 ) == 0 ) ? (void)atomicAdd( &mb_steps[what], 1ull ) : (void)0 )
__device__ unsigned long long mb_sweep_cycles;
// cycles spent inside the sweeps proper (the rest of lahqr is the deflation scan, the shifts and the start vector)
__shared__ long long mb_t0[32];
#define PB_SWEEP_PROFILE( end ) \
  do { \
    if ( ( threadIdx.x & 31 ) == 0 ) { \
      if ( end ) atomicAdd( &mb_sweep_cycles, (unsigned long long)( clock64() - mb_t0[threadIdx.x >> 5] ) ); \
      else mb_t0[threadIdx.x >> 5] = clock64(); \
    } \
  } while ( 0 )
#include "pb_small.cuh"

__global__ void k_lat( double* p, long long* out ) {
  __shared__ double sm[64];
  double x = p[0], y = p[1];
  long long t0 = clock64();
#pragma unroll 1
  for ( int i = 0; i < 64; ++i ) {
    x = fma( x, y, y );
    x = fma( x, y, y );
    x = fma( x, y, y );
    x = fma( x, y, y );
  }
  long long t1 = clock64();
  out[0] = t1 - t0;  // 256 dependent DFMA
  double a0 = x, a1 = p[2], a2 = p[3], t;
  t0 = clock64();
#pragma unroll 1
  for ( int i = 0; i < 64; ++i ) {
    pb::larfg3_fast( 3, a0, a1, a2, t );
    a1 += t;
    a2 += a0;
  }
  t1 = clock64();
  out[1] = t1 - t0;  // 64 reflectors (+ 1 DADD each)
  sm[threadIdx.x] = a1;
  __syncwarp();
  t0 = clock64();
  double z = a2;
#pragma unroll 1
  for ( int i = 0; i < 64; ++i ) {
    sm[( threadIdx.x + 1 ) & 31] = z;
    __syncwarp();
    z = sm[threadIdx.x] + 1.0;
    __syncwarp();
  }
  t1 = clock64();
  out[2] = t1 - t0;  // 64 x (STS, syncwarp, LDS, DADD, syncwarp)
  p[4 + threadIdx.x] = z + a0 + a1 + t;
}

// one matrix per warp: Hessenberg H (m x m, ld m+1) and V in shared memory
// warps >= nwork of a CTA are "noise": independent DFMA streams (what the strip phase of a neighbouring CTA looks like
// to the scheduler) until the working warps are done
__device__ double mb_sink;
template <bool WANTT>
__global__ void k_lahqr( int m, const double* Hin, long long* cyc, double* wout, int nwork ) {
  extern __shared__ double sm[];
  __shared__ int done;
  if ( threadIdx.x == 0 ) done = 0;
  __syncthreads();
  if ( (int)( threadIdx.x >> 5 ) >= nwork ) {
    double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    const double x = 1.0000001, y = 1e-9;
    while ( *( (volatile int*)&done ) < nwork ) {
#pragma unroll
      for ( int q = 0; q < 16; ++q ) {
        a0 = fma( a0, x, y ); a1 = fma( a1, x, y ); a2 = fma( a2, x, y ); a3 = fma( a3, x, y );
        a4 = fma( a4, x, y ); a5 = fma( a5, x, y ); a6 = fma( a6, x, y ); a7 = fma( a7, x, y );
      }
    }
    if ( a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 0.123 ) mb_sink = a0;
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = m + 1;
  double* H = sm + (size_t)warp * ( 2 * m * ld + 2 * m );
  double* V = H + m * ld;
  double* wr = V + m * ld;
  double* wi = wr + m;
  const double* src = Hin + (size_t)( blockIdx.x * nwork + warp ) * m * m;
  for ( int q = lane; q < m * m; q += 32 ) {
    const int c = q / m, r = q - c * m;
    H[r + c * ld] = ( r - c <= 1 ) ? src[q] : 0.0;
    V[r + c * ld] = ( r == c ) ? 1.0 : 0.0;
  }
  __syncwarp();
  const long long t0 = clock64();
  const int info = pb::lahqr<pb::WarpGroup, true>( pb::WarpGroup(), WANTT, WANTT, m, 0, m - 1, H, ld, wr, wi, 0, m - 1, V, ld );
  const long long t1 = clock64();
  if ( lane == 0 ) cyc[blockIdx.x * nwork + warp] = ( t1 - t0 ) + ( info != 0 ? ( 1ll << 50 ) : 0 );
  if ( lane == 0 ) atomicAdd( &done, 1 );
  if ( blockIdx.x == 0 && warp == 0 )
    for ( int q = lane; q < m; q += 32 ) {
      wout[q] = wr[q];
      wout[m + q] = wi[q];
    }
}

int main() {
  double* p;
  long long* out;
  cudaMalloc( &p, 64 * sizeof( double ) );
  cudaMalloc( &out, 65536 * sizeof( long long ) );
  double hp[4] = { 0.999, 1e-3, 0.7, -0.3 };
  cudaMemcpy( p, hp, sizeof( hp ), cudaMemcpyHostToDevice );
  k_lat<<<1, 32>>>( p, out );
  long long ho[3];
  cudaMemcpy( ho, out, sizeof( ho ), cudaMemcpyDeviceToHost );
  printf( "DFMA dependent latency %.1f cycles; larfg3_fast %.1f cycles; STS+sync+LDS+DADD+sync %.1f cycles\n", ho[0] / 256.0, ho[1] / 64.0, ho[2] / 64.0 );
  const int sizes[3] = { 32, 48, 64 };
  const int nsm = 148;
  for ( int si = 0; si < 1; ++si ) {
    const int m = sizes[si];
    for ( int wantt = 0; wantt < 2; ++wantt ) {
      for ( int cfg = 0; cfg < 5; ++cfg ) {
        const int wps = cfg == 0 ? 1 : cfg == 1 ? 8 : 1;            // working warps per SM
        const int noise = cfg == 2 ? 3 : cfg == 3 ? 8 : cfg == 4 ? 15 : 0;  // FP64-streaming warps next to them
        const size_t smem = sizeof( double ) * (size_t)wps * ( 2 * m * ( m + 1 ) + 2 * m );
        if ( smem > 200 * 1024 ) continue;
        const int nmat = nsm * wps;
        std::vector<double> h( (size_t)nmat * m * m );
        srand( 7 );
        for ( auto& x : h ) {
          double u = 0;
          for ( int q = 0; q < 12; ++q ) u += rand() / (double)RAND_MAX;
          x = u - 6.0;
        }
        double *Hin, *wout;
        cudaMalloc( &Hin, h.size() * sizeof( double ) );
        cudaMalloc( &wout, 2 * m * sizeof( double ) );
        cudaMemcpy( Hin, h.data(), h.size() * sizeof( double ), cudaMemcpyHostToDevice );
        auto kern = wantt ? k_lahqr<true> : k_lahqr<false>;
        cudaFuncSetAttribute( kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem );
        int one = 1, zero = 0;
        unsigned long long z4[4] = { 0, 0, 0, 0 };
        cudaMemcpyToSymbol( mb_steps, z4, sizeof( z4 ) );
        cudaMemcpyToSymbol( mb_sweep_cycles, z4, sizeof( unsigned long long ) );
        cudaMemcpyToSymbol( mb_do_count, &one, sizeof( int ) );
        kern<<<nsm, 32 * ( wps + noise ), smem>>>( m, Hin, out, wout, wps );
        cudaDeviceSynchronize();
        unsigned long long st[4];
        cudaMemcpyFromSymbol( st, mb_steps, sizeof( st ) );
        unsigned long long swc = 0;
        cudaMemcpyFromSymbol( &swc, mb_sweep_cycles, sizeof( swc ) );
        cudaMemcpyToSymbol( mb_do_count, &zero, sizeof( int ) );
        cudaEvent_t e0, e1;
        cudaEventCreate( &e0 );
        cudaEventCreate( &e1 );
        cudaEventRecord( e0 );
        kern<<<nsm, 32 * ( wps + noise ), smem>>>( m, Hin, out, wout, wps );
        cudaEventRecord( e1 );
        cudaEventSynchronize( e1 );
        float ms;
        cudaEventElapsedTime( &ms, e0, e1 );
        std::vector<long long> hc( nmat );
        cudaMemcpy( hc.data(), out, nmat * sizeof( long long ), cudaMemcpyDeviceToHost );
        double tot = 0;
        int bad = 0;
        for ( auto c : hc ) {
          if ( c >= ( 1ll << 50 ) ) ++bad, c -= ( 1ll << 50 );
          tot += (double)c;
        }
        const double steps = (double)( st[1] + st[2] );
        std::vector<double> w( 2 * m );
        cudaMemcpy( w.data(), wout, 2 * m * sizeof( double ), cudaMemcpyDeviceToHost );
        double tr = 0;
        for ( int q = 0; q < m; ++q ) tr += w[q];
        printf( "m %2d wantt %d warps/SM %d noise %2d: %.3f ms, %.0f steps/matrix, %.0f sweeps/matrix, %.0f cycles/step (avg over warps; %.0f inside the sweeps, counting run), failed %d, trace(w) %.6f (%s)\n", m, wantt, wps, noise, ms,
                steps / nmat, (double)st[3] / nmat, tot / steps, (double)swc / steps, bad, tr, cudaGetErrorString( cudaGetLastError() ) );
        cudaFree( Hin );
        cudaFree( wout );
      }
    }
  }
  return 0;
}
