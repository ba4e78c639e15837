// pb_coop.cuh -- Hessenberg reduction and explicit Q for LARGE matrices (n > 1024) on a GROUP of cooperating
// CTAs (device only). One CTA per matrix leaves the GPU idle when a batch holds a few large matrices
// (BASELINE.json configs[4]: 3072 x 3072) and takes seconds per matrix; here a group of S CTAs (37, 74 or all 148
// SMs, one CTA per SM, cooperative launch) works on one matrix at a time.
//
// Replaces the dgehrd / dorghr stages of the reference's dgeev_ call (reference src/lapack-wrapper.hpp:47-60,
// src/paladin.hpp:248-254). Same reflectors as LAPACK's unblocked dgehd2 and the same single-sweep schedule as
// pb_hess.cuh (the rank-2 update of step j-1 is applied in the sweep that accumulates A v and A^T v of step j, so the
// trailing matrix -- 72 MB at n = 3072, resident in the 126 MB L2 -- is read once and written once per column), with
// the COLUMNS of the trailing matrix dealt round-robin to the CTAs of the group:
//   * column j (brought up to date) and its reflector are computed REDUNDANTLY by every CTA from the same data in the
//     same order, hence bit-identical everywhere: no barrier between reflector generation and the sweep;
//   * in the sweep a thread owns rows t, t+256, ...: accumulators of (A v)(r) live in registers, (A^T v)(c) is one
//     shuffle reduction per warp and column -- no shared-memory atomics (measured and rejected: double-buffering the
//     column batches in registers, 1026 vs 910 ms per 8 matrices of 3072^2; groups of 4 CTAs for n = 512, 387 vs 133 ms
//     per 592 matrices for the one-CTA fused kernel: the per-column barriers and reductions dominate small groups);
//   * the per-CTA partial sums of A v are combined by a deterministic reduce-scatter through global memory:
//     two group barriers per column (sense-reversing counter in global memory, co-residency guaranteed by the
//     cooperative launch).
// Explicit Q: column c of Q only depends on reflectors j < c and is owned by one CTA from start to finish, so the
// backward accumulation needs no group barrier at all.
#pragma once

#include "pb_common.cuh"
#include "pb_hess.cuh"

#if defined( __CUDACC__ )

namespace pb {

struct CoopGroup {
  unsigned* bar;  // two words in global memory: arrivals, generation (zero before the launch)
  int size, rank;
};

__device__ __forceinline__ void coop_sync( const CoopGroup& cg ) {
  __syncthreads();
  if ( cg.size > 1 && threadIdx.x == 0 ) {
    volatile unsigned* gen = cg.bar + 1;
    __threadfence();
    const unsigned g0 = *gen;
    if ( atomicAdd( cg.bar, 1u ) == (unsigned)cg.size - 1 ) {
      *( (volatile unsigned*)cg.bar ) = 0u;
      __threadfence();
      *gen = g0 + 1;
    } else {
      while ( *gen == g0 ) {
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// 512 threads per CTA (one CTA per SM): twice the loads in flight of the 256-thread kernels -- the column sweep of a small
// group is bound by how much memory traffic one SM keeps in flight
constexpr int kCoopThreads = 512;
constexpr int kCoopWarps = kCoopThreads / 32;
constexpr int kCoopMaxK = 9;  // rows per thread: n <= 512 * kCoopMaxK = 4608

__device__ __forceinline__ double coop_cta_sum( double x, double* red ) {
  x = warp_sum( x );
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ( ( threadIdx.x & 31 ) == 0 ) red[w] = x;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for ( int i = 0; i < kCoopWarps; ++i ) t += red[i];
  return t;
}
__device__ __forceinline__ double coop_cta_max( double x, double* red ) {
  x = warp_max( x );
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ( ( threadIdx.x & 31 ) == 0 ) red[w] = x;
  __syncthreads();
  double t = red[0];
#pragma unroll
  for ( int i = 1; i < kCoopWarps; ++i ) t = fmax( t, red[i] );
  return t;
}

// shared memory (doubles): vprev, ypp, vcur, zpp, ycur (also column j), zcur : 6 n ; red: 64 ; zw: warps x (n / S + 2)
__host__ __device__ inline size_t coop_hess_smem_doubles( int n, int S ) { return (size_t)6 * n + 64 + (size_t)kCoopWarps * ( n / S + 2 ); }

// One sweep over this CTA's columns c0, c0 + S, ... (ncols of them): applies the rank-2 update of the previous step,
// accumulates (A v)(r) for the thread's rows r = tid + 512 k in registers and (A^T v)(c) per column. Columns are taken CB
// at a time with all their loads issued before the first store, so that CB * K loads per thread are in flight (a
// column-at-a-time loop exposes one L2 round trip per column). K rows per thread: n <= 512 K.
template <int K, int CB>
__device__ __forceinline__ void coop_sweep( int nrow, int S, int c0, int ncols, int ihi, bool have_prev, double* A, int lda,
                                            const double* vprev, const double* ypp, const double* vcur, const double* zpp, double* zw,
                                            int maxc, double* ymine ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double yacc[K];
#pragma unroll
  for ( int k = 0; k < K; ++k ) yacc[k] = 0.0;
  for ( int ci0 = 0; ci0 < ncols; ci0 += CB ) {
    double a[CB][K];
#pragma unroll
    for ( int b = 0; b < CB; ++b ) {
      const double* col = A + (size_t)( c0 + imin( ci0 + b, ncols - 1 ) * S ) * lda;
#pragma unroll
      for ( int k = 0; k < K; ++k ) a[b][k] = __ldcg( col + imin( tid + kCoopThreads * k, nrow - 1 ) );
    }
#pragma unroll
    for ( int b = 0; b < CB; ++b ) {
      if ( ci0 + b < ncols ) {
        const int c = c0 + ( ci0 + b ) * S;
        double* col = A + (size_t)c * lda;
        const double zc = zpp[c], vpc = vprev[c];
        const double vcc = ( c <= ihi ) ? vcur[c] : 0.0;
        double zsum = 0.0;
#pragma unroll
        for ( int k = 0; k < K; ++k ) {
          const int r = tid + kCoopThreads * k;
          if ( r < nrow ) {
            const double x = a[b][k] - ( vprev[r] * zc + ypp[r] * vpc );
            yacc[k] += x * vcc;
            zsum += vcur[r] * x;
            if ( have_prev ) __stcg( col + r, x );
          }
        }
        zsum = warp_sum( zsum );
        if ( lane == 0 ) zw[warp * maxc + ci0 + b] = zsum;
      }
    }
  }
#pragma unroll
  for ( int k = 0; k < K; ++k ) {
    const int r = tid + kCoopThreads * k;
    if ( r < nrow ) __stcg( ymine + r, yacc[k] );
  }
}

// ypart: S * n doubles, yz: 2 * n doubles of global scratch owned by the group
__device__ inline void coop_hess( const CoopGroup& cg, int n, int ilo, int ihi, double* A, int lda, double* tau, double* sm,
                                  double* ypart, double* yz ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = cg.size, rank = cg.rank;
  double* vprev = sm;
  double* ypp = sm + n;
  double* vcur = sm + 2 * n;
  double* zpp = sm + 3 * n;
  double* ycur = sm + 4 * n;
  double* zcur = sm + 5 * n;
  double* red = sm + 6 * n;
  double* zw = red + 64;
  const int maxc = n / S + 2;
  double* yfull = yz;
  double* zfull = yz + n;
  const int nrow = ihi + 1;
  const int kcnt = ( nrow + kCoopThreads - 1 ) / kCoopThreads;
  for ( int r = tid; r < n; r += kCoopThreads ) {
    vprev[r] = 0.0;
    ypp[r] = 0.0;
    zpp[r] = 0.0;
    vcur[r] = 0.0;
  }
  if ( rank == 0 )
    for ( int j = tid; j < n; j += kCoopThreads )
      if ( j < ilo || j >= ihi ) tau[j] = 0.0;
  __syncthreads();
  bool have_prev = false;
  for ( int j = ilo; j <= ihi - 1; ++j ) {
    // (a) column j up to date with respect to step j-1, in shared memory (every CTA)
    double* colj = ycur;
    {
      const double zj = zpp[j];
      const double* src = A + (size_t)j * lda;
      for ( int r = tid; r < nrow; r += kCoopThreads ) colj[r] = __ldcg( src + r ) - ( vprev[r] * zj + ypp[r] );
    }
    __syncthreads();
    // (b) reflector from column j, rows j+1..ihi
    double tj = 0.0, beta;
    {
      const double alpha = colj[j + 1];
      double mx = 0.0;
      for ( int r = j + 2 + tid; r <= ihi; r += kCoopThreads ) mx = fmax( mx, fabs( colj[r] ) );
      mx = coop_cta_max( mx, red );
      beta = alpha;
      double sc = 0.0;
      if ( mx != 0.0 ) {
        double ss = 0.0;
        for ( int r = j + 2 + tid; r <= ihi; r += kCoopThreads ) {
          const double x = colj[r] / mx;
          ss += x * x;
        }
        double xnorm = mx * sqrt( coop_cta_sum( ss, red ) );
        double al = alpha;
        beta = -dsign( dlapy2( al, xnorm ), al );
        const double safmin = kSafmin / kEps;
        int knt = 0;
        double pre = 1.0;
        if ( fabs( beta ) < safmin ) {
          const double rsafmn = 1.0 / safmin;
          do {
            ++knt;
            pre *= rsafmn;
            xnorm *= rsafmn;
            beta *= rsafmn;
            al *= rsafmn;
          } while ( fabs( beta ) < safmin && knt < 20 );
          beta = -dsign( dlapy2( al, xnorm ), al );
        }
        tj = ( beta - al ) / beta;
        sc = pre / ( al - beta );
        for ( int q = 0; q < knt; ++q ) beta *= safmin;
      }
      __syncthreads();
      const bool owner = ( j % S ) == rank;
      double* dst = A + (size_t)j * lda;
      for ( int r = tid; r < n; r += kCoopThreads ) {
        double v = 0.0;
        if ( r == j + 1 ) v = 1.0;
        else if ( r >= j + 2 && r <= ihi ) v = colj[r] * sc;
        vcur[r] = v;
        if ( owner && r < nrow ) __stcg( dst + r, r <= j ? colj[r] : ( r == j + 1 ? beta : v ) );
      }
      if ( owner && tid == 0 ) tau[j] = tj;
    }
    __syncthreads();
    // (c) sweep over this CTA's columns c > j: apply step j-1, accumulate y_j (rows, registers) and z_j (columns)
    int c0 = j + 1;
    {
      const int m = c0 % S;
      c0 += ( rank - m + S ) % S;
    }
    const int ncols = ( c0 < n ) ? ( n - 1 - c0 ) / S + 1 : 0;
    if ( kcnt <= 3 ) coop_sweep<3, 8>( nrow, S, c0, ncols, ihi, have_prev, A, lda, vprev, ypp, vcur, zpp, zw, maxc, ypart + (size_t)rank * n );
    else if ( kcnt <= 6 ) coop_sweep<6, 4>( nrow, S, c0, ncols, ihi, have_prev, A, lda, vprev, ypp, vcur, zpp, zw, maxc, ypart + (size_t)rank * n );
    else coop_sweep<9, 4>( nrow, S, c0, ncols, ihi, have_prev, A, lda, vprev, ypp, vcur, zpp, zw, maxc, ypart + (size_t)rank * n );
    __syncthreads();
    // (d) partial results to global memory, reduce-scatter over the group
    for ( int ci = tid; ci < ncols; ci += kCoopThreads ) {
      double z = 0.0;
#pragma unroll
      for ( int w = 0; w < kCoopWarps; ++w ) z += zw[w * maxc + ci];
      __stcg( zfull + c0 + ci * S, z );
    }
    coop_sync( cg );
    {
      const int rows_per = ( nrow + S - 1 ) / S;
      const int r0 = rank * rows_per, r1 = imin( nrow, r0 + rows_per );
      if ( S > 32 ) {
        // few rows, many partial vectors: a warp per row, lanes over the partials
        for ( int r = r0 + warp; r < r1; r += kCoopWarps ) {
          double s = 0.0;
          for ( int p = lane; p < S; p += 32 ) s += __ldcg( ypart + (size_t)p * n + r );
          s = warp_sum( s );
          if ( lane == 0 ) __stcg( yfull + r, s );
        }
      } else {
        // many rows, few partial vectors: a thread per row (coalesced), the S loads of a row in flight together
        for ( int r = r0 + tid; r < r1; r += kCoopThreads ) {
          double s = 0.0;
          for ( int p = 0; p < S; ++p ) s += __ldcg( ypart + (size_t)p * n + r );
          __stcg( yfull + r, s );
        }
      }
    }
    coop_sync( cg );
    double av = 0.0;
    for ( int r = tid; r < n; r += kCoopThreads ) {
      const double y = ( r < nrow ) ? __ldcg( yfull + r ) : 0.0;
      ycur[r] = y;
      zcur[r] = ( r > j ) ? __ldcg( zfull + r ) : 0.0;
      av += vcur[r] * y;
    }
    const double alpha2 = coop_cta_sum( av, red );
    const double half = 0.5 * tj * tj * alpha2;
    __syncthreads();
    for ( int r = tid; r < n; r += kCoopThreads ) {
      const double v = vcur[r];
      vprev[r] = v;
      ypp[r] = ( r < nrow ) ? tj * ycur[r] - half * v : 0.0;
      zpp[r] = ( r > j ) ? tj * zcur[r] - half * v : 0.0;
    }
    have_prev = true;
    __syncthreads();
  }
}

// explicit Q, columns dealt round-robin to the CTAs of the group, warps take the CTA's columns round-robin. The
// reflectors are applied kCoopQGroup at a time in compact WY form, H(jlo) ... H(jhi) = I - V T V^T (dlarft forward /
// columnwise): V (4 vectors, zero outside their support) and T are staged in shared memory (sm: 4 n + 64 doubles), a
// column then needs two passes over its live part per GROUP instead of per reflector (u = V^T q; q -= V (T u)).
constexpr int kCoopQGroup = 4;

__device__ inline void coop_form_q( const CoopGroup& cg, int n, int ilo, int ihi, const double* A, int lda, const double* tau,
                                    double* Q, int ldq, double* sm ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = cg.size, rank = cg.rank;
  double* vs = sm;                          // kCoopQGroup x n
  double* red = sm + kCoopQGroup * n;       // 64 doubles of reduction scratch
  __shared__ double Ts[kCoopQGroup * kCoopQGroup];
  for ( int c = rank + warp * S; c < n; c += kCoopWarps * S )
    for ( int r = lane; r < n; r += 32 ) __stcg( Q + r + (size_t)c * ldq, ( r == c ) ? 1.0 : 0.0 );
  __syncthreads();
  for ( int jhi = ihi - 2; jhi >= ilo; jhi -= kCoopQGroup ) {
    const int ng = imin( kCoopQGroup, jhi - ilo + 1 );
    const int jlo = jhi - ng + 1;
    // stage the group's vectors (ascending reflector index), rows jlo+1..ihi
    for ( int g = 0; g < kCoopQGroup; ++g ) {
      const int j = jlo + g;
      for ( int r = jlo + 1 + tid; r <= ihi; r += kCoopThreads )
        vs[g * n + r] = ( g >= ng || r <= j ) ? 0.0 : ( r == j + 1 ? 1.0 : __ldcg( A + r + (size_t)j * lda ) );
    }
    __syncthreads();
    // T: T(i,i) = tau_i, T(0:i, i) = -tau_i T(0:i,0:i) (V(:,0:i)^T v_i)
    double gram[kCoopQGroup][kCoopQGroup];
#pragma unroll
    for ( int a = 0; a < kCoopQGroup; ++a )
#pragma unroll
      for ( int b = 0; b < kCoopQGroup; ++b ) gram[a][b] = 0.0;
#pragma unroll
    for ( int b = 1; b < kCoopQGroup; ++b )
#pragma unroll
      for ( int a = 0; a < b; ++a ) {
        double p = 0.0;
        if ( b < ng )
          for ( int r = jlo + 1 + tid; r <= ihi; r += kCoopThreads ) p += vs[a * n + r] * vs[b * n + r];
        gram[a][b] = coop_cta_sum( p, red );
      }
    if ( tid == 0 ) {
      double T[kCoopQGroup][kCoopQGroup];
      for ( int a = 0; a < kCoopQGroup; ++a )
        for ( int b = 0; b < kCoopQGroup; ++b ) T[a][b] = 0.0;
      for ( int i = 0; i < ng; ++i ) {
        const double ti = __ldcg( tau + jlo + i );
        T[i][i] = ti;
        for ( int a = 0; a < i; ++a ) {
          double acc = 0.0;
          for ( int b = a; b < i; ++b ) acc += T[a][b] * gram[b][i];
          T[a][i] = -ti * acc;
        }
      }
      for ( int a = 0; a < kCoopQGroup; ++a )
        for ( int b = 0; b < kCoopQGroup; ++b ) Ts[a * kCoopQGroup + b] = T[a][b];
    }
    __syncthreads();
    const int nr = ihi - jlo;  // rows jlo+1..ihi
    // this CTA's columns in jlo+1..ihi: c = rank (mod S); its oi-th column overall goes to warp oi % warps
    int oi = ( jlo + 1 - rank + S - 1 ) / S;
    if ( oi < 0 ) oi = 0;
    oi += ( warp - ( oi % kCoopWarps ) + kCoopWarps ) % kCoopWarps;
    for ( int c = rank + oi * S; c <= ihi; c += kCoopWarps * S ) {
      double* col = Q + (size_t)c * ldq + ( jlo + 1 );
      const double* v = vs + ( jlo + 1 );
      double u[kCoopQGroup];
#pragma unroll
      for ( int g = 0; g < kCoopQGroup; ++g ) u[g] = 0.0;
      for ( int r0 = 0; r0 < nr; r0 += 256 ) {
        double q[8];
#pragma unroll
        for ( int k = 0; k < 8; ++k ) q[k] = __ldcg( col + imin( r0 + lane + 32 * k, nr - 1 ) );
#pragma unroll
        for ( int k = 0; k < 8; ++k ) {
          const int r = r0 + lane + 32 * k;
          if ( r < nr ) {
#pragma unroll
            for ( int g = 0; g < kCoopQGroup; ++g ) u[g] += v[g * n + r] * q[k];
          }
        }
      }
#pragma unroll
      for ( int g = 0; g < kCoopQGroup; ++g ) u[g] = warp_sum( u[g] );
      double w[kCoopQGroup];
#pragma unroll
      for ( int a = 0; a < kCoopQGroup; ++a ) {
        double acc = 0.0;
#pragma unroll
        for ( int b = 0; b < kCoopQGroup; ++b )
          if ( b >= a ) acc += Ts[a * kCoopQGroup + b] * u[b];
        w[a] = acc;
      }
      for ( int r0 = 0; r0 < nr; r0 += 256 ) {
        double q[8];
#pragma unroll
        for ( int k = 0; k < 8; ++k ) q[k] = __ldcg( col + imin( r0 + lane + 32 * k, nr - 1 ) );
#pragma unroll
        for ( int k = 0; k < 8; ++k ) {
          const int r = r0 + lane + 32 * k;
          if ( r < nr ) {
            double x = q[k];
#pragma unroll
            for ( int g = 0; g < kCoopQGroup; ++g ) x -= v[g * n + r] * w[g];
            __stcg( col + r, x );
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace pb

#endif  // __CUDACC__
