// pb_hess.cuh -- Hessenberg reduction and explicit Q for one matrix per CTA (device only).
//
// Replaces the dgehrd / dorghr stages of the reference's dgeev_ call (reference
// src/lapack-wrapper.hpp:47-60, src/paladin.hpp:248-254). Same Householder reflectors as LAPACK's
// unblocked dgehd2 (v(j+1) = 1, stored below the subdiagonal, tau[j]); the schedule is different:
//
// hess_fused: the matrix (n^2 doubles, 2 MiB at n = 512) lives in global memory / L2 and does not fit
//   on one SM, so the reduction is bound by how often it streams past the SM. Step j needs
//   y = A v and z = A^T v over the whole trailing matrix before it can update it (3 sweeps per column
//   in the textbook form). Here the rank-2 form  A <- A - v z'^T - y' v^T  (y' = tau y - (tau^2 a/2) v,
//   z' = tau z - (tau^2 a/2) v, a = v^T A v) of step j-1 is applied IN THE SAME sweep that accumulates
//   y and z of step j: column j is brought up to date first (look-ahead), its reflector is generated,
//   then every trailing element is read once, updated, written once and immediately used for both
//   products. One read + one write of the trailing matrix per column: 16 n^3 / 2 bytes in total.
//   A warp owns a column at a time (lanes along rows: coalesced 256-byte accesses), z(c) is a warp
//   shuffle reduction, y(r) accumulates in registers across the warp's columns and is combined across
//   warps once per step through shared memory.
//
// form_q: column c of Q = H(ilo) ... H(ihi-2) only depends on reflectors j < c. A warp builds one column
//   entirely in registers (start from e_c, apply j = c-1 ... ilo), and stores it once: no barriers, no
//   intermediate traffic; the reflector vectors are re-read from L1/L2.
#pragma once

#include "pb_common.cuh"

#if defined( __CUDACC__ )

namespace pb {

constexpr int kHessThreads = 256;
constexpr int kHessWarps = kHessThreads / 32;

// shared memory (doubles): vprev, ypp, vcur, zpp, ycur, zcur : 6 n ; ypart: kHessWarps * n ; red: 64
__host__ __device__ inline size_t hess_smem_doubles( int n ) { return (size_t)( 6 + kHessWarps ) * n + 64; }  // >= form_q's kFormQGroup*32*MAXCH + kFormQGroup

__device__ __forceinline__ double cta_sum( double x, double* red ) {
  x = warp_sum( x );
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ( ( threadIdx.x & 31 ) == 0 ) red[w] = x;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for ( int i = 0; i < kHessWarps; ++i ) t += red[i];
  return t;
}
__device__ __forceinline__ double cta_max( double x, double* red ) {
  x = warp_max( x );
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ( ( threadIdx.x & 31 ) == 0 ) red[w] = x;
  __syncthreads();
  double t = red[0];
#pragma unroll
  for ( int i = 1; i < kHessWarps; ++i ) t = fmax( t, red[i] );
  return t;
}

// MAXCH: row chunks of 32 a lane may own (n <= 32 * MAXCH)
template <int MAXCH>
__device__ void hess_fused( int n, int ilo, int ihi, double* A, int lda, double* tau, double* sm ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* vprev = sm;
  double* ypp = sm + n;
  double* vcur = sm + 2 * n;
  double* zpp = sm + 3 * n;
  double* ycur = sm + 4 * n;
  double* zcur = sm + 5 * n;
  double* ypart = sm + 6 * n;
  double* red = sm + ( 6 + kHessWarps ) * n;
  const int nrow = ihi + 1;                    // rows 0..ihi take part
  const int nch = ( nrow + 31 ) >> 5;

  for ( int r = tid; r < n; r += kHessThreads ) {
    vprev[r] = 0.0;
    ypp[r] = 0.0;
    zpp[r] = 0.0;
    vcur[r] = 0.0;
  }
  for ( int j = tid; j < n; j += kHessThreads )
    if ( j < ilo || j >= ihi ) tau[j] = 0.0;  // only ilo..ihi-1 are generated below (ihi-1 is trivially 0)
  __syncthreads();
  bool have_prev = false;

  for ( int j = ilo; j <= ihi - 1; ++j ) {
    // (a) column j up to date with respect to step j-1 (v_{j-1}(j) = 1)
    if ( have_prev ) {
      const double zj = zpp[j];
      for ( int r = tid; r < nrow; r += kHessThreads ) A[r + (size_t)j * lda] -= vprev[r] * zj + ypp[r];
    }
    __syncthreads();
    // (b) reflector from A(j+1:ihi, j)
    double tj = 0.0, beta;
    {
      const double alpha = A[( j + 1 ) + (size_t)j * lda];
      double mx = 0.0;
      for ( int r = j + 2 + tid; r <= ihi; r += kHessThreads ) mx = fmax( mx, fabs( A[r + (size_t)j * lda] ) );
      mx = cta_max( mx, red );
      beta = alpha;
      double sc = 0.0;
      if ( mx != 0.0 ) {
        double ss = 0.0;
        for ( int r = j + 2 + tid; r <= ihi; r += kHessThreads ) {
          const double x = A[r + (size_t)j * lda] / mx;
          ss += x * x;
        }
        double xnorm = mx * sqrt( cta_sum( ss, red ) );
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
      // v_j into shared memory (0 outside j+1..ihi) and below the subdiagonal of A
      for ( int r = tid; r < n; r += kHessThreads ) {
        double v = 0.0;
        if ( r == j + 1 ) v = 1.0;
        else if ( r >= j + 2 && r <= ihi ) {
          v = A[r + (size_t)j * lda] * sc;
          A[r + (size_t)j * lda] = v;
        }
        vcur[r] = v;
      }
      if ( tid == 0 ) {
        A[( j + 1 ) + (size_t)j * lda] = beta;
        tau[j] = tj;
      }
    }
    __syncthreads();

    // (c) one sweep over columns j+1..n-1: apply step j-1, accumulate y_j (rows) and z_j (columns)
    double yacc[MAXCH];
#pragma unroll
    for ( int k = 0; k < MAXCH; ++k ) yacc[k] = 0.0;
    for ( int c = j + 1 + warp; c < n; c += kHessWarps ) {
      double* col = A + (size_t)c * lda;
      const double zc = zpp[c], vpc = vprev[c];
      const double vcc = ( c <= ihi ) ? vcur[c] : 0.0;
      // all loads of the column first (clamped addresses, no branches), so that they are in flight together
      double a[MAXCH];
#pragma unroll
      for ( int k = 0; k < MAXCH; ++k ) a[k] = __ldcg( col + imin( lane + 32 * k, nrow - 1 ) );
      double zsum = 0.0;
#pragma unroll
      for ( int k = 0; k < MAXCH; ++k ) {
        const int r = lane + 32 * k;
        const int rc = imin( r, nrow - 1 );
        const bool ok = r < nrow;
        const double x = a[k] - ( vprev[rc] * zc + ypp[rc] * vpc );
        a[k] = x;
        yacc[k] += ok ? x * vcc : 0.0;
        zsum += ok ? vcur[rc] * x : 0.0;
      }
      if ( have_prev ) {
#pragma unroll
        for ( int k = 0; k < MAXCH; ++k ) {
          const int r = lane + 32 * k;
          if ( r < nrow ) col[r] = a[k];
        }
      }
      zsum = warp_sum( zsum );
      if ( lane == 0 ) zcur[c] = zsum;
    }
    // (d) combine: y = sum over warps, alpha = v^T y, then y', z'
#pragma unroll
    for ( int k = 0; k < MAXCH; ++k ) {
      const int r = lane + 32 * k;
      if ( k < nch && r < nrow ) ypart[warp * n + r] = yacc[k];
    }
    __syncthreads();
    double av = 0.0;
    for ( int r = tid; r < nrow; r += kHessThreads ) {
      double y = 0.0;
#pragma unroll
      for ( int w = 0; w < kHessWarps; ++w ) y += ypart[w * n + r];
      ycur[r] = y;
      av += vcur[r] * y;
    }
    const double alpha2 = cta_sum( av, red );
    const double half = 0.5 * tj * tj * alpha2;
    __syncthreads();
    for ( int r = tid; r < n; r += kHessThreads ) {
      const double v = vcur[r];
      vprev[r] = v;
      ypp[r] = ( r < nrow ) ? tj * ycur[r] - half * v : 0.0;
      zpp[r] = ( r > j ) ? tj * zcur[r] - half * v : 0.0;
    }
    have_prev = true;
    __syncthreads();
  }
  // the last generated step is j = ihi-1 with tau = 0: nothing is left to apply
}

// Q (n x n, ldq) = H(ilo) ... H(ihi-2) from the reflectors stored in A / tau, by backward accumulation:
// for j = ihi-2 ... ilo the block Q(j+1:ihi, j+1:ihi) <- (I - tau v v^T) Q(...). Same streaming pattern as the
// reduction itself: a warp owns a column at a time (loads it once into registers, stores it once); kFormQGroup
// consecutive reflectors are applied per sweep from shared memory (their vectors are staged there, zero outside their
// support), which divides the traffic by kFormQGroup: every later reflector of the group only needs the
// already-updated column, which is still in registers.
constexpr int kFormQGroup = 8;

template <int MAXCH>
__device__ void form_q( int n, int ilo, int ihi, const double* __restrict__ A, int lda, const double* __restrict__ tau,
                        double* __restrict__ Q, int ldq, double* sm ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* vs = sm;                               // kFormQGroup vectors of length 32 * MAXCH
  double* ts = sm + kFormQGroup * 32 * MAXCH;    // their tau
  // identity
  for ( int c = warp; c < n; c += kHessWarps )
    for ( int r = lane; r < n; r += 32 ) Q[r + (size_t)c * ldq] = ( r == c ) ? 1.0 : 0.0;
  __syncthreads();
  int j = ihi - 2;
  while ( j >= ilo ) {
    const int ng = imin( kFormQGroup, j - ilo + 1 );   // reflectors j, j-1, ..., j-ng+1, applied in this order
    const int jlast = j - ng + 1;
    for ( int q = tid; q < ng * 32 * MAXCH; q += kHessThreads ) {
      const int g = q / ( 32 * MAXCH ), r = q - g * ( 32 * MAXCH );
      const int jj = j - g;
      double v = 0.0;
      if ( r == jj + 1 ) v = 1.0;
      else if ( r >= jj + 2 && r <= ihi ) v = A[r + (size_t)jj * lda];
      vs[q] = v;
    }
    if ( tid < ng ) ts[tid] = tau[j - tid];
    __syncthreads();
    const int k0 = ( jlast + 1 ) >> 5;  // first 32-row chunk that any reflector of the group touches
    // columns jlast+1 .. ihi (a column left of a reflector's support is simply not changed by it: v^T q = 0)
    for ( int c = jlast + 1 + warp; c <= ihi; c += kHessWarps ) {
      double* col = Q + (size_t)c * ldq;
      double q[MAXCH];
#pragma unroll
      for ( int k = 0; k < MAXCH; ++k ) q[k] = ( k >= k0 ) ? col[imin( lane + 32 * k, n - 1 )] : 0.0;
      for ( int g = 0; g < ng; ++g ) {
        const double* v = vs + g * 32 * MAXCH + lane;
        double w = 0.0;
#pragma unroll
        for ( int k = 0; k < MAXCH; ++k ) w += v[32 * k] * q[k];
        w = warp_sum( w ) * ts[g];
#pragma unroll
        for ( int k = 0; k < MAXCH; ++k ) q[k] -= w * v[32 * k];
      }
#pragma unroll
      for ( int k = 0; k < MAXCH; ++k ) {
        const int r = lane + 32 * k;
        if ( k >= k0 && r <= ihi ) col[r] = q[k];
      }
    }
    __syncthreads();
    j -= ng;
  }
}

// ---- any n (n <= ~4700: 6n + 64 doubles of shared memory): the same single-sweep reduction with the row
// accumulators of A v in shared memory (atomicAdd) instead of registers, row chunks walked eight at a time ------
__host__ __device__ inline size_t hess_big_smem_doubles( int n ) { return (size_t)6 * n + 64; }

__device__ inline void hess_fused_big( int n, int ilo, int ihi, double* A, int lda, double* tau, double* sm ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* vprev = sm;
  double* ypp = sm + n;
  double* vcur = sm + 2 * n;
  double* zpp = sm + 3 * n;
  double* ycur = sm + 4 * n;
  double* zcur = sm + 5 * n;
  double* red = sm + 6 * n;
  const int nrow = ihi + 1;
  const int nch = ( nrow + 31 ) >> 5;
  for ( int r = tid; r < n; r += kHessThreads ) {
    vprev[r] = 0.0;
    ypp[r] = 0.0;
    zpp[r] = 0.0;
    vcur[r] = 0.0;
    ycur[r] = 0.0;
  }
  for ( int j = tid; j < n; j += kHessThreads )
    if ( j < ilo || j >= ihi ) tau[j] = 0.0;
  __syncthreads();
  bool have_prev = false;
  for ( int j = ilo; j <= ihi - 1; ++j ) {
    if ( have_prev ) {
      const double zj = zpp[j];
      for ( int r = tid; r < nrow; r += kHessThreads ) A[r + (size_t)j * lda] -= vprev[r] * zj + ypp[r];
    }
    __syncthreads();
    double tj = 0.0, beta;
    {
      const double alpha = A[( j + 1 ) + (size_t)j * lda];
      double mx = 0.0;
      for ( int r = j + 2 + tid; r <= ihi; r += kHessThreads ) mx = fmax( mx, fabs( A[r + (size_t)j * lda] ) );
      mx = cta_max( mx, red );
      beta = alpha;
      double sc = 0.0;
      if ( mx != 0.0 ) {
        double ss = 0.0;
        for ( int r = j + 2 + tid; r <= ihi; r += kHessThreads ) {
          const double x = A[r + (size_t)j * lda] / mx;
          ss += x * x;
        }
        double xnorm = mx * sqrt( cta_sum( ss, red ) );
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
      for ( int r = tid; r < n; r += kHessThreads ) {
        double v = 0.0;
        if ( r == j + 1 ) v = 1.0;
        else if ( r >= j + 2 && r <= ihi ) {
          v = A[r + (size_t)j * lda] * sc;
          A[r + (size_t)j * lda] = v;
        }
        vcur[r] = v;
      }
      if ( tid == 0 ) {
        A[( j + 1 ) + (size_t)j * lda] = beta;
        tau[j] = tj;
      }
    }
    __syncthreads();
    for ( int c = j + 1 + warp; c < n; c += kHessWarps ) {
      double* col = A + (size_t)c * lda;
      const double zc = zpp[c], vpc = vprev[c];
      const double vcc = ( c <= ihi ) ? vcur[c] : 0.0;
      double zsum = 0.0;
      for ( int kb = 0; kb < nch; kb += 8 ) {
        double a[8];
#pragma unroll
        for ( int u = 0; u < 8; ++u ) a[u] = __ldcg( col + imin( lane + 32 * ( kb + u ), nrow - 1 ) );
#pragma unroll
        for ( int u = 0; u < 8; ++u ) {
          const int r = lane + 32 * ( kb + u );
          const int rc = imin( r, nrow - 1 );
          const bool ok = r < nrow;
          const double x = a[u] - ( vprev[rc] * zc + ypp[rc] * vpc );
          a[u] = x;
          if ( ok && vcc != 0.0 ) atomicAdd( ycur + r, x * vcc );
          zsum += ok ? vcur[rc] * x : 0.0;
        }
        if ( have_prev ) {
#pragma unroll
          for ( int u = 0; u < 8; ++u ) {
            const int r = lane + 32 * ( kb + u );
            if ( r < nrow ) col[r] = a[u];
          }
        }
      }
      zsum = warp_sum( zsum );
      if ( lane == 0 ) zcur[c] = zsum;
    }
    __syncthreads();
    double av = 0.0;
    for ( int r = tid; r < nrow; r += kHessThreads ) av += vcur[r] * ycur[r];
    const double alpha2 = cta_sum( av, red );
    const double half = 0.5 * tj * tj * alpha2;
    __syncthreads();
    for ( int r = tid; r < n; r += kHessThreads ) {
      const double v = vcur[r];
      vprev[r] = v;
      ypp[r] = ( r < nrow ) ? tj * ycur[r] - half * v : 0.0;
      zpp[r] = ( r > j ) ? tj * zcur[r] - half * v : 0.0;
      ycur[r] = 0.0;
    }
    have_prev = true;
    __syncthreads();
  }
}

// explicit Q for any n: one reflector per sweep, v in shared memory (sm: n doubles), two passes over each column
// (dot product, update; the second pass hits L1/L2)
__device__ inline void form_q_big( int n, int ilo, int ihi, const double* __restrict__ A, int lda, const double* __restrict__ tau,
                                   double* __restrict__ Q, int ldq, double* sm ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* vs = sm;
  for ( int c = warp; c < n; c += kHessWarps )
    for ( int r = lane; r < n; r += 32 ) Q[r + (size_t)c * ldq] = ( r == c ) ? 1.0 : 0.0;
  __syncthreads();
  for ( int j = ihi - 2; j >= ilo; --j ) {
    const double tj = tau[j];
    if ( tj == 0.0 ) continue;
    for ( int r = j + 1 + tid; r <= ihi; r += kHessThreads ) vs[r] = ( r == j + 1 ) ? 1.0 : A[r + (size_t)j * lda];
    __syncthreads();
    const int nr = ihi - j;  // rows j+1..ihi
    for ( int c = j + 1 + warp; c <= ihi; c += kHessWarps ) {
      double* col = Q + (size_t)c * ldq + ( j + 1 );
      const double* v = vs + ( j + 1 );
      double w = 0.0;
      for ( int r0 = 0; r0 < nr; r0 += 256 ) {
        double q[8];
#pragma unroll
        for ( int u = 0; u < 8; ++u ) q[u] = col[imin( r0 + lane + 32 * u, nr - 1 )];
#pragma unroll
        for ( int u = 0; u < 8; ++u ) {
          const int r = r0 + lane + 32 * u;
          w += ( r < nr ) ? v[imin( r, nr - 1 )] * q[u] : 0.0;
        }
      }
      w = warp_sum( w ) * tj;
      for ( int r0 = 0; r0 < nr; r0 += 256 ) {
        double q[8];
#pragma unroll
        for ( int u = 0; u < 8; ++u ) q[u] = col[imin( r0 + lane + 32 * u, nr - 1 )];
#pragma unroll
        for ( int u = 0; u < 8; ++u ) {
          const int r = r0 + lane + 32 * u;
          if ( r < nr ) col[r] = q[u] - w * v[r];
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace pb

#endif  // __CUDACC__
