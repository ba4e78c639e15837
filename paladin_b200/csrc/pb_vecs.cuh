// pb_vecs.cuh -- right and left eigenvectors for one matrix per CTA (device only), the dtrevc3 'R'/'L','B' + dgebak 'B' +
// normalisation tail of the reference's dgeev_ call (reference src/lapack-wrapper.hpp:47-60,
// src/paladin.hpp:248-254; the packed layout is consumed at src/paladin.hpp:256-333).
//
//   vecs_solve     : every warp back-substitutes its own eigenvector(s) of the quasi-triangular T
//                    (pb::trevc_right_solve instantiated for a warp), all n solves of a matrix in flight
//                    on the CTA's warps; the triangular factor X (n x n, upper) goes to a work buffer.
//   vecs_gemm      : V_raw = Z X on the FP64 tensor pipe (mma.sync m8n8k4 f64), tiles of Z and X staged
//                    through shared memory; k runs only over the non-zero part of each X column block.
//   vecs_finish    : one warp per column (pair): undo balancing (row scaling + permutation), unit 2-norm,
//                    rotate complex pairs so that the largest component is real; single read, single write.
#pragma once

#include "pb_common.cuh"
#include "pb_vec.cuh"

#if defined( __CUDACC__ )

namespace pb {

constexpr int kVecThreads = 256;
constexpr int kVecWarps = kVecThreads / 32;

// ---- 1. triangular solves ---------------------------------------------------------------------------
// X(:, ki) (real) or X(:, ki-1), X(:, ki) (pair, real and imaginary parts) <- eigenvector of T; entries
// below the block are zeroed so that the GEMM may treat X as dense.
// part / nparts: this CTA is one of nparts that share the matrix (large matrices, big_vec_kernel): it takes every
// nparts-th group of kVecWarps blocks; every part computes the (identical) column norms itself.
__device__ inline void vecs_solve( int n, const double* T, int ldt, double* X, int ldx, double* cnorm, bool left, int part = 0,
                                   int nparts = 1 ) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int stride = kVecWarps * nparts, mine = part * kVecWarps + warp;
  // column 1-norms of the strictly upper part (warp per column, lanes along rows)
  for ( int j = warp; j < n; j += kVecWarps ) {
    double s = 0.0;
    for ( int i = lane; i < j; i += 32 ) s += fabs( T[i + (size_t)j * ldt] );
    s = warp_sum( s );
    if ( lane == 0 ) cnorm[j] = s;
  }
  __syncthreads();
  WarpGroup g;
  // every warp walks the same block list and takes every kVecWarps-th block (the expensive blocks first)
  int blk = 0;
  if ( !left ) {
    int ki = n - 1;
    while ( ki >= 0 ) {
      const bool cplx = ( ki > 0 && T[ki + (size_t)( ki - 1 ) * ldt] != 0.0 );
      if ( ( blk % stride ) == mine ) {
        if ( !cplx ) {
          double* xr = X + (size_t)ki * ldx;
          trevc_right_solve( g, n, T, ldt, ki, false, cnorm, xr, xr );
          for ( int r = ki + 1 + lane; r < n; r += 32 ) xr[r] = 0.0;
        } else {
          double* xr = X + (size_t)( ki - 1 ) * ldx;
          double* xi = X + (size_t)ki * ldx;
          trevc_right_solve( g, n, T, ldt, ki, true, cnorm, xr, xi );
          for ( int r = ki + 1 + lane; r < n; r += 32 ) {
            xr[r] = 0.0;
            xi[r] = 0.0;
          }
        }
      }
      ki -= cplx ? 2 : 1;
      blk += 1;
    }
  } else {
    int ki = 0;
    while ( ki < n ) {
      const bool cplx = ( ki < n - 1 && T[( ki + 1 ) + (size_t)ki * ldt] != 0.0 );
      if ( ( blk % stride ) == mine ) {
        double* xr = X + (size_t)ki * ldx;
        double* xi = cplx ? X + (size_t)( ki + 1 ) * ldx : xr;
        trevc_left_solve( g, n, T, ldt, ki, cplx, cnorm, xr, xi );
        for ( int r = lane; r < ki; r += 32 ) {
          xr[r] = 0.0;
          if ( cplx ) xi[r] = 0.0;
        }
      }
      ki += cplx ? 2 : 1;
      blk += 1;
    }
  }
  __syncthreads();
}

// ---- 2. C = Z X, X upper (quasi-)triangular -----------------------------------------------------------
constexpr int kGemmBM = 128, kGemmBN = 64, kGemmBK = 16;
constexpr int kGemmLdA = kGemmBM + 4;  // As[k][m], padded: the 4 k-rows of a fragment hit different bank groups
constexpr int kGemmLdB = kGemmBK + 4;  // Bs[n][k], same idea for the 4 columns x 4 k of a half warp
__host__ __device__ inline size_t vecs_gemm_smem_doubles() { return (size_t)kGemmBK * kGemmLdA + (size_t)kGemmBN * kGemmLdB; }

__device__ __forceinline__ void dmma_m8n8k4( double& c0, double& c1, double a, double b ) {
  asm volatile( "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"( c0 ), "+d"( c1 )
                : "d"( a ), "d"( b ) );
}

// C (n x n, ldc) = Z (n x n, ldz) * X (n x n, ldx); column block [c0, c0+BN) of X is zero below row c0+BN
// lower = true: X is lower (quasi-)triangular (left eigenvectors): column block [n0, n0+BN) is zero above row n0-1
__device__ inline void vecs_gemm( int n, const double* __restrict__ Z, int ldz, const double* __restrict__ X, int ldx,
                                  double* __restrict__ C, int ldc, double* sm, bool lower, int part = 0, int nparts = 1 ) {
  double* As = sm;                                 // [BK][LdA]
  double* Bs = sm + (size_t)kGemmBK * kGemmLdA;    // [BN][LdB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;         // 4 x 2 warps, each 32 x 32
  const int gid = lane >> 2, tig = lane & 3;
  for ( int n0 = 0; n0 < n; n0 += kGemmBN ) {
    const int kend = lower ? n : imin( n, n0 + kGemmBN + 1 );
    const int kbeg = lower ? ( imax( 0, n0 - 1 ) / kGemmBK ) * kGemmBK : 0;
    for ( int m0 = 0; m0 < n; m0 += kGemmBM ) {
      if ( nparts > 1 && ( ( n0 / kGemmBN ) * ( ( n + kGemmBM - 1 ) / kGemmBM ) + m0 / kGemmBM ) % nparts != part ) continue;
      double acc[4][4][2];
#pragma unroll
      for ( int i = 0; i < 4; ++i )
#pragma unroll
        for ( int j = 0; j < 4; ++j ) acc[i][j][0] = acc[i][j][1] = 0.0;
      for ( int k0 = kbeg; k0 < kend; k0 += kGemmBK ) {
        // stage Z(m0:m0+BM, k0:k0+BK) and X(k0:k0+BK, n0:n0+BN), zero padded
        double za[8], xb[4];
#pragma unroll
        for ( int u = 0; u < 8; ++u ) {
          const int q = tid + u * kVecThreads;  // 0..2047 = BK*BM
          const int kk = q >> 7, mm = q & 127;
          const int gr = m0 + mm, gk = k0 + kk;
          za[u] = ( gr < n && gk < kend ) ? Z[gr + (size_t)gk * ldz] : 0.0;
        }
#pragma unroll
        for ( int u = 0; u < 4; ++u ) {
          const int q = tid + u * kVecThreads;  // 0..1023 = BN*BK
          const int nn = q >> 4, kk = q & 15;
          const int gc = n0 + nn, gk = k0 + kk;
          xb[u] = ( gc < n && gk < kend ) ? X[gk + (size_t)gc * ldx] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for ( int u = 0; u < 8; ++u ) {
          const int q = tid + u * kVecThreads;
          As[( q >> 7 ) * kGemmLdA + ( q & 127 )] = za[u];
        }
#pragma unroll
        for ( int u = 0; u < 4; ++u ) {
          const int q = tid + u * kVecThreads;
          Bs[( q >> 4 ) * kGemmLdB + ( q & 15 )] = xb[u];
        }
        __syncthreads();
#pragma unroll
        for ( int ks = 0; ks < kGemmBK; ks += 4 ) {
          double a[4], b[4];
#pragma unroll
          for ( int i = 0; i < 4; ++i ) a[i] = As[( ks + tig ) * kGemmLdA + wm * 32 + i * 8 + gid];
#pragma unroll
          for ( int j = 0; j < 4; ++j ) b[j] = Bs[( wn * 32 + j * 8 + gid ) * kGemmLdB + ks + tig];
#pragma unroll
          for ( int i = 0; i < 4; ++i )
#pragma unroll
            for ( int j = 0; j < 4; ++j ) dmma_m8n8k4( acc[i][j][0], acc[i][j][1], a[i], b[j] );
        }
      }
      // C fragment: rows gid, columns 2*tig, 2*tig+1 of each 8x8 tile
#pragma unroll
      for ( int i = 0; i < 4; ++i )
#pragma unroll
        for ( int j = 0; j < 4; ++j ) {
          const int r = m0 + wm * 32 + i * 8 + gid;
          const int c = n0 + wn * 32 + j * 8 + 2 * tig;
          if ( r < n ) {
            if ( c < n ) C[r + (size_t)c * ldc] = acc[i][j][0];
            if ( c + 1 < n ) C[r + (size_t)( c + 1 ) * ldc] = acc[i][j][1];
          }
        }
      __syncthreads();
    }
  }
}

// ---- 3. un-balance + normalise -------------------------------------------------------------------------
// perm: n ints of shared/global scratch (final row of each source row after dgebak's swaps)
template <int MAXCH>
__device__ void vecs_finish( int n, int ilo, int ihi, const double* scale, const double* wi, const double* __restrict__ C,
                             int ldc, double* __restrict__ V, int ldv, int* perm, bool left, int part = 0, int nparts = 1 ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int stride = kVecWarps * nparts, mine = part * kVecWarps + warp;
  // dgebak 'B','R': rows ilo..ihi are scaled, then row swaps are applied for i = ilo-1..0 and ihi+1..n-1.
  // dest[p] = row where source row p ends up.
  if ( tid == 0 ) {
    for ( int r = 0; r < n; ++r ) perm[r] = r;  // perm[position] = source row currently at that position
    for ( int ii = 0; ii < n; ++ii ) {
      int i = ii;
      if ( i >= ilo && i <= ihi ) continue;
      if ( i < ilo ) i = ilo - 1 - ii;
      const int k = (int)scale[i];
      if ( k == i ) continue;
      const int tmp = perm[i];
      perm[i] = perm[k];
      perm[k] = tmp;
    }
  }
  __syncthreads();
  const int nch = ( n + 31 ) >> 5;
  // every warp walks the same block list
  int blk = 0;
  int c = 0;
  while ( c < n ) {
    const bool pair = ( wi[c] > 0.0 && c + 1 < n );
    if ( ( blk % stride ) == mine ) {
      double a[MAXCH], b[MAXCH];
      // position-major: element k of this lane is destination row r = lane + 32 k, read from source row perm[r]
#pragma unroll
      for ( int k = 0; k < MAXCH; ++k ) {
        const int r = lane + 32 * k;
        a[k] = 0.0;
        b[k] = 0.0;
        if ( k < nch && r < n ) {
          const int src = perm[r];
          double s = ( src >= ilo && src <= ihi && ilo != ihi ) ? scale[src] : 1.0;
          if ( left ) s = 1.0 / s;  // dgebak 'L' divides
          a[k] = C[src + (size_t)c * ldc] * s;
          if ( pair ) b[k] = C[src + (size_t)( c + 1 ) * ldc] * s;
        }
      }
      // overflow-safe joint 2-norm
      double mx = 0.0;
#pragma unroll
      for ( int k = 0; k < MAXCH; ++k ) mx = fmax( mx, fmax( fabs( a[k] ), fabs( b[k] ) ) );
      mx = warp_max( mx );
      double scl = 0.0;
      if ( mx > 0.0 && mx <= DBL_MAX ) {
        const double inv = 1.0 / mx;
        double ss = 0.0;
#pragma unroll
        for ( int k = 0; k < MAXCH; ++k ) {
          const double p = a[k] * inv, q = b[k] * inv;
          ss += p * p + q * q;
        }
        ss = warp_sum( ss );
        scl = 1.0 / ( mx * sqrt( ss ) );
      }
#pragma unroll
      for ( int k = 0; k < MAXCH; ++k ) {
        a[k] *= scl;
        b[k] *= scl;
      }
      if ( pair ) {
        // component of largest modulus (first index on ties) becomes real
        double best = -1.0;
        int bestr = n;
#pragma unroll
        for ( int k = 0; k < MAXCH; ++k ) {
          const int r = lane + 32 * k;
          const double w = a[k] * a[k] + b[k] * b[k];
          if ( k < nch && r < n && w > best ) {
            best = w;
            bestr = r;
          }
        }
        const double gbest = warp_max( best );
        int cand = ( best == gbest ) ? bestr : n;
#pragma unroll
        for ( int o = 16; o > 0; o >>= 1 ) cand = ::min( cand, __shfl_xor_sync( 0xffffffffu, cand, o ) );
        if ( cand < n ) {
          const int kk = cand >> 5, ll = cand & 31;
          double pa = 0.0, pb2 = 0.0;
#pragma unroll
          for ( int k = 0; k < MAXCH; ++k )
            if ( k == kk ) {
              pa = a[k];
              pb2 = b[k];
            }
          pa = __shfl_sync( 0xffffffffu, pa, ll );
          pb2 = __shfl_sync( 0xffffffffu, pb2, ll );
          double cs, sn, rr;
          dlartg( pa, pb2, cs, sn, rr );
#pragma unroll
          for ( int k = 0; k < MAXCH; ++k ) {
            const double x = a[k], y = b[k];
            a[k] = cs * x + sn * y;
            b[k] = cs * y - sn * x;
            if ( lane + 32 * k == cand ) b[k] = 0.0;
          }
        }
      }
#pragma unroll
      for ( int k = 0; k < MAXCH; ++k ) {
        const int r = lane + 32 * k;
        if ( k < nch && r < n ) {
          V[r + (size_t)c * ldv] = a[k];
          if ( pair ) V[r + (size_t)( c + 1 ) * ldv] = b[k];
        }
      }
    }
    c += pair ? 2 : 1;
    blk += 1;
  }
}

// any n: the same un-balance + normalise with the column(s) streamed from memory in several passes (L2 hits)
__device__ inline void vecs_finish_big( int n, int ilo, int ihi, const double* scale, const double* wi, const double* __restrict__ C,
                                        int ldc, double* __restrict__ V, int ldv, int* perm, bool left, int part = 0, int nparts = 1 ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int stride = kVecWarps * nparts, mine = part * kVecWarps + warp;
  if ( tid == 0 ) {
    for ( int r = 0; r < n; ++r ) perm[r] = r;
    for ( int ii = 0; ii < n; ++ii ) {
      int i = ii;
      if ( i >= ilo && i <= ihi ) continue;
      if ( i < ilo ) i = ilo - 1 - ii;
      const int k = (int)scale[i];
      if ( k == i ) continue;
      const int tmp = perm[i];
      perm[i] = perm[k];
      perm[k] = tmp;
    }
  }
  __syncthreads();
  int blk = 0;
  int c = 0;
  while ( c < n ) {
    const bool pair = ( wi[c] > 0.0 && c + 1 < n );
    if ( ( blk % stride ) == mine ) {
      const double* ca = C + (size_t)c * ldc;
      const double* cb = C + (size_t)( pair ? c + 1 : c ) * ldc;
      auto rowscale = [&]( int src ) {
        const double f = ( src >= ilo && src <= ihi && ilo != ihi ) ? scale[src] : 1.0;
        return left ? 1.0 / f : f;
      };
      double mx = 0.0;
      for ( int r = lane; r < n; r += 32 ) {
        const int src = perm[r];
        const double s = rowscale( src );
        mx = fmax( mx, fabs( ca[src] * s ) );
        if ( pair ) mx = fmax( mx, fabs( cb[src] * s ) );
      }
      mx = warp_max( mx );
      double scl = 0.0;
      if ( mx > 0.0 && mx <= DBL_MAX ) {
        const double inv = 1.0 / mx;
        double ss = 0.0;
        for ( int r = lane; r < n; r += 32 ) {
          const int src = perm[r];
          const double s = rowscale( src ) * inv;
          const double p = ca[src] * s, q = pair ? cb[src] * s : 0.0;
          ss += p * p + q * q;
        }
        ss = warp_sum( ss );
        scl = 1.0 / ( mx * sqrt( ss ) );
      }
      double cs = 1.0, sn = 0.0;
      int cand = n;
      if ( pair ) {
        double best = -1.0;
        int bestr = n;
        for ( int r = lane; r < n; r += 32 ) {
          const int src = perm[r];
          const double s = rowscale( src );
          const double a = ca[src] * s * scl, b = cb[src] * s * scl;
          const double w = a * a + b * b;
          if ( w > best ) {
            best = w;
            bestr = r;
          }
        }
        const double gbest = warp_max( best );
        cand = ( best == gbest ) ? bestr : n;
#pragma unroll
        for ( int o = 16; o > 0; o >>= 1 ) cand = ::min( cand, __shfl_xor_sync( 0xffffffffu, cand, o ) );
        if ( cand < n ) {
          const int src = perm[cand];
          const double s = rowscale( src );
          double rr;
          dlartg( ca[src] * s * scl, cb[src] * s * scl, cs, sn, rr );
        }
      }
      for ( int r = lane; r < n; r += 32 ) {
        const int src = perm[r];
        const double s = rowscale( src );
        const double a = ca[src] * s * scl;
        if ( pair ) {
          const double b = cb[src] * s * scl;
          V[r + (size_t)c * ldv] = cs * a + sn * b;
          V[r + (size_t)( c + 1 ) * ldv] = ( r == cand ) ? 0.0 : cs * b - sn * a;
        } else {
          V[r + (size_t)c * ldv] = a;
        }
      }
    }
    c += pair ? 2 : 1;
    blk += 1;
  }
}

}  // namespace pb

#endif  // __CUDACC__
