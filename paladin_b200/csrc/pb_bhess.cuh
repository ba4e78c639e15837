// pb_bhess.cuh -- BLOCKED Hessenberg reduction and blocked explicit Q for one matrix per CTA (device only, sm_100a).
//
// Replaces the dgehrd / dorghr stages of the reference's dgeev_ call (reference src/lapack-wrapper.hpp:47-60,
// src/paladin.hpp:248-254) for 64 < n <= 1024 with the published LAPACK 3.12 blocked algorithm (dgehrd = dlahr2 panels +
// block-reflector updates, dorghr = backward accumulation of the block reflectors), scheduled for one B200 CTA:
//
//   panel (bh_panel, NB = 32 columns):  per column one streamed product  y = A(k+1:ihi, c+1:ihi) v  -- the only part that
//       reads the trailing matrix, 8 (n-k)(n-c) bytes per column, 8 n^3 / 3 bytes per reduction instead of the 8 n^3 of the
//       unblocked single-sweep kernel (pb_hess.cuh) -- plus thin (n x 32) corrections with the panel's V (kept below the
//       subdiagonal of A), Y = A V T (global scratch, L2-resident) and T (shared memory). Warps own columns, lanes own rows:
//       256-byte coalesced segments, row accumulators in registers, only as many 32-row chunks as the trailing block has.
//   Y top rows (bh_ytop):               Y(0:k, :) = A(0:k, k+1:ihi) (V T)   on the FP64 tensor pipe (mma.sync m8n8k4 f64,
//       fragments straight from global memory: the strip is read once).
//   trailing update (bh_tile_pass):     A <- A - Y V^T (all rows) and A <- (I - V T^T V^T) A (rows k+1..ihi) fused per tile of
//       8 columns: the tile (full column height) is staged in shared memory by TMA (cp.async.bulk.tensor.2d, mbarrier
//       completion, two stages: tile i+1 lands while tile i is updated), both rank-32 updates run as DMMA on the staged tile
//       (C fragments in shared memory, A fragments Y / V from L2), the tile is written back once.
//   explicit Q (bh_form_q):             Q = H(ilo) ... H(ihi-2) by backward accumulation of the same block reflectors
//       (T of every panel is kept in the scratch): the same tile pass with Q as the target.
// Reflectors are LAPACK's (v(c+1) = 1, tau[c], beta on the subdiagonal), so H equals dgehrd's up to rounding.
#pragma once

#include "pb_common.cuh"
#include "pb_hess.cuh"
#include "pb_geev.cuh"

#if defined( __CUDACC__ )

#include <cooperative_groups.h>
#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint, nothing links libcuda)

namespace pb {

constexpr int kTickHessUpdate = 7;  // phase-profile slot: Y top rows + trailing updates (the panels are counted as kTickHess)
constexpr int kBhNB = 32;        // panel width
constexpr int kBhTC = 16;        // columns per update tile (two DMMA n-tiles)
constexpr int kBhMaxStages = 8;  // row blocks of a tile that can be resident at once
#ifndef PB_BH_BOXROWS
#define PB_BH_BOXROWS 256
#endif
#ifndef PB_BH_MAXRES
#define PB_BH_MAXRES 4
#endif
constexpr int kBhBoxRows = PB_BH_BOXROWS;  // rows per TMA box (box = kBhBoxRows x 1 column of doubles)
constexpr int kBhMaxResident = PB_BH_MAXRES;  // a tile whose row blocks number at most this many stays on chip between its two sweeps
constexpr int kBhThreads = kHessThreads;
constexpr int kBhWarps = kBhThreads / 32;
constexpr int kBhLdT = kBhNB + 1;  // T(q, p) at Ts[q * kBhLdT + p]

// an update tile is staged in ROW BLOCKS of kBhBoxRows rows x kBhTC columns (one TMA box per column): doubles per stage
constexpr int kBhStageDoubles = kBhBoxRows * kBhTC;
__host__ __device__ inline int bh_row_blocks( int n ) { return ( n + kBhBoxRows - 1 ) / kBhBoxRows; }
// number of panels of a reduction with n_refl reflector columns
__host__ __device__ inline int bh_num_panels( int n ) { return n > 2 ? ( n - 2 + kBhNB - 1 ) / kBhNB : 0; }
// global scratch per matrix (doubles): Y (ldy x NB), VT (ldy x NB), T of every panel (NB x NB each); ldy = n rounded up to even
__host__ __device__ inline size_t bh_scratch_doubles( int n ) {
  const size_t ldy = (size_t)( ( n + 1 ) & ~1 );
  return 2 * ldy * kBhNB + (size_t)bh_num_panels( n ) * kBhNB * kBhNB + 16;
}
// stages of the tile ring: every row block of the tallest tile resident when that fits (then nothing is re-read), two otherwise
__host__ __device__ inline int bh_stages( int n ) {
  const int nrb = bh_row_blocks( n );
  return nrb <= kBhMaxResident ? ( nrb < 2 ? 2 : nrb ) : 2;
}
// shared memory (doubles): vs, bs (2 n) | ws, wt, w2s, vrow (4 x 32) | red (64) | Ts (32 x 33) |
//   union { ypart: 8 x min(n, 1024) ; stages x (256 x 16) + Ws: 32 x 16 + Wt: 32 x 16 }
__host__ __device__ inline size_t bh_smem_doubles( int n ) {
  const size_t u1 = (size_t)kBhWarps * ( n < 1024 ? n : 1024 );
  const size_t u2 = (size_t)bh_stages( n ) * kBhStageDoubles + 2 * (size_t)kBhNB * kBhTC;
  return 2 * (size_t)n + 4 * kBhNB + 64 + (size_t)kBhNB * kBhLdT + ( u1 > u2 ? u1 : u2 ) + 32;
}

// ---- PTX helpers: mbarrier, TMA tensor load, proxy fence, FP64 MMA ---------------------------------------------------
__device__ __forceinline__ uint32_t bh_smem_addr( const void* p ) { return (uint32_t)__cvta_generic_to_shared( p ); }
__device__ __forceinline__ void bh_mbar_init( uint64_t* bar, int count ) {
  asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( bh_smem_addr( bar ) ), "r"( count ) : "memory" );
}
__device__ __forceinline__ void bh_mbar_expect_tx( uint64_t* bar, uint32_t bytes ) {
  asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( bh_smem_addr( bar ) ), "r"( bytes ) : "memory" );
}
__device__ __forceinline__ bool bh_mbar_try_wait( uint64_t* bar, uint32_t parity ) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"( ok )
      : "r"( bh_smem_addr( bar ) ), "r"( parity )
      : "memory" );
  return ok != 0;
}
// bounded wait: a TMA that never completes (bad descriptor) must end in a trap, not in a hung GPU
__device__ __forceinline__ void bh_mbar_wait( uint64_t* bar, uint32_t parity ) {
  const long long t0 = clock64();
  while ( !bh_mbar_try_wait( bar, parity ) ) {
    if ( clock64() - t0 > 4000000000LL ) __trap();
  }
}
__device__ __forceinline__ void bh_fence_proxy_async() { asm volatile( "fence.proxy.async;" ::: "memory" ); }
// one box (kBhBoxRows rows x 1 column) of the matrix behind `tmap` -> shared memory, completion counted on `bar`
__device__ __forceinline__ void bh_tma_load_box( void* dst, const CUtensorMap* tmap, int row0, int col, uint64_t* bar ) {
  asm volatile( "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                    bh_smem_addr( dst ) ),
                "l"( reinterpret_cast<uint64_t>( tmap ) ), "r"( row0 ), "r"( col ), "r"( bh_smem_addr( bar ) )
                : "memory" );
}
// bulk prefetch of a contiguous global range into L2 (address 16-byte aligned, size a multiple of 16)
__device__ __forceinline__ void bh_prefetch_l2( const void* p, uint32_t bytes ) {
  asm volatile( "cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"( p ), "r"( bytes ) : "memory" );
}
__device__ __forceinline__ void bh_dmma( double& c0, double& c1, double a, double b ) {
  asm volatile( "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"( c0 ), "+d"( c1 ) : "d"( a ), "d"( b ) );
}

// V(r, p) of the panel that starts at column k: unit lower trapezoidal, stored below the subdiagonal of Av
__device__ __forceinline__ double bh_v( const double* Av, int ldv, int k, int ihi, int r, int p ) {
  const int c = k + p;
  if ( r > c + 1 && r <= ihi ) return __ldcg( Av + r + (size_t)c * ldv );
  return r == c + 1 ? 1.0 : 0.0;
}

// the CTAs that share one matrix (a thread-block cluster for n > 1024; {0, 1} = one CTA per matrix)
struct BhPart {
  int rank = 0, size = 1;
};
__device__ __forceinline__ void bh_part_sync( const BhPart& pt ) {
  if ( pt.size > 1 ) cooperative_groups::this_cluster().sync();
  else __syncthreads();
}

// per-CTA working set
struct BhSmem {
  double *vs, *bs, *ws, *wt, *w2s, *vrow, *red, *Ts, *un;  // un: union of ypart / (tiles, Wp, Wt)
  uint64_t* bars;                                          // kBhMaxStages mbarriers (static shared memory of the kernel)
  uint32_t phase;                                          // parity bits of the barriers
  int nstages;                                             // stages of the tile ring
};
__device__ __forceinline__ BhSmem bh_carve( double* sm, int n, uint64_t* bars ) {
  BhSmem s;
  double* p = sm;
  s.vs = p; p += n;
  s.bs = p; p += n;
  s.ws = p; p += kBhNB;
  s.wt = p; p += kBhNB;
  s.w2s = p; p += kBhNB;
  s.vrow = p; p += kBhNB;
  s.red = p; p += 64;
  s.Ts = p; p += kBhNB * kBhLdT;
  // TMA destinations must be 128-byte aligned
  p = reinterpret_cast<double*>( ( reinterpret_cast<uintptr_t>( p ) + 127 ) & ~(uintptr_t)127 );
  s.un = p;
  s.bars = bars;
  s.phase = 0;
  s.nstages = bh_stages( n );
  return s;
}

// ---- streamed product of one panel column: acc over rows rb + lane + 32 ch of  sum_cc A(r, cc) vs[cc] -----------------
// A warp takes NC columns per iteration (CH * NC loads in flight per lane whatever is left of the trailing block).
// The loads of a lane wait for HBM (the batch does not fit L2 and every column step streams the whole trailing block), so lane
// 0 of every warp runs a bulk L2 prefetch kBhPfDist iterations ahead of its own loads, and at the end of a column step
// requests the first iterations of the NEXT step (c_first + 1 ...): they arrive while the thin panel phases run.
constexpr int kBhPfDist = 4;
template <int CH, int NC>
__device__ __forceinline__ void bh_gemv_cols( const double* A, int lda, int rb, int rlast, int c_first, int c_last, const double* vs,
                                              double* ypart_w, int r_first, BhPart pt = BhPart() ) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[CH];
#pragma unroll
  for ( int u = 0; u < CH; ++u ) acc[u] = 0.0;
  const int span = rlast - rb;  // clamp for the (masked) rows past the block
  const bool pf = ( lda & 1 ) == 0 && lane == 0;
  const uint32_t pf_bytes = (uint32_t)( ( ( span + 2 ) & ~1 ) * sizeof( double ) );
  const int stride = pt.size * kBhWarps * NC;
  for ( int cc0 = c_first + ( pt.rank * kBhWarps + warp ) * NC; cc0 <= c_last; cc0 += stride ) {
    if ( pf ) {
      const int cp = cc0 + kBhPfDist * stride;
#pragma unroll
      for ( int q = 0; q < NC; ++q )
        if ( cp + q <= c_last ) bh_prefetch_l2( A + (size_t)( cp + q ) * lda + rb, pf_bytes );
    }
    double a[NC][CH], vc[NC];
#pragma unroll
    for ( int q = 0; q < NC; ++q ) {
      const int cc = imin( cc0 + q, c_last );
      vc[q] = ( cc0 + q <= c_last ) ? vs[cc] : 0.0;
      const double* col = A + (size_t)cc * lda + rb;
#pragma unroll
      for ( int u = 0; u < CH; ++u ) a[q][u] = __ldcg( col + imin( lane + 32 * u, span ) );
    }
#pragma unroll
    for ( int q = 0; q < NC; ++q )
#pragma unroll
      for ( int u = 0; u < CH; ++u ) acc[u] += a[q][u] * vc[q];
  }
  if ( pf ) {
    // the next column step starts one column further right
    for ( int it = 0; it < kBhPfDist; ++it ) {
      const int cp = c_first + 1 + ( ( it * pt.size + pt.rank ) * kBhWarps + warp ) * NC;
#pragma unroll
      for ( int q = 0; q < NC; ++q )
        if ( cp + q <= c_last ) bh_prefetch_l2( A + (size_t)( cp + q ) * lda + rb, pf_bytes );
    }
  }
#pragma unroll
  for ( int u = 0; u < CH; ++u ) {
    const int r = rb + lane + 32 * u;
    if ( r >= r_first && r <= rlast ) ypart_w[r] = acc[u];
  }
}

// dot product of a stored column segment with a shared-memory vector, rows r_first..rlast, all loads of the lane in flight
// together (the thin V^T x products of the panel are latency-bound otherwise: shared memory leaves almost no L1)
template <int CH>
__device__ __forceinline__ double bh_dot_chunks( const double* col, const double* x, int r_first, int rlast ) {
  const int lane = threadIdx.x & 31;
  const int rb = r_first & ~31;
  const int span = rlast - rb;
  double a[CH];
#pragma unroll
  for ( int u = 0; u < CH; ++u ) a[u] = __ldcg( col + rb + imin( lane + 32 * u, span ) );
  double s = 0.0;
#pragma unroll
  for ( int u = 0; u < CH; ++u ) {
    const int r = rb + lane + 32 * u;
    s += ( r >= r_first && r <= rlast ) ? a[u] * x[r] : 0.0;
  }
  return warp_sum( s );
}
template <int MAXCH>
__device__ __forceinline__ double bh_dot( const double* col, const double* x, int r_first, int rlast ) {
  if ( r_first > rlast ) return 0.0;
  const int nch = ( rlast - ( r_first & ~31 ) ) / 32 + 1;
  if ( MAXCH > 16 && nch > 16 ) return bh_dot_chunks<( MAXCH > 16 ? 32 : 4 )>( col, x, r_first, rlast );
  if ( nch > 8 ) return bh_dot_chunks<16>( col, x, r_first, rlast );
  if ( nch > 4 ) return bh_dot_chunks<8>( col, x, r_first, rlast );
  return bh_dot_chunks<4>( col, x, r_first, rlast );
}

// ---- one panel: columns k .. k+ib-1 (dlahr2). On exit V below the subdiagonal of A, tau, T in sh.Ts, Y(k+1:ihi, 0:ib) in
// global memory; the columns right of the panel and the rows above it are untouched. -------------------------------------
template <int MAXCH>
__device__ void bh_panel( int n, int ihi, int k, int ib, double* A, int lda, double* tau, double* Y, int ldy, BhSmem& sh ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* ypart = sh.un;
  for ( int q = tid; q < kBhNB * kBhLdT; q += kBhThreads ) sh.Ts[q] = 0.0;
  for ( int j = 0; j < ib; ++j ) {
    const int c = k + j;
    // (a) b = A(k+1:ihi, c) - Y(k+1:ihi, 0:j) V(c, 0:j)^T
    if ( tid < kBhNB ) sh.vrow[tid] = ( tid < j ) ? bh_v( A, lda, k, ihi, c, tid ) : 0.0;
    __syncthreads();
    for ( int r = k + 1 + tid; r <= ihi; r += kBhThreads ) {
      double x = A[r + (size_t)c * lda];
      const double* yr = Y + r;
#pragma unroll 8
      for ( int jj = 0; jj < j; ++jj ) x -= yr[(size_t)jj * ldy] * sh.vrow[jj];
      sh.bs[r] = x;
    }
    __syncthreads();
    if ( j > 0 ) {
      // (b) b <- (I - V T^T V^T) b : w = V^T b (a warp per column of V), w <- T^T w, b -= V w
      for ( int jj = warp; jj < j; jj += kBhWarps ) {
        const int cj = k + jj;
        const double s = bh_dot<MAXCH>( A + (size_t)cj * lda, sh.bs, cj + 2, ihi ) + sh.bs[cj + 1];  // v(cj+1) = 1
        if ( lane == 0 ) sh.ws[jj] = s;
      }
      __syncthreads();
      if ( tid < j ) {
        double s = 0.0;
        for ( int q = 0; q <= tid; ++q ) s += sh.Ts[q * kBhLdT + tid] * sh.ws[q];
        sh.wt[tid] = s;
      }
      __syncthreads();
      for ( int r = k + 1 + tid; r <= ihi; r += kBhThreads ) {
        // V(r, jj) = A(r, k+jj) for jj < r-k-1, 1 for jj == r-k-1, 0 beyond
        const int d = r - k - 1;
        const int jl = imin( j, d );
        double x = sh.bs[r];
        const double* vr = A + r + (size_t)k * lda;
#pragma unroll 8
        for ( int jj = 0; jj < jl; ++jj ) x -= vr[(size_t)jj * lda] * sh.wt[jj];
        if ( d < j ) x -= sh.wt[d];
        sh.bs[r] = x;
      }
      __syncthreads();
    }
    // (c) reflector from b(c+1:ihi): alpha = b(c+1), x = b(c+2:ihi)   (LAPACK dlarfg)
    double tj = 0.0, beta, sc = 0.0;
    {
      const double alpha = sh.bs[c + 1];
      double mx = 0.0;
      for ( int r = c + 2 + tid; r <= ihi; r += kBhThreads ) mx = fmax( mx, fabs( sh.bs[r] ) );
      mx = cta_max( mx, sh.red );
      beta = alpha;
      if ( mx != 0.0 ) {
        double ss = 0.0;
        for ( int r = c + 2 + tid; r <= ihi; r += kBhThreads ) {
          const double x = sh.bs[r] / mx;
          ss += x * x;
        }
        double xnorm = mx * sqrt( cta_sum( ss, sh.red ) );
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
    }
    __syncthreads();
    // column c of A: H part (rows k+1..c), beta, v ; v also to shared memory (zero outside c+1..ihi)
    for ( int r = tid; r < n; r += kBhThreads ) {
      double v = 0.0;
      if ( r >= k + 1 && r <= c ) A[r + (size_t)c * lda] = sh.bs[r];
      else if ( r == c + 1 ) {
        v = 1.0;
        A[r + (size_t)c * lda] = beta;
      } else if ( r >= c + 2 && r <= ihi ) {
        v = sh.bs[r] * sc;
        A[r + (size_t)c * lda] = v;
      }
      sh.vs[r] = v;
    }
    if ( tid == 0 ) tau[c] = tj;
    __syncthreads();
    // (d) y = A(k+1:ihi, c+1:ihi) v  (streamed; only the 32-row chunks the block still has) and w2 = V(c+1:ihi, 0:j)^T v
    {
      const int rb = ( k + 1 ) & ~31;
      const int nch = ( ihi - rb ) / 32 + 1;
      double* yw = ypart + (size_t)warp * n;
      if ( MAXCH > 24 && nch > 24 ) bh_gemv_cols<( MAXCH > 24 ? 32 : 4 ), 1>( A, lda, rb, ihi, c + 1, ihi, sh.vs, yw, k + 1 );
      else if ( MAXCH > 16 && nch > 16 ) bh_gemv_cols<( MAXCH > 16 ? 24 : 4 ), 1>( A, lda, rb, ihi, c + 1, ihi, sh.vs, yw, k + 1 );
      else if ( nch > 12 ) bh_gemv_cols<16, 1>( A, lda, rb, ihi, c + 1, ihi, sh.vs, yw, k + 1 );
      else if ( nch > 8 ) bh_gemv_cols<12, 1>( A, lda, rb, ihi, c + 1, ihi, sh.vs, yw, k + 1 );
      else if ( nch > 4 ) bh_gemv_cols<8, 2>( A, lda, rb, ihi, c + 1, ihi, sh.vs, yw, k + 1 );
      else if ( nch > 2 ) bh_gemv_cols<4, 4>( A, lda, rb, ihi, c + 1, ihi, sh.vs, yw, k + 1 );
      else bh_gemv_cols<2, 8>( A, lda, rb, ihi, c + 1, ihi, sh.vs, yw, k + 1 );
      for ( int jj = warp; jj < j; jj += kBhWarps ) {
        const double s = bh_dot<MAXCH>( A + (size_t)( k + jj ) * lda, sh.vs, c + 1, ihi );
        if ( lane == 0 ) sh.w2s[jj] = s;
      }
    }
    __syncthreads();
    // (e) Y(:, j) = tau (y - Y(:, 0:j) w2) ; T(0:j, j) = -tau T(0:j, 0:j) w2, T(j, j) = tau
    for ( int r = k + 1 + tid; r <= ihi; r += kBhThreads ) {
      double y = 0.0;
#pragma unroll
      for ( int w = 0; w < kBhWarps; ++w ) y += ypart[(size_t)w * n + r];  // (every warp stored its rows, zeros included)
      const double* yr = Y + r;
#pragma unroll 8
      for ( int jj = 0; jj < j; ++jj ) y -= yr[(size_t)jj * ldy] * sh.w2s[jj];
      Y[r + (size_t)j * ldy] = tj * y;
    }
    if ( tid < j ) {
      double s = 0.0;
      for ( int p = tid; p < j; ++p ) s += sh.Ts[tid * kBhLdT + p] * sh.w2s[p];
      sh.ws[tid] = -tj * s;  // (read of column < j, write of column j: no hazard, but keep the value out of Ts until all
    }                        //  threads of this phase are done with w2s)
    __syncthreads();
    if ( tid < j ) sh.Ts[tid * kBhLdT + j] = sh.ws[tid];
    if ( tid == 0 ) sh.Ts[j * kBhLdT + j] = tj;
    __syncthreads();
  }
}

// ---- VT = V T (rows k+1..ihi) to global memory; T of the panel to its slot of the scratch ----------------------------------
__device__ inline void bh_vt( int ihi, int k, int ib, const double* A, int lda, double* VT, int ldy, double* Tsave, const BhSmem& sh,
                              BhPart pt = BhPart() ) {
  const int tid = threadIdx.x;
  if ( pt.rank == 0 )
    for ( int q = tid; q < kBhNB * kBhNB; q += kBhThreads ) Tsave[q] = sh.Ts[( q >> 5 ) * kBhLdT + ( q & 31 )];
  for ( int r = k + 1 + pt.rank * kBhThreads + tid; r <= ihi; r += kBhThreads * pt.size ) {
    double v[kBhNB];
#pragma unroll
    for ( int q = 0; q < kBhNB; ++q ) v[q] = ( q < ib ) ? bh_v( A, lda, k, ihi, r, q ) : 0.0;
    for ( int p = 0; p < ib; ++p ) {
      double s = 0.0;
#pragma unroll
      for ( int q = 0; q < kBhNB; ++q ) s += ( q <= p ) ? v[q] * sh.Ts[q * kBhLdT + p] : 0.0;
      VT[r + (size_t)p * ldy] = s;
    }
  }
  bh_part_sync( pt );
}

// ---- Y(0:k, 0:ib) = A(0:k, k+1:ihi) VT(k+1:ihi, 0:ib)   (DMMA, fragments from global memory) -------------------------------
__device__ inline void bh_ytop( int ihi, int k, int ib, const double* A, int lda, const double* VT, double* Y, int ldy, BhPart pt = BhPart() ) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int ntile = ( k + 1 + 7 ) >> 3;
  for ( int rt = pt.rank * kBhWarps + warp; rt < ntile; rt += kBhWarps * pt.size ) {
    const int r = 8 * rt + gid;
    const bool rok = r <= k;
    const double* arow = A + imin( r, k );
    double acc[4][2];
#pragma unroll
    for ( int nt = 0; nt < 4; ++nt ) acc[nt][0] = acc[nt][1] = 0.0;
    for ( int c4 = k + 1; c4 <= ihi; c4 += 16 ) {
      double a[4], b[4][4];
#pragma unroll
      for ( int u = 0; u < 4; ++u ) {
        const int cc = c4 + 4 * u + tig;
        const bool cok = cc <= ihi;
        const int ccl = imin( cc, ihi );
        a[u] = ( rok && cok ) ? __ldcg( arow + (size_t)ccl * lda ) : 0.0;
#pragma unroll
        for ( int nt = 0; nt < 4; ++nt ) {
          const int p = 8 * nt + gid;
          b[u][nt] = ( cok && p < ib ) ? __ldcg( VT + ccl + (size_t)p * ldy ) : 0.0;
        }
      }
#pragma unroll
      for ( int u = 0; u < 4; ++u )
#pragma unroll
        for ( int nt = 0; nt < 4; ++nt ) bh_dmma( acc[nt][0], acc[nt][1], a[u], b[u][nt] );
    }
    if ( rok ) {
#pragma unroll
      for ( int nt = 0; nt < 4; ++nt ) {
        const int p = 8 * nt + 2 * tig;
        if ( p < ib ) Y[r + (size_t)p * ldy] = acc[nt][0];
        if ( p + 1 < ib ) Y[r + (size_t)( p + 1 ) * ldy] = acc[nt][1];
      }
    }
  }
  bh_part_sync( pt );
}

// ---- tile pass: columns [c_begin, c_end) of the target M in tiles of kBhTC columns, staged by TMA in row blocks ------------------
//   R update (r_rows > 0):  M(0:r_rows, cc) -= Y(0:r_rows, 0:ib) V(cc, 0:ib)^T          for cc <= r_col_hi
//   L update (do_l):        M(k+1:ihi, cc)  <- (I - V op(T) V^T) M(k+1:ihi, cc)          op(T) = T^T (TRANS_T) or T
//   rows [st_lo, st_hi) of every tile column are written back.
// A tile is worked on in row blocks of kBhBoxRows rows (one TMA box per column and block, sh.nstages blocks on chip):
//   sweep A  per block: R update (DMMA, C fragments in shared memory, Y fragments from L2), then the block's share of
//            W = V^T M (every warp owns one 8 x 8 tile of the 32 x 16 result and walks all rows: no reduction);
//   then     Wt = -op(T) W;
//   sweep B  per block: M -= V Wt (DMMA), write back.
// When the tile's blocks all fit the ring (n <= 4 kBhBoxRows) they stay resident between the sweeps; otherwise sweep A
// writes its blocks back and sweep B streams them again (they come from L2).
// V lives below the subdiagonal of Av (panel start k, ib columns). tmap describes M (box kBhBoxRows x 1); without a tensor
// map (odd leading dimension) the blocks are staged by plain loads.
template <bool TRANS_T>
__device__ void bh_tile_pass( int n, int ihi, int k, int ib, double* M, int ldm, const CUtensorMap* tmap, const double* Av, int ldv,
                              const double* Y, int ldy, int r_rows, int r_col_hi, bool do_l, int c_begin, int c_end, int st_lo, int st_hi,
                              BhSmem& sh, BhPart pt = BhPart() ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int NS = sh.nstages;
  double* ring = sh.un;                                      // [NS][kBhTC][kBhBoxRows]
  double* Ws = ring + (size_t)NS * kBhStageDoubles;           // [kBhNB][kBhTC]
  double* Wt = Ws + kBhNB * kBhTC;                            // [kBhNB][kBhTC]
  const int ntiles = ( c_end - c_begin + kBhTC - 1 ) / kBhTC;
  if ( ntiles <= 0 ) return;
  // row blocks the pass touches: R rows [0, r_rows), L rows [k+1, ihi], stored rows [st_lo, st_hi)
  const int row_lo = imin( st_lo, imin( r_rows > 0 ? 0 : n, do_l ? k + 1 : n ) );
  const int row_hi = imax( st_hi - 1, imax( r_rows - 1, do_l ? ihi : -1 ) );
  const int rb0 = row_lo / kBhBoxRows, rb1 = row_hi / kBhBoxRows;
  const int nA = rb1 - rb0 + 1;
  const bool resident = nA <= NS;
  const int lb0 = do_l ? ( k + 1 ) / kBhBoxRows : 0, lb1 = do_l ? ihi / kBhBoxRows : -1;  // blocks of sweep B
  const int nB = ( do_l && !resident ) ? lb1 - lb0 + 1 : 0;
  const int nseq = nA + nB;  // loads per tile, in this order: sweep A blocks, then (streaming mode) sweep B blocks
  // global data this CTA wrote through the generic proxy is about to be read by the async proxy (TMA)
  bh_fence_proxy_async();
  __syncthreads();
  auto block_of = [&]( int q ) { return q < nA ? rb0 + q : lb0 + ( q - nA ); };
  auto issue = [&]( int cc0, int q ) {  // thread 0 only: load q of the tile that starts at column cc0
    const int stage = q % NS;
    double* dst = ring + (size_t)stage * kBhStageDoubles;
    const int row0 = block_of( q ) * kBhBoxRows;
    bh_mbar_expect_tx( &sh.bars[stage], (uint32_t)( kBhStageDoubles * sizeof( double ) ) );
    for ( int c = 0; c < kBhTC; ++c ) bh_tma_load_box( dst + (size_t)c * kBhBoxRows, tmap, row0, cc0 + c, &sh.bars[stage] );
  };
  // block q of the current tile is on chip after this
  auto acquire = [&]( int cc0, int q ) -> double* {
    const int stage = q % NS;
    double* St = ring + (size_t)stage * kBhStageDoubles;
    if ( tmap != nullptr ) {
      bh_mbar_wait( &sh.bars[stage], ( sh.phase >> stage ) & 1u );
      sh.phase ^= 1u << stage;
    } else {
      const int row0 = block_of( q ) * kBhBoxRows;
      __syncthreads();
      for ( int e = tid; e < kBhStageDoubles; e += kBhThreads ) {
        const int c = e / kBhBoxRows, r = row0 + ( e - c * kBhBoxRows );
        St[e] = ( cc0 + c < n && r < n ) ? M[r + (size_t)( cc0 + c ) * ldm] : 0.0;
      }
      __syncthreads();
    }
    return St;
  };
  // rows [lo, hi) of a block back to global memory, then the stage is free for load q + NS
  auto store_rows = [&]( const double* St, int row0, int cc0, int lo, int hi ) {
    lo = imax( lo, row0 );
    hi = imin( hi, row0 + kBhBoxRows );
    const int nr = hi - lo;
    if ( nr <= 0 ) return;
    for ( int e = tid; e < kBhTC * nr; e += kBhThreads ) {
      const int c = e / nr, r = lo + ( e - c * nr );
      if ( cc0 + c < c_end ) M[r + (size_t)( cc0 + c ) * ldm] = St[(size_t)c * kBhBoxRows + ( r - row0 )];
    }
  };
  // (sweep B re-reads what sweep A stored: its loads are never requested before sweep A has released its last block)
  auto release = [&]( int cc0, int q ) {
    bh_fence_proxy_async();
    __syncthreads();
    const int lim = q < nA ? nA : nseq;
    if ( tmap != nullptr && tid == 0 && q + NS < lim ) issue( cc0, q + NS );
  };
  // M(rows of the block) -= A B with A = rows x 32 fragments from L2 (Y or V), B = 32 x 16 in registers (two n-tiles)
  auto rank_update = [&]( double* St, int row0, int r_first, int r_last, const double ( &bfr )[2][8], bool from_y ) {
    const int t0 = imax( r_first, row0 ) >> 3, t1 = imin( r_last, row0 + kBhBoxRows - 1 ) >> 3;  // 8-row tiles (global numbering)
    auto load_a = [&]( int rt, double* a ) {
      const int r = 8 * rt + gid;
#pragma unroll
      for ( int s2 = 0; s2 < 8; ++s2 ) {
        const int p = 4 * s2 + tig;
        if ( from_y ) a[s2] = ( r <= r_last && p < ib ) ? __ldcg( Y + imin( r, r_last ) + (size_t)p * ldy ) : 0.0;
        else a[s2] = ( p < ib ) ? bh_v( Av, ldv, k, ihi, r, p ) : 0.0;
      }
    };
    for ( int rt = t0 + warp; rt <= t1; rt += kBhWarps ) {
      double a[8];
      load_a( rt, a );
      double* cp = St + (size_t)( 2 * tig ) * kBhBoxRows + ( 8 * rt + gid - row0 );
      double c00 = cp[0], c01 = cp[kBhBoxRows], c10 = cp[8 * kBhBoxRows], c11 = cp[9 * kBhBoxRows];
#pragma unroll
      for ( int s2 = 0; s2 < 8; ++s2 ) {
        bh_dmma( c00, c01, a[s2], bfr[0][s2] );
        bh_dmma( c10, c11, a[s2], bfr[1][s2] );
      }
      cp[0] = c00;
      cp[kBhBoxRows] = c01;
      cp[8 * kBhBoxRows] = c10;
      cp[9 * kBhBoxRows] = c11;
    }
  };
  if ( tmap != nullptr && tid == 0 && pt.rank < ntiles )
    for ( int q = 0; q < imin( NS, nA ); ++q ) issue( c_begin + pt.rank * kBhTC, q );
  for ( int t = pt.rank; t < ntiles; t += pt.size ) {
    const int cc0 = c_begin + t * kBhTC;
    const bool do_r = r_rows > 0 && cc0 <= r_col_hi;
    double bf[2][8];  // B fragments: -V(cc, :) during sweep A, Wt during sweep B
    if ( do_r ) {
#pragma unroll
      for ( int nt = 0; nt < 2; ++nt ) {
        const int cc = cc0 + 8 * nt + gid;
        const bool cok = cc < c_end && cc <= r_col_hi;
#pragma unroll
        for ( int s2 = 0; s2 < 8; ++s2 ) {
          const int p = 4 * s2 + tig;
          bf[nt][s2] = ( cok && p < ib ) ? -bh_v( Av, ldv, k, ihi, cc, p ) : 0.0;
        }
      }
    }
    // W(p, c): warp w owns the 8 x 8 tile (mt = w & 3, nt = w >> 2), four interleaved accumulator pairs over the rows
    const int wmt = warp & 3, wnt = warp >> 2;
    double wacc[2][2];
#pragma unroll
    for ( int h = 0; h < 2; ++h ) wacc[h][0] = wacc[h][1] = 0.0;
    // ---- sweep A ----
    for ( int q = 0; q < nA; ++q ) {
      const int row0 = ( rb0 + q ) * kBhBoxRows;
      double* St = acquire( cc0, q );
      if ( do_r && row0 < r_rows ) {
        rank_update( St, row0, 0, r_rows - 1, bf, true );
        __syncthreads();
      }
      if ( do_l ) {
        const int g0 = imax( k + 1, row0 ) >> 2, g1 = imin( ihi, row0 + kBhBoxRows - 1 ) >> 2;  // groups of 4 rows
        const int p = 8 * wmt + gid;
        for ( int g = g0; g <= g1; g += 2 ) {
          double av[2], bb[2];
#pragma unroll
          for ( int h = 0; h < 2; ++h ) {
            const int gg = g + h;
            const int r = 4 * imin( gg, g1 ) + tig;
            const bool ok = gg <= g1;
            bb[h] = ok ? St[(size_t)( 8 * wnt + gid ) * kBhBoxRows + ( r - row0 )] : 0.0;
            av[h] = ( ok && p < ib ) ? bh_v( Av, ldv, k, ihi, r, p ) : 0.0;
          }
#pragma unroll
          for ( int h = 0; h < 2; ++h ) bh_dmma( wacc[h][0], wacc[h][1], av[h], bb[h] );
        }
      }
      if ( !resident ) {
        if ( do_r ) store_rows( St, row0, cc0, imax( st_lo, 0 ), imin( st_hi, r_rows ) );
        release( cc0, q );
      }
    }
    if ( do_l ) {
      {
        double* wp = Ws + ( 8 * wmt + gid ) * kBhTC + 8 * wnt + 2 * tig;
        wp[0] = wacc[0][0] + wacc[1][0];
        wp[1] = wacc[0][1] + wacc[1][1];
      }
      __syncthreads();
      for ( int e = tid; e < kBhNB * kBhTC; e += kBhThreads ) {
        const int p = e / kBhTC, c = e - p * kBhTC;
        double acc = 0.0;
        if ( TRANS_T ) {
          for ( int q2 = 0; q2 <= p; ++q2 ) acc += sh.Ts[q2 * kBhLdT + p] * Ws[q2 * kBhTC + c];
        } else {
          for ( int q2 = p; q2 < ib; ++q2 ) acc += sh.Ts[p * kBhLdT + q2] * Ws[q2 * kBhTC + c];
        }
        Wt[e] = ( p < ib ) ? -acc : 0.0;
      }
      __syncthreads();
#pragma unroll
      for ( int nt = 0; nt < 2; ++nt )
#pragma unroll
        for ( int s2 = 0; s2 < 8; ++s2 ) bf[nt][s2] = Wt[( 4 * s2 + tig ) * kBhTC + 8 * nt + gid];
      // ---- sweep B ----
      if ( !resident && tmap != nullptr && tid == 0 )
        for ( int q = nA; q < imin( nA + NS, nseq ); ++q ) issue( cc0, q );
      for ( int blk = lb0; blk <= lb1; ++blk ) {
        const int row0 = blk * kBhBoxRows;
        const int q = resident ? blk - rb0 : nA + ( blk - lb0 );
        double* St = resident ? ring + (size_t)( q % NS ) * kBhStageDoubles : acquire( cc0, q );
        rank_update( St, row0, k + 1, ihi, bf, false );
        if ( !resident ) {
          __syncthreads();
          store_rows( St, row0, cc0, imax( st_lo, k + 1 ), imin( st_hi, ihi + 1 ) );
          release( cc0, q );
        }
      }
    }
    if ( resident ) {
      __syncthreads();
      for ( int q = 0; q < nA; ++q ) store_rows( ring + (size_t)( q % NS ) * kBhStageDoubles, ( rb0 + q ) * kBhBoxRows, cc0, st_lo, st_hi );
      bh_fence_proxy_async();
      __syncthreads();
    }
    // the loads of the next tile
    if ( tmap != nullptr && tid == 0 && t + pt.size < ntiles ) {
      if ( resident ) {
        for ( int q = 0; q < nseq; ++q ) issue( cc0 + pt.size * kBhTC, q );
      } else {
        for ( int q = 0; q < imin( NS, nA ); ++q ) issue( cc0 + pt.size * kBhTC, q );
      }
    }
  }
}

// ---- the reduction ------------------------------------------------------------------------------------------------------
// scratch: bh_scratch_doubles(n) doubles of global memory owned by this matrix; tmapA: tensor map of A (or nullptr)
template <int MAXCH>
__device__ void bh_gehrd( const CtaGroup& g, int n, int ilo, int ihi, double* A, int lda, double* tau, double* scratch, const CUtensorMap* tmapA, BhSmem& sh ) {
  const int tid = threadIdx.x;
  const int ldy = ( n + 1 ) & ~1;
  double* Y = scratch;
  double* VT = Y + (size_t)ldy * kBhNB;
  double* Tall = VT + (size_t)ldy * kBhNB;
  for ( int j = tid; j < n; j += kBhThreads ) tau[j] = 0.0;
  __syncthreads();
  int pidx = 0;
  for ( int k = ilo; k <= ihi - 2; k += kBhNB, ++pidx ) {
    const int ib = imin( kBhNB, ihi - 2 - k + 1 );
    bh_panel<MAXCH>( n, ihi, k, ib, A, lda, tau, Y, ldy, sh );
    g.tick( kTickHess );
    bh_vt( ihi, k, ib, A, lda, VT, ldy, Tall + (size_t)pidx * kBhNB * kBhNB, sh );
    bh_ytop( ihi, k, ib, A, lda, VT, Y, ldy );
    // rows above the panel in the panel's own columns (dgehrd's dtrmm + daxpy step)
    bh_tile_pass<true>( n, ihi, k, ib, A, lda, tmapA, A, lda, Y, ldy, k + 1, k + ib - 1, false, k + 1, k + ib, 0, k + 1, sh );
    // trailing columns: right update over all rows (columns <= ihi), left update over rows k+1..ihi (columns up to n-1)
    bh_tile_pass<true>( n, ihi, k, ib, A, lda, tmapA, A, lda, Y, ldy, ihi + 1, ihi, true, k + ib, n, 0, ihi + 1, sh );
    g.tick( kTickHessUpdate );
  }
}

// Q (n x n, ldq) = H(ilo) ... H(ihi-2): identity, then the block reflectors of the reduction applied from the last panel to
// the first to Q(k+1:ihi, k+1:ihi)
__device__ inline void bh_form_q( int n, int ilo, int ihi, const double* A, int lda, double* Q, int ldq, const double* scratch,
                                  const CUtensorMap* tmapQ, BhSmem& sh ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ldy = ( n + 1 ) & ~1;
  const double* Tall = scratch + (size_t)2 * ldy * kBhNB;
  for ( int c = warp; c < n; c += kBhWarps )
    for ( int r = lane; r < n; r += 32 ) Q[r + (size_t)c * ldq] = ( r == c ) ? 1.0 : 0.0;
  __syncthreads();
  const int nrefl = ihi - 2 - ilo + 1;
  if ( nrefl <= 0 ) return;
  const int npan = ( nrefl + kBhNB - 1 ) / kBhNB;
  for ( int pidx = npan - 1; pidx >= 0; --pidx ) {
    const int k = ilo + pidx * kBhNB;
    const int ib = imin( kBhNB, ihi - 2 - k + 1 );
    const double* Tp = Tall + (size_t)pidx * kBhNB * kBhNB;
    for ( int q = tid; q < kBhNB * kBhNB; q += kBhThreads ) sh.Ts[( q >> 5 ) * kBhLdT + ( q & 31 )] = Tp[q];
    __syncthreads();
    bh_tile_pass<false>( n, ihi, k, ib, Q, ldq, tmapQ, A, lda, nullptr, 0, 0, -1, true, k + 1, ihi + 1, k + 1, ihi + 1, sh );
  }
}

// =====================================================================================================================
// n > 1024: the same blocked reduction with ONE MATRIX SHARED BY A THREAD-BLOCK CLUSTER of S CTAs (S = 1 when the batch
// holds enough large matrices to fill the GPU with one CTA each, up to 16 when it holds a few).
//   panel: every CTA runs the thin O(n * 32) steps redundantly (bit-identical: same data, same order), the streamed
//          product y = A(k+1:ihi, c+1:ihi) v is split by COLUMNS over the cluster (rows in segments of 1024 = 32 register
//          chunks), the partial sums meet in global memory (L2), and every CTA adds them up in the same order;
//          rank 0 alone writes column c of A, tau and Y(:, j). Two cluster barriers per column.
//   trailing update / Y top rows / V T / explicit Q: tiles (or rows) dealt round-robin to the CTAs, a cluster barrier
//          between passes; update tiles are streamed in row blocks (two stages on chip).
// Everything another CTA may have written is read through L2 (ld.global.cg / TMA).
// =====================================================================================================================
constexpr int kBhSeg = 1024;  // rows per segment of the streamed product (32 chunks of 32 rows in registers)

// extra global scratch of the cluster path: S partial products of n rows (behind the bh_scratch_doubles block)
__host__ __device__ inline size_t bhc_scratch_doubles( int n, int smax ) {
  return bh_scratch_doubles( n ) + (size_t)smax * ( ( n + 1 ) & ~1 ) + 16;
}

__device__ inline void bhc_panel( int n, int ihi, int k, int ib, double* A, int lda, double* tau, double* Y, int ldy, double* gpart, BhSmem& sh,
                                  const BhPart& pt ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* ypart = sh.un;  // [kBhWarps][kBhSeg]
  for ( int q = tid; q < kBhNB * kBhLdT; q += kBhThreads ) sh.Ts[q] = 0.0;
  for ( int j = 0; j < ib; ++j ) {
    const int c = k + j;
    // (a) b = A(k+1:ihi, c) - Y(k+1:ihi, 0:j) V(c, 0:j)^T
    if ( tid < kBhNB ) sh.vrow[tid] = ( tid < j ) ? bh_v( A, lda, k, ihi, c, tid ) : 0.0;
    __syncthreads();
    for ( int r = k + 1 + tid; r <= ihi; r += kBhThreads ) {
      double x = __ldcg( A + r + (size_t)c * lda );
      const double* yr = Y + r;
#pragma unroll 8
      for ( int jj = 0; jj < j; ++jj ) x -= __ldcg( yr + (size_t)jj * ldy ) * sh.vrow[jj];
      sh.bs[r] = x;
    }
    __syncthreads();
    if ( j > 0 ) {
      // (b) b <- (I - V T^T V^T) b
      for ( int jj = warp; jj < j; jj += kBhWarps ) {
        const int cj = k + jj;
        double s = sh.bs[cj + 1];
        for ( int r0 = cj + 2; r0 <= ihi; r0 += kBhSeg - 32 ) s += bh_dot<32>( A + (size_t)cj * lda, sh.bs, r0, imin( ihi, r0 + kBhSeg - 33 ) );  // (31 rows of alignment slack)
        if ( lane == 0 ) sh.ws[jj] = s;
      }
      __syncthreads();
      if ( tid < j ) {
        double s = 0.0;
        for ( int q = 0; q <= tid; ++q ) s += sh.Ts[q * kBhLdT + tid] * sh.ws[q];
        sh.wt[tid] = s;
      }
      __syncthreads();
      for ( int r = k + 1 + tid; r <= ihi; r += kBhThreads ) {
        const int d = r - k - 1;
        const int jl = imin( j, d );
        double x = sh.bs[r];
        const double* vr = A + r + (size_t)k * lda;
#pragma unroll 8
        for ( int jj = 0; jj < jl; ++jj ) x -= __ldcg( vr + (size_t)jj * lda ) * sh.wt[jj];
        if ( d < j ) x -= sh.wt[d];
        sh.bs[r] = x;
      }
      __syncthreads();
    }
    // (c) reflector from b(c+1:ihi)   (LAPACK dlarfg)
    double tj = 0.0, beta, sc = 0.0;
    {
      const double alpha = sh.bs[c + 1];
      double mx = 0.0;
      for ( int r = c + 2 + tid; r <= ihi; r += kBhThreads ) mx = fmax( mx, fabs( sh.bs[r] ) );
      mx = cta_max( mx, sh.red );
      beta = alpha;
      if ( mx != 0.0 ) {
        double ss = 0.0;
        for ( int r = c + 2 + tid; r <= ihi; r += kBhThreads ) {
          const double x = sh.bs[r] / mx;
          ss += x * x;
        }
        double xnorm = mx * sqrt( cta_sum( ss, sh.red ) );
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
    }
    __syncthreads();
    // v to shared memory only: column c of A is written (by rank 0) after the cluster has finished reading it
    for ( int r = tid; r < n; r += kBhThreads ) {
      double v = 0.0;
      if ( r == c + 1 ) v = 1.0;
      else if ( r >= c + 2 && r <= ihi ) v = sh.bs[r] * sc;
      sh.vs[r] = v;
    }
    __syncthreads();
    // (d) this CTA's share of y = A(k+1:ihi, c+1:ihi) v, a segment of rows at a time, to gpart[rank]; w2 = V^T v
    {
      const int rb = ( k + 1 ) & ~31;
      double* gp = gpart + (size_t)pt.rank * ldy;
      for ( int r0 = rb; r0 <= ihi; r0 += kBhSeg ) {
        const int r1 = imin( ihi, r0 + kBhSeg - 1 );
        const int nch = ( r1 - r0 ) / 32 + 1;
        double* yw = ypart + (size_t)warp * kBhSeg - r0;  // indexed by the global row
        if ( nch > 24 ) bh_gemv_cols<32, 1>( A, lda, r0, r1, c + 1, ihi, sh.vs, yw, k + 1, pt );
        else if ( nch > 16 ) bh_gemv_cols<24, 1>( A, lda, r0, r1, c + 1, ihi, sh.vs, yw, k + 1, pt );
        else if ( nch > 8 ) bh_gemv_cols<16, 1>( A, lda, r0, r1, c + 1, ihi, sh.vs, yw, k + 1, pt );
        else if ( nch > 4 ) bh_gemv_cols<8, 2>( A, lda, r0, r1, c + 1, ihi, sh.vs, yw, k + 1, pt );
        else bh_gemv_cols<4, 4>( A, lda, r0, r1, c + 1, ihi, sh.vs, yw, k + 1, pt );
        __syncthreads();
        for ( int r = imax( r0, k + 1 ) + tid; r <= r1; r += kBhThreads ) {
          double y = 0.0;
#pragma unroll
          for ( int w = 0; w < kBhWarps; ++w ) y += ypart[(size_t)w * kBhSeg + ( r - r0 )];
          __stcg( gp + r, y );
        }
        __syncthreads();
      }
      for ( int jj = warp; jj < j; jj += kBhWarps ) {
        double s = 0.0;
        for ( int r0 = c + 1; r0 <= ihi; r0 += kBhSeg - 32 ) s += bh_dot<32>( A + (size_t)( k + jj ) * lda, sh.vs, r0, imin( ihi, r0 + kBhSeg - 33 ) );
        if ( lane == 0 ) sh.w2s[jj] = s;
      }
    }
    bh_part_sync( pt );  // every share of y is in L2; nobody reads column c of A any more
    // (e) Y(:, j) = tau (y - Y(:, 0:j) w2) (every CTA the same sum, rank 0 stores) ; column c of A, tau ; T(:, j)
    for ( int r = k + 1 + tid; r <= ihi; r += kBhThreads ) {
      double y = 0.0;
      for ( int s2 = 0; s2 < pt.size; ++s2 ) y += __ldcg( gpart + (size_t)s2 * ldy + r );
      const double* yr = Y + r;
#pragma unroll 8
      for ( int jj = 0; jj < j; ++jj ) y -= __ldcg( yr + (size_t)jj * ldy ) * sh.w2s[jj];
      if ( pt.rank == 0 ) __stcg( Y + r + (size_t)j * ldy, tj * y );
    }
    if ( pt.rank == 0 ) {
      for ( int r = k + 1 + tid; r <= ihi; r += kBhThreads ) {
        double x;
        if ( r <= c ) x = sh.bs[r];
        else if ( r == c + 1 ) x = beta;
        else x = sh.vs[r];
        __stcg( A + r + (size_t)c * lda, x );
      }
      if ( tid == 0 ) tau[c] = tj;
    }
    if ( tid < j ) {
      double s = 0.0;
      for ( int p = tid; p < j; ++p ) s += sh.Ts[tid * kBhLdT + p] * sh.w2s[p];
      sh.ws[tid] = -tj * s;
    }
    __syncthreads();
    if ( tid < j ) sh.Ts[tid * kBhLdT + j] = sh.ws[tid];
    if ( tid == 0 ) sh.Ts[j * kBhLdT + j] = tj;
    bh_part_sync( pt );  // column c, Y(:, j) visible to the cluster
  }
}

__device__ inline void bhc_gehrd( const CtaGroup& g, int n, int ilo, int ihi, double* A, int lda, double* tau, double* scratch,
                                  const CUtensorMap* tmapA, BhSmem& sh, const BhPart& pt ) {
  const int tid = threadIdx.x;
  const int ldy = ( n + 1 ) & ~1;
  double* Y = scratch;
  double* VT = Y + (size_t)ldy * kBhNB;
  double* Tall = VT + (size_t)ldy * kBhNB;
  double* gpart = scratch + bh_scratch_doubles( n );
  if ( pt.rank == 0 )
    for ( int j = tid; j < n; j += kBhThreads ) tau[j] = 0.0;
  bh_part_sync( pt );
  int pidx = 0;
  for ( int k = ilo; k <= ihi - 2; k += kBhNB, ++pidx ) {
    const int ib = imin( kBhNB, ihi - 2 - k + 1 );
    bhc_panel( n, ihi, k, ib, A, lda, tau, Y, ldy, gpart, sh, pt );
    g.tick( kTickHess );
    bh_vt( ihi, k, ib, A, lda, VT, ldy, Tall + (size_t)pidx * kBhNB * kBhNB, sh, pt );
    bh_ytop( ihi, k, ib, A, lda, VT, Y, ldy, pt );
    bh_tile_pass<true>( n, ihi, k, ib, A, lda, tmapA, A, lda, Y, ldy, k + 1, k + ib - 1, false, k + 1, k + ib, 0, k + 1, sh, pt );
    bh_tile_pass<true>( n, ihi, k, ib, A, lda, tmapA, A, lda, Y, ldy, ihi + 1, ihi, true, k + ib, n, 0, ihi + 1, sh, pt );
    bh_part_sync( pt );
    g.tick( kTickHessUpdate );
  }
}

__device__ inline void bhc_form_q( int n, int ilo, int ihi, const double* A, int lda, double* Q, int ldq, const double* scratch,
                                   const CUtensorMap* tmapQ, BhSmem& sh, const BhPart& pt ) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ldy = ( n + 1 ) & ~1;
  const double* Tall = scratch + (size_t)2 * ldy * kBhNB;
  for ( int c = pt.rank * kBhWarps + warp; c < n; c += kBhWarps * pt.size )
    for ( int r = lane; r < n; r += 32 ) Q[r + (size_t)c * ldq] = ( r == c ) ? 1.0 : 0.0;
  bh_part_sync( pt );
  const int nrefl = ihi - 2 - ilo + 1;
  if ( nrefl <= 0 ) return;
  const int npan = ( nrefl + kBhNB - 1 ) / kBhNB;
  for ( int pidx = npan - 1; pidx >= 0; --pidx ) {
    const int k = ilo + pidx * kBhNB;
    const int ib = imin( kBhNB, ihi - 2 - k + 1 );
    const double* Tp = Tall + (size_t)pidx * kBhNB * kBhNB;
    for ( int q = tid; q < kBhNB * kBhNB; q += kBhThreads ) sh.Ts[( q >> 5 ) * kBhLdT + ( q & 31 )] = __ldcg( Tp + q );
    __syncthreads();
    bh_tile_pass<false>( n, ihi, k, ib, Q, ldq, tmapQ, A, lda, nullptr, 0, 0, -1, true, k + 1, ihi + 1, k + 1, ihi + 1, sh, pt );
    bh_part_sync( pt );
  }
}

}  // namespace pb

#endif  // __CUDACC__
