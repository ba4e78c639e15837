// pb_small.cuh -- group-parallel dense kernels for matrices that are small enough to be worked on
// in place by one warp or one CTA (shared or global memory, generic pointers, column-major):
//   gebal  : permutation + diagonal scaling                 (LAPACK dgebal 'B', reached from dgeev)
//   gehd2  : unblocked Householder Hessenberg reduction     (LAPACK dgehd2)
//   orghr  : explicit Q from the stored reflectors          (LAPACK dorghr/dorg2r)
//   lahqr  : Francis double-shift QR with Ahues-Tisseur deflation, exceptional shifts and dlanv2
//            standardisation                                 (LAPACK dlahqr)
// These restate the published LAPACK 3.12 algorithms that the reference reaches through its single
// numerical call, dgeev_ (reference src/lapack-wrapper.hpp:47-60, src/paladin.hpp:248-254), so that
// eigenvalue order and the conjugate-pair convention consumed at src/paladin.hpp:262-263 are kept.
// All indices are 0-based and inclusive (ilo..ihi).
#pragma once

#include "pb_common.cuh"

namespace pb {

#define PB_A( i, j ) A[( i ) + (size_t)( j ) * lda]
#define PB_H( i, j ) H[( i ) + (size_t)( j ) * ldh]
#define PB_Z( i, j ) Z[( i ) + (size_t)( j ) * ldz]
#define PB_Q( i, j ) Q[( i ) + (size_t)( j ) * ldq]

// ------------------------------------------------------------------------------------------------
// gebal. icnt: 2*n ints of scratch (row / column off-diagonal nonzero counts).
// On exit scale[] holds permutation indices outside ilo..ihi (0-based) and scaling factors inside.
// ------------------------------------------------------------------------------------------------
// work: optional 4 n doubles visible to the group. With it every sweep of the scaling loop starts from row / column sums
// computed in two coalesced passes over the matrix (rows: a thread owns rows t, t + nt, ...; columns: one group
// reduction per column) and keeps using them until the first row / column of the sweep is actually rescaled; from there on
// the sums are taken per index as without work. A dense matrix that needs no scaling (the N(0,1) workloads) costs two
// passes instead of n strided row walks: 91 -> about 15 ms at n = 3072 on one CTA.
template <class G>
PB_D void gebal( const G& g, int n, double* A, int lda, int* ilo_out, int* ihi_out, double* scale,
                 int* icnt, double* work = nullptr ) {
  const int t = g.tid(), nt = g.size();
  int* rowcnt = icnt;
  int* colcnt = icnt + n;
  int k = 0, l = n - 1;
  if ( n == 0 ) {
    *ilo_out = 0;
    *ihi_out = -1;
    return;
  }
  // off-diagonal nonzero counts
  for ( int r = t; r < n; r += nt ) {
    int cnt = 0;
    for ( int c = 0; c < n; ++c ) cnt += ( c != r && PB_A( r, c ) != 0.0 ) ? 1 : 0;
    rowcnt[r] = cnt;
  }
  // (columns one after the other, the group along the rows: a thread-per-column walk reads 32 different cache lines per
  // warp access and alone cost 40 ms of the 84 ms prologue at n = 3072)
  for ( int c = 0; c < n; ++c ) {
    int cnt = 0;
    for ( int r = t; r < n; r += nt ) cnt += ( c != r && PB_A( r, c ) != 0.0 ) ? 1 : 0;
    cnt = g.isum( cnt );
    if ( t == 0 ) colcnt[c] = cnt;
  }
  g.sync();

  // ---- rows isolating an eigenvalue are pushed down ----
  bool noconv = true;
  bool all_isolated = false;
  while ( noconv && !all_isolated ) {
    noconv = false;
    for ( int i = l; i >= 0; --i ) {
      if ( rowcnt[i] != 0 ) continue;  // uniform: every thread reads the same shared value
      g.sync();
      if ( t == 0 ) scale[l] = (double)i;
      if ( i != l ) {
        for ( int r = t; r <= l; r += nt ) {
          const double tmp = PB_A( r, i );
          PB_A( r, i ) = PB_A( r, l );
          PB_A( r, l ) = tmp;
        }
        g.sync();
        for ( int c = k + t; c < n; c += nt ) {
          const double tmp = PB_A( i, c );
          PB_A( i, c ) = PB_A( l, c );
          PB_A( l, c ) = tmp;
        }
        if ( t == 0 ) {
          int tmp = rowcnt[i];
          rowcnt[i] = rowcnt[l];
          rowcnt[l] = tmp;
          tmp = colcnt[i];
          colcnt[i] = colcnt[l];
          colcnt[l] = tmp;
        }
      }
      g.sync();
      noconv = true;
      if ( l == 0 ) {
        all_isolated = true;
        break;
      }
      // column l leaves the active set
      for ( int r = t; r < l; r += nt )
        if ( PB_A( r, l ) != 0.0 ) rowcnt[r] -= 1;
      l -= 1;
      g.sync();
    }
  }
  if ( all_isolated ) {
    *ilo_out = 0;
    *ihi_out = 0;
    return;
  }

  // ---- columns isolating an eigenvalue are pushed left ----
  noconv = true;
  while ( noconv ) {
    noconv = false;
    const int jend = l;
    for ( int j = k; j <= jend; ++j ) {
      if ( j < k || colcnt[j] != 0 ) continue;
      g.sync();
      if ( t == 0 ) scale[k] = (double)j;
      if ( j != k ) {
        for ( int r = t; r <= l; r += nt ) {
          const double tmp = PB_A( r, j );
          PB_A( r, j ) = PB_A( r, k );
          PB_A( r, k ) = tmp;
        }
        g.sync();
        for ( int c = k + t; c < n; c += nt ) {
          const double tmp = PB_A( j, c );
          PB_A( j, c ) = PB_A( k, c );
          PB_A( k, c ) = tmp;
        }
        if ( t == 0 ) {
          const int tmp = colcnt[j];
          colcnt[j] = colcnt[k];
          colcnt[k] = tmp;
        }
      }
      g.sync();
      noconv = true;
      // row k leaves the active set
      for ( int c = k + 1 + t; c <= l; c += nt )
        if ( PB_A( k, c ) != 0.0 ) colcnt[c] -= 1;
      k += 1;
      g.sync();
    }
  }

  // ---- diagonal scaling of rows/columns k..l (power-of-two factors) ----
  for ( int i = k + t; i <= l; i += nt ) scale[i] = 1.0;
  g.sync();
  const double sclfac = 2.0, factor = 0.95;
  const double sfmin1 = kSafmin / kUlp, sfmax1 = 1.0 / sfmin1;
  const double sfmin2 = sfmin1 * sclfac, sfmax2 = 1.0 / sfmin2;
  // overflow-safe sums of squares: entries are divided by a power of two near the largest magnitude
  double amax = 0.0;
  for ( int c = k; c <= l; ++c )
    for ( int r = k + t; r <= l; r += nt ) amax = dmax( amax, fabs( PB_A( r, c ) ) );
  amax = g.max( amax );
  if ( amax == 0.0 ) {
    *ilo_out = k;
    *ihi_out = l;
    return;
  }
  int ex = 0;
  (void)frexp( amax, &ex );
  const double down = ldexp( 1.0, -ex ), up = ldexp( 1.0, ex );

  noconv = true;
  int guard = 0;
  while ( noconv && guard++ < 200 ) {
    noconv = false;
    bool pre = ( work != nullptr ) && ( l - k + 1 >= 128 );
    if ( pre ) {
      double* pcs = work;
      double* prs = work + n;
      double* pca = work + 2 * n;
      double* pra = work + 3 * n;
      // rows: thread-owned, columns walked in order (coalesced across the group)
      for ( int r0 = k; r0 <= l; r0 += nt ) {
        const int r = r0 + t;
        double s2 = 0.0, mx = 0.0;
        if ( r <= l ) {
          for ( int c = k; c < n; ++c ) {
            const double a = PB_A( r, c );
            if ( c <= l ) {
              const double x = a * down;
              s2 += x * x;
            }
            mx = dmax( mx, fabs( a ) );
          }
          prs[r] = s2;
          pra[r] = mx;
        }
      }
      // columns: one reduction each
      for ( int c = k; c <= l; ++c ) {
        double s2 = 0.0, mx = 0.0;
        for ( int r = t; r <= l; r += nt ) {
          const double a = PB_A( r, c );
          if ( r >= k ) {
            const double x = a * down;
            s2 += x * x;
          }
          mx = dmax( mx, fabs( a ) );
        }
        s2 = g.sum( s2 );
        mx = g.max( mx );
        if ( t == 0 ) {
          pcs[c] = s2;
          pca[c] = mx;
        }
      }
      g.sync();
    }
    for ( int i = k; i <= l; ++i ) {
      double cs = 0.0, rs = 0.0, ca = 0.0, ra = 0.0;
      double c, r;
      if ( pre ) {
        // the totals of this sweep's pre-pass: every thread reads them, no reduction (and no barrier) per index
        c = sqrt( work[i] ) * up;
        r = sqrt( work[n + i] ) * up;
        ca = work[2 * n + i];
        ra = work[3 * n + i];
      } else {
        for ( int rr = k + t; rr <= l; rr += nt ) {
          const double x = PB_A( rr, i ) * down;
          cs += x * x;
        }
        for ( int cc = k + t; cc <= l; cc += nt ) {
          const double x = PB_A( i, cc ) * down;
          rs += x * x;
        }
        for ( int rr = t; rr <= l; rr += nt ) ca = dmax( ca, fabs( PB_A( rr, i ) ) );
        for ( int cc = k + t; cc < n; cc += nt ) ra = dmax( ra, fabs( PB_A( i, cc ) ) );
        c = sqrt( g.sum( cs ) ) * up;
        r = sqrt( g.sum( rs ) ) * up;
        ca = g.max( ca );
        ra = g.max( ra );
      }
      if ( c == 0.0 || r == 0.0 ) continue;
      double gg = r / sclfac;
      double f = 1.0;
      const double s = c + r;
      while ( c < gg && dmax( f, dmax( c, ca ) ) < sfmax2 && dmin( r, dmin( gg, ra ) ) > sfmin2 ) {
        f *= sclfac;
        c *= sclfac;
        ca *= sclfac;
        r /= sclfac;
        gg /= sclfac;
        ra /= sclfac;
      }
      gg = c / sclfac;
      while ( gg >= r && dmax( r, ra ) < sfmax2 && dmin( dmin( f, c ), dmin( gg, ca ) ) > sfmin2 ) {
        f /= sclfac;
        c /= sclfac;
        gg /= sclfac;
        ca /= sclfac;
        r *= sclfac;
        ra *= sclfac;
      }
      if ( ( c + r ) >= factor * s ) continue;
      const double sci = scale[i];
      if ( f < 1.0 && sci < 1.0 ) {
        if ( f * sci <= sfmin1 ) continue;
      }
      if ( f > 1.0 && sci > 1.0 ) {
        if ( sci >= sfmax1 / f ) continue;
      }
      const double ginv = 1.0 / f;
      pre = false;  // the precomputed sums are stale from here on
      g.sync();
      if ( t == 0 ) scale[i] = sci * f;
      for ( int cc = k + t; cc < n; cc += nt ) PB_A( i, cc ) *= ginv;
      g.sync();
      for ( int rr = t; rr <= l; rr += nt ) PB_A( rr, i ) *= f;
      g.sync();
      noconv = true;
    }
  }
  *ilo_out = k;
  *ihi_out = l;
}

// ------------------------------------------------------------------------------------------------
// Householder generation for a column segment x = A(r0+1:r1, c), alpha = A(r0, c)  (dlarfg).
// Leaves v (scaled x) in place, returns beta and tau; A(r0,c) is NOT modified (caller stores beta).
// ------------------------------------------------------------------------------------------------
template <class G>
PB_D void larfg_col( const G& g, double* A, int lda, int r0, int r1, int c, double* beta_out,
                     double* tau_out ) {
  const int t = g.tid(), nt = g.size();
  double alpha = PB_A( r0, c );
  if ( r1 <= r0 ) {
    *beta_out = alpha;
    *tau_out = 0.0;
    return;
  }
  double mx = 0.0;
  for ( int r = r0 + 1 + t; r <= r1; r += nt ) mx = dmax( mx, fabs( PB_A( r, c ) ) );
  mx = g.max( mx );
  if ( mx == 0.0 ) {
    *beta_out = alpha;
    *tau_out = 0.0;
    return;
  }
  double ss = 0.0;
  for ( int r = r0 + 1 + t; r <= r1; r += nt ) {
    const double x = PB_A( r, c ) / mx;
    ss += x * x;
  }
  double xnorm = mx * sqrt( g.sum( ss ) );
  double beta = -dsign( dlapy2( alpha, xnorm ), alpha );
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
      alpha *= rsafmn;
    } while ( fabs( beta ) < safmin && knt < 20 );
    beta = -dsign( dlapy2( alpha, xnorm ), alpha );
  }
  const double tau = ( beta - alpha ) / beta;
  const double sc = pre / ( alpha - beta );
  g.sync();
  for ( int r = r0 + 1 + t; r <= r1; r += nt ) PB_A( r, c ) *= sc;
  for ( int j = 0; j < knt; ++j ) beta *= safmin;
  g.sync();
  *beta_out = beta;
  *tau_out = tau;
}

// ------------------------------------------------------------------------------------------------
// gehd2: A(0:n,0:n) -> upper Hessenberg in rows/cols ilo..ihi, reflectors below the subdiagonal.
// wrk: n doubles of scratch visible to the whole group.
// ------------------------------------------------------------------------------------------------
template <class G>
PB_D void gehd2( const G& g, int n, int ilo, int ihi, double* A, int lda, double* tau, double* wrk ) {
  const int t = g.tid(), nt = g.size();
  for ( int j = ilo; j < ihi; ++j ) {
    double beta, tj;
    larfg_col( g, A, lda, j + 1, ihi, j, &beta, &tj );
    if ( t == 0 ) tau[j] = tj;
    if ( tj != 0.0 ) {
      // v = [1; A(j+2:ihi, j)]
      // right: A(0:ihi, j+1:ihi) -= tau * (A v) v^T
      for ( int r = t; r <= ihi; r += nt ) {
        double w = PB_A( r, j + 1 );
        for ( int c = j + 2; c <= ihi; ++c ) w += PB_A( r, c ) * PB_A( c, j );
        w *= tj;
        PB_A( r, j + 1 ) -= w;
        for ( int c = j + 2; c <= ihi; ++c ) PB_A( r, c ) -= w * PB_A( c, j );
      }
      g.sync();
      // left: A(j+1:ihi, j+1:n) -= tau * v (v^T A)
      for ( int c = j + 1 + t; c < n; c += nt ) {
        double w = PB_A( j + 1, c );
        for ( int r = j + 2; r <= ihi; ++r ) w += PB_A( r, j ) * PB_A( r, c );
        w *= tj;
        PB_A( j + 1, c ) -= w;
        for ( int r = j + 2; r <= ihi; ++r ) PB_A( r, c ) -= w * PB_A( r, j );
      }
    }
    g.sync();
    if ( t == 0 ) PB_A( j + 1, j ) = beta;
    g.sync();
  }
  (void)wrk;
}

// ------------------------------------------------------------------------------------------------
// orghr: Q (n x n, ldq) from reflectors stored below the subdiagonal of A (columns ilo..ihi-1).
// Q = H(ilo) H(ilo+1) ... H(ihi-1); rows/cols outside ilo..ihi are identity (dorghr semantics).
// ------------------------------------------------------------------------------------------------
template <class G>
PB_D void orghr( const G& g, int n, int ilo, int ihi, const double* A, int lda, const double* tau,
                 double* Q, int ldq ) {
  const int t = g.tid(), nt = g.size();
  for ( int c = 0; c < n; ++c )
    for ( int r = t; r < n; r += nt ) PB_Q( r, c ) = ( r == c ) ? 1.0 : 0.0;
  g.sync();
  // apply H(j) from the left to Q(j+1:ihi, j+1:ihi), j descending (backward accumulation)
  for ( int j = ihi - 1; j >= ilo; --j ) {
    const double tj = tau[j];
    if ( tj == 0.0 ) continue;
    for ( int c = j + 1 + t; c <= ihi; c += nt ) {
      double w = PB_Q( j + 1, c );
      for ( int r = j + 2; r <= ihi; ++r ) w += PB_A( r, j ) * PB_Q( r, c );
      w *= tj;
      PB_Q( j + 1, c ) -= w;
      for ( int r = j + 2; r <= ihi; ++r ) PB_Q( r, c ) -= w * PB_A( r, j );
    }
    g.sync();
  }
}

// H(i, j) with separate row and column strides (hrs = 1 everywhere except where two matrices share one array, see lahqr)
#define PB_HH( i, j ) H[(size_t)( i ) * hrs + (size_t)( j ) * ldh]

// ------------------------------------------------------------------------------------------------
// One Francis double-shift sweep on ONE WARP with the data a step depends on held in registers.
//
// LAPACK's loop (the generic code in lahqr below) is one dependent chain per bulge position k: bulge column from memory ->
// reflector -> row update -> barrier -> column update -> barrier -> next bulge column: 1,300 (eigenvalues only) to 1,900
// (Schur form + vectors) cycles per step on a B200 warp whatever the order (32..64), of which the reflector is 280 and
// the rest is branches, address arithmetic and shared-memory round trips (tools/mb_lahqr.cu). Here
//   * every lane tracks the 4 x 3 diagonal block the bulge lives in (rows k..k+3, columns k..k+2, and the bulge column) in
//     registers, redundantly and bit-identically, so reflector k+1 follows from reflector k by register arithmetic alone;
//   * lane j owns COLUMN j for the row updates right of the block (rows k, k+1 carried in registers from step to step:
//     one load, one store per step), lane r owns ROW r for the column updates above the block and row r of Z (likewise);
//     the block's own entries never go through memory: they enter it from the owning lane by a shuffle (column k+3) and
//     leave it to the lane that owns row k in registers;
//   * these updates run ONE STEP BEHIND the reflectors (the updates of step k-1 are independent of the generation of
//     reflector k and are issued next to it), and nothing written in a step is read in the same step: one __syncwarp per step.
// The matrix goes through exactly LAPACK's sequence of reflections; only the association of the (exactly orthogonal)
// products differs, i.e. rounding. Requires i2 <= 32 SLOTS - 1 and ihiz <= 32 SLOTS - 1.
//
// Lane-explicit code: NL = 1 on the device (this thread is lane g.tid()), NL = 32 in the host emulation (tests/hostsim:
// one thread walks the 32 lanes phase by phase, which is what the warp does in lockstep).
// ------------------------------------------------------------------------------------------------
template <class G, class = void>
struct GroupLanes {
  static constexpr int value = 0;
};
template <class G>
struct GroupLanes<G, decltype( (void)G::kLanes )> {
  static constexpr int value = G::kLanes;
};

struct SweepStepGuarded {
  static constexpr bool value = false;
};
struct SweepStepInside {
  static constexpr bool value = true;
};
// SPLIT: the steps whose block lies strictly inside the active part (k <= i - 4) run in a loop of their own without the
// end-of-block guards (about a sixth of the instructions of a step; worth it where the sweep is bound by issue slots: the
// values-only n <= 64 kernel at twelve warps per SM)
template <class G, int SLOTS, bool WANTZ, bool FASTREFL, bool SPLIT = false>
PB_D void francis_sweep_lanes( const G& g, int l, int i, int m, double v0, double v1, double v2, int i1, int i2, double* H, int ldh,
                               int iloz, int ihiz, double* Z, int ldz, int hrs = 1 ) {
  constexpr int NL = GroupLanes<G>::value;
  static_assert( NL == 1 || NL == 32, "one lane per thread, or all 32 on one host thread" );
#define PB_FOR_LANES for ( int li = 0; li < NL; ++li )
#define PB_LANE ( NL == 1 ? g.tid() : li )
  // (entries below the first subdiagonal are zero at the start of a sweep and never stored by it: not even read)
  auto hg = [&]( int r, int c ) -> double { return ( r <= i && c <= i && r <= c + 1 ) ? PB_HH( r, c ) : 0.0; };
  // ---- tracked block (uniform) ----
  double b0 = v0, b1 = v1, b2 = v2;
  double c00 = hg( m, m ), c01 = hg( m, m + 1 ), c02 = hg( m, m + 2 );
  double c10 = hg( m + 1, m ), c11 = hg( m + 1, m + 1 ), c12 = hg( m + 1, m + 2 );
  double c20 = hg( m + 2, m ), c21 = hg( m + 2, m + 1 ), c22 = hg( m + 2, m + 2 );
  double c30 = hg( m + 3, m ), c31 = hg( m + 3, m + 1 ), c32 = hg( m + 3, m + 2 );
  // ---- per-lane carried entries: x = H(k, j), H(k+1, j) of the lane's column j; y = H(r, k), H(r, k+1) of the lane's
  // row r; z likewise for Z (k = the next step to be applied to memory) ----
  double x0[NL][SLOTS], x1[NL][SLOTS], y0[NL][SLOTS], y1[NL][SLOTS], z0[NL][SLOTS], z1[NL][SLOTS];
  PB_FOR_LANES {
#pragma unroll
    for ( int s = 0; s < SLOTS; ++s ) {
      const int idx = PB_LANE + 32 * s;
      const bool xa = idx >= m && idx <= i2;
      x0[li][s] = xa ? PB_HH( m, idx ) : 0.0;
      x1[li][s] = xa ? PB_HH( m + 1, idx ) : 0.0;
      const bool ya = idx >= i1 && idx <= m - 1;
      y0[li][s] = ya ? PB_HH( idx, m ) : 0.0;
      y1[li][s] = ya ? PB_HH( idx, m + 1 ) : 0.0;
      if ( WANTZ ) {
        const bool za = idx >= iloz && idx <= ihiz;
        z0[li][s] = za ? PB_Z( idx, m ) : 0.0;
        z1[li][s] = za ? PB_Z( idx, m + 1 ) : 0.0;
      }
    }
  }
  // entry of the lane that owns column c (uniform result)
  auto from_owner = [&]( const double ( &a )[NL][SLOTS], int c ) -> double {
    const int s = c >> 5;
#if defined( __CUDA_ARCH__ )
    double v = a[0][0];
    if ( SLOTS > 1 && s == 1 ) v = a[0][SLOTS - 1];
    return __shfl_sync( 0xffffffffu, v, c & 31 );
#else
    return a[NL == 1 ? 0 : ( c & 31 )][s < SLOTS ? s : 0];
#endif
  };
  // pending step (reflector generated, stored matrix not yet updated) and the finished top row of its block
  int p = m, pnr = 0;
  double pw2 = 0.0, pw3 = 0.0, pt1 = 0.0, pbeta = 0.0, pf00 = 0.0, pf01 = 0.0, pf02 = 0.0, qf01 = 0.0, qf02 = 0.0;
  // reflector of step k from the tracked bulge column, then step k on the tracked block (column k+3 joins it)
  auto step = [&]( auto inside, int k ) {
    constexpr bool kInside = decltype( inside )::value;  // k + 4 <= i: nothing of the block or of column k+3 is outside
    const int nr = kInside ? 3 : imin( 3, i - k + 1 );
    double a0 = b0, a1 = b1, a2 = b2, t1;
    if ( FASTREFL ) larfg3_bf( nr, a0, a1, a2, t1 );
    else dlarfg3( nr, a0, a1, a2, t1 );
    if ( nr < 3 ) a2 = 0.0;
    const bool din = kInside || k + 3 <= i;
    double d0 = from_owner( x0, k + 3 ), d1 = from_owner( x1, k + 3 );
    d0 = din ? d0 : 0.0;
    d1 = din ? d1 : 0.0;
    double d2 = kInside ? PB_HH( k + 2, k + 3 ) : hg( k + 2, k + 3 );
    const double d3 = kInside ? PB_HH( k + 3, k + 3 ) : hg( k + 3, k + 3 ), d4 = kInside ? PB_HH( k + 4, k + 3 ) : hg( k + 4, k + 3 );
    const double w2 = a1, w3 = a2, t2 = t1 * w2, t3 = t1 * w3;
    double s = c00 + w2 * c10 + w3 * c20;
    c00 -= s * t1;
    c10 -= s * t2;
    c20 -= s * t3;
    s = c01 + w2 * c11 + w3 * c21;
    c01 -= s * t1;
    c11 -= s * t2;
    c21 -= s * t3;
    s = c02 + w2 * c12 + w3 * c22;
    c02 -= s * t1;
    c12 -= s * t2;
    c22 -= s * t3;
    s = d0 + w2 * d1 + w3 * d2;
    d1 -= s * t2;
    d2 -= s * t3;
    s = c00 + w2 * c01 + w3 * c02;
    qf01 = pf01;
    qf02 = pf02;
    pf00 = c00 - s * t1;
    pf01 = c01 - s * t2;
    pf02 = c02 - s * t3;
    s = c10 + w2 * c11 + w3 * c12;
    c10 -= s * t1;
    c11 -= s * t2;
    c12 -= s * t3;
    s = c20 + w2 * c21 + w3 * c22;
    c20 -= s * t1;
    c21 -= s * t2;
    c22 -= s * t3;
    s = c30 + w2 * c31 + w3 * c32;
    c30 -= s * t1;
    c31 -= s * t2;
    c32 -= s * t3;
    b0 = c10;
    b1 = c20;
    b2 = c30;
    c00 = c11; c01 = c12; c02 = d1;
    c10 = c21; c11 = c22; c12 = d2;
    c20 = c31; c21 = c32; c22 = d3;
    c30 = 0.0; c31 = 0.0; c32 = d4;
    p = k;
    pnr = nr;
    pw2 = w2;
    pw3 = w3;
    pt1 = t1;
    pbeta = a0;
  };
  // the updates of step P (parameters W2, W3, T1; QF = top-row entries of the block of step P-1) on the stored matrix
  auto apply = [&]( auto inside, int P, double W2, double W3, double T1, double BETA, double F00, double QF1, double QF2 ) {
    const double T2 = T1 * W2, T3 = T1 * W3;
    const bool third = decltype( inside )::value || P + 2 <= i;
    PB_FOR_LANES {
#pragma unroll
      for ( int s = 0; s < SLOTS; ++s ) {
        const int idx = PB_LANE + 32 * s;
        {  // rows P..P+2 of column idx
          const bool act = idx <= i2 && ( idx >= P + 3 || idx > i );
          const double e0 = x0[li][s], e1 = x1[li][s];
          const double e2 = ( act && third ) ? PB_HH( P + 2, idx ) : 0.0;
          const double sum = e0 + W2 * e1 + W3 * e2;
          if ( act ) PB_HH( P, idx ) = e0 - sum * T1;
          x0[li][s] = e1 - sum * T2;
          x1[li][s] = e2 - sum * T3;
        }
        {  // columns P..P+2 of row idx (row P-1 joins with the entries the tracked block finished one step earlier)
          const bool act = idx >= i1 && idx <= P - 1;
          const bool join = P > m && idx == P - 1;
          const double e0 = join ? QF1 : y0[li][s], e1 = join ? QF2 : y1[li][s];
          const double e2 = ( act && third ) ? PB_HH( idx, P + 2 ) : 0.0;
          const double sum = e0 + W2 * e1 + W3 * e2;
          if ( act ) PB_HH( idx, P ) = e0 - sum * T1;
          y0[li][s] = e1 - sum * T2;
          y1[li][s] = e2 - sum * T3;
        }
        if ( WANTZ ) {
          const bool act = idx >= iloz && idx <= ihiz;
          const double e0 = z0[li][s], e1 = z1[li][s];
          const double e2 = ( act && third ) ? PB_Z( idx, P + 2 ) : 0.0;
          const double sum = e0 + W2 * e1 + W3 * e2;
          if ( act ) PB_Z( idx, P ) = e0 - sum * T1;
          z0[li][s] = e1 - sum * T2;
          z1[li][s] = e2 - sum * T3;
        }
      }
    }
    if ( g.tid() == 0 ) {
      PB_HH( P, P ) = F00;
      if ( P > m ) PB_HH( P, P - 1 ) = BETA;
    }
  };
  // ---- first step: nothing pending yet ----
  using Guarded = SweepStepGuarded;
  using Inside = SweepStepInside;
  step( Guarded(), m );
  if ( g.tid() == 0 && m > l ) PB_HH( m, m - 1 ) = PB_HH( m, m - 1 ) * ( 1.0 - pt1 );
  int k = m + 1;
  if constexpr ( SPLIT ) {
    for ( ; k <= i - 4; ++k ) {
      const int P = p;
      const double W2 = pw2, W3 = pw3, T1 = pt1, BETA = pbeta, F00 = pf00, QF1 = qf01, QF2 = qf02;
      apply( Inside(), P, W2, W3, T1, BETA, F00, QF1, QF2 );
      step( Inside(), k );
      g.sync();
    }
  }
  for ( ; k <= i - 1; ++k ) {
    const int P = p;
    const double W2 = pw2, W3 = pw3, T1 = pt1, BETA = pbeta, F00 = pf00, QF1 = qf01, QF2 = qf02;
    apply( Guarded(), P, W2, W3, T1, BETA, F00, QF1, QF2 );
    step( Guarded(), k );
    g.sync();
  }
  apply( Guarded(), p, pw2, pw3, pt1, pbeta, pf00, qf01, qf02 );
  // ---- what is still in registers after the last step (p = i-1, a 2 x 2 reflection) ----
  PB_FOR_LANES {
#pragma unroll
    for ( int s = 0; s < SLOTS; ++s ) {
      const int idx = PB_LANE + 32 * s;
      if ( idx > i && idx <= i2 ) PB_HH( i, idx ) = x0[li][s];
      if ( idx >= i1 && idx <= i - 2 ) PB_HH( idx, i ) = y0[li][s];
      if ( WANTZ ) {
        if ( idx >= iloz && idx <= ihiz ) PB_Z( idx, i ) = z0[li][s];
      }
    }
  }
  if ( g.tid() == 0 ) {
    PB_HH( i - 1, i ) = pf01;
    PB_HH( i, i - 1 ) = b0;
    PB_HH( i, i ) = c00;
  }
  g.sync();
#undef PB_FOR_LANES
#undef PB_LANE
}

// ------------------------------------------------------------------------------------------------
// lahqr. H upper Hessenberg in rows/cols ilo..ihi of an n x n array. With wantt the full Schur form
// is produced (updates extend over 0..n-1), otherwise only eigenvalues. With wantz the same
// transformations are accumulated into rows iloz..ihiz of Z.
// Returns 0, or (i+1) when the QR iteration failed to converge for eigenvalue index i.
// ------------------------------------------------------------------------------------------------
// FASTREFL: generate the 3-element reflectors with larfg3_fast (one sqrt, two reciprocals) instead of the
// rescaling dlarfg3; used where the result only steers the iteration (shift computation).
// HOOK: called (by every thread, after a barrier) each time the bottom block [l..i] of the active part has converged;
// when it returns true the iteration stops and -(l+1) is returned: rows/columns l.. are in Schur form, the rest is
// still upper Hessenberg (used by the early-deflation window, pb_aed.cuh).
struct LahqrNoHook {
  static constexpr bool kActive = false;
  PB_HD bool operator()( int, int ) const { return false; }
};
template <class G, bool FASTREFL = false, class HOOK = LahqrNoHook, bool SPLIT = false>
// hrs / band_only: H(i, j) at H[i * hrs + j * ldh], and only the entries with i <= j + 1 exist (two Hessenberg matrices share
// one array, the second one transposed: the values-only n <= 64 kernel, paladin_gpu.cu::small_qr_pair_kernel). Needs the
// register-resident sweep (a single warp, wantt == false): it is the one that never stores a bulge.
PB_D int lahqr( const G& g, bool wantt, bool wantz, int n, int ilo, int ihi, double* H, int ldh,
                double* wr, double* wi, int iloz, int ihiz, double* Z, int ldz, const HOOK& hook = HOOK(), int hrs = 1,
                bool band_only = false ) {
  const int t = g.tid(), nt = g.size();
  // the Schur-vector loops start half a group away from the column loops they follow: in a group wider than the
  // block both run side by side
  const int tz = ( t + nt / 2 ) % nt;
  if ( n == 0 || ihi < ilo ) return 0;
  if ( ilo == ihi ) {
    if ( t == 0 ) {
      wr[ilo] = PB_HH( ilo, ilo );
      wi[ilo] = 0.0;
    }
    g.sync();
    return 0;
  }
  // clear out the trash below the subdiagonal
  if ( !band_only ) {
    for ( int j = ilo + t; j <= ihi - 3; j += nt ) {
      PB_HH( j + 2, j ) = 0.0;
      PB_HH( j + 3, j ) = 0.0;
    }
    if ( t == 0 && ilo <= ihi - 2 ) PB_HH( ihi, ihi - 2 ) = 0.0;
  }
  g.sync();

  const int nh = ihi - ilo + 1;
  const double smlnum = kSafmin * ( (double)nh / kUlp );
  const double dat1 = 0.75, dat2 = -0.4375;
  const int kexsh = 10;
  int i1 = 0, i2 = n - 1;
  const int itmax = 30 * imax( 10, nh );
  int kdefl = 0;

  int i = ihi;
  while ( i >= ilo ) {
    int l = ilo;
    bool converged = false;
    for ( int its = 0; its <= itmax; ++its ) {
      // ---- look for a single small subdiagonal element (largest k in l+1..i that qualifies) ----
      int kfound = l;
      for ( int k = i - t; k > l; k -= nt ) {
        const double hkk1 = fabs( PB_HH( k, k - 1 ) );
        bool small = false;
        if ( hkk1 <= smlnum ) {
          small = true;
        } else {
          double tst = fabs( PB_HH( k - 1, k - 1 ) ) + fabs( PB_HH( k, k ) );
          if ( tst == 0.0 ) {
            if ( k - 2 >= ilo ) tst += fabs( PB_HH( k - 1, k - 2 ) );
            if ( k + 1 <= ihi ) tst += fabs( PB_HH( k + 1, k ) );
          }
          if ( hkk1 <= kUlp * tst ) {
            const double hk1k = fabs( PB_HH( k - 1, k ) );
            const double ab = dmax( hkk1, hk1k ), ba = dmin( hkk1, hk1k );
            const double dd = fabs( PB_HH( k - 1, k - 1 ) - PB_HH( k, k ) );
            const double aa = dmax( fabs( PB_HH( k, k ) ), dd ), bb = dmin( fabs( PB_HH( k, k ) ), dd );
            const double s = aa + ab;
            if ( ba * ( ab / s ) <= dmax( smlnum, kUlp * ( bb * ( aa / s ) ) ) ) small = true;
          }
        }
        if ( small ) {
          kfound = k;
          break;  // this thread's k values only decrease from here
        }
      }
      l = g.imax_( kfound );
      if ( l > ilo ) {
        if ( t == 0 ) PB_HH( l, l - 1 ) = 0.0;
      }
      g.sync();
      if ( l >= i - 1 ) {
        converged = true;
        break;
      }
      kdefl += 1;
      if ( !wantt ) {
        i1 = l;
        i2 = i;
      }
      double h11, h21, h12, h22;
      if ( kdefl % ( 2 * kexsh ) == 0 ) {
        const double s = fabs( PB_HH( i, i - 1 ) ) + fabs( PB_HH( i - 1, i - 2 ) );
        h11 = dat1 * s + PB_HH( i, i );
        h12 = dat2 * s;
        h21 = s;
        h22 = h11;
      } else if ( kdefl % kexsh == 0 ) {
        const double s = fabs( PB_HH( l + 1, l ) ) + fabs( PB_HH( l + 2, l + 1 ) );
        h11 = dat1 * s + PB_HH( l, l );
        h12 = dat2 * s;
        h21 = s;
        h22 = h11;
      } else {
        h11 = PB_HH( i - 1, i - 1 );
        h21 = PB_HH( i, i - 1 );
        h12 = PB_HH( i - 1, i );
        h22 = PB_HH( i, i );
      }
      double rt1r, rt1i, rt2r, rt2i;
      {
        const double s = fabs( h11 ) + fabs( h12 ) + fabs( h21 ) + fabs( h22 );
        if ( s == 0.0 ) {
          rt1r = rt1i = rt2r = rt2i = 0.0;
        } else {
          h11 /= s;
          h21 /= s;
          h12 /= s;
          h22 /= s;
          const double tr = ( h11 + h22 ) / 2.0;
          const double det = ( h11 - tr ) * ( h22 - tr ) - h12 * h21;
          const double rtdisc = sqrt( fabs( det ) );
          if ( det >= 0.0 ) {
            rt1r = tr * s;
            rt2r = rt1r;
            rt1i = rtdisc * s;
            rt2i = -rt1i;
          } else {
            rt1r = tr + rtdisc;
            rt2r = tr - rtdisc;
            if ( fabs( rt1r - h22 ) <= fabs( rt2r - h22 ) ) {
              rt1r = rt1r * s;
              rt2r = rt1r;
            } else {
              rt2r = rt2r * s;
              rt1r = rt2r;
            }
            rt1i = rt2i = 0.0;
          }
        }
      }
      // ---- look for two consecutive small subdiagonal elements (largest m in l..i-2) ----
      int mfound = l;
      for ( int m = i - 2 - t; m > l; m -= nt ) {
        double h21s = fabs( PB_HH( m + 1, m ) );
        double s = fabs( PB_HH( m, m ) - rt2r ) + fabs( rt2i ) + h21s;
        h21s = PB_HH( m + 1, m ) / s;
        double v0 = h21s * PB_HH( m, m + 1 ) + ( PB_HH( m, m ) - rt1r ) * ( ( PB_HH( m, m ) - rt2r ) / s ) -
                    rt1i * ( rt2i / s );
        double v1 = h21s * ( PB_HH( m, m ) + PB_HH( m + 1, m + 1 ) - rt1r - rt2r );
        double v2 = h21s * PB_HH( m + 2, m + 1 );
        s = fabs( v0 ) + fabs( v1 ) + fabs( v2 );
        v0 /= s;
        v1 /= s;
        v2 /= s;
        const double h00 = fabs( PB_HH( m - 1, m - 1 ) ), h11a = fabs( PB_HH( m, m ) ),
                     h22a = fabs( PB_HH( m + 1, m + 1 ) );
        if ( fabs( PB_HH( m, m - 1 ) ) * ( fabs( v1 ) + fabs( v2 ) ) <= kUlp * fabs( v0 ) * ( h00 + h11a + h22a ) ) {
          mfound = m;
          break;
        }
      }
      const int m = g.imax_( mfound );
      double v0, v1, v2;
      {
        double h21s = fabs( PB_HH( m + 1, m ) );
        double s = fabs( PB_HH( m, m ) - rt2r ) + fabs( rt2i ) + h21s;
        h21s = PB_HH( m + 1, m ) / s;
        v0 = h21s * PB_HH( m, m + 1 ) + ( PB_HH( m, m ) - rt1r ) * ( ( PB_HH( m, m ) - rt2r ) / s ) -
             rt1i * ( rt2i / s );
        v1 = h21s * ( PB_HH( m, m ) + PB_HH( m + 1, m + 1 ) - rt1r - rt2r );
        v2 = h21s * PB_HH( m + 2, m + 1 );
        s = fabs( v0 ) + fabs( v1 ) + fabs( v2 );
        v0 /= s;
        v1 /= s;
        v2 /= s;
      }
      g.sync();  // every thread has finished reading H for the shift vector before the sweep writes
      // ---- double-shift QR sweep ----
#ifndef PB_LAHQR_NO_LANES
      if constexpr ( GroupLanes<G>::value > 0 ) {
        // one warp, everything the updates touch within 64 columns / rows: the register-resident sweep
        const int top = wantz ? imax( i2, ihiz ) : i2;
        if ( top < 64 ) {
#ifdef PB_AED_COUNT
          for ( int k = m; k <= i - 1; ++k ) PB_AED_COUNT( wantt ? 1 : 2 );
          PB_AED_COUNT( 3 );
#endif
#ifdef PB_SWEEP_PROFILE  // (defined by tools/mb_lahqr.cu only: cycles inside the sweeps vs. the scans / shifts around them)
          PB_SWEEP_PROFILE( 0 );
#endif
          if ( top < 32 ) {
            if ( wantz ) francis_sweep_lanes<G, 1, true, FASTREFL, SPLIT>( g, l, i, m, v0, v1, v2, i1, i2, H, ldh, iloz, ihiz, Z, ldz, hrs );
            else francis_sweep_lanes<G, 1, false, FASTREFL, SPLIT>( g, l, i, m, v0, v1, v2, i1, i2, H, ldh, iloz, ihiz, Z, ldz, hrs );
          } else {
            if ( wantz ) francis_sweep_lanes<G, 2, true, FASTREFL, SPLIT>( g, l, i, m, v0, v1, v2, i1, i2, H, ldh, iloz, ihiz, Z, ldz, hrs );
            else francis_sweep_lanes<G, 2, false, FASTREFL, SPLIT>( g, l, i, m, v0, v1, v2, i1, i2, H, ldh, iloz, ihiz, Z, ldz, hrs );
          }
#ifdef PB_SWEEP_PROFILE
          PB_SWEEP_PROFILE( 1 );
#endif
          continue;
        }
      }
#endif
      for ( int k = m; k <= i - 1; ++k ) {
#ifdef PB_AED_COUNT
        PB_AED_COUNT( wantt ? 1 : 2 );
        if ( k == m ) PB_AED_COUNT( 3 );
#endif
        const int nr = imin( 3, i - k + 1 );
        double a0, a1, a2 = 0.0, t1;
        if ( k > m ) {
          a0 = PB_HH( k, k - 1 );
          a1 = PB_HH( k + 1, k - 1 );
          if ( nr == 3 ) a2 = PB_HH( k + 2, k - 1 );
        } else {
          a0 = v0;
          a1 = v1;
          a2 = v2;
        }
        if ( FASTREFL ) larfg3_fast( nr, a0, a1, a2, t1 );
        else dlarfg3( nr, a0, a1, a2, t1 );
        const double w2 = a1, t2 = t1 * w2;
        // column k-1 is rewritten by thread 0 only after the barrier that follows the left
        // application, i.e. once every thread has read it
#define PB_FIX_BULGE_COLUMN()                                \
  if ( t == 0 ) {                                            \
    if ( k > m ) {                                           \
      PB_HH( k, k - 1 ) = a0;                                 \
      PB_HH( k + 1, k - 1 ) = 0.0;                            \
      if ( k < i - 1 ) PB_HH( k + 2, k - 1 ) = 0.0;           \
    } else if ( m > l ) {                                    \
      PB_HH( k, k - 1 ) = PB_HH( k, k - 1 ) * ( 1.0 - t1 );    \
    }                                                        \
  }
        if ( nr == 3 ) {
          const double w3 = a2, t3 = t1 * w3;
          for ( int j = k + t; j <= i2; j += nt ) {
            const double sum = PB_HH( k, j ) + w2 * PB_HH( k + 1, j ) + w3 * PB_HH( k + 2, j );
            PB_HH( k, j ) -= sum * t1;
            PB_HH( k + 1, j ) -= sum * t2;
            PB_HH( k + 2, j ) -= sum * t3;
          }
          g.sync();
          PB_FIX_BULGE_COLUMN();
          const int jend = imin( k + 3, i );
          for ( int j = i1 + t; j <= jend; j += nt ) {
            const double sum = PB_HH( j, k ) + w2 * PB_HH( j, k + 1 ) + w3 * PB_HH( j, k + 2 );
            PB_HH( j, k ) -= sum * t1;
            PB_HH( j, k + 1 ) -= sum * t2;
            PB_HH( j, k + 2 ) -= sum * t3;
          }
          if ( wantz ) {
            for ( int j = iloz + tz; j <= ihiz; j += nt ) {
              const double sum = PB_Z( j, k ) + w2 * PB_Z( j, k + 1 ) + w3 * PB_Z( j, k + 2 );
              PB_Z( j, k ) -= sum * t1;
              PB_Z( j, k + 1 ) -= sum * t2;
              PB_Z( j, k + 2 ) -= sum * t3;
            }
          }
        } else if ( nr == 2 ) {
          for ( int j = k + t; j <= i2; j += nt ) {
            const double sum = PB_HH( k, j ) + w2 * PB_HH( k + 1, j );
            PB_HH( k, j ) -= sum * t1;
            PB_HH( k + 1, j ) -= sum * t2;
          }
          g.sync();
          PB_FIX_BULGE_COLUMN();
          for ( int j = i1 + t; j <= i; j += nt ) {
            const double sum = PB_HH( j, k ) + w2 * PB_HH( j, k + 1 );
            PB_HH( j, k ) -= sum * t1;
            PB_HH( j, k + 1 ) -= sum * t2;
          }
          if ( wantz ) {
            for ( int j = iloz + tz; j <= ihiz; j += nt ) {
              const double sum = PB_Z( j, k ) + w2 * PB_Z( j, k + 1 );
              PB_Z( j, k ) -= sum * t1;
              PB_Z( j, k + 1 ) -= sum * t2;
            }
          }
        }
        g.sync();
#undef PB_FIX_BULGE_COLUMN
      }
    }
    if ( !converged ) return i + 1;

    if ( l == i ) {
      if ( t == 0 ) {
        wr[i] = PB_HH( i, i );
        wi[i] = 0.0;
      }
    } else if ( l == i - 1 ) {
      double a = PB_HH( i - 1, i - 1 ), b = PB_HH( i - 1, i ), c = PB_HH( i, i - 1 ), d = PB_HH( i, i );
      double r1r, r1i, r2r, r2i, cs, sn;
      dlanv2( a, b, c, d, r1r, r1i, r2r, r2i, cs, sn );
      g.sync();
      if ( t == 0 ) {
        PB_HH( i - 1, i - 1 ) = a;
        PB_HH( i - 1, i ) = b;
        PB_HH( i, i - 1 ) = c;
        PB_HH( i, i ) = d;
        wr[i - 1] = r1r;
        wi[i - 1] = r1i;
        wr[i] = r2r;
        wi[i] = r2i;
      }
      if ( wantt ) {
        for ( int j = i + 1 + t; j <= i2; j += nt ) {
          const double x = PB_HH( i - 1, j ), y = PB_HH( i, j );
          PB_HH( i - 1, j ) = cs * x + sn * y;
          PB_HH( i, j ) = cs * y - sn * x;
        }
        for ( int j = i1 + t; j < i - 1; j += nt ) {
          const double x = PB_HH( j, i - 1 ), y = PB_HH( j, i );
          PB_HH( j, i - 1 ) = cs * x + sn * y;
          PB_HH( j, i ) = cs * y - sn * x;
        }
      }
      if ( wantz ) {
        for ( int j = iloz + tz; j <= ihiz; j += nt ) {
          const double x = PB_Z( j, i - 1 ), y = PB_Z( j, i );
          PB_Z( j, i - 1 ) = cs * x + sn * y;
          PB_Z( j, i ) = cs * y - sn * x;
        }
      }
    }
    g.sync();
    if ( HOOK::kActive ) {
      if ( hook( l, i ) ) return -( l + 1 );
    }
    kdefl = 0;
    i = l - 1;
  }
  return 0;
}

}  // namespace pb
