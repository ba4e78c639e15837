// pb_aed.cuh -- aggressive early deflation for the windowed multishift QR (pb_msqr.cuh).
//
// Part of the dhseqr stage of the reference's dgeev_ call (reference src/lapack-wrapper.hpp:47-60,
// src/paladin.hpp:248-254). LAPACK's dlaqr0 interleaves small-bulge sweeps with AED (dlaqr2/dlaqr3,
// Braman/Byers/Mathias 2002): the trailing jw x jw block T of the active Hessenberg block is brought to
// real Schur form T = V S V^T, the "spike" s V(0,:) (s = subdiagonal entry left of the window) is tested
// entry by entry from the bottom, Schur blocks whose spike entries are negligible are deflated, the others
// are swapped upwards (dtrexc/dlaexc) so the test can continue, and what is left is returned to
// Hessenberg form. The eigenvalues of the undeflated part are the shifts of the next sweep.
//
// This is a restatement for a deflation window that lives entirely in shared memory: T and V are the
// window buffers of MsqrWork, one warp (device) or the group (host simulator) runs the Schur
// factorisation, the block swaps and the re-reduction with warp-level barriers only; the off-window parts
// of H and the Schur vectors are then updated by the whole CTA as x <- V^T x per line
// (ms_update_strips). The small Sylvester equations of the block swaps are solved by Gaussian elimination
// with complete pivoting on the Kronecker form (what dlasy2 does, without its special cases).
// 0-based inclusive indices.
#pragma once

#include "pb_common.cuh"
#include "pb_small.cuh"

namespace pb {

#define PB_AT( i, j ) T[( i ) + ( j ) * ldt]
#define PB_QQ( i, j ) Q[( i ) + ( j ) * ldq]

// ---- TL X - X TR = scale B for n1, n2 in {1,2}; D (4x4, ld 4) holds [TL B; . TR] ------------------------
// Returns X (ld 2) and scale <= 1 chosen so that X cannot overflow (dlasy2 semantics, isgn = -1).
PB_HD void aed_sylvester( int n1, int n2, const double* D, double* X, double& scale ) {
  const int nn = n1 * n2;
  double K[16], b[4];
  int colperm[4];
  for ( int q = 0; q < 16; ++q ) K[q] = 0.0;
  double smin = 0.0;
  for ( int r = 0; r < n1; ++r )
    for ( int c = 0; c < n1; ++c ) smin = dmax( smin, fabs( D[r + 4 * c] ) );
  for ( int r = 0; r < n2; ++r )
    for ( int c = 0; c < n2; ++c ) smin = dmax( smin, fabs( D[( n1 + r ) + 4 * ( n1 + c )] ) );
  const double smlnum = kSafmin / kUlp;
  smin = dmax( kUlp * smin, smlnum );
  // equation (i,j), unknown (k,l), both numbered row + n1 * column
  for ( int j = 0; j < n2; ++j )
    for ( int i = 0; i < n1; ++i ) {
      const int e = i + n1 * j;
      for ( int k = 0; k < n1; ++k ) K[e + 4 * ( k + n1 * j )] += D[i + 4 * k];
      for ( int k = 0; k < n2; ++k ) K[e + 4 * ( i + n1 * k )] -= D[( n1 + k ) + 4 * ( n1 + j )];
      b[e] = D[i + 4 * ( n1 + j )];
    }
  for ( int q = 0; q < nn; ++q ) colperm[q] = q;
  for ( int p = 0; p < nn; ++p ) {
    // complete pivoting
    int pr = p, pc = p;
    double best = -1.0;
    for ( int c = p; c < nn; ++c )
      for ( int r = p; r < nn; ++r )
        if ( fabs( K[r + 4 * c] ) > best ) {
          best = fabs( K[r + 4 * c] );
          pr = r;
          pc = c;
        }
    if ( pr != p ) {
      for ( int c = 0; c < nn; ++c ) {
        const double tmp = K[p + 4 * c];
        K[p + 4 * c] = K[pr + 4 * c];
        K[pr + 4 * c] = tmp;
      }
      const double tmp = b[p];
      b[p] = b[pr];
      b[pr] = tmp;
    }
    if ( pc != p ) {
      for ( int r = 0; r < nn; ++r ) {
        const double tmp = K[r + 4 * p];
        K[r + 4 * p] = K[r + 4 * pc];
        K[r + 4 * pc] = tmp;
      }
      const int ti = colperm[p];
      colperm[p] = colperm[pc];
      colperm[pc] = ti;
    }
    if ( fabs( K[p + 4 * p] ) < smin ) K[p + 4 * p] = smin;  // perturbed pivot (the swap is verified afterwards)
    for ( int r = p + 1; r < nn; ++r ) {
      const double f = K[r + 4 * p] / K[p + 4 * p];
      K[r + 4 * p] = 0.0;
      for ( int c = p + 1; c < nn; ++c ) K[r + 4 * c] -= f * K[p + 4 * c];
      b[r] -= f * b[p];
    }
  }
  scale = 1.0;
  bool risky = false;
  double bmax = 0.0;
  for ( int p = 0; p < nn; ++p ) {
    risky = risky || ( ( 8.0 * smlnum ) * fabs( b[p] ) > fabs( K[p + 4 * p] ) );
    bmax = dmax( bmax, fabs( b[p] ) );
  }
  if ( risky && bmax > 0.0 ) {
    scale = 0.125 / bmax;
    for ( int p = 0; p < nn; ++p ) b[p] *= scale;
  }
  double y[4];
  for ( int p = nn - 1; p >= 0; --p ) {
    double acc = b[p];
    for ( int c = p + 1; c < nn; ++c ) acc -= K[p + 4 * c] * y[c];
    y[p] = acc / K[p + 4 * p];
  }
  for ( int p = 0; p < nn; ++p ) {
    const int u = colperm[p];  // unknown number = row + n1 * column
    X[( u % n1 ) + 2 * ( u / n1 )] = y[p];
  }
}

// 3-element reflector H = I - tau u u^T on a private 4x4 array (ld 4)
PB_HD void aed_d_left( double* D, int r0, int c0, int nc, const double* u, double tau ) {
  for ( int c = c0; c < c0 + nc; ++c ) {
    const double sum = tau * ( u[0] * D[r0 + 4 * c] + u[1] * D[r0 + 1 + 4 * c] + u[2] * D[r0 + 2 + 4 * c] );
    D[r0 + 4 * c] -= sum * u[0];
    D[r0 + 1 + 4 * c] -= sum * u[1];
    D[r0 + 2 + 4 * c] -= sum * u[2];
  }
}
PB_HD void aed_d_right( double* D, int c0, int r0, int nr, const double* u, double tau ) {
  for ( int r = r0; r < r0 + nr; ++r ) {
    const double sum = tau * ( u[0] * D[r + 4 * c0] + u[1] * D[r + 4 * ( c0 + 1 )] + u[2] * D[r + 4 * ( c0 + 2 )] );
    D[r + 4 * c0] -= sum * u[0];
    D[r + 4 * ( c0 + 1 )] -= sum * u[1];
    D[r + 4 * ( c0 + 2 )] -= sum * u[2];
  }
}

// the same on the shared arrays, work split over the group; no barrier inside
template <class G>
PB_D void aed_refl_rows( const G& g, double* T, int ldt, int r0, int c_lo, int c_hi, const double* u, double tau ) {
  for ( int c = c_lo + g.tid(); c <= c_hi; c += g.size() ) {
    const double sum = tau * ( u[0] * PB_AT( r0, c ) + u[1] * PB_AT( r0 + 1, c ) + u[2] * PB_AT( r0 + 2, c ) );
    PB_AT( r0, c ) -= sum * u[0];
    PB_AT( r0 + 1, c ) -= sum * u[1];
    PB_AT( r0 + 2, c ) -= sum * u[2];
  }
}
template <class G>
PB_D void aed_refl_cols( const G& g, double* T, int ldt, int c0, int r_lo, int r_hi, const double* u, double tau ) {
  for ( int r = r_lo + g.tid(); r <= r_hi; r += g.size() ) {
    const double sum = tau * ( u[0] * PB_AT( r, c0 ) + u[1] * PB_AT( r, c0 + 1 ) + u[2] * PB_AT( r, c0 + 2 ) );
    PB_AT( r, c0 ) -= sum * u[0];
    PB_AT( r, c0 + 1 ) -= sum * u[1];
    PB_AT( r, c0 + 2 ) -= sum * u[2];
  }
}
template <class G>
PB_D void aed_rot_rows( const G& g, double* T, int ldt, int r0, int r1, int c_lo, int c_hi, double cs, double sn ) {
  for ( int c = c_lo + g.tid(); c <= c_hi; c += g.size() ) {
    const double x = PB_AT( r0, c ), y = PB_AT( r1, c );
    PB_AT( r0, c ) = cs * x + sn * y;
    PB_AT( r1, c ) = cs * y - sn * x;
  }
}
template <class G>
PB_D void aed_rot_cols( const G& g, double* T, int ldt, int c0, int c1, int r_lo, int r_hi, double cs, double sn ) {
  for ( int r = r_lo + g.tid(); r <= r_hi; r += g.size() ) {
    const double x = PB_AT( r, c0 ), y = PB_AT( r, c1 );
    PB_AT( r, c0 ) = cs * x + sn * y;
    PB_AT( r, c1 ) = cs * y - sn * x;
  }
}

// standardise the 2x2 diagonal block at (j, j) of the quasi-triangular T (dlanv2 + the rotation on the rest)
template <class G>
PB_D void aed_standardise( const G& g, int n, double* T, int ldt, double* Q, int ldq, int nq, int j ) {
  double a = PB_AT( j, j ), b = PB_AT( j, j + 1 ), c = PB_AT( j + 1, j ), d = PB_AT( j + 1, j + 1 );
  double r1r, r1i, r2r, r2i, cs, sn;
  dlanv2( a, b, c, d, r1r, r1i, r2r, r2i, cs, sn );
  g.sync();
  if ( g.tid() == 0 ) {
    PB_AT( j, j ) = a;
    PB_AT( j, j + 1 ) = b;
    PB_AT( j + 1, j ) = c;
    PB_AT( j + 1, j + 1 ) = d;
  }
  if ( j + 2 <= n - 1 ) aed_rot_rows( g, T, ldt, j, j + 1, j + 2, n - 1, cs, sn );
  if ( j >= 1 ) aed_rot_cols( g, T, ldt, j, j + 1, 0, j - 1, cs, sn );
  aed_rot_cols( g, Q, ldq, j, j + 1, 0, nq - 1, cs, sn );
  g.sync();
}

// ---- dlaexc: swap the adjacent diagonal blocks (order n1 at j1, order n2 behind it) of the quasi-triangular T
// (order n), accumulating into the nq rows of Q. Returns false (T, Q untouched) when the swap would be too
// inaccurate. Every thread evaluates the small dense part redundantly. ------------------------------------
#ifndef PB_AED_COUNT
#define PB_AED_COUNT( what )
#endif
template <class G>
PB_D bool aed_laexc( const G& g, int n, double* T, int ldt, double* Q, int ldq, int nq, int j1, int n1, int n2 ) {
  PB_AED_COUNT( 0 );
  if ( n == 0 || n1 == 0 || n2 == 0 || j1 + n1 > n - 1 ) return true;
  const int j2 = j1 + 1, j3 = j1 + 2, j4 = j1 + 3;
  if ( n1 == 1 && n2 == 1 ) {
    const double t11 = PB_AT( j1, j1 ), t22 = PB_AT( j2, j2 );
    double cs, sn, tmp;
    dlartg( PB_AT( j1, j2 ), t22 - t11, cs, sn, tmp );
    g.sync();
    if ( j3 <= n - 1 ) aed_rot_rows( g, T, ldt, j1, j2, j3, n - 1, cs, sn );
    if ( j1 >= 1 ) aed_rot_cols( g, T, ldt, j1, j2, 0, j1 - 1, cs, sn );
    if ( g.tid() == 0 ) {
      PB_AT( j1, j1 ) = t22;
      PB_AT( j2, j2 ) = t11;
    }
    aed_rot_cols( g, Q, ldq, j1, j2, 0, nq - 1, cs, sn );
    g.sync();
    return true;
  }
  const int nd = n1 + n2;
  double D[16], X[4] = { 0, 0, 0, 0 };
  double dnorm = 0.0;
  for ( int c = 0; c < 4; ++c )
    for ( int r = 0; r < 4; ++r ) {
      const double v = ( r < nd && c < nd ) ? PB_AT( j1 + r, j1 + c ) : 0.0;
      D[r + 4 * c] = v;
      dnorm = dmax( dnorm, fabs( v ) );
    }
  const double smlnum = kSafmin / kUlp;
  const double thresh = dmax( 10.0 * kUlp * dnorm, smlnum );
  double scale;
  aed_sylvester( n1, n2, D, X, scale );
  double u1[3], u2[3], tau1 = 0.0, tau2 = 0.0;
  g.sync();  // every thread has read the diagonal blocks
  if ( n1 == 1 && n2 == 2 ) {
    // reflector H with (scale, X11, X12) H = (0, 0, *)
    double al = X[2], x0 = scale, x1 = X[0];
    dlarfg3( 3, al, x0, x1, tau1 );
    u1[0] = x0; u1[1] = x1; u1[2] = 1.0;
    const double t11 = D[0];
    aed_d_left( D, 0, 0, 3, u1, tau1 );
    aed_d_right( D, 0, 0, 3, u1, tau1 );
    if ( dmax( fabs( D[2] ), dmax( fabs( D[2 + 4] ), fabs( D[2 + 8] - t11 ) ) ) > thresh ) return false;
    aed_refl_rows( g, T, ldt, j1, j1, n - 1, u1, tau1 );
    g.sync();
    aed_refl_cols( g, T, ldt, j1, 0, j2, u1, tau1 );
    aed_refl_cols( g, Q, ldq, j1, 0, nq - 1, u1, tau1 );
    g.sync();
    if ( g.tid() == 0 ) {
      PB_AT( j3, j1 ) = 0.0;
      PB_AT( j3, j2 ) = 0.0;
      PB_AT( j3, j3 ) = t11;
    }
    g.sync();
  } else if ( n1 == 2 && n2 == 1 ) {
    // reflector H with H (-X11, -X21, scale)^T = (*, 0, 0)^T
    double al = -X[0], x0 = -X[1], x1 = scale;
    dlarfg3( 3, al, x0, x1, tau1 );
    u1[0] = 1.0; u1[1] = x0; u1[2] = x1;
    const double t33 = D[2 + 8];
    aed_d_left( D, 0, 0, 3, u1, tau1 );
    aed_d_right( D, 0, 0, 3, u1, tau1 );
    if ( dmax( fabs( D[1] ), dmax( fabs( D[2] ), fabs( D[0] - t33 ) ) ) > thresh ) return false;
    aed_refl_cols( g, T, ldt, j1, 0, j3, u1, tau1 );
    aed_refl_cols( g, Q, ldq, j1, 0, nq - 1, u1, tau1 );
    g.sync();
    aed_refl_rows( g, T, ldt, j1, j2, n - 1, u1, tau1 );
    g.sync();
    if ( g.tid() == 0 ) {
      PB_AT( j1, j1 ) = t33;
      PB_AT( j2, j1 ) = 0.0;
      PB_AT( j3, j1 ) = 0.0;
    }
    g.sync();
  } else {
    // two reflectors with H2 H1 (-X; scale I) = (*; 0)
    double al = -X[0], x0 = -X[1], x1 = scale;
    dlarfg3( 3, al, x0, x1, tau1 );
    u1[0] = 1.0; u1[1] = x0; u1[2] = x1;
    const double temp = -tau1 * ( X[2] + u1[1] * X[3] );
    double al2 = -temp * u1[1] - X[3], y0 = -temp * u1[2], y1 = scale;
    dlarfg3( 3, al2, y0, y1, tau2 );
    u2[0] = 1.0; u2[1] = y0; u2[2] = y1;
    aed_d_left( D, 0, 0, 4, u1, tau1 );
    aed_d_right( D, 0, 0, 4, u1, tau1 );
    aed_d_left( D, 1, 0, 4, u2, tau2 );
    aed_d_right( D, 1, 0, 4, u2, tau2 );
    if ( dmax( dmax( fabs( D[2] ), fabs( D[2 + 4] ) ), dmax( fabs( D[3] ), fabs( D[3 + 4] ) ) ) > thresh ) return false;
    aed_refl_rows( g, T, ldt, j1, j1, n - 1, u1, tau1 );
    g.sync();
    aed_refl_cols( g, T, ldt, j1, 0, j4, u1, tau1 );
    aed_refl_cols( g, Q, ldq, j1, 0, nq - 1, u1, tau1 );
    g.sync();
    aed_refl_rows( g, T, ldt, j2, j1, n - 1, u2, tau2 );
    g.sync();
    aed_refl_cols( g, T, ldt, j2, 0, j4, u2, tau2 );
    aed_refl_cols( g, Q, ldq, j2, 0, nq - 1, u2, tau2 );
    g.sync();
    if ( g.tid() == 0 ) {
      PB_AT( j3, j1 ) = 0.0;
      PB_AT( j3, j2 ) = 0.0;
      PB_AT( j4, j1 ) = 0.0;
      PB_AT( j4, j2 ) = 0.0;
    }
    g.sync();
  }
  if ( n2 == 2 ) aed_standardise( g, n, T, ldt, Q, ldq, nq, j1 );
  if ( n1 == 2 ) aed_standardise( g, n, T, ldt, Q, ldq, nq, j1 + n2 );
  return true;
}

// ---- dtrexc, upward direction only: the block whose first row is ifst is moved up until its first row is ilst
// (ilst <= ifst, both at block starts). Returns the row the block ended at (== ilst unless a swap was rejected).
template <class G>
PB_D int aed_trexc_up( const G& g, int n, double* T, int ldt, double* Q, int ldq, int nq, int ifst, int ilst ) {
  int nbf = 1;
  if ( ifst < n - 1 && PB_AT( ifst + 1, ifst ) != 0.0 ) nbf = 2;
  int here = ifst;
  while ( here > ilst ) {
    int nbnext = 1;
    if ( here >= 2 && PB_AT( here - 1, here - 2 ) != 0.0 ) nbnext = 2;
    if ( nbf == 1 || nbf == 2 ) {
      if ( !aed_laexc( g, n, T, ldt, Q, ldq, nq, here - nbnext, nbnext, nbf ) ) return here;
      here -= nbnext;
      if ( nbf == 2 && PB_AT( here + 1, here ) == 0.0 ) nbf = 3;  // the 2x2 block split into two 1x1 blocks
    } else {
      // two 1x1 blocks are moved one after the other
      if ( !aed_laexc( g, n, T, ldt, Q, ldq, nq, here - nbnext, nbnext, 1 ) ) return here;
      if ( nbnext == 1 ) {
        aed_laexc( g, n, T, ldt, Q, ldq, nq, here, nbnext, 1 );
        here -= 1;
      } else {
        if ( PB_AT( here, here - 1 ) == 0.0 ) nbnext = 1;  // the 2x2 block ahead split
        if ( nbnext == 2 ) {
          if ( !aed_laexc( g, n, T, ldt, Q, ldq, nq, here - 1, 2, 1 ) ) return here;
          here -= 2;
        } else {
          aed_laexc( g, n, T, ldt, Q, ldq, nq, here, 1, 1 );
          aed_laexc( g, n, T, ldt, Q, ldq, nq, here - 1, 1, 1 );
          here -= 2;
        }
      }
    }
  }
  return here;
}

// ---- deflation test of one Schur block [l..i] (order 1 or 2) at the bottom of the undeflated part ---------
PB_HD bool aed_block_deflatable( const double* T, int ldt, const double* V, int ldv, double s, double smlnum, int l, int i ) {
  if ( l == i ) {
    double foo = fabs( T[i + i * ldt] );
    if ( foo == 0.0 ) foo = fabs( s );
    return fabs( s * V[i * ldv] ) <= dmax( smlnum, kUlp * foo );
  }
  double foo = fabs( T[i + i * ldt] ) + sqrt( fabs( T[i + l * ldt] ) ) * sqrt( fabs( T[l + i * ldt] ) );
  if ( foo == 0.0 ) foo = fabs( s );
  return dmax( fabs( s * V[i * ldv] ), fabs( s * V[l * ldv] ) ) <= dmax( smlnum, kUlp * foo );
}

// stops the window's QR iteration at the first undeflatable block (no-reordering mode)
struct AedStopHook {
  static constexpr bool kActive = true;
  const double* T;
  const double* V;
  int ldt, ldv;
  double s, smlnum;
  int* ns;  // private to the calling thread
  PB_HD bool operator()( int l, int i ) const {
    if ( !aed_block_deflatable( T, ldt, V, ldv, s, smlnum, l, i ) ) return true;
    *ns = l;
    return false;
  }
};

// ---- the on-chip part of one AED step -------------------------------------------------------------------
// in : T (jw x jw upper Hessenberg copy of the trailing block), V = I, s = spike entry H(kwtop, kwtop-1)
// out: T Hessenberg in its leading ns x ns block and quasi-triangular behind it, V the accumulated orthogonal
//      factor, sr/si[ns..jw): the deflated eigenvalues; with max_moves > 0 also sr/si[0..ns): eigenvalues of the
//      undeflated part (shift candidates). out[0] = ns, out[1] = 0 or the lahqr failure code (then nothing may be
//      used), out[2] = 1 when T/V changed H, out[3] = 1 when sr/si[0..ns) are valid.
// max_moves: how many undeflatable blocks may be swapped to the top (LAPACK: unlimited); the test stops at the first
//      undeflatable block beyond that. On random matrices the converged eigenvalues sit at the bottom of the Schur
//      form anyway (measured at n = 512: 21 sweeps without any reordering against 20 with all 9525 block swaps).
//      With max_moves == 0 the window's QR iteration itself stops at the first undeflatable block: the leading part
//      is never brought to Schur form (and yields no shifts).
// work: >= 2 jw doubles visible to the group.
template <class G>
PB_D void aed_window( const G& g, int jw, int max_moves, double* T, int ldt, double* V, int ldv, double s, double* sr, double* si,
                      double* work, int* out, long long* t_qr_done = nullptr ) {
  const int t = g.tid(), nt = g.size();
  const double smlnum = kSafmin * ( (double)jw / kUlp );
  int ns = jw;    // T(0:ns, 0:ns) is the part not (yet) deflated
  // (the window's reflectors come from the fast 3-element generator, as in the bulge chase)
  int infqr;
  bool have_shifts = true;
  if ( max_moves == 0 ) {
    const AedStopHook hook{ T, V, ldt, ldv, s, smlnum, &ns };
    infqr = lahqr<G, true, AedStopHook>( g, true, true, jw, 0, jw - 1, T, ldt, sr, si, 0, jw - 1, V, ldv, hook );
    if ( infqr < 0 ) {
      infqr = 0;
      have_shifts = false;
    }
  } else {
    infqr = lahqr<G, true>( g, true, true, jw, 0, jw - 1, T, ldt, sr, si, 0, jw - 1, V, ldv );
  }
  g.sync();
#if defined( __CUDA_ARCH__ )
  if ( t_qr_done != nullptr ) *t_qr_done = clock64();  // (phase profile: what follows is the spike test / re-reduction)
#endif
  if ( infqr != 0 ) {
    if ( t == 0 ) {
      out[0] = 0;
      out[1] = infqr;
      out[2] = 0;
      out[3] = 0;
    }
    g.sync();
    return;
  }
  int ilst = 0;   // undeflatable blocks are collected in T(0:ilst, 0:ilst)
  int moves = 0;  // undeflatable blocks moved out of the way so far
  while ( max_moves > 0 && ilst < ns ) {
    const bool bulge = ( ns >= 2 ) && ( T[( ns - 1 ) + ( ns - 2 ) * ldt] != 0.0 );
    const int bl = bulge ? ns - 2 : ns - 1;
    if ( aed_block_deflatable( T, ldt, V, ldv, s, smlnum, bl, ns - 1 ) ) {
      ns = bl;
    } else {
      if ( moves++ >= max_moves ) break;
      const int at = aed_trexc_up( g, jw, T, ldt, V, ldv, jw, bl, ilst );
      ilst = at + ( bulge ? 2 : 1 );
    }
    PB_CHECK_UNIFORM( g, ns * 1000 + ilst, "aed deflation loop" );
  }
  if ( ns == 0 ) s = 0.0;
  // eigenvalues of the reordered Schur form
  g.sync();
  if ( t == 0 && have_shifts ) {
    int i = jw - 1;
    while ( i >= 0 ) {
      if ( i == 0 || T[i + ( i - 1 ) * ldt] == 0.0 ) {
        sr[i] = T[i + i * ldt];
        si[i] = 0.0;
        i -= 1;
      } else {
        double aa = T[( i - 1 ) + ( i - 1 ) * ldt], cc = T[i + ( i - 1 ) * ldt], bb = T[( i - 1 ) + i * ldt], dd = T[i + i * ldt];
        double cs, sn;
        dlanv2( aa, bb, cc, dd, sr[i - 1], si[i - 1], sr[i], si[i], cs, sn );
        i -= 2;
      }
    }
  }
  g.sync();
  int changed = 0;
  if ( ns < jw || s == 0.0 ) {
    changed = 1;
    if ( ns > 1 && s != 0.0 ) {
      // reflect the spike back: (I - tau w w^T) maps V(0, 0:ns)^T onto beta e_0
      double* w = work;
      double* tau = work + jw;
      for ( int c = t; c < ns; c += nt ) w[c] = V[c * ldv];
      g.sync();
      double beta, tw;
      larfg_col( g, w, jw, 0, ns - 1, 0, &beta, &tw );
      if ( t == 0 ) w[0] = 1.0;
      // nothing below the first subdiagonal survives
      for ( int q = t; q < jw * jw; q += nt ) {
        const int c = q / jw, r = q - c * jw;
        if ( r - c >= 2 ) T[r + c * ldt] = 0.0;
      }
      g.sync();
      for ( int c = t; c < jw; c += nt ) {
        double sum = 0.0;
        for ( int r = 0; r < ns; ++r ) sum += w[r] * T[r + c * ldt];
        sum *= tw;
        for ( int r = 0; r < ns; ++r ) T[r + c * ldt] -= sum * w[r];
      }
      g.sync();
      for ( int r = t; r < ns; r += nt ) {
        double sum = 0.0;
        for ( int c = 0; c < ns; ++c ) sum += T[r + c * ldt] * w[c];
        sum *= tw;
        for ( int c = 0; c < ns; ++c ) T[r + c * ldt] -= sum * w[c];
      }
      for ( int r = t; r < jw; r += nt ) {
        double sum = 0.0;
        for ( int c = 0; c < ns; ++c ) sum += V[r + c * ldv] * w[c];
        sum *= tw;
        for ( int c = 0; c < ns; ++c ) V[r + c * ldv] -= sum * w[c];
      }
      g.sync();
      gehd2( g, jw, 0, ns - 1, T, ldt, tau, w );
      g.sync();
      // V <- V Q, Q = H(0) H(1) ... H(ns-2) from the reflectors gehd2 left below the subdiagonal
      for ( int r = t; r < jw; r += nt ) {
        for ( int j = 0; j + 1 < ns; ++j ) {
          const double tj = tau[j];
          if ( tj == 0.0 ) continue;
          double sum = V[r + ( j + 1 ) * ldv];
          for ( int c = j + 2; c < ns; ++c ) sum += V[r + c * ldv] * T[c + j * ldt];
          sum *= tj;
          V[r + ( j + 1 ) * ldv] -= sum;
          for ( int c = j + 2; c < ns; ++c ) V[r + c * ldv] -= sum * T[c + j * ldt];
        }
      }
      g.sync();
    }
  }
  if ( t == 0 ) {
    out[0] = ns;
    out[1] = 0;
    out[2] = changed;
    out[3] = have_shifts ? 1 : 0;
  }
  g.sync();
}

#undef PB_AT
#undef PB_QQ

}  // namespace pb
