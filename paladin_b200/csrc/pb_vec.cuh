// pb_vec.cuh -- eigenvectors of a real quasi-triangular Schur factor and the dgeev epilogue:
//   dladiv / dlaln2      : scalar helpers (robust complex division, scaled 1x1 / 2x2 solves)
//   colnorms             : 1-norms of the strictly upper columns of T          (dtrevc3 prologue)
//   trevc_right_solve    : one right eigenvector (or conjugate pair) of T by back substitution
//   trevc_right_inplace  : all right eigenvectors, back-transformed in place by Z (dtrevc3 'R','B')
//   gebak_right          : undo balancing on the rows of V                      (dgebak 'B','R')
//   normalize_vectors    : unit 2-norm, complex pairs rotated so the largest component is real
//                          (tail of dgeev)
// These restate the published LAPACK 3.12 algorithms reached through the reference's only numerical
// call, dgeev_ (reference src/lapack-wrapper.hpp:47-60, src/paladin.hpp:248-254). The real-packed
// pair layout they produce is what the reference unpacks at src/paladin.hpp:256-333.
// 0-based indices; column-major; G is a warp or a CTA (pb_common.cuh).
#pragma once

#include "pb_common.cuh"

namespace pb {

#define PB_T( i, j ) T[( i ) + (size_t)( j ) * ldt]
#define PB_V( i, j ) V[( i ) + (size_t)( j ) * ldv]

// (a + ib) / (c + id) with Smith's scaling (published dladiv contract: no spurious overflow for
// representable quotients)
PB_HD void dladiv( double a, double b, double c, double d, double& p, double& q ) {
  if ( fabs( d ) < fabs( c ) ) {
    const double e = d / c;
    const double f = c + d * e;
    p = ( a + b * e ) / f;
    q = ( b - a * e ) / f;
  } else {
    const double e = c / d;
    const double f = d + c * e;
    p = ( b + a * e ) / f;
    q = ( -a + b * e ) / f;
  }
}

// dlaln2 with ca = 1, d1 = d2 = 1:  solves (A - (wr + i wi) I) X = scale * B, A is na x na (na = 1, 2),
// X/B have nw columns (nw = 1 real, nw = 2 = real and imaginary parts). ltrans solves with A^T.
// a = {a11, a21, a12, a22}; b, x = column-major 2 x 2 (element (r,c) at r + 2 c).
PB_HD int dlaln2( bool ltrans, int na, int nw, double smin, const double* a, const double* b, double wr,
                  double wi, double* x, double& scale, double& xnorm ) {
  const double smlnum = 2.0 * kSafmin;
  const double bignum = 1.0 / smlnum;
  const double smini = dmax( smin, smlnum );
  int info = 0;
  scale = 1.0;
  if ( na == 1 ) {
    if ( nw == 1 ) {
      double csr = a[0] - wr;
      double cnorm = fabs( csr );
      if ( cnorm < smini ) {
        csr = smini;
        cnorm = smini;
        info = 1;
      }
      const double bnorm = fabs( b[0] );
      if ( cnorm < 1.0 && bnorm > 1.0 ) {
        if ( bnorm > bignum * cnorm ) scale = 1.0 / bnorm;
      }
      x[0] = ( b[0] * scale ) / csr;
      xnorm = fabs( x[0] );
    } else {
      double csr = a[0] - wr;
      double csi = -wi;
      double cnorm = fabs( csr ) + fabs( csi );
      if ( cnorm < smini ) {
        csr = smini;
        csi = 0.0;
        cnorm = smini;
        info = 1;
      }
      const double bnorm = fabs( b[0] ) + fabs( b[2] );
      if ( cnorm < 1.0 && bnorm > 1.0 ) {
        if ( bnorm > bignum * cnorm ) scale = 1.0 / bnorm;
      }
      dladiv( scale * b[0], scale * b[2], csr, csi, x[0], x[2] );
      xnorm = fabs( x[0] ) + fabs( x[2] );
    }
    return info;
  }
  // 2 x 2: crv = {c11, c21, c12, c22}
  double crv[4];
  crv[0] = a[0] - wr;
  crv[3] = a[3] - wr;
  if ( ltrans ) {
    crv[2] = a[1];
    crv[1] = a[2];
  } else {
    crv[1] = a[1];
    crv[2] = a[2];
  }
  // pivot tables (complete pivoting): element order after bringing entry icmax to (1,1)
  const int ipivot[4][4] = { { 0, 1, 2, 3 }, { 1, 0, 3, 2 }, { 2, 3, 0, 1 }, { 3, 2, 1, 0 } };
  const bool zswap[4] = { false, false, true, true };
  const bool rswap[4] = { false, true, false, true };
  if ( nw == 1 ) {
    double cmax = 0.0;
    int icmax = -1;
    for ( int j = 0; j < 4; ++j )
      if ( fabs( crv[j] ) > cmax ) {
        cmax = fabs( crv[j] );
        icmax = j;
      }
    if ( cmax < smini ) {
      const double bnorm = dmax( fabs( b[0] ), fabs( b[1] ) );
      if ( smini < 1.0 && bnorm > 1.0 ) {
        if ( bnorm > bignum * smini ) scale = 1.0 / bnorm;
      }
      const double temp = scale / smini;
      x[0] = temp * b[0];
      x[1] = temp * b[1];
      xnorm = temp * bnorm;
      return 1;
    }
    const double ur11 = crv[icmax];
    const double cr21 = crv[ipivot[icmax][1]];
    const double ur12 = crv[ipivot[icmax][2]];
    const double cr22 = crv[ipivot[icmax][3]];
    const double ur11r = 1.0 / ur11;
    const double lr21 = ur11r * cr21;
    double ur22 = cr22 - ur12 * lr21;
    if ( fabs( ur22 ) < smini ) {
      ur22 = smini;
      info = 1;
    }
    double br1, br2;
    if ( rswap[icmax] ) {
      br1 = b[1];
      br2 = b[0];
    } else {
      br1 = b[0];
      br2 = b[1];
    }
    br2 = br2 - lr21 * br1;
    const double bbnd = dmax( fabs( br1 * ( ur22 * ur11r ) ), fabs( br2 ) );
    if ( bbnd > 1.0 && fabs( ur22 ) < 1.0 ) {
      if ( bbnd >= bignum * fabs( ur22 ) ) scale = 1.0 / bbnd;
    }
    const double xr2 = ( br2 * scale ) / ur22;
    const double xr1 = ( scale * br1 ) * ur11r - xr2 * ( ur11r * ur12 );
    if ( zswap[icmax] ) {
      x[0] = xr2;
      x[1] = xr1;
    } else {
      x[0] = xr1;
      x[1] = xr2;
    }
    xnorm = dmax( fabs( xr1 ), fabs( xr2 ) );
    if ( xnorm > 1.0 && cmax > 1.0 ) {
      if ( xnorm > bignum / cmax ) {
        const double temp = cmax / bignum;
        x[0] *= temp;
        x[1] *= temp;
        xnorm *= temp;
        scale *= temp;
      }
    }
    return info;
  }
  // complex 2 x 2
  double civ[4] = { -wi, 0.0, 0.0, -wi };
  double cmax = 0.0;
  int icmax = -1;
  for ( int j = 0; j < 4; ++j )
    if ( fabs( crv[j] ) + fabs( civ[j] ) > cmax ) {
      cmax = fabs( crv[j] ) + fabs( civ[j] );
      icmax = j;
    }
  if ( cmax < smini ) {
    const double bnorm = dmax( fabs( b[0] ) + fabs( b[2] ), fabs( b[1] ) + fabs( b[3] ) );
    if ( smini < 1.0 && bnorm > 1.0 ) {
      if ( bnorm > bignum * smini ) scale = 1.0 / bnorm;
    }
    const double temp = scale / smini;
    x[0] = temp * b[0];
    x[1] = temp * b[1];
    x[2] = temp * b[2];
    x[3] = temp * b[3];
    xnorm = temp * bnorm;
    return 1;
  }
  const double ur11 = crv[icmax], ui11 = civ[icmax];
  const double cr21 = crv[ipivot[icmax][1]], ci21 = civ[ipivot[icmax][1]];
  const double ur12 = crv[ipivot[icmax][2]], ui12 = civ[ipivot[icmax][2]];
  const double cr22 = crv[ipivot[icmax][3]], ci22 = civ[ipivot[icmax][3]];
  double ur11r, ui11r, lr21, li21, ur12s, ui12s, ur22, ui22;
  if ( icmax == 0 || icmax == 3 ) {
    // off-diagonals of pivoted C are real
    if ( fabs( ur11 ) > fabs( ui11 ) ) {
      const double temp = ui11 / ur11;
      ur11r = 1.0 / ( ur11 * ( 1.0 + temp * temp ) );
      ui11r = -temp * ur11r;
    } else {
      const double temp = ur11 / ui11;
      ui11r = -1.0 / ( ui11 * ( 1.0 + temp * temp ) );
      ur11r = -temp * ui11r;
    }
    lr21 = cr21 * ur11r;
    li21 = cr21 * ui11r;
    ur12s = ur12 * ur11r;
    ui12s = ur12 * ui11r;
    ur22 = cr22 - ur12 * lr21;
    ui22 = ci22 - ur12 * li21;
  } else {
    // diagonals of pivoted C are real
    ur11r = 1.0 / ur11;
    ui11r = 0.0;
    lr21 = cr21 * ur11r;
    li21 = ci21 * ur11r;
    ur12s = ur12 * ur11r;
    ui12s = ui12 * ur11r;
    ur22 = cr22 - ur12 * lr21 + ui12 * li21;
    ui22 = -ur12 * li21 - ui12 * lr21;
  }
  double u22abs = fabs( ur22 ) + fabs( ui22 );
  if ( u22abs < smini ) {
    ur22 = smini;
    ui22 = 0.0;
    info = 1;
  }
  double br1, br2, bi1, bi2;
  if ( rswap[icmax] ) {
    br2 = b[0];
    br1 = b[1];
    bi2 = b[2];
    bi1 = b[3];
  } else {
    br1 = b[0];
    br2 = b[1];
    bi1 = b[2];
    bi2 = b[3];
  }
  br2 = br2 - lr21 * br1 + li21 * bi1;
  bi2 = bi2 - li21 * br1 - lr21 * bi1;
  const double bbnd =
      dmax( ( fabs( br1 ) + fabs( bi1 ) ) * ( u22abs * ( fabs( ur11r ) + fabs( ui11r ) ) ), fabs( br2 ) + fabs( bi2 ) );
  if ( bbnd > 1.0 && u22abs < 1.0 ) {
    if ( bbnd >= bignum * u22abs ) {
      scale = 1.0 / bbnd;
      br1 *= scale;
      bi1 *= scale;
      br2 *= scale;
      bi2 *= scale;
    }
  }
  double xr2, xi2;
  dladiv( br2, bi2, ur22, ui22, xr2, xi2 );
  const double xr1 = ur11r * br1 - ui11r * bi1 - ur12s * xr2 + ui12s * xi2;
  const double xi1 = ui11r * br1 + ur11r * bi1 - ui12s * xr2 - ur12s * xi2;
  if ( zswap[icmax] ) {
    x[0] = xr2;
    x[1] = xr1;
    x[2] = xi2;
    x[3] = xi1;
  } else {
    x[0] = xr1;
    x[1] = xr2;
    x[2] = xi1;
    x[3] = xi2;
  }
  xnorm = dmax( fabs( xr1 ) + fabs( xi1 ), fabs( xr2 ) + fabs( xi2 ) );
  if ( xnorm > 1.0 && cmax > 1.0 ) {
    if ( xnorm > bignum / cmax ) {
      const double temp = cmax / bignum;
      x[0] *= temp;
      x[1] *= temp;
      x[2] *= temp;
      x[3] *= temp;
      xnorm *= temp;
      scale *= temp;
    }
  }
  return info;
}

// cnorm[j] = sum_{i<j} |T(i,j)|
template <class G>
PB_D void colnorms( const G& g, int n, const double* T, int ldt, double* cnorm ) {
  const int t = g.tid(), nt = g.size();
  for ( int j = t; j < n; j += nt ) {
    double s = 0.0;
    for ( int i = 0; i < j; ++i ) s += fabs( PB_T( i, j ) );
    cnorm[j] = s;
  }
  g.sync();
}

// Block structure of eigenvalue ki of quasi-triangular T: 0 real, -1 second row of a 2x2 block
// (pair is ki-1, ki), +1 first row of a 2x2 block.
PB_HD int trevc_block_kind( int n, const double* T, int ldt, int ki ) {
  if ( ki > 0 && PB_T( ki, ki - 1 ) != 0.0 ) return -1;
  if ( ki < n - 1 && PB_T( ki + 1, ki ) != 0.0 ) return 1;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// One right eigenvector of T by back substitution (the solve phase of dtrevc3, right side).
// ki must be a real eigenvalue (kind 0) or the SECOND index of a pair (kind -1).
// Real:    xr[0..ki] receives x (xr[ki] = 1 before scaling).
// Complex: xr[0..ki], xi[0..ki] receive real and imaginary parts for the eigenvalue wr + i wi
//          (wi > 0), i.e. packed columns ki-1 and ki of the LAPACK layout.
// Every thread of the group must call; xr/xi/cnorm are visible to the whole group.
// ------------------------------------------------------------------------------------------------
template <class G>
PB_D void trevc_right_solve( const G& g, int n, const double* T, int ldt, int ki, bool cplx, const double* cnorm,
                             double* xr, double* xi ) {
  const int t = g.tid(), nt = g.size();
  const double smlnum = kSafmin * ( (double)n / kUlp );
  const double bignum = ( 1.0 - kUlp ) / smlnum;
  const double wr = PB_T( ki, ki );
  double wi = 0.0;
  if ( cplx ) wi = sqrt( fabs( PB_T( ki, ki - 1 ) ) ) * sqrt( fabs( PB_T( ki - 1, ki ) ) );
  const double smin = dmax( kUlp * ( fabs( wr ) + fabs( wi ) ), smlnum );

  if ( !cplx ) {
    for ( int k = t; k < ki; k += nt ) xr[k] = -PB_T( k, ki );
    if ( t == 0 ) xr[ki] = 1.0;
    g.sync();
    int j = ki - 1;
    while ( j >= 0 ) {
      const bool two = ( j > 0 && PB_T( j, j - 1 ) != 0.0 );
      if ( !two ) {
        double x[4], scale, xnorm;
        const double a[4] = { PB_T( j, j ), 0.0, 0.0, 0.0 };
        const double b[4] = { xr[j], 0.0, 0.0, 0.0 };
        dlaln2( false, 1, 1, smin, a, b, wr, 0.0, x, scale, xnorm );
        if ( xnorm > 1.0 ) {
          if ( cnorm[j] > bignum / xnorm ) {
            x[0] /= xnorm;
            scale /= xnorm;
          }
        }
        g.sync();  // everyone has read xr[j]
        if ( scale != 1.0 )
          for ( int k = t; k <= ki; k += nt ) xr[k] *= scale;
        g.sync();
        if ( t == 0 ) xr[j] = x[0];
        for ( int k = t; k < j; k += nt ) xr[k] -= x[0] * PB_T( k, j );
        g.sync();
        j -= 1;
      } else {
        double x[4], scale, xnorm;
        const double a[4] = { PB_T( j - 1, j - 1 ), PB_T( j, j - 1 ), PB_T( j - 1, j ), PB_T( j, j ) };
        const double b[4] = { xr[j - 1], xr[j], 0.0, 0.0 };
        dlaln2( false, 2, 1, smin, a, b, wr, 0.0, x, scale, xnorm );
        if ( xnorm > 1.0 ) {
          const double beta = dmax( cnorm[j - 1], cnorm[j] );
          if ( beta > bignum / xnorm ) {
            x[0] /= xnorm;
            x[1] /= xnorm;
            scale /= xnorm;
          }
        }
        g.sync();
        if ( scale != 1.0 )
          for ( int k = t; k <= ki; k += nt ) xr[k] *= scale;
        g.sync();
        if ( t == 0 ) {
          xr[j - 1] = x[0];
          xr[j] = x[1];
        }
        for ( int k = t; k < j - 1; k += nt ) xr[k] -= x[0] * PB_T( k, j - 1 ) + x[1] * PB_T( k, j );
        g.sync();
        j -= 2;
      }
    }
    return;
  }

  // complex pair (ki-1, ki)
  {
    double a1, b2;  // xr[ki-1], xi[ki]
    if ( fabs( PB_T( ki - 1, ki ) ) >= fabs( PB_T( ki, ki - 1 ) ) ) {
      a1 = 1.0;
      b2 = wi / PB_T( ki - 1, ki );
    } else {
      a1 = -wi / PB_T( ki, ki - 1 );
      b2 = 1.0;
    }
    for ( int k = t; k < ki - 1; k += nt ) {
      xr[k] = -a1 * PB_T( k, ki - 1 );
      xi[k] = -b2 * PB_T( k, ki );
    }
    if ( t == 0 ) {
      xr[ki - 1] = a1;
      xi[ki] = b2;
      xr[ki] = 0.0;
      xi[ki - 1] = 0.0;
    }
    g.sync();
  }
  int j = ki - 2;
  while ( j >= 0 ) {
    const bool two = ( j > 0 && PB_T( j, j - 1 ) != 0.0 );
    if ( !two ) {
      double x[4], scale, xnorm;
      const double a[4] = { PB_T( j, j ), 0.0, 0.0, 0.0 };
      const double b[4] = { xr[j], 0.0, xi[j], 0.0 };
      dlaln2( false, 1, 2, smin, a, b, wr, wi, x, scale, xnorm );
      if ( xnorm > 1.0 ) {
        if ( cnorm[j] > bignum / xnorm ) {
          x[0] /= xnorm;
          x[2] /= xnorm;
          scale /= xnorm;
        }
      }
      g.sync();
      if ( scale != 1.0 )
        for ( int k = t; k <= ki; k += nt ) {
          xr[k] *= scale;
          xi[k] *= scale;
        }
      g.sync();
      if ( t == 0 ) {
        xr[j] = x[0];
        xi[j] = x[2];
      }
      for ( int k = t; k < j; k += nt ) {
        const double tk = PB_T( k, j );
        xr[k] -= x[0] * tk;
        xi[k] -= x[2] * tk;
      }
      g.sync();
      j -= 1;
    } else {
      double x[4], scale, xnorm;
      const double a[4] = { PB_T( j - 1, j - 1 ), PB_T( j, j - 1 ), PB_T( j - 1, j ), PB_T( j, j ) };
      const double b[4] = { xr[j - 1], xr[j], xi[j - 1], xi[j] };
      dlaln2( false, 2, 2, smin, a, b, wr, wi, x, scale, xnorm );
      if ( xnorm > 1.0 ) {
        const double beta = dmax( cnorm[j - 1], cnorm[j] );
        if ( beta > bignum / xnorm ) {
          const double rec = 1.0 / xnorm;
          x[0] *= rec;
          x[1] *= rec;
          x[2] *= rec;
          x[3] *= rec;
          scale *= rec;
        }
      }
      g.sync();
      if ( scale != 1.0 )
        for ( int k = t; k <= ki; k += nt ) {
          xr[k] *= scale;
          xi[k] *= scale;
        }
      g.sync();
      if ( t == 0 ) {
        xr[j - 1] = x[0];
        xr[j] = x[1];
        xi[j - 1] = x[2];
        xi[j] = x[3];
      }
      for ( int k = t; k < j - 1; k += nt ) {
        const double t1 = PB_T( k, j - 1 ), t2 = PB_T( k, j );
        xr[k] -= x[0] * t1 + x[1] * t2;
        xi[k] -= x[2] * t1 + x[3] * t2;
      }
      g.sync();
      j -= 2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// All right eigenvectors of T, back-transformed in place: on entry V = Z (Schur vectors), on exit
// V(:,k) holds the real-packed eigenvectors of Z T Z^T (dtrevc3 'R','B', one vector at a time).
// xr, xi: n doubles each; cnorm: n doubles. Each vector is scaled so that its largest component
// (|re|+|im| for pairs) is 1, as LAPACK does.
// ------------------------------------------------------------------------------------------------
template <class G>
PB_D void trevc_right_inplace( const G& g, int n, const double* T, int ldt, double* V, int ldv, double* cnorm,
                               double* xr, double* xi ) {
  const int t = g.tid(), nt = g.size();
  colnorms( g, n, T, ldt, cnorm );
  int ki = n - 1;
  while ( ki >= 0 ) {
    const bool cplx = ( ki > 0 && PB_T( ki, ki - 1 ) != 0.0 );
    trevc_right_solve( g, n, T, ldt, ki, cplx, cnorm, xr, xi );
    if ( !cplx ) {
      // V(:,ki) = V(:,0:ki-1) x(0:ki-1) + x(ki) V(:,ki)
      double emax = 0.0;
      for ( int r = t; r < n; r += nt ) {
        double s = xr[ki] * PB_V( r, ki );
        for ( int c = 0; c < ki; ++c ) s += PB_V( r, c ) * xr[c];
        PB_V( r, ki ) = s;
        emax = dmax( emax, fabs( s ) );
      }
      emax = g.max( emax );
      const double remax = 1.0 / emax;
      for ( int r = t; r < n; r += nt ) PB_V( r, ki ) *= remax;
      g.sync();
      ki -= 1;
    } else {
      double emax = 0.0;
      for ( int r = t; r < n; r += nt ) {
        double sr = xr[ki - 1] * PB_V( r, ki - 1 );
        double si = xi[ki] * PB_V( r, ki );
        for ( int c = 0; c < ki - 1; ++c ) {
          const double v = PB_V( r, c );
          sr += v * xr[c];
          si += v * xi[c];
        }
        PB_V( r, ki - 1 ) = sr;
        PB_V( r, ki ) = si;
        emax = dmax( emax, fabs( sr ) + fabs( si ) );
      }
      emax = g.max( emax );
      const double remax = 1.0 / emax;
      for ( int r = t; r < n; r += nt ) {
        PB_V( r, ki - 1 ) *= remax;
        PB_V( r, ki ) *= remax;
      }
      g.sync();
      ki -= 2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One LEFT eigenvector of T (the solve phase of dtrevc3, left side): forward substitution with the
// transposed blocks, dot-product form (column segments of T, contiguous).
// ki must be a real eigenvalue or the FIRST index of a pair (pair = ki, ki+1).
// Real:    xr[ki..n-1] receives y (xr[ki] = 1 before scaling).
// Complex: xr[ki..n-1], xi[ki..n-1] = real and imaginary parts (packed columns ki and ki+1).
// ------------------------------------------------------------------------------------------------
template <class G>
PB_D void trevc_left_solve( const G& g, int n, const double* T, int ldt, int ki, bool cplx, const double* cnorm, double* xr,
                            double* xi ) {
  const int t = g.tid(), nt = g.size();
  const double smlnum = kSafmin * ( (double)n / kUlp );
  const double bignum = ( 1.0 - kUlp ) / smlnum;
  const double wr = PB_T( ki, ki );
  double wi = 0.0;
  if ( cplx ) wi = sqrt( fabs( PB_T( ki, ki + 1 ) ) ) * sqrt( fabs( PB_T( ki + 1, ki ) ) );
  const double smin = dmax( kUlp * ( fabs( wr ) + fabs( wi ) ), smlnum );
  double vmax = 1.0, vcrit = bignum;

  if ( !cplx ) {
    for ( int k = ki + 1 + t; k < n; k += nt ) xr[k] = -PB_T( ki, k );
    if ( t == 0 ) xr[ki] = 1.0;
    g.sync();
    int j = ki + 1;
    while ( j < n ) {
      const bool two = ( j < n - 1 && PB_T( j + 1, j ) != 0.0 );
      const double beta = two ? dmax( cnorm[j], cnorm[j + 1] ) : cnorm[j];
      if ( beta > vcrit ) {
        const double rec = 1.0 / vmax;
        for ( int k = ki + t; k < n; k += nt ) xr[k] *= rec;
        vmax = 1.0;
        vcrit = bignum;
        g.sync();
      }
      double d0 = 0.0, d1 = 0.0;
      for ( int k = ki + 1 + t; k < j; k += nt ) {
        const double xk = xr[k];
        d0 += PB_T( k, j ) * xk;
        if ( two ) d1 += PB_T( k, j + 1 ) * xk;
      }
      d0 = g.sum( d0 );
      if ( two ) d1 = g.sum( d1 );
      double x[4], scale, xnorm;
      if ( !two ) {
        const double a[4] = { PB_T( j, j ), 0.0, 0.0, 0.0 };
        const double b[4] = { xr[j] - d0, 0.0, 0.0, 0.0 };
        dlaln2( false, 1, 1, smin, a, b, wr, 0.0, x, scale, xnorm );
      } else {
        const double a[4] = { PB_T( j, j ), PB_T( j + 1, j ), PB_T( j, j + 1 ), PB_T( j + 1, j + 1 ) };
        const double b[4] = { xr[j] - d0, xr[j + 1] - d1, 0.0, 0.0 };
        dlaln2( true, 2, 1, smin, a, b, wr, 0.0, x, scale, xnorm );
      }
      g.sync();
      if ( scale != 1.0 )
        for ( int k = ki + t; k < n; k += nt ) xr[k] *= scale;
      g.sync();
      if ( t == 0 ) {
        xr[j] = x[0];
        if ( two ) xr[j + 1] = x[1];
      }
      vmax = dmax( fabs( x[0] ), vmax );
      if ( two ) vmax = dmax( fabs( x[1] ), vmax );
      vcrit = bignum / vmax;
      g.sync();
      j += two ? 2 : 1;
    }
    return;
  }

  // complex pair (ki, ki+1), eigenvalue wr - i wi of T^T
  {
    double a1, b2;  // xr[ki], xi[ki+1]
    if ( fabs( PB_T( ki, ki + 1 ) ) >= fabs( PB_T( ki + 1, ki ) ) ) {
      a1 = wi / PB_T( ki, ki + 1 );
      b2 = 1.0;
    } else {
      a1 = 1.0;
      b2 = -wi / PB_T( ki + 1, ki );
    }
    for ( int k = ki + 2 + t; k < n; k += nt ) {
      xr[k] = -a1 * PB_T( ki, k );
      xi[k] = -b2 * PB_T( ki + 1, k );
    }
    if ( t == 0 ) {
      xr[ki] = a1;
      xi[ki + 1] = b2;
      xr[ki + 1] = 0.0;
      xi[ki] = 0.0;
    }
    g.sync();
  }
  int j = ki + 2;
  while ( j < n ) {
    const bool two = ( j < n - 1 && PB_T( j + 1, j ) != 0.0 );
    const double beta = two ? dmax( cnorm[j], cnorm[j + 1] ) : cnorm[j];
    if ( beta > vcrit ) {
      const double rec = 1.0 / vmax;
      for ( int k = ki + t; k < n; k += nt ) {
        xr[k] *= rec;
        xi[k] *= rec;
      }
      vmax = 1.0;
      vcrit = bignum;
      g.sync();
    }
    double dr0 = 0.0, di0 = 0.0, dr1 = 0.0, di1 = 0.0;
    for ( int k = ki + 2 + t; k < j; k += nt ) {
      const double t0 = PB_T( k, j ), xrk = xr[k], xik = xi[k];
      dr0 += t0 * xrk;
      di0 += t0 * xik;
      if ( two ) {
        const double t1 = PB_T( k, j + 1 );
        dr1 += t1 * xrk;
        di1 += t1 * xik;
      }
    }
    dr0 = g.sum( dr0 );
    di0 = g.sum( di0 );
    if ( two ) {
      dr1 = g.sum( dr1 );
      di1 = g.sum( di1 );
    }
    double x[4], scale, xnorm;
    if ( !two ) {
      const double a[4] = { PB_T( j, j ), 0.0, 0.0, 0.0 };
      const double b[4] = { xr[j] - dr0, 0.0, xi[j] - di0, 0.0 };
      dlaln2( false, 1, 2, smin, a, b, wr, -wi, x, scale, xnorm );
    } else {
      const double a[4] = { PB_T( j, j ), PB_T( j + 1, j ), PB_T( j, j + 1 ), PB_T( j + 1, j + 1 ) };
      const double b[4] = { xr[j] - dr0, xr[j + 1] - dr1, xi[j] - di0, xi[j + 1] - di1 };
      dlaln2( true, 2, 2, smin, a, b, wr, -wi, x, scale, xnorm );
    }
    g.sync();
    if ( scale != 1.0 )
      for ( int k = ki + t; k < n; k += nt ) {
        xr[k] *= scale;
        xi[k] *= scale;
      }
    g.sync();
    if ( t == 0 ) {
      xr[j] = x[0];
      xi[j] = x[2];
      if ( two ) {
        xr[j + 1] = x[1];
        xi[j + 1] = x[3];
      }
    }
    vmax = dmax( dmax( fabs( x[0] ), fabs( x[2] ) ), vmax );
    if ( two ) vmax = dmax( dmax( fabs( x[1] ), fabs( x[3] ) ), vmax );
    vcrit = bignum / vmax;
    g.sync();
    j += two ? 2 : 1;
  }
}

// All left eigenvectors, back-transformed in place: on entry V = Z, on exit V(:,k) = real-packed left eigenvectors
// of Z T Z^T (dtrevc3 'L','B', one vector at a time; ascending order keeps columns > ki untouched until used).
template <class G>
PB_D void trevc_left_inplace( const G& g, int n, const double* T, int ldt, double* V, int ldv, double* cnorm, double* xr,
                              double* xi ) {
  const int t = g.tid(), nt = g.size();
  colnorms( g, n, T, ldt, cnorm );
  int ki = 0;
  while ( ki < n ) {
    const bool cplx = ( ki < n - 1 && PB_T( ki + 1, ki ) != 0.0 );
    trevc_left_solve( g, n, T, ldt, ki, cplx, cnorm, xr, xi );
    if ( !cplx ) {
      double emax = 0.0;
      for ( int r = t; r < n; r += nt ) {
        double s = xr[ki] * PB_V( r, ki );
        for ( int c = ki + 1; c < n; ++c ) s += PB_V( r, c ) * xr[c];
        PB_V( r, ki ) = s;
        emax = dmax( emax, fabs( s ) );
      }
      emax = g.max( emax );
      const double remax = 1.0 / emax;
      for ( int r = t; r < n; r += nt ) PB_V( r, ki ) *= remax;
      g.sync();
      ki += 1;
    } else {
      double emax = 0.0;
      for ( int r = t; r < n; r += nt ) {
        double sr = xr[ki] * PB_V( r, ki );
        double si = xi[ki + 1] * PB_V( r, ki + 1 );
        for ( int c = ki + 2; c < n; ++c ) {
          const double v = PB_V( r, c );
          sr += v * xr[c];
          si += v * xi[c];
        }
        PB_V( r, ki ) = sr;
        PB_V( r, ki + 1 ) = si;
        emax = dmax( emax, fabs( sr ) + fabs( si ) );
      }
      emax = g.max( emax );
      const double remax = 1.0 / emax;
      for ( int r = t; r < n; r += nt ) {
        PB_V( r, ki ) *= remax;
        PB_V( r, ki + 1 ) *= remax;
      }
      g.sync();
      ki += 2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// gebak 'B','R' on the rows of V (n x m). scale[] as produced by gebal (0-based permutation indices
// outside ilo..ihi, scaling factors inside).
// ------------------------------------------------------------------------------------------------
// left = true: dgebak 'B','L' (rows are divided by the scaling factors instead of multiplied)
template <class G>
PB_D void gebak_vectors( const G& g, bool left, int n, int ilo, int ihi, const double* scale, int m, double* V, int ldv ) {
  const int t = g.tid(), nt = g.size();
  if ( n == 0 || m == 0 ) return;
  if ( ilo != ihi ) {
    for ( int c = 0; c < m; ++c )
      for ( int r = ilo + t; r <= ihi; r += nt ) PB_V( r, c ) *= left ? 1.0 / scale[r] : scale[r];
  }
  g.sync();
  for ( int ii = 0; ii < n; ++ii ) {
    int i = ii;
    if ( i >= ilo && i <= ihi ) continue;
    if ( i < ilo ) i = ilo - 1 - ii;
    const int k = (int)scale[i];
    if ( k == i ) continue;
    for ( int c = t; c < m; c += nt ) {
      const double tmp = PB_V( i, c );
      PB_V( i, c ) = PB_V( k, c );
      PB_V( k, c ) = tmp;
    }
    g.sync();
  }
}

template <class G>
PB_D void gebak_right( const G& g, int n, int ilo, int ihi, const double* scale, int m, double* V, int ldv ) {
  gebak_vectors( g, false, n, ilo, ihi, scale, m, V, ldv );
}

// overflow-safe 2-norm of a column by the group
template <class G>
PB_D double col_nrm2( const G& g, int n, const double* x ) {
  const int t = g.tid(), nt = g.size();
  double mx = 0.0;
  for ( int r = t; r < n; r += nt ) mx = dmax( mx, fabs( x[r] ) );
  mx = g.max( mx );
  if ( mx == 0.0 || !( mx <= DBL_MAX ) ) return mx;
  double ss = 0.0;
  for ( int r = t; r < n; r += nt ) {
    const double q = x[r] / mx;
    ss += q * q;
  }
  return mx * sqrt( g.sum( ss ) );
}

// ------------------------------------------------------------------------------------------------
// Normalisation done by dgeev after dgebak: real vectors to unit 2-norm; pairs (re, im) to unit
// joint 2-norm, then rotated so that the component of largest modulus is real (imag part set to 0).
// ------------------------------------------------------------------------------------------------
template <class G>
PB_D void normalize_vectors( const G& g, int n, const double* wi, double* V, int ldv ) {
  const int t = g.tid(), nt = g.size();
  for ( int i = 0; i < n; ++i ) {
    if ( wi[i] == 0.0 ) {
      const double nrm = col_nrm2( g, n, &PB_V( 0, i ) );
      const double scl = 1.0 / nrm;
      for ( int r = t; r < n; r += nt ) PB_V( r, i ) *= scl;
      g.sync();
    } else if ( wi[i] > 0.0 && i + 1 < n ) {
      const double n1 = col_nrm2( g, n, &PB_V( 0, i ) );
      const double n2 = col_nrm2( g, n, &PB_V( 0, i + 1 ) );
      const double scl = 1.0 / dlapy2( n1, n2 );
      double best = -1.0;
      int bestk = n;
      for ( int r = t; r < n; r += nt ) {
        const double a = PB_V( r, i ) * scl, b = PB_V( r, i + 1 ) * scl;
        PB_V( r, i ) = a;
        PB_V( r, i + 1 ) = b;
        const double w = a * a + b * b;
        if ( w > best ) {  // strictly greater: keeps the first maximum of this thread's rows
          best = w;
          bestk = r;
        }
      }
      const double gbest = g.max( best );
      const int k = g.imin_( best == gbest ? bestk : n );  // first index attaining the maximum
      g.sync();
      if ( k < n ) {
        double cs, sn, r;
        dlartg( PB_V( k, i ), PB_V( k, i + 1 ), cs, sn, r );
        g.sync();
        for ( int rr = t; rr < n; rr += nt ) {
          const double x = PB_V( rr, i ), y = PB_V( rr, i + 1 );
          PB_V( rr, i ) = cs * x + sn * y;
          PB_V( rr, i + 1 ) = cs * y - sn * x;
        }
        g.sync();
        if ( t == 0 ) PB_V( k, i + 1 ) = 0.0;
        g.sync();
      }
    }
  }
}

}  // namespace pb
