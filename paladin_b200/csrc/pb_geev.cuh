// pb_geev.cuh -- the whole dgeev pipeline executed by ONE thread group (a warp for n <= 64, a CTA as
// the size-generic path) on one matrix held in memory the group can address (shared or global).
//
// Restates the driver logic of LAPACK 3.12 dgeev (third party to the reference; reached through
// reference src/lapack-wrapper.hpp:47-60 from src/paladin.hpp:248-254):
//   scale check (dlange 'M' / dlascl) -> gebal 'B' -> Hessenberg -> [orghr] -> QR iteration ->
//   [trevc 'R','B' -> gebak 'B','R' -> normalise] -> undo scaling of the eigenvalues.
// Output conventions consumed by the reference unpack loop (src/paladin.hpp:256-333): conjugate
// pairs consecutive, positive imaginary part first, wi[k+1] == -wi[k] exactly, vectors real-packed.
#pragma once

#include "pb_common.cuh"
#include "pb_small.cuh"
#include "pb_vec.cuh"
#include "pb_msqr.cuh"

namespace pb {

// x *= cto/cfrom without intermediate over/underflow (published dlascl multiplication schedule)
template <class G>
PB_D void lascl( const G& g, double cfrom, double cto, int rows, int cols, double* A, int lda ) {
  const int t = g.tid(), nt = g.size();
  const double smlnum = kSafmin, bignum = 1.0 / smlnum;
  double cfromc = cfrom, ctoc = cto;
  bool done = false;
  int guard = 0;
  while ( !done && guard++ < 64 ) {
    const double cfrom1 = cfromc * smlnum;
    double mul;
    if ( cfrom1 == cfromc ) {
      mul = ctoc / cfromc;
      done = true;
    } else {
      const double cto1 = ctoc / bignum;
      if ( cto1 == ctoc ) {
        mul = ctoc;
        done = true;
        cfromc = 1.0;
      } else if ( fabs( cfrom1 ) > fabs( ctoc ) && ctoc != 0.0 ) {
        mul = smlnum;
        done = false;
        cfromc = cfrom1;
      } else if ( fabs( cto1 ) > fabs( cfromc ) ) {
        mul = bignum;
        done = false;
        ctoc = cto1;
      } else {
        mul = ctoc / cfromc;
        done = true;
      }
    }
    for ( int c = 0; c < cols; ++c )
      for ( int r = t; r < rows; r += nt ) A[r + (size_t)c * lda] *= mul;
  }
  g.sync();
}

// Scratch carved by the caller: dwork >= 5*n doubles, iwork >= 2*n+2 ints, visible to the group.
// Returns the LAPACK info value (0, or i > 0 when the QR iteration failed: wr/wi[i..n) are valid).
// On exit with wantvr, V (n x n, ldv) holds the right eigenvectors. A is destroyed.
// QR-stage policies: plain double-shift iteration (n <= 64 and the size-generic path) ...
struct QrLahqr {
  template <class G>
  PB_D int operator()( const G& g, bool wantt, bool wantz, int n, int ilo, int ihi, double* H, int ldh, double* wr,
                       double* wi, double* Z, int ldz ) const {
    // (a single warp runs the register-resident sweep, whose step loop wants the branch-free reflector)
    return lahqr<G, ( GroupLanes<G>::value > 0 )>( g, wantt, wantz, n, ilo, ihi, H, ldh, wr, wi, 0, n - 1, Z, ldz );
  }
};

// ... or windowed multishift sweeps with delayed reflector application (pb_msqr.cuh)
struct QrMultishift {
  MsqrCfg cfg;
  MsqrWork wk;
  template <class G>
  PB_D int operator()( const G& g, bool wantt, bool wantz, int n, int ilo, int ihi, double* H, int ldh, double* wr,
                       double* wi, double* Z, int ldz ) const {
    return hqr_multishift( g, wantt, wantz, n, ilo, ihi, H, ldh, wr, wi, 0, n - 1, Z, ldz, cfg, wk );
  }
};

// phase-profile slots (CtaGroup::tick)
constexpr int kTickBalance = 0, kTickHess = 1, kTickOrghr = 2, kTickQr = 3, kTickTrevc = 4, kTickBak = 5, kTickVecGemm = 6;

// scale check (dlange 'M' + dlascl) and balancing. Returns false when the matrix holds a non-finite entry.
struct GeevScaling {
  bool scalea;
  double anrm, cscale;
};

template <class G>
PB_D bool geev_prologue( const G& g, int n, double* A, int lda, GeevScaling& sc, int* ilo, int* ihi, double* scale,
                         int* iwork, double* work4n = nullptr ) {
  const int t = g.tid(), nt = g.size();
  double anrm = 0.0;
  bool bad = false;
  for ( int c = 0; c < n; ++c )
    for ( int r = t; r < n; r += nt ) {
      const double a = fabs( PB_A( r, c ) );
      if ( !( a <= DBL_MAX ) ) bad = true;
      anrm = dmax( anrm, a );
    }
  anrm = g.max( anrm );
  bad = g.any( bad );
  if ( bad ) return false;
  const double smlnum0 = sqrt( kSafmin ) / kUlp;
  const double bignum0 = 1.0 / smlnum0;
  sc.scalea = false;
  sc.anrm = anrm;
  sc.cscale = 1.0;
  if ( anrm > 0.0 && anrm < smlnum0 ) {
    sc.scalea = true;
    sc.cscale = smlnum0;
  } else if ( anrm > bignum0 ) {
    sc.scalea = true;
    sc.cscale = bignum0;
  }
  if ( sc.scalea ) lascl( g, anrm, sc.cscale, n, n, A, lda );
  gebal( g, n, A, lda, ilo, ihi, scale, iwork, work4n );
  g.sync();
  return true;
}

// undo the scaling on the converged eigenvalues (and on the isolated head when the iteration failed)
template <class G>
PB_D void geev_unscale( const G& g, int n, const GeevScaling& sc, int info, int ilo, double* wr, double* wi ) {
  if ( sc.scalea ) {
    lascl( g, sc.cscale, sc.anrm, n - info, 1, wr + info, n );
    lascl( g, sc.cscale, sc.anrm, n - info, 1, wi + info, n );
    if ( info > 0 ) {
      lascl( g, sc.cscale, sc.anrm, ilo, 1, wr, n );
      lascl( g, sc.cscale, sc.anrm, ilo, 1, wi, n );
    }
  }
  g.sync();
}

template <class G>
PB_D void geev_isolated_eigenvalues( const G& g, int n, int ilo, int ihi, const double* A, int lda, double* wr, double* wi ) {
  const int t = g.tid(), nt = g.size();
  for ( int i = t; i < n; i += nt )
    if ( i < ilo || i > ihi ) {
      wr[i] = PB_A( i, i );
      wi[i] = 0.0;
    }
  g.sync();
}

// wantvl / wantvr: left / right eigenvectors into VL / VR (n x n). With both, the Schur vectors are accumulated in VL
// and copied to VR before the triangular solves, as dgeev does.
template <class G, class QR = QrLahqr>
PB_D int geev_group( const G& g, int n, double* A, int lda, double* wr, double* wi, bool wantvl, double* VL, int ldvl, bool wantvr,
                     double* VR, int ldvr, double* dwork, int* iwork, const QR& qr = QR() ) {
  const int t = g.tid(), nt = g.size();
  if ( n == 0 ) return 0;
  double* tau = dwork;
  double* scale = dwork + n;
  double* cnorm = dwork + 2 * n;
  double* xr = dwork + 3 * n;
  double* xi = dwork + 4 * n;
  GeevScaling sc;
  int ilo, ihi;
  if ( !geev_prologue( g, n, A, lda, sc, &ilo, &ihi, scale, iwork ) ) return -4;  // non-finite input (argument 4 of dgeev)
  g.tick( kTickBalance );
  gehd2( g, n, ilo, ihi, A, lda, tau, cnorm );
  g.tick( kTickHess );
  const bool wantz = wantvl || wantvr;
  double* Z = wantvl ? VL : VR;
  const int ldz = wantvl ? ldvl : ldvr;
  if ( wantz ) orghr( g, n, ilo, ihi, A, lda, tau, Z, ldz );
  g.sync();
  g.tick( kTickOrghr );
  geev_isolated_eigenvalues( g, n, ilo, ihi, A, lda, wr, wi );
  int info = qr( g, wantz, wantz, n, ilo, ihi, A, lda, wr, wi, Z, ldz );
  g.sync();
  g.tick( kTickQr );
  if ( info == 0 && wantz ) {
    if ( wantvl && wantvr ) {
      for ( int c = 0; c < n; ++c )
        for ( int r = t; r < n; r += nt ) VR[r + (size_t)c * ldvr] = VL[r + (size_t)c * ldvl];
      g.sync();
    }
    // quasi-triangular T: the reflector storage below the first subdiagonal is never read again
    if ( wantvl ) trevc_left_inplace( g, n, A, lda, VL, ldvl, cnorm, xr, xi );
    if ( wantvr ) trevc_right_inplace( g, n, A, lda, VR, ldvr, cnorm, xr, xi );
    g.tick( kTickTrevc );
    if ( wantvl ) {
      gebak_vectors( g, true, n, ilo, ihi, scale, n, VL, ldvl );
      g.sync();
      normalize_vectors( g, n, wi, VL, ldvl );
    }
    if ( wantvr ) {
      gebak_vectors( g, false, n, ilo, ihi, scale, n, VR, ldvr );
      g.sync();
      normalize_vectors( g, n, wi, VR, ldvr );
    }
    g.tick( kTickBak );
  }
  geev_unscale( g, n, sc, info, ilo, wr, wi );
  return info;
}

}  // namespace pb
