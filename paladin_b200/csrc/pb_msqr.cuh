// pb_msqr.cuh -- QR iteration for mid-size Hessenberg matrices (one CTA per matrix, matrix in global
// memory / L2): small-bulge multishift sweeps with DELAYED reflector application.
//
// Replaces, for 64 < n, the dhseqr stage of the reference's dgeev_ call (reference
// src/lapack-wrapper.hpp:47-60, src/paladin.hpp:248-254); LAPACK itself switches from the plain
// double-shift dlahqr to the small-bulge multishift dlaqr0/dlaqr5 for n > 75. This is not dlaqr5:
//
//   * a chain of nb tightly packed 3x3 bulges (2 nb shifts = eigenvalues of the trailing 2nb x 2nb
//     block) is chased through a W x W diagonal WINDOW of H held on chip (shared memory);
//   * inside the window every macro-step advances all bulges by one row (reflector generation,
//     row updates, column updates: three barriers), so the latency-bound part touches only on-chip data;
//   * the 3-element Householder reflectors are RECORDED, not accumulated into a dense orthogonal
//     factor: the off-window parts of H (right of the window, above it) and the Schur vectors Z are
//     then updated strip by strip: a tile of lines (columns of the right strip, rows of the upper strip
//     and of Z) is staged on chip once, each thread streams the recorded reflectors over ITS line with
//     a three-register sliding window (5 FMAs per reflector and line, no barriers), and the tile is
//     written back once. Flops are those of the unblocked algorithm (no dense-U overhead) while every
//     off-window element moves through the memory system once per window instead of once per bulge.
//   * when the active block is <= ws_small rows, or splits, the block is finished by the plain
//     double-shift iteration (lahqr) inside the window with its transformations accumulated into a
//     small orthogonal U that is applied to the strips as x <- U^T x per line.
//
// Deflation criterion, exceptional shifts and 2x2 standardisation are those of pb::lahqr (LAPACK dlahqr /
// dlanv2), so the pair convention consumed by the reference (src/paladin.hpp:262-263) is kept.
// 0-based inclusive indices. G is a CTA (device) or a host group (tests/hostsim).
#pragma once

#include "pb_common.cuh"
#include "pb_small.cuh"
#include "pb_strips.cuh"
#include "pb_aed.cuh"
#if defined( __CUDACC__ )
#include <cooperative_groups.h>
#endif

namespace pb {

// H entries another CTA of the cluster may have updated are read through L2
#define PB_HLD( i, j ) PB_LDCG( &PB_H( i, j ) )

constexpr int kMsMaxBulges = 16;
// ints of MsqrWork::ibuf: kpos, kfirst, kcount (kMsMaxBulges each), then
//   +0 shift count  +1 sweeps  +2 windows  +3 small solves  +4 scratch  +5 steady-state windows  +6 strip work queue
//   +8..+11 AED results (ns, lahqr code, changed, shifts valid)  +12 eigenvalues deflated by AED  +13 sweeps skipped  +14 AED calls
constexpr int kMsIbufInts = 3 * kMsMaxBulges + 16 + 8;  // (+15: result code of the speculative shift computation)
constexpr int kReflStride = 6;  // w2, w3, t1, t2, t3, pad  (three 16-byte loads)
// phase-profile slots (CtaGroup::tick)
constexpr int kTickMsScan = 8, kTickMsShifts = 9, kTickMsChase = 10, kTickMsStrips = 11, kTickMsSmall = 12, kTickMsSmallU = 13, kTickMsAed = 14, kTickMsAedU = 15;

struct MsqrCfg {
  int nb;        // bulges per chain, <= kMsMaxBulges
  int W;         // window order, >= 3*nb + 4
  int ws_small;  // active blocks of order <= ws_small are finished inside the window (<= W)
  int tl;        // lines per strip tile
  int nchains;   // chains (of nb bulges) per sweep
  int nshift;    // shifts computed per sweep (<= 2*nb*nchains); chains beyond nshift/(2 nb) reuse them cyclically
  int aed = 0;   // order of the aggressive-early-deflation window (<= W - 1); 0 = no AED
  int aed_moves = 1 << 20;  // undeflatable blocks an AED step may reorder (see pb::aed_window)
  int nibble = 14;          // a sweep is skipped when AED deflated more than nibble % of its window (LAPACK: 14)
  int spec_shifts = 0;      // 1: the shifts of the next sweep are computed from the block as it is BEFORE the AED step
                            // (on the device: by another warp, while the AED window is being factored)
  int itmax = 0;            // > 0: cap on the number of sweeps (LAPACK: 30 * max(10, nh)); tests force info > 0 with it
};

// on-chip work space (shared memory on the device)
struct MsqrWork {
  double* Hw;     // W x ldw   window of H
  double* U;      // W x ldw   accumulated transformations of a small solve / scratch for shifts
  double* tile;   // W x ldt   strip tile, element (pos, line) at pos*ldt + line; the x <- U^T x update
                  //           uses lines [0, tl/2) as input and [tl/2, tl) as output
  double* refl;   // nb * W * kReflStride
  double* sr;     // 2*nb*nchains shifts (real parts)
  double* si;     //                     (imaginary parts)
  int* ibuf;      // kMsIbufInts ints: kpos, kfirst, kcount, counters
  int ldw, ldt;
  // thread-block cluster that shares the strip updates of this matrix (device, large matrices): the leader (rank 0)
  // runs the whole iteration, the others wait in ms_cluster_helper for its commands
  int csize = 1, crank = 0;
  int overlap = 1;             // cluster: the leader chases the next window while the helpers finish the strips (needs gscratch)
  double* gscratch = nullptr;  // optional global scratch (>= nb * W * kReflStride doubles): the leader publishes the recorded
                               // reflectors there and the helpers read them through L2 instead of all pulling them out of
                               // the leader's shared memory
};

// command block the leader publishes in its ibuf (read by the helpers through distributed shared memory)
constexpr int kMsCmd = 3 * kMsMaxBulges + 16;  // cmd (0 = exit, 1 = strips of a steady-state window), ws, wlen, l, i
constexpr int kMsCmdInts = 8;

PB_HD size_t ms_even( size_t x ) { return ( x + 1 ) & ~(size_t)1; }  // keeps every sub-buffer 16-byte aligned
PB_HD size_t msqr_work_doubles( const MsqrCfg& c, int ldw, int ldt ) {
  return 2 * ms_even( (size_t)c.W * ldw ) + ms_even( (size_t)c.W * ldt ) + (size_t)c.nb * c.W * kReflStride + 4 * (size_t)c.nb * c.nchains;
}

// ---- deflation scan: largest k in (lo, hi] with negligible H(k,k-1); lo when none -----------------
template <class G>
PB_D int ms_find_split( const G& g, int ilo, int ihi, int lo, int hi, const double* H, int ldh, double smlnum ) {
  const int t = g.tid(), nt = g.size();
  int kfound = lo;
  for ( int k = hi - t; k > lo; k -= nt ) {
    const double hkk1 = fabs( PB_HLD( k, k - 1 ) );
    bool small = false;
    if ( hkk1 <= smlnum ) {
      small = true;
    } else {
      double tst = fabs( PB_HLD( k - 1, k - 1 ) ) + fabs( PB_HLD( k, k ) );
      if ( tst == 0.0 ) {
        if ( k - 2 >= ilo ) tst += fabs( PB_HLD( k - 1, k - 2 ) );
        if ( k + 1 <= ihi ) tst += fabs( PB_HLD( k + 1, k ) );
      }
      if ( hkk1 <= kUlp * tst ) {
        const double hk1k = fabs( PB_HLD( k - 1, k ) );
        const double ab = dmax( hkk1, hk1k ), ba = dmin( hkk1, hk1k );
        const double dd = fabs( PB_HLD( k - 1, k - 1 ) - PB_HLD( k, k ) );
        const double aa = dmax( fabs( PB_HLD( k, k ) ), dd ), bb = dmin( fabs( PB_HLD( k, k ) ), dd );
        const double s = aa + ab;
        if ( ba * ( ab / s ) <= dmax( smlnum, kUlp * ( bb * ( aa / s ) ) ) ) small = true;
      }
    }
    if ( small ) {
      kfound = k;
      break;
    }
  }
  return g.imax_( kfound );
}

constexpr int kWinBatch = 8;
// warps that run the in-window dense problems (pb::WarpsGroup): Schur factorisations with vectors have three 48-wide
// loops per Francis step, the eigenvalue-only shift computation two
// (measured at n = 512, two CTAs per SM: 1 warp 263 ms per 592 matrices, 3/2 warps 275 ms -- the Francis step is bound by
// the latency of its reflector, not by its loops, and a named barrier costs more than __syncwarp)
#ifndef PB_SMALL_WARPS
#define PB_SMALL_WARPS 1
#endif
#ifndef PB_SHIFT_WARPS
#define PB_SHIFT_WARPS 1
#endif
constexpr int kSmallWarps = PB_SMALL_WARPS, kShiftWarps = PB_SHIFT_WARPS;
#if defined( __CUDACC__ )
template <int NW>
__device__ __forceinline__ auto small_group( double* red ) {
  if constexpr ( NW == 1 ) return WarpGroup();
  else return WarpsGroup{ NW, red };
}
#endif

// ---- window <-> global ----------------------------------------------------------------------------
// Only the band r - c <= 3 is meaningful below the diagonal (Hessenberg + parked bulges); whatever else
// the array holds there (Householder vectors of the reduction) is neither read nor written.
template <class G>
PB_D void ms_load_window( const G& g, const double* __restrict__ H, int ldh, int ws, int wlen, double* __restrict__ Hw, int ldw ) {
  const int t = g.tid(), nt = g.size();
  const int total = wlen * wlen;
  for ( int q0 = t; q0 < total; q0 += nt * kWinBatch ) {
    double v[kWinBatch];
#pragma unroll
    for ( int u = 0; u < kWinBatch; ++u ) {
      const int q = imin( q0 + u * nt, total - 1 );
      const int c = q / wlen, r = q - c * wlen;
      v[u] = PB_HLD( ws + r, ws + c );
    }
#pragma unroll
    for ( int u = 0; u < kWinBatch; ++u ) {
      const int q = q0 + u * nt;
      if ( q < total ) {
        const int c = q / wlen, r = q - c * wlen;
        Hw[r + c * ldw] = ( r - c <= 3 ) ? v[u] : 0.0;
      }
    }
  }
  g.sync();
}
template <class G>
PB_D void ms_store_window( const G& g, double* H, int ldh, int ws, int wlen, const double* Hw, int ldw ) {
  const int t = g.tid(), nt = g.size();
  for ( int q = t; q < wlen * wlen; q += nt ) {
    const int c = q / wlen, r = q - c * wlen;
    if ( r - c <= 3 ) PB_H( ws + r, ws + c ) = Hw[r + c * ldw];
  }
  g.sync();
}

// ---- strip tiles ----------------------------------------------------------------------------------
// kind 0: lines are COLUMNS j0.. of M, positions are rows ws..   (strip right of the window)
// kind 1: lines are ROWS r0.. of M, positions are columns ws..   (strip above the window, Schur vectors)
// Loads are issued in batches of kTileBatch before the dependent shared-memory stores: the compiler may not
// reorder a global load above a store through a generic pointer, so a load/store-per-iteration loop would
// expose the full memory latency for every element.
constexpr int kTileBatch = 8;

template <class G>
PB_D void ms_tile_load( const G& g, int kind, const double* __restrict__ M, int ldm, int ws, int wlen, int l0, int nl,
                        double* __restrict__ tile, int ldt ) {
  const int t = g.tid(), nt = g.size();
  const int total = wlen * nl;
  for ( int q0 = t; q0 < total; q0 += nt * kTileBatch ) {
    double v[kTileBatch];
    int dst[kTileBatch];
#pragma unroll
    for ( int u = 0; u < kTileBatch; ++u ) {
      const int q = imin( q0 + u * nt, total - 1 );
      int line, pos;
      if ( kind == 0 ) {
        line = q / wlen;
        pos = q - line * wlen;
        v[u] = PB_LDCG( &M[( ws + pos ) + (size_t)( l0 + line ) * ldm] );
      } else {
        pos = q / nl;
        line = q - pos * nl;
        v[u] = PB_LDCG( &M[( l0 + line ) + (size_t)( ws + pos ) * ldm] );
      }
      dst[u] = pos * ldt + line;
    }
#pragma unroll
    for ( int u = 0; u < kTileBatch; ++u )
      if ( q0 + u * nt < total ) tile[dst[u]] = v[u];
  }
  g.sync();
}
template <class G>
PB_D void ms_tile_store( const G& g, int kind, double* __restrict__ M, int ldm, int ws, int wlen, int l0, int nl,
                         const double* __restrict__ tile, int ldt ) {
  const int t = g.tid(), nt = g.size();
  if ( kind == 0 ) {
    for ( int q = t; q < wlen * nl; q += nt ) {
      const int line = q / wlen, pos = q - line * wlen;
      M[( ws + pos ) + (size_t)( l0 + line ) * ldm] = tile[pos * ldt + line];
    }
  } else {
    for ( int q = t; q < wlen * nl; q += nt ) {
      const int pos = q / nl, line = q - pos * nl;
      M[( l0 + line ) + (size_t)( ws + pos ) * ldm] = tile[pos * ldt + line];
    }
  }
  g.sync();
}

// every thread streams the recorded reflectors over its own lines (no barriers inside)
template <class G>
PB_D void ms_tile_apply_reflectors( const G& g, int nl, int wlen, int ws, int nbulges, const int* kfirst, const int* kcount,
                                    const double* refl, int W, double* tile, int ldt ) {
  const int t = g.tid(), nt = g.size();
  for ( int line = t; line < nl; line += nt ) {
    double* x = tile + line;
    for ( int b = 0; b < nbulges; ++b ) {
      const int cnt = kcount[b];
      if ( cnt == 0 ) continue;
      const int p = kfirst[b] - ws;
      const double* rf = refl + (size_t)b * W * kReflStride;
      double x0 = x[p * ldt];
      double x1 = x[( p + 1 ) * ldt];
      double x2 = ( p + 2 < wlen ) ? x[( p + 2 ) * ldt] : 0.0;
      for ( int s = 0; s < cnt; ++s ) {
        const double w2 = rf[0], w3 = rf[1], t1 = rf[2], t2 = rf[3], t3 = rf[4];
        rf += kReflStride;
        const double sum = x0 + w2 * x1 + w3 * x2;
        x[( p + s ) * ldt] = x0 - sum * t1;
        x0 = x1 - sum * t2;
        x1 = x2 - sum * t3;
        x2 = ( p + s + 3 < wlen ) ? x[( p + s + 3 ) * ldt] : 0.0;
      }
      x[( p + cnt ) * ldt] = x0;
      if ( p + cnt + 1 < wlen ) x[( p + cnt + 1 ) * ldt] = x1;
    }
  }
  g.sync();
}

// x <- U^T x for every line of the tile (U is m x m, ldu), result in tile2
template <class G>
PB_D void ms_tile_apply_u( const G& g, int nl, int m, const double* U, int ldu, const double* tile, double* tile2,
                           int ldt ) {
  const int t = g.tid(), nt = g.size();
  for ( int line = t; line < nl; line += nt ) {
    const double* x = tile + line;
    double* y = tile2 + line;
    for ( int c0 = 0; c0 < m; c0 += 8 ) {
      double acc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
      const int cw = imin( 8, m - c0 );
      for ( int p = 0; p < m; ++p ) {
        const double xv = x[p * ldt];
        const double* up = U + p;
#pragma unroll
        for ( int c = 0; c < 8; ++c )
          if ( c < cw ) acc[c] += xv * up[(size_t)( c0 + c ) * ldu];
      }
#pragma unroll
      for ( int c = 0; c < 8; ++c )
        if ( c < cw ) y[( c0 + c ) * ldt] = acc[c];
    }
  }
  g.sync();
}

// Applies either the recorded reflectors (U == nullptr) or x <- U^T x to every off-window strip.
template <class G>
PB_D void ms_update_strips( const G& g, bool wantt, bool wantz, int n, int l, int i, int ws, int wlen, double* H, int ldh,
                            int iloz, int ihiz, double* Z, int ldz, const MsqrCfg& cfg, const MsqrWork& wk, int nbulges,
                            const double* U, int crank = 0, int csize = 1 ) {
  const int we = ws + wlen - 1;
  const int i1 = wantt ? 0 : l, i2 = wantt ? n - 1 : i;
  const int* kfirst = wk.ibuf + kMsMaxBulges;
  const int* kcount = wk.ibuf + 2 * kMsMaxBulges;
  for ( int pass = 0; pass < 3; ++pass ) {
    double* M;
    int ldm, kind, lo, hi;
    if ( pass == 0 ) {  // right of the window: columns we+1..i2
      M = H; ldm = ldh; kind = 0; lo = we + 1; hi = i2;
    } else if ( pass == 1 ) {  // above the window: rows i1..ws-1
      M = H; ldm = ldh; kind = 1; lo = i1; hi = ws - 1;
    } else {
      if ( !wantz ) break;
      M = Z; ldm = ldz; kind = 1; lo = iloz; hi = ihiz;
    }
    const int tl = ( U == nullptr ) ? cfg.tl : cfg.tl / 2;
    for ( int l0 = lo + crank * tl; l0 <= hi; l0 += tl * csize ) {
      const int nl = imin( tl, hi - l0 + 1 );
      ms_tile_load( g, kind, M, ldm, ws, wlen, l0, nl, wk.tile, wk.ldt );
      if ( U == nullptr ) {
        ms_tile_apply_reflectors( g, nl, wlen, ws, nbulges, kfirst, kcount, wk.refl, cfg.W, wk.tile, wk.ldt );
        ms_tile_store( g, kind, M, ldm, ws, wlen, l0, nl, wk.tile, wk.ldt );
      } else {
        ms_tile_apply_u( g, nl, wlen, U, wk.ldw, wk.tile, wk.tile + tl, wk.ldt );
        ms_tile_store( g, kind, M, ldm, ws, wlen, l0, nl, wk.tile + tl, wk.ldt );
      }
    }
  }
}

#if defined( __CUDACC__ )
// ---- leader side of the cluster hand-off: publish a command, meet the helpers (they copy what they need from this
// CTA's shared memory), and meet them again when everybody's share is done (pb::ms_cluster_helper) ------------------------
// cmd 1: strips of a steady-state window (register path); 2: x <- U^T x with wk.U; 3: recorded reflectors, tile path
__device__ __forceinline__ void ms_cluster_post( const MsqrWork& wk, int cmd, int ws, int wlen, int l, int i, int nbulges ) {
  if ( wk.gscratch != nullptr && cmd != 2 ) {
    const int nrefl2 = kStripNB * kStripW * kReflStride / 2;
    for ( int q = threadIdx.x; q < nrefl2; q += blockDim.x )
      __stcg( reinterpret_cast<double2*>( wk.gscratch ) + q, reinterpret_cast<const double2*>( wk.refl )[q] );
  }
  if ( threadIdx.x == 0 ) {
    int* c = wk.ibuf + kMsCmd;
    c[0] = cmd; c[1] = ws; c[2] = wlen; c[3] = l; c[4] = i; c[5] = nbulges;
  }
  cooperative_groups::this_cluster().sync();
}
__device__ __forceinline__ void ms_cluster_join() { cooperative_groups::this_cluster().sync(); }
#endif

// the off-window update of one window / block, shared with the cluster when there is one
template <class G>
PB_D void ms_update_strips_shared( const G& g, bool wantt, bool wantz, int n, int l, int i, int ws, int wlen, double* H, int ldh,
                                   int iloz, int ihiz, double* Z, int ldz, const MsqrCfg& cfg, const MsqrWork& wk, int nbulges,
                                   const double* U ) {
#if defined( __CUDACC__ )
  if constexpr ( G::kIsCta ) {
    if ( wk.csize > 1 ) {
      ms_cluster_post( wk, U != nullptr ? 2 : 3, ws, wlen, l, i, nbulges );
      ms_update_strips( g, wantt, wantz, n, l, i, ws, wlen, H, ldh, iloz, ihiz, Z, ldz, cfg, wk, nbulges, U, 0, wk.csize );
      ms_cluster_join();
      return;
    }
  }
#endif
  ms_update_strips( g, wantt, wantz, n, l, i, ws, wlen, H, ldh, iloz, ihiz, Z, ldz, cfg, wk, nbulges, U );
}


// ---- finish a small diagonal block [l..i] (order <= W) inside the window ---------------------------
// Returns 0 or the lahqr failure index (global, 1-based like LAPACK's info).
template <class G>
PB_D int ms_small_solve( const G& g, bool wantt, bool wantz, int n, int l, int i, double* H, int ldh, double* wr, double* wi,
                         int iloz, int ihiz, double* Z, int ldz, const MsqrCfg& cfg, const MsqrWork& wk ) {
  const int t = g.tid(), nt = g.size();
  const int m = i - l + 1;
  if ( m == 1 ) {
    if ( t == 0 ) {
      wr[i] = PB_HLD( i, i );
      wi[i] = 0.0;
    }
    g.sync();
    return 0;
  }
  const int ldw = wk.ldw;
  if ( t == 0 ) wk.ibuf[3 * kMsMaxBulges + 3] += 1;
  ms_load_window( g, H, ldh, l, m, wk.Hw, ldw );
  const bool acc = wantt || wantz;
  if ( acc ) {
    for ( int q = t; q < m * m; q += nt ) {
      const int c = q / m, r = q - c * m;
      wk.U[r + c * ldw] = ( r == c ) ? 1.0 : 0.0;
    }
    g.sync();
  }
  int info = 0;
  bool warp_done = false;
#if defined( __CUDACC__ )
  if constexpr ( G::kIsCta ) {
    // a block of <= 48 rows keeps a few warps busy at best: they run it behind their own named barrier
    if ( ( threadIdx.x >> 5 ) < kSmallWarps ) {
      const int r = lahqr( small_group<kSmallWarps>( g.red ), acc, acc, m, 0, m - 1, wk.Hw, ldw, wr + l, wi + l, 0, m - 1, wk.U, ldw );
      if ( threadIdx.x == 0 ) wk.ibuf[3 * kMsMaxBulges + 4] = r;
    }
    __syncthreads();
    info = wk.ibuf[3 * kMsMaxBulges + 4];
    warp_done = true;
  }
#endif
  if ( !warp_done ) info = lahqr( g, acc, acc, m, 0, m - 1, wk.Hw, ldw, wr + l, wi + l, 0, m - 1, wk.U, ldw );
  g.sync();
  g.tick( kTickMsSmall );
  if ( info != 0 ) return l + info;
  if ( !acc ) return 0;
  // Schur form of the block back to H (strictly below the first subdiagonal stays untouched)
  for ( int q = t; q < m * m; q += nt ) {
    const int c = q / m, r = q - c * m;
    if ( r - c <= 1 ) PB_H( l + r, l + c ) = wk.Hw[r + c * ldw];
    else if ( r - c <= 3 ) PB_H( l + r, l + c ) = 0.0;
  }
  g.sync();
  ms_update_strips_shared( g, wantt, wantz, n, l, i, l, m, H, ldh, iloz, ihiz, Z, ldz, cfg, wk, 0, wk.U );
  g.tick( kTickMsSmallU );
  return 0;
}

// ---- shift list -> pairs ----------------------------------------------------------------------------
// in: candidates ks..ns-1 of sr/si (any arrays visible to the group); out: at most `keep` of them (the smallest
// moduli) at the front of wk.sr / wk.si, arranged in pairs (conjugates adjacent, reals two by two). Returns the count.
template <class G>
PB_D int ms_arrange_shifts( const G& g, double* sr, double* si, int ks, int ns, int keep, const MsqrWork& wk ) {
  const int t = g.tid();
  // one thread: sort by decreasing modulus (helps convergence a little), then arrange into pairs
  if ( t == 0 ) {
    bool sorted = false;
    for ( int k = ns - 1; k >= ks + 1 && !sorted; --k ) {
      sorted = true;
      for ( int q = ks; q < k; ++q ) {
        if ( fabs( sr[q] ) + fabs( si[q] ) < fabs( sr[q + 1] ) + fabs( si[q + 1] ) ) {
          sorted = false;
          double tmp = sr[q]; sr[q] = sr[q + 1]; sr[q + 1] = tmp;
          tmp = si[q]; si[q] = si[q + 1]; si[q + 1] = tmp;
        }
      }
    }
    // conjugate pairs are adjacent after the sort (equal modulus); rotate triples where a real shift
    // would split a pair
    for ( int q = ns - 1; q >= ks + 2; q -= 2 ) {
      if ( si[q] != -si[q - 1] ) {
        double tmp = sr[q]; sr[q] = sr[q - 1]; sr[q - 1] = sr[q - 2]; sr[q - 2] = tmp;
        tmp = si[q]; si[q] = si[q - 1]; si[q - 1] = si[q - 2]; si[q - 2] = tmp;
      }
    }
    // keep the trailing (smallest) ones; an odd count drops the first shift
    int first = imax( ks, ns - keep );
    if ( ( ( ns - first ) & 1 ) != 0 ) first += 1;
    // a pair that is neither conjugate nor real-real (cannot happen with exact conjugates) is made real
    for ( int q = first; q + 1 < ns; q += 2 ) {
      if ( si[q] != -si[q + 1] ) {
        si[q] = 0.0;
        si[q + 1] = 0.0;
      }
    }
    // compact to the front of the work arrays
    const int cnt = ns - first;
    for ( int q = 0; q < cnt; ++q ) {
      wk.sr[q] = sr[first + q];
      wk.si[q] = si[first + q];
    }
    wk.ibuf[3 * kMsMaxBulges] = cnt;
  }
  g.sync();
  return wk.ibuf[3 * kMsMaxBulges];
}

// ---- shifts: eigenvalues of the trailing ns x ns block of [l..i] -----------------------------------
// Leaves ns_out (even, >= 2) shifts in wk.sr / wk.si, arranged in pairs (conjugates adjacent, reals two by two).
template <class G>
PB_D int ms_compute_shifts( const G& g, int l, int i, const double* H, int ldh, int ns, bool exceptional, const MsqrWork& wk ) {
  const int t = g.tid(), nt = g.size();
  double* sr = wk.sr;
  double* si = wk.si;
  int ks = 0;  // first valid shift
  if ( exceptional ) {
    if ( t == 0 ) {
      for ( int q = ns - 1; q >= 1; q -= 2 ) {
        const int r = i - ( ns - 1 - q );  // row of H matched with shift q
        const double ss = fabs( PB_HLD( r, r - 1 ) ) + fabs( PB_HLD( r - 1, r - 2 ) );
        double aa = 0.75 * ss + PB_HLD( r, r ), bb = ss, cc = -0.4375 * ss, dd = aa;
        double cs, sn;
        dlanv2( aa, bb, cc, dd, sr[q - 1], si[q - 1], sr[q], si[q], cs, sn );
      }
    }
    g.sync();
  } else {
    const int ldw = wk.ldw;
    double* S = wk.U;
    const int r0 = i - ns + 1;
    for ( int q = t; q < ns * ns; q += nt ) {
      const int c = q / ns, r = q - c * ns;
      S[r + c * ldw] = ( r - c <= 1 ) ? PB_HLD( r0 + r, r0 + c ) : 0.0;
    }
    g.sync();
    int inf = 0;
    bool warp_done = false;
#if defined( __CUDACC__ )
    if constexpr ( G::kIsCta ) {
      // a 2nb x 2nb problem keeps two warps busy at best: they run it behind their own named barrier
      if ( ( threadIdx.x >> 5 ) < kShiftWarps ) {
        const int r = lahqr<decltype( small_group<kShiftWarps>( g.red ) ), true>( small_group<kShiftWarps>( g.red ), false, false, ns, 0, ns - 1, S, ldw, sr, si, 0, ns - 1, (double*)nullptr, 1 );
        if ( threadIdx.x == 0 ) wk.ibuf[3 * kMsMaxBulges + 4] = r;
      }
      __syncthreads();
      inf = wk.ibuf[3 * kMsMaxBulges + 4];
      warp_done = true;
    }
#endif
    if ( !warp_done ) inf = lahqr( g, false, false, ns, 0, ns - 1, S, ldw, sr, si, 0, ns - 1, (double*)nullptr, 1 );
    g.sync();
    PB_CHECK_UNIFORM( g, inf, "shift lahqr info" );
    ks = inf;  // shifts ks..ns-1 converged
    if ( ns - ks < 2 ) {
      // fall back on the trailing 2x2 block
      if ( t == 0 ) {
        double aa = PB_HLD( i - 1, i - 1 ), bb = PB_HLD( i - 1, i ), cc = PB_HLD( i, i - 1 ), dd = PB_HLD( i, i );
        double cs, sn;
        dlanv2( aa, bb, cc, dd, sr[ns - 2], si[ns - 2], sr[ns - 1], si[ns - 1], cs, sn );
      }
      g.sync();
      ks = ns - 2;
    }
  }
  return ms_arrange_shifts( g, sr, si, ks, ns, ns, wk );
}

#if defined( __CUDACC__ )
// Device version of the macro-step loop of one window: warp b owns bulge b (reflector generation and both
// updates are warp-local), two CTA barriers per macro-step separate the row phase from the column phase.
// Every thread tracks all bulge positions in registers. Requires nbulges <= number of warps, nbulges <= 8.
__device__ inline bool ms_window_steps_fast( int l, int i, int ws, int we, int nbulges, int* kpos_s, int* kcount_s, double* Hw, int ldw,
                                             double* refl, int W, const double* sr, const double* si ) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int kp[8], kc = 0;
#pragma unroll
  for ( int b = 0; b < 8; ++b ) kp[b] = ( b < nbulges ) ? kpos_s[b] : ( i + 1 );
  bool progressed = false;
  // steady-state window (all 8 bulges tightly packed at the window start, window strictly inside the block): every
  // bulge moves in every macro-step for exactly W - 1 - 3*8 steps, no scheduling logic is needed
  bool lockstep = ( nbulges == 8 ) && ( we - ws + 1 == W ) && ( i > we );
#pragma unroll
  for ( int b = 0; b < 8; ++b ) lockstep = lockstep && ( kp[b] == ws + 1 + 3 * ( 7 - b ) );
  int lock_left = W - 1 - 3 * 8;
#define PB_W( r, c ) Hw[( ( r ) - ws ) + ( ( c ) - ws ) * ldw]
  for ( ;; ) {
    unsigned moving = 0;
    if ( lockstep ) {
      moving = ( lock_left > 0 ) ? 0xffu : 0u;
      lock_left -= 1;
    } else {
    bool ahead_moves = false, have_ahead = false;
    int ahead_pos = 0;
#pragma unroll
    for ( int b = 0; b < 8; ++b ) {
      const int k = kp[b];
      bool mv = false;
      if ( b < nbulges && k <= i - 1 ) {
        const bool fits = imin( k + 3, i ) <= we;
        bool clear = true;
        if ( have_ahead ) clear = ( ahead_pos >= k + 4 ) || ( ahead_pos == k + 3 && ahead_moves );
        mv = fits && clear;
        have_ahead = true;
        ahead_pos = k;
        ahead_moves = mv;
      }
      if ( mv ) moving |= 1u << b;
    }
    }
    if ( moving == 0 ) break;
    progressed = true;
    int k = 0;
#pragma unroll
    for ( int b = 0; b < 8; ++b )
      if ( b == warp ) k = kp[b];
    const bool mine = ( warp < nbulges ) && ( ( moving >> warp ) & 1u );
    const bool three = ( k + 2 <= i );
    double w2 = 0.0, w3 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    if ( mine ) {
      const int nr = three ? 3 : 2;
      double a0, a1, a2 = 0.0;
      if ( k == l ) {
        const double rt1r = sr[2 * warp], rt1i = si[2 * warp], rt2r = sr[2 * warp + 1], rt2i = si[2 * warp + 1];
        const double h00 = PB_W( l, l ), h10 = PB_W( l + 1, l ), h01 = PB_W( l, l + 1 ), h11 = PB_W( l + 1, l + 1 );
        double s = fabs( h00 - rt2r ) + fabs( rt2i ) + fabs( h10 );
        if ( s == 0.0 ) {
          a0 = a1 = a2 = 0.0;
        } else {
          const double h21s = h10 / s;
          a0 = h21s * h01 + ( h00 - rt1r ) * ( ( h00 - rt2r ) / s ) - rt1i * ( rt2i / s );
          a1 = h21s * ( h00 + h11 - rt1r - rt2r );
          a2 = three ? h21s * PB_W( l + 2, l + 1 ) : 0.0;
          s = fabs( a0 ) + fabs( a1 ) + fabs( a2 );
          if ( s != 0.0 ) {
            const double rs = 1.0 / s;
            a0 *= rs; a1 *= rs; a2 *= rs;
          }
        }
      } else {
        a0 = PB_W( k, k - 1 );
        a1 = PB_W( k + 1, k - 1 );
        if ( three ) a2 = PB_W( k + 2, k - 1 );
      }
      larfg3_fast( nr, a0, a1, a2, t1 );
      w2 = a1;
      w3 = a2;
      t2 = t1 * w2;
      t3 = t1 * w3;
      __syncwarp();  // every lane has read the bulge column before lane 0 rewrites it
      if ( lane == 0 ) {
        if ( k > l ) {
          PB_W( k, k - 1 ) = a0;
          PB_W( k + 1, k - 1 ) = 0.0;
          if ( three ) PB_W( k + 2, k - 1 ) = 0.0;
        }
        double* rf = refl + ( (size_t)warp * W + kc ) * kReflStride;
        rf[0] = w2; rf[1] = w3; rf[2] = t1; rf[3] = t2; rf[4] = t3;
      }
      // row update: rows k..k+2, columns k..we
      for ( int j = k + lane; j <= we; j += 32 ) {
        const double x0 = PB_W( k, j ), x1 = PB_W( k + 1, j ), x2 = three ? PB_W( k + 2, j ) : 0.0;
        const double sum = x0 + w2 * x1 + w3 * x2;
        PB_W( k, j ) = x0 - sum * t1;
        PB_W( k + 1, j ) = x1 - sum * t2;
        if ( three ) PB_W( k + 2, j ) = x2 - sum * t3;
      }
    }
    __syncthreads();
    if ( mine ) {
      // column update: columns k..k+2, rows ws..min(k+3, i)
      const int rend = imin( k + 3, i );
      for ( int r = ws + lane; r <= rend; r += 32 ) {
        const double x0 = PB_W( r, k ), x1 = PB_W( r, k + 1 ), x2 = three ? PB_W( r, k + 2 ) : 0.0;
        const double sum = x0 + w2 * x1 + w3 * x2;
        PB_W( r, k ) = x0 - sum * t1;
        PB_W( r, k + 1 ) = x1 - sum * t2;
        if ( three ) PB_W( r, k + 2 ) = x2 - sum * t3;
      }
      kc += 1;
    }
    __syncthreads();
#pragma unroll
    for ( int b = 0; b < 8; ++b )
      if ( ( moving >> b ) & 1u ) kp[b] += 1;
  }
#undef PB_W
  if ( lane == 0 && warp < nbulges ) {
#pragma unroll
    for ( int b = 0; b < 8; ++b )
      if ( b == warp ) kpos_s[b] = kp[b];
    kcount_s[warp] = kc;
  }
  __syncthreads();
  return progressed;
}
#endif  // __CUDACC__

// ---- one chain of nbulges bulges swept over the active block [l..i] --------------------------------
template <class G>
PB_D void ms_chase_chain( const G& g, bool wantt, bool wantz, int n, int l, int i, double* H, int ldh, int iloz, int ihiz,
                          double* Z, int ldz, const MsqrCfg& cfg, const MsqrWork& wk, int nbulges, const double* sr,
                          const double* si ) {
  const int t = g.tid(), nt = g.size();
  int* kpos = wk.ibuf;                        // next step of each bulge (l = not yet introduced; > i-1 = finished)
  int* kfirst = wk.ibuf + kMsMaxBulges;
  int* kcount = wk.ibuf + 2 * kMsMaxBulges;
  double* Hw = wk.Hw;
  const int ldw = wk.ldw;
  g.sync();  // the previous chain's last look at kpos is over
  for ( int b = t; b < nbulges; b += nt ) kpos[b] = l;
  g.sync();
  int ws = l;
  bool pending = false;  // cluster: the helpers may still be working on the strips of the previous window
  for ( ;; ) {
    // all bulges gone?
    int first_live = -1, last_live = -1;
    for ( int b = 0; b < nbulges; ++b )
      if ( kpos[b] <= i - 1 ) {
        if ( first_live < 0 ) first_live = b;
        last_live = b;
      }
    PB_CHECK_UNIFORM( g, first_live * 100 + last_live, "live bulges" );
    if ( first_live < 0 ) break;
    // the trailing live bulge fixes the window start (its next step k reads column k-1)
    ws = imax( l, kpos[last_live] - 1 );
    const int we = imin( i, ws + cfg.W - 1 );
    const int wlen = we - ws + 1;
    g.sync();
    ms_load_window( g, H, ldh, ws, wlen, Hw, ldw );
    for ( int b = t; b < nbulges; b += nt ) {
      kfirst[b] = kpos[b];
      kcount[b] = 0;
    }
    g.sync();
    bool progressed = false;
    bool done_fast = false;
#if defined( __CUDACC__ )
    if constexpr ( G::kIsCta ) {
      if ( nbulges <= 8 && nbulges <= ( g.size() >> 5 ) ) {
        progressed = ms_window_steps_fast( l, i, ws, we, nbulges, kpos, kcount, Hw, ldw, wk.refl, cfg.W, sr, si );
        done_fast = true;
      }
    }
#endif
#define PB_W( r, c ) Hw[( ( r ) - ws ) + ( ( c ) - ws ) * ldw]
    for ( ; !done_fast; ) {
      // which bulges step now (every thread evaluates the same shared state)
      unsigned moving = 0;
      bool ahead_moves = false;
      int ahead_pos = 0;
      bool have_ahead = false;
      for ( int b = 0; b < nbulges; ++b ) {
        const int k = kpos[b];
        bool mv = false;
        if ( k <= i - 1 ) {
          const bool fits = imin( k + 3, i ) <= we;
          bool clear = true;
          if ( have_ahead ) clear = ( ahead_pos >= k + 4 ) || ( ahead_pos == k + 3 && ahead_moves );
          mv = fits && clear;
          have_ahead = true;
          ahead_pos = k;
          ahead_moves = mv;
        }
        if ( mv ) moving |= 1u << b;
      }
      PB_CHECK_UNIFORM( g, moving, "moving" );
      if ( moving == 0 ) break;
      progressed = true;
      // phase 1: reflectors (one thread per bulge)
      for ( int b = t; b < nbulges; b += nt ) {
        if ( !( moving & ( 1u << b ) ) ) continue;
        const int k = kpos[b];
        const int nr = imin( 3, i - k + 1 );
        double a0, a1, a2 = 0.0, t1;
        if ( k == l ) {
          const double rt1r = sr[2 * b], rt1i = si[2 * b], rt2r = sr[2 * b + 1], rt2i = si[2 * b + 1];
          double h21s = fabs( PB_W( l + 1, l ) );
          double s = fabs( PB_W( l, l ) - rt2r ) + fabs( rt2i ) + h21s;
          if ( s == 0.0 ) {
            a0 = 0.0; a1 = 0.0; a2 = 0.0;
          } else {
            h21s = PB_W( l + 1, l ) / s;
            a0 = h21s * PB_W( l, l + 1 ) + ( PB_W( l, l ) - rt1r ) * ( ( PB_W( l, l ) - rt2r ) / s ) - rt1i * ( rt2i / s );
            a1 = h21s * ( PB_W( l, l ) + PB_W( l + 1, l + 1 ) - rt1r - rt2r );
            a2 = ( nr == 3 ) ? h21s * PB_W( l + 2, l + 1 ) : 0.0;
            s = fabs( a0 ) + fabs( a1 ) + fabs( a2 );
            if ( s != 0.0 ) {
              a0 /= s; a1 /= s; a2 /= s;
            }
          }
        } else {
          a0 = PB_W( k, k - 1 );
          a1 = PB_W( k + 1, k - 1 );
          if ( nr == 3 ) a2 = PB_W( k + 2, k - 1 );
        }
        larfg3_fast( nr, a0, a1, a2, t1 );
        if ( k > l ) {
          PB_W( k, k - 1 ) = a0;
          PB_W( k + 1, k - 1 ) = 0.0;
          if ( nr == 3 ) PB_W( k + 2, k - 1 ) = 0.0;
        }
        double* rf = wk.refl + ( (size_t)b * cfg.W + kcount[b] ) * kReflStride;
        rf[0] = a1;
        rf[1] = a2;
        rf[2] = t1;
        rf[3] = t1 * a1;
        rf[4] = t1 * a2;
      }
      g.sync();
      // phase 2: row updates inside the window, all moving bulges at once (disjoint rows)
      for ( int q = t; q < nbulges * wlen; q += nt ) {
        const int b = q / wlen, jj = q - b * wlen;
        if ( !( moving & ( 1u << b ) ) ) continue;
        const int k = kpos[b];
        const int j = ws + jj;
        if ( j < k ) continue;
        const double* rf = wk.refl + ( (size_t)b * cfg.W + kcount[b] ) * kReflStride;
        const bool three = ( k + 2 <= i );
        const double x0 = PB_W( k, j ), x1 = PB_W( k + 1, j ), x2 = three ? PB_W( k + 2, j ) : 0.0;
        const double sum = x0 + rf[0] * x1 + rf[1] * x2;
        PB_W( k, j ) = x0 - sum * rf[2];
        PB_W( k + 1, j ) = x1 - sum * rf[3];
        if ( three ) PB_W( k + 2, j ) = x2 - sum * rf[4];
      }
      g.sync();
      // phase 3: column updates inside the window (disjoint columns)
      for ( int q = t; q < nbulges * wlen; q += nt ) {
        const int b = q / wlen, rr = q - b * wlen;
        if ( !( moving & ( 1u << b ) ) ) continue;
        const int k = kpos[b];
        const int r = ws + rr;
        if ( r > imin( k + 3, i ) ) continue;
        const double* rf = wk.refl + ( (size_t)b * cfg.W + kcount[b] ) * kReflStride;
        const bool three = ( k + 2 <= i );
        const double x0 = PB_W( r, k ), x1 = PB_W( r, k + 1 ), x2 = three ? PB_W( r, k + 2 ) : 0.0;
        const double sum = x0 + rf[0] * x1 + rf[1] * x2;
        PB_W( r, k ) = x0 - sum * rf[2];
        PB_W( r, k + 1 ) = x1 - sum * rf[3];
        if ( three ) PB_W( r, k + 2 ) = x2 - sum * rf[4];
      }
      g.sync();
      for ( int b = t; b < nbulges; b += nt )
        if ( moving & ( 1u << b ) ) {
          kpos[b] += 1;
          kcount[b] += 1;
        }
      g.sync();
    }
#undef PB_W
    if ( !progressed ) break;  // cannot happen for W >= 3*nb + 4; never spin
    if ( t == 0 ) wk.ibuf[3 * kMsMaxBulges + 6] = 0;  // work queue of the strip phase (visible after the barrier below)
    ms_store_window( g, H, ldh, ws, wlen, Hw, ldw );
    if ( t == 0 ) {
      wk.ibuf[3 * kMsMaxBulges + 2] += 1;
      bool reg = ( nbulges == cfg.nb ) && ( wlen == cfg.W );
      for ( int b = 0; b < nbulges; ++b )
        reg = reg && ( kfirst[b] - ws == 1 + 3 * ( cfg.nb - 1 - b ) ) && ( kcount[b] == cfg.W - 1 - 3 * cfg.nb );
      if ( reg ) wk.ibuf[3 * kMsMaxBulges + 5] += 1;  // steady-state windows (diagnostic)
    }
    g.tick( kTickMsChase );
#if defined( __CUDACC__ )
    if constexpr ( G::kIsCta ) {
      // device fast path: lines in registers, warps independent (pb_strips.cuh); needs the compile-time window order
      // steady-state windows only (about five in six); the others take the tile path below
      if ( cfg.W == kStripW && cfg.nb == kStripNB && cfg.tl >= 16 * ( g.size() >> 5 ) &&
           strip_window_is_regular( nbulges, wlen, ws, kfirst, kcount ) ) {
        // hand the window to the cluster. With the reflectors published in global memory the leader does not wait for the
        // helpers: it takes its own share -- which holds the right-strip batches 0..3, i.e. every column the next window
        // can reach -- and goes on to chase the next window while the helpers finish this one; they are joined before the
        // next hand-off (or at the end of the chain).
        if ( wk.csize > 1 ) {
          if ( pending ) {
            ms_cluster_join();
            pending = false;
          }
          ms_cluster_post( wk, 1, ws, wlen, l, i, nbulges );
        }
        ms_strips_regs( wantt, wantz, n, l, i, ws, wlen, H, ldh, iloz, ihiz, Z, ldz, wk.refl, wk.tile, wk.ldt,
                        wk.ibuf + 3 * kMsMaxBulges + 6, 0, wk.csize, wk.csize > 2 && wk.gscratch != nullptr && wk.overlap != 0 );
        if ( wk.csize > 1 ) {
          if ( wk.gscratch != nullptr && wk.overlap ) pending = true;
          else ms_cluster_join();
        }
      } else {
        if ( pending ) {
          ms_cluster_join();
          pending = false;
        }
        ms_update_strips_shared( g, wantt, wantz, n, l, i, ws, wlen, H, ldh, iloz, ihiz, Z, ldz, cfg, wk, nbulges, (const double*)nullptr );
      }
    } else
#endif
    {
      ms_update_strips( g, wantt, wantz, n, l, i, ws, wlen, H, ldh, iloz, ihiz, Z, ldz, cfg, wk, nbulges, (const double*)nullptr );
    }
    g.sync();
    g.tick( kTickMsStrips );
  }
#if defined( __CUDACC__ )
  if constexpr ( G::kIsCta ) {
    if ( pending ) ms_cluster_join();
  }
#endif
  (void)pending;
}

#if defined( __CUDACC__ )
// kept out of line: the window code has its own register allocation and stack frame, the sweep code around it keeps its own
__device__ __noinline__ void aed_window_warps( double* red, int jw, int max_moves, double* T, int ldt, double* V, int ldv, double s,
                                               double* sr, double* si, double* work, int* out, long long* t_qr_done ) {
  aed_window( small_group<kSmallWarps>( red ), jw, max_moves, T, ldt, V, ldv, s, sr, si, work, out, t_qr_done );
}
#endif

// ---- aggressive early deflation on the trailing jw x jw window of the active block [ktop..kbot] -----------
// nd_out: eigenvalues deflated (now final in wr/wi, H split below row kbot - nd); ns_out: shift candidates left in
// wk.sr / wk.si (arranged in pairs, at most `keep`).
// spec_ns > 0: the eigenvalues of the trailing spec_ns x spec_ns block AS IT IS BEFORE this AED step are computed as well
// (speculative shifts of the next sweep; on the device by a second warp while the first one factors the AED window);
// *spec_out receives how many of them are left in wk.sr / wk.si (arranged in pairs), 0 when the AED step itself left shifts.
template <class G>
PB_D void ms_aed( const G& g, bool wantt, bool wantz, int n, int ktop, int kbot, int jw, int keep, double* H, int ldh, double* wr,
                  double* wi, int iloz, int ihiz, double* Z, int ldz, const MsqrCfg& cfg, const MsqrWork& wk, int* nd_out,
                  int* ns_out, int spec_ns = 0, int* spec_out = nullptr ) {
  const int t = g.tid(), nt = g.size();
  const int kwtop = kbot - jw + 1;
  const int ldw = wk.ldw;
  const double s = ( kwtop > ktop ) ? PB_HLD( kwtop, kwtop - 1 ) : 0.0;
  // speculative shift block: a copy of the trailing block in the strip-tile buffer, behind the AED scratch
  double* S2 = wk.tile + 4 * cfg.W;
  double* sr2 = S2 + (size_t)cfg.W * ldw;
  double* si2 = sr2 + cfg.W;
  if ( spec_ns > 0 ) {
    const int r0 = kbot - spec_ns + 1;
    for ( int q = t; q < spec_ns * spec_ns; q += nt ) {
      const int c = q / spec_ns, r = q - c * spec_ns;
      S2[r + c * ldw] = ( r - c <= 1 ) ? PB_HLD( r0 + r, r0 + c ) : 0.0;
    }
  }
  ms_load_window( g, H, ldh, kwtop, jw, wk.Hw, ldw );
  for ( int q = t; q < jw * jw; q += nt ) {
    const int c = q / jw, r = q - c * jw;
    wk.U[r + c * ldw] = ( r == c ) ? 1.0 : 0.0;
  }
  if ( t == 0 ) wk.ibuf[3 * kMsMaxBulges + 14] += 1;
  g.sync();
  int* out = wk.ibuf + 3 * kMsMaxBulges + 8;
  double* er = wk.tile;            // eigenvalues of the window (2 jw doubles), then 2 jw doubles of scratch
  double* ei = wk.tile + jw;
  bool warp_done = false;
  int spec_inf = 0;
  long long t_qr_done = 0;
#if defined( __CUDACC__ )
  if constexpr ( G::kIsCta ) {
    // a window of <= 48 rows keeps a few warps busy at best: they run it behind their own named barrier
    if ( ( threadIdx.x >> 5 ) < kSmallWarps ) aed_window_warps( g.red, jw, cfg.aed_moves, wk.Hw, ldw, wk.U, ldw, s, er, ei, wk.tile + 2 * jw, out, &t_qr_done );
    else if ( spec_ns > 0 && ( threadIdx.x >> 5 ) == kSmallWarps ) {
      // the warp next to them: eigenvalues of the speculative shift block, side by side with the window
      const int r = lahqr<WarpGroup, true>( WarpGroup(), false, false, spec_ns, 0, spec_ns - 1, S2, ldw, sr2, si2, 0, spec_ns - 1, (double*)nullptr, 1 );
      if ( ( threadIdx.x & 31 ) == 0 ) wk.ibuf[3 * kMsMaxBulges + 15] = r;
    }
    __syncthreads();
    if ( spec_ns > 0 ) spec_inf = wk.ibuf[3 * kMsMaxBulges + 15];
    warp_done = true;
  }
#endif
  if ( !warp_done ) {
    if ( spec_ns > 0 ) spec_inf = lahqr<G, true>( g, false, false, spec_ns, 0, spec_ns - 1, S2, ldw, sr2, si2, 0, spec_ns - 1, (double*)nullptr, 1 );
    aed_window( g, jw, cfg.aed_moves, wk.Hw, ldw, wk.U, ldw, s, er, ei, wk.tile + 2 * jw, out );
  }
  g.sync();
  const int ns = out[0], infqr = out[1], changed = out[2];
  PB_CHECK_UNIFORM( g, ns * 4 + changed, "aed result" );
  g.tick( kTickMsAed );
#if defined( __CUDACC__ )
  if constexpr ( G::kIsCta ) {
    // slot 3 (unused by this kernel otherwise): the part of the window phase that follows its QR iteration
    if ( g.stats != nullptr && threadIdx.x == 0 && t_qr_done != 0 ) g.stats[3] += g.last - t_qr_done;
  }
#endif
  if ( spec_out != nullptr ) *spec_out = 0;
  if ( infqr != 0 ) {  // the window's own QR iteration failed: nothing was written back, no shifts from here
    *nd_out = 0;
    *ns_out = 0;
    return;
  }
  // deflated eigenvalues are final
  for ( int q = ns + t; q < jw; q += nt ) {
    wr[kwtop + q] = er[q];
    wi[kwtop + q] = ei[q];
  }
  if ( t == 0 ) wk.ibuf[3 * kMsMaxBulges + 12] += jw - ns;
  // shift candidates: eigenvalues of the undeflated part
  int nshift = 0;
  if ( ns >= 2 && out[3] != 0 ) nshift = ms_arrange_shifts( g, er, ei, 0, ns, keep, wk );
  else if ( spec_ns > 0 && spec_out != nullptr && spec_ns - spec_inf >= 2 ) *spec_out = ms_arrange_shifts( g, sr2, si2, spec_inf, spec_ns, spec_ns, wk );
  else g.sync();
  if ( changed ) {
    if ( t == 0 && kwtop > ktop ) PB_H( kwtop, kwtop - 1 ) = s * wk.U[0];
    for ( int q = t; q < jw * jw; q += nt ) {
      const int c = q / jw, r = q - c * jw;
      if ( r - c <= 1 ) PB_H( kwtop + r, kwtop + c ) = wk.Hw[r + c * ldw];
    }
    g.sync();
    ms_update_strips_shared( g, wantt, wantz, n, ktop, kbot, kwtop, jw, H, ldh, iloz, ihiz, Z, ldz, cfg, wk, 0, wk.U );
  }
  *nd_out = jw - ns;
  *ns_out = nshift;
}

#if defined( __CUDACC__ )
// ---- cluster helpers (ranks 1..csize-1): strips of the leader's steady-state windows -------------------------------------
__device__ inline void ms_cluster_helper( const CtaGroup& g, bool wantt, bool wantz, int n, double* H, int ldh, int iloz, int ihiz,
                                          double* Z, int ldz, const MsqrCfg& cfg, const MsqrWork& wk ) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int* lcmd = cluster.map_shared_rank( wk.ibuf + kMsCmd, 0 );
  const double* lrefl = cluster.map_shared_rank( wk.refl, 0 );
  const int nrefl = cfg.nb * cfg.W * kReflStride;
  for ( ;; ) {
    cluster.sync();
    const int cmd = lcmd[0];
    if ( cmd == 0 ) {
      cluster.sync();  // the leader's shared memory stays alive until every helper has read the command
      return;
    }
    const int ws = lcmd[1], wlen = lcmd[2], l = lcmd[3], i = lcmd[4], nbulges = lcmd[5];
    if ( cmd == 2 ) {
      const double* lU = cluster.map_shared_rank( wk.U, 0 );
      for ( int q = threadIdx.x; q < wlen * wk.ldw; q += blockDim.x ) wk.U[q] = lU[q];
      __syncthreads();
      ms_update_strips( g, wantt, wantz, n, l, i, ws, wlen, H, ldh, iloz, ihiz, Z, ldz, cfg, wk, 0, wk.U, wk.crank, wk.csize );
    } else {
      if ( wk.gscratch != nullptr ) {
        for ( int q = threadIdx.x; q < nrefl / 2; q += blockDim.x )
          reinterpret_cast<double2*>( wk.refl )[q] = __ldcg( reinterpret_cast<const double2*>( wk.gscratch ) + q );
      } else {
        for ( int q = threadIdx.x; q < nrefl / 2; q += blockDim.x )
          reinterpret_cast<double2*>( wk.refl )[q] = reinterpret_cast<const double2*>( lrefl )[q];
      }
      if ( cmd == 3 ) {
        const int* lk = cluster.map_shared_rank( wk.ibuf + kMsMaxBulges, 0 );
        if ( threadIdx.x < 2 * kMsMaxBulges ) wk.ibuf[kMsMaxBulges + threadIdx.x] = lk[threadIdx.x];
      }
      if ( threadIdx.x == 0 ) wk.ibuf[3 * kMsMaxBulges + 6] = 0;
      __syncthreads();
      if ( cmd == 1 )
        ms_strips_regs( wantt, wantz, n, l, i, ws, wlen, H, ldh, iloz, ihiz, Z, ldz, wk.refl, wk.tile, wk.ldt,
                        wk.ibuf + 3 * kMsMaxBulges + 6, wk.crank, wk.csize, wk.csize > 2 && wk.gscratch != nullptr && wk.overlap != 0 );
      else
        ms_update_strips( g, wantt, wantz, n, l, i, ws, wlen, H, ldh, iloz, ihiz, Z, ldz, cfg, wk, nbulges, (const double*)nullptr,
                          wk.crank, wk.csize );
    }
    cluster.sync();
  }
}
// leader: releases the helpers for good
__device__ inline void ms_cluster_release( const MsqrWork& wk ) {
  if ( wk.csize <= 1 ) return;
  __syncthreads();
  if ( threadIdx.x == 0 ) wk.ibuf[kMsCmd] = 0;
  cooperative_groups::this_cluster().sync();
  cooperative_groups::this_cluster().sync();
}
#endif

// ------------------------------------------------------------------------------------------------
// Driver. H upper Hessenberg in rows/cols ilo..ihi (isolated eigenvalues outside are the caller's
// business, as in pb::lahqr). Returns 0 or LAPACK's info (i+1: eigenvalues i+1..ihi converged).
// ------------------------------------------------------------------------------------------------
template <class G>
PB_D int hqr_multishift( const G& g, bool wantt, bool wantz, int n, int ilo, int ihi, double* H, int ldh, double* wr,
                         double* wi, int iloz, int ihiz, double* Z, int ldz, const MsqrCfg& cfg, const MsqrWork& wk ) {
  const int t = g.tid(), nt = g.size();
  if ( n == 0 || ihi < ilo ) return 0;
  // clear the band below the subdiagonal that parked bulges use
  for ( int j = ilo + t; j <= ihi - 2; j += nt ) {
    PB_H( j + 2, j ) = 0.0;
    if ( j + 3 <= ihi ) PB_H( j + 3, j ) = 0.0;
  }
  g.sync();
  const int nh = ihi - ilo + 1;
  const double smlnum = kSafmin * ( (double)nh / kUlp );
  const int itmax = cfg.itmax > 0 ? cfg.itmax : 30 * imax( 10, nh );
  int kbot = ihi;
  int it = 0, nodefl = 0;
  while ( kbot >= ilo ) {
    const int ktop = ms_find_split( g, ilo, ihi, ilo, kbot, H, ldh, smlnum );
    if ( ktop > ilo ) {
      if ( t == 0 ) PB_H( ktop, ktop - 1 ) = 0.0;
    }
    g.sync();
    const int m = kbot - ktop + 1;
    PB_CHECK_UNIFORM( g, ktop, "ktop" );
    g.tick( kTickMsScan );
    if ( m <= cfg.ws_small ) {
      const int info = ms_small_solve( g, wantt, wantz, n, ktop, kbot, H, ldh, wr, wi, iloz, ihiz, Z, ldz, cfg, wk );
      PB_CHECK_UNIFORM( g, info, "small-solve info" );
      if ( info != 0 ) return info;
      kbot = ktop - 1;
      nodefl = 0;
      continue;
    }
    if ( ++it > itmax ) return kbot + 1;
    // the shift block is factored inside the U buffer, hence ns <= W
    const int nshift_cfg = imin( cfg.nshift > 0 ? cfg.nshift : 2 * cfg.nb * cfg.nchains, 2 * cfg.nb * cfg.nchains );
    int ns_aed = 0, ns_spec = 0, spec_want = 0;
    if ( cfg.aed > 0 && cfg.aed_moves == 0 && cfg.spec_shifts != 0 && ( ( nodefl + 1 ) % 6 ) != 0 ) {
      const int want0 = imax( 2, imin( imin( nshift_cfg, cfg.W & ~1 ), ( ( m - 1 ) / 3 ) * 2 ) & ~1 );
      if ( cfg.spec_shifts == 1 ) {
        ns_spec = ms_compute_shifts( g, ktop, kbot, H, ldh, want0, false, wk );
        g.tick( kTickMsShifts );
      } else {
        spec_want = imin( want0, m );  // side by side with the AED window (ms_aed)
      }
    }
    if ( cfg.aed > 0 ) {
      const int jw = imin( imin( cfg.aed, cfg.W - 1 ), m );
      int nd = 0;
      ms_aed( g, wantt, wantz, n, ktop, kbot, jw, nshift_cfg, H, ldh, wr, wi, iloz, ihiz, Z, ldz, cfg, wk, &nd, &ns_aed, spec_want, &ns_spec );
      PB_CHECK_UNIFORM( g, nd * 100 + ns_aed, "aed nd/ns" );
      g.tick( kTickMsAedU );
      if ( nd > 0 ) {
        kbot -= nd;
        nodefl = 0;
        // enough deflations (LAPACK's nibble = 14 %), or what is left is a small block: no sweep this time
        if ( 100 * nd > cfg.nibble * jw || kbot - ktop + 1 <= cfg.ws_small ) {
          if ( t == 0 ) wk.ibuf[3 * kMsMaxBulges + 13] += 1;
          continue;
        }
      }
    }
    const int mm = kbot - ktop + 1;
    if ( t == 0 ) wk.ibuf[3 * kMsMaxBulges + 1] += 1;
    nodefl += 1;
    const int nshift_want = imax( 2, imin( imin( nshift_cfg, cfg.W & ~1 ), ( ( mm - 1 ) / 3 ) * 2 ) & ~1 );
    int ns;
    const bool exceptional = ( nodefl % 6 ) == 0;
    if ( !exceptional && ns_aed >= imax( 2, nshift_want / 2 ) ) {
      // AED left enough shift candidates in wk.sr / wk.si (the smallest moduli are at the end)
      ns = imin( ns_aed, nshift_want );
      if ( ns < ns_aed ) {
        g.sync();
        if ( t == 0 )
          for ( int q = 0; q < ns; ++q ) {
            wk.sr[q] = wk.sr[ns_aed - ns + q];
            wk.si[q] = wk.si[ns_aed - ns + q];
          }
        g.sync();
      }
    } else if ( !exceptional && ns_spec >= 2 ) {
      // speculative shifts (eigenvalues of the trailing block before the AED step): keep the smallest moduli
      ns = imin( ns_spec, nshift_want );
      if ( ns < ns_spec ) {
        g.sync();
        if ( t == 0 )
          for ( int q = 0; q < ns; ++q ) {
            wk.sr[q] = wk.sr[ns_spec - ns + q];
            wk.si[q] = wk.si[ns_spec - ns + q];
          }
        g.sync();
      }
    } else {
      ns = ms_compute_shifts( g, ktop, kbot, H, ldh, nshift_want, exceptional, wk );
    }
    PB_CHECK_UNIFORM( g, ns, "ns" );
    g.tick( kTickMsShifts );
    // chains of nb bulges; with fewer computed shifts than chains * 2 nb the set is reused cyclically
    const int nchain_run = ( nshift_cfg < 2 * cfg.nb * cfg.nchains && ns >= 2 * cfg.nb ) ? cfg.nchains : ( ns + 2 * cfg.nb - 1 ) / ( 2 * cfg.nb );
    for ( int ch = 0; ch < nchain_run; ++ch ) {
      const int s0 = ( ch * 2 * cfg.nb ) % ns;
      const int nbul = imin( cfg.nb, ( ns - s0 ) / 2 );
      if ( nbul <= 0 ) continue;
      ms_chase_chain( g, wantt, wantz, n, ktop, kbot, H, ldh, iloz, ihiz, Z, ldz, cfg, wk, nbul, wk.sr + s0, wk.si + s0 );
    }
    // a sweep that exposes a split anywhere in the block resets the exceptional-shift counter
    const int split = ms_find_split( g, ilo, ihi, ktop, kbot, H, ldh, smlnum );
    if ( split > ktop ) nodefl = 0;
  }
  return 0;
}

}  // namespace pb
