// pb_strips.cuh -- off-window strip updates of the multishift QR with every line held in registers
// (device only). Fast path of pb::ms_update_strips (pb_msqr.cuh) for the recorded-reflector mode.
//
// A "line" is a column of H right of the window, a row of H above it, or a row of the Schur vectors Z;
// inside the window's index range it has W entries. One thread owns one line: it loads the W entries into
// registers, streams every recorded 3-element reflector of the window over them (5 FMAs each, operands at
// compile-time register indices, reflector parameters broadcast from shared memory with 16-byte loads) and
// stores the line once. Warps work on batches of 32 lines independently: there is no CTA barrier inside the
// strip phase.
//   rows (upper strip, Z): lane = row, so every global access of a warp is one contiguous 256-byte segment;
//   columns (right strip): a column's W entries are contiguous in memory; the warp reads them with lanes
//     along the rows (coalesced) and transposes through its private 32-line slice of the shared tile.
// Each bulge's reflectors act on consecutive positions p, p+1, ..., so the unrolled step list is entered at
// position p through a jump table and left after the bulge's last step (uniform branches only).
#pragma once

#include "pb_common.cuh"

#if defined( __CUDACC__ )

namespace pb {

// x[P], x[P+1], x[P+2] <- reflector at rf (w2,w3 | t1,t2 | t3,pad as three 16-byte words)
#define PB_STRIP_STEP( P )                                              \
  case P: {                                                             \
    const double2 w = rf[0], tt = rf[1];                                \
    const double t3 = reinterpret_cast<const double*>( rf )[4];         \
    rf += 3;                                                            \
    const double sum = x[P] + w.x * x[P + 1] + w.y * x[P + 2];          \
    x[P] -= sum * tt.x;                                                 \
    x[P + 1] -= sum * tt.y;                                             \
    x[P + 2] -= sum * t3;                                               \
    if ( --cnt == 0 ) break;                                            \
  }

#define PB_STRIP_STEPS_8( B ) \
  PB_STRIP_STEP( B + 0 ) PB_STRIP_STEP( B + 1 ) PB_STRIP_STEP( B + 2 ) PB_STRIP_STEP( B + 3 ) \
  PB_STRIP_STEP( B + 4 ) PB_STRIP_STEP( B + 5 ) PB_STRIP_STEP( B + 6 ) PB_STRIP_STEP( B + 7 )

// Window order of the device path (compile time: the line registers are indexed statically). PB_WINDOW = 49 keeps
// one CTA per SM (49 + 2 doubles of line state per thread); 41 leaves room for two CTAs per SM.
// positions 0..W-1, steps may start at 0..W-2 (x[W], x[W+1] are padding that only a final 2-row reflector with
// w3 = t3 = 0 touches)
#ifndef PB_WINDOW
#define PB_WINDOW 49
#endif
constexpr int kStripW = PB_WINDOW;
static_assert( kStripW == 41 || kStripW == 49, "the step table below is written for W = 41 or 49" );

__device__ __forceinline__ void strip_apply_line( double ( &x )[kStripW + 2], int nbulges, const int* kfirst, const int* kcount,
                                                  int ws, const double* refl ) {
  for ( int b = 0; b < nbulges; ++b ) {
    int cnt = kcount[b];
    if ( cnt == 0 ) continue;
    const int p = kfirst[b] - ws;
    const double2* rf = reinterpret_cast<const double2*>( refl + (size_t)b * kStripW * 6 );
    switch ( p ) {
      PB_STRIP_STEPS_8( 0 )
      PB_STRIP_STEPS_8( 8 )
      PB_STRIP_STEPS_8( 16 )
      PB_STRIP_STEPS_8( 24 )
      PB_STRIP_STEPS_8( 32 )
#if PB_WINDOW == 49
      PB_STRIP_STEPS_8( 40 )
#endif
      default:
        break;
    }
  }
}

// Steady-state window: all kStripNB bulges enter tightly packed (bulge b at position 1 + 3 (NB-1-b)) and advance
// kStripD rows in lockstep. Every position is then known at compile time, and the steps are issued macro-step by
// macro-step: the NB reflectors of one macro-step touch disjoint positions, so a thread has NB independent
// 5-FMA groups in flight instead of one serial chain (the bulge-major order of the generic path is equivalent:
// a step of bulge b only depends on EARLIER macro-steps of the bulges ahead of it).
constexpr int kStripNB = 8, kStripD = kStripW - 1 - 3 * kStripNB;

__device__ __forceinline__ bool strip_window_is_regular( int nbulges, int wlen, int ws, const int* kfirst, const int* kcount ) {
  if ( nbulges != kStripNB || wlen != kStripW ) return false;
  bool ok = true;
#pragma unroll
  for ( int b = 0; b < kStripNB; ++b ) ok = ok && ( kfirst[b] - ws == 1 + 3 * ( kStripNB - 1 - b ) ) && ( kcount[b] == kStripD );
  return ok;
}

__device__ __forceinline__ void strip_apply_line_regular( double ( &x )[kStripW + 2], const double* refl ) {
#pragma unroll
  for ( int t = 0; t < kStripD; ++t ) {
#pragma unroll
    for ( int b = 0; b < kStripNB; ++b ) {
      const int P = 1 + 3 * ( kStripNB - 1 - b ) + t;
      const double2* rf = reinterpret_cast<const double2*>( refl + ( (size_t)b * kStripW + t ) * 6 );
      const double2 w = rf[0], tt = rf[1];
      const double t3 = reinterpret_cast<const double*>( rf )[4];
      const double sum = x[P] + w.x * x[P + 1] + w.y * x[P + 2];
      x[P] -= sum * tt.x;
      x[P + 1] -= sum * tt.y;
      x[P + 2] -= sum * t3;
    }
  }
}

// Streaming form of the steady-state update: at macro-step t only positions 1+t .. 24+t are live (the trailing bulge
// has left everything below, the leading one has not reached anything above), so a line never needs more than
// 3*NB + kStripPF positions in registers: position 24+PF+1+t is fetched while macro-step t runs and position 1+t is
// stored right after it. Positions 0 and W-1 are never touched by a steady-state window and are neither loaded
// nor stored. LD(p) / ST(p, v) access position p of the thread's line.
#ifndef PB_STRIP_PF
#define PB_STRIP_PF 10
#endif
constexpr int kStripPF = PB_STRIP_PF;

template <class LD, class ST>
__device__ __forceinline__ void strip_regular_stream( LD ld, ST st, const double* refl ) {
  constexpr int kTop = 3 * kStripNB;  // highest position touched by macro-step 0
  double x[kStripW + 2];
#pragma unroll
  for ( int p = 1; p <= kTop + kStripPF; ++p )
    if ( p <= kStripW - 2 ) x[p] = ld( p );
#pragma unroll
  for ( int t = 0; t < kStripD; ++t ) {
    const int pn = kTop + kStripPF + 1 + t;
    if ( pn <= kStripW - 2 ) x[pn] = ld( pn );
#pragma unroll
    for ( int b = 0; b < kStripNB; ++b ) {
      const int P = 1 + 3 * ( kStripNB - 1 - b ) + t;
      const double2* rf = reinterpret_cast<const double2*>( refl + ( (size_t)b * kStripW + t ) * 6 );
      const double2 w = rf[0], tt = rf[1];
      const double t3 = reinterpret_cast<const double*>( rf )[4];
      const double sum = x[P] + w.x * x[P + 1] + w.y * x[P + 2];
      x[P] -= sum * tt.x;
      x[P + 1] -= sum * tt.y;
      x[P + 2] -= sum * t3;
    }
    st( 1 + t, x[1 + t] );
  }
#pragma unroll
  for ( int p = 1 + kStripD; p <= kStripW - 2; ++p ) st( p, x[p] );
}

// Every off-window strip of one window move. refl / kfirst / kcount live in shared memory; tile is the
// W x ldt shared tile (ldt >= 16 * warps + 1, odd); the caller synchronises the CTA before and after.
__device__ __forceinline__ void prefetch_l2( const void* p ) { asm volatile( "prefetch.global.L2 [%0];" ::"l"( p ) ); }

// Steady-state windows only. counter: one int of shared memory, zeroed by the caller before the CTA barrier that
// precedes this call (work queue of the row batches).
// crank / csize: this CTA is one of csize (a thread-block cluster, pb_msqr.cuh) that share the strips of the window:
// right-strip batches and row batches are dealt round-robin to the CTAs; counter is CTA-local either way.
__device__ inline void ms_strips_regs( bool wantt, bool wantz, int n, int l, int i, int ws, int wlen, double* __restrict__ H,
                                       int ldh, int iloz, int ihiz, double* __restrict__ Z, int ldz, const double* refl,
                                       double* tile, int ldt, int* counter, int crank = 0, int csize = 1, bool lookahead = false ) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // lookahead (overlapped cluster hand-off): the leader only takes right-strip batches 0 and 1 -- every column the next
  // window can reach -- and leaves the rest of the window's strips to the other csize - 1 CTAs
  const int rfirst = lookahead ? 2 : 0;
  if ( lookahead ) {
    if ( crank == 0 ) {
      csize = 0;  // (marks the leader below)
    } else {
      crank -= 1;
      csize -= 1;
    }
  }
  const int we = ws + wlen - 1;
  const int i1 = wantt ? 0 : l, i2 = wantt ? n - 1 : i;
  const int nright = imax( 0, i2 - we ), nup = imax( 0, ws - i1 ), nz = wantz ? ( ihiz - iloz + 1 ) : 0;
  const int bright = ( nright + 31 ) >> 5, bup = ( nup + 31 ) >> 5, bz = ( nz + 31 ) >> 5;
  // ---- right strip: 32 columns per batch, even warps only. A column's W entries are contiguous in memory: the warp
  // reads them with lanes along the rows (coalesced) and transposes through a 32-line slice of the shared tile
  // (slice w/2 spans the 16-line slices of warps w and w+1; odd warps never use theirs here), then lane = column.
  if ( ( warp & 1 ) == 0 ) {
    double* wt = tile + 32 * ( warp >> 1 );
    const int rstride = csize > 0 ? ( ( nwarps + 1 ) >> 1 ) * csize : ( 1 << 20 );
    const int rlimit = csize > 0 ? bright : imin( bright, 2 );
    for ( int bt = ( csize > 0 ? rfirst + ( warp >> 1 ) + crank * ( ( nwarps + 1 ) >> 1 ) : ( warp >> 1 ) ); bt < rlimit; bt += rstride ) {
      const int j0 = we + 1 + 32 * bt;
      const int nl = imin( 32, i2 - j0 + 1 );
      if ( lane < nl ) {  // L2 prefetch of the next batch of this warp
        const int jn = j0 + 32 * rstride + lane;
        if ( jn <= i2 ) {
          const double* col = H + (size_t)jn * ldh + ws;
#pragma unroll
          for ( int q = 0; q < kStripW; q += 4 ) prefetch_l2( col + q );
        }
      }
      for ( int c0 = 0; c0 < nl; c0 += 8 ) {
        double v0[8], v1[8];
#pragma unroll
        for ( int u = 0; u < 8; ++u ) {
          const int cc = imin( c0 + u, nl - 1 );
          const double* col = H + (size_t)( j0 + cc ) * ldh + ws;
          v0[u] = __ldcg( col + lane );
          v1[u] = __ldcg( col + imin( 32 + lane, kStripW - 1 ) );
        }
#pragma unroll
        for ( int u = 0; u < 8; ++u ) {
          if ( c0 + u < nl ) {
            wt[lane * ldt + c0 + u] = v0[u];
            if ( 32 + lane < kStripW ) wt[( 32 + lane ) * ldt + c0 + u] = v1[u];
          }
        }
      }
      __syncwarp();
      {
        double* mine = wt + lane;
        const bool owner = lane < nl;
        strip_regular_stream( [&]( int p ) { return mine[p * ldt]; },
                              [&]( int p, double v ) {
                                if ( owner ) mine[p * ldt] = v;
                              },
                              refl );
      }
      __syncwarp();
      for ( int cc = 0; cc < nl; ++cc ) {
        double* col = H + (size_t)( j0 + cc ) * ldh + ws;
        __stcg( col + lane, wt[lane * ldt + cc] );
        if ( 32 + lane < kStripW ) __stcg( col + 32 + lane, wt[( 32 + lane ) * ldt + cc] );
      }
      __syncwarp();
    }
  }
  // ---- rows of the strip above the window and of Z: lane = row, every access is a contiguous 256-byte segment;
  // batches are handed out through a shared counter (odd warps start here at once, even warps join after their
  // share of the right strip)
  const int nrow_batches = csize > 0 ? bup + bz : 0;
  auto row_batch = [&]( int bt, double*& M, int& ldm, int& r0, int& nl ) {
    if ( bt < bup ) {
      M = H;
      ldm = ldh;
      r0 = i1 + 32 * bt;
      nl = imin( 32, ws - r0 );
    } else {
      M = Z;
      ldm = ldz;
      r0 = iloz + 32 * ( bt - bup );
      nl = imin( 32, ihiz - r0 + 1 );
    }
  };
  for ( ;; ) {
    int bt = 0;
    if ( lane == 0 ) bt = atomicAdd( counter, 1 );
    bt = crank + csize * __shfl_sync( 0xffffffffu, bt, 0 );
    if ( bt >= nrow_batches ) break;
    double* M;
    int ldm, r0, nl;
    row_batch( bt, M, ldm, r0, nl );
    {  // L2 prefetch of a batch a few positions ahead in the queue
      const int bn = bt + nwarps * csize;
      if ( bn < nrow_batches ) {
        double* Mn;
        int ldn, rn, nn;
        row_batch( bn, Mn, ldn, rn, nn );
        const double* base = Mn + rn + (size_t)ws * ldn;
        for ( int p = 1 + ( lane >> 2 ); p <= kStripW - 2; p += 8 ) prefetch_l2( base + (size_t)p * ldn + 8 * ( lane & 3 ) );
      }
    }
    const int r = r0 + imin( lane, nl - 1 );
    double* row = M + r + (size_t)ws * ldm;
    const bool owner = lane < nl;
    strip_regular_stream( [&]( int p ) { return __ldcg( row + (size_t)p * ldm ); },
                          [&]( int p, double v ) {
                            if ( owner ) __stcg( row + (size_t)p * ldm, v );
                          },
                          refl );
  }
}

}  // namespace pb

#endif  // __CUDACC__
