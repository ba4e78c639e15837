// hostsim.cpp -- TEST INFRASTRUCTURE ONLY. Never linked into libpaladin_gpu.so.
//
// Instantiates the package's group-templated device code (paladin_b200/csrc/*.cuh) on the host so
// that the numerical LOGIC of the kernels -- deflation tests, shift strategy, index arithmetic,
// conventions -- can be compared with LAPACK in the GPU-less build container:
//   hs_*  : pb::HostGroup, one thread, barriers are no-ops (fast; logic only)
//   ht_*  : ThreadGroup, nt real threads with a pthread barrier behind sync(); built a second time
//           with -fsanitize=thread (libhostsim_tsan.so) it reports every missing barrier as a data
//           race, which is how the kernels' synchronisation is checked without a GPU.
#include <dlfcn.h>
#include <pthread.h>

#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

long g_aed_counts[4];
int g_aed_moves = 1 << 20, g_spec_shifts = 0;
#define PB_AED_COUNT( what ) ( g.tid() == 0 ? (void)++g_aed_counts[what] : (void)0 )
#define PB_CHECK_UNIFORM( g, value, tag ) ( g ).check_uniform( (long)( value ), tag )
#include "../../paladin_b200/csrc/pb_geev.cuh"

namespace {

// Barrier with divergence diagnostics: every thread publishes the call site of its sync(); when some
// thread fails to arrive within 20 s the sites are printed (resolve them with addr2line -e
// libhostsim.so) and the process aborts instead of hanging. (Sites are not compared on arrival: the
// compiler duplicates code paths, so equal source lines can have different return addresses.)
struct ThreadShared {
  pthread_mutex_t mu;
  pthread_cond_t cv;
  int nt, arrived = 0;
  unsigned long gen = 0;
  std::vector<void*> site;
  std::vector<double> d;
  std::vector<int> i;
  explicit ThreadShared( int n ) : nt( n ), site( n, nullptr ), d( n ), i( n ) {
    pthread_mutex_init( &mu, nullptr );
    pthread_cond_init( &cv, nullptr );
  }
  ~ThreadShared() {
    pthread_cond_destroy( &cv );
    pthread_mutex_destroy( &mu );
  }
  void report( const char* why ) {
    std::fprintf( stderr, "hostsim barrier: %s\n", why );
    for ( int k = 0; k < nt; ++k ) {
      Dl_info inf;
      unsigned long off = 0;
      if ( dladdr( site[k], &inf ) && inf.dli_fbase ) off = (unsigned long)( (char*)site[k] - (char*)inf.dli_fbase );
      std::fprintf( stderr, "  thread %d last sync site offset 0x%lx\n", k, off );
    }
    std::abort();
  }
  void wait( int me, void* where ) {
    pthread_mutex_lock( &mu );
    site[me] = where;
    const unsigned long my_gen = gen;
    if ( ++arrived == nt ) {
      arrived = 0;
      ++gen;
      pthread_cond_broadcast( &cv );
    } else {
      timespec ts;
      clock_gettime( CLOCK_REALTIME, &ts );
      ts.tv_sec += 20;
      while ( gen == my_gen ) {
        if ( pthread_cond_timedwait( &cv, &mu, &ts ) != 0 && gen == my_gen ) report( "a thread never arrived (divergent control flow)" );
      }
    }
    pthread_mutex_unlock( &mu );
  }
};

struct ThreadGroup {
  static constexpr bool kIsCta = false;
  ThreadShared* s;
  int me;
  int tid() const { return me; }
  int size() const { return s->nt; }
  __attribute__( ( noinline ) ) void sync() const { s->wait( me, __builtin_return_address( 0 ) ); }
  void tick( int ) const {}
  template <class F>
  double reduce_d( double x, F f ) const {
    s->d[me] = x;
    sync();
    double r = s->d[0];
    for ( int k = 1; k < s->nt; ++k ) r = f( r, s->d[k] );
    sync();
    return r;
  }
  template <class F>
  int reduce_i( int x, F f ) const {
    s->i[me] = x;
    sync();
    int r = s->i[0];
    for ( int k = 1; k < s->nt; ++k ) r = f( r, s->i[k] );
    sync();
    return r;
  }
  double sum( double x ) const {
    return reduce_d( x, []( double a, double b ) { return a + b; } );
  }
  double max( double x ) const {
    return reduce_d( x, []( double a, double b ) { return std::fmax( a, b ); } );
  }
  int imax_( int x ) const {
    return reduce_i( x, []( int a, int b ) { return a > b ? a : b; } );
  }
  int imin_( int x ) const {
    return reduce_i( x, []( int a, int b ) { return a < b ? a : b; } );
  }
  int isum( int x ) const {
    return reduce_i( x, []( int a, int b ) { return a + b; } );
  }
  bool any( bool x ) const { return imax_( x ? 1 : 0 ) != 0; }
  void check_uniform( long v, const char* tag ) const {
    if ( const char* only = std::getenv( "PB_ONLY_CHECK" ) )
      if ( std::strcmp( only, tag ) != 0 ) return;
    const int lo = imin_( (int)v ), hi = imax_( (int)v );
    if ( lo != hi ) {
      std::fprintf( stderr, "hostsim: value '%s' differs between threads (%d..%d)\n", tag, lo, hi );
      std::abort();
    }
  }
};

// run f(group) on nt threads; returns thread 0's result
int run_threads( int nt, const std::function<int( const ThreadGroup& )>& f ) {
  ThreadShared sh( nt );
  std::vector<int> res( nt, 0 );
  std::vector<std::thread> th;
  for ( int k = 0; k < nt; ++k ) th.emplace_back( [&, k] { res[k] = f( ThreadGroup{ &sh, k } ); } );
  for ( auto& x : th ) x.join();
  return res[0];
}

}  // namespace

extern "C" {

void hs_gebal( int n, double* A, int* ilo, int* ihi, double* scale ) {
  pb::HostGroup g;
  std::vector<int> icnt( 2 * n + 2 );
  pb::gebal( g, n, A, n, ilo, ihi, scale, icnt.data() );
}
// with the 4 n doubles of work: precomputed row / column sums (used from 128 active rows on)
void hs_gebal_work( int n, double* A, int* ilo, int* ihi, double* scale ) {
  pb::HostGroup g;
  std::vector<int> icnt( 2 * n + 2 );
  std::vector<double> work( 4 * n + 4 );
  pb::gebal( g, n, A, n, ilo, ihi, scale, icnt.data(), work.data() );
}
int ht_gebal_work( int nt, int n, double* A, int* ilo, int* ihi, double* scale ) {
  std::vector<int> icnt( 2 * n + 2 );
  std::vector<double> work( 4 * n + 4 );
  return run_threads( nt, [&]( const ThreadGroup& g ) {
    int lo = 0, hi = 0;
    pb::gebal( g, n, A, n, &lo, &hi, scale, icnt.data(), work.data() );
    if ( g.tid() == 0 ) {
      *ilo = lo;
      *ihi = hi;
    }
    return 0;
  } );
}

void hs_gehd2( int n, int ilo, int ihi, double* A, double* tau ) {
  pb::HostGroup g;
  std::vector<double> wrk( n + 1 );
  pb::gehd2( g, n, ilo, ihi, A, n, tau, wrk.data() );
}

void hs_orghr( int n, int ilo, int ihi, const double* A, const double* tau, double* Q ) {
  pb::HostGroup g;
  pb::orghr( g, n, ilo, ihi, A, n, tau, Q, n );
}

int hs_lahqr( int wantt, int wantz, int n, int ilo, int ihi, double* H, double* wr, double* wi, double* Z ) {
  pb::HostGroup g;
  return pb::lahqr( g, wantt != 0, wantz != 0, n, ilo, ihi, H, n, wr, wi, 0, n - 1, Z, n );
}

void hs_trevc_right( int n, const double* T, double* V ) {
  pb::HostGroup g;
  std::vector<double> w( 3 * n + 3 );
  pb::trevc_right_inplace( g, n, T, n, V, n, w.data(), w.data() + n, w.data() + 2 * n );
}

void hs_trevc_left( int n, const double* T, double* V ) {
  pb::HostGroup g;
  std::vector<double> w( 3 * n + 3 );
  pb::trevc_left_inplace( g, n, T, n, V, n, w.data(), w.data() + n, w.data() + 2 * n );
}

void hs_gebak_left( int n, int ilo, int ihi, const double* scale, double* V ) {
  pb::HostGroup g;
  pb::gebak_vectors( g, true, n, ilo, ihi, scale, n, V, n );
}

void hs_gebak_right( int n, int ilo, int ihi, const double* scale, double* V ) {
  pb::HostGroup g;
  pb::gebak_right( g, n, ilo, ihi, scale, n, V, n );
}

int hs_dlaln2( int ltrans, int na, int nw, double smin, const double* a, const double* b, double wr, double wi,
               double* x, double* scale, double* xnorm ) {
  return pb::dlaln2( ltrans != 0, na, nw, smin, a, b, wr, wi, x, *scale, *xnorm );
}

int hs_geev_lr( int n, double* A, double* wr, double* wi, int wantvl, double* VL, int wantvr, double* VR ) {
  pb::HostGroup g;
  std::vector<double> dw( 5 * n + 5 );
  std::vector<int> iw( 2 * n + 2 );
  return pb::geev_group( g, n, A, n, wr, wi, wantvl != 0, VL, n, wantvr != 0, VR, n, dw.data(), iw.data() );
}

int ht_geev_lr( int nt, int n, double* A, double* wr, double* wi, int wantvl, double* VL, int wantvr, double* VR ) {
  std::vector<double> dw( 5 * n + 5 );
  std::vector<int> iw( 2 * n + 2 );
  return run_threads( nt, [&]( const ThreadGroup& g ) {
    return pb::geev_group( g, n, A, n, wr, wi, wantvl != 0, VL, n, wantvr != 0, VR, n, dw.data(), iw.data() );
  } );
}

// the lane-explicit warp code (pb_small.cuh::francis_sweep_lanes) with one host thread walking the 32 lanes
struct HostWarpGroup : pb::HostGroup {
  static constexpr int kLanes = 32;
};
int hs_geev_warp( int n, double* A, double* wr, double* wi, int wantvr, double* V ) {
  HostWarpGroup g;
  std::vector<double> dw( 5 * n + 5 );
  std::vector<int> iw( 2 * n + 2 );
  return pb::geev_group( g, n, A, n, wr, wi, false, (double*)nullptr, n, wantvr != 0, V, n, dw.data(), iw.data() );
}
int hs_lahqr_warp( int wantt, int wantz, int n, int ilo, int ihi, double* H, double* wr, double* wi, double* Z ) {
  HostWarpGroup g;
  return pb::lahqr<HostWarpGroup, true>( g, wantt != 0, wantz != 0, n, ilo, ihi, H, n, wr, wi, 0, n - 1, Z, n );
}

// two upper Hessenberg matrices in one n x (n + 3) array (paladin_gpu.cu::small_qr_pair_kernel): matrix 0 at cell (r, c + 3),
// matrix 1 transposed at cell (c, r); eigenvalues only, first matrix 0 to the end, then matrix 1 (whose cells must have survived)
int hs_lahqr_pair( int n, const double* H0, const double* H1, double* wr0, double* wi0, double* wr1, double* wi1 ) {
  HostWarpGroup g;
  const int lds = n | 1;
  std::vector<double> S( (size_t)lds * ( n + 3 ), 1e300 );  // (a cell nobody owns must never be read: poison)
  for ( int c = 0; c < n; ++c )
    for ( int r = 0; r < n && r <= c + 1; ++r ) {
      S[r + (size_t)( c + 3 ) * lds] = H0[r + (size_t)c * n];
      S[c + (size_t)r * lds] = H1[r + (size_t)c * n];
    }
  const int i0 = pb::lahqr<HostWarpGroup, true, pb::LahqrNoHook, true>( g, false, false, n, 0, n - 1, S.data() + 3 * (size_t)lds, lds, wr0, wi0, 0, n - 1, (double*)nullptr, 1,
                                                 pb::LahqrNoHook(), 1, true );
  const int i1 = pb::lahqr<HostWarpGroup, true, pb::LahqrNoHook, true>( g, false, false, n, 0, n - 1, S.data(), 1, wr1, wi1, 0, n - 1, (double*)nullptr, 1, pb::LahqrNoHook(), lds,
                                                 true );
  return i0 * 1000 + i1;
}

int hs_geev( int n, double* A, double* wr, double* wi, int wantvr, double* V ) {
  pb::HostGroup g;
  std::vector<double> dw( 5 * n + 5 );
  std::vector<int> iw( 2 * n + 2 );
  return pb::geev_group( g, n, A, n, wr, wi, false, (double*)nullptr, n, wantvr != 0, V, n, dw.data(), iw.data() );
}

struct MsBuffers {
  std::vector<double> d;
  std::vector<int> i;
  pb::MsqrWork wk;
  MsBuffers( const pb::MsqrCfg& c ) {
    const int ldw = c.W | 1, ldt = c.tl | 1;
    d.assign( pb::msqr_work_doubles( c, ldw, ldt ) + 64, 0.0 );
    i.assign( pb::kMsIbufInts, 0 );
    double* p = d.data();
    wk.Hw = p; p += pb::ms_even( (size_t)c.W * ldw );
    wk.U = p; p += pb::ms_even( (size_t)c.W * ldw );
    wk.tile = p; p += pb::ms_even( (size_t)c.W * ldt );
    wk.refl = p; p += (size_t)c.nb * c.W * pb::kReflStride;
    wk.sr = p; p += 2 * c.nb * c.nchains;
    wk.si = p;
    wk.ibuf = i.data();
    wk.ldw = ldw;
    wk.ldt = ldt;
  }
};

// block reordering of the AED step (pb_aed.cuh): move the block at ifst up to ilst, Q accumulated
int hs_trexc_up( int n, double* T, double* Q, int ifst, int ilst ) {
  pb::HostGroup g;
  return pb::aed_trexc_up( g, n, T, n, Q, n, n, ifst, ilst );
}

int g_ms_counters[16];
int hs_msqr( int wantt, int wantz, int n, int ilo, int ihi, double* H, double* wr, double* wi, double* Z, int nb, int W,
             int ws_small, int tl, int nchains ) {
  pb::HostGroup g;
  pb::MsqrCfg c{ nb, W, ws_small, tl, nchains % 100, ( nchains / 100 ) % 100, nchains / 10000, g_aed_moves, 14, g_spec_shifts };
  MsBuffers buf( c );
  const int r = pb::hqr_multishift( g, wantt != 0, wantz != 0, n, ilo, ihi, H, n, wr, wi, 0, n - 1, Z, n, c, buf.wk );
  for ( int k = 0; k < 16; ++k ) g_ms_counters[k] = buf.i[3 * pb::kMsMaxBulges + k];
  return r;
}
long* hs_aed_counts() { return g_aed_counts; }
void hs_aed_moves( int v ) { g_aed_moves = v; }
void hs_spec_shifts( int v ) { g_spec_shifts = v; }
const int* hs_ms_counters() { return g_ms_counters; }

int ht_msqr( int nt, int wantt, int wantz, int n, int ilo, int ihi, double* H, double* wr, double* wi, double* Z, int nb,
             int W, int ws_small, int tl, int nchains ) {
  pb::MsqrCfg c{ nb, W, ws_small, tl, nchains % 100, ( nchains / 100 ) % 100, nchains / 10000, g_aed_moves, 14, g_spec_shifts };
  MsBuffers buf( c );
  return run_threads( nt, [&]( const ThreadGroup& g ) {
    return pb::hqr_multishift( g, wantt != 0, wantz != 0, n, ilo, ihi, H, n, wr, wi, 0, n - 1, Z, n, c, buf.wk );
  } );
}

int hs_geev_ms( int n, double* A, double* wr, double* wi, int wantvr, double* V, int nb, int W, int ws_small, int tl,
                int nchains ) {
  pb::HostGroup g;
  std::vector<double> dw( 5 * n + 5 );
  std::vector<int> iw( 2 * n + 2 );
  pb::MsqrCfg c{ nb, W, ws_small, tl, nchains % 100, ( nchains / 100 ) % 100, nchains / 10000, g_aed_moves, 14, g_spec_shifts };
  MsBuffers buf( c );
  return pb::geev_group( g, n, A, n, wr, wi, false, (double*)nullptr, n, wantvr != 0, V, n, dw.data(), iw.data(), pb::QrMultishift{ c, buf.wk } );
}

int ht_geev_ms( int nt, int n, double* A, double* wr, double* wi, int wantvr, double* V, int nb, int W, int ws_small, int tl,
                int nchains ) {
  std::vector<double> dw( 5 * n + 5 );
  std::vector<int> iw( 2 * n + 2 );
  pb::MsqrCfg c{ nb, W, ws_small, tl, nchains % 100, ( nchains / 100 ) % 100, nchains / 10000, g_aed_moves, 14, g_spec_shifts };
  MsBuffers buf( c );
  return run_threads( nt, [&]( const ThreadGroup& g ) {
    return pb::geev_group( g, n, A, n, wr, wi, false, (double*)nullptr, n, wantvr != 0, V, n, dw.data(), iw.data(), pb::QrMultishift{ c, buf.wk } );
  } );
}

// ---- threaded versions (race checking under -fsanitize=thread) ----
int ht_geev( int nt, int n, double* A, double* wr, double* wi, int wantvr, double* V ) {
  std::vector<double> dw( 5 * n + 5 );
  std::vector<int> iw( 2 * n + 2 );
  return run_threads( nt, [&]( const ThreadGroup& g ) {
    return pb::geev_group( g, n, A, n, wr, wi, false, (double*)nullptr, n, wantvr != 0, V, n, dw.data(), iw.data() );
  } );
}

}  // extern "C"
