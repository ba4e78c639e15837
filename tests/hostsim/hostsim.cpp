// hostsim.cpp -- TEST INFRASTRUCTURE ONLY. Never linked into libpaladin_gpu.so.
//
// Instantiates the package's group-templated device code (paladin_b200/csrc/*.cuh) on the host so
// that the numerical LOGIC of the kernels -- deflation tests, shift strategy, index arithmetic,
// conventions -- can be compared with LAPACK in the GPU-less build container:
//   hs_*  : pb::HostGroup, one thread, barriers are no-ops (fast; logic only)
//   ht_*  : ThreadGroup, nt real threads with a pthread barrier behind sync(); built a second time
//           with -fsanitize=thread (libhostsim_tsan.so) it reports every missing barrier as a data
//           race, which is how the kernels' synchronisation is checked without a GPU.
#include <pthread.h>

#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "../../paladin_b200/csrc/pb_geev.cuh"

namespace {

struct ThreadShared {
  pthread_barrier_t bar;
  int nt;
  std::vector<double> d;
  std::vector<int> i;
  explicit ThreadShared( int n ) : nt( n ), d( n ), i( n ) { pthread_barrier_init( &bar, nullptr, n ); }
  ~ThreadShared() { pthread_barrier_destroy( &bar ); }
};

struct ThreadGroup {
  ThreadShared* s;
  int me;
  int tid() const { return me; }
  int size() const { return s->nt; }
  void sync() const { pthread_barrier_wait( &s->bar ); }
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

void hs_gebak_right( int n, int ilo, int ihi, const double* scale, double* V ) {
  pb::HostGroup g;
  pb::gebak_right( g, n, ilo, ihi, scale, n, V, n );
}

int hs_dlaln2( int ltrans, int na, int nw, double smin, const double* a, const double* b, double wr, double wi,
               double* x, double* scale, double* xnorm ) {
  return pb::dlaln2( ltrans != 0, na, nw, smin, a, b, wr, wi, x, *scale, *xnorm );
}

int hs_geev( int n, double* A, double* wr, double* wi, int wantvr, double* V ) {
  pb::HostGroup g;
  std::vector<double> dw( 5 * n + 5 );
  std::vector<int> iw( 2 * n + 2 );
  return pb::geev_group( g, n, A, n, wr, wi, wantvr != 0, V, n, dw.data(), iw.data() );
}

// ---- threaded versions (race checking under -fsanitize=thread) ----
int ht_geev( int nt, int n, double* A, double* wr, double* wi, int wantvr, double* V ) {
  std::vector<double> dw( 5 * n + 5 );
  std::vector<int> iw( 2 * n + 2 );
  return run_threads( nt, [&]( const ThreadGroup& g ) {
    return pb::geev_group( g, n, A, n, wr, wi, wantvr != 0, V, n, dw.data(), iw.data() );
  } );
}

}  // extern "C"
