// pb_common.cuh -- thread-group abstraction + scalar LAPACK-style helpers shared by every kernel.
//
// All numerical device code in this package is written against a "group" G (a warp or a whole CTA)
// that offers tid()/size()/sync()/reductions. The same templates are instantiated
//   * with WarpGroup / CtaGroup by nvcc for sm_100a (the product), and
//   * with HostGroup (one thread, sync = no-op) by g++ in tests/hostsim (a TEST-ONLY logic
//     simulator used in the GPU-less build container; it is never linked into the product library).
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PB_D __device__ __forceinline__
#define PB_DN __device__ __noinline__
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_D inline
#define PB_DN inline
#define PB_HD inline
#endif

// Loads of matrix data that another CTA of a thread-block cluster may have written (pb_msqr.cuh): straight from L2, never
// through the L1 / read-only path (a `const __restrict__` pointer would otherwise allow ld.global.nc).
#if defined(__CUDA_ARCH__)
#define PB_LDCG( ptr ) __ldcg( ptr )
#else
#define PB_LDCG( ptr ) ( *( ptr ) )
#endif

// Uniformity checks of group-wide control values: compiled only into the threaded host simulator
// (tests/hostsim defines PB_CHECK_UNIFORM as a call into its barrier), nothing on the device.
#ifndef PB_CHECK_UNIFORM
#define PB_CHECK_UNIFORM( g, value, tag )
#endif

namespace pb {

// ---- machine constants (IEEE double; same values LAPACK's dlamch returns) -------------------------
constexpr double kEps = 1.1102230246251565e-16;     // dlamch('E') relative machine eps (2^-53)
constexpr double kUlp = 2.2204460492503131e-16;     // dlamch('P') = eps*base
constexpr double kSafmin = 2.2250738585072014e-308;  // dlamch('S')
constexpr double kSafmax = 1.0 / kSafmin;

PB_HD double dsign( double a, double b ) { return b >= 0.0 ? fabs( a ) : -fabs( a ); }  // Fortran SIGN
PB_HD double dmax( double a, double b ) { return a > b ? a : b; }
PB_HD double dmin( double a, double b ) { return a < b ? a : b; }
PB_HD int imin( int a, int b ) { return a < b ? a : b; }
PB_HD int imax( int a, int b ) { return a > b ? a : b; }

// dlapy2: sqrt(x^2+y^2) without unnecessary overflow
PB_HD double dlapy2( double x, double y ) {
  const double xa = fabs( x ), ya = fabs( y );
  const double w = dmax( xa, ya ), z = dmin( xa, ya );
  if ( z == 0.0 || w > DBL_MAX ) return w;
  const double q = z / w;
  return w * sqrt( 1.0 + q * q );
}

// ---- groups ------------------------------------------------------------------------------------
struct HostGroup {
  static constexpr bool kIsCta = false;
  PB_HD int tid() const { return 0; }
  PB_HD int size() const { return 1; }
  PB_HD void sync() const {}
  PB_HD double sum( double x ) const { return x; }
  PB_HD double max( double x ) const { return x; }
  PB_HD int imax_( int x ) const { return x; }
  PB_HD int imin_( int x ) const { return x; }
  PB_HD int isum( int x ) const { return x; }
  PB_HD bool any( bool x ) const { return x; }
  PB_HD void tick( int ) const {}
  PB_HD void check_uniform( long, const char* ) const {}
};

#if defined(__CUDACC__)
PB_D double warp_sum( double x ) {
#pragma unroll
  for ( int o = 16; o > 0; o >>= 1 ) x += __shfl_xor_sync( 0xffffffffu, x, o );
  return x;
}
PB_D double warp_max( double x ) {
#pragma unroll
  for ( int o = 16; o > 0; o >>= 1 ) x = fmax( x, __shfl_xor_sync( 0xffffffffu, x, o ) );
  return x;
}
PB_D int warp_imax( int x ) {
#pragma unroll
  for ( int o = 16; o > 0; o >>= 1 ) x = ::max( x, __shfl_xor_sync( 0xffffffffu, x, o ) );
  return x;
}
PB_D int warp_isum( int x ) {
#pragma unroll
  for ( int o = 16; o > 0; o >>= 1 ) x += __shfl_xor_sync( 0xffffffffu, x, o );
  return x;
}

// one warp; every lane receives bit-identical reduction results (xor butterflies)
struct WarpGroup {
  static constexpr bool kIsCta = false;
  static constexpr int kLanes = 1;  // one lane per thread: lane-explicit code (pb_small.cuh::francis_sweep_lanes) applies
  PB_D int tid() const { return threadIdx.x & 31; }
  PB_D int size() const { return 32; }
  PB_D void sync() const { __syncwarp(); }
  PB_D double sum( double x ) const { return warp_sum( x ); }
  PB_D double max( double x ) const { return warp_max( x ); }
  PB_D int imax_( int x ) const { return warp_imax( x ); }
  PB_D int imin_( int x ) const { return -warp_imax( -x ); }
  PB_D int isum( int x ) const { return warp_isum( x ); }
  PB_D bool any( bool x ) const { return __any_sync( 0xffffffffu, x ); }
  PB_D void tick( int ) const {}
};

// the first nw warps of a CTA, synchronised by named barrier 1 (the CTA's other warps must not use that barrier
// meanwhile); red points at >= 2 * nw doubles of shared scratch. A 32..48-row problem in shared memory offers 48-way
// parallelism in each of its three update loops: three warps run them side by side where one warp needs two passes.
struct WarpsGroup {
  static constexpr bool kIsCta = false;
  int nw;
  double* red;
  PB_D int tid() const { return threadIdx.x; }
  PB_D int size() const { return 32 * nw; }
  PB_D void sync() const { asm volatile( "bar.sync 1, %0;" ::"r"( 32 * nw ) : "memory" ); }
  PB_D double sum( double x ) const {
    x = warp_sum( x );
    sync();
    if ( ( threadIdx.x & 31 ) == 0 ) red[threadIdx.x >> 5] = x;
    sync();
    double t = 0.0;
    for ( int i = 0; i < nw; ++i ) t += red[i];
    return t;
  }
  PB_D double max( double x ) const {
    x = warp_max( x );
    sync();
    if ( ( threadIdx.x & 31 ) == 0 ) red[threadIdx.x >> 5] = x;
    sync();
    double t = red[0];
    for ( int i = 1; i < nw; ++i ) t = fmax( t, red[i] );
    return t;
  }
  PB_D int imax_( int x ) const {
    x = warp_imax( x );
    int* ired = reinterpret_cast<int*>( red );
    sync();
    if ( ( threadIdx.x & 31 ) == 0 ) ired[threadIdx.x >> 5] = x;
    sync();
    int t = ired[0];
    for ( int i = 1; i < nw; ++i ) t = ::max( t, ired[i] );
    return t;
  }
  PB_D int imin_( int x ) const { return -imax_( -x ); }
  PB_D int isum( int x ) const {
    x = warp_isum( x );
    int* ired = reinterpret_cast<int*>( red );
    sync();
    if ( ( threadIdx.x & 31 ) == 0 ) ired[threadIdx.x >> 5] = x;
    sync();
    int t = 0;
    for ( int i = 0; i < nw; ++i ) t += ired[i];
    return t;
  }
  PB_D bool any( bool x ) const { return imax_( x ? 1 : 0 ) != 0; }
  PB_D void tick( int ) const {}
};

// whole thread block; red points at >= 64 doubles of shared scratch owned by the group
struct CtaGroup {
  static constexpr bool kIsCta = true;
  double* red;
  // optional phase profile: thread 0 adds the cycles since the previous tick to stats[slot]
  long long* stats = nullptr;
  mutable long long last = 0;
  PB_D void tick( int slot ) const {
    if ( stats != nullptr && threadIdx.x == 0 ) {
      const long long now = clock64();
      stats[slot] += now - last;
      last = now;
    }
  }
  PB_D int tid() const { return threadIdx.x; }
  PB_D int size() const { return blockDim.x; }
  PB_D void sync() const { __syncthreads(); }
  PB_D double sum( double x ) const {
    x = warp_sum( x );
    const int w = threadIdx.x >> 5, nw = ( blockDim.x + 31 ) >> 5;
    __syncthreads();
    if ( ( threadIdx.x & 31 ) == 0 ) red[w] = x;
    __syncthreads();
    double t = 0.0;
    for ( int i = 0; i < nw; ++i ) t += red[i];
    return t;
  }
  PB_D double max( double x ) const {
    x = warp_max( x );
    const int w = threadIdx.x >> 5, nw = ( blockDim.x + 31 ) >> 5;
    __syncthreads();
    if ( ( threadIdx.x & 31 ) == 0 ) red[w] = x;
    __syncthreads();
    double t = red[0];
    for ( int i = 1; i < nw; ++i ) t = fmax( t, red[i] );
    return t;
  }
  PB_D int imax_( int x ) const {
    x = warp_imax( x );
    int* ired = reinterpret_cast<int*>( red );
    const int w = threadIdx.x >> 5, nw = ( blockDim.x + 31 ) >> 5;
    __syncthreads();
    if ( ( threadIdx.x & 31 ) == 0 ) ired[w] = x;
    __syncthreads();
    int t = ired[0];
    for ( int i = 1; i < nw; ++i ) t = ::max( t, ired[i] );
    return t;
  }
  PB_D int imin_( int x ) const { return -imax_( -x ); }
  PB_D int isum( int x ) const {
    x = warp_isum( x );
    int* ired = reinterpret_cast<int*>( red );
    const int w = threadIdx.x >> 5, nw = ( blockDim.x + 31 ) >> 5;
    __syncthreads();
    if ( ( threadIdx.x & 31 ) == 0 ) ired[w] = x;
    __syncthreads();
    int t = 0;
    for ( int i = 0; i < nw; ++i ) t += ired[i];
    return t;
  }
  PB_D bool any( bool x ) const { return __syncthreads_or( x ? 1 : 0 ) != 0; }
};
#endif  // __CUDACC__

// ---- dlanv2: Schur factorisation of a real 2x2 nonsymmetric matrix in standardised form -----------
// Follows the published LAPACK 3.12 algorithm (third-party to the reference, reached from dgeev via
// dhseqr/dlahqr). Outputs the rotation (cs,sn) and both eigenvalues; a,b,c,d are overwritten.
PB_HD void dlanv2( double& a, double& b, double& c, double& d, double& rt1r, double& rt1i, double& rt2r,
                   double& rt2i, double& cs, double& sn ) {
  const double multpl = 4.0;
  const double eps = kUlp;
  // safmn2 = 2^int(log(safmin/eps)/log(2)/2) = 2^-485
  const double safmn2 = 1.0010415475915505e-146;
  const double safmx2 = 1.0 / safmn2;
  if ( c == 0.0 ) {
    cs = 1.0;
    sn = 0.0;
  } else if ( b == 0.0 ) {
    cs = 0.0;
    sn = 1.0;
    const double temp = d;
    d = a;
    a = temp;
    b = -c;
    c = 0.0;
  } else if ( ( a - d ) == 0.0 && dsign( 1.0, b ) != dsign( 1.0, c ) ) {
    cs = 1.0;
    sn = 0.0;
  } else {
    double temp = a - d;
    double p = 0.5 * temp;
    const double bcmax = dmax( fabs( b ), fabs( c ) );
    const double bcmis = dmin( fabs( b ), fabs( c ) ) * dsign( 1.0, b ) * dsign( 1.0, c );
    double scale = dmax( fabs( p ), bcmax );
    double z = ( p / scale ) * p + ( bcmax / scale ) * bcmis;
    if ( z >= multpl * eps ) {
      // real eigenvalues
      z = p + dsign( sqrt( scale ) * sqrt( z ), p );
      a = d + z;
      d = d - ( bcmax / z ) * bcmis;
      const double tau = dlapy2( c, z );
      cs = z / tau;
      sn = c / tau;
      b = b - c;
      c = 0.0;
    } else {
      // complex or nearly equal real eigenvalues: make diagonal elements equal
      int count = 0;
      double sigma = b + c;
      for ( ;; ) {
        ++count;
        scale = dmax( fabs( temp ), fabs( sigma ) );
        if ( scale >= safmx2 ) {
          sigma *= safmn2;
          temp *= safmn2;
          if ( count <= 20 ) continue;
          break;
        } else if ( scale <= safmn2 ) {
          sigma *= safmx2;
          temp *= safmx2;
          if ( count <= 20 ) continue;
          break;
        } else {
          p = 0.5 * temp;
          double tau = dlapy2( sigma, temp );
          cs = sqrt( 0.5 * ( 1.0 + fabs( sigma ) / tau ) );
          sn = -( p / ( tau * cs ) ) * dsign( 1.0, sigma );
          const double aa = a * cs + b * sn;
          const double bb = -a * sn + b * cs;
          const double cc = c * cs + d * sn;
          const double dd = -c * sn + d * cs;
          a = aa * cs + cc * sn;
          b = bb * cs + dd * sn;
          c = -aa * sn + cc * cs;
          d = -bb * sn + dd * cs;
          temp = 0.5 * ( a + d );
          a = temp;
          d = temp;
          if ( c != 0.0 ) {
            if ( b != 0.0 ) {
              if ( dsign( 1.0, b ) == dsign( 1.0, c ) ) {
                // real eigenvalues: reduce to upper triangular form
                const double sab = sqrt( fabs( b ) );
                const double sac = sqrt( fabs( c ) );
                p = dsign( sab * sac, c );
                tau = 1.0 / sqrt( fabs( b + c ) );
                a = temp + p;
                d = temp - p;
                b = b - c;
                c = 0.0;
                const double cs1 = sab * tau;
                const double sn1 = sac * tau;
                temp = cs * cs1 - sn * sn1;
                sn = cs * sn1 + sn * cs1;
                cs = temp;
              }
            } else {
              b = -c;
              c = 0.0;
              temp = cs;
              cs = -sn;
              sn = temp;
            }
          }
          break;
        }
      }
    }
  }
  rt1r = a;
  rt2r = d;
  if ( c == 0.0 ) {
    rt1i = 0.0;
    rt2i = 0.0;
  } else {
    rt1i = sqrt( fabs( b ) ) * sqrt( fabs( c ) );
    rt2i = -rt1i;
  }
}

// ---- dlarfg on a tiny vector held in scalars (n = 2 or 3), as used by the Francis sweep -----------
// in: alpha = x0, x1, x2 ; out: beta (returned in alpha), v1, v2 (x scaled), tau
PB_HD void dlarfg3( int n, double& alpha, double& x1, double& x2, double& tau ) {
  if ( n <= 1 ) {
    tau = 0.0;
    return;
  }
  double xnorm = ( n == 3 ) ? dlapy2( x1, x2 ) : fabs( x1 );
  if ( xnorm == 0.0 ) {
    tau = 0.0;
    return;
  }
  double beta = -dsign( dlapy2( alpha, xnorm ), alpha );
  const double safmin = kSafmin / kEps;
  int knt = 0;
  if ( fabs( beta ) < safmin ) {
    const double rsafmn = 1.0 / safmin;
    do {
      ++knt;
      x1 *= rsafmn;
      x2 *= rsafmn;
      beta *= rsafmn;
      alpha *= rsafmn;
    } while ( fabs( beta ) < safmin && knt < 20 );
    xnorm = ( n == 3 ) ? dlapy2( x1, x2 ) : fabs( x1 );
    beta = -dsign( dlapy2( alpha, xnorm ), alpha );
  }
  tau = ( beta - alpha ) / beta;
  const double sc = 1.0 / ( alpha - beta );
  x1 *= sc;
  x2 *= sc;
  for ( int j = 0; j < knt; ++j ) beta *= safmin;
  alpha = beta;
}

// 3-element Householder reflector, fast path: one square root and two independent divisions; falls back on the
// rescaling dlarfg3 when the entries are so large or small that the plain sum of squares could over/underflow.
// Same contract as dlarfg3 (alpha <- beta, x <- v, tau); for n < 3 the unused entry is returned as 0.
PB_HD void larfg3_fast( int n, double& a0, double& a1, double& a2, double& t1 ) {
  if ( n < 3 ) a2 = 0.0;
  if ( n <= 1 ) {
    t1 = 0.0;
    return;
  }
  const double m = dmax( fabs( a0 ), dmax( fabs( a1 ), fabs( a2 ) ) );
  if ( !( m > 1.0e-140 && m < 1.0e140 ) ) {
    dlarfg3( n, a0, a1, a2, t1 );
    if ( n < 3 ) a2 = 0.0;
    return;
  }
  const double xn2 = a1 * a1 + a2 * a2;
  if ( xn2 == 0.0 ) {
    t1 = 0.0;
    return;
  }
#if defined( __CUDA_ARCH__ )
  // one reciprocal square root and one reciprocal on the critical path (the Francis step is bound by this latency):
  // norm = |x| , beta = -sign(a0) norm, tau = (beta - a0) / beta = 1 + |a0| / norm, v = x / (a0 - beta) = sign(a0) x / (|a0| + norm)
  const double nn = a0 * a0 + xn2;
  const double rn = rsqrt( nn );
  const double nrm = nn * rn;
  const double aa = fabs( a0 );
  const double rs = dsign( 1.0 / ( aa + nrm ), a0 );
  const double beta = -dsign( nrm, a0 );
  t1 = 1.0 + aa * rn;
#else
  const double beta = -dsign( sqrt( a0 * a0 + xn2 ), a0 );
  const double rb = 1.0 / beta, rs = 1.0 / ( a0 - beta );
  t1 = ( beta - a0 ) * rb;
#endif
  a1 *= rs;
  a2 *= rs;
  a0 = beta;
}

// Branch-free reflector for the register-resident Francis sweep (pb_small.cuh::francis_sweep_lanes): the step loop must
// stay one basic block so that the updates of the previous step are scheduled into the latency of this chain. The
// entries are scaled by the power of two that brings the largest one into [1, 2) -- no over/underflow of the squares
// whatever the input, no rescaling loop --, the reciprocal square root and the reciprocal start from the hardware
// approximations (MUFU.RSQ64H / MUFU.RCP64H, 2^-23) and are refined in registers (one third-order / two second-order
// steps: full precision up to the last rounding, no special-case paths: the arguments are in [1, 12] and [1, 6]).
// A column below 2^-1022 in magnitude is treated as zero; x with |x|^2 underflowing against a0^2 (ratio below 1e-154) gives
// tau = 0 exactly as dlarfg does for x = 0. Same contract as larfg3_fast.
PB_HD void larfg3_bf( int n, double& a0, double& a1, double& a2, double& t1 ) {
#if defined( __CUDA_ARCH__ )
  a2 = n < 3 ? 0.0 : a2;
  a1 = n < 2 ? 0.0 : a1;
  const double mx = fmax( fabs( a0 ), fmax( fabs( a1 ), fabs( a2 ) ) );
  const int ex = imin( __double2hiint( mx ) & 0x7ff00000, 0x7fd00000 );  // (clamped: the scale itself must stay normal)
  const double sc = __hiloint2double( 0x7fe00000 - ex, 0 );  // 2^(1023 - E): mx * sc in [1, 2)
  const double si = __hiloint2double( ex, 0 );               // 2^(E - 1023)  (0 for a zero or subnormal column)
  const double x0 = a0 * sc, x1 = a1 * sc, x2 = a2 * sc;
  const double xn2 = x1 * x1 + x2 * x2;
  const double nn = x0 * x0 + xn2;  // in [1, 12]
  double y;
  asm( "rsqrt.approx.ftz.f64 %0, %1;" : "=d"( y ) : "d"( nn ) );
  {
    const double e = fma( -nn * y, y, 1.0 );
    y = fma( y * e, fma( e, 0.375, 0.5 ), y );
  }
  double nrm = nn * y;
  nrm = fma( fma( -nrm, nrm, nn ), 0.5 * y, nrm );  // sqrt(nn), last-bit correction
  const double aa = fabs( x0 );
  const double den = aa + nrm;  // in [1, 6]
  double r;
  asm( "rcp.approx.ftz.f64 %0, %1;" : "=d"( r ) : "d"( den ) );
  {
    double e = fma( -den, r, 1.0 );
    r = fma( r, fma( e, e, e ), r );
    e = fma( -den, r, 1.0 );
    r = fma( r, e, r );
  }
  const bool refl = xn2 > 0.0;
  const double rs = a0 >= 0.0 ? r : -r;
  const double bt = a0 >= 0.0 ? -nrm : nrm;
  t1 = refl ? fma( aa, y, 1.0 ) : 0.0;
  a1 = refl ? x1 * rs : 0.0;
  a2 = refl ? x2 * rs : 0.0;
  a0 = refl ? bt * si : a0;
#else
  larfg3_fast( n, a0, a1, a2, t1 );
  if ( n < 3 ) a2 = 0.0;
#endif
}

// dlartg (LAPACK 3.10+ la_lartg semantics for the unscaled range): r has the sign of f
PB_HD void dlartg( double f, double g, double& c, double& s, double& r ) {
  if ( g == 0.0 ) {
    c = 1.0;
    s = 0.0;
    r = f;
  } else if ( f == 0.0 ) {
    c = 0.0;
    s = dsign( 1.0, g );
    r = fabs( g );
  } else {
    const double d = dlapy2( f, g );
    c = fabs( f ) / d;
    r = dsign( d, f );
    s = g / r;
  }
}

}  // namespace pb
