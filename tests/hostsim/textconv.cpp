// TEST-ONLY: the text <-> binary64 conversions of paladin_b200/csrc/pb_text.cuh compiled for the host and checked against
// the C library calls the reference makes (atof = strtod, operator<<(double) = printf "%g").
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

#include "../../paladin_b200/csrc/pb_text.cuh"

namespace {
std::string random_decimal( std::mt19937_64& g, int mode ) {
  char buf[80];
  std::uniform_real_distribution<double> u( -1.0, 1.0 );
  std::normal_distribution<double> nrm;
  switch ( mode ) {
    case 0: std::snprintf( buf, sizeof buf, "%.17g", nrm( g ) ); break;                       // the benchmark's own format
    case 1: {                                                                                // any finite double, 17 digits
      uint64_t b = g();
      double d;
      std::memcpy( &d, &b, 8 );
      if ( !std::isfinite( d ) ) d = 1.5;
      std::snprintf( buf, sizeof buf, "%.17g", d );
      break;
    }
    case 2: {                                                                                // short spellings
      std::snprintf( buf, sizeof buf, "%.*g", 1 + (int)( g() % 16 ), nrm( g ) * std::pow( 10.0, (double)( (int)( g() % 40 ) - 20 ) ) );
      break;
    }
    case 3: {                                                                                // random digit strings, random exponents
      int nd = 1 + (int)( g() % 19 );
      std::string s = ( g() & 1 ) ? "-" : "";
      int pt = (int)( g() % ( nd + 1 ) );
      for ( int i = 0; i < nd; ++i ) {
        if ( i == pt ) s += '.';
        s += (char)( '0' + g() % 10 );
      }
      if ( g() % 3 ) s += ( g() & 1 ? "e" : "E" ) + std::to_string( (int)( g() % 700 ) - 350 );
      return s;
    }
    case 4: {                                                                                // halfway points between doubles
      uint64_t b = ( g() >> 12 ) | ( (uint64_t)( 1023 - 60 + g() % 120 ) << 52 );
      double lo;
      std::memcpy( &lo, &b, 8 );
      const double hi = std::nextafter( lo, INFINITY );
      // exact midpoint printed with enough digits is longer than 19 digits -> mostly refused; use 19 digits near it
      std::snprintf( buf, sizeof buf, "%.18e", lo + ( hi - lo ) * 0.5 );
      break;
    }
    default: {                                                                               // subnormals and the range ends
      const double d = std::ldexp( u( g ), (int)( g() % 80 ) - 1100 + ( g() & 1 ? 2100 : 0 ) );
      std::snprintf( buf, sizeof buf, "%.17g", d );
    }
  }
  return buf;
}
}  // namespace

extern "C" {

int tc_text_to_double( const char* s, int len, double* out ) { return pb::text_to_double( s, len, out ) ? 1 : 0; }
int tc_text_to_index( const char* s, int len, int* out ) { return pb::text_to_index( s, len, out ) ? 1 : 0; }
int tc_format_g6( double x, char* buf ) { return pb::format_g6( x, buf ); }

// returns mismatches; *refused receives the number of refusals
long tc_check_parse( unsigned long long seed, long count, int mode, long* refused, char* first_bad ) {
  std::mt19937_64 g( seed );
  long bad = 0, ref = 0;
  for ( long i = 0; i < count; ++i ) {
    const std::string s = random_decimal( g, mode );
    double got = 0.0;
    if ( !pb::text_to_double( s.data(), (int)s.size(), &got ) ) {
      ++ref;
      continue;
    }
    const double want = std::atof( s.c_str() );
    if ( std::memcmp( &got, &want, 8 ) != 0 ) {
      if ( bad == 0 && first_bad ) std::snprintf( first_bad, 120, "%s -> %a, atof %a", s.c_str(), got, want );
      ++bad;
    }
  }
  *refused = ref;
  return bad;
}

long tc_check_format( unsigned long long seed, long count, int mode, long* refused, char* first_bad ) {
  std::mt19937_64 g( seed );
  std::normal_distribution<double> nrm;
  long bad = 0, ref = 0;
  for ( long i = 0; i < count; ++i ) {
    double x;
    switch ( mode ) {
      case 0: x = nrm( g ); break;                                    // eigenvector components / eigenvalues
      case 1: {                                                        // any bit pattern
        uint64_t b = g();
        std::memcpy( &x, &b, 8 );
        break;
      }
      case 2: {                                                        // 6-digit ties and near-ties at every decade
        const long d = 100000 + (long)( g() % 900000 );
        const int e = (int)( g() % 60 ) - 30;
        x = ( (double)d + 0.5 ) * std::pow( 10.0, (double)e );
        if ( g() & 1 ) x = std::nextafter( x, ( g() & 2 ) ? INFINITY : -INFINITY );
        if ( g() & 4 ) x = -x;
        break;
      }
      case 3: {                                                        // short decimals: exact representable ties like 0.5, 2.5e-1
        x = (double)( (long)( g() % 20000001 ) - 10000000 ) / (double)( 1 << ( g() % 12 ) );
        break;
      }
      default: {                                                       // around the %g notation switches and powers of ten
        const int e = (int)( g() % 24 ) - 12;
        x = std::pow( 10.0, (double)e ) * ( 1.0 + ( (double)( (long)( g() % 2001 ) - 1000 ) ) * 1e-9 );
        if ( g() & 1 ) x = 9.999995 * std::pow( 10.0, (double)e ) + nrm( g ) * 1e-12 * std::pow( 10.0, (double)e );
      }
    }
    char got[24], want[40];
    const int n = pb::format_g6( x, got );
    if ( n == 0 ) {
      ++ref;
      continue;
    }
    const int m = std::snprintf( want, sizeof want, "%g", x );
    if ( n != m || std::memcmp( got, want, n ) != 0 ) {
      if ( bad == 0 && first_bad ) std::snprintf( first_bad, 120, "%a (%.17g) -> %.*s, printf %s", x, x, n, got, want );
      ++bad;
    }
  }
  *refused = ref;
  return bad;
}
}
