// pb_text.cuh -- text <-> binary64 conversions of the paladin file formats, bit-exact with the C library calls the reference
// makes, usable on the device and (the same source) on the host:
//
//   text_to_double : the value field of a MatrixMarket entry line, converted by the reference with atof
//                    (reference src/square-matrix.hpp:114). Plain decimal spellings with at most 19 significant digits take
//                    the Eisel-Lemire path (correctly rounded like strtod; D. Lemire, "Number parsing at a gigabyte per
//                    second", 2021 -- the algorithm std::from_chars uses); everything else (hex floats, inf / nan, longer
//                    digit strings, out-of-range exponents) is REFUSED: the caller converts those fields with the C library.
//   text_to_index  : the row / column fields (atoi, src/square-matrix.hpp:112-113): optional sign, digits; refused otherwise.
//   format_g6      : a double as operator<<(double) prints it at the default precision of 6 significant digits, i.e. printf's
//                    "%g" (reference src/spectrum-util.hpp:122, :150-151, :186-187): the decimal expansion is cut from a 192-bit
//                    product with the same power table; a value too close to a rounding boundary for that precision to decide
//                    (or an exact tie the table cannot represent exactly) is REFUSED and formatted by the C library.
//
// Refusals never produce a wrong byte: both directions fall back on the reference's own library calls.
#pragma once

#include <cstdint>

#include "pb_common.cuh"
#include "pb_pow5_table.h"

namespace pb {

#if defined( __CUDACC__ )
__device__ const uint64_t kPow5Dev[] = { PB_POW5_TABLE_BODY };
#endif
static const uint64_t kPow5Host[] = { PB_POW5_TABLE_BODY };

PB_HD uint64_t pow5_word( int q, int lo ) {
#if defined( __CUDA_ARCH__ )
  return kPow5Dev[2 * ( q - PB_POW5_QMIN ) + lo];
#else
  return kPow5Host[2 * ( q - PB_POW5_QMIN ) + lo];
#endif
}

struct U128 {
  uint64_t hi, lo;
};
PB_HD U128 mul64( uint64_t a, uint64_t b ) {
#if defined( __CUDA_ARCH__ )
  return U128{ __umul64hi( a, b ), a * b };
#else
  const unsigned __int128 p = (unsigned __int128)a * b;
  return U128{ (uint64_t)( p >> 64 ), (uint64_t)p };
#endif
}
PB_HD int clz64( uint64_t x ) {
#if defined( __CUDA_ARCH__ )
  return __clzll( (long long)x );
#else
  return __builtin_clzll( x );
#endif
}
PB_HD double bits_to_double( uint64_t b ) {
#if defined( __CUDA_ARCH__ )
  return __longlong_as_double( (long long)b );
#else
  double d;
  __builtin_memcpy( &d, &b, 8 );
  return d;
#endif
}
PB_HD uint64_t double_to_bits( double d ) {
#if defined( __CUDA_ARCH__ )
  return (uint64_t)__double_as_longlong( d );
#else
  uint64_t b;
  __builtin_memcpy( &b, &d, 8 );
  return b;
#endif
}

// ---- decimal (w * 10^q, w != 0 with at most 19 digits) -> nearest binary64, ties to even ------------------------------------
// Returns false when the result is not certain (never observed for w < 10^19 inside the table range; kept as a guard) or
// the exponent is outside the table.
PB_HD bool decimal_to_double( uint64_t w, int q, bool neg, double* out ) {
  if ( q < PB_POW5_QMIN || q > PB_POW5_QMAX ) return false;
  const int lz = clz64( w );
  w <<= lz;
  // 128-bit product with the truncated power of five; the second word only when the first cannot decide (55 bits needed)
  U128 p = mul64( w, pow5_word( q, 0 ) );
  const uint64_t precision_mask = 0xFFFFFFFFFFFFFFFFull >> 55;
  if ( ( p.hi & precision_mask ) == precision_mask ) {
    const U128 p2 = mul64( w, pow5_word( q, 1 ) );
    p.lo += p2.hi;
    if ( p2.hi > p.lo ) p.hi++;
    if ( p.lo == 0xFFFFFFFFFFFFFFFFull && ( q < -27 || q > 55 ) ) return false;  // conservative guard of the original algorithm
  }
  const int upperbit = (int)( p.hi >> 63 );
  uint64_t mant = p.hi >> ( upperbit + 64 - 52 - 3 );
  int power2 = ( ( ( 152170 + 65536 ) * q ) >> 16 ) + 63 + upperbit - lz + 1023;
  if ( power2 <= 0 ) {  // subnormal
    if ( -power2 + 1 >= 64 ) {
      *out = neg ? -0.0 : 0.0;
      return true;
    }
    mant >>= -power2 + 1;
    mant += ( mant & 1 );
    mant >>= 1;
    power2 = ( mant < ( 1ull << 52 ) ) ? 0 : 1;
    *out = bits_to_double( ( (uint64_t)power2 << 52 ) | ( mant & ~( 1ull << 52 ) ) | ( neg ? 0x8000000000000000ull : 0 ) );
    return true;
  }
  // round to even when exactly between two floats (only possible for small |q|)
  if ( p.lo <= 1 && q >= -4 && q <= 23 && ( mant & 3 ) == 1 ) {
    if ( ( mant << ( upperbit + 64 - 52 - 3 ) ) == p.hi ) mant &= ~1ull;
  }
  mant += ( mant & 1 );
  mant >>= 1;
  if ( mant >= ( 2ull << 52 ) ) {
    mant = 1ull << 52;
    power2++;
  }
  mant &= ~( 1ull << 52 );
  if ( power2 >= 0x7FF ) return false;  // overflow: the C library decides (inf + errno)
  *out = bits_to_double( ( (uint64_t)power2 << 52 ) | mant | ( neg ? 0x8000000000000000ull : 0 ) );
  return true;
}

// ---- value field -> double. Grammar taken here: [+-] digits [. digits] [ (e|E) [+-] digits ], at least one digit, nothing
// else in the field (leading blanks are accepted like atof does). Returns false = "convert this field with atof". ------------
PB_HD bool text_to_double( const char* p, int len, double* out ) {
  int i = 0;
  while ( i < len && ( p[i] == ' ' || ( p[i] >= '\t' && p[i] <= '\r' ) ) ) ++i;
  bool neg = false;
  if ( i < len && ( p[i] == '-' || p[i] == '+' ) ) neg = ( p[i++] == '-' );
  uint64_t w = 0;
  int ndig = 0;      // significant digits taken into w
  int nseen = 0;     // digits seen at all
  int scale = 0;     // decimal exponent contributed by the digits' positions
  bool point = false;
  for ( ; i < len; ++i ) {
    const char c = p[i];
    if ( c >= '0' && c <= '9' ) {
      ++nseen;
      if ( w == 0 && c == '0' ) {  // leading zero: only moves the point
        if ( point ) --scale;
        continue;
      }
      if ( ndig == 19 ) return false;  // more digits than a uint64 holds exactly: the C library rounds those
      w = w * 10 + (uint64_t)( c - '0' );
      ++ndig;
      if ( point ) --scale;
    } else if ( c == '.' && !point ) {
      point = true;
    } else {
      break;
    }
  }
  if ( nseen == 0 ) return false;
  int ex = 0;
  if ( i < len && ( p[i] == 'e' || p[i] == 'E' ) ) {
    ++i;
    bool eneg = false;
    if ( i < len && ( p[i] == '-' || p[i] == '+' ) ) eneg = ( p[i++] == '-' );
    int nd = 0;
    for ( ; i < len && p[i] >= '0' && p[i] <= '9'; ++i ) {
      if ( ex < 100000 ) ex = ex * 10 + ( p[i] - '0' );
      ++nd;
    }
    if ( nd == 0 ) return false;
    if ( eneg ) ex = -ex;
  }
  if ( i != len ) return false;  // trailing characters: atof would stop there, let it
  if ( w == 0 ) {
    *out = neg ? -0.0 : 0.0;
    return true;
  }
  return decimal_to_double( w, ex + scale, neg, out );
}

// ---- index field -> int (atoi): blanks, optional sign, 1..9 digits, nothing else --------------------------------------------
PB_HD bool text_to_index( const char* p, int len, int* out ) {
  int i = 0;
  while ( i < len && ( p[i] == ' ' || ( p[i] >= '\t' && p[i] <= '\r' ) ) ) ++i;
  bool neg = false;
  if ( i < len && ( p[i] == '-' || p[i] == '+' ) ) neg = ( p[i++] == '-' );
  int v = 0, nd = 0;
  for ( ; i < len && p[i] >= '0' && p[i] <= '9'; ++i ) {
    if ( nd == 9 ) return false;
    v = v * 10 + ( p[i] - '0' );
    ++nd;
  }
  if ( nd == 0 || i != len ) return false;
  *out = neg ? -v : v;
  return true;
}

// ---- "%g" at 6 significant digits --------------------------------------------------------------------------------------------
// buf needs 16 bytes. Returns the number of characters, or 0 = refused (non-finite, or the rounding could not be decided).
PB_HD int format_g6( double x, char* buf ) {
  const uint64_t bits = double_to_bits( x );
  const bool neg = ( bits >> 63 ) != 0;
  const int bexp = (int)( ( bits >> 52 ) & 0x7FF );
  uint64_t frac = bits & 0x000FFFFFFFFFFFFFull;
  if ( bexp == 0x7FF ) return 0;
  int n = 0;
  if ( neg ) buf[n++] = '-';
  if ( bexp == 0 && frac == 0 ) {
    buf[n++] = '0';
    return n;
  }
  // x = m * 2^e2, m normalised to 64 bits (top bit set)
  int e2;
  uint64_t m;
  if ( bexp == 0 ) {
    m = frac;
    e2 = -1074;
  } else {
    m = frac | ( 1ull << 52 );
    e2 = bexp - 1075;
  }
  const int lz = clz64( m );
  m <<= lz;
  e2 -= lz;  // x = m * 2^e2, 2^63 <= m < 2^64
  // decimal exponent estimate k = floor(log10 x): log2 x in [e2 + 63, e2 + 64)
  int k = ( ( e2 + 63 ) * 78913 ) >> 18;  // floor((e2+63) * log10(2)), exact for this range
  uint32_t D = 0;
  for ( int attempt = 0; attempt < 2; ++attempt ) {
    // D.f = x * 10^(5-k) = m * 2^e2 * 5^q * 2^q with q = 5 - k
    const int q = 5 - k;
    if ( q < PB_POW5_QMIN || q > PB_POW5_QMAX ) return 0;
    const uint64_t th = pow5_word( q, 0 ), tl = pow5_word( q, 1 );
    // 5^q = T * 2^t, T = th:tl in [2^127, 2^128), t = floor(log2 5^q) - 127
    const int t = ( ( 217706 * q ) >> 16 ) - q - 127;  // floor(log2 10^q) - q: the verified exponent formula of the parser
    // P = m * T (192 bits: p2:p1:p0)
    const U128 a = mul64( m, th );
    const U128 b = mul64( m, tl );
    uint64_t p0 = b.lo;
    uint64_t p1 = a.lo + b.hi;
    uint64_t p2 = a.hi + ( p1 < a.lo ? 1 : 0 );
    // value = P * 2^s, s = e2 + t + q  (negative: the integer part has at most 21 bits)
    const int s = e2 + t + q;
    const int sh = -s;  // 192 - 21 >= sh >= 192 - 22 is the normal case; always 128 < sh < 192 here
    if ( sh <= 128 || sh >= 192 ) return 0;
    const int sh2 = sh - 128;  // bits of p2 below the binary point
    const uint64_t ip = p2 >> sh2;
    const uint64_t fmask = ( 1ull << sh2 ) - 1;
    const uint64_t f2 = p2 & fmask;  // fraction: f2 : p1 : p0 over 2^sh
    const uint64_t half = 1ull << ( sh2 - 1 );
    if ( ip >= 1000000 ) {  // estimate one too small
      ++k;
      continue;
    }
    if ( ip < 100000 ) {  // estimate one too large
      --k;
      continue;
    }
    // rounding: compare the fraction with 1/2. The product is exact when T is (0 <= q <= 55); otherwise T carries a relative
    // error below 2^-127 (truncated for q > 55, rounded up for q < 0), i.e. below 2^65 units of p0: undecidable when the
    // fraction is that close to 1/2.
    const bool exact = ( q >= 0 && q <= 55 );
    bool up;
    if ( exact ) {
      if ( f2 != half ) up = f2 > half;
      else if ( p1 != 0 || p0 != 0 ) up = true;
      else up = ( ip & 1 ) != 0;  // exact tie: round half to even, like the C library in round-to-nearest mode
    } else {
      // distance of the fraction from 1/2 in units of p1, decided only outside a margin of 4 units (twice the error bound)
      if ( f2 > half ) up = true;
      else if ( f2 == half ) {
        if ( p1 < 4 ) return 0;
        up = true;
      } else if ( half - f2 >= 2 ) up = false;
      else {
        if ( p1 > 0xFFFFFFFFFFFFFFFFull - 4 ) return 0;
        up = false;
      }
    }
    D = (uint32_t)ip + ( up ? 1u : 0u );
    if ( D == 1000000 ) {
      D = 100000;
      ++k;
    }
    break;
  }
  if ( D < 100000 || D >= 1000000 ) return 0;
  // digits, trailing zeros stripped
  char dg[6];
  uint32_t v = D;
  for ( int i = 5; i >= 0; --i ) {
    dg[i] = (char)( '0' + v % 10 );
    v /= 10;
  }
  int nd = 6;
  while ( nd > 1 && dg[nd - 1] == '0' ) --nd;
  if ( k >= -4 && k < 6 ) {
    if ( k >= 0 ) {
      for ( int i = 0; i <= k; ++i ) buf[n++] = ( i < nd ) ? dg[i] : '0';
      if ( nd > k + 1 ) {
        buf[n++] = '.';
        for ( int i = k + 1; i < nd; ++i ) buf[n++] = dg[i];
      }
    } else {
      buf[n++] = '0';
      buf[n++] = '.';
      for ( int i = 0; i < -k - 1; ++i ) buf[n++] = '0';
      for ( int i = 0; i < nd; ++i ) buf[n++] = dg[i];
    }
  } else {
    buf[n++] = dg[0];
    if ( nd > 1 ) {
      buf[n++] = '.';
      for ( int i = 1; i < nd; ++i ) buf[n++] = dg[i];
    }
    buf[n++] = 'e';
    int ke = k;
    if ( ke < 0 ) {
      buf[n++] = '-';
      ke = -ke;
    } else {
      buf[n++] = '+';
    }
    if ( ke >= 100 ) buf[n++] = (char)( '0' + ke / 100 );
    buf[n++] = (char)( '0' + ( ke / 10 ) % 10 );
    buf[n++] = (char)( '0' + ke % 10 );
  }
  return n;
}

// decimal digits of a positive int (at most 10), returns the count
PB_HD int format_uint( uint32_t v, char* buf ) {
  char tmp[10];
  int n = 0;
  do {
    tmp[n++] = (char)( '0' + v % 10 );
    v /= 10;
  } while ( v != 0 );
  for ( int i = 0; i < n; ++i ) buf[i] = tmp[n - 1 - i];
  return n;
}

}  // namespace pb
