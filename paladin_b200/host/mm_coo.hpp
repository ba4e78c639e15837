// mm_coo.hpp -- MatrixMarket "coordinate real general" reader that stops at COO triplets: the dense scatter the
// reference performs while parsing (src/square-matrix.hpp:115) happens on the GPU (include/paladin_gpu.h).
//
// Same accepted format and the same conversions as the reference reader:
//   banner line consumed unconditionally, then every line starting with '%'   (src/square-matrix.hpp:132-144)
//   size line split at single spaces, empty tokens dropped, std::stoi           (:151-163)
//   nnz entry lines "i j v": two fields ended by ' ', one by '\n', converted with atoi / atoi / atof (:101-117)
// Where the reference has undefined or silently wrong behaviour (index fields longer than 5 characters, value
// fields longer than 31, which set the stream's failbit and zero every later entry; non-square or overfull
// size lines behind assert()) this reader reports an error instead.
#ifndef PALADIN_GPU_HOST_MM_COO_HPP_
#define PALADIN_GPU_HOST_MM_COO_HPP_

#include <charconv>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "cli.hpp"

namespace paladin {

struct CooMatrix {
  int nrows_ = 0;
  int nnz_ = 0;
  std::vector<int> row, col;  // 1-based, file order
  std::vector<double> val;
};

inline int count_nonempty_lines( const std::string& filename ) {
  std::ifstream in( filename );
  std::string line;
  int n = 0;
  while ( std::getline( in, line ) && !line.empty() ) ++n;
  return n;
}

namespace detail {

inline const char* line_end( const char* p, const char* end ) {
  while ( p < end && *p != '\n' ) ++p;
  return p;
}
inline const char* skip_line( const char* p, const char* end ) {
  p = line_end( p, end );
  return p < end ? p + 1 : end;
}

// size line "n n nnz" of a MatrixMarket text; *body receives the first entry line. with_banner: the first line is
// consumed unconditionally (full reader); otherwise only '%' lines are skipped (the reference's measure scan,
// src/square-matrix.hpp:187-228, never looks at the banner).
inline void parse_size_line( const char* text, const char* end, bool with_banner, int& n, int& nnz, const char** body,
                             const std::string& what ) {
  const char* p = text;
  if ( with_banner ) {
    const std::string banner( p, line_end( p, end ) );
    const std::vector<std::string> tok = split_string( banner );
    if ( tok.size() < 5 || ( tok[0] != "%MatrixMarket" && tok[0] != "%%MatrixMarket" ) || tok[1] != "matrix" ||
         tok[2] != "coordinate" || tok[3] != "real" || tok[4] != "general" )
      throw std::runtime_error( what + ": not a 'MatrixMarket matrix coordinate real general' file" );
    p = skip_line( p, end );
  }
  while ( p < end && *p == '%' ) p = skip_line( p, end );
  const std::string sizes( p, line_end( p, end ) );
  const std::vector<std::string> tok = split_string( sizes );
  if ( tok.size() < 3 ) throw std::runtime_error( what + ": malformed size line" );
  n = std::stoi( tok[0] );
  const int cols = std::stoi( tok[1] );
  nnz = std::stoi( tok[2] );
  if ( n != cols ) throw std::runtime_error( what + ": matrix is not square" );
  if ( n < 0 || nnz < 0 || static_cast<long long>( nnz ) > static_cast<long long>( n ) * n )
    throw std::runtime_error( what + ": entry count exceeds n*n" );
  if ( body ) *body = skip_line( p, end );
}

}  // namespace detail

namespace detail {

// atoi on a field known to be short: optional blanks, optional sign, digits (same value as atoi for every input atoi
// defines)
inline int field_to_int( const char* p, const char* end ) {
  while ( p < end && ( *p == ' ' || ( *p >= '\t' && *p <= '\r' ) ) ) ++p;
  bool neg = false;
  if ( p < end && ( *p == '+' || *p == '-' ) ) neg = ( *p++ == '-' );
  int v = 0;
  while ( p < end && *p >= '0' && *p <= '9' ) v = v * 10 + ( *p++ - '0' );
  return neg ? -v : v;
}

// atof on a field: std::from_chars (correctly rounded, like strtod, several times faster) for plain decimal numbers,
// the C library for everything else it accepts (hex floats, inf/nan spellings, out-of-range values, odd signs), so
// the result is bit-identical to the reference's atof call (src/square-matrix.hpp:114).
inline double field_to_double( const char* p, const char* end ) {
  const char* q = p;
  while ( q < end && ( *q == ' ' || ( *q >= '\t' && *q <= '\r' ) ) ) ++q;
  const char* num = q;
  if ( num < end && *num == '+' ) ++num;  // from_chars takes no '+'
  const char* dig = ( num < end && *num == '-' && num == q ) ? num + 1 : num;
  const bool plain = dig < end && ( ( *dig >= '0' && *dig <= '9' ) || *dig == '.' ) &&
                     !( dig + 1 < end && *dig == '0' && ( dig[1] == 'x' || dig[1] == 'X' ) );
  if ( plain ) {
    double v = 0.0;
    const std::from_chars_result r = std::from_chars( num, end, v, std::chars_format::general );
    if ( r.ec == std::errc() ) return v;
  }
  char buf[40];
  const std::size_t len = std::min<std::size_t>( static_cast<std::size_t>( end - p ), sizeof( buf ) - 1 );
  std::memcpy( buf, p, len );
  buf[len] = '\0';
  return std::atof( buf );
}

}  // namespace detail

inline std::string slurp( const std::string& path ) {
  std::ifstream in( path, std::ios::binary | std::ios::ate );
  if ( !in ) throw std::runtime_error( "Couldn't open the matrix file, " + path );
  const std::streamoff size = in.tellg();
  if ( size < 0 ) {  // not seekable: stream it
    in.clear();
    std::ostringstream ss;
    ss << in.rdbuf();
    return ss.str();
  }
  std::string text( static_cast<std::size_t>( size ), '\0' );
  in.seekg( 0 );
  in.read( text.data(), size );
  text.resize( static_cast<std::size_t>( in.gcount() ) );
  return text;
}

// dimension and entry count from the size line only (what the load measures need)
inline void read_matrix_sizes_from_mm_file( const std::string& path, int& n, int& nnz ) {
  std::ifstream in( path, std::ios::binary );
  if ( !in ) throw std::runtime_error( "Couldn't open the matrix file, " + path );
  std::string head, line;
  while ( in.peek() == '%' ) std::getline( in, line );
  std::getline( in, line );
  detail::parse_size_line( line.data(), line.data() + line.size(), false, n, nnz, nullptr, path );
}

inline CooMatrix parse_mm_text( const std::string& text, const std::string& what ) {
  CooMatrix m;
  const char* end = text.data() + text.size();
  const char* p = nullptr;
  detail::parse_size_line( text.data(), end, true, m.nrows_, m.nnz_, &p, what );
  m.row.resize( m.nnz_ );
  m.col.resize( m.nnz_ );
  m.val.resize( m.nnz_ );
  for ( int k = 0; k < m.nnz_; ++k ) {
    const char* f[3];
    std::size_t len[3];
    for ( int part = 0; part < 3; ++part ) {
      const char stop = part < 2 ? ' ' : '\n';
      f[part] = p;
      while ( p < end && *p != stop ) ++p;
      len[part] = static_cast<std::size_t>( p - f[part] );
      if ( p < end ) ++p;
    }
    if ( len[0] > 5 || len[1] > 5 || len[2] > 31 )
      throw std::runtime_error( what + ": entry " + std::to_string( k + 1 ) +
                                " has a field longer than the reference reader's buffers (5/5/31 characters)" );
    m.row[k] = detail::field_to_int( f[0], f[0] + len[0] );
    m.col[k] = detail::field_to_int( f[1], f[1] + len[1] );
    m.val[k] = detail::field_to_double( f[2], f[2] + len[2] );
  }
  return m;
}

inline CooMatrix read_coo_from_mm_file( const std::string& path ) { return parse_mm_text( slurp( path ), path ); }

}  // namespace paladin

#endif
