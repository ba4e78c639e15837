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

#include <cstdlib>
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

inline std::string slurp( const std::string& path ) {
  std::ifstream in( path, std::ios::binary );
  if ( !in ) throw std::runtime_error( "Couldn't open the matrix file, " + path );
  std::ostringstream ss;
  ss << in.rdbuf();
  return ss.str();
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
  char field[40];
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
    for ( int part = 0; part < 3; ++part ) {
      std::memcpy( field, f[part], len[part] );
      field[len[part]] = '\0';
      if ( part == 0 ) m.row[k] = std::atoi( field );
      else if ( part == 1 ) m.col[k] = std::atoi( field );
      else m.val[k] = std::atof( field );
    }
  }
  return m;
}

inline CooMatrix read_coo_from_mm_file( const std::string& path ) { return parse_mm_text( slurp( path ), path ); }

}  // namespace paladin

#endif
