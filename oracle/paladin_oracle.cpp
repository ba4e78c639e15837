/*
 * paladin_oracle.cpp -- TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product.
 *
 * CPU restatement (plain C/C++, C linkage) of the pieces of the reference hot
 * path that must be matched bit-exactly. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this library, and only as a checker.
 *
 *   oracle_mm_header / oracle_mm_entries : MatrixMarket "coordinate real general"
 *        parse of the strict format the reference accepts
 *        (reference src/square-matrix.hpp:132-163 header/size line, :101-117 entries).
 *   oracle_scatter_dense : zero-filled column-major dense scatter, 1-based
 *        indices, later entries overwrite earlier ones
 *        (reference src/square-matrix.hpp:70-71 and :115).
 *   oracle_measure : the six load measures in double arithmetic
 *        (reference src/square-matrix.hpp:236-268, enum order src/measure-util.hpp:40).
 *   oracle_partition : descending std::sort by measure (unstable, libstdc++ introsort)
 *        followed by greedy first-minimum pod colouring
 *        (reference src/measure-util.hpp:104-111, src/paladin.hpp:141-155).
 */
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace {

/* skip to the character after the next '\n' (or to end) */
const char* next_line( const char* p, const char* end ) {
  while ( p < end && *p != '\n' ) ++p;
  return p < end ? p + 1 : end;
}

}  // namespace

extern "C" {

/*
 * Parses banner, comment lines and the size line. Returns the byte offset of the
 * first entry line, or a negative value on malformed input:
 *   -1 bad banner, -2 bad size line, -3 not square, -4 nnz > n*n.
 * Reference: read_header (src/square-matrix.hpp:132-144) consumes the first line
 * unconditionally, then every following line that starts with '%';
 * allocate_matrix (:151-163) splits the next line on single spaces, dropping
 * empty tokens (split_string, src/command-line-parser.hpp:39-50).
 */
long oracle_mm_header( const char* text, long len, int* n, int* nnz ) {
  const char* p = text;
  const char* end = text + len;
  {
    const char* q = next_line( p, end );
    std::string head( p, q );
    while ( !head.empty() && ( head.back() == '\n' || head.back() == '\r' ) ) head.pop_back();
    std::vector<std::string> tok;
    size_t s = 0;
    while ( s <= head.size() ) {
      size_t e = head.find( ' ', s );
      if ( e == std::string::npos ) e = head.size();
      if ( e > s ) tok.push_back( head.substr( s, e - s ) );
      s = e + 1;
    }
    if ( tok.size() < 5 ) return -1;
    if ( !( tok[0] == "%MatrixMarket" || tok[0] == "%%MatrixMarket" ) ) return -1;
    if ( tok[1] != "matrix" || tok[2] != "coordinate" || tok[3] != "real" || tok[4] != "general" ) return -1;
    p = q;
  }
  while ( p < end && *p == '%' ) p = next_line( p, end );
  const char* q = next_line( p, end );
  std::string line( p, q );
  std::vector<std::string> tok;
  size_t s = 0;
  while ( s <= line.size() ) {
    size_t e = line.find( ' ', s );
    if ( e == std::string::npos ) e = line.size();
    if ( e > s ) tok.push_back( line.substr( s, e - s ) );
    s = e + 1;
  }
  if ( tok.size() < 3 ) return -2;
  const int rows = std::atoi( tok[0].c_str() );
  const int cols = std::atoi( tok[1].c_str() );
  const int nz = std::atoi( tok[2].c_str() );
  if ( rows <= 0 ) return -2;
  if ( rows != cols ) return -3;
  if ( static_cast<long long>( nz ) > static_cast<long long>( rows ) * rows || nz < 0 ) return -4;
  *n = rows;
  *nnz = nz;
  return static_cast<long>( q - text );
}

/*
 * Parses nnz entry lines "i j v\n" starting at text+offset, single-space
 * separated (reference src/square-matrix.hpp:108-114: getline(' '), getline(' '),
 * getline('\n') then atoi/atoi/atof). Returns number of entries parsed.
 */
int oracle_mm_entries( const char* text, long len, long offset, int nnz, int* ii, int* jj, double* vv ) {
  const char* p = text + offset;
  const char* end = text + len;
  int k = 0;
  for ( ; k < nnz && p < end; ++k ) {
    const char* a = p;
    while ( p < end && *p != ' ' ) ++p;
    std::string si( a, p );
    if ( p < end ) ++p;
    const char* b = p;
    while ( p < end && *p != ' ' ) ++p;
    std::string sj( b, p );
    if ( p < end ) ++p;
    const char* c = p;
    while ( p < end && *p != '\n' ) ++p;
    std::string sv( c, p );
    if ( p < end ) ++p;
    ii[k] = std::atoi( si.c_str() );
    jj[k] = std::atoi( sj.c_str() );
    vv[k] = std::atof( sv.c_str() );
  }
  return k;
}

/* reference src/square-matrix.hpp:70-71 (zero fill) and :115 (overwrite store) */
void oracle_scatter_dense( int n, int nnz, const int* ii, const int* jj, const double* vv, double* dense ) {
  const long long nelem = static_cast<long long>( n ) * n;
  for ( long long e = 0; e < nelem; ++e ) dense[e] = 0.0;
  for ( int k = 0; k < nnz; ++k ) {
    dense[( ii[k] - 1 ) + static_cast<long long>( n ) * ( jj[k] - 1 )] = vv[k];
  }
}

/* type: 0 dim, 1 nnz, 2 dcb, 3 zds, 4 sps, 5 spc -- reference src/square-matrix.hpp:243-260 */
double oracle_measure( int n, int nnz_i, int type ) {
  const double dim = static_cast<double>( n );
  const double nnz = static_cast<double>( nnz_i );
  switch ( type ) {
    case 0: return dim;
    case 1: return nnz;
    case 2: return dim * dim * dim;
    case 3: return dim * dim * nnz;
    case 4: return nnz / dim / dim;
    case 5: return nnz * dim;
    default: return 0.0;
  }
}

/*
 * order[s]  = index (in listing order) of the matrix at sorted position s
 * color[s]  = pod of the matrix at sorted position s
 * The sort MUST be std::sort from the same libstdc++ with the same strict
 * "greater on .second" comparator on the same element type, because the tie
 * order of an unstable sort is implementation-defined
 * (reference src/measure-util.hpp:104-111). Greedy colouring: first minimum of
 * the running pod loads (src/paladin.hpp:149-155).
 */
void oracle_partition( int M, const double* measures, int P, int* order, int* color ) {
  typedef std::pair<std::string, double> Pair;
  std::vector<Pair> listing;
  listing.reserve( M );
  for ( int i = 0; i < M; ++i ) listing.push_back( Pair( std::to_string( i ), measures[i] ) );
  struct Cmp {
    inline bool operator()( const Pair& l, const Pair& r ) { return l.second > r.second; }
  };
  std::sort( listing.begin(), listing.end(), Cmp() );
  std::vector<double> load( P, 0 );
  for ( int s = 0; s < M; ++s ) {
    order[s] = std::atoi( listing[s].first.c_str() );
    int best = static_cast<int>( std::distance( load.begin(), std::min_element( load.begin(), load.end() ) ) );
    load[best] += listing[s].second;
    color[s] = best;
  }
}

}  // extern "C"
