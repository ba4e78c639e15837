// balance.hpp -- load measures and the static balancer, bit-identical to the reference:
//   measures        reference src/square-matrix.hpp:236-268 (double arithmetic, enum order of
//                   src/measure-util.hpp:40, option strings :48-69, descriptions :76-97)
//   sort            std::sort, descending by measure, NOT stable (src/measure-util.hpp:104-111). Tie order of
//                   an unstable sort is implementation-defined, so this code calls the same libstdc++
//                   std::sort on the same element type (pair<string,double>) with the same strict comparator.
//   pod colouring   greedy: each matrix in sorted order goes to the first pod of minimal load
//                   (src/paladin.hpp:149-155). A pod is an MPI rank in the reference and a GPU here.
#ifndef PALADIN_GPU_HOST_BALANCE_HPP_
#define PALADIN_GPU_HOST_BALANCE_HPP_

#include <algorithm>
#include <iostream>
#include <string>
#include <utility>
#include <vector>

namespace paladin {

enum class LoadPredictor { DIMENSION, NUMNONZEROS, DIMCUBED, NNZDIMSQRD, SPARSITY, SPARSENCUBED };

using NameMeasurePairT = std::pair<std::string, double>;

inline LoadPredictor string_to_measure_type( const std::string& s ) {
  static const std::pair<const char*, LoadPredictor> table[] = {
      { "dim", LoadPredictor::DIMENSION },    { "nnz", LoadPredictor::NUMNONZEROS }, { "dcb", LoadPredictor::DIMCUBED },
      { "zds", LoadPredictor::NNZDIMSQRD },   { "sps", LoadPredictor::SPARSITY },    { "spc", LoadPredictor::SPARSENCUBED } };
  for ( const auto& e : table )
    if ( s == e.first ) return e.second;
  std::cerr << "Invalid argument: Invalid load string." << '\n';  // same message, same fallback as the reference
  return LoadPredictor::NUMNONZEROS;
}

inline std::string measure_type_description( LoadPredictor t ) {
  switch ( t ) {
    case LoadPredictor::DIMENSION: return "matrix dimension";
    case LoadPredictor::NUMNONZEROS: return "number of nonzeros";
    case LoadPredictor::DIMCUBED: return "matrix dimension cubed";
    case LoadPredictor::NNZDIMSQRD: return "number nonzeros * dim squared";
    case LoadPredictor::SPARSITY: return "matrix sparsity = nnz / dim / dim";
    case LoadPredictor::SPARSENCUBED: return "sparsity * dim cubed";
  }
  return std::string();
}

inline double matrix_measure( int n, int nonzeros, LoadPredictor t ) {
  const double dim = static_cast<double>( n );
  const double nnz = static_cast<double>( nonzeros );
  switch ( t ) {
    case LoadPredictor::DIMENSION: return dim;
    case LoadPredictor::NUMNONZEROS: return nnz;
    case LoadPredictor::DIMCUBED: return dim * dim * dim;
    case LoadPredictor::NNZDIMSQRD: return dim * dim * nnz;
    case LoadPredictor::SPARSITY: return nnz / dim / dim;
    case LoadPredictor::SPARSENCUBED: return nnz * dim;
  }
  return 0.0;
}

inline void sort_name_measure_pairs( std::vector<NameMeasurePairT>& listing ) {
  struct ByMeasureDescending {
    inline bool operator()( const NameMeasurePairT& a, const NameMeasurePairT& b ) { return a.second > b.second; }
  };
  std::sort( listing.begin(), listing.end(), ByMeasureDescending() );
}

// colour[s] = pod of the matrix at sorted position s
inline std::vector<int> color_pods( const std::vector<NameMeasurePairT>& sorted, int pods ) {
  std::vector<double> load( pods, 0.0 );
  std::vector<int> colour;
  colour.reserve( sorted.size() );
  for ( const auto& entry : sorted ) {
    const int lightest = static_cast<int>( std::min_element( load.begin(), load.end() ) - load.begin() );
    load[lightest] += entry.second;
    colour.push_back( lightest );
  }
  return colour;
}

}  // namespace paladin

#endif
