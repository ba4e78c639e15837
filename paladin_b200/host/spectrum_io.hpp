// spectrum_io.hpp -- eigenpair unpack and the spectrum / eigenvector writers, same bytes as the reference:
//   unpack   LAPACK real-packed (wr, wi, V) -> complex values and vectors. A conjugate pair is recognised iff
//            wi[k] == -wi[k+1] && wi[k] != 0; its vectors are V(:,k) +- i V(:,k+1) with the conjugate built as
//            (re, -im) so that signed zeros survive; the last column is handled apart
//            (reference src/paladin.hpp:256-333).
//   spectrum one line per eigenvalue "re im", default ostream precision (6 significant digits), LAPACK order
//            (src/spectrum-util.hpp:119-124).
//   vectors  MatrixMarket "coordinate complex general": banner, "N N nnz", then "row col re im" for every entry
//            with |x|^2 > threshold * max|x|^2 (squared magnitudes against the un-squared threshold), vector index
//            outermost (src/spectrum-util.hpp:134-196).
#ifndef PALADIN_GPU_HOST_SPECTRUM_IO_HPP_
#define PALADIN_GPU_HOST_SPECTRUM_IO_HPP_

#include <charconv>
#include <cmath>
#include <complex>
#include <cstdio>
#include <string>
#include <vector>

namespace paladin {

class SpectralDecomposition {
  std::vector<std::complex<double>> eigenvalues_;
  std::vector<std::vector<std::complex<double>>> rightEigenvectors_, leftEigenvectors_;
  double vectorWriteThreshold_;

 public:
  explicit SpectralDecomposition( int n = 0, double threshold = 1.0e-3 ) : vectorWriteThreshold_( threshold ) {
    eigenvalues_.reserve( n );
  }

  int size() const { return static_cast<int>( eigenvalues_.size() ); }
  const std::vector<std::complex<double>>& eigenvalues() const { return eigenvalues_; }
  const std::vector<std::vector<std::complex<double>>>& right_eigenvectors() const { return rightEigenvectors_; }

  // wr, wi: n values; vr / vl: n x n column-major real-packed vectors or nullptr. The reference pairs the left
  // vectors exactly like the right ones: (re, +im) for the first eigenvalue of a pair, (re, -im) for its conjugate
  // (src/paladin.hpp:265-274).
  void unpack( int n, const double* wr, const double* wi, const double* vr, const double* vl = nullptr ) {
    eigenvalues_.clear();
    rightEigenvectors_.clear();
    leftEigenvectors_.clear();
    auto column = [&]( int c, int imag_col, double imag_sign ) {
      for ( int side = 0; side < 2; ++side ) {
        const double* src = side == 0 ? vr : vl;
        if ( src == nullptr ) continue;
        std::vector<std::complex<double>> v( n );
        for ( int r = 0; r < n; ++r ) {
          const double re = src[r + static_cast<std::size_t>( n ) * c];
          const double im = imag_col < 0 ? 0.0 : imag_sign * src[r + static_cast<std::size_t>( n ) * imag_col];
          v[r] = std::complex<double>( re, im );
        }
        ( side == 0 ? rightEigenvectors_ : leftEigenvectors_ ).push_back( std::move( v ) );
      }
    };
    bool second_of_pair = false;
    for ( int c = 0; c < n; ++c ) {
      eigenvalues_.emplace_back( wr[c], wi[c] );
      if ( second_of_pair ) {
        second_of_pair = false;
        continue;  // vector already emitted together with its conjugate
      }
      const bool pair = ( c + 1 < n ) && wi[c] == -wi[c + 1] && wi[c] != 0.0;
      if ( vr != nullptr || vl != nullptr ) {
        if ( pair ) {
          column( c, c + 1, 1.0 );
          column( c, c + 1, -1.0 );
        } else {
          column( c, -1, 0.0 );
        }
      }
      second_of_pair = pair;
    }
  }

  // "%g" is what operator<<(double) prints at the default precision of 6 (the reference never changes it)
  // std::to_chars( general, 6 ) is specified to give printf's "%.6g" characters and is about twice as fast as snprintf
  // (compared on 2e7 random, denormal, short-decimal and tie values); non-finite values keep the printf spelling.
  static void append_number( std::string& out, double x ) {
    char buf[40];
    if ( std::isfinite( x ) ) {
      const std::to_chars_result r = std::to_chars( buf, buf + sizeof( buf ), x, std::chars_format::general, 6 );
      out.append( buf, static_cast<std::size_t>( r.ptr - buf ) );
      return;
    }
    const int len = std::snprintf( buf, sizeof( buf ), "%g", x );
    out.append( buf, static_cast<std::size_t>( len ) );
  }
  static void append_number( std::string& out, long x ) {
    char buf[24];
    const std::to_chars_result r = std::to_chars( buf, buf + sizeof( buf ), x );
    out.append( buf, static_cast<std::size_t>( r.ptr - buf ) );
  }

  void write_eigenvalues( std::string& out ) const {
    for ( const auto& ev : eigenvalues_ ) {
      append_number( out, ev.real() );
      out += ' ';
      append_number( out, ev.imag() );
      out += '\n';
    }
  }

  // max2 / kept: the filter pass of the writer (largest |x|^2 per vector, number of entries kept) when it has already
  // been evaluated on the device (paladin_gpu_batch_vector_filter); nullptr = evaluate it here
  void write_right_eigenvectors( std::string& out, const double* max2 = nullptr, long long kept = 0 ) const {
    write_vectors( out, rightEigenvectors_, max2, kept );
  }
  void write_left_eigenvectors( std::string& out, const double* max2 = nullptr, long long kept = 0 ) const {
    write_vectors( out, leftEigenvectors_, max2, kept );
  }

 private:
  void write_vectors( std::string& out, const std::vector<std::vector<std::complex<double>>>& vecs, const double* max2,
                      long long kept_in ) const {
    const int n = size();
    std::vector<double> biggest( vecs.size(), 0.0 );
    long kept = 0;
    if ( max2 != nullptr && !vecs.empty() ) {
      for ( std::size_t v = 0; v < vecs.size(); ++v ) biggest[v] = max2[v];
      kept = static_cast<long>( kept_in );
    } else {
      for ( std::size_t v = 0; v < vecs.size(); ++v ) {
        for ( const auto& x : vecs[v] ) biggest[v] = std::norm( x ) > biggest[v] ? std::norm( x ) : biggest[v];
        for ( const auto& x : vecs[v] )
          if ( std::norm( x ) > vectorWriteThreshold_ * biggest[v] ) ++kept;
      }
    }
    out += "%%MatrixMarket matrix coordinate complex general\n";
    append_number( out, static_cast<long>( n ) );
    out += ' ';
    append_number( out, static_cast<long>( n ) );
    out += ' ';
    append_number( out, kept );
    out += '\n';
    for ( std::size_t v = 0; v < vecs.size(); ++v )
      for ( int r = 0; r < n; ++r ) {
        const std::complex<double>& x = vecs[v][r];
        if ( std::norm( x ) > vectorWriteThreshold_ * biggest[v] ) {
          append_number( out, static_cast<long>( r + 1 ) );
          out += ' ';
          append_number( out, static_cast<long>( v + 1 ) );
          out += ' ';
          append_number( out, x.real() );
          out += ' ';
          append_number( out, x.imag() );
          out += '\n';
        }
      }
  }
};

}  // namespace paladin

#endif
