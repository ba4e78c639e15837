/*
 * ref_probe -- TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product.
 *
 * Thin driver around the UNMODIFIED reference headers (included from
 * /root/reference/src at build time, never copied) that exposes two pieces of
 * reference state the parity tests need bit-exactly:
 *
 *   ref_probe partition <paladin args...>
 *       With PALADIN_SHIM_FAKE_SIZE=P in the environment, prints the reference's
 *       sorted path list and pod colouring (protected members of
 *       paladin::Paladin, reference src/paladin.hpp:58-60, filled at :141-155).
 *       Output lines after the marker "@@PARTITION": "<color> <path>".
 *
 *   ref_probe dense <matrix-file> <out.bin>
 *       Runs reference read_matrix_from_mm_file (src/square-matrix.hpp:170-180)
 *       and dumps int32 n, int32 nnz, then n*n float64 column-major.
 */
#include <paladin.hpp>

#include <cstdint>
#include <cstdio>
#include <cstring>

namespace {

class PartitionPeek : public paladin::Paladin {
 public:
  PartitionPeek( int& argc, char* argv[], const paladin::MpiComm& comm ) : paladin::Paladin( argc, argv, comm ) {}
  void dump() const {
    std::cout << "@@PARTITION " << allMatrixPaths.size() << '\n';
    for ( std::size_t i = 0; i < allMatrixPaths.size(); ++i ) {
      std::cout << podColors[i] << ' ' << allMatrixPaths[i] << '\n';
    }
  }
};

}  // namespace

int main( int argc, char* argv[] ) {
  if ( argc < 2 ) {
    std::fprintf( stderr, "usage: ref_probe partition|dense ...\n" );
    return 2;
  }
  const std::string mode = argv[1];
  if ( mode == "partition" ) {
    int subArgc = argc - 1;
    char** subArgv = argv + 1;  // subArgv[0] plays the role of the program name
    paladin::MpiComm comm( subArgc, subArgv );
    PartitionPeek peek( subArgc, subArgv, comm );
    peek.dump();
    return 0;
  }
  if ( mode == "dense" && argc == 4 ) {
    paladin::SquareMatrix A = paladin::read_matrix_from_mm_file( argv[2] );
    FILE* f = std::fopen( argv[3], "wb" );
    if ( !f ) return 1;
    std::int32_t hdr[2] = {A.nrows_, A.nnz_};
    std::fwrite( hdr, sizeof( std::int32_t ), 2, f );
    std::fwrite( A.mat_.data(), sizeof( double ), A.mat_.size(), f );
    std::fclose( f );
    return 0;
  }
  std::fprintf( stderr, "ref_probe: bad arguments\n" );
  return 2;
}
