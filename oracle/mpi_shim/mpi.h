/*
 * mpi.h -- TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product.
 *
 * A tiny stand-in for the handful of MPI entry points the reference paladin
 * headers name (reference src/mpi-util.hpp:36-72, src/paladin.hpp:159-210 and
 * :391-399), so that the UNMODIFIED reference sources under /root/reference
 * compile and run in an image that has no MPI installation.
 *
 * Two modes, selected by environment variables at MPI_Init time:
 *
 *   (1) real multi-process mode: launched by oracle/mpi_shim/shim_mpirun.cpp,
 *       which forks P ranks connected to rank 0 by pipes and exports
 *           PALADIN_SHIM_SIZE, PALADIN_SHIM_RANK,
 *           PALADIN_SHIM_FDS   (rank 0 : "r1:w1,r2:w2,..." one pair per leaf)
 *           PALADIN_SHIM_FD_R / PALADIN_SHIM_FD_W (leaf: from-root / to-root).
 *       All communication in the reference is root<->leaf and issued in the
 *       same program order on every rank, so FIFO pipes are sufficient.
 *
 *   (2) fake-size mode: PALADIN_SHIM_FAKE_SIZE=P with a single process.
 *       MPI_Comm_size reports P, rank is 0 and every communication call is a
 *       no-op. Rank 0 of the reference computes the whole partition
 *       (src/paladin.hpp:149-155) without needing the leaves, which is what
 *       the partition oracle (oracle/ref_partition_dump.cpp) relies on.
 *
 *   With neither set: one rank, size 1.
 */
#ifndef PALADIN_ORACLE_MPI_SHIM_H_
#define PALADIN_ORACLE_MPI_SHIM_H_

#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
struct MPI_Status {
  int count_bytes;
};

#define MPI_COMM_WORLD 0
#define MPI_CHAR 1
#define MPI_INT 4
#define MPI_DOUBLE 8
#define MPI_MIN 100
#define MPI_MAX 101
#define MPI_SUM 102
#define MPI_STATUS_IGNORE ( (MPI_Status*)0 )
#define MPI_SUCCESS 0

namespace paladin_shim {

struct State {
  int size = 1, rank = 0;
  bool fake = false;
  std::vector<int> from_leaf, to_leaf;  // rank 0: indexed by leaf rank (entry 0 unused)
  int from_root = -1, to_root = -1;     // leaves
  long pending_len = -1;                // leaf: length header already consumed by MPI_Probe
};

inline State& st() {
  static State s;
  return s;
}

inline void die( const char* what ) {
  std::fprintf( stderr, "[mpi shim] fatal: %s\n", what );
  std::_Exit( 3 );
}

inline void wr( int fd, const void* buf, size_t n ) {
  const char* p = static_cast<const char*>( buf );
  while ( n > 0 ) {
    ssize_t k = ::write( fd, p, n );
    if ( k <= 0 ) die( "pipe write" );
    p += k;
    n -= static_cast<size_t>( k );
  }
}

inline void rd( int fd, void* buf, size_t n ) {
  char* p = static_cast<char*>( buf );
  while ( n > 0 ) {
    ssize_t k = ::read( fd, p, n );
    if ( k <= 0 ) die( "pipe read" );
    p += k;
    n -= static_cast<size_t>( k );
  }
}

inline size_t type_bytes( MPI_Datatype t ) { return static_cast<size_t>( t ); }

inline void send_to_leaf( const void* buf, int count, int dest ) {
  State& s = st();
  if ( s.fake || s.size == 1 ) return;
  long len = count;
  wr( s.to_leaf[dest], &len, sizeof( len ) );
  wr( s.to_leaf[dest], buf, static_cast<size_t>( count ) );
}

}  // namespace paladin_shim

inline int MPI_Init( int*, char*** ) {
  using namespace paladin_shim;
  State& s = st();
  if ( const char* f = std::getenv( "PALADIN_SHIM_FAKE_SIZE" ) ) {
    s.fake = true;
    s.size = std::atoi( f );
    s.rank = 0;
    return MPI_SUCCESS;
  }
  const char* sz = std::getenv( "PALADIN_SHIM_SIZE" );
  const char* rk = std::getenv( "PALADIN_SHIM_RANK" );
  if ( !sz || !rk ) return MPI_SUCCESS;
  s.size = std::atoi( sz );
  s.rank = std::atoi( rk );
  if ( s.size == 1 ) return MPI_SUCCESS;
  if ( s.rank == 0 ) {
    s.from_leaf.assign( s.size, -1 );
    s.to_leaf.assign( s.size, -1 );
    const char* fds = std::getenv( "PALADIN_SHIM_FDS" );
    if ( !fds ) die( "PALADIN_SHIM_FDS missing on rank 0" );
    std::string all( fds );
    size_t pos = 0;
    for ( int leaf = 1; leaf < s.size; ++leaf ) {
      size_t comma = all.find( ',', pos );
      std::string pair = all.substr( pos, comma == std::string::npos ? std::string::npos : comma - pos );
      size_t colon = pair.find( ':' );
      if ( colon == std::string::npos ) die( "bad PALADIN_SHIM_FDS" );
      s.from_leaf[leaf] = std::atoi( pair.substr( 0, colon ).c_str() );
      s.to_leaf[leaf] = std::atoi( pair.substr( colon + 1 ).c_str() );
      pos = ( comma == std::string::npos ) ? all.size() : comma + 1;
    }
  } else {
    const char* r = std::getenv( "PALADIN_SHIM_FD_R" );
    const char* w = std::getenv( "PALADIN_SHIM_FD_W" );
    if ( !r || !w ) die( "leaf pipe fds missing" );
    s.from_root = std::atoi( r );
    s.to_root = std::atoi( w );
  }
  return MPI_SUCCESS;
}

inline int MPI_Finalize() { return MPI_SUCCESS; }

inline int MPI_Comm_size( MPI_Comm, int* size ) {
  *size = paladin_shim::st().size;
  return MPI_SUCCESS;
}

inline int MPI_Comm_rank( MPI_Comm, int* rank ) {
  *rank = paladin_shim::st().rank;
  return MPI_SUCCESS;
}

inline int MPI_Barrier( MPI_Comm ) {
  using namespace paladin_shim;
  State& s = st();
  if ( s.fake || s.size == 1 ) return MPI_SUCCESS;
  char token = 'b';
  if ( s.rank == 0 ) {
    for ( int leaf = 1; leaf < s.size; ++leaf ) rd( s.from_leaf[leaf], &token, 1 );
    for ( int leaf = 1; leaf < s.size; ++leaf ) wr( s.to_leaf[leaf], &token, 1 );
  } else {
    wr( s.to_root, &token, 1 );
    rd( s.from_root, &token, 1 );
  }
  return MPI_SUCCESS;
}

/* root is always 0 in the reference */
inline int MPI_Bcast( void* buf, int count, MPI_Datatype type, int /*root*/, MPI_Comm ) {
  using namespace paladin_shim;
  State& s = st();
  if ( s.fake || s.size == 1 ) return MPI_SUCCESS;
  const size_t nbytes = static_cast<size_t>( count ) * type_bytes( type );
  if ( s.rank == 0 ) {
    for ( int leaf = 1; leaf < s.size; ++leaf ) wr( s.to_leaf[leaf], buf, nbytes );
  } else {
    rd( s.from_root, buf, nbytes );
  }
  return MPI_SUCCESS;
}

inline int MPI_Reduce( const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op,
                       int /*root*/, MPI_Comm ) {
  using namespace paladin_shim;
  State& s = st();
  if ( type != MPI_DOUBLE ) die( "MPI_Reduce: only MPI_DOUBLE is used by the reference" );
  const double* in = static_cast<const double*>( sendbuf );
  if ( s.fake || s.size == 1 ) {
    std::memcpy( recvbuf, sendbuf, sizeof( double ) * static_cast<size_t>( count ) );
    return MPI_SUCCESS;
  }
  if ( s.rank == 0 ) {
    double* out = static_cast<double*>( recvbuf );
    for ( int i = 0; i < count; ++i ) out[i] = in[i];
    std::vector<double> tmp( static_cast<size_t>( count ) );
    for ( int leaf = 1; leaf < s.size; ++leaf ) {
      rd( s.from_leaf[leaf], tmp.data(), sizeof( double ) * tmp.size() );
      for ( int i = 0; i < count; ++i ) {
        if ( op == MPI_MIN ) out[i] = tmp[i] < out[i] ? tmp[i] : out[i];
        if ( op == MPI_MAX ) out[i] = tmp[i] > out[i] ? tmp[i] : out[i];
        if ( op == MPI_SUM ) out[i] += tmp[i];
      }
    }
  } else {
    wr( s.to_root, in, sizeof( double ) * static_cast<size_t>( count ) );
  }
  return MPI_SUCCESS;
}

inline int MPI_Probe( int /*source*/, int /*tag*/, MPI_Comm, MPI_Status* status ) {
  using namespace paladin_shim;
  State& s = st();
  if ( s.pending_len < 0 ) {
    long len = 0;
    rd( s.from_root, &len, sizeof( len ) );
    s.pending_len = len;
  }
  if ( status ) status->count_bytes = static_cast<int>( s.pending_len );
  return MPI_SUCCESS;
}

inline int MPI_Get_count( const MPI_Status* status, MPI_Datatype type, int* count ) {
  *count = static_cast<int>( static_cast<size_t>( status->count_bytes ) / paladin_shim::type_bytes( type ) );
  return MPI_SUCCESS;
}

inline int MPI_Recv( void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm,
                     MPI_Status* status ) {
  using namespace paladin_shim;
  State& s = st();
  if ( s.pending_len < 0 ) MPI_Probe( source, tag, comm, status );
  const size_t want = static_cast<size_t>( count ) * type_bytes( type );
  if ( static_cast<size_t>( s.pending_len ) != want ) die( "MPI_Recv: size mismatch" );
  rd( s.from_root, buf, want );
  s.pending_len = -1;
  return MPI_SUCCESS;
}

/* the removed-in-MPI-3 C++ binding used at reference src/mpi-util.hpp:60 */
namespace MPI {
struct Datatype {
  int bytes;
};
static const Datatype CHAR = {1};
struct Intracomm {
  void Send( const void* buf, int count, const Datatype& type, int dest, int /*tag*/ ) const {
    paladin_shim::send_to_leaf( buf, count * type.bytes, dest );
  }
};
static const Intracomm COMM_WORLD = {};
}  // namespace MPI

#endif /* PALADIN_ORACLE_MPI_SHIM_H_ */
