/*
 * shim_mpirun -- TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product.
 *
 * "mpirun -np P prog args..." for binaries compiled against oracle/mpi_shim/mpi.h.
 * Forks P copies of prog; rank 0 is wired to every leaf rank by a pair of pipes
 * whose descriptors are published through the environment (see mpi.h).
 * Replaces the `mpirun -np N` of the reference's ctest harness
 * (reference test/compare_spectra.cmake:8-13) in an image without MPI.
 */
#include <sys/wait.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main( int argc, char** argv ) {
  if ( argc < 4 || std::strcmp( argv[1], "-np" ) != 0 ) {
    std::fprintf( stderr, "usage: %s -np P prog [args...]\n", argv[0] );
    return 2;
  }
  const int np = std::atoi( argv[2] );
  if ( np < 1 ) {
    std::fprintf( stderr, "shim_mpirun: bad rank count\n" );
    return 2;
  }
  // down[leaf] : root -> leaf, up[leaf] : leaf -> root
  std::vector<int> downR( np, -1 ), downW( np, -1 ), upR( np, -1 ), upW( np, -1 );
  for ( int leaf = 1; leaf < np; ++leaf ) {
    int d[2], u[2];
    if ( pipe( d ) != 0 || pipe( u ) != 0 ) {
      std::perror( "pipe" );
      return 2;
    }
    downR[leaf] = d[0];
    downW[leaf] = d[1];
    upR[leaf] = u[0];
    upW[leaf] = u[1];
  }
  std::string rootFds;
  for ( int leaf = 1; leaf < np; ++leaf ) {
    if ( leaf > 1 ) rootFds += ",";
    rootFds += std::to_string( upR[leaf] ) + ":" + std::to_string( downW[leaf] );
  }

  std::vector<pid_t> kids;
  for ( int rank = 0; rank < np; ++rank ) {
    pid_t pid = fork();
    if ( pid < 0 ) {
      std::perror( "fork" );
      return 2;
    }
    if ( pid == 0 ) {
      setenv( "PALADIN_SHIM_SIZE", std::to_string( np ).c_str(), 1 );
      setenv( "PALADIN_SHIM_RANK", std::to_string( rank ).c_str(), 1 );
      if ( rank == 0 ) {
        setenv( "PALADIN_SHIM_FDS", rootFds.c_str(), 1 );
        for ( int leaf = 1; leaf < np; ++leaf ) {
          close( downR[leaf] );
          close( upW[leaf] );
        }
      } else {
        setenv( "PALADIN_SHIM_FD_R", std::to_string( downR[rank] ).c_str(), 1 );
        setenv( "PALADIN_SHIM_FD_W", std::to_string( upW[rank] ).c_str(), 1 );
        for ( int leaf = 1; leaf < np; ++leaf ) {
          close( downW[leaf] );
          close( upR[leaf] );
          if ( leaf != rank ) {
            close( downR[leaf] );
            close( upW[leaf] );
          }
        }
      }
      execvp( argv[3], argv + 3 );
      std::perror( "execvp" );
      _exit( 127 );
    }
    kids.push_back( pid );
  }
  for ( int leaf = 1; leaf < np; ++leaf ) {
    close( downR[leaf] );
    close( downW[leaf] );
    close( upR[leaf] );
    close( upW[leaf] );
  }
  int worst = 0;
  for ( pid_t k : kids ) {
    int status = 0;
    waitpid( k, &status, 0 );
    int code = WIFEXITED( status ) ? WEXITSTATUS( status ) : 128 + WTERMSIG( status );
    if ( code > worst ) worst = code;
  }
  return worst;
}
