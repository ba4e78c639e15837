// paladin_host.cpp -- host side of the GPU path with the reference's command line, listing format, load
// measures, static balancer, unpack logic, output formats and timing report
// (reference src/paladin.hpp:57-421, src/compute-write-with-timings.cpp:29-36), calling libpaladin_gpu.so
// through include/paladin_gpu.h instead of dgeev_.
//
// A "pod" is an MPI rank in the reference and a GPU here: the balancer runs with numRanks = number of GPUs, every
// pod is driven by one host thread, pods never exchange data (the reference's compute() has no communication
// either). Extra keys: --gpus=N (default: all usable devices), --batch=B (matrices per library call),
// --inflight=K (batches of one GPU in flight at once, default 2), --dump-partition=FILE (write the sorted listing with pod colours and exit; needs no GPU).
//
// Built twice: as the executable paladin-gpu and (with -DPALADIN_HOST_LIBRARY) as libpaladin_host.so, whose
// extern "C" hooks let tests/ drive the parser, balancer and writers without a GPU.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <future>
#include <iostream>
#include <deque>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/paladin_gpu.h"
#include "balance.hpp"
#include "cli.hpp"
#include "mm_coo.hpp"
#include "spectrum_io.hpp"

namespace paladin {

struct Options {
  std::string listing = "no matrices given!", rootdir = ".", dump_partition;
  int repeats = 1, gpus = -1, batch = 256;
  int inflight = 2;  // --inflight=K: batches of one GPU whose library calls / writers may overlap (1 = strictly one after the other)
  bool showdist = false, right = false, left = false;
  bool dynamic = false;  // --dynamic: GPUs pull batches of the sorted list from a shared queue (default: the reference's static pods)
  LoadPredictor measure = LoadPredictor::NUMNONZEROS;
  double threshold = 1.0e-3;
};

struct Plan {
  std::vector<std::string> paths;  // sorted by measure, descending
  std::vector<double> measures;
  std::vector<int> colour;         // pod of each sorted entry
};

inline void print_header() {
  std::cout << "\n\npaladin-gpu: many independent FP64 eigendecompositions on B200 (sm_100a)\n"
            << "drop-in for the paladin executable (listing, --measure, --left/--right, output files unchanged)\n"
            << "library: " << paladin_gpu_version() << "\n\n";
}

inline Options parse_options( int argc, char* argv[], bool quiet ) {
  CommandLineParser clp( argc, argv );
  Options o;
  o.listing = clp.getValue( "listing", o.listing );
  o.rootdir = clp.getValue( "rootdir", "." );
  o.repeats = std::stoi( clp.getValue( "repeats", "1" ) );
  o.showdist = clp.checkExists( "showdist" );
  o.right = clp.checkExists( "right" );
  o.left = clp.checkExists( "left" );
  o.measure = string_to_measure_type( clp.getValue( "measure", "nnz" ) );
  o.threshold = std::stod( clp.getValue( "write-threshold", "1e-3" ) );
  o.gpus = std::stoi( clp.getValue( "gpus", "-1" ) );
  o.batch = std::max( 1, std::stoi( clp.getValue( "batch", "256" ) ) );
  o.inflight = std::max( 1, std::stoi( clp.getValue( "inflight", "2" ) ) );
  o.dump_partition = clp.getValue( "dump-partition", "" );
  o.dynamic = clp.checkExists( "dynamic" );
  if ( !quiet ) {
    bool bad = false;
    for ( const auto& k : clp.get_keys() ) {
      static const char* known[] = { "listing", "repeats", "showdist", "measure", "rootdir", "left", "right",
                                     "write-threshold", "gpus", "batch", "dump-partition", "dynamic", "inflight" };
      if ( std::none_of( std::begin( known ), std::end( known ), [&]( const char* s ) { return k == s; } ) ) {
        std::cout << "\n-- POTENTIAL ERROR: key " << k << " is not a recognized option!";
        bad = true;
      }
    }
    if ( bad )
      std::cout << "\n-- Allowable keys: listing, repeats, showdist, measure, rootdir, left, right, write-threshold, "
                   "gpus, batch, dump-partition, dynamic, inflight\n\n";
  }
  return o;
}

// listing -> measures -> std::sort -> greedy colouring (reference src/paladin.hpp:123-155)
inline Plan make_plan( const Options& o, int pods ) {
  std::ifstream listing( o.listing );
  if ( !listing ) {
    std::cerr << "Couldn't open the matrix listing file, " << o.listing << ", exiting!" << '\n';
    std::exit( 1 );
  }
  std::vector<NameMeasurePairT> entries;
  std::string line;
  while ( std::getline( listing, line ) && !line.empty() ) {  // stops at the first empty line, like the reference
    const std::string path = o.rootdir + "/" + line;
    int n = 0, nnz = 0;
    try {
      read_matrix_sizes_from_mm_file( path, n, nnz );
    } catch ( const std::exception& e ) {
      std::cerr << e.what() << ", exiting!" << '\n';
      std::exit( 1 );
    }
    entries.emplace_back( path, matrix_measure( n, nnz, o.measure ) );
  }
  sort_name_measure_pairs( entries );
  Plan p;
  p.colour = color_pods( entries, pods );
  for ( const auto& e : entries ) {
    p.paths.push_back( e.first );
    p.measures.push_back( e.second );
  }
  return p;
}

// --dynamic: consecutive batches of the sorted list, (first, count)
inline std::vector<std::pair<std::size_t, std::size_t>> dynamic_chunks( std::size_t total, std::size_t batch ) {
  std::vector<std::pair<std::size_t, std::size_t>> chunks;
  for ( std::size_t first = 0; first < total; first += batch ) chunks.emplace_back( first, std::min( batch, total - first ) );
  return chunks;
}

inline void write_file( const std::string& path, const std::string& text ) {
  std::FILE* f = std::fopen( path.c_str(), "wb" );
  if ( !f ) {
    std::cerr << "Couldn't open the output file, " << path << '\n';
    return;
  }
  std::fwrite( text.data(), 1, text.size(), f );
  std::fclose( f );
}

struct PodTimes {
  double read = 0.0, eig = 0.0, total = 0.0;
};

// A parsed batch: the COO triplets of `count` consecutive pod entries, concatenated in listing order.
struct ParsedBatch {
  std::size_t first = 0, count = 0;
  std::vector<int> n;
  std::vector<int64_t> off;
  std::vector<int> ci, cj;
  std::vector<double> cv;
  double seconds = 0.0;  // wall time of the parse
};

// Files of one batch are parsed by `threads` workers (the reference spreads this work over its MPI ranks; here the
// host cores not driving a GPU do it).
inline ParsedBatch parse_batch( const std::vector<std::string>& pod, std::size_t first, std::size_t count, int threads ) {
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<CooMatrix> mats( count );
  std::vector<std::string> errors( count );
  // files are handed out one at a time (their sizes differ), and the concatenation of the triplets -- hundreds of MB
  // per batch -- is spread over the same workers instead of being a serial tail
  const std::size_t nthreads = std::max<std::size_t>( 1, std::min<std::size_t>( threads, count ) );
  auto run_on_pool = [&]( const std::function<void( std::size_t )>& item ) {
    std::atomic<std::size_t> cursor{ 0 };
    auto loop = [&]() {
      for ( std::size_t k = cursor.fetch_add( 1 ); k < count; k = cursor.fetch_add( 1 ) ) item( k );
    };
    std::vector<std::thread> pool;
    for ( std::size_t w = 1; w < nthreads; ++w ) pool.emplace_back( loop );
    loop();
    for ( auto& th : pool ) th.join();
  };
  run_on_pool( [&]( std::size_t k ) {
    try {
      mats[k] = read_coo_from_mm_file( pod[first + k] );
    } catch ( const std::exception& e ) {
      errors[k] = e.what();
    }
  } );
  for ( const auto& e : errors )
    if ( !e.empty() ) {
      std::cerr << e << ", exiting!" << '\n';
      std::exit( 1 );
    }
  ParsedBatch b;
  b.first = first;
  b.count = count;
  b.n.resize( count );
  b.off.assign( count + 1, 0 );
  for ( std::size_t k = 0; k < count; ++k ) {
    b.n[k] = mats[k].nrows_;
    b.off[k + 1] = b.off[k] + mats[k].nnz_;
  }
  b.ci.resize( b.off[count] );
  b.cj.resize( b.off[count] );
  b.cv.resize( b.off[count] );
  run_on_pool( [&]( std::size_t k ) {
    std::copy( mats[k].row.begin(), mats[k].row.end(), b.ci.begin() + b.off[k] );
    std::copy( mats[k].col.begin(), mats[k].col.end(), b.cj.begin() + b.off[k] );
    std::copy( mats[k].val.begin(), mats[k].val.end(), b.cv.begin() + b.off[k] );
    CooMatrix().row.swap( mats[k].row );
    CooMatrix().col.swap( mats[k].col );
    CooMatrix().val.swap( mats[k].val );
  } );
  b.seconds = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
  return b;
}

// One pod = one GPU: parse (next batch, in the background) -> batched library call -> unpack + write. Up to
// --inflight batches are between their library call and the end of their writer at any time, each on a host thread of
// its own: every batch owns a CUDA stream (include/paladin_gpu.h "Thread model"), so the copies and the HBM-bound
// Hessenberg kernels of one batch run under the latency-bound QR kernels of another, and the text
// writers of finished batches run beside both. The files do not depend on the schedule (one file set per matrix,
// every matrix solved on its own). "read" = time spent parsing, "eig." = time inside the library, as in the
// reference's table; because the stages overlap their sum can exceed "total".
inline int run_pod( int device, const std::vector<std::string>& pod, const Options& o, PodTimes& times, int parse_threads ) {
  using clock = std::chrono::steady_clock;
  const auto t_start = clock::now();
  const char jobvr = o.right ? 'V' : 'N';
  const char jobvl = o.left ? 'V' : 'N';
  int failures = 0;
  bool gpu_failed = false;
  std::mutex mu;  // times.eig, std::cerr
  std::vector<std::size_t> firsts;
  for ( std::size_t first = 0; first < pod.size(); first += o.batch ) firsts.push_back( first );

  // library call + writers of one batch; returns the number of matrices with info != 0, -1 when the library failed
  auto solve_and_write = [&]( std::shared_ptr<ParsedBatch> curp, bool write ) -> int {
    const ParsedBatch& cur = *curp;
    const std::size_t count = cur.count, first = cur.first;
    // results of the whole batch, concatenated (the staged C-ABI): eigenvalue offsets woff, vector offsets voff
    std::vector<std::size_t> woff( count + 1, 0 ), voff( count + 1, 0 );
    for ( std::size_t k = 0; k < count; ++k ) {
      const std::size_t nk = static_cast<std::size_t>( cur.n[k] );
      woff[k + 1] = woff[k] + nk;
      voff[k + 1] = voff[k] + nk * nk;
    }
    const std::size_t tn = woff[count], tnn = voff[count];
    std::vector<double> wr( tn, 0.0 ), wi( tn, 0.0 ), vr( o.right ? tnn : 0, 0.0 ), vl( o.left ? tnn : 0, 0.0 );
    // write filter of the eigenvector files, evaluated on the device (paladin_gpu_batch_vector_filter)
    std::vector<double> max2r( o.right ? tn : 0, 0.0 ), max2l( o.left ? tn : 0, 0.0 );
    std::vector<long long> keptr( o.right ? count : 0, 0 ), keptl( o.left ? count : 0, 0 );
    std::vector<int> info( count, 0 );
    const auto t_eig = clock::now();
    paladin_gpu_batch* gb = nullptr;
    int rc = paladin_gpu_batch_create( device, static_cast<int>( count ), cur.n.data(), cur.off.data(), jobvl, jobvr, &gb );
    if ( rc == PALADIN_GPU_OK ) rc = paladin_gpu_batch_upload( gb, cur.ci.data() + cur.off[0], cur.cj.data() + cur.off[0], cur.cv.data() + cur.off[0] );
    if ( rc == PALADIN_GPU_OK ) rc = paladin_gpu_batch_scatter( gb );
    if ( rc == PALADIN_GPU_OK ) rc = paladin_gpu_batch_solve( gb );
    if ( rc == PALADIN_GPU_OK && o.right && count > 0 ) rc = paladin_gpu_batch_vector_filter( gb, 'R', o.threshold, max2r.data(), keptr.data() );
    if ( rc == PALADIN_GPU_OK && o.left && count > 0 ) rc = paladin_gpu_batch_vector_filter( gb, 'L', o.threshold, max2l.data(), keptl.data() );
    if ( rc == PALADIN_GPU_OK )
      rc = paladin_gpu_batch_download( gb, wr.data(), wi.data(), o.left ? vl.data() : nullptr, o.right ? vr.data() : nullptr, info.data() );
    const std::string why = rc == PALADIN_GPU_OK ? std::string() : std::string( paladin_gpu_last_error() );  // per-thread message
    paladin_gpu_batch_destroy( gb );
    {
      std::lock_guard<std::mutex> lock( mu );
      times.eig += std::chrono::duration<double>( clock::now() - t_eig ).count();
      if ( rc != PALADIN_GPU_OK ) std::cerr << "the GPU eigensolver failed on device " << device << ": " << why << '\n';
    }
    if ( rc != PALADIN_GPU_OK ) return -1;
    if ( !write ) return 0;  // like the reference, only the first repeat produces output
    // unpack + text formatting: one matrix per worker at a time (the reference formats on all of its ranks at once; with
    // vectors this stage costs far more host time per matrix than the GPU spends on the decomposition)
    std::atomic<int> bad{ 0 };
    std::atomic<std::size_t> cursor{ 0 };
    auto write_some = [&]() {
    for ( std::size_t k = cursor.fetch_add( 1 ); k < count; k = cursor.fetch_add( 1 ) ) {
      const int nk = cur.n[k];
      const bool ok = info[k] == 0;
      if ( !ok ) {
        ++bad;
        std::lock_guard<std::mutex> lock( mu );
        std::cerr << "warning: " << pod[first + k] << ": info = " << info[k]
                  << ( info[k] < 0 ? " (input rejected, no output files written)" : " (QR iteration did not converge, eigenvalues before that index are NaN)" ) << '\n';
      }
      // info < 0: the library refused the input (COO index outside [1, n], non-finite entry) and computed nothing --
      // there is no spectrum to write; the run ends with a non-zero exit status
      if ( info[k] < 0 ) continue;
      SpectralDecomposition s( nk, o.threshold );
      s.unpack( nk, wr.data() + woff[k], wi.data() + woff[k], o.right && ok ? vr.data() + voff[k] : nullptr,
                o.left && ok ? vl.data() + voff[k] : nullptr );
      std::string text;
      s.write_eigenvalues( text );
      write_file( pod[first + k] + ".spectrum", text );
      if ( o.left ) {
        text.clear();
        if ( ok ) s.write_left_eigenvectors( text, max2l.data() + woff[k], keptl[k] );
        else s.write_left_eigenvectors( text );
        write_file( pod[first + k] + ".leftvecs", text );
      }
      if ( o.right ) {
        text.clear();
        if ( ok ) s.write_right_eigenvectors( text, max2r.data() + woff[k], keptr[k] );
        else s.write_right_eigenvectors( text );
        write_file( pod[first + k] + ".rightvecs", text );
      }
    }
    };
    std::vector<std::thread> writers;
    const std::size_t nw = std::min<std::size_t>( static_cast<std::size_t>( std::max( 1, parse_threads ) ), count );
    for ( std::size_t w = 1; w < nw; ++w ) writers.emplace_back( write_some );
    write_some();
    for ( auto& w : writers ) w.join();
    return bad.load();
  };

  std::deque<std::future<int>> flying;
  auto land = [&]( std::size_t keep ) {
    while ( flying.size() > keep ) {
      const int r = flying.front().get();
      flying.pop_front();
      if ( r < 0 ) gpu_failed = true;
      else failures += r;
    }
  };
  for ( int rep = 0; rep < o.repeats && !gpu_failed; ++rep ) {
    std::future<ParsedBatch> next;
    if ( !firsts.empty() )
      next = std::async( std::launch::async, parse_batch, std::cref( pod ), firsts[0], std::min<std::size_t>( o.batch, pod.size() ), parse_threads );
    for ( std::size_t bi = 0; bi < firsts.size(); ++bi ) {
      auto cur = std::make_shared<ParsedBatch>( next.get() );
      times.read += cur->seconds;
      if ( bi + 1 < firsts.size() )
        next = std::async( std::launch::async, parse_batch, std::cref( pod ), firsts[bi + 1],
                           std::min<std::size_t>( o.batch, pod.size() - firsts[bi + 1] ), parse_threads );
      land( static_cast<std::size_t>( o.inflight ) - 1 );
      if ( gpu_failed ) {
        if ( next.valid() ) next.wait();
        break;
      }
      flying.push_back( std::async( std::launch::async, solve_and_write, cur, rep == 0 ) );
    }
    land( 0 );  // a repeat is timed as a whole
  }
  land( 0 );
  if ( gpu_failed ) return -1;
  times.total = std::chrono::duration<double>( clock::now() - t_start ).count() / o.repeats;
  times.read /= o.repeats;
  times.eig /= o.repeats;
  return failures;
}

inline void display_timings( const std::vector<PodTimes>& t, int repeats ) {
  std::cout << '\n' << "Timings averaged over " << repeats << " runs:" << '\n';
  std::cout << "                      runtimes (s)\t\t\t  fractions\n";
  std::cout << "        -----------------------------------------\t-------------\n";
  std::cout << "pod\tread\t\teig.\t\ttotal\t\tread\teig.\n";
  PodTimes mn = t[0], mx = t[0], sum;
  for ( std::size_t r = 0; r < t.size(); ++r ) {
    std::printf( "%zu\t%0.3e\t%0.3e\t%0.3e\t%0.2f\t%0.2f\n", r, t[r].read, t[r].eig, t[r].total, t[r].read / t[r].total,
                 t[r].eig / t[r].total );
    mn.read = std::min( mn.read, t[r].read ); mn.eig = std::min( mn.eig, t[r].eig ); mn.total = std::min( mn.total, t[r].total );
    mx.read = std::max( mx.read, t[r].read ); mx.eig = std::max( mx.eig, t[r].eig ); mx.total = std::max( mx.total, t[r].total );
    sum.read += t[r].read; sum.eig += t[r].eig; sum.total += t[r].total;
  }
  const double p = static_cast<double>( t.size() );
  std::cout << "---------------------------------------------------------------------\n";
  std::printf( "min\t%0.3e\t%0.3e\t%0.3e\n", mn.read, mn.eig, mn.total );
  std::printf( "avg\t%0.3e\t%0.3e\t%0.3e\t%0.2f\t%0.2f\n", sum.read / p, sum.eig / p, sum.total / p, sum.read / sum.total,
               sum.eig / sum.total );
  std::printf( "max\t%0.3e\t%0.3e\t%0.3e\n", mx.read, mx.eig, mx.total );
  std::cout << '\n' << "load imbalance = " << 100 * ( mx.total / ( sum.total / p ) - 1.0 ) << "%" << '\n';
}

inline int host_main( int argc, char* argv[] ) {
  const Options o = parse_options( argc, argv, false );
  print_header();
  int pods = o.gpus;
  if ( pods <= 0 ) pods = o.dump_partition.empty() ? paladin_gpu_device_count() : 1;
  if ( pods <= 0 ) {
    std::cerr << "paladin-gpu: no usable sm_100 device (there is no CPU fallback)\n";
    return 3;
  }
  const int total = count_nonempty_lines( o.listing );
  std::cout << "\nCommand line parsed!" << '\n';
  std::cout << "GPUs (pods)        : " << pods << '\n';
  std::cout << "Matrix listing     : " << o.listing << '\n';
  std::cout << "Number of matrices : " << total << '\n';
  std::cout << "Load measure type  : " << measure_type_description( o.measure ) << '\n';
  std::cout << "Timing repeats     : " << o.repeats << '\n' << '\n';
  const Plan plan = make_plan( o, pods );
  if ( !o.dump_partition.empty() ) {
    std::ofstream out( o.dump_partition );
    for ( std::size_t s = 0; s < plan.paths.size(); ++s ) out << s << " " << plan.colour[s] << " " << plan.paths[s] << '\n';
    return 0;
  }
  if ( o.dynamic ) {
    // SURVEY.md 8(f) rank 4, behind a flag: the static colouring balances the reference's cost model (nnz, n^3, ...), which
    // is not what a matrix costs on a GPU. Here the sorted list (largest measure first) is cut into batches and every GPU
    // thread takes the next batch when it is done with its own. The output files do not depend on who computed them.
    const std::vector<std::pair<std::size_t, std::size_t>> chunks = dynamic_chunks( plan.paths.size(), static_cast<std::size_t>( o.batch ) );
    std::cout << "dynamic balancing: " << chunks.size() << " batches of up to " << o.batch << " matrices on " << pods << " GPUs" << '\n';
    std::atomic<std::size_t> next_chunk( 0 );
    std::vector<PodTimes> dtimes( pods );
    std::vector<int> dstatus( pods, 0 );
    std::vector<std::size_t> taken( pods, 0 );
    std::vector<std::thread> dworkers;
    const int dparse = std::max( 1, static_cast<int>( std::thread::hardware_concurrency() ) / pods - 1 );
    for ( int r = 0; r < pods; ++r )
      dworkers.emplace_back( [&, r] {
        const auto t0 = std::chrono::steady_clock::now();
        for ( ;; ) {
          const std::size_t c = next_chunk.fetch_add( 1 );
          if ( c >= chunks.size() ) break;
          const std::vector<std::string> part( plan.paths.begin() + chunks[c].first, plan.paths.begin() + chunks[c].first + chunks[c].second );
          PodTimes t;
          const int st = run_pod( r, part, o, t, dparse );
          if ( st < 0 ) {
            dstatus[r] = st;
            break;
          }
          dstatus[r] += st;
          dtimes[r].read += t.read;
          dtimes[r].eig += t.eig;
          taken[r] += chunks[c].second;
        }
        dtimes[r].total = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count() / o.repeats;
      } );
    for ( auto& w : dworkers ) w.join();
    int dfailed = 0;
    for ( int r = 0; r < pods; ++r ) {
      std::cout << "pod " << r << ": " << taken[r] << " matrices" << '\n';
      if ( dstatus[r] < 0 ) return 4;
      dfailed += dstatus[r];
    }
    display_timings( dtimes, o.repeats );
    if ( dfailed > 0 ) {
      std::cerr << "paladin-gpu: " << dfailed << " matrices were rejected or did not converge (see the warnings above)\n";
      return 5;
    }
    return 0;
  }
  std::vector<std::vector<std::string>> pod( pods );
  for ( std::size_t s = 0; s < plan.paths.size(); ++s ) pod[plan.colour[s]].push_back( plan.paths[s] );
  for ( int r = 0; r < pods; ++r ) {
    std::cout << "pod " << r << ": " << pod[r].size() << " matrices" << '\n';
    if ( o.showdist )
      for ( std::size_t s = 0; s < plan.paths.size(); ++s )
        if ( plan.colour[s] == r ) std::printf( "%s%0.1e%s%s\n", "    ", plan.measures[s], "   ", plan.paths[s].c_str() );
  }
  std::vector<PodTimes> times( pods );
  std::vector<int> status( pods, 0 );
  std::vector<std::thread> workers;
  const int parse_threads = std::max( 1, static_cast<int>( std::thread::hardware_concurrency() ) / pods - 1 );
  for ( int r = 0; r < pods; ++r ) workers.emplace_back( [&, r] { status[r] = run_pod( r, pod[r], o, times[r], parse_threads ); } );
  for ( auto& w : workers ) w.join();
  int failed = 0;
  for ( int r = 0; r < pods; ++r ) {
    if ( status[r] < 0 ) return 4;
    failed += status[r];
  }
  display_timings( times, o.repeats );
  if ( failed > 0 ) {
    std::cerr << "paladin-gpu: " << failed << " matrices were rejected or did not converge (see the warnings above)\n";
    return 5;
  }
  return 0;
}

}  // namespace paladin

#ifndef PALADIN_HOST_LIBRARY
int main( int argc, char* argv[] ) { return paladin::host_main( argc, argv ); }
#else
// ---- test hooks (CPU only) ----------------------------------------------------------------------------------------
extern "C" {

// returns nnz (>= 0) or -1; call with null arrays first to size them
int ph_parse_mm( const char* path, int* n, int cap, int* row, int* col, double* val ) {
  try {
    const paladin::CooMatrix m = paladin::read_coo_from_mm_file( path );
    *n = m.nrows_;
    if ( row && cap >= m.nnz_ ) {
      std::copy( m.row.begin(), m.row.end(), row );
      std::copy( m.col.begin(), m.col.end(), col );
      std::copy( m.val.begin(), m.val.end(), val );
    }
    return m.nnz_;
  } catch ( const std::exception& ) {
    return -1;
  }
}

// parse_batch over `count` newline-separated paths on `threads` workers; returns the total entry count (>= 0) or -1.
// Call with null arrays first to size them; n[count], off[count + 1].
long long ph_parse_batch( const char* paths, int count, int threads, int* n, long long* off, long long cap, int* row, int* col, double* val ) {
  try {
    std::vector<std::string> pod;
    std::istringstream in( paths );
    for ( std::string line; std::getline( in, line ); ) pod.push_back( line );
    if ( static_cast<int>( pod.size() ) != count ) return -1;
    const paladin::ParsedBatch b = paladin::parse_batch( pod, 0, pod.size(), threads );
    const long long total = static_cast<long long>( b.off[pod.size()] );
    if ( row && cap >= total ) {
      std::copy( b.n.begin(), b.n.end(), n );
      for ( std::size_t k = 0; k <= pod.size(); ++k ) off[k] = static_cast<long long>( b.off[k] );
      std::copy( b.ci.begin(), b.ci.end(), row );
      std::copy( b.cj.begin(), b.cj.end(), col );
      std::copy( b.cv.begin(), b.cv.end(), val );
    }
    return total;
  } catch ( const std::exception& ) {
    return -1;
  }
}

double ph_measure( int n, int nnz, const char* kind ) { return paladin::matrix_measure( n, nnz, paladin::string_to_measure_type( kind ) ); }

// order[s] = listing index at sorted position s, colour[s] = its pod
void ph_partition( int count, const double* measures, int pods, int* order, int* colour ) {
  std::vector<paladin::NameMeasurePairT> entries;
  for ( int k = 0; k < count; ++k ) entries.emplace_back( std::to_string( k ), measures[k] );
  paladin::sort_name_measure_pairs( entries );
  const std::vector<int> c = paladin::color_pods( entries, pods );
  for ( int s = 0; s < count; ++s ) {
    order[s] = std::stoi( entries[s].first );
    colour[s] = c[s];
  }
}

// formats into caller buffers; returns the byte count needed
// which: 0 spectrum, 1 right-vector file (vr = right vectors), 2 left-vector file (vr = left vectors)
long ph_format( int n, const double* wr, const double* wi, const double* vr, double threshold, int which, char* out, long cap ) {
  paladin::SpectralDecomposition s( n, threshold );
  s.unpack( n, wr, wi, which == 2 ? nullptr : vr, which == 2 ? vr : nullptr );
  std::string str;
  if ( which == 0 ) s.write_eigenvalues( str );
  else if ( which == 1 ) s.write_right_eigenvectors( str );
  else s.write_left_eigenvectors( str );
  if ( out && cap >= static_cast<long>( str.size() ) ) std::memcpy( out, str.data(), str.size() );
  return static_cast<long>( str.size() );
}

int ph_main( int argc, char** argv ) { return paladin::host_main( argc, argv ); }
}
#endif
