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
#include <cstdlib>
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
  bool host_text = false;  // --host-text: parse and format on the host cores (default: the text ends run on the GPU)
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
  o.host_text = clp.checkExists( "host-text" );
  if ( !quiet ) {
    bool bad = false;
    for ( const auto& k : clp.get_keys() ) {
      static const char* known[] = { "listing", "repeats", "showdist", "measure", "rootdir", "left", "right",
                                     "write-threshold", "gpus", "batch", "dump-partition", "dynamic", "inflight", "host-text" };
      if ( std::none_of( std::begin( known ), std::end( known ), [&]( const char* s ) { return k == s; } ) ) {
        std::cout << "\n-- POTENTIAL ERROR: key " << k << " is not a recognized option!";
        bad = true;
      }
    }
    if ( bad )
      std::cout << "\n-- Allowable keys: listing, repeats, showdist, measure, rootdir, left, right, write-threshold, "
                   "gpus, batch, dump-partition, dynamic, inflight, host-text\n\n";
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

// Page-locked host buffers for the text that goes to / comes from the GPU (uploads and downloads run as DMA at PCIe speed
// instead of the staged copies of pageable memory). Buffers are kept and reused: pinning gigabytes costs more than a batch.
class PinnedPool {
  struct Item { char* p; std::size_t cap; };
  std::mutex mu_;
  std::vector<Item> free_;
 public:
  static PinnedPool& instance() {
    static PinnedPool pool;
    return pool;
  }
  // returns nullptr when the buffer cannot be pinned (the caller falls back on pageable memory)
  char* acquire( std::size_t bytes, std::size_t* cap ) {
    {
      std::lock_guard<std::mutex> lock( mu_ );
      std::size_t best = free_.size();
      for ( std::size_t q = 0; q < free_.size(); ++q )
        if ( free_[q].cap >= bytes && ( best == free_.size() || free_[q].cap < free_[best].cap ) ) best = q;
      if ( best != free_.size() ) {
        const Item it = free_[best];
        free_.erase( free_.begin() + static_cast<std::ptrdiff_t>( best ) );
        *cap = it.cap;
        return it.p;
      }
    }
    void* p = nullptr;
    const std::size_t want = bytes + bytes / 8 + 4096;  // a little head room: the next batch is rarely exactly as large
    if ( paladin_gpu_host_alloc( &p, want ) != PALADIN_GPU_OK ) return nullptr;
    *cap = want;
    return static_cast<char*>( p );
  }
  void release( char* p, std::size_t cap ) {
    std::lock_guard<std::mutex> lock( mu_ );
    free_.push_back( Item{ p, cap } );
  }
};

// a text buffer: pinned when it fits the pool's budget, pageable otherwise
struct HostBuf {
  char* p = nullptr;
  std::size_t cap = 0;
  bool pinned = false;
  HostBuf() = default;
  explicit HostBuf( std::size_t bytes, bool want_pinned = true ) {
    if ( want_pinned && bytes <= ( std::size_t( 1 ) << 30 ) ) p = PinnedPool::instance().acquire( bytes, &cap );
    pinned = p != nullptr;
    if ( !p ) {
      p = new char[bytes];
      cap = bytes;
    }
  }
  HostBuf( HostBuf&& o ) noexcept : p( o.p ), cap( o.cap ), pinned( o.pinned ) { o.p = nullptr; }
  HostBuf& operator=( HostBuf&& o ) noexcept {
    reset();
    p = o.p; cap = o.cap; pinned = o.pinned;
    o.p = nullptr;
    return *this;
  }
  HostBuf( const HostBuf& ) = delete;
  HostBuf& operator=( const HostBuf& ) = delete;
  void reset() {
    if ( p ) {
      if ( pinned ) PinnedPool::instance().release( p, cap );
      else delete[] p;
    }
    p = nullptr;
  }
  ~HostBuf() { reset(); }
  char* get() const { return p; }
};

// A parsed batch: the COO triplets of `count` consecutive pod entries, concatenated in listing order.
struct ParsedBatch {
  std::size_t first = 0, count = 0;
  std::vector<int> n;
  std::vector<int64_t> off;
  std::vector<int> ci, cj;
  std::vector<double> cv;
  // text mode (the GPU parses): the entry lines of every file, packed; toff[k] .. toff[k+1] is matrix k's share
  bool is_text = false;
  HostBuf text;                   // (only when the whole text is held on the host; the default path streams file by file)
  std::vector<int64_t> toff;
  std::vector<long> body_start;   // text mode: where the entry lines start in every file
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

// Text mode: the files of a batch are only READ here -- header and size line parsed, the entry lines copied as they are into
// one packed buffer -- the conversion of the entries happens on the GPU (paladin_gpu_batch_upload_text).
inline ParsedBatch read_text_batch( const std::vector<std::string>& pod, std::size_t first, std::size_t count, int threads ) {
  const auto t0 = std::chrono::steady_clock::now();
  ParsedBatch b;
  b.first = first;
  b.count = count;
  b.is_text = true;
  b.n.assign( count, 0 );
  b.off.assign( count + 1, 0 );
  b.toff.assign( count + 1, 0 );
  std::vector<long> body_start( count, 0 ), body_len( count, 0 );
  std::vector<int> nnz( count, 0 );
  std::vector<std::string> errors( count );
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
  // pass 1: header + size line (the first bytes of the file), file size
  run_on_pool( [&]( std::size_t k ) {
    const std::string& path = pod[first + k];
    std::FILE* f = std::fopen( path.c_str(), "rb" );
    if ( !f ) {
      errors[k] = "Couldn't open the matrix file, " + path;
      return;
    }
    std::string head( 4096, '\0' );
    std::size_t got = std::fread( head.data(), 1, head.size(), f );
    std::fseek( f, 0, SEEK_END );
    const long size = std::ftell( f );
    // (comment blocks longer than the first read: take the whole file)
    while ( got < static_cast<std::size_t>( size ) ) {
      const char* p = head.data();
      const char* e = p + got;
      bool complete = false;
      const char* q = detail::skip_line( p, e );
      while ( q < e && *q == '%' ) q = detail::skip_line( q, e );
      if ( q < e && detail::line_end( q, e ) < e ) complete = true;
      if ( complete ) break;
      head.resize( std::min<std::size_t>( head.size() * 4, static_cast<std::size_t>( size ) ) );
      std::fseek( f, 0, SEEK_SET );
      got = std::fread( head.data(), 1, head.size(), f );
    }
    std::fclose( f );
    try {
      const char* body = nullptr;
      detail::parse_size_line( head.data(), head.data() + got, true, b.n[k], nnz[k], &body, path );
      body_start[k] = static_cast<long>( body - head.data() );
      body_len[k] = std::max( 0L, size - body_start[k] );
    } catch ( const std::exception& e ) {
      errors[k] = e.what();
    }
  } );
  for ( const auto& e : errors )
    if ( !e.empty() ) {
      std::cerr << e << ", exiting!" << '\n';
      std::exit( 1 );
    }
  for ( std::size_t k = 0; k < count; ++k ) {
    b.off[k + 1] = b.off[k] + nnz[k];
    b.toff[k + 1] = b.toff[k] + body_len[k];
  }
  // the entry lines themselves are read later, file by file, straight into page-locked staging buffers on their way to the
  // device (stream_text_to_device): no batch-sized host copy of the text exists
  b.body_start = body_start;
  b.seconds = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
  return b;
}

// Text mode, second half: the entry lines of every file go disk -> page-locked slot (16 MiB per reader thread) -> device.
inline int stream_text_to_device( paladin_gpu_batch* gb, const std::vector<std::string>& pod, const ParsedBatch& b, int threads ) {
  const std::size_t count = b.count;
  int rc = paladin_gpu_batch_text_begin( gb, static_cast<std::size_t>( b.toff[count] ) );
  if ( rc != PALADIN_GPU_OK ) return rc;
  std::atomic<std::size_t> cursor{ 0 };
  std::atomic<int> status{ PALADIN_GPU_OK };
  auto loop = [&]() {
    constexpr std::size_t kSlot = std::size_t( 16 ) << 20;
    HostBuf slot( kSlot );
    for ( std::size_t k = cursor.fetch_add( 1 ); k < count; k = cursor.fetch_add( 1 ) ) {
      const std::size_t len = static_cast<std::size_t>( b.toff[k + 1] - b.toff[k] );
      if ( len == 0 ) continue;
      std::FILE* f = std::fopen( pod[b.first + k].c_str(), "rb" );
      std::size_t pos = 0;
      if ( f ) std::fseek( f, b.body_start[k], SEEK_SET );
      while ( pos < len ) {
        const std::size_t chunk = std::min( kSlot, len - pos );
        const std::size_t got = f ? std::fread( slot.get(), 1, chunk, f ) : 0;
        if ( got < chunk ) std::memset( slot.get() + got, '\n', chunk - got );  // (a file that shrank since its size was taken)
        const int r = paladin_gpu_batch_text_put( gb, static_cast<std::size_t>( b.toff[k] ) + pos, slot.get(), chunk );
        if ( r != PALADIN_GPU_OK ) {
          status = r;
          break;
        }
        pos += chunk;
      }
      if ( f ) std::fclose( f );
    }
  };
  const std::size_t nthreads = std::max<std::size_t>( 1, std::min<std::size_t>( static_cast<std::size_t>( threads ), count ) );
  std::vector<std::thread> pool;
  for ( std::size_t w = 1; w < nthreads; ++w ) pool.emplace_back( loop );
  loop();
  for ( auto& th : pool ) th.join();
  return status.load();
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
    std::vector<double> wr( tn, 0.0 ), wi( tn, 0.0 ), vr( o.right && !cur.is_text ? tnn : 0, 0.0 ), vl( o.left && !cur.is_text ? tnn : 0, 0.0 );
    // write filter of the eigenvector files, evaluated on the device (paladin_gpu_batch_vector_filter)
    std::vector<double> max2r( o.right ? tn : 0, 0.0 ), max2l( o.left ? tn : 0, 0.0 );  // (host-side writers only)
    std::vector<long long> keptr( o.right ? count : 0, 0 ), keptl( o.left ? count : 0, 0 );
    std::vector<int> info( count, 0 );
    const auto t_eig = clock::now();
    static const bool timing = std::getenv( "PALADIN_GPU_TIMING" ) != nullptr;  // per-batch stage times on stderr
    auto lap_t = clock::now();
    std::string laps;
    auto lap = [&]( const char* what ) {
      if ( !timing ) return;
      const auto now = clock::now();
      char tmp[64];
      std::snprintf( tmp, sizeof tmp, " %s %.0fms", what, 1e3 * std::chrono::duration<double>( now - lap_t ).count() );
      laps += tmp;
      lap_t = now;
    };
    paladin_gpu_batch* gb = nullptr;
    int rc = paladin_gpu_batch_create( device, static_cast<int>( count ), cur.n.data(), cur.off.data(), jobvl, jobvr, &gb );
    // text mode: the device converts the entry lines; the few matrices it hands back go through the host reader
    lap( "create" );
    if ( rc == PALADIN_GPU_OK && cur.is_text ) {
      std::vector<int> need( count, 0 );
      rc = stream_text_to_device( gb, pod, cur, parse_threads );
      lap( "read+h2d" );
      if ( rc == PALADIN_GPU_OK ) rc = paladin_gpu_batch_upload_text( gb, nullptr, cur.toff.data(), need.data() );
      for ( std::size_t k = 0; k < count && rc == PALADIN_GPU_OK; ++k ) {
        if ( !need[k] ) continue;
        try {
          const CooMatrix m = read_coo_from_mm_file( pod[first + k] );
          rc = paladin_gpu_batch_upload_coo_one( gb, static_cast<int>( k ), m.row.data(), m.col.data(), m.val.data() );
        } catch ( const std::exception& e ) {
          std::cerr << e.what() << ", exiting!" << '\n';
          std::exit( 1 );
        }
      }
    } else if ( rc == PALADIN_GPU_OK ) {
      rc = paladin_gpu_batch_upload( gb, cur.ci.data() + cur.off[0], cur.cj.data() + cur.off[0], cur.cv.data() + cur.off[0] );
    }
    lap( "upload" );
    if ( rc == PALADIN_GPU_OK ) rc = paladin_gpu_batch_scatter( gb );
    lap( "scatter" );
    if ( rc == PALADIN_GPU_OK ) rc = paladin_gpu_batch_solve( gb );
    lap( "solve" );
    if ( !cur.is_text ) {
      if ( rc == PALADIN_GPU_OK && o.right && count > 0 ) rc = paladin_gpu_batch_vector_filter( gb, 'R', o.threshold, max2r.data(), keptr.data() );
      if ( rc == PALADIN_GPU_OK && o.left && count > 0 ) rc = paladin_gpu_batch_vector_filter( gb, 'L', o.threshold, max2l.data(), keptl.data() );
      lap( "filter" );
    }
    if ( rc == PALADIN_GPU_OK )
      rc = paladin_gpu_batch_download( gb, wr.data(), wi.data(), o.left && !cur.is_text ? vl.data() : nullptr, o.right && !cur.is_text ? vr.data() : nullptr,
                                       info.data() );
    lap( "download" );
    std::atomic<int> bad{ 0 };
    // Output files of matrices [k0, k1) on the writer threads. sides: bit 0 = right-vector file, bit 1 = left-vector file,
    // bit 2 = spectrum file. dev_boff != nullptr: the bodies of the ONE vector side in `sides` were formatted on the device
    // (offsets relative to k0, fetched through paladin_gpu_batch_fetch_text) except where dev_host[k - k0] is set; otherwise
    // the vector files come from the unpacked vectors on the host.
    auto write_files = [&]( std::size_t k0, std::size_t k1, int sides, const std::vector<int64_t>* dev_boff, const std::vector<int>* dev_host ) {
      std::atomic<std::size_t> cursor{ k0 };
      auto write_some = [&]() {
        for ( std::size_t k = cursor.fetch_add( 1 ); k < k1; k = cursor.fetch_add( 1 ) ) {
          const int nk = cur.n[k];
          const bool ok = info[k] == 0;
          if ( !ok && ( sides & 4 ) ) {
            ++bad;
            std::lock_guard<std::mutex> lock( mu );
            std::cerr << "warning: " << pod[first + k] << ": info = " << info[k]
                      << ( info[k] < 0 ? " (input rejected, no output files written)" : " (QR iteration did not converge, eigenvalues before that index are NaN)" ) << '\n';
          }
          // info < 0: the library refused the input (COO index outside [1, n], non-finite entry) and computed nothing --
          // there is no spectrum to write; the run ends with a non-zero exit status
          if ( info[k] < 0 ) continue;
          const bool on_device = dev_boff != nullptr && ok && !( *dev_host )[k - k0];
          const bool want_r = ( sides & 1 ) != 0, want_l = ( sides & 2 ) != 0;
          const bool have_vr = !vr.empty(), have_vl = !vl.empty();
          SpectralDecomposition s( nk, o.threshold );
          s.unpack( nk, wr.data() + woff[k], wi.data() + woff[k], want_r && ok && !on_device && have_vr ? vr.data() + voff[k] : nullptr,
                    want_l && ok && !on_device && have_vl ? vl.data() + voff[k] : nullptr );
          std::string text;
          if ( sides & 4 ) {
            s.write_eigenvalues( text );
            write_file( pod[first + k] + ".spectrum", text );
          }
          for ( int side = 0; side < 2; ++side ) {
            const bool is_r = side == 0;
            if ( !( is_r ? want_r : want_l ) ) continue;
            const std::string path = pod[first + k] + ( is_r ? ".rightvecs" : ".leftvecs" );
            const long long kept_k = ( is_r ? keptr : keptl )[k];
            if ( on_device ) {
              std::string head = "%%MatrixMarket matrix coordinate complex general\n";
              head += std::to_string( nk ) + " " + std::to_string( nk ) + " " + std::to_string( kept_k ) + "\n";
              std::FILE* f = std::fopen( path.c_str(), "wb" );
              if ( !f ) {
                std::cerr << "Couldn't open the output file, " << path << '\n';
                continue;
              }
              std::fwrite( head.data(), 1, head.size(), f );
              // the body streams from the device through this thread's page-locked staging buffer
              constexpr std::size_t kSlot = std::size_t( 16 ) << 20;
              thread_local HostBuf slot;
              if ( slot.get() == nullptr ) slot = HostBuf( kSlot );
              std::size_t pos = static_cast<std::size_t>( ( *dev_boff )[k - k0] );
              const std::size_t end = static_cast<std::size_t>( ( *dev_boff )[k - k0 + 1] );
              while ( pos < end ) {
                const std::size_t chunk = std::min( kSlot, end - pos );
                if ( paladin_gpu_batch_fetch_text( gb, pos, chunk, slot.get() ) != PALADIN_GPU_OK ) {
                  std::lock_guard<std::mutex> lock( mu );
                  std::cerr << "fetching the formatted vectors failed: " << paladin_gpu_last_error() << '\n';
                  ++bad;
                  break;
                }
                std::fwrite( slot.get(), 1, chunk, f );
                pos += chunk;
              }
              std::fclose( f );
            } else {
              text.clear();
              const bool have = is_r ? have_vr : have_vl;
              const double* m2 = ( is_r ? max2r : max2l ).data() + woff[k];
              if ( is_r ) {
                if ( ok && have ) s.write_right_eigenvectors( text, m2, kept_k );
                else s.write_right_eigenvectors( text );
              } else {
                if ( ok && have ) s.write_left_eigenvectors( text, m2, kept_k );
                else s.write_left_eigenvectors( text );
              }
              write_file( path, text );
            }
          }
        }
      };
      std::vector<std::thread> writers;
      const std::size_t nw = std::min<std::size_t>( static_cast<std::size_t>( std::max( 1, parse_threads ) ), k1 - k0 );
      for ( std::size_t w = 1; w < nw; ++w ) writers.emplace_back( write_some );
      write_some();
      for ( auto& w : writers ) w.join();
    };
    std::string why = rc == PALADIN_GPU_OK ? std::string() : std::string( paladin_gpu_last_error() );  // per-thread message
    if ( rc == PALADIN_GPU_OK && write && cur.is_text && ( o.right || o.left ) && count > 0 ) {
      // the bodies of the eigenvector files come back as text, a range of matrices at a time (48 bytes per component always
      // suffice; ranges of at most 24 Mi components keep the page-locked buffers near 1 GiB)
      const std::size_t range_cap = std::size_t( 24 ) << 20;
      std::size_t k0 = 0;
      while ( k0 < count && rc == PALADIN_GPU_OK ) {
        std::size_t k1 = k0, entries = 0;
        while ( k1 < count && ( k1 == k0 || entries + ( voff[k1 + 1] - voff[k1] ) <= range_cap ) ) {
          entries += voff[k1 + 1] - voff[k1];
          ++k1;
        }
        const std::size_t cnt = k1 - k0;
        std::vector<int64_t> boffr( cnt + 1, 0 ), boffl( cnt + 1, 0 );
        std::vector<int> hostr( cnt, 0 ), hostl( cnt, 0 );
        // the device keeps the formatted text of one side at a time: format right, write, format left, write
        for ( int side = 0; side < 2 && rc == PALADIN_GPU_OK; ++side ) {
          const bool is_r = side == 0;
          if ( ( is_r && !o.right ) || ( !is_r && !o.left ) ) continue;
          rc = paladin_gpu_batch_format_vectors( gb, is_r ? 'R' : 'L', o.threshold, static_cast<int>( k0 ), static_cast<int>( cnt ), nullptr, 0,
                                                 ( is_r ? boffr : boffl ).data(), ( is_r ? keptr : keptl ).data() + k0, ( is_r ? hostr : hostl ).data() );
          lap( "format" );
          if ( rc != PALADIN_GPU_OK ) break;
          bool refused = false;
          for ( std::size_t q = 0; q < cnt; ++q ) refused = refused || ( is_r ? hostr : hostl )[q];
          if ( refused && vr.empty() && vl.empty() ) {
            // (a rounding the device refused: those matrices are formatted by the host writer from the vectors themselves)
            if ( o.right ) vr.assign( tnn, 0.0 );
            if ( o.left ) vl.assign( tnn, 0.0 );
            rc = paladin_gpu_batch_download( gb, wr.data(), wi.data(), o.left ? vl.data() : nullptr, o.right ? vr.data() : nullptr, info.data() );
            if ( rc == PALADIN_GPU_OK && o.right ) rc = paladin_gpu_batch_vector_filter( gb, 'R', o.threshold, max2r.data(), keptr.data() );
            if ( rc == PALADIN_GPU_OK && o.left ) rc = paladin_gpu_batch_vector_filter( gb, 'L', o.threshold, max2l.data(), keptl.data() );
            if ( rc != PALADIN_GPU_OK ) break;
          }
          write_files( k0, k1, ( is_r ? 1 : 2 ) | ( ( is_r || !o.right ) ? 4 : 0 ), is_r ? &boffr : &boffl, is_r ? &hostr : &hostl );
          lap( "write" );
        }
        k0 = k1;
      }
      if ( rc != PALADIN_GPU_OK ) why = paladin_gpu_last_error();
    }
    paladin_gpu_batch_destroy( gb );
    lap( "destroy" );
    {
      std::lock_guard<std::mutex> lock( mu );
      times.eig += std::chrono::duration<double>( clock::now() - t_eig ).count();
      if ( rc != PALADIN_GPU_OK ) std::cerr << "the GPU eigensolver failed on device " << device << ": " << why << '\n';
    }
    if ( rc != PALADIN_GPU_OK ) return -1;
    if ( write && !( cur.is_text && ( o.right || o.left ) && count > 0 ) ) {
      // host-side writers (or no vector files at all): unpack + text formatting, one matrix per worker at a time
      write_files( 0, count, ( o.right ? 1 : 0 ) | ( o.left ? 2 : 0 ) | 4, nullptr, nullptr );
      lap( "write" );
    }
    if ( timing ) {
      std::lock_guard<std::mutex> lock( mu );
      std::cerr << "batch@" << first << " (" << count << " matrices, read " << static_cast<int>( 1e3 * cur.seconds ) << "ms):" << laps << '\n';
    }
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
      next = std::async( std::launch::async, o.host_text ? parse_batch : read_text_batch, std::cref( pod ), firsts[0], std::min<std::size_t>( o.batch, pod.size() ), parse_threads );
    for ( std::size_t bi = 0; bi < firsts.size(); ++bi ) {
      auto cur = std::make_shared<ParsedBatch>( next.get() );
      times.read += cur->seconds;
      if ( bi + 1 < firsts.size() )
        next = std::async( std::launch::async, o.host_text ? parse_batch : read_text_batch, std::cref( pod ), firsts[bi + 1],
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
  // the CUDA contexts (and the library's per-device state) come up while the listing pre-pass reads the size lines
  std::thread warm;
  if ( o.dump_partition.empty() )
    warm = std::thread( [pods] {
      for ( int r = 0; r < pods; ++r ) paladin_gpu_init( r );
    } );
  struct Joiner {
    std::thread& t;
    ~Joiner() {
      if ( t.joinable() ) t.join();
    }
  } warm_joiner{ warm };
  const int total = count_nonempty_lines( o.listing );
  std::cout << "\nCommand line parsed!" << '\n';
  std::cout << "GPUs (pods)        : " << pods << '\n';
  std::cout << "Matrix listing     : " << o.listing << '\n';
  std::cout << "Number of matrices : " << total << '\n';
  std::cout << "Load measure type  : " << measure_type_description( o.measure ) << '\n';
  std::cout << "Timing repeats     : " << o.repeats << '\n' << '\n';
  const Plan plan = make_plan( o, pods );
  if ( warm.joinable() ) warm.join();
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
