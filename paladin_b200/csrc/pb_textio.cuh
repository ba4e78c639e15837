// pb_textio.cuh -- the text ends of the path on the device:
//
//   parse  : MatrixMarket entry lines "i j v\n" -> COO triplets, the conversions of the reference reader
//            (reference src/square-matrix.hpp:101-117: two fields ended by ' ', one by '\n', then atoi / atoi / atof), bit-exact.
//            The raw text of a batch is uploaded as it is on disk; chunks of 8 KiB are staged in shared memory with 16-byte
//            loads, line starts are ranked with a block scan on top of a per-matrix chunk scan, the thread that owns a line
//            start converts the line (pb_text.cuh). A field the device refuses to convert with certainty goes on a fallback
//            list (the host converts it with the C library); a line with fewer than two blanks, an over-long field or a
//            matrix with fewer lines than entries marks the whole matrix for the host reader.
//   format : eigenvector files, entry lines "row col re im\n" for every component with |x|^2 > threshold * max |x|^2
//            (reference src/spectrum-util.hpp:134-196; numbers as operator<<(double) prints them, pb_text.cuh::format_g6).
//            One thread formats one component into a fixed slot, a three-kernel scan of the line lengths places the lines,
//            a copy kernel compacts them: the bytes of the file body in file order. A number the device refuses to round
//            with certainty marks the matrix for the host writer.
#pragma once

#include "pb_common.cuh"
#include "pb_text.cuh"

#if defined( __CUDACC__ )

namespace pb {

constexpr int kTxtChunk = 8192;   // bytes of text per CTA
constexpr int kTxtThreads = 256;  // 32 bytes per thread
constexpr int kTxtHalo = 96;      // a line that starts in the chunk may end this far behind it (longest legal line: 5+1+5+1+31+1)
constexpr int kTxtMaxFallback = 1 << 20;

struct TextChunk {
  int matrix;
  int index;            // chunk index inside the matrix
  long long begin;      // byte offset of the chunk in the batch text
  int bytes;            // bytes of the chunk (<= kTxtChunk)
};
struct TextFallback {
  int matrix;
  int line;             // entry index inside the matrix
  long long offset;     // byte offset of the line start in the batch text
};
struct TextArgs {
  const char* text;
  const long long* t_begin;  // per matrix: byte range of its entry lines
  const long long* t_end;
  const int64_t* nnz_off;
  const TextChunk* chunks;
  const long long* chunk_first;  // per matrix: index of its first chunk
  int* chunk_lines;          // per chunk: line starts in it, then (after the scan) the rank of its first line start
  int* coo_i;
  int* coo_j;
  double* coo_v;
  int* needs_host;           // per matrix
  int* lines_found;          // per matrix
  TextFallback* fallback;
  int* n_fallback;
};

__device__ __forceinline__ bool txt_is_start( const char* text, long long pos, long long mbegin ) { return pos == mbegin || text[pos - 1] == '\n'; }

// pass 1: line starts per chunk
__global__ void __launch_bounds__( kTxtThreads ) text_count_kernel( TextArgs a ) {
  const TextChunk ch = a.chunks[blockIdx.x];
  const long long mb = a.t_begin[ch.matrix];
  int cnt = 0;
  const int lo = threadIdx.x * 32;
  for ( int q = lo; q < lo + 32 && q < ch.bytes; ++q ) cnt += txt_is_start( a.text, ch.begin + q, mb ) ? 1 : 0;
  cnt = warp_isum( cnt );
  __shared__ int red[kTxtThreads / 32];
  if ( ( threadIdx.x & 31 ) == 0 ) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if ( threadIdx.x == 0 ) {
    int t = 0;
    for ( int w = 0; w < kTxtThreads / 32; ++w ) t += red[w];
    a.chunk_lines[blockIdx.x] = t;
  }
}

// pass 2: exclusive scan of the chunk counts inside every matrix (one warp per matrix)
__global__ void __launch_bounds__( 32 ) text_scan_kernel( TextArgs a, int count ) {
  const int k = blockIdx.x;
  if ( k >= count ) return;
  const long long c0 = a.chunk_first[k], c1 = a.chunk_first[k + 1];
  const int lane = threadIdx.x;
  int carry = 0;
  for ( long long c = c0; c < c1; c += 32 ) {
    const long long q = c + lane;
    const int v = q < c1 ? a.chunk_lines[q] : 0;
    int s = v;
#pragma unroll
    for ( int o = 1; o < 32; o <<= 1 ) {
      const int t = __shfl_up_sync( 0xffffffffu, s, o );
      if ( lane >= o ) s += t;
    }
    if ( q < c1 ) a.chunk_lines[q] = carry + s - v;
    carry += __shfl_sync( 0xffffffffu, s, 31 );
  }
  if ( lane == 0 ) {
    a.lines_found[k] = carry;
    const long long nnz = a.nnz_off[k + 1] - a.nnz_off[k];
    if ( carry < nnz ) a.needs_host[k] = 1;  // fewer lines than entries: whatever the reference reader makes of that, it makes it
  }
}

// pass 3: every line start is converted by the thread that owns its byte
__global__ void __launch_bounds__( kTxtThreads ) text_parse_kernel( TextArgs a ) {
  __shared__ __align__( 16 ) char buf[kTxtChunk + kTxtHalo + 32];
  __shared__ int wsum[kTxtThreads / 32];
  const TextChunk ch = a.chunks[blockIdx.x];
  const long long mb = a.t_begin[ch.matrix], me = a.t_end[ch.matrix];
  const long long nnz = a.nnz_off[ch.matrix + 1] - a.nnz_off[ch.matrix];
  // stage the chunk and its halo: 16-byte loads from the aligned address below the chunk start
  const long long abase = ch.begin & ~15LL;
  const int shift = (int)( ch.begin - abase );
  const long long want_end = ( ch.begin + ch.bytes + kTxtHalo < me ) ? ch.begin + ch.bytes + kTxtHalo : me;
  const int nload = (int)( ( want_end - abase + 15 ) >> 4 );
  for ( int q = threadIdx.x; q < nload; q += kTxtThreads )
    reinterpret_cast<uint4*>( buf )[q] = reinterpret_cast<const uint4*>( a.text + abase )[q];  // (the batch text is padded by 16 bytes)
  __syncthreads();
  const char* cb = buf + shift;                  // cb[q] = text[ch.begin + q]
  const int avail = (int)( want_end - ch.begin );  // valid bytes from the chunk start
  const char prev = ( ch.begin > mb ) ? a.text[ch.begin - 1] : '\n';
  // line starts in this thread's 32 bytes
  const int lo = threadIdx.x * 32;
  unsigned starts = 0;
  for ( int q = 0; q < 32; ++q ) {
    const int p = lo + q;
    if ( p >= ch.bytes ) break;
    const char before = p == 0 ? prev : cb[p - 1];
    if ( before == '\n' ) starts |= 1u << q;
  }
  int cnt = __popc( starts );
  // exclusive scan over the block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int s = cnt;
#pragma unroll
  for ( int o = 1; o < 32; o <<= 1 ) {
    const int t = __shfl_up_sync( 0xffffffffu, s, o );
    if ( lane >= o ) s += t;
  }
  if ( lane == 31 ) wsum[warp] = s;
  __syncthreads();
  int base = a.chunk_lines[blockIdx.x];
  for ( int w = 0; w < warp; ++w ) base += wsum[w];
  int line = base + s - cnt;
  while ( starts ) {
    const int q = __ffs( starts ) - 1;
    starts &= starts - 1;
    const int p0 = lo + q;
    const int L = line++;
    if ( L >= nnz ) continue;  // lines past the entry count are never read by the reference
    // fields: up to the first blank, up to the second blank, up to the newline (or the end of the text)
    int p = p0;
    int f0 = p, l0, f1, l1, f2, l2;
    while ( p < avail && cb[p] != ' ' && cb[p] != '\n' ) ++p;
    l0 = p - f0;
    bool broken = !( p < avail && cb[p] == ' ' );
    f1 = ++p;
    while ( !broken && p < avail && cb[p] != ' ' && cb[p] != '\n' ) ++p;
    l1 = p - f1;
    broken = broken || !( p < avail && cb[p] == ' ' );
    f2 = ++p;
    while ( !broken && p < avail && cb[p] != '\n' ) ++p;
    l2 = p - f2;
    // a line that runs out of the staged halo without a newline is either the last line of the text (fine) or too long
    if ( !broken && p >= avail && ch.begin + avail < me ) broken = true;
    if ( broken || l0 > 5 || l1 > 5 || l2 > 31 ) {
      a.needs_host[ch.matrix] = 1;  // the reference reads fields, not lines: only its own reader knows what this means
      continue;
    }
    int vi = 0, vj = 0;
    double vv = 0.0;
    const bool ok = text_to_index( cb + f0, l0, &vi ) && text_to_index( cb + f1, l1, &vj ) && text_to_double( cb + f2, l2, &vv );
    const long long e = a.nnz_off[ch.matrix] + L;
    if ( ok ) {
      a.coo_i[e] = vi;
      a.coo_j[e] = vj;
      a.coo_v[e] = vv;
    } else {
      const int slot = atomicAdd( a.n_fallback, 1 );
      if ( slot < kTxtMaxFallback ) a.fallback[slot] = TextFallback{ ch.matrix, L, ch.begin + p0 };
      else a.needs_host[ch.matrix] = 1;
    }
  }
}

// entries converted on the host, patched into the device arrays
__global__ void text_patch_kernel( int count, const long long* where, const int* vi, const int* vj, const double* vv, int* coo_i, int* coo_j,
                                   double* coo_v ) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if ( q >= count ) return;
  coo_i[where[q]] = vi[q];
  coo_j[where[q]] = vj[q];
  coo_v[where[q]] = vv[q];
}

// ---- formatting ---------------------------------------------------------------------------------------------------------------
constexpr int kFmtSlot = 48;       // bytes per formatted line (longest: 10 + 1 + 10 + 1 + 13 + 1 + 13 + 1 = 50 only for n >= 10^9)
constexpr int kFmtThreads = 256;
constexpr int kFmtPerBlock = 1024;  // entries per block of the length scan

struct FmtArgs {
  const int* n;            // orders of the matrices of this call
  const int64_t* a_off;    // dense offsets (device layout)
  const int64_t* w_off;
  const long long* e_off;  // per matrix: first entry (vector-major, n^2 per matrix) inside this call
  int count;
  const double* vecs;      // packed vectors (right or left)
  const double* wi;
  const double* max2;      // largest |x|^2 per unpacked vector
  const int* info;
  double threshold;
  char* slots;             // kFmtSlot bytes per entry
  unsigned char* lens;     // line length per entry (0 = not written)
  int* needs_host;         // per matrix
};

// entry e of matrix k: vector v = e / n (output order), row r = e % n
__global__ void __launch_bounds__( kFmtThreads ) fmt_lines_kernel( FmtArgs a, long long total ) {
  const long long e = (long long)blockIdx.x * kFmtThreads + threadIdx.x;
  if ( e >= total ) return;
  // matrix of this entry: binary search in e_off
  int lo = 0, hi = a.count - 1;
  while ( lo < hi ) {
    const int mid = ( lo + hi + 1 ) >> 1;
    if ( a.e_off[mid] <= e ) lo = mid;
    else hi = mid - 1;
  }
  const int k = lo;
  const int n = a.n[k];
  const long long q = e - a.e_off[k];
  const int v = (int)( q / n ), r = (int)( q - (long long)v * n );
  unsigned char len = 0;
  if ( a.info[k] == 0 ) {
    const double* wi = a.wi + a.w_off[k];
    // the reference decides "pair" sequentially (skipNext, src/paladin.hpp:257-289); with LAPACK's output convention
    // (wi[c] > 0 opens a pair, wi[c + 1] == -wi[c] closes it) column v is the second of a pair iff wi[v] < 0 and
    // wi[v - 1] == -wi[v], and the first iff wi[v] > 0 and wi[v + 1] == -wi[v] -- checked against the host unpack in tests
    int role = 0;
    if ( wi[v] != 0.0 ) {
      // replay from the start of the run of non-zero imaginary parts that contains v
      int s0 = v;
      while ( s0 > 0 && wi[s0 - 1] != 0.0 ) --s0;
      bool second = false;
      for ( int c = s0; c <= v; ++c ) {
        if ( second ) {
          role = 2;
          second = false;
        } else {
          const bool pair = ( c + 1 < n ) && ( wi[c] == -wi[c + 1] ) && ( wi[c] != 0.0 );
          role = pair ? 1 : 0;
          second = pair;
        }
      }
    }
    const double* V = a.vecs + a.a_off[k];
    const double re = V[r + (size_t)( role == 2 ? v - 1 : v ) * n];
    double im = 0.0;
    if ( role == 1 ) im = 1.0 * V[r + (size_t)( v + 1 ) * n];
    else if ( role == 2 ) im = -1.0 * V[r + (size_t)v * n];
    const double nrm = __dadd_rn( __dmul_rn( re, re ), __dmul_rn( im, im ) );
    const double cut = __dmul_rn( a.threshold, a.max2[a.w_off[k] + v] );
    if ( nrm > cut ) {
      char* out = a.slots + e * kFmtSlot;
      int p = format_uint( (uint32_t)( r + 1 ), out );
      out[p++] = ' ';
      p += format_uint( (uint32_t)( v + 1 ), out + p );
      out[p++] = ' ';
      const int l1 = format_g6( re, out + p );
      p += l1;
      out[p++] = ' ';
      const int l2 = format_g6( im, out + p );
      p += l2;
      out[p++] = '\n';
      if ( l1 == 0 || l2 == 0 ) {
        a.needs_host[k] = 1;
        p = 1;  // keeps the scan well defined; the body of this matrix is discarded
      }
      len = (unsigned char)p;
    }
  }
  a.lens[e] = len;
}

// block sums of the line lengths
__global__ void __launch_bounds__( kFmtThreads ) fmt_blocksum_kernel( const unsigned char* lens, long long total, unsigned long long* bsum ) {
  const long long base = (long long)blockIdx.x * kFmtPerBlock;
  unsigned s = 0;
  for ( int q = threadIdx.x; q < kFmtPerBlock; q += kFmtThreads )
    if ( base + q < total ) s += lens[base + q];
  s = (unsigned)warp_isum( (int)s );
  __shared__ unsigned red[kFmtThreads / 32];
  if ( ( threadIdx.x & 31 ) == 0 ) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if ( threadIdx.x == 0 ) {
    unsigned t = 0;
    for ( int w = 0; w < kFmtThreads / 32; ++w ) t += red[w];
    bsum[blockIdx.x] = t;
  }
}

// exclusive scan of the block sums (one CTA, sequential over tiles of 1024), total to bsum[nblocks]
__global__ void __launch_bounds__( 1024 ) fmt_scan_kernel( unsigned long long* bsum, long long nblocks ) {
  __shared__ unsigned long long wtot[32];
  __shared__ unsigned long long carry_s;
  if ( threadIdx.x == 0 ) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for ( long long b0 = 0; b0 < nblocks; b0 += 1024 ) {
    const long long q = b0 + threadIdx.x;
    const unsigned long long v = q < nblocks ? bsum[q] : 0ull;
    unsigned long long s = v;
#pragma unroll
    for ( int o = 1; o < 32; o <<= 1 ) {
      const unsigned long long t = __shfl_up_sync( 0xffffffffu, s, o );
      if ( lane >= o ) s += t;
    }
    if ( lane == 31 ) wtot[warp] = s;
    __syncthreads();
    unsigned long long wbase = 0;
    for ( int w = 0; w < warp; ++w ) wbase += wtot[w];
    const unsigned long long carry = carry_s;
    if ( q < nblocks ) bsum[q] = carry + wbase + s - v;
    __syncthreads();
    if ( threadIdx.x == 1023 ) carry_s = carry + wbase + s;
    __syncthreads();
  }
  if ( threadIdx.x == 0 ) bsum[nblocks] = carry_s;
}

// byte offset of every matrix body inside the compacted text: the scan value at its first entry
__global__ void fmt_bodyoff_kernel( FmtArgs a, long long total, long long nblocks, const unsigned long long* bsum, long long* body_off ) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if ( k > a.count ) return;
  if ( k == a.count ) {
    body_off[k] = (long long)bsum[nblocks];
    return;
  }
  const long long e = a.e_off[k];
  if ( e >= total ) {
    body_off[k] = (long long)bsum[nblocks];
    return;
  }
  const long long blk = e / kFmtPerBlock;
  unsigned long long off = bsum[blk];
  for ( long long q = blk * kFmtPerBlock; q < e; ++q ) off += a.lens[q];
  body_off[k] = (long long)off;
}

// compaction: every line to its place
__global__ void __launch_bounds__( kFmtThreads ) fmt_compact_kernel( FmtArgs a, long long total, const unsigned long long* bsum, char* out ) {
  __shared__ unsigned wsum[kFmtThreads / 32];
  __shared__ unsigned carry_s;
  const long long base = (long long)blockIdx.x * kFmtPerBlock;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ( threadIdx.x == 0 ) carry_s = 0;
  __syncthreads();
  const unsigned long long boff = bsum[blockIdx.x];
  for ( int q0 = 0; q0 < kFmtPerBlock; q0 += kFmtThreads ) {
    const long long e = base + q0 + threadIdx.x;
    const unsigned len = e < total ? a.lens[e] : 0u;
    unsigned s = len;
#pragma unroll
    for ( int o = 1; o < 32; o <<= 1 ) {
      const unsigned t = __shfl_up_sync( 0xffffffffu, s, o );
      if ( lane >= o ) s += t;
    }
    if ( lane == 31 ) wsum[warp] = s;
    __syncthreads();
    unsigned wb = 0;
    for ( int w = 0; w < warp; ++w ) wb += wsum[w];
    const unsigned carry = carry_s;
    const unsigned long long off = boff + carry + wb + s - len;
    if ( e < total ) {
      const char* src = a.slots + e * kFmtSlot;
      for ( unsigned c = 0; c < len; ++c ) out[off + c] = src[c];
    }
    __syncthreads();
    if ( threadIdx.x == kFmtThreads - 1 ) carry_s = carry + wb + s;
    __syncthreads();
  }
}

}  // namespace pb

#endif  // __CUDACC__
