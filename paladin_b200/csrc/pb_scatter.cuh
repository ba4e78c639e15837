// pb_scatter.cuh -- COO -> dense on the device.
//
// Replaces the store at reference src/square-matrix.hpp:115 (mat_[(i-1)+n*(j-1)] = v, applied in
// file order so that a later duplicate overwrites an earlier one) together with the zero fill of
// the SquareMatrix constructor (src/square-matrix.hpp:70-71). Output must be bit-identical.
//
// Algorithmic bytes per matrix: 16*nnz (read i, j, v) + 8*n^2 (write every dense element once).
//
// Fast path (scatter_sorted_kernel): MatrixMarket files are almost always written in increasing
// column-major position. For such input entry k owns the half-open range (p[k-1], p[k]]: it writes
// zeros into the gap and its value at p[k], so every dense element is written exactly once and no
// separate memset pass exists -- DRAM traffic equals the algorithmic bytes. Dense, sorted input
// degenerates into a straight vectorised copy of the value array (16-byte loads and stores).
// Any entry that is not strictly after its predecessor raises the matrix's flag; flagged matrices
// are redone from scratch by the order-exact fallback below (their partial output is discarded).
//
// Fallback (zero_flagged_kernel, scatter_mark/elect/write): last-writer-wins without sorting.
//   mark : atomicMax( slot[p], k+1 ) on the zero-filled dense slots viewed as u64 -> the slot holds
//          the index of the last entry addressing it
//   elect: entry k lost iff slot[p] != k+1; losers are tagged by negating their (1-based) row index
//   write: winners store their value bits into the slot; losers restore their row index
// Each pass is its own launch, so no value is ever compared with an index.
#pragma once

#include <cstdint>

namespace pb {

constexpr int kScatterThreads = 256;
constexpr int kScatterPerThread = 8;                                   // entries per thread (4 pairs)
constexpr int kScatterChunk = kScatterThreads * kScatterPerThread;      // entries per CTA

// flags[] bits
constexpr int kFlagUnsorted = 1;
constexpr int kFlagBadIndex = 2;

struct ScatterChunk {
  int matrix;     // matrix id
  int first;      // first entry of the chunk, relative to the matrix's first entry
  int count;      // entries in the chunk (0 only for an empty matrix)
  int nchunks;    // chunks of this matrix
};

struct ScatterArgs {
  const ScatterChunk* chunks;
  const int* n;                 // order per matrix
  const int64_t* nnz_off;       // count+1
  const int64_t* a_off;         // dense offset per matrix (doubles)
  int* coo_i;                   // device copy (library-owned; losers are tagged in place, then restored)
  const int* coo_j;
  const double* coo_v;
  double* dense;
  int* flags;                   // per matrix
};

#if defined( __CUDACC__ )

__device__ __forceinline__ void fill_zero_warp( double* dst, int64_t start, int64_t len, int lane ) {
  // all lanes of the warp participate; coalesced 8-byte stores
  for ( int64_t q = lane; q < len; q += 32 ) dst[start + q] = 0.0;
}

// gap (lo, hi) exclusive on both ends is zero-filled; small gaps by the owning lane, large ones by
// the whole warp (one lane's gap at a time)
__device__ __forceinline__ void fill_gap( double* dst, int64_t lo, int64_t hi, bool active, int lane ) {
  const int64_t len = active ? ( hi - lo - 1 ) : 0;
  const bool big = len > 16;
  if ( active && !big )
    for ( int64_t q = 0; q < len; ++q ) dst[lo + 1 + q] = 0.0;
  unsigned m = __ballot_sync( 0xffffffffu, big );
  while ( m ) {
    const int src = __ffs( m ) - 1;
    m &= m - 1;
    const int64_t s = __shfl_sync( 0xffffffffu, lo, src );
    const int64_t l = __shfl_sync( 0xffffffffu, len, src );
    fill_zero_warp( dst, s + 1, l, lane );
  }
}

__global__ void __launch_bounds__( kScatterThreads ) scatter_sorted_kernel( ScatterArgs a ) {
  const ScatterChunk ch = a.chunks[blockIdx.x];
  const int n = a.n[ch.matrix];
  const int64_t e0 = a.nnz_off[ch.matrix];
  const int nnz = (int)( a.nnz_off[ch.matrix + 1] - e0 );
  double* dst = a.dense + a.a_off[ch.matrix];
  const int64_t nn = (int64_t)n * n;
  const int lane = threadIdx.x & 31;

  if ( ch.count == 0 ) {  // empty matrix: one chunk zero-fills it
    for ( int64_t q = threadIdx.x; q < nn; q += blockDim.x ) dst[q] = 0.0;
    return;
  }
  const int* ci = a.coo_i + e0;
  const int* cj = a.coo_j + e0;
  const double* cv = a.coo_v + e0;
  int flag = 0;
  const bool vec_ok = ( ( ( e0 + ch.first ) & 1 ) == 0 );  // pair loads need 8/16-byte alignment

  // each thread walks pairs (k, k+1); pairs are strided over the CTA for coalesced loads
  for ( int pr = threadIdx.x; pr * 2 < kScatterChunk; pr += kScatterThreads ) {
    const int k = ch.first + pr * 2;
    const bool have0 = ( pr * 2 ) < ch.count;
    const bool have1 = ( pr * 2 + 1 ) < ch.count;
    int i0 = 1, j0 = 1, i1 = 1, j1 = 1;
    double v0 = 0.0, v1 = 0.0;
    if ( have1 && vec_ok ) {
      const int2 ii = *reinterpret_cast<const int2*>( ci + k );
      const int2 jj = *reinterpret_cast<const int2*>( cj + k );
      const double2 vv = *reinterpret_cast<const double2*>( cv + k );
      i0 = ii.x; i1 = ii.y; j0 = jj.x; j1 = jj.y; v0 = vv.x; v1 = vv.y;
    } else {
      if ( have0 ) { i0 = ci[k]; j0 = cj[k]; v0 = cv[k]; }
      if ( have1 ) { i1 = ci[k + 1]; j1 = cj[k + 1]; v1 = cv[k + 1]; }
    }
    bool ok0 = have0, ok1 = have1;
    if ( have0 && ( i0 < 1 || i0 > n || j0 < 1 || j0 > n ) ) { flag |= kFlagBadIndex; ok0 = false; }
    if ( have1 && ( i1 < 1 || i1 > n || j1 < 1 || j1 > n ) ) { flag |= kFlagBadIndex; ok1 = false; }
    const int64_t p0 = (int64_t)( i0 - 1 ) + (int64_t)n * ( j0 - 1 );
    const int64_t p1 = (int64_t)( i1 - 1 ) + (int64_t)n * ( j1 - 1 );
    // predecessor position (-1 before the first entry of the matrix)
    int64_t pp = -1;
    if ( have0 && k > 0 ) {
      const int ip = ci[k - 1], jp = cj[k - 1];
      if ( ip < 1 || ip > n || jp < 1 || jp > n ) {
        flag |= kFlagBadIndex;
        ok0 = false;
      } else {
        pp = (int64_t)( ip - 1 ) + (int64_t)n * ( jp - 1 );
      }
    }
    if ( ok0 && p0 <= pp ) { flag |= kFlagUnsorted; ok0 = false; }
    if ( ok1 && ( !ok0 || p1 <= p0 ) ) { flag |= kFlagUnsorted; ok1 = false; }
    // gaps before each entry
    fill_gap( dst, pp, p0, ok0, lane );
    fill_gap( dst, p0, p1, ok1, lane );
    // the values: one 16-byte store when the pair is adjacent and aligned
    if ( ok0 && ok1 && p1 == p0 + 1 && ( ( reinterpret_cast<uintptr_t>( dst + p0 ) & 15 ) == 0 ) ) {
      *reinterpret_cast<double2*>( dst + p0 ) = make_double2( v0, v1 );
    } else {
      if ( ok0 ) dst[p0] = v0;
      if ( ok1 ) dst[p1] = v1;
    }
    // tail after the last entry of the matrix
    const bool last0 = ok0 && ( k == nnz - 1 );
    const bool last1 = ok1 && ( k + 1 == nnz - 1 );
    fill_gap( dst, last1 ? p1 : p0, nn, last0 || last1, lane );
  }
  if ( flag ) atomicOr( a.flags + ch.matrix, flag );
}

// ---- fallback, only for matrices whose flag has kFlagUnsorted (and no bad index) -----------------
__device__ __forceinline__ bool chunk_is_flagged( const ScatterArgs& a, const ScatterChunk& ch ) {
  return a.flags[ch.matrix] == kFlagUnsorted;
}

__global__ void __launch_bounds__( kScatterThreads ) zero_flagged_kernel( ScatterArgs a ) {
  const ScatterChunk ch = a.chunks[blockIdx.x];
  if ( !chunk_is_flagged( a, ch ) ) return;
  const int n = a.n[ch.matrix];
  const int64_t nn = (int64_t)n * n;
  const int c = ch.first / kScatterChunk;
  const int64_t lo = nn * c / ch.nchunks, hi = nn * ( c + 1 ) / ch.nchunks;
  double* dst = a.dense + a.a_off[ch.matrix];
  for ( int64_t q = lo + threadIdx.x; q < hi; q += blockDim.x ) dst[q] = 0.0;
}

template <int PASS>
__global__ void __launch_bounds__( kScatterThreads ) scatter_fallback_kernel( ScatterArgs a ) {
  const ScatterChunk ch = a.chunks[blockIdx.x];
  if ( !chunk_is_flagged( a, ch ) ) return;
  const int n = a.n[ch.matrix];
  const int64_t e0 = a.nnz_off[ch.matrix];
  unsigned long long* slot = reinterpret_cast<unsigned long long*>( a.dense + a.a_off[ch.matrix] );
  for ( int q = threadIdx.x; q < ch.count; q += blockDim.x ) {
    const int k = ch.first + q;
    int i = a.coo_i[e0 + k];
    const int j = a.coo_j[e0 + k];
    const int ia = i < 0 ? -i : i;
    const int64_t p = (int64_t)( ia - 1 ) + (int64_t)n * ( j - 1 );
    if ( PASS == 0 ) {
      atomicMax( slot + p, (unsigned long long)k + 1ull );
    } else if ( PASS == 1 ) {
      if ( slot[p] != (unsigned long long)k + 1ull ) a.coo_i[e0 + k] = -i;
    } else {
      if ( i > 0 )
        slot[p] = (unsigned long long)__double_as_longlong( a.coo_v[e0 + k] );
      else
        a.coo_i[e0 + k] = -i;
    }
  }
}

#endif  // __CUDACC__

}  // namespace pb
