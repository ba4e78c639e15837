// paladin_gpu.cu -- kernels' launch code and the C-ABI declared in include/paladin_gpu.h.
//
// Drop-in boundary: the reference's extern "C" dgeev_ (reference src/lapack-wrapper.hpp:47-60, called
// at src/paladin.hpp:248-254) plus the dense scatter of its MatrixMarket reader
// (src/square-matrix.hpp:101-117). Everything here runs on the device; there is no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/paladin_gpu.h"
#include "pb_geev.cuh"
#include "pb_hess.cuh"
#include "pb_bhess.cuh"
#include "pb_coop.cuh"
#include "pb_vecs.cuh"
#include "pb_scatter.cuh"
#include "pb_textio.cuh"

namespace {

thread_local std::string g_err;

int fail( int code, const std::string& msg ) {
  g_err = msg;
  return code;
}

#define PB_CUDA( call )                                                                           \
  do {                                                                                            \
    cudaError_t e__ = ( call );                                                                   \
    if ( e__ != cudaSuccess ) {                                                                   \
      return fail( e__ == cudaErrorMemoryAllocation ? PALADIN_GPU_ENOMEM : PALADIN_GPU_ECUDA,     \
                   std::string( #call ) + ": " + cudaGetErrorString( e__ ) );                     \
    }                                                                                             \
  } while ( 0 )

constexpr int kMaxDevices = 64;
constexpr int kSmallMax = 64;       // warp-per-matrix path for n <= kSmallMax
constexpr int kSmallBucket = 8;     // shared-memory buckets of 8 rows
constexpr int kGenericThreads = 256;

struct DeviceCtx {
  std::mutex mu;
  bool ready = false;
  cudaStream_t stream = nullptr;
  int sm_count = 0;
  int max_smem_optin = 0;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  bool attrs_set = false;  // per-function launch attributes raised to the device maximum (set_kernel_attributes)
};

DeviceCtx g_ctx[kMaxDevices];

int usable_devices() {
  int cnt = 0;
  if ( cudaGetDeviceCount( &cnt ) != cudaSuccess ) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for ( int d = 0; d < cnt && d < kMaxDevices; ++d ) {
    int major = 0;
    if ( cudaDeviceGetAttribute( &major, cudaDevAttrComputeCapabilityMajor, d ) == cudaSuccess && major == 10 ) ok = d + 1;
  }
  return ok;  // devices 0..ok-1 are addressable; entries check the capability again
}

int ensure_ctx( int device ) {
  if ( device < 0 || device >= kMaxDevices ) return fail( PALADIN_GPU_ENODEVICE, "device index out of range" );
  int cnt = 0;
  if ( cudaGetDeviceCount( &cnt ) != cudaSuccess || device >= cnt ) {
    cudaGetLastError();
    return fail( PALADIN_GPU_ENODEVICE, "no CUDA device " + std::to_string( device ) + " (there is no CPU fallback)" );
  }
  int major = 0;
  PB_CUDA( cudaDeviceGetAttribute( &major, cudaDevAttrComputeCapabilityMajor, device ) );
  if ( major != 10 )
    return fail( PALADIN_GPU_ENODEVICE, "device is not compute capability 10.x (library is built for sm_100a only)" );
  DeviceCtx& c = g_ctx[device];
  std::lock_guard<std::mutex> lk( c.mu );
  PB_CUDA( cudaSetDevice( device ) );
  if ( !c.ready ) {
    PB_CUDA( cudaStreamCreateWithFlags( &c.stream, cudaStreamNonBlocking ) );
    PB_CUDA( cudaDeviceGetAttribute( &c.sm_count, cudaDevAttrMultiProcessorCount, device ) );
    PB_CUDA( cudaDeviceGetAttribute( &c.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device ) );
    PB_CUDA( cudaEventCreate( &c.t0 ) );
    PB_CUDA( cudaEventCreate( &c.t1 ) );
    {
      // device memory of the batches comes from the stream-ordered allocator and is kept by its pool between batches:
      // cudaMalloc / cudaFree synchronise the whole device, which stalls the OTHER batches in flight on it
      cudaMemPool_t pool;
      PB_CUDA( cudaDeviceGetDefaultMemPool( &pool, device ) );
      unsigned long long keep = ~0ull;
      PB_CUDA( cudaMemPoolSetAttribute( pool, cudaMemPoolAttrReleaseThreshold, &keep ) );
    }
    c.ready = true;
  }
  return PALADIN_GPU_OK;
}

// ---- device view of a batch ------------------------------------------------------------------
struct BatchView {
  const int* n;
  const int64_t* a_off;  // dense offsets (doubles)
  const int64_t* w_off;  // eigenvalue offsets (doubles)
  double* A;
  double* V;   // right vectors / Schur vectors (nullptr when not wanted)
  double* VL;  // left vectors (nullptr when not wanted)
  double* wr;
  double* wi;
  int* info;
  double* dwork;  // 5 doubles per eigenvalue slot (generic path)
  int* iwork;     // 2 ints per eigenvalue slot + 2 per matrix
  int wantvr;
  int wantvl;
};

// ---- warp-per-matrix path, n <= 64: the matrix (and Z) live in shared memory ---------------------
// one warp per CTA; dynamic shared memory = lds*nmax doubles (x2 with vectors) + 5*nmax doubles + ints
__global__ void __launch_bounds__( 32 ) geev_warp_kernel( BatchView bv, const int* ids, int nmax, int lds ) {
  extern __shared__ double smem[];
  const int id = ids[blockIdx.x];
  const int n = bv.n[id];
  const int lane = threadIdx.x;
  const size_t msz = (size_t)lds * nmax;
  double* As = smem;
  double* p = As + msz;
  double* Vs = p;
  if ( bv.wantvr ) p += msz;
  double* Ls = p;
  if ( bv.wantvl ) p += msz;
  double* dw = p;
  int* iw = reinterpret_cast<int*>( dw + 5 * nmax );
  const double* Ag = bv.A + bv.a_off[id];
  const int nn = n * n;
  for ( int q = lane; q < nn; q += 32 ) {
    const int c = q / n, r = q - c * n;
    As[r + c * lds] = Ag[q];
  }
  __syncwarp();
  pb::WarpGroup g;
  double* wr = bv.wr + bv.w_off[id];
  double* wi = bv.wi + bv.w_off[id];
  const int info = pb::geev_group( g, n, As, lds, wr, wi, bv.wantvl != 0, Ls, lds, bv.wantvr != 0, Vs, lds, dw, iw );
  __syncwarp();
  if ( info == 0 ) {
    if ( bv.wantvr ) {
      double* Vg = bv.V + bv.a_off[id];
      for ( int q = lane; q < nn; q += 32 ) {
        const int c = q / n, r = q - c * n;
        Vg[q] = Vs[r + c * lds];
      }
    }
    if ( bv.wantvl ) {
      double* Lg = bv.VL + bv.a_off[id];
      for ( int q = lane; q < nn; q += 32 ) {
        const int c = q / n, r = q - c * n;
        Lg[q] = Ls[r + c * lds];
      }
    }
  }
  if ( lane == 0 ) bv.info[id] = info;
}

// the same with a small CTA (2 or 4 warps) per matrix: the shared-memory footprint of a matrix caps the warp-per-matrix kernel
// at 6 warps per SM for n = 64 (33 KB each); more warps on the SAME matrix raise the occupancy without raising the footprint
// (PALADIN_GPU_SMALL_WARPS)
__global__ void __launch_bounds__( 128 ) geev_smallcta_kernel( BatchView bv, const int* ids, int nmax, int lds ) {
  extern __shared__ double smem[];
  __shared__ double red[64];
  const int id = ids[blockIdx.x];
  const int n = bv.n[id];
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t msz = (size_t)lds * nmax;
  double* As = smem;
  double* p = As + msz;
  double* Vs = p;
  if ( bv.wantvr ) p += msz;
  double* Ls = p;
  if ( bv.wantvl ) p += msz;
  double* dw = p;
  int* iw = reinterpret_cast<int*>( dw + 5 * nmax );
  const double* Ag = bv.A + bv.a_off[id];
  const int nn = n * n;
  for ( int q = tid; q < nn; q += nt ) {
    const int c = q / n, r = q - c * n;
    As[r + c * lds] = Ag[q];
  }
  __syncthreads();
  pb::CtaGroup g{ red };
  double* wr = bv.wr + bv.w_off[id];
  double* wi = bv.wi + bv.w_off[id];
  const int info = pb::geev_group( g, n, As, lds, wr, wi, bv.wantvl != 0, Ls, lds, bv.wantvr != 0, Vs, lds, dw, iw );
  __syncthreads();
  if ( info == 0 ) {
    if ( bv.wantvr ) {
      double* Vg = bv.V + bv.a_off[id];
      for ( int q = tid; q < nn; q += nt ) {
        const int c = q / n, r = q - c * n;
        Vg[q] = Vs[r + c * lds];
      }
    }
    if ( bv.wantvl ) {
      double* Lg = bv.VL + bv.a_off[id];
      for ( int q = tid; q < nn; q += nt ) {
        const int c = q / n, r = q - c * n;
        Lg[q] = Ls[r + c * lds];
      }
    }
  }
  if ( tid == 0 ) bv.info[id] = info;
}

// ---- n <= 64, eigenvalues only, in two launches ---------------------------------------------------------------------
// The one-launch kernel above keeps the full matrix in shared memory from the balancing to the last QR sweep: 33 KB per
// matrix at n = 64, i.e. 6 warps per SM for an iteration that is one dependent chain per warp. Only the reduction needs the
// full matrix. Stage 1 (small_hess_kernel, one warp per matrix as before) ends with the Hessenberg matrix back in A and
// ilo / ihi / the scaling in the per-matrix meta record; stage 2 (small_qr_pair_kernel) runs the iteration with TWO
// Hessenberg matrices in one nmax x (nmax + 3) array, one warp each: matrix 0 at (r, c + 3), matrix 1 transposed at (c, r) --
// an upper Hessenberg matrix has no entry with r > c + 1, so the first occupies cells with col - row >= 2 and the second
// cells with col - row <= 1 -- which doubles the warps in flight (12 per SM at n = 64). The register-resident sweep never
// stores a bulge, so neither matrix ever touches the other's cells (pb::lahqr, hrs / band_only).
struct MidMeta;  // (defined with the mid-size path below)
#ifndef PB_SMALL_HESS_THREADS
#define PB_SMALL_HESS_THREADS 64
#endif
constexpr int kSmallHessThreads = PB_SMALL_HESS_THREADS;
__global__ void __launch_bounds__( kSmallHessThreads ) small_hess_kernel( BatchView bv, const int* ids, MidMeta* meta, int nmax, int lds );
__global__ void __launch_bounds__( 64 ) small_qr_pair_kernel( BatchView bv, const int* ids, int count, const MidMeta* meta, int nmax, int lds );

// ---- size-generic path: one CTA per matrix, in place in global memory ----------------------------
__global__ void __launch_bounds__( kGenericThreads ) geev_cta_kernel( BatchView bv, const int* ids ) {
  __shared__ double red[64];
  const int id = ids[blockIdx.x];
  const int n = bv.n[id];
  pb::CtaGroup g{ red };
  double* A = bv.A + bv.a_off[id];
  double* V = bv.wantvr ? bv.V + bv.a_off[id] : nullptr;
  double* VL = bv.wantvl ? bv.VL + bv.a_off[id] : nullptr;
  double* wr = bv.wr + bv.w_off[id];
  double* wi = bv.wi + bv.w_off[id];
  double* dw = bv.dwork + 5 * bv.w_off[id];
  int* iw = bv.iwork + 2 * bv.w_off[id] + 2 * id;
  const int info = pb::geev_group( g, n, A, n, wr, wi, bv.wantvl != 0, VL, n, bv.wantvr != 0, V, n, dw, iw );
  if ( threadIdx.x == 0 ) bv.info[id] = info;
}

// ---- mid-size path: one CTA per matrix, matrix in global memory, QR by windowed multishift sweeps ----
constexpr int kMidThreads = 256;
constexpr int kNumTicks = 16;

// AED window without block reordering (pb::aed_window). Its order is set per matrix (mid_aed_order): measured on B200,
// 592 x n=512: QR stage 392 ms without AED, 263 / 271 / 277 ms with windows of 32 / 40 / 48 rows; 148 x n=1024: 682 ms
// without, 402 / 410 ms with 32 / 48 (the strips grow with n, the window's own QR iteration does not).
__host__ __device__ inline pb::MsqrCfg mid_cfg() { return pb::MsqrCfg{ 8, pb::kStripW, pb::kStripW - 1, 128, 2, 32, 32, 0 }; }
__host__ __device__ inline int mid_aed_order( int n ) { return n <= 1024 ? 32 : pb::kStripW - 1; }

constexpr int kMidMaxFusedN = 1024;  // warp-level kernels keep a matrix column in 32 * MAXCH registers per warp
constexpr int kMidMaxBigN = 4608;    // streaming variants: 6 n + 64 doubles of shared memory

// per-matrix state handed from stage to stage (global memory)
struct MidMeta {
  int ilo, ihi, info, scalea;
  double anrm, cscale;
};

// (kSmallHessThreads threads per matrix: the reduction's row / column updates are 64 wide at n = 64, and the shared-memory
// footprint -- the limit on matrices per SM -- does not grow with the threads)
__global__ void __launch_bounds__( kSmallHessThreads ) small_hess_kernel( BatchView bv, const int* ids, MidMeta* meta, int nmax, int lds ) {
  extern __shared__ double smem[];
  __shared__ double red[64];
  const int id = ids[blockIdx.x];
  const int n = bv.n[id];
  const int lane = threadIdx.x;
  constexpr int kT = kSmallHessThreads;
  double* As = smem;
  double* dw = As + (size_t)lds * nmax;
  int* iw = reinterpret_cast<int*>( dw + 5 * nmax );
  double* Ag = bv.A + bv.a_off[id];
  const int nn = n * n;
  for ( int q = lane; q < nn; q += kT ) {
    const int c = q / n, r = q - c * n;
    As[r + c * lds] = Ag[q];
  }
  __syncthreads();
  pb::CtaGroup g{ red };
  double* wr = bv.wr + bv.w_off[id];
  double* wi = bv.wi + bv.w_off[id];
  MidMeta m;
  m.ilo = 0;
  m.ihi = n - 1;
  m.info = 0;
  m.scalea = 0;
  m.anrm = 0.0;
  m.cscale = 1.0;
  if ( n > 0 ) {
    pb::GeevScaling sc;
    if ( !pb::geev_prologue( g, n, As, lds, sc, &m.ilo, &m.ihi, dw + nmax, iw ) ) {
      m.info = PALADIN_GPU_INFO_NONFINITE;
    } else {
      m.scalea = sc.scalea ? 1 : 0;
      m.anrm = sc.anrm;
      m.cscale = sc.cscale;
      pb::gehd2( g, n, m.ilo, m.ihi, As, lds, dw, dw + 2 * nmax );
      __syncthreads();
      pb::geev_isolated_eigenvalues( g, n, m.ilo, m.ihi, As, lds, wr, wi );
      for ( int q = lane; q < nn; q += kT ) {
        const int c = q / n, r = q - c * n;
        if ( r <= c + 1 ) Ag[q] = As[r + c * lds];
      }
    }
  }
  if ( lane == 0 ) {
    meta[id] = m;
    if ( m.info != 0 ) bv.info[id] = m.info;
  }
}

__global__ void __launch_bounds__( 64 ) small_qr_pair_kernel( BatchView bv, const int* ids, int count, const MidMeta* meta, int nmax, int lds ) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = 2 * blockIdx.x + warp;
  if ( k >= count ) return;  // (the warps of a pair never meet at a barrier)
  const int id = ids[k];
  const int n = bv.n[id];
  const MidMeta m = meta[id];
  if ( m.info != 0 || n == 0 ) {
    if ( lane == 0 && m.info == 0 ) bv.info[id] = 0;
    return;
  }
  // warp 0: H(r, c) at cell (r, c + 3); warp 1: H(r, c) at cell (c, r)   (cell (row, col) at smem[row + col * lds])
  double* H = warp == 0 ? smem + 3 * (size_t)lds : smem;
  const int hrs = warp == 0 ? 1 : lds, ldh = warp == 0 ? lds : 1;
  const double* Ag = bv.A + bv.a_off[id];
  const int nn = n * n;
  for ( int q = lane; q < nn; q += 32 ) {
    const int c = q / n, r = q - c * n;
    if ( r <= c + 1 ) H[(size_t)r * hrs + (size_t)c * ldh] = Ag[q];
  }
  __syncwarp();
  pb::WarpGroup g;
  double* wr = bv.wr + bv.w_off[id];
  double* wi = bv.wi + bv.w_off[id];
  const int info = pb::lahqr<pb::WarpGroup, true, pb::LahqrNoHook, true>( g, false, false, n, m.ilo, m.ihi, H, ldh, wr, wi, 0, n - 1, (double*)nullptr, 1, pb::LahqrNoHook(), hrs, true );
  __syncwarp();
  pb::GeevScaling sc{ m.scalea != 0, m.anrm, m.cscale };
  pb::geev_unscale( g, n, sc, info, m.ilo, wr, wi );
  if ( lane == 0 ) bv.info[id] = info;
}

struct MidArgs {
  BatchView bv;
  const int* ids;
  int count;
  MidMeta* meta;
  long long* stats;
  int* counter;    // work queue head of the persistent vectors kernel
  double* xwork;   // slots of xn_max^2 doubles for the triangular eigenvector factor
  size_t xslot;    // doubles per slot
  int xn_max;      // largest order a slot can hold (larger matrices use the in-place path)
  int64_t xhalf;   // > 0: the three-launch path addresses X like the dense matrices (xwork + a_off[id]); the second factor
                   // (both sides wanted) lives xhalf doubles further
  int refl_global;  // clusters: reflectors through global memory (L2) instead of the leader's shared memory
  int aed, aed_moves, nibble;  // tuning overrides of mid_cfg() (PALADIN_GPU_AED, _AED_MOVES, _AED_NIBBLE); -1 = default
  int itmax;                   // PALADIN_GPU_QR_ITMAX: cap on the QR sweeps of a matrix (0 = LAPACK's 30 * max(10, nh))
  int spec_shifts;             // PALADIN_GPU_SPEC_SHIFTS: 0 = shifts after the AED step, 1 = before it, 2 = side by side with it (second warp)
  // blocked Hessenberg reduction (pb_bhess.cuh): global scratch per matrix, tensor maps of A (2 id) and of the Schur-vector
  // array (2 id + 1), and whether a matrix has them (even leading dimension, encoder succeeded)
  double* hwork;
  const int64_t* h_off;
  const CUtensorMap* tmaps;
  const int* tmap_ok;
  int hess_tma;                // 0: stage the update tiles with plain loads (PALADIN_GPU_HESS=notma)
  int formq_unblocked;         // 1: explicit Q by the register-resident unblocked accumulation (pb::form_q) instead of the block reflectors
};

__device__ __forceinline__ void mid_pointers( const BatchView& bv, int id, int n, double*& A, double*& V, double*& wr, double*& wi,
                                              double*& dw, int*& iw ) {
  A = bv.A + bv.a_off[id];
  // Schur vectors are accumulated in the right-vector array when it exists, else in the left-vector array
  V = bv.wantvr ? bv.V + bv.a_off[id] : ( bv.wantvl ? bv.VL + bv.a_off[id] : nullptr );
  wr = bv.wr + bv.w_off[id];
  wi = bv.wi + bv.w_off[id];
  dw = bv.dwork + 5 * bv.w_off[id];
  iw = bv.iwork + 2 * bv.w_off[id] + 2 * id;
}

__device__ __forceinline__ void stats_begin( pb::CtaGroup& g, long long* stats, long long* lstats ) {
  if ( threadIdx.x < kNumTicks ) lstats[threadIdx.x] = 0;
  __syncthreads();
  if ( stats != nullptr ) {
    g.stats = lstats;
    g.last = clock64();
  }
}
__device__ __forceinline__ void stats_end( long long* stats, const long long* lstats ) {
  __syncthreads();
  if ( stats != nullptr && threadIdx.x < kNumTicks && lstats[threadIdx.x] != 0 )
    atomicAdd( reinterpret_cast<unsigned long long*>( stats ) + threadIdx.x, (unsigned long long)lstats[threadIdx.x] );
}

// ---- stage 1: scale check, balancing, Hessenberg reduction, explicit Q ----------------------------------
inline size_t prep_smem_bytes( int nmax ) {
  // a group may mix sizes on both sides of kMidMaxFusedN: size for the largest matrix the fused kernel can meet
  return sizeof( double ) * pb::hess_smem_doubles( std::min( nmax, kMidMaxFusedN ) );
}

// MAXCH = 16: n <= 512, two CTAs per SM (<= 128 registers); MAXCH = 32: n <= 1024; MAXCH = 0: any n, unblocked
// group code in global memory (slow, correctness path for the largest matrices)
template <int MAXCH>
__global__ void __launch_bounds__( pb::kHessThreads, ( MAXCH == 16 ? 2 : 1 ) ) mid_prep_kernel( MidArgs a ) {
  extern __shared__ double smem[];
  __shared__ double red[64];
  __shared__ long long lstats[kNumTicks];
  const int id = a.ids[blockIdx.x];
  const int n = a.bv.n[id];
  pb::CtaGroup g{ red };
  stats_begin( g, a.stats, lstats );
  double *A, *V, *wr, *wi, *dw;
  int* iw;
  mid_pointers( a.bv, id, n, A, V, wr, wi, dw, iw );
  double* tau = dw;
  double* scale = dw + n;
  pb::GeevScaling sc;
  int ilo = 0, ihi = n - 1, info = 0;
  // (the dynamic shared memory is free until the reduction starts: row / column sums of the balancing sweeps)
  if ( !pb::geev_prologue( g, n, A, n, sc, &ilo, &ihi, scale, iw, MAXCH > 0 || n <= kMidMaxBigN ? smem : nullptr ) ) {
    info = PALADIN_GPU_INFO_NONFINITE;
    sc.scalea = false;
    sc.anrm = 0.0;
    sc.cscale = 1.0;
  } else {
    g.tick( pb::kTickBalance );
    if constexpr ( MAXCH > 0 ) {
      pb::hess_fused<MAXCH>( n, ilo, ihi, A, n, tau, smem );
    } else {
      if ( n <= kMidMaxBigN ) pb::hess_fused_big( n, ilo, ihi, A, n, tau, smem );
      else pb::gehd2( g, n, ilo, ihi, A, n, tau, dw + 2 * n );
    }
    __syncthreads();
    g.tick( pb::kTickHess );
    if ( a.bv.wantvr || a.bv.wantvl ) {
      if constexpr ( MAXCH > 0 ) {
        pb::form_q<MAXCH>( n, ilo, ihi, A, n, tau, V, n, smem );
      } else {
        if ( n <= kMidMaxBigN ) pb::form_q_big( n, ilo, ihi, A, n, tau, V, n, smem );
        else pb::orghr( g, n, ilo, ihi, A, n, tau, V, n );
      }
    }
    __syncthreads();
    g.tick( pb::kTickOrghr );
    pb::geev_isolated_eigenvalues( g, n, ilo, ihi, A, n, wr, wi );
  }
  if ( threadIdx.x == 0 ) {
    MidMeta m;
    m.ilo = ilo;
    m.ihi = ihi;
    m.info = info;
    m.scalea = sc.scalea ? 1 : 0;
    m.anrm = sc.anrm;
    m.cscale = sc.cscale;
    a.meta[id] = m;
    a.bv.info[id] = info;
  }
  stats_end( a.stats, lstats );
}


// ---- stage 1, blocked: scale check, balancing, dlahr2-panel Hessenberg reduction with DMMA / TMA-staged trailing updates,
// blocked explicit Q (pb_bhess.cuh). MAXCH = 16: n <= 512, two CTAs per SM; MAXCH = 32: n <= 1024 -----------------------
inline size_t bprep_smem_bytes( int nmax ) { return sizeof( double ) * pb::bh_smem_doubles( nmax ); }

template <int MAXCH>
#ifndef PB_BH_OCC
#define PB_BH_OCC 2
#endif
__global__ void __launch_bounds__( pb::kBhThreads, ( MAXCH == 16 ? PB_BH_OCC : 1 ) ) mid_prep_blocked_kernel( MidArgs a ) {
  extern __shared__ __align__( 128 ) double smem[];
  __shared__ double red[64];
  __shared__ long long lstats[kNumTicks];
  __shared__ __align__( 8 ) uint64_t bars[pb::kBhMaxStages];
  const int id = a.ids[blockIdx.x];
  const int n = a.bv.n[id];
  pb::CtaGroup g{ red };
  stats_begin( g, a.stats, lstats );
  if ( threadIdx.x == 0 ) {
    for ( int q = 0; q < pb::kBhMaxStages; ++q ) pb::bh_mbar_init( &bars[q], 1 );
    asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
  }
  __syncthreads();
  double *A, *V, *wr, *wi, *dw;
  int* iw;
  mid_pointers( a.bv, id, n, A, V, wr, wi, dw, iw );
  double* tau = dw;
  double* scale = dw + n;
  pb::GeevScaling sc;
  int ilo = 0, ihi = n - 1, info = 0;
  if ( !pb::geev_prologue( g, n, A, n, sc, &ilo, &ihi, scale, iw, smem ) ) {
    info = PALADIN_GPU_INFO_NONFINITE;
    sc.scalea = false;
    sc.anrm = 0.0;
    sc.cscale = 1.0;
  } else {
    g.tick( pb::kTickBalance );
    __syncthreads();
    pb::BhSmem sh = pb::bh_carve( smem, n, bars );
    const bool tma = a.hess_tma != 0 && a.tmap_ok[id] != 0;
    double* scratch = a.hwork + a.h_off[id];
    pb::bh_gehrd<MAXCH>( g, n, ilo, ihi, A, n, tau, scratch, tma ? a.tmaps + 2 * (size_t)id : nullptr, sh );
    __syncthreads();
    g.tick( pb::kTickHess );
    if ( a.bv.wantvr || a.bv.wantvl ) {
      if ( a.formq_unblocked ) pb::form_q<MAXCH>( n, ilo, ihi, A, n, tau, V, n, smem );
      else pb::bh_form_q( n, ilo, ihi, A, n, V, n, scratch, tma ? a.tmaps + 2 * (size_t)id + 1 : nullptr, sh );
    }
    __syncthreads();
    g.tick( pb::kTickOrghr );
    pb::geev_isolated_eigenvalues( g, n, ilo, ihi, A, n, wr, wi );
  }
  if ( threadIdx.x == 0 ) {
    MidMeta m;
    m.ilo = ilo;
    m.ihi = ihi;
    m.info = info;
    m.scalea = sc.scalea ? 1 : 0;
    m.anrm = sc.anrm;
    m.cscale = sc.cscale;
    a.meta[id] = m;
    a.bv.info[id] = info;
  }
  stats_end( a.stats, lstats );
}


// ---- stage 1, blocked, n > 1024: one matrix per thread-block cluster of S CTAs (S = 1 ... 16 by how many large matrices the
// launch holds), pb_bhess.cuh::bhc_* -------------------------------------------------------------------------------------------
inline size_t bcprep_smem_bytes( int nmax ) { return sizeof( double ) * pb::bh_smem_doubles( nmax ); }

__global__ void __launch_bounds__( pb::kBhThreads, 1 ) big_prep_blocked_kernel( MidArgs a, int csize ) {
  extern __shared__ __align__( 128 ) double smem[];
  __shared__ double red[64];
  __shared__ long long lstats[kNumTicks];
  __shared__ __align__( 8 ) uint64_t bars[pb::kBhMaxStages];
  pb::BhPart pt;
  pt.size = csize;
  pt.rank = csize > 1 ? (int)cooperative_groups::this_cluster().block_rank() : 0;
  const int id = a.ids[blockIdx.x / csize];
  const int n = a.bv.n[id];
  pb::CtaGroup g{ red };
  stats_begin( g, pt.rank == 0 ? a.stats : nullptr, lstats );
  if ( threadIdx.x == 0 ) {
    for ( int q = 0; q < pb::kBhMaxStages; ++q ) pb::bh_mbar_init( &bars[q], 1 );
    asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
  }
  __syncthreads();
  double *A, *V, *wr, *wi, *dw;
  int* iw;
  mid_pointers( a.bv, id, n, A, V, wr, wi, dw, iw );
  double* tau = dw;
  double* scale = dw + n;
  if ( pt.rank == 0 ) {
    // scale check + balancing: one CTA per matrix (a few passes over it)
    pb::GeevScaling sc;
    int ilo = 0, ihi = n - 1, info = 0;
    if ( !pb::geev_prologue( g, n, A, n, sc, &ilo, &ihi, scale, iw, smem ) ) {
      info = PALADIN_GPU_INFO_NONFINITE;
      sc.scalea = false;
      sc.anrm = 0.0;
      sc.cscale = 1.0;
    }
    if ( threadIdx.x == 0 ) {
      MidMeta m;
      m.ilo = ilo;
      m.ihi = ihi;
      m.info = info;
      m.scalea = sc.scalea ? 1 : 0;
      m.anrm = sc.anrm;
      m.cscale = sc.cscale;
      a.meta[id] = m;
      a.bv.info[id] = info;
    }
    __threadfence();
  }
  pb::bh_part_sync( pt );
  g.tick( pb::kTickBalance );
  const int4 mi = __ldcg( reinterpret_cast<const int4*>( &a.meta[id] ) );
  const int ilo = mi.x, ihi = mi.y, info = mi.z;
  if ( info == 0 ) {
    pb::BhSmem sh = pb::bh_carve( smem, n, bars );
    const bool tma = a.hess_tma != 0 && a.tmap_ok[id] != 0;
    double* scratch = a.hwork + a.h_off[id];
    pb::bhc_gehrd( g, n, ilo, ihi, A, n, tau, scratch, tma ? a.tmaps + 2 * (size_t)id : nullptr, sh, pt );
    if ( a.bv.wantvr || a.bv.wantvl ) pb::bhc_form_q( n, ilo, ihi, A, n, V, n, scratch, tma ? a.tmaps + 2 * (size_t)id + 1 : nullptr, sh, pt );
    g.tick( pb::kTickOrghr );
    if ( pt.rank == 0 ) pb::geev_isolated_eigenvalues( g, n, ilo, ihi, A, n, wr, wi );
  }
  stats_end( pt.rank == 0 ? a.stats : nullptr, lstats );
}


// ---- stage 1 for large matrices (kMidMaxFusedN < n <= kMidMaxBigN): groups of S cooperating CTAs, one matrix per group
// at a time (pb_coop.cuh); cooperative launch, one CTA per SM ------------------------------------------------------------
constexpr int kCoopMaxGroups = 37;
// smallest order whose eigenvector stage runs as three grid-wide launches instead of the persistent one-CTA-per-matrix
// kernel (tuning: PALADIN_GPU_BIGVEC_MIN_N). Measured on B200, vectors stage per step: 1184 x n=200 9.9 -> 7.3 ms,
// 592 x n=512 51.8 -> 35.2 ms, 148 x n=1024 154 -> 63.5 ms: the solves run at twice the warps per SM.
inline int bigvec_min_n() {
  static const int v = std::getenv( "PALADIN_GPU_BIGVEC_MIN_N" ) ? std::max( 64, std::atoi( std::getenv( "PALADIN_GPU_BIGVEC_MIN_N" ) ) ) : 64;
  return v;
}
// smallest order that takes the cooperative stage 1 (tuning: PALADIN_GPU_COOP_MIN_N)
inline int coop_min_n() {
  static const int v = std::getenv( "PALADIN_GPU_COOP_MIN_N" ) ? std::max( 64, std::atoi( std::getenv( "PALADIN_GPU_COOP_MIN_N" ) ) ) : kMidMaxFusedN;
  return v;
}
struct CoopArgs {
  unsigned* bars;  // 2 words per group, 2 more for the whole grid
  double* ypart;   // S * nmax doubles per group
  double* yz;      // 2 * nmax doubles per group
  int S, G, nmax;
};

__global__ void __launch_bounds__( pb::kCoopThreads, 1 ) big_prep_kernel( MidArgs a, CoopArgs ca ) {
  extern __shared__ double smem[];
  __shared__ double red[64];
  __shared__ long long lstats[kNumTicks];
  const int group = blockIdx.x / ca.S, rank = blockIdx.x % ca.S;
  const pb::CoopGroup cg{ ca.bars + 2 * group, ca.S, rank };
  pb::CtaGroup g{ red };
  stats_begin( g, rank == 0 ? a.stats : nullptr, lstats );
  double* ypart = ca.ypart + (size_t)group * ca.S * ca.nmax;
  double* yz = ca.yz + (size_t)group * 2 * ca.nmax;
  // scale check and balancing: one CTA per matrix (a few passes over it), all matrices of the launch side by side
  for ( int k = blockIdx.x; k < a.count; k += gridDim.x ) {
    const int id = a.ids[k];
    const int n = a.bv.n[id];
    double *A, *V, *wr, *wi, *dw;
    int* iw;
    mid_pointers( a.bv, id, n, A, V, wr, wi, dw, iw );
    pb::GeevScaling sc;
    int ilo = 0, ihi = n - 1, info = 0;
    if ( !pb::geev_prologue( g, n, A, n, sc, &ilo, &ihi, dw + n, iw, smem ) ) {
      info = PALADIN_GPU_INFO_NONFINITE;
      sc.scalea = false;
      sc.anrm = 0.0;
      sc.cscale = 1.0;
    }
    if ( threadIdx.x == 0 ) {
      MidMeta m;
      m.ilo = ilo;
      m.ihi = ihi;
      m.info = info;
      m.scalea = sc.scalea ? 1 : 0;
      m.anrm = sc.anrm;
      m.cscale = sc.cscale;
      a.meta[id] = m;
      a.bv.info[id] = info;
    }
    __syncthreads();
  }
  pb::coop_sync( pb::CoopGroup{ ca.bars + 2 * ca.G, (int)gridDim.x, (int)blockIdx.x } );
  g.tick( pb::kTickBalance );
  for ( int k = group; k < a.count; k += ca.G ) {
    const int id = a.ids[k];
    const int n = a.bv.n[id];
    double *A, *V, *wr, *wi, *dw;
    int* iw;
    mid_pointers( a.bv, id, n, A, V, wr, wi, dw, iw );
    double* tau = dw;
    const int4 mi = __ldcg( reinterpret_cast<const int4*>( &a.meta[id] ) );
    const int ilo = mi.x, ihi = mi.y, info = mi.z;
    if ( info == 0 ) {
      pb::coop_hess( cg, n, ilo, ihi, A, n, tau, smem, ypart, yz );
      pb::coop_sync( cg );
      g.tick( pb::kTickHess );
      if ( a.bv.wantvr || a.bv.wantvl ) pb::coop_form_q( cg, n, ilo, ihi, A, n, tau, V, n, smem );
      if ( rank == 0 ) pb::geev_isolated_eigenvalues( g, n, ilo, ihi, A, n, wr, wi );
    }
    pb::coop_sync( cg );
    g.tick( pb::kTickOrghr );
  }
  stats_end( rank == 0 ? a.stats : nullptr, lstats );
}

// ---- stage 2: QR iteration (windowed multishift sweeps) --------------------------------------------------
inline size_t qr_smem_bytes() {
  const pb::MsqrCfg c = mid_cfg();
  return sizeof( double ) * pb::msqr_work_doubles( c, c.W | 1, c.tl | 1 ) + sizeof( int ) * pb::kMsIbufInts;
}

// (measured and rejected: four 128-thread CTAs per SM -- window inside the tile buffer, U and reflectors sharing a buffer,
// two bulges per warp in the chase: 45 KB per CTA -- QR stage 407 ms per 592 matrices of n=512 against 271 ms with two
// 256-thread CTAs; the same 16 warps per SM either way, and the chase slows down by more than the window phases gain)
// CLUSTER: launched with a thread-block cluster per matrix (large matrices, few of them): rank 0 runs the iteration and
// shares the strip updates of its steady-state windows with the other CTAs of the cluster (pb::ms_cluster_helper)
template <bool CLUSTER>
__global__ void __launch_bounds__( kMidThreads, 2 ) mid_qr_kernel( MidArgs a ) {
  extern __shared__ double smem[];
  __shared__ double red[64];
  __shared__ long long lstats[kNumTicks];
  pb::MsqrCfg c = mid_cfg();
  if ( a.aed >= 0 ) c.aed = a.aed;
  if ( a.aed_moves >= 0 ) c.aed_moves = a.aed_moves;
  if ( a.nibble >= 0 ) c.nibble = a.nibble;
  c.itmax = a.itmax;
  c.spec_shifts = a.spec_shifts;
  const int ldw = c.W | 1, ldt = c.tl | 1;
  pb::MsqrWork wk;
  double* p = smem;
  wk.Hw = p; p += pb::ms_even( (size_t)c.W * ldw );
  wk.U = p; p += pb::ms_even( (size_t)c.W * ldw );
  wk.tile = p; p += pb::ms_even( (size_t)c.W * ldt );
  wk.refl = p; p += (size_t)c.nb * c.W * pb::kReflStride;  // 16-byte aligned: read as double2
  wk.sr = p; p += 2 * c.nb * c.nchains;
  wk.si = p; p += 2 * c.nb * c.nchains;
  wk.ibuf = reinterpret_cast<int*>( p );
  wk.ldw = ldw;
  wk.ldt = ldt;
  if ( threadIdx.x < pb::kMsIbufInts ) wk.ibuf[threadIdx.x] = 0;
  int mat = blockIdx.x;
  if constexpr ( CLUSTER ) {
    cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
    wk.csize = (int)cl.num_blocks();
    wk.crank = (int)cl.block_rank();
    mat = blockIdx.x / wk.csize;
  }
  const int id = a.ids[mat];
  const int n = a.bv.n[id];
  if ( a.aed < 0 ) c.aed = mid_aed_order( n );
  pb::CtaGroup g{ red };
  stats_begin( g, wk.crank == 0 ? a.stats : nullptr, lstats );
  const MidMeta m = a.meta[id];
  if ( m.info != 0 ) return;  // (the whole cluster)
  double *A, *V, *wr, *wi, *dw;
  int* iw;
  mid_pointers( a.bv, id, n, A, V, wr, wi, dw, iw );
  const bool wantz = a.bv.wantvr != 0 || a.bv.wantvl != 0;
  if constexpr ( CLUSTER ) {
    // (tau, scale | 3 n doubles free until the eigenvector stage: the reflectors of a window, 2352 doubles, for n >= 784)
    if ( a.refl_global && 3 * n >= c.nb * c.W * pb::kReflStride + 1 ) {
      wk.gscratch = dw + 2 * n;
      if ( reinterpret_cast<uintptr_t>( wk.gscratch ) & 15 ) wk.gscratch += 1;  // read and written as double2
    }
    wk.overlap = a.refl_global >= 2 ? 0 : 1;  // PALADIN_GPU_REFL_GLOBAL=2: published reflectors, blocking hand-off
    if ( wk.crank != 0 ) {
      pb::ms_cluster_helper( g, wantz, wantz, n, A, n, 0, n - 1, V, n, c, wk );
      return;
    }
  }
  const int info = pb::hqr_multishift( g, wantz, wantz, n, m.ilo, m.ihi, A, n, wr, wi, 0, n - 1, V, n, c, wk );
  if constexpr ( CLUSTER ) pb::ms_cluster_release( wk );
  __syncthreads();
  g.tick( pb::kTickQr );
  pb::GeevScaling sc{ m.scalea != 0, m.anrm, m.cscale };
  pb::geev_unscale( g, n, sc, info, m.ilo, wr, wi );
  if ( threadIdx.x == 0 ) {
    a.meta[id].info = info;
    a.bv.info[id] = info;
  }
  stats_end( wk.crank == 0 ? a.stats : nullptr, lstats );
}

// ---- stage 3: eigenvectors (persistent CTAs, one work slot each) -----------------------------------------
inline size_t vec_smem_bytes( int nmax ) { return sizeof( double ) * pb::vecs_gemm_smem_doubles() + sizeof( int ) * nmax; }

__global__ void __launch_bounds__( pb::kVecThreads, 2 ) mid_vec_kernel( MidArgs a ) {
  extern __shared__ double smem[];
  __shared__ double red[64];
  __shared__ long long lstats[kNumTicks];
  __shared__ int s_next;
  int* perm = reinterpret_cast<int*>( smem + pb::vecs_gemm_smem_doubles() );
  pb::CtaGroup g{ red };
  stats_begin( g, a.stats, lstats );
  double* X = a.xwork + (size_t)blockIdx.x * a.xslot;
  for ( ;; ) {
    __syncthreads();
    if ( threadIdx.x == 0 ) s_next = atomicAdd( a.counter, 1 );
    __syncthreads();
    const int k = s_next;
    if ( k >= a.count ) break;
    const int id = a.ids[k];
    const int n = a.bv.n[id];
    const MidMeta m = a.meta[id];
    if ( m.info != 0 ) continue;
    double *A, *V, *wr, *wi, *dw;
    int* iw;
    mid_pointers( a.bv, id, n, A, V, wr, wi, dw, iw );
    double* scale = dw + n;
    double* cnorm = dw + 2 * n;
    const bool wantvr = a.bv.wantvr != 0, wantvl = a.bv.wantvl != 0;
    double* VLout = wantvl ? a.bv.VL + a.bv.a_off[id] : nullptr;
    if ( n <= a.xn_max ) {
      // V = Schur vectors Z (right-vector array if it exists, else the left-vector array); T = A
      double* Y = X + (size_t)n * n + 16;  // second half of the slot (both sides wanted)
      if ( wantvr ) pb::vecs_solve( n, A, n, X, n, cnorm, false );
      if ( wantvl ) pb::vecs_solve( n, A, n, wantvr ? Y : X, n, cnorm, true );
      g.tick( pb::kTickTrevc );
      // T is dead once every solve of this matrix finished
      if ( wantvr ) {
        pb::vecs_gemm( n, V, n, X, n, A, n, smem, false );         // A <- Z X
        __syncthreads();
        if ( wantvl ) {
          pb::vecs_gemm( n, V, n, Y, n, X, n, smem, true );        // X <- Z Y   (X is free again)
          __syncthreads();
        }
      } else {
        pb::vecs_gemm( n, V, n, X, n, A, n, smem, true );          // A <- Z Y
        __syncthreads();
      }
      g.tick( pb::kTickVecGemm );
      // Z is dead: the right vectors go where it was
      for ( int side = 0; side < 2; ++side ) {
        const bool left = side == 1;
        if ( left ? !wantvl : !wantvr ) continue;
        const double* src = ( left && wantvr ) ? X : A;
        double* dst = left ? VLout : V;
        if ( n <= 512 ) pb::vecs_finish<16>( n, m.ilo, m.ihi, scale, wi, src, n, dst, n, perm, left );
        else if ( n <= kMidMaxFusedN ) pb::vecs_finish<32>( n, m.ilo, m.ihi, scale, wi, src, n, dst, n, perm, left );
        else pb::vecs_finish_big( n, m.ilo, m.ihi, scale, wi, src, n, dst, n, perm, left );
        __syncthreads();
      }
      g.tick( pb::kTickBak );
    } else {
      // in-place path for matrices no work slot can hold
      if ( wantvl && wantvr ) {
        for ( size_t q = threadIdx.x; q < (size_t)n * n; q += blockDim.x ) VLout[q] = V[q];
        __syncthreads();
      }
      double* Lz = wantvl ? VLout : nullptr;
      if ( wantvl ) pb::trevc_left_inplace( g, n, A, n, Lz, n, cnorm, dw + 3 * n, dw + 4 * n );
      if ( wantvr ) pb::trevc_right_inplace( g, n, A, n, V, n, cnorm, dw + 3 * n, dw + 4 * n );
      g.tick( pb::kTickTrevc );
      if ( wantvl ) {
        pb::gebak_vectors( g, true, n, m.ilo, m.ihi, scale, n, Lz, n );
        __syncthreads();
        pb::normalize_vectors( g, n, wi, Lz, n );
      }
      if ( wantvr ) {
        pb::gebak_vectors( g, false, n, m.ilo, m.ihi, scale, n, V, n );
        __syncthreads();
        pb::normalize_vectors( g, n, wi, V, n );
      }
      g.tick( pb::kTickBak );
    }
  }
  stats_end( a.stats, lstats );
}


// ---- stage 3 for large matrices: every matrix is shared by P CTAs; the three steps are separate launches (the kernel
// boundary is the barrier between "all solves done" -> GEMM -> normalise). STEP 0: triangular solves, 1: first GEMM,
// 2: second GEMM (both sides wanted), 3: un-balance + normalise ------------------------------------------------------------
// (the solves are latency bound -- one dependent chain per warp -- and run at four CTAs per SM / 64 registers; measured
// and rejected at 592 x n=512: the vector being solved in shared memory, 64 KB per CTA hence three CTAs per SM, 48.3 ms
// against 34.8 ms; L1 prefetch of the next columns of T, no change: warps in flight are what counts)
template <int STEP>
__global__ void __launch_bounds__( pb::kVecThreads, STEP == 0 ? 4 : 2 ) big_vec_kernel( MidArgs a, const int4* items ) {
  extern __shared__ double smem[];
  int* perm = reinterpret_cast<int*>( smem + pb::vecs_gemm_smem_doubles() );
  // work item = (matrix of the group, part, parts of that matrix): a matrix is shared by a number of CTAs proportional to n^3
  const int4 item = items[blockIdx.x];
  const int k = item.x, part = item.y, P = item.z;
  const int id = a.ids[k];
  const int n = a.bv.n[id];
  const MidMeta m = a.meta[id];
  if ( m.info != 0 ) return;
  double *A, *V, *wr, *wi, *dw;
  int* iw;
  mid_pointers( a.bv, id, n, A, V, wr, wi, dw, iw );
  double* scale = dw + n;
  double* cnorm = dw + 2 * n;
  const bool wantvr = a.bv.wantvr != 0, wantvl = a.bv.wantvl != 0;
  double* VLout = wantvl ? a.bv.VL + a.bv.a_off[id] : nullptr;
  double* X = a.xhalf > 0 ? a.xwork + a.bv.a_off[id] : a.xwork + (size_t)k * a.xslot;
  double* Y = a.xhalf > 0 ? X + a.xhalf : X + (size_t)n * n + 16;
  if constexpr ( STEP == 0 ) {
    if ( wantvr ) pb::vecs_solve( n, A, n, X, n, cnorm, false, part, P );
    if ( wantvl ) pb::vecs_solve( n, A, n, wantvr ? Y : X, n, cnorm, true, part, P );
  } else if constexpr ( STEP == 1 ) {
    if ( wantvr ) pb::vecs_gemm( n, V, n, X, n, A, n, smem, false, part, P );  // A <- Z X
    else pb::vecs_gemm( n, V, n, X, n, A, n, smem, true, part, P );            // A <- Z Y
  } else if constexpr ( STEP == 2 ) {
    pb::vecs_gemm( n, V, n, Y, n, X, n, smem, true, part, P );                 // X <- Z Y (X is free again)
  } else {
    for ( int side = 0; side < 2; ++side ) {
      const bool left = side == 1;
      if ( left ? !wantvl : !wantvr ) continue;
      const double* src = ( left && wantvr ) ? X : A;
      double* dst = left ? VLout : V;
      if ( n <= 512 ) pb::vecs_finish<16>( n, m.ilo, m.ihi, scale, wi, src, n, dst, n, perm, left, part, P );
      else if ( n <= kMidMaxFusedN ) pb::vecs_finish<32>( n, m.ilo, m.ihi, scale, wi, src, n, dst, n, perm, left, part, P );
      else pb::vecs_finish_big( n, m.ilo, m.ihi, scale, wi, src, n, dst, n, perm, left, part, P );
      __syncthreads();
    }
  }
}

// ---- eigenvector write filter (reference src/spectrum-util.hpp:134-196): per eigenvalue index the largest |x|^2 of the
// unpacked complex vector and the number of entries above threshold * that. One warp per column. -----------------------
__global__ void __launch_bounds__( 256 ) vector_filter_kernel( BatchView bv, const double* vecs, double threshold, double* max2,
                                                               unsigned long long* kept, int count ) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.y;
  if ( k >= count ) return;
  const int n = bv.n[k];
  const int c = blockIdx.x * 8 + warp;
  if ( c >= n || bv.info[k] != 0 ) return;
  const double* wi = bv.wi + bv.w_off[k];
  // the reference decides "pair" sequentially (skipNext, src/paladin.hpp:257-289): replay it up to column c
  int role = 0;  // 0 real, 1 first of a pair, 2 second of a pair
  {
    bool second = false;
    for ( int q = 0; q <= c; ++q ) {
      if ( second ) {
        role = 2;
        second = false;
      } else {
        const bool pair = ( q + 1 < n ) && ( wi[q] == -wi[q + 1] ) && ( wi[q] != 0.0 );
        role = pair ? 1 : 0;
        second = pair;
      }
    }
  }
  const double* V = vecs + bv.a_off[k];
  const double* re = V + (size_t)( role == 2 ? c - 1 : c ) * n;
  const double* im = role == 0 ? nullptr : V + (size_t)( role == 2 ? c : c + 1 ) * n;
  double mx = 0.0;
  for ( int r = lane; r < n; r += 32 ) {
    const double x = re[r], y = im ? im[r] : 0.0;
    const double nrm = __dadd_rn( __dmul_rn( x, x ), __dmul_rn( y, y ) );
    mx = nrm > mx ? nrm : mx;
  }
  mx = pb::warp_max( mx );
  const double cut = __dmul_rn( threshold, mx );
  int cnt = 0;
  for ( int r = lane; r < n; r += 32 ) {
    const double x = re[r], y = im ? im[r] : 0.0;
    const double nrm = __dadd_rn( __dmul_rn( x, x ), __dmul_rn( y, y ) );
    cnt += ( nrm > cut ) ? 1 : 0;
  }
  cnt = pb::warp_isum( cnt );
  if ( lane == 0 ) {
    max2[bv.w_off[k] + c] = mx;
    atomicAdd( kept + k, (unsigned long long)cnt );
  }
}

// Per-function, per-device launch attributes are set ONCE (under the context mutex) to the largest dynamic shared-memory
// size the device offers and never lowered: batches of one device are solved from several host threads
// (paladin-gpu --inflight), and a set-then-launch pair with a batch-dependent size could be undercut by another thread
// between the two calls.
template <class K>
cudaError_t raise_smem_limit( K kernel, int optin ) {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes( &fa, kernel );
  if ( e != cudaSuccess ) return e;
  return cudaFuncSetAttribute( kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes );
}

int set_kernel_attributes( DeviceCtx& c ) {
  if ( c.attrs_set ) return PALADIN_GPU_OK;
  const int o = c.max_smem_optin;
  PB_CUDA( raise_smem_limit( geev_warp_kernel, o ) );
  PB_CUDA( raise_smem_limit( small_hess_kernel, o ) );
  PB_CUDA( raise_smem_limit( small_qr_pair_kernel, o ) );
  PB_CUDA( raise_smem_limit( geev_smallcta_kernel, o ) );
  PB_CUDA( raise_smem_limit( mid_prep_kernel<0>, o ) );
  PB_CUDA( raise_smem_limit( mid_prep_kernel<16>, o ) );
  PB_CUDA( raise_smem_limit( mid_prep_kernel<32>, o ) );
  PB_CUDA( raise_smem_limit( mid_prep_blocked_kernel<16>, o ) );
  PB_CUDA( raise_smem_limit( mid_prep_blocked_kernel<32>, o ) );
  PB_CUDA( raise_smem_limit( big_prep_kernel, o ) );
  PB_CUDA( raise_smem_limit( big_prep_blocked_kernel, o ) );
  PB_CUDA( cudaFuncSetAttribute( big_prep_blocked_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1 ) );
  PB_CUDA( raise_smem_limit( mid_qr_kernel<false>, o ) );
  PB_CUDA( raise_smem_limit( mid_qr_kernel<true>, o ) );
  PB_CUDA( cudaFuncSetAttribute( mid_qr_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1 ) );
  PB_CUDA( raise_smem_limit( mid_vec_kernel, o ) );
  PB_CUDA( raise_smem_limit( big_vec_kernel<0>, o ) );
  PB_CUDA( raise_smem_limit( big_vec_kernel<1>, o ) );
  PB_CUDA( raise_smem_limit( big_vec_kernel<2>, o ) );
  PB_CUDA( raise_smem_limit( big_vec_kernel<3>, o ) );
  c.attrs_set = true;
  return PALADIN_GPU_OK;
}

struct StageTimer {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  bool used = false;
  int launches = 0;
};

}  // namespace

// ---- the opaque batch ------------------------------------------------------------------------------
struct paladin_gpu_batch {
  int device = 0;
  int count = 0;
  char jobvl = 'N', jobvr = 'N';
  std::vector<int> n;
  std::vector<int64_t> nnz_off, a_off, w_off;
  int64_t total_nnz = 0, total_nn = 0, total_n = 0;
  // device
  int* d_n = nullptr;
  int64_t *d_nnz_off = nullptr, *d_a_off = nullptr, *d_w_off = nullptr;
  int *d_coo_i = nullptr, *d_coo_j = nullptr;
  double* d_coo_v = nullptr;
  double *d_A = nullptr, *d_V = nullptr, *d_VL = nullptr, *d_wr = nullptr, *d_wi = nullptr, *d_dwork = nullptr;
  int *d_info = nullptr, *d_iwork = nullptr, *d_flags = nullptr, *d_ids = nullptr, *d_ids_f = nullptr;
  pb::ScatterChunk* d_chunks = nullptr;
  int nchunks = 0;
  // host-side plan
  std::vector<int> flags;                    // scatter flags per matrix
  std::vector<int> ids;                      // all ids grouped by class
  struct Group { int first, count, nmax, kind; };  // kind 0 = warp bucket, 1 = CTA per matrix
  long long* d_stats = nullptr;              // phase profile of the CTA-per-matrix kernels (cycles)
  MidMeta* d_meta = nullptr;
  int* d_counter = nullptr;
  double* d_xwork = nullptr;
  size_t xslot = 0;
  int xslots = 0, xn_max = 0;
  int64_t xhalf = 0;
  // blocked Hessenberg reduction (64 < n <= 1024): scratch, its offsets, tensor maps (A, Schur vectors) per matrix
  double* d_hwork = nullptr;
  int64_t* d_h_off = nullptr;
  CUtensorMap* d_tmaps = nullptr;
  int* d_tmap_ok = nullptr;
  // text ends (pb_textio.cuh): raw MatrixMarket text of the batch and the parser's tables; formatter scratch
  char* d_text = nullptr;
  size_t text_cap = 0;
  size_t text_resident = 0;                  // bytes placed with paladin_gpu_batch_text_begin / _put (upload_text with text == NULL parses them)
  void* d_text_tables = nullptr;             // chunks, ranges, counters (one allocation)
  size_t text_tables_cap = 0;
  char *d_fmt_slots = nullptr, *d_fmt_out = nullptr;
  unsigned char* d_fmt_lens = nullptr;
  unsigned long long* d_fmt_bsum = nullptr;
  long long* d_fmt_eoff = nullptr;           // e_off (count + 1), body_off (count + 1)
  int* d_fmt_flags = nullptr;
  long long fmt_cap_entries = 0;
  size_t fmt_text_bytes = 0;                 // bytes of formatted text kept in d_fmt_out (paladin_gpu_batch_fetch_text)
  double* d_max2 = nullptr;                  // scratch of paladin_gpu_batch_vector_filter (allocated with the batch: no
  unsigned long long* d_kept = nullptr;      // cudaMalloc / cudaFree -- device-wide synchronisations -- per call)
  int4* d_vec_items = nullptr;               // work items of the three-launch eigenvector stage
  int n_vec_items = 0;
  // cooperative stage 1 of large matrices
  unsigned* d_coop_bars = nullptr;
  double *d_coop_ypart = nullptr, *d_coop_yz = nullptr;
  int coop_S = 0, coop_G = 0, coop_nmax = 0;
  std::vector<Group> groups;
  StageTimer st[PALADIN_GPU_NUM_STAGES];
  bool scattered = false;
  cudaStream_t stream = nullptr;             // every batch owns its stream: batches of one device overlap
  cudaStream_t side[3] = { nullptr, nullptr, nullptr };  // side streams: the QR launches of the size classes run side by side
  cudaEvent_t ev_fork = nullptr, ev_join[3] = { nullptr, nullptr, nullptr };
  cudaEvent_t t0 = nullptr, t1 = nullptr;    // stopwatch (paladin_gpu_batch_timer_*)
};

namespace {

// stream the calling thread's allocations are ordered on (the batch's own stream, set by every entry that allocates)
thread_local cudaStream_t g_alloc_stream = nullptr;

template <class T>
int dev_alloc( T** p, size_t count ) {
  *p = nullptr;
  if ( count == 0 ) count = 1;
  PB_CUDA( cudaMallocAsync( reinterpret_cast<void**>( p ), count * sizeof( T ), g_alloc_stream ) );
  return PALADIN_GPU_OK;
}
inline void dev_free( void* p, cudaStream_t st ) {
  if ( p ) cudaFreeAsync( p, st );
}


// cuTensorMapEncodeTiled through the runtime's driver-entry-point query (nothing links libcuda). One map per dense
// matrix: 2-D, column-major n x n doubles, box = kBhBoxRows rows x 1 column, no swizzle, zero fill out of bounds.
typedef CUresult ( *EncodeTiledFn )( CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill );
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if ( cudaGetDriverEntryPoint( "cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q ) != cudaSuccess || q != cudaDriverEntryPointSuccess ) {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>( p );
  }();
  return fn;
}
bool encode_matrix_map( CUtensorMap* map, double* base, int n ) {
  EncodeTiledFn enc = tensor_map_encoder();
  if ( !enc || ( n & 1 ) || ( reinterpret_cast<uintptr_t>( base ) & 15 ) ) return false;
  const cuuint64_t gdim[2] = { (cuuint64_t)n, (cuuint64_t)n };
  const cuuint64_t gstr[1] = { (cuuint64_t)n * sizeof( double ) };
  const cuuint32_t box[2] = { (cuuint32_t)pb::kBhBoxRows, 1u };
  const cuuint32_t estr[2] = { 1u, 1u };
  return enc( map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE ) == CUDA_SUCCESS;
}

void free_batch( paladin_gpu_batch* b ) {
  if ( !b ) return;
  cudaSetDevice( b->device );
  dev_free( b->d_n, b->stream ); dev_free( b->d_nnz_off, b->stream ); dev_free( b->d_a_off, b->stream ); dev_free( b->d_w_off, b->stream );
  dev_free( b->d_coo_i, b->stream ); dev_free( b->d_coo_j, b->stream ); dev_free( b->d_coo_v, b->stream );
  dev_free( b->d_A, b->stream ); dev_free( b->d_V, b->stream ); dev_free( b->d_VL, b->stream ); dev_free( b->d_wr, b->stream ); dev_free( b->d_wi, b->stream ); dev_free( b->d_dwork, b->stream );
  dev_free( b->d_stats, b->stream ); dev_free( b->d_meta, b->stream ); dev_free( b->d_counter, b->stream ); dev_free( b->d_xwork, b->stream );
  dev_free( b->d_coop_bars, b->stream ); dev_free( b->d_coop_ypart, b->stream ); dev_free( b->d_coop_yz, b->stream ); dev_free( b->d_vec_items, b->stream );
  dev_free( b->d_info, b->stream ); dev_free( b->d_iwork, b->stream ); dev_free( b->d_flags, b->stream ); dev_free( b->d_text, b->stream ); dev_free( b->d_text_tables, b->stream ); dev_free( b->d_fmt_slots, b->stream ); dev_free( b->d_fmt_out, b->stream ); dev_free( b->d_fmt_lens, b->stream );
  dev_free( b->d_fmt_bsum, b->stream ); dev_free( b->d_fmt_eoff, b->stream ); dev_free( b->d_fmt_flags, b->stream );
  dev_free( b->d_hwork, b->stream ); dev_free( b->d_h_off, b->stream ); dev_free( b->d_tmaps, b->stream ); dev_free( b->d_tmap_ok, b->stream );
  dev_free( b->d_ids, b->stream ); dev_free( b->d_ids_f, b->stream ); dev_free( b->d_max2, b->stream ); dev_free( b->d_kept, b->stream ); dev_free( b->d_chunks, b->stream );
  for ( auto& s : b->st ) {
    if ( s.e0 ) cudaEventDestroy( s.e0 );
    if ( s.e1 ) cudaEventDestroy( s.e1 );
  }
  if ( b->t0 ) cudaEventDestroy( b->t0 );
  if ( b->t1 ) cudaEventDestroy( b->t1 );
  if ( b->stream ) cudaStreamDestroy( b->stream );
  for ( int q = 0; q < 3; ++q ) {
    if ( b->side[q] ) cudaStreamDestroy( b->side[q] );
    if ( b->ev_join[q] ) cudaEventDestroy( b->ev_join[q] );
  }
  if ( b->ev_fork ) cudaEventDestroy( b->ev_fork );
  delete b;
}

#define PB_TRY( expr )                    \
  do {                                    \
    int rc__ = ( expr );                  \
    if ( rc__ != PALADIN_GPU_OK ) return rc__; \
  } while ( 0 )

void stage_begin( paladin_gpu_batch* b, int s, cudaStream_t st ) {
  if ( !b->st[s].used ) {
    cudaEventRecord( b->st[s].e0, st );
    b->st[s].used = true;
  }
}
void stage_end( paladin_gpu_batch* b, int s, cudaStream_t st, int launches ) {
  cudaEventRecord( b->st[s].e1, st );
  b->st[s].launches += launches;
}
void stage_reset( paladin_gpu_batch* b, int s ) {
  b->st[s].used = false;
  b->st[s].launches = 0;
}

pb::ScatterArgs scatter_args( paladin_gpu_batch* b ) {
  pb::ScatterArgs a;
  a.chunks = b->d_chunks;
  a.n = b->d_n;
  a.nnz_off = b->d_nnz_off;
  a.a_off = b->d_a_off;
  a.coo_i = b->d_coo_i;
  a.coo_j = b->d_coo_j;
  a.coo_v = b->d_coo_v;
  a.dense = b->d_A;
  a.flags = b->d_flags;
  return a;
}

BatchView batch_view( paladin_gpu_batch* b ) {
  BatchView v;
  v.n = b->d_n;
  v.a_off = b->d_a_off;
  v.w_off = b->d_w_off;
  v.A = b->d_A;
  v.V = b->d_V;
  v.wr = b->d_wr;
  v.wi = b->d_wi;
  v.info = b->d_info;
  v.dwork = b->d_dwork;
  v.iwork = b->d_iwork;
  v.wantvr = ( b->jobvr == 'V' ) ? 1 : 0;
  v.wantvl = ( b->jobvl == 'V' ) ? 1 : 0;
  v.VL = b->d_VL;
  return v;
}

size_t warp_smem_bytes( int nmax, int nvec ) {
  const int lds = nmax | 1;
  return sizeof( double ) * ( (size_t)lds * nmax * ( 1 + nvec ) + 5 * (size_t)nmax ) + sizeof( int ) * ( 2 * (size_t)nmax + 2 );
}

}  // namespace

// =====================================================================================================
extern "C" {

int paladin_gpu_device_count( void ) { return usable_devices(); }

int paladin_gpu_init( int device ) { return ensure_ctx( device ); }

int paladin_gpu_shutdown( int device ) {
  if ( device < 0 || device >= kMaxDevices ) return fail( PALADIN_GPU_ENODEVICE, "device index out of range" );
  DeviceCtx& c = g_ctx[device];
  std::lock_guard<std::mutex> lk( c.mu );
  if ( c.ready ) {
    cudaSetDevice( device );
    cudaStreamSynchronize( c.stream );
    cudaStreamDestroy( c.stream );
    cudaEventDestroy( c.t0 );
    cudaEventDestroy( c.t1 );
    c.stream = nullptr;
    c.ready = false;
  }
  return PALADIN_GPU_OK;
}

const char* paladin_gpu_last_error( void ) { return g_err.c_str(); }

const char* paladin_gpu_version( void ) { return "paladin_gpu 0.1 sm_100a"; }

int paladin_gpu_host_alloc( void** ptr, size_t bytes ) {
  if ( !ptr ) return fail( PALADIN_GPU_EINVAL, "null pointer" );
  *ptr = nullptr;
  if ( usable_devices() == 0 ) return fail( PALADIN_GPU_ENODEVICE, "no CUDA device (there is no CPU fallback)" );
  PB_CUDA( cudaHostAlloc( ptr, bytes ? bytes : 1, cudaHostAllocPortable ) );
  return PALADIN_GPU_OK;
}

int paladin_gpu_host_free( void* ptr ) {
  if ( ptr ) PB_CUDA( cudaFreeHost( ptr ) );
  return PALADIN_GPU_OK;
}

int paladin_gpu_timer_start( int device ) {
  PB_TRY( ensure_ctx( device ) );
  DeviceCtx& c = g_ctx[device];
  std::lock_guard<std::mutex> lk( c.mu );
  PB_CUDA( cudaSetDevice( device ) );
  PB_CUDA( cudaEventRecord( c.t0, c.stream ) );
  return PALADIN_GPU_OK;
}

int paladin_gpu_timer_stop( int device, float* ms ) {
  if ( !ms ) return fail( PALADIN_GPU_EINVAL, "null pointer" );
  PB_TRY( ensure_ctx( device ) );
  DeviceCtx& c = g_ctx[device];
  std::lock_guard<std::mutex> lk( c.mu );
  PB_CUDA( cudaSetDevice( device ) );
  PB_CUDA( cudaEventRecord( c.t1, c.stream ) );
  PB_CUDA( cudaEventSynchronize( c.t1 ) );
  PB_CUDA( cudaEventElapsedTime( ms, c.t0, c.t1 ) );
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_create( int device, int count, const int* n, const int64_t* nnz_offsets, char jobvl, char jobvr,
                              paladin_gpu_batch** out ) {
  if ( !out ) return fail( PALADIN_GPU_EINVAL, "out is null" );
  *out = nullptr;
  if ( count < 0 || ( count > 0 && ( !n || !nnz_offsets ) ) ) return fail( PALADIN_GPU_EINVAL, "bad count / null sizes" );
  if ( ( jobvl != 'N' && jobvl != 'V' ) || ( jobvr != 'N' && jobvr != 'V' ) )
    return fail( PALADIN_GPU_EINVAL, "jobvl/jobvr must be 'N' or 'V'" );
  PB_TRY( ensure_ctx( device ) );
  DeviceCtx& c = g_ctx[device];
  std::lock_guard<std::mutex> lk( c.mu );
  PB_CUDA( cudaSetDevice( device ) );
  PB_TRY( set_kernel_attributes( c ) );

  paladin_gpu_batch* b = new paladin_gpu_batch;
  if ( cudaStreamCreateWithFlags( &b->stream, cudaStreamNonBlocking ) != cudaSuccess ) {
    delete b;
    return fail( PALADIN_GPU_ECUDA, "cudaStreamCreate failed" );
  }
  g_alloc_stream = b->stream;
  b->device = device;
  b->count = count;
  b->jobvl = jobvl;
  b->jobvr = jobvr;
  b->n.assign( n, n + count );
  b->nnz_off.assign( count + 1, 0 );
  b->a_off.assign( count + 1, 0 );
  b->w_off.assign( count + 1, 0 );
  for ( int k = 0; k < count; ++k ) {
    if ( n[k] < 0 || nnz_offsets[k + 1] < nnz_offsets[k] || nnz_offsets[k + 1] - nnz_offsets[k] > 0x7fffffffLL ) {
      cudaStreamDestroy( b->stream );
      delete b;
      return fail( PALADIN_GPU_EINVAL, "negative order or decreasing nnz_offsets at matrix " + std::to_string( k ) );
    }
    b->nnz_off[k] = nnz_offsets[k] - nnz_offsets[0];
    // every dense matrix starts on a 16-byte boundary so that paired stores stay aligned
    // ... and consecutive matrices are staggered by 33 cache lines: with power-of-two sizes (2 MiB at n = 512)
    // every CTA would otherwise walk the same memory-channel pattern in lockstep
    int64_t nn = (int64_t)n[k] * n[k];
    b->a_off[k + 1] = b->a_off[k] + ( ( nn + 1 ) & ~1LL ) + ( nn >= 4096 ? 528 : 0 );
    b->w_off[k + 1] = b->w_off[k] + n[k];
  }
  if ( count > 0 ) b->nnz_off[count] = nnz_offsets[count] - nnz_offsets[0];
  b->total_nnz = b->nnz_off[count];
  b->total_nn = b->a_off[count];
  b->total_n = b->w_off[count];

  // scatter chunk table
  std::vector<pb::ScatterChunk> chunks;
  for ( int k = 0; k < count; ++k ) {
    const int nnz = (int)( b->nnz_off[k + 1] - b->nnz_off[k] );
    const int nch = std::max( 1, ( nnz + pb::kScatterChunk - 1 ) / pb::kScatterChunk );
    for ( int c2 = 0; c2 < nch; ++c2 ) {
      pb::ScatterChunk ch;
      ch.matrix = k;
      ch.first = c2 * pb::kScatterChunk;
      ch.count = std::max( 0, std::min( pb::kScatterChunk, nnz - ch.first ) );
      ch.nchunks = nch;
      chunks.push_back( ch );
    }
  }
  b->nchunks = (int)chunks.size();

  // solve plan: warp buckets for n <= kSmallMax (when the bucket fits in shared memory), generic otherwise
  {
    std::vector<std::vector<int>> bucket( kSmallMax / kSmallBucket + 1 );
    std::vector<int> generic;
    for ( int k = 0; k < count; ++k ) {
      if ( n[k] == 0 ) continue;
      if ( n[k] <= kSmallMax )
        bucket[( n[k] + kSmallBucket - 1 ) / kSmallBucket].push_back( k );
      else
        generic.push_back( k );
    }
    std::stable_sort( generic.begin(), generic.end(), [&]( int x, int y ) { return n[x] > n[y]; } );
    for ( size_t q = 0; q < bucket.size(); ++q ) {
      if ( bucket[q].empty() ) continue;
      paladin_gpu_batch::Group g{ (int)b->ids.size(), (int)bucket[q].size(), (int)q * kSmallBucket, 0 };
      b->ids.insert( b->ids.end(), bucket[q].begin(), bucket[q].end() );
      b->groups.push_back( g );
    }
    if ( !generic.empty() ) {
      paladin_gpu_batch::Group g{ (int)b->ids.size(), (int)generic.size(), n[generic[0]], 1 };
      b->ids.insert( b->ids.end(), generic.begin(), generic.end() );
      b->groups.push_back( g );
    }
  }

  int rc = PALADIN_GPU_OK;
  auto A = [&]( int r ) { if ( rc == PALADIN_GPU_OK ) rc = r; };
  A( dev_alloc( &b->d_n, count ) );
  A( dev_alloc( &b->d_nnz_off, count + 1 ) );
  A( dev_alloc( &b->d_a_off, count + 1 ) );
  A( dev_alloc( &b->d_w_off, count + 1 ) );
  A( dev_alloc( &b->d_coo_i, b->total_nnz + 2 ) );
  A( dev_alloc( &b->d_coo_j, b->total_nnz + 2 ) );
  A( dev_alloc( &b->d_coo_v, b->total_nnz + 2 ) );
  A( dev_alloc( &b->d_A, b->total_nn ) );
  if ( jobvr == 'V' ) A( dev_alloc( &b->d_V, b->total_nn ) );
  if ( jobvl == 'V' ) A( dev_alloc( &b->d_VL, b->total_nn ) );
  A( dev_alloc( &b->d_wr, b->total_n ) );
  A( dev_alloc( &b->d_wi, b->total_n ) );
  A( dev_alloc( &b->d_dwork, 5 * b->total_n ) );
  A( dev_alloc( &b->d_iwork, 2 * b->total_n + 2 * count ) );
  A( dev_alloc( &b->d_info, count ) );
  A( dev_alloc( &b->d_flags, count ) );
  A( dev_alloc( &b->d_ids, b->ids.size() ) );
  A( dev_alloc( &b->d_ids_f, b->ids.size() ) );
  A( dev_alloc( &b->d_chunks, chunks.size() ) );
  if ( jobvr == 'V' || jobvl == 'V' ) {
    A( dev_alloc( &b->d_max2, b->total_n ) );
    A( dev_alloc( &b->d_kept, count ) );
  }
  A( dev_alloc( &b->d_stats, kNumTicks ) );
  A( dev_alloc( &b->d_meta, count ) );
  A( dev_alloc( &b->d_counter, 1 ) );
  {
    // blocked Hessenberg reduction: per-matrix scratch (Y, V T, T of every panel) and TMA tensor maps of the dense matrix
    // and of the Schur-vector array (box = 256 rows x 1 column of doubles; needs an even leading dimension)
    std::vector<int64_t> h_off( count + 1, 0 );
    for ( int k = 0; k < count; ++k )
      h_off[k + 1] = h_off[k] + ( ( n[k] > kSmallMax && n[k] <= kMidMaxBigN ) ? (int64_t)( ( ( n[k] <= kMidMaxFusedN ? pb::bh_scratch_doubles( n[k] ) : pb::bhc_scratch_doubles( n[k], 16 ) ) + 1 ) & ~(size_t)1 ) : 0 );
    if ( h_off[count] > 0 ) {
      A( dev_alloc( &b->d_hwork, (size_t)h_off[count] ) );
      A( dev_alloc( &b->d_h_off, count + 1 ) );
      A( dev_alloc( &b->d_tmaps, 2 * (size_t)count ) );
      A( dev_alloc( &b->d_tmap_ok, count ) );
      if ( rc == PALADIN_GPU_OK ) {
        std::vector<CUtensorMap> maps( 2 * (size_t)count );
        std::vector<int> ok( count, 0 );
        std::memset( maps.data(), 0, sizeof( CUtensorMap ) * maps.size() );
        double* qbase = jobvr == 'V' ? b->d_V : ( jobvl == 'V' ? b->d_VL : nullptr );
        for ( int k = 0; k < count; ++k ) {
          if ( !( n[k] > kSmallMax && n[k] <= kMidMaxBigN ) || ( n[k] & 1 ) ) continue;
          bool good = encode_matrix_map( &maps[2 * (size_t)k], b->d_A + b->a_off[k], n[k] );
          if ( good && qbase ) good = encode_matrix_map( &maps[2 * (size_t)k + 1], qbase + b->a_off[k], n[k] );
          ok[k] = good ? 1 : 0;
        }
        if ( cudaStreamSynchronize( b->stream ) != cudaSuccess || cudaMemcpy( b->d_h_off, h_off.data(), sizeof( int64_t ) * ( count + 1 ), cudaMemcpyHostToDevice ) != cudaSuccess ||
             cudaMemcpy( b->d_tmaps, maps.data(), sizeof( CUtensorMap ) * maps.size(), cudaMemcpyHostToDevice ) != cudaSuccess ||
             cudaMemcpy( b->d_tmap_ok, ok.data(), sizeof( int ) * count, cudaMemcpyHostToDevice ) != cudaSuccess )
          rc = fail( PALADIN_GPU_ECUDA, "cudaMemcpy failed" );
      }
    }
  }
  {
    // work slots of the eigenvector stage: two resident CTAs per SM at most
    int nmid = 0, nmax = 0;
    for ( const auto& g : b->groups )
      if ( g.kind == 1 ) {
        nmid += g.count;
        nmax = std::max( nmax, g.nmax );
      }
    if ( ( jobvr == 'V' || jobvl == 'V' ) && nmid > 0 ) {
      // one slot holds the triangular factor(s) of the largest matrix (two when both sides are wanted); at most two
      // resident CTAs per SM, at most 24 GB in total (fewer slots for very large matrices)
      b->xn_max = nmax;
      if ( bigvec_min_n() <= kSmallMax ) {
        // every matrix of the group takes the three-launch path: X is laid out like the dense matrices themselves
        b->xhalf = b->total_nn + 16;
        b->xslots = nmid;
        // work items: parts per matrix proportional to n^3, about four items per SM in total, at most 64 per matrix
        for ( const auto& g : b->groups ) {
          if ( g.kind != 1 ) continue;
          double total = 0.0;
          for ( int q = 0; q < g.count; ++q ) total += std::pow( (double)n[b->ids[g.first + q]], 3.0 );
          const double unit = std::max( 1.0, total / ( 4.0 * c.sm_count ) );
          std::vector<int4> items;
          for ( int q = 0; q < g.count; ++q ) {
            const double w = std::pow( (double)n[b->ids[g.first + q]], 3.0 );
            const int parts = std::max( 1, std::min( 64, (int)std::ceil( w / unit - 1e-9 ) ) );
            for ( int p2 = 0; p2 < parts; ++p2 ) items.push_back( make_int4( q, p2, parts, 0 ) );
          }
          b->n_vec_items = (int)items.size();
          A( dev_alloc( &b->d_vec_items, items.size() ) );
          if ( rc == PALADIN_GPU_OK && ( cudaStreamSynchronize( b->stream ) != cudaSuccess || cudaMemcpy( b->d_vec_items, items.data(), sizeof( int4 ) * items.size(), cudaMemcpyHostToDevice ) != cudaSuccess ) )
            rc = fail( PALADIN_GPU_ECUDA, "cudaMemcpy failed" );
        }
        A( dev_alloc( &b->d_xwork, (size_t)b->xhalf * ( jobvr == 'V' && jobvl == 'V' ? 2 : 1 ) ) );
      } else {
        b->xslot = ( (size_t)nmax * nmax + 16 ) * ( jobvr == 'V' && jobvl == 'V' ? 2 : 1 ) + 528;
        const size_t budget = (size_t)24 << 30;
        const int by_mem = (int)std::max<size_t>( 1, budget / ( b->xslot * sizeof( double ) ) );
        b->xslots = std::max( 1, std::min( std::min( nmid, 2 * c.sm_count ), by_mem ) );
        A( dev_alloc( &b->d_xwork, b->xslot * b->xslots ) );
      }
    }
  }
  {
    // scratch of the cooperative Hessenberg stage: the largest matrix that takes it fixes the group size
    int nbigmax = 0;
    for ( int k = 0; k < count; ++k )
      if ( n[k] > coop_min_n() && n[k] <= kMidMaxBigN ) nbigmax = std::max( nbigmax, n[k] );
    static const bool no_coop = std::getenv( "PALADIN_GPU_NO_COOP" ) != nullptr;
    if ( nbigmax > 0 && !no_coop ) {
      // groups per launch are chosen per size class at solve time (1, 2 or 4); the scratch covers every choice
      b->coop_G = kCoopMaxGroups;
      b->coop_S = c.sm_count;
      b->coop_nmax = nbigmax;
      A( dev_alloc( &b->d_coop_bars, 2 * (size_t)b->coop_G + 2 ) );
      A( dev_alloc( &b->d_coop_ypart, (size_t)c.sm_count * nbigmax ) );
      A( dev_alloc( &b->d_coop_yz, (size_t)b->coop_G * 2 * nbigmax ) );
      if ( rc == PALADIN_GPU_OK && ( cudaStreamSynchronize( b->stream ) != cudaSuccess || cudaMemset( b->d_coop_bars, 0, sizeof( unsigned ) * ( 2 * b->coop_G + 2 ) ) != cudaSuccess ) )
        rc = fail( PALADIN_GPU_ECUDA, "cudaMemset failed" );
    }
  }
  if ( rc != PALADIN_GPU_OK ) {
    free_batch( b );
    return rc;
  }
  for ( auto& s : b->st ) {
    cudaEventCreate( &s.e0 );
    cudaEventCreate( &s.e1 );
  }
  cudaEventCreate( &b->t0 );
  cudaEventCreate( &b->t1 );
  cudaEventCreateWithFlags( &b->ev_fork, cudaEventDisableTiming );
  bool side_ok = true;
  for ( int q = 0; q < 3; ++q ) {
    cudaEventCreateWithFlags( &b->ev_join[q], cudaEventDisableTiming );
    side_ok = side_ok && cudaStreamCreateWithFlags( &b->side[q], cudaStreamNonBlocking ) == cudaSuccess;
  }
  if ( !side_ok ) {
    free_batch( b );
    return fail( PALADIN_GPU_ECUDA, "cudaStreamCreate failed" );
  }
  cudaStream_t st = b->stream;
  auto H2D = [&]( void* d, const void* h, size_t bytes ) {
    if ( rc == PALADIN_GPU_OK && bytes ) {
      cudaError_t e = cudaMemcpyAsync( d, h, bytes, cudaMemcpyHostToDevice, st );
      if ( e != cudaSuccess ) rc = fail( PALADIN_GPU_ECUDA, cudaGetErrorString( e ) );
    }
  };
  H2D( b->d_n, b->n.data(), sizeof( int ) * count );
  H2D( b->d_nnz_off, b->nnz_off.data(), sizeof( int64_t ) * ( count + 1 ) );
  H2D( b->d_a_off, b->a_off.data(), sizeof( int64_t ) * ( count + 1 ) );
  H2D( b->d_w_off, b->w_off.data(), sizeof( int64_t ) * ( count + 1 ) );
  H2D( b->d_ids, b->ids.data(), sizeof( int ) * b->ids.size() );
  H2D( b->d_chunks, chunks.data(), sizeof( pb::ScatterChunk ) * chunks.size() );
  if ( rc == PALADIN_GPU_OK ) {
    cudaError_t e = cudaMemsetAsync( b->d_info, 0, sizeof( int ) * std::max( count, 1 ), st );
    // eigenvalues of matrices that are never solved (bad COO index, non-finite entry) read as NaN, not as whatever the
    // allocation held before
    if ( e == cudaSuccess ) e = cudaMemsetAsync( b->d_wr, 0xff, sizeof( double ) * std::max<int64_t>( b->total_n, 1 ), st );
    if ( e == cudaSuccess ) e = cudaMemsetAsync( b->d_wi, 0xff, sizeof( double ) * std::max<int64_t>( b->total_n, 1 ), st );
    if ( e == cudaSuccess ) e = cudaStreamSynchronize( st );
    if ( e != cudaSuccess ) rc = fail( PALADIN_GPU_ECUDA, cudaGetErrorString( e ) );
  }
  if ( rc != PALADIN_GPU_OK ) {
    free_batch( b );
    return rc;
  }
  b->flags.assign( count, 0 );
  *out = b;
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_upload( paladin_gpu_batch* b, const int* coo_i, const int* coo_j, const double* coo_v ) {
  if ( !b ) return fail( PALADIN_GPU_EINVAL, "null batch" );
  if ( b->total_nnz > 0 && ( !coo_i || !coo_j || !coo_v ) ) return fail( PALADIN_GPU_EINVAL, "null COO arrays" );
  DeviceCtx& c = g_ctx[b->device];
  (void)c;
  PB_CUDA( cudaSetDevice( b->device ) );
  if ( b->total_nnz > 0 ) {
    PB_CUDA( cudaMemcpyAsync( b->d_coo_i, coo_i, sizeof( int ) * b->total_nnz, cudaMemcpyHostToDevice, b->stream ) );
    PB_CUDA( cudaMemcpyAsync( b->d_coo_j, coo_j, sizeof( int ) * b->total_nnz, cudaMemcpyHostToDevice, b->stream ) );
    PB_CUDA( cudaMemcpyAsync( b->d_coo_v, coo_v, sizeof( double ) * b->total_nnz, cudaMemcpyHostToDevice, b->stream ) );
  }
  PB_CUDA( cudaStreamSynchronize( b->stream ) );
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_scatter( paladin_gpu_batch* b ) {
  if ( !b ) return fail( PALADIN_GPU_EINVAL, "null batch" );
  DeviceCtx& c = g_ctx[b->device];
  (void)c;
  PB_CUDA( cudaSetDevice( b->device ) );
  cudaStream_t st = b->stream;
  stage_reset( b, PALADIN_GPU_STAGE_SCATTER );
  if ( b->count == 0 ) {
    b->scattered = true;
    return PALADIN_GPU_OK;
  }
  PB_CUDA( cudaMemsetAsync( b->d_flags, 0, sizeof( int ) * b->count, st ) );
  PB_CUDA( cudaMemsetAsync( b->d_info, 0, sizeof( int ) * b->count, st ) );
  pb::ScatterArgs a = scatter_args( b );
  stage_begin( b, PALADIN_GPU_STAGE_SCATTER, st );
  pb::scatter_sorted_kernel<<<b->nchunks, pb::kScatterThreads, 0, st>>>( a );
  stage_end( b, PALADIN_GPU_STAGE_SCATTER, st, 1 );
  PB_CUDA( cudaGetLastError() );
  PB_CUDA( cudaMemcpyAsync( b->flags.data(), b->d_flags, sizeof( int ) * b->count, cudaMemcpyDeviceToHost, st ) );
  PB_CUDA( cudaStreamSynchronize( st ) );
  bool any_unsorted = false, any_bad = false;
  for ( int k = 0; k < b->count; ++k ) {
    if ( b->flags[k] & pb::kFlagBadIndex ) any_bad = true;
    else if ( b->flags[k] & pb::kFlagUnsorted ) any_unsorted = true;
  }
  if ( any_unsorted ) {
    // order-exact fallback for the flagged matrices (separate launches = global barriers between passes)
    pb::zero_flagged_kernel<<<b->nchunks, pb::kScatterThreads, 0, st>>>( a );
    pb::scatter_fallback_kernel<0><<<b->nchunks, pb::kScatterThreads, 0, st>>>( a );
    pb::scatter_fallback_kernel<1><<<b->nchunks, pb::kScatterThreads, 0, st>>>( a );
    pb::scatter_fallback_kernel<2><<<b->nchunks, pb::kScatterThreads, 0, st>>>( a );
    stage_end( b, PALADIN_GPU_STAGE_SCATTER, st, 4 );
    PB_CUDA( cudaGetLastError() );
    PB_CUDA( cudaStreamSynchronize( st ) );
  }
  if ( any_bad ) {
    std::vector<int> info( b->count, 0 );
    for ( int k = 0; k < b->count; ++k )
      if ( b->flags[k] & pb::kFlagBadIndex ) info[k] = PALADIN_GPU_INFO_BAD_INDEX;
    PB_CUDA( cudaMemcpyAsync( b->d_info, info.data(), sizeof( int ) * b->count, cudaMemcpyHostToDevice, st ) );
    PB_CUDA( cudaStreamSynchronize( st ) );
  }
  b->scattered = true;
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_upload_dense( paladin_gpu_batch* b, const double* dense ) {
  if ( !b ) return fail( PALADIN_GPU_EINVAL, "null batch" );
  if ( b->total_nn > 0 && !dense ) return fail( PALADIN_GPU_EINVAL, "null dense array" );
  DeviceCtx& c = g_ctx[b->device];
  (void)c;
  PB_CUDA( cudaSetDevice( b->device ) );
  int64_t src = 0;
  for ( int k = 0; k < b->count; ++k ) {
    const int64_t nn = (int64_t)b->n[k] * b->n[k];
    if ( nn )
      PB_CUDA( cudaMemcpyAsync( b->d_A + b->a_off[k], dense + src, sizeof( double ) * nn, cudaMemcpyHostToDevice, b->stream ) );
    src += nn;
  }
  PB_CUDA( cudaMemsetAsync( b->d_info, 0, sizeof( int ) * std::max( b->count, 1 ), b->stream ) );
  PB_CUDA( cudaStreamSynchronize( b->stream ) );
  std::fill( b->flags.begin(), b->flags.end(), 0 );
  b->scattered = true;
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_get_dense( paladin_gpu_batch* b, int k, double* out ) {
  if ( !b || k < 0 || k >= b->count || !out ) return fail( PALADIN_GPU_EINVAL, "bad arguments" );
  DeviceCtx& c = g_ctx[b->device];
  (void)c;
  PB_CUDA( cudaSetDevice( b->device ) );
  const int64_t nn = (int64_t)b->n[k] * b->n[k];
  if ( nn ) PB_CUDA( cudaMemcpyAsync( out, b->d_A + b->a_off[k], sizeof( double ) * nn, cudaMemcpyDeviceToHost, b->stream ) );
  PB_CUDA( cudaStreamSynchronize( b->stream ) );
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_solve( paladin_gpu_batch* b ) {
  if ( !b ) return fail( PALADIN_GPU_EINVAL, "null batch" );
  if ( !b->scattered ) return fail( PALADIN_GPU_EINVAL, "solve called before scatter / upload_dense" );
  g_alloc_stream = b->stream;
  DeviceCtx& c = g_ctx[b->device];
  (void)c;
  PB_CUDA( cudaSetDevice( b->device ) );
  cudaStream_t st = b->stream;
  for ( int s = 1; s < PALADIN_GPU_NUM_STAGES; ++s ) stage_reset( b, s );
  BatchView bv = batch_view( b );
  const bool wantvr = b->jobvr == 'V';
  const bool wantvl = b->jobvl == 'V';
  bool any_bad = false;
  for ( int k = 0; k < b->count; ++k ) any_bad = any_bad || ( b->flags[k] & pb::kFlagBadIndex );
  std::vector<int> gid;  // ids of the group that are solved (size-sorted like the group itself)
  for ( const auto& g : b->groups ) {
    const int* ids = b->d_ids + g.first;
    gid.assign( b->ids.begin() + g.first, b->ids.begin() + g.first + g.count );
    if ( any_bad ) {
      // rare: drop matrices whose scatter failed. The filtered list goes to a scratch copy of the id array (d_ids itself
      // is never modified), and every class boundary below is derived from it.
      gid.clear();
      for ( int q = 0; q < g.count; ++q )
        if ( !( b->flags[b->ids[g.first + q]] & pb::kFlagBadIndex ) ) gid.push_back( b->ids[g.first + q] );
      ids = b->d_ids_f + g.first;
      if ( !gid.empty() ) PB_CUDA( cudaMemcpyAsync( b->d_ids_f + g.first, gid.data(), sizeof( int ) * gid.size(), cudaMemcpyHostToDevice, st ) );
      PB_CUDA( cudaStreamSynchronize( st ) );
    }
    const int cnt = (int)gid.size();
    if ( cnt == 0 ) continue;
    if ( g.kind == 0 ) {
      const size_t smem = warp_smem_bytes( g.nmax, ( wantvr ? 1 : 0 ) + ( wantvl ? 1 : 0 ) );
      if ( smem <= (size_t)c.max_smem_optin ) {
        stage_begin( b, PALADIN_GPU_STAGE_SMALL, st );
        static const int small_warps = std::getenv( "PALADIN_GPU_SMALL_WARPS" ) ? std::max( 1, std::min( 4, std::atoi( std::getenv( "PALADIN_GPU_SMALL_WARPS" ) ) ) ) : 1;
        // eigenvalues only: reduction and iteration as two launches, the iteration with two matrices per shared array
        // (PALADIN_GPU_SMALL_SPLIT=0: the one-launch kernel)
        static const bool small_split = !( std::getenv( "PALADIN_GPU_SMALL_SPLIT" ) && std::atoi( std::getenv( "PALADIN_GPU_SMALL_SPLIT" ) ) == 0 );
        int nl = 1;
        if ( small_warps > 1 && g.nmax >= 32 ) geev_smallcta_kernel<<<cnt, 32 * small_warps, smem, st>>>( bv, ids, g.nmax, g.nmax | 1 );
        else if ( small_split && !wantvr && !wantvl && g.nmax >= 8 ) {
          const int lds = g.nmax | 1;
          small_hess_kernel<<<cnt, kSmallHessThreads, smem, st>>>( bv, ids, b->d_meta, g.nmax, lds );
          small_qr_pair_kernel<<<( cnt + 1 ) / 2, 64, sizeof( double ) * (size_t)lds * ( g.nmax + 3 ), st>>>( bv, ids, cnt, b->d_meta, g.nmax, lds );
          nl = 2;
        } else geev_warp_kernel<<<cnt, 32, smem, st>>>( bv, ids, g.nmax, g.nmax | 1 );
        stage_end( b, PALADIN_GPU_STAGE_SMALL, st, nl );
        PB_CUDA( cudaGetLastError() );
        continue;
      }
    }
    static const bool force_generic = std::getenv( "PALADIN_GPU_FORCE_GENERIC" ) != nullptr;
    if ( g.kind == 1 && !force_generic ) {
      MidArgs ma;
      ma.bv = bv;
      ma.ids = ids;
      ma.count = cnt;
      ma.meta = b->d_meta;
      ma.stats = b->d_stats;
      ma.counter = b->d_counter;
      ma.xwork = b->d_xwork;
      ma.xslot = b->xslot;
      ma.xn_max = b->xn_max;
      ma.xhalf = b->xhalf;
      static const int env_aed = std::getenv( "PALADIN_GPU_AED" ) ? std::atoi( std::getenv( "PALADIN_GPU_AED" ) ) : -1;
      static const int env_moves = std::getenv( "PALADIN_GPU_AED_MOVES" ) ? std::atoi( std::getenv( "PALADIN_GPU_AED_MOVES" ) ) : -1;
      ma.aed = std::min( env_aed, pb::kStripW - 1 );
      ma.aed_moves = env_moves;
      static const int env_nibble = std::getenv( "PALADIN_GPU_AED_NIBBLE" ) ? std::atoi( std::getenv( "PALADIN_GPU_AED_NIBBLE" ) ) : -1;
      ma.nibble = env_nibble;
      static const int env_itmax = std::getenv( "PALADIN_GPU_QR_ITMAX" ) ? std::max( 0, std::atoi( std::getenv( "PALADIN_GPU_QR_ITMAX" ) ) ) : 0;
      ma.itmax = env_itmax;
      static const int env_spec = std::getenv( "PALADIN_GPU_SPEC_SHIFTS" ) ? std::atoi( std::getenv( "PALADIN_GPU_SPEC_SHIFTS" ) ) : 0;
      ma.spec_shifts = env_spec;
      // PALADIN_GPU_HESS: "fused" = the unblocked single-sweep kernel (pb_hess.cuh), "notma" = blocked, tiles staged by plain
      // loads; default = blocked with TMA-staged tiles (odd orders always stage by plain loads)
      static const std::string env_hess = std::getenv( "PALADIN_GPU_HESS" ) ? std::getenv( "PALADIN_GPU_HESS" ) : "";
      const bool hess_blocked = env_hess != "fused" && b->d_hwork != nullptr;
      ma.hwork = b->d_hwork;
      ma.h_off = b->d_h_off;
      ma.tmaps = b->d_tmaps;
      ma.tmap_ok = b->d_tmap_ok;
      ma.hess_tma = env_hess == "notma" ? 0 : 1;
      static const bool env_formq_old = std::getenv( "PALADIN_GPU_FORMQ" ) != nullptr && std::string( std::getenv( "PALADIN_GPU_FORMQ" ) ) == "unblocked";
      ma.formq_unblocked = env_formq_old ? 1 : 0;
      static const int env_rg = std::getenv( "PALADIN_GPU_REFL_GLOBAL" ) ? std::atoi( std::getenv( "PALADIN_GPU_REFL_GLOBAL" ) ) : 1;  // 3072^2 x 2, clusters of 16: QR stage 1373 -> 1275 ms
      ma.refl_global = env_rg;
      PB_CUDA( cudaMemsetAsync( b->d_stats, 0, sizeof( long long ) * kNumTicks, st ) );
      PB_CUDA( cudaMemsetAsync( b->d_counter, 0, sizeof( int ), st ) );
      const size_t s2 = qr_smem_bytes(), s3 = vec_smem_bytes( g.nmax );
      stage_begin( b, PALADIN_GPU_STAGE_HESSENBERG, st );
      {
        // the group's ids are sorted by decreasing n: [0, nbig) need the unblocked path, [nbig, nmid) the 32-chunk
        // kernel, the rest the 16-chunk kernel
        int nbig = 0, nmid = 0;
        for ( int q = 0; q < cnt; ++q ) {
          const int nq = b->n[gid[q]];
          if ( nq > ( b->coop_S > 0 ? coop_min_n() : kMidMaxFusedN ) ) nbig = q + 1;
          if ( nq > 512 ) nmid = q + 1;
        }
        nbig = std::min( nbig, cnt );
        nmid = std::max( std::min( nmid, cnt ), nbig );
        int launches = 0;
        int nhuge = 0;  // [0, nhuge): beyond the streaming kernels (n > kMidMaxBigN)
        for ( int q = 0; q < nbig; ++q )
          if ( b->n[gid[q]] > kMidMaxBigN ) nhuge = q + 1;
        // n > 1024: blocked reduction, a cluster of S CTAs per matrix (S by how many large matrices there are: the GPU's SMs
        // divided among them, a power of two up to 16; PALADIN_GPU_BHESS_CLUSTER forces it, PALADIN_GPU_HESS=fused selects
        // the unblocked cooperative groups instead)
        bool big_done = false;
        // (measured on B200: the cluster-shared blocked reduction is correct for every order up to 4608 but its thin panel
        // steps and tile fragments are latency-bound -- 8 x 3072^2: 1091 ms against 486 ms, mixed batch of 512: 8.0 s against
        // 2.8 s for the unblocked cooperative groups -- so it only runs on request: PALADIN_GPU_HESS=blocked_big)
        if ( hess_blocked && env_hess == "blocked_big" && nbig > nhuge ) {
          static const int env_bc = std::getenv( "PALADIN_GPU_BHESS_CLUSTER" ) ? std::atoi( std::getenv( "PALADIN_GPU_BHESS_CLUSTER" ) ) : 0;
          const int nmat = nbig - nhuge;
          int S = 1;
          while ( S < 16 && 2 * S * nmat <= c.sm_count ) S *= 2;
          if ( env_bc > 0 ) S = std::min( 16, env_bc );
          MidArgs mc = ma;
          mc.ids = ids + nhuge;
          mc.count = nmat;
          const size_t sb = bcprep_smem_bytes( b->n[gid[nhuge]] );
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3( nmat * S );
          cfg.blockDim = dim3( pb::kBhThreads );
          cfg.dynamicSmemBytes = sb;
          cfg.stream = st;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeClusterDimension;
          attr[0].val.clusterDim.x = S;
          attr[0].val.clusterDim.y = 1;
          attr[0].val.clusterDim.z = 1;
          cfg.attrs = attr;
          cfg.numAttrs = S > 1 ? 1 : 0;
          PB_CUDA( cudaLaunchKernelEx( &cfg, big_prep_blocked_kernel, mc, S ) );
          ++launches;
          big_done = true;
        }
        if ( !big_done && b->coop_S > 0 && nbig > nhuge ) {
          // large matrices: groups of cooperating CTAs, one matrix per group at a time. The size-sorted list is cut into
          // classes: all SMs on one matrix above n = 2048, two groups down to 1400, four groups below.
          static const int env_groups = std::getenv( "PALADIN_GPU_COOP_GROUPS" ) ? std::max( 1, std::atoi( std::getenv( "PALADIN_GPU_COOP_GROUPS" ) ) ) : 0;
          int q0 = nhuge;
          while ( q0 < nbig ) {
            const int nfirst = b->n[gid[q0]];
            const int kind = nfirst > 2048 ? 0 : ( nfirst > 1400 ? 1 : 2 );
            int q1 = q0;
            while ( q1 < nbig ) {
              const int nq = b->n[gid[q1]];
              if ( ( nq > 2048 ? 0 : ( nq > 1400 ? 1 : 2 ) ) != kind ) break;
              ++q1;
            }
            // a single matrix takes all SMs (latency); a class with many matrices is better off with more, smaller
            // groups (throughput: the per-column barriers and reductions cost the same whatever the group size)
            const int ncls = q1 - q0;
            // (3072^2 x 8: 910 ms with 4 groups of 37; the sweep of a small group is bound by its SMs' L2 bandwidth, the
            // fixed per-column cost is amortised over more columns per CTA)
            const int gmin = kind == 0 ? 1 : ( kind == 1 ? 2 : 4 );
            const int gcap = std::max( 1, c.sm_count / 8 );
            int G = std::min( std::max( ncls, gmin ), gcap );
            if ( ncls > G ) {
              // several rounds: the same number of rounds with as few (hence as large) groups as possible, so that the last
              // round is not mostly idle (23 matrices: 12 groups of 12 CTAs instead of 18 + 5 on groups of 8)
              const int rounds = ( ncls + G - 1 ) / G;
              G = ( ncls + rounds - 1 ) / rounds;
            }
            if ( env_groups > 0 ) {
              G = std::min( env_groups, kCoopMaxGroups );
              q1 = nbig;
            }
            const int S = std::max( 1, c.sm_count / G );
            MidArgs mc = ma;
            mc.ids = ids + q0;
            mc.count = q1 - q0;
            CoopArgs ca{ b->d_coop_bars, b->d_coop_ypart, b->d_coop_yz, S, G, b->coop_nmax };
            const size_t sc = sizeof( double ) * pb::coop_hess_smem_doubles( nfirst, S );
            void* kargs[] = { &mc, &ca };
            PB_CUDA( cudaLaunchCooperativeKernel( (const void*)big_prep_kernel, dim3( S * G ), dim3( pb::kCoopThreads ), kargs, sc, st ) );
            ++launches;
            q0 = q1;
          }
        }
        const int n0 = ( b->coop_S > 0 || big_done ) ? nhuge : nbig;  // what is left for the one-CTA streaming / unblocked kernel
        if ( n0 > 0 ) {
          MidArgs m0 = ma;
          const size_t sbig = sizeof( double ) * pb::hess_big_smem_doubles( std::min( g.nmax, kMidMaxBigN ) );
          mid_prep_kernel<0><<<n0, pb::kHessThreads, sbig, st>>>( m0 );
          ++launches;
        }
        if ( nmid > nbig ) {
          MidArgs m1 = ma;
          m1.ids = ids + nbig;
          if ( hess_blocked ) mid_prep_blocked_kernel<32><<<nmid - nbig, pb::kBhThreads, bprep_smem_bytes( 1024 ), st>>>( m1 );
          else mid_prep_kernel<32><<<nmid - nbig, pb::kHessThreads, prep_smem_bytes( 1024 ), st>>>( m1 );
          ++launches;
        }
        if ( cnt > nmid ) {
          MidArgs m2 = ma;
          m2.ids = ids + nmid;
          if ( hess_blocked ) mid_prep_blocked_kernel<16><<<cnt - nmid, pb::kBhThreads, bprep_smem_bytes( 512 ), st>>>( m2 );
          else mid_prep_kernel<16><<<cnt - nmid, pb::kHessThreads, prep_smem_bytes( 512 ), st>>>( m2 );
          ++launches;
        }
        stage_end( b, PALADIN_GPU_STAGE_HESSENBERG, st, launches );
      }
      // debugging / tests: stop after stage 1 so that H (+ reflectors) and Q can be inspected (get_dense / download)
      if ( std::getenv( "PALADIN_GPU_DEBUG_STOP_AFTER_PREP" ) != nullptr ) {
        PB_CUDA( cudaGetLastError() );
        continue;
      }
      stage_begin( b, PALADIN_GPU_STAGE_QR, st );
      {
        // large matrices ([0, nqbig) of the size-sorted list) get a thread-block cluster each when they are few: the
        // leader's strip updates are shared by the cluster
        int nqbig = 0;
        for ( int q = 0; q < cnt; ++q )
          if ( b->n[gid[q]] > kMidMaxFusedN ) nqbig = q + 1;
        // cluster size per size class of the sorted list (n > 2048: 16 for up to four matrices, else 8; > 1400: 8; else 4
        // CTAs), halved together until the clusters fit one and a half times the 2 x SM-count resident CTAs (the tail of a
        // class may queue; with a tighter limit the largest class of a mixed batch loses its clusters of 8: 62 against 74
        // matrices/s on 256 mixed matrices); every class is its own launch on its own stream, the remaining matrices (one
        // CTA each) run next to them. The iteration is bound by its leader CTA:
        // measured on 3072^2 (QR stage, ms): 2 matrices 1015 / 1059 with clusters of 16 / 8, 4: 1019 / 1073, 8: 1158 / 1088,
        // 16: 2154 / 1396 (/ 1821 with 4), 32: 1483 with 8, 48: 2762 with 8, 64: 2745 with 4 -- a cluster of 16 only pays
        // while it has the SMs to itself
        const int slots = 3 * c.sm_count;
        struct QrClass { int first, count, csize; };
        QrClass cls[3] = { { 0, 0, 16 }, { 0, 0, 8 }, { 0, 0, 4 } };
        for ( int q = 0; q < nqbig; ++q ) {
          const int nq = b->n[gid[q]];
          QrClass& k = cls[nq > 2048 ? 0 : ( nq > 1400 ? 1 : 2 )];
          if ( k.count == 0 ) k.first = q;
          k.count += 1;
        }
        if ( cls[0].count > 4 ) cls[0].csize = 8;
        static const int env_cluster = std::getenv( "PALADIN_GPU_CLUSTER" ) ? std::atoi( std::getenv( "PALADIN_GPU_CLUSTER" ) ) : 0;
        for ( ;; ) {
          int demand = 0;
          for ( const QrClass& k : cls ) demand += k.count * k.csize;
          if ( demand <= slots ) break;
          bool any = false;
          for ( QrClass& k : cls )
            if ( k.csize > 1 ) {
              k.csize /= 2;
              any = true;
            }
          if ( !any ) break;
        }
        if ( env_cluster > 0 )
          for ( QrClass& k : cls ) k.csize = env_cluster;
        int launches = 0, nclustered = 0, nside = 0;
        PB_CUDA( cudaEventRecord( b->ev_fork, st ) );
        for ( const QrClass& k : cls ) {
          if ( k.count == 0 || k.csize <= 1 ) continue;
          if ( k.first != nclustered ) break;  // (classes are contiguous prefixes of the sorted list)
          cudaStream_t sq = st;
          if ( nside > 0 ) {
            sq = b->side[nside - 1];
            PB_CUDA( cudaStreamWaitEvent( sq, b->ev_fork, 0 ) );
          }
          MidArgs mq = ma;
          mq.ids = ids + k.first;
          mq.count = k.count;
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3( k.count * k.csize );
          cfg.blockDim = dim3( kMidThreads );
          cfg.dynamicSmemBytes = s2;
          cfg.stream = sq;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeClusterDimension;
          attr[0].val.clusterDim.x = k.csize;
          attr[0].val.clusterDim.y = 1;
          attr[0].val.clusterDim.z = 1;
          cfg.attrs = attr;
          cfg.numAttrs = 1;
          PB_CUDA( cudaLaunchKernelEx( &cfg, mid_qr_kernel<true>, mq ) );
          ++launches;
          ++nside;
          nclustered = k.first + k.count;
        }
        if ( cnt > nclustered ) {
          cudaStream_t sq = st;
          if ( nside > 0 ) {
            sq = b->side[nside - 1];
            PB_CUDA( cudaStreamWaitEvent( sq, b->ev_fork, 0 ) );
          }
          MidArgs mq = ma;
          mq.ids = ids + nclustered;
          mq.count = cnt - nclustered;
          mid_qr_kernel<false><<<cnt - nclustered, kMidThreads, s2, sq>>>( mq );
          ++launches;
          ++nside;
        }
        for ( int q = 0; q + 1 < nside; ++q ) {
          PB_CUDA( cudaEventRecord( b->ev_join[q], b->side[q] ) );
          PB_CUDA( cudaStreamWaitEvent( st, b->ev_join[q], 0 ) );
        }
        stage_end( b, PALADIN_GPU_STAGE_QR, st, launches );
      }
      if ( wantvr || wantvl ) {
        stage_begin( b, PALADIN_GPU_STAGE_VECTORS, st );
        int launches = 0;
        // large matrices first ([0, nvbig) of the size-sorted list): each one shared by P CTAs, `xslots` of them at a time
        int nvbig = 0;
        for ( int q = 0; q < cnt; ++q )
          if ( b->n[gid[q]] > bigvec_min_n() ) nvbig = q + 1;
        if ( nvbig > 0 ) {
          if ( b->d_vec_items != nullptr && nvbig == cnt && !any_bad ) {
            // every matrix of the group, parts proportional to n^3 (work items built at batch_create)
            MidArgs mv = ma;
            const int ni = b->n_vec_items;
            big_vec_kernel<0><<<ni, pb::kVecThreads, s3, st>>>( mv, b->d_vec_items );
            big_vec_kernel<1><<<ni, pb::kVecThreads, s3, st>>>( mv, b->d_vec_items );
            if ( wantvr && wantvl ) big_vec_kernel<2><<<ni, pb::kVecThreads, s3, st>>>( mv, b->d_vec_items );
            big_vec_kernel<3><<<ni, pb::kVecThreads, s3, st>>>( mv, b->d_vec_items );
            launches += ( wantvr && wantvl ) ? 4 : 3;
          } else {
            // uniform parts (tuning paths, filtered id lists): items built on the fly
            for ( int q0 = 0; q0 < nvbig; q0 += b->xslots ) {
              const int nq = std::min( b->xslots, nvbig - q0 );
              const int P = std::max( 1, std::min( 64, ( 4 * c.sm_count ) / nq ) );
              std::vector<int4> items;
              for ( int q = 0; q < nq; ++q )
                for ( int p2 = 0; p2 < P; ++p2 ) items.push_back( make_int4( q, p2, P, 0 ) );
              int4* d_items = nullptr;
              PB_TRY( dev_alloc( &d_items, items.size() ) );
              PB_CUDA( cudaMemcpyAsync( d_items, items.data(), sizeof( int4 ) * items.size(), cudaMemcpyHostToDevice, st ) );
              MidArgs mv = ma;
              mv.ids = ids + q0;
              mv.count = nq;
              big_vec_kernel<0><<<nq * P, pb::kVecThreads, s3, st>>>( mv, d_items );
              big_vec_kernel<1><<<nq * P, pb::kVecThreads, s3, st>>>( mv, d_items );
              if ( wantvr && wantvl ) big_vec_kernel<2><<<nq * P, pb::kVecThreads, s3, st>>>( mv, d_items );
              big_vec_kernel<3><<<nq * P, pb::kVecThreads, s3, st>>>( mv, d_items );
              launches += ( wantvr && wantvl ) ? 4 : 3;
              PB_CUDA( cudaStreamSynchronize( st ) );
              dev_free( d_items, st );
            }
          }
        }
        if ( cnt > nvbig ) {
          MidArgs mv = ma;
          mv.ids = ids + nvbig;
          mv.count = cnt - nvbig;
          mid_vec_kernel<<<std::max( 1, std::min( cnt - nvbig, b->xslots ) ), pb::kVecThreads, s3, st>>>( mv );
          ++launches;
        }
        stage_end( b, PALADIN_GPU_STAGE_VECTORS, st, launches );
      }
      PB_CUDA( cudaGetLastError() );
      continue;
    }
    stage_begin( b, PALADIN_GPU_STAGE_GENERIC, st );
    geev_cta_kernel<<<cnt, kGenericThreads, 0, st>>>( bv, ids );
    stage_end( b, PALADIN_GPU_STAGE_GENERIC, st, 1 );
    PB_CUDA( cudaGetLastError() );
  }
  PB_CUDA( cudaStreamSynchronize( st ) );
  b->scattered = false;  // the dense matrices have been destroyed
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_download( paladin_gpu_batch* b, double* wr, double* wi, double* vl, double* vr, int* info ) {
  if ( !b ) return fail( PALADIN_GPU_EINVAL, "null batch" );
  DeviceCtx& c = g_ctx[b->device];
  (void)c;
  PB_CUDA( cudaSetDevice( b->device ) );
  cudaStream_t st = b->stream;
  if ( wr && b->total_n ) PB_CUDA( cudaMemcpyAsync( wr, b->d_wr, sizeof( double ) * b->total_n, cudaMemcpyDeviceToHost, st ) );
  if ( wi && b->total_n ) PB_CUDA( cudaMemcpyAsync( wi, b->d_wi, sizeof( double ) * b->total_n, cudaMemcpyDeviceToHost, st ) );
  if ( info && b->count ) PB_CUDA( cudaMemcpyAsync( info, b->d_info, sizeof( int ) * b->count, cudaMemcpyDeviceToHost, st ) );
  for ( int side = 0; side < 2; ++side ) {
    double* host = side == 0 ? vr : vl;
    const double* dev = side == 0 ? b->d_V : b->d_VL;
    if ( !host || ( side == 0 ? b->jobvr : b->jobvl ) != 'V' ) continue;
    // device matrices are padded and staggered; the host layout is the plain concatenation
    bool padded = false;
    for ( int k = 0; k < b->count; ++k ) padded = padded || ( b->n[k] & 1 ) || ( (int64_t)b->n[k] * b->n[k] >= 4096 );
    bool uniform = b->count > 0;
    for ( int k = 1; k < b->count; ++k ) uniform = uniform && ( b->n[k] == b->n[0] );
    if ( !padded ) {
      if ( b->total_nn ) PB_CUDA( cudaMemcpyAsync( host, dev, sizeof( double ) * b->total_nn, cudaMemcpyDeviceToHost, st ) );
    } else if ( uniform && b->n[0] > 0 ) {
      // equal sizes: one strided copy (device pitch = padded matrix, host pitch = packed matrix)
      const size_t nn = (size_t)b->n[0] * b->n[0];
      PB_CUDA( cudaMemcpy2DAsync( host, nn * sizeof( double ), dev, (size_t)( b->a_off[1] - b->a_off[0] ) * sizeof( double ),
                                  nn * sizeof( double ), b->count, cudaMemcpyDeviceToHost, st ) );
    } else {
      int64_t dst = 0;
      for ( int k = 0; k < b->count; ++k ) {
        const int64_t nn = (int64_t)b->n[k] * b->n[k];
        if ( nn ) PB_CUDA( cudaMemcpyAsync( host + dst, dev + b->a_off[k], sizeof( double ) * nn, cudaMemcpyDeviceToHost, st ) );
        dst += nn;
      }
    }
  }
  PB_CUDA( cudaStreamSynchronize( st ) );
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_vector_filter( paladin_gpu_batch* b, char side, double threshold, double* max2, long long* kept ) {
  if ( !b || !max2 || !kept ) return fail( PALADIN_GPU_EINVAL, "bad arguments" );
  if ( side != 'R' && side != 'L' ) return fail( PALADIN_GPU_EINVAL, "side must be 'R' or 'L'" );
  const double* vecs = side == 'R' ? b->d_V : b->d_VL;
  if ( ( side == 'R' ? b->jobvr : b->jobvl ) != 'V' || !vecs ) return fail( PALADIN_GPU_EINVAL, "the batch holds no vectors of that side" );
  PB_CUDA( cudaSetDevice( b->device ) );
  if ( b->count == 0 ) return PALADIN_GPU_OK;
  cudaStream_t st = b->stream;
  double* d_max2 = b->d_max2;
  unsigned long long* d_kept = b->d_kept;
  int nmax = 0;
  for ( int k = 0; k < b->count; ++k ) nmax = std::max( nmax, b->n[k] );
  cudaError_t e = cudaMemsetAsync( d_max2, 0, sizeof( double ) * std::max<int64_t>( b->total_n, 1 ), st );
  if ( e == cudaSuccess ) e = cudaMemsetAsync( d_kept, 0, sizeof( unsigned long long ) * b->count, st );
  if ( e == cudaSuccess && nmax > 0 ) {
    BatchView bv = batch_view( b );
    // (grid.y is limited to 65535 matrices per launch)
    for ( int k0 = 0; k0 < b->count && e == cudaSuccess; k0 += 65535 ) {
      BatchView bk = bv;
      bk.n += k0; bk.a_off += k0; bk.w_off += k0; bk.info += k0;
      const int cnt = std::min( 65535, b->count - k0 );
      vector_filter_kernel<<<dim3( ( nmax + 7 ) / 8, cnt ), 256, 0, st>>>( bk, vecs, threshold, d_max2, d_kept + k0, cnt );
      e = cudaGetLastError();
    }
  }
  if ( e == cudaSuccess && b->total_n ) e = cudaMemcpyAsync( max2, d_max2, sizeof( double ) * b->total_n, cudaMemcpyDeviceToHost, st );
  if ( e == cudaSuccess ) e = cudaMemcpyAsync( kept, d_kept, sizeof( long long ) * b->count, cudaMemcpyDeviceToHost, st );
  if ( e == cudaSuccess ) e = cudaStreamSynchronize( st );
  if ( e != cudaSuccess ) return fail( PALADIN_GPU_ECUDA, cudaGetErrorString( e ) );
  return PALADIN_GPU_OK;
}

// ---- text ends --------------------------------------------------------------------------------------------------------------
int paladin_gpu_batch_text_begin( paladin_gpu_batch* b, size_t total_bytes ) {
  if ( !b ) return fail( PALADIN_GPU_EINVAL, "null batch" );
  PB_CUDA( cudaSetDevice( b->device ) );
  cudaStream_t st = b->stream;
  g_alloc_stream = st;
  if ( b->text_cap < total_bytes + 64 ) {
    dev_free( b->d_text, st );
    b->d_text = nullptr;
    b->text_cap = 0;
    PB_CUDA( cudaMallocAsync( reinterpret_cast<void**>( &b->d_text ), total_bytes + 64, st ) );
    b->text_cap = total_bytes + 64;
  }
  PB_CUDA( cudaMemsetAsync( b->d_text + total_bytes, '\n', 64, st ) );
  PB_CUDA( cudaStreamSynchronize( st ) );
  b->text_resident = total_bytes;
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_text_put( paladin_gpu_batch* b, size_t offset, const char* src, size_t bytes ) {
  if ( !b || ( bytes > 0 && !src ) ) return fail( PALADIN_GPU_EINVAL, "bad arguments" );
  if ( offset + bytes > b->text_resident ) return fail( PALADIN_GPU_EINVAL, "range outside the text reserved with paladin_gpu_batch_text_begin" );
  if ( bytes == 0 ) return PALADIN_GPU_OK;
  PB_CUDA( cudaSetDevice( b->device ) );
  PB_CUDA( cudaMemcpy( b->d_text + offset, src, bytes, cudaMemcpyHostToDevice ) );
  return PALADIN_GPU_OK;
}
int paladin_gpu_batch_upload_text( paladin_gpu_batch* b, const char* text, const int64_t* text_off, int* needs_host ) {
  if ( !b ) return fail( PALADIN_GPU_EINVAL, "null batch" );
  if ( b->count > 0 && !text_off ) return fail( PALADIN_GPU_EINVAL, "null text offsets" );
  const bool resident = text == nullptr;  // the text was placed on the device with paladin_gpu_batch_text_begin / _put
  PB_CUDA( cudaSetDevice( b->device ) );
  cudaStream_t st = b->stream;
  g_alloc_stream = st;
  const int count = b->count;
  if ( needs_host )
    for ( int k = 0; k < count; ++k ) needs_host[k] = 0;
  if ( count == 0 ) return PALADIN_GPU_OK;
  for ( int k = 0; k < count; ++k )
    if ( text_off[k + 1] < text_off[k] ) return fail( PALADIN_GPU_EINVAL, "decreasing text offsets" );
  const long long t0 = text_off[0];
  const size_t total = (size_t)( text_off[count] - t0 );
  if ( resident ) {
    if ( t0 != 0 || total > b->text_resident ) return fail( PALADIN_GPU_EINVAL, "text offsets do not match the text placed on the device" );
  } else {
    if ( b->text_cap < total + 64 ) {
      dev_free( b->d_text, st );
      b->d_text = nullptr;
      b->text_cap = 0;
      PB_CUDA( cudaMallocAsync( reinterpret_cast<void**>( &b->d_text ), total + 64, st ) );
      b->text_cap = total + 64;
    }
    if ( total ) PB_CUDA( cudaMemcpyAsync( b->d_text, text + t0, total, cudaMemcpyHostToDevice, st ) );
    PB_CUDA( cudaMemsetAsync( b->d_text + total, '\n', 64, st ) );
  }
  // chunk table
  std::vector<pb::TextChunk> chunks;
  std::vector<long long> cfirst( count + 1, 0 ), tb( count ), te( count );
  for ( int k = 0; k < count; ++k ) {
    tb[k] = text_off[k] - t0;
    te[k] = text_off[k + 1] - t0;
    cfirst[k] = (long long)chunks.size();
    int idx = 0;
    for ( long long p = tb[k]; p < te[k]; p += pb::kTxtChunk )
      chunks.push_back( pb::TextChunk{ k, idx++, p, (int)std::min<long long>( pb::kTxtChunk, te[k] - p ) } );
  }
  cfirst[count] = (long long)chunks.size();
  const size_t nchunks = chunks.size();
  // one allocation: chunks | chunk_first | t_begin | t_end | chunk_lines | needs_host | lines_found | n_fallback | fallback
  auto up16 = []( size_t x ) { return ( x + 15 ) & ~(size_t)15; };
  const size_t o_chunks = 0;
  const size_t o_cfirst = o_chunks + up16( sizeof( pb::TextChunk ) * std::max<size_t>( nchunks, 1 ) );
  const size_t o_tb = o_cfirst + up16( sizeof( long long ) * ( count + 1 ) );
  const size_t o_te = o_tb + up16( sizeof( long long ) * count );
  const size_t o_lines = o_te + up16( sizeof( long long ) * count );
  const size_t o_need = o_lines + up16( sizeof( int ) * std::max<size_t>( nchunks, 1 ) );
  const size_t o_found = o_need + up16( sizeof( int ) * count );
  const size_t o_nfb = o_found + up16( sizeof( int ) * count );
  const size_t o_fb = o_nfb + 16;
  const size_t need = o_fb + sizeof( pb::TextFallback ) * (size_t)pb::kTxtMaxFallback;
  if ( b->text_tables_cap < need ) {
    dev_free( b->d_text_tables, st );
    b->d_text_tables = nullptr;
    b->text_tables_cap = 0;
    PB_CUDA( cudaMallocAsync( &b->d_text_tables, need, st ) );
    b->text_tables_cap = need;
  }
  char* tbl = reinterpret_cast<char*>( b->d_text_tables );
  if ( nchunks ) PB_CUDA( cudaMemcpyAsync( tbl + o_chunks, chunks.data(), sizeof( pb::TextChunk ) * nchunks, cudaMemcpyHostToDevice, st ) );
  PB_CUDA( cudaMemcpyAsync( tbl + o_cfirst, cfirst.data(), sizeof( long long ) * ( count + 1 ), cudaMemcpyHostToDevice, st ) );
  PB_CUDA( cudaMemcpyAsync( tbl + o_tb, tb.data(), sizeof( long long ) * count, cudaMemcpyHostToDevice, st ) );
  PB_CUDA( cudaMemcpyAsync( tbl + o_te, te.data(), sizeof( long long ) * count, cudaMemcpyHostToDevice, st ) );
  PB_CUDA( cudaMemsetAsync( tbl + o_need, 0, o_fb - o_need, st ) );
  pb::TextArgs ta;
  ta.text = b->d_text;
  ta.t_begin = reinterpret_cast<const long long*>( tbl + o_tb );
  ta.t_end = reinterpret_cast<const long long*>( tbl + o_te );
  ta.nnz_off = b->d_nnz_off;
  ta.chunks = reinterpret_cast<const pb::TextChunk*>( tbl + o_chunks );
  ta.chunk_first = reinterpret_cast<const long long*>( tbl + o_cfirst );
  ta.chunk_lines = reinterpret_cast<int*>( tbl + o_lines );
  ta.coo_i = b->d_coo_i;
  ta.coo_j = b->d_coo_j;
  ta.coo_v = b->d_coo_v;
  ta.needs_host = reinterpret_cast<int*>( tbl + o_need );
  ta.lines_found = reinterpret_cast<int*>( tbl + o_found );
  ta.fallback = reinterpret_cast<pb::TextFallback*>( tbl + o_fb );
  ta.n_fallback = reinterpret_cast<int*>( tbl + o_nfb );
  if ( nchunks ) {
    pb::text_count_kernel<<<(unsigned)nchunks, pb::kTxtThreads, 0, st>>>( ta );
    pb::text_scan_kernel<<<count, 32, 0, st>>>( ta, count );
    pb::text_parse_kernel<<<(unsigned)nchunks, pb::kTxtThreads, 0, st>>>( ta );
  } else {
    pb::text_scan_kernel<<<count, 32, 0, st>>>( ta, count );
  }
  PB_CUDA( cudaGetLastError() );
  std::vector<int> need_h( count, 0 );
  int nfb = 0;
  PB_CUDA( cudaMemcpyAsync( need_h.data(), tbl + o_need, sizeof( int ) * count, cudaMemcpyDeviceToHost, st ) );
  PB_CUDA( cudaMemcpyAsync( &nfb, tbl + o_nfb, sizeof( int ), cudaMemcpyDeviceToHost, st ) );
  PB_CUDA( cudaStreamSynchronize( st ) );
  if ( nfb > 0 ) {
    // fields the device refused: the C library converts them (atoi / atoi / atof on the field, reference
    // src/square-matrix.hpp:112-114), the results are patched into the device arrays
    const int take = std::min( nfb, pb::kTxtMaxFallback );
    std::vector<pb::TextFallback> fb( take );
    PB_CUDA( cudaMemcpy( fb.data(), tbl + o_fb, sizeof( pb::TextFallback ) * take, cudaMemcpyDeviceToHost ) );
    std::vector<long long> where;
    std::vector<int> vi, vj;
    std::vector<double> vv;
    char line[112];
    for ( const pb::TextFallback& f : fb ) {
      if ( need_h[f.matrix] ) continue;
      const char* p = resident ? line : text + t0 + f.offset;
      const char* end = resident ? line : text + t0 + te[f.matrix];
      if ( resident ) {
        // the line itself comes back from the device (legal lines are at most 5 + 1 + 5 + 1 + 31 + 1 bytes)
        const size_t nb = (size_t)std::min<long long>( (long long)sizeof( line ), te[f.matrix] - f.offset );
        PB_CUDA( cudaMemcpy( line, b->d_text + f.offset, nb, cudaMemcpyDeviceToHost ) );
        end = line + nb;
      }
      const char* fld[3];
      size_t len[3];
      for ( int part = 0; part < 3; ++part ) {
        const char stop = part < 2 ? ' ' : '\n';
        fld[part] = p;
        while ( p < end && *p != stop ) ++p;
        len[part] = (size_t)( p - fld[part] );
        if ( p < end ) ++p;
      }
      char buf[3][40];
      for ( int part = 0; part < 3; ++part ) {
        const size_t l = std::min<size_t>( len[part], 39 );
        std::memcpy( buf[part], fld[part], l );
        buf[part][l] = '\0';
      }
      where.push_back( b->nnz_off[f.matrix] + f.line );
      vi.push_back( std::atoi( buf[0] ) );
      vj.push_back( std::atoi( buf[1] ) );
      vv.push_back( std::atof( buf[2] ) );
    }
    const int np = (int)where.size();
    if ( np > 0 ) {
      long long* d_where = nullptr;
      int *d_vi = nullptr, *d_vj = nullptr;
      double* d_vv = nullptr;
      int rc = dev_alloc( &d_where, np );
      if ( rc == PALADIN_GPU_OK ) rc = dev_alloc( &d_vi, np );
      if ( rc == PALADIN_GPU_OK ) rc = dev_alloc( &d_vj, np );
      if ( rc == PALADIN_GPU_OK ) rc = dev_alloc( &d_vv, np );
      cudaError_t e = cudaSuccess;
      if ( rc == PALADIN_GPU_OK ) {
        e = cudaMemcpy( d_where, where.data(), sizeof( long long ) * np, cudaMemcpyHostToDevice );
        if ( e == cudaSuccess ) e = cudaMemcpy( d_vi, vi.data(), sizeof( int ) * np, cudaMemcpyHostToDevice );
        if ( e == cudaSuccess ) e = cudaMemcpy( d_vj, vj.data(), sizeof( int ) * np, cudaMemcpyHostToDevice );
        if ( e == cudaSuccess ) e = cudaMemcpy( d_vv, vv.data(), sizeof( double ) * np, cudaMemcpyHostToDevice );
        if ( e == cudaSuccess ) {
          pb::text_patch_kernel<<<( np + 255 ) / 256, 256, 0, st>>>( np, d_where, d_vi, d_vj, d_vv, b->d_coo_i, b->d_coo_j, b->d_coo_v );
          e = cudaStreamSynchronize( st );
        }
      }
      dev_free( d_where, st ); dev_free( d_vi, st ); dev_free( d_vj, st ); dev_free( d_vv, st );
      if ( rc != PALADIN_GPU_OK ) return rc;
      if ( e != cudaSuccess ) return fail( PALADIN_GPU_ECUDA, cudaGetErrorString( e ) );
    }
  }
  if ( needs_host )
    for ( int k = 0; k < count; ++k ) needs_host[k] = need_h[k];
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_upload_coo_one( paladin_gpu_batch* b, int k, const int* coo_i, const int* coo_j, const double* coo_v ) {
  if ( !b || k < 0 || k >= b->count ) return fail( PALADIN_GPU_EINVAL, "bad arguments" );
  const int64_t nnz = b->nnz_off[k + 1] - b->nnz_off[k];
  if ( nnz > 0 && ( !coo_i || !coo_j || !coo_v ) ) return fail( PALADIN_GPU_EINVAL, "null COO arrays" );
  PB_CUDA( cudaSetDevice( b->device ) );
  if ( nnz > 0 ) {
    PB_CUDA( cudaMemcpyAsync( b->d_coo_i + b->nnz_off[k], coo_i, sizeof( int ) * nnz, cudaMemcpyHostToDevice, b->stream ) );
    PB_CUDA( cudaMemcpyAsync( b->d_coo_j + b->nnz_off[k], coo_j, sizeof( int ) * nnz, cudaMemcpyHostToDevice, b->stream ) );
    PB_CUDA( cudaMemcpyAsync( b->d_coo_v + b->nnz_off[k], coo_v, sizeof( double ) * nnz, cudaMemcpyHostToDevice, b->stream ) );
  }
  PB_CUDA( cudaStreamSynchronize( b->stream ) );
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_get_coo( paladin_gpu_batch* b, int* coo_i, int* coo_j, double* coo_v ) {
  if ( !b ) return fail( PALADIN_GPU_EINVAL, "null batch" );
  PB_CUDA( cudaSetDevice( b->device ) );
  if ( b->total_nnz > 0 ) {
    if ( coo_i ) PB_CUDA( cudaMemcpyAsync( coo_i, b->d_coo_i, sizeof( int ) * b->total_nnz, cudaMemcpyDeviceToHost, b->stream ) );
    if ( coo_j ) PB_CUDA( cudaMemcpyAsync( coo_j, b->d_coo_j, sizeof( int ) * b->total_nnz, cudaMemcpyDeviceToHost, b->stream ) );
    if ( coo_v ) PB_CUDA( cudaMemcpyAsync( coo_v, b->d_coo_v, sizeof( double ) * b->total_nnz, cudaMemcpyDeviceToHost, b->stream ) );
  }
  PB_CUDA( cudaStreamSynchronize( b->stream ) );
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_format_vectors( paladin_gpu_batch* b, char side, double threshold, int first, int count, char* text, size_t cap,
                                      int64_t* body_off, long long* kept, int* needs_host ) {
  if ( !b || !body_off || !kept || !needs_host ) return fail( PALADIN_GPU_EINVAL, "bad arguments" );
  if ( side != 'R' && side != 'L' ) return fail( PALADIN_GPU_EINVAL, "side must be 'R' or 'L'" );
  if ( first < 0 || count < 0 || first + count > b->count ) return fail( PALADIN_GPU_EINVAL, "matrix range outside the batch" );
  const double* vecs = side == 'R' ? b->d_V : b->d_VL;
  if ( ( side == 'R' ? b->jobvr : b->jobvl ) != 'V' || !vecs ) return fail( PALADIN_GPU_EINVAL, "the batch holds no vectors of that side" );
  PB_CUDA( cudaSetDevice( b->device ) );
  body_off[0] = 0;
  if ( count == 0 ) return PALADIN_GPU_OK;
  cudaStream_t st = b->stream;
  g_alloc_stream = st;
  // 1. the write filter: largest |x|^2 per vector, kept entries per matrix (bit-exact with the reference writer's first pass)
  {
    int nmax = 0;
    for ( int k = first; k < first + count; ++k ) nmax = std::max( nmax, b->n[k] );
    PB_CUDA( cudaMemsetAsync( b->d_kept + first, 0, sizeof( unsigned long long ) * count, st ) );
    if ( nmax > 0 ) {
      BatchView bv = batch_view( b );
      for ( int k0 = first; k0 < first + count; k0 += 65535 ) {
        BatchView bk = bv;
        bk.n += k0; bk.a_off += k0; bk.w_off += k0; bk.info += k0;
        const int cnt = std::min( 65535, first + count - k0 );
        vector_filter_kernel<<<dim3( ( nmax + 7 ) / 8, cnt ), 256, 0, st>>>( bk, vecs, threshold, b->d_max2, b->d_kept + k0, cnt );
      }
      PB_CUDA( cudaGetLastError() );
    }
    PB_CUDA( cudaMemcpyAsync( kept, b->d_kept + first, sizeof( long long ) * count, cudaMemcpyDeviceToHost, st ) );
  }
  // 2. scratch of the formatter: groups of consecutive matrices with at most fmt_cap_entries vector components
  const long long want_entries = 48LL << 20;
  long long biggest = 0, range_entries = 0;
  for ( int k = first; k < first + count; ++k ) {
    biggest = std::max( biggest, (long long)b->n[k] * b->n[k] );
    range_entries += (long long)b->n[k] * b->n[k];
  }
  const long long cap_entries = std::max( std::max( std::min( want_entries, range_entries ), biggest ), 1LL );
  if ( b->fmt_cap_entries < cap_entries ) {
    dev_free( b->d_fmt_slots, st ); dev_free( b->d_fmt_out, st ); dev_free( b->d_fmt_lens, st ); dev_free( b->d_fmt_bsum, st ); dev_free( b->d_fmt_eoff, st ); dev_free( b->d_fmt_flags, st );
    b->d_fmt_slots = b->d_fmt_out = nullptr;
    b->d_fmt_lens = nullptr;
    b->d_fmt_bsum = nullptr;
    b->d_fmt_eoff = nullptr;
    b->d_fmt_flags = nullptr;
    b->fmt_cap_entries = 0;
    PB_CUDA( cudaMallocAsync( reinterpret_cast<void**>( &b->d_fmt_slots ), (size_t)cap_entries * pb::kFmtSlot, st ) );
    PB_CUDA( cudaMallocAsync( reinterpret_cast<void**>( &b->d_fmt_out ), (size_t)cap_entries * pb::kFmtSlot, st ) );
    PB_CUDA( cudaMallocAsync( reinterpret_cast<void**>( &b->d_fmt_lens ), (size_t)cap_entries, st ) );
    PB_CUDA( cudaMallocAsync( reinterpret_cast<void**>( &b->d_fmt_bsum ), sizeof( unsigned long long ) * ( (size_t)( cap_entries / pb::kFmtPerBlock ) + 2 ), st ) );
    PB_CUDA( cudaMallocAsync( reinterpret_cast<void**>( &b->d_fmt_eoff ), sizeof( long long ) * 2 * ( (size_t)b->count + 1 ), st ) );
    PB_CUDA( cudaMallocAsync( reinterpret_cast<void**>( &b->d_fmt_flags ), sizeof( int ) * (size_t)b->count, st ) );
    b->fmt_cap_entries = cap_entries;
  }
  PB_CUDA( cudaMemsetAsync( b->d_fmt_flags + first, 0, sizeof( int ) * count, st ) );
  size_t written = 0;
  std::vector<long long> eoff, boff;
  int k0 = first;
  const int kend = first + count;
  while ( k0 < kend ) {
    int k1 = k0;
    long long entries = 0;
    while ( k1 < kend && ( k1 == k0 || entries + (long long)b->n[k1] * b->n[k1] <= b->fmt_cap_entries ) ) {
      entries += (long long)b->n[k1] * b->n[k1];
      ++k1;
    }
    const int cnt = k1 - k0;
    eoff.assign( cnt + 1, 0 );
    for ( int q = 0; q < cnt; ++q ) eoff[q + 1] = eoff[q] + (long long)b->n[k0 + q] * b->n[k0 + q];
    if ( entries > 0 ) {
      PB_CUDA( cudaMemcpyAsync( b->d_fmt_eoff, eoff.data(), sizeof( long long ) * ( cnt + 1 ), cudaMemcpyHostToDevice, st ) );
      pb::FmtArgs fa;
      fa.n = b->d_n + k0;
      fa.a_off = b->d_a_off + k0;
      fa.w_off = b->d_w_off + k0;
      fa.e_off = b->d_fmt_eoff;
      fa.count = cnt;
      fa.vecs = vecs;
      fa.wi = b->d_wi;
      fa.max2 = b->d_max2;
      fa.info = b->d_info + k0;
      fa.threshold = threshold;
      fa.slots = b->d_fmt_slots;
      fa.lens = b->d_fmt_lens;
      fa.needs_host = b->d_fmt_flags + k0;
      const long long nblocks = ( entries + pb::kFmtPerBlock - 1 ) / pb::kFmtPerBlock;
      long long* d_boff = b->d_fmt_eoff + ( b->count + 1 );
      pb::fmt_lines_kernel<<<(unsigned)( ( entries + pb::kFmtThreads - 1 ) / pb::kFmtThreads ), pb::kFmtThreads, 0, st>>>( fa, entries );
      pb::fmt_blocksum_kernel<<<(unsigned)nblocks, pb::kFmtThreads, 0, st>>>( b->d_fmt_lens, entries, b->d_fmt_bsum );
      pb::fmt_scan_kernel<<<1, 1024, 0, st>>>( b->d_fmt_bsum, nblocks );
      pb::fmt_bodyoff_kernel<<<( cnt + 1 + 127 ) / 128, 128, 0, st>>>( fa, entries, nblocks, b->d_fmt_bsum, d_boff );
      pb::fmt_compact_kernel<<<(unsigned)nblocks, pb::kFmtThreads, 0, st>>>( fa, entries, b->d_fmt_bsum, b->d_fmt_out );
      PB_CUDA( cudaGetLastError() );
      boff.assign( cnt + 1, 0 );
      PB_CUDA( cudaMemcpyAsync( boff.data(), d_boff, sizeof( long long ) * ( cnt + 1 ), cudaMemcpyDeviceToHost, st ) );
      PB_CUDA( cudaStreamSynchronize( st ) );
      const size_t bytes = (size_t)boff[cnt];
      if ( text == nullptr ) {
        // the compacted text stays on the device for paladin_gpu_batch_fetch_text: only possible for a single group
        if ( k0 != first || k1 != kend ) return fail( PALADIN_GPU_EINVAL, "range too large to keep the formatted text on the device (48 Mi components at most)" );
      } else {
        if ( written + bytes > cap ) return fail( PALADIN_GPU_EINVAL, "text buffer too small for the formatted vectors (48 bytes per component always suffice)" );
        if ( bytes ) PB_CUDA( cudaMemcpyAsync( text + written, b->d_fmt_out, bytes, cudaMemcpyDeviceToHost, st ) );
      }
      for ( int q = 0; q < cnt; ++q ) body_off[k0 - first + q + 1] = (int64_t)( written + (size_t)boff[q + 1] );
      PB_CUDA( cudaStreamSynchronize( st ) );
      written += bytes;
    } else {
      for ( int q = 0; q < cnt; ++q ) body_off[k0 - first + q + 1] = (int64_t)written;
    }
    k0 = k1;
  }
  PB_CUDA( cudaMemcpyAsync( needs_host, b->d_fmt_flags + first, sizeof( int ) * count, cudaMemcpyDeviceToHost, st ) );
  PB_CUDA( cudaStreamSynchronize( st ) );
  b->fmt_text_bytes = text == nullptr ? written : 0;
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_fetch_text( paladin_gpu_batch* b, size_t offset, size_t bytes, char* dst ) {
  if ( !b || ( bytes > 0 && !dst ) ) return fail( PALADIN_GPU_EINVAL, "bad arguments" );
  if ( offset + bytes > b->fmt_text_bytes ) return fail( PALADIN_GPU_EINVAL, "range outside the formatted text kept on the device" );
  if ( bytes == 0 ) return PALADIN_GPU_OK;
  PB_CUDA( cudaSetDevice( b->device ) );
  // (a copy of its own: several writer threads fetch their files side by side; the legacy-blocking semantics of cudaMemcpy
  // do not apply to the batch's non-blocking stream, whose work has been synchronised by the format call)
  PB_CUDA( cudaMemcpy( dst, b->d_fmt_out + offset, bytes, cudaMemcpyDeviceToHost ) );
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_stage_ms( paladin_gpu_batch* b, float* ms, int* launches ) {
  if ( !b || !ms ) return fail( PALADIN_GPU_EINVAL, "bad arguments" );
  DeviceCtx& c = g_ctx[b->device];
  (void)c;
  PB_CUDA( cudaSetDevice( b->device ) );
  for ( int s = 0; s < PALADIN_GPU_NUM_STAGES; ++s ) {
    ms[s] = 0.0f;
    if ( launches ) launches[s] = b->st[s].launches;
    if ( b->st[s].used ) {
      cudaEventSynchronize( b->st[s].e1 );
      cudaEventElapsedTime( &ms[s], b->st[s].e0, b->st[s].e1 );
    }
  }
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_phase_cycles( paladin_gpu_batch* b, long long* cycles ) {
  if ( !b || !cycles ) return fail( PALADIN_GPU_EINVAL, "bad arguments" );
  DeviceCtx& c = g_ctx[b->device];
  (void)c;
  PB_CUDA( cudaSetDevice( b->device ) );
  PB_CUDA( cudaMemcpy( cycles, b->d_stats, sizeof( long long ) * kNumTicks, cudaMemcpyDeviceToHost ) );
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_timer_start( paladin_gpu_batch* b ) {
  if ( !b ) return fail( PALADIN_GPU_EINVAL, "null batch" );
  PB_CUDA( cudaSetDevice( b->device ) );
  PB_CUDA( cudaEventRecord( b->t0, b->stream ) );
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_timer_stop( paladin_gpu_batch* b, float* ms ) {
  if ( !b || !ms ) return fail( PALADIN_GPU_EINVAL, "bad arguments" );
  PB_CUDA( cudaSetDevice( b->device ) );
  PB_CUDA( cudaEventRecord( b->t1, b->stream ) );
  PB_CUDA( cudaEventSynchronize( b->t1 ) );
  PB_CUDA( cudaEventElapsedTime( ms, b->t0, b->t1 ) );
  return PALADIN_GPU_OK;
}

int paladin_gpu_batch_destroy( paladin_gpu_batch* b ) {
  if ( !b ) return PALADIN_GPU_OK;
  free_batch( b );
  return PALADIN_GPU_OK;
}

int paladin_gpu_geev_batch( int device, int count, const int* n, const int64_t* nnz_offsets, const int* coo_i,
                            const int* coo_j, const double* coo_v, char jobvl, char jobvr, double* const* wr,
                            double* const* wi, double* const* vl, double* const* vr, int* info ) {
  if ( count < 0 || ( count > 0 && ( !wr || !wi || !info ) ) ) return fail( PALADIN_GPU_EINVAL, "null output arrays" );
  if ( jobvr == 'V' && count > 0 && !vr ) return fail( PALADIN_GPU_EINVAL, "vr is null but jobvr == 'V'" );
  if ( jobvl == 'V' && count > 0 && !vl ) return fail( PALADIN_GPU_EINVAL, "vl is null but jobvl == 'V'" );
  paladin_gpu_batch* b = nullptr;
  PB_TRY( paladin_gpu_batch_create( device, count, n, nnz_offsets, jobvl, jobvr, &b ) );
  int rc = paladin_gpu_batch_upload( b, coo_i ? coo_i + nnz_offsets[0] : nullptr, coo_j ? coo_j + nnz_offsets[0] : nullptr,
                                     coo_v ? coo_v + nnz_offsets[0] : nullptr );
  if ( rc == PALADIN_GPU_OK ) rc = paladin_gpu_batch_scatter( b );
  if ( rc == PALADIN_GPU_OK ) rc = paladin_gpu_batch_solve( b );
  if ( rc == PALADIN_GPU_OK ) {
    // results straight into the caller's per-matrix buffers
    cudaSetDevice( device );
    cudaStream_t st = b->stream;
    cudaError_t e = cudaMemcpyAsync( info, b->d_info, sizeof( int ) * count, cudaMemcpyDeviceToHost, st );
    for ( int k = 0; k < count && e == cudaSuccess; ++k ) {
      const int nk = b->n[k];
      if ( nk == 0 ) continue;
      e = cudaMemcpyAsync( wr[k], b->d_wr + b->w_off[k], sizeof( double ) * nk, cudaMemcpyDeviceToHost, st );
      if ( e == cudaSuccess ) e = cudaMemcpyAsync( wi[k], b->d_wi + b->w_off[k], sizeof( double ) * nk, cudaMemcpyDeviceToHost, st );
      if ( e == cudaSuccess && jobvr == 'V' && vr[k] )
        e = cudaMemcpyAsync( vr[k], b->d_V + b->a_off[k], sizeof( double ) * nk * nk, cudaMemcpyDeviceToHost, st );
      if ( e == cudaSuccess && jobvl == 'V' && vl[k] )
        e = cudaMemcpyAsync( vl[k], b->d_VL + b->a_off[k], sizeof( double ) * nk * nk, cudaMemcpyDeviceToHost, st );
    }
    if ( e == cudaSuccess ) e = cudaStreamSynchronize( st );
    if ( e != cudaSuccess ) rc = fail( PALADIN_GPU_ECUDA, cudaGetErrorString( e ) );
  }
  paladin_gpu_batch_destroy( b );
  return rc;
}

void paladin_gpu_dgeev_( const char* jobvl, const char* jobvr, int* n, double* a, int* lda, double* wr, double* wi,
                         double* vl, int* ldvl, double* vr, int* ldvr, double* work, int* lwork, int* info ) {
  if ( !info ) return;
  *info = 0;
  const char jl = jobvl ? ( *jobvl == 'v' ? 'V' : ( *jobvl == 'n' ? 'N' : *jobvl ) ) : '?';
  const char jr = jobvr ? ( *jobvr == 'v' ? 'V' : ( *jobvr == 'n' ? 'N' : *jobvr ) ) : '?';
  if ( jl != 'N' && jl != 'V' ) { *info = -1; return; }
  if ( jr != 'N' && jr != 'V' ) { *info = -2; return; }
  if ( !n || *n < 0 ) { *info = -3; return; }
  if ( !lda || *lda < std::max( 1, *n ) ) { *info = -5; return; }
  if ( !ldvl || *ldvl < 1 || ( jl == 'V' && *ldvl < *n ) ) { *info = -9; return; }
  if ( !ldvr || *ldvr < 1 || ( jr == 'V' && *ldvr < *n ) ) { *info = -11; return; }
  if ( lwork && *lwork == -1 ) {  // workspace query (reference src/paladin.hpp:247-249): the device path owns its workspace
    if ( work ) work[0] = 1.0;
    return;
  }
  const int N = *n;
  if ( N == 0 ) return;
  int device = 0;
  if ( const char* env = std::getenv( "PALADIN_GPU_DEVICE" ) ) device = std::atoi( env );
  const int64_t offs[2] = { 0, 0 };
  paladin_gpu_batch* b = nullptr;
  int rc = paladin_gpu_batch_create( device, 1, &N, offs, jl, jr, &b );
  if ( rc != PALADIN_GPU_OK ) {
    std::fprintf( stderr, "paladin_gpu_dgeev_: %s\n", paladin_gpu_last_error() );
    *info = -1000 - rc;  // no CPU fallback: fail loudly
    return;
  }
  std::vector<double> dense( (size_t)N * N ), w( 2 * (size_t)N ), VR( jr == 'V' ? (size_t)N * N : 0 ), VL( jl == 'V' ? (size_t)N * N : 0 );
  for ( int c2 = 0; c2 < N; ++c2 ) std::memcpy( &dense[(size_t)c2 * N], a + (size_t)c2 * *lda, sizeof( double ) * N );
  rc = paladin_gpu_batch_upload_dense( b, dense.data() );
  if ( rc == PALADIN_GPU_OK ) rc = paladin_gpu_batch_solve( b );
  if ( rc == PALADIN_GPU_OK )
    rc = paladin_gpu_batch_download( b, w.data(), w.data() + N, jl == 'V' ? VL.data() : nullptr, jr == 'V' ? VR.data() : nullptr, info );
  paladin_gpu_batch_destroy( b );
  if ( rc != PALADIN_GPU_OK ) {
    std::fprintf( stderr, "paladin_gpu_dgeev_: %s\n", paladin_gpu_last_error() );
    *info = -1000 - rc;
    return;
  }
  std::memcpy( wr, w.data(), sizeof( double ) * N );
  std::memcpy( wi, w.data() + N, sizeof( double ) * N );
  if ( jr == 'V' && vr )
    for ( int c2 = 0; c2 < N; ++c2 ) std::memcpy( vr + (size_t)c2 * *ldvr, &VR[(size_t)c2 * N], sizeof( double ) * N );
  if ( jl == 'V' && vl )
    for ( int c2 = 0; c2 < N; ++c2 ) std::memcpy( vl + (size_t)c2 * *ldvl, &VL[(size_t)c2 * N], sizeof( double ) * N );
}

}  // extern "C"
