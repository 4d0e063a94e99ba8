// Fast triangular solve over batch rows (LUTransform.backward and its input gradient), fp32.
//
//   X E^T = R   (each row r: E x_r = rhs_r),  E = D x D triangular,  E(i,j) = TRANS ? T[j*D+i] : T[i*D+j]
//
// Two kernels:
//   usf_tri_diag_inv : inverts the 32x32 diagonal blocks of E once (nblk CTAs).
//   usf_trsm_fast    : one CTA owns 32 batch rows and keeps them RESIDENT IN SHARED MEMORY (32 x D fp32,
//                      100 KB at D=784) while it walks the 32-wide column blocks in dependency order:
//                        acc   = X[:, solved blocks] * E[block j, solved blocks]^T   (E tiles streamed from L2,
//                                register-prefetched one tile ahead, transposed through smem)
//                        X_j   = (R_j - acc) * inv(E_jj)^T                           (dense 32x32x32 product)
//                      so there is no serial substitution and no global re-read of X.
// Bound: FFMA (32*D^2/2 FMA per CTA) once latency is hidden; replaces the first-cut kernel in usf_simt.cu
// (one global round trip + two __syncthreads per 32x32 tile, ~1.3 ms per solve at D=784) by ~0.1 ms.
#include "usf_common.cuh"

namespace usf {
namespace {

constexpr int NB = 32;        // column block
constexpr int ROWS = 32;      // batch rows per CTA
constexpr int THREADS = 256;   // 2 k-halves x (32 columns x 4 row groups); each thread owns 8 rows x 1 column
constexpr int EP = NB + 4;      // smem row pitch of the 32x32 tiles: 16-byte aligned rows, conflict-free float4 reads

template <bool TRANS>
__device__ __forceinline__ float tri_elem(const float* __restrict__ T, int64_t D, int64_t i, int64_t j) {
  return TRANS ? T[j * D + i] : T[i * D + j];
}

// Dinv[jb][r][c] = (E_jj)^{-1}[r][c] for every diagonal block jb (identity-padded past D).
template <bool LOWER, bool UNIT, bool TRANS>
__device__ __forceinline__ void tri_diag_inv_body(const float* __restrict__ T, int64_t D, float* __restrict__ Dinv) {
  __shared__ float Es[NB][NB + 1];
  __shared__ float Ys[NB][NB + 1];   // Ys[row][col] = inverse
  const int jb = blockIdx.x, c = threadIdx.x;
  const int64_t j0 = (int64_t)jb * NB;
  for (int r = 0; r < NB; ++r) {
    const int64_t gi = j0 + r, gj = j0 + c;
    float v = (r == c) ? 1.f : 0.f;
    if (gi < D && gj < D) {
      const bool in_tri = LOWER ? (c <= r) : (c >= r);
      v = in_tri ? tri_elem<TRANS>(T, D, gi, gj) : 0.f;
      if (UNIT && r == c) v = 1.f;
    }
    Es[r][c] = v;
  }
  __syncthreads();
  // thread c solves E y = e_c (column c of the inverse)
  if (LOWER) {
    for (int r = 0; r < NB; ++r) {
      float s = (r == c) ? 1.f : 0.f;
      for (int k = 0; k < r; ++k) s = fmaf(-Es[r][k], Ys[k][c], s);
      Ys[r][c] = UNIT ? s : s / Es[r][r];
    }
  } else {
    for (int r = NB - 1; r >= 0; --r) {
      float s = (r == c) ? 1.f : 0.f;
      for (int k = r + 1; k < NB; ++k) s = fmaf(-Es[r][k], Ys[k][c], s);
      Ys[r][c] = UNIT ? s : s / Es[r][r];
    }
  }
  __syncthreads();
  for (int r = 0; r < NB; ++r) Dinv[((int64_t)jb * NB + r) * NB + c] = Ys[r][c];
}

template <bool LOWER, bool UNIT, bool TRANS>
__global__ void __launch_bounds__(NB)
usf_tri_diag_inv_kernel(const float* __restrict__ T, int64_t D, float* __restrict__ Dinv) {
  tri_diag_inv_body<LOWER, UNIT, TRANS>(T, D, Dinv);
}

// Both factors of an LU pair at once (usf_lu_inverse): y = 0 -> U^T (lower, non-unit), y = 1 -> L^T (upper, unit).
__global__ void __launch_bounds__(NB)
usf_lu_diag_inv_kernel(const float* __restrict__ L_raw, const float* __restrict__ U_raw, int64_t D,
                       float* __restrict__ DinvU, float* __restrict__ DinvL) {
  if (blockIdx.y == 0) tri_diag_inv_body<true, false, true>(U_raw, D, DinvU);
  else tri_diag_inv_body<false, true, true>(L_raw, D, DinvL);
}

constexpr int KC = 128;         // k-chunk streamed per iteration (4 column blocks): amortises the L2 latency of E tiles
constexpr int EKP = NB + 4;     // pitch of the k-major E chunk  Es[kk][c]  (16-byte aligned rows: one float4 = 4 columns)
constexpr int XPAD = 4;         // Xs row pitch = Dp + 4: four different rows read by one warp fall into different banks
constexpr int KSPLIT = 4;       // k-splits of a chunk (32 k each): 64 threads x (4 rows x 4 columns) per split

// IDENT: the right-hand sides are the rows of the identity (X = E^{-T} row by row, i.e. the inverse of a triangular
// matrix).  Row r of the solution then vanishes before (LOWER) / after (upper) column r, so a CTA skips the column
// blocks on that side of its own rows and starts every accumulation at them: half the work of a dense solve.
//
// Accumulation: thread (ks, rq, cq) owns rows rq + 8 i (a warp's four row values are neighbours: with the row pitch
// Dp + 4 their float4 reads fall into different banks) x columns 4 cq .. +3 of the 32 x 32 block over the k-quarter
// ks of every chunk: per 4 k it reads 4 float4 of X (4 k of one row each, warp-broadcast) and 4 float4 of E (4 columns
// at one k each) for 64 FMAs -- 8 shared-memory reads per 64 FMAs; the first form of this kernel (8 rows x 1 column per
// thread) needed 9 per 32 and was bound by them.
template <bool LOWER, bool TRANS, bool IDENT>
__device__ __forceinline__ void trsm_fast_body(const float* __restrict__ T, int64_t D, int Dp,
                                               const float* __restrict__ Dinv, const float* rhs, int64_t ldr,
                                               const float* __restrict__ bias, float* X, int64_t ldx, int64_t B) {
  extern __shared__ __align__(16) float sm[];
  const int XP = Dp + XPAD;
  float* Xs = sm;                                                                       // [ROWS][XP]
  float* Es = sm + (size_t)ROWS * XP;                                                    // [KC][EKP]   Es[kk][c]
  float(*Ts)[EP] = reinterpret_cast<float(*)[EP]>(Es + KC * EKP);                        // [ROWS][EP] (16B aligned)
  float(*Ds)[EP] = reinterpret_cast<float(*)[EP]>(Es + KC * EKP + ROWS * EP);            // [NB][EP]
  float(*Ps)[ROWS][EP] = reinterpret_cast<float(*)[ROWS][EP]>(Es + KC * EKP + (ROWS + NB) * EP);   // [KSPLIT] partial sums
  const int tid = threadIdx.x;
  const int ks = tid >> 6, rq = (tid & 63) >> 3, cq = tid & 7;      // accumulation: k-split, row quad, column quad
  const int kh = tid >> 7, c = tid & 31, rg = (tid & 127) >> 5;      // diagonal-block stage: column c, rows rg*8 + kh*4 .. +3
  const int64_t r0 = (int64_t)blockIdx.x * ROWS;
  const int nblk = Dp / NB;

  // resident right-hand sides: Xs = rhs - bias (zero padded)
  for (int e = tid; e < ROWS * Dp; e += THREADS) {
    const int r = e / Dp, k = e - r * Dp;
    const int64_t gr = r0 + r;
    float v = 0.f;
    if (IDENT) {
      v = (gr < B && gr == k) ? 1.f : 0.f;
    } else if (gr < B && k < D) {
      v = rhs[gr * ldr + k];
      if (bias != nullptr) v -= bias[k];
    }
    Xs[(size_t)r * XP + k] = v;
  }
  __syncthreads();

  // E chunk: rows of block jb, columns [k0, k0+KC) restricted to the already-solved range [lo, hi)
  auto load_chunk = [&](int jb, int k0, int lo, int hi, float (&reg)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int e = tid + i * THREADS;              // 0..4095
      int cc, kk;
      if (TRANS) { kk = e >> 5; cc = e & 31; } else { cc = e >> 7; kk = e & 127; }
      const int64_t gi = (int64_t)jb * NB + cc;
      const int gk = k0 + kk;
      reg[i] = (gi < D && gk >= lo && gk < hi && gk < D) ? tri_elem<TRANS>(T, D, gi, gk) : 0.f;
    }
  };
  auto store_chunk = [&](const float (&reg)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int e = tid + i * THREADS;
      int cc, kk;
      if (TRANS) { kk = e >> 5; cc = e & 31; } else { cc = e >> 7; kk = e & 127; }
      Es[kk * EKP + cc] = reg[i];
    }
  };

  const int own = (int)blockIdx.x;                 // (ROWS == NB: the column block of this CTA's own rows)
  for (int step = 0; step < nblk; ++step) {
    const int jb = LOWER ? step : nblk - 1 - step;
    if (IDENT && (LOWER ? jb < own : jb > own)) continue;      // still zero there (uniform over the CTA)
    // solved column range feeding block jb
    const int lo = LOWER ? (IDENT ? own * NB : 0) : (jb + 1) * NB;
    const int hi = LOWER ? jb * NB : (IDENT ? (own + 1) * NB : Dp);
    const int nchunk = (hi - lo + KC - 1) / KC;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float reg[16];
    if (nchunk > 0) load_chunk(jb, lo, lo, hi, reg);
    // the inverse of the diagonal block is independent of the accumulation: fetch it early
    float dreg[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) dreg[i] = Dinv[(int64_t)jb * NB * NB + tid + i * THREADS];
    for (int q = 0; q < nchunk; ++q) {
      const int k0 = lo + q * KC;
      __syncthreads();                 // previous chunk fully consumed
      store_chunk(reg);
      __syncthreads();
      if (q + 1 < nchunk) load_chunk(jb, k0 + KC, lo, hi, reg);   // prefetch the next chunk (overlaps the FMAs below)
      const int kw = hi - k0 < KC ? hi - k0 : KC;                  // multiple of 32
      if (ks * 32 < kw) {                                          // this split's k-quarter is part of the chunk
        const float* xrow = Xs + (size_t)rq * XP + k0 + ks * 32;        // rows rq, rq + 8, rq + 16, rq + 24
        const float* erow = Es + (ks * 32) * EKP + cq * 4;
#pragma unroll 2
        for (int kk = 0; kk < 32; kk += 4) {
          float4 xv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xrow + (size_t)(8 * i) * XP + kk);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 ev = *reinterpret_cast<const float4*>(erow + (kk + u) * EKP);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float xs = u == 0 ? xv[i].x : (u == 1 ? xv[i].y : (u == 2 ? xv[i].z : xv[i].w));
              acc[i][0] = fmaf(xs, ev.x, acc[i][0]);
              acc[i][1] = fmaf(xs, ev.y, acc[i][1]);
              acc[i][2] = fmaf(xs, ev.z, acc[i][2]);
              acc[i][3] = fmaf(xs, ev.w, acc[i][3]);
            }
          }
        }
      }
    }
    // partial sums of the k-splits, then Ts = R_j - sum ; Ds = inverse of the diagonal block
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(&Ps[ks][rq + 8 * i][cq * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * THREADS;
      Ds[e >> 5][e & 31] = dreg[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * THREADS, r = e >> 5, cc = e & 31;
      Ts[r][cc] = Xs[(size_t)r * XP + jb * NB + cc] - (Ps[0][r][cc] + Ps[1][r][cc]) - (Ps[2][r][cc] + Ps[3][r][cc]);
    }
    __syncthreads();
    // x[r][c] = sum_c' Dinv[c][c'] * t[r][c']   (each k-half group takes 4 of the thread's 8 rows)
    float xo[4] = {0.f, 0.f, 0.f, 0.f};
    const int rb = rg * 8 + kh * 4;
#pragma unroll
    for (int k2 = 0; k2 < NB; k2 += 4) {
      const float4 dv = *reinterpret_cast<const float4*>(&Ds[c][k2]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 tv = *reinterpret_cast<const float4*>(&Ts[rb + i][k2]);
        xo[i] = fmaf(tv.x, dv.x, xo[i]);
        xo[i] = fmaf(tv.y, dv.y, xo[i]);
        xo[i] = fmaf(tv.z, dv.z, xo[i]);
        xo[i] = fmaf(tv.w, dv.w, xo[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) Xs[(size_t)(rb + i) * XP + jb * NB + c] = xo[i];
    __syncthreads();
  }

  for (int e = tid; e < ROWS * Dp; e += THREADS) {
    const int r = e / Dp, k = e - r * Dp;
    const int64_t gr = r0 + r;
    if (gr < B && k < D) X[gr * ldx + k] = Xs[(size_t)r * XP + k];
  }
}

template <bool LOWER, bool TRANS>
__global__ void __launch_bounds__(THREADS)
usf_trsm_fast_kernel(const float* __restrict__ T, int64_t D, int Dp, const float* __restrict__ Dinv,
                     const float* rhs, int64_t ldr, const float* __restrict__ bias, float* X, int64_t ldx, int64_t B) {
  trsm_fast_body<LOWER, TRANS, false>(T, D, Dp, Dinv, rhs, ldr, bias, X, ldx, B);
}

// Z = U^{-1} (y = 0) and W = L^{-1} (y = 1), row by row, in one launch (usf_lu_inverse).
__global__ void __launch_bounds__(THREADS, 1)
usf_lu_tri_inverse_kernel(const float* __restrict__ L_raw, const float* __restrict__ U_raw, int64_t D, int Dp,
                          const float* __restrict__ DinvU, const float* __restrict__ DinvL, float* Z, float* W) {
  if (blockIdx.y == 0) trsm_fast_body<true, true, true>(U_raw, D, Dp, DinvU, nullptr, 0, nullptr, Z, D, D);
  else trsm_fast_body<false, true, true>(L_raw, D, Dp, DinvL, nullptr, 0, nullptr, W, D, D);
}

size_t fast_smem_bytes(int Dp) {
  return sizeof(float) * ((size_t)ROWS * (Dp + XPAD) + KC * EKP + (ROWS + NB) * EP + (size_t)KSPLIT * ROWS * EP);
}

}  // namespace

// floats of scratch needed by one triangular solve of size D (the diagonal-block inverses)
int64_t trsm_fast_scratch_floats(int64_t D) { return round_up(D, NB) * NB; }

bool trsm_fast_supported(int64_t D) { return fast_smem_bytes((int)round_up(D, NB)) <= 200 * 1024; }

int trsm_rows_fast(const float* T, int64_t D, bool lower, bool unit, bool trans, const float* rhs, int64_t ldr,
                   const float* bias, float* X, int64_t ldx, int64_t B, float* scratch, cudaStream_t stream) {
  if (B <= 0 || D <= 0) return USF_OK;
  const int Dp = (int)round_up(D, NB);
  const int nblk = Dp / NB;
  const size_t smem = fast_smem_bytes(Dp);
#define USF_DINV(L, U, TT) usf_tri_diag_inv_kernel<L, U, TT><<<nblk, NB, 0, stream>>>(T, D, scratch)
  if (lower) {
    if (unit) { if (trans) USF_DINV(true, true, true); else USF_DINV(true, true, false); }
    else      { if (trans) USF_DINV(true, false, true); else USF_DINV(true, false, false); }
  } else {
    if (unit) { if (trans) USF_DINV(false, true, true); else USF_DINV(false, true, false); }
    else      { if (trans) USF_DINV(false, false, true); else USF_DINV(false, false, false); }
  }
#undef USF_DINV
  USF_LAUNCH_CHECK("usf_tri_diag_inv_kernel");
  dim3 grid((unsigned)ceil_div(B, ROWS));
#define USF_FAST(L, TT)                                                                                              \
  do {                                                                                                               \
    USF_CUDA(cudaFuncSetAttribute(usf_trsm_fast_kernel<L, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    usf_trsm_fast_kernel<L, TT><<<grid, THREADS, smem, stream>>>(T, D, Dp, scratch, rhs, ldr, bias, X, ldx, B);        \
  } while (0)
  if (lower) { if (trans) USF_FAST(true, true); else USF_FAST(true, false); }
  else       { if (trans) USF_FAST(false, true); else USF_FAST(false, false); }
#undef USF_FAST
  USF_LAUNCH_CHECK("usf_trsm_fast_kernel");
  return USF_OK;
}

// A = (L U)^{-1} (row-major, dense), L = unit lower / U = upper triangle of the raw factors: the two triangular
// inverses in ONE launch (identity right-hand sides: half the work of a solve each, and the two run side by side),
// then A = U^{-1} L^{-1} by the fp32 GEMM.  scratch: 2 D^2 + 2 round_up(D, 32) * 32 floats.
int64_t lu_inverse_scratch_floats(int64_t D) { return 2 * D * D + 2 * round_up(D, NB) * NB; }

int lu_inverse(const float* L_raw, const float* U_raw, int64_t D, float* A, float* scratch, cudaStream_t stream) {
  // A == nullptr: the factors' inverses only (Z = U^{-1} at scratch, W = L^{-1} at scratch + D^2): the caller
  // multiplies them (the host side uses the 3xTF32 tensor-core GEMM where its shapes allow)
  if (D <= 0) return USF_OK;
  const int Dp = (int)round_up(D, NB);
  const int nblk = Dp / NB;
  float* Z = scratch;
  float* W = Z + D * D;
  float* DinvU = W + D * D;
  float* DinvL = DinvU + (int64_t)Dp * NB;
  usf_lu_diag_inv_kernel<<<dim3(nblk, 2), NB, 0, stream>>>(L_raw, U_raw, D, DinvU, DinvL);
  USF_LAUNCH_CHECK("usf_lu_diag_inv_kernel");
  const size_t smem = fast_smem_bytes(Dp);
  USF_CUDA(cudaFuncSetAttribute(usf_lu_tri_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  usf_lu_tri_inverse_kernel<<<dim3(nblk, 2), THREADS, smem, stream>>>(L_raw, U_raw, D, Dp, DinvU, DinvL, Z, W);
  USF_LAUNCH_CHECK("usf_lu_tri_inverse_kernel");
  if (A == nullptr) return USF_OK;
  // A[i][j] = sum_k Z[i][k] W[k][j]
  return simt_gemm_plain(Z, D, 0, W, D, 1, D, D, D, A, D, 0, stream);
}

}  // namespace usf
