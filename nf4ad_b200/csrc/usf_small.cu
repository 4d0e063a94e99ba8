// Whole-stack kernel for small event shapes (D + context <= 64, layer widths <= 128): ONE launch runs the complete
// launch chain of usf_stack_run -- every affine map, conditioner layer, coupling update and the base log-density --
// for a tile of 64 rows that never leaves shared memory.
//
// Why: the ADBench / GMM shapes (BASELINE.json C4 D = 6..64, C1 D = 2; the reference's own test fixtures, D = 20 / 32)
// are launch-latency bound as a chain: 8-43 kernels of a few microseconds each for 28..260 bytes per row (round 1:
// 0.4 % of the HBM roofline at D = 6).  Their weights are a few KB to a few hundred KB and sit in L2; per row tile each
// layer's packed fp32 weights are staged into shared memory once and used by all 64 rows.
//
//   * arithmetic: fp32 FFMA, precise tanhf / expf -- the <= 1e-4 tier; every precision request is served by it when the
//     stack is eligible (the tensor cores have nothing to win at these widths);
//   * deterministic: a row's log-det / log-density terms are summed by one half-warp in a fixed order -- no atomics;
//   * 256 threads = 16 row groups x 16 column lanes; a thread owns a 4 x 8 register tile (rows 4 ty..+3, columns
//     tx + 16 j), operands are read as float4 along K from padded rows (conflict-free quarter-warp accesses);
//   * shared memory is sized from the widest layer of the stack (22 KB for D = 6 -> several CTAs per SM hide the
//     barrier latency of the tiny layers; 200 KB for 128-wide hidden layers).
//
// Layout conventions are those of the packed descriptors (include/usflow_b200.h): activation row
// [a-part | pad | b-part at b_off], last conditioner layer packed as ONE tile [s(64) | t(64)] (affine) or [t(128)]
// (additive), final map in natural column order.
#include "usf_common.cuh"

namespace usf {
namespace {

constexpr int SS_ROWS = 64;
constexpr int SS_THREADS = 256;
constexpr int SS_MAX_OPS = 96;
constexpr int SS_MAXW = 128;

enum SsKind : int { SS_AFFINE = 0, SS_HIDDEN = 1, SS_COUPLING = 2, SS_FINAL = 3 };

struct SsOp {
  const float* W;      // (N, ldw) fp32, K-major rows
  const float* bias;   // (N)
  int N, K, ldw;
  int kind;
  int first;           // conditioner layer whose input is the activation row (columns [0, K)) instead of the hidden buffer
  int b_off, Db, affine;
  float clamp;
};

struct SsArgs {
  int n_ops;
  int d_in;            // input columns (D + context)
  int D;
  int lda;             // shared-memory row stride of the activation / hidden buffers (floats)
  int wbuf_floats;     // capacity of the weight staging buffer
  int inverse;
  int base_kind;       // -1: none
  int64_t B;
  const float* x;
  int64_t ldx;
  float* out_lp;
  float* out_y;
  int64_t ldy;
  float* out_ladj;
  float acc_init;
  const float* loc;
  const float* inv_scale;
  SsOp ops[SS_MAX_OPS];
};

__device__ __forceinline__ float half_warp_sum(float v) {
  // the 16 column lanes of one row group are 16 consecutive lanes of a warp
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

__global__ void __launch_bounds__(SS_THREADS) usf_small_stack_kernel(const __grid_constant__ SsArgs args) {
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  const int lda = args.lda;
  float* act[2] = {smem, smem + SS_ROWS * lda};
  float* hid[2] = {smem + 2 * SS_ROWS * lda, smem + 3 * SS_ROWS * lda};
  float* wbuf = smem + 4 * SS_ROWS * lda;
  float* bbuf = wbuf + args.wbuf_floats;
  float* racc = bbuf + SS_MAXW;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t n_tiles = (args.B + SS_ROWS - 1) / SS_ROWS;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * SS_ROWS;
    const int nrows = (int)((args.B - row0) < SS_ROWS ? (args.B - row0) : SS_ROWS);
    __syncthreads();   // the previous tile's last reads of the buffers are done
    {
      // x rows -> act[0] (columns beyond d_in up to the next multiple of 4 are zero: the first map reads K rounded up)
      const int d4 = (args.d_in + 3) & ~3;
      for (int i = tid; i < SS_ROWS * d4; i += SS_THREADS) {
        const int r = i / d4, c = i - r * d4;
        act[0][r * lda + c] = (r < nrows && c < args.d_in) ? args.x[(row0 + r) * args.ldx + c] : 0.f;
      }
      if (tid < SS_ROWS) racc[tid] = args.acc_init;
    }
    int cur = 0, hp = 0;
    for (int oi = 0; oi < args.n_ops; ++oi) {
      const SsOp& op = args.ops[oi];
      const int N = op.N, K = op.K;
      const int K4 = (K + 3) & ~3, ldw = K4 + 4, N16 = (N + 15) & ~15;
      __syncthreads();   // previous layer: output complete, its weights no longer read
      // stage this layer's weights (zero beyond N / K) and bias
      for (int i = tid; i < N16 * (K4 >> 2); i += SS_THREADS) {
        const int n = i / (K4 >> 2), k = (i - n * (K4 >> 2)) << 2;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < N) {
          const float* src = op.W + (size_t)n * op.ldw + k;
          if (k + 4 <= K && (op.ldw & 3) == 0) {
            w = *reinterpret_cast<const float4*>(src);
          } else {
            w.x = k < K ? src[0] : 0.f;
            w.y = k + 1 < K ? src[1] : 0.f;
            w.z = k + 2 < K ? src[2] : 0.f;
            w.w = k + 3 < K ? src[3] : 0.f;
          }
        }
        *reinterpret_cast<float4*>(wbuf + n * ldw + k) = w;
      }
      for (int i = tid; i < N16; i += SS_THREADS) bbuf[i] = (i < N && op.bias != nullptr) ? op.bias[i] : 0.f;
      __syncthreads();

      const float* in = (op.kind == SS_AFFINE || op.kind == SS_FINAL || op.first) ? act[cur] : hid[hp];
      const int jn = N16 >> 4;   // 16-column groups of this layer (<= 8), uniform over the CTA
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      const float* arow = in + (ty * 4) * lda;
      const float* wrow = wbuf + tx * ldw;
      for (int k = 0; k < K4; k += 4) {
        float4 a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(arow + i * lda + k);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < jn) {
            const float4 w = *reinterpret_cast<const float4*>(wrow + j * 16 * ldw + k);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc[i][j] = fmaf(a[i].x, w.x, acc[i][j]);
              acc[i][j] = fmaf(a[i].y, w.y, acc[i][j]);
              acc[i][j] = fmaf(a[i].z, w.z, acc[i][j]);
              acc[i][j] = fmaf(a[i].w, w.w, acc[i][j]);
            }
          }
        }
      }

      if (op.kind == SS_AFFINE || op.kind == SS_HIDDEN) {
        float* out = op.kind == SS_AFFINE ? act[cur ^ 1] : hid[op.first ? 0 : (hp ^ 1)];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < jn) {
            const int n = tx + 16 * j;
            const float b = bbuf[n];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float v = acc[i][j] + b;
              if (op.kind == SS_HIDDEN) v = fmaxf(v, 0.f);
              out[(ty * 4 + i) * lda + n] = v;
            }
          }
        }
        if (op.kind == SS_AFFINE) cur ^= 1;
        else hp = op.first ? 0 : (hp ^ 1);
      } else if (op.kind == SS_COUPLING) {
        // affine: columns tx + 16 j, j < 4 are s of coordinate c = tx + 16 j and j + 4 its t; additive: all 8 are t
        float* u_base = act[cur] + op.b_off;
        float lsum[4] = {0.f, 0.f, 0.f, 0.f};
        const int jc = op.affine ? 4 : 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < jc) {
            const int c = tx + 16 * j;
            if (c < op.Db) {
              const float bs = op.affine ? bbuf[c] : 0.f;
              const float bt = op.affine ? bbuf[64 + c] : bbuf[c];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float* up = u_base + (ty * 4 + i) * lda + c;
                const float u = *up;
                // (j + 4) & 7 keeps the index in range when this branch is compiled for the additive case
                const float t = (op.affine ? acc[i][(j + 4) & 7] : acc[i][j]) + bt;
                if (op.affine) {
                  const float ls = op.clamp * tanhf(acc[i][j] + bs);
                  lsum[i] += ls;
                  *up = args.inverse ? (u - t) * expf(-ls) : fmaf(u, expf(ls), t);
                } else {
                  *up = args.inverse ? u - t : u + t;
                }
              }
            }
          }
        }
        if (op.affine) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float s = half_warp_sum(lsum[i]);
            if (tx == 0) racc[ty * 4 + i] += args.inverse ? -s : s;   // one writer per row, fixed order: deterministic
          }
        }
        hp = 0;
      } else {  // SS_FINAL: natural column order, optional store, optional base log-density
        float lsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < jn) {
            const int n = tx + 16 * j;
            if (n < args.D) {
              const float b = bbuf[n];
              const float lc = args.base_kind >= 0 ? args.loc[n] : 0.f;
              const float is = args.base_kind >= 0 ? args.inv_scale[n] : 0.f;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r = ty * 4 + i;
                const float z = acc[i][j] + b;
                if (args.out_y != nullptr && r < nrows) args.out_y[(row0 + r) * args.ldy + n] = z;
                const float d = (z - lc) * is;
                lsum[i] += args.base_kind == 0 ? -0.5f * d * d : -fabsf(d);
              }
            }
          }
        }
        if (args.base_kind >= 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float s = half_warp_sum(lsum[i]);
            if (tx == 0) racc[ty * 4 + i] += s;
          }
        }
      }
    }
    __syncthreads();
    if (tid < nrows) {
      if (args.out_lp != nullptr) args.out_lp[row0 + tid] = racc[tid];
      if (args.out_ladj != nullptr) args.out_ladj[row0 + tid] = racc[tid];
    }
  }
}

bool small_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("USF_SMALL_STACK"); on = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return on == 1;
}

// Fills `a` from the descriptor; false when the stack is not one this kernel runs.
bool build_args(const usf_stack_desc* st, SsArgs* a) {
  if (st == nullptr || st->D <= 0 || st->D + st->ctx_dim > 64 || st->n_blocks < 0) return false;
  int n = 0, maxw = (st->D + st->ctx_dim + 3) & ~3, maxwb = 0;
  auto push = [&](const usf_linear_desc& L, int N, int kind, int first, const usf_block_desc* blk) -> bool {
    if (n >= SS_MAX_OPS || L.W == nullptr || N <= 0 || N > SS_MAXW || L.K <= 0 || L.K > SS_MAXW || L.ldw < L.K) return false;
    SsOp& o = a->ops[n++];
    o.W = L.W; o.bias = L.bias; o.N = N; o.K = L.K; o.ldw = L.ldw; o.kind = kind; o.first = first;
    o.b_off = blk ? blk->b_off : 0; o.Db = blk ? blk->Db : 0; o.affine = blk ? blk->affine : 0; o.clamp = blk ? blk->clamp : 0.f;
    const int n16 = (N + 15) & ~15, k4 = (L.K + 3) & ~3;
    if (n16 > maxw) maxw = n16;
    if (k4 > maxw) maxw = k4;
    if (n16 * (k4 + 4) > maxwb) maxwb = n16 * (k4 + 4);
    return true;
  };
  for (int b = 0; b < st->n_blocks; ++b) {
    const usf_block_desc& blk = st->blocks[b];
    if (blk.n_mlp < 1 || blk.n_mlp > USF_MAX_MLP) return false;
    if (!push(blk.G, blk.G.N, SS_AFFINE, 0, &blk)) return false;
    if (blk.b_off + blk.Db > blk.G.N || blk.Db <= 0 || blk.mlp[0].K > blk.b_off) return false;
    for (int l = 0; l < blk.n_mlp; ++l) {
      const bool last = l == blk.n_mlp - 1;
      // the coupling tile must be the single [s(64) | t(64)] / [t(128)] tile of the fp32 packing
      if (last && (blk.mlp[l].N != 128 || blk.C != (blk.affine ? 64 : 128) || blk.Db > blk.C)) return false;
      if (!push(blk.mlp[l], blk.mlp[l].N, last ? SS_COUPLING : SS_HIDDEN, l == 0 ? 1 : 0, &blk)) return false;
    }
  }
  if (st->G_final.N < st->D) return false;
  if (!push(st->G_final, st->D, SS_FINAL, 0, nullptr)) return false;
  a->n_ops = n;
  a->lda = maxw + 4;
  a->wbuf_floats = (maxwb + 3) & ~3;
  return true;
}

size_t smem_bytes(const SsArgs& a) {
  return sizeof(float) * ((size_t)4 * SS_ROWS * a.lda + a.wbuf_floats + SS_MAXW + SS_ROWS);
}

}  // namespace

bool small_stack_supported(const usf_stack_desc* st, int precision) {
  if (precision != USF_PREC_FP32 || !small_enabled()) return false;
  static thread_local SsArgs probe;
  return build_args(st, &probe) && smem_bytes(probe) <= 220 * 1024;
}

int small_stack_run(const usf_stack_desc* st, const float* x, int64_t ldx, int64_t B, float* out_logprob, float* out_y,
                    int64_t ldy, float* out_ladj, cudaStream_t stream) {
  static thread_local SsArgs a;
  USF_CHECK_ARG(build_args(st, &a), "small_stack_run: unsupported stack");
  a.d_in = st->D + st->ctx_dim;
  a.D = st->D;
  a.inverse = st->inverse;
  a.base_kind = out_logprob != nullptr ? st->base_kind : -1;
  a.B = B;
  a.x = x;
  a.ldx = ldx;
  a.out_lp = out_logprob;
  a.out_y = out_y;
  a.ldy = ldy;
  a.out_ladj = out_ladj;
  a.acc_init = out_logprob != nullptr ? st->const_term : 0.f;
  a.loc = st->loc;
  a.inv_scale = st->inv_scale;
  const size_t smem = smem_bytes(a);
  static thread_local size_t attr_bytes = 0;
  if (smem > 48 * 1024 && smem > attr_bytes) {
    USF_CUDA(cudaFuncSetAttribute(usf_small_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    attr_bytes = 220 * 1024;
  }
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, usf_small_stack_kernel, SS_THREADS, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  const int64_t tiles = ceil_div(B, SS_ROWS);
  int64_t grid = (int64_t)num_sms() * per_sm;
  if (grid > tiles) grid = tiles;
  usf_small_stack_kernel<<<(unsigned)grid, SS_THREADS, smem, stream>>>(a);
  USF_LAUNCH_CHECK("usf_small_stack_kernel");
  return USF_OK;
}

}  // namespace usf
