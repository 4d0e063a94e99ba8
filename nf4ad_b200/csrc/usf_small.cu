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
//   * latency: the next layer's packed weights (and the next tile's input rows) travel into the second half of a
//     double-buffered staging area with cp.async while the current layer computes -- one __syncthreads per layer, no
//     global-memory latency on the layer-to-layer critical path; only the output columns a coupling really uses
//     (round_up(Db, 16) of s and of t out of the packed 64 + 64) are staged and multiplied;
//   * shared memory is sized from the widest layer of the stack (a few KB for D = 6 -> two CTAs per SM; ~200 KB for
//     128-wide hidden layers).
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
  const float* W;      // (N, ldw) fp32, K-major rows (ldw a multiple of 8, pad columns zero: usf_pack_matrix)
  const float* bias;   // (N)
  int N, K, ldw;       // N: columns this kernel computes (a coupling: 2 x Db16 (affine) / Db16 (additive))
  int kind;
  int first;           // conditioner layer whose input is the activation row (columns [0, K)) instead of the hidden buffer
  int b_off, Db, affine;
  int t_row;           // affine coupling: first packed row of the shift parameters (64)
  int woff;            // resident mode: offset (floats) of this layer's staged weights, the bias follows them
  float clamp;
};

struct SsArgs {
  int n_ops;
  int d_in;            // input columns (D + context)
  int D;
  int lda;             // shared-memory row stride of the activation buffers (floats)
  int ldh;             // ... of the hidden buffers
  int ldxs;            // ... of the input staging buffers
  int wbuf_floats;     // capacity of ONE weight staging buffer (resident mode: of the whole weight area)
  int resident;        // every layer's weights fit in shared memory together: staged once per CTA, not once per row tile
  int inverse;
  int base_kind;       // -1: none
  int x_vec;           // input rows can be fetched in 16-byte pieces
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
__device__ __forceinline__ uint32_t s_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Requests one layer's weights + bias into a staging buffer (asynchronously; rows beyond N are zeroed with plain stores).
// A coupling layer only brings the rows it uses: [0, Db16) of s and [t_row, t_row + Db16) of t, packed back to back.
__device__ __forceinline__ void stage_op(const SsOp& op, float* wbuf, float* bbuf, int tid) {
  const int K4 = (op.K + 3) & ~3, ldw = K4 + 4, N = op.N, N16 = (N + 15) & ~15, kq = K4 >> 2;
  const int half = op.kind == SS_COUPLING && op.affine ? (N >> 1) : N;   // rows of the first (or only) source range
  for (int i = tid; i < N16 * kq; i += SS_THREADS) {
    const int n = i / kq, k = (i - n * kq) << 2;
    float* dst = wbuf + n * ldw + k;
    if (n < N) {
      const int src_row = n < half ? n : op.t_row + (n - half);
      cp_async16(dst, op.W + (size_t)src_row * op.ldw + k);
    } else {
      *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (int i = tid; i < N16; i += SS_THREADS) {
    if (i < N && op.bias != nullptr) cp_async4(bbuf + i, op.bias + (i < half ? i : op.t_row + (i - half)));
    else bbuf[i] = 0.f;
  }
}

// Requests the input rows of one tile into a staging buffer (columns [d_in, d4) and rows beyond the batch are zeroed).
__device__ __forceinline__ void stage_x(const SsArgs& args, int64_t tile, float* xbuf, int tid) {
  const int64_t row0 = tile * SS_ROWS;
  const int nrows = (int)((args.B - row0) < SS_ROWS ? (args.B - row0) : SS_ROWS);
  const int d4 = (args.d_in + 3) & ~3;
  if (args.x_vec) {
    const int q = d4 >> 2;
    for (int i = tid; i < SS_ROWS * q; i += SS_THREADS) {
      const int r = i / q, c = (i - r * q) << 2;
      float* dst = xbuf + r * args.ldxs + c;
      if (r < nrows) cp_async16(dst, args.x + (row0 + r) * args.ldx + c);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    for (int i = tid; i < SS_ROWS * d4; i += SS_THREADS) {
      const int r = i / d4, c = i - r * d4;
      float* dst = xbuf + r * args.ldxs + c;
      if (r < nrows && c < args.d_in) cp_async4(dst, args.x + (row0 + r) * args.ldx + c);
      else *dst = 0.f;
    }
  }
}

struct SsLayerCtx {
  const float* in; int ldi;
  const float* wbuf; const float* bbuf; int ldw, K4, jn;
  float* actc;        // current activation buffer (coupling: updated in place)
  float* act_out;     // the other activation buffer (affine map output)
  float* hid_out;     // hidden buffer this layer writes
  int lda, ldh;
  float* racc; const float* locs; const float* iscs;
  int64_t row0; int nrows, tx, ty;
};

// One layer on this thread's 4 x (16 JN) register tile: GEMM against the staged weights, then the layer's epilogue.
template <int JN>
__device__ __forceinline__ void ss_layer(const SsArgs& args, const SsOp& op, const SsLayerCtx& cx) {
  const int jn = cx.jn, tx = cx.tx, ty = cx.ty, lda = cx.lda;
  const float* bbuf = cx.bbuf;
  float acc[4][JN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < JN; ++j) acc[i][j] = 0.f;
  const float* arow = cx.in + (ty * 4) * cx.ldi;
  const float* wrow = cx.wbuf + tx * cx.ldw;
  for (int k = 0; k < cx.K4; k += 4) {
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(arow + i * cx.ldi + k);
#pragma unroll
    for (int j = 0; j < JN; ++j) {
      if (JN <= 2 || j < jn) {
        const float4 w = *reinterpret_cast<const float4*>(wrow + j * 16 * cx.ldw + k);
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
    float* out = op.kind == SS_AFFINE ? cx.act_out : cx.hid_out;
    const int ldo = op.kind == SS_AFFINE ? lda : cx.ldh;
#pragma unroll
    for (int j = 0; j < JN; ++j) {
      if (j < jn) {
        const int n = tx + 16 * j;
        const float b = bbuf[n];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v = acc[i][j] + b;
          if (op.kind == SS_HIDDEN) v = fmaxf(v, 0.f);
          out[(ty * 4 + i) * ldo + n] = v;
        }
      }
    }
  } else if (op.kind == SS_COUPLING) {
    // staged columns: [s of coordinates 0..Db16) | t of the same] (affine) or [t] (additive); this thread's
    // coordinates are c = tx + 16 j
    float* u_base = cx.actc + op.b_off;
    float lsum[4] = {0.f, 0.f, 0.f, 0.f};
    const int jh = op.affine ? (jn >> 1) : jn;    // column groups of one parameter
    constexpr int JH = JN > 4 ? 4 : JN;           // (at most 64 coordinates = 4 groups)
#pragma unroll
    for (int j = 0; j < JH; ++j) {
      if (j < jh) {
        const int c = tx + 16 * j;
        if (c < op.Db) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float* up = u_base + (ty * 4 + i) * lda + c;
            const float u = *up;
            if (op.affine) {
              // the shift of coordinate group j is accumulator group j + jh (jh = 1 .. 4); indices kept in range
              const float tj = jh == 1 ? acc[i][(j + 1) % JN]
                                       : (jh == 2 ? acc[i][(j + 2) % JN] : (jh == 3 ? acc[i][(j + 3) % JN] : acc[i][(j + 4) % JN]));
              const float t = tj + bbuf[jh * 16 + c];
              const float ls = op.clamp * tanhf(acc[i][j] + bbuf[c]);
              lsum[i] += ls;
              *up = args.inverse ? (u - t) * expf(-ls) : fmaf(u, expf(ls), t);
            } else {
              const float t = acc[i][j] + bbuf[c];
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
        if (tx == 0) cx.racc[ty * 4 + i] += args.inverse ? -s : s;   // one writer per row, fixed order: deterministic
      }
    }
  } else {  // SS_FINAL: natural column order, optional store, optional base log-density
    float lsum[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr int JF = JN > 4 ? 4 : JN;           // D <= 64
#pragma unroll
    for (int j = 0; j < JF; ++j) {
      if (j < jn) {
        const int n = tx + 16 * j;
        if (n < args.D) {
          const float b = bbuf[n], lc = cx.locs[n], is = cx.iscs[n];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = ty * 4 + i;
            const float z = acc[i][j] + b;
            if (args.out_y != nullptr && r < cx.nrows) args.out_y[(cx.row0 + r) * args.ldy + n] = z;
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
        if (tx == 0) cx.racc[ty * 4 + i] += s;
      }
    }
  }
}

__global__ void __launch_bounds__(SS_THREADS, 2) usf_small_stack_kernel(const __grid_constant__ SsArgs args) {
  extern __shared__ float4 smem4[];
  float* smem = reinterpret_cast<float*>(smem4);
  const int lda = args.lda, ldh = args.ldh;
  float* const act0 = smem;
  float* const act1 = act0 + SS_ROWS * lda;
  float* const hid0 = act1 + SS_ROWS * lda;
  float* const hid1 = hid0 + SS_ROWS * ldh;
  float* const xb0 = hid1 + SS_ROWS * ldh;
  float* const xb1 = xb0 + SS_ROWS * args.ldxs;
  float* const wb0 = xb1 + SS_ROWS * args.ldxs;
  float* const wb1 = wb0 + (args.resident ? 0 : args.wbuf_floats);
  float* const bb0 = wb0 + (args.resident ? 1 : 2) * args.wbuf_floats;
  float* const bb1 = bb0 + SS_MAXW;
  float* const racc = bb1 + SS_MAXW;
  float* const locs = racc + SS_ROWS;
  float* const iscs = locs + 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t n_tiles = (args.B + SS_ROWS - 1) / SS_ROWS;
  if ((int64_t)blockIdx.x >= n_tiles) return;

  if (tid < 64) {
    locs[tid] = (args.base_kind >= 0 && tid < args.D) ? args.loc[tid] : 0.f;
    iscs[tid] = (args.base_kind >= 0 && tid < args.D) ? args.inv_scale[tid] : 0.f;
  }
  // in flight before the first layer: this CTA's first input tile and the first layer's weights -- or, when they all
  // fit, EVERY layer's weights, which then stay for all the row tiles of this CTA
  stage_x(args, blockIdx.x, xb0, tid);
  if (args.resident) {
    for (int oi = 0; oi < args.n_ops; ++oi) {
      const SsOp& op = args.ops[oi];
      const int n16 = (op.N + 15) & ~15, ldw = ((op.K + 3) & ~3) + 4;
      stage_op(op, wb0 + op.woff, wb0 + op.woff + n16 * ldw, tid);
    }
  } else {
    stage_op(args.ops[0], wb0, bb0, tid);
  }
  cp_async_commit();

  int sbuf = 0;   // staging buffer holding the CURRENT layer's weights
  int xsel = 0;   // staging buffer holding the CURRENT tile's input rows
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * SS_ROWS;
    const int nrows = (int)((args.B - row0) < SS_ROWS ? (args.B - row0) : SS_ROWS);
    if (tid < SS_ROWS) racc[tid] = args.acc_init;   // (its previous values were stored by these same threads)
    int cur = 0, hp = 0;
    for (int oi = 0; oi < args.n_ops; ++oi) {
      const SsOp& op = args.ops[oi];
      cp_async_wait_all();
      __syncthreads();   // this layer's weights (and, for the first layer, the tile's rows) have landed; the previous
                         // layer's output is complete and its weights are no longer read
      {
        // next in line: the next layer of this tile, or the first layer + the input rows of this CTA's next tile
        const bool last_op = oi + 1 == args.n_ops;
        const int64_t next_tile = tile + gridDim.x;
        if (!last_op) {
          if (!args.resident) stage_op(args.ops[oi + 1], sbuf ? wb0 : wb1, sbuf ? bb0 : bb1, tid);
        } else if (next_tile < n_tiles) {
          if (!args.resident) stage_op(args.ops[0], sbuf ? wb0 : wb1, sbuf ? bb0 : bb1, tid);
          stage_x(args, next_tile, xsel ? xb0 : xb1, tid);
        }
        cp_async_commit();
      }
      float* const actc = cur ? act1 : act0;
      const int N = op.N, K4 = (op.K + 3) & ~3, ldw = K4 + 4, N16 = (N + 15) & ~15;
      const float* wbuf = args.resident ? wb0 + op.woff : (sbuf ? wb1 : wb0);
      const float* bbuf = args.resident ? wbuf + N16 * ldw : (sbuf ? bb1 : bb0);
      const float* in;
      int ldi = lda;
      if (oi == 0) { in = xsel ? xb1 : xb0; ldi = args.ldxs; }
      else if (op.kind == SS_AFFINE || op.kind == SS_FINAL || op.first) in = actc;
      else { in = hp ? hid1 : hid0; ldi = ldh; }
      // the layer itself, compiled for 1 / 2 / 4 / 8 column groups of 16 (a 6-wide layer must not pay for the 128-wide
      // register tile: with one body predicated over 8 groups the tiny stacks were instruction-bound)
      const int jn = N16 >> 4;
      SsLayerCtx cx{in, ldi, wbuf, bbuf, ldw, K4, jn, actc, cur ? act0 : act1, (op.first || hp) ? hid0 : hid1, lda, ldh,
                    racc, locs, iscs, row0, nrows, tx, ty};
      if (jn <= 1) ss_layer<1>(args, op, cx);
      else if (jn == 2) ss_layer<2>(args, op, cx);
      else if (jn <= 4) ss_layer<4>(args, op, cx);
      else ss_layer<8>(args, op, cx);
      if (op.kind == SS_AFFINE) cur ^= 1;
      else if (op.kind == SS_HIDDEN) hp = (op.first || hp) ? 0 : 1;
      else if (op.kind == SS_COUPLING) hp = 0;
      sbuf ^= 1;
    }
    xsel ^= 1;
    __syncthreads();
    if (tid < nrows) {
      if (args.out_lp != nullptr) args.out_lp[row0 + tid] = racc[tid];
      if (args.out_ladj != nullptr) args.out_ladj[row0 + tid] = racc[tid];
    }
  }
  cp_async_wait_all();
}

bool small_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("USF_SMALL_STACK"); on = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return on == 1;
}

// Fills `a` from the descriptor; false when the stack is not one this kernel runs.
bool build_args(const usf_stack_desc* st, SsArgs* a) {
  if (st == nullptr || st->D <= 0 || st->D + st->ctx_dim > 64 || st->n_blocks < 0) return false;
  int n = 0, maxa = 16, maxh = 16, maxwb = 0, total = 0;
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  // stored: 0 = the output stays in registers (coupling, final), 1 = activation buffer, 2 = hidden buffer;
  // in_hidden: the layer reads the hidden buffer
  auto push = [&](const usf_linear_desc& L, int N, int kind, int first, const usf_block_desc* blk, int stored, bool in_hidden) -> bool {
    if (n >= SS_MAX_OPS || L.W == nullptr || !aligned(L.W) || (L.ldw & 3) != 0 || N <= 0 || N > SS_MAXW || L.K <= 0 ||
        L.K > SS_MAXW || L.ldw < ((L.K + 3) & ~3))
      return false;
    SsOp& o = a->ops[n++];
    o.W = L.W; o.bias = L.bias; o.N = N; o.K = L.K; o.ldw = L.ldw; o.kind = kind; o.first = first;
    o.b_off = blk ? blk->b_off : 0; o.Db = blk ? blk->Db : 0; o.affine = blk ? blk->affine : 0; o.clamp = blk ? blk->clamp : 0.f;
    o.t_row = 64;
    const int n16 = (N + 15) & ~15, k4 = (L.K + 3) & ~3;
    if (stored == 1 && n16 > maxa) maxa = n16;
    if (stored == 2 && n16 > maxh) maxh = n16;
    if (in_hidden) { if (k4 > maxh) maxh = k4; } else { if (k4 > maxa) maxa = k4; }
    if (n16 * (k4 + 4) > maxwb) maxwb = n16 * (k4 + 4);
    o.woff = total;
    total += n16 * (k4 + 4) + n16;
    return true;
  };
  for (int b = 0; b < st->n_blocks; ++b) {
    const usf_block_desc& blk = st->blocks[b];
    if (blk.n_mlp < 1 || blk.n_mlp > USF_MAX_MLP) return false;
    if (!push(blk.G, blk.G.N, SS_AFFINE, 0, &blk, 1, false)) return false;
    if (blk.b_off + blk.Db > blk.G.N || blk.Db <= 0 || blk.Db > 64 || blk.mlp[0].K > blk.b_off) return false;
    for (int l = 0; l < blk.n_mlp; ++l) {
      const bool last = l == blk.n_mlp - 1;
      if (!last) {
        if (!push(blk.mlp[l], blk.mlp[l].N, SS_HIDDEN, l == 0 ? 1 : 0, &blk, 2, l > 0)) return false;
        continue;
      }
      // the coupling tile must be the single [s(64) | t(64)] / [t(128)] tile of the fp32 packing; only the
      // round_up(Db, 16) columns of each parameter that carry coordinates are computed
      if (blk.mlp[l].N != 128 || blk.C != (blk.affine ? 64 : 128)) return false;
      const int db16 = (blk.Db + 15) & ~15;
      if (!push(blk.mlp[l], blk.affine ? 2 * db16 : db16, SS_COUPLING, l == 0 ? 1 : 0, &blk, 0, l > 0)) return false;
    }
  }
  if (st->G_final.N < st->D) return false;
  if (!push(st->G_final, st->D, SS_FINAL, 0, nullptr, 0, false)) return false;
  a->n_ops = n;
  a->lda = maxa + 4;
  a->ldh = maxh + 4;
  a->ldxs = ((st->D + st->ctx_dim + 3) & ~3) + 4;
  // all layers resident when they fit beside the row buffers (then no weight travels per row tile)
  a->resident = (size_t)total * sizeof(float) <= 96 * 1024 ? 1 : 0;
  a->wbuf_floats = a->resident ? ((total + 3) & ~3) : ((maxwb + 3) & ~3);
  return true;
}

size_t smem_bytes(const SsArgs& a) {
  return sizeof(float) * ((size_t)2 * SS_ROWS * a.lda + (size_t)2 * SS_ROWS * a.ldh + (size_t)2 * SS_ROWS * a.ldxs +
                          (size_t)(a.resident ? 1 : 2) * a.wbuf_floats + 2 * SS_MAXW + SS_ROWS + 128);
}

}  // namespace

bool small_stack_supported(const usf_stack_desc* st, int precision) {
  if (precision != USF_PREC_FP32 || !small_enabled()) return false;
  static thread_local SsArgs probe;
  return build_args(st, &probe) && smem_bytes(probe) <= 227 * 1024;
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
  a.x_vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ldx & 3) == 0 && (a.d_in & 3) == 0) ? 1 : 0;
  a.out_lp = out_logprob;
  a.out_y = out_y;
  a.ldy = ldy;
  a.out_ladj = out_ladj;
  a.acc_init = out_logprob != nullptr ? st->const_term : 0.f;
  a.loc = st->loc;
  a.inv_scale = st->inv_scale;
  const size_t smem = smem_bytes(a);
  static thread_local bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    USF_CUDA(cudaFuncSetAttribute(usf_small_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    attr_set = true;
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
