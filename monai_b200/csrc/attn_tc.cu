// Windowed multi-head self-attention on tcgen05 tensor cores.
//
// Reference: WindowAttention.forward (monai/networks/nets/swin_unetr.py:509-532): per window of n <= 343 tokens and head
// (head_dim 16):  softmax(q k^T * scale + relative_position_bias[:n,:n] + shift_mask) v.
//
// One persistent CTA per SM works on tiles = (window, head, 128-query row tile).  Per tile
//   S[128 x n_pad] = Q K^T                      1 tcgen05.mma per key half   (SS form, K = 16 = the whole head)
//                  + I[128 x 128] * B'          8 tcgen05.mma per key half   (TS form: A = identity held in TMEM)
// where B'[i][j] = log2(e) * (bias[i][j] + mask[i][j]) is an fp16 B operand that stays RESIDENT in shared memory: it depends
// on (head, row tile, mask type) only, so tiles are scheduled (mask type, head, row tile)-major and a CTA reloads it a
// handful of times per launch.  Adding the bias with the tensor core (1.0 * fp16 value into the fp32 accumulator: exact)
// removes the per-element table lookup and mask test that bounded the mma.sync kernel (swin.cu) -- with head_dim 16 the
// tensor pipe is otherwise idle.  Padded keys carry B' = -30000 (P = 0).  Scores are in log2 units: the caller folds
// scale * log2(e) into the q rows of the qkv projection.
//   softmax: the two key halves of a tile are INDEPENDENT pipelines (flash-attention style): each half has its own exact row
//            maximum m_h (pass 1 over TMEM), its own P_h = ex2(S_h - m_h) -- fp32 MUFU, packed to fp16 pairs and stored with
//            tcgen05.st over the TMEM columns of S_h that the thread has already read (pass 2) -- and its own accumulator
//   O_h[128 x 32] = P_h [V | 1 | 0]_h           n_pad/32 tcgen05.mma in TS form (A = P_h from TMEM), V read in place as an
//            MN-major B operand (NC8 rows are 16-byte vectors of 8 dims); the ones column accumulates the row sums l_h in fp32;
//   epilogue: O = (a0 O0 + a1 O1) / (a0 l0 + a1 l1), a_h = 2^(m_h - max(m0, m1))  -> fp16 NC8.
// Because no maximum is shared between the halves, the tensor pipe computes PV_h(i) and S_h(i+1) for one half while the softmax
// warps of the OTHER half exponentiate: TMEM has room for one S tile only (2 x 176 + 2 x 32 + 64 identity columns), and the
// earlier single-maximum version (both halves needed before any exponential) left every softmax warp idle for the ~1 000 cycles
// the S MMAs of the next tile take -- 8 300 cycles per tile against a MUFU floor of 2 816 (ncu: 33 % issue-active, MUFU 38 %,
// tensor 25 %).
// History of the P path (stage-1 shape, batch 8, ms; profiles/r02_attention_phase_db.jsonl, timelines read with
// profiles/read_attn_trace.py from B200_ATTN_TRACE dumps): P through a shared-memory A image (SS MMAs, N = 32: bound by the 4 KB
// A read, 59 cycles each) 0.473 -> P in TMEM (TS MMAs) 0.445 -> Q / K / V double-buffered in the 90 KB the P image freed
// (the refill of the single buffers sat on the critical path of each half: softmax -> PV -> S) 0.418 -> S1 held half a period
// behind S0 so that the exponential passes of the two halves do not share the MUFU 0.385 (shifted windows 0.533 -> 0.406).
// Q, K, V tiles are 1-D bulk copies of NC8 rows (contiguous per 8-channel chunk); the producer runs up to two tiles ahead.
//
// Warp roles (576 threads): warp 0 = copy producer, warp 1 = TMEM owner + MMA issuer, warps 2-9 = softmax of key half 0,
// warps 10-17 = softmax of key half 1 (+ the epilogue); the two threads of a (query row, half) split its 16-column chunks and
// exchange their maxima through shared memory and a 64-thread named barrier.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "tc05.cuh"
#include "../../include/monai_b200.h"

namespace b200 {

constexpr int kAtNPadMax = 352;                       // keys per window, padded (n <= 343 -> 352)
constexpr int kAtKChunk = kAtNPadMax * 16;            // bytes of one 8-dim chunk of K / V in shared memory
constexpr int kAtBiasBytes = 16 * kAtNPadMax * 16;    // 16 chunks of 8 query rows
constexpr int kAtIdBytes = 16 * 2048;                 // identity operand image: [k chunk of 8][128 rows][16 B]
constexpr int kAtQBytes = 2 * 2048, kAtKBytes = 2 * kAtKChunk, kAtVBytes = 4 * kAtKChunk;   // one buffer of each (two of each are kept)
constexpr int kAtColS1 = 176, kAtColO0 = 352, kAtColI = 384, kAtColO1 = 448;   // TMEM columns: S half 0 at 0, S half 1, O of half 0 (32), identity (64), O of half 1 (32)
constexpr int kAtSmem = kAtBiasBytes + kAtIdBytes + 2 * (kAtQBytes + kAtKBytes + kAtVBytes) + 8 * 128 * 4 + 256 + 128;
constexpr float kAtPadBias = -30000.f;
constexpr int kAtPhaseDefault = 175;                  // cycles per 32 padded keys (1 925 for 352 keys ~ half a tile period); see B200_ATTN_PHASE

struct AttnTcParams {
  const __half* qkv; __half* out; const __half* bias; const int32_t* sched;
  int N, C8, heads, nW, n, n_pad, nrt, ntypes;
  int phase_delay;      // cycles the S MMAs of key half 1 are held behind those of half 0 (0 = none): keeps the two softmax pipelines in anti-phase
  long long* trace;     // debug timeline (B200_ATTN_TRACE): clock64 per (tile, role, event) of CTA 0, else null
};

struct AttnTile { int ty, h, rt, b, w; };

// flattened tile index -> (mask type, head, row tile, batch item, window); order: type, (head, row tile), batch, window
__device__ __forceinline__ AttnTile attn_decode(const AttnTcParams& p, long long f) {
  const int32_t* cnt = p.sched;
  const int32_t* start = p.sched + 8;
  const int32_t* win = p.sched + 16;
  AttnTile t;
  t.ty = 0;
  const long long hr_n = (long long)p.heads * p.nrt;
  for (;;) {
    const long long blk = (long long)cnt[t.ty] * p.N * hr_n;
    if (f < blk || t.ty + 1 >= p.ntypes) break;
    f -= blk; ++t.ty;
  }
  const int c = cnt[t.ty];
  const long long per = (long long)c * p.N;
  const int hr = (int)(f / per);
  const int l2 = (int)(f % per);
  t.h = hr / p.nrt; t.rt = hr % p.nrt;
  t.b = l2 / c;
  t.w = win[start[t.ty] + l2 % c];
  return t;
}

// 2^(a - m), 2^(b - m) as packed fp16.  (ex2.approx.f16x2 is split by ptxas into one MUFU per half plus a PRMT, so the fp32
// MUFU form costs the same MUFU slots, one instruction less, and keeps the exponent argument in fp32.)
__device__ __forceinline__ uint32_t exp2_pack(float a, float b, float m) {
  float ea, eb;
  // volatile: keeps the exponentials BEHIND the tcgen05.ld of the next chunk in program order (the prefetch must be issued first)
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(a - m));
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(b - m));
  const __half2 h = __floats2half2_rn(ea, eb);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// running maximum of 16 fp32 TMEM values (FMNMX3: two values per instruction)
__device__ __forceinline__ float max16(const uint32_t (&v)[16], float m) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) asm("max.f32 %0, %0, %1, %2;" : "+f"(m) : "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])));
  return m;
}

constexpr int kAtThreads = 64 + 512;        // producer, MMA issuer, 16 softmax warps (2 key halves x 2 threads per query row)

template <int NPAD, bool TRACE = false>
__global__ void __launch_bounds__(kAtThreads, 1) window_attention_tc_kernel(AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = tc::align_smem128(smem_raw);   // keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* s_bias = smem;
  uint8_t* s_id = s_bias + kAtBiasBytes;              // identity operand image, copied to TMEM once
  uint8_t* s_q = s_id + kAtIdBytes;                   // [2 buffers][2 chunks][128 rows][16 B]
  uint8_t* s_k = s_q + 2 * kAtQBytes;                 // [2 buffers][2 chunks][n_pad keys][16 B]
  uint8_t* s_v = s_k + 2 * kAtKBytes;                 // [2 buffers][4 chunk slots: V dims 0-7, 8-15, ones column, zeros]
  float* s_max = reinterpret_cast<float*>(s_v + 2 * kAtVBytes);   // [half][sub][128]: maxima exchanged by the two threads of a (row, half)
  float* s_hmax = s_max + 4 * 128;                                // [tile parity][half][128]: half maxima for the epilogue
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_hmax + 4 * 128);
  uint64_t* qk_full = bars + 0;     // [2]  Q / K (+ bias) of a tile have landed in buffer b
  uint64_t* qk_empty = bars + 2;    // [2]  the S MMAs that read buffer b are done
  uint64_t* v_full = bars + 4;      // [2]
  uint64_t* v_empty = bars + 6;     // [2]  the PV MMAs that read V buffer b are done
  uint64_t* s_full = bars + 8;      // [2]  per key half
  uint64_t* s_empty = bars + 10;    // [2]
  uint64_t* p_full = bars + 12;     // [2]
  uint64_t* pv_done = bars + 14;    // [2]
  uint64_t* o_empty = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto mark = [&](long long it, int role, int ev) {
    if (TRACE && blockIdx.x == 0 && lane == 0 && it < 64) p.trace[(it * 3 + role) * 8 + ev] = clock64();
  };
  // NPAD (keys per window, padded to a multiple of 32) is a template parameter: the softmax passes are straight-line code
  constexpr int n_pad = NPAD, NH = NPAD / 2;
  constexpr int nchunk = NH / 16;                    // 16-key chunks of a half
  constexpr int kCMax = (nchunk + 1) / 2;            // chunks of the larger of the two per-thread shares
  const int n = p.n;
  const long long T = (long long)p.nW * n;
  const long long total = (long long)p.N * p.nW * p.heads * p.nrt;
  const long long lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    tc::mbar_init(o_empty, 4);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&qk_full[i], 1); tc::mbar_init(&qk_empty[i], 1); tc::mbar_init(&v_full[i], 1); tc::mbar_init(&v_empty[i], 1);
      tc::mbar_init(&s_full[i], 1); tc::mbar_init(&s_empty[i], 8); tc::mbar_init(&p_full[i], 8); tc::mbar_init(&pv_done[i], 1);
    }
    tc::fence_barrier_init();
  }
  // zero the identity image and Q / K / V (rows the bulk copies never write must be finite)
  {
    const uint4 z = make_uint4(0, 0, 0, 0);
    uint4* zq = reinterpret_cast<uint4*>(s_id);
    const int nz = (kAtIdBytes + 2 * kAtQBytes + 2 * kAtKBytes + 2 * kAtVBytes) / 16;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) zq[i] = z;
  }
  __syncthreads();
  {
    for (int j = threadIdx.x; j < 2 * kAtNPadMax; j += blockDim.x) {      // the ones column of both V buffers
      __half* ones = reinterpret_cast<__half*>(s_v + (j / kAtNPadMax) * kAtVBytes + 2 * kAtKChunk);
      ones[(j % kAtNPadMax) * 8] = __float2half_rn(1.f);
    }
    __half* id = reinterpret_cast<__half*>(s_id);   // [k chunk of 8][row][8]: element (row r, k = r) = 1
    for (int r = threadIdx.x; r < 128; r += blockDim.x) id[((r >> 3) * 128 + r) * 8 + (r & 7)] = __float2half_rn(1.f);
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== copy producer: runs up to two tiles ahead of the MMAs =====================
    if (lane == 0) {
      int last_combo = -1;
      int it = 0;
      for (long long f = lo; f < hi; ++f, ++it) {
        const AttnTile t = attn_decode(p, f);
        const int combo = (t.ty * p.heads + t.h) * p.nrt + t.rt;
        const int rows = min(128, n - t.rt * 128);
        const int b = it & 1;
        const uint32_t use = (uint32_t)((it >> 1) & 1);
        tc::mbar_wait(&qk_empty[b], use ^ 1u);                  // the S MMAs that read this buffer two tiles ago are done
        uint32_t bias_bytes = 0u;
        if (combo != last_combo) {
          // the bias image is single-buffered: the S MMAs of the PREVIOUS tile (the other Q / K buffer) still read the old one
          if (it > 0) tc::mbar_wait(&qk_empty[b ^ 1], (uint32_t)(((it - 1) >> 1) & 1));
          bias_bytes = (uint32_t)(16 * n_pad * 16);
        }
        tc::mbar_arrive_expect_tx(&qk_full[b], bias_bytes + 2u * rows * 16u + 2u * n * 16u);
        if (bias_bytes) tc::bulk_load(s_bias, p.bias + (long long)combo * (16 * n_pad * 8), bias_bytes, &qk_full[b]);
        last_combo = combo;
        const __half* base = p.qkv + (long long)t.b * (3 * p.C8) * T * 8;
        const long long row0 = (long long)t.w * n;
        for (int c = 0; c < 2; ++c) {
          tc::bulk_load(s_q + b * kAtQBytes + c * 2048, base + ((long long)(2 * t.h + c) * T + row0 + t.rt * 128) * 8, rows * 16, &qk_full[b]);
          tc::bulk_load(s_k + b * kAtKBytes + c * kAtKChunk, base + ((long long)(p.C8 + 2 * t.h + c) * T + row0) * 8, n * 16, &qk_full[b]);
        }
        tc::mbar_wait(&v_empty[b], use ^ 1u);                   // the PV MMAs that read this V buffer two tiles ago are done
        tc::mbar_arrive_expect_tx(&v_full[b], 2u * n * 16u);
        for (int c = 0; c < 2; ++c)
          tc::bulk_load(s_v + b * kAtVBytes + c * kAtKChunk, base + ((long long)(2 * p.C8 + 2 * t.h + c) * T + row0) * 8, n * 16, &v_full[b]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = tc::elect_one();
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc_s = tc::make_idesc_f16(128, NH);
    const uint32_t idesc_pv = tc::make_idesc_f16(128, 32) | (1u << 16);   // B (= V) is MN-major
    const uint32_t q_a = tc::smem_u32(s_q), k_a = tc::smem_u32(s_k), v_a = tc::smem_u32(s_v), b_a = tc::smem_u32(s_bias), i_a = tc::smem_u32(s_id);
    // identity -> TMEM (8 K-steps of 16: 8 columns each)
    if (leader)
      for (int s = 0; s < 8; ++s) tc::tmem_cp_128x256b(tm + kAtColI + 8 * s, tc::make_desc_kmajor_noswz(i_a + s * 4096, 2048, 128));
    __syncwarp();
    // S = Q K^T + I B' for one key half of the tile whose Q / K sit in buffer b
    auto issue_s = [&](int hf, int b) {
      const uint32_t ts = tm + hf * kAtColS1;
      const uint64_t qd = tc::make_desc_kmajor_noswz(q_a + b * kAtQBytes, 2048, 128);
      const uint64_t kd = tc::make_desc_kmajor_noswz(k_a + b * kAtKBytes + hf * NH * 16, kAtKChunk, 128);
      if (leader) tc::mma_f16_ss(ts, qd, kd, idesc_s, 0u);
      for (int s = 0; s < 8; ++s) {
        const uint64_t bd = tc::make_desc_kmajor_noswz(b_a + (2 * s) * n_pad * 16 + hf * NH * 16, n_pad * 16, 128);
        if (leader) tc::mma_f16_ts(ts, tm + kAtColI + 8 * s, bd, idesc_s, 1u);
      }
      if (leader) tc::mma_commit(&s_full[hf]);
      __syncwarp();
    };
    // O_hf = P_hf [V | 1 | 0]_hf: every key half has its own accumulator (its probabilities are scaled by its own row maximum).
    // TS form: P (fp16 pairs) was stored over the S columns of its own half by the softmax threads -- chunk s sits at the start
    // of the S range of the thread that wrote it (the thread with the larger share, sub 0 of half 0 / sub 1 of half 1, owns
    // kCMax chunks).  With P in shared memory (SS form) the N = 32 MMA was bound by the 4 KB A read: 59 cycles against 39.
    auto issue_pv = [&](int hf, int b) {
      const uint32_t to = tm + (hf ? kAtColO1 : kAtColO0);
      const int cnt0 = hf ? nchunk - kCMax : kCMax;
      for (int s = 0; s < nchunk; ++s) {
        const int ks = hf * nchunk + s;
        // MN-major B: 8 keys x 16 B (8 dims) per core matrix, next 8 keys +128 B (LBO), next 8 dims +chunk (SBO)
        const uint64_t vd = tc::make_desc_kmajor_noswz(v_a + b * kAtVBytes + ks * 256, 128, kAtKChunk);
        const int pcol = s < cnt0 ? 8 * s : 16 * cnt0 + 8 * (s - cnt0);
        if (leader) tc::mma_f16_ts(to, tm + hf * kAtColS1 + pcol, vd, idesc_pv, s != 0 ? 1u : 0u);
      }
      if (leader) tc::mma_commit(&pv_done[hf]);
      __syncwarp();
    };
    // Ping-pong between the key halves, across tiles: the issue order is  PV0(i), S0(i+1), PV1(i), S1(i+1) -- while the softmax
    // warps of one half exponentiate, the tensor pipe works for the other half.  The two halves are equal loops
    // (softmax -> PV -> S -> softmax) that only meet on the MUFU; started back to back they stay ~1 100 cycles apart and their
    // exponential passes overlap half of the time at half rate each, so S1 is held `phase_delay` cycles behind S0 (anti-phase).
    // Measured alternatives that were NOT faster: issuing S0(i+1) before PV0(i); all sixteen softmax warps on one half at a time;
    // part of the exponentials as an FMA-pipe polynomial (profiles/r02_attention_poly_ab.jsonl); L2 prefetch of the next rows.
    long long t_s0 = 0;
    auto hold_half1 = [&]() {
      if (p.phase_delay > 0) while (clock64() - t_s0 < p.phase_delay) { }
    };
    const long long ntile = hi - lo;
    if (ntile > 0) {
      tc::mbar_wait(&qk_full[0], 0u);
      tc::fence_after_sync();
      t_s0 = clock64();
      issue_s(0, 0);
      hold_half1();
      issue_s(1, 0);
      if (leader) tc::mma_commit(&qk_empty[0]);
      __syncwarp();
    }
    for (long long it = 0; it < ntile; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const int b = (int)(it & 1), nb = b ^ 1;
      const bool more = it + 1 < ntile;
      tc::mbar_wait(&v_full[b], (uint32_t)((it >> 1) & 1));
      tc::mbar_wait(&p_full[0], ph);
      tc::mbar_wait(o_empty, ph ^ 1);               // the epilogue has read O0 / O1 of the previous tile
      tc::fence_after_sync();
      mark(it, 0, 0);
      issue_pv(0, b);
      if (more) {
        tc::mbar_wait(&qk_full[nb], (uint32_t)(((it + 1) >> 1) & 1));   // Q, K (and bias) of tile it+1 have landed
        tc::mbar_wait(&s_empty[0], ph);             // the softmax threads of half 0 have read S0 of tile it
        tc::fence_after_sync();
        mark(it, 0, 1);
        t_s0 = clock64();
        issue_s(0, nb);
        mark(it, 0, 2);
      }
      tc::mbar_wait(&p_full[1], ph);
      tc::fence_after_sync();
      mark(it, 0, 3);
      issue_pv(1, b);
      if (leader) tc::mma_commit(&v_empty[b]);
      __syncwarp();
      if (more) {
        tc::mbar_wait(&s_empty[1], ph);
        tc::fence_after_sync();
        hold_half1();
        mark(it, 0, 4);
        issue_s(1, nb);
        mark(it, 0, 5);
        if (leader) tc::mma_commit(&qk_empty[nb]);
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax + epilogue (warps 2..17) =====================
    // Warps 2-9 own key half 0, warps 10-17 key half 1; the two threads of a (query row, half) split its 16-column chunks.
    const int jj = (warp - 2) >> 2;
    const int hf = jj >> 1, sub = jj & 1;
    const int q = warp & 3;                   // TMEM lane quarter
    const int row = q * 32 + lane;
    // contiguous chunk ranges; the thread that also runs the epilogue (half 1 / sub 0) takes the smaller share
    const bool big = (sub == 0) != (hf == 1);                     // sub 0 of half 0 and sub 1 of half 1 take the larger share
    const int cnt = big ? kCMax : nchunk - kCMax;                 // warp-uniform
    const int c_lo = sub ? nchunk - cnt : 0;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t ts = tlane + hf * kAtColS1 + c_lo * 16;        // this thread's first S column (and first P column)
    const int bar_id = 1 + hf * 4 + q;
    const bool tr = TRACE && sub == 0 && q == 2;
    int it = 0;
    for (long long f = lo; f < hi; ++f, ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      // ---- pass 1: exact maximum of this key half of the row (two independent running maxima)
      float m = -INFINITY, m2 = -INFINITY;
      tc::mbar_wait(&s_full[hf], ph);
      tc::fence_after_sync();
      if (tr) mark(it, 1 + hf, 0);
      {
        uint32_t va[16], vb[16];
        if (cnt > 0) tc::tmem_ld16(ts, va);
#pragma unroll
        for (int k = 0; k < kCMax; k += 2) {
          if (k < cnt) {
            tc::tmem_ld_wait16(va);
            if (k + 1 < cnt) tc::tmem_ld16(ts + (k + 1) * 16, vb);
            m = max16(va, m);
          }
          if (k + 1 < kCMax && k + 1 < cnt) {
            tc::tmem_ld_wait16(vb);
            if (k + 2 < cnt) tc::tmem_ld16(ts + (k + 2) * 16, va);
            m2 = max16(vb, m2);
          }
        }
      }
      m = fmaxf(m, m2);
      if (tr) mark(it, 1 + hf, 1);
      s_max[(hf * 2 + sub) * 128 + row] = m;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      m = fmaxf(m, s_max[(hf * 2 + (sub ^ 1)) * 128 + row]);
      if (sub == 0) s_hmax[((it & 1) * 2 + hf) * 128 + row] = m;     // read by the epilogue of this tile
      if (tr) mark(it, 1 + hf, 2);
      // ---- pass 2: P = 2^(S - m) as fp16 pairs, stored over the S columns this thread has already read: chunk k (8 columns)
      //      lands inside S chunk k / 2 of the same thread, so neither the partner thread nor the load in flight is touched
      {
        uint32_t va[16], vb[16];
        auto emit = [&](const uint32_t (&v)[16], int k) {
          const uint32_t u0 = exp2_pack(__uint_as_float(v[0]), __uint_as_float(v[1]), m);
          const uint32_t u1 = exp2_pack(__uint_as_float(v[2]), __uint_as_float(v[3]), m);
          const uint32_t u2 = exp2_pack(__uint_as_float(v[4]), __uint_as_float(v[5]), m);
          const uint32_t u3 = exp2_pack(__uint_as_float(v[6]), __uint_as_float(v[7]), m);
          const uint32_t u4 = exp2_pack(__uint_as_float(v[8]), __uint_as_float(v[9]), m);
          const uint32_t u5 = exp2_pack(__uint_as_float(v[10]), __uint_as_float(v[11]), m);
          const uint32_t u6 = exp2_pack(__uint_as_float(v[12]), __uint_as_float(v[13]), m);
          const uint32_t u7 = exp2_pack(__uint_as_float(v[14]), __uint_as_float(v[15]), m);
          tc::tmem_st8(ts + 8 * k, u0, u1, u2, u3, u4, u5, u6, u7);
        };
        if (cnt > 0) tc::tmem_ld16(ts, va);
#pragma unroll
        for (int k = 0; k < kCMax; k += 2) {
          if (k < cnt) {
            tc::tmem_ld_wait16(va);
            if (k + 1 < cnt) tc::tmem_ld16(ts + (k + 1) * 16, vb);
            emit(va, k);
          }
          if (k + 1 < kCMax && k + 1 < cnt) {
            tc::tmem_ld_wait16(vb);
            if (k + 2 < cnt) tc::tmem_ld16(ts + (k + 2) * 16, va);
            emit(vb, k + 1);
          }
        }
      }
      if (tr) mark(it, 1 + hf, 4);
      tc::tmem_st_wait();            // P has landed in TMEM
      tc::fence_before_sync();       // ... and this thread's TMEM reads of this S half are complete
      __syncwarp();                  // one arrival per warp (the barriers count 8): per-thread arrivals serialise on one word
      if (lane == 0) { tc::mbar_arrive(&p_full[hf]); tc::mbar_arrive(&s_empty[hf]); }
      if (hf == 1 && sub == 0) {
        // ---- epilogue: combine the two halves (flash-attention style) and normalise:  O = (a0 O0 + a1 O1) / (a0 l0 + a1 l1),
        //      a_h = 2^(m_h - max(m0, m1)), l_h = the ones column of O_h
        const AttnTile t = attn_decode(p, f);
        tc::mbar_wait(&pv_done[0], ph);
        tc::mbar_wait(&pv_done[1], ph);
        tc::fence_after_sync();
        if (tr) mark(it, 2, 5);
        uint32_t o0[16], o1[16], l0[8], l1[8];
        tc::tmem_ld16(tlane + kAtColO0, o0);
        tc::tmem_ld8(tlane + kAtColO0 + 16, l0);
        tc::tmem_ld16(tlane + kAtColO1, o1);
        tc::tmem_ld8(tlane + kAtColO1 + 16, l1);
        const float m0 = s_hmax[((it & 1) * 2 + 0) * 128 + row], m1 = m;
        tc::tmem_ld_wait();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(o_empty);
        const int r = t.rt * 128 + row;
        if (r < n) {
          const float mm = fmaxf(m0, m1);
          float a0, a1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a0) : "f"(m0 - mm));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a1) : "f"(m1 - mm));
          const float inv = 1.f / (a0 * __uint_as_float(l0[0]) + a1 * __uint_as_float(l1[0]));
          a0 *= inv; a1 *= inv;
          __half* ob = p.out + (long long)t.b * p.C8 * T * 8 + ((long long)t.w * n + r) * 8;
#pragma unroll
          for (int dt = 0; dt < 2; ++dt) {
            uint4 hv;
            __half2* hp = reinterpret_cast<__half2*>(&hv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int k = dt * 8 + 2 * j;
              hp[j] = __floats2half2_rn(a0 * __uint_as_float(o0[k]) + a1 * __uint_as_float(o1[k]),
                                        a0 * __uint_as_float(o0[k + 1]) + a1 * __uint_as_float(o1[k + 1]));
            }
            *reinterpret_cast<uint4*>(ob + (long long)(2 * t.h + dt) * T * 8) = hv;
          }
        }
        if (tr) mark(it, 2, 6);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// B' operand images: [type][head][row tile][16 chunks of 8 query rows][n_pad keys][8] fp16 (see the header comment)
__global__ void attn_bias_pack_kernel(const float* __restrict__ table, const int32_t* __restrict__ region, __half* __restrict__ out,
                                      int heads, int n, int n_pad, int nrt, int ntypes, int ws0, int ws1, int ws2) {
  const long long total = (long long)ntypes * heads * nrt * 16 * n_pad * 8;
  const int s1 = 2 * ws2 - 1, s0 = (2 * ws1 - 1) * s1;
  const int lin_c = (ws0 - 1) * s0 + (ws1 - 1) * s1 + (ws2 - 1);
  constexpr float kLog2e = 1.4426950408889634f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int e = (int)(r % 8); r /= 8;
    const int j = (int)(r % n_pad); r /= n_pad;
    const int c = (int)(r % 16); r /= 16;
    const int rt = (int)(r % nrt); r /= nrt;
    const int h = (int)(r % heads); r /= heads;
    const int ty = (int)r;
    const int ig = rt * 128 + c * 8 + e;
    float v;
    if (j >= n) v = kAtPadBias;
    else if (ig >= n) v = 0.f;
    else {
      // tokens keep their coordinates in the MODULE window (relative_position_index[:n, :n], swin_unetr.py:514-516)
      const int li = (ig / (ws1 * ws2)) * s0 + ((ig / ws2) % ws1) * s1 + ig % ws2;
      const int lj = (j / (ws1 * ws2)) * s0 + ((j / ws2) % ws1) * s1 + j % ws2;
      v = table[(long long)(li - lj + lin_c) * heads + h] * kLog2e;
      if (region && region[(long long)ty * n + ig] != region[(long long)ty * n + j]) v += -100.f * kLog2e;   // compute_mask, swin_unetr.py:779-816
    }
    out[i] = __float2half_rn(v);
  }
}

}  // namespace b200

using namespace b200;

static bool attn_tc_shape_ok(int heads, int n, int ntypes) {
  return heads > 0 && n > 0 && n <= kAtNPadMax && ntypes >= 1 && ntypes <= 8;
}

extern "C" long long b200_window_attention_tc_bias_bytes(int heads, int n, int ntypes) {
  if (!attn_tc_shape_ok(heads, n, ntypes) || n > kAtNPadMax) return -1;
  const int n_pad = (n + 31) / 32 * 32, nrt = (n + 127) / 128;
  return (long long)ntypes * heads * nrt * 16 * n_pad * 16;
}

extern "C" int b200_window_attention_tc_pack_bias(const float* table, int heads, int n, int ws0, int ws1, int ws2,
                                                  const int32_t* region_types, int ntypes, void* packed, void* stream) {
  B200_REQUIRE(table && packed, "window_attention_tc_pack_bias: null pointer");
  B200_REQUIRE(attn_tc_shape_ok(heads, n, ntypes) && n <= kAtNPadMax, "window_attention_tc: unsupported shape (n = %d, types = %d)", n, ntypes);
  B200_REQUIRE(ws0 > 0 && ws1 > 0 && ws2 > 0 && n <= ws0 * ws1 * ws2, "window_attention_tc: window of %d tokens exceeds the module window", n);
  B200_REQUIRE(ntypes == 1 || region_types, "window_attention_tc_pack_bias: several mask types need their region rows");
  const int n_pad = (n + 31) / 32 * 32, nrt = (n + 127) / 128;
  const long long total = (long long)ntypes * heads * nrt * 16 * n_pad * 8;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 8192);
  attn_bias_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(table, region_types, (__half*)packed, heads, n, n_pad, nrt, ntypes, ws0, ws1, ws2);
  B200_LAUNCH_CHECK("attn_bias_pack_kernel");
  return B200_OK;
}

extern "C" int b200_window_attention_tc(const void* qkv, int N, int C, int heads, int nW, int n, const void* packed_bias,
                                        const int32_t* sched, int ntypes, void* out, void* stream) {
  B200_REQUIRE(qkv && out && packed_bias && sched, "window_attention_tc: null pointer");
  B200_REQUIRE(N > 0 && nW > 0, "window_attention_tc: empty problem");
  B200_REQUIRE(C == heads * 16, "window_attention_tc: head_dim must be 16 (C = %d, heads = %d)", C, heads);
  B200_REQUIRE(attn_tc_shape_ok(heads, n, ntypes) && n <= kAtNPadMax, "window_attention_tc: unsupported shape (n = %d, types = %d)", n, ntypes);
  AttnTcParams p;
  p.qkv = (const __half*)qkv; p.out = (__half*)out; p.bias = (const __half*)packed_bias; p.sched = sched;
  p.N = N; p.C8 = C / 8; p.heads = heads; p.nW = nW; p.n = n; p.n_pad = (n + 31) / 32 * 32; p.nrt = (n + 127) / 128; p.ntypes = ntypes;
  const long long total = (long long)N * nW * heads * p.nrt;
  dim3 grid((unsigned)std::min<long long>(total, num_sms()));
  void (*kern)(AttnTcParams) = nullptr;
  switch (p.n_pad) {
#define B200_ATTN_CASE(NP) case NP: kern = window_attention_tc_kernel<NP>; break;
    B200_ATTN_CASE(32) B200_ATTN_CASE(64) B200_ATTN_CASE(96) B200_ATTN_CASE(128) B200_ATTN_CASE(160) B200_ATTN_CASE(192)
    B200_ATTN_CASE(224) B200_ATTN_CASE(256) B200_ATTN_CASE(288) B200_ATTN_CASE(320) B200_ATTN_CASE(352)
#undef B200_ATTN_CASE
  }
  B200_REQUIRE(kern != nullptr, "window_attention_tc: no kernel for %d padded keys", p.n_pad);
  // per-device attribute: set on every call (cheap), so a second GPU in the same process works
  B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
  p.trace = nullptr;
  // B200_ATTN_PHASE: cycles per 32 keys by which the S MMAs of key half 1 trail those of half 0 (see the MMA issuer)
  static const int phase_q = [] { const char* e = std::getenv("B200_ATTN_PHASE"); return e ? std::atoi(e) : kAtPhaseDefault; }();
  p.phase_delay = phase_q * (p.n_pad / 32);
  // debug: B200_ATTN_TRACE=<file> records the phase timeline of CTA 0 of every n_pad = 352 launch (last launch wins)
  static const char* trace_path = std::getenv("B200_ATTN_TRACE");
  if (trace_path && p.n_pad == 352) {
    static long long* trace_buf = nullptr;
    if (!trace_buf) B200_CUDA(cudaMalloc(&trace_buf, 64 * 3 * 8 * sizeof(long long)));
    B200_CUDA(cudaMemsetAsync(trace_buf, 0, 64 * 3 * 8 * sizeof(long long), (cudaStream_t)stream));
    p.trace = trace_buf;
    kern = window_attention_tc_kernel<352, true>;
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    kern<<<grid, kAtThreads, kAtSmem, (cudaStream_t)stream>>>(p);
    B200_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    std::vector<long long> host(64 * 3 * 8);
    B200_CUDA(cudaMemcpy(host.data(), trace_buf, host.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    if (FILE* fp = std::fopen(trace_path, "wb")) { std::fwrite(host.data(), sizeof(long long), host.size(), fp); std::fclose(fp); }
    return B200_OK;
  }
  kern<<<grid, kAtThreads, kAtSmem, (cudaStream_t)stream>>>(p);
  B200_LAUNCH_CHECK("window_attention_tc_kernel");
  return B200_OK;
}
