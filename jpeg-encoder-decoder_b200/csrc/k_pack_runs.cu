// k_pack_runs.cu — pass 2 of the batched encoder on sm_100a: from the token runs of k_pixels_to_tokens to the un-stuffed
// scan bits (reference main/encoder.c:434-502).
//   k_runs_prepare    per job: the first DC token of every run (encoder.c:168-177 predicts across the whole plane, a run only
//                     knows its own blocks) with its histogram entry; exclusive prefix of the runs' token counts inside each
//                     scan = where each run goes in scan order
//   k_compact_tokens  per batch of 16 consecutive runs (one warp): copies the runs' tokens into scan order (ws.tok2), resolved to
//                     (code word, length), and adds their code bits (code length + magnitude bits [+ ZRL codes]) to the totals
//                     of the token chunks they land in
//   k_scan_tchunks    per job: exclusive prefix of the chunk bits inside each of the three scans, scan placement in the job's
//                     scratch area, clearing of the words two chunks share (JB_FUSE_SCAN = 1 runs it in the CTA of
//                     k_compact_tokens that finishes the job last instead: measured slower, profiles/r2b_summary.md)
//   k_pack_tchunks    per chunk of 256 tokens (one warp): 8 consecutive tokens per lane are concatenated in registers, a warp
//                     scan gives the bit offsets, the chunk's bits are assembled in shared memory and flushed as big-endian
//                     words (the first and the last word of a chunk are OR-ed into place)
// Byte stuffing, headers and layout stay with k_count_ff (+ layout_job) / k_stuff (k_entropy.cu).
#include "jpegb200_internal.cuh"
#include "walk.cuh"
#include <type_traits>

namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr int PR_WARPS = 8;                 // warps per CTA
#ifndef JB_FUSE_SCAN
#define JB_FUSE_SCAN 0
#endif
constexpr int PR_TOK = JB_TCHUNK / 32;      // tokens per lane
static_assert(PR_TOK == 8, "the register concatenation below is written for 8 tokens per lane");
// a token is at most 3 ZRL codes + one code + 11 magnitude bits = 75 bits
constexpr int PR_STAGE_WORDS = (31 + JB_TCHUNK * 75 + 31) / 32 + 2;

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t* total) {
  const int lane = threadIdx.x & 31;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += n;
  }
  *total = __shfl_sync(FULL, inc, 31);
  return inc - v;
}

// run r of the job -> scan (0 Y, 1 Cb, 2 Cr); scans 1 and 2 share the chroma tables
__device__ __forceinline__ int run_scan(uint32_t r, uint32_t nrc) { return r < 2u * nrc ? 0 : (r < 3u * nrc ? 1 : 2); }

// enc[t][i]: the resolved form of table index i (0..255 AC symbols, 256..271 DC categories, 0 above: void tokens) of table set
// t (0 luma, 1 chroma): a token's category is a function of its index (AC: i & 15, DC: i - 256), so the shift that makes room
// for the magnitude bits is applied here, once per CTA, instead of once per token:
//   enc = code << (category + 5) | (code length + category)        resolved token = enc | magnitude bits << 5
__device__ __forceinline__ void load_enc(const JbWs& ws, uint32_t job, uint32_t (*enc)[512]) {
  const uint32_t* g = ws.enc + (size_t)job * 4 * 256;
  for (int k = threadIdx.x; k < 2 * 512; k += blockDim.x) {
    const int t = k >> 9, i = k & 511;
    const uint32_t e = i < 256 ? g[(2 * t + 1) * 256 + i] : (i < 272 ? g[(2 * t) * 256 + (i - 256)] : 0u);
    const uint32_t cat = i < 256 ? (uint32_t)(i & 15) : (uint32_t)(i - 256);
    enc[t][i] = e ? ((e >> 5) << (cat + 5)) | ((e & 31u) + cat) : 0u;
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_runs_prepare(JbWs ws) {
  __shared__ uint32_t h[32];
  __shared__ uint32_t wsum[9];
  const JbJob job = ws.jobs[blockIdx.x];
  JbJobState* st = ws.state + blockIdx.x;
  if (st->error & JB_ERR_TOKENS) return;        // the job overflowed its token budget: its run records are not valid
  const uint32_t nrc = jb_runs_chroma(job.w, job.h), nr = 4u * nrc;
  const JbRun* runs = ws.runs + job.run_off;
  if (threadIdx.x < 32) h[threadIdx.x] = 0;
  __syncthreads();
  for (uint32_t r = threadIdx.x; r < nr; r += 256) {
    const int plane = run_scan(r, nrc);
    const bool first = r == 0 || r == 2u * nrc || r == 3u * nrc;
    const JbRun run = runs[r];
    const int dc = (int)(short)(run.dc & 0xFFFFu);
    const int prev = first ? 0 : (int)(short)(runs[r - 1].dc >> 16);
    const int diff = dc - prev;
    const int cat = 32 - __clz(abs(diff));
    ws.tok[run.tok] = jb_token(diff, cat, 256 + cat, 0);
    atomicAdd(&h[(plane ? 16 : 0) + cat], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 32 && h[threadIdx.x]) {
    int* G = ws.hist + (size_t)blockIdx.x * 4 * 257;
    atomicAdd(&G[(threadIdx.x >= 16 ? 2 * 257 : 0) + (threadIdx.x & 15)], (int)h[threadIdx.x]);
  }
  // where each run goes in scan order
  uint32_t start = 0;
  for (int s = 0; s < 3; s++) {
    const uint32_t r0 = s == 0 ? 0u : (s == 1 ? 2u * nrc : 3u * nrc), n = s == 0 ? 2u * nrc : nrc;
    uint32_t carry = 0;
    for (uint32_t b = 0; b < n; b += 256) {
      const uint32_t k = b + threadIdx.x;
      uint32_t v = k < n ? runs[r0 + k].ntok : 0, total;
      const uint32_t ex = cta_exclusive_scan(v, wsum, &total);
      if (k < n) ws.run_base[job.run_off + r0 + k] = start + carry + ex;
      carry += total;
    }
    if (threadIdx.x == 0) { st->tok_total[s] = carry; st->tok_start[s] = start; }
    // the rest of the scan's last token chunk reads as tokens without bits: k_pack_tchunks loads whole chunks and tests no token
    // against the end of its scan
    {
      const uint32_t end = start + carry, pad = (0u - end) & (uint32_t)(JB_TCHUNK - 1);
      for (uint32_t k = threadIdx.x; k < pad; k += 256) ws.tok2[job.tok_off + end + k] = 0;
    }
    start = (start + carry + JB_TCHUNK - 1) & ~(uint32_t)(JB_TCHUNK - 1);
  }
}

// ---------------------------------------------------------------------------------------------
// Per job, once all its runs are compacted: exclusive prefix of the chunk bits inside each scan, placement of the scans in
// the job's scratch area, clearing of the words two chunks share.  Runs in the CTA that finishes the job's compaction last.
__device__ __forceinline__ void scan_tchunks(const JbWs& ws, uint32_t jobid, const JbJob& job, uint32_t* wsum /*[9]*/, uint32_t* s_word /*[4]*/) {
  JbJobState* st = ws.state + jobid;
  const uint32_t* cbits = ws.tchunk_bits + job.tchunk_off;
  uint32_t* cbase = ws.tchunk_base + job.tchunk_off;
  uint32_t seg_bits[3];
  for (int s = 0; s < 3; s++) {
    const uint32_t c0 = st->tok_start[s] / JB_TCHUNK, n = (st->tok_total[s] + JB_TCHUNK - 1) / JB_TCHUNK;
    uint32_t carry = 0;
    for (uint32_t b = 0; b < n; b += 256) {
      const uint32_t k = b + threadIdx.x;
      uint32_t v = k < n ? __ldcg(cbits + c0 + k) : 0, total;            // accumulated by atomics of other CTAs: read at L2
      const uint32_t ex = cta_exclusive_scan(v, wsum, &total);
      if (k < n) cbase[c0 + k] = carry + ex;
      carry += total;
    }
    seg_bits[s] = carry;
  }
  if (threadIdx.x == 0) {
    uint32_t w = 0;
    for (int s = 0; s < 3; s++) {
      st->seg_bits[s] = seg_bits[s];
      st->seg_word[s] = w;
      s_word[s] = w;
      w = (w + (seg_bits[s] + 31) / 32 + 1 + 3) & ~3u;     // +1 slack word, 16-byte aligned scans
    }
    s_word[3] = w;
    if (w > job.scratch_cap) atomicOr(&st->error, (uint32_t)JB_ERR_SCRATCH);
  }
  __syncthreads();
  if (s_word[3] > job.scratch_cap) return;
  // clear every word that two chunks (or a chunk and the scan end) may share
  for (int s = 0; s < 3; s++) {
    const uint32_t c0 = st->tok_start[s] / JB_TCHUNK, n = (st->tok_total[s] + JB_TCHUNK - 1) / JB_TCHUNK;
    uint32_t* scr = ws.scratch + job.scratch_off + s_word[s];
    for (uint32_t k = threadIdx.x; k <= n; k += 256) {
      const uint32_t bit = k < n ? cbase[c0 + k] : seg_bits[s];
      scr[bit >> 5] = 0;
    }
  }
}

// A warp takes a *batch* of CP_BATCH consecutive runs of one scan: their records arrive with one load per lane and reach the
// run loop by shuffles, their destinations are contiguous in scan order, and the code bits of the tokens are kept per lane for
// the current token chunk and the next one and leave with one warp reduction per chunk (r2: one warp per run paid 40
// warp-instructions of set-up per run and 70 per 256-token step, 61 % of the kernel, on 164 000 runs per 64-frame wave of which
// half - the chroma runs - hold 38 tokens).
#ifndef JB_CP_BATCH
#define JB_CP_BATCH 16
#endif
#ifndef JB_CP_STEP
#define JB_CP_STEP 4
#endif
constexpr int CP_BATCH = JB_CP_BATCH;      // runs per batch
constexpr int CP_STEP = JB_CP_STEP;        // 32-token slices whose loads are in flight together
// CTAs per SM the register allocation must allow (r2: 1 (46 registers) 255, 5: 261, 6: 262, 7: 250 Gpix/s together with the same
// setting for k_pack_tchunks; more resident warps hide the token loads, until the registers spill)
#ifndef JB_COMPACT_MIN_CTAS
#define JB_COMPACT_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(PR_WARPS * 32, JB_COMPACT_MIN_CTAS) k_compact_tokens(JbWs ws) {
  __shared__ uint32_t enc[2][512];
  if (ws.state[blockIdx.y].error & JB_ERR_TOKENS) return;
  const JbJob job = ws.jobs[blockIdx.y];
  const uint32_t nrc = jb_runs_chroma(job.w, job.h);
  load_enc(ws, blockIdx.y, enc);
  __syncthreads();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* tok2 = ws.tok2 + job.tok_off;
  uint32_t* cbits = ws.tchunk_bits + job.tchunk_off;
  const uint32_t nb_y = (2u * nrc + CP_BATCH - 1) / CP_BATCH, nb_c = (nrc + CP_BATCH - 1) / CP_BATCH, nbatch = nb_y + 2u * nb_c;
  for (uint32_t b = blockIdx.x * PR_WARPS + warp; b < nbatch; b += gridDim.x * PR_WARPS) {
    const int s = b < nb_y ? 0 : (b < nb_y + nb_c ? 1 : 2);
    const uint32_t first = (b - (s == 0 ? 0u : s == 1 ? nb_y : nb_y + nb_c)) * CP_BATCH;
    const uint32_t r0 = s == 0 ? 0u : (s == 1 ? 2u * nrc : 3u * nrc), ns = s == 0 ? 2u * nrc : nrc;
    const uint32_t nvalid = min((uint32_t)CP_BATCH, ns - first);
    uint32_t src_l = 0, n_l = 0, d_l = 0;
    if (lane < nvalid) {
      const uint2 rec = __ldg(reinterpret_cast<const uint2*>(ws.runs + job.run_off + r0 + first + lane));     // first token, tokens
      src_l = rec.x; n_l = rec.y;
      d_l = __ldg(ws.run_base + job.run_off + r0 + first + lane);
    }
    const uint32_t* e = enc[s ? 1 : 0];
    const uint32_t zrl_len = e[0xF0] & 31u;
    // code bits of the tokens that land in chunk cA (accA) and cA + 1 (accB), per lane
    uint32_t cA = __shfl_sync(FULL, d_l, 0) / JB_TCHUNK, accA = 0, accB = 0;
    // (run, step) pairs in one loop, software-pipelined: the loads of the next step - of this run or of the next one - are in
    // flight while the current step is resolved and stored
    uint32_t r = 0, k0 = 0;
    uint32_t src = __shfl_sync(FULL, src_l, 0), n = __shfl_sync(FULL, n_l, 0), d = __shfl_sync(FULL, d_l, 0);
    uint32_t t[CP_STEP];
#pragma unroll
    for (int j = 0; j < CP_STEP; j++) t[j] = 32u * j + lane < n ? __ldg(ws.tok + src + lane + 32 * j) : JB_TOKEN_VOID;
    while (r < nvalid) {
      uint32_t r2 = r, k2 = k0 + 32 * CP_STEP, src2 = src, n2 = n, d2 = d;
      if (k2 >= n) {
        r2 = r + 1; k2 = 0;
        src2 = __shfl_sync(FULL, src_l, r2 & 31); n2 = __shfl_sync(FULL, n_l, r2 & 31); d2 = __shfl_sync(FULL, d_l, r2 & 31);
        if (r2 >= nvalid) n2 = 0;
      }
      uint32_t tn[CP_STEP];
#pragma unroll
      for (int j = 0; j < CP_STEP; j++) tn[j] = k2 + 32u * j + lane < n2 ? __ldg(ws.tok + src2 + k2 + lane + 32 * j) : JB_TOKEN_VOID;
      if (k0 < n) {
        const uint32_t cs = (d + k0) / JB_TCHUNK;       // destinations are contiguous: the chunk of a step's first token is cA or cA + 1
        if (cs != cA) {
          accA = __reduce_add_sync(FULL, accA);
          if (lane == 0 && accA) atomicAdd(&cbits[cA], accA);
          accA = accB; accB = 0; cA = cs;
        }
        const uint32_t rem = n - k0;                    // tokens of the run from k0 on
        uint32_t* dp = tok2 + d + k0 + lane;
        // a step spans at most two chunks; most lie inside one (warp-uniform test), and then no token asks which chunk it is in
        const bool one_chunk = (d + k0 + min(rem, 32u * CP_STEP) - 1u) / JB_TCHUNK == cA;
        auto slices = [&](auto single) {
#pragma unroll
          for (int j = 0; j < CP_STEP; j++) {
            if (32u * j >= rem) break;                    // warp-uniform
            const uint32_t tk = t[j], ent = e[(tk >> 15) & 0x1FFu], z = tk >> 24;
            // resolved token: code word << 5 | length; a token that carries ZRLs keeps its raw form behind the escape length 31
            if (32u * j + lane < rem) dp[32 * j] = z ? (tk << 5) | 31u : ent | ((tk & 0x7FFu) << 5);
            const uint32_t len = (ent & 31u) + z * zrl_len;                               // a void token: no bits
            if constexpr (decltype(single)::value) {
              accA += len;
            } else {
              const bool in_b = (d + k0 + 32u * j + lane) / JB_TCHUNK != cA;
              accA += in_b ? 0u : len;
              accB += in_b ? len : 0u;
            }
          }
        };
        if (one_chunk) slices(std::true_type());
        else slices(std::false_type());
      }
      r = r2; k0 = k2; src = src2; n = n2; d = d2;
#pragma unroll
      for (int j = 0; j < CP_STEP; j++) t[j] = tn[j];
    }
    accA = __reduce_add_sync(FULL, accA);
    accB = __reduce_add_sync(FULL, accB);
    if (lane == 0) {
      if (accA) atomicAdd(&cbits[cA], accA);
      if (accB) atomicAdd(&cbits[cA + 1], accB);
    }
  }
#if JB_FUSE_SCAN
  // the CTA that finishes the job last scans its chunks (one launch and one thin kernel less per wave)
  __shared__ uint32_t s_wsum[9], s_word[4], s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&ws.state[blockIdx.y].ctas_compacted, 1u) == gridDim.x - 1u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  scan_tchunks(ws, blockIdx.y, job, s_wsum, s_word);
#endif
}

#if !JB_FUSE_SCAN
__global__ void __launch_bounds__(256) k_scan_tchunks(JbWs ws) {
  __shared__ uint32_t s_wsum[9], s_word[4];
  if (ws.state[blockIdx.x].error & JB_ERR_TOKENS) return;
  scan_tchunks(ws, blockIdx.x, ws.jobs[blockIdx.x], s_wsum, s_word);
}
#endif

// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// OR `len` (<= 32) bits into the MSB-first bit image at bit position pos (rare path: lanes whose tokens carry ZRLs).
__device__ __forceinline__ void or_bits_s(uint32_t* img, uint32_t pos, uint32_t bits, uint32_t len) {
  if (!len) return;
  const uint32_t sh = pos & 31, wi = pos >> 5;
  const uint64_t v = ((uint64_t)bits << (64 - len)) >> sh;
  atomicOr(img + wi, (uint32_t)(v >> 32));
  if ((uint32_t)v) atomicOr(img + wi + 1, (uint32_t)v);
}

// Concatenate the lane's 8 code words (MSB first) behind the sbit & 31 bits that precede them in their first word, in a
// W-word register accumulator, and store the words into the bit image: the first and the last word are shared with the
// neighbouring lanes (OR), the others are the lane's own.  W = 9 holds any 8 tokens (8 x 27 + 31 bits).
template <int W>
__device__ __forceinline__ void concat_store(uint32_t* stage, const uint32_t (&word)[PR_TOK], const uint32_t (&len)[PR_TOK], uint32_t sbit, uint32_t nbits) {
  uint32_t A[W];
#pragma unroll
  for (int q = 0; q < W; q++) A[q] = 0;
  uint32_t tot = sbit & 31;
#pragma unroll
  for (int j = 0; j < PR_TOK; j++) {
#pragma unroll
    for (int q = 0; q < W - 1; q++) A[q] = __funnelshift_l(A[q + 1], A[q], len[j]);
    A[W - 1] = (A[W - 1] << len[j]) | (len[j] ? word[j] : 0u);
    tot += len[j];
  }
  const uint32_t pad = (32u - (tot & 31u)) & 31u;
#pragma unroll
  for (int q = 0; q < W - 1; q++) A[q] = __funnelshift_l(A[q + 1], A[q], pad);
  A[W - 1] <<= pad;
  const int nw = nbits ? (int)((tot + pad) >> 5) : 0;
  uint32_t* dst = stage + (sbit >> 5) - (W - nw);
#pragma unroll
  for (int q = 0; q < W; q++) {
    if (q >= W - nw) {
      if (q == W - nw || q == W - 1) atomicOr(dst + q, A[q]);
      else dst[q] = A[q];
    }
  }
}

// The same with up to two ZRL codes (runs of 16 zeros, encoder.c:470-476) in front of a token: they form one more code word of
// at most 32 bits.  About one token in 500 carries a ZRL, but two chunks in five hold one, and the OR-into-the-image path for
// such a lane cost the whole warp some 300 instructions; here the warp pays 8 more appends only in those chunks.
__device__ __forceinline__ void concat_store_z4(uint32_t* stage, const uint32_t (&word)[PR_TOK], const uint32_t (&len)[PR_TOK], uint32_t sbit, uint32_t nbits,
                                                uint32_t zr, uint32_t zrl_code, uint32_t zrl_len) {
  constexpr int W = 4;
  uint32_t A[W] = {0, 0, 0, 0};
  uint32_t tot = sbit & 31;
  const uint32_t two = (zrl_code << (zrl_len & 31u)) | zrl_code;
#pragma unroll
  for (int j = 0; j < PR_TOK; j++) {
    const uint32_t z = (zr >> (2 * j)) & 3u;
    const uint32_t pl = z * zrl_len, pw = z == 2u ? two : zrl_code;     // z = 0: length 0, the word does not matter
    if (pl == 32u) {                                                     // two 16-bit ZRL codes: a whole word moves up
#pragma unroll
      for (int q = 0; q < W - 1; q++) A[q] = A[q + 1];
      A[W - 1] = pw;
    } else {
#pragma unroll
      for (int q = 0; q < W - 1; q++) A[q] = __funnelshift_l(A[q + 1], A[q], pl);
      A[W - 1] = (A[W - 1] << pl) | (pl ? pw : 0u);
    }
    tot += pl;
#pragma unroll
    for (int q = 0; q < W - 1; q++) A[q] = __funnelshift_l(A[q + 1], A[q], len[j]);
    A[W - 1] = (A[W - 1] << len[j]) | (len[j] ? word[j] : 0u);
    tot += len[j];
  }
  const uint32_t pad = (32u - (tot & 31u)) & 31u;
#pragma unroll
  for (int q = 0; q < W - 1; q++) A[q] = __funnelshift_l(A[q + 1], A[q], pad);
  A[W - 1] <<= pad;
  const int nw = nbits ? (int)((tot + pad) >> 5) : 0;
  uint32_t* dst = stage + (sbit >> 5) - (W - nw);
#pragma unroll
  for (int q = 0; q < W; q++) {
    if (q >= W - nw) {
      if (q == W - nw || q == W - 1) atomicOr(dst + q, A[q]);
      else dst[q] = A[q];
    }
  }
}

#ifndef JB_PACK_MIN_CTAS
#define JB_PACK_MIN_CTAS 6
#endif
__global__ void __launch_bounds__(PR_WARPS * 32, JB_PACK_MIN_CTAS) k_pack_tchunks(JbWs ws) {
  __shared__ uint32_t stage_all[PR_WARPS][PR_STAGE_WORDS];
  const JbJob job = ws.jobs[blockIdx.y];
  const JbJobState* st = ws.state + blockIdx.y;
  if (st->error) return;
  uint32_t nchunk[3], total_chunks = 0;
  for (int s = 0; s < 3; s++) { nchunk[s] = (st->tok_total[s] + JB_TCHUNK - 1) / JB_TCHUNK; total_chunks += nchunk[s]; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* stage = stage_all[warp];
  for (uint32_t f = blockIdx.x * PR_WARPS + warp; f < total_chunks; f += gridDim.x * PR_WARPS) {
    const int s = f < nchunk[0] ? 0 : (f < nchunk[0] + nchunk[1] ? 1 : 2);
    const uint32_t c = f - (s == 0 ? 0u : s == 1 ? nchunk[0] : nchunk[0] + nchunk[1]);
    const uint32_t cg = st->tok_start[s] / JB_TCHUNK + c;                      // chunk id inside the job
    const uint32_t base_bits = ws.tchunk_base[job.tchunk_off + cg], total = ws.tchunk_bits[job.tchunk_off + cg];
    const uint32_t phase = base_bits & 31;
    uint32_t* gw = ws.scratch + job.scratch_off + st->seg_word[s] + (base_bits >> 5);   // word that holds the chunk's first bit
    const bool first_shared = phase != 0 || total < 32;
    const uint32_t last_word = (phase + total - 1) >> 5;
    const uint4* tp = reinterpret_cast<const uint4*>(ws.tok2 + job.tok_off + (size_t)cg * JB_TCHUNK) + 2 * lane;
    uint32_t t8[PR_TOK];
    {                                       // the tail of a scan's last chunk holds zeros (k_runs_prepare)
      const uint4 a = __ldg(tp), b = __ldg(tp + 1);
      t8[0] = a.x; t8[1] = a.y; t8[2] = a.z; t8[3] = a.w; t8[4] = b.x; t8[5] = b.y; t8[6] = b.z; t8[7] = b.w;
    }
    // resolved tokens (k_compact_tokens): code word << 5 | length; length 31 escapes to a raw token that carries ZRLs
    const uint32_t* enc_ac = ws.enc + ((size_t)blockIdx.y * 4 + (s ? 3 : 1)) * 256;
    uint32_t word[PR_TOK], len[PR_TOK], zr = 0, nbits = 0, zrl_code = 0, zrl_len = 0;
#pragma unroll
    for (int j = 0; j < PR_TOK; j++) {
      const uint32_t t = t8[j];
      word[j] = t >> 5;
      len[j] = t & 31u;
      if (len[j] == 31u) {                                    // rare: decode the raw token with the table in global memory
        const uint32_t raw = t >> 5, ent = __ldg(enc_ac + ((raw >> 15) & 0xFFu)), cat = (raw >> 11) & 15u, z = raw >> 24;
        const uint32_t ez = __ldg(enc_ac + 0xF0);
        zrl_code = ez >> 5; zrl_len = ez & 31u;
        word[j] = ((ent >> 5) << cat) | (raw & 0x7FFu);
        len[j] = (ent & 31u) + cat;
        zr |= z << (2 * j);
        nbits += z * zrl_len;
      }
      nbits += len[j];
    }
    uint32_t chunk_total;
    const uint32_t ex = warp_excl_scan(nbits, &chunk_total);
    const uint32_t endbit = phase + chunk_total, nwr = (endbit + 31) >> 5;
    for (uint32_t k = lane; k < nwr + 1; k += 32) stage[k] = 0;
    __syncwarp();
    const uint32_t sbit = phase + ex;
    // the lanes' bits rarely exceed four words (5 bits per token on photographic content): the short accumulator shifts
    // 3 words per token instead of 8
    const bool wide = __any_sync(FULL, (sbit & 31u) + nbits > 128u);
    const bool zrls = __any_sync(FULL, zr != 0);
    const bool zrl3 = __any_sync(FULL, ((zr & (zr >> 1)) & 0x5555u) != 0);      // a token behind 48 or more zeros: three ZRL codes
    if (!wide && !zrl3) {
      if (zrls) concat_store_z4(stage, word, len, sbit, nbits, zr, __shfl_sync(FULL, zrl_code, __ffs(__ballot_sync(FULL, zr != 0)) - 1),
                                __shfl_sync(FULL, zrl_len, __ffs(__ballot_sync(FULL, zr != 0)) - 1));
      else concat_store<4>(stage, word, len, sbit, nbits);
    } else if (zr == 0) {
      concat_store<9>(stage, word, len, sbit, nbits);
    } else {
      uint32_t pos = sbit;
#pragma unroll
      for (int j = 0; j < PR_TOK; j++) {
        for (uint32_t z = (zr >> (2 * j)) & 3u; z; z--) { or_bits_s(stage, pos, zrl_code, zrl_len); pos += zrl_len; }
        or_bits_s(stage, pos, word[j], len[j]);
        pos += len[j];
      }
    }
    __syncwarp();
    // flush: every word of the image that holds bits of the chunk
    for (uint32_t k = lane; k < nwr; k += 32) {
      const uint32_t w = __byte_perm(stage[k], 0, 0x0123);     // first bit of the stream = MSB of the first byte
      if ((k == 0 && first_shared) || (k == last_word && (endbit & 31))) atomicOr(gw + k, w);
      else gw[k] = w;
    }
    __syncwarp();
  }
}

}  // namespace

// CTAs per job: enough to fill the GPU a few times over, few enough that the table load of a CTA is amortised over many items
static uint32_t item_ctas(int njobs, uint32_t max_items) {
  const uint32_t want = (max_items + PR_WARPS - 1) / PR_WARPS, cap = (uint32_t)((148 * 8 * 2 + njobs - 1) / njobs);
  return want < cap ? (want ? want : 1) : (cap ? cap : 1);
}
void jb_launch_runs_prepare(const JbWs& ws, int njobs, cudaStream_t st) { k_runs_prepare<<<njobs, 256, 0, st>>>(ws); }
void jb_launch_compact_tokens(const JbWs& ws, int njobs, uint32_t max_runs, cudaStream_t st) {
  k_compact_tokens<<<dim3(item_ctas(njobs, (max_runs + CP_BATCH - 1) / CP_BATCH + 3), njobs), PR_WARPS * 32, 0, st>>>(ws);
#if !JB_FUSE_SCAN
  k_scan_tchunks<<<njobs, 256, 0, st>>>(ws);
#endif
}
void jb_launch_pack_tchunks(const JbWs& ws, int njobs, uint32_t max_tchunks, cudaStream_t st) {
  k_pack_tchunks<<<dim3(item_ctas(njobs, max_tchunks), njobs), PR_WARPS * 32, 0, st>>>(ws);
}
