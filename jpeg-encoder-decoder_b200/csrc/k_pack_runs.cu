// k_pack_runs.cu — pass 2 of the batched encoder on sm_100a: from the token stream of k_pixels_to_tokens to the un-stuffed
// scan bits (reference main/encoder.c:434-502).  The unit of work is a run (consecutive blocks of one scan whose tokens
// are contiguous, jpegb200_internal.cuh); one warp streams one run, no CTA barrier is involved.
//   k_run_bits    bits of every run = sum over its tokens of (code length + magnitude bits [+ ZRL codes])
//   k_scan_runs   per job: exclusive prefix of the run bits inside each of the three scans, scan placement in the
//                 job's scratch area, clearing of the words that two runs share          (as k_scan for chunks)
//   k_pack_runs   per run: 8 consecutive tokens per lane and step are concatenated in registers, a warp scan gives the
//                 bit offsets, the step's bits are assembled in shared memory and flushed as big-endian words
//                 (the first and the last word of a run are OR-ed into place)
// Byte stuffing, headers and layout stay with k_count_ff / k_layout / k_stuff (k_entropy.cu).
#include "jpegb200_internal.cuh"
#include "walk.cuh"

namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr int PR_WARPS = 8;                 // runs per CTA
constexpr int PR_TOK = 8;                   // tokens per lane and step
constexpr int PR_STEP = 32 * PR_TOK;
// a token is at most 3 ZRL codes + one code + 11 magnitude bits = 75 bits
constexpr int PR_STAGE_WORDS = (31 + PR_STEP * 75 + 31) / 32 + 2;

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

// run r of the job -> table set (0 luma, 1 chroma) and scan (0 Y, 1 Cb, 2 Cr)
__device__ __forceinline__ int run_scan(uint32_t r, uint32_t nrc) { return r < 2u * nrc ? 0 : (r < 3u * nrc ? 1 : 2); }

// cost[t][i]: code length of table index i (0..255 AC symbols, 256..271 DC categories) of table set t
__device__ __forceinline__ void load_enc(const JbWs& ws, uint32_t job, uint32_t (*enc)[272]) {
  const uint32_t* g = ws.enc + (size_t)job * 4 * 256;
  for (int k = threadIdx.x; k < 2 * 272; k += blockDim.x) {
    const int t = k / 272, i = k - t * 272;
    enc[t][i] = i < 256 ? g[(2 * t + 1) * 256 + i] : g[(2 * t) * 256 + (i - 256)];
  }
}

__global__ void __launch_bounds__(PR_WARPS * 32) k_run_bits(JbWs ws) {
  __shared__ uint32_t enc[2][272];
  const JbJob job = ws.jobs[blockIdx.y];
  const uint32_t nrc = jb_runs_chroma(job.w, job.h), nr = 4u * nrc;
  load_enc(ws, blockIdx.y, enc);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t r = blockIdx.x * PR_WARPS + warp; r < nr; r += gridDim.x * PR_WARPS) {
    JbRun* run = ws.runs + job.run_off + r;
    const uint32_t* e = enc[run_scan(r, nrc) ? 1 : 0];
    const uint32_t zrl_len = e[0xF0] & 31u;
    const uint32_t* tok = ws.tok + run->tok;
    const uint32_t n = run->ntok;
    uint32_t sum = 0;
    for (uint32_t k = lane; k < n; k += 32) {
      const uint32_t t = __ldg(tok + k);
      sum += (e[(t >> 15) & 0x1FF] & 31u) + ((t >> 11) & 15u) + (t >> 24) * zrl_len;
    }
    sum = __reduce_add_sync(FULL, sum);
    if (lane == 0) run->bits = sum;
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scan_runs(JbWs ws) {
  __shared__ uint32_t wsum[9];
  __shared__ uint32_t s_word[4];
  const JbJob job = ws.jobs[blockIdx.x];
  JbJobState* st = ws.state + blockIdx.x;
  const uint32_t nrc = jb_runs_chroma(job.w, job.h);
  const JbRun* runs = ws.runs + job.run_off;
  uint32_t* base = ws.run_base + job.run_off;
  uint32_t seg_bits[3];
  for (int s = 0; s < 3; s++) {
    const uint32_t r0 = s == 0 ? 0u : (s == 1 ? 2u * nrc : 3u * nrc), n = s == 0 ? 2u * nrc : nrc;
    uint32_t carry = 0;
    for (uint32_t b = 0; b < n; b += 256) {
      const uint32_t k = b + threadIdx.x;
      uint32_t v = k < n ? runs[r0 + k].bits : 0, total;
      const uint32_t ex = cta_exclusive_scan(v, wsum, &total);
      if (k < n) base[r0 + k] = carry + ex;
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
  // clear every word that two runs (or a run and the scan end) may share
  for (int s = 0; s < 3; s++) {
    const uint32_t r0 = s == 0 ? 0u : (s == 1 ? 2u * nrc : 3u * nrc), n = s == 0 ? 2u * nrc : nrc;
    uint32_t* scr = ws.scratch + job.scratch_off + s_word[s];
    for (uint32_t k = threadIdx.x; k <= n; k += 256) {
      const uint32_t bit = k < n ? base[r0 + k] : seg_bits[s];
      scr[bit >> 5] = 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// OR `len` (<= 32) bits into the MSB-first bit image at bit position pos (rare path: lanes whose tokens carry ZRLs).
__device__ __forceinline__ void or_bits_s(uint32_t* img, uint32_t pos, uint32_t bits, uint32_t len) {
  if (!len) return;
  const uint32_t sh = pos & 31, wi = pos >> 5;
  const uint64_t v = ((uint64_t)bits << (64 - len)) >> sh;
  atomicOr(img + wi, (uint32_t)(v >> 32));
  if ((uint32_t)v) atomicOr(img + wi + 1, (uint32_t)v);
}

__global__ void __launch_bounds__(PR_WARPS * 32) k_pack_runs(JbWs ws) {
  __shared__ uint32_t enc[2][272];
  __shared__ uint32_t stage_all[PR_WARPS][PR_STAGE_WORDS];
  const JbJob job = ws.jobs[blockIdx.y];
  const JbJobState* st = ws.state + blockIdx.y;
  if (st->error) return;
  const uint32_t nrc = jb_runs_chroma(job.w, job.h), nr = 4u * nrc;
  load_enc(ws, blockIdx.y, enc);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t r = blockIdx.x * PR_WARPS + warp; r < nr; r += gridDim.x * PR_WARPS) {
  const JbRun run = ws.runs[job.run_off + r];
  const int s = run_scan(r, nrc);
  const uint32_t* e = enc[s ? 1 : 0];
  uint32_t* stage = stage_all[warp];
  const uint32_t zrl_code = e[0xF0] >> 5, zrl_len = e[0xF0] & 31u;
  const uint32_t base_bits = ws.run_base[job.run_off + r], total = run.bits, ntok = run.ntok;
  const uint32_t phase = base_bits & 31;
  uint32_t* gw = ws.scratch + job.scratch_off + st->seg_word[s] + (base_bits >> 5);   // word that holds the run's first bit
  const bool first_shared = phase != 0 || total < 32;
  const uint32_t last_word = (phase + total - 1) >> 5;
  const uint32_t* tok = ws.tok + run.tok;

  uint32_t rbase = phase;            // run-relative bit position where the step starts (bit 0 = MSB of gw[0])
  if (lane == 0) stage[0] = 0;
  for (uint32_t r0 = 0; r0 < ntok; r0 += PR_STEP) {
    const uint32_t first = r0 + lane * PR_TOK;
    uint32_t word[PR_TOK], len[PR_TOK], zr = 0, nbits = 0;
#pragma unroll
    for (int j = 0; j < PR_TOK; j++) {
      const bool live = first + j < ntok;
      const uint32_t t = live ? __ldg(tok + first + j) : 0u;
      const uint32_t ent = e[(t >> 15) & 0x1FF];
      const uint32_t cat = (t >> 11) & 15u;
      word[j] = ((ent >> 5) << cat) | (t & 0x7FFu);
      len[j] = live ? (ent & 31u) + cat : 0u;
      const uint32_t z = t >> 24;
      zr |= z << (2 * j);
      nbits += len[j] + z * zrl_len;
    }
    uint32_t step_total;
    const uint32_t ex = warp_excl_scan(nbits, &step_total);
    const uint32_t r_in = rbase & 31, endbit = r_in + step_total, nwr = (endbit + 31) >> 5;
    for (uint32_t k = lane + 1; k < nwr + 1; k += 32) stage[k] = 0;        // stage[0] carries the previous step's tail
    __syncwarp();
    const uint32_t sbit = r_in + ex;
    if (zr == 0) {
      uint32_t A[9];
#pragma unroll
      for (int q = 0; q < 9; q++) A[q] = 0;
      uint32_t tot = sbit & 31;
#pragma unroll
      for (int j = 0; j < PR_TOK; j++) {
#pragma unroll
        for (int q = 0; q < 8; q++) A[q] = __funnelshift_l(A[q + 1], A[q], len[j]);
        A[8] = (A[8] << len[j]) | (len[j] ? word[j] : 0u);
        tot += len[j];
      }
      const uint32_t pad = (32u - (tot & 31u)) & 31u;
#pragma unroll
      for (int q = 0; q < 8; q++) A[q] = __funnelshift_l(A[q + 1], A[q], pad);
      A[8] <<= pad;
      const int nw = nbits ? (int)((tot + pad) >> 5) : 0;
      uint32_t* dst = stage + (sbit >> 5) - (9 - nw);
#pragma unroll
      for (int q = 0; q < 9; q++) {
        if (q >= 9 - nw) {
          if (q == 9 - nw || q == 8) atomicOr(dst + q, A[q]);
          else dst[q] = A[q];
        }
      }
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
    // flush: complete words of the image; the very last word of the run even if partial
    const bool last_step = r0 + PR_STEP >= ntok;
    const uint32_t full = endbit >> 5, rem = endbit & 31, w0 = rbase >> 5;
    const uint32_t nflush = full + ((last_step && rem) ? 1u : 0u);
    for (uint32_t k = lane; k < nflush; k += 32) {
      const uint32_t w = __byte_perm(stage[k], 0, 0x0123);     // first bit of the stream = MSB of the first byte
      const uint32_t g = w0 + k;
      if ((g == 0 && first_shared) || (g == last_word && k == full)) atomicOr(gw + g, w);
      else gw[g] = w;
    }
    const uint32_t tail = (!last_step && rem) ? stage[full] : 0u;
    __syncwarp();
    if (lane == 0) stage[0] = tail;
    rbase += step_total;
  }
  __syncwarp();
  }
}

}  // namespace

// CTAs per job: enough to fill the GPU a few times over, few enough that the table load of a CTA is amortised over many runs
static uint32_t run_ctas(int njobs, uint32_t max_runs) {
  const uint32_t want = (max_runs + PR_WARPS - 1) / PR_WARPS, cap = (uint32_t)((148 * 8 * 2 + njobs - 1) / njobs);
  return want < cap ? want : (cap ? cap : 1);
}
void jb_launch_run_bits(const JbWs& ws, int njobs, uint32_t max_runs, cudaStream_t st) {
  k_run_bits<<<dim3(run_ctas(njobs, max_runs), njobs), PR_WARPS * 32, 0, st>>>(ws);
}
void jb_launch_scan_runs(const JbWs& ws, int njobs, cudaStream_t st) { k_scan_runs<<<njobs, 256, 0, st>>>(ws); }
void jb_launch_pack_runs(const JbWs& ws, int njobs, uint32_t max_runs, cudaStream_t st) {
  k_pack_runs<<<dim3(run_ctas(njobs, max_runs), njobs), PR_WARPS * 32, 0, st>>>(ws);
}
