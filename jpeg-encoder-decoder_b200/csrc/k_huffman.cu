// k_huffman.cu — stage 2 of the encoder on sm_100a: symbol statistics and the four per-image
// optimal Huffman tables (reference main/encoder.c:180-381).
//
//   k_symbol_stats   one thread per 8x8 block: DC prediction (encoder.c:168-177), DC category and
//                    AC run/size symbols (encoder.c:303-358) accumulated in shared-memory histograms,
//                    then merged into the job's four 257-bin histograms.
//   k_build_huffman  one warp per table: exact replay of init_huff_table (encoder.c:180-301) — the
//                    two-least-frequent selection with its (freq asc, index desc) tie-break, chain
//                    merging, the 16-bit length limiter, the reserved code point, sym_sorted and
//                    the canonical codes — leaving every field of huff_code as the reference does.
#include "jpegb200_internal.cuh"
#include "walk.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(JB_CHUNK_BLOCKS) k_symbol_stats(JbWs ws, int dc_from_raw, int store_dc_diff) {
  __shared__ JbChunkTokens ct;
  __shared__ int h[8][JB_CHUNK_HIST];                    // one histogram per warp: 16 DC categories, then 256 AC symbols
  const JbJob job = ws.jobs[blockIdx.y];
  const uint32_t cy = jb_chunks(jb_nby(job.w, job.h)), cc = jb_chunks(jb_nbc(job.w, job.h));
  uint32_t c = blockIdx.x;
  if (c >= cy + 2 * cc) return;
  const int s = c < cy ? 0 : (c < cy + cc ? 1 : 2);
  c -= (s == 0 ? 0 : s == 1 ? cy : cy + cc);
  const JbSeg seg = jb_seg(job, s);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int k = tid; k < 8 * JB_CHUNK_HIST; k += JB_CHUNK_BLOCKS) (&h[0][0])[k] = 0;
  const uint32_t total = jb_stage_chunk(ws, seg, c, dc_from_raw, store_dc_diff, ct);   // its barriers also cover the zeroing

  // every thread takes an equal run of consecutive tokens; h[w][0..15] DC categories, h[w][16..271] AC symbols
  const uint32_t share = (total + JB_CHUNK_BLOCKS - 1) / JB_CHUNK_BLOCKS, first = min(total, tid * share), count = min(share, total - first);
  const int16_t* coef = ws.coef + seg.coef0 + (size_t)c * JB_CHUNK_BLOCKS * 64;
  int* hw = h[warp];
  JbCursor cur = jb_cursor_init(ct, first);
  for (uint32_t j = 0; j < share; j++) {
    if (j < count) {
      const JbToken t = jb_next_token(ct, cur, coef);
      atomicAdd(&hw[t.idx < 256 ? 16 + t.idx : t.idx - 256], 1);
      if (t.zrl) atomicAdd(&hw[16 + 0xF0], t.zrl);
    }
  }
  __syncthreads();
  // per-chunk counts (k_chunk_bits turns them into the chunk's bit length once the tables exist) and the job's totals
  int* ch = ws.chunk_hist + (size_t)(seg.chunk0 + c) * JB_CHUNK_HIST;
  int* g = ws.hist + (size_t)blockIdx.y * 4 * 257 + (s ? 2 * 257 : 0);
  for (int k = tid; k < JB_CHUNK_HIST; k += JB_CHUNK_BLOCKS) {
    int v = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) v += h[w][k];
    ch[k] = v;
    if (v) atomicAdd(&g[k < 16 ? k : 257 + (k - 16)], v);
  }
}

// ---------------------------------------------------------------------------------------------
// One warp per table.  Shared-memory state per warp mirrors huff_code plus a tail[] array that
// replaces the reference's walk to the end of v1's chain.
struct TabSmem {
  int freq[257 + 31];    // padded so that every lane can read 9 strided entries
  int len[257 + 31];
  int next[257];
  int tail[257];
  int grp[257 + 31];     // representative (chain head) of each symbol
  int clf[32];           // code_len_freq
  int sorted[256], slen[256], code[256];
};

// Warps (= tables) per CTA.  r2 measured 16 (a quarter of the CTAs, so that a multi-lane batch leaves more SMs to the other
// lanes' k_pixels_to_tokens): 53 us against 32 us per 64-job wave, and a slower step - the merge loop is issue-dense and four
// warps per scheduler slow each other down more than the freed SMs are worth.
#ifndef JB_TAB_WARPS
#define JB_TAB_WARPS 4
#endif
constexpr int TAB_WARPS = JB_TAB_WARPS;

// warp-wide minimum: one REDUX for 32-bit keys, a shuffle tree for 64-bit keys
__device__ __forceinline__ unsigned warp_min(unsigned v) { return __reduce_min_sync(0xFFFFFFFFu, v); }
__device__ __forceinline__ unsigned long long warp_min(unsigned long long v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xFFFFFFFFu, v, o);
    v = t < v ? t : v;
  }
  return v;
}

// KeyT = unsigned when every frequency is below 2^23 (crops up to 2^23 pixels), else unsigned long long.
template <typename KeyT>
__global__ void __launch_bounds__(TAB_WARPS * 32) k_build_huffman(JbWs ws, int ntables) {
  extern __shared__ __align__(16) unsigned char tab_smem_raw[];
  TabSmem* sm_all = reinterpret_cast<TabSmem*>(tab_smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * TAB_WARPS + warp;
  if (t >= ntables) return;
  TabSmem& sm = sm_all[warp];
  const int job = t >> 2, which = t & 3;
  const int* hist = ws.hist + (size_t)t * 257;
  JbHuff* hc = ws.huff + t;
  JbJobState* state = ws.state + job;
  (void)which;

  // Symbol k = lane + 32*j lives in slot j of lane `lane`: frequency, code length and chain head stay in registers for
  // the whole merge loop; only the chain links (next/tail, touched by one lane, never read back by the selection) are
  // in shared memory.
  int rf[9], rl[9], rg[9];
#pragma unroll
  for (int j = 0; j < 9; j++) {
    const int k = lane + 32 * j;
    int f = 0;
    if (k < 256) f = hist[k];
    else if (k == 256) f = 1;                          // reserved code point, encoder.c:367
    rf[j] = f;
    rl[j] = 0;
    rg[j] = k;
    if (k < 257) { sm.next[k] = -1; sm.tail[k] = k; }
  }
  __syncwarp();

  // -- merge loop (encoder.c:190-228).  key = (frequency << 9) | (511 - index): smallest key = least frequent, ties to
  //    the LATER index, exactly the reference's ascending scan with '<=' (encoder.c:196-207).
  constexpr KeyT NONE = ~(KeyT)0;
  for (;;) {
    KeyT best1 = NONE, best2 = NONE;
#pragma unroll
    for (int j = 0; j < 9; j++) {
      const KeyT key = rf[j] ? (((KeyT)(unsigned)rf[j] << 9) | (KeyT)(511 - (lane + 32 * j))) : NONE;
      if (key < best1) { best2 = best1; best1 = key; }
      else if (key < best2) best2 = key;
    }
    const KeyT g1 = warp_min(best1);
    const KeyT g2 = warp_min(best1 == g1 ? best2 : best1);      // keys are unique: the index is part of them
    if (g2 == NONE) break;
    const int v1 = 511 - (int)(g1 & 511u), v2 = 511 - (int)(g2 & 511u);
    const int fsum = (int)(g1 >> 9) + (int)(g2 >> 9);
    // every member of both chains gets one bit longer and now belongs to v1; v1 carries the merged frequency
#pragma unroll
    for (int j = 0; j < 9; j++) {
      const int k = lane + 32 * j;
      if (rg[j] == v1 || rg[j] == v2) { rl[j]++; rg[j] = v1; }
      if (k == v1) rf[j] = fsum;
      if (k == v2) rf[j] = 0;
    }
    if (lane == 0) {
      sm.next[sm.tail[v1]] = v2;
      sm.tail[v1] = sm.tail[v2];
    }
  }
#pragma unroll
  for (int j = 0; j < 9; j++) {
    const int k = lane + 32 * j;
    sm.freq[k] = rf[j];
    sm.len[k] = rl[j];
  }
  __syncwarp();

  // The rest is short and strictly sequential: lane 0 replays it, the warp copies results out.
  // Out of contract (unreachable from pixels, only through jpegb200_debug_build_tables): (a) a histogram without any used
  // symbol - the reference's scan for the longest used length (encoder.c:255-257) then runs off code_len_freq[] into
  // next[256] and its code loop (:283-300) runs off sym_sorted[]; (b) all 256 symbols used - encoder.c:277 then reads
  // sym_sorted[256], i.e. sym_code_len[0], and zeroes the length of the symbol with that number.  The encoder's
  // histograms always hold a DC category and an EOB, and its walker emits at most 162 of the 256 AC symbols; the kernel
  // stays inside its arrays for both cases instead of replaying the overruns (tests exclude them).
  int* clf = sm.clf;
  int* sorted = sm.sorted;
  int* slen = sm.slen;
  int* code = sm.code;
  for (int k = lane; k < 256; k += 32) { sorted[k] = -1; slen[k] = 0; code[k] = -1; }
  clf[lane] = 0;
  __syncwarp();
  // code_len_freq (encoder.c:230-236): histogram of the lengths of all 257 symbols
  {
    bool overflow = false;
    for (int k = lane; k < 257; k += 32) {
      int l = sm.len[k];
      if (l > 31) { overflow = true; l = 31; }         // the reference would write past code_len_freq[32]
      if (l) atomicAdd(&clf[l], 1);
    }
    if (overflow) atomicOr(&state->error, (uint32_t)JB_ERR_CODELEN);
  }
  // sym_sorted (encoder.c:262-268): symbols 0..255 ordered by (pre-limit length, symbol).  Rank of symbol k = lane + 32 j:
  // symbols of shorter lengths + symbols of its own length in earlier slots j + those in lower lanes of its slot
  // (MATCH.ANY groups the lanes of a slot by length).  r2: a ballot per (length 1..31, slot) was 52 % of the kernel's time.
  {
    int* cnt = sm.tail;                 // dead after the merge loop: [0..31] symbols seen so far per length, [32..63] first rank per length
    cnt[lane] = 0;
    __syncwarp();
    int within[8], lens[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int l = sm.len[lane + 32 * j];
      if (l > 31) l = 0;                 // never equal to a length the reference's loop visits: not sorted
      lens[j] = l;
      const unsigned peers = __match_any_sync(0xFFFFFFFFu, l);
      const int seen = cnt[l];
      within[j] = seen + __popc(peers & ((1u << lane) - 1u));
      __syncwarp();
      if (l && lane == __ffs(peers) - 1) cnt[l] = seen + __popc(peers);
      __syncwarp();
    }
    const int c = lane ? cnt[lane] : 0;
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += n;
    }
    cnt[32 + lane] = inc - c;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (lens[j]) sorted[cnt[32 + lens[j]] + within[j]] = lane + 32 * j;
  }
  __syncwarp();
  if (lane == 0) {
    for (int i = 31; i > 16; i--)                        // encoder.c:239-254
      while (clf[i] > 0) {
        int j = i - 2;
        while (clf[j] <= 0) j--;
        clf[i] -= 2; clf[i - 1]++; clf[j + 1] += 2; clf[j]--;
      }
    { int i = 16; while (i > 0 && clf[i] == 0) i--; clf[i]--; }   // encoder.c:255-258
    int k = 0, cd = 0;
    for (int l = 1; l <= 16; l++) {                      // encoder.c:271-276 and :280-300
      for (int c = 0; c < clf[l] && k < 256 && sorted[k] >= 0; c++) { slen[sorted[k]] = l; code[sorted[k]] = cd++; k++; }
      cd <<= 1;
    }
    if (k < 256 && sorted[k] < 0) sorted[255] = 0;       // encoder.c:277 writes sym_code_len[-1] == sym_sorted[255]
  }
  __syncwarp();
  for (int k = lane; k < 257; k += 32) {
    hc->sym_freq[k] = sm.freq[k];
    hc->code_len[k] = sm.len[k];
    hc->next[k] = sm.next[k];
  }
  hc->code_len_freq[lane] = clf[lane];
  for (int k = lane; k < 256; k += 32) {
    hc->sym_sorted[k] = sorted[k];
    hc->sym_code_len[k] = slen[k];
    hc->sym_code[k] = code[k];
    // the entropy kernels' packed form (what k_pack_tables derives from a huff_code)
    const int l = slen[k];
    ws.enc[(size_t)t * 256 + k] = (l > 0 && l <= 16) ? ((uint32_t)code[k] << 5) | (uint32_t)l : 0u;
  }
}

// Pack (code, length) of every symbol for the entropy kernels; also used after the drop-in
// write_jpg uploaded caller-provided tables.
__global__ void k_pack_tables(JbWs ws, int ntables) {
  const int t = blockIdx.x;
  if (t >= ntables) return;
  const JbHuff* hc = ws.huff + t;
  const int k = threadIdx.x;
  int l = hc->sym_code_len[k];
  uint32_t e = 0;
  if (l > 0 && l <= 16) e = ((uint32_t)hc->sym_code[k] << 5) | (uint32_t)l;
  ws.enc[(size_t)t * 256 + k] = e;
}

}  // namespace

void jb_launch_symbol_stats(const JbWs& ws, int njobs, uint32_t max_chunks, int dc_from_raw, int store_dc_diff, cudaStream_t st) {
  k_symbol_stats<<<dim3(max_chunks, njobs), JB_CHUNK_BLOCKS, 0, st>>>(ws, dc_from_raw, store_dc_diff);
}
// Also writes the packed tables (ws.enc): k_pack_tables is needed only for caller-provided huff_codes.
void jb_launch_build_huffman(const JbWs& ws, int njobs, bool wide_keys, cudaStream_t st) {
  static bool opted[64][2] = {};          // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  const int smem = (int)sizeof(TabSmem) * TAB_WARPS;
  if (!opted[dev][wide_keys ? 1 : 0]) {
    if (wide_keys) cudaFuncSetAttribute(k_build_huffman<unsigned long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    else cudaFuncSetAttribute(k_build_huffman<unsigned>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    opted[dev][wide_keys ? 1 : 0] = true;
  }
  int nt = njobs * 4;
  if (wide_keys) k_build_huffman<unsigned long long><<<(nt + TAB_WARPS - 1) / TAB_WARPS, TAB_WARPS * 32, smem, st>>>(ws, nt);
  else k_build_huffman<unsigned><<<(nt + TAB_WARPS - 1) / TAB_WARPS, TAB_WARPS * 32, smem, st>>>(ws, nt);
}
void jb_launch_pack_tables(const JbWs& ws, int njobs, cudaStream_t st) {
  k_pack_tables<<<njobs * 4, 256, 0, st>>>(ws, njobs * 4);
}
