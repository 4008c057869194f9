// sketch.cu — MashMap 3.5 query-fragment sketch on the GPU (sm_100a), batched over fragments.
//
// Replaces skch::CommonFunc::sketchSequence (src/map/include/commonFunc.hpp:217-323) as called per
// fragment by MappingCore::getSeedHits (src/map/include/mappingCore.hpp:61-76):
//   * makeUpperCaseAndValidDNA (commonFunc.hpp:110-142) fused into the fragment load;
//   * canonical k-mer hash = min(MurmurHash3_x64_128(fwd k-mer ASCII, seed 42).lo64,
//     MurmurHash3_x64_128(revcomp ASCII).lo64) (commonFunc.hpp:173-182,260-275; murmur3.h:226-303),
//     palindromes (hashFwd == hashBwd) and k-mers covering a non-ACGT base are dropped;
//   * bottom-s DISTINCT hashes with first / last position and the sign of the +-1 strand tally,
//     ascending by hash (closed form of the heap + hash-map loop at :277-321).
//
// B200 mapping: one CTA per fragment; the fragment is staged once in shared memory, every thread
// hashes k-mers from shared memory (two 64-bit Murmur3 evaluations per position, all in registers),
// the (hash, position) pairs are bitonic-sorted in shared memory, run heads are ranked with a
// ballot/popc block scan and the first s runs are folded and written as 32-byte MinmerInfo records.
#include "wfb_rt.h"
#include "../../include/wfmash_b200.h"

#include <string.h>
#include <string>
#include <vector>

void wfb_set_last_error_(const std::string& s); /* wfa_host.cu */
void wfb_count_launch_();
struct SkErr { SkErr& operator=(const std::string& s) { wfb_set_last_error_(s); return *this; } SkErr& operator=(const char* s) { wfb_set_last_error_(s); return *this; } };
static SkErr sk_last_error;

#ifndef WFB_EMU
#define SK_CHECK(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      sk_last_error = std::string(#call) + ": " + cudaGetErrorString(e_);                \
      rc = (e_ == cudaErrorMemoryAllocation) ? WFB_ENOMEM : WFB_ECUDA;                   \
      goto done;                                                                         \
    }                                                                                    \
  } while (0)
#endif



WFB_DEV uint64_t sk_rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
WFB_DEV uint64_t sk_fmix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

/* MurmurHash3_x64_128 (murmur3.h:226-303) of len <= 32 bytes given as four little-endian words
 * w[0..3] (unused high bytes zero), seed 42; returns the low 64 bits (commonFunc.hpp:173-182). */
WFB_DEV uint64_t sk_murmur3_lo64(const uint64_t w[4], int len) {
  const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
  uint64_t h1 = 42, h2 = 42;
  const int nblocks = len >> 4;
  int wi = 0;
  for (int i = 0; i < nblocks; ++i) {
    uint64_t k1 = w[wi], k2 = w[wi + 1];
    wi += 2;
    k1 *= c1; k1 = sk_rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = sk_rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= c2; k2 = sk_rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = sk_rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  }
  const int rem = len & 15;
  if (rem > 8) { uint64_t k2 = w[wi + 1]; k2 *= c2; k2 = sk_rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
  if (rem > 0) { uint64_t k1 = w[wi]; k1 *= c1; k1 = sk_rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
  h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
  h1 += h2; h2 += h1;
  h1 = sk_fmix64(h1); h2 = sk_fmix64(h2);
  h1 += h2;
  return h1;
}

/* makeUpperCaseAndValidDNA (commonFunc.hpp:110-142): a..z -> A..Z, anything but A,C,G,T -> N */
WFB_DEV uint8_t sk_clean_base(uint8_t c) {
  if (c > 96 && c < 123) c -= 32;
  return (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : (uint8_t)'N';
}
WFB_DEV uint8_t sk_comp(uint8_t c) { /* reverseComplement LUT (commonFunc.hpp:74-83) on cleaned bases */
  return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
}

#define SK_INVALID_VAL 0xFFFFFFFFu

/* dynamic shared memory: keys[N] u64 | vals[N] u32 | seq[len_pad] u8 */
WFB_KERNEL(wfb_sketch_kernel, const uint8_t* seq_base, const wfb_frag_t* frags, int nfrags, int ksize, int ssize, int npow2_max,
           wfb_minmer_t* out, int32_t* out_count
#ifdef WFB_EMU
           , unsigned char* smem_emu
#endif
) {
  WFB_KERNEL_PROLOGUE
#ifndef WFB_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
#else
  unsigned char* smem = smem_emu;
#endif
  WFB_SHARED int sh_warp[32];
  uint64_t* keys = (uint64_t*)smem;
  uint32_t* vals = (uint32_t*)(smem + (size_t)npow2_max * 8);
  uint8_t* sseq = (uint8_t*)(smem + (size_t)npow2_max * 12);
  for (int f = bid; f < nfrags; f += nblocks) {
    const wfb_frag_t fr = frags[f];
    const int len = fr.len;
    const int nk = len - ksize + 1;
    WFB_SYNC();
    if (nk <= 0) {
      if (WFB_TID == 0) out_count[f] = 0;
      continue;
    }
    int N = 1;
    while (N < nk) N <<= 1;
    const uint8_t* src = seq_base + fr.seq_offset;
    for (int i = WFB_TID; i < len; i += WFB_NT) sseq[i] = sk_clean_base(wfb_ldg8(src + i));
    WFB_SYNC();
    /* hash every k-mer start (commonFunc.hpp:252-311) */
    for (int i = WFB_TID; i < N; i += WFB_NT) {
      uint64_t key = ~0ULL;
      uint32_t val = SK_INVALID_VAL;
      if (i < nk) {
        uint64_t wf[4] = {0, 0, 0, 0}, wr[4] = {0, 0, 0, 0};
        bool ambig = false;
        for (int j = 0; j < ksize; ++j) {
          const uint8_t b = sseq[i + j];
          ambig |= (b == 'N');
          wf[j >> 3] |= (uint64_t)b << ((j & 7) * 8);
          const int jr = ksize - 1 - j; /* revcomp byte index of this base */
          wr[jr >> 3] |= (uint64_t)sk_comp(b) << ((jr & 7) * 8);
        }
        if (!ambig) {
          const uint64_t hf = sk_murmur3_lo64(wf, ksize), hb = sk_murmur3_lo64(wr, ksize);
          if (hf != hb) {
            key = hf < hb ? hf : hb;
            val = ((uint32_t)i << 1) | (hf < hb ? 1u : 0u);
          }
        }
      }
      keys[i] = key;
      vals[i] = val;
    }
    WFB_SYNC();
    /* bitonic sort by (hash, position) */
    for (int k = 2; k <= N; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = WFB_TID; t < (N >> 1); t += WFB_NT) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const int p = i | j;
          const bool up = (i & k) == 0;
          const uint64_t ka = keys[i], kb = keys[p];
          const uint32_t va = vals[i], vb = vals[p];
          const bool gt = (ka > kb) || (ka == kb && va > vb);
          if (gt == up) {
            keys[i] = kb; keys[p] = ka;
            vals[i] = vb; vals[p] = va;
          }
        }
        WFB_SYNC();
      }
    }
    /* rank run heads: each thread owns a contiguous chunk */
    const int per = (N + WFB_NT - 1) / WFB_NT;
    const int b0 = WFB_TID * per, b1 = min(N, b0 + per);
    int cnt = 0;
    for (int i = b0; i < b1; ++i) {
      const bool head = vals[i] != SK_INVALID_VAL && (i == 0 || keys[i] != keys[i - 1]);
      cnt += head ? 1 : 0;
    }
    int incl = cnt;
#ifndef WFB_EMU
    {
      const int lane = wfb_lane(), wid = WFB_TID >> 5;
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      if (lane == 31) sh_warp[wid] = incl;
      WFB_SYNC();
      if (wid == 0) {
        const int nw = (WFB_NT + 31) >> 5;
        int w = lane < nw ? sh_warp[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, w, o);
          if (lane >= o) w += y;
        }
        sh_warp[lane] = w;
      }
      WFB_SYNC();
      incl += wid ? sh_warp[wid - 1] : 0;
    }
    const int total = sh_warp[((WFB_NT + 31) >> 5) - 1];
#else
    (void)sh_warp;
    const int total = incl;
#endif
    int rank = incl - cnt; /* exclusive */
    wfb_minmer_t* o = out + (size_t)f * ssize;
    for (int i = b0; i < b1 && rank < ssize; ++i) {
      if (vals[i] == SK_INVALID_VAL) break; /* invalid entries sort last */
      if (i != 0 && keys[i] == keys[i - 1]) continue;
      /* fold the run (:285,302-303,317): first pos, last pos, sign of the strand tally */
      const uint64_t h = keys[i];
      int tally = 0;
      int e = i;
      while (e < N && keys[e] == h && vals[e] != SK_INVALID_VAL) {
        tally += (vals[e] & 1u) ? 1 : -1;
        ++e;
      }
      wfb_minmer_t m;
      m.hash = h;
      m.wpos = (int64_t)(vals[i] >> 1);
      m.wpos_end = (int64_t)(vals[e - 1] >> 1);
      m.seqId = fr.seq_id;
      m.strand = (int16_t)(tally > 0 ? 1 : (tally == 0 ? 0 : -1));
      m.pad_ = 0;
      o[rank] = m;
      ++rank;
    }
    if (WFB_TID == 0) out_count[f] = min(total, ssize);
  }
}

extern "C" int wfb_sketch_fragments(int device, const char* seq_base, int64_t seq_bytes, const wfb_frag_t* frags, int32_t n,
                                    int32_t kmer_size, int32_t sketch_size, wfb_minmer_t* out, int32_t* out_count,
                                    double* kernel_ms) {
  if (n < 0 || kmer_size <= 0 || kmer_size > 32 || sketch_size <= 0 || (n > 0 && (!seq_base || !frags || !out || !out_count))) {
    sk_last_error = "bad argument";
    return WFB_EINVAL;
  }
  if (kernel_ms) *kernel_ms = 0.0;
  if (n == 0) return WFB_OK;
  int maxlen = 0;
  for (int i = 0; i < n; ++i) {
    if (frags[i].len < 0 || frags[i].seq_offset < 0 || frags[i].seq_offset + frags[i].len > seq_bytes) {
      sk_last_error = "fragment out of range";
      return WFB_EINVAL;
    }
    if (frags[i].len > maxlen) maxlen = frags[i].len;
  }
  int npow2 = 1;
  while (npow2 < maxlen - kmer_size + 1) npow2 <<= 1;
  const size_t smem = (size_t)npow2 * 12 + (((size_t)maxlen + 15) & ~(size_t)15) + 16;
  if (smem > 200 * 1024) { sk_last_error = "fragment too long for the shared-memory sketch kernel (max ~16 kb)"; return WFB_EINVAL; }
  int rc = WFB_OK;
#ifndef WFB_EMU
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    sk_last_error = "no CUDA device (this library has no CPU path)";
    return WFB_ENODEV;
  }
  uint8_t* d_seq = nullptr;
  wfb_frag_t* d_frags = nullptr;
  wfb_minmer_t* d_out = nullptr;
  int32_t* d_cnt = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  SK_CHECK(cudaSetDevice(device));
  SK_CHECK(cudaMalloc(&d_seq, (size_t)seq_bytes + 16));
  SK_CHECK(cudaMalloc(&d_frags, sizeof(wfb_frag_t) * (size_t)n));
  SK_CHECK(cudaMalloc(&d_out, sizeof(wfb_minmer_t) * (size_t)n * sketch_size));
  SK_CHECK(cudaMalloc(&d_cnt, sizeof(int32_t) * (size_t)n));
  SK_CHECK(cudaMemcpy(d_seq, seq_base, (size_t)seq_bytes, cudaMemcpyHostToDevice));
  SK_CHECK(cudaMemcpy(d_frags, frags, sizeof(wfb_frag_t) * (size_t)n, cudaMemcpyHostToDevice));
  SK_CHECK(cudaMemset(d_out, 0, sizeof(wfb_minmer_t) * (size_t)n * sketch_size));
  SK_CHECK(cudaFuncSetAttribute(wfb_sketch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SK_CHECK(cudaEventCreate(&e0));
  SK_CHECK(cudaEventCreate(&e1));
  {
    cudaDeviceProp prop;
    SK_CHECK(cudaGetDeviceProperties(&prop, device));
    const int grid = n < prop.multiProcessorCount * 8 ? n : prop.multiProcessorCount * 8;
    SK_CHECK(cudaEventRecord(e0, 0));
    wfb_sketch_kernel<<<grid, 256, smem, 0>>>(d_seq, d_frags, n, kmer_size, sketch_size, npow2, d_out, d_cnt);
    wfb_count_launch_();
    SK_CHECK(cudaEventRecord(e1, 0));
    SK_CHECK(cudaGetLastError());
    SK_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (kernel_ms) *kernel_ms = ms;
  }
  SK_CHECK(cudaMemcpy(out, d_out, sizeof(wfb_minmer_t) * (size_t)n * sketch_size, cudaMemcpyDeviceToHost));
  SK_CHECK(cudaMemcpy(out_count, d_cnt, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
done:
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(d_seq);
  cudaFree(d_frags);
  cudaFree(d_out);
  cudaFree(d_cnt);
#else
  (void)device;
  std::vector<unsigned char> sm(smem + 64);
  memset(out, 0, sizeof(wfb_minmer_t) * (size_t)n * sketch_size);
  for (int b = 0; b < 1; ++b)
    wfb_sketch_kernel(b, 1, (const uint8_t*)seq_base, frags, n, kmer_size, sketch_size, npow2, out, out_count, sm.data());
#endif
  return rc;
}
